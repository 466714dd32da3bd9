CS="compute-sanitizer --target-processes all --error-exitcode 7 --print-limit 20"
( timeout 900 $CS --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -k "(substeps_track_oracle and shape0) or (test_scalars and shape0 and kw0) or (ibm_substeps and shape0) or (resident_channel and shape0) or (poisson_solve and shape5) or rk3_step_host" 2>&1 | tail -25 ) > gpurun_out/r2_sanitizer_memcheck_1gpu.txt 2>&1
echo "memcheck 1gpu rc=$?" >> gpurun_out/r2_sanitizer_memcheck_1gpu.txt
( timeout 900 $CS --tool racecheck python -m pytest tests/test_gpu_parity.py -x -q -k "(test_fused_advection_subgrid and shape2 and kw0) or (poisson_solve and shape2) or (test_scalars and shape0 and kw0 and 0-) or (test_closure and shape0)" 2>&1 | tail -25 ) > gpurun_out/r2_sanitizer_racecheck_1gpu.txt 2>&1
( timeout 1200 $CS --tool memcheck python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/mgpu_worker.py quick 2>&1 | tail -30 ) > gpurun_out/r2_sanitizer_memcheck_2gpu.txt 2>&1
tail -5 gpurun_out/r2_sanitizer_*.txt
