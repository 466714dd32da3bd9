python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_pytest_d.log; tail -3 gpurun_out/r2_pytest_d.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for v in "" "UDGPU_FILL_FUSED=0"; do
env $v $TR bench.py --gpus 2 --steps 60 --warmup 5 --no-parity --no-1024 > gpurun_out/r2_bench_n2_c_$v.json 2> gpurun_out/r2_bench_n2_c.err
done
