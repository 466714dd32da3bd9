python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest_c.log
tail -3 gpurun_out/r2_pytest_c.log
for v in "" "UDGPU_XSTREAMS=1" "UDGPU_XCHUNKS=1" "UDGPU_XCHUNKS=4"; do
env $v python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 60 --warmup 5 --no-parity > gpurun_out/r2_bench_n2_b_$v.json 2> gpurun_out/r2_bench_n2_b.err
tail -c 300 gpurun_out/r2_bench_n2_b.err
done
