python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_pytest_e.log; tail -3 gpurun_out/r2_pytest_e.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 60 --warmup 5 > gpurun_out/r2_bench_n2_d.json 2> gpurun_out/r2_bench_n2_d.err
UDGPU_XMODE=store $TR bench.py --gpus 2 --steps 12 --warmup 3 --grid 1024,1024,1024 --no-parity --no-1024 --no-cpu > gpurun_out/r2_bench_n2_1024_store.json 2>> gpurun_out/r2_bench_n2_d.err
tail -c 300 gpurun_out/r2_bench_n2_d.err
