python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest_b.log
python bench.py --steps 60 --warmup 5 > gpurun_out/r2_bench_n1_a.json 2> gpurun_out/r2_bench_n1_a.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 60 --warmup 5 > gpurun_out/r2_bench_n2_a.json 2> gpurun_out/r2_bench_n2_a.err
tail -3 gpurun_out/r2_pytest_b.log; tail -c 600 gpurun_out/r2_bench_n1_a.err; tail -c 600 gpurun_out/r2_bench_n2_a.err
