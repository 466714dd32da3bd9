TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 8 --steps 60 --warmup 5 > gpurun_out/r2_bench_n8_a.json 2> gpurun_out/r2_bench_n8_a.err
tail -c 300 gpurun_out/r2_bench_n8_a.err
UDGPU_XMODE=store $TR bench.py --gpus 8 --steps 60 --warmup 5 --no-1024 --no-parity > gpurun_out/r2_bench_n8_store.json 2> gpurun_out/r2_bench_n8_b.err
UDGPU_XCHUNKS=2 $TR bench.py --gpus 8 --steps 60 --warmup 5 --no-1024 --no-parity > gpurun_out/r2_bench_n8_c2.json 2> gpurun_out/r2_bench_n8_b.err
UDGPU_XCHUNKS=1 $TR bench.py --gpus 8 --steps 60 --warmup 5 --no-1024 --no-parity > gpurun_out/r2_bench_n8_c1.json 2> gpurun_out/r2_bench_n8_b.err
