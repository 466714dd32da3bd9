TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 8 --workload ibm --grid 512,512,256 --steps 60 --warmup 5 --no-cpu > gpurun_out/r2_bench_ibm_n8.json 2> gpurun_out/r2_n8_y.err
$TR bench.py --gpus 8 --workload scalars --grid 512,512,512 --steps 30 --warmup 5 --no-cpu --no-parity > gpurun_out/r2_bench_scalars_n8.json 2>> gpurun_out/r2_n8_y.err
tail -c 300 gpurun_out/r2_n8_y.err
