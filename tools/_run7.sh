TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
B="bench.py --gpus 8 --no-parity --no-1024 --no-cpu"
UDGPU_XMODE=store $TR $B --grid 1024,1024,1024 --steps 12 --warmup 3 > gpurun_out/r2_n8_1024_storefill.json 2> gpurun_out/r2_n8_x.err
UDGPU_XMODE=store UDGPU_FILL_FUSED=0 $TR $B --grid 1024,1024,1024 --steps 12 --warmup 3 > gpurun_out/r2_n8_1024_store.json 2>> gpurun_out/r2_n8_x.err
$TR $B --grid 1024,1024,1024 --steps 12 --warmup 3 > gpurun_out/r2_n8_1024_ce.json 2>> gpurun_out/r2_n8_x.err
$TR $B --steps 60 --warmup 5 > gpurun_out/r2_n8_weak_storefill.json 2>> gpurun_out/r2_n8_x.err
UDGPU_FILL_FUSED=0 $TR $B --steps 60 --warmup 5 > gpurun_out/r2_n8_weak_store.json 2>> gpurun_out/r2_n8_x.err
tail -c 400 gpurun_out/r2_n8_x.err
