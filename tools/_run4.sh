TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$TR tools/trace_substep.py --out gpurun_out/trace_n8 2> gpurun_out/trace.err
$TR tools/trace_substep.py --grid 1024,1024,1024 --out gpurun_out/trace_n8 2>> gpurun_out/trace.err
UDGPU_XCHUNKS=2 $TR tools/trace_substep.py --out gpurun_out/trace_n8_c2 2>> gpurun_out/trace.err
tail -c 500 gpurun_out/trace.err
