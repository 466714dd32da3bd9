TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 4 --no-cpu > gpurun_out/r2_bench_final_n4.json 2> gpurun_out/r2_bench_final_n4.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_final_n4.json').read().strip().split('\n')[-1])
print(d['value'], d['ms_per_step'], d['parity']['ok'], d['grid1024']['value'], d['grid1024']['ms_per_step'])"
