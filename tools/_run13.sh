python bench.py --workload thermo --grid 512,512,256 --steps 30 --warmup 5 --no-cpu --no-parity > gpurun_out/r2_bench_thermo_n1c.json 2> gpurun_out/r2_bench_thermo_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_thermo_n1c.json').read().strip().split('\n')[-1])
print(d['value'], d['ms_per_step'], d['gpu_launches'], {k:round(v['ms'],4) for k,v in d['roofline']['families'].items()})"
