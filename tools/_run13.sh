python -m pytest tests/test_thermo.py tests/test_gpu_parity.py tests/test_gpu_golden.py -q -x -k "thermo or glue or masscorr or forces or bottom" 2>&1 | tail -4
python bench.py --workload thermo --grid 512,512,256 --steps 30 --warmup 5 --no-cpu --no-parity > gpurun_out/r2_bench_thermo_n1b.json 2> gpurun_out/r2_bench_thermo_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_thermo_n1b.json').read().strip().split('\n')[-1])
print(d['value'], d['ms_per_step'], {k:round(v['ms'],4) for k,v in d['roofline']['families'].items()})"
