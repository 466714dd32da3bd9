python -m pytest tests -q -x -m gpu 2>&1 | tail -8 > gpurun_out/r2_pytest_f.log
tail -8 gpurun_out/r2_pytest_f.log
python bench.py --steps 30 --warmup 5 > gpurun_out/r2_bench_n1_zseg.json 2> gpurun_out/r2_bench_n1_zseg.err
tail -c 1500 gpurun_out/r2_bench_n1_zseg.json
