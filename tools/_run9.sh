# launch list + DRAM traffic of one 256^3 substep with the final kernels, and the IBM config at N = 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_n1.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-parity --no-1024 > gpurun_out/r2_b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_momtend|k_closure|k_fillps|k_tderive|k_rfft|k_zsolve" -s 27 -c 9 -o gpurun_out/r2_substep_full python tools/ab_variants.py --steps 1 base > /dev/null 2>&1
python bench.py --workload ibm --grid 512,512,256 --steps 30 --warmup 5 --no-cpu > gpurun_out/r2_bench_ibm_n1.json 2> gpurun_out/r2_bench_ibm_n1.err
tail -c 300 gpurun_out/r2_bench_ibm_n1.err
python bench.py --workload scalars --steps 30 --warmup 5 --no-cpu > gpurun_out/r2_bench_scalars_n1.json 2>> gpurun_out/r2_bench_ibm_n1.err
ls -la gpurun_out/r2_substep_full.ncu-rep
