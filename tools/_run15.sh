python bench.py > gpurun_out/r2_bench_final_n1.json 2> gpurun_out/r2_bench_final_n1.err
tail -c 200 gpurun_out/r2_bench_final_n1.err
python bench.py --impl reference > gpurun_out/r2_bench_final_reference_arm.json 2>> gpurun_out/r2_bench_final_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-parity --no-1024 > gpurun_out/r2_b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_momtend|k_closure|k_fillps|k_tderive|k_rfft|k_zsolve" -s 27 -c 9 -o gpurun_out/r2_substep_final -f python tools/ab_variants.py --steps 1 base > /dev/null 2>&1
python bench.py --workload thermo --grid 512,512,256 --steps 30 --warmup 5 --no-cpu --no-parity > gpurun_out/r2_bench_thermo_n1.json 2> gpurun_out/r2_bench_thermo_n1.err
tail -c 300 gpurun_out/r2_bench_thermo_n1.err
python bench.py --workload ibm --grid 512,512,256 --steps 30 --warmup 5 --no-cpu --no-parity > gpurun_out/r2_bench_ibm_n1b.json 2>> gpurun_out/r2_bench_thermo_n1.err
ls -la gpurun_out/r2_substep_final.ncu-rep
