ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_thermo.csv python bench.py --workload thermo --grid 256,256,256 --steps 3 --warmup 3 --no-cpu --no-parity --no-1024 > gpurun_out/r2_b_ncu_thermo.log 2>&1
tail -2 gpurun_out/r2_b_ncu_thermo.log | cut -c1-300
