CS="compute-sanitizer --target-processes all --error-exitcode 7 --print-limit 20"
( timeout 600 $CS --tool racecheck python -m pytest tests/test_gpu_parity.py -x -q -k "(line_local and (shape0 or shape2 or shape3)) or (segmented_z and shape0 and L16) or (segmented_z and shape4 and L8)" 2>&1 | tail -12 ) > gpurun_out/r2_sanitizer_racecheck_round2_kernels.txt 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2_sanitizer_racecheck_round2_kernels.txt
( timeout 600 $CS --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_thermo.py -x -q -k "(line_local and (shape1 or shape4 or shape5 or shape6)) or (segmented_z and (shape2 or shape4)) or (thermo_matches_reference and 0-) or (substeps_track_oracle and 0-shape0 and (ibm or buoycorr or masscorr))" 2>&1 | tail -12 ) > gpurun_out/r2_sanitizer_memcheck_round2_kernels.txt 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2_sanitizer_memcheck_round2_kernels.txt
tail -6 gpurun_out/r2_sanitizer_racecheck_round2_kernels.txt gpurun_out/r2_sanitizer_memcheck_round2_kernels.txt
