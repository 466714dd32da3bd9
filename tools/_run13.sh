python -m pytest tests/test_thermo.py -q -x -m gpu 2>&1 | tail -30 > gpurun_out/r2_thermo_pytest.log
tail -30 gpurun_out/r2_thermo_pytest.log
