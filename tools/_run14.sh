python -m pytest tests -q -x -m gpu 2>&1 | tail -6 > gpurun_out/r2_pytest_g.log
tail -6 gpurun_out/r2_pytest_g.log
python -c "import __graft_entry__ as g; g.smoke()"
