python -m pytest tests/test_thermo.py -q -x -m gpu 2>&1 | tail -12
