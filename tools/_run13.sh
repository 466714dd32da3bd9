python -m pytest tests/test_gpu_parity.py -q -x -k "line_local or poisson or segmented" 2>&1 | tail -4
python tools/ab_variants.py --size 256 --steps 30 xline0:UDGPU_XLINE=0 xline1 xline1b > gpurun_out/r2_ab12_xline_pairs.jsonl 2> gpurun_out/r2_ab8.err
cut -c1-300 gpurun_out/r2_ab12_xline_pairs.jsonl
