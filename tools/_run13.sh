python tools/ab_variants.py --size 256 --steps 30 unroll2 unroll2b > gpurun_out/r2_ab14_momtend_unroll2.jsonl 2> gpurun_out/r2_ab8.err
cut -c1-300 gpurun_out/r2_ab14_momtend_unroll2.jsonl
python -m pytest tests/test_gpu_parity.py -q -x -k "fused_advection_subgrid" 2>&1 | tail -2
