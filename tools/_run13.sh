python tools/ab_variants.py --size 256 --steps 30 pf0:UDGPU_CLOSURE_PF=0 default > gpurun_out/r2_ab8b_closure_pf.jsonl 2> gpurun_out/r2_ab8.err
cut -c1-300 gpurun_out/r2_ab8b_closure_pf.jsonl
python tools/ab_variants.py --size 256 --steps 20 --nsv 4 spf0 spf1:UDGPU_SCALAR_PF=1 spf2:UDGPU_SCALAR_PF=2 spf3:UDGPU_SCALAR_PF=3 > gpurun_out/r2_ab9_scalar_pf.jsonl 2>> gpurun_out/r2_ab8.err
cut -c1-300 gpurun_out/r2_ab9_scalar_pf.jsonl
python -m pytest tests/test_gpu_parity.py -q -x -k "scalars or closure" 2>&1 | tail -3
