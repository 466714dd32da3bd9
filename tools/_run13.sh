python tools/ab_variants.py --size 256 --steps 30 kc16 kc32:UDGPU_CLOSURE_KC=32 kc64:UDGPU_CLOSURE_KC=64 > gpurun_out/r2_ab15_closure_kc.jsonl 2> gpurun_out/r2_ab8.err
cut -c1-300 gpurun_out/r2_ab15_closure_kc.jsonl
