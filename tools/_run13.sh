python tools/ab_variants.py --size 256 --steps 30 zpf0 zpf1:UDGPU_ZSEG_PF=1 zpf2:UDGPU_ZSEG_PF=2 zpf3:UDGPU_ZSEG_PF=3 zpf4:UDGPU_ZSEG_PF=4 > gpurun_out/r2_ab10_zseg_pf.jsonl 2> gpurun_out/r2_ab8.err
cut -c1-300 gpurun_out/r2_ab10_zseg_pf.jsonl
