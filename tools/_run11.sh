python tools/ab_variants.py --size 256 --steps 30 stream:UDGPU_ZSEG=0 V0 V1:UDGPU_ZSEG_V=1 V2:UDGPU_ZSEG_V=2 V3:UDGPU_ZSEG_V=3 TW32V2:UDGPU_ZSEG_V=2,UDGPU_ZSEG_TW=32 V0again > gpurun_out/r2_ab7b_zseg.jsonl 2> gpurun_out/r2_ab7.err
cut -c1-330 gpurun_out/r2_ab7b_zseg.jsonl
