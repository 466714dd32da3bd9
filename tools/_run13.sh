python tools/ab_variants.py --size 256 --steps 30 base p25:UDGPU_L2PERSIST=0.25 p50:UDGPU_L2PERSIST=0.5 p75:UDGPU_L2PERSIST=0.75 p100:UDGPU_L2PERSIST=1.0 > gpurun_out/r2_ab13_l2persist.jsonl 2> gpurun_out/r2_ab13.err
cut -c1-300 gpurun_out/r2_ab13_l2persist.jsonl; grep udgpu gpurun_out/r2_ab13.err | head -3
