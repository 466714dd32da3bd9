ncu --set full --clock-control none --import-source on -k regex:"k_rfft" -s 24 -c 4 -o gpurun_out/r2_xline -f python tools/ab_variants.py --size 256 --steps 2 V0 > gpurun_out/r2_ncu_xline.log 2>&1
tail -2 gpurun_out/r2_ncu_xline.log
