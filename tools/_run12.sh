ncu --set full --clock-control none --import-source on -k regex:"k_zsolve" -s 6 -c 3 -o gpurun_out/r2_zseg -f python tools/ab_variants.py --size 256 --steps 2 V0 > gpurun_out/r2_ncu_zseg.log 2>&1
UDGPU_ZSEG=0 ncu --set full --clock-control none -k regex:"k_zsolve" -s 6 -c 2 -o gpurun_out/r2_zstream -f python tools/ab_variants.py --size 256 --steps 2 V0 >> gpurun_out/r2_ncu_zseg.log 2>&1
tail -3 gpurun_out/r2_ncu_zseg.log
