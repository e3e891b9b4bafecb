cd /root/repo
H264B_CABAC_LOOP=2 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest4_loop2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest4_loop2.log
tail -5 gpurun_out/r2_pytest4_loop2.log
EXP_VARIANTS=2:0:1,2:0:0 python tools/cabac_exp2.py > gpurun_out/r2_exp2_d.log 2>&1; echo "rc=$?"
cat gpurun_out/r2_exp2_d.log
export EXP_SLICES=18944 EXP_MEAN_BINS=30000 EXP_K=20000 EXP_ONLY=18944 EXP_VARIANTS=2:0:1
python tools/cabac_exp2.py > gpurun_out/r2_lone2_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:cabac_decode -c 2 -o gpurun_out/r2_lone_loop2c python tools/cabac_exp2.py > gpurun_out/r2_lone2_ncu.log 2>&1
cat gpurun_out/r2_lone2_plain.log; tail -3 gpurun_out/r2_lone2_ncu.log
