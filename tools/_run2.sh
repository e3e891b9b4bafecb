cd /root/repo
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest2.log
tail -5 gpurun_out/r2_pytest2.log
H264B_CABAC_LOOP=2 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest2_loop2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest2_loop2.log
tail -5 gpurun_out/r2_pytest2_loop2.log
python tools/cabac_exp2.py > gpurun_out/r2_exp2_b.log 2>&1; echo "rc=$?"
cat gpurun_out/r2_exp2_b.log
export EXP_SLICES=18944 EXP_MEAN_BINS=30000 EXP_K=20000 EXP_ONLY=18944 EXP_VARIANTS=2:0:1
python tools/cabac_exp2.py > gpurun_out/r2_lone2_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:cabac_decode -c 2 -o gpurun_out/r2_lone_loop2 python tools/cabac_exp2.py > gpurun_out/r2_lone2_ncu.log 2>&1
cat gpurun_out/r2_lone2_plain.log; tail -3 gpurun_out/r2_lone2_ncu.log
