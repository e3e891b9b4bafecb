set -x
cd /root/repo
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
tail -5 gpurun_out/r2_pytest1.log
python tools/cabac_exp2.py > gpurun_out/r2_exp2_a.log 2>&1; echo "rc=$?"
cat gpurun_out/r2_exp2_a.log
export EXP_SLICES=18944 EXP_MEAN_BINS=30000 EXP_K=20000 EXP_ONLY=18944 EXP_VARIANTS=1:0:1
python tools/cabac_exp2.py > gpurun_out/r2_lone_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:cabac_decode -c 2 -o gpurun_out/r2_lone_loop1 python tools/cabac_exp2.py > gpurun_out/r2_lone_ncu.log 2>&1
tail -3 gpurun_out/r2_lone_plain.log gpurun_out/r2_lone_ncu.log
