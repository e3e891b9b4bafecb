cd /root/repo
for cfg in "64 64" "460 460" "460 1024"; do
  set -- $cfg
  EXP_ACTIVE=$1 EXP_NCTX=$2 EXP_VARIANTS=2:0:3 python tools/cabac_exp2.py >> gpurun_out/r2_exp2_nctx2.log 2>&1; echo "rc=$?"
done
cat gpurun_out/r2_exp2_nctx2.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
