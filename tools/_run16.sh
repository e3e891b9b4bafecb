cd /root/repo
for c in 50 100 25; do echo "=== carveout $c"; H264B_CABAC_CARVEOUT=$c EXP_VARIANTS=2:0:3 python tools/cabac_exp2.py 2>&1 | tail -5; done
echo "=== 460 active of 1024 rows with n_ctx_used"; EXP_ACTIVE=460 EXP_NCTX=1024 EXP_VARIANTS=2:0:3 python tools/cabac_exp2.py 2>&1 | tail -5
