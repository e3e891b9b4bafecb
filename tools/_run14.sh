cd /root/repo
EXP_VARIANTS=2:0:1,2:0:3 python tools/cabac_exp2.py > gpurun_out/r2_exp2_g.log 2>&1; cat gpurun_out/r2_exp2_g.log
EXP_SLICES=75776 EXP_MEAN_BINS=70000 EXP_K=70000 EXP_ONLY=75776 EXP_VARIANTS=2:0:3 python tools/cabac_exp2.py > gpurun_out/r2_ll_plain.log 2>&1 && \
EXP_SLICES=75776 EXP_MEAN_BINS=70000 EXP_K=70000 EXP_ONLY=75776 EXP_VARIANTS=2:0:3 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_ll.csv python tools/cabac_exp2.py > gpurun_out/r2_ll_ncu.log 2>&1
python tools/parse_launches.py gpurun_out/r2_ll.csv 2>/dev/null | tail -20 || tail -20 gpurun_out/r2_ll.csv
