cd /root/repo
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest5.log
tail -5 gpurun_out/r2_pytest5.log
EXP_VARIANTS=2:0:1,1:0:1 python tools/cabac_exp2.py > gpurun_out/r2_exp2_e.log 2>&1; echo "rc=$?"
cat gpurun_out/r2_exp2_e.log
python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; tail -c 1500 gpurun_out/r2_bench_a.json; tail -3 gpurun_out/r2_bench_a.err
