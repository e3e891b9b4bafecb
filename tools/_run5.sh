cd /root/repo
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest6.log
tail -5 gpurun_out/r2_pytest6.log
EXP_VARIANTS=2:0:3,2:0:1 python tools/cabac_exp2.py > gpurun_out/r2_exp2_f.log 2>&1; echo "rc=$?"
cat gpurun_out/r2_exp2_f.log
python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; tail -c 1500 gpurun_out/r2_bench_b.json; tail -3 gpurun_out/r2_bench_b.err
