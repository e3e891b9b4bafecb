cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_v7.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_v7.log
BENCH_CONFIGS_SKIP_C4=1 timeout 600 python tools/bench_configs.py > gpurun_out/r2_configs_v7.txt 2>&1; echo "configs rc=$?"; cat gpurun_out/r2_configs_v7.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-probe > gpurun_out/r2_bench_v7.json 2> gpurun_out/r2_bench_v7.err; echo "bench rc=$?"; python tools/show_bench.py gpurun_out/r2_bench_v7.json 2>/dev/null | head -1
