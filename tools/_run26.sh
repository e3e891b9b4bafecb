cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_v6.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_v6.log
timeout 300 python tools/scan_dense_probe.py 2>&1 | tail -2
BENCH_CONFIGS_SKIP_C4=1 timeout 600 python tools/bench_configs.py > gpurun_out/r2_configs_v6.txt 2>&1; echo "configs rc=$?"; cat gpurun_out/r2_configs_v6.txt
echo "--- loop 0"; H264B_CABAC_LOOP=0 BENCH_CONFIGS_SKIP_C4=1 timeout 600 python tools/bench_configs.py 2>&1 | grep -A2 'configs\[1\]'
echo "--- loop 1"; H264B_CABAC_LOOP=1 BENCH_CONFIGS_SKIP_C4=1 timeout 600 python tools/bench_configs.py 2>&1 | grep -A2 'configs\[1\]'
