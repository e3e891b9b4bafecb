cd /root/repo
timeout 300 python tools/scan_dense_probe.py > gpurun_out/r2_dense_v3.log 2>&1; echo "dense rc=$?"; tail -2 gpurun_out/r2_dense_v3.log
BENCH_CONFIGS_SKIP_C4=1 timeout 600 python tools/bench_configs.py > gpurun_out/r2_configs_v3.txt 2>&1; echo "configs rc=$?"; cat gpurun_out/r2_configs_v3.txt
timeout 900 python bench.py --steps 4 --warmup 3 --no-probe > gpurun_out/r2_bench_v3.json 2> gpurun_out/r2_bench_v3.err; echo "bench rc=$?"; python tools/show_bench.py gpurun_out/r2_bench_v3.json 2>/dev/null | head -40 || tail -3 gpurun_out/r2_bench_v3.err
