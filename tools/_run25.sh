cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q -k "scan or annexb or stream or nal or golden or config" > gpurun_out/r2_pytest_v5.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_v4.log
timeout 300 python tools/scan_dense_probe.py 2>&1 | tail -2
timeout 900 python bench.py --steps 4 --warmup 3 --no-probe > gpurun_out/r2_bench_v5.json 2> gpurun_out/r2_bench_v5.err; echo "bench rc=$?"; python tools/show_bench.py gpurun_out/r2_bench_v5.json 2>/dev/null | head -3
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:'annexb|order_|nal_|scan_|chunk_' --log-file gpurun_out/r2_bench_scan_launches3.csv python bench.py --steps 2 --warmup 1 --no-probe --no-dense > /dev/null 2>&1; echo "ncu rc=$?"
