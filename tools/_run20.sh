cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_lb.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_lb.log
timeout 300 python tools/scan_dense_probe.py > gpurun_out/r2_dense_lb.log 2>&1; echo "dense rc=$?"; cat gpurun_out/r2_dense_lb.log | tail -5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_dense_lb_launches.csv python tools/scan_dense_probe.py > /dev/null 2>&1; echo "ncu rc=$?"
