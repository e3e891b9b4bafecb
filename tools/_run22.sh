cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_img.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_pytest_img.log
timeout 300 python tools/scan_dense_probe.py > gpurun_out/r2_dense_img.log 2>&1; echo "dense rc=$?"; cat gpurun_out/r2_dense_img.log | tail -5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_dense_img_launches.csv python tools/scan_dense_probe.py > /dev/null 2>&1; echo "ncu rc=$?"
