cd /root/repo
timeout 900 ncu --set full --clock-control none --import-source on -k regex:annexb_dirty -s 3 -c 1 -o gpurun_out/r2_dense_dirty_v6 python tools/scan_dense_probe.py > gpurun_out/r2_dense_dirty_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2_dense_dirty_ncu.log
