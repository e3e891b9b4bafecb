cd /root/repo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:'annexb|order_|nal_|scan_|chunk_|slice_select' --log-file gpurun_out/r2_bench_scan_launches.csv python bench.py --steps 2 --warmup 1 --no-probe --no-dense > /dev/null 2>&1; echo "ncu rc=$?"
