cd /root/repo
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1200 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_final.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_ncu_launch_list_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-probe --no-dense --verify-slices 8 > gpurun_out/r2_ll.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cabac_decode -s 3 -c 1 -o gpurun_out/r2_ncu_cabac_final python bench.py --steps 2 --warmup 1 --no-cpu --no-probe --no-dense --verify-slices 8 > gpurun_out/r2_ncu_cabac_final.log 2>&1; echo "cabac capture rc=$?"
