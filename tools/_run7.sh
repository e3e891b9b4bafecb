cd /root/repo
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest7.log
tail -4 gpurun_out/r2_pytest7.log
python bench.py --steps 4 --warmup 3 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; echo "bench rc=$?"
tail -5 gpurun_out/r2_bench_c.err
python tools/show_bench.py gpurun_out/r2_bench_c.json 2>/dev/null | head -60 || head -c 3000 gpurun_out/r2_bench_c.json
python bench.py --steps 2 --warmup 1 --no-cpu --no-probe --no-dense --verify-slices 8 > gpurun_out/r2_bench_plain.json 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cabac_decode -s 3 -c 1 -o gpurun_out/r2_ncu_cabac python bench.py --steps 2 --warmup 1 --no-cpu --no-probe --no-dense --verify-slices 8 > gpurun_out/r2_ncu_cabac.log 2>&1
tail -2 gpurun_out/r2_ncu_cabac.log
