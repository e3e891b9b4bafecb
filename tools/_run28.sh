cd /root/repo
for L in 0 2; do
echo "--- no rounds loop, loop $L"; H264B_LIB=/root/repo/h264decode_b200/build_lib_norounds.so H264B_CABAC_LOOP=$L BENCH_CONFIGS_SKIP_C4=1 timeout 600 python tools/bench_configs.py 2>&1 | grep -A2 'configs\[1\]'
done
