cd /root/repo
echo "--- r1 library"; H264B_LIB=/root/repo/h264decode_b200/build_lib_r1.so BENCH_CONFIGS_SKIP_C4=1 timeout 600 python tools/bench_configs.py 2>&1 | grep -A2 'configs\[1\]'
echo "--- current, loop 0, carveout default"; H264B_CABAC_CARVEOUT=-1 H264B_CABAC_LOOP=0 BENCH_CONFIGS_SKIP_C4=1 timeout 600 python tools/bench_configs.py 2>&1 | grep -A2 'configs\[1\]'
echo "--- current, loop 0, W=2"; H264B_CABAC_W=2 H264B_CABAC_LOOP=0 BENCH_CONFIGS_SKIP_C4=1 timeout 600 python tools/bench_configs.py 2>&1 | grep -A2 'configs\[1\]'
echo "--- current, loop 0, map 0"; H264B_CABAC_MAP=0 H264B_CABAC_LOOP=0 BENCH_CONFIGS_SKIP_C4=1 timeout 600 python tools/bench_configs.py 2>&1 | grep -A2 'configs\[1\]'
