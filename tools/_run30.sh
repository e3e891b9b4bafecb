cd /root/repo
for v in d7 d8; do echo "--- $v"; H264B_LIB=/root/repo/h264decode_b200/build_lib_$v.so timeout 300 python tools/scan_dense_probe.py 2>&1 | tail -2 | head -1; done
echo "--- shipped"; timeout 300 python tools/scan_dense_probe.py 2>&1 | tail -2 | head -1
