cd /root/repo
echo "=== shipped"; EXP_VARIANTS=2:0:3 python tools/cabac_exp2.py 2>&1 | tail -4
echo "=== no rounds loop"; H264B_LIB=/root/repo/h264decode_b200/build_lib_norounds.so EXP_VARIANTS=2:0:3 python tools/cabac_exp2.py 2>&1 | tail -4
