cd /root/repo
N=$1
python bench.py --config 4 --gpus $N --steps 2 --warmup 1 > gpurun_out/r2_c4_n$N.json 2> gpurun_out/r2_c4_n$N.err; echo "c4 N=$N rc=$?"; tail -2 gpurun_out/r2_c4_n$N.err
