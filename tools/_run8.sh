cd /root/repo
python -m pytest tests/test_gpu_configs.py -m gpu -x -q > gpurun_out/r2_pytest9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest9.log
tail -15 gpurun_out/r2_pytest9.log
python bench.py --config 4 --gpus 1 --streams 512 --steps 2 --warmup 1 > gpurun_out/r2_c4b_512.json 2> gpurun_out/r2_c4b_512.err; echo "rc=$?"; tail -3 gpurun_out/r2_c4b_512.err; cut -c1-1800 gpurun_out/r2_c4b_512.json
python bench.py --config 4 --gpus 1 --steps 2 --warmup 1 > gpurun_out/r2_c4b_n1.json 2> gpurun_out/r2_c4b_n1.err; echo "rc=$?"; tail -3 gpurun_out/r2_c4b_n1.err; cut -c1-2500 gpurun_out/r2_c4b_n1.json
