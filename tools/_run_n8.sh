cd /root/repo
nvidia-smi topo -m > gpurun_out/r2_topo_n8.txt 2>&1
python bench.py --config 4 --gpus 8 --steps 2 --warmup 1 > gpurun_out/r2_c4_n8.json 2> gpurun_out/r2_c4_n8.err; echo "c4 rc=$?"; tail -2 gpurun_out/r2_c4_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_n8.err
