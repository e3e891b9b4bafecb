cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q -k "scheduler" > gpurun_out/r2_pytest_sched.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_sched.log
H264B_SCHED_TRACE=1 timeout 900 python bench.py --config 4 --gpus 1 --streams 512 --steps 2 --warmup 1 > gpurun_out/r2_c4_512_passes.json 2> gpurun_out/r2_c4_512_passes.err; echo "c4 rc=$?"; grep scheduler gpurun_out/r2_c4_512_passes.err | tail -3
python -c "
import json
for l in open('gpurun_out/r2_c4_512_passes.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['makespan_ms'], d['slice_completion_ms'], d['results_verified'], d['longest_slice_floor_ms'])"
bash tools/_run34.sh 1
