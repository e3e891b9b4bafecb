cd /root/repo
N=$1
timeout 1500 python bench.py --config 4 --gpus $N --steps 2 --warmup 1 > gpurun_out/r2_c4p_n$N.json 2> gpurun_out/r2_c4p_n$N.err; echo "c4 N=$N rc=$?"; tail -2 gpurun_out/r2_c4p_n$N.err
python -c "
import json
for l in open('gpurun_out/r2_c4p_n$N.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['makespan_ms'], d['slice_completion_ms'], d['device_busy_ms'], d['results_verified'], d['longest_slice_floor_ms'], d['device_jobs'])"
