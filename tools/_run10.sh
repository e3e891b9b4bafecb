cd /root/repo
H264B_SCHED_TRACE=1 python bench.py --config 4 --gpus 1 --streams 512 --steps 2 --warmup 1 > gpurun_out/r2_c4d_512.json 2> gpurun_out/r2_c4d_512.err; echo "rc=$?"; tail -16 gpurun_out/r2_c4d_512.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_c4d_512.json').read().strip().splitlines()[-1])
print(d["makespan_ms"], d["slice_completion_ms"], d["results_verified"])
PY
