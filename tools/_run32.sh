cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "scan" > gpurun_out/r2_pytest_carry.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_pytest_carry.log
timeout 1200 python bench.py --no-cpu > gpurun_out/r2_bench_e2e24.json 2> gpurun_out/r2_bench_e2e24.err; echo "bench rc=$?"; python tools/show_bench.py gpurun_out/r2_bench_e2e24.json | head -1; python -c "
import json
for l in open('gpurun_out/r2_bench_e2e24.json'):
    if l.startswith('{'): print(json.dumps(json.loads(l)['e2e'])[:900])"
