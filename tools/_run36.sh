cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q -k "scheduler" > gpurun_out/r2_pytest_sched.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_sched.log
bash tools/_run34.sh 1
