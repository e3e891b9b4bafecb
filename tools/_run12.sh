cd /root/repo
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest12.log
tail -12 gpurun_out/r2_pytest12.log
python tools/mb_type_exp.py > gpurun_out/r2_mb_type_exp.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2_mb_type_exp.log
