cd /root/repo
H264B_TRACE=1 timeout 600 python tools/e2e_trace.py > gpurun_out/r2_e2e_trace.txt 2> gpurun_out/r2_e2e_trace.err; echo rc=$?
grep 'trace' gpurun_out/r2_e2e_trace.err | tail -12; tail -14 gpurun_out/r2_e2e_trace.txt
