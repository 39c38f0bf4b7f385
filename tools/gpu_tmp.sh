G2S_HOST_PROF=1 python tools/gpu_hostthreads.py cfg4 0.1 2>&1 | grep "host_threads=16\|fill_records" | tail -4
