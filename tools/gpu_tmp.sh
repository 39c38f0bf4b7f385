for mb in 16 1000; do echo "== G2S_FLUSH_MB=$mb"
G2S_FLUSH_MB=$mb G2S_HOST_PROF=1 python tools/gpu_hostthreads.py cfg4 0.1 2>&1 | grep "host_threads=16\|host_threads= 8\|fill_records" | tail -4
done
