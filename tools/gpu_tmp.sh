grep -m1 "model name" /proc/cpuinfo; grep -c processor /proc/cpuinfo; grep -m1 -o "clflushopt" /proc/cpuinfo
for mb in 0 16 1000; do echo "== G2S_FLUSH_MB=$mb"
G2S_FLUSH_MB=$mb G2S_HOST_PROF=1 python tools/gpu_hostthreads.py cfg2 1.0 2>&1 | grep "host_threads=16\|fill_records" | tail -2
G2S_FLUSH_MB=$mb G2S_HOST_PROF=1 python tools/gpu_hostthreads.py cfg4 0.1 2>&1 | grep "host_threads=16\|fill_records" | tail -2
done
