mkdir -p gpurun_out
SECONDS=0
timeout 1000 python bench.py --workload cfg4 --scale 1.0 --steps 3 --warmup 3 --cpu-seconds 10 > gpurun_out/cfg4_full.json 2> gpurun_out/cfg4_full.err
echo "rc=$? after ${SECONDS}s"
tail -5 gpurun_out/cfg4_full.err | cut -c1-300; cut -c1-2500 gpurun_out/cfg4_full.json
