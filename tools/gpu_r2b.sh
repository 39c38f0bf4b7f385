# Round 2, second GPU job: -m gpu (incl. the multi-engine tests), the default bench line (cfg4), cfg2 / cfg3 lines.
# usage: gpurun --timeout 2400 -- 'bash tools/gpu_r2b.sh <tag>'
TAG=${1:-r2b}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.csv
nproc > gpurun_out/${TAG}_nproc.txt
(time timeout 1500 python -m pytest tests -m gpu -x -q --durations=10) > gpurun_out/${TAG}_pytest.log 2>&1
tail -18 gpurun_out/${TAG}_pytest.log
G2S_HOST_PROF=1 timeout 600 python bench.py > gpurun_out/${TAG}_bench_cfg4.json 2> gpurun_out/${TAG}_bench_cfg4.err
cut -c1-600 gpurun_out/${TAG}_bench_cfg4.json; tail -3 gpurun_out/${TAG}_bench_cfg4.err
for spec in "cfg2 1.0 10 3" "cfg3 1.0 3 2"; do
  set -- $spec
  G2S_HOST_PROF=1 timeout 600 python bench.py --workload $1 --scale $2 --steps $3 --warmup $4 --cpu-seconds 5 > gpurun_out/${TAG}_bench_$1.json 2> gpurun_out/${TAG}_bench_$1.err
  cut -c1-300 gpurun_out/${TAG}_bench_$1.json; tail -2 gpurun_out/${TAG}_bench_$1.err
done
