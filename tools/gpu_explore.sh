# Exploration round on one B200: parity tests, then the larger named workloads (bounded by timeouts).
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_explore.sh <tag>'
TAG=${1:-x}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | head -2
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${TAG}_pytest.log; tail -3 gpurun_out/${TAG}_pytest.log
for spec in "cfg2 1.0 10 3" "cfg4 0.1 5 3" "cfg5 0.2 5 3" "cfg3 0.3 2 1" "cfg3 1.0 2 1"; do
  set -- $spec
  timeout 420 python bench.py --workload $1 --scale $2 --steps $3 --warmup $4 --cpu-seconds 5 > gpurun_out/${TAG}_bench_$1_$2.json 2> gpurun_out/${TAG}_bench_$1_$2.err
  echo "== $spec rc=$?"; tail -4 gpurun_out/${TAG}_bench_$1_$2.err; cut -c1-900 gpurun_out/${TAG}_bench_$1_$2.json
done
