# Exploration round on one B200: the named workloads through bench.py (bounded by timeouts).
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_explore.sh <tag> [pytest]'
TAG=${1:-x}
mkdir -p gpurun_out
if [ "$2" = "pytest" ]; then timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${TAG}_pytest.log; tail -3 gpurun_out/${TAG}_pytest.log; fi
for spec in "cfg2 1.0 10 3" "cfg4 0.1 5 3" "cfg5 0.2 5 3" "cfg3 1.0 3 1"; do
  set -- $spec
  G2S_HOST_PROF=1 timeout 300 python bench.py --workload $1 --scale $2 --steps $3 --warmup $4 --cpu-seconds 5 > gpurun_out/${TAG}_bench_$1_$2.json 2> gpurun_out/${TAG}_bench_$1_$2.err
  echo "== $spec rc=$?"; tail -2 gpurun_out/${TAG}_bench_$1_$2.err | cut -c1-400
  python - gpurun_out/${TAG}_bench_$1_$2.json <<'P'
import json,sys
d=json.load(open(sys.argv[1]))
print(' kernel ms',round(d['ms_per_step'],3),'gaps/s',round(d['value']),'cells/s %.3g'%d['dp_cells_per_s'],'| e2e ms',round(d['e2e']['ms_per_step'],3),'gaps/s',round(d['e2e']['value']),d['e2e']['output_identical_every_step'],{k:round(v,3) for k,v in d['e2e']['breakdown_ms_last_step'].items()})
P
done
