# wide tier with clusters of 1 / 2 / 4 / 8 CTAs per gap on cfg3 (A/B), then the GPU parity tests with the automatic choice
TAG=${1:-r2l}
mkdir -p gpurun_out
for c in 1 2 4 8; do
  G2S_CLUSTER=$c G2S_HOST_PROF=1 timeout 300 python bench.py --workload cfg3 --steps 3 --warmup 2 --cpu-seconds 2 > gpurun_out/${TAG}_cfg3_c$c.json 2> gpurun_out/${TAG}_cfg3_c$c.err
  echo "== cluster $c"; grep "pass tier" gpurun_out/${TAG}_cfg3_c$c.err | tail -3; tail -2 gpurun_out/${TAG}_cfg3_c$c.err | cut -c1-300
  python - gpurun_out/${TAG}_cfg3_c$c.json <<'P'
import json,sys
try:
    d=json.load(open(sys.argv[1])); s=d['phase_cycles_slowest_gap']
    print('  ms', round(d['ms_per_step'],2), {k:round(v/1e6,2) for k,v in s.items() if not k.startswith('n_')}, d['e2e']['output_identical_every_step'], d['run']['records_checked_against_oracle'])
except Exception as e: print('  failed', e)
P
done
(time timeout 1200 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -x -q --durations=4) > gpurun_out/${TAG}_pytest.log 2>&1
tail -8 gpurun_out/${TAG}_pytest.log
