# wide-tier iteration job: parity of the tiers that use the general level step, then the cfg3 / cfg2 bench lines
TAG=${1:-w0}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -x -q) > gpurun_out/${TAG}_pytest.log 2>&1
tail -4 gpurun_out/${TAG}_pytest.log
for spec in "cfg3 1.0 3 2" "cfg2 1.0 10 3"; do
  set -- $spec
  G2S_HOST_PROF=1 timeout 600 python bench.py --workload $1 --scale $2 --steps $3 --warmup $4 --cpu-seconds 1 > gpurun_out/${TAG}_bench_$1.json 2> gpurun_out/${TAG}_bench_$1.err
  python - gpurun_out/${TAG}_bench_$1.json <<'P'
import json,sys
d=json.load(open(sys.argv[1])); print(d['config']['workload'][:5], 'ms', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],2), d['roofline']['kernel'], d['roofline']['tier_ms_per_step'], d['e2e']['output_identical_every_step'], d['run']['records_checked_against_oracle'])
print(' slowest', {k:round(v/1e6,2) for k,v in d['phase_cycles_slowest_gap'].items() if v>1e5}); print(' sum', {k:round(v/1e6,1) for k,v in d['phase_cycles_sum'].items() if v>1e5}); print(' e2e', d['e2e']['breakdown_ms_last_step'])
P
done
