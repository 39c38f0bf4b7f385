TAG=${1:-w7}
(timeout 900 python -m pytest tests/test_gpu_own_graphs.py tests/test_h5_import.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15)
for c in 2 1; do
G2S_CLUSTER=$c G2S_HOST_PROF=1 timeout 300 python bench.py --workload cfg3 --steps 3 --warmup 2 --cpu-seconds 1 2> gpurun_out/${TAG}_c$c.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cluster $c ms', round(d['ms_per_step'],3), d['run']['records_checked_against_oracle'], {k:round(v/1e6,3) for k,v in d['phase_cycles_slowest_gap'].items() if v>1e3}); print({k:round(v/1e6,1) for k,v in d['phase_cycles_sum'].items() if v>1e5})"
done
timeout 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --cpu-seconds 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cfg2 ms', round(d['ms_per_step'],4), d['run']['records_checked_against_oracle'])"
