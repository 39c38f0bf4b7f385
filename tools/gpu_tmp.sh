timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for sh in 0 2; do G2S_THIN_SHAPE=$sh timeout 300 python tools/gpu_check_oracle.py cfg2 1.0 2>&1 | tail -2; done
timeout 300 python tools/gpu_check_oracle.py cfg3 0.3 2>&1 | tail -2
G2S_THIN_SHAPE=0 timeout 300 python tools/gpu_check_oracle.py cfg4 0.05 2>&1 | tail -2
timeout 300 python tools/gpu_check_oracle.py cfg5 0.1 2>&1 | tail -2
bash tools/gpu_variants.sh "base"
timeout 600 python bench.py --workload cfg4 --scale 1.0 --steps 3 --warmup 3 --cpu-seconds 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); t=d['phase_cycles_sum']; tot=t['total']
print('cfg4 full: kernel ms',round(d['ms_per_step'],3),'e2e ms',round(d['e2e']['ms_per_step'],3),d['config']['fills_equal_to_removed_bases'],{k:round(v/tot,3) for k,v in t.items() if not k.startswith('n_')})"
