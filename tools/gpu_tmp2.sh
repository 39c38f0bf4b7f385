for v in base oldsite base oldsite; do
if [ "$v" = base ]; then unset G2S_LIB; else export G2S_LIB=$PWD/gap2seq_b200/lib/libg2s_gpu_$v.so; fi
python bench.py --workload cfg2 --scale 1.0 --steps 10 --warmup 3 --cpu-seconds 1 > gpurun_out/x_$v.json 2>/dev/null
python - gpurun_out/x_$v.json $v <<'P'
import json,sys
d=json.load(open(sys.argv[1])); t=d['phase_cycles_sum']; s=d['phase_cycles_slowest_gap']
print(sys.argv[2],'kernel ms',round(d['ms_per_step'],3),'sum M: layout',round(t['fwd_layout']/1e6),'scout',round(t['fwd_scout']/1e6),'node',round(t['fwd_node']/1e6),'general',round(t['fwd_general']/1e6),'forward',round(t['forward']/1e6),'| slowest layout',round(s['fwd_layout']/1e6,2),'fwd',round(s['forward']/1e6,2),'batches',int(t['n_narrow_batches']))
P
done
