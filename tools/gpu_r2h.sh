TAG=${1:-r2h}
mkdir -p gpurun_out
for v in base t1024_ws0_wl0 t1024_ws1_wl0 t512_ws1_wl0 t512_ws1_wl1 t256_ws1_wl1 t256_ws1_wl0; do
  if [ "$v" = base ]; then unset G2S_LIB; else export G2S_LIB=$PWD/gap2seq_b200/lib/libg2s_gpu_$v.so; fi
  G2S_HOST_PROF=1 timeout 300 python bench.py --workload cfg3 --steps 3 --warmup 2 --cpu-seconds 1 > gpurun_out/${TAG}_cfg3_$v.json 2> gpurun_out/${TAG}_cfg3_$v.err
  echo "== $v"; grep "pass tier" gpurun_out/${TAG}_cfg3_$v.err | tail -3
  python - gpurun_out/${TAG}_cfg3_$v.json <<'P'
import json,sys
d=json.load(open(sys.argv[1])); s=d['phase_cycles_slowest_gap']
print('  ms', round(d['ms_per_step'],2), {k:round(v/1e6,2) for k,v in s.items() if not k.startswith('n_')}, d['e2e']['output_identical_every_step'], d['run']['records_checked_against_oracle'])
P
done
unset G2S_LIB
(time timeout 900 python -m pytest tests/test_gpu_form2.py -m gpu -x -q --durations=5) > gpurun_out/${TAG}_pytest_form2.log 2>&1
tail -12 gpurun_out/${TAG}_pytest_form2.log
