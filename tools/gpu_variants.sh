# usage: bash tools/gpu_variants.sh "<variant> <variant> ..."   ("base" = the default library)
mkdir -p gpurun_out
VARIANTS=$1
PYT=$2
for v in $VARIANTS; do
  if [ "$v" = base ]; then unset G2S_LIB; else export G2S_LIB=$PWD/gap2seq_b200/lib/libg2s_gpu_$v.so; fi
  echo "=== variant $v"
  if [ "$PYT" = "pytest" ]; then timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2; fi
  for spec in "cfg2 1.0 10 3" "cfg4 0.1 5 3"; do  # benches
    set -- $spec
    timeout 300 python bench.py --workload $1 --scale $2 --steps $3 --warmup $4 --cpu-seconds 1 > gpurun_out/var_${v}_$1.json 2> gpurun_out/var_${v}_$1.err || tail -5 gpurun_out/var_${v}_$1.err
    python - gpurun_out/var_${v}_$1.json <<'P'
import json,sys
d=json.load(open(sys.argv[1]))
s=d['phase_cycles_slowest_gap']
print(' ',d['config']['workload'][:4],'kernel ms',round(d['ms_per_step'],3),'| e2e ms',round(d['e2e']['ms_per_step'],3),d['e2e']['output_identical_every_step'],'| slowest gap Mcyc:',{k:round(v/1e6,2) for k,v in s.items() if not k.startswith('n_')})
P
  done
done
