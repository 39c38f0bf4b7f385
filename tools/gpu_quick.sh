TAG=${1:-q}
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_chain.py -m gpu -x -q) 2>&1 | tail -2
for spec in "cfg4 1.0 10 3" "cfg5 1.0 5 3"; do
  set -- $spec
  G2S_HOST_PROF=1 timeout 600 python bench.py --workload $1 --scale $2 --steps $3 --warmup $4 --cpu-seconds 3 > gpurun_out/${TAG}_bench_$1.json 2> gpurun_out/${TAG}_bench_$1.err
  python - gpurun_out/${TAG}_bench_$1.json <<'P'
import json,sys
d=json.load(open(sys.argv[1])); print(d['config']['workload'][:5], 'ms', round(d['ms_per_step'],4), 'value', round(d['value']), 'e2e ms', round(d['e2e']['ms_per_step'],2), round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],4), d['e2e']['output_identical_every_step'], d['run']['records_checked_against_oracle'])
P
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chain -s 6 -c 2 -o gpurun_out/${TAG}_chain_cfg4 python bench.py --steps 1 --warmup 3 --cpu-seconds 0.1 > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i gpurun_out/${TAG}_chain_cfg4.ncu-rep --page details 2>/dev/null | grep -E "k_chain|Duration|Executed Ipc Active|Avg. Active Threads" | head -8
