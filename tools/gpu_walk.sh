TAG=${1:-r2m}
mkdir -p gpurun_out
for ws in 8 32 16 4 2; do
  G2S_WALK_STRIDE=$ws G2S_HOST_PROF=1 timeout 300 python bench.py --steps 10 --warmup 3 --cpu-seconds 1 > gpurun_out/${TAG}_cfg4_ws$ws.json 2> gpurun_out/${TAG}_cfg4_ws$ws.err
  python - gpurun_out/${TAG}_cfg4_ws$ws.json $ws <<'P'
import json,sys
d=json.load(open(sys.argv[1])); print('stride', sys.argv[2], 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],2), d['e2e']['output_identical_every_step'], d['run']['records_checked_against_oracle'])
P
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chain -s 6 -c 2 -o gpurun_out/${TAG}_chain_cfg4 python bench.py --steps 1 --warmup 3 --cpu-seconds 0.1 > gpurun_out/${TAG}_ncu.log 2>&1
