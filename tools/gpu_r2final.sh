# Round 2, closing job on 2 GPUs: the whole -m gpu suite, bench lines of every named workload on one GPU, the 2-GPU line,
# the reference arm, ncu captures (chain kernels on cfg4, launch list).
TAG=${1:-r2z}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_smi.csv; nproc > gpurun_out/${TAG}_nproc.txt
(time timeout 1500 python -m pytest tests -m gpu -x -q --durations=5) > gpurun_out/${TAG}_pytest.log 2>&1
tail -9 gpurun_out/${TAG}_pytest.log
for spec in "cfg4 1.0 10 3" "cfg5 1.0 5 3" "cfg1 1.0 10 3" "cfg2 1.0 10 3" "cfg3 1.0 3 2"; do
  set -- $spec
  G2S_HOST_PROF=1 timeout 600 python bench.py --workload $1 --scale $2 --steps $3 --warmup $4 > gpurun_out/${TAG}_bench_$1.json 2> gpurun_out/${TAG}_bench_$1.err
  python - gpurun_out/${TAG}_bench_$1.json <<'P'
import json,sys
d=json.load(open(sys.argv[1])); print(d['config']['workload'][:5], 'ms', round(d['ms_per_step'],4), 'value', round(d['value']), 'e2e ms', round(d['e2e']['ms_per_step'],2), round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],4), d['roofline']['kernel'], d['e2e']['output_identical_every_step'], d['run']['records_checked_against_oracle'], 'cpu', round(d['cpu_baseline']['value']))
P
done
G2S_HOST_PROF=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
python - gpurun_out/${TAG}_bench_n2.json <<'P'
import json,sys
d=json.load(open(sys.argv[1])); print('N=2', 'ms', round(d['ms_per_step'],4), 'value', round(d['value']), 'e2e ms', round(d['e2e']['ms_per_step'],2), round(d['e2e']['value']), d['e2e']['output_identical_every_step'], d['run']['strong_scaling_named_workload'])
P
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
cut -c1-900 gpurun_out/${TAG}_bench_reference.json
timeout 300 python tools/own_graphs_bench.py cfg2 1.0 > gpurun_out/${TAG}_own_graphs_cfg2.json 2> gpurun_out/${TAG}_own_graphs_cfg2.err; cut -c1-400 gpurun_out/${TAG}_own_graphs_cfg2.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chain -s 6 -c 2 -o gpurun_out/${TAG}_chain_cfg4 python bench.py --steps 1 --warmup 3 --cpu-seconds 0.1 > gpurun_out/${TAG}_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches_cfg4.csv python bench.py --steps 2 --warmup 1 --cpu-seconds 0.1 > /dev/null 2>&1
ncu -i gpurun_out/${TAG}_chain_cfg4.ncu-rep --page details 2>/dev/null | grep -E "k_chain|Duration" | head -4
