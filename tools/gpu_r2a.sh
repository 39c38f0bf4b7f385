# Round 2, first GPU job: the new full-size parity tests, bench on the named workloads (hash counters), launch list and
# full ncu captures of the thin kernel on cfg4 and the wide-tier kernel on cfg3.
# usage: gpurun --timeout 2400 -- 'bash tools/gpu_r2a.sh <tag>'
TAG=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.csv
nproc > gpurun_out/${TAG}_nproc.txt
(time timeout 1500 python -m pytest tests -m gpu -x -q --durations=15) > gpurun_out/${TAG}_pytest.log 2>&1
tail -25 gpurun_out/${TAG}_pytest.log
for spec in "cfg2 1.0 10 3" "cfg4 1.0 5 3" "cfg3 1.0 3 2" "cfg5 0.2 5 3"; do
  set -- $spec
  G2S_HOST_PROF=1 timeout 600 python bench.py --workload $1 --scale $2 --steps $3 --warmup $4 --cpu-seconds 5 > gpurun_out/${TAG}_bench_$1.json 2> gpurun_out/${TAG}_bench_$1.err
  cut -c1-400 gpurun_out/${TAG}_bench_$1.json
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fill_thin -s 2 -c 1 -o gpurun_out/${TAG}_thin_cfg4 python bench.py --workload cfg4 --scale 1.0 --steps 1 --warmup 3 --cpu-seconds 0.1 > gpurun_out/${TAG}_ncu_thin_cfg4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fill_cta -s 2 -c 1 -o gpurun_out/${TAG}_cta_cfg3 python bench.py --workload cfg3 --scale 1.0 --steps 1 --warmup 2 --cpu-seconds 0.1 > gpurun_out/${TAG}_ncu_cta_cfg3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches_cfg3.csv python bench.py --workload cfg3 --scale 1.0 --steps 1 --warmup 1 --cpu-seconds 0.1 > /dev/null 2>&1
ls -la gpurun_out | grep ${TAG}
