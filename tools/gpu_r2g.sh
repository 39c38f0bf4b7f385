TAG=${1:-r2g}
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q --durations=6) > gpurun_out/${TAG}_pytest.log 2>&1
tail -12 gpurun_out/${TAG}_pytest.log
for spec in "cfg3 1.0 3 2" "cfg4 1.0 10 3" "cfg2 1.0 10 3"; do
  set -- $spec
  G2S_HOST_PROF=1 timeout 600 python bench.py --workload $1 --scale $2 --steps $3 --warmup $4 --cpu-seconds 3 > gpurun_out/${TAG}_bench_$1.json 2> gpurun_out/${TAG}_bench_$1.err
  cut -c1-250 gpurun_out/${TAG}_bench_$1.json; grep "pass tier\|chain tier\|g2s_fill_records" gpurun_out/${TAG}_bench_$1.err | tail -5
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fill_cta -s 2 -c 1 -o gpurun_out/${TAG}_cta_cfg3 python bench.py --workload cfg3 --steps 1 --warmup 2 --cpu-seconds 0.1 > gpurun_out/${TAG}_ncu_cta_cfg3.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_cta_cfg3.log
