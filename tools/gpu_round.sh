# One GPU round: parity tests, bench (both arms), launch list, one full ncu capture of the thin-tier kernel.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag> [full]'
TAG=${1:-v}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log
G2S_HOST_PROF=1 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
if [ "$2" = "full" ]; then
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${TAG}_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fill_thin -s 1 -c 1 -o gpurun_out/${TAG}_fill_thin python bench.py --steps 2 --warmup 1 > gpurun_out/${TAG}_ncu_full.log 2>&1
python bench.py --workload cfg3 --scale 1.0 --steps 3 --warmup 1 --cpu-seconds 5 > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err
python bench.py --workload cfg4 --scale 0.1 --steps 5 --warmup 3 --cpu-seconds 5 > gpurun_out/${TAG}_bench_cfg4_scale0.1.json 2> gpurun_out/${TAG}_bench_cfg4.err
python bench.py --workload cfg5 --scale 0.2 --steps 5 --warmup 3 --cpu-seconds 5 > gpurun_out/${TAG}_bench_cfg5_scale0.2.json 2> gpurun_out/${TAG}_bench_cfg5.err
python __graft_entry__.py smoke 2>&1 | tail -2
fi
tail -3 gpurun_out/${TAG}_pytest.log; tail -4 gpurun_out/${TAG}_bench.err; cut -c1-600 gpurun_out/${TAG}_bench.json; cut -c1-400 gpurun_out/${TAG}_bench_reference.json
if [ "$2" = "full" ]; then
timeout 900 python bench.py --workload cfg4 --scale 1.0 --steps 3 --warmup 3 --cpu-seconds 10 > gpurun_out/${TAG}_bench_cfg4_full.json 2> gpurun_out/${TAG}_bench_cfg4_full.err
cut -c1-300 gpurun_out/${TAG}_bench_cfg4_full.json
fi
