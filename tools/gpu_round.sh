set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/v3_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/v3_bench.json 2> gpurun_out/v3_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/v3_bench_reference.json 2> gpurun_out/v3_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/v3_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/v3_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fill_warp -s 1 -c 1 -o gpurun_out/v3_fill_warp python bench.py --steps 2 --warmup 1 > gpurun_out/v3_ncu_full.log 2>&1
python bench.py --workload cfg3 --scale 0.1 --steps 3 --warmup 1 --cpu-seconds 5 > gpurun_out/v3_bench_cfg3_s01.json 2> gpurun_out/v3_bench_cfg3.err
tail -3 gpurun_out/v3_pytest.log; cat gpurun_out/v3_bench.json | cut -c1-600
