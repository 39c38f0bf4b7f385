TAG=${1:-r2f}
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_chain.py -m gpu -x -q) > gpurun_out/${TAG}_pytest.log 2>&1
tail -4 gpurun_out/${TAG}_pytest.log
G2S_HOST_PROF=1 timeout 600 python bench.py --cpu-seconds 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
cut -c1-300 gpurun_out/${TAG}_bench_n1.json; grep "g2s_fill_records" gpurun_out/${TAG}_bench_n1.err | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chain -s 6 -c 2 -o gpurun_out/${TAG}_chain_cfg4 python bench.py --steps 1 --warmup 3 --cpu-seconds 0.1 > gpurun_out/${TAG}_ncu_chain_cfg4.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_chain_cfg4.log
