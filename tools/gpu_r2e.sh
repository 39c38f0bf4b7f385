# Round 2: chain tier v2 (4 walks per warp, two-slot probes, filtered finish) on 2 GPUs: tests, N=1 / N=2 lines, ncu.
TAG=${1:-r2e}
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_multi.py tests/test_gpu_fullsize.py -m gpu -x -q --durations=5) > gpurun_out/${TAG}_pytest.log 2>&1
tail -10 gpurun_out/${TAG}_pytest.log
G2S_HOST_PROF=1 timeout 600 python bench.py --cpu-seconds 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
cut -c1-300 gpurun_out/${TAG}_bench_n1.json; grep "g2s_fill_records" gpurun_out/${TAG}_bench_n1.err | tail -2
G2S_HOST_PROF=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
cut -c1-300 gpurun_out/${TAG}_bench_n2.json; grep "g2s_fill_records\|GPU [01]:\|replicated" gpurun_out/${TAG}_bench_n2.err | tail -4; tail -3 gpurun_out/${TAG}_bench_n2.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chain -s 6 -c 2 -o gpurun_out/${TAG}_chain_cfg4 python bench.py --steps 1 --warmup 3 --cpu-seconds 0.1 > gpurun_out/${TAG}_ncu_chain_cfg4.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_chain_cfg4.log
