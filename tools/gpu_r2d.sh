# Round 2, 2-GPU job: multi-engine tests on two real GPUs, N=1 and N=2 bench lines, CLI with -gpus 2.
TAG=${1:-r2d}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
(time timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_chain.py -m gpu -x -q --durations=5) > gpurun_out/${TAG}_pytest_multi.log 2>&1
tail -12 gpurun_out/${TAG}_pytest_multi.log
G2S_HOST_PROF=1 timeout 600 python bench.py --cpu-seconds 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
cut -c1-300 gpurun_out/${TAG}_bench_n1.json; grep "g2s_fill_records" gpurun_out/${TAG}_bench_n1.err | tail -2
G2S_HOST_PROF=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
cut -c1-300 gpurun_out/${TAG}_bench_n2.json; grep "g2s_fill_records\|GPU [01]:\|replicated" gpurun_out/${TAG}_bench_n2.err | tail -6; tail -3 gpurun_out/${TAG}_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 --ref-gaps 200 > gpurun_out/${TAG}_bench_ref_n2.json 2> gpurun_out/${TAG}_bench_ref_n2.err
cut -c1-600 gpurun_out/${TAG}_bench_ref_n2.json
