# Round 2, third GPU job: chain tier on the device. usage: gpurun --timeout 2400 -- 'bash tools/gpu_r2c.sh <tag>'
TAG=${1:-r2c}
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_gpu_chain.py tests/test_gpu_parity.py -m gpu -x -q --durations=5) > gpurun_out/${TAG}_pytest_chain.log 2>&1
tail -12 gpurun_out/${TAG}_pytest_chain.log
G2S_HOST_PROF=1 timeout 600 python bench.py > gpurun_out/${TAG}_bench_cfg4.json 2> gpurun_out/${TAG}_bench_cfg4.err
cut -c1-400 gpurun_out/${TAG}_bench_cfg4.json; grep -v "^g2s: pass\|subgraph" gpurun_out/${TAG}_bench_cfg4.err | tail -4
(time timeout 1500 python -m pytest tests -m gpu -x -q --durations=8) > gpurun_out/${TAG}_pytest.log 2>&1
tail -14 gpurun_out/${TAG}_pytest.log
for spec in "cfg2 1.0 10 3" "cfg5 1.0 5 3" "cfg1 1.0 10 3"; do
  set -- $spec
  G2S_HOST_PROF=1 timeout 600 python bench.py --workload $1 --scale $2 --steps $3 --warmup $4 --cpu-seconds 3 > gpurun_out/${TAG}_bench_$1.json 2> gpurun_out/${TAG}_bench_$1.err
  cut -c1-300 gpurun_out/${TAG}_bench_$1.json; tail -2 gpurun_out/${TAG}_bench_$1.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chain -s 6 -c 2 -o gpurun_out/${TAG}_chain_cfg4 python bench.py --steps 1 --warmup 3 --cpu-seconds 0.1 > gpurun_out/${TAG}_ncu_chain_cfg4.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_chain_cfg4.log
ls -la gpurun_out | grep ${TAG}
