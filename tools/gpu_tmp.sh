timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/gpu_check_oracle.py cfg3 1.0 2>&1 | tail -2
timeout 300 python tools/gpu_check_oracle.py cfg2 1.0 2>&1 | tail -2
G2S_HOST_PROF=1 timeout 200 python bench.py --workload cfg3 --scale 1.0 --steps 3 --warmup 1 --cpu-seconds 1 2>&1 >/dev/null | grep "g2s: pass\|fill_records" | tail -3
bash tools/gpu_variants.sh "base"
