mkdir -p gpurun_out
timeout 300 python tools/gpu_stress.py cfg3 1.0 60 2>&1 | tail -12
timeout 120 python tools/gpu_stress.py cfg3 0.3 100 2>&1 | tail -12
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
