timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for sh in 0 2; do G2S_THIN_SHAPE=$sh timeout 300 python tools/gpu_check_oracle.py cfg2 1.0 2>&1 | tail -2; done
timeout 300 python tools/gpu_check_oracle.py cfg3 0.3 2>&1 | tail -2
timeout 300 python tools/gpu_check_oracle.py cfg4 0.05 2>&1 | tail -2
bash tools/gpu_variants.sh "base"
