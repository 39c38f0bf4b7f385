TAG=${1:-r2n}
mkdir -p gpurun_out
G2S_HOST_PROF=1 timeout 600 python bench.py --cpu-seconds 2 > gpurun_out/${TAG}_bench_cfg4.json 2> gpurun_out/${TAG}_bench_cfg4.err
grep "g2s_fill_records" gpurun_out/${TAG}_bench_cfg4.err | tail -4
python - gpurun_out/${TAG}_bench_cfg4.json <<'P'
import json,sys
d=json.load(open(sys.argv[1])); print('ms', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],2), round(d['e2e']['value']), d['e2e']['output_identical_every_step'], d['run']['records_checked_against_oracle'])
P
(time timeout 900 python -m pytest tests/test_gpu_cli.py tests/test_gpu_parity.py tests/test_gpu_multi.py -m gpu -x -q) > gpurun_out/${TAG}_pytest.log 2>&1
tail -4 gpurun_out/${TAG}_pytest.log
