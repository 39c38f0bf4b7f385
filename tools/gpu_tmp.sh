G2S_HOST_PROF=1 timeout 600 python bench.py --workload cfg4 --scale 1.0 --steps 3 --warmup 3 --cpu-seconds 1 2>gpurun_out/t.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('cfg4 full: kernel ms',round(d['ms_per_step'],3),'e2e ms',round(d['e2e']['ms_per_step'],3),d['e2e']['output_identical_every_step'])"
grep "fill_records" gpurun_out/t.err | tail -2
G2S_HOST_PROF=1 timeout 600 python bench.py --steps 10 --warmup 3 --cpu-seconds 1 2>gpurun_out/t.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('cfg2: kernel ms',round(d['ms_per_step'],3),'e2e ms',round(d['e2e']['ms_per_step'],3),d['e2e']['output_identical_every_step'])"
grep "fill_records" gpurun_out/t.err | tail -2
