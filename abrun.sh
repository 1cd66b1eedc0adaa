timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_field.py tests/test_gpu_render.py -m gpu -q --timeout 600 > gpurun_out/t43.log 2>&1; echo EXIT $? >> gpurun_out/t43.log; tail -3 gpurun_out/t43.log
for f in "" ""; do python bench.py --steps 200 --warmup 300 --no-cpu-baseline $f > gpurun_out/bench43.json 2> gpurun_out/bench43.err; python -c "
import json
d=json.loads(open('gpurun_out/bench43.json').read().strip().splitlines()[-1])
p=d['phase_ms_per_step']
print('$f', round(d['ms_per_step'],4), round(d['value']/1e6,1), int(d['per_step']['rays']), 'gen %.3f inf %.3f loss %.3f fb %.3f opt %.3f'%(p['sample_generation'],p['inference'],p['loss_compaction'],p['forward_backward'],p['optimizer']), round(d['e2e']['value']/1e6,1), d['render']['ms_per_frame'])
"; done
