timeout 600 python -m pytest tests/test_gpu_march.py -m gpu -q --timeout 600 > gpurun_out/t29.log 2>&1; echo EXIT $? >> gpurun_out/t29.log; tail -3 gpurun_out/t29.log
for rep in 1 2 3; do for pt in 700; do python bench.py --steps 200 --warmup 300 --pretrain $pt --no-cpu-baseline > gpurun_out/bench29.json 2> gpurun_out/bench29.err; python -c "
import json
d=json.loads(open('gpurun_out/bench29.json').read().strip().splitlines()[-1])
p=d['phase_ms_per_step']
print($pt, round(d['ms_per_step'],4), round(d['value']/1e6,1), int(d['per_step']['rays']), int(d['per_step']['samples_before_compaction']), 'gen %.3f inf %.3f loss %.3f fb %.3f'%(p['sample_generation'],p['inference'],p['loss_compaction'],p['forward_backward']), round(d['e2e']['value']/1e6,1))
"; done; done
