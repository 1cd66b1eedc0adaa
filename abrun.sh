timeout 900 python -m pytest tests/test_gpu_march.py tests/test_gpu_testbed.py -m gpu -q --timeout 900 -s > gpurun_out/t37.log 2>&1; echo EXIT $? >> gpurun_out/t37.log; grep -a "lane waste\|psnr from disk\|srgb epilogue" gpurun_out/t37.log; tail -4 gpurun_out/t37.log
for f in "" "--no-sort" "" "--no-sort"; do python bench.py --steps 200 --warmup 300 --no-cpu-baseline --no-render $f > gpurun_out/bench37.json 2> gpurun_out/bench37.err; python -c "
import json
d=json.loads(open('gpurun_out/bench37.json').read().strip().splitlines()[-1])
p=d['phase_ms_per_step']
print('$f', round(d['ms_per_step'],4), round(d['value']/1e6,1), int(d['per_step']['rays']), int(d['per_step']['samples_before_compaction']), 'gen %.3f inf %.3f loss %.3f fb %.3f'%(p['sample_generation'],p['inference'],p['loss_compaction'],p['forward_backward']), round(d['e2e']['value']/1e6,1))
"; done
