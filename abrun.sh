timeout 900 python -m pytest tests/test_gpu_march.py tests/test_gpu_testbed.py -m gpu -q --timeout 900 > gpurun_out/t26.log 2>&1; echo EXIT $? >> gpurun_out/t26.log; tail -4 gpurun_out/t26.log
bash tests/golden/make_ref_tcnn_goldens.sh gpurun_out 2>&1 | tail -6
for rep in 1 2; do python bench.py --steps 200 --warmup 300 --no-cpu-baseline > gpurun_out/bench26.json 2> gpurun_out/bench26.err; python -c "
import json
d=json.loads(open('gpurun_out/bench26.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],4), round(d['value']/1e6,1), int(d['per_step']['rays']), d['phase_ms_per_step'], round(d['e2e']['value']/1e6,1))
"; done
