timeout 600 python -m pytest tests/test_gpu_field.py -m gpu -q --timeout 600 -k "module" > gpurun_out/t40.log 2>&1; echo EXIT $? >> gpurun_out/t40.log; tail -3 gpurun_out/t40.log
timeout 600 python tools/bench_fields.py > gpurun_out/bench_fields.jsonl 2> gpurun_out/bench_fields.err; cat gpurun_out/bench_fields.jsonl | cut -c1-420; tail -3 gpurun_out/bench_fields.err
python bench.py --steps 100 --warmup 300 --no-cpu-baseline > gpurun_out/bench40.json 2> gpurun_out/bench40.err; python -c "
import json
d=json.loads(open('gpurun_out/bench40.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d.get('render'), d.get('quality'))
"; tail -3 gpurun_out/bench40.err
