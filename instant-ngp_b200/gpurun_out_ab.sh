set -x
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/t13.log 2>&1; echo EXIT $? >> gpurun_out/t13.log
for n in 5 4 3; do
  sed -i "s/constexpr uint32_t FWD_RAYS_CTAS_PER_SM = [0-9]*;/constexpr uint32_t FWD_RAYS_CTAS_PER_SM = $n;/" instant-ngp_b200/csrc/render.cu
  python -c "import sys; sys.path.insert(0,'instant-ngp_b200'); import build; build.build()" > gpurun_out/build_$n.log 2>&1
  python bench.py --steps 100 --warmup 300 --no-cpu-baseline > gpurun_out/bench_ctas$n.json 2> gpurun_out/bench_ctas$n.err
done
tail -3 gpurun_out/t13.log
for n in 5 4 3; do python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_ctas$n.json').read().strip().splitlines()[-1])
print($n, d['ms_per_step'], d['value'], d.get('phase_ms_per_step'), d['e2e']['value'])
"; done
