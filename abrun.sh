timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/t36.log 2>&1; echo EXIT $? >> gpurun_out/t36.log; tail -12 gpurun_out/t36.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 3000 gpurun_out/bench_default.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1200 gpurun_out/bench_ref.json
