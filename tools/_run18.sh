python -m pytest tests/test_gpu_villain.py -x -q -k "inplace or config5 or swapping or tiled" > gpurun_out/r2_tests_i.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests_i.log
tail -3 gpurun_out/r2_tests_i.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --configs c2,c5 > gpurun_out/r2_bench_i.json 2> gpurun_out/r2_bench_i.err; tail -3 gpurun_out/r2_bench_i.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_i.json'))
for c in d['configs']:
    print(c['name'], 'hot us=%.2f frac=%.3f' % (c['ms_per_step']*1e3, c['roofline']['frac']), 'cold us=%.2f frac=%.3f' % (c['cold']['ms_per_step']*1e3, c['cold']['roofline_frac']), c['clocks'])
PY
