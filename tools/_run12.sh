python -m pytest tests/test_gpu_villain.py tests/test_gpu_villain_decoupled.py -x -q > gpurun_out/r2_diet_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_diet_tests.log
tail -3 gpurun_out/r2_diet_tests.log
{
KB_OVERLAP=1 KB_OBSIN=1 python tools/kbench.py
KB_OVERLAP=1 KB_OBS=0 python tools/kbench.py
KB_L=64 KB_CHAINS=2048 KB_OVERLAP=1 KB_OBSIN=1 python tools/kbench.py
} 2>&1 | grep -v "^+"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --configs c2 > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; tail -3 gpurun_out/r2_bench_d.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_d.json'))
for c in d['configs']:
    print(c['name'], 'hot us=%.2f frac=%.3f' % (c['ms_per_step']*1e3, c['roofline']['frac']), 'cold us=%.2f frac=%.3f' % (c['cold']['ms_per_step']*1e3, c['cold']['roofline_frac']), c['clocks'])
PY
