python -m pytest tests/test_gpu_villain.py tests/test_gpu_villain_decoupled.py tests/test_gpu_checkpoint.py -x -q > gpurun_out/r2_tests_f.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests_f.log
tail -3 gpurun_out/r2_tests_f.log
KB_L=128 KB_CHAINS=1024 KB_OVERLAP=1 KB_OBSIN=1 KB_THERM=50 python tools/kbench.py
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --configs c2,c4,c5 > gpurun_out/r2_bench_f.json 2> gpurun_out/r2_bench_f.err; tail -3 gpurun_out/r2_bench_f.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_f.json'))
for c in d['configs']:
    print(c['name'], 'hot us=%.2f frac=%.3f' % (c['ms_per_step']*1e3, c['roofline']['frac']), 'cold us=%.2f frac=%.3f' % (c['cold']['ms_per_step']*1e3, c['cold']['roofline_frac']), c['clocks'])
PY
