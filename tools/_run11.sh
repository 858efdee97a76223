python -m pytest tests/test_gpu_villain.py -x -q -k "inplace or config5" > gpurun_out/r2_inplace_tests3.log 2>&1; echo "rc=$?" >> gpurun_out/r2_inplace_tests3.log
tail -3 gpurun_out/r2_inplace_tests3.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --configs c2,c5 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; tail -3 gpurun_out/r2_bench_c.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_c.json'))
for c in d['configs']:
    print(c['name'], 'hot us=%.2f frac=%.3f' % (c['ms_per_step']*1e3, c['roofline']['frac']), 'cold us=%.2f frac=%.3f' % (c['cold']['ms_per_step']*1e3, c['cold']['roofline_frac']), c['clocks'])
PY
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --configs c2,c5 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 60 --csv --log-file gpurun_out/r2_c5_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --configs c2,c5 > /dev/null 2>&1
tail -20 gpurun_out/r2_c5_launches.csv | cut -c1-200
