python -m pytest tests/test_gpu_worldline.py tests/test_gpu_villain.py -x -q -k "worldline or soak or inplace or config5" > gpurun_out/r2_tests_l.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests_l.log
tail -3 gpurun_out/r2_tests_l.log
python tools/kbench2.py c3 2>&1 | grep -v "^+"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --configs c2,c3,c5 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for c in d['configs'][1:]:
    print(c['name'], 'hot us=%.2f frac=%.3f' % (c['ms_per_step']*1e3, c['roofline']['frac']), 'cold us=%.2f frac=%.3f' % (c['cold']['ms_per_step']*1e3, c['cold']['roofline_frac']))"
