python -m pytest tests/test_gpu_villain.py -x -q -k "inplace or config5 or swapping or tiled" > gpurun_out/r2_tests_k.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests_k.log
tail -3 gpurun_out/r2_tests_k.log
KB_OBS=0 python tools/kbench_c5_swap.py
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --configs c2,c5 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for c in d['configs'][1:]:
    print(c['name'], 'hot us=%.2f frac=%.3f' % (c['ms_per_step']*1e3, c['roofline']['frac']), 'cold us=%.2f frac=%.3f' % (c['cold']['ms_per_step']*1e3, c['cold']['roofline_frac']))"
