SVB_VILLAIN_KERNEL=smem python -m pytest tests/test_gpu_villain.py -x -q -k "cluster or overlapped or general or arriving" > gpurun_out/r2_csparse_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_csparse_tests.log
tail -3 gpurun_out/r2_csparse_tests.log
{
for sp in 0 1; do
  SVB_VILLAIN_KERNEL=smem SVB_VILLAIN_SPARSE=$sp KB_L=128 KB_CHAINS=1024 KB_OVERLAP=1 KB_OBSIN=1 python tools/kbench.py
  SVB_VILLAIN_KERNEL=smem SVB_VILLAIN_SPARSE=$sp KB_L=128 KB_CHAINS=1024 KB_OVERLAP=0 KB_OBS=0 python tools/kbench.py
done
SVB_VILLAIN_KERNEL=stream KB_L=128 KB_CHAINS=1024 KB_OVERLAP=1 KB_OBSIN=1 python tools/kbench.py
for sp in 0 1; do
SVB_VILLAIN_SPARSE=$sp python bench.py --steps 20 --warmup 5 --no-cpu-baseline --configs c2 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench c2 sparse=$sp', d['ms_per_step']*1e3, d['roofline']['frac'], d['clocks'], 'cold', d['configs'][0]['cold']['ms_per_step']*1e3)"
done
} 2>&1 | grep -v "^+" > gpurun_out/r2_csparse_kbench.txt
cat gpurun_out/r2_csparse_kbench.txt
