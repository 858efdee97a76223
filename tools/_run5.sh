SVB_VILLAIN_KERNEL=stream python -m pytest tests/test_gpu_villain.py -x -q > gpurun_out/r2_stream2_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_stream2_tests.log
tail -3 gpurun_out/r2_stream2_tests.log
{
for k in smem stream; do
  SVB_VILLAIN_KERNEL=$k KB_OVERLAP=1 KB_OBSIN=1 python tools/kbench.py
  SVB_VILLAIN_KERNEL=$k KB_L=64 KB_CHAINS=2048 KB_OVERLAP=1 KB_OBSIN=1 python tools/kbench.py
  SVB_VILLAIN_KERNEL=$k KB_L=128 KB_CHAINS=1024 KB_OVERLAP=1 KB_OBSIN=1 KB_THERM=50 python tools/kbench.py
  SVB_VILLAIN_KERNEL=$k python tools/kbench_c5_swap.py
done
} 2>&1 | grep -v "^+" > gpurun_out/r2_stream2_kbench.txt
cat gpurun_out/r2_stream2_kbench.txt
