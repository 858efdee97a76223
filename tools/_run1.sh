set -x
python -m pytest tests/test_gpu_villain.py -x -q > gpurun_out/r2_stream_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_stream_tests.log
tail -5 gpurun_out/r2_stream_tests.log
{
for k in smem stream; do
  SVB_VILLAIN_KERNEL=$k KB_OVERLAP=1 KB_OBSIN=1 python tools/kbench.py
  SVB_VILLAIN_KERNEL=$k KB_OVERLAP=0 KB_OBS=0 python tools/kbench.py
done
for v in t256m3 t128m8 t128m6 t128m5; do
  SVB200_LIB=$PWD/variants/libsvb200_$v.so KB_OVERLAP=1 KB_OBSIN=1 python tools/kbench.py
done
for k in smem stream; do
  SVB_VILLAIN_KERNEL=$k KB_L=128 KB_CHAINS=1024 KB_OVERLAP=1 KB_OBSIN=1 python tools/kbench.py
  SVB_VILLAIN_KERNEL=$k KB_L=64 KB_CHAINS=2048 KB_OVERLAP=1 KB_OBSIN=1 python tools/kbench.py
  SVB_VILLAIN_KERNEL=$k python tools/kbench_c5_swap.py
done
} > gpurun_out/r2_stream_kbench.txt 2>&1
cat gpurun_out/r2_stream_kbench.txt
