python -m pytest tests/test_gpu_villain.py tests/test_gpu_checkpoint.py -x -q > gpurun_out/r2_sparse_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_sparse_tests.log
tail -4 gpurun_out/r2_sparse_tests.log
{
for sp in 0 1; do
  SVB_VILLAIN_SPARSE=$sp KB_OVERLAP=1 KB_OBSIN=1 python tools/kbench.py
  SVB_VILLAIN_SPARSE=$sp KB_OVERLAP=1 KB_OBS=0 python tools/kbench.py
  SVB_VILLAIN_SPARSE=$sp KB_OVERLAP=0 KB_OBS=0 python tools/kbench.py
  SVB_VILLAIN_SPARSE=$sp KB_L=64 KB_CHAINS=2048 KB_OVERLAP=1 KB_OBSIN=1 python tools/kbench.py
  SVB_VILLAIN_SPARSE=$sp KB_L=16 KB_CHAINS=16384 KB_OVERLAP=1 KB_OBSIN=1 python tools/kbench.py
done
} 2>&1 | grep -v "^+" > gpurun_out/r2_sparse_kbench.txt
cat gpurun_out/r2_sparse_kbench.txt
