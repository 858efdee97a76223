python -m pytest tests/test_gpu_villain.py -x -q > gpurun_out/r2_tile_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tile_tests.log
tail -3 gpurun_out/r2_tile_tests.log
SVB_VILLAIN_PASS=stream python -m pytest tests/test_gpu_villain.py -x -q -k "tiled or swapping or cluster or general" > gpurun_out/r2_tile_tests_b.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tile_tests_b.log
tail -3 gpurun_out/r2_tile_tests_b.log
{
python tools/kbench_c5_swap.py
SVB_VILLAIN_PASS=stream python tools/kbench_c5_swap.py
python tools/kbench_c5_swap.py 1024 16
python tools/kbench_c5_swap.py 128 1024
SVB_VILLAIN_KERNEL=smem python tools/kbench_c5_swap.py 128 1024
} 2>&1 | grep -v "^+" > gpurun_out/r2_tile_kbench.txt
cat gpurun_out/r2_tile_kbench.txt
