{
python tools/kbench_c5_swap.py
KB_OBS=0 python tools/kbench_c5_swap.py
KB_OBS=0 SVB_VILLAIN_PASS=stream python tools/kbench_c5_swap.py
KB_OBS=0 SVB_VILLAIN_KERNEL=smem python tools/kbench_c5_swap.py
KB_OBS=0 python tools/kbench_c5_swap.py 128 1024
KB_OBS=0 KB_THERM=0 python tools/kbench_c5_swap.py
} 2>&1 | grep -v "^+" > gpurun_out/r2_tile_kbench2.txt
cat gpurun_out/r2_tile_kbench2.txt
