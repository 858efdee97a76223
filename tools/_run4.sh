set -x
NCU="ncu --set full --clock-control none --import-source on"
prof() {  # name regex skip sites kernelpat -- cmd...
  name=$1; re=$2; skip=$3; sites=$4; pat=$5; shift 5
  env "$@" > gpurun_out/plain_$name.log 2>&1 && $NCU -k regex:$re -s $skip -c 1 -o gpurun_out/$name env "$@" > gpurun_out/ncu_$name.log 2>&1
  python tools/ncu_summary.py gpurun_out/$name.ncu-rep $sites > gpurun_out/${name}_summary.txt 2>&1
  NCU_LINES_TOP=70 python tools/ncu_lines.py gpurun_out/$name.ncu-rep $pat $sites > gpurun_out/${name}_lines.txt 2>&1
}
prof r2_c2_sparse villain_smem_filtered 30 4194304 villain_smem_filtered_kernelILi32ELi8ELi1ELb1ELb1ELi0ELb1 KB_OVERLAP=1 KB_OBSIN=1 python tools/kbench.py
prof r2_c4_cluster_sparse villain_cluster 20 16777216 villain_cluster_kernelILi128ELi4ELi512ELi1ELb1ELi0ELb1 KB_L=128 KB_CHAINS=1024 KB_OVERLAP=1 KB_OBSIN=1 python tools/kbench.py
rm -f gpurun_out/r2_c4_cluster_sparse.ncu-rep
prof r2_c5_tiled villain_tiled_filtered 8 16777216 villain_tiled_filtered_kernel SVB_VILLAIN_KERNEL=smem python tools/kbench_c5_swap.py
rm -f gpurun_out/r2_c5_tiled.ncu-rep
prof r2_c5_stream villain_stream_pass 8 8388608 villain_stream_pass_kernelILb1 X=1 python tools/kbench_c5_swap.py
rm -f gpurun_out/r2_c5_stream.ncu-rep
ls -la gpurun_out/
