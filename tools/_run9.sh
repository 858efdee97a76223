set -x
NCU="ncu --set full --clock-control none --import-source on"
prof() {  # name regex skip count sites kernelpat -- cmd...
  name=$1; re=$2; skip=$3; sites=$4; pat=$5; shift 5
  env "$@" > gpurun_out/plain_$name.log 2>&1 && $NCU -k regex:$re -s $skip -c 1 -o gpurun_out/$name env "$@" > gpurun_out/ncu_$name.log 2>&1
  python tools/ncu_summary.py gpurun_out/$name.ncu-rep $sites > gpurun_out/${name}_summary.txt 2>&1
  NCU_LINES_TOP=90 python tools/ncu_lines.py gpurun_out/$name.ncu-rep $pat $sites > gpurun_out/${name}_lines.txt 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
}
prof r2_c2_sparse_therm villain_smem_filtered 20 4194304 villain_smem_filtered_kernelILi32ELi8ELi1ELb1ELb1ELi0ELb1 KB_OVERLAP=1 KB_OBSIN=1 python tools/kbench.py
prof r2_c5_tile villain_tile_pass 8 8388608 villain_tile_pass_kernelILb1 KB_OBS=0 python tools/kbench_c5_swap.py
rm -f gpurun_out/r2_c5_tile.ncu-rep
prof r2_c5_dn2 villain_stream_dn2 3 16777216 villain_stream_dn2_kernel KB_THERM=5 python -c "
import torch, sys
sys.path.insert(0,'.')
import supervillain_b200 as svb
from supervillain_b200 import ops
S = svb.Villain(svb.Lattice2D(4096), 0.5)
phi, n = svb.BatchedEnsemble(S, 1)._start('hot', 1)
st = ops.VillainInplaceSweeps(phi, n, 0.5, seed=1)
a = torch.zeros((1, 6), dtype=torch.float64, device='cuda'); b = torch.zeros_like(a)
for k in range(12): st.step(k, 1, obs=a, obs_in=b)
torch.cuda.synchronize()
"
rm -f gpurun_out/r2_c5_dn2.ncu-rep
ls -la gpurun_out | tail -20
