"""oracle/ -- CPU restatements of the reference's algorithms for the hot path.

TEST INFRASTRUCTURE.  Everything under this directory exists to CHECK the CUDA path and to be
timed as the CPU baseline; it is never the thing shipped.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / `--impl reference` legs may import, call, link or execute it.
The product package (supervillain_b200/) must never import from here and fails loudly when its
CUDA library is missing.

Parity status: PINNED -- every restatement here is checked against golden vectors produced by
running the reference itself (tests/golden/make_golden.py, run in the build container where
/root/reference is mounted) and against the reference's own algebraic identities
(test/test_delta_s.py, test/test_lattice_kernels.py, test/test_vortex_sparse.py patterns).
"""
