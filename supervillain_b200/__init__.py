"""supervillain_b200 -- B200-native (sm_100a) batched Metropolis sampling of the Villain model.

A drop-in for the hot path of evanberkowitz/supervillain: the same lattice / action / generator /
ensemble interface, with the sweeps, the action and the observable reductions executed by
hand-written CUDA kernels in libsvb200.so (C ABI in include/svb200.h, bound with ctypes in
`_lib.py`).  PyTorch is used for device memory and streams only.  There is no CPU fallback.
"""
from . import _lib
from . import ops
from . import lattice
from . import action
from . import generator
from . import batch
from . import hostpath
from . import sharding
from . import analysis
from .batch import Batch, Configurations
from .ensemble import Ensemble, BatchedEnsemble
from .lattice import Lattice, Lattice2D, Form, d, delta
from .action import Villain, Worldline

__version__ = '0.1.0'

__all__ = ['ops', 'lattice', 'action', 'generator', 'batch', 'Batch', 'Configurations', 'Ensemble', 'BatchedEnsemble',
           'Lattice', 'Lattice2D', 'Form', 'd', 'delta', 'Villain', 'Worldline']
