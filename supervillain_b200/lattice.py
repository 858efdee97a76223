"""Host-side mirror of the reference's lattice interface for the hot path (D = 2).

Mirrors `supervillain.lattice.{Lattice, Lattice2D, Form, d, delta}` and `Form.face_sum /
Form.coface_sum` (supervillain/lattice/compact.py:60-86, 665-752, 848-890, 973-1037;
two_dimensional.py:9-26).  The operators run on the GPU through svb_form_op; a host `Form` is
copied to the device, transformed there, and copied back, so the call sites of the reference
(`d(phi)`, `n.face_sum()`, ...) keep working unchanged.  For resident data use
`supervillain_b200.ops.form_op` on device tensors directly.
"""
from functools import cached_property
from math import comb

import numpy as np
import torch

from . import ops


def _dimension(n):
    """FFT-convention coordinates of a periodic direction (lattice/__init__.py:4-9)."""
    idx = np.arange(n)
    return np.where(idx <= n // 2, idx, idx - n)


class Lattice:
    """A periodic hypercubic lattice with N sites per direction.  Only D = 2 is supported on the GPU path."""

    def __init__(self, D, N):
        if D != 2:
            raise NotImplementedError('supervillain_b200 implements the D=2 hot path only '
                                      '(every configuration named in BASELINE.json is D=2)')
        if int(N) != N or N < 3:
            raise ValueError('N must be an integer >= 3')
        self.D = int(D)
        self.N = int(N)
        self.components = {0: [()], 1: [(0,), (1,)], 2: [(0, 1)]}
        self.comp_index = {p: {c: i for i, c in enumerate(cs)} for p, cs in self.components.items()}

    def __str__(self):
        return f'Lattice(D={self.D}, N={self.N})'

    __repr__ = __str__

    def __eq__(self, other):
        return isinstance(other, Lattice) and (self.D, self.N) == (other.D, other.N)

    def __hash__(self):
        return hash((self.D, self.N))

    @property
    def dim(self):
        return self.D

    @cached_property
    def sites(self):
        return self.N ** self.D

    @cached_property
    def links(self):
        return self.D * self.sites

    @cached_property
    def dims(self):
        return (self.N,) * self.D

    @cached_property
    def origin(self):
        return (0,) * self.D

    @cached_property
    def cells_of_degree(self):
        return {p: comb(self.D, p) * self.sites for p in range(self.D + 1)}

    @cached_property
    def coords(self):
        c = _dimension(self.N)
        return np.stack(np.meshgrid(c, c, indexing='ij'), axis=0)

    @cached_property
    def coordinates(self):
        c = _dimension(self.N)
        return np.stack([a.flatten() for a in np.meshgrid(c, c, indexing='ij')], axis=1)

    @cached_property
    def colour_map(self):
        """(N, N) colour of every site; the kernels' `site_colour` computes the same function."""
        c0, c1 = self.coords
        parity = np.mod(c0 + c1, 2)
        if self.N % 2 == 0:
            return parity
        return 2 * ((c0 >= 0) != (c1 >= 0)).astype(parity.dtype) + parity

    @cached_property
    def checkerboarding(self):
        """Tuple of `np.where` index tuples, one per colour, row-major (compact.py:192-239)."""
        ncol = 2 if self.N % 2 == 0 else 4
        return tuple(np.where(self.colour_map == c) for c in range(ncol))

    def zeros(self, p, dtype=float):
        return Form(np.zeros((comb(self.D, p),) + self.dims, dtype=dtype), degree=p, lattice=self)

    form = zeros

    def mod(self, x):
        return _dimension(self.N)[np.mod(np.asarray(x), self.N)]


class Lattice2D(Lattice):
    """`Lattice(D=2, N=n)` (two_dimensional.py:9-26)."""

    def __init__(self, n):
        super().__init__(2, n)

    def __str__(self):
        return f'Lattice2D({self.N})'

    __repr__ = __str__


class Form(np.ndarray):
    """A p-form stored compactly as (C(D,p), N, N): an ndarray carrying `degree` and `lattice`."""

    __batch_tag__ = 'Form'

    @classmethod
    def spatial_shape(cls, *, degree, lattice):
        return (comb(lattice.D, degree),) + (lattice.N,) * lattice.D

    def __new__(cls, input_array, *, degree, lattice, dtype=None):
        obj = np.asarray(input_array, dtype=dtype).view(cls)
        obj.degree = degree
        obj.lattice = lattice
        return obj

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self.degree = getattr(obj, 'degree', None)
        self.lattice = getattr(obj, 'lattice', None)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        forms = [x for x in inputs if isinstance(x, Form)]
        raw = tuple(np.asarray(x) for x in inputs)
        if kwargs.get('out') is not None:
            kwargs['out'] = tuple(np.asarray(o) for o in kwargs['out'])
        result = getattr(ufunc, method)(*raw, **kwargs)
        if (len({f.degree for f in forms}) == 1 and isinstance(result, np.ndarray)
                and result.shape == forms[0].shape):
            return Form(result, degree=forms[0].degree, lattice=forms[0].lattice)
        return result

    def face_sum(self):
        return _host_op('face_sum', self)

    def coface_sum(self):
        return _host_op('coface_sum', self)

    def __repr__(self):
        return f'Form(degree={self.degree}, shape={self.shape}, lattice={self.lattice})'


_TORCH_OF = {np.dtype('float64'): torch.float64, np.dtype('float32'): torch.float32,
             np.dtype('int32'): torch.int32, np.dtype('int64'): torch.int64}


def _host_op(op, f):
    """Run a form operator on the GPU for a host Form (or a batch (chains, C, N, N) of them)."""
    if not isinstance(f, Form):
        raise TypeError(f'{op} needs a Form (an array carrying its degree); got {type(f).__name__}')
    p, lat = f.degree, f.lattice
    arr = np.asarray(f)
    if arr.dtype == np.bool_:
        arr = arr.astype(np.int64)
    if arr.dtype not in _TORCH_OF:
        raise TypeError(f'{op}: unsupported dtype {arr.dtype}')
    if (op in ('d', 'coface_sum') and p == lat.D) or (op in ('delta', 'face_sum') and p == 0):
        return 0          # the ends of the complex (compact.py:999-1000, 1035-1036, 865-866, 888-889)
    batched = arr.ndim == 4
    dev = torch.from_numpy(np.ascontiguousarray(arr if batched else arr[None])).cuda()
    out = ops.form_op(op, p, dev)
    res = out.cpu().numpy()
    out_degree = p + 1 if op in ('d', 'coface_sum') else p - 1
    return Form(res if batched else res[0], degree=out_degree, lattice=lat)


def d(f):
    """Exterior derivative (compact.py:973-1001), evaluated on the GPU."""
    return _host_op('d', f)


def delta(f):
    """Codifferential (compact.py:1008-1037), evaluated on the GPU."""
    return _host_op('delta', f)
