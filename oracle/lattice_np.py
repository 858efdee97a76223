"""CPU restatement (numpy) of the reference's D=2 lattice form calculus.

TEST INFRASTRUCTURE ONLY -- this module is the *checker* for the CUDA path.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.

Fields are plain ndarrays shaped (C, N, N) -- the reference's compact layout
(supervillain/lattice/compact.py:665-716) with the component axis first -- or with any number
of leading batch axes, (..., C, N, N).  Axis -2 is lattice direction 0, axis -1 is direction 1.

Each function cites the reference lines it restates.  The arithmetic ORDER of the reference's
numba kernels (supervillain/lattice/_kernels.py:19-46) is kept so float results are bit-equal.
"""
import numpy as np

TWO_PI = 2 * np.pi


def fft_coordinates(N):
    """Index -> FFT-convention coordinate, [0, 1, ..., N//2, -(N-1)//2, ..., -1].

    Restates supervillain/lattice/__init__.py:4-9 (`_dimension`).
    """
    idx = np.arange(N)
    return np.where(idx <= N // 2, idx, idx - N)


def colour_map(N):
    """(N, N) int array: checkerboard colour of every site.

    Restates supervillain/lattice/compact.py:192-239 for D=2: even N has two colours given by the
    parity of the coordinate sum; odd N has four, enumerated as 2*b + parity where b=0 when the
    two FFT coordinates fall in the same sign class (both >=0 or both <0) and b=1 otherwise
    (`_hyperoctant_pair_mask`, compact.py:36-53).
    """
    c = fft_coordinates(N)
    c0, c1 = np.meshgrid(c, c, indexing='ij')
    parity = np.mod(c0 + c1, 2)
    if N % 2 == 0:
        return parity.astype(np.int64)
    mixed = ((c0 >= 0) != (c1 >= 0)).astype(np.int64)
    return 2 * mixed + parity


def n_colours(N):
    return 2 if N % 2 == 0 else 4


def colour_sites(N):
    """Tuple (one entry per colour) of (x0, x1) index arrays in row-major order.

    Same content and order as `Lattice.checkerboarding` (compact.py:231-239), which is built from
    `np.where` and therefore row-major within a colour.
    """
    cmap = colour_map(N)
    return tuple(np.where(cmap == c) for c in range(n_colours(N)))


# ---------------------------------------------------------------------------------------------
# d, delta, face_sum, coface_sum.  Incidence rows for D=2 (compact.py:144-174):
#   d,0:     (0,0,0,+) (1,0,1,+)          d,1:        (0,1,0,+) (0,0,1,-)
#   delta,1: (0,0,0,+) (0,1,1,+)          delta,2:    (0,0,1,-) (1,0,0,+)
#   face_sum / coface_sum: the same rows with sign +1.
# Kernel bodies (_kernels.py:31-45):
#   d:      res[out] += sign * (F[in][x+e] - F[in][x])
#   delta:  res[out] -= sign * (F[in][x] - F[in][x-e])
#   sums:   res[out] += F[in][x];  res[out] += F[in][x +/- e]      (two separate adds)
# ---------------------------------------------------------------------------------------------

def _fwd(a, axis):
    """a[x + e_axis] (periodic); axis is 0 or 1 and refers to the last two array axes."""
    return np.roll(a, -1, axis=axis - 2)


def _bwd(a, axis):
    """a[x - e_axis] (periodic)."""
    return np.roll(a, +1, axis=axis - 2)


def d0(phi):
    """d of a 0-form (..., 1, N, N) -> 1-form (..., 2, N, N).  compact.py:973-1001."""
    f = phi[..., 0, :, :]
    out = np.zeros(phi.shape[:-3] + (2,) + phi.shape[-2:], dtype=phi.dtype)
    out[..., 0, :, :] += (_fwd(f, 0) - f)
    out[..., 1, :, :] += (_fwd(f, 1) - f)
    return out


def d1(n):
    """d of a 1-form (..., 2, N, N) -> 2-form (..., 1, N, N).  Rows (0,1,0,+), (0,0,1,-)."""
    n0 = n[..., 0, :, :]
    n1 = n[..., 1, :, :]
    out = np.zeros(n.shape[:-3] + (1,) + n.shape[-2:], dtype=n.dtype)
    out[..., 0, :, :] += (_fwd(n1, 0) - n1)
    out[..., 0, :, :] += -(_fwd(n0, 1) - n0)
    return out


def delta1(m):
    """delta of a 1-form -> 0-form.  compact.py:1008-1037; rows (0,0,0,+), (0,1,1,+)."""
    m0 = m[..., 0, :, :]
    m1 = m[..., 1, :, :]
    out = np.zeros(m.shape[:-3] + (1,) + m.shape[-2:], dtype=m.dtype)
    out[..., 0, :, :] -= (m0 - _bwd(m0, 0))
    out[..., 0, :, :] -= (m1 - _bwd(m1, 1))
    return out


def delta2(v):
    """delta of a 2-form -> 1-form.  Rows (0,0,1,-), (1,0,0,+)."""
    f = v[..., 0, :, :]
    out = np.zeros(v.shape[:-3] + (2,) + v.shape[-2:], dtype=v.dtype)
    out[..., 0, :, :] -= -(f - _bwd(f, 1))
    out[..., 1, :, :] -= (f - _bwd(f, 0))
    return out


def face_sum1(F):
    """face_sum of a 1-form -> 0-form.  compact.py:848-867; _kernels.py:37-45 (two adds per row)."""
    F0 = F[..., 0, :, :]
    F1 = F[..., 1, :, :]
    out = np.zeros(F.shape[:-3] + (1,) + F.shape[-2:], dtype=F.dtype)
    g = out[..., 0, :, :]
    g += F0
    g += _bwd(F0, 0)
    g += F1
    g += _bwd(F1, 1)
    return out


def face_sum2(F):
    """face_sum of a 2-form -> 1-form.  Rows (0,0,1,+), (1,0,0,+)."""
    f = F[..., 0, :, :]
    out = np.zeros(F.shape[:-3] + (2,) + F.shape[-2:], dtype=F.dtype)
    out[..., 0, :, :] += f
    out[..., 0, :, :] += _bwd(f, 1)
    out[..., 1, :, :] += f
    out[..., 1, :, :] += _bwd(f, 0)
    return out


def coface_sum0(F):
    """coface_sum of a 0-form -> 1-form.  compact.py:869-890; rows (0,0,0,+), (1,0,1,+)."""
    f = F[..., 0, :, :]
    out = np.zeros(F.shape[:-3] + (2,) + F.shape[-2:], dtype=F.dtype)
    out[..., 0, :, :] += f
    out[..., 0, :, :] += _fwd(f, 0)
    out[..., 1, :, :] += f
    out[..., 1, :, :] += _fwd(f, 1)
    return out


def coface_sum1(F):
    """coface_sum of a 1-form -> 2-form.  Rows (0,1,0,+), (0,0,1,+)."""
    F0 = F[..., 0, :, :]
    F1 = F[..., 1, :, :]
    out = np.zeros(F.shape[:-3] + (1,) + F.shape[-2:], dtype=F.dtype)
    g = out[..., 0, :, :]
    g += F1
    g += _fwd(F1, 0)
    g += F0
    g += _fwd(F0, 1)
    return out


_OPS = {
    ('d', 0): d0, ('d', 1): d1,
    ('delta', 1): delta1, ('delta', 2): delta2,
    ('face_sum', 1): face_sum1, ('face_sum', 2): face_sum2,
    ('coface_sum', 0): coface_sum0, ('coface_sum', 1): coface_sum1,
}


def form_op(op, degree, F):
    """Dispatch by (operator name, input degree); raises KeyError at the ends of the complex,
    where the reference returns the scalar 0 (compact.py:999-1000, 1035-1036, 865-866, 888-889)."""
    return _OPS[(op, degree)](np.asarray(F))


def correlation(f, g):
    """Translation-averaged cross-correlation, compact.py:465-536:
    fft2(conj(fft2 f) * fft2 g, ortho) / sqrt(N^2), over the last two axes."""
    N2 = f.shape[-1] * f.shape[-2]
    Ff = np.fft.fftn(f, axes=(-2, -1), norm='ortho')
    Fg = np.fft.fftn(g, axes=(-2, -1), norm='ortho')
    return np.fft.fftn(Ff.conj() * Fg, axes=(-2, -1), norm='ortho') / np.sqrt(N2)


def autocorrelation(data, mean=None, cutoff=1e-16):
    """supervillain.analysis.autocorrelation (supervillain/analysis/autocorrelation.py:7-59), restated."""
    data = np.asarray(data, dtype=np.float64)
    if mean is None:
        mean = data.mean()                                                  # :43-44
    Delta = data - mean                                                     # :46
    plus = np.fft.fft(Delta, norm='backward')                               # :48
    minus = np.fft.ifft(Delta, norm='forward')                              # :49
    C = np.fft.fft(plus * minus, norm='backward').real / (len(Delta)) ** 2  # :51
    if np.abs(C[0]) < cutoff:                                               # :52-53
        raise ValueError('The fluctuations are too small to reliably determine an autocorrelation.')
    C /= C[0]                                                               # :54
    clamped = np.clip(C, 0, None)                                           # :56
    minIdx = np.argmin(clamped)                                             # :57
    return C, int(np.ceil(C[:minIdx].sum()))                                # :58


def block_mean(data, width, weight=None):
    """supervillain.analysis.Blocking._block (supervillain/analysis/blocking.py:54-66; drop = len % width, :35), restated
    for a scalar column (T,) or a batch of columns (series, T) -> (..., blocks)."""
    data = np.asarray(data, dtype=np.float64)
    T = data.shape[-1]
    drop = T % width
    w = np.ones(T) if weight is None else np.asarray(weight, dtype=np.float64)
    return (data[..., drop:] * w[drop:]).reshape(*data.shape[:-1], -1, width).mean(axis=-1)


def bootstrap_mean(data, indices, weight=None):
    """supervillain.analysis.Bootstrap._resample (supervillain/analysis/bootstrap.py:57-67), restated for a scalar column
    (T,) or a batch of columns (series, T); indices (T, draws) -> (..., draws)."""
    data = np.asarray(data, dtype=np.float64)
    T = data.shape[-1]
    w = (np.ones(T) if weight is None else np.asarray(weight, dtype=np.float64))[indices]          # (T, draws)
    return (w * data[..., indices]).mean(axis=-2) / w.mean(axis=0)


def fft_coordinate(a, N):
    """The displacement an index stands for (supervillain/lattice/__init__.py:4-9): 0 .. N//2, then -(N-1)//2 .. -1."""
    return a if a <= N // 2 else a - N


def _taxicab(links, alpha, beta, comp_t, comp_s, offset, sign_s):
    """mean_x exp(alpha * S(x; D) + beta * |D|) for every displacement D = (Dt, Dx), where S sums `links` along the taxicab
    path from x: first |Dt| steps in time on component comp_t at column x1 -- rows x0 + offset + i (i = 0 .. Dt-1, sign +)
    for Dt > 0, rows x0 + offset - i (i = 1 .. |Dt|, sign -) for Dt < 0 -- then |Dx| steps in space on component comp_s at
    row x0 + Dt -- columns x1 + offset + j (sign sign_s) for Dx > 0, x1 + offset - j (sign -sign_s) for Dx < 0."""
    links = np.asarray(links, dtype=np.float64)
    N = links.shape[-1]
    out = np.zeros((N, N))
    x0, x1 = np.meshgrid(np.arange(N), np.arange(N), indexing='ij')
    for a in range(N):
        for b in range(N):
            Dt, Dx = fft_coordinate(a, N), fft_coordinate(b, N)
            S = np.zeros((N, N))
            for i in (range(Dt) if Dt > 0 else range(-1, Dt - 1, -1)):
                S += (1 if Dt > 0 else -1) * links[comp_t, (x0 + offset + i) % N, x1]
            for j in (range(Dx) if Dx > 0 else range(-1, Dx - 1, -1)):
                S += (sign_s if Dx > 0 else -sign_s) * links[comp_s, (x0 + Dt) % N, (x1 + offset + j) % N]
            out[a, b] = np.exp(alpha * S + beta * (abs(Dt) + abs(Dx))).mean()
    return out


def spin_spin_worldline(links, kappa):
    """Spin_Spin.Worldline (supervillain/observable/spin.py:50-224): links = m - delta(v)/W; the taxicab path from x goes
    Dt in time along (0, .) links then Dx in space along (1, .) links, P = +1 along / -1 against a link, and the
    reweighting factor is exp(-(2 P.links + |P|) / (2 kappa)), averaged over x."""
    return _taxicab(links, -1.0 / kappa, -0.5 / kappa, 0, 1, 0, +1)


def vortex_vortex_villain(links, kappa):
    """Vortex_Vortex.Villain (supervillain/observable/vortex.py:63-189): links = d(phi) - 2 pi n; the dual taxicab path
    changes n by +-1 on the (1, .) links of rows x0 + 1 .. x0 + Dt (column x1) and by -+1 on the (0, .) links of columns
    x1 + 1 .. x1 + Dx (row x0 + Dt); dS = -2 pi kappa (change.links - pi |P|) and V = mean_x exp(-dS)."""
    return _taxicab(links, 2 * np.pi * kappa, -2 * np.pi ** 2 * kappa, 1, 0, 1, -1)
