"""Import the upstream reference (`supervillain`), unmodified.

TEST / BASELINE INFRASTRUCTURE ONLY.  Two places can provide it:
  * /root/reference (the build container) -- used to validate the CPU restatements in this directory and to generate the
    golden vectors committed under tests/golden/;
  * oracle/_ref/ -- the copy `oracle/stage_reference.py` stages from /root/reference at build time (git-ignored, ships to
    the GPU box with the snapshot).  On the GPU box this is the only one; it serves bench.py's CPU-baseline / reference
    arm and the `-m gpu` tests in which the reference's own Ensemble drives the GPU generators.
Nothing on the product path may call this.
"""
import importlib
import os
import sys

from . import stage_reference

REFERENCE_ROOT = os.environ.get('SVB_REFERENCE_ROOT', '/root/reference')
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'stubs')


def root():
    """The directory that holds the reference's `supervillain` package: the staged copy if present (identical files, and the
    one that exists on the GPU box), else the live tree; None if neither is there."""
    if stage_reference.staged():
        return stage_reference.DEST
    if os.path.isdir(os.path.join(REFERENCE_ROOT, 'supervillain')):
        return REFERENCE_ROOT
    return None


def available():
    return root() is not None


def import_reference():
    """Return the `supervillain` module of the reference, stubbing h5py/matplotlib if absent."""
    where = root()
    if where is None:
        raise ImportError(f'reference not found: neither {stage_reference.DEST} (python -m oracle.stage_reference) nor {REFERENCE_ROOT}')
    for name in ('h5py', 'matplotlib'):
        try:
            importlib.import_module(name)
        except ImportError:
            if _STUBS not in sys.path:
                sys.path.append(_STUBS)  # appended: a real install always wins
    if where not in sys.path:
        sys.path.insert(0, where)
    if 'NUMBA_CACHE_DIR' not in os.environ:
        # a fresh numba cache directory per process: the live tree is read-only, and a cache written by another process
        # has failed to load here ("NRT_adapt_ndarray_to_python: descr is NULL")
        import tempfile
        os.environ['NUMBA_CACHE_DIR'] = tempfile.mkdtemp(prefix='svb_numba_cache_')
    return importlib.import_module('supervillain')
