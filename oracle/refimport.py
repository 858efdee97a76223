"""Import the upstream reference (`supervillain`) from /root/reference.

TEST INFRASTRUCTURE ONLY.  This is used in the build container to (a) validate the CPU
restatements in this directory against the real reference and (b) generate the golden vectors
committed under tests/golden/.  /root/reference does not exist on the GPU box, so nothing on
the product path, in `-m gpu` tests, in smoke() or in bench.py may call this.
"""
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get('SVB_REFERENCE_ROOT', '/root/reference')
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'stubs')


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'supervillain'))


def import_reference():
    """Return the `supervillain` module of the reference, stubbing h5py/matplotlib if absent."""
    if not available():
        raise ImportError(f'reference tree not found at {REFERENCE_ROOT}')
    for name in ('h5py', 'matplotlib'):
        try:
            importlib.import_module(name)
        except ImportError:
            if _STUBS not in sys.path:
                sys.path.append(_STUBS)  # appended: a real install always wins
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if 'NUMBA_CACHE_DIR' not in os.environ:
        # the reference tree is read-only, so numba needs a cache directory elsewhere; a fresh one per process, because a
        # cache written by another process has failed to load here ("NRT_adapt_ndarray_to_python: descr is NULL")
        import tempfile
        os.environ['NUMBA_CACHE_DIR'] = tempfile.mkdtemp(prefix='svb_numba_cache_')
    return importlib.import_module('supervillain')
