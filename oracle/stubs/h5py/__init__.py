"""Stub of h5py: just enough names for `import supervillain` to succeed where h5py is absent.

Test infrastructure only (used by oracle/refimport.py when generating golden vectors in the
build container).  No HDF5 functionality.
"""


class Group:
    pass


class File(Group):
    def __init__(self, *args, **kwargs):
        raise RuntimeError('h5py stub: real HDF5 I/O is not available in this environment')


class Dataset:
    pass

__svb_stub__ = True          # lets tests tell this stub from a real h5py
