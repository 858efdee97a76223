"""Draw-major column storage with the reference's lossless-cast contract.

Mirrors `supervillain.batch.Batch` (supervillain/batch.py:42-267) and
`supervillain.configurations.Configurations` (supervillain/configurations.py:10-126): a column is
an array of shape (draw, ...); `batch[i]` is one draw (wrapped in `cls`, e.g. `Form`), a slice is a
sub-Batch, and stores that would lose information raise TypeError (batch.py:206-227).
"""
import numbers
import warnings

import numpy as np


class Batch:
    def __init__(self, draws_or_data, *, cls=None, shape=None, dtype=None, **item_kwargs):
        if isinstance(draws_or_data, numbers.Integral) and not isinstance(draws_or_data, bool):
            if cls is not None:
                spatial = cls.spatial_shape(**item_kwargs)
            elif shape is None:
                raise ValueError('Batch(draws, …) requires shape= when cls is None.')
            else:
                spatial = tuple(shape)
            data = np.zeros((int(draws_or_data),) + spatial, dtype=float if dtype is None else dtype)
        else:
            data = np.asarray(draws_or_data) if dtype is None else self._checked_array(draws_or_data, dtype)
        self._data = data
        self.cls = cls
        self.dtype = data.dtype
        self._item_kwargs = item_kwargs

    @classmethod
    def from_data(cls, data, *, dtype=None, **kwargs):
        return cls(data, dtype=dtype, **kwargs)

    @property
    def array(self):
        return self._data

    @staticmethod
    def as_array(column):
        return column.array if isinstance(column, Batch) else column

    @property
    def shape(self):
        return self._data.shape

    def __len__(self):
        return len(self._data)

    def __getitem__(self, index):
        if isinstance(index, numbers.Integral) and not isinstance(index, bool):
            item = self._data[index]
            return item if self.cls is None else self.cls(item, dtype=self.dtype, **self._item_kwargs)
        if type(index) is slice:
            return Batch(self._data[index], cls=self.cls, dtype=self.dtype, **self._item_kwargs)
        return self._data[index]

    def __setitem__(self, index, item):
        self._data[index] = self._checked_array(item, self.dtype)

    @staticmethod
    def _checked_array(data, dtype):
        """Cast to `dtype` only when every value survives the round trip (batch.py:206-227)."""
        arr = np.asarray(data)
        dtype = np.dtype(dtype)
        if arr.dtype == dtype:
            return arr
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            out = arr.astype(dtype)
        if not np.array_equal(out, arr):
            raise TypeError(f'Batch cannot store {arr.dtype} data as {dtype} without loss '
                            f'(the values do not round-trip); convert it explicitly first.')
        return out

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def __repr__(self):
        name = self.cls.__name__ if self.cls is not None else 'ndarray'
        return f'Batch(shape={self.shape}, cls={name}, dtype={self.dtype})'


class Configurations:
    """A dict of equally long Batch columns; `cfgs[i]` is a dict of per-draw values."""

    def __init__(self, dictionary):
        self.__dict__['fields'] = dictionary

    def __str__(self):
        return str(self.fields)

    def __contains__(self, name):
        return name in self.fields

    def __getitem__(self, index):
        if isinstance(index, numbers.Integral) and not isinstance(index, bool):
            return {k: v[index] for k, v in self.fields.items()}
        if isinstance(index, (slice, list)):
            return Configurations({k: v[index] for k, v in self.fields.items()})
        raise ValueError(f'Not sure how to select configurations given a {type(index)}.')

    def __setitem__(self, index, new):
        for key, value in new.items():
            self.fields[key][index] = value

    def __len__(self):
        lengths = {len(v) for v in self.fields.values() if hasattr(v, '__len__')}
        if len(lengths) > 1:
            raise ValueError('Configurations have no consistent length')
        return lengths.pop() if lengths else None

    def items(self):
        return self.fields.items()

    def __getattr__(self, name):
        try:
            return self.__dict__['fields'][name]
        except KeyError:
            raise AttributeError(name) from None

    def __setattr__(self, name, value):
        if name == 'fields':
            self.__dict__['fields'] = value
        elif name in self.fields:
            self.fields[name] = value
        else:
            self.__dict__[name] = value

    def __ior__(self, value):
        self.fields |= value
        return self

    def copy(self):
        return Configurations(self.fields.copy())
