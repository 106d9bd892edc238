"""Minimal in-memory stand-in for the subset of h5py that quflow_b200.simulation uses (h5py is not installed in this
image).  Files persist in a module-level dict keyed by path, so reopening works inside one test process; it also
touches the path on disk so that ``os.path.exists`` behaves.  Resizable datasets, attrs, groups, ``in``, indexing."""
import os

import numpy as np

_FILES = {}


class Attrs(dict):
    pass


class Dataset:
    """Resizable along axis 0 with amortised O(1) appends (capacity doubling), like a chunked HDF5 dataset whose append
    writes one new chunk instead of rewriting the file."""

    def __init__(self, shape, dtype, maxshape=None, chunks=None):
        self._buf = np.zeros(shape, dtype=dtype)
        self._n = shape[0] if len(shape) else 0
        self.maxshape, self.chunks = maxshape, chunks
        self.attrs = Attrs()

    @property
    def _a(self):
        return self._buf[:self._n] if self._buf.ndim else self._buf

    shape = property(lambda self: self._a.shape)
    dtype = property(lambda self: self._buf.dtype)

    def resize(self, size, axis=0):
        assert axis == 0 and self.maxshape is not None and self.maxshape[0] is None, "dataset is not resizable along this axis"
        if size > self._buf.shape[0]:
            cap = max(size, 2 * self._buf.shape[0])
            nb = np.zeros((cap,) + self._buf.shape[1:], dtype=self._buf.dtype)
            nb[:self._n] = self._buf[:self._n]
            self._buf = nb
        elif size > self._n:
            self._buf[self._n:size] = 0
        self._n = size

    def __getitem__(self, idx):
        return self._a[idx]

    def __setitem__(self, idx, value):
        self._a[idx] = value


class Group:
    def __init__(self):
        self.items_, self.attrs = {}, Attrs()

    def _walk(self, path, create=False):
        node = self
        for part in [p for p in path.split("/") if p]:
            if part not in node.items_:
                if not create:
                    raise KeyError(path)
                node.items_[part] = Group()
            node = node.items_[part]
        return node

    def __getitem__(self, path):
        return self._walk(path)

    def __contains__(self, path):
        try:
            self._walk(path)
            return True
        except KeyError:
            return False

    def keys(self):
        return self.items_.keys()

    def create_group(self, path):
        return self._walk(path, create=True)

    def create_dataset(self, path, shape, dtype=None, maxshape=None, chunks=None):
        parts = [p for p in path.split("/") if p]
        parent = self._walk("/".join(parts[:-1]), create=True)
        ds = Dataset(shape, dtype, maxshape, chunks)
        parent.items_[parts[-1]] = ds
        return ds


class File(Group):
    def __new__(cls, filename, mode="r"):
        filename = str(filename)
        if mode == "w" or (mode in ("a", "r+") and filename not in _FILES and mode == "a"):
            obj = super().__new__(cls)
            Group.__init__(obj)
            _FILES[filename] = obj
            open(filename, "wb").close()
        elif filename not in _FILES:
            raise OSError(f"no such file {filename}")
        obj = _FILES[filename]
        obj.mode = mode
        return obj

    def __init__(self, filename, mode="r"):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def close(self):
        pass
