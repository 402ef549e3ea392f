"""FAST5 (HDF5) access behind one small h5py-shaped interface.

The reference talks to FAST5 files through h5py directly (resquiggle.py:80,
328, 809, 953, 973; nanoraw_helper.py:116).  Neither h5py nor libhdf5 exists in
this image, so the host side goes through ``open_fast5`` which dispatches on the
selected backend:

* ``'h5py'``  -- the real thing, import-guarded (used when h5py is installed);
* ``'mem'``   -- an in-process store with the subset of the h5py object model the
  reference's hot path touches (groups, datasets, attrs, ``in``, ``del``,
  ``values()``, context manager).  File names of the form ``mem://...`` always
  use it.  It keeps h5py's refusal to store ``None`` attributes (SURVEY.md
  section 8b quirks) so reads that fail in the reference fail here too.

* ``'native'`` -- ``hdf5_min.py``: this repo's own reader / writer for the slice of HDF5 that
  FAST5 files use (what ``'auto'`` falls back to when h5py is missing), same object model.

The object model is deliberately h5py's so the same caller code runs on both.
"""
import threading

import numpy as np

_BACKEND = ['auto']


def set_backend(name):
    assert name in ('auto', 'h5py', 'mem', 'native')
    _BACKEND[0] = name


_H5PY = []


def _have_h5py():
    if not _H5PY:  # asked on every open: a failed import is not cached by Python and costs half a millisecond
        try:
            import h5py  # noqa: F401
            _H5PY.append(True)
        except Exception:
            _H5PY.append(False)
    return _H5PY[0]


def _split(path):
    return [p for p in path.split('/') if p]


class MemAttrs(dict):
    _hook = None  # called on every change (the native backend marks the file dirty)

    def __setitem__(self, key, value):
        if value is None:
            # h5py: np.asarray(None) is object dtype, which has no HDF5 type
            raise TypeError("Object dtype dtype('O') has no native HDF5 equivalent")
        dict.__setitem__(self, key, value)
        if self._hook:
            self._hook()

    def __delitem__(self, key):
        dict.__delitem__(self, key)
        if self._hook:
            self._hook()


class MemDataset(object):
    def __init__(self, data, compression=None):
        self._data = np.array(data)
        self.attrs = MemAttrs()
        self.compression = compression

    @property
    def value(self):  # h5py < 3 spelling used by the reference (resquiggle.py:333)
        return self._data

    def __getitem__(self, key):
        if isinstance(key, tuple) and len(key) == 0:
            return self._data
        return self._data[key]

    def __len__(self):
        return len(self._data)

    @property
    def shape(self):
        return self._data.shape

    @property
    def dtype(self):
        return self._data.dtype


class LazyDataset(MemDataset):
    """dataset whose contents are decoded from the file on first access (native HDF5 backend)"""

    def __init__(self, loader, dtype, shape, compression=None):
        self._loader = loader
        self._cache = None
        self._dtype = dtype
        self._shape = tuple(shape)
        self.attrs = MemAttrs()
        self.compression = compression

    @property
    def _data(self):
        if self._cache is None:
            self._cache = np.asarray(self._loader())
        return self._cache

    def __len__(self):
        return self._shape[0]

    @property
    def shape(self):
        return self._shape

    @property
    def dtype(self):
        return self._dtype


class _PackedLoader(object):
    """a dataset that exists as its gzip chunk (zlib stream): inflated only if somebody reads it"""

    def __init__(self, payload, dtype, shape):
        self.payload, self.dtype, self.shape = payload, dtype, shape

    def __call__(self):
        import zlib
        return np.frombuffer(zlib.decompress(self.payload), dtype=self.dtype).reshape(self.shape)

    def stored_chunk(self):
        return self.payload


def create_packed_dataset(group, name, shape, dtype, payload):
    """group.create_dataset(name, data=..., compression='gzip') for data that is already one gzip chunk (a zlib
    stream of the C-order bytes): the native and mem backends keep the chunk as it is, h5py writes it with
    write_direct_chunk.  Returns the dataset."""
    dtype = np.dtype(dtype)
    shape = tuple(int(v) for v in shape)
    if isinstance(group, (MemGroup, MemFile)) or hasattr(group, '_children'):
        parts = _split(name)
        parent = group._walk('/'.join(parts[:-1]), create=True)
        if parts[-1] in parent._children:
            raise ValueError('Unable to create dataset (name already exists)')
        ds = LazyDataset(_PackedLoader(bytes(payload), dtype, shape), dtype, shape, compression='gzip')
        parent._children[parts[-1]] = parent._adopt(ds)
        return ds
    ds = group.create_dataset(name, shape=shape, dtype=dtype, chunks=shape, compression='gzip')  # h5py
    ds.id.write_direct_chunk((0,) * len(shape), bytes(payload))
    return ds


class MemGroup(object):
    def __init__(self, writable=True):
        self._children = {}
        self.attrs = MemAttrs()
        self._writable = [writable]
        self._hook = None

    def _set_hook(self, hook):
        """`hook()` is called whenever this subtree changes"""
        self._hook = hook
        self.attrs._hook = hook
        for child in self._children.values():
            if isinstance(child, MemGroup):
                child._set_hook(hook)
            else:
                child.attrs._hook = hook

    def _adopt(self, node):
        if self._hook:
            if isinstance(node, MemGroup):
                node._set_hook(self._hook)
            else:
                node.attrs._hook = self._hook
            self._hook()
        return node

    def _walk(self, path, create=False):
        node = self
        for part in _split(path):
            if not isinstance(node, MemGroup):
                raise KeyError(path)
            if part not in node._children:
                if not create:
                    raise KeyError("Unable to open object (component not found: %s)" % part)
                node._children[part] = node._adopt(MemGroup())
            node = node._children[part]
        return node

    def __getitem__(self, path):
        return self._walk(path)

    def __contains__(self, path):
        try:
            self._walk(path)
            return True
        except KeyError:
            return False

    def __delitem__(self, path):
        parts = _split(path)
        parent = self._walk('/'.join(parts[:-1]))
        del parent._children[parts[-1]]
        if self._hook:
            self._hook()

    def create_group(self, path):
        parts = _split(path)
        parent = self._walk('/'.join(parts[:-1]), create=True)
        if parts[-1] in parent._children:
            raise ValueError('Unable to create group (name already exists)')
        parent._children[parts[-1]] = parent._adopt(MemGroup())
        return parent._children[parts[-1]]

    def require_group(self, path):
        return self._walk(path, create=True)

    def create_dataset(self, path, data=None, compression=None, **_):
        parts = _split(path)
        parent = self._walk('/'.join(parts[:-1]), create=True)
        if parts[-1] in parent._children:
            raise ValueError('Unable to create dataset (name already exists)')
        parent._children[parts[-1]] = parent._adopt(MemDataset(data, compression))
        return parent._children[parts[-1]]

    def keys(self):
        return list(self._children.keys())

    def values(self):
        return list(self._children.values())

    def items(self):
        return list(self._children.items())


class MemStore(object):
    """name -> root group; one per process (tests / bench only)."""

    def __init__(self):
        self._files = {}
        self._lock = threading.Lock()

    def create(self, name):
        with self._lock:
            self._files[name] = MemGroup()
            return self._files[name]

    def get(self, name):
        return self._files[name]

    def __contains__(self, name):
        return name in self._files

    def clear(self):
        self._files.clear()


MEM_STORE = MemStore()


class MemFile(object):
    """h5py.File look-alike over MEM_STORE."""

    def __init__(self, name, mode='r'):
        if name not in MEM_STORE:
            if mode in ('r', 'r+'):
                raise IOError("Unable to open file (unable to open file: name = '%s')" % name)
            MEM_STORE.create(name)
        if mode == 'w':
            MEM_STORE.create(name)
        self._root = MEM_STORE.get(name)
        self.filename = name
        self.mode = mode
        self._open = True

    def __getattr__(self, item):
        return getattr(self._root, item)

    def __getitem__(self, path):
        return self._root[path]

    def __contains__(self, path):
        return path in self._root

    def __delitem__(self, path):
        del self._root[path]

    def flush(self):
        pass

    def close(self):
        self._open = False

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def open_fast5(name, mode='r'):
    backend = _BACKEND[0]
    if name.startswith('mem://') or backend == 'mem':
        return MemFile(name, mode)
    if backend == 'h5py' or (backend == 'auto' and _have_h5py()):
        import h5py
        return h5py.File(name, mode)
    # no h5py / libhdf5 in this environment: the built-in reader / writer (hdf5_min.py)
    from . import hdf5_min
    return hdf5_min.NativeFile(name, mode)


def read_dataset(ds):
    """dataset contents on both h5py >= 3 (no .value) and the mem backend."""
    return ds[()]
