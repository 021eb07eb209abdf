"""In-memory stand-in for the slice of ``h5py`` that HyMD's ``file_io.py`` uses (test infrastructure).

``h5py`` is not installed in this image, so the H5MD layer is exercised over this
module instead: ``File`` / ``Group`` / ``Dataset`` objects with ``create_group``, ``create_dataset(name, shape,
dtype)``, numpy-style ``__getitem__`` / ``__setitem__``, ``.attrs``, ``.name``, ``keys()``, ``in``, truthiness of an
open file.  Two h5py rules that shape the reference's code are enforced so that a port cannot get away without them:
point selections (index lists) must be strictly increasing, and a closed file is falsy.  ``tree(file)`` flattens a
file into ``{path: array, path@attr: value}`` for golden comparisons.

The same module is handed to the REFERENCE's unmodified ``hymd/file_io.py`` by
``tests/golden/make_file_io_golden.py`` (as ``sys.modules["h5py"]``) and to ``hymd_b200.file_io`` by the tests
(``file_io.set_backend``), so both write into the same kind of object."""
from __future__ import annotations

import numpy as np


def _check_selection(key):
    keys = key if isinstance(key, tuple) else (key,)
    for k in keys:
        if isinstance(k, (list, np.ndarray)) and np.asarray(k).dtype != bool:
            a = np.asarray(k).reshape(-1)
            if a.size > 1 and not np.all(np.diff(a) > 0):
                raise TypeError("Indexing elements must be in increasing order")      # h5py's own message


class Attrs(dict):
    def __setitem__(self, key, value):
        if isinstance(value, str):
            value = value                       # h5py stores str as variable-length UTF-8, reads back str
        elif isinstance(value, bytes):
            value = np.bytes_(value)
        elif not isinstance(value, np.generic):
            value = np.asarray(value)
            if value.ndim == 0:
                value = value[()]
        super().__setitem__(key, value)


class Dataset:
    def __init__(self, name, shape, dtype):
        self.name = name
        self.data = np.zeros(tuple(int(s) for s in shape), dtype=np.dtype(dtype))
        self.attrs = Attrs()

    shape = property(lambda self: self.data.shape)
    dtype = property(lambda self: self.data.dtype)

    def __getitem__(self, key):
        _check_selection(key)
        return self.data[key]

    def __setitem__(self, key, value):
        _check_selection(key)
        self.data[key] = value

    def __len__(self):
        return len(self.data)


class Group:
    def __init__(self, name, root=None):
        self.name = name
        self.members = {}
        self.attrs = Attrs()
        self._root = root if root is not None else self

    file = property(lambda self: self._root)

    def _resolve(self, path, create=False):
        node = self._root if path.startswith("/") else self
        parts = [p for p in path.split("/") if p]
        for p in parts[:-1]:
            if p not in node.members:
                if not create:
                    raise KeyError(path)
                node.members[p] = Group((node.name.rstrip("/") + "/" + p), self._root)
            node = node.members[p]
        return node, (parts[-1] if parts else "")

    def create_group(self, name):
        node, leaf = self._resolve(name, create=True)
        if leaf in node.members:
            raise ValueError(f"Unable to create group (name already exists): {name}")
        g = Group(node.name.rstrip("/") + "/" + leaf, self._root)
        node.members[leaf] = g
        return g

    def create_dataset(self, name, shape=None, dtype=None, data=None):
        node, leaf = self._resolve(name, create=True)
        if leaf in node.members:
            raise ValueError(f"Unable to create dataset (name already exists): {name}")
        if data is not None:
            data = np.asarray(data, dtype=dtype)
            shape = data.shape
            dtype = data.dtype
        if isinstance(shape, (int, np.integer)):
            shape = (shape,)
        d = Dataset(node.name.rstrip("/") + "/" + leaf, shape, dtype if dtype is not None else "float32")
        if data is not None:
            d.data[...] = data
        node.members[leaf] = d
        return d

    def __getitem__(self, path):
        node, leaf = self._resolve(path)
        return node.members[leaf] if leaf else node

    def __contains__(self, path):
        try:
            self[path]
            return True
        except KeyError:
            return False

    def keys(self):
        return self.members.keys()

    def __iter__(self):
        return iter(self.members)


class File(Group):
    def __init__(self, path=None, mode="r", driver=None, comm=None, **kw):
        super().__init__("/")
        self.filename, self.mode, self.driver = str(path), mode, driver
        self._open = True

    def close(self):
        self._open = False

    def flush(self):
        pass

    def __bool__(self):
        return self._open

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def tree(node, out=None):
    """Flatten: datasets as ``path -> array``, attributes as ``path@name -> value``, groups as ``path/ -> None``."""
    out = {} if out is None else out
    for k, v in node.attrs.items():
        out[f"{node.name}@{k}"] = v
    if isinstance(node, Dataset):
        out[node.name] = node.data
        return out
    if node.name != "/":
        out[node.name + "/"] = None
    for child in node.members.values():
        tree(child, out)
    return out
