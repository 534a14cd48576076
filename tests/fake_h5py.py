"""In-memory stand-in for the subset of h5py the reference writer and DeviceLBMCaseWriter use (File, create_dataset with
shape / maxshape / chunks / compression, Dataset.resize / item assignment / shape, File.attrs, close, keys / getitem).
h5py is absent from the build image and from the GPU boxes; with this the HDF5 code paths run in the tests, and the
UNMODIFIED reference `io/lbm_writer.py` can be executed here (tests/golden/gen/make_ref_loop_fixture.py).
Finished files are kept in FILES[path] so that a test (or a reader) can open them again."""
import numpy as np

FILES = {}


class Dataset:
    def __init__(self, name, arr, maxshape=None, chunks=None, compression=None, **kw):
        self.name, self.arr, self.maxshape, self.chunks, self.compression = name, arr, maxshape, chunks, compression
        self.n_resize = 0

    @property
    def shape(self):
        return self.arr.shape

    @property
    def dtype(self):
        return self.arr.dtype

    def resize(self, size, axis=None):
        if self.maxshape is None:
            raise TypeError("Only chunked datasets can be resized")
        shape = list(self.arr.shape)
        if axis is None:
            shape = list(size)
        else:
            shape[axis] = size
        new = np.zeros(shape, self.arr.dtype)
        sl = tuple(slice(0, min(a, b)) for a, b in zip(self.arr.shape, shape))
        new[sl] = self.arr[sl]
        self.arr = new
        self.n_resize += 1

    def __setitem__(self, key, value):
        self.arr[key] = value

    def __getitem__(self, key):
        return self.arr[key]

    def __array__(self, dtype=None, copy=None):
        return self.arr if dtype is None else self.arr.astype(dtype)


class File:
    def __init__(self, path, mode="r", libver=None):
        self.path, self.mode, self.libver = path, mode, libver
        if mode == "r":
            src = FILES[path]
            self.datasets, self.attrs, self.closed = src.datasets, src.attrs, False
        else:
            self.datasets, self.attrs, self.closed = {}, {}, False
            FILES[path] = self
            open(path, "wb").close()   # callers test for / remove the file

    def create_dataset(self, name, shape=None, dtype=None, data=None, **kw):
        if self.closed:
            raise ValueError("file is closed")
        if data is not None:
            arr = np.array(data, dtype=None if dtype is None else np.dtype(dtype))
        else:
            arr = np.zeros(shape, np.dtype(dtype or "f4"))
        ds = Dataset(name, arr, **kw)
        self.datasets[name] = ds
        return ds

    def keys(self):
        return self.datasets.keys()

    def __getitem__(self, name):
        return self.datasets[name]

    def __contains__(self, name):
        return name in self.datasets

    def close(self):
        self.closed = True

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
