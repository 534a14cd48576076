"""h5lite: the package's dependency-free HDF5 writer / reader (01-lbm-2d_b200/h5lite.py).

h5py / libhdf5 are not installed here, so the format knowledge is pinned in two steps: (1) the READER parses a file
written by libhdf5 itself -- the MATLAB 7.3 file scipy ships with its test data -- and returns the values scipy's own
tests expect for it; the structures it walks there (superblock 0, root symbol-table entry, group B-tree, local heap,
symbol table node, version-1 object header, dataspace / IEEE datatype / layout messages, a string attribute) are also
compared byte for byte with what the WRITER emits for the same content; (2) the writer's files are read back by that
reader: contiguous datasets, the extensible one-chunk-per-frame dataset across one, two and three B-tree levels, numeric
and variable-length string attributes."""
import importlib
import json
import os
import struct

import numpy as np
import pytest

h5 = importlib.import_module("01-lbm-2d_b200.h5lite")


def _libhdf5_file():
    try:
        import scipy.io.matlab
    except Exception:
        return None
    p = os.path.join(os.path.dirname(scipy.io.matlab.__file__), "tests", "data", "testhdf5_7.4_GLNX86.mat")
    return p if os.path.exists(p) else None


needs_real_file = pytest.mark.skipif(_libhdf5_file() is None, reason="scipy's HDF5 test file is not installed")


@needs_real_file
def test_reader_parses_a_file_written_by_libhdf5():
    d = h5.read(_libhdf5_file())
    # scipy/io/matlab/tests/test_mio.py: testdouble = arange(0, 2 pi + pi / 4, pi / 4), stored as a (9, 1) MATLAB column
    assert d["testdouble"].dtype == np.float64 and d["testdouble"].shape == (9, 1)
    assert np.array_equal(d["testdouble"].ravel(), np.arange(9) * (np.pi / 4))
    assert d["dataset_attrs"]["testdouble"]["MATLAB_class"] == b"double"


@needs_real_file
def test_writer_emits_the_bytes_libhdf5_emits_for_the_same_content():
    r = h5._Reader(_libhdf5_file())
    (name, header), = [e for m in r.messages(r.root_header) if m[0] == h5.MSG_SYMTAB
                      for e in r.group_entries(*struct.unpack_from("<QQ", m[2], 0))]
    msgs = {t: body for t, _, body in r.messages(header)}
    assert msgs[h5.MSG_DATATYPE][:20] == h5._dt_message(np.dtype("<f8"))          # IEEE double, little endian
    assert msgs[h5.MSG_DATASPACE][:24] == h5._space_message((9, 1))               # version-1 simple dataspace
    attr = msgs[h5.MSG_ATTRIBUTE]
    mine = h5._attribute_message("MATLAB_class", struct.pack("<BBBBI", 0x13, 0, 0, 0, 6), h5._space_message(()), b"double")
    assert mine[8:8 + len(attr)].rstrip(b"\0") == attr.rstrip(b"\0")            # version-1 attribute with a scalar dataspace
    p = r.at(header)
    assert bytes(r.buf[p:p + 2]) == h5._object_header([])[:2] and bytes(r.buf[p + 12:p + 16]) == b"\0" * 4   # 16-byte version-1 prefix
    sb = bytes(r.buf[512:512 + 96])
    mine = h5.Writer._superblock(0x60, 0x180, 0x260, 4168)
    assert mine[:20] == sb[:20] and mine[32:40] == sb[32:40] and mine[48:56] == sb[48:56]   # versions, sizes, K values, UNDEFs
    assert struct.unpack_from("<II", sb, 72) == (1, 0) == struct.unpack_from("<II", mine, 72)   # root entry: cached group


@pytest.mark.parametrize("n_frames", [0, 1, 5, 64, 65, 64 * 64 + 3])
def test_round_trip_of_a_case_file(tmp_path, n_frames):
    rng = np.random.default_rng(n_frames)
    th, tw = (3, 5) if n_frames > 100 else (12, 20)
    path = str(tmp_path / "case.h5")
    w = h5.Writer(path)
    sm = rng.standard_normal((2, th, tw)).astype(np.float32)
    w.create_dataset("static_mask", sm)
    dset = w.create_appendable("turbulence", (9, th, tw), "f4")
    with pytest.raises(ValueError):
        h5.read(path)                                 # unfinished: no reader takes it for a finished case
    frames = rng.standard_normal((n_frames, 9, th, tw)).astype(np.float32)
    for f in frames:
        dset.append(f)
    w.create_dataset("mean_vel_field", np.arange(9 * th * tw, dtype=np.float32).reshape(9, th, tw))
    w.create_dataset("counts", np.arange(7, dtype=np.int64))
    w.set_attr("config_json", json.dumps({"name": "urban", "名": "值", "nu": 0.007}))
    w.set_attr("stats_min", np.linspace(-1, 1, 9))
    w.set_attr("stats_mean", np.arange(9, dtype=np.float32))
    w.close()
    w.close()
    d = h5.read(path)
    assert set(d) == {"attrs", "static_mask", "turbulence", "mean_vel_field", "counts"}
    assert d["turbulence"].shape == (n_frames, 9, th, tw) and d["turbulence"].dtype == np.float32
    assert np.array_equal(d["turbulence"], frames) and np.array_equal(d["static_mask"], sm)
    assert d["counts"].dtype == np.int64 and np.array_equal(d["counts"], np.arange(7))
    assert json.loads(d["attrs"]["config_json"]) == {"name": "urban", "名": "值", "nu": 0.007}
    assert d["attrs"]["stats_min"].dtype == np.float64 and np.array_equal(d["attrs"]["stats_min"], np.linspace(-1, 1, 9))
    assert d["attrs"]["stats_mean"].dtype == np.float32
    with open(path, "rb") as f:
        raw = f.read()
    assert struct.unpack_from("<Q", raw, 40)[0] == len(raw)      # end-of-file address = file size (libhdf5 checks it)


def test_structure_of_the_extensible_dataset(tmp_path):
    """What libhdf5 needs to treat `turbulence` the way the reference declares it (writer:112-119): unlimited first
    dimension, chunk = one frame, and a chunk B-tree whose keys are the frame offsets in increasing order."""
    path = str(tmp_path / "c.h5")
    w = h5.Writer(path)
    dset = w.create_appendable("turbulence", (9, 4, 6), "f4")
    for i in range(130):
        dset.append(np.full((9, 4, 6), i, np.float32))
    w.close()
    r = h5._Reader(path)
    (name, header), = [e for m in r.messages(r.root_header) if m[0] == h5.MSG_SYMTAB
                      for e in r.group_entries(*struct.unpack_from("<QQ", m[2], 0))]
    msgs = {t: body for t, _, body in r.messages(header)}
    shape, maxshape = r.shape_of(msgs[h5.MSG_DATASPACE])
    assert shape == (130, 9, 4, 6) and maxshape == (h5.UNDEF, 9, 4, 6)
    lay = msgs[h5.MSG_LAYOUT]
    assert lay[0] == 3 and lay[1] == 2 and lay[2] == 5 and struct.unpack_from("<5I", lay, 11) == (1, 9, 4, 6, 4)
    root = r.at(struct.unpack_from("<Q", lay, 3)[0])
    assert bytes(r.buf[root:root + 4]) == b"TREE" and r.buf[root + 4] == 1 and r.buf[root + 5] == 1   # chunk tree, level 1
    chunks = r.chunks(struct.unpack_from("<Q", lay, 3)[0], 5)
    assert [c[0] for c in chunks] == [(i, 0, 0, 0, 0) for i in range(130)]
    assert all(c[1] == 9 * 4 * 6 * 4 and c[2] == 0 and c[3] % 8 == 0 for c in chunks)
    assert [c[3] for c in chunks] == sorted(c[3] for c in chunks)      # frames lie in the file in arrival order


def test_rejects_what_it_cannot_store(tmp_path):
    w = h5.Writer(str(tmp_path / "x.h5"))
    with pytest.raises(TypeError):
        w.create_dataset("c", np.zeros(3, np.complex64))
    w.create_dataset("a", np.zeros(3, np.float32))
    with pytest.raises(ValueError):
        w.create_dataset("a", np.zeros(3, np.float32))
    with pytest.raises(ValueError):
        w.create_appendable("g/x", (2,))
    w.abort()


def test_gzip_frames_use_the_standard_deflate_filter(tmp_path):
    """`compression: gzip` in the case config: every frame is one deflate-compressed chunk, announced by a version-1
    filter pipeline message (filter id 1) -- the layout h5py produces for compression="gzip"."""
    import zlib

    path = str(tmp_path / "z.h5")
    w = h5.Writer(path)
    dset = w.create_appendable("turbulence", (9, 8, 10), "f4", gzip=4)
    frames = np.repeat(np.arange(70, dtype=np.float32), 9 * 8 * 10).reshape(70, 9, 8, 10)     # compressible
    for f in frames:
        dset.append(f)
    w.close()
    assert os.path.getsize(path) < frames.nbytes // 4
    d = h5.read(path)
    assert np.array_equal(d["turbulence"], frames)
    r = h5._Reader(path)
    (name, header), = [e for m in r.messages(r.root_header) if m[0] == h5.MSG_SYMTAB
                      for e in r.group_entries(*struct.unpack_from("<QQ", m[2], 0))]
    msgs = {t: body for t, _, body in r.messages(header)}
    assert r.filters_of(msgs[h5.MSG_FILTERS]) == [(1, (4,))]
    lay = msgs[h5.MSG_LAYOUT]
    offs, nbytes, fmask, addr = r.chunks(struct.unpack_from("<Q", lay, 3)[0], 5)[69]
    assert offs[0] == 69 and fmask == 0
    assert zlib.decompress(bytes(r.buf[addr:addr + nbytes])) == frames[69].tobytes()


def test_random_case_files_round_trip():
    """Property test: any mix of contiguous datasets (float / integer dtypes, ranks 0-4, empty ones), one or two appendable
    datasets with or without deflate, numeric and string attributes comes back exactly, whatever order things were created in."""
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")
    import tempfile

    dtypes = ["f4", "f8", "i4", "i8", "u1", "i2"]

    @hyp.settings(max_examples=40, deadline=None)
    @hyp.given(st.data())
    def run(data):
        rng = np.random.default_rng(data.draw(st.integers(0, 2**31)))
        n_plain = data.draw(st.integers(0, 11))
        plain = {}
        for i in range(n_plain):
            shape = tuple(data.draw(st.lists(st.integers(0, 5), min_size=0, max_size=4)))
            dt = data.draw(st.sampled_from(dtypes))
            plain[f"d{i:02d}_{data.draw(st.sampled_from(['a', 'Z', '_', 'm']))}"] = (rng.standard_normal(shape) * 100).astype(dt)
        n_app = data.draw(st.integers(0, 2))
        apps = {}
        for i in range(n_app):
            fshape = tuple(data.draw(st.lists(st.integers(1, 4), min_size=1, max_size=3)))
            n_frames = data.draw(st.sampled_from([0, 1, 2, 63, 64, 65, 130]))
            apps[f"t{i}"] = ((rng.standard_normal((n_frames,) + fshape) * 3).astype("f4"), data.draw(st.sampled_from([None, 1, 6])))
        attrs = {"s": data.draw(st.text(max_size=60)), "v": rng.standard_normal(data.draw(st.integers(1, 9))),
                 "n": np.arange(data.draw(st.integers(1, 5)), dtype=np.int32)}
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, "x.h5")
            w = h5.Writer(path)
            handles = {k: w.create_appendable(k, v[0].shape[1:], "f4", gzip=v[1]) for k, v in apps.items()}
            names = list(plain)
            for step in range(max([len(names)] + [len(v[0]) for v in apps.values()])):   # interleave creations and appends
                if step < len(names):
                    w.create_dataset(names[step], plain[names[step]])
                for k, v in apps.items():
                    if step < len(v[0]):
                        handles[k].append(v[0][step])
            for k, v in attrs.items():
                w.set_attr(k, v)
            w.close()
            got = h5.read(path)
            assert set(got) == set(plain) | set(apps) | {"attrs"}
            for k, v in plain.items():
                assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v), k
            for k, v in apps.items():
                assert got[k].shape == v[0].shape and np.array_equal(got[k], v[0]), k
            assert got["attrs"]["s"] == attrs["s"] and np.array_equal(got["attrs"]["v"], attrs["v"])
            assert got["attrs"]["n"].dtype == np.int32 and np.array_equal(got["attrs"]["n"], attrs["n"])
            del got

    run()
