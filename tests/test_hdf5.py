"""Dataset.read (benchmark/dataset.ml:76-102) over ann-benchmarks HDF5 files without libhdf5: the package's own
reader (hdf5min.py) against files laid out the way h5py lays them out (written by the same module's writer —
no HDF5 library exists in this image to produce or cross-check them; the byte layout follows the HDF5 File
Format Specification and is spot-checked below field by field)."""
import struct

import numpy as np
import pytest

import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200.hdf5min import Hdf5File, Hdf5Unsupported, write_hdf5, SIGNATURE


def _ann_file(tmp_path, name="a.hdf5", **kw):
    rng = np.random.default_rng(5)
    d = {"train": rng.standard_normal((300, 24)).astype(np.float32),
         "test": rng.standard_normal((40, 24)).astype(np.float32),
         "distances": np.sort(rng.random((40, 10)).astype(np.float32), axis=1),
         "neighbors": rng.integers(0, 300, (40, 10)).astype(np.int32)}
    p = str(tmp_path / name)
    write_hdf5(p, d, {"distance": "euclidean", "dimension": 24, "point_type": "float"}, **kw)
    return p, d


@pytest.mark.parametrize("kw", [dict(), dict(vlen_strings=False), dict(chunk_rows=64), dict(chunk_rows=37, deflate=True),
                                dict(chunk_rows=50, deflate=True, shuffle=True)])
def test_round_trip(tmp_path, kw):
    p, d = _ann_file(tmp_path, **kw)
    with Hdf5File(p) as f:
        assert f.keys() == sorted(d)
        assert f.attrs["distance"] == "euclidean" and f.attrs["dimension"] == 24 and f.attrs["point_type"] == "float"
        for k, v in d.items():
            assert f.shape(k) == v.shape
            a = f[k]
            assert a.dtype == v.dtype and np.array_equal(a, v)
        assert np.array_equal(f.read("train", 17), d["train"][:17])
        assert "nothing" not in f
        with pytest.raises(KeyError):
            f["nothing"]


def test_dataset_read_mirrors_the_reference(tmp_path):
    """Dataset.read ?limit_train ?limit_test f: train / test / distances cropped to the first rows, distance string kept."""
    p, d = _ann_file(tmp_path)
    ds = H.Dataset.read_hdf5(p, limit_train=100, limit_test=7)
    assert ds.distance == "euclidean"
    assert np.array_equal(ds.train, d["train"][:100]) and np.array_equal(ds.test, d["test"][:7])
    assert np.array_equal(ds.test_distances, d["distances"][:7]) and np.array_equal(ds.test_ids, d["neighbors"][:7])
    full = H.Dataset.read_hdf5(p)
    assert full.train.shape == (300, 24) and full.test_distances.shape == (40, 10)
    # an unknown metric name is kept, not rejected (Distance.of_string, dataset.ml:10-12)
    q = str(tmp_path / "b.hdf5")
    write_hdf5(q, d, {"distance": "hamming"})
    assert H.Dataset.read_hdf5(q).distance == "hamming"
    write_hdf5(q, d, {})
    with pytest.raises(ValueError):
        H.Dataset.read_hdf5(q)


def test_layout_fields_follow_the_specification(tmp_path):
    """Spot checks of the bytes against the HDF5 File Format Specification (superblock v0, symbol table entry,
    object header v1, data layout v3): what any HDF5 library would need to find in these places."""
    p, d = _ann_file(tmp_path)
    b = open(p, "rb").read()
    assert b[:8] == SIGNATURE and b[8] == 0                       # superblock version 0
    assert b[13] == 8 and b[14] == 8                              # sizes of offsets / lengths
    assert struct.unpack_from("<HH", b, 16) == (4, 16)            # group leaf / internal node K
    base, free, eof, drv = struct.unpack_from("<QQQQ", b, 24)
    assert base == 0 and free == 2**64 - 1 and drv == 2**64 - 1 and eof == len(b)
    name_off, root, cache = struct.unpack_from("<QQI", b, 56)
    assert cache == 1 and b[root] == 1                            # cached symbol table; object header version 1
    btree, heap = struct.unpack_from("<QQ", b, 56 + 24)
    assert b[btree:btree + 4] == b"TREE" and b[heap:heap + 4] == b"HEAP"
    nmsg, = struct.unpack_from("<H", b, root + 2)
    mtype, msize = struct.unpack_from("<HH", b, root + 16)
    assert mtype == 0x0011 and msize == 16 and nmsg == 4          # symbol table message + three attributes
    assert struct.unpack_from("<QQ", b, root + 24) == (btree, heap)
    snod, = struct.unpack_from("<Q", b, btree + 24 + 8)
    assert b[snod:snod + 4] == b"SNOD" and struct.unpack_from("<H", b, snod + 6)[0] == 4


def test_errors(tmp_path):
    p = str(tmp_path / "x.bin")
    open(p, "wb").write(b"not hdf5 at all" * 400)
    with pytest.raises(ValueError):
        Hdf5File(p)
    good, _ = _ann_file(tmp_path)
    b = bytearray(open(good, "rb").read())
    q = str(tmp_path / "trunc.hdf5")
    open(q, "wb").write(b[:2000])
    with pytest.raises(ValueError):
        with Hdf5File(q) as f:
            f["train"]
    b[8] = 7                                                      # unknown superblock version
    open(q, "wb").write(b)
    with pytest.raises(Hdf5Unsupported):
        Hdf5File(q)


def test_hdf5_to_search_pipeline_shapes(tmp_path):
    """The reader's output feeds the C ABI unchanged: C-contiguous float32 [n][dim] (a Lacaml.S.mat dim x n)."""
    p, d = _ann_file(tmp_path)
    ds = H.Dataset.read_hdf5(p)
    for a in (ds.train, ds.test, ds.test_distances):
        assert a.flags["C_CONTIGUOUS"] and a.dtype == np.float32
