"""The C-ABI library loads and exports exactly what include/hnsw_b200.h declares; argument
validation works without a GPU; compute entry points fail loudly without one (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "hnsw_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hnswb200_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    names = _declared()
    assert len(names) >= 20
    L = capi.lib()
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/hnsw_b200.h but not exported"
    assert sorted(capi.SIGNATURES) == names, "python binding and header disagree"


def test_no_oracle_in_product():
    """The product never links, loads, includes or imports the oracle (comments may cite it)."""
    pkg = os.path.join(ROOT, "ocaml-hnsw_b200")
    pat = re.compile(r"liboracle|import\s+oracle|from\s+oracle|#include\s*[\"<][^\n]*oracle|oracle/|orc_[a-z]+")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h", ".ml", ".c")) or f == "Makefile":
                for line in open(os.path.join(dirpath, f)):
                    code = line.split("//")[0].split("#")[0] if not line.lstrip().startswith("#include") else line
                    assert not pat.search(code), (dirpath, f, line)
    assert "oracle" not in os.popen(f"ldd {capi.LIB_PATH}").read()


def test_argument_validation_without_gpu():
    h = C.c_void_p()
    L = capi.lib()
    assert L.hnswb200_create(C.byref(h), 0, 0, 16, 100, 0, 0) == capi.EINVAL
    assert b"dim" in L.hnswb200_last_error()
    assert L.hnswb200_create(C.byref(h), 128, 7, 16, 100, 0, 0) == capi.EINVAL
    assert L.hnswb200_create(C.byref(h), 128, 0, 1, 100, 0, 0) == capi.EINVAL      # level_mult = 1/ln 1
    assert L.hnswb200_create(None, 128, 0, 16, 100, 0, 0) == capi.EINVAL
    assert L.hnswb200_search(None, None, 1, 1, 1, 0, None, None) == capi.EINVAL
    with pytest.raises(ValueError, match="unequal shapes"):
        H.Recall.compute(np.zeros((2, 3), np.float32), np.zeros((2, 2), np.float32))


def test_recall_compute_host():          # benchmark/dataset.ml:105-127
    exp = np.array([[1, 2, 3], [1, 2, 3]], np.float32)
    got = np.array([[1, 2, 3.5], [np.nan, 1, 3]], np.float32)
    assert H.Recall.compute(exp, got) == pytest.approx((2 / 3 + 2 / 3) / 2)


def test_compute_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.HnswB200Error, match="no CUDA device"):
        H.Ohnsw.Hgraph(128)
    with pytest.raises(capi.HnswB200Error, match="no CUDA device"):
        H.brute_force_knn_l2(np.zeros((4, 8), np.float32), np.zeros((2, 8), np.float32), 2)


def test_graph_file_roundtrip(tmp_path):
    g = H.FlatGraph(3, 1, 2, [np.array([0, 1, 3, 4]), np.array([0, 0, 0, 0])],
                    [np.array([1, 0, 2, 1], np.int32), np.zeros(0, np.int32)], np.array([0, 0, 1], np.int32))
    vec = np.arange(6, dtype=np.float32).reshape(3, 2)
    p = str(tmp_path / "g.hnswb200")
    H.write_graph(p, g, dim=2, M=3, vectors=vec)
    g2, meta, v2 = H.read_graph(p)
    assert (g2.n, g2.max_layer, g2.entry, meta["dim"], meta["M"]) == (3, 1, 2, 2, 3)
    assert g2.row(0, 1).tolist() == [0, 2] and np.array_equal(v2, vec) and g2.levels.tolist() == [0, 0, 1]
    assert g2.is_symmetric(0)


def test_fbin_roundtrip_and_dot(tmp_path):
    a = np.random.default_rng(0).random((37, 5), dtype=np.float32)
    H.write_fbin(tmp_path / "x.fbin", a)
    assert np.array_equal(H.read_fbin(tmp_path / "x.fbin"), a)
    assert H.read_fbin(tmp_path / "x.fbin", max_rows=10).shape == (10, 5)
    g = H.FlatGraph(3, 0, 1, [np.array([0, 1, 3, 4])], [np.array([1, 0, 2, 1], np.int32)], np.zeros(3, np.int32))
    dot = H.to_dot(g, 0, positions=[(0, 0), (1, 0), (2, 0)], highlight=[2])
    assert "0 -- 1;" in dot and "1 -- 2;" in dot and dot.count("--") == 2 and "doublecircle" in dot
    p = tmp_path / "g.bin"
    H.write_graph(p, g, dim=5, M=2, vectors=a[:3])
    g2, meta, vec = H.read_graph(p)
    assert g2.entry == 1 and meta["dim"] == 5 and np.array_equal(vec, a[:3]) and np.array_equal(g2.nbrs[0], g.nbrs[0])


def test_pinned_buffer_registry():
    """capi.is_pinned: a slice is pinned when its bytes lie inside a buffer registered with host_register (what lets
    the sharded host mirror hand a query slice to the kernel without a copy)."""
    a = np.zeros((100, 8), np.float32)
    capi._PINNED[a.ctypes.data] = a.nbytes            # what host_register records (no GPU here)
    try:
        assert capi.is_pinned(a) and capi.is_pinned(a[10:20]) and capi.is_pinned(a[99:])
        assert not capi.is_pinned(a[:, :4])            # not contiguous
        assert not capi.is_pinned(np.zeros((4, 8), np.float32))
    finally:
        capi._PINNED.pop(a.ctypes.data)
    assert not capi.is_pinned(a)
