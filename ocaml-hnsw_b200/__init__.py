"""hnsw_b200 — host-side mirror of lehy/ocaml-hnsw's build + k-NN API over libhnsw_b200.so.

    from ocaml_hnsw_b200 import Ohnsw, Hnsw, Dataset, Recall

`Ohnsw` mirrors lib/ohnsw.ml, `Hnsw.Ba` mirrors lib/hnsw.ml's Bigarray instantiation,
`Dataset` / `Recall` mirror benchmark/dataset.ml.  Everything computes on the GPU through the
C ABI declared in include/hnsw_b200.h; there is no CPU fallback.
"""
from . import _capi as capi
from . import ohnsw as Ohnsw
from . import hnsw as Hnsw
from . import graphio
from .dataset import Dataset, Recall, brute_force_knn_l2, read_fbin, sift_like, write_fbin
from .graphio import FlatGraph, read_graph, to_dot, write_graph

__all__ = ["capi", "Ohnsw", "Hnsw", "Dataset", "Recall", "brute_force_knn_l2", "sift_like", "graphio",
           "FlatGraph", "read_graph", "write_graph", "to_dot", "read_fbin", "write_fbin"]
