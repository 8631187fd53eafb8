"""Host-side mirror of `Hnsw.Ba` = MakeBatch(EuclideanBa) (lib/hnsw.ml:729-778,817-819), the
functor instantiation north_star names, over the same C ABI with the HNSW_BA flavour:
ties accepted in the beam (lib/hnsw.ml:494-506), M links for a new node on every layer with
caps 2M / M (lib/hnsw.ml:753-758), distances-only results padded with +inf (:769-777).

Q5 (lib/hnsw.ml:519-525: nearest_k keeps the k FARTHEST of the ef set when ef > k) is a defect
of the reference that the benchmark never exercises (ef = k); it is not reproduced: knn returns
the k nearest."""
import numpy as np

from . import _capi as capi
from . import ohnsw


class Ba:
    @staticmethod
    def build(values, *, num_neighbours, num_neighbours_build, levels=None, seed=0, device=0):
        """Hnsw.Ba.build (lib/hnsw.ml:753-761)."""
        values = capi.as_mat(values)
        h = ohnsw.Hgraph(values.shape[1], capi.L2, num_neighbours, num_neighbours_build, seed, device,
                         flavour=capi.FLAVOUR_HNSW_BA)
        lv = None if levels is None else np.ascontiguousarray(levels, np.int32)
        capi.check(capi.lib().hnswb200_build(h._h, capi.ptr(values), values.shape[0], capi.ptr(lv)))
        return h

    @staticmethod
    def knn(hgraph, point, *, num_neighbours_search=5, num_neighbours):
        """Hnsw.Ba.knn (lib/hnsw.ml:763-767) -> [(node, distance_to_target)], nearest first.  Nodes are numbered
        from 1 as in Hnsw.Ba (lib/hnsw.ml:313-325: node k is `Mat.col m k`), i.e. row index + 1."""
        t = np.asarray(point, np.float32)[None, :]
        ef = max(num_neighbours_search, num_neighbours)
        ids, d = ohnsw.knn_batch_bigarray(hgraph, t, k=num_neighbours, ef=ef)
        return [(int(i) + 1, float(x)) for i, x in zip(ids[0], d[0]) if i >= 0]

    @staticmethod
    def knn_batch(hgraph, batch, *, num_neighbours_search, num_neighbours):
        """Hnsw.Ba.knn_batch (lib/hnsw.ml:769-777) -> distances [nq][k], +inf padded."""
        ef = max(num_neighbours_search, num_neighbours)
        _, d = ohnsw.knn_batch_bigarray(hgraph, batch, k=num_neighbours, ef=ef)
        return d
