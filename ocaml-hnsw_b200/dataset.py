"""Mirror of benchmark/dataset.ml: synthetic datasets, brute-force ground truth, recall."""
import numpy as np

from . import _capi as capi


def brute_force_knn_l2(train, test, k, device=0, return_ids=False, metric=capi.L2):
    """brute_force_knn_l2 (benchmark/dataset.ml:15-30) on the GPU: distances [nq][k] ascending
    (the reference returns only distances; ids on request)."""
    import ctypes as C
    train, test = capi.as_mat(train), capi.as_mat(test, train.shape[1])
    ids = np.empty((test.shape[0], k), np.int32)
    d = np.empty((test.shape[0], k), np.float32)
    capi.check(capi.lib().hnswb200_bruteforce_knn(capi.ptr(train), train.shape[0], capi.ptr(test), test.shape[0],
                                                  train.shape[1], k, metric, device, capi.ptr(ids), capi.ptr(d)))
    return (ids, d) if return_ids else d


class Dataset:
    """Dataset.t (benchmark/dataset.ml:32-45)."""

    def __init__(self, train, test, test_distances, distance="euclidean", test_ids=None):
        self.train, self.test, self.test_distances, self.distance = train, test, test_distances, distance
        self.test_ids = test_ids

    @staticmethod
    def read(train_fbin, test_fbin, k, limit_train=None, limit_test=None, device=0):
        """Dataset.read (benchmark/dataset.ml:76-102) over .fbin files; the ground-truth distances are
        recomputed with the exact GPU scan (the reference takes them from the HDF5 `distances` set)."""
        train, test = read_fbin(train_fbin, limit_train), read_fbin(test_fbin, limit_test)
        ids, d = brute_force_knn_l2(train, test, k, device, return_ids=True)
        return Dataset(train, test, d, test_ids=ids)

    @staticmethod
    def read_hdf5(path, limit_train=None, limit_test=None):
        """Dataset.read (benchmark/dataset.ml:76-102) on an ann-benchmarks HDF5 file: the `train`, `test` and
        `distances` float32 datasets and the `distance` attribute of the root group, cropped like the reference's
        ?limit_train / ?limit_test (sub_right on the Fortran-layout matrix = the first rows here).  No libhdf5
        in this image: the file is parsed by hdf5min.py."""
        from .hdf5min import Hdf5File
        with Hdf5File(path) as f:
            distance = f.attrs.get("distance")
            if not isinstance(distance, str):
                raise ValueError(f"{path}: no `distance` string attribute")
            # Distance.of_string (dataset.ml:10-12): "euclidean" -> Euclidean, anything else -> Unknown x (kept as is)
            train = np.ascontiguousarray(f.read("train", limit_train), np.float32)
            test = np.ascontiguousarray(f.read("test", limit_test), np.float32)
            dists = np.ascontiguousarray(f.read("distances", limit_test), np.float32)
            ids = np.ascontiguousarray(f.read("neighbors", limit_test), np.int32) if "neighbors" in f else None
        return Dataset(train, test, dists, distance, test_ids=ids)

    @staticmethod
    def random(dim, num_train, num_test, k, seed=(1234, 4321), device=0):
        """Dataset.random (benchmark/dataset.ml:47-58): Lacaml.S.Mat.random = uniform [-1, 1)."""
        train = (np.random.default_rng(seed[0]).random((num_train, dim), dtype=np.float32) * 2 - 1)
        test = (np.random.default_rng(seed[1]).random((num_test, dim), dtype=np.float32) * 2 - 1)
        ids, d = brute_force_knn_l2(train, test, k, device, return_ids=True)
        return Dataset(train, test, d, test_ids=ids)


def write_fbin(path, a):
    """Raw vector file: int32 n, int32 dim, then n x dim float32 (the big-ann-benchmarks .fbin layout).
    The reference reads ann-benchmarks HDF5 (benchmark/dataset.ml:76-102); this image has no HDF5
    library, so real data sets come in through this format (scripts/hdf5_to_fbin.py converts where
    h5py exists)."""
    a = np.ascontiguousarray(a, np.float32)
    with open(path, "wb") as f:
        np.array(a.shape, np.int32).tofile(f)
        a.tofile(f)


def read_fbin(path, max_rows=None):
    with open(path, "rb") as f:
        n, dim = np.fromfile(f, np.int32, 2).tolist()
        if max_rows is not None:
            n = min(n, max_rows)
        a = np.fromfile(f, np.float32, n * dim)
    if a.size != n * dim:
        raise ValueError("truncated .fbin file")
    return a.reshape(n, dim)


def sift_like(n, dim, latent=16, seed=1234, noise=0.05, proj_seed=99):
    """SURVEY.md 8d generator (b): low intrinsic dimension, integer-valued like SIFT.
    latent-d standard normal x fixed random projection + noise, affine to [0,218], rounded."""
    proj = np.random.default_rng(proj_seed).standard_normal((latent, dim)).astype(np.float32)
    rng = np.random.default_rng(seed)
    out = np.empty((n, dim), np.float32)
    step = 1 << 18
    for s in range(0, n, step):
        m = min(step, n - s)
        z = rng.standard_normal((m, latent)).astype(np.float32)
        x = z @ proj + noise * rng.standard_normal((m, dim)).astype(np.float32)
        out[s:s + m] = x
    scale = 4.0 * np.sqrt(latent)            # ~4 sigma of the projected coordinates
    out = np.clip((out / scale + 1.0) * 109.0, 0, 218)
    return np.rint(out).astype(np.float32)


class Recall:
    @staticmethod
    def compute(expected, got, epsilon=1e-8):
        """Recall.compute (benchmark/dataset.ml:105-127): fraction of returned distances
        <= the true k-th distance + epsilon; NaN never counts."""
        import ctypes as C
        e = capi.as_mat(expected)
        g = capi.as_mat(got)
        if e.shape != g.shape:
            raise ValueError("Recall.compute: arrrays have unequal shapes")     # dataset.ml:112 (sic)
        out = C.c_double()
        capi.check(capi.lib().hnswb200_recall(capi.ptr(e), capi.ptr(g), e.shape[0], e.shape[1], epsilon, C.byref(out)))
        return out.value

    @staticmethod
    def ids(expected_ids, got_ids):
        """id-set recall@k against exact ids."""
        k = expected_ids.shape[1]
        hit = 0
        for a, b in zip(expected_ids, got_ids):
            hit += len(set(a.tolist()) & set(b[b >= 0].tolist()))
        return hit / (k * expected_ids.shape[0])
