"""Flat graph exchange format: the parity vehicle between a graph built by
the reference (exported by a functor over Hnsw_algo.KNN_HGRAPH / a walker over Ohnsw.Hgraph.t,
see INTEGRATION.md), the oracle, and the GPU layout.

File layout (little endian):
    magic    8 bytes  b"HNSWB200"
    header   int64 x 8: version(1), n, dim, id_base, max_layer, entry_point, M, has_levels
    per layer l = 0..max_layer:  int64 offsets[n+1], int32 nbrs[offsets[n]]   (list order kept)
    levels   int32[n]            (if has_levels)
    vectors  float32[n][dim]     (optional, if the file continues)
"""
import numpy as np

MAGIC = b"HNSWB200"


class FlatGraph:
    def __init__(self, n, max_layer, entry, offsets, nbrs, levels=None):
        self.n, self.max_layer, self.entry = int(n), int(max_layer), int(entry)
        self.offsets, self.nbrs, self.levels = offsets, nbrs, levels

    def row(self, layer, node):
        o = self.offsets[layer]
        return self.nbrs[layer][o[node]:o[node + 1]]

    def degree(self, layer):
        return np.diff(self.offsets[layer])

    def is_symmetric(self, layer):
        """Graph.Test.invariant (lib/ohnsw.ml:217-225)."""
        o, a = self.offsets[layer], self.nbrs[layer]
        src = np.repeat(np.arange(self.n, dtype=np.int64), np.diff(o))
        fwd = set(zip(src.tolist(), a.tolist()))
        return all((b, s) in fwd for s, b in fwd)


def write_graph(path, g, dim, M, id_base=0, vectors=None):
    with open(path, "wb") as f:
        f.write(MAGIC)
        has_levels = g.levels is not None
        np.array([1, g.n, dim, id_base, g.max_layer, g.entry, M, int(has_levels)], np.int64).tofile(f)
        for l in range(g.max_layer + 1):
            np.ascontiguousarray(g.offsets[l], np.int64).tofile(f)
            np.ascontiguousarray(g.nbrs[l], np.int32).tofile(f)
        if has_levels:
            np.ascontiguousarray(g.levels, np.int32).tofile(f)
        if vectors is not None:
            np.ascontiguousarray(vectors, np.float32).tofile(f)


def read_graph(path):
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError("not an HNSWB200 graph file")
        ver, n, dim, id_base, max_layer, entry, M, has_levels = np.fromfile(f, np.int64, 8).tolist()
        if ver != 1:
            raise ValueError("unsupported graph file version")
        offs, nbrs = [], []
        for _ in range(max_layer + 1):
            o = np.fromfile(f, np.int64, n + 1)
            offs.append(o)
            nbrs.append(np.fromfile(f, np.int32, int(o[-1])))
        levels = np.fromfile(f, np.int32, n) if has_levels else None
        rest = np.fromfile(f, np.float32)
        vectors = rest.reshape(n, dim) if rest.size == n * dim else None
    return FlatGraph(n, max_layer, entry, offs, nbrs, levels), dict(dim=dim, M=M, id_base=id_base), vectors


def to_dot(g, layer=0, positions=None, highlight=None):
    """Graphviz dump of one layer (test/test.ml:6-56 show_hgraph / show_neighbours): nodes, one
    undirected edge per symmetric link, optional 2-D positions and a highlighted node set."""
    out = ["graph hnsw {", "  node [shape=circle, fontsize=8];"]
    o, a = g.offsets[layer], g.nbrs[layer]
    present = [i for i in range(g.n) if layer == 0 or (g.levels is not None and g.levels[i] >= layer) or o[i + 1] > o[i]]
    hl = set(highlight or [])
    for i in present:
        attrs = []
        if positions is not None:
            attrs.append(f'pos="{float(positions[i][0]):.3f},{float(positions[i][1]):.3f}!"')
        if i in hl:
            attrs.append('color=red')
        if i == g.entry:
            attrs.append('shape=doublecircle')
        out.append(f"  {i} [{', '.join(attrs)}];" if attrs else f"  {i};")
    for i in present:
        for j in a[o[i]:o[i + 1]].tolist():
            if i < j:
                out.append(f"  {i} -- {j};")
    out.append("}")
    return "\n".join(out)
