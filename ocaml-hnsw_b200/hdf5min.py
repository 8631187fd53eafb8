"""A small reader (and a writer for the same subset) of HDF5 files as ann-benchmarks publishes them, so that
Dataset.read (benchmark/dataset.ml:76-102: `train`, `test`, `distances` float32 datasets and the `distance`
string attribute of the root group) works in an image without libhdf5 / h5py.

Implemented from the HDF5 File Format Specification (version 3.0), the parts h5py's defaults produce:
  * superblock versions 0-1 (root symbol-table entry) and 2-3 (root object header address);
  * old-style groups (symbol table message -> v1 B-tree of SNOD nodes + local heap) and new-style groups with
    compact storage (link messages in the object header);
  * object headers version 1 and 2 ("OHDR"), continuation blocks;
  * dataspace v1/v2 (simple), datatypes: fixed-point, IEEE float, fixed-length string, variable-length string
    (global heap);
  * data layout v3: compact, contiguous, chunked (v1 B-tree chunk index) with the deflate and shuffle filters;
    layout v1/v2 contiguous;
  * attribute messages v1-v3.
Anything else (dense groups in fractal heaps, layout v4 chunk indexes, other filters, compound types) raises
Hdf5Unsupported with the name of the feature.  Not part of the hot path: host-side ingestion only."""
import struct
import zlib

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class Hdf5Unsupported(ValueError):
    pass


class _Type:
    def __init__(self, kind, dtype=None, size=0, base=None):
        self.kind, self.dtype, self.size, self.base = kind, dtype, size, base     # kind: "num" | "str" | "vlen_str"


class Hdf5File:
    """f = Hdf5File(path); f.keys(); f["train"] -> numpy array; f.attrs -> dict; f.shape("train") without reading."""

    def __init__(self, path):
        self.path = path
        self._f = open(path, "rb")
        self._parse_superblock()
        self._root = self._object(self._root_addr)
        self.attrs = {k: self._attr_value(v) for k, v in self._root["attrs"].items()}
        self._links = self._group_links(self._root)

    def close(self):
        self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- low level -------------------------------------------------------------------------------------
    def _read(self, addr, n):
        self._f.seek(self._base + addr)
        b = self._f.read(n)
        if len(b) != n:
            raise ValueError(f"{self.path}: truncated file (wanted {n} bytes at {addr})")
        return b

    def _off(self, b, o):          # an "offset"-sized field
        return int.from_bytes(b[o:o + self._so], "little")

    def _len(self, b, o):          # a "length"-sized field
        return int.from_bytes(b[o:o + self._sl], "little")

    def _parse_superblock(self):
        self._base = 0
        start = None
        for cand in (0, 512, 1024, 2048, 4096):      # the superblock may sit at 0 or at a power of two >= 512
            self._f.seek(cand)
            if self._f.read(8) == SIGNATURE:
                start = cand
                break
        if start is None:
            raise ValueError(f"{self.path}: not an HDF5 file")
        self._f.seek(start)
        head = self._f.read(128)
        ver = head[8]
        if ver in (0, 1):
            self._so, self._sl = head[13], head[14]
            o = 24 + (4 if ver == 1 else 0)
            base = int.from_bytes(head[o:o + self._so], "little")
            o += 4 * self._so                          # base, free-space, end of file, driver info
            # root group symbol table entry: link name offset, object header address, cache type, reserved, scratch
            self._root_addr = int.from_bytes(head[o + self._so:o + 2 * self._so], "little")
        elif ver in (2, 3):
            self._so, self._sl = head[9], head[10]
            o = 12
            base = int.from_bytes(head[o:o + self._so], "little")
            self._root_addr = int.from_bytes(head[o + 3 * self._so:o + 4 * self._so], "little")
        else:
            raise Hdf5Unsupported(f"superblock version {ver}")
        if self._so != 8 or self._sl != 8:
            raise Hdf5Unsupported(f"offset/length sizes {self._so}/{self._sl}")
        self._base = base if base != UNDEF else 0
        if start and self._base == 0:
            self._base = start

    # ---- object headers --------------------------------------------------------------------------------
    def _object(self, addr):
        """-> dict(messages=[(type, flags, bytes)], attrs={name: (type, shape, raw)})"""
        msgs = []
        head = self._read(addr, 16)
        if head[:4] == b"OHDR":
            self._messages_v2(addr, msgs)
        else:
            if head[0] != 1:
                raise Hdf5Unsupported(f"object header version {head[0]}")
            nmsg, = struct.unpack_from("<H", head, 2)
            size, = struct.unpack_from("<I", head, 8)
            self._messages_v1(addr + 16, size, msgs, nmsg)
        obj = {"messages": msgs, "attrs": {}}
        for t, fl, d in msgs:
            if t == 0x000C:
                name, val = self._attribute(d)
                obj["attrs"][name] = val
            elif t == 0x0015:                             # attribute info: dense storage if a fractal heap is named
                o = 2 + (2 if d[1] & 1 else 0)
                if self._off(d, o) != UNDEF:
                    raise Hdf5Unsupported("dense attribute storage (fractal heap)")
        return obj

    def _messages_v1(self, addr, size, out, budget):
        blocks = [(addr, size)]
        while blocks:
            a, sz = blocks.pop(0)
            buf = self._read(a, sz)
            o = 0
            while o + 8 <= sz and len(out) < budget + 64:
                t, n, fl = struct.unpack_from("<HHB", buf, o)
                d = buf[o + 8:o + 8 + n]
                o += 8 + n
                if t == 0x0010:
                    blocks.append((self._off(d, 0), self._len(d, self._so)))
                elif t != 0:
                    out.append((t, fl, d))

    def _messages_v2(self, addr, out):
        head = self._read(addr, 64)
        flags = head[5]
        o = 6
        if flags & 0x20:
            o += 16
        if flags & 0x10:
            o += 4
        szf = 1 << (flags & 3)
        size0 = int.from_bytes(head[o:o + szf], "little")
        o += szf
        blocks = [(addr + o, size0)]
        track = bool(flags & 0x04)
        while blocks:
            a, sz = blocks.pop(0)
            buf = self._read(a, sz)
            p = 0
            while p + 4 <= sz:
                t = buf[p]
                n, = struct.unpack_from("<H", buf, p + 1)
                fl = buf[p + 3]
                p += 4 + (2 if track else 0)
                d = buf[p:p + n]
                p += n
                if t == 0x10:
                    ca, cl = self._off(d, 0), self._len(d, self._so)
                    if self._read(ca, 4) != b"OCHK":
                        raise ValueError("bad object header continuation block")
                    blocks.append((ca + 4, cl - 8))    # minus signature and checksum
                elif t != 0:
                    out.append((t, fl, d))

    # ---- groups ------------------------------------------------------------------------------------------
    def _group_links(self, obj):
        links = {}
        for t, fl, d in obj["messages"]:
            if t == 0x0011:                               # symbol table: v1 B-tree + local heap
                btree, heap = self._off(d, 0), self._off(d, self._so)
                hb = self._read(heap, 8 + 3 * 8)
                if hb[:4] != b"HEAP":
                    raise ValueError("bad local heap")
                data_addr = self._off(hb, 8 + 2 * self._sl)
                data_size = self._len(hb, 8)
                names = self._read(data_addr, data_size)
                self._walk_group_btree(btree, names, links)
            elif t == 0x0006:                             # link message (compact new-style group)
                ver, lf = d[0], d[1]
                o = 2
                ltype = 0
                if lf & 0x08:
                    ltype = d[o]; o += 1
                if lf & 0x04:
                    o += 8
                if lf & 0x10:
                    o += 1
                nl = 1 << (lf & 3)
                ln = int.from_bytes(d[o:o + nl], "little"); o += nl
                name = d[o:o + ln].decode("utf-8"); o += ln
                if ltype == 0:
                    links[name] = self._off(d, o)
            elif t == 0x0002:                             # link info: dense storage if a fractal heap is named
                o = 2 + (8 if d[1] & 1 else 0)
                if self._off(d, o) != UNDEF:
                    raise Hdf5Unsupported("dense link storage (fractal heap); re-save the file with fewer than 9 root members or libver='earliest'")
        return links

    def _walk_group_btree(self, addr, names, links):
        node = self._read(addr, 8 + 2 * self._so)
        if node[:4] != b"TREE" or node[4] != 0:
            raise ValueError("bad group B-tree node")
        level, used = node[5], struct.unpack_from("<H", node, 6)[0]
        body = self._read(addr + 8 + 2 * self._so, (2 * used + 1) * 8)
        for i in range(used):
            child = self._off(body, (2 * i + 1) * 8)
            if level > 0:
                self._walk_group_btree(child, names, links)
                continue
            sn = self._read(child, 8)
            if sn[:4] != b"SNOD":
                raise ValueError("bad symbol table node")
            nsym, = struct.unpack_from("<H", sn, 6)
            ents = self._read(child + 8, nsym * 40)
            for k in range(nsym):
                e = ents[k * 40:(k + 1) * 40]
                no, oa = self._off(e, 0), self._off(e, 8)
                end = names.index(b"\0", no)
                links[names[no:end].decode("utf-8")] = oa

    def keys(self):
        return sorted(self._links)

    def __contains__(self, name):
        return name in self._links

    # ---- datatypes, dataspaces, attributes ---------------------------------------------------------------
    def _datatype(self, d):
        cls, ver = d[0] & 0x0F, d[0] >> 4
        bits = d[1] | (d[2] << 8) | (d[3] << 16)
        size, = struct.unpack_from("<I", d, 4)
        if cls == 0:                                       # fixed point
            if size not in (1, 2, 4, 8):
                raise Hdf5Unsupported(f"{size}-byte integer")
            dt = np.dtype(("i" if bits & 0x08 else "u") + str(size)).newbyteorder(">" if bits & 1 else "<")
            return _Type("num", dt, size)
        if cls == 1:                                       # IEEE float
            if size not in (2, 4, 8):
                raise Hdf5Unsupported(f"{size}-byte float")
            return _Type("num", np.dtype("f" + str(size)).newbyteorder(">" if bits & 1 else "<"), size)
        if cls == 3:
            return _Type("str", None, size)
        if cls == 9:
            if (bits & 0x0F) != 1:
                raise Hdf5Unsupported("variable-length sequence datatype")
            return _Type("vlen_str", None, size)
        raise Hdf5Unsupported(f"datatype class {cls}")

    def _dataspace(self, d):
        ver, rank, flags = d[0], d[1], d[2]
        if ver == 1:
            o = 8
        elif ver == 2:
            if d[3] == 2:
                return None                                # null dataspace
            o = 4
        else:
            raise Hdf5Unsupported(f"dataspace version {ver}")
        return tuple(self._len(d, o + 8 * i) for i in range(rank))

    def _attribute(self, d):
        ver = d[0]
        nsz, tsz, ssz = struct.unpack_from("<HHH", d, 2)
        if ver == 1:
            pad = lambda n: (n + 7) & ~7
            o = 8
        elif ver == 2:
            pad = lambda n: n
            o = 8
        elif ver == 3:
            pad = lambda n: n
            o = 9
        else:
            raise Hdf5Unsupported(f"attribute message version {ver}")
        name = d[o:o + nsz].split(b"\0")[0].decode("utf-8"); o += pad(nsz)
        ty = self._datatype(d[o:o + tsz]); o += pad(tsz)
        shape = self._dataspace(d[o:o + ssz]); o += pad(ssz)
        return name, (ty, shape, d[o:])

    def _global_heap_object(self, addr, index):
        head = self._read(addr, 16)
        if head[:4] != b"GCOL":
            raise ValueError("bad global heap collection")
        size = self._len(head, 8)
        buf = self._read(addr, size)
        o = 16
        while o + 16 <= size:
            idx, = struct.unpack_from("<H", buf, o)
            osz = self._len(buf, o + 8)
            if idx == 0:
                break
            if idx == index:
                return buf[o + 16:o + 16 + osz]
            o += 16 + ((osz + 7) & ~7)
        raise ValueError("global heap object not found")

    def _attr_value(self, v):
        ty, shape, raw = v
        if shape is None:
            return None
        count = int(np.prod(shape)) if shape else 1
        if ty.kind == "num":
            a = np.frombuffer(raw[:count * ty.size], ty.dtype).astype(ty.dtype.newbyteorder("="))
            return a.reshape(shape) if shape else a[0].item()
        if ty.kind == "str":
            vals = [raw[i * ty.size:(i + 1) * ty.size].split(b"\0")[0].decode("utf-8") for i in range(count)]
        else:
            vals = []
            for i in range(count):
                e = raw[i * 16:(i + 1) * 16]
                ln, = struct.unpack_from("<I", e, 0)
                gaddr, gidx = self._off(e, 4), struct.unpack_from("<I", e, 12)[0]
                vals.append(self._global_heap_object(gaddr, gidx)[:ln].decode("utf-8"))
        return vals[0] if not shape else np.array(vals, dtype=object).reshape(shape)

    # ---- datasets ------------------------------------------------------------------------------------------
    def _dataset_info(self, name):
        if name not in self._links:
            raise KeyError(name)
        obj = self._object(self._links[name])
        ty = shape = layout = None
        filters = []
        for t, fl, d in obj["messages"]:
            if t == 0x0001:
                shape = self._dataspace(d)
            elif t == 0x0003:
                ty = self._datatype(d)
            elif t == 0x0008:
                layout = d
            elif t == 0x000B:
                filters = self._filters(d)
        if ty is None or shape is None or layout is None:
            raise ValueError(f"{name}: not a dataset")
        if ty.kind != "num":
            raise Hdf5Unsupported(f"{name}: non-numeric dataset")
        return ty, shape, layout, filters

    def shape(self, name):
        return self._dataset_info(name)[1]

    def _filters(self, d):
        ver, n = d[0], d[1]
        o = 8 if ver == 1 else 2
        out = []
        for _ in range(n):
            fid, = struct.unpack_from("<H", d, o); o += 2
            nlen = 0
            if ver == 1 or fid >= 256:
                nlen, = struct.unpack_from("<H", d, o); o += 2
            fl, ncd = struct.unpack_from("<HH", d, o); o += 4
            o += (nlen + 7) & ~7 if ver == 1 else nlen
            cd = struct.unpack_from("<%dI" % ncd, d, o); o += 4 * ncd
            if ver == 1 and ncd % 2:
                o += 4
            if fid not in (1, 2):
                if fl & 1:
                    continue                                # optional filter
                raise Hdf5Unsupported(f"filter id {fid}")
            out.append((fid, cd))
        return out

    def __getitem__(self, name):
        return self.read(name)

    def read(self, name, max_rows=None):
        """The dataset as a numpy array (native byte order); `max_rows` crops the first dimension (contiguous
        data is then only read that far — Dataset.read's ?limit_train / ?limit_test)."""
        ty, shape, lay, filters = self._dataset_info(name)
        rows = shape[0] if shape else 1
        if max_rows is not None and shape:
            rows = min(rows, max_rows)
        out_shape = ((rows,) + tuple(shape[1:])) if shape else ()
        row_elems = int(np.prod(shape[1:])) if len(shape) > 1 else 1
        ver = lay[0]
        if ver in (1, 2):
            rank, cls = lay[1], lay[2]
            if cls != 1:
                raise Hdf5Unsupported(f"data layout version {ver} class {cls}")
            addr = self._off(lay, 8)
            return self._contiguous(addr, ty, out_shape, rows * row_elems)
        if ver != 3:
            raise Hdf5Unsupported(f"data layout version {ver}" + (" (libver='latest' chunk index)" if ver == 4 else ""))
        cls = lay[1]
        if cls == 0:
            n, = struct.unpack_from("<H", lay, 2)
            a = np.frombuffer(lay[4:4 + n], ty.dtype)
            return a[:rows * row_elems].astype(ty.dtype.newbyteorder("=")).reshape(out_shape)
        if cls == 1:
            return self._contiguous(self._off(lay, 2), ty, out_shape, rows * row_elems)
        if cls == 2:
            nd = lay[2]
            btree = self._off(lay, 3)
            cdims = struct.unpack_from("<%dI" % nd, lay, 3 + self._so)
            full = np.zeros(shape, ty.dtype.newbyteorder("="))
            if btree != UNDEF:
                self._walk_chunks(btree, nd, cdims[:-1], ty, filters, full, rows)
            return full[:rows] if shape else full
        raise Hdf5Unsupported(f"data layout class {cls}")

    def _contiguous(self, addr, ty, out_shape, count):
        if addr == UNDEF:
            return np.zeros(out_shape, ty.dtype.newbyteorder("="))
        self._f.seek(self._base + addr)
        a = np.fromfile(self._f, ty.dtype, count)
        if a.size != count:
            raise ValueError(f"{self.path}: truncated dataset")
        return a.astype(ty.dtype.newbyteorder("="), copy=False).reshape(out_shape)

    def _walk_chunks(self, addr, nd, cshape, ty, filters, full, rows):
        node = self._read(addr, 8 + 2 * self._so)
        if node[:4] != b"TREE" or node[4] != 1:
            raise ValueError("bad chunk B-tree node")
        level, used = node[5], struct.unpack_from("<H", node, 6)[0]
        ksz = 8 + 8 * nd
        body = self._read(addr + 8 + 2 * self._so, used * (ksz + self._so) + ksz)
        for i in range(used):
            k = body[i * (ksz + self._so):]
            csize, fmask = struct.unpack_from("<II", k, 0)
            offs = struct.unpack_from("<%dQ" % nd, k, 8)[:-1]
            child = self._off(k, ksz)
            if level > 0:
                self._walk_chunks(child, nd, cshape, ty, filters, full, rows)
                continue
            if offs and offs[0] >= rows:
                continue
            raw = self._read(child, csize)
            for j, (fid, cd) in reversed(list(enumerate(filters))):
                if fmask & (1 << j):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    es = cd[0] if cd else ty.size
                    raw = np.frombuffer(raw, np.uint8).reshape(es, -1).T.tobytes()
            chunk = np.frombuffer(raw, ty.dtype, int(np.prod(cshape))).reshape(cshape)
            sl_full = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cshape, full.shape))
            sl_chunk = tuple(slice(0, s.stop - s.start) for s in sl_full)
            full[sl_full] = chunk[sl_chunk]


# ---- writer (the same subset, the way h5py's defaults lay a small file out) ---------------------------------
def write_hdf5(path, datasets, attrs=None, vlen_strings=True, chunk_rows=None, deflate=False, shuffle=False):
    """Write `datasets` {name: 2-D / 1-D float32 | int32 | float64 array} and root `attrs` {name: str | int | float}
    as an HDF5 file: superblock v0, old-style root group (symbol table, v1 B-tree, local heap), v1 object headers,
    contiguous layout; string attributes as variable-length UTF-8 strings in a global heap (what
    `f.attrs["distance"] = "euclidean"` produces in h5py) or, with vlen_strings=False, fixed-length strings.
    chunk_rows = r stores every dataset in chunks of r rows behind a v1 B-tree (at most 64 chunks), optionally
    through the shuffle and deflate filters (`create_dataset(..., chunks=..., compression="gzip", shuffle=True)`)."""
    attrs = attrs or {}
    names = sorted(datasets)
    if len(names) > 8:
        raise ValueError("write_hdf5: at most 8 datasets (one symbol table node)")
    buf = bytearray()

    def align():
        while len(buf) % 8:
            buf.append(0)

    def dtype_msg(a):
        dt = a.dtype
        if dt.kind == "f":
            exp, man, bias = {4: (8, 23, 127), 8: (11, 52, 1023)}[dt.itemsize]
            return struct.pack("<BBBBI", 0x11, 0x20, 8 * dt.itemsize - 1, 0, dt.itemsize) + \
                struct.pack("<HHBBBBI", 0, 8 * dt.itemsize, man, exp, 0, man, bias)
        if dt.kind in "iu":
            return struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
        raise ValueError(f"write_hdf5: dtype {dt}")

    def space_msg(shape):
        return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", s) for s in shape)

    def message(t, data, flags=0):
        data = bytes(data) + b"\0" * (-len(data) % 8)
        return struct.pack("<HHB3x", t, len(data), flags) + data

    def object_header(msgs):
        body = b"".join(msgs)
        return struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body)) + body

    def pad8(b):
        return bytes(b) + b"\0" * (-len(b) % 8)

    buf += b"\0" * 96                                        # superblock, filled in last
    # global heap for variable-length strings
    gheap_addr, gobjs = None, []
    str_attrs = {k: v for k, v in attrs.items() if isinstance(v, str)}
    if vlen_strings and str_attrs:
        gheap_addr = len(buf)
        body = bytearray()
        for i, (k, v) in enumerate(sorted(str_attrs.items()), 1):
            b = v.encode("utf-8")
            gobjs.append((k, i, len(b)))
            body += struct.pack("<HH4xQ", i, 1, len(b)) + pad8(b)
        size = max(4096, 16 + len(body) + 16)
        free = size - 16 - len(body)
        body += struct.pack("<HH4xQ", 0, 0, free) + b"\0" * (free - 16)
        buf += b"GCOL" + struct.pack("<B3xQ", 1, size) + body
    align()

    def attr_msg(name, v):
        nm = name.encode("utf-8") + b"\0"
        if isinstance(v, str):
            if vlen_strings:
                idx, ln = next((i, l) for k, i, l in gobjs if k == name)
                ty = struct.pack("<BBBBI", 0x19, 0x01, 0x01, 0, 16) + struct.pack("<BBBBI", 0x10, 0, 0, 0, 1) + struct.pack("<HH", 0, 8)
                data = struct.pack("<IQI", ln, gheap_addr, idx)
            else:
                b = v.encode("utf-8") + b"\0"
                ty = struct.pack("<BBBBI", 0x13, 0x00, 0, 0, len(b))
                data = b
            sp = struct.pack("<BBBB4x", 1, 0, 0, 0)
        else:
            a = np.asarray(v, np.float64 if isinstance(v, float) else np.int64)
            ty, sp, data = dtype_msg(a), struct.pack("<BBBB4x", 1, 0, 0, 0), a.tobytes()
        return message(0x000C, struct.pack("<BxHHH", 1, len(nm), len(ty), len(sp)) + pad8(nm) + pad8(ty) + pad8(sp) + data)

    # raw data, then one object header per dataset
    data_addr, obj_addr = {}, {}
    arrays = {}
    for n in names:
        a = np.ascontiguousarray(datasets[n])
        if a.dtype.byteorder == ">":
            a = a.astype(a.dtype.newbyteorder("<"))
        arrays[n] = a
        align()
        data_addr[n] = len(buf)
        buf += a.tobytes()
    chunk_tree = {}
    if chunk_rows:
        for n in names:                                      # re-lay the raw data out as chunks + one leaf B-tree node
            a = arrays[n] if arrays[n].ndim > 1 else arrays[n].reshape(-1, 1)
            cshape = (chunk_rows,) + a.shape[1:]
            entries = []
            for r0 in range(0, a.shape[0], chunk_rows):
                c = np.zeros(cshape, a.dtype)
                part = a[r0:r0 + chunk_rows]
                c[:len(part)] = part
                raw = c.tobytes()
                if shuffle:
                    raw = np.frombuffer(raw, np.uint8).reshape(-1, a.dtype.itemsize).T.tobytes()
                if deflate:
                    raw = zlib.compress(raw, 4)
                align()
                entries.append((len(raw), r0, len(buf)))
                buf.extend(raw)
            if len(entries) > 64:
                raise ValueError("write_hdf5: more than 64 chunks")
            align()
            nd = a.ndim + 1
            chunk_tree[n] = (len(buf), cshape)
            node = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(entries), UNDEF, UNDEF)
            for size, r0, addr in entries:
                node += struct.pack("<II", size, 0) + struct.pack("<%dQ" % nd, r0, *([0] * (nd - 1))) + struct.pack("<Q", addr)
            node += struct.pack("<II", 0, 0) + struct.pack("<%dQ" % nd, (a.shape[0] + chunk_rows - 1) // chunk_rows * chunk_rows, *([0] * (nd - 1)))
            buf.extend(node)
    for n in names:
        a = arrays[n]
        align()
        obj_addr[n] = len(buf)
        if chunk_rows:
            a2 = a if a.ndim > 1 else a.reshape(-1, 1)
            taddr, cshape = chunk_tree[n]
            lay = struct.pack("<BBBQ", 3, 2, a2.ndim + 1, taddr) + struct.pack("<%dI" % (a2.ndim + 1), *cshape, a.dtype.itemsize)
            msgs = [message(0x0001, space_msg(a2.shape)), message(0x0003, dtype_msg(a), 1), message(0x0008, lay)]
            fl = []
            if shuffle:
                fl.append(struct.pack("<HHHH", 2, 0, 0, 1) + struct.pack("<II", a.dtype.itemsize, 0))
            if deflate:
                fl.append(struct.pack("<HHHH", 1, 0, 0, 1) + struct.pack("<II", 4, 0))
            if fl:
                msgs.append(message(0x000B, struct.pack("<BB6x", 1, len(fl)) + b"".join(fl)))
            buf += object_header(msgs)
        else:
            buf += object_header([message(0x0001, space_msg(a.shape)), message(0x0003, dtype_msg(a), 1),
                                  message(0x0008, struct.pack("<BBQQ", 3, 1, data_addr[n], a.nbytes))])
    # local heap with the link names (offset 0 holds the empty string), symbol table node, B-tree, root header
    align()
    heap_data = bytearray(b"\0" * 8)
    name_off = {}
    for n in names:
        name_off[n] = len(heap_data)
        heap_data += pad8(n.encode("utf-8") + b"\0")
    free_off = len(heap_data)
    heap_data += struct.pack("<QQ", 1, 16)                   # one free block: no next block (1), 16 bytes
    heap_data_addr = len(buf)
    buf += heap_data
    align()
    heap_addr = len(buf)
    buf += b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free_off, heap_data_addr)
    align()
    snod_addr = len(buf)
    snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
    for n in names:
        snod += struct.pack("<QQII16x", name_off[n], obj_addr[n], 0, 0)
    snod += b"\0" * (40 * (8 - len(names)))
    buf += snod
    align()
    btree_addr = len(buf)
    last = name_off[names[-1]] if names else 0
    buf += b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod_addr, last)
    buf += b"\0" * (2 * 16 * 16)                             # room for the remaining keys / children of a 2K-entry node
    align()
    root_addr = len(buf)
    buf += object_header([message(0x0011, struct.pack("<QQ", btree_addr, heap_addr))] +
                         [attr_msg(k, v) for k, v in sorted(attrs.items())])
    align()
    eof = len(buf)
    sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, root_addr, 1, 0) + struct.pack("<QQ", btree_addr, heap_addr)
    assert len(sb) == 96
    buf[:96] = sb
    with open(path, "wb") as f:
        f.write(buf)
