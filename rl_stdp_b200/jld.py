"""Read-only importer for the reference's `.jld` result / elite files.

The reference saves its optimisation results with JLD.jl (`save("results.jld", "results", results)`,
Julia/Modules/save_param.jl:4-32; `Actors/cartpoleElites/*.jld`): HDF5 files (superblock version 0
behind a 512-byte Julia header) whose datasets are small, uncompressed arrays of Float64 / Int64 /
strings tied together by HDF5 object references — a `Dict{String,Any}` is a two-field compound
(keys, values) of references, an `Array{Array{Float64}}` an array of references.  There is no HDF5
library in this image, so this module parses exactly the subset of the HDF5 file format those
files use (version-1 object headers, compact / contiguous / chunked+deflate layouts, fixed-point,
floating-point, string, reference and compound datatypes, old-style symbol-table groups and inline
link messages, and fractal-heap link storage — direct blocks, scanned without the name index — for
files with more than eight top-level variables).

    from rl_stdp_b200.jld import load
    results = load("results_LIFbase888.jld")["results"]      # {'weights': [array, ...], 'fit': array, ...}
    genes = results["weights"][0]                             # -> lif.set_weight(arch, genes)

Arrays come back with Julia's shape (column-major data, transposed view of the HDF5 dataset).
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"


class JldError(ValueError):
    pass


class _Ref(int):
    """An HDF5 object reference (address of an object header)."""


class _File:
    def __init__(self, data: bytes):
        self.d = data
        off = data.find(_SIG)
        if off < 0:
            raise JldError("not an HDF5 / JLD file")
        ver = data[off + 8]
        if ver not in (0, 1):
            raise JldError(f"HDF5 superblock version {ver} is not supported (the reference's files are version 0)")
        if data[off + 13] != 8 or data[off + 14] != 8:
            raise JldError("only 8-byte offsets / lengths are supported")
        p = off + 24 + (4 if ver == 1 else 0)
        self.base = struct.unpack_from("<Q", data, p)[0]
        root = p + 32  # root group symbol table entry
        self.root_ohdr = struct.unpack_from("<Q", data, root + 8)[0]
        self._cache = {}

    # ---------------------------------------------------------------- object headers (version 1)
    def messages(self, addr: int):
        d, a = self.d, addr + self.base
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", d, a)
        if ver != 1:
            raise JldError(f"object header version {ver} at {addr} is not supported")
        blocks = [(a + 16, a + 16 + hsize)]
        out = []
        while blocks and len(out) < nmsg:
            pos, end = blocks.pop(0)
            while pos + 8 <= end and len(out) < nmsg:
                mtype, size, flags = struct.unpack_from("<HHB", d, pos)
                body = pos + 8
                out.append((mtype, flags, body, size))
                if mtype == 0x10:  # continuation
                    ca, cl = struct.unpack_from("<QQ", d, body)
                    blocks.append((ca + self.base, ca + self.base + cl))
                pos = body + size
        return out

    # ---------------------------------------------------------------- datatypes
    def datatype(self, pos: int):
        """-> (kind, size, extra, end_pos) with kind in int/uint/float/string/ref/compound/other."""
        d = self.d
        cv, b0, b1, b2, size = struct.unpack_from("<BBBBI", d, pos)
        cls, ver = cv & 0x0F, cv >> 4
        p = pos + 8
        if cls == 0:
            return ("int" if (b0 & 0x08) else "uint"), size, None, p + 4
        if cls == 1:
            return "float", size, None, p + 12
        if cls == 3:
            return "string", size, None, p
        if cls == 7:
            return "ref", size, None, p
        if cls == 6:
            nmem = b0 | (b1 << 8)
            members = []
            for _ in range(nmem):
                e = d.index(b"\x00", p)
                name = d[p:e].decode("utf-8", "replace")
                if ver >= 3:
                    p = e + 1
                    nb = 1 if size < 256 else (2 if size < 65536 else 4)
                    moff = int.from_bytes(d[p:p + nb], "little")
                    p += nb
                else:
                    p += ((e - p) // 8 + 1) * 8
                    moff = struct.unpack_from("<I", d, p)[0]
                    p += 4
                    if ver == 1:
                        p += 28  # dimensionality, reserved, permutation, reserved, 4 dimension sizes
                mk, ms, mx, p = self.datatype(p)
                members.append((name, moff, mk, ms, mx))
            return "compound", size, members, p
        if cls == 9:  # variable length: a string (type 1) or a sequence of the base type
            is_str = (b0 & 0x0F) == 1
            bk, bs, bx, p = self.datatype(p)
            return "vlen", size, (is_str, (bk, bs, bx)), p
        return "other", size, None, p

    def _shared_datatype(self, body: int):
        ver = self.d[body]
        addr = struct.unpack_from("<Q", self.d, body + (8 if ver == 1 else 2))[0]
        for mtype, flags, b, _ in self.messages(addr):
            if mtype == 0x03 and not (flags & 0x02):
                return self.datatype(b)
        raise JldError("committed datatype without a datatype message")

    # ---------------------------------------------------------------- datasets
    def _chunks(self, btree: int, rank: int, out: list):
        d, a = self.d, btree + self.base
        if d[a:a + 4] != b"TREE":
            raise JldError("chunk index is not a version-1 B-tree")
        ntype, level, nent = struct.unpack_from("<BBH", d, a + 4)
        p = a + 24
        for _ in range(nent):
            csize, fmask = struct.unpack_from("<II", d, p)
            offs = struct.unpack_from(f"<{rank + 1}Q", d, p + 8)
            p += 8 + 8 * (rank + 1)
            child = struct.unpack_from("<Q", d, p)[0]
            p += 8
            if level == 0:
                out.append((offs[:rank], csize, fmask, child))
            else:
                self._chunks(child, rank, out)

    def raw_dataset(self, addr: int):
        """-> (h5 dims, (kind, size, extra), bytes)"""
        d = self.d
        dims, dt, data, filters = (), None, None, []
        layout = None
        for mtype, flags, body, size in self.messages(addr):
            if mtype == 0x01:
                ver, rank, fl = d[body], d[body + 1], d[body + 2]
                p = body + (8 if ver == 1 else 4)
                dims = struct.unpack_from(f"<{rank}Q", d, p) if rank else ()
                if ver >= 2 and d[body + 3] == 2:  # null dataspace: an empty Julia array
                    dims = (0,)
            elif mtype == 0x03:
                dt = self._shared_datatype(body)[:3] if (flags & 0x02) else self.datatype(body)[:3]
            elif mtype == 0x0B:
                ver, nf = d[body], d[body + 1]
                p = body + (8 if ver == 1 else 2)
                for _ in range(nf):
                    fid, nlen, _, ncv = struct.unpack_from("<HHHH", d, p)
                    p += 8
                    if ver == 1 or fid >= 256:
                        p += (nlen + 7) // 8 * 8 if ver == 1 else nlen
                    p += 4 * ncv + (4 if (ver == 1 and ncv % 2) else 0)
                    filters.append(fid)
            elif mtype == 0x08:
                layout = body
        if dt is None or layout is None:
            raise JldError(f"object at {addr} is not a dataset")
        n = int(np.prod(dims)) if dims else 1
        esize = dt[1]
        ver, cls = d[layout], d[layout + 1]
        if ver != 3:
            raise JldError(f"data layout version {ver} is not supported")
        if cls == 0:
            sz = struct.unpack_from("<H", d, layout + 2)[0]
            data = d[layout + 4:layout + 4 + sz]
        elif cls == 1:
            a, sz = struct.unpack_from("<QQ", d, layout + 2)
            data = b"" if a == 0xFFFFFFFFFFFFFFFF else d[a + self.base:a + self.base + sz]
        elif cls == 2:
            rank1 = d[layout + 2]
            bt = struct.unpack_from("<Q", d, layout + 3)[0]
            cdims = struct.unpack_from(f"<{rank1}I", d, layout + 11)[:-1]
            rank = rank1 - 1
            full = np.zeros(dims, dtype=np.uint8).reshape(-1) if False else None
            arr = np.zeros(tuple(dims) + (esize,), np.uint8)
            chunks = []
            if bt != 0xFFFFFFFFFFFFFFFF:
                self._chunks(bt, rank, chunks)
            for offs, csize, fmask, a in chunks:
                raw = d[a + self.base:a + self.base + csize]
                for fi, fid in reversed(list(enumerate(filters))):
                    if fmask & (1 << fi):
                        continue
                    if fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 2:  # shuffle
                        raw = np.frombuffer(raw, np.uint8).reshape(esize, -1).T.tobytes()
                    elif fid == 3:  # fletcher32 checksum trailer
                        raw = raw[:-4]
                    else:
                        raise JldError(f"HDF5 filter {fid} is not supported")
                c = np.frombuffer(raw, np.uint8).reshape(tuple(cdims) + (esize,))
                sl = tuple(slice(o, min(o + cd, dm)) for o, cd, dm in zip(offs, cdims, dims))
                arr[sl] = c[tuple(slice(0, s.stop - s.start) for s in sl)]
            data = arr.tobytes()
        else:
            raise JldError(f"data layout class {cls} is not supported")
        if len(data) < n * esize:
            raise JldError(f"dataset at {addr}: {len(data)} bytes for {n} x {esize}")
        return tuple(dims), dt, data[:n * esize]

    # ---------------------------------------------------------------- global heap (variable-length data)
    def _heap_object(self, coll: int, index: int) -> bytes:
        d, a = self.d, coll + self.base
        if d[a:a + 4] != b"GCOL":
            raise JldError("bad global heap collection")
        csize = struct.unpack_from("<Q", d, a + 8)[0]
        p, end = a + 16, a + csize
        while p + 16 <= end:
            idx, _, _, osize = struct.unpack_from("<HHIQ", d, p)
            if idx == index:
                return d[p + 16:p + 16 + osize]
            if idx == 0:
                break
            p += 16 + (osize + 7) // 8 * 8
        raise JldError(f"global heap object {index} not found")

    # ---------------------------------------------------------------- values
    def _decode(self, dims, dt, data):
        kind, size, extra = dt
        if kind == "vlen":
            is_str, base = extra
            n = int(np.prod(dims)) if dims else 1
            vals = []
            for i in range(n):
                ln, coll, idx = struct.unpack_from("<IQI", data, i * size)
                raw = self._heap_object(coll, idx) if (ln and coll) else b""
                if is_str:
                    vals.append(raw[:ln].decode("utf-8", "replace"))
                else:
                    vals.append(self._decode((ln,), base, raw[:ln * base[1]]))
            if not dims:
                return vals[0]
            return vals if len(dims) == 1 else np.array(vals, dtype=object).reshape(dims).T
        if kind in ("int", "uint", "float"):
            code = {"int": "i", "uint": "u", "float": "f"}[kind]
            a = np.frombuffer(data, dtype=f"<{code}{size}")
            if not dims:
                return a[0].item()
            return a.reshape(dims).T.copy()  # HDF5 dims are Julia's reversed
        if kind == "string":
            n = int(np.prod(dims)) if dims else 1
            vals = [data[i * size:(i + 1) * size].split(b"\x00")[0].decode("utf-8", "replace") for i in range(n)]
            return vals[0] if not dims else vals
        if kind == "ref":
            refs = np.frombuffer(data, "<u8")
            vals = [None if r == 0 else self.load(int(r)) for r in refs]
            if not dims:
                return vals[0]
            if len(dims) == 1:
                return vals
            return np.array(vals, dtype=object).reshape(dims).T
        if kind == "compound":
            n = int(np.prod(dims)) if dims else 1
            recs = []
            for i in range(n):
                rec = {}
                for name, moff, mk, ms, mx in extra:
                    # JLD appends "_" to the field names of a Julia struct
                    rec[name[:-1] if name.endswith("_") else name] = self._decode(
                        (), (mk, ms, mx), data[i * size + moff:i * size + moff + ms])
                # JLD's wrapper for Dict (`AssociativeWrapper`: keys, values) -> a Python dict
                if set(rec) == {"keys", "values"} and isinstance(rec["keys"], list):
                    rec = dict(zip(rec["keys"], rec["values"]))
                recs.append(rec)
            return recs[0] if not dims else recs
        raise JldError(f"datatype {kind} is not supported")

    def load(self, addr: int):
        if addr in self._cache:
            return self._cache[addr]
        if any(m[0] == 0x08 for m in self.messages(addr)):
            val = self._decode(*self.raw_dataset(addr))
        else:
            val = self.group(addr)
        self._cache[addr] = val
        return val

    # ---------------------------------------------------------------- groups
    def _link(self, body: int, out: dict):
        d = self.d
        fl = d[body + 1]
        p = body + 2
        ltype = 0
        if fl & 0x08:
            ltype = d[p]; p += 1
        if fl & 0x04:
            p += 8
        if fl & 0x10:
            p += 1
        nb = 1 << (fl & 0x03)
        ln = int.from_bytes(d[p:p + nb], "little"); p += nb
        name = d[p:p + ln].decode("utf-8", "replace"); p += ln
        if ltype == 0:
            out[name] = struct.unpack_from("<Q", d, p)[0]
            return p + 8
        if ltype == 1:  # soft link: skipped
            return p + 2 + struct.unpack_from("<H", d, p)[0]
        return p + 2 + struct.unpack_from("<H", d, p)[0]

    def _dense_links(self, body: int, out: dict):
        """Link Info message: links kept in a fractal heap (groups with more than 8 entries).
        Supported: a heap of direct blocks (one root direct block, or one indirect block of them)."""
        d = self.d
        fl = d[body + 1]
        p = body + 2 + (8 if fl & 1 else 0)
        heap, bt = struct.unpack_from("<QQ", d, p)
        if heap == 0xFFFFFFFFFFFFFFFF:
            return
        h = heap + self.base
        if d[h:h + 4] != b"FRHP":
            raise JldError("bad fractal heap")
        max_direct = struct.unpack_from("<Q", d, h + 120)[0]
        max_heap_bits = struct.unpack_from("<H", d, h + 128)[0]
        root_block = struct.unpack_from("<Q", d, h + 132)[0]
        cur_rows = struct.unpack_from("<H", d, h + 140)[0]
        off_bytes = (max_heap_bits + 7) // 8
        width = struct.unpack_from("<H", d, h + 110)[0]
        start_size = struct.unpack_from("<Q", d, h + 112)[0]
        io_filter_len = struct.unpack_from("<H", d, h + 7)[0]
        blocks = []  # (first heap offset, size, file address) of the direct blocks
        if cur_rows == 0:
            blocks.append((0, start_size, root_block + self.base))
        else:
            ib = root_block + self.base
            if d[ib:ib + 4] != b"FHIB":
                raise JldError("bad fractal-heap indirect block")
            q = ib + 5 + 8 + off_bytes
            hoff = 0
            for row in range(cur_rows):
                bsize = start_size if row < 2 else start_size << (row - 1)
                if bsize > max_direct:
                    raise JldError("fractal heap with nested indirect blocks is not supported")
                for _ in range(width):
                    ba = struct.unpack_from("<Q", d, q)[0]
                    q += 8 + (12 if io_filter_len else 0)
                    if ba != 0xFFFFFFFFFFFFFFFF:
                        blocks.append((hoff, bsize, ba + self.base))
                    hoff += bsize
        # The name index (version-2 B-tree) is not needed to enumerate: the managed objects of a
        # direct block are the link messages themselves, packed from the start of its payload.
        has_cksum = bool(d[h + 9] & 0x02)
        for b0, bs, ba in blocks:
            if d[ba:ba + 4] != b"FHDB":
                raise JldError("bad fractal-heap direct block")
            q = ba + 5 + 8 + off_bytes + (4 if has_cksum else 0)
            while q < ba + bs and d[q] == 1:  # link message version 1; free space is zero-filled
                q = self._link(q, out)

    def links(self, addr: int):
        d = self.d
        out = {}
        for mtype, flags, body, size in self.messages(addr):
            if mtype == 0x06:  # link message
                self._link(body, out)
            elif mtype == 0x02:
                self._dense_links(body, out)
            elif mtype == 0x11:  # symbol table: version-1 B-tree of symbol nodes + local heap
                bt, heap = struct.unpack_from("<QQ", d, body)
                ha = heap + self.base
                if d[ha:ha + 4] != b"HEAP":
                    raise JldError("bad local heap")
                hdata = struct.unpack_from("<Q", d, ha + 24)[0] + self.base
                self._symnodes(bt, hdata, out)
        return out

    def _symnodes(self, bt: int, hdata: int, out: dict):
        d, a = self.d, bt + self.base
        if d[a:a + 4] == b"TREE":
            level, nent = struct.unpack_from("<BH", d, a + 5)
            p = a + 24 + 8
            for _ in range(nent):
                child = struct.unpack_from("<Q", d, p)[0]
                self._symnodes(child, hdata, out)
                p += 16
        elif d[a:a + 4] == b"SNOD":
            nsym = struct.unpack_from("<H", d, a + 6)[0]
            p = a + 8
            for _ in range(nsym):
                noff, oh = struct.unpack_from("<QQ", d, p)
                e = d.index(b"\x00", hdata + noff)
                out[d[hdata + noff:e].decode("utf-8", "replace")] = oh
                p += 40
        else:
            raise JldError("bad group B-tree node")

    def group(self, addr: int):
        return {k: _Lazy(self, v) for k, v in self.links(addr).items()}


class _Lazy:
    """A named child of a group, loaded on first use (groups such as `_refs` hold every object twice)."""

    def __init__(self, f: _File, addr: int):
        self._f, self._addr = f, addr

    def get(self):
        return self._f.load(self._addr)


def load(path: str, name: str | None = None):
    """`JLD.load(path[, name])`: the top-level variables of a .jld file as a dict (or one of them).
    JLD's own bookkeeping entries (`_refs`, `_types`, `_creator`, `_require`) are left out."""
    with open(path, "rb") as fh:
        f = _File(fh.read())
    links = {k: v for k, v in f.links(f.root_ohdr).items() if not k.startswith("_")}
    if name is not None:
        if name not in links:
            raise KeyError(name)
        return f.load(links[name])
    return {k: f.load(v) for k, v in links.items()}
