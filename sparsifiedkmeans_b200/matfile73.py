"""Minimal reader (and writer, for tests and for users without MATLAB) of MATLAB -v7.3 MAT-files.

The reference's out-of-core mode reads a single dense variable from a -v7.3 file through `matfile`
(private/sampleAndMixFromLargeFile.m:60-107, kmeans_sparsified.m:180-207).  A -v7.3 file is an HDF5 file behind a
512-byte MATLAB header; this image has neither h5py nor libhdf5, so the subset of the HDF5 file format that MATLAB
writes for a numeric matrix is parsed here directly (HDF5 File Format Specification 2.0/3.0):

  superblock version 0/1, 8-byte offsets and lengths, base address = user block (512)
  groups as symbol tables: version-1 B-tree (node type 0) + local heap + SNOD symbol nodes
  version-1 object headers with continuation blocks
  messages: dataspace (v1/v2), datatype (fixed point / IEEE floating point), data layout v3 (compact, contiguous,
  chunked) and v1/v2, filter pipeline v1/v2, attribute v1/v2/v3 (MATLAB_class), symbol table
  chunk index: version-1 B-tree (node type 1), any depth; filters: deflate (zlib), shuffle, fletcher32

MATLAB stores a p x n matrix (column-major) as an HDF5 dataset of shape (n, p) (row-major): the bytes are the same.
`open_matrix(path)` returns the single variable as a (p, n) array without loading it: a contiguous dataset is
memory-mapped in place; a chunked / compressed one is decoded chunk by chunk into a memory-mapped temporary file (so
the matrix still never has to fit in RAM).  Anything outside the subset raises `MatFileError` naming what was met.

Pinned against a file written by MATLAB itself (scipy's test fixture testhdf5_7.4_GLNX86.mat, tests/test_matfile73.py);
the chunked + deflate path, which MATLAB uses for large arrays, is exercised through `write_matrix` below (same
structures: superblock 0, v1 headers, v1 B-trees) -- no MATLAB-written chunked file is available in this image.
"""
from __future__ import annotations

import mmap
import os
import struct
import tempfile
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"


class MatFileError(Exception):
    pass


# ------------------------------------------------------------------------------------------------------------------
# reader
# ------------------------------------------------------------------------------------------------------------------
class _Dataset:
    def __init__(self):
        self.shape = None
        self.dtype = None
        self.layout = None          # ("contiguous", addr, size) | ("compact", bytes) | ("chunked", btree_addr, chunk_dims)
        self.filters = []           # [(id, client values)]
        self.attrs = {}
        self.is_group = False
        self.group = None           # (btree, heap)


class H5Reader:
    def __init__(self, path: str):
        self.path = path
        self.f = open(path, "rb")
        size = os.fstat(self.f.fileno()).st_size
        if size < 64:
            raise MatFileError("file too small to be an HDF5 / MATLAB -v7.3 file")
        self.buf = mmap.mmap(self.f.fileno(), 0, access=mmap.ACCESS_READ)
        self.base = None
        off = 0
        while off + 8 <= size:                                   # the superblock sits at 0, 512, 1024, ...
            if self.buf[off:off + 8] == SIGNATURE:
                self.base = off
                break
            off = 512 if off == 0 else off * 2
        if self.base is None:
            head = bytes(self.buf[:19])
            if head.startswith(b"MATLAB 5.0"):
                raise MatFileError("this is a MATLAB v5/v7 MAT-file, not -v7.3 (HDF5); scipy.io.loadmat reads it")
            raise MatFileError("no HDF5 signature found: not a MATLAB -v7.3 file")
        self._superblock()

    def close(self):
        try:
            self.buf.close()
        finally:
            self.f.close()

    # -- primitives --
    def _at(self, addr: int, n: int) -> bytes:
        a = self.base + addr
        if addr == UNDEF or a + n > len(self.buf):
            raise MatFileError(f"address {addr:#x}+{n} outside the file")
        return self.buf[a:a + n]

    def _superblock(self):
        b = self.buf
        o = self.base + 8
        ver = b[o]
        if ver not in (0, 1):
            raise MatFileError(f"HDF5 superblock version {ver} (MATLAB writes 0); not supported")
        so, sl = b[o + 5], b[o + 6]
        if so != 8 or sl != 8:
            raise MatFileError(f"offsets/lengths of {so}/{sl} bytes; only 8/8 is supported")
        o2 = o + 16 + (4 if ver == 1 else 0)                      # v1 adds indexed-storage K + reserved
        base_addr, _free, _eof, _drv = struct.unpack_from("<QQQQ", b, o2)
        # MATLAB writes base address = size of the user block; addresses in the file are relative to it
        if base_addr not in (0, self.base):
            raise MatFileError(f"base address {base_addr} does not match the superblock position {self.base}")
        ent = o2 + 32
        _name_off, self.root_header, cache, _res = struct.unpack_from("<QQII", b, ent)
        self.root_cache = None
        if cache == 1:
            self.root_cache = struct.unpack_from("<QQ", b, ent + 24)

    # -- object headers --
    def _messages(self, addr: int):
        hdr = self._at(addr, 16)
        if hdr[:4] == b"OHDR":
            raise MatFileError("version-2 object headers (libver='latest') are not supported; MATLAB writes version 1")
        ver, _r, nmsg, _refs, hsize = struct.unpack_from("<BBHII", hdr, 0)
        if ver != 1:
            raise MatFileError(f"object header version {ver} not supported")
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            a, n = blocks.pop(0)
            data = self._at(a, n)
            o = 0
            while o + 8 <= n and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", data, o)
                body = data[o + 8:o + 8 + msize]
                o += 8 + msize
                if mtype == 0x10:                                # continuation
                    ca, cl = struct.unpack_from("<QQ", body, 0)
                    blocks.append((ca, cl))
                out.append((mtype, body))
        return out

    @staticmethod
    def _dataspace(body):
        ver, rank, flags = body[0], body[1], body[2]
        if ver == 1:
            o = 8
        elif ver == 2:
            o = 4
            if body[3] == 2:                                     # null dataspace
                return ()
        else:
            raise MatFileError(f"dataspace message version {ver} not supported")
        return tuple(struct.unpack_from("<" + "Q" * rank, body, o)) if rank else ()

    @staticmethod
    def _datatype(body):
        cv, b0, _b1, _b2, size = struct.unpack_from("<BBBBI", body, 0)
        cls = cv & 0x0F
        order = ">" if (b0 & 1) else "<"
        if cls == 0:                                             # fixed point
            signed = bool(b0 & 0x08)
            if size not in (1, 2, 4, 8):
                raise MatFileError(f"integer of {size} bytes not supported")
            return np.dtype(order + ("i" if signed else "u") + str(size)), size
        if cls == 1:                                             # IEEE floating point
            if size not in (4, 8):
                raise MatFileError(f"floating-point type of {size} bytes not supported")
            return np.dtype(order + "f" + str(size)), size
        if cls == 3:                                             # fixed-length string (attributes)
            return np.dtype(f"S{size}"), size
        return None, size                                        # something this reader does not decode

    @staticmethod
    def _layout(body):
        ver = body[0]
        if ver == 3:
            cls = body[1]
            if cls == 0:
                n = struct.unpack_from("<H", body, 2)[0]
                return ("compact", bytes(body[4:4 + n]))
            if cls == 1:
                addr, size = struct.unpack_from("<QQ", body, 2)
                return ("contiguous", addr, size)
            if cls == 2:
                nd = body[2]
                addr = struct.unpack_from("<Q", body, 3)[0]
                dims = struct.unpack_from("<" + "I" * nd, body, 11)
                return ("chunked", addr, tuple(dims))            # last entry = element size
            raise MatFileError(f"data layout class {cls} not supported")
        if ver in (1, 2):
            nd, cls = body[1], body[2]
            o = 8
            addr = None
            if cls != 0:
                addr = struct.unpack_from("<Q", body, o)[0]
                o += 8
            dims = struct.unpack_from("<" + "I" * nd, body, o)
            o += 4 * nd
            if cls == 0:
                n = struct.unpack_from("<I", body, o)[0]
                return ("compact", bytes(body[o + 4:o + 4 + n]))
            if cls == 1:
                return ("contiguous", addr, None)
            return ("chunked", addr, tuple(dims))
        raise MatFileError(f"data layout message version {ver} not supported (HDF5 1.10 layouts need libhdf5)")

    @staticmethod
    def _filters(body):
        ver, nf = body[0], body[1]
        out = []
        o = 8 if ver == 1 else 2
        for _ in range(nf):
            fid = struct.unpack_from("<H", body, o)[0]
            o += 2
            nlen = 0
            if ver == 1 or fid >= 256:
                nlen = struct.unpack_from("<H", body, o)[0]
                o += 2
            _flags, ncv = struct.unpack_from("<HH", body, o)
            o += 4
            if ver == 1:
                nlen = (nlen + 7) & ~7
            o += nlen
            cv = struct.unpack_from("<" + "I" * ncv, body, o)
            o += 4 * ncv
            if ver == 1 and ncv % 2:
                o += 4
            out.append((fid, cv))
        return out

    def _attribute(self, body):
        ver = body[0]
        if ver == 1:
            nsz, tsz, ssz = struct.unpack_from("<HHH", body, 2)
            o = 8
            pad = lambda v: (v + 7) & ~7                              # noqa: E731
        elif ver in (2, 3):
            nsz, tsz, ssz = struct.unpack_from("<HHH", body, 2)
            o = 8 + (1 if ver == 3 else 0)
            pad = lambda v: v                                         # noqa: E731
        else:
            return None, None
        name = bytes(body[o:o + nsz]).split(b"\0")[0].decode("ascii", "replace")
        o += pad(nsz)
        dt, esize = self._datatype(body[o:o + tsz])
        o += pad(tsz)
        shape = self._dataspace(body[o:o + ssz])
        o += pad(ssz)
        if dt is None:
            return name, None
        count = int(np.prod(shape)) if shape else 1
        raw = bytes(body[o:o + count * esize])
        if dt.kind == "S":
            return name, raw.split(b"\0")[0].decode("ascii", "replace")
        val = np.frombuffer(raw, dtype=dt, count=count)
        return name, (val.reshape(shape) if shape else val[0])

    def object(self, addr: int) -> _Dataset:
        d = _Dataset()
        for mtype, body in self._messages(addr):
            if mtype == 0x01:
                d.shape = self._dataspace(body)
            elif mtype == 0x03:
                d.dtype, _ = self._datatype(body)
            elif mtype == 0x08:
                d.layout = self._layout(body)
            elif mtype == 0x0B:
                d.filters = self._filters(body)
            elif mtype == 0x0C:
                k, v = self._attribute(body)
                if k is not None:
                    d.attrs[k] = v
            elif mtype == 0x11:
                d.is_group = True
                d.group = struct.unpack_from("<QQ", body, 0)
            elif mtype in (0x02, 0x06):
                d.is_group = True                                # new-style group (link info / link): not walked
        return d

    # -- groups (symbol tables) --
    def _heap_data(self, heap_addr: int):
        h = self._at(heap_addr, 32)
        if h[:4] != b"HEAP":
            raise MatFileError("local heap signature missing")
        size, _free, data_addr = struct.unpack_from("<QQQ", h, 8)
        return data_addr, size

    def _group_entries(self, btree_addr: int, heap_addr: int):
        data_addr, hsize = self._heap_data(heap_addr)
        heap = self._at(data_addr, hsize)
        out = []

        def walk(addr):
            node = self._at(addr, 24)
            if node[:4] == b"SNOD":
                nsym = struct.unpack_from("<H", node, 6)[0]
                ents = self._at(addr + 8, 40 * nsym)
                for i in range(nsym):
                    noff, ohdr = struct.unpack_from("<QQ", ents, 40 * i)
                    name = bytes(heap[noff:heap.find(b"\0", noff)]).decode("ascii", "replace")
                    out.append((name, ohdr))
                return
            if node[:4] != b"TREE":
                raise MatFileError("group B-tree signature missing")
            ntype, _level, used = struct.unpack_from("<BBH", node, 4)
            if ntype != 0:
                raise MatFileError("expected a group B-tree node")
            body = self._at(addr + 24, (2 * used + 1) * 8)
            for i in range(used):
                walk(struct.unpack_from("<Q", body, 8 + 16 * i)[0])

        walk(btree_addr)
        return out

    def variables(self):
        """[(name, object header address)] of the root group, MATLAB's bookkeeping groups (#refs#, #subsystem#) left out."""
        root = self.object(self.root_header)
        if root.group is None and self.root_cache is None:
            raise MatFileError("the root group is not a symbol table (new-style groups need libhdf5)")
        bt, hp = root.group if root.group is not None else self.root_cache
        return [(n, a) for n, a in self._group_entries(bt, hp) if not n.startswith("#")]

    # -- data --
    def _chunks(self, btree_addr: int, ndims: int):
        """yields (offsets, stored size, filter mask, address) of every chunk (version-1 B-tree, node type 1)"""
        keysz = 8 + 8 * ndims

        def walk(addr):
            node = self._at(addr, 24)
            if node[:4] != b"TREE":
                raise MatFileError("chunk B-tree signature missing")
            ntype, level, used = struct.unpack_from("<BBH", node, 4)
            if ntype != 1:
                raise MatFileError("expected a chunk B-tree node")
            body = self._at(addr + 24, used * (keysz + 8) + keysz)
            for i in range(used):
                o = i * (keysz + 8)
                csize, mask = struct.unpack_from("<II", body, o)
                offs = struct.unpack_from("<" + "Q" * ndims, body, o + 8)
                child = struct.unpack_from("<Q", body, o + keysz)[0]
                if level > 0:
                    yield from walk(child)
                else:
                    yield offs[:-1], csize, mask, child

        if btree_addr != UNDEF:
            yield from walk(btree_addr)

    @staticmethod
    def _unfilter(raw: bytes, filters, mask: int, esize: int) -> bytes:
        for i in reversed(range(len(filters))):                  # the pipeline is undone back to front
            if mask & (1 << i):
                continue
            fid, _cv = filters[i]
            if fid == 1:
                raw = zlib.decompress(raw)
            elif fid == 2:                                       # shuffle: byte planes -> elements
                n = len(raw) // esize
                a = np.frombuffer(raw, dtype=np.uint8, count=n * esize).reshape(esize, n)
                raw = a.T.tobytes() + raw[n * esize:]
            elif fid == 3:                                       # fletcher32 checksum appended
                raw = raw[:-4]
            else:
                raise MatFileError(f"HDF5 filter {fid} not supported (deflate, shuffle, fletcher32 are)")
        return raw

    def read_matrix(self, addr: int, tmpdir: str | None = None, in_memory_limit: int = 1 << 28):
        """The 2-D numeric dataset at `addr` as a (p, n) array in MATLAB's orientation (a transposed view of the stored
        (n, p) rows): memory-mapped in place when contiguous, decoded into a temporary memory-mapped file otherwise."""
        d = self.object(addr)
        if d.is_group and d.layout is None:
            cls = d.attrs.get("MATLAB_class", "?")
            raise MatFileError(f"the variable is a group (MATLAB class '{cls}', e.g. sparse / cell / struct), not a dense numeric matrix")
        if d.dtype is None or d.dtype.kind not in "fiu":
            raise MatFileError("the variable is not numeric")
        if d.shape is None or len(d.shape) != 2:
            raise MatFileError(f"Error reading file; returned bad size for matrix: {d.shape}")   # kmeans_sparsified.m:204
        n, p = d.shape
        kind = d.layout[0]
        nbytes = n * p * d.dtype.itemsize
        if kind == "contiguous":
            if d.layout[1] == UNDEF:                             # never written: fill value (zeros)
                return np.zeros((n, p), dtype=d.dtype.newbyteorder("=")).T
            a = np.memmap(self.path, dtype=d.dtype, mode="r", offset=self.base + d.layout[1], shape=(n, p))
            return a.T
        if kind == "compact":
            return np.frombuffer(d.layout[1], dtype=d.dtype, count=n * p).reshape(n, p).T
        # chunked
        _, bt, cdims = d.layout
        if len(cdims) != 3:
            raise MatFileError("chunk rank does not match the matrix rank")
        c0, c1, esize = cdims
        native = d.dtype.newbyteorder("=")
        if nbytes <= in_memory_limit:
            out = np.zeros((n, p), dtype=native)
        else:
            fd, name = tempfile.mkstemp(prefix="skm_v73_", suffix=".bin", dir=tmpdir)
            os.close(fd)
            out = np.memmap(name, dtype=native, mode="w+", shape=(n, p))
            os.unlink(name)                                      # the mapping keeps the file alive; nothing to clean up
        for offs, csize, mask, caddr in self._chunks(bt, 3):
            raw = self._unfilter(bytes(self._at(caddr, csize)), d.filters, mask, esize)
            blk = np.frombuffer(raw, dtype=d.dtype, count=c0 * c1).reshape(c0, c1)
            r0, q0 = offs
            r1, q1 = min(r0 + c0, n), min(q0 + c1, p)
            out[r0:r1, q0:q1] = blk[:r1 - r0, :q1 - q0]
        return out.T


def open_matrix(path: str, tmpdir: str | None = None):
    """The single variable of a -v7.3 MAT-file as a (p, n) array (private/sampleAndMixFromLargeFile.m:60-66: more than
    one variable is an error).  Returns (array, variable name)."""
    r = H5Reader(path)
    try:
        names = r.variables()
        if len(names) != 1:
            raise MatFileError("Expected a single variable")     # sampleAndMixFromLargeFile.m:62
        name, addr = names[0]
        return r.read_matrix(addr, tmpdir=tmpdir), name          # a contiguous matrix is its own mapping of the file
    finally:
        r.close()


# ------------------------------------------------------------------------------------------------------------------
# writer (tests, and a way to produce -v7.3 input without MATLAB): one dense 2-D variable, contiguous or chunked
# ------------------------------------------------------------------------------------------------------------------
def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


def _msg(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _datatype_msg(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    size = dt.itemsize
    if dt.kind == "f":
        if size == 8:
            props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            b0, b1 = 0x20, 63
        else:
            props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            b0, b1 = 0x20, 31
        return struct.pack("<BBBBI", 0x11, b0, b1, 0, size) + props
    if dt.kind in "iu":
        b0 = 0x08 if dt.kind == "i" else 0
        return struct.pack("<BBBBI", 0x10, b0, 0, 0, size) + struct.pack("<HH", 0, 8 * size)
    raise MatFileError(f"cannot write dtype {dt}")


def _string_attr(name: str, value: str, version: int = 1) -> bytes:
    nb = name.encode() + b"\0"
    vb = value.encode()
    dtype = struct.pack("<BBBBI", 0x13, 0, 0, 0, len(vb))                  # class 3, null-terminated ASCII
    if version == 1:
        space = struct.pack("<BBB5x", 1, 0, 0)                             # scalar, dataspace version 1
        body = struct.pack("<BxHHH", 1, len(nb), len(dtype), len(space)) + _pad8(nb) + _pad8(dtype) + _pad8(space) + vb
    else:                                                                  # versions 2 / 3: no padding (3 adds the name's character set)
        space = struct.pack("<BBBB", 2, 0, 0, 0)                           # scalar, dataspace version 2
        body = struct.pack("<BBHHH", version, 0, len(nb), len(dtype), len(space)) + (b"\0" if version == 3 else b"")
        body += nb + dtype + space + vb
    return _msg(0x0C, body)


def write_matrix(path: str, X, name: str = "X", chunks=None, compress: int | None = None, shuffle: bool = False,
                 matlab_class: str | None = None, message_versions: str = "old", continuation: bool = False) -> None:
    """Write the (p, n) matrix X as the variable `name` of a -v7.3 MAT-file the way MATLAB lays it out (HDF5 dataset of
    shape (n, p) behind a 512-byte header).  chunks=(rows, cols) of the stored (n, p) array switches to chunked storage;
    compress = zlib level adds the deflate filter (MATLAB's default for large arrays), shuffle the byte-shuffle filter.
    message_versions: "old" = what HDF5 1.6/1.8 in its default (earliest) format and MATLAB write (dataspace 1, layout 3,
    filter pipeline 1, attribute 1); "new" = the later versions of the same messages (dataspace 2, filter pipeline 2,
    attribute 3) inside the same version-1 object headers; "v1layout" = the HDF5 1.4 layout message (version 1)."""
    X = np.asarray(X)
    if X.ndim != 2:
        raise MatFileError("write_matrix needs a 2-D array")
    A = np.ascontiguousarray(X.T)                                           # stored rows = MATLAB columns
    A = A.astype(A.dtype.newbyteorder("<"), copy=False)
    n, p = A.shape
    es = A.dtype.itemsize
    if matlab_class is None:
        matlab_class = {"f8": "double", "f4": "single"}.get(A.dtype.str[1:], A.dtype.name)
    if (compress is not None or shuffle) and chunks is None:
        chunks = (max(1, min(n, (1 << 20) // max(1, p * es))), p)
    out = bytearray()
    base = 512

    def alloc(b: bytes) -> int:                                             # append, 8-byte aligned; address relative to base
        out.extend(b"\0" * (-len(out) % 8))
        a = len(out)
        out.extend(b)
        return a

    out.extend(b"\0" * 96)                                                  # superblock placeholder (56 + 40-byte root entry)
    # ---- raw data ----
    filters = []
    if shuffle:
        filters.append((2, (es,)))
    if compress is not None:
        filters.append((1, (int(compress),)))
    if chunks is None:
        data_addr = alloc(A.tobytes())
        layout = struct.pack("<BBQQ", 3, 1, data_addr, n * p * es)
    else:
        c0, c1 = int(chunks[0]), int(chunks[1])
        recs = []
        for r0 in range(0, n, c0):
            for q0 in range(0, p, c1):
                blk = np.zeros((c0, c1), dtype=A.dtype)
                part = A[r0:r0 + c0, q0:q0 + c1]
                blk[:part.shape[0], :part.shape[1]] = part
                raw = blk.tobytes()
                if shuffle:
                    raw = np.frombuffer(raw, dtype=np.uint8).reshape(-1, es).T.tobytes()
                if compress is not None:
                    raw = zlib.compress(raw, int(compress))
                recs.append(((r0, q0, 0), len(raw), alloc(raw)))
        # version-1 B-tree of the chunks: leaves of <= 64 entries, internal levels above as needed
        K2 = 64
        end_key = (((n + c0 - 1) // c0) * c0, 0, 0)

        def node(level, entries, last_key):
            body = bytearray(b"TREE" + struct.pack("<BBHQQ", 1, level, len(entries), UNDEF, UNDEF))
            for key, size, child in entries:
                body += struct.pack("<II", size, 0) + struct.pack("<QQQ", *key) + struct.pack("<Q", child)
            body += struct.pack("<II", 0, 0) + struct.pack("<QQQ", *last_key)
            return alloc(bytes(body))

        level = 0
        cur = recs
        while True:
            groups = [cur[i:i + K2] for i in range(0, len(cur), K2)] or [[]]
            nxt = []
            for gi, g in enumerate(groups):
                last = groups[gi + 1][0][0] if gi + 1 < len(groups) else end_key
                nxt.append((g[0][0] if g else (0, 0, 0), 0, node(level, g, last)))
            if len(nxt) == 1:
                btree = nxt[0][2]
                break
            cur = nxt
            level += 1
        layout = struct.pack("<BBBQ", 3, 2, 3, btree) + struct.pack("<III", c0, c1, es)
    # ---- dataset object header ----
    newer = message_versions == "new"
    if newer:
        msgs = _msg(0x01, struct.pack("<BBBBQQ", 2, 2, 0, 1, n, p))           # dataspace version 2, simple
    else:
        msgs = _msg(0x01, struct.pack("<BBB5xQQ", 1, 2, 0, n, p))
    msgs += _msg(0x03, _datatype_msg(A.dtype), flags=1)
    msgs += _msg(0x05, struct.pack("<BBBB", 2, 2, 0, 0))                    # fill value: version 2, undefined
    if filters and newer:
        fb = struct.pack("<BB", 2, len(filters))                           # version 2: no name for predefined filters, no padding
        for fid, cv in filters:
            fb += struct.pack("<HHH", fid, 1, len(cv)) + b"".join(struct.pack("<I", v) for v in cv)
        msgs += _msg(0x0B, fb)
    elif filters:
        fb = struct.pack("<BB6x", 1, len(filters))
        for fid, cv in filters:
            fb += struct.pack("<HHHH", fid, 0, 1, len(cv)) + b"".join(struct.pack("<I", v) for v in cv)
            if len(cv) % 2:
                fb += b"\0" * 4
        msgs += _msg(0x0B, fb)
    if message_versions == "v1layout":
        if chunks is None:
            layout = struct.pack("<BBB5xQ", 1, 2, 1, data_addr) + struct.pack("<II", n, p)
        else:
            layout = struct.pack("<BBB5xQ", 1, 3, 2, btree) + struct.pack("<III", c0, c1, es)
    msgs += _msg(0x08, layout)
    attr = _string_attr("MATLAB_class", matlab_class, version=3 if newer else 1)
    nmsg = 5 + (1 if filters else 0)
    if continuation:
        # the attribute lives in a continuation block, as in headers that grew after they were created (MATLAB adds
        # its attributes after the dataset exists)
        cont_addr = alloc(attr)
        msgs += _msg(0x10, struct.pack("<QQ", cont_addr, len(attr)))
        nmsg += 1
    else:
        msgs += attr
    ds_hdr = alloc(struct.pack("<BBHII4x", 1, 0, nmsg, 1, len(msgs)) + msgs)
    # ---- root group: local heap, symbol node, B-tree, object header ----
    heap_data = _pad8(b"\0" * 8 + name.encode() + b"\0")
    heap_data += b"\0" * (-len(heap_data) % 16 + 16)
    name_off = 8
    hd_addr = alloc(heap_data)
    heap_addr = alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), UNDEF, hd_addr))
    snod = b"SNOD" + struct.pack("<BBH", 1, 0, 1) + struct.pack("<QQII16x", name_off, ds_hdr, 0, 0)
    snod += b"\0" * (40 * 7)                                                # room for 2*K leaf entries (K = 4)
    snod_addr = alloc(snod)
    bt = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod_addr, name_off)
    bt += b"\0" * (16 * 31)                                                 # room for 2*K internal entries (K = 16)
    bt_addr = alloc(bt)
    root_msgs = _msg(0x11, struct.pack("<QQ", bt_addr, heap_addr))
    root_hdr = alloc(struct.pack("<BBHII4x", 1, 0, 1, 1, len(root_msgs)) + root_msgs)
    eof = len(out)
    sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
    sb += struct.pack("<QQQQ", base, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", bt_addr, heap_addr)
    out[:len(sb)] = sb
    import time
    text = f"MATLAB 7.3 MAT-file, Platform: sparsifiedkmeans_b200, Created on: {time.asctime()} HDF5 schema 1.00 ."
    header = text.encode("ascii").ljust(116) + b"\0" * 8 + struct.pack("<H", 0x0200) + b"IM"
    with open(path, "wb") as f:
        f.write(header.ljust(512, b"\0"))
        f.write(bytes(out))
