"""Minimal HDF5 container writer/reader for dolfin-style XDMF checkpoints (no HDF5 library exists in this image).

Writes exactly the on-disk structures that libhdf5 1.10 / dolfin 2019.1.0 produce for
``XDMFFile.write_checkpoint(f, name, t, XDMFFile.Encoding.HDF5, append)`` (reference call sites: fedm/file_io.py:594-600,
fedm_gd.py:271; consumer: tests/integrated_tests/testing_utils.py:21-24 reads ``h5[key][subkey]["vector"]``):

* superblock version 0, 8-byte offsets/lengths, group leaf K = 4, internal K = 16;
* "old-style" groups: version-1 object header with one symbol-table message -> version-1 B-tree (node type 0) ->
  symbol-table nodes (SNOD, <= 8 entries) + local heap (HEAP) holding the link names;
* data sets: version-1 object header (256-byte message block) with dataspace v1 (with max dims), datatype v1 (little-endian
  fixed point / IEEE double), fill-value v2, layout v3 class 1 (contiguous), modification time v1 and the scalar-like
  attribute ``partition`` (uint64[1]) dolfin attaches to every data set; raw data contiguous, little endian.

Appending never rewrites data: new objects go to the end of the file, only the (few hundred bytes of) link tables of the
groups that gained a member are written again and the two addresses that refer to them are patched.

The reader walks the same structures strictly (signatures, versions, sizes, name order) and works on files written by
libhdf5 itself: it is pinned on the reference's golden ``*.h5`` files (tests/test_hdf5_lite.py).
"""
import struct
import time
from pathlib import Path

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"
LEAF_K, INTERNAL_K = 4, 16
SNOD_CAP = 2 * LEAF_K
TREE_NODE_BYTES = 24 + 2 * INTERNAL_K * 16 + 8          # 544
SNOD_BYTES = 8 + SNOD_CAP * 40                            # 328
DSET_HEADER_BLOCK = 256


def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


# ---- message encodings -----------------------------------------------------------------------------------------------
def _datatype_body(dtype):
    dt = np.dtype(dtype)
    if dt.kind in "iu":
        bits = (0x08 if dt.kind == "i" else 0x00)         # byte order LE, padding 0, signed flag
        return struct.pack("<B3BI", 0x10, bits, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt == np.float64:
        # class 1 v1; bits: LE, mantissa normalisation 2 (implied msb), sign bit position 63
        return (struct.pack("<B3BI", 0x11, 0x20, 0x3F, 0x00, 8) +
                struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023))
    raise TypeError(f"hdf5_lite: unsupported dtype {dt}")


def _dataspace_body(shape):
    rank = len(shape)
    return struct.pack("<BBB5x", 1, rank, 1) + b"".join(struct.pack("<Q", int(d)) for d in shape) * 2


def _attribute_partition():
    name = _pad8(b"partition\0")
    dtb = _datatype_body(np.uint64)
    dsb = struct.pack("<BBB5x", 1, 1, 1) + struct.pack("<QQ", 1, 1)
    return struct.pack("<BBHHH", 1, 0, 10, len(dtb), len(dsb)) + name + _pad8(dtb) + _pad8(dsb) + struct.pack("<Q", 0)


def _message(mtype, body, flags=0):
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _dataset_header(arr, data_addr, mtime):
    msgs = [
        _message(0x01, _dataspace_body(arr.shape)),
        _message(0x03, _datatype_body(arr.dtype), flags=1),
        _message(0x05, struct.pack("<BBBBI", 2, 2, 2, 1, 0), flags=1),
        _message(0x08, struct.pack("<BBQQ", 3, 1, data_addr, arr.nbytes)),
        _message(0x12, struct.pack("<B3xI", 1, int(mtime) & 0xFFFFFFFF)),
        _message(0x0C, _attribute_partition()),
    ]
    used = sum(len(m) for m in msgs)
    nil = DSET_HEADER_BLOCK - used - 8
    assert nil >= 0
    msgs.append(struct.pack("<HHB3x", 0, nil, 0) + b"\0" * nil)
    return struct.pack("<BBHII4x", 1, 0, len(msgs), 1, DSET_HEADER_BLOCK) + b"".join(msgs)


def _group_header(btree, heap):
    return struct.pack("<BBHII4x", 1, 0, 1, 1, 24) + _message(0x11, struct.pack("<QQ", btree, heap))


# ---- writer ----------------------------------------------------------------------------------------------------------
class _Group:
    def __init__(self):
        self.header_addr = None          # object header (fixed once written)
        self.members = {}                # name -> _Group | int (data-set object-header address)
        self.dirty = True                # link tables must be (re)written
        self.btree = self.heap = None


class H5CheckpointFile:
    """One HDF5 file holding /<name>/<name>_<k>/{mesh/{topology,geometry},cell_dofs,x_cell_dofs,cells,vector}."""

    def __init__(self, path):
        self.path = Path(path)
        self.root = None
        self.eof = 0

    # -- low level
    def _alloc(self, fh, data):
        addr = self.eof
        fh.seek(addr)
        fh.write(data)
        self.eof = addr + len(data)
        pad = -self.eof % 8
        if pad:
            fh.write(b"\0" * pad)
            self.eof += pad
        return addr

    def _write_links(self, fh, grp):
        """(Re)writes heap + SNODs + B-tree of a group whose member list changed -- or whose child groups moved their
        tables (the parent's symbol-table entry caches them); children first.  Returns True if anything was written."""
        moved = False
        for m in grp.members.values():
            if isinstance(m, _Group):
                moved |= self._write_links(fh, m)
        if not (grp.dirty or moved or grp.header_addr is None):
            return False
        names = sorted(grp.members)                       # symbol-table order = strcmp order of the names
        heap_data = bytearray(8)
        off = {}
        for nm in names:
            off[nm] = len(heap_data)
            heap_data += _pad8(nm.encode() + b"\0")
        size = max(88, len(heap_data) + 16)
        size += -size % 8
        free = len(heap_data)
        heap_data += struct.pack("<QQ", 1, size - free) + b"\0" * (size - free - 16)
        heap_addr = self.eof
        heap = b"HEAP" + struct.pack("<B3xQQQ", 0, size, free, heap_addr + 32) + bytes(heap_data)
        self._alloc(fh, heap)
        # leaves
        entries = []
        for nm in names:
            m = grp.members[nm]
            if isinstance(m, _Group):
                entries.append(struct.pack("<QQII", off[nm], m.header_addr, 1, 0) + struct.pack("<QQ", m.btree, m.heap))
            else:
                entries.append(struct.pack("<QQII", off[nm], m, 0, 0) + b"\0" * 16)
        level = []                                        # (address, largest key) of the current tree level
        for i in range(0, max(len(names), 1), SNOD_CAP):
            chunk = entries[i:i + SNOD_CAP]
            snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(chunk)) + b"".join(chunk)
            snod += b"\0" * (SNOD_BYTES - len(snod))
            last = off[names[min(i + SNOD_CAP, len(names)) - 1]] if names else 0
            level.append((self._alloc(fh, snod), last))
        depth = 0
        while True:
            nodes = []
            for i in range(0, len(level), 2 * INTERNAL_K):
                chunk = level[i:i + 2 * INTERNAL_K]
                first_key = 0 if i == 0 else level[i - 1][1]
                body = struct.pack("<Q", first_key) + b"".join(struct.pack("<QQ", a, k) for a, k in chunk)
                nodes.append((chunk, body))
            addrs = [self.eof + j * TREE_NODE_BYTES for j in range(len(nodes))]
            new_level = []
            for j, (chunk, body) in enumerate(nodes):
                left = addrs[j - 1] if j > 0 else UNDEF
                right = addrs[j + 1] if j + 1 < len(nodes) else UNDEF
                node = b"TREE" + struct.pack("<BBHQQ", 0, depth, len(chunk), left, right) + body
                node += b"\0" * (TREE_NODE_BYTES - len(node))
                new_level.append((self._alloc(fh, node), chunk[-1][1]))
            level = new_level
            depth += 1
            if len(level) == 1:
                break
        grp.btree, grp.heap = level[0][0], heap_addr
        if grp.header_addr is None:
            grp.header_addr = self._alloc(fh, _group_header(grp.btree, grp.heap))
        else:                                             # patch the symbol-table message of the existing header
            fh.seek(grp.header_addr + 16 + 8)
            fh.write(struct.pack("<QQ", grp.btree, grp.heap))
        grp.dirty = False
        return True

    def _finish(self, fh):
        self._write_links(fh, self.root)
        root = self.root
        sb = (SIGNATURE + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", LEAF_K, INTERNAL_K, 0) +
              struct.pack("<QQQQ", 0, UNDEF, self.eof, UNDEF) +
              struct.pack("<QQII", 0, root.header_addr, 1, 0) + struct.pack("<QQ", root.btree, root.heap))
        assert len(sb) == 96
        fh.seek(0)
        fh.write(sb)

    def _load_existing(self):
        rd = H5Reader(self.path)
        self.eof = rd.eof

        def conv(addr):
            g = _Group()
            g.header_addr, g.dirty = addr, False
            g.btree, g.heap = rd.group_tables(addr)
            for nm, (child, is_group) in rd.group_members(addr).items():
                g.members[nm] = conv(child) if is_group else child
            return g

        self.root = conv(rd.root_header)

    # -- public
    def add_checkpoint(self, name, datasets, append):
        """datasets: dict path -> ndarray with paths relative to /<name>/<name>_<k>/ (e.g. 'mesh/topology', 'vector')."""
        fresh = not (append and self.path.exists() and self.path.stat().st_size > 0)
        if fresh:
            self.root, self.eof = _Group(), 96
        elif self.root is None:
            self._load_existing()
        self.path.parent.mkdir(parents=True, exist_ok=True)
        with open(self.path, "wb" if fresh else "r+b") as fh:
            if fresh:
                fh.write(b"\0" * 96)
                self.root.header_addr = self._alloc(fh, _group_header(0, 0))       # root header sits at 96 like libhdf5's
                self.root.dirty = True
            top = self.root.members.get(name)
            if top is None:
                top = self.root.members[name] = _Group()
                self.root.dirty = True
            k = len(top.members)
            cp = top.members[f"{name}_{k}"] = _Group()
            top.dirty = True
            now = time.time()
            for rel, arr in datasets.items():
                arr = np.ascontiguousarray(arr)
                if arr.dtype.byteorder == ">":
                    arr = arr.astype(arr.dtype.newbyteorder("<"))
                parts = rel.split("/")
                g = cp
                for p in parts[:-1]:
                    if p not in g.members:
                        g.members[p] = _Group()
                    g = g.members[p]
                hdr_addr = self.eof
                data_addr = hdr_addr + 16 + DSET_HEADER_BLOCK
                self._alloc(fh, _dataset_header(arr, data_addr, now))
                assert self.eof == data_addr
                self._alloc(fh, arr.tobytes())
                g.members[parts[-1]] = hdr_addr
            self._finish(fh)
        return k


# ---- reader ----------------------------------------------------------------------------------------------------------
class H5Reader:
    """Strict reader for the subset above (superblock v0, old-style groups, contiguous little-endian data sets)."""

    def __init__(self, path):
        self.buf = Path(path).read_bytes()
        b = self.buf
        if b[:8] != SIGNATURE:
            raise ValueError("not an HDF5 file")
        if b[8] != 0 or b[13] != 8 or b[14] != 8:
            raise ValueError("hdf5_lite: only superblock version 0 with 8-byte offsets is supported")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", b, 16)
        base, _, self.eof, _ = struct.unpack_from("<QQQQ", b, 24)
        if base != 0 or self.eof != len(b):
            raise ValueError("hdf5_lite: base address / end-of-file mismatch")
        _, self.root_header, cache, _ = struct.unpack_from("<QQII", b, 56)

    def _u(self, fmt, off):
        return struct.unpack_from("<" + fmt, self.buf, off)

    def messages(self, addr):
        ver, _, nmsg, _, hsize = self._u("BBHII", addr)
        if ver != 1:
            raise ValueError(f"hdf5_lite: object header version {ver} at {addr}")
        blocks, out = [(addr + 16, addr + 16 + hsize)], []
        while blocks and len(out) < nmsg:
            o, end = blocks.pop(0)
            while o < end and len(out) < nmsg:
                mtype, msize, mflags = self._u("HHB", o)
                body = self.buf[o + 8:o + 8 + msize]
                if mtype == 0x10:
                    ca, cl = struct.unpack("<QQ", body[:16])
                    blocks.append((ca, ca + cl))
                out.append((mtype, mflags, body))
                o += 8 + msize
        return out

    def group_tables(self, addr):
        for mtype, _, body in self.messages(addr):
            if mtype == 0x11:
                return struct.unpack("<QQ", body[:16])
        raise ValueError(f"object at {addr} is not an old-style group")

    def is_group(self, addr):
        return any(m[0] == 0x11 for m in self.messages(addr))

    def _heap_name(self, heap, off):
        if self.buf[heap:heap + 4] != b"HEAP":
            raise ValueError("bad local heap signature")
        size, _, daddr = self._u("QQQ", heap + 8)
        if off >= size:
            raise ValueError("link name offset outside the heap")
        end = self.buf.index(b"\0", daddr + off)
        return self.buf[daddr + off:end].decode()

    def group_members(self, addr):
        btree, heap = self.group_tables(addr)
        out = {}

        def node(a):
            if self.buf[a:a + 4] != b"TREE":
                raise ValueError("bad B-tree signature")
            ntype, level, nent = self._u("BBH", a + 4)
            if ntype != 0 or nent > 2 * self.internal_k:
                raise ValueError("bad group B-tree node")
            for i in range(nent):
                child, = self._u("Q", a + 24 + 16 * i + 8)
                if level > 0:
                    node(child)
                    continue
                if self.buf[child:child + 4] != b"SNOD":
                    raise ValueError("bad symbol-table node signature")
                _, _, nsym = self._u("BBH", child + 4)
                if nsym > 2 * self.leaf_k:
                    raise ValueError("symbol-table node overflow")
                for j in range(nsym):
                    lno, oh, cache, _ = self._u("QQII", child + 8 + 40 * j)
                    out[self._heap_name(heap, lno)] = (oh, cache == 1 or self.is_group(oh))

        node(btree)
        names = list(out)
        if names != sorted(names):
            raise ValueError("symbol table not in name order")
        return out

    def dataset(self, addr):
        shape = dtype = layout = None
        attrs = {}
        for mtype, _, body in self.messages(addr):
            if mtype == 0x01:
                ver, rank, flags = struct.unpack_from("<BBB", body, 0)
                shape = struct.unpack_from(f"<{rank}Q", body, 8 if ver == 1 else 4)
            elif mtype == 0x03:
                dtype = self._dtype(body)
            elif mtype == 0x08:
                ver, cls = struct.unpack_from("<BB", body, 0)
                if ver != 3 or cls != 1:
                    raise ValueError("hdf5_lite: only contiguous version-3 layouts are supported")
                layout = struct.unpack_from("<QQ", body, 2)
            elif mtype == 0x0C:
                ver, _, nsz, tsz, ssz = struct.unpack_from("<BBHHH", body, 0)
                o = 8
                nm = body[o:o + nsz].split(b"\0")[0].decode()
                o += nsz + (-nsz % 8)
                adt = self._dtype(body[o:o + tsz])
                o += tsz + (-tsz % 8)
                rank = body[o + 1]
                ashape = struct.unpack_from(f"<{rank}Q", body, o + 8)
                o += ssz + (-ssz % 8)
                attrs[nm] = np.frombuffer(body, dtype=adt, count=int(np.prod(ashape)), offset=o).reshape(ashape).copy()
        if shape is None or dtype is None or layout is None:
            raise ValueError(f"object at {addr} is not a contiguous data set")
        daddr, nbytes = layout
        count = int(np.prod(shape))
        if nbytes != count * dtype.itemsize or daddr + nbytes > len(self.buf):
            raise ValueError("data set size mismatch")
        return np.frombuffer(self.buf, dtype=dtype, count=count, offset=daddr).reshape(shape).copy(), attrs

    @staticmethod
    def _dtype(body):
        cv, b0, _, _, size = struct.unpack_from("<B3BI", body, 0)
        cls = cv & 0x0F
        if b0 & 1:
            raise ValueError("hdf5_lite: big-endian data")
        if cls == 0:
            return np.dtype(("<i" if b0 & 0x08 else "<u") + str(size))
        if cls == 1 and size == 8:
            return np.dtype("<f8")
        if cls == 1 and size == 4:
            return np.dtype("<f4")
        raise ValueError(f"hdf5_lite: datatype class {cls} size {size}")

    def tree(self, addr=None):
        """Nested dict: groups -> dicts, data sets -> ndarrays."""
        addr = self.root_header if addr is None else addr
        out = {}
        for nm, (child, is_group) in self.group_members(addr).items():
            out[nm] = self.tree(child) if is_group else self.dataset(child)[0]
        return out


def read_h5(path, key):
    """Same result as the reference's tests/integrated_tests/testing_utils.py:21-24 (list of the 'vector' arrays)."""
    t = H5Reader(path).tree()
    return [np.asarray(t[key][sub]["vector"]) for sub in t[key]]
