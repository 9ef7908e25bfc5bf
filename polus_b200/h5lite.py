"""Minimal HDF5 reader / writer for the reference's weight files (`<name>.h5`, polus/models.py:42-48,127-133):
a root group holding N plain numeric datasets `weight0 .. weightN-1`, written with
`h5py.File(path, "w").create_dataset("weight%d" % i, data=w)`.  h5py is not installable in this image, so the two
directions of that exchange are implemented here from the HDF5 File Format Specification (version 0 superblock, the
layout h5py's default `libver="earliest"` produces and every HDF5 library reads):

    superblock v0 -> root symbol-table entry -> object header v1 -> Symbol Table message
        -> v1 B-tree (group nodes) + local heap (link names) -> symbol nodes (SNOD) -> one object header per dataset
        -> Dataspace (simple, v1) + Datatype (fixed-point / IEEE float, little endian) + Data Layout v3 (contiguous)

Reader: also accepts a user block (non-zero base address), compact layout, header continuation blocks, big-endian
scalars types and dataspace v2 -- what the HDF5 library may emit for such files; chunked / filtered datasets, groups
below the root and non-numeric types are refused with a clear error (the reference never writes them).
Pinned against a file written by the real HDF5 library (tests/test_h5lite_cpu.py)."""
import struct

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF
_LEAF_K = 64    # symbols per SNOD = 2 * leaf K   (superblock field; the library honours what the file declares)
_INT_K = 16     # children per group B-tree node = 2 * internal K


class H5Error(ValueError):
    pass


# ======================================================================================================== reader
class _File:
    def __init__(self, buf):
        self.b = buf
        start = 0
        while True:   # the superblock sits at 0, 512, 1024, 2048 ... (user block in front)
            if buf[start:start + 8] == _SIG:
                break
            start = 512 if start == 0 else start * 2
            if start + 8 > len(buf):
                raise H5Error("not an HDF5 file (no superblock signature)")
        self.sb = start
        ver = buf[start + 8]
        if ver not in (0, 1):
            raise H5Error(f"HDF5 superblock version {ver} is not supported (h5py's default writes version 0)")
        self.osz, self.lsz = buf[start + 13], buf[start + 14]
        if self.osz != 8 or self.lsz != 8:
            raise H5Error("only 8-byte offsets / lengths are supported")
        p = start + 16
        self.leaf_k, self.int_k = struct.unpack_from("<HH", buf, p)
        p += 4 + 4                      # K values, file consistency flags
        if ver == 1:
            p += 4                      # indexed-storage K + reserved
        self.base = struct.unpack_from("<Q", buf, p)[0]
        p += 8 * 4                      # base, free-space info, end of file, driver info
        self.root_ste = p

    def u(self, off, n=8):
        return int.from_bytes(self.b[off:off + n], "little")

    def at(self, addr):
        return self.base + addr

    # ---- object header v1: list of (type, flags, payload offset, payload size)
    def messages(self, addr):
        o = self.at(addr)
        if self.b[o] != 1:
            raise H5Error(f"object header version {self.b[o]} is not supported (need version 1)")
        nmsg = self.u(o + 2, 2)
        size = self.u(o + 8, 4)
        blocks = [(o + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = self.u(p, 2), self.u(p + 2, 2), self.b[p + 4]
                body = p + 8
                if mtype == 0x0010:     # continuation: another block of messages elsewhere
                    blocks.append((self.at(self.u(body)), self.u(body + 8)))
                out.append((mtype, flags, body, msize))
                p = body + msize
        return out

    def group_entries(self, btree_addr, heap_addr):
        h = self.at(heap_addr)
        if self.b[h:h + 4] != b"HEAP":
            raise H5Error("bad local heap signature")
        heap_data = self.at(self.u(h + 24))

        def name_at(off):
            s = heap_data + off
            e = self.b.index(b"\x00", s)
            return self.b[s:e].decode("utf-8")
        out = []

        def walk(addr):
            t = self.at(addr)
            if self.b[t:t + 4] == b"SNOD":
                n = self.u(t + 6, 2)
                for i in range(n):
                    e = t + 8 + 40 * i
                    out.append((name_at(self.u(e)), self.u(e + 8)))
                return
            if self.b[t:t + 4] != b"TREE" or self.b[t + 4] != 0:
                raise H5Error("bad group B-tree node")
            used = self.u(t + 6, 2)
            for i in range(used):
                walk(self.u(t + 24 + 8 + 16 * i))   # key0, child0, key1, child1 ...
        walk(btree_addr)
        return out


_FLOAT = {(4, "<"): "<f4", (8, "<"): "<f8", (2, "<"): "<f2", (4, ">"): ">f4", (8, ">"): ">f8", (2, ">"): ">f2"}


def _dtype_of(f, body):
    cls_ver = f.b[body]
    cls, bits0 = cls_ver & 0x0F, f.b[body + 1]
    size = f.u(body + 4, 4)
    order = ">" if (bits0 & 1) else "<"
    if cls == 0:    # fixed point
        signed = bool(bits0 & 0x08)
        return np.dtype(f"{order}{'i' if signed else 'u'}{size}")
    if cls == 1:    # floating point (IEEE layouts only: the sizes decide)
        key = (size, order)
        if key not in _FLOAT:
            raise H5Error(f"unsupported floating-point size {size}")
        return np.dtype(_FLOAT[key])
    raise H5Error(f"datatype class {cls} is not supported (numeric datasets only)")


def _read_dataset(f, addr):
    shape = dtype = None
    data = None
    layout = None
    for mtype, flags, body, msize in f.messages(addr):
        if mtype == 0x0001:     # dataspace
            ver, rank, fl = f.b[body], f.b[body + 1], f.b[body + 2]
            p = body + (8 if ver == 1 else 4)
            if ver == 2 and f.b[body + 3] == 2:
                raise H5Error("null dataspace")
            shape = tuple(f.u(p + 8 * i) for i in range(rank))
        elif mtype == 0x0003:
            dtype = _dtype_of(f, body)
        elif mtype == 0x0008:   # data layout
            ver = f.b[body]
            if ver == 3:
                cls = f.b[body + 1]
                if cls == 1:
                    layout = ("contiguous", f.u(body + 2), f.u(body + 10))
                elif cls == 0:
                    n = f.u(body + 2, 2)
                    layout = ("compact", body + 4, n)
                else:
                    raise H5Error("chunked datasets are not supported (the reference writes contiguous arrays)")
            elif ver in (1, 2):
                rank, cls = f.b[body + 1], f.b[body + 2]
                if cls != 1:
                    raise H5Error("only contiguous layout is supported for layout message versions 1 / 2")
                layout = ("contiguous", f.u(body + 8), None)
            else:
                raise H5Error(f"data layout message version {ver}")
        elif mtype == 0x000B:
            raise H5Error("filtered (compressed) datasets are not supported")
    if shape is None or dtype is None or layout is None:
        raise H5Error("object is not a dataset (a sub-group?)")
    count = int(np.prod(shape)) if shape else 1
    nbytes = count * dtype.itemsize
    if layout[0] == "contiguous":
        if layout[1] == _UNDEF:
            data = np.zeros(shape, dtype)   # never written: fill value 0
        else:
            o = f.at(layout[1])
            data = np.frombuffer(f.b, dtype=dtype, count=count, offset=o).reshape(shape)
    else:
        data = np.frombuffer(f.b, dtype=dtype, count=count, offset=layout[1]).reshape(shape)
    del nbytes
    return np.array(data, dtype=dtype.newbyteorder("="))


def read_h5(path):
    """{dataset name: numpy array} of the root group, in the file's B-tree (= lexicographic) order."""
    with open(path, "rb") as fh:
        f = _File(fh.read())
    ste = f.root_ste
    cache = f.u(ste + 16, 4)
    if cache == 1:
        btree, heap = f.u(ste + 24), f.u(ste + 32)
    else:   # not cached: read the Symbol Table message of the root object header
        btree = heap = None
        for mtype, flags, body, msize in f.messages(f.u(ste + 8)):
            if mtype == 0x0011:
                btree, heap = f.u(body), f.u(body + 8)
        if btree is None:
            raise H5Error("root group without a symbol table (new-style groups are not supported)")
    return {name: _read_dataset(f, addr) for name, addr in f.group_entries(btree, heap)}


def read_weights(path):
    """The reference's loader (polus/models.py:44-47): [f['weight%d' % i][:] for i in range(len(f.keys()))]."""
    d = read_h5(path)
    return [d[f"weight{i}"] for i in range(len(d))]


# ======================================================================================================== writer
def _pad8(n):
    return (n + 7) & ~7


def _msg(mtype, payload, flags=0):
    body = payload + b"\x00" * (_pad8(len(payload)) - len(payload))
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _object_header(messages):
    body = b"".join(messages)
    # version 1, reserved, number of messages, object reference count, header size, 4 bytes of padding to an 8-byte boundary
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


def _datatype_msg(dt):
    dt = np.dtype(dt)
    if dt.kind == "f":
        # class 1 version 1; bits: little endian, lo/hi/internal pad 0, mantissa normalisation 2 (implied msb), sign position
        spec = {2: (15, 10, 5, 0, 10, 15), 4: (31, 23, 8, 0, 23, 127), 8: (63, 52, 11, 0, 52, 1023)}[dt.itemsize]
        sign, epos, esize, mpos, msize, bias = spec
        head = struct.pack("<BBBBI", 0x11, 0x20, sign, 0, dt.itemsize)
        props = struct.pack("<HHBBBBI", 0, dt.itemsize * 8, epos, esize, mpos, msize, bias)
        return _msg(0x0003, head + props, flags=1)
    if dt.kind in "iu":
        head = struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize)
        props = struct.pack("<HH", 0, dt.itemsize * 8)
        return _msg(0x0003, head + props, flags=1)
    raise H5Error(f"cannot store dtype {dt} (float / integer arrays only)")


def _dataspace_msg(shape):
    if len(shape) == 0:
        return _msg(0x0001, struct.pack("<BBB5x", 1, 0, 0))
    # version 1, rank, flags = 1 (maximum dimensions present: h5py writes them equal to the current ones)
    return _msg(0x0001, struct.pack("<BBB5x", 1, len(shape), 1) + b"".join(struct.pack("<Q", int(s)) for s in shape) * 2)


def write_h5(path, named_arrays):
    """Write [(name, array)...] as contiguous datasets of the root group (version-0 superblock, symbol-table group)."""
    items = [(str(n), np.require(np.asarray(a), requirements="C")) for n, a in named_arrays]
    for n, a in items:
        if a.dtype.byteorder == ">":
            raise H5Error("big-endian arrays are not written")
    names = sorted(n for n, _ in items)              # group B-trees are ordered by link name (strcmp)
    if len(set(names)) != len(names):
        raise H5Error("duplicate dataset names")
    per_snod = 2 * _LEAF_K
    n_snod = max(1, (len(names) + per_snod - 1) // per_snod)
    if n_snod > 2 * _INT_K:
        raise H5Error(f"too many datasets for a single-level group B-tree ({len(names)} > {2 * _INT_K * per_snod})")

    # ---- local heap data: offset 0 = "" (the root's own name), then every link name, 8-byte aligned
    heap_off, heap = {}, bytearray(b"\x00" * 8)
    for n in names:
        heap_off[n] = len(heap)
        raw = n.encode("utf-8") + b"\x00"
        heap += raw + b"\x00" * (_pad8(len(raw)) - len(raw))
    free_off = len(heap)
    heap += struct.pack("<QQ", 1, 16)                # one free block at the end: next = 1 (none), size 16 (itself)
    heap_size = len(heap)

    # ---- layout of the file (addresses relative to base address 0)
    SB = 8 + 8 + 4 + 4 + 8 * 4 + 40                  # superblock v0 incl. the root symbol-table entry = 96 bytes
    addr = SB
    root_ohdr_addr = addr
    root_ohdr = _object_header([_msg(0x0011, struct.pack("<QQ", 0, 0))])   # placeholder, same size as the final one
    addr += len(root_ohdr)
    btree_addr = addr
    btree_size = 24 + (2 * _INT_K + 1) * 8 + 2 * _INT_K * 8
    addr += btree_size
    heap_hdr_addr = addr
    addr += 32
    heap_data_addr = addr
    addr += heap_size
    snod_size = 8 + per_snod * 40
    snod_addr = [addr + i * snod_size for i in range(n_snod)]
    addr += n_snod * snod_size
    by_name = dict(items)
    ohdr_addr, ohdr_bytes, data_addr = {}, {}, {}
    for n in names:                                   # object headers first (placeholders fix their sizes) ...
        a = by_name[n]
        hdr = _object_header([_dataspace_msg(a.shape), _datatype_msg(a.dtype), _msg(0x0005, struct.pack("<BBBB", 2, 2, 2, 0)),
                              _msg(0x0008, struct.pack("<BBQQ", 3, 1, 0, a.nbytes)),
                              _msg(0x0000, b"\x00" * 8)])           # a NIL message: room for a later attribute, like the library leaves
        ohdr_addr[n] = addr
        ohdr_bytes[n] = len(hdr)
        addr += len(hdr)
    for n in names:                                   # ... then the raw data, 8-byte aligned
        addr = _pad8(addr)
        data_addr[n] = addr if by_name[n].nbytes else _UNDEF
        addr += by_name[n].nbytes
    eof = addr

    out = bytearray(eof)
    # superblock
    sb = _SIG + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", _LEAF_K, _INT_K, 0)
    sb += struct.pack("<QQQQ", 0, _UNDEF, eof, _UNDEF)
    sb += struct.pack("<QQII", 0, root_ohdr_addr, 1, 0) + struct.pack("<QQ", btree_addr, heap_hdr_addr)   # root entry, cached
    assert len(sb) == SB
    out[0:SB] = sb
    root_ohdr = _object_header([_msg(0x0011, struct.pack("<QQ", btree_addr, heap_hdr_addr))])
    out[root_ohdr_addr:root_ohdr_addr + len(root_ohdr)] = root_ohdr
    # B-tree: one level-0 node whose children are the symbol nodes; key i+1 = heap offset of the LAST name in child i
    groups = [names[i * per_snod:(i + 1) * per_snod] for i in range(n_snod)] if names else [[]]
    used = n_snod if names else 0
    bt = bytearray(btree_size)
    bt[0:24] = b"TREE" + struct.pack("<BBH", 0, 0, used) + struct.pack("<QQ", _UNDEF, _UNDEF)
    p = 24
    struct.pack_into("<Q", bt, p, 0)
    p += 8
    for i in range(used):
        struct.pack_into("<Q", bt, p, snod_addr[i])
        struct.pack_into("<Q", bt, p + 8, heap_off[groups[i][-1]])
        p += 16
    out[btree_addr:btree_addr + btree_size] = bt
    # local heap
    out[heap_hdr_addr:heap_hdr_addr + 32] = b"HEAP" + struct.pack("<B3x", 0) + struct.pack("<QQQ", heap_size, free_off, heap_data_addr)
    out[heap_data_addr:heap_data_addr + heap_size] = heap
    # symbol nodes
    for i in range(n_snod):
        sn = bytearray(snod_size)
        g = groups[i] if names else []
        sn[0:8] = b"SNOD" + struct.pack("<BBH", 1, 0, len(g))
        for k, n in enumerate(g):
            struct.pack_into("<QQII16x", sn, 8 + 40 * k, heap_off[n], ohdr_addr[n], 0, 0)
        out[snod_addr[i]:snod_addr[i] + snod_size] = sn
    # datasets
    for n in names:
        a = by_name[n]
        hdr = _object_header([_dataspace_msg(a.shape), _datatype_msg(a.dtype), _msg(0x0005, struct.pack("<BBBB", 2, 2, 2, 0)),
                              _msg(0x0008, struct.pack("<BBQQ", 3, 1, data_addr[n], a.nbytes)), _msg(0x0000, b"\x00" * 8)])
        assert len(hdr) == ohdr_bytes[n]
        out[ohdr_addr[n]:ohdr_addr[n] + len(hdr)] = hdr
        if a.nbytes:
            out[data_addr[n]:data_addr[n] + a.nbytes] = a.tobytes()
    with open(path, "wb") as fh:
        fh.write(out)


def write_weights(path, weights):
    """The reference's writer (polus/models.py:130-133): dataset 'weight%d' % i for every array of get_weights()."""
    write_h5(path, [(f"weight{i}", np.asarray(w)) for i, w in enumerate(weights)])
