"""A small native HDF5 reader / writer for FAST5 files.

h5py and libhdf5 are absent from this image, and the re-squiggle path only needs a narrow slice of
HDF5: open a single-read FAST5 (resquiggle.py:328-341, 809-846), read int16 / compound datasets
and numeric / string attributes, add `/Analyses/RawGenomeCorrected_000/...` with gzip datasets and
attributes (resquiggle.py:80-135, 942-988), save.  This module implements that slice from the
published HDF5 File Format Specification (version 3.0), in pure Python + numpy + zlib:

reader
    superblock 0/1 (symbol-table root) and 2/3 (object-header root), user blocks; object headers
    v1 and v2 with continuation blocks; old-style groups (v1 B-tree + local heap + symbol nodes) and
    new-style compact groups (link messages); datasets with compact / contiguous / chunked (v1
    B-tree, and single-chunk v4) layout; deflate, shuffle and fletcher32 filters; fixed-point,
    floating-point, fixed and variable-length string, compound, enum and array datatypes;
    attribute messages v1-v3; global heap collections.
    Not read (raises Unsupported): dense groups / dense attributes (fractal heaps), shared
    (committed) datatypes, references, virtual / external storage, v4 chunk indices other than the
    single chunk.

writer
    always the most conservative on-disk forms, the ones libhdf5 itself writes by default
    ("earliest" format): superblock 0, v1 object headers, symbol-table groups, v1 attribute and
    datatype messages, contiguous datasets or a single gzip chunk indexed by a v1 B-tree, one global
    heap collection for variable-length strings.

Modifying a file ('r+') re-serialises the whole object tree into a temporary file and renames it
over the original on close: FAST5 files hold one read (about a megabyte), so this costs little and
the file is never left half-written.  A file that contains anything the reader cannot represent
faithfully is refused when opened for writing instead of being rewritten lossily.

The object model (groups, datasets, `.attrs`, `create_group`, `create_dataset(..., compression=
'gzip')`, `in`, `del`, `values()`) is the in-memory backend's (fast5_io.MemGroup), i.e. h5py's.
"""
import os
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b'\x89HDF\r\n\x1a\n'


class Unsupported(Exception):
    """the file uses an HDF5 feature outside the supported slice"""


class VlenStr(object):
    """marker datatype: variable-length string (element = length, global heap address, index)"""

    def __init__(self, utf8=True):
        self.utf8 = utf8
        self.itemsize = 16


# =====================================================================================  reader
class _Buf(object):
    def __init__(self, data, base=0, osz=8, lsz=8):
        self.d = data
        self.base = base
        self.osz = osz
        self.lsz = lsz

    def u(self, off, n):
        return int.from_bytes(self.d[off:off + n], 'little')

    def offset(self, off):
        v = self.u(off, self.osz)
        return None if v == (1 << (8 * self.osz)) - 1 else v + self.base

    def length(self, off):
        return self.u(off, self.lsz)


def _pad8(n):
    return (n + 7) & ~7


def _cstr(data, off):
    end = data.index(b'\x00', off)
    return bytes(data[off:end])


class Reader(object):
    def __init__(self, data):
        self.data = data
        pos = 0
        while True:
            if data[pos:pos + 8] == SIGNATURE:
                break
            pos = 512 if pos == 0 else pos * 2
            if pos + 8 > len(data):
                raise IOError('not an HDF5 file (no signature)')
        self.sb_off = pos
        ver = data[pos + 8]
        self.sb_version = ver
        if ver in (0, 1):
            osz, lsz = data[pos + 13], data[pos + 14]
            self.b = _Buf(data, 0, osz, lsz)
            p = pos + 24 + (4 if ver == 1 else 0)
            self.b.base = self.b.u(p, osz)  # addresses are relative to the base address (user block size)
            p += 4 * osz
            self.root_addr = self.b.offset(p + osz)  # root group symbol table entry: object header address
            self.chunk_k = self.b.u(pos + 24, 2) if ver == 1 else 32
        elif ver in (2, 3):
            osz, lsz = data[pos + 9], data[pos + 10]
            self.b = _Buf(data, 0, osz, lsz)
            self.b.base = self.b.u(pos + 12, osz)
            self.root_addr = self.b.offset(pos + 12 + 3 * osz)
            self.chunk_k = 32
        else:
            raise Unsupported('superblock version %d' % ver)
        self._gcol = {}
        self.lossy = []  # reasons why re-writing this file would not be faithful

    # ------------------------------------------------------------------ object headers
    def messages(self, addr):
        """[(type, flags, bytes)] of the object header at addr (continuations followed)"""
        d, b = self.data, self.b
        out = []
        if d[addr:addr + 4] == b'OHDR':
            if d[addr + 4] != 2:
                raise Unsupported('object header version %d' % d[addr + 4])
            flags = d[addr + 5]
            p = addr + 6
            if flags & 0x20:
                p += 16
            if flags & 0x10:
                p += 4
            nsz = 1 << (flags & 3)
            size0 = b.u(p, nsz)
            p += nsz
            blocks = [(p, size0)]
            track_order = bool(flags & 0x04)
            while blocks:
                p, size = blocks.pop(0)
                end = p + size
                while p + 4 <= end:
                    mtype, msize, mflags = d[p], b.u(p + 1, 2), d[p + 3]
                    p += 4 + (2 if track_order else 0)
                    body = bytes(d[p:p + msize])
                    p += msize
                    if mtype == 0x10:
                        caddr, clen = b.offset_from(body, 0), b.u_from(body, b.osz, b.lsz)
                        if d[caddr:caddr + 4] != b'OCHK':
                            raise IOError('bad object header continuation')
                        blocks.append((caddr + 4, clen - 8))
                    elif mtype != 0:
                        out.append((mtype, mflags, body))
            return out
        if d[addr] != 1:
            raise Unsupported('object header version %d' % d[addr])
        nmsg = b.u(addr + 2, 2)
        size0 = b.u(addr + 8, 4)
        blocks = [(addr + 16, size0)]
        while blocks and nmsg > 0:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and nmsg > 0:
                mtype, msize, mflags = b.u(p, 2), b.u(p + 2, 2), d[p + 4]
                body = bytes(d[p + 8:p + 8 + msize])
                p += 8 + msize
                nmsg -= 1
                if mtype == 0x10:
                    blocks.append((b.offset_from(body, 0), b.u_from(body, b.osz, b.lsz)))
                elif mtype != 0:
                    out.append((mtype, mflags, body))
        return out

    # ------------------------------------------------------------------ groups
    def _heap_data(self, addr):
        d, b = self.data, self.b
        if d[addr:addr + 4] != b'HEAP':
            raise IOError('bad local heap')
        return b.offset(addr + 8 + 2 * b.lsz)

    def _btree_group(self, addr, heap, out):
        d, b = self.data, self.b
        if d[addr:addr + 4] != b'TREE' or d[addr + 4] != 0:
            raise IOError('bad group B-tree node')
        level, n = d[addr + 5], b.u(addr + 6, 2)
        p = addr + 8 + 2 * b.osz
        for i in range(n):
            child = b.offset(p + b.lsz)
            p += b.lsz + b.osz
            if level > 0:
                self._btree_group(child, heap, out)
            else:
                if d[child:child + 4] != b'SNOD':
                    raise IOError('bad symbol table node')
                ns = b.u(child + 6, 2)
                q = child + 8
                for _ in range(ns):
                    name = _cstr(d, heap + b.u(q, b.osz)).decode('utf-8')
                    cache = b.u(q + 2 * b.osz, 4)
                    if cache == 2:
                        raise Unsupported('soft link %r' % name)
                    out.append((name, b.offset(q + b.osz)))
                    q += 2 * b.osz + 24

    def links(self, msgs):
        """[(name, object header address)] of a group given its header messages, or None"""
        b = self.b
        out, is_group = [], False
        for mtype, mflags, body in msgs:
            if mtype == 0x11:  # symbol table
                is_group = True
                btree, heap = b.offset_from(body, 0), b.offset_from(body, b.osz)
                self._btree_group(btree, self._heap_data(heap), out)
            elif mtype == 0x02:  # link info
                is_group = True
                flags = body[1]
                p = 2 + (8 if flags & 1 else 0)
                if b.offset_from(body, p) is not None:
                    raise Unsupported('dense link storage (fractal heap)')
            elif mtype == 0x06:  # link
                is_group = True
                flags = body[1]
                p = 2
                ltype = 0
                if flags & 0x08:
                    ltype = body[p]
                    p += 1
                if flags & 0x04:
                    p += 8
                if flags & 0x10:
                    p += 1
                nsz = 1 << (flags & 3)
                nlen = int.from_bytes(body[p:p + nsz], 'little')
                p += nsz
                name = body[p:p + nlen].decode('utf-8')
                p += nlen
                if ltype != 0:
                    raise Unsupported('soft / external link %r' % name)
                out.append((name, b.offset_from(body, p)))
        return out if is_group else None

    # ------------------------------------------------------------------ datatypes
    def datatype(self, body, p=0):
        """-> (numpy dtype | VlenStr, bytes consumed)"""
        cls, ver = body[p] & 0x0f, body[p] >> 4
        bits = body[p + 1] | (body[p + 2] << 8) | (body[p + 3] << 16)
        size = int.from_bytes(body[p + 4:p + 8], 'little')
        q = p + 8
        if cls == 0:  # fixed point
            order = '>' if bits & 1 else '<'
            kind = 'i' if bits & 8 else 'u'
            return np.dtype('%s%s%d' % (order, kind, size)), q + 4 - p
        if cls == 1:  # floating point
            order = '>' if bits & 1 else '<'
            if size not in (2, 4, 8):
                raise Unsupported('float of %d bytes' % size)
            return np.dtype('%sf%d' % (order, size)), q + 12 - p
        if cls == 3:  # fixed-length string
            return np.dtype('S%d' % size), q - p
        if cls == 6:  # compound
            n = bits & 0xffff
            names, formats, offsets = [], [], []
            for _ in range(n):
                name = _cstr(body, q)
                q += _pad8(len(name) + 1) if ver < 3 else len(name) + 1
                if ver < 3:
                    off = int.from_bytes(body[q:q + 4], 'little')
                    q += 4
                else:
                    nb = max(1, (max(size - 1, 1).bit_length() + 7) // 8)
                    off = int.from_bytes(body[q:q + nb], 'little')
                    q += nb
                dims = ()
                if ver == 1:
                    rank = body[q]
                    dims = tuple(int.from_bytes(body[q + 12 + 4 * k:q + 16 + 4 * k], 'little') for k in range(rank))
                    q += 28
                mt, used = self.datatype(body, q)
                q += used
                if isinstance(mt, VlenStr):
                    raise Unsupported('variable-length string inside a compound')
                names.append(name.decode('utf-8'))
                formats.append((mt, dims) if dims else mt)
                offsets.append(off)
            return np.dtype({'names': names, 'formats': formats, 'offsets': offsets, 'itemsize': size}), q - p
        if cls == 8:  # enum -> its base integer type
            n = bits & 0xffff
            base, used = self.datatype(body, q)
            q += used
            for _ in range(n):
                name = _cstr(body, q)
                q += _pad8(len(name) + 1) if ver < 3 else len(name) + 1
            q += n * base.itemsize
            self.lossy.append('enum datatype')
            return base, q - p
        if cls == 9:  # variable length
            base, used = self.datatype(body, q)
            if (bits & 0x0f) != 1:
                raise Unsupported('variable-length sequence')
            return VlenStr(utf8=((bits >> 8) & 0x0f) == 1), q + used - p
        if cls == 10:  # array
            rank = body[q]
            q += 4 if ver < 3 else 1
            dims = tuple(int.from_bytes(body[q + 4 * k:q + 4 * k + 4], 'little') for k in range(rank))
            q += 4 * rank * (2 if ver < 3 else 1)
            base, used = self.datatype(body, q)
            return np.dtype((base, dims)), q + used - p
        raise Unsupported('datatype class %d' % cls)

    def dataspace(self, body):
        """-> shape tuple, or None for a null dataspace"""
        ver, rank, flags = body[0], body[1], body[2]
        if ver == 1:
            p = 8
        elif ver == 2:
            if body[3] == 2:
                return None
            p = 4
        else:
            raise Unsupported('dataspace version %d' % ver)
        L = self.b.lsz
        return tuple(int.from_bytes(body[p + L * k:p + L * k + L], 'little') for k in range(rank))

    # ------------------------------------------------------------------ global heap (variable-length strings)
    def gcol_object(self, addr, index):
        if addr not in self._gcol:
            d, b = self.data, self.b
            if d[addr:addr + 4] != b'GCOL':
                raise IOError('bad global heap collection')
            size = b.length(addr + 8)
            objs = {}
            p, end = addr + 8 + b.lsz, addr + size
            while p + 8 + b.lsz <= end:
                idx = b.u(p, 2)
                if idx == 0:
                    break
                osize = b.length(p + 8)
                objs[idx] = bytes(d[p + 8 + b.lsz:p + 8 + b.lsz + osize])
                p += 8 + b.lsz + _pad8(osize)
            self._gcol[addr] = objs
        return self._gcol[addr][index]

    def decode(self, raw, dtype, shape):
        """raw element bytes -> numpy array / list of str of `shape`"""
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        if isinstance(dtype, VlenStr):
            vals = []
            step = 8 + self.b.osz
            for k in range(n):
                e = raw[k * step:(k + 1) * step]
                ln = int.from_bytes(e[0:4], 'little')
                a = int.from_bytes(e[4:4 + self.b.osz], 'little')
                idx = int.from_bytes(e[4 + self.b.osz:8 + self.b.osz], 'little')
                s = b'' if (ln == 0 or a == 0) else self.gcol_object(a + self.b.base, idx)[:ln]
                vals.append(s.decode('utf-8', 'surrogateescape'))
            if not shape:
                return vals[0]
            arr = np.empty(n, dtype=object)
            arr[:] = vals
            return arr.reshape(shape)
        arr = np.frombuffer(raw, dtype=dtype, count=n)
        return arr.reshape(shape) if shape else arr[0]

    # ------------------------------------------------------------------ attributes
    def attributes(self, msgs):
        out = []
        for mtype, mflags, body in msgs:
            if mtype == 0x15:  # attribute info: dense storage
                flags = body[1]
                p = 2 + (2 if flags & 1 else 0)
                if self.b.offset_from(body, p) is not None:
                    raise Unsupported('dense attribute storage (fractal heap)')
            if mtype != 0x0C:
                continue
            if mflags & 0x02:
                raise Unsupported('shared attribute message')
            ver = body[0]
            nsz, tsz, ssz = (int.from_bytes(body[2 + 2 * k:4 + 2 * k], 'little') for k in range(3))
            if ver == 1:
                p = 8
                name = _cstr(body, p)
                p += _pad8(nsz)
                tb, p2 = body[p:p + tsz], p + _pad8(tsz)
                sb, p3 = body[p2:p2 + ssz], p2 + _pad8(ssz)
            elif ver in (2, 3):
                if body[1] & 0x03:
                    raise Unsupported('attribute with shared datatype / dataspace')
                p = 8 + (1 if ver == 3 else 0)
                name = _cstr(body, p)
                p += nsz
                tb, p2 = body[p:p + tsz], p + tsz
                sb, p3 = body[p2:p2 + ssz], p2 + ssz
            else:
                raise Unsupported('attribute version %d' % ver)
            dtype, _ = self.datatype(tb)
            shape = self.dataspace(sb)
            if shape is None:
                self.lossy.append('null dataspace attribute %r' % name)
                continue
            out.append((name.decode('utf-8'), self.decode(body[p3:], dtype, shape)))
        return out

    # ------------------------------------------------------------------ datasets
    def _chunks_btree(self, addr, rank, out):
        d, b = self.data, self.b
        if d[addr:addr + 4] != b'TREE' or d[addr + 4] != 1:
            raise IOError('bad chunk B-tree node')
        level, n = d[addr + 5], b.u(addr + 6, 2)
        p = addr + 8 + 2 * b.osz
        ksz = 8 + 8 * (rank + 1)
        for _ in range(n):
            csize, cmask = b.u(p, 4), b.u(p + 4, 4)
            offs = tuple(b.u(p + 8 + 8 * k, 8) for k in range(rank))
            child = b.offset(p + ksz)
            p += ksz + b.osz
            if level > 0:
                self._chunks_btree(child, rank, out)
            else:
                out.append((offs, csize, cmask, child))

    @staticmethod
    def _unfilter(raw, filters, mask, itemsize):
        for k in range(len(filters) - 1, -1, -1):
            if mask & (1 << k):
                continue
            fid, cd = filters[k]
            if fid == 1:
                raw = zlib.decompress(raw)
            elif fid == 2:
                sz = cd[0] if cd else itemsize
                n = len(raw) // sz
                a = np.frombuffer(raw, dtype=np.uint8, count=n * sz).reshape(sz, n)
                raw = a.T.tobytes() + raw[n * sz:]
            elif fid == 3:
                raw = raw[:-4]
            else:
                raise Unsupported('filter %d' % fid)
        return raw

    def dataset(self, msgs):
        """-> (dtype, shape, loader, compression) or None when the header is not a dataset"""
        b, d = self.b, self.data
        dtype = shape = layout = None
        filters = []
        for mtype, mflags, body in msgs:
            if mtype in (0x01, 0x03) and mflags & 0x02:
                raise Unsupported('shared datatype / dataspace message')
            if mtype == 0x01:
                shape = self.dataspace(body)
                if shape is None:
                    raise Unsupported('null dataspace dataset')
            elif mtype == 0x03:
                dtype, _ = self.datatype(body)
            elif mtype == 0x08:
                layout = body
            elif mtype == 0x0B:
                ver, nf = body[0], body[1]
                p = 8 if ver == 1 else 2
                for _ in range(nf):
                    fid = int.from_bytes(body[p:p + 2], 'little')
                    p += 2
                    nlen = 0
                    if ver == 1 or fid >= 256:
                        nlen = int.from_bytes(body[p:p + 2], 'little')
                        p += 2
                    ncd = int.from_bytes(body[p + 2:p + 4], 'little')
                    p += 4
                    p += _pad8(nlen) if ver == 1 else nlen
                    cd = [int.from_bytes(body[p + 4 * k:p + 4 * k + 4], 'little') for k in range(ncd)]
                    p += 4 * ncd
                    if ver == 1 and ncd % 2:
                        p += 4
                    filters.append((fid, cd))
        if dtype is None or shape is None or layout is None:
            return None
        itemsize = dtype.itemsize
        nbytes = int(np.prod(shape, dtype=np.int64)) * itemsize if shape else itemsize
        ver = layout[0]
        compression = 'gzip' if any(f[0] == 1 for f in filters) else None
        if ver in (1, 2):  # HDF5 1.6 and older: version, dimensionality, class, reserved(5), address, dimensions
            ndim, cls = layout[1], layout[2]
            p = 8
            addr12 = None
            if cls != 0:
                addr12 = b.offset_from(layout, p)
                p += b.osz
            dims12 = tuple(int.from_bytes(layout[p + 4 * k:p + 4 * k + 4], 'little') for k in range(ndim))
            p += 4 * ndim
            if cls == 2:   # rewritten as the version-3 message the code below understands
                layout = bytes([3, 2, ndim]) + (UNDEF if addr12 is None else addr12 - b.base).to_bytes(b.osz, 'little') + \
                    b''.join(int(x).to_bytes(4, 'little') for x in dims12)
            elif cls == 1:
                layout = bytes([3, 1]) + (UNDEF if addr12 is None else addr12 - b.base).to_bytes(b.osz, 'little') + \
                    int(nbytes).to_bytes(b.lsz, 'little')
            else:
                size = int.from_bytes(layout[p:p + 4], 'little')
                layout = bytes([3, 0]) + size.to_bytes(2, 'little') + layout[p + 4:p + 4 + size]
            ver = 3
        if ver not in (3, 4):
            raise Unsupported('data layout version %d' % ver)
        cls = layout[1]
        if cls == 0:  # compact
            size = int.from_bytes(layout[2:4], 'little')
            raw0 = layout[4:4 + size]
            return dtype, shape, (lambda: self.decode(raw0, dtype, shape)), compression
        if cls == 1:  # contiguous
            addr = b.offset_from(layout, 2)

            def load_contig():
                raw = bytes(nbytes) if addr is None else bytes(d[addr:addr + nbytes])
                return self.decode(raw, dtype, shape)
            return dtype, shape, load_contig, compression
        if cls != 2:
            raise Unsupported('data layout class %d' % cls)
        if ver == 3:
            ndim = layout[2]
            btree = b.offset_from(layout, 3)
            cdims = tuple(int.from_bytes(layout[3 + b.osz + 4 * k:7 + b.osz + 4 * k], 'little') for k in range(ndim - 1))
            single = None
        else:
            flags, ndim, enc = layout[2], layout[3], layout[4]
            cdims = tuple(int.from_bytes(layout[5 + enc * k:5 + enc * (k + 1)], 'little') for k in range(ndim - 1))
            p = 5 + enc * ndim
            itype = layout[p]
            p += 1
            if itype != 1:
                raise Unsupported('chunk index type %d' % itype)
            csize, cmask = nbytes, 0
            if flags & 0x02:
                csize = int.from_bytes(layout[p:p + b.lsz], 'little')
                cmask = int.from_bytes(layout[p + b.lsz:p + b.lsz + 4], 'little')
                p += b.lsz + 4
            single = (csize, cmask, b.offset_from(layout, p))
            btree = None
        rank = len(shape)

        def load_chunked():
            if isinstance(dtype, VlenStr):
                raise Unsupported('chunked variable-length strings')
            full = np.zeros(shape, dtype=dtype)
            chunks = []
            if single is not None:
                if single[2] is not None:
                    chunks.append(((0,) * rank, single[0], single[1], single[2]))
            elif btree is not None:
                self._chunks_btree(btree, rank, chunks)
            for offs, csize, cmask, addr in chunks:
                raw = self._unfilter(bytes(d[addr:addr + csize]), filters, cmask, itemsize)
                block = np.frombuffer(raw, dtype=dtype, count=int(np.prod(cdims, dtype=np.int64))).reshape(cdims)
                sel_dst = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
                sel_src = tuple(slice(0, sl.stop - sl.start) for sl in sel_dst)
                full[sel_dst] = block[sel_src]
            return full

        def stored_chunk():
            """the compressed bytes of a dataset held as one deflate chunk covering it, for verbatim re-writing"""
            if [f[0] for f in filters] != [1] or tuple(cdims) != tuple(shape) or isinstance(dtype, VlenStr):
                return None
            chunks = []
            if single is not None:
                if single[2] is not None:
                    chunks.append(((0,) * rank, single[0], single[1], single[2]))
            elif btree is not None:
                self._chunks_btree(btree, rank, chunks)
            if len(chunks) != 1 or chunks[0][2] != 0:
                return None
            return bytes(d[chunks[0][3]:chunks[0][3] + chunks[0][1]])
        load_chunked.stored_chunk = stored_chunk
        return dtype, shape, load_chunked, compression


def _offset_from(self, body, p):
    v = int.from_bytes(body[p:p + self.osz], 'little')
    return None if v == (1 << (8 * self.osz)) - 1 else v + self.base


def _u_from(self, body, p, n):
    return int.from_bytes(body[p:p + n], 'little')


_Buf.offset_from = _offset_from
_Buf.u_from = _u_from


# =====================================================================================  writer
def _dt_message(dt):
    """numpy dtype (or VlenStr) -> datatype message bytes (version 1 encodings)"""
    if isinstance(dt, VlenStr):
        base = struct.pack('<BBBBI', 0x13, 0x10 if dt.utf8 else 0x00, 0, 0, 1)
        return struct.pack('<BBBBI', 0x19, 0x01, 0x01 if dt.utf8 else 0x00, 0, 16) + base
    dt = np.dtype(dt)
    if dt.subdtype is not None:
        raise Unsupported('array datatype %r' % (dt,))
    if dt.kind in 'iub':
        be = 1 if dt.byteorder == '>' else 0
        signed = 8 if dt.kind == 'i' else 0
        return struct.pack('<BBBBIHH', 0x10, be | signed, 0, 0, dt.itemsize, 0, 8 * dt.itemsize)
    if dt.kind == 'f':
        be = 1 if dt.byteorder == '>' else 0
        spec = {2: (15, 10, 5, 0, 10, 15), 4: (31, 23, 8, 0, 23, 127), 8: (63, 52, 11, 0, 52, 1023)}[dt.itemsize]
        sign, eloc, esz, mloc, msz, bias = spec
        return struct.pack('<BBBBIHHBBBBI', 0x11, 0x20 | be, sign, 0, dt.itemsize, 0, 8 * dt.itemsize,
                           eloc, esz, mloc, msz, bias)
    if dt.kind == 'S':
        return struct.pack('<BBBBI', 0x13, 0x01, 0, 0, max(dt.itemsize, 1))  # null-padded, ASCII (as h5py)
    if dt.kind == 'V' and dt.names:
        out = struct.pack('<BBBBI', 0x16, len(dt.names) & 0xff, len(dt.names) >> 8, 0, dt.itemsize)
        for name in dt.names:
            mt, off = dt.fields[name][0], dt.fields[name][1]
            nb = name.encode('utf-8') + b'\x00'
            out += nb + b'\x00' * (_pad8(len(nb)) - len(nb))
            out += struct.pack('<IB3xI4x16x', off, 0, 0)
            out += _dt_message(mt)
        return out
    raise Unsupported('numpy dtype %r' % (dt,))


def _ds_message(shape):
    if shape == ():
        return struct.pack('<BBB5x', 1, 0, 0)
    out = struct.pack('<BBB5x', 1, len(shape), 1)
    for s in shape:
        out += struct.pack('<Q', s)
    for s in shape:
        out += struct.pack('<Q', s)  # maximum dimensions = current dimensions
    return out


def _as_array(value):
    """Python / numpy value -> (array-like for storage, datatype, shape); str becomes a VlenStr scalar"""
    if isinstance(value, str):
        return value, VlenStr(True), ()
    if isinstance(value, bytes):
        value = np.bytes_(value)
    arr = np.asarray(value)
    if arr.dtype.kind == 'U':
        if arr.shape == ():
            return str(arr[()]), VlenStr(True), ()
        arr = np.char.encode(arr, 'utf-8')
    if arr.dtype.kind == 'O':
        if arr.shape != () and all(isinstance(v, str) for v in arr.ravel()):
            return arr, VlenStr(True), arr.shape
        raise TypeError("Object dtype dtype('O') has no native HDF5 equivalent")
    if arr.dtype.kind == 'b':
        arr = arr.astype(np.uint8)  # h5py writes an enum; a plain byte reads back the same truth values
    if arr.dtype.kind == 'S' and arr.dtype.itemsize == 0:
        arr = arr.astype('S1')
    return np.ascontiguousarray(arr), arr.dtype, arr.shape


class Writer(object):
    """serialises a MemGroup tree (fast5_io object model) into HDF5 bytes"""
    GROUP_LEAF_K = 4
    GROUP_INT_K = 16
    CHUNK_K = 32
    SUPERBLOCK = 96

    def __init__(self):
        self.buf = bytearray(self.SUPERBLOCK)
        self.strings = []      # variable-length string payloads, global heap index = position + 1
        self.gcol_addr = None

    def alloc(self, data, align=8):
        while len(self.buf) % align:
            self.buf.append(0)
        addr = len(self.buf)
        self.buf += data
        return addr

    # ---- variable-length strings: collected first so that the collection's address is known
    def _collect(self, group):
        from . import fast5_io as f5
        for node in [group] + list(group._children.values()):
            if node is not group and isinstance(node, f5.MemGroup):
                self._collect(node)
                continue
            for v in node.attrs.values():
                val, dt, shape = _as_array(v)
                if isinstance(dt, VlenStr):
                    n = 1 if shape == () else int(np.prod(shape))
                    self._n_strings += n
                    self._string_bytes += sum(_pad8(len(s.encode('utf-8', 'surrogateescape'))) + 16
                                              for s in ([val] if shape == () else list(val.ravel())))

    def _vlen_element(self, s):
        raw = s.encode('utf-8', 'surrogateescape')
        self.strings.append(raw)
        return struct.pack('<IQI', len(raw), self.gcol_addr, len(self.strings))

    def _attr_message(self, name, value):
        val, dt, shape = _as_array(value)
        if isinstance(dt, VlenStr):
            data = b''.join(self._vlen_element(s) for s in ([val] if shape == () else list(val.ravel())))
        else:
            data = val.tobytes()
        nb = name.encode('utf-8') + b'\x00'
        tb, sb = _dt_message(dt), _ds_message(shape)
        body = struct.pack('<BxHHH', 1, len(nb), len(tb), len(sb))
        for part in (nb, tb, sb):
            body += part + b'\x00' * (_pad8(len(part)) - len(part))
        body += data
        if len(body) > 65528:
            raise Unsupported('attribute %r larger than 64 KB' % name)
        return (0x0C, 0, body)

    def _object_header(self, msgs):
        body = b''
        for mtype, mflags, data in msgs:
            data = data + b'\x00' * (_pad8(len(data)) - len(data))
            body += struct.pack('<HHB3x', mtype, len(data), mflags) + data
        head = struct.pack('<BxHII4x', 1, len(msgs), 1, len(body))
        return self.alloc(head + body)

    def _dataset(self, ds):
        comp = None
        if getattr(ds, '_cache', 0) is None and hasattr(ds._loader, 'stored_chunk'):
            comp = ds._loader.stored_chunk()  # untouched since it was read: its chunk is copied as it is
        if comp is not None:
            dt, shape, raw = ds.dtype, ds.shape, b''
        else:
            val, dt, shape = _as_array(ds._data)
            if isinstance(dt, VlenStr):
                raise Unsupported('variable-length string dataset')
            raw = val.tobytes()
        msgs = [(0x01, 0, _ds_message(shape)), (0x03, 1, _dt_message(dt))]
        nelem = int(np.prod(shape, dtype=np.int64)) if shape else 1
        if ds.compression and nelem > 0 and shape != ():
            if comp is None:
                comp = zlib.compress(raw, 4)
            caddr = self.alloc(comp)
            rank = len(shape)
            ksz = 8 + 8 * (rank + 1)
            node = bytearray(24 + (2 * self.CHUNK_K + 1) * ksz + 2 * self.CHUNK_K * 8)
            node[0:8] = struct.pack('<4sBBH', b'TREE', 1, 0, 1)
            node[8:24] = struct.pack('<QQ', UNDEF, UNDEF)
            p = 24
            node[p:p + 8] = struct.pack('<II', len(comp), 0)            # key 0: the chunk at the origin
            p += ksz
            node[p:p + 8] = struct.pack('<Q', caddr)
            p += 8
            node[p:p + 8] = struct.pack('<II', 0, 0)                    # key 1: one chunk past the end
            node[p + 8:p + 8 + 8 * rank] = b''.join(struct.pack('<Q', s) for s in shape)
            baddr = self.alloc(bytes(node))
            msgs.append((0x05, 0, struct.pack('<BBBB', 2, 3, 2, 0)))    # fill value: incremental, if set, undefined
            msgs.append((0x0B, 0, struct.pack('<BB6xHHHH8sI4x', 1, 1, 1, 8, 1, 1, b'deflate\x00', 4)))
            layout = struct.pack('<BBBQ', 3, 2, rank + 1, baddr)
            layout += b''.join(struct.pack('<I', s) for s in shape) + struct.pack('<I', dt.itemsize)
            msgs.append((0x08, 0, layout))
        else:
            addr = self.alloc(raw) if raw and nelem > 0 else UNDEF
            msgs.append((0x05, 0, struct.pack('<BBBB', 2, 2, 2, 0)))    # fill value: late, if set, undefined
            msgs.append((0x08, 0, struct.pack('<BBQQ', 3, 1, addr, len(raw) if nelem > 0 else 0)))
        for k, v in ds.attrs.items():
            msgs.append(self._attr_message(k, v))
        return self._object_header(msgs)

    def _group(self, group):
        """-> (object header address, B-tree address, heap address)"""
        from . import fast5_io as f5
        entries = []
        for name, node in group._children.items():
            if isinstance(node, f5.MemGroup):
                haddr, baddr, laddr = self._group(node)
                entries.append((name.encode('utf-8'), haddr, 1, struct.pack('<QQ', baddr, laddr)))
            else:
                entries.append((name.encode('utf-8'), self._dataset(node), 0, b'\x00' * 16))
        entries.sort(key=lambda e: e[0])
        # local heap: "" at offset 0, then the names
        heap = bytearray(8)
        name_off = {}
        for nm, _, _, _ in entries:
            name_off[nm] = len(heap)
            heap += nm + b'\x00'
            while len(heap) % 8:
                heap.append(0)
        free_off = len(heap)
        heap += struct.pack('<QQ', 1, 16) + b'\x00' * 0                 # one free block closing the segment
        daddr = self.alloc(bytes(heap))
        laddr = self.alloc(struct.pack('<4sB3xQQQ', b'HEAP', 0, len(heap), free_off, daddr))
        # symbol table nodes of up to 2 * leaf K entries, then the B-tree over them
        cap = 2 * self.GROUP_LEAF_K
        leaves = []
        for lo in range(0, len(entries), cap):
            part = entries[lo:lo + cap]
            node = bytearray(8 + cap * 40)
            node[0:8] = struct.pack('<4sBxH', b'SNOD', 1, len(part))
            for k, (nm, haddr, cache, scratch) in enumerate(part):
                node[8 + 40 * k:48 + 40 * k] = struct.pack('<QQI4x', name_off[nm], haddr, cache) + scratch
            leaves.append((self.alloc(bytes(node)), name_off[part[-1][0]]))
        level = 0
        nodes = leaves
        cap_i = 2 * self.GROUP_INT_K
        while True:
            parents = []
            low_key = 0  # heap offset of the name every entry of the node sorts after: "" for the leftmost node
            for lo in range(0, max(len(nodes), 1), cap_i):
                part = nodes[lo:lo + cap_i]
                node = bytearray(24 + (cap_i + 1) * 8 + cap_i * 8)
                node[0:8] = struct.pack('<4sBBH', b'TREE', 0, level, len(part))
                node[8:24] = struct.pack('<QQ', UNDEF, UNDEF)
                p = 24
                node[p:p + 8] = struct.pack('<Q', low_key)
                low_key = part[-1][1] if part else 0
                p += 8
                for addr, last_key in part:
                    node[p:p + 16] = struct.pack('<QQ', addr, last_key)
                    p += 16
                parents.append((self.alloc(bytes(node)), part[-1][1] if part else 0))
            if len(parents) == 1:
                baddr = parents[0][0]
                break
            # siblings of one level point at each other
            for k, (addr, _) in enumerate(parents):
                left = parents[k - 1][0] if k > 0 else UNDEF
                right = parents[k + 1][0] if k + 1 < len(parents) else UNDEF
                self.buf[addr + 8:addr + 24] = struct.pack('<QQ', left, right)
            nodes = parents
            level += 1
        msgs = [(0x11, 0, struct.pack('<QQ', baddr, laddr))]
        for k, v in group.attrs.items():
            msgs.append(self._attr_message(k, v))
        return self._object_header(msgs), baddr, laddr

    def serialise(self, root):
        self._n_strings = 0
        self._string_bytes = 0
        self._collect(root)
        if self._n_strings:
            size = max(4096, _pad8(16 + self._string_bytes + 16))
            self.gcol_addr = self.alloc(bytes(size))
            self._gcol_size = size
        haddr, baddr, laddr = self._group(root)
        if self._n_strings:
            col = bytearray(struct.pack('<4sB3xQ', b'GCOL', 1, self._gcol_size))
            for k, raw in enumerate(self.strings):
                col += struct.pack('<HH4xQ', k + 1, 1, len(raw)) + raw + b'\x00' * (_pad8(len(raw)) - len(raw))
            free = self._gcol_size - len(col)
            assert free >= 0
            if free >= 16:
                col += struct.pack('<HH4xQ', 0, 0, free)
            self.buf[self.gcol_addr:self.gcol_addr + len(col)] = col
        while len(self.buf) % 8:
            self.buf.append(0)
        sb = SIGNATURE + struct.pack('<BBBBBBBBHHI', 0, 0, 0, 0, 0, 8, 8, 0, self.GROUP_LEAF_K, self.GROUP_INT_K, 0)
        sb += struct.pack('<QQQQ', 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack('<QQI4xQQ', 0, haddr, 1, baddr, laddr)
        assert len(sb) == self.SUPERBLOCK
        self.buf[0:self.SUPERBLOCK] = sb
        return bytes(self.buf)


# =====================================================================================  file object
# A group whose members are decoded from the file when somebody first asks for them: a read-only open of a FAST5
# file touches four or five of its twenty-odd objects (resquiggle.py:328-341).
def _lazy_group_class():
    from . import fast5_io as f5

    class LazyGroup(f5.MemGroup):
        def __init__(self, build, links):
            f5.MemGroup.__init__(self)
            self.__dict__.pop('_children', None)
            self._build, self._links, self._members = build, links, None

        @property
        def _children(self):
            if self._members is None:
                self._members = {}
                for name, child in self._links:
                    self._members[name] = self._build(child)
            return self._members

        @_children.setter
        def _children(self, value):  # MemGroup.__init__ assigns an empty dict
            pass

    return LazyGroup


_LAZY_GROUP = []


def load_tree(data, lazy=False):
    """HDF5 bytes -> (fast5_io.MemGroup tree with lazily decoded datasets, reasons a rewrite would be lossy).
    lazy: groups decode their members on first access as well (read-only opens); `lossy` is then only complete for
    what has been visited, which is why files opened for writing are always decoded whole."""
    from . import fast5_io as f5
    rd = Reader(data)
    seen = {}
    if lazy and not _LAZY_GROUP:
        _LAZY_GROUP.append(_lazy_group_class())

    def build(addr):
        if addr in seen:
            rd.lossy.append('hard link to an object that is already linked elsewhere')
        seen[addr] = True
        msgs = rd.messages(addr)
        links = rd.links(msgs)
        if links is not None:
            if lazy:
                node = _LAZY_GROUP[0](build, links)
            else:
                node = f5.MemGroup()
                for name, child in links:
                    node._children[name] = build(child)
        else:
            ds = rd.dataset(msgs)
            if ds is None:
                raise Unsupported('object at %d is neither a group nor a dataset' % addr)
            dtype, shape, loader, compression = ds
            node = f5.LazyDataset(loader, dtype, shape, compression)
        for k, v in rd.attributes(msgs):
            dict.__setitem__(node.attrs, k, v)
        return node

    return build(rd.root_addr), rd.lossy


def dump_tree(root):
    return Writer().serialise(root)


class NativeFile(object):
    """h5py.File look-alike over this module: 'r', 'r+', 'w', 'a'"""

    def __init__(self, name, mode='r'):
        from . import fast5_io as f5
        self.filename = name
        self.mode = mode
        self._open = True
        exists = os.path.exists(name)
        if mode in ('r', 'r+') and not exists:
            raise IOError("Unable to open file (unable to open file: name = '%s', errno = 2)" % name)
        if mode == 'w' or (mode == 'a' and not exists):
            self._root = f5.MemGroup()
            self._dirty = True
        else:
            with open(name, 'rb') as fp:
                data = fp.read()
            self._root, lossy = load_tree(data, lazy=(mode == 'r'))
            self._dirty = False
            if mode != 'r' and lossy:
                raise Unsupported('refusing to rewrite %s: %s' % (name, '; '.join(sorted(set(lossy)))))
        self._writable = mode != 'r'
        if self._writable:
            self._root._set_hook(self._touch)

    def _touch(self):
        self._dirty = True

    def __getattr__(self, item):
        return getattr(self._root, item)

    def __getitem__(self, path):
        return self._root[path]

    def __contains__(self, path):
        return path in self._root

    def __delitem__(self, path):
        if not self._writable:
            raise KeyError("Couldn't delete link (no write intent on file)")
        del self._root[path]
        self._dirty = True

    def flush(self):
        if self._writable and self._dirty:
            data = dump_tree(self._root)
            tmp = '%s.tmp%d' % (self.filename, os.getpid())
            with open(tmp, 'wb') as fp:
                fp.write(data)
            os.replace(tmp, self.filename)
            self._dirty = False

    def close(self):
        if self._open:
            self.flush()
            self._open = False

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
