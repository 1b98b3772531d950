"""Host-side mirror of the reference's container API, package `minnow` (go/writer.go,
go/reader.go, go/group.go): same method names, argument meaning, panics-as-exceptions and
on-disk bytes.  The container bookkeeping (48-byte header, user headers, group starts, tail
arrays, group tails) is host logic exactly as in the reference; every block of an IntGroup /
FloatGroup -- and the min-offset + bit-packing of the group tails (go/group.go:215-232) -- goes
through the C ABI of libminnow_b200 (`Context`), i.e. runs on the GPU.  There is no CPU codec
here.

    wr = minnow.Create("halos.minnow", ctx)        # go/writer.go:32
    wr.Header(np.int64([1, 2, 3]))                 # :42
    wr.FloatGroup(n, (0.0, 125.0), 0.001)          # :72   (periodic, like the reference)
    wr.Data(x0); wr.Data(x1)                       # :90   one block per call
    wr.DataBlocks(xs)                              # extension: all blocks of the group in ONE GPU call
    wr.Close()                                     # :107

    rd = minnow.Open("halos.minnow", ctx)          # go/reader.go:28
    rd.Data(b)                                     # :114  (jitter policy: rd.jitter)
"""
import io
import struct

import numpy as np

from .capi import FloatDesc, Jitter, JITTER_CENTER, array_bytes, float_group_pixels

Magic = 0xacedad          # go/minnow.go:7-8
Version = 1

# group type codes, go/group.go:11-24
(Int64Group, Int32Group, Int16Group, Int8Group, Uint64Group, Uint32Group, Uint16Group, Uint8Group,
 Float64Group, Float32Group, IntGroup, FloatGroup) = range(12)
_FIXED = {Int64Group: "<i8", Int32Group: "<i4", Int16Group: "<i2", Int8Group: "<i1", Uint64Group: "<u8",
          Uint32Group: "<u4", Uint16Group: "<u2", Uint8Group: "<u1", Float64Group: "<f8", Float32Group: "<f4"}


def _type_match(x, gt):
    """TypeMatch, go/group.go:43-71: the block's element type must be the group's."""
    want = np.dtype("<i8") if gt == IntGroup else np.dtype("<f4") if gt == FloatGroup else np.dtype(_FIXED[gt])
    x = np.asarray(x)
    if x.dtype != want:
        raise TypeError("Data of type %s written to group of type %d (wants %s)." % (x.dtype, gt, want))
    return np.ascontiguousarray(x)


class _Group:
    """blockIndex (go/block_index.go) + the per-group state the tail needs."""

    def __init__(self, gt, start_block, n):
        self.gt, self.start_block, self.N = gt, start_block, int(n)
        self.sizes = []            # byte size of every block: offsets are their running sum
        self.mins, self.bits = [], []
        self.low = self.high = np.float32(0)
        self.pixels, self.periodic = 0, 1

    def block_offset(self, b):     # go/block_index.go:25-35: offsets[b - start - 1], kept as a running-sum array
        if len(getattr(self, "_ends", ())) != len(self.sizes):
            self._ends = np.cumsum(np.asarray(self.sizes, np.int64)) if self.sizes else np.zeros(0, np.int64)
        k = b - self.start_block
        return int(self._ends[k - 1]) if k > 0 else 0


class Writer:
    """minnow.Writer, go/writer.go."""

    def __init__(self, f, ctx):
        self.f, self.ctx = f, ctx
        self.headers = self.blocks = 0
        self.curr = None
        self.groups, self.header_offsets, self.header_sizes = [], [], []
        self.group_blocks, self.group_offsets = [], []
        self.f.write(b"\0" * 48)                               # placeholder header, go/writer.go:37

    def Header(self, x):                                       # go/writer.go:42-52
        data = x if isinstance(x, (bytes, bytearray)) else np.ascontiguousarray(x).tobytes()
        self.header_offsets.append(self.f.tell())
        self.header_sizes.append(len(data))
        self.f.write(data)
        self.headers += 1
        self.curr = None
        return self.headers - 1

    def _new_group(self, g):                                   # go/writer.go:78-87
        self.curr = g
        self.groups.append(g)
        self.group_blocks.append(0)
        self.group_offsets.append(self.f.tell())

    def FixedSizeGroup(self, group_type, N):                   # :56-58
        if group_type not in _FIXED:
            raise ValueError("Unrecognized group type, %d." % group_type)
        self._new_group(_Group(group_type, self.blocks, N))

    def IntGroup(self, N):                                     # :62-64
        self._new_group(_Group(IntGroup, self.blocks, N))

    def FloatGroup(self, N, lim, dx):                          # :72-75 (always periodic)
        g = _Group(FloatGroup, self.blocks, N)
        g.low, g.high = np.float32(lim[0]), np.float32(lim[1])
        g.pixels = float_group_pixels(g.low, g.high, np.float32(dx))
        g.periodic = 1
        g.log10 = g.clamp = 0
        self._new_group(g)

    def Data(self, x):                                         # :90-104
        if self.curr is None:
            raise RuntimeError("Data written to minnow.Writer without assigning Group first.")
        x = _type_match(x, self.curr.gt)
        if len(x) != self.curr.N:
            raise ValueError("block of %d elements written to a group of N = %d" % (len(x), self.curr.N))
        self._blocks(x, 1)
        return self.blocks - 1

    def DataBlocks(self, x):
        """All of x (len = k * N) as k consecutive blocks of the current group: one GPU call
        instead of k (the bytes on disk are those of k Data calls)."""
        if self.curr is None:
            raise RuntimeError("Data written to minnow.Writer without assigning Group first.")
        x = _type_match(np.asarray(x).reshape(-1), self.curr.gt)
        if self.curr.N == 0 or len(x) % self.curr.N:
            raise ValueError("%d elements are not whole blocks of N = %d" % (len(x), self.curr.N))
        self._blocks(x, len(x) // self.curr.N)
        return self.blocks - 1

    def _blocks(self, x, k):
        g = self.curr
        if g.gt in _FIXED:                                     # fixedSizeGroup.writeData, go/group.go:150-153
            self.f.write(x.tobytes())
            g.sizes += [x.itemsize * g.N] * k
        else:
            if g.gt == IntGroup:                               # intGroup.writeData, go/group.go:242-255
                mins, bits, offs, data = self.ctx.encode_int_group(x, g.N, k)
            else:                                              # floatGroup.writeData, go/group.go:312-327
                d = FloatDesc.make(g.low, g.high, g.pixels, g.periodic, getattr(g, "log10", 0), getattr(g, "clamp", 0))
                mins, bits, offs, data = self.ctx.encode_float_group(d, x, g.N, k)
            self.f.write(data.tobytes())
            g.mins += [int(m) for m in mins]
            g.bits += [int(b) for b in bits]
            g.sizes += [array_bytes(int(b), g.N) for b in bits]
        self.group_blocks[-1] += k
        self.blocks += k

    def EncodedBlock(self, mn, bits, data):
        """One block of the current IntGroup / FloatGroup that was encoded elsewhere (minh.Writer.Block encodes all
        columns of a block in one GPU call): what writeData records, go/group.go:249-254."""
        g = self.curr
        if g is None or g.gt in _FIXED:
            raise RuntimeError("EncodedBlock needs a current IntGroup or FloatGroup")
        self.f.write(bytes(data))
        g.mins.append(int(mn)); g.bits.append(int(bits)); g.sizes.append(array_bytes(int(bits), g.N))
        self.group_blocks[-1] += 1
        self.blocks += 1
        return self.blocks - 1

    def _write_int_array_tail(self, x):
        """`write` closure of intGroup.writeTail, go/group.go:216-224: min, bits, packed (x - min)."""
        x = np.asarray(x, np.int64)
        if len(x) == 0:                                        # int64Min of an empty slice is 0, Bits is 0
            self.f.write(struct.pack("<qq", 0, 0))
            return
        mins, bits, _, data = self.ctx.encode_int_group(x, len(x), 1)
        self.f.write(struct.pack("<qq", int(mins[0]), int(bits[0])))
        self.f.write(data.tobytes())

    def Close(self):                                           # go/writer.go:107-141
        tail_start = self.f.tell()
        for arr in (self.header_offsets, self.header_sizes, self.group_offsets,
                    [g.gt for g in self.groups], self.group_blocks):
            self.f.write(np.asarray(arr, "<i8").tobytes())
        for g in self.groups:
            self.f.write(struct.pack("<qqq", g.N, g.start_block, len(g.sizes)))     # N, startBlock, blocks
            if g.gt in (IntGroup, FloatGroup):                                         # go/group.go:215-232
                self._write_int_array_tail(g.mins)
                self._write_int_array_tail(g.bits)
            if g.gt == FloatGroup:                                                     # go/group.go:328-334
                self.f.write(struct.pack("<ffqB", g.low, g.high, g.pixels, g.periodic))
        self.f.seek(0)
        self.f.write(struct.pack("<QQQQQq", Magic, Version, len(self.groups), self.headers, self.blocks, tail_start))
        self.f.flush()
        if not isinstance(self.f, io.BytesIO):
            self.f.close()


def Create(fname, ctx):
    """minnow.Create, go/writer.go:32-39.  fname: a path or a writable, seekable binary file object."""
    f = fname if hasattr(fname, "write") else open(fname, "w+b")
    return Writer(f, ctx)


class Reader:
    """minnow.Reader, go/reader.go."""

    def __init__(self, f, ctx, jitter=None):
        self.f, self.ctx = f, ctx
        self.jitter = jitter if jitter is not None else Jitter.make(JITTER_CENTER)
        hd = f.read(48)
        if len(hd) < 48:
            raise ValueError("not a minnow file: shorter than its 48-byte header")
        magic, version, groups, headers, blocks, tail_start = struct.unpack("<QQQQQq", hd)
        if magic != Magic:                                     # go/reader.go:39-41
            raise ValueError("not a minnow file. Magic number is %x, not %x." % (magic, Magic))
        if version != Version:                                 # :42-46
            raise ValueError("file was written with minnow version %d, but this code has version %d." % (version, Version))
        self.groups_n, self.headers, self.blocks = groups, headers, blocks
        f.seek(tail_start)                                     # :55

        def i64s(n):
            return np.frombuffer(f.read(8 * n), "<i8").astype(np.int64)
        self.header_offsets, self.header_sizes = i64s(headers), i64s(headers)
        self.group_offsets, self.group_types, group_blocks = i64s(groups), i64s(groups), i64s(groups)
        self.groups = [self._group_from_tail(int(gt)) for gt in self.group_types]       # :74-76
        self.block_index = np.repeat(np.arange(groups), group_blocks)                   # :78-85

    def _read_int_array_tail(self, n):
        """`read` closure of newIntGroupFromTail, go/group.go:191-198."""
        mn, bits = struct.unpack("<qq", self.f.read(16))
        if n == 0:
            return []
        data = np.frombuffer(self.f.read(array_bytes(bits, n)), np.uint8)
        out = self.ctx.decode_int_blocks(data, np.zeros(1, np.int64), np.array([mn], np.int64),
                                         np.array([bits], np.int64), n)
        return [int(v) for v in np.asarray(out).reshape(-1)]

    def _group_from_tail(self, gt):                            # go/group.go:93-103
        if gt not in _FIXED and gt not in (IntGroup, FloatGroup):
            raise ValueError("Unrecognized group type, %d." % gt)
        N, start, blocks = struct.unpack("<qqq", self.f.read(24))
        g = _Group(gt, start, N)
        if gt in _FIXED:                                       # go/group.go:124-137
            g.sizes = [np.dtype(_FIXED[gt]).itemsize * N] * blocks
            return g
        g.mins = self._read_int_array_tail(blocks)             # go/group.go:186-213
        g.bits = self._read_int_array_tail(blocks)
        g.sizes = [array_bytes(b, N) for b in g.bits]
        if gt == FloatGroup:                                   # go/group.go:336-344
            g.low, g.high, g.pixels, g.periodic = struct.unpack("<ffqB", self.f.read(17))
        return g

    def Header(self, i, dtype=None):                           # go/reader.go:91-100
        self.f.seek(int(self.header_offsets[i]))
        raw = self.f.read(int(self.header_sizes[i]))
        if dtype is None:
            return raw
        dt = np.dtype(dtype)
        if len(raw) % dt.itemsize:
            raise ValueError("Header buffer has element size %d, but written header has size %d." % (dt.itemsize, len(raw)))
        return np.frombuffer(raw, dt).copy()

    def HeaderSize(self, i):                                   # :103-105
        return int(self.header_sizes[i])

    def Blocks(self):                                          # :108-110
        return int(self.blocks)

    def DataType(self, b):                                     # :130-132
        return int(self.group_types[self.block_index[b]])

    def DataLen(self, b):                                      # :135-137
        return int(self.groups[self.block_index[b]].N)

    def Data(self, b, log10=False):                            # :114-127
        """log10: the block is a minh Log column -- float32(math.Pow(10, float64(x))) of go/minh/minh.go:315-319
        is applied on the device by the decode kernel (mnw_float_desc.log10)."""
        i = int(self.block_index[b])
        g = self.groups[i]
        self.f.seek(int(self.group_offsets[i]) + g.block_offset(b))
        k = b - g.start_block
        raw = self.f.read(g.sizes[k])
        if g.gt in _FIXED:                                     # fixedSizeGroup.readData
            x = np.frombuffer(raw, _FIXED[g.gt]).copy()
            return self.ctx.pow10_f32(x) if (log10 and g.gt == Float32Group) else x
        data = np.frombuffer(raw, np.uint8)
        meta = (np.zeros(1, np.int64), np.array([g.mins[k]], np.int64), np.array([g.bits[k]], np.int64))
        if g.gt == IntGroup:                                   # intGroup.readData, go/group.go:257-263
            return np.asarray(self.ctx.decode_int_blocks(data, *meta, g.N)).reshape(-1)
        d = FloatDesc.make(g.low, g.high, g.pixels, g.periodic, 1 if log10 else 0)     # floatGroup.readData, :299-310
        jit = Jitter.make(self.jitter.mode, self.jitter.seed, self.jitter.block_id0 + b)
        return np.asarray(self.ctx.decode_float_blocks(d, data, *meta, g.N, jitter=jit)).reshape(-1)

    def Close(self):
        if not isinstance(self.f, io.BytesIO):
            self.f.close()


def Open(fname, ctx, jitter=None):
    """minnow.Open, go/reader.go:28-88.  fname: a path, bytes, or a readable, seekable binary file object."""
    if isinstance(fname, (bytes, bytearray)):
        f = io.BytesIO(bytes(fname))
    else:
        f = fname if hasattr(fname, "read") else open(fname, "rb")
    return Reader(f, ctx, jitter)
