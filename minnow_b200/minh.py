"""Host-side mirror of the reference's halo-catalogue format, package `minh`
(go/minh/minh.go): same names, argument meaning and on-disk bytes.  Every column of a block
is its own minnow group with one block (go/minh/minh.go:121-137); Float columns take the
log10 / clamp pre-transform of processFloatGroup (:141-149) fused into the quantise kernel."""
import struct

import numpy as np

from . import minnow
from .capi import FloatDesc, array_bytes, float_group_pixels

Magic = 0xbaff1ed         # go/minh/minh.go:13-16
Version = 0
basicFileType, boundaryFileType = 0, 1                                                       # :18-21

Column = np.dtype([("Type", "<i8"), ("Log", "<i4"), ("Low", "<f4"), ("High", "<f4"), ("Dx", "<f4"),
                   ("Buffer", "S232")])                                                    # :50-55
assert Column.itemsize == 256


def columns(specs):
    """specs: (Type[, Log, Low, High, Dx]) per column -> Column array"""
    a = np.zeros(len(specs), Column)
    for i, c in enumerate(specs):
        c = tuple(c) + (0,) * (5 - len(c))
        a["Type"][i], a["Log"][i], a["Low"][i], a["High"][i], a["Dx"][i] = c
    return a


class Writer:
    def __init__(self, fname, ctx, file_type=basicFileType):                                 # Create / create, :71-86
        self.ctx = ctx
        self.f = minnow.Create(fname, ctx)
        self.f.Header(struct.pack("<qqq", Magic, Version, file_type))
        self.blocks, self.block_sizes = 0, []
        self.l = self.boundary = np.float32(0)
        self.cells = 0

    def Header(self, names, text, cols):                                                     # :88-93
        self.cols = cols if isinstance(cols, np.ndarray) else columns(cols)
        self.f.Header(text.encode("ascii"))
        self.f.Header("$".join(names).encode("ascii"))
        self.f.Header(self.cols.tobytes())

    def Geometry(self, L, boundary, cells):                                                  # :95-97
        self.l, self.boundary, self.cells = np.float32(L), np.float32(boundary), int(cells)

    def Block(self, cols):                                                                   # :99-139
        if len(cols) != len(self.cols):
            raise ValueError("Expected %d columns, got %d." % (len(self.cols), len(cols)))
        N = len(cols[0])
        for i, x in enumerate(cols):
            if len(x) != N:
                raise ValueError("len(cols[%d]) = %d instead of %d" % (i, len(x), N))
        self.block_sizes.append(N)
        self.blocks += 1
        w = self.f
        # every IntGroup / FloatGroup column of the block goes to the GPU in ONE call (mnw_encode_columns);
        # processFloatGroup (:141-149) runs inside the quantise kernels (desc.log10, desc.clamp)
        quant, descs = [], {}
        for i, (x, col) in enumerate(zip(cols, self.cols)):
            t = int(col["Type"])
            if t == minnow.IntGroup:
                quant.append((i, (np.asarray(x, np.int64), None)))
            elif t == minnow.FloatGroup:
                lo, hi = np.float32(col["Low"]), np.float32(col["High"])
                d = FloatDesc.make(lo, hi, float_group_pixels(lo, hi, np.float32(col["Dx"])), 1, 1 if col["Log"] != 0 else 0, 1)
                quant.append((i, (np.asarray(x, np.float32), d)))
            elif t not in minnow._FIXED:
                raise ValueError("Unrecognized group type, %d." % t)
        enc = {}
        if quant and N > 0:
            mins, bits, packed = self.ctx.encode_columns([q for _, q in quant])
            enc = {i: (mins[j], bits[j], packed[j]) for j, (i, _) in enumerate(quant)}
        for i, (x, col) in enumerate(zip(cols, self.cols)):
            t = int(col["Type"])
            if t in minnow._FIXED:
                w.FixedSizeGroup(t, N)
                w.Data(np.asarray(x, minnow._FIXED[t]))
            elif t == minnow.IntGroup:
                w.IntGroup(N)
                if i in enc: w.EncodedBlock(*enc[i])
                else: w.Data(np.asarray(x, np.int64))
            else:
                w.FloatGroup(N, (col["Low"], col["High"]), col["Dx"])
                w.curr.log10 = 1 if col["Log"] != 0 else 0
                w.curr.clamp = 1
                if i in enc: w.EncodedBlock(*enc[i])
                else: w.Data(np.asarray(x, np.float32))

    def Close(self):                                                                         # :151-156
        self.f.Header(struct.pack("<ffq", self.l, self.boundary, self.cells))
        self.f.Header(struct.pack("<q", self.blocks))
        self.f.Header(np.asarray(self.block_sizes, "<i8").tobytes())
        self.f.Close()


def Create(fname, ctx):
    return Writer(fname, ctx)


class BoundaryWriter(Writer):
    """minh.BoundaryWriter, go/minh/boundary.go: a catalogue split into cells^3 cells, every cell stored together with the
    points of its ghost layers.  Coordinates bins the points on the GPU (the index lists stay on the device); Column
    gathers and encodes every cell's values there."""

    def __init__(self, fname, ctx):                                                          # CreateBoundary, :25-29
        super().__init__(fname, ctx, boundaryFileType)
        self.names, self.colspecs = [], []
        self.cell_sizes = self._index = None

    def Header(self, text):                                                                  # :31-33
        self.f.Header(text.encode("ascii"))

    def Block(self, cols):                                                                   # :35-37
        raise RuntimeError("Block() cannot be called for BoundaryWriter. Use")

    def Coordinates(self, x, y, z):                                                          # :39-51
        self.cell_sizes = self.ctx.boundary_coordinates(x, y, z, self.l, self.boundary, self.cells)
        self._index = None
        self._encoded_column("boundary", columns([(minnow.IntGroup,)])[0], *self.ctx.boundary_encode_flags())   # boundaryColumn, :227-246
        self.block_sizes = [int(n) for n in self.cell_sizes]
        self.blocks = len(self.cell_sizes)

    def _encoded_column(self, name, col, mins, bits, offs, data):
        self.colspecs.append(col)
        self.names.append(name)
        w, t = self.f, int(col["Type"])
        for i, N in enumerate(self.cell_sizes):
            N = int(N)
            if t == minnow.IntGroup:
                w.IntGroup(N)
            else:
                w.FloatGroup(N, (col["Low"], col["High"]), col["Dx"])
            w.EncodedBlock(mins[i], bits[i], data[offs[i]:offs[i] + array_bytes(int(bits[i]), N)])

    def Column(self, name, col, x):                                                          # :184-225
        if self.cell_sizes is None:
            raise RuntimeError("Column called before Coordinates")
        col = np.asarray(col, Column).reshape(1)[0] if not isinstance(col, np.void) else col
        t = int(col["Type"])
        if t == minnow.IntGroup:
            self._encoded_column(name, col, *self.ctx.boundary_encode_column(np.asarray(x, np.int64)))
        elif t == minnow.FloatGroup:
            lo, hi = np.float32(col["Low"]), np.float32(col["High"])
            d = FloatDesc.make(lo, hi, float_group_pixels(lo, hi, np.float32(col["Dx"])), 1, 1 if col["Log"] != 0 else 0, 1)
            self._encoded_column(name, col, *self.ctx.boundary_encode_column(np.asarray(x, np.float32), d))
        elif t in (minnow.Int64Group, minnow.Float32Group):       # fixed-size columns: a plain gather and copy, as in the format
            if self._index is None:
                self._index = self.ctx.boundary_index()[0]
            x = np.asarray(x, minnow._FIXED[t])
            self.colspecs.append(col)
            self.names.append(name)
            start = 0
            for N in self.cell_sizes:
                self.f.FixedSizeGroup(t, int(N))
                self.f.Data(x[self._index[start:start + int(N)]])
                start += int(N)
        else:
            raise ValueError("Can't write column with type flag %d" % t)

    def Close(self):                                                                         # :249-256
        self.f.Header("$".join(self.names).encode("ascii"))
        self.f.Header(np.asarray(self.colspecs, Column).tobytes() if self.colspecs else b"")
        self.f.Header(struct.pack("<ffq", self.l, self.boundary, self.cells))
        self.f.Header(struct.pack("<q", self.blocks))
        self.f.Header(np.asarray(self.block_sizes, "<i8").tobytes())
        self.f.Close()


def CreateBoundary(fname, ctx):
    return BoundaryWriter(fname, ctx)


class Reader:
    def __init__(self, fname, ctx, jitter=None):                                             # Open, :184-226
        self.f = minnow.Open(fname, ctx, jitter)
        magic, version, self.file_type = struct.unpack("<qqq", self.f.Header(0))
        if magic != Magic:
            raise ValueError("not a minh file. Expected magic number %d, but got %d." % (Magic, magic))
        if version < Version:
            raise ValueError("written with minh version %d, but reader is version %d." % (version, Version))
        self.Text = self.f.Header(1).decode("ascii")
        self.Names = self.f.Header(2).decode("ascii").split("$")
        self.Columns = self.f.Header(3, Column)
        self.L, self.Boundary, self.Cells = struct.unpack("<ffq", self.f.Header(4))
        self.Blocks = struct.unpack("<q", self.f.Header(5))[0]
        self.BlockLengths = [int(v) for v in self.f.Header(6, "<i8")]
        self.Length = sum(self.BlockLengths)

    def _index(self, name, b):                                                               # :267-283
        if name not in self.Names:
            raise KeyError("Name %s not in Reader.Names = %s." % (name, self.Names))
        c = self.Names.index(name)
        return c, (c + b * len(self.Columns) if self.file_type == basicFileType else c * self.Blocks + b)

    def IntBlock(self, b, names):                                                            # :267-294
        out = {}
        for name in names:
            c, idx = self._index(name, b)
            if self.f.DataType(idx) == minnow.FloatGroup or self.f.DataType(idx) in (minnow.Float32Group, minnow.Float64Group):
                raise TypeError("Column '%s' does not hold integers" % name)
            out[name] = self.f.Data(idx)
        return out

    def FloatBlock(self, b, names):                                                          # :296-323
        out = {}
        for name in names:
            c, idx = self._index(name, b)
            t = self.f.DataType(idx)
            if t not in (minnow.FloatGroup, minnow.Float32Group):
                raise TypeError("Column '%s' does not hold float32" % name)
            # float32(math.Pow(10, float64(x))) of a Log column (:315-319) runs in the decode kernel
            out[name] = self.f.Data(idx, log10=bool(self.Columns[c]["Log"] != 0))
        return out

    def Ints(self, names):                                                                   # :228-240
        parts = [self.IntBlock(b, names) for b in range(self.Blocks)]
        return {n: np.concatenate([p[n] for p in parts]) if parts else np.zeros(0, np.int64) for n in names}

    def Floats(self, names):                                                                 # :242-254
        parts = [self.FloatBlock(b, names) for b in range(self.Blocks)]
        return {n: np.concatenate([p[n] for p in parts]) if parts else np.zeros(0, np.float32) for n in names}

    def Close(self):
        self.f.Close()


def Open(fname, ctx, jitter=None):
    return Reader(fname, ctx, jitter)
