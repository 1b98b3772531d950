"""Host-side mirror of the reference's particle-snapshot format, package `minp`
(go/minp/minp.go): same names, argument meaning and on-disk bytes.  Writer.Vectors is one
call into the C ABI (mnw_minp_encode_vectors: limits, pixels, sub-cell gather, encode of the
three axis groups); Reader.Vectors is one call of mnw_decode_vec3_subcells (decode, periodic
wrap, sub-cell scatter)."""
import struct

import numpy as np

from . import minnow
from .capi import FloatDesc, Jitter, JITTER_CENTER, array_bytes

Magic = 0xbadf00d         # go/minp/minp.go:10-15
Version = 0
basicFileType = 2         # iota in the third spec of the const block

Header = np.dtype([("Z", "<f8"), ("Scale", "<f8"), ("OmegaM", "<f8"), ("OmegaL", "<f8"), ("H100", "<f8"),
                   ("L", "<f8"), ("Epsilon", "<f8"), ("NSide", "<i8"), ("NTotal", "<i8"),
                   ("UniformMp", "<f8")])                                                   # :24-30
Cell = np.dtype([("FileIndex", "<i8"), ("FileCells", "<i8"), ("SubCells", "<i8")])         # :32-34


def NFile(cell, nside):                                                                      # Cell.NFile, :36-42
    fc = int(cell["FileCells"])
    if nside < 0 or fc <= 0 or nside % fc:
        raise ValueError("NSide = %d not a valid combination with FileCells = %d" % (nside, fc))
    return nside // fc


class Writer:
    def __init__(self, fname, ctx):                                                          # Create, :62-67
        self.ctx = ctx
        self.f = minnow.Create(fname, ctx)
        self.f.Header(struct.pack("<qqq", Magic, Version, basicFileType))

    def Header(self, hd, raw_hd, c, dx, periodic):                                           # :69-84
        self.hd = np.asarray(hd, Header).reshape(1)[0]
        self.c = np.asarray(c, Cell).reshape(1)[0]
        self.f.Header(self.hd.tobytes())
        self.f.Header(bytes(raw_hd))
        self.f.Header(self.c.tobytes())
        self.f.Header(struct.pack("<d", dx))
        self.f.Header(struct.pack("<B", 1 if periodic else 0))
        self.periodic, self.dx = bool(periodic), np.float32(dx)

    def Vectors(self, vec):                                                                  # :86-119
        vec = np.ascontiguousarray(vec, np.float32).reshape(-1, 3)
        nfile = NFile(self.c, int(self.hd["NSide"]))
        sub = int(self.c["SubCells"])
        if nfile ** 3 != len(vec):
            raise ValueError("len(vec) = %d, but NSide = %d and FileCells = %d" %
                             (len(vec), self.hd["NSide"], self.c["FileCells"]))
        nsub3, sc3 = (nfile // sub) ** 3, sub ** 3
        descs, mins, bits, offs, streams = self.ctx.minp_encode_vectors(vec, nfile, sub, self.periodic,
                                                                        float(self.hd["L"]), self.dx)
        self.EncodedVectors(descs, mins, bits, streams)

    def EncodedVectors(self, descs, mins, bits, streams):
        """The three FloatGroups of one Vectors call from bytes that were encoded elsewhere -- by a device-resident
        mnw_minp_encode_vectors_dev batch, or by another rank of a sharded snapshot (its bytes read back from
        groupOffset + base_r): what Vectors records for them, go/minp/minp.go:112-118."""
        nfile = NFile(self.c, int(self.hd["NSide"]))
        sub = int(self.c["SubCells"])
        nsub3, sc3 = (nfile // sub) ** 3, sub ** 3
        w = self.f
        for k in range(3):                   # one FloatGroup per axis, sub^3 blocks each (:112-118)
            g = minnow._Group(minnow.FloatGroup, w.blocks, nsub3)
            g.low, g.high = np.float32(descs[k].low), np.float32(descs[k].high)
            g.pixels, g.periodic = int(descs[k].pixels), 1
            w._new_group(g)
            w.f.write(bytes(streams[k]) if isinstance(streams[k], (bytes, bytearray, memoryview)) else np.asarray(streams[k]).tobytes())
            sl = slice(k * sc3, (k + 1) * sc3)
            g.mins, g.bits = [int(m) for m in mins[sl]], [int(b) for b in bits[sl]]
            g.sizes = [array_bytes(b, nsub3) for b in g.bits]
            w.group_blocks[-1] += sc3
            w.blocks += sc3

    def Close(self):
        self.f.Close()


def Create(fname, ctx):
    return Writer(fname, ctx)


class Reader:
    def __init__(self, fname, ctx, jitter=None):                                             # Open, :142-173
        self.ctx = ctx
        self.jitter = jitter if jitter is not None else Jitter.make(JITTER_CENTER)
        self.f = minnow.Open(fname, ctx, self.jitter)
        magic, version, ftype = struct.unpack("<qqq", self.f.Header(0))
        if magic != Magic:
            raise ValueError("Not a minp file. Magic number is %d, not %d" % (magic, Magic))
        if version != Version:
            raise ValueError("File version = %d, but code version = %d." % (version, Version))
        if ftype != basicFileType:
            raise ValueError("File type = %d" % ftype)
        self.Header = self.f.Header(1, Header)[0]
        self.RawHeader = self.f.Header(2)
        self.c = self.f.Header(3, Cell)[0]
        self.Dx = struct.unpack("<d", self.f.Header(4))[0]
        self.Periodic = self.f.Header(5) != b"\0"
        self.FileIndex, self.FileCells = int(self.c["FileIndex"]), int(self.c["FileCells"])

    def Vectors(self):                                                                       # :175-207
        nfile = NFile(self.c, int(self.Header["NSide"]))
        sub = int(self.c["SubCells"])
        sc3, nsub3 = sub ** 3, (nfile // sub) ** 3
        if self.f.Blocks() != 3 * sc3:
            raise ValueError("Expected %d sub-cells, but got %d" % (3 * sub, self.f.Blocks()))
        descs, data3, mins, bits, offs = [], [], [], [], []
        for k in range(3):
            g = self.f.groups[k]
            descs.append(FloatDesc.make(g.low, g.high, g.pixels, g.periodic))
            self.f.f.seek(int(self.f.group_offsets[k]))
            data3.append(np.frombuffer(self.f.f.read(int(sum(g.sizes))), np.uint8))
            mins += g.mins; bits += g.bits
            ends = np.cumsum(np.asarray(g.sizes, np.int64))                                  # blockOffset, go/block_index.go:25-35
            offs += [0] + [int(e) for e in ends[:-1]]
        return self.ctx.decode_vec3_subcells(descs, data3, np.array(offs, np.int64), np.array(mins, np.int64),
                                             np.array(bits, np.int64), nfile, sub,
                                             wrap_L=float(np.float32(self.Header["L"])) if self.Periodic else 0.0,
                                             jitter=self.jitter)

    def IDs(self):                                                                           # :210-230 (no stored data)
        nfile, nside = NFile(self.c, int(self.Header["NSide"])), int(self.Header["NSide"])
        fc, fi = self.FileCells, self.FileIndex
        fx, fy, fz = fi % fc, (fi // fc) % fc, fi // (fc * fc)
        j = np.arange(nfile, dtype=np.int64)
        ix, iy, iz = j + fx * nfile, j + fy * nfile, j + fz * nfile
        return (ix[None, None, :] + iy[None, :, None] * nside + iz[:, None, None] * nside * nside).reshape(-1)

    def N(self):
        return self.f.Blocks() // 3

    def Close(self):
        self.f.Close()


def Open(fname, ctx, jitter=None):
    return Reader(fname, ctx, jitter)
