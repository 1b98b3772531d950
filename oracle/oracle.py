"""ctypes front-end for the CPU oracle (oracle/minnow_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(minnow_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libminnow_oracle.so")

INT64, INT32, INT16, INT8, UINT64, UINT32, UINT16, UINT8, FLOAT64, FLOAT32, INT_GROUP, FLOAT_GROUP = range(12)
FIXED_DTYPES = [np.int64, np.int32, np.int16, np.int8, np.uint64, np.uint32,
                np.uint16, np.uint8, np.float64, np.float32]


def build(force=False):
    src = os.path.join(_HERE, "minnow_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _declare(_lib)
    return _lib


_p = C.c_void_p
_i64 = C.c_int64
_u64 = C.c_uint64
_f32 = C.c_float
_f64 = C.c_double
_int = C.c_int


def _declare(L):
    sig = {
        "orc_go_log10": (_f64, [_f64]),
        "orc_go_log2": (_f64, [_f64]),
        "orc_precision_needed": (_i64, [_u64]),
        "orc_array_bytes": (_i64, [_i64, _i64]),
        "orc_pack": (None, [_int, _p, _i64, _p]),
        "orc_unpack": (None, [_int, _p, _i64, _p]),
        "orc_bits": (_i64, [_p, _i64]),
        "orc_bound": (None, [_p, _i64, _i64, _i64]),
        "orc_periodic_min": (_i64, [_p, _i64, _i64]),
        "orc_quantize": (None, [_p, _i64, _f32, _f32, _i64, _p]),
        "orc_float_group_pixels": (_i64, [_f32, _f32, _f32]),
        "orc_minh_process_float": (None, [_p, _i64, C.c_int32, _f32, _f32]),
        "orc_jitter_hash32": (C.c_uint32, [_u64, _u64, _u64]),
        "orc_int_block_encode": (_i64, [_p, _i64, _p, _p, _p]),
        "orc_int_block_decode": (None, [_p, _i64, _i64, _i64, _p]),
        "orc_float_block_encode": (_i64, [_p, _i64, _f32, _f32, _i64, _int, _p, _p, _p]),
        "orc_float_block_decode": (None, [_p, _i64, _i64, _i64, _f32, _f32, _i64, _int, _int, _u64, _u64, _p, _p]),
        "orc_writer_create": (_p, []),
        "orc_writer_header": (_i64, [_p, _p, _i64]),
        "orc_writer_fixed_size_group": (None, [_p, _i64, _i64]),
        "orc_writer_int_group": (None, [_p, _i64]),
        "orc_writer_float_group": (None, [_p, _i64, _f32, _f32, _f32]),
        "orc_writer_data": (_i64, [_p, _p]),
        "orc_writer_close": (_i64, [_p, C.POINTER(_p)]),
        "orc_free": (None, [_p]),
        "orc_reader_open": (_p, [_p, _i64]),
        "orc_reader_close": (None, [_p]),
        "orc_reader_groups": (_i64, [_p]),
        "orc_reader_headers": (_i64, [_p]),
        "orc_reader_blocks": (_i64, [_p]),
        "orc_reader_header_size": (_i64, [_p, _i64]),
        "orc_reader_data_type": (_i64, [_p, _i64]),
        "orc_reader_data_len": (_i64, [_p, _i64]),
        "orc_reader_header": (_int, [_p, _i64, _p, _i64]),
        "orc_reader_block_min": (_i64, [_p, _i64]),
        "orc_reader_block_bits": (_i64, [_p, _i64]),
        "orc_reader_block_file_offset": (_i64, [_p, _i64]),
        "orc_reader_float_params": (_int, [_p, _i64, C.POINTER(_f32), C.POINTER(_f32), C.POINTER(_i64), C.POINTER(_int)]),
        "orc_reader_data": (_int, [_p, _i64, _p, _int, _u64, _p]),
        "orc_get_sub_cell": (None, [_p, _p, _p, _p, _i64, _i64, _i64]),
        "orc_set_sub_cell": (None, [_p, _p, _p, _p, _i64, _i64, _i64]),
        "orc_minp_limits": (None, [_p, _i64, _int, _f64, _p, _p]),
        "orc_minp_write": (_i64, [_p, _p, _i64, _p, _f64, _int, _p, _i64, C.POINTER(_p)]),
        "orc_minp_read": (_i64, [_p, _i64, _p, _p, C.POINTER(_f64), C.POINTER(_int), _p, _i64, _int, _u64]),
        "orc_minp_ids": (None, [_i64, _p, _p]),
        "orc_minh_create": (_p, []),
        "orc_minh_header": (None, [_p, _p, _i64, _p, _i64, _p, _i64]),
        "orc_minh_geometry": (None, [_p, _f32, _f32, _i64]),
        "orc_minh_block": (_int, [_p, _p, _i64, _i64]),
        "orc_minh_close": (_i64, [_p, C.POINTER(_p)]),
        "orc_bench_minp_encode": (_i64, [_p, _i64, _i64, _p, _p, _p, _p, _p, _p, _p, _i64, _int]),
        "orc_bench_minp_decode": (None, [_p, _i64, _i64, _i64, _p, _p, _p, _p, _p, _int, _f32, _int, _u64, _p, _int]),
        "orc_bench_group_encode": (_i64, [_int, _p, _i64, _i64, _f32, _f32, _i64, _int, _int, _p, _p, _p, _p, _i64, _int]),
        "orc_bench_group_decode": (None, [_int, _p, _i64, _i64, _i64, _p, _p, _p, _f32, _f32, _i64, _int, _int, _u64, _p, _int]),
        "orc_bnd_region": (_int, [_f32, _f32, _i64, _f32, _i64, _f32]),
        "orc_bnd_idx_reg": (None, [_f32, _f32, _i64, _f32, _p, _p, _p]),
        "orc_bnd_host_cells": (_int, [_i64, _p, _p, _p]),
        "orc_boundary_bin": (_i64, [_p, _p, _p, _i64, _f32, _f32, _i64, _f32, _p, _p, _p]),
        "orc_max_threads": (_int, []),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args


def _ptr(a):
    return a.ctypes.data_as(_p) if a is not None else None


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


# ---- package bit -----------------------------------------------------------
def precision_needed(mx):
    return int(lib().orc_precision_needed(int(mx) & 0xFFFFFFFFFFFFFFFF))


def array_bytes(bits, length):
    return int(lib().orc_array_bytes(bits, length))


def pack(bits, x):
    x = _c(x, np.uint64)
    out = np.zeros(array_bytes(bits, len(x)), dtype=np.uint8)
    lib().orc_pack(bits, _ptr(x), len(x), _ptr(out))
    return out


def unpack(bits, data, n):
    data = _c(data, np.uint8)
    out = np.zeros(n, dtype=np.uint64)
    lib().orc_unpack(bits, _ptr(data), n, _ptr(out))
    return out


def bits_of(x):
    x = _c(x, np.uint64)
    return int(lib().orc_bits(_ptr(x), len(x)))


# ---- go/group.go helpers ----------------------------------------------------
def periodic_min(x, pixels):
    x = _c(x, np.int64)
    return int(lib().orc_periodic_min(_ptr(x), len(x), pixels))


def bound(x, mn, pixels):
    x = _c(x, np.int64).copy()
    lib().orc_bound(_ptr(x), len(x), mn, pixels)
    return x


def quantize(x, low, high, pixels):
    x = _c(x, np.float32)
    q = np.zeros(len(x), dtype=np.int64)
    lib().orc_quantize(_ptr(x), len(x), low, high, pixels, _ptr(q))
    return q


def float_group_pixels(lo, hi, dx):
    return int(lib().orc_float_group_pixels(lo, hi, dx))


def minh_process_float(x, is_log, low, high):
    x = _c(x, np.float32).copy()
    lib().orc_minh_process_float(_ptr(x), len(x), int(is_log), low, high)
    return x


def jitter_hash32(seed, block, i):
    return int(lib().orc_jitter_hash32(seed, block, i))


def int_block_encode(x):
    """-> (min, bits, packed bytes)"""
    x = _c(x, np.int64)
    out = np.zeros(max(8 * len(x), 1), dtype=np.uint8)
    mn, bt = _i64(), _i64()
    nb = lib().orc_int_block_encode(_ptr(x), len(x), C.byref(mn), C.byref(bt), _ptr(out))
    return mn.value, bt.value, out[:nb if bt.value else 0].copy()


def int_block_decode(data, n, mn, bits):
    data = _c(data, np.uint8)
    out = np.zeros(n, dtype=np.int64)
    lib().orc_int_block_decode(_ptr(data), n, mn, bits, _ptr(out))
    return out


def float_block_encode(x, low, high, pixels, periodic=1):
    x = _c(x, np.float32)
    out = np.zeros(max(8 * len(x), 1), dtype=np.uint8)
    mn, bt = _i64(), _i64()
    nb = lib().orc_float_block_encode(_ptr(x), len(x), low, high, pixels, periodic,
                                      C.byref(mn), C.byref(bt), _ptr(out))
    return mn.value, bt.value, out[:nb if bt.value else 0].copy()


def float_block_decode(data, n, mn, bits, low, high, pixels, periodic=1,
                       jitter_mode=0, seed=0, block_id=0, u=None):
    data = _c(data, np.uint8)
    out = np.zeros(n, dtype=np.float32)
    if u is not None:
        u = _c(u, np.float64)
    lib().orc_float_block_decode(_ptr(data), n, mn, bits, low, high, pixels, periodic,
                                 jitter_mode, seed, block_id, _ptr(u), _ptr(out))
    return out


# ---- container --------------------------------------------------------------
class Writer:
    """Mirror of minnow.Writer (go/writer.go) writing to memory."""

    def __init__(self):
        self.h = lib().orc_writer_create()

    def header(self, data):
        b = data.tobytes() if isinstance(data, np.ndarray) else bytes(data)
        return int(lib().orc_writer_header(self.h, b, len(b)))

    def fixed_size_group(self, gt, N):
        self._dtype, self._N = FIXED_DTYPES[gt], N
        lib().orc_writer_fixed_size_group(self.h, gt, N)

    def int_group(self, N):
        self._dtype, self._N = np.int64, N
        lib().orc_writer_int_group(self.h, N)

    def float_group(self, N, lim, dx):
        self._dtype, self._N = np.float32, N
        lib().orc_writer_float_group(self.h, N, lim[0], lim[1], dx)

    def data(self, x):
        x = _c(x, self._dtype)
        assert len(x) == self._N
        return int(lib().orc_writer_data(self.h, _ptr(x)))

    def close(self):
        p = _p()
        n = lib().orc_writer_close(self.h, C.byref(p))
        out = C.string_at(p, n)
        lib().orc_free(p)
        self.h = None
        return out


class Reader:
    """Mirror of minnow.Reader (go/reader.go) over a bytes image."""

    def __init__(self, image):
        self._img = np.frombuffer(image, dtype=np.uint8).copy()
        self.h = lib().orc_reader_open(_ptr(self._img), len(self._img))
        if not self.h:
            raise ValueError("not a minnow file")
        L = lib()
        self.groups = L.orc_reader_groups(self.h)
        self.headers = L.orc_reader_headers(self.h)
        self.blocks = L.orc_reader_blocks(self.h)

    def header_size(self, i):
        return int(lib().orc_reader_header_size(self.h, i))

    def header(self, i):
        n = self.header_size(i)
        buf = np.zeros(max(n, 1), dtype=np.uint8)
        if lib().orc_reader_header(self.h, i, _ptr(buf), n):
            raise ValueError("header read failed")
        return buf[:n].tobytes()

    def data_type(self, b):
        return int(lib().orc_reader_data_type(self.h, b))

    def data_len(self, b):
        return int(lib().orc_reader_data_len(self.h, b))

    def block_min(self, b):
        return int(lib().orc_reader_block_min(self.h, b))

    def block_bits(self, b):
        return int(lib().orc_reader_block_bits(self.h, b))

    def block_file_offset(self, b):
        return int(lib().orc_reader_block_file_offset(self.h, b))

    def float_params(self, b):
        lo, hi, px, per = _f32(), _f32(), _i64(), _int()
        if lib().orc_reader_float_params(self.h, b, C.byref(lo), C.byref(hi), C.byref(px), C.byref(per)):
            raise ValueError("not a float group")
        return lo.value, hi.value, px.value, per.value

    def data(self, b, jitter_mode=0, seed=0, u=None):
        gt = self.data_type(b)
        dt = np.int64 if gt == INT_GROUP else np.float32 if gt == FLOAT_GROUP else FIXED_DTYPES[gt]
        out = np.zeros(self.data_len(b), dtype=dt)
        if u is not None:
            u = _c(u, np.float64)
        if lib().orc_reader_data(self.h, b, _ptr(out), jitter_mode, seed, _ptr(u)):
            raise ValueError("block read failed")
        return out

    def close(self):
        lib().orc_reader_close(self.h)
        self.h = None


# ---- minp -------------------------------------------------------------------
MINP_HEADER = np.dtype([("Z", "<f8"), ("Scale", "<f8"), ("OmegaM", "<f8"), ("OmegaL", "<f8"),
                        ("H100", "<f8"), ("L", "<f8"), ("Epsilon", "<f8"),
                        ("NSide", "<i8"), ("NTotal", "<i8"), ("UniformMp", "<f8")])
MINP_CELL = np.dtype([("FileIndex", "<i8"), ("FileCells", "<i8"), ("SubCells", "<i8")])


def minp_write(hd, raw_hd, cell, dx, periodic, vec):
    hd = np.asarray(hd, dtype=MINP_HEADER).reshape(1)
    cell = np.asarray(cell, dtype=MINP_CELL).reshape(1)
    vec = _c(vec, np.float32).reshape(-1, 3)
    raw = bytes(raw_hd)
    p = _p()
    n = lib().orc_minp_write(_ptr(hd), raw, len(raw), _ptr(cell), float(dx), int(periodic),
                             _ptr(vec), len(vec), C.byref(p))
    if n < 0:
        raise ValueError("orc_minp_write failed: %d" % n)
    out = C.string_at(p, n)
    lib().orc_free(p)
    return out


def minp_read(image, jitter_mode=0, seed=0):
    img = np.frombuffer(image, dtype=np.uint8).copy()
    hd = np.zeros(1, dtype=MINP_HEADER)
    cell = np.zeros(1, dtype=MINP_CELL)
    dx, per = _f64(), _int()
    rc = lib().orc_minp_read(_ptr(img), len(img), _ptr(hd), _ptr(cell), C.byref(dx), C.byref(per),
                             None, 0, jitter_mode, seed)
    if rc < 0:
        raise ValueError("orc_minp_read failed: %d" % rc)
    nfile = int(hd["NSide"][0] // cell["FileCells"][0])
    out = np.zeros((nfile ** 3, 3), dtype=np.float32)
    rc = lib().orc_minp_read(_ptr(img), len(img), _ptr(hd), _ptr(cell), C.byref(dx), C.byref(per),
                             _ptr(out), len(out), jitter_mode, seed)
    if rc < 0:
        raise ValueError("orc_minp_read failed: %d" % rc)
    return hd[0], cell[0], dx.value, bool(per.value), out


def minp_ids(nside, cell):
    cell = np.asarray(cell, dtype=MINP_CELL).reshape(1)
    nfile = nside // int(cell["FileCells"][0])
    out = np.zeros(nfile ** 3, dtype=np.int64)
    lib().orc_minp_ids(nside, _ptr(cell), _ptr(out))
    return out


def minp_limits(vec, periodic, L):
    vec = _c(vec, np.float32).reshape(-1, 3)
    mn = np.zeros(3, np.float32)
    mx = np.zeros(3, np.float32)
    lib().orc_minp_limits(_ptr(vec), len(vec), int(periodic), float(L), _ptr(mn), _ptr(mx))
    return mn, mx


def get_sub_cell(vec, sc, sub_cells, n_sub):
    vec = _c(vec, np.float32).reshape(-1, 3)
    sb = [np.zeros(n_sub ** 3, np.float32) for _ in range(3)]
    lib().orc_get_sub_cell(_ptr(vec), _ptr(sb[0]), _ptr(sb[1]), _ptr(sb[2]), sc, sub_cells, n_sub)
    return sb


# ---- minh -------------------------------------------------------------------
MINH_COLUMN = np.dtype([("Type", "<i8"), ("Log", "<i4"), ("Low", "<f4"), ("High", "<f4"),
                        ("Dx", "<f4"), ("Buffer", "S232")])
assert MINH_COLUMN.itemsize == 256


def minh_columns(cols):
    """cols: list of (type, log, low, high, dx)."""
    a = np.zeros(len(cols), dtype=MINH_COLUMN)
    for i, c in enumerate(cols):
        c = tuple(c) + (0,) * (5 - len(c))
        a["Type"][i], a["Log"][i], a["Low"][i], a["High"][i], a["Dx"][i] = c
    return a


class MinhWriter:
    """Mirror of minh.Writer (go/minh/minh.go:41-156) writing to memory."""

    def __init__(self):
        self.h = lib().orc_minh_create()

    def header(self, names, text, cols):
        self.cols = minh_columns(cols) if not isinstance(cols, np.ndarray) else cols
        jn = "$".join(names).encode("ascii")
        tx = text.encode("ascii")
        lib().orc_minh_header(self.h, jn, len(jn), tx, len(tx), _ptr(self.cols), len(self.cols))

    def geometry(self, L, boundary, cells):
        lib().orc_minh_geometry(self.h, L, boundary, cells)

    def block(self, cols):
        arrs = []
        for c, x in zip(self.cols, cols):
            t = int(c["Type"])
            dt = np.int64 if t == INT_GROUP else np.float32 if t == FLOAT_GROUP else FIXED_DTYPES[t]
            arrs.append(_c(x, dt))
        N = len(arrs[0])
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        if lib().orc_minh_block(self.h, ptrs, len(arrs), N):
            raise ValueError("column count mismatch")

    def close(self):
        p = _p()
        n = lib().orc_minh_close(self.h, C.byref(p))
        out = C.string_at(p, n)
        lib().orc_free(p)
        self.h = None
        return out


# ---- timed legs (bench.py only) ---------------------------------------------
def bench_minp_encode(vec, nfile, sub_cells, low, high, pixels, threads=0):
    vec = _c(vec, np.float32)
    nsub3 = (nfile // sub_cells) ** 3
    nb = 3 * sub_cells ** 3
    stride = 8 * nsub3
    mins = np.zeros(nb, np.int64)
    bits = np.zeros(nb, np.int64)
    nbytes = np.zeros(nb, np.int64)
    out = np.zeros(nb * stride, np.uint8)
    low, high, pixels = _c(low, np.float32), _c(high, np.float32), _c(pixels, np.int64)
    total = lib().orc_bench_minp_encode(_ptr(vec), nfile, sub_cells, _ptr(low), _ptr(high), _ptr(pixels),
                                        _ptr(mins), _ptr(bits), _ptr(nbytes), _ptr(out), stride, threads)
    return mins, bits, nbytes, out, stride, int(total)


def bench_minp_decode(packed, stride, nfile, sub_cells, low, high, pixels, mins, bits,
                      periodic, Lbox, jitter_mode=1, seed=0, threads=0):
    out = np.zeros((nfile ** 3, 3), np.float32)
    low, high, pixels = _c(low, np.float32), _c(high, np.float32), _c(pixels, np.int64)
    lib().orc_bench_minp_decode(_ptr(packed), stride, nfile, sub_cells, _ptr(low), _ptr(high), _ptr(pixels),
                                _ptr(mins), _ptr(bits), int(periodic), Lbox, jitter_mode, seed,
                                _ptr(out), threads)
    return out


def bench_group_encode(x, n, nblocks, desc=None, threads=0):
    """nblocks blocks of one group; desc = None for an IntGroup (x int64) or (low, high, pixels, is_log, clamp).
    -> mins, bits, nbytes, packed (block b at b * stride), stride, total"""
    kind = 0 if desc is None else 1
    x = _c(x, np.int64 if kind == 0 else np.float32)
    low, high, pixels, is_log, clamp = desc if desc is not None else (0.0, 0.0, 0, 0, 0)
    stride = 8 * n + 8
    mins, bits, nbytes = (np.zeros(nblocks, np.int64) for _ in range(3))
    out = np.zeros(nblocks * stride, np.uint8)
    total = lib().orc_bench_group_encode(kind, _ptr(x), n, nblocks, low, high, pixels, int(is_log), int(clamp),
                                         _ptr(mins), _ptr(bits), _ptr(nbytes), _ptr(out), stride, threads)
    return mins, bits, nbytes, out, stride, int(total)


def bench_group_decode(packed, stride, n, mins, bits, desc=None, sel=None, jitter_mode=1, seed=0, threads=0):
    kind = 0 if desc is None else 1
    low, high, pixels, is_log, _ = desc if desc is not None else (0.0, 0.0, 0, 0, 0)
    s = _c(sel, np.int64) if sel is not None else None
    nsel = len(s) if s is not None else len(mins)
    out = np.zeros((nsel, n), np.int64 if kind == 0 else np.float32)
    lib().orc_bench_group_decode(kind, _ptr(packed), stride, n, nsel, _ptr(s) if s is not None else None, _ptr(mins), _ptr(bits),
                                 low, high, pixels, int(is_log), jitter_mode, seed, _ptr(out), threads)
    return out


# ---- minh BoundaryWriter.Coordinates (go/minh/boundary.go:39-180) -----------------------------
def bnd_region(l, boundary, cells, scaled, ix, x):
    return int(lib().orc_bnd_region(l, boundary, cells, scaled, ix, x))


def bnd_idx_reg(l, boundary, cells, scaled, vec):
    v = _c(vec, np.float32)
    idx, reg = np.zeros(3, np.int64), np.zeros(3, np.int32)
    lib().orc_bnd_idx_reg(l, boundary, cells, scaled, _ptr(v), _ptr(idx), _ptr(reg))
    return idx.tolist(), reg.tolist()


def bnd_host_cells(cells, idx, reg):
    i, r, out = _c(idx, np.int64), _c(reg, np.int32), np.zeros(8, np.int64)
    n = lib().orc_bnd_host_cells(cells, _ptr(i), _ptr(r), _ptr(out))
    return out[:n].tolist()


def boundary_bin(x, y, z, l, boundary, cells, scaled=-1.0, want_index=True):
    """-> sizes [cells^3], idx (points of every cell in insertion order, cell after cell), flags"""
    x, y, z = _c(x, np.float32), _c(y, np.float32), _c(z, np.float32)
    sizes = np.zeros(cells ** 3, np.int64)
    total = lib().orc_boundary_bin(_ptr(x), _ptr(y), _ptr(z), len(x), l, boundary, cells, scaled, _ptr(sizes), None, None)
    if not want_index:
        return sizes, None, None
    idx, flags = np.zeros(total, np.int64), np.zeros(total, np.int8)
    lib().orc_boundary_bin(_ptr(x), _ptr(y), _ptr(z), len(x), l, boundary, cells, scaled, _ptr(sizes), _ptr(idx), _ptr(flags))
    return sizes, idx, flags


def max_threads():
    return int(lib().orc_max_threads())
