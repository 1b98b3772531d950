"""ctypes binding of include/minnow_cuda.h (libminnow_b200.so)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libminnow_b200.so")
_lib = None

JITTER_CENTER, JITTER_HASH, JITTER_STREAM = 0, 1, 2


class MinnowError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("minnow_b200 error %d: %s" % (code, msg))
        self.code = code


class FloatDesc(C.Structure):
    """mnw_float_desc"""
    _fields_ = [("low", C.c_float), ("high", C.c_float), ("pixels", C.c_int64), ("periodic", C.c_uint8),
                ("log10", C.c_uint8), ("clamp", C.c_uint8), ("reserved", C.c_uint8 * 5)]

    @classmethod
    def make(cls, low, high, pixels, periodic=1, log10=0, clamp=0):
        d = cls()
        d.low, d.high, d.pixels = float(low), float(high), int(pixels)
        d.periodic, d.log10, d.clamp = int(periodic), int(log10), int(clamp)
        return d


class Column(C.Structure):
    """mnw_column"""
    _fields_ = [("is_float", C.c_int32), ("reserved", C.c_int32), ("desc", FloatDesc)]


class Jitter(C.Structure):
    """mnw_jitter"""
    _fields_ = [("mode", C.c_int32), ("reserved", C.c_int32), ("seed", C.c_uint64), ("block_id0", C.c_uint64),
                ("u_stream", C.c_void_p)]

    @classmethod
    def make(cls, mode=JITTER_CENTER, seed=0, block_id0=0, u_stream=None):
        j = cls()
        j.mode, j.seed, j.block_id0 = int(mode), int(seed), int(block_id0)
        j.u_stream = u_stream
        return j


_p, _i64, _u64, _int, _f32 = C.c_void_p, C.c_int64, C.c_uint64, C.c_int, C.c_float
_FD = C.POINTER(FloatDesc)
_JT = C.POINTER(Jitter)

# name -> (restype, argtypes); every symbol declared in include/minnow_cuda.h
SIGNATURES = {
    "mnw_create": (_int, [_int, C.POINTER(_p)]),
    "mnw_destroy": (None, [_p]),
    "mnw_last_error": (C.c_char_p, [_p]),
    "mnw_sync": (_int, [_p]),
    "mnw_stream": (_p, [_p]),
    "mnw_version": (C.c_char_p, []),
    "mnw_launch_count": (_i64, [_p]),
    "mnw_selftest_fastdiv": (_int, [_p, _FD, C.c_uint32, _u64, C.POINTER(_u64), C.POINTER(_u64)]),
    "mnw_minp_encode_vectors": (_int, [_p, _p, _i64, _i64, _int, _f32, _f32, _FD, _p, _p, _p, _p, _i64, _p]),
    "mnw_encode_int_group_gather": (_int, [_p, _p, _i64, _p, _i64, _p, _p, _p, _p, _p, _i64, C.POINTER(_i64)]),
    "mnw_encode_columns": (_int, [_p, _i64, _p, _p, _i64, _p, _p, _p, _p, _i64]),
    "mnw_encode_columns_dev": (_int, [_p, _i64, _p, _p, _i64, _p, _p, _p, _p, _i64]),
    "mnw_decode_columns_dev": (_int, [_p, _i64, _p, _p, _i64, _p, _p, _p, _i64, _p, _p]),
    "mnw_encode_float_group_gather": (_int, [_p, _FD, _p, _i64, _p, _i64, _p, _p, _p, _p, _p, _i64, C.POINTER(_i64)]),
    "mnw_precision_needed": (_int, [_u64]),
    "mnw_array_bytes": (_i64, [_int, _i64]),
    "mnw_pack": (_int, [_p, _int, _p, _i64, _p]),
    "mnw_unpack": (_int, [_p, _int, _p, _i64, _p]),
    "mnw_bits": (_int, [_p, _p, _i64, C.POINTER(_int)]),
    "mnw_float_group_pixels": (_i64, [_f32, _f32, _f32]),
    "mnw_encode_int_group": (_int, [_p, _p, _i64, _i64, _p, _p, _p, _p, _p, _i64, C.POINTER(_i64)]),
    "mnw_encode_float_group": (_int, [_p, _FD, _p, _i64, _i64, _p, _p, _p, _p, _p, _i64, C.POINTER(_i64)]),
    "mnw_jitter_hash32": (C.c_uint32, [_u64, _u64, _u64]),
    "mnw_decode_int_blocks": (_int, [_p, _p, _i64, _p, _p, _p, _i64, _i64, _p, _p]),
    "mnw_decode_float_blocks": (_int, [_p, _FD, _p, _i64, _p, _p, _p, _i64, _i64, _p, _JT, _p]),
    "mnw_encode_vec3_subcells": (_int, [_p, _FD, _p, _i64, _i64, _p, _p, _p, _p, _i64, _p]),
    "mnw_decode_vec3_subcells": (_int, [_p, _FD, _p, _p, _p, _p, _p, _i64, _i64, _f32, _JT, _p]),
    "mnw_scan_offsets": (_int, [_p, _p, _i64, _i64, _p, C.POINTER(_i64)]),
    "mnw_encode_int_group_dev": (_int, [_p, _p, _i64, _i64, _p, _p, _p, _p, _i64, _p]),
    "mnw_encode_float_group_dev": (_int, [_p, _FD, _p, _i64, _i64, _p, _p, _p, _p, _i64, _p]),
    "mnw_decode_int_blocks_dev": (_int, [_p, _p, _i64, _p, _p, _p, _i64, _i64, _p, _p]),
    "mnw_decode_float_blocks_dev": (_int, [_p, _FD, _p, _i64, _p, _p, _p, _i64, _i64, _p, _JT, _p]),
    "mnw_encode_vec3_subcells_dev": (_int, [_p, _FD, _int, _p, _i64, _i64, _i64, _p, _p, _p, _p, _i64, _p]),
    "mnw_decode_vec3_subcells_dev": (_int, [_p, _FD, _int, _p, _i64, _p, _p, _p, _i64, _i64, _i64, _f32, _JT, _p]),
    "mnw_minp_encode_vectors_dev": (_int, [_p, _p, _i64, _i64, _i64, _int, _f32, _f32, _p, _p, _p, _p, _p, _i64, _p]),
    "mnw_minp_decode_vectors_dev": (_int, [_p, _p, _p, _i64, _p, _p, _p, _i64, _i64, _i64, _int, _f32, _JT, _p]),
    "mnw_text_parse_block": (_int, [_p, _p, _i64, C.c_char, C.c_char, _int, _p, _int, _p, C.POINTER(_i64), C.POINTER(_i64)]),
    "mnw_text_columns": (_int, [_p, _p, _p, _p]),
    "mnw_text_columns_dev": (_int, [_p, C.POINTER(_p), C.POINTER(_p)]),
    "mnw_regrid_insert": (_int, [_p, _p, _p, _i64, _i64, _i64, _p]),
    "mnw_regrid_insert_dev": (_int, [_p, _p, _p, _i64, _i64, _i64, _p]),
    "mnw_boundary_coordinates": (_int, [_p, _p, _p, _p, _i64, _f32, _f32, _i64, _p, C.POINTER(_i64)]),
    "mnw_boundary_index": (_int, [_p, _p, _p]),
    "mnw_boundary_encode_int_column": (_int, [_p, _p, _i64, _p, _p, _p, _p, _i64, C.POINTER(_i64)]),
    "mnw_boundary_encode_float_column": (_int, [_p, _FD, _p, _i64, _p, _p, _p, _p, _i64, C.POINTER(_i64)]),
    "mnw_boundary_encode_flags": (_int, [_p, _p, _p, _p, _p, _i64, C.POINTER(_i64)]),
    "mnw_comm_unique_id": (_int, [_p]),
    "mnw_comm_init": (_int, [_p, _p, _int, _int]),
    "mnw_comm_destroy": (_int, [_p]),
    "mnw_comm_size": (_int, [_p]),
    "mnw_comm_rank": (_int, [_p]),
    "mnw_allgather_sizes": (_int, [_p, _p, _i64, _p]),
    "mnw_sharded_offsets_dev": (_int, [_p, _p, _i64, _p, _p, _p]),
    "mnw_pipe_create": (_int, [_int, _int, C.POINTER(_p)]),
    "mnw_pipe_destroy": (None, [_p]),
    "mnw_pipe_last_error": (C.c_char_p, [_p]),
    "mnw_pipe_minp_encode_vectors": (_int, [_p, _p, _i64, _i64, _int, _f32, _f32, _FD, _p, _p, _p, _p, _i64, _p, C.POINTER(_i64)]),
    "mnw_pipe_minp_decode_vectors": (_int, [_p, _FD, _p, _p, _p, _p, _p, _i64, _i64, _f32, _JT, _p, C.POINTER(_i64)]),
    "mnw_pipe_poll": (_int, [_p]),
    "mnw_pipe_wait": (_int, [_p, _i64]),
    "mnw_pipe_drain": (_int, [_p]),
    "mnw_vec3_limits": (_int, [_p, _p, _i64, _i64, _p, _p]),
    "mnw_vec3_limits_dev": (_int, [_p, _p, _i64, _i64, _p, _p]),
    "mnw_scan_offsets_dev": (_int, [_p, _p, _i64, _i64, _p, _p]),
    "mnw_profile": (_int, [_p, _int]),
    "mnw_profile_summary": (_int, [_p, C.c_char_p, _i64]),
    "mnw_selftest_log10": (_int, [_p, C.c_uint32, _u64, C.POINTER(_u64)]),
    "mnw_selftest_pow10": (_int, [_p, C.c_uint32, _u64, C.POINTER(_u64)]),
    "mnw_pow10_f32": (_int, [_p, _p, _i64, _p]),
    "mnw_last_path": (_int, [_p]),
    "mnw_force_generic": (None, [_p, _int]),
}


def library_path():
    return _LIB_PATH


def load_library():
    """Loads libminnow_b200.so.  Fails loudly when it has not been built: there
    is no fallback implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(nvcc, sm_100a). minnow_b200 has no CPU fallback." % _LIB_PATH)
        lib = C.CDLL(_LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(lib, name)
            f.restype, f.argtypes = res, args
        _lib = lib
    return _lib


def precision_needed(mx):
    r = load_library().mnw_precision_needed(int(mx) & 0xFFFFFFFFFFFFFFFF)
    if r < 0:
        raise MinnowError(r, "bit.PrecisionNeeded is undefined for 2^64-1")
    return r


def array_bytes(bits, n):
    return int(load_library().mnw_array_bytes(bits, n))


def float_group_pixels(lo, hi, dx):
    return int(load_library().mnw_float_group_pixels(lo, hi, dx))


def jitter_hash32(seed, block, i):
    return int(load_library().mnw_jitter_hash32(seed, block, i))


def _go_parse_float(tok):
    """strconv.ParseFloat(tok, 64) for the fields mnw_text_parse_block leaves to the host: half-way cases, subnormals,
    more than 19 digits on a rounding boundary (Python's float() is correctly rounded too), hexadecimal floats."""
    t = tok.decode("ascii")
    if "_" in t:
        raise ValueError("strconv.ParseFloat: parsing %r: invalid syntax" % t)
    if t.lower().lstrip("+-").startswith("0x"):
        return float.fromhex(t)
    return float(t)


def _np(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(_p)
    if hasattr(a, "data_ptr"):      # torch tensor (device or pinned host)
        return _p(a.data_ptr())
    return _p(int(a))


class Pipe:
    """mnw_pipe: a ring of `depth` slots through which one host thread streams minp files (upload, kernels and download
    of different files overlap).  encode / decode return a ticket; the caller's arrays must stay alive and untouched
    until wait(ticket)."""

    def __init__(self, device=0, depth=4):
        self.lib = load_library()
        h = _p()
        rc = self.lib.mnw_pipe_create(device, depth, C.byref(h))
        if rc:
            raise MinnowError(rc, self.lib.mnw_last_error(None).decode())
        self.h, self.depth = h, depth

    def close(self):
        if getattr(self, "h", None):
            self.lib.mnw_pipe_destroy(self.h)
            self.h = None

    __del__ = close

    def _check(self, rc):
        if rc:
            raise MinnowError(rc, self.lib.mnw_pipe_last_error(self.h).decode())

    def encode(self, aos, nfile, subcells, periodic, L, dx, desc_out, mins, bits, offsets, out, out_axis_stride, out_len):
        t = _i64(-1)
        self._check(self.lib.mnw_pipe_minp_encode_vectors(self.h, _ptr(aos), nfile, subcells, int(bool(periodic)), float(L), float(dx),
                                                          desc_out, _ptr(mins), _ptr(bits), _ptr(offsets), _ptr(out), out_axis_stride,
                                                          _ptr(out_len), C.byref(t)))
        return t.value

    def decode(self, desc3, data_ptrs3, data_len, offsets, mins, bits, nfile, subcells, wrap_L, jitter, aos_out):
        t = _i64(-1)
        self._check(self.lib.mnw_pipe_minp_decode_vectors(self.h, desc3, data_ptrs3, _ptr(data_len), _ptr(offsets), _ptr(mins),
                                                          _ptr(bits), nfile, subcells, float(wrap_L), C.byref(jitter), _ptr(aos_out),
                                                          C.byref(t)))
        return t.value

    def poll(self):
        self._check(self.lib.mnw_pipe_poll(self.h))

    def wait(self, ticket):
        self._check(self.lib.mnw_pipe_wait(self.h, ticket))

    def drain(self):
        self._check(self.lib.mnw_pipe_drain(self.h))


class Context:
    """One mnw_ctx: a CUDA stream plus grow-only scratch.  Not thread-safe,
    like a minnow.Writer (go/writer.go) -- use one per thread."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = _p()
        rc = self.lib.mnw_create(device, C.byref(h))
        if rc:
            raise MinnowError(rc, self.lib.mnw_last_error(None).decode())
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.mnw_destroy(self.h)
            self.h = None

    __del__ = close

    def _check(self, rc):
        if rc:
            raise MinnowError(rc, self.lib.mnw_last_error(self.h).decode())

    def sync(self):
        self._check(self.lib.mnw_sync(self.h))

    @property
    def stream(self):
        return self.lib.mnw_stream(self.h)

    @property
    def launch_count(self):
        return int(self.lib.mnw_launch_count(self.h))

    @property
    def last_path(self):
        return int(self.lib.mnw_last_path(self.h))

    def profile(self, on=True):
        self._check(self.lib.mnw_profile(self.h, int(on)))

    def profile_summary(self):
        import json
        buf = C.create_string_buffer(1 << 16)
        self._check(self.lib.mnw_profile_summary(self.h, buf, len(buf)))
        return json.loads(buf.value.decode())

    # ---- multi-GPU (one context per GPU and rank) ---------------------------------------------
    @staticmethod
    def comm_unique_id():
        """rank 0: the 128-byte NCCL id to hand to the other ranks"""
        buf = C.create_string_buffer(128)
        lib = load_library()
        rc = lib.mnw_comm_unique_id(buf)
        if rc:
            raise MinnowError(rc, lib.mnw_last_error(None).decode())
        return buf.raw

    def comm_init(self, uid, nranks, rank):
        self._check(self.lib.mnw_comm_init(self.h, C.create_string_buffer(bytes(uid), 128), nranks, rank))

    def comm_destroy(self):
        self._check(self.lib.mnw_comm_destroy(self.h))

    @property
    def comm_size(self):
        return int(self.lib.mnw_comm_size(self.h))

    def allgather_sizes(self, local, count, out):
        self._check(self.lib.mnw_allgather_sizes(self.h, _ptr(local), count, _ptr(out)))

    def sharded_offsets_dev(self, local_nbytes, count, all_nbytes, all_offsets, total):
        self._check(self.lib.mnw_sharded_offsets_dev(self.h, _ptr(local_nbytes), count, _ptr(all_nbytes), _ptr(all_offsets), _ptr(total)))

    def scan_offsets_dev(self, nbytes, nblocks, base, offsets, total):
        self._check(self.lib.mnw_scan_offsets_dev(self.h, _ptr(nbytes), nblocks, base, _ptr(offsets), _ptr(total)))

    def selftest_fastdiv(self, desc, first_bits=0, count=1 << 32):
        """-> (mismatches, accepted) of the fast quantiser against the IEEE divide"""
        bad, acc = _u64(0), _u64(0)
        self._check(self.lib.mnw_selftest_fastdiv(self.h, C.byref(desc), first_bits, count, C.byref(bad), C.byref(acc)))
        return bad.value, acc.value

    def selftest_log10(self, first_bits=0, count=1 << 32):
        """-> mismatches of go_log10_f32 against the restated Go math.Log10 over float32 bit patterns"""
        bad = _u64(0)
        self._check(self.lib.mnw_selftest_log10(self.h, first_bits, count, C.byref(bad)))
        return bad.value

    def selftest_pow10(self, first_bits=0, count=1 << 32):
        bad = _u64(0)
        self._check(self.lib.mnw_selftest_pow10(self.h, first_bits, count, C.byref(bad)))
        return bad.value

    def pow10_f32(self, x):
        """float32(math.Pow(10, float64(x))), go/minh/minh.go:315-319"""
        x = _np(x, np.float32).reshape(-1)
        out = np.empty_like(x)
        self._check(self.lib.mnw_pow10_f32(self.h, _ptr(x), len(x), _ptr(out)))
        return out

    def force_generic(self, on=True):
        self.lib.mnw_force_generic(self.h, int(on))

    # ---- package bit ---------------------------------------------------------
    def pack(self, bits, x):
        """bit.BufferedArray / NewArray (go/bit/bit.go:84-142)"""
        x = _np(x, np.uint64)
        out = np.zeros(array_bytes(bits, len(x)) if 1 <= bits <= 64 else 0, np.uint8)
        self._check(self.lib.mnw_pack(self.h, bits, _ptr(x), len(x), _ptr(out)))
        return out

    def unpack(self, bits, data, n):
        """(*Array).Slice (go/bit/bit.go:29-82)"""
        data = _np(data, np.uint8)
        out = np.zeros(n, np.uint64)
        self._check(self.lib.mnw_unpack(self.h, bits, _ptr(data), n, _ptr(out)))
        return out

    def bits(self, x):
        """ArrayBuffer.Bits (go/bit/bit.go:151-159)"""
        x = _np(x, np.uint64)
        b = _int()
        self._check(self.lib.mnw_bits(self.h, _ptr(x), len(x), C.byref(b)))
        return b.value

    # ---- groups, host buffers --------------------------------------------------
    def _encode_group(self, fn, pre, x, esz, n, nblocks, starts):
        total = len(x)
        mins, bits, offs = (np.zeros(nblocks, np.int64) for _ in range(3))
        out = np.zeros(8 * total + 8, np.uint8)
        ln = _i64()
        st = _np(starts, np.int64) if starts is not None else None
        self._check(fn(self.h, *pre, _ptr(x), n, nblocks, _ptr(st), _ptr(mins), _ptr(bits), _ptr(offs),
                       _ptr(out), len(out), C.byref(ln)))
        return mins, bits, offs, out[:ln.value].copy()

    def encode_int_group(self, x, n=None, nblocks=None, starts=None):
        """nblocks x intGroup.writeData (go/group.go:242-255) -> (mins, bits, offsets, bytes)"""
        x = _np(x, np.int64).reshape(-1)
        if starts is not None:
            nblocks, n = len(starts) - 1, 0
        return self._encode_group(self.lib.mnw_encode_int_group, (), x, 8, n, nblocks, starts)

    def encode_float_group(self, desc, x, n=None, nblocks=None, starts=None):
        """nblocks x floatGroup.writeData (go/group.go:312-327)"""
        x = _np(x, np.float32).reshape(-1)
        if starts is not None:
            nblocks, n = len(starts) - 1, 0
        return self._encode_group(self.lib.mnw_encode_float_group, (C.byref(desc),), x, 4, n, nblocks, starts)

    def encode_group_gather(self, col, idx, starts, desc=None):
        """Blocks gathered from one column (BoundaryWriter.Column, go/minh/boundary.go:184-225):
        block b = col[idx[starts[b]:starts[b+1]]]; desc = None for an int64 column.
        -> (mins, bits, offsets, data)"""
        is_int = desc is None
        col = _np(col, np.int64 if is_int else np.float32)
        idx, starts = _np(idx, np.int64), _np(starts, np.int64)
        nb = len(starts) - 1
        mins, bits, offs = (np.zeros(nb, np.int64) for _ in range(3))
        out = np.zeros(8 * len(idx) + 64, np.uint8)
        ln = _i64(0)
        if is_int:
            rc = self.lib.mnw_encode_int_group_gather(self.h, _ptr(col), len(col), _ptr(idx), nb, _ptr(starts), _ptr(mins),
                                                      _ptr(bits), _ptr(offs), _ptr(out), len(out), C.byref(ln))
        else:
            rc = self.lib.mnw_encode_float_group_gather(self.h, C.byref(desc), _ptr(col), len(col), _ptr(idx), nb, _ptr(starts),
                                                        _ptr(mins), _ptr(bits), _ptr(offs), _ptr(out), len(out), C.byref(ln))
        self._check(rc)
        return mins, bits, offs, out[:ln.value].copy()

    def encode_columns(self, columns):
        """All quantised columns of one minh block in one call (minh.Writer.Block, go/minh/minh.go:99-139).
        columns: list of (x, desc) with desc = None for an IntGroup column (x int64) or a FloatDesc (x float32);
        all of one length.  -> mins, bits (int64 arrays) and the list of packed byte arrays, one per column."""
        nc = len(columns)
        arrs = [_np(x, np.float32 if d is not None else np.int64).reshape(-1) for x, d in columns]
        n = len(arrs[0]) if nc else 0
        if any(len(a) != n for a in arrs):
            raise ValueError("columns of one block must have one length")
        cols = (Column * max(nc, 1))()
        ptrs = (C.c_void_p * max(nc, 1))()
        for i, ((x, d), a) in enumerate(zip(columns, arrs)):
            cols[i].is_float = 0 if d is None else 1
            if d is not None:
                cols[i].desc = d
            ptrs[i] = a.ctypes.data if n else None
        stride = 8 * n + 16   # ArrayBytes(64, n): a NaN in a float column makes a block of more than 32 bits
        mins, bits, nbytes = (np.zeros(nc, np.int64) for _ in range(3))
        out = np.empty(max(nc * stride, 1), np.uint8)
        self._check(self.lib.mnw_encode_columns(self.h, nc, cols, ptrs, n, _ptr(mins), _ptr(bits), _ptr(nbytes), _ptr(out), stride))
        return mins, bits, [out[i * stride:i * stride + int(nbytes[i])].copy() for i in range(nc)]

    def encode_columns_dev(self, columns, n, mins, bits, nbytes, out, out_col_stride):
        """mnw_encode_columns_dev: columns = [(device tensor, FloatDesc or None)], outputs device tensors"""
        nc = len(columns)
        cols = (Column * max(nc, 1))()
        ptrs = (C.c_void_p * max(nc, 1))()
        for i, (x, d) in enumerate(columns):
            cols[i].is_float = 0 if d is None else 1
            if d is not None:
                cols[i].desc = d
            ptrs[i] = x.data_ptr() if hasattr(x, "data_ptr") else int(x)
        self._check(self.lib.mnw_encode_columns_dev(self.h, nc, cols, ptrs, n, _ptr(mins), _ptr(bits), _ptr(nbytes), _ptr(out), out_col_stride))

    def decode_columns_dev(self, descs, data, offsets, mins, bits, n, jitter, outs):
        """mnw_decode_columns_dev: descs = [FloatDesc or None per column], outs = [device tensor per column]"""
        nc = len(descs)
        cols = (Column * max(nc, 1))()
        ptrs = (C.c_void_p * max(nc, 1))()
        for i, d in enumerate(descs):
            cols[i].is_float = 0 if d is None else 1
            if d is not None:
                cols[i].desc = d
            ptrs[i] = outs[i].data_ptr() if hasattr(outs[i], "data_ptr") else int(outs[i])
        self._check(self.lib.mnw_decode_columns_dev(self.h, nc, cols, _ptr(data), int(data.numel()), _ptr(offsets), _ptr(mins), _ptr(bits), n,
                                                    C.byref(jitter) if jitter is not None else None, ptrs))

    # ---- text -> columns ------------------------------------------------------------------------
    def text_parse_block(self, buf, icols, fcols, sep=b" ", comment=b"#"):
        """text.Reader.Block for one block of bytes (go/text/text.go:181-200): -> (int64 [len(icols), rows],
        float32 [len(fcols), rows]); icols / fcols: column numbers in any order.  The few floats the device leaves to
        the host's parser (see the header) are converted here with Python's float()."""
        buf = bytes(buf)
        io_, fo_ = np.argsort(icols, kind="stable"), np.argsort(fcols, kind="stable")
        ic = np.ascontiguousarray(np.asarray(icols, np.int32)[io_]) if len(icols) else np.zeros(0, np.int32)
        fc = np.ascontiguousarray(np.asarray(fcols, np.int32)[fo_]) if len(fcols) else np.zeros(0, np.int32)
        rows, nfb = _i64(0), _i64(0)
        self._check(self.lib.mnw_text_parse_block(self.h, buf, len(buf), sep, comment, len(ic), _ptr(ic), len(fc), _ptr(fc),
                                                  C.byref(rows), C.byref(nfb)))
        iout, fout = np.zeros((len(ic), rows.value), np.int64), np.zeros((len(fc), rows.value), np.float32)
        fb = np.zeros((nfb.value, 3), np.int64)
        self._check(self.lib.mnw_text_columns(self.h, _ptr(iout), _ptr(fout), _ptr(fb)))
        for r, c, where in fb:
            off, ln = int(where) & ((1 << 40) - 1), int(where) >> 40
            try:
                fout[c, r] = np.float32(_go_parse_float(buf[off:off + ln]))
            except ValueError as exc:                     # the reference panics with strconv's error
                raise MinnowError(-4, "text: %s" % exc)
        inv_i, inv_f = np.empty_like(io_), np.empty_like(fo_)
        inv_i[io_], inv_f[fo_] = np.arange(len(io_)), np.arange(len(fo_))
        return iout[inv_i], fout[inv_f]

    # ---- Lagrangian re-gridding ---------------------------------------------------------------
    def regrid_insert(self, ids, vec, ncell, nside, grid_dev, dev=False):
        """vectorGrid.Insert over a batch (go/minp/snapshot/grid.go:118-137,206-211) into the device grid [ncell^3][nside^3][3]"""
        if dev:
            self._check(self.lib.mnw_regrid_insert_dev(self.h, _ptr(ids), _ptr(vec), ids.numel(), ncell, nside, _ptr(grid_dev)))
        else:
            ids, vec = _np(ids, np.int64).reshape(-1), _np(vec, np.float32).reshape(-1)
            self._check(self.lib.mnw_regrid_insert(self.h, _ptr(ids), _ptr(vec), len(ids), ncell, nside, _ptr(grid_dev)))

    # ---- minh BoundaryWriter ------------------------------------------------------------------
    def boundary_coordinates(self, x, y, z, L, boundary, cells):
        """BoundaryWriter.Coordinates (go/minh/boundary.go:39-51) -> sizes [cells^3]; the index lists stay on the device"""
        x, y, z = (_np(a, np.float32).reshape(-1) for a in (x, y, z))
        sizes, total = np.zeros(cells ** 3, np.int64), _i64(0)
        self._check(self.lib.mnw_boundary_coordinates(self.h, _ptr(x), _ptr(y), _ptr(z), len(x), float(L), float(boundary), cells,
                                                      _ptr(sizes), C.byref(total)))
        self._bnd_total, self._bnd_cells = total.value, cells
        return sizes

    def boundary_index(self):
        idx, flags = np.zeros(self._bnd_total, np.int64), np.zeros(self._bnd_total, np.int64)
        self._check(self.lib.mnw_boundary_index(self.h, _ptr(idx), _ptr(flags)))
        return idx, flags

    def _boundary_encode(self, fn, pre, col, ncol):
        nb = self._bnd_cells ** 3
        mins, bits, offs = (np.zeros(nb, np.int64) for _ in range(3))
        out = np.zeros(8 * self._bnd_total + 64, np.uint8)
        ln = _i64(0)
        args = pre + ((_ptr(col), ncol) if col is not None else ())
        self._check(fn(self.h, *args, _ptr(mins), _ptr(bits), _ptr(offs), _ptr(out), len(out), C.byref(ln)))
        return mins, bits, offs, out[:ln.value].copy()

    def boundary_encode_column(self, col, desc=None):
        """BoundaryWriter.Column for an IntGroup (desc None) or FloatGroup column -> (mins, bits, offsets, data), one block per cell"""
        if desc is None:
            col = _np(col, np.int64).reshape(-1)
            return self._boundary_encode(self.lib.mnw_boundary_encode_int_column, (), col, len(col))
        col = _np(col, np.float32).reshape(-1)
        return self._boundary_encode(self.lib.mnw_boundary_encode_float_column, (C.byref(desc),), col, len(col))

    def boundary_encode_flags(self):
        return self._boundary_encode(self.lib.mnw_boundary_encode_flags, (), None, 0)

    def decode_int_blocks(self, data, offsets, mins, bits, n, sel=None):
        """intGroup.readData per selected block (go/group.go:257-263)"""
        data = _np(data, np.uint8)
        offsets, mins, bits = _np(offsets, np.int64), _np(mins, np.int64), _np(bits, np.int64)
        s = _np(sel, np.int64) if sel is not None else None
        nsel = len(s) if s is not None else len(offsets)
        out = np.zeros((nsel, n), np.int64)
        self._check(self.lib.mnw_decode_int_blocks(self.h, _ptr(data), len(data), _ptr(offsets), _ptr(mins),
                                                   _ptr(bits), n, nsel, _ptr(s), _ptr(out)))
        return out

    def decode_float_blocks(self, desc, data, offsets, mins, bits, n, sel=None, jitter=None, u=None):
        """floatGroup.readData per selected block (go/group.go:299-310)"""
        data = _np(data, np.uint8)
        offsets, mins, bits = _np(offsets, np.int64), _np(mins, np.int64), _np(bits, np.int64)
        s = _np(sel, np.int64) if sel is not None else None
        nsel = len(s) if s is not None else len(offsets)
        out = np.zeros((nsel, n), np.float32)
        jit = jitter if jitter is not None else Jitter.make()
        if u is not None:
            u = _np(u, np.float64)
            jit = Jitter.make(JITTER_STREAM, u_stream=u.ctypes.data)
        self._check(self.lib.mnw_decode_float_blocks(self.h, C.byref(desc), _ptr(data), len(data), _ptr(offsets),
                                                     _ptr(mins), _ptr(bits), n, nsel, _ptr(s), C.byref(jit),
                                                     _ptr(out)))
        return out

    def scan_offsets(self, nbytes, base=0):
        """blockIndex.addBlock / blockOffset (go/block_index.go:16-35)"""
        nbytes = _np(nbytes, np.int64)
        out = np.zeros(len(nbytes), np.int64)
        tot = _i64()
        self._check(self.lib.mnw_scan_offsets(self.h, _ptr(nbytes), len(nbytes), base, _ptr(out), C.byref(tot)))
        return out, tot.value

    # ---- minp ---------------------------------------------------------------------
    def encode_vec3_subcells(self, descs, aos, nfile, subcells):
        """body of minp.Writer.Vectors (go/minp/minp.go:112-118).
        -> (mins, bits, offsets) each [3*subcells^3], [bytes_x, bytes_y, bytes_z]"""
        aos = _np(aos, np.float32).reshape(-1)
        assert len(aos) == 3 * nfile ** 3
        nb = 3 * subcells ** 3
        mins, bits, offs = (np.zeros(nb, np.int64) for _ in range(3))
        stride = 8 * nfile ** 3 + 8
        out = np.zeros(3 * stride, np.uint8)
        lens = np.zeros(3, np.int64)
        d3 = (FloatDesc * 3)(*descs)
        self._check(self.lib.mnw_encode_vec3_subcells(self.h, d3, _ptr(aos), nfile, subcells, _ptr(mins),
                                                      _ptr(bits), _ptr(offs), _ptr(out), stride, _ptr(lens)))
        return mins, bits, offs, [out[k * stride:k * stride + lens[k]].copy() for k in range(3)]

    def minp_encode_vectors(self, aos, nfile, subcells, periodic, L, dx):
        """minp.Writer.Vectors for one file (go/minp/minp.go:86-119), one upload.
        -> (descs[3], mins, bits, offsets, [bytes_x, bytes_y, bytes_z])"""
        aos = _np(aos, np.float32).reshape(-1)
        assert len(aos) == 3 * nfile ** 3
        nb = 3 * subcells ** 3
        mins, bits, offs = (np.zeros(nb, np.int64) for _ in range(3))
        stride = 8 * nfile ** 3 + 8
        out = np.zeros(3 * stride, np.uint8)
        lens = np.zeros(3, np.int64)
        d3 = (FloatDesc * 3)()
        self._check(self.lib.mnw_minp_encode_vectors(self.h, _ptr(aos), nfile, subcells, int(bool(periodic)), float(L), float(dx),
                                                     d3, _ptr(mins), _ptr(bits), _ptr(offs), _ptr(out), stride, _ptr(lens)))
        descs = [FloatDesc.make(d.low, d.high, d.pixels, d.periodic) for d in d3]
        return descs, mins, bits, offs, [out[k * stride:k * stride + lens[k]].copy() for k in range(3)]

    def decode_vec3_subcells(self, descs, data3, offsets, mins, bits, nfile, subcells, wrap_L=0.0, jitter=None):
        """body of minp.Reader.Vectors (go/minp/minp.go:191-206) -> [nfile^3, 3] float32"""
        data3 = [_np(d, np.uint8) for d in data3]
        ptrs = (C.c_void_p * 3)(*[d.ctypes.data for d in data3])
        lens = np.array([len(d) for d in data3], np.int64)
        offsets, mins, bits = _np(offsets, np.int64), _np(mins, np.int64), _np(bits, np.int64)
        out = np.zeros((nfile ** 3, 3), np.float32)
        d3 = (FloatDesc * 3)(*descs)
        jit = jitter if jitter is not None else Jitter.make()
        self._check(self.lib.mnw_decode_vec3_subcells(self.h, d3, ptrs, _ptr(lens), _ptr(offsets), _ptr(mins),
                                                      _ptr(bits), nfile, subcells, wrap_L, C.byref(jit), _ptr(out)))
        return out

    # ---- device-resident variants (torch tensors or raw device addresses) ------------
    def encode_float_group_dev(self, desc, x, n, nblocks, mins, bits, offsets, out, out_cap, out_len):
        self._check(self.lib.mnw_encode_float_group_dev(self.h, C.byref(desc), _ptr(x), n, nblocks, _ptr(mins),
                                                        _ptr(bits), _ptr(offsets), _ptr(out), out_cap, _ptr(out_len)))

    def encode_int_group_dev(self, x, n, nblocks, mins, bits, offsets, out, out_cap, out_len):
        self._check(self.lib.mnw_encode_int_group_dev(self.h, _ptr(x), n, nblocks, _ptr(mins), _ptr(bits),
                                                      _ptr(offsets), _ptr(out), out_cap, _ptr(out_len)))

    def decode_float_blocks_dev(self, desc, data, data_len, offsets, mins, bits, n, nsel, sel, jitter, out):
        self._check(self.lib.mnw_decode_float_blocks_dev(self.h, C.byref(desc), _ptr(data), data_len, _ptr(offsets),
                                                         _ptr(mins), _ptr(bits), n, nsel, _ptr(sel),
                                                         C.byref(jitter), _ptr(out)))

    def decode_int_blocks_dev(self, data, data_len, offsets, mins, bits, n, nsel, sel, out):
        self._check(self.lib.mnw_decode_int_blocks_dev(self.h, _ptr(data), data_len, _ptr(offsets), _ptr(mins),
                                                       _ptr(bits), n, nsel, _ptr(sel), _ptr(out)))

    def minp_encode_vectors_dev(self, aos, nfile, subcells, nfiles, periodic, L, dx, desc_dev, mins, bits, offsets, out,
                                out_axis_stride, out_len):
        """minp.Writer.Vectors for nfiles resident cubes, no host round trip; desc_dev: device buffer of 24 * 3 * nfiles bytes"""
        self._check(self.lib.mnw_minp_encode_vectors_dev(self.h, _ptr(aos), nfile, subcells, nfiles, int(bool(periodic)), float(L),
                                                         float(dx), _ptr(desc_dev), _ptr(mins), _ptr(bits), _ptr(offsets), _ptr(out),
                                                         out_axis_stride, _ptr(out_len)))

    def minp_decode_vectors_dev(self, desc_dev, data, data_axis_stride, offsets, mins, bits, nfile, subcells, nfiles, periodic,
                                L, jitter, aos_out):
        self._check(self.lib.mnw_minp_decode_vectors_dev(self.h, _ptr(desc_dev), _ptr(data), data_axis_stride, _ptr(offsets),
                                                         _ptr(mins), _ptr(bits), nfile, subcells, nfiles, int(bool(periodic)), float(L),
                                                         C.byref(jitter), _ptr(aos_out)))

    def vec3_limits(self, aos, nfiles=1, dev=False):
        """minp.Writer.Vectors limits of non-periodic fields (go/minp/minp.go:92-95) -> lo, hi [nfiles, 3]"""
        lo, hi = np.zeros((nfiles, 3), np.float32), np.zeros((nfiles, 3), np.float32)
        if dev:
            npart = aos.numel() // (3 * nfiles)
            self._check(self.lib.mnw_vec3_limits_dev(self.h, _ptr(aos), npart, nfiles, _ptr(lo), _ptr(hi)))
        else:
            aos = _np(aos, np.float32).reshape(-1)
            npart = len(aos) // (3 * nfiles)
            self._check(self.lib.mnw_vec3_limits(self.h, _ptr(aos), npart, nfiles, _ptr(lo), _ptr(hi)))
        return lo, hi

    def encode_vec3_subcells_dev(self, descs, aos, nfile, subcells, nfiles, mins, bits, offsets, out,
                                 out_axis_stride, out_len):
        """descs: 3 FloatDesc (shared) or 3*nfiles (per file)"""
        d3 = (FloatDesc * len(descs))(*descs)
        self._check(self.lib.mnw_encode_vec3_subcells_dev(self.h, d3, int(len(descs) != 3), _ptr(aos), nfile, subcells, nfiles,
                                                          _ptr(mins), _ptr(bits), _ptr(offsets), _ptr(out),
                                                          out_axis_stride, _ptr(out_len)))

    def decode_vec3_subcells_dev(self, descs, data, data_axis_stride, offsets, mins, bits, nfile, subcells,
                                 nfiles, wrap_L, jitter, aos_out):
        d3 = (FloatDesc * len(descs))(*descs)
        self._check(self.lib.mnw_decode_vec3_subcells_dev(self.h, d3, int(len(descs) != 3), _ptr(data), data_axis_stride, _ptr(offsets),
                                                          _ptr(mins), _ptr(bits), nfile, subcells, nfiles, wrap_L,
                                                          C.byref(jitter), _ptr(aos_out)))
