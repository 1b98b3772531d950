"""GPU parity tests proper: the CUDA path, called through the C ABI
(include/minnow_cuda.h via minnow_b200.capi), against the CPU oracle on the
same inputs.  Integer, byte and index results must be bit-exact; decoded floats
are bit-exact given the jitter stream.  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest

import minnow_b200 as mb
from helpers import oracle_float_group, oracle_int_group, uniform_starts

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = mb.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module", params=["auto", "generic"])
def anyctx(request):
    """Both device paths: the one the library picks, and the generic one forced."""
    c = mb.Context(0)
    c.force_generic(request.param == "generic")
    yield c
    c.close()


# ---- package bit ------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 7, 123, 4096, 4097, 100000])
def test_pack_unpack_all_widths(ctx, orc, n):
    rng = np.random.default_rng(n)
    data = rng.integers(0, 2 ** 63, n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, n, dtype=np.uint64)
    for bits in range(1, 65):
        mask = np.uint64((1 << bits) - 1) if bits < 64 else np.uint64(2 ** 64 - 1)
        want = orc.pack(bits, data)
        got = ctx.pack(bits, data)
        assert got.tobytes() == want.tobytes(), (n, bits)
        back = ctx.unpack(bits, want, n)
        assert np.array_equal(back, data & mask), (n, bits)   # go/bit/bit_test.go:9-31


def test_bits_and_array_buffer_kat(ctx, orc):
    for n, bits in ((10, 4), (5, 3), (1, 0), (20, 5)):          # go/bit/bit_test.go:33-69
        x = np.arange(n, dtype=np.uint64)
        assert ctx.bits(x) == bits == orc.bits_of(x)
    rng = np.random.default_rng(3)
    for k in (0, 1, 13, 31, 32, 33, 47, 50, 63):
        x = rng.integers(0, 2 ** k, 5000, dtype=np.uint64) if k else np.zeros(5000, np.uint64)
        assert ctx.bits(x) == orc.bits_of(x)
    with pytest.raises(mb.MinnowError):
        ctx.pack(65, np.zeros(3, np.uint64))                    # go/bit/bit.go:85-87 panics


# ---- IntGroup ------------------------------------------------------------------------
def test_int_group_reference_vectors(ctx, orc):
    # go/minnow_test.go:242-268 TestBitIntRecord
    for blocks in ([[100, 101, 102, 104]], [[1024, 1024, 1024], [0, 1023, 500]], [[-1000000, -500000]]):
        x = np.array(blocks, np.int64)
        nb, n = x.shape
        mins, bits, offs, data = ctx.encode_int_group(x, n, nb)
        omins, obits, ooffs, odata = oracle_int_group(orc, x.reshape(-1), uniform_starts(n, nb))
        assert mins.tolist() == omins.tolist() and bits.tolist() == obits.tolist() and offs.tolist() == ooffs.tolist()
        assert data.tobytes() == odata.tobytes()
        back = ctx.decode_int_blocks(data, offs, mins, bits, n)
        assert np.array_equal(back, x)
    mins, bits, offs, data = ctx.encode_int_group(np.array([[1024] * 3, [0, 1023, 500]], np.int64), 3, 2)
    assert data.tobytes() == bytes([0x00, 0xfc, 0x4f, 0x1f]) and bits.tolist() == [0, 10]


@pytest.mark.parametrize("n,nb", [(1, 1), (1, 9), (5, 3), (31, 4), (32, 4), (33, 4), (4096, 3), (4097, 2),
                                  (16384, 2), (16385, 2), (65536, 5), (100003, 2)])
def test_int_group_random(anyctx, orc, n, nb):
    rng = np.random.default_rng(n * 31 + nb)
    blocks = []
    for b in range(nb):
        width = int(rng.integers(0, 64))
        base = int(rng.integers(-2 ** 62, 2 ** 62))
        lo = rng.integers(0, 2 ** width, n, dtype=np.uint64).astype(np.int64) if width else np.zeros(n, np.int64)
        with np.errstate(over="ignore"):
            blocks.append((np.int64(base) + lo).astype(np.int64))
    x = np.concatenate(blocks)
    mins, bits, offs, data = anyctx.encode_int_group(x, n, nb)
    omins, obits, ooffs, odata = oracle_int_group(orc, x, uniform_starts(n, nb))
    assert np.array_equal(mins, omins) and np.array_equal(bits, obits) and np.array_equal(offs, ooffs)
    assert data.tobytes() == odata.tobytes()
    back = anyctx.decode_int_blocks(data, offs, mins, bits, n)
    # Go's own round trip is lossy where PrecisionNeeded under-counts (>= 2^49); compare with the oracle's decode
    for b in range(nb):
        end = offs[b + 1] if b + 1 < nb else len(data)
        assert np.array_equal(back[b], orc.int_block_decode(data[offs[b]:end], n, int(mins[b]), int(bits[b])))
        if bits[b] < 48:
            assert np.array_equal(back[b], blocks[b])
    sel = np.array([nb - 1, 0, nb // 2], np.int64)
    part = anyctx.decode_int_blocks(data, offs, mins, bits, n, sel=sel)
    assert np.array_equal(part, back[sel])


def test_int_group_ragged_and_empty(ctx, orc):
    rng = np.random.default_rng(11)
    lens = [0, 1, 40, 0, 4096, 5000, 3, 0]
    starts = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    x = rng.integers(-10 ** 6, 10 ** 6, starts[-1]).astype(np.int64)
    mins, bits, offs, data = ctx.encode_int_group(x, starts=starts)
    omins, obits, ooffs, odata = oracle_int_group(orc, x, starts)
    assert np.array_equal(mins, omins) and np.array_equal(bits, obits) and np.array_equal(offs, ooffs)
    assert data.tobytes() == odata.tobytes()
    m, b, o, d = ctx.encode_int_group(np.zeros(0, np.int64), 0, 0)
    assert len(m) == 0 and len(d) == 0


# ---- FloatGroup ------------------------------------------------------------------------
def test_float_group_reference_vectors(ctx, orc):
    # go/minnow_test.go:270-310 TestQFloatRecord (limit -50..100; dx 1 -> 150 px; dx 10 -> 15 px)
    for px, blocks, want in ((150, [[-50, 0, 50, 49], [25, 25, 25, 25]], bytes([0x00, 0x19, 0x79, 0x0c])),
                             (15, [[-50, 0, 50, 49, 0], [1, 2, 3, 4, 5], [0, 20, 0, 20, 0]],
                              bytes([0x50, 0x9a, 0x05, 0x88, 0x00]))):
        x = np.array(blocks, np.float32)
        nb, n = x.shape
        d = mb.FloatDesc.make(-50, 100, px)
        mins, bits, offs, data = ctx.encode_float_group(d, x, n, nb)
        assert data.tobytes() == want                                    # SURVEY 8c.4 hexdumps
        om, ob, oo, od = oracle_float_group(orc, x.reshape(-1), uniform_starts(n, nb), -50, 100, px)
        assert np.array_equal(mins, om) and np.array_equal(bits, ob) and np.array_equal(offs, oo)
        dx = 150.0 / px
        for mode in (mb.JITTER_CENTER, mb.JITTER_HASH):
            back = ctx.decode_float_blocks(d, data, offs, mins, bits, n, jitter=mb.Jitter.make(mode, 5))
            assert np.all(np.abs(back - x) <= dx)                        # float32sEq, go/minnow_test.go:328


def _float_cases(rng, n, nb, low, high, pixels):
    span = high - low
    dx = span / pixels
    out = []
    for b in range(nb):
        kind = b % 6
        if kind == 0:      # anywhere in the box: arc too wide, periodicMin = 0
            x = rng.uniform(low, high, n)
        elif kind == 1:    # tight cluster in the middle
            c = rng.uniform(low + 0.3 * span, low + 0.6 * span); x = c + rng.normal(0, 0.01 * span, n)
        elif kind == 2:    # cluster straddling the periodic edge (wraps)
            x = low + np.mod(rng.normal(0, 0.02 * span, n), span)
        elif kind == 3:    # all equal: 0 bits
            x = np.full(n, rng.uniform(low, high))
        elif kind == 4:    # values ON pixel edges (division rounding matters)
            x = low + rng.integers(0, pixels, n) * np.float32(dx)
        else:              # slightly less than half the box wide
            x = low + np.mod(rng.uniform(0.7, 1.19, n) * span, span)
        out.append(np.clip(x, low, np.nextafter(np.float32(high), np.float32(-np.inf))).astype(np.float32))
    return np.concatenate(out)


@pytest.mark.parametrize("n,nb,low,high,dx", [
    (1, 6, 0.0, 125.0, 0.001), (37, 12, 0.0, 125.0, 0.001), (4096, 12, 0.0, 125.0, 0.001),
    (65536, 6, 0.0, 125.0, 0.001), (100003, 6, -50.0, 100.0, 0.37), (262144, 6, 0.0, 1000.0, 0.005),
    (5000, 12, 0.0, 250.0, 1.0), (5000, 6, 10.0, 14.0, 0.01), (3000, 6, -3.0, 3.0, 1e-6)])
def test_float_group_random(anyctx, orc, n, nb, low, high, dx):
    rng = np.random.default_rng(int(n * 7 + nb))
    pixels = mb.float_group_pixels(low, high, dx)
    assert pixels == orc.float_group_pixels(low, high, dx)
    x = _float_cases(rng, n, nb, low, high, pixels)
    d = mb.FloatDesc.make(low, high, pixels)
    mins, bits, offs, data = anyctx.encode_float_group(d, x, n, nb)
    om, ob, oo, od = oracle_float_group(orc, x, uniform_starts(n, nb), low, high, pixels)
    assert np.array_equal(mins, om), (mins, om)
    assert np.array_equal(bits, ob) and np.array_equal(offs, oo)
    assert data.tobytes() == od.tobytes()
    # decode: bit-exact for each jitter policy, and within dx of the input
    for mode in (mb.JITTER_CENTER, mb.JITTER_HASH):
        back = anyctx.decode_float_blocks(d, data, offs, mins, bits, n, jitter=mb.Jitter.make(mode, 99, 1000))
        for b in range(nb):
            end = offs[b + 1] if b + 1 < nb else len(data)
            want = orc.float_block_decode(data[offs[b]:end], n, int(mins[b]), int(bits[b]), low, high, pixels, 1,
                                          mode, 99, 1000 + b)
            assert back[b].tobytes() == want.tobytes(), (mode, b)
        span = high - low
        err = np.abs(back.reshape(-1) - x)
        err = np.minimum(err, span - err)                       # periodic
        assert np.all(err <= np.float32(span) / np.float32(pixels) * 1.0001 + 1e-6 * span)
    u = rng.uniform(0, 1, (nb, n))
    back = anyctx.decode_float_blocks(d, data, offs, mins, bits, n, u=u)
    for b in (0, nb - 1):
        end = offs[b + 1] if b + 1 < nb else len(data)
        want = orc.float_block_decode(data[offs[b]:end], n, int(mins[b]), int(bits[b]), low, high, pixels, 1, 2, 0, 0, u[b])
        assert back[b].tobytes() == want.tobytes()


def test_float_group_out_of_range_values_take_exact_path(anyctx, orc):
    """Values at or beyond the limits give pixel indices outside [0, pixels): the
    order-independent arc statistic does not apply and the exact sequential
    periodicMin must be reproduced (go/group.go:384-409)."""
    rng = np.random.default_rng(77)
    low, high, pixels = 0.0, 100.0, 1000
    n, blocks = 777, []
    for b in range(24):
        c = rng.uniform(0, 100)
        x = np.mod(c + rng.normal(0, 3.0, n), 100.0)
        k = rng.integers(1, 6)
        idx = rng.integers(0, n, k)
        x[idx] = rng.choice([100.0, 100.05, -0.05, -0.2, 100.3, 199.0, -99.0, 250.0], k)
        if b % 4 == 0:
            x[0] = 100.0                                        # first element is the out-of-range one
        blocks.append(x.astype(np.float32))
    x = np.concatenate(blocks)
    d = mb.FloatDesc.make(low, high, pixels)
    mins, bits, offs, data = anyctx.encode_float_group(d, x, n, len(blocks))
    om, ob, oo, od = oracle_float_group(orc, x, uniform_starts(n, len(blocks)), low, high, pixels)
    assert np.array_equal(mins, om) and np.array_equal(bits, ob) and np.array_equal(offs, oo)
    assert data.tobytes() == od.tobytes()


def test_float_group_nonperiodic_log_clamp(anyctx, orc):
    rng = np.random.default_rng(8)
    n, nb = 3001, 4
    # minh Log column: masses log-uniform 1e10..1e15, Low 10 High 15 Dx 0.01
    x = (10.0 ** rng.uniform(9.5, 15.5, n * nb)).astype(np.float32)
    pixels = mb.float_group_pixels(10, 15, 0.01)
    d = mb.FloatDesc.make(10, 15, pixels, periodic=1, log10=1, clamp=1)
    mins, bits, offs, data = anyctx.encode_float_group(d, x, n, nb)
    om, ob, oo, od = oracle_float_group(orc, x, uniform_starts(n, nb), 10, 15, pixels, 1, 1, 1)
    assert np.array_equal(mins, om) and np.array_equal(bits, ob) and data.tobytes() == od.tobytes()
    # non-periodic group (the format supports it: floatGroup.periodic, go/group.go:273)
    y = rng.uniform(-5, 5, n * nb).astype(np.float32)
    d2 = mb.FloatDesc.make(-5, 5, 2000, periodic=0)
    mins, bits, offs, data = anyctx.encode_float_group(d2, y, n, nb)
    om, ob, oo, od = oracle_float_group(orc, y, uniform_starts(n, nb), -5, 5, 2000, 0)
    assert np.array_equal(mins, om) and np.array_equal(bits, ob) and data.tobytes() == od.tobytes()


def test_float_group_ragged(ctx, orc):
    rng = np.random.default_rng(12)
    lens = [5, 0, 3, 8000, 1, 4097]
    starts = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    x = rng.uniform(0, 125, starts[-1]).astype(np.float32)
    d = mb.FloatDesc.make(0, 125, 125000)
    mins, bits, offs, data = ctx.encode_float_group(d, x, starts=starts)
    om, ob, oo, od = oracle_float_group(orc, x, starts, 0, 125, 125000)
    assert np.array_equal(mins, om) and np.array_equal(bits, ob) and np.array_equal(offs, oo)
    assert data.tobytes() == od.tobytes()


# ---- block index ------------------------------------------------------------------------
def test_scan_offsets(ctx):
    rng = np.random.default_rng(4)
    for n in (0, 1, 5, 1024, 1025, 70000):
        sizes = rng.integers(0, 10 ** 6, n).astype(np.int64)
        offs, total = ctx.scan_offsets(sizes, base=48)
        want = 48 + np.concatenate([[0], np.cumsum(sizes)[:-1]]) if n else np.zeros(0)
        assert np.array_equal(offs, want) and total == int(sizes.sum())


# ---- minp ------------------------------------------------------------------------------------
def _make_vectors(rng, nfile, L, kind):
    g = np.stack(np.meshgrid(np.arange(nfile), np.arange(nfile), np.arange(nfile), indexing="ij"), -1)
    g = g.transpose(2, 1, 0, 3).reshape(-1, 3)                  # x fastest
    if kind == "grid":                                          # go/minp/minp_test.go:128-151 makeVectors
        v = (g * np.float32(L / nfile)).astype(np.float32)
    else:
        v = np.mod(g * (L / nfile) + rng.normal(0, 0.02 * L, g.shape), L).astype(np.float32)
        v[v >= L] = 0
    return np.ascontiguousarray(v, np.float32)


@pytest.mark.parametrize("nside,subcells", [(1, 1), (2, 1), (8, 1), (10, 1), (2, 2), (10, 2), (10, 5), (32, 2), (64, 4)])
@pytest.mark.parametrize("periodic", [False, True])
def test_minp_vectors(anyctx, orc, nside, subcells, periodic):
    # go/minp/minp_test.go:7-73 TestVecReaderWriter shapes (+ two larger ones)
    rng = np.random.default_rng(nside * 10 + subcells)
    L, dx = 100.0, 0.1
    vec = _make_vectors(rng, nside, L, "grid" if nside <= 10 else "noisy")
    hd = np.zeros(1, orc.MINP_HEADER); hd["L"] = L; hd["NSide"] = nside; hd["NTotal"] = nside ** 3
    hd["Z"], hd["Scale"], hd["OmegaM"], hd["OmegaL"], hd["H100"], hd["Epsilon"], hd["UniformMp"] = 1, .5, .27, .73, .7, 2, 1e10
    cell = np.array([(0, 1, subcells)], orc.MINP_CELL)
    img = orc.minp_write(hd, bytes(range(130)), cell, dx, periodic, vec)
    r = orc.Reader(img)
    mn, mx = orc.minp_limits(vec, periodic, L)
    sc3 = subcells ** 3
    descs = []
    for k in range(3):
        px = mb.float_group_pixels(float(mn[k]), float(mx[k]), np.float32(dx))
        assert r.float_params(k * sc3)[2] == px
        descs.append(mb.FloatDesc.make(float(mn[k]), float(mx[k]), px))
    mins, bits, offs, streams = anyctx.encode_vec3_subcells(descs, vec, nside, subcells)
    for b in range(3 * sc3):
        assert mins[b] == r.block_min(b) and bits[b] == r.block_bits(b), b
    for k in range(3):
        start = r.block_file_offset(k * sc3)
        for sc in range(sc3):
            assert start + offs[k * sc3 + sc] == r.block_file_offset(k * sc3 + sc)
        assert img[start:start + len(streams[k])] == streams[k].tobytes()
    r.close()
    for mode in (mb.JITTER_CENTER, mb.JITTER_HASH):
        got = anyctx.decode_vec3_subcells(descs, streams, offs, mins, bits, nside, subcells,
                                          wrap_L=L if periodic else 0.0, jitter=mb.Jitter.make(mode, 31))
        _, _, _, _, want = orc.minp_read(img, mode, 31)
        if all(d.pixels > 0 for d in descs):
            assert got.tobytes() == want.tobytes()
        else:   # pixels == 0 (one particle, non-periodic): dx = 0/0; Go and CUDA both give NaN, payloads differ
            assert np.array_equal(np.isnan(got), np.isnan(want))
            assert np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)])
        if all(d.pixels > 0 for d in descs) and not periodic:    # vectorsEq, go/minp/minp_test.go:116-126
            assert np.all(np.abs(got - vec) <= dx * 1.001)


# ---- gathered blocks (BoundaryWriter.Column, go/minh/boundary.go:184-225) ------------------------
def test_group_gather_matches_encode_of_gathered_copy(ctx, orc):
    rng = np.random.default_rng(44)
    ncol = 20000
    lens = [0, 5, 4097, 1, 300, 9000]
    starts = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    idx = rng.integers(0, ncol, starts[-1]).astype(np.int64)          # ghost layers repeat halos: duplicates allowed
    ids = rng.integers(-10 ** 15, 10 ** 15, ncol).astype(np.int64)
    pos = rng.uniform(0, 125, ncol).astype(np.float32)
    mins, bits, offs, data = ctx.encode_group_gather(ids, idx, starts)
    om, ob, oo, od = oracle_int_group(orc, ids[idx], starts)
    assert np.array_equal(mins, om) and np.array_equal(bits, ob) and np.array_equal(offs, oo) and data.tobytes() == od.tobytes()
    d = mb.FloatDesc.make(0, 125, 125000)
    mins, bits, offs, data = ctx.encode_group_gather(pos, idx, starts, d)
    om, ob, oo, od = oracle_float_group(orc, pos[idx], starts, 0, 125, 125000)
    assert np.array_equal(mins, om) and np.array_equal(bits, ob) and np.array_equal(offs, oo) and data.tobytes() == od.tobytes()
    with pytest.raises(mb.MinnowError):
        ctx.encode_group_gather(ids, np.array([0, ncol], np.int64), np.array([0, 2], np.int64))


def test_encode_columns_matches_per_column_calls(ctx):
    """mnw_encode_columns (one call per minh.Writer.Block) = the per-column group encodes, byte for byte: mixed int64 /
    float32 columns, a log10 + clamp column, a constant column (0 bits), an int column wider than 32 bits"""
    rng = np.random.default_rng(77)
    for n in (1, 4097, 70001):
        px = mb.float_group_pixels(0.0, 125.0, 0.001)
        cols = [
            (rng.integers(10 ** 9, 10 ** 9 + 10 ** 6, n).astype(np.int64), None),
            ((rng.random(n) * 125.0).astype(np.float32), mb.FloatDesc.make(0.0, 125.0, px)),
            (np.power(10.0, 10.0 + 5.0 * rng.random(n)).astype(np.float32),
             mb.FloatDesc.make(10.0, 15.0, mb.float_group_pixels(10.0, 15.0, 0.01), 1, 1, 1)),
            (np.full(n, 7, np.int64), None),
            (rng.integers(-2 ** 60, 2 ** 60, n).astype(np.int64), None),
            ((rng.random(n) * 300.0 - 100.0).astype(np.float32), mb.FloatDesc.make(0.0, 125.0, px, 1, 0, 1)),   # clamped
        ]
        mins, bits, packed = ctx.encode_columns(cols)
        for c, (x, d) in enumerate(cols):
            if d is None:
                m, b, o, data = ctx.encode_int_group(x, n, 1)
            else:
                m, b, o, data = ctx.encode_float_group(d, x, n, 1)
            assert (int(mins[c]), int(bits[c])) == (int(m[0]), int(b[0])), (n, c)
            assert packed[c].tobytes() == data.tobytes(), (n, c)
    assert ctx.encode_columns([])[2] == []
