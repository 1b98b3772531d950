"""GPU parity of the fused single-read minp kernels (k_fused_vec3 / k_decode_vec3),
called through the C ABI, against the CPU oracle: block (min, bits), byte offsets,
packed bytes and decoded float32 values must all be bit-identical.  Covers every
sub-cell size the fused path takes (16^3, 32^3, 64^3 with an 8-CTA cluster), narrow and
wide periodic arcs, blocks of 0 bits, blocks wider than 16 bits (the list-driven second
pass), pixel indices equal to `pixels`, and inputs it must hand to the exact generic path
(NaN / out-of-range values).  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest

import minnow_b200 as mb

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = mb.Context(0)
    yield c
    c.close()


def lagrangian(rng, nfile, L, sigma, wrap=True):
    j = np.arange(nfile, dtype=np.float64) * (L / nfile)
    g = np.stack(np.meshgrid(j, j, j, indexing="ij"), -1).transpose(2, 1, 0, 3).reshape(-1, 3)   # x fastest
    v = g + rng.normal(0.0, sigma, g.shape)
    if wrap:
        v = np.mod(v, L)
    v = v.astype(np.float32)
    if wrap:
        v[v >= L] = 0
    return np.ascontiguousarray(v)


def check(ctx, orc, vec, nfile, subcells, low, high, pixels, wrap_L, expect_path=1):
    descs = [mb.FloatDesc.make(low[k], high[k], pixels[k]) for k in range(3)]
    mins, bits, offs, streams = ctx.encode_vec3_subcells(descs, vec, nfile, subcells)
    assert ctx.last_path == expect_path
    omins, obits, onbytes, packed, stride, total = orc.bench_minp_encode(vec, nfile, subcells, low, high, pixels, threads=0)
    assert np.array_equal(mins, omins), np.flatnonzero(mins != omins)[:8]
    assert np.array_equal(bits, obits), np.flatnonzero(bits != obits)[:8]
    sc3 = subcells ** 3
    for k in range(3):
        sl = slice(k * sc3, (k + 1) * sc3)
        want_offs = np.concatenate([[0], np.cumsum(onbytes[sl])[:-1]])
        assert np.array_equal(offs[sl], want_offs)
        want = b"".join(packed[t * stride:t * stride + onbytes[t]].tobytes() for t in range(k * sc3, (k + 1) * sc3))
        got = streams[k].tobytes()
        assert len(got) == len(want)
        if got != want:
            a, b = np.frombuffer(got, np.uint8), np.frombuffer(want, np.uint8)
            first = int(np.flatnonzero(a != b)[0])
            raise AssertionError("axis %d: packed bytes differ first at byte %d of %d" % (k, first, len(want)))
    for mode in (mb.JITTER_CENTER, mb.JITTER_HASH):
        got = ctx.decode_vec3_subcells(descs, streams, offs, mins, bits, nfile, subcells, wrap_L=wrap_L,
                                       jitter=mb.Jitter.make(mode, 11))
        want = orc.bench_minp_decode(packed, stride, nfile, subcells, low, high, pixels, omins, obits,
                                     wrap_L > 0, wrap_L, mode, 11, threads=0)
        assert got.tobytes() == want.tobytes(), mode
    return mins, bits


@pytest.mark.parametrize("nfile,subcells", [(16, 1), (32, 2), (64, 4), (32, 1), (64, 2), (64, 1), (128, 2)])
def test_fused_positions(ctx, orc, nfile, subcells):
    """periodic positions: narrow arcs, sub-cells that straddle the box edge (periodicMin != 0)"""
    rng = np.random.default_rng(nfile + subcells)
    L, dx = 100.0, 0.01
    vec = lagrangian(rng, nfile, L, 0.8)
    px = mb.float_group_pixels(0.0, L, dx)
    mins, bits = check(ctx, orc, vec, nfile, subcells, [0.0] * 3, [L] * 3, [px] * 3, L)
    if subcells >= 4:   # sub-cells narrower than half the box: re-origined arcs
        assert (mins != 0).any() and bits.max() < 16


@pytest.mark.parametrize("nfile,subcells", [(32, 2), (64, 2), (64, 1)])
def test_fused_velocities(ctx, orc, nfile, subcells):
    """non-periodic field through the periodic code path (go/writer.go:74): limits from the data,
    the maximum lands on pixel index == pixels, every arc is wide"""
    rng = np.random.default_rng(100 + nfile)
    vec = (300.0 * rng.standard_normal((nfile ** 3, 3))).astype(np.float32)
    lo, hi = orc.minp_limits(vec, False, 0.0)
    px = [mb.float_group_pixels(float(lo[k]), float(hi[k]), 1.0) for k in range(3)]
    glo, ghi = ctx.vec3_limits(vec)
    assert np.array_equal(glo[0], lo) and np.array_equal(ghi[0], hi)
    check(ctx, orc, vec, nfile, subcells, lo.tolist(), hi.tolist(), px, 0.0)


@pytest.mark.parametrize("dx,sigma", [(0.0001, 3.0), (0.00001, 0.5), (1.0, 0.5), (40.0, 0.5)])
def test_fused_bit_widths(ctx, orc, dx, sigma):
    """> 16 bits (second pass from global memory), P > 65536 with narrow arcs, tiny P"""
    rng = np.random.default_rng(7)
    L, nfile, subcells = 100.0, 64, 2
    vec = lagrangian(rng, nfile, L, sigma)
    px = mb.float_group_pixels(0.0, L, dx)
    mins, bits = check(ctx, orc, vec, nfile, subcells, [0.0] * 3, [L] * 3, [px] * 3, L)
    if dx == 0.0001:
        assert bits.max() > 16
    if dx == 40.0:
        assert px == 3


def test_fused_constant_and_edges(ctx, orc):
    """blocks of 0 bits, values at 0, just below `high` (pixel index == pixels), mixed per axis"""
    rng = np.random.default_rng(9)
    L, nfile, subcells, dx = 1000.0, 32, 2, 0.005
    px = mb.float_group_pixels(0.0, L, dx)
    vec = lagrangian(rng, nfile, L, 2.0)
    vec[:, 1] = 123.456                                   # y: every block is one value -> 0 bits
    top = np.nextafter(np.float32(L), np.float32(0))      # quantises to index == pixels
    vec[5::97, 2] = top
    vec[11::89, 2] = 0.0
    vec[0, 0] = 0.0
    mins, bits = check(ctx, orc, vec, nfile, subcells, [0.0] * 3, [L] * 3, [px] * 3, L)
    sc3 = subcells ** 3
    assert (bits[sc3:2 * sc3] == 0).all()
    # first element of a block exactly at index == pixels: the exact generic path takes over
    vec2 = vec.copy()
    vec2[0, 2] = top
    check(ctx, orc, vec2, nfile, subcells, [0.0] * 3, [L] * 3, [px] * 3, L)


@pytest.mark.parametrize("bad", [np.nan, np.inf, -3.0, 2500.0])
def test_fused_hands_bad_values_to_exact_path(ctx, orc, bad):
    """values outside [low, high] make periodicMin order dependent: the redo by the generic
    kernels must still match the reference's sequential scan bit for bit"""
    rng = np.random.default_rng(13)
    L, nfile, subcells, dx = 1000.0, 32, 2, 0.005
    px = mb.float_group_pixels(0.0, L, dx)
    vec = lagrangian(rng, nfile, L, 2.0)
    vec[777, 0] = bad
    vec[20000, 2] = bad
    descs = [mb.FloatDesc.make(0.0, L, px) for _ in range(3)]
    mins, bits, offs, streams = ctx.encode_vec3_subcells(descs, vec, nfile, subcells)
    omins, obits, onbytes, packed, stride, total = orc.bench_minp_encode(vec, nfile, subcells, [0.0] * 3, [L] * 3, [px] * 3)
    assert np.array_equal(mins, omins) and np.array_equal(bits, obits)
    sc3 = subcells ** 3
    for k in range(3):
        want = b"".join(packed[t * stride:t * stride + onbytes[t]].tobytes() for t in range(k * sc3, (k + 1) * sc3))
        assert streams[k].tobytes() == want


def test_fused_many_files_dev(ctx, orc):
    """device-resident entry point with several files, per-file descriptors and look-back chains"""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(21)
    nfile, subcells, nfiles = 32, 2, 5
    sc3, n3 = subcells ** 3, nfile ** 3
    vecs = [(200.0 * (f + 1) * rng.standard_normal((n3, 3))).astype(np.float32) for f in range(nfiles)]
    dev = torch.device("cuda", 0)
    aos = torch.from_numpy(np.stack(vecs)).to(dev)
    lo, hi = ctx.vec3_limits(aos, nfiles, dev=True)
    descs, px = [], []
    for f in range(nfiles):
        olo, ohi = orc.minp_limits(vecs[f], False, 0.0)
        assert np.array_equal(lo[f], olo) and np.array_equal(hi[f], ohi)
        for k in range(3):
            px.append(mb.float_group_pixels(float(lo[f, k]), float(hi[f, k]), 0.5))
            descs.append(mb.FloatDesc.make(lo[f, k], hi[f, k], px[-1]))
    nb = nfiles * 3 * sc3
    stride = 4 * n3 + 256
    i64 = dict(dtype=torch.int64, device=dev)
    mins, bits, offs = (torch.zeros(nb, **i64) for _ in range(3))
    out_len = torch.zeros(3 * nfiles, **i64)
    out = torch.zeros(3 * nfiles * stride, dtype=torch.uint8, device=dev)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        ctx.encode_vec3_subcells_dev(descs, aos, nfile, subcells, nfiles, mins, bits, offs, out, stride, out_len)
        dec = torch.zeros((nfiles, n3, 3), dtype=torch.float32, device=dev)
        jit = mb.Jitter.make(mb.JITTER_HASH, 5, 1000)
        ctx.decode_vec3_subcells_dev(descs, out, stride, offs, mins, bits, nfile, subcells, nfiles, 0.0, jit, dec)
    ctx.sync()
    assert ctx.last_path == 1
    mins, bits, offs, out_len, out, dec = (t.cpu().numpy() for t in (mins, bits, offs, out_len, out, dec))
    for f in range(nfiles):
        lo3, hi3, px3 = lo[f].tolist(), hi[f].tolist(), px[3 * f:3 * f + 3]
        om, ob, onb, packed, ostride, _ = orc.bench_minp_encode(vecs[f], nfile, subcells, lo3, hi3, px3)
        sl = slice(f * 3 * sc3, (f + 1) * 3 * sc3)
        assert np.array_equal(mins[sl], om) and np.array_equal(bits[sl], ob)
        for k in range(3):
            want = b"".join(packed[t * ostride:t * ostride + onb[t]].tobytes() for t in range(k * sc3, (k + 1) * sc3))
            base = (3 * f + k) * stride
            assert out_len[3 * f + k] == len(want)
            assert out[base:base + len(want)].tobytes() == want
        # jitter block ids: block_id0 + index of the block in the batch
        for k in range(3):
            for sc in range(sc3):
                t = k * sc3 + sc
                want = orc.float_block_decode(packed[t * ostride:t * ostride + onb[t]], (nfile // subcells) ** 3, int(om[t]),
                                              int(ob[t]), lo3[k], hi3[k], px3[k], 1, 1, 5, 1000 + f * 3 * sc3 + t)
                ns = nfile // subcells
                cube = dec[f].reshape(nfile, nfile, nfile, 3)
                z0, y0, x0 = ns * (sc // (subcells * subcells)), ns * ((sc // subcells) % subcells), ns * (sc % subcells)
                got = cube[z0:z0 + ns, y0:y0 + ns, x0:x0 + ns, k].reshape(-1)
                assert got.tobytes() == want.tobytes(), (f, k, sc)


def test_fast_quantiser_matches_ieee_divide(ctx):
    """quantize_fast (reciprocal multiply + 2 FMA corrections) against __fdiv_rn: exhaustive over all
    2^32 float32 inputs for the benchmark's position grid, and over 2^28 inputs for other grids"""
    px = mb.float_group_pixels(0.0, 1000.0, 0.005)
    bad, acc = ctx.selftest_fastdiv(mb.FloatDesc.make(0.0, 1000.0, px))
    assert bad == 0 and acc > 10 ** 8
    # a velocity-like grid (negative low, dx = 1), exhaustive too: both the F2I form and the RM(y + 2^23) form
    pv = mb.float_group_pixels(-1543.21, 1622.5, 1.0)
    bad, acc = ctx.selftest_fastdiv(mb.FloatDesc.make(-1543.21, 1622.5, pv))
    assert bad == 0 and acc > 10 ** 7
    rng = np.random.default_rng(3)
    for _ in range(24):
        lo = float(np.float32(rng.uniform(-2000, 2000)))
        hi = float(np.float32(lo + 10 ** rng.uniform(-2, 5)))
        pixels = int(rng.integers(2, 2 ** int(rng.integers(2, 30))))
        first = int(rng.integers(0, 2 ** 32 - 2 ** 28))
        bad, acc = ctx.selftest_fastdiv(mb.FloatDesc.make(lo, hi, pixels), first, 1 << 28)
        assert bad == 0, (lo, hi, pixels)
    # mantissa of dx all ones / just above a power of two
    for dxbits in (0x3c7fffff, 0x3c800001, 0x3cffffff, 0x3b000000):
        dx = float(np.array([dxbits], np.uint32).view(np.float32)[0])
        pixels = 100000
        hi = float(np.float32(dx) * np.float32(pixels))
        bad, acc = ctx.selftest_fastdiv(mb.FloatDesc.make(0.0, hi, pixels), 0x3f800000, 1 << 28)
        assert bad == 0, hex(dxbits)


@pytest.mark.parametrize("periodic", [True, False])
def test_minp_encode_vectors_single_upload(ctx, orc, periodic):
    """mnw_minp_encode_vectors = limits + pixels + encode of minp.Writer.Vectors (go/minp/minp.go:86-119)
    in one call: same group parameters and bytes as the oracle's writer"""
    rng = np.random.default_rng(77)
    nfile, subcells, L, dx = 32, 2, 250.0, 0.01
    vec = lagrangian(rng, nfile, L, 1.5) if periodic else (150.0 * rng.standard_normal((nfile ** 3, 3))).astype(np.float32)
    descs, mins, bits, offs, streams = ctx.minp_encode_vectors(vec, nfile, subcells, periodic, L, dx)
    lo, hi = orc.minp_limits(vec, periodic, L)
    px = [orc.float_group_pixels(float(lo[k]), float(hi[k]), np.float32(dx)) for k in range(3)]
    for k in range(3):
        assert (np.float32(descs[k].low), np.float32(descs[k].high), descs[k].pixels) == (lo[k], hi[k], px[k])
    om, ob, onb, packed, stride, _ = orc.bench_minp_encode(vec, nfile, subcells, lo.tolist(), hi.tolist(), px)
    assert np.array_equal(mins, om) and np.array_equal(bits, ob)
    sc3 = subcells ** 3
    for k in range(3):
        want = b"".join(packed[t * stride:t * stride + onb[t]].tobytes() for t in range(k * sc3, (k + 1) * sc3))
        assert streams[k].tobytes() == want


# ---- 64^3 sub-cells: the warp-specialised k_pipe_vec3 path -------------------------------------
def test_pipe_edges_64(ctx, orc):
    """64^3 blocks over several units per file: 0-bit blocks, index == pixels folding (thread-local
    exact redo), a wrap-around arc and a first element at index == pixels (generic redo)"""
    rng = np.random.default_rng(64)
    L, nfile, subcells, dx = 1000.0, 128, 2, 0.005
    px = mb.float_group_pixels(0.0, L, dx)
    vec = lagrangian(rng, nfile, L, 2.0)
    vec[:, 1] = 321.125                                   # y: 0 bits everywhere
    top = np.nextafter(np.float32(L), np.float32(0))      # quantises to index == pixels
    vec[5::9973, 2] = top
    vec[11::8191, 2] = 0.0
    mins, bits = check(ctx, orc, vec, nfile, subcells, [0.0] * 3, [L] * 3, [px] * 3, L)
    sc3 = subcells ** 3
    assert (bits[sc3:2 * sc3] == 0).all()
    vec2 = vec.copy()
    vec2[0, 2] = top
    check(ctx, orc, vec2, nfile, subcells, [0.0] * 3, [L] * 3, [px] * 3, L)


@pytest.mark.parametrize("bad", [np.nan, -np.inf, -3.0, 2500.0])
def test_pipe_bad_values_64(ctx, orc, bad):
    rng = np.random.default_rng(65)
    L, nfile, subcells, dx = 1000.0, 64, 1, 0.005
    px = mb.float_group_pixels(0.0, L, dx)
    vec = lagrangian(rng, nfile, L, 2.0)
    vec[77777, 0] = bad
    vec[200000, 2] = bad
    descs = [mb.FloatDesc.make(0.0, L, px) for _ in range(3)]
    mins, bits, offs, streams = ctx.encode_vec3_subcells(descs, vec, nfile, subcells)
    omins, obits, onbytes, packed, stride, total = orc.bench_minp_encode(vec, nfile, subcells, [0.0] * 3, [L] * 3, [px] * 3)
    assert np.array_equal(mins, omins) and np.array_equal(bits, obits)
    for k in range(3):
        assert streams[k].tobytes() == packed[k * stride:k * stride + onbytes[k]].tobytes()


@pytest.mark.parametrize("dx,sigma", [(0.05, 3.0), (0.0005, 0.5), (2.0, 30.0), (30.0, 60.0)])
def test_pipe_bit_widths_64(ctx, orc, dx, sigma):
    """narrow and wide arcs, > 16 bit blocks (repack list) and tiny pixel counts on 64^3 blocks"""
    rng = np.random.default_rng(int(dx * 1e4) + 66)
    L, nfile, subcells = 100.0, 128, 2
    vec = lagrangian(rng, nfile, L, sigma)
    px = mb.float_group_pixels(0.0, L, dx)
    mins, bits = check(ctx, orc, vec, nfile, subcells, [0.0] * 3, [L] * 3, [px] * 3, L)
    if dx == 0.0005:
        assert bits.max() > 16


def test_pipe_many_files_dev_64(ctx, orc):
    """device entry point, 3 files x 8 units of 64^3 with per-file limits: ticket order, look-back chains
    across clusters and units, non-periodic wide blocks"""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(67)
    nfile, subcells, nfiles = 128, 2, 3
    sc3, n3 = subcells ** 3, nfile ** 3
    vecs = [(150.0 * (f + 1) * rng.standard_normal((n3, 3))).astype(np.float32) for f in range(nfiles)]
    dev = torch.device("cuda:0")
    aos = torch.from_numpy(np.concatenate(vecs)).to(dev)
    lo, hi = ctx.vec3_limits(aos, nfiles, dev=True)
    descs, pxs = [], []
    for f in range(nfiles):
        for k in range(3):
            px = mb.float_group_pixels(float(lo[f, k]), float(hi[f, k]), 1.0)
            descs.append(mb.FloatDesc.make(lo[f, k], hi[f, k], px))
            pxs.append(px)
    nb = nfiles * 3 * sc3
    stride = 4 * (nfile // subcells) ** 3 * sc3
    i64 = dict(dtype=torch.int64, device=dev)
    mins, bits, offs = torch.zeros(nb, **i64), torch.zeros(nb, **i64), torch.zeros(nb, **i64)
    out_len = torch.zeros(3 * nfiles, **i64)
    out = torch.zeros(3 * nfiles * stride, dtype=torch.uint8, device=dev)
    ctx.encode_vec3_subcells_dev(descs, aos, nfile, subcells, nfiles, mins, bits, offs, out, stride, out_len)
    torch.cuda.synchronize()
    mins, bits, offs, out_len, out = (t.cpu().numpy() for t in (mins, bits, offs, out_len, out))
    for f in range(nfiles):
        lo3 = [d.low for d in descs[3 * f:3 * f + 3]]
        hi3 = [d.high for d in descs[3 * f:3 * f + 3]]
        om, ob, onb, packed, ostride, _ = orc.bench_minp_encode(vecs[f], nfile, subcells, lo3, hi3, pxs[3 * f:3 * f + 3])
        sl = slice(f * 3 * sc3, (f + 1) * 3 * sc3)
        assert np.array_equal(mins[sl], om) and np.array_equal(bits[sl], ob)
        for k in range(3):
            want = b"".join(packed[t * ostride:t * ostride + onb[t]].tobytes() for t in range(k * sc3, (k + 1) * sc3))
            assert out_len[3 * f + k] == len(want)
            got = out[(3 * f + k) * stride:(3 * f + k) * stride + len(want)].tobytes()
            assert got == want
            assert np.array_equal(offs[sl][k * sc3:(k + 1) * sc3], np.concatenate([[0], np.cumsum(onb[k * sc3:(k + 1) * sc3])[:-1]]))


@pytest.mark.parametrize("case", ["edges", "bad", "widths", "files"])
def test_pipe_cooperative_schedule_64(ctx, orc, case, monkeypatch):
    """the cluster-free cooperative schedule of k_pipe_vec3 (default only for batches of >= 256 units) on the small
    parity cases: forced with MNW_PIPE_COOP_MIN=1"""
    monkeypatch.setenv("MNW_PIPE_COOP_MIN", "1")
    if case == "edges":
        test_pipe_edges_64(ctx, orc)
    elif case == "bad":
        test_pipe_bad_values_64(ctx, orc, np.nan)
        test_pipe_bad_values_64(ctx, orc, 2500.0)
    elif case == "widths":
        test_pipe_bit_widths_64(ctx, orc, 0.05, 3.0)
        test_pipe_bit_widths_64(ctx, orc, 30.0, 60.0)
    else:
        test_pipe_many_files_dev_64(ctx, orc)


def test_pipe_cooperative_large_batch_roundtrip(ctx):
    """a batch large enough for the cooperative schedule by default (4 files x 64 units of 64^3): encode, decode
    with CENTER jitter, and check the domain's size-independent properties: every decoded value within dx/2 of
    its input, offsets = running sum of ArrayBytes(bits, n), out_len = last offset + size"""
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda:0")
    nfile, subcells, nfiles, L, dx = 256, 4, 4, 1000.0, 0.005
    sc3, n3, nsub3 = subcells ** 3, nfile ** 3, (nfile // subcells) ** 3
    g = torch.Generator(device=dev); g.manual_seed(5)
    aos = torch.rand((nfiles, n3, 3), generator=g, device=dev, dtype=torch.float32) * 3.0
    j = torch.arange(nfile, device=dev, dtype=torch.float32) * (L / nfile)
    grid = torch.stack(torch.meshgrid(j, j, j, indexing="ij")[::-1], dim=-1).reshape(n3, 3)
    aos = torch.remainder(aos + grid[None], L).contiguous()
    aos[aos >= L] = 0.0
    px = mb.float_group_pixels(0.0, L, dx)
    descs = [mb.FloatDesc.make(0.0, L, px) for _ in range(3)]
    nb = nfiles * 3 * sc3
    stride = 4 * nsub3 * sc3
    i64 = dict(dtype=torch.int64, device=dev)
    mins, bits, offs = (torch.zeros(nb, **i64) for _ in range(3))
    out_len = torch.zeros(3 * nfiles, **i64)
    out = torch.zeros(3 * nfiles * stride, dtype=torch.uint8, device=dev)
    ctx.encode_vec3_subcells_dev(descs, aos, nfile, subcells, nfiles, mins, bits, offs, out, stride, out_len)
    dec = torch.empty_like(aos)
    ctx.decode_vec3_subcells_dev(descs, out, stride, offs, mins, bits, nfile, subcells, nfiles, L, mb.Jitter.make(mb.JITTER_CENTER, 0), dec)
    torch.cuda.synchronize()
    assert ctx.last_path == 1
    d = (dec - aos).abs()
    d = torch.minimum(d, L - d)
    assert float(d.max()) <= 0.5 * dx * 1.01 + 1e-4
    nbytes = (bits * nsub3 + 7) // 8
    o = offs.reshape(3 * nfiles, sc3); nbs = nbytes.reshape(3 * nfiles, sc3)
    assert torch.equal(o, torch.cumsum(nbs, 1) - nbs)
    assert torch.equal(out_len, o[:, -1] + nbs[:, -1])
    assert int(bits.max()) <= 16 and int(bits.min()) >= 1


@pytest.mark.parametrize("nfile,subcells", [(128, 1), (256, 2)])
def test_pipe_128_subcells(ctx, orc, nfile, subcells):
    """128^3 sub-cells: 64 CTAs per unit in the cooperative k_pipe_vec3 schedule, k_decode_vec3<128>"""
    rng = np.random.default_rng(128 + subcells)
    L, dx = 500.0, 0.01
    vec = lagrangian(rng, nfile, L, 1.5)
    px = mb.float_group_pixels(0.0, L, dx)
    mins, bits = check(ctx, orc, vec, nfile, subcells, [0.0] * 3, [L] * 3, [px] * 3, L)
    if subcells == 2:
        assert bits.max() <= 16   # packed by k_pipe_vec3 itself, not by the > 16 bit list


def test_pipe_128_velocities(ctx, orc):
    rng = np.random.default_rng(1280)
    nfile, subcells = 128, 1
    vec = (250.0 * rng.standard_normal((nfile ** 3, 3))).astype(np.float32)
    lo, hi = orc.minp_limits(vec, False, 0.0)
    px = [mb.float_group_pixels(float(lo[k]), float(hi[k]), 1.0) for k in range(3)]
    vec[12345, 1] = np.nan                      # one bad value: the generic redo has to take over
    descs = [mb.FloatDesc.make(lo[k], hi[k], px[k]) for k in range(3)]
    mins, bits, offs, streams = ctx.encode_vec3_subcells(descs, vec, nfile, subcells)
    omins, obits, onbytes, packed, stride, total = orc.bench_minp_encode(vec, nfile, subcells, lo.tolist(), hi.tolist(), px)
    assert np.array_equal(mins, omins) and np.array_equal(bits, obits)
    for k in range(3):
        assert streams[k].tobytes() == packed[k * stride:k * stride + onbytes[k]].tobytes()
