"""bench_configs.py -- part of bench.py: the BASELINE.json configurations beside the headline one (C2).

  C1  configs[0]: one minnow file, 2^20 halos, int64 ID + 3 float32 positions (1 kpc/h pixels in a 125 Mpc/h box),
      16 blocks of 65536 per group                         -> IntGroup + 3 FloatGroups, encode + decode
  C3  configs[2]: minh halo catalogue, 40 mixed columns: ONE block of 2^22 rows (the full 10^8-row catalogue is 24 such
      blocks, each processed exactly like this one)        -> 30 quantised columns on the GPU, 10 fixed-size columns
      are a plain copy in the format (go/group.go:150-153) and are not timed
  C4  configs[3]: random access, 10^4 selected blocks of 16^3 = 4096 values out of the 3 x 32768 blocks of one axis-major
      512^3 snapshot file (SubCells = 32)                  -> decode latency + GB/s

Each returns {"workload", "value", "unit", "ms", "roofline", "e2e", "cpu_baseline", "verified"}: value = uncompressed
GB/s on device-resident data (CUDA events on the library's stream, L2 flushed between repetitions where the data would
fit it), roofline = the dominant kernel against the measured copy bandwidth, e2e = the same work through the
host-pointer C ABI with pinned host buffers (copies timed), cpu_baseline = the oracle port (test infrastructure: the
thing timed here as the BASELINE, never the product) on 1 core and on all cores, verified = GPU bytes and decoded
values against the oracle on this very input (not timed)."""
import ctypes as C
import threading
import time

import numpy as np


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


class Timer:
    """CUDA-event timing of closures on the library's stream, with an L2 flush (a 256 MB write) before every
    repetition when asked for."""

    def __init__(self, torch, ctx, stream, dev, flush):
        self.torch, self.ctx, self.stream = torch, ctx, stream
        self.scratch = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if flush else None

    def __call__(self, fn, reps=7):
        torch, best = self.torch, []
        for r in range(reps + 2):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(self.stream):
                if self.scratch is not None:
                    self.scratch.fill_(r)
                a.record(self.stream); fn(); b.record(self.stream)
            self.ctx.sync()
            if r >= 2:
                best.append(a.elapsed_time(b))
        return float(np.median(best))


def profile_of(ctx, stream, torch, fn):
    ctx.profile(True)
    with torch.cuda.stream(stream):
        fn()
    ctx.sync()
    ctx.profile(False)
    return ctx.profile_summary()


def roofline_of(prof, algo_bytes, peak):
    """prof: [{"kernel", "launches", "ms"}] of ONE pass of the work; algo_bytes: {kernel: algorithmic bytes of that pass}"""
    prof = [p for p in prof if p["ms"] > 0]
    if not prof:
        return None
    top = max(prof, key=lambda p: p["ms"])
    a = algo_bytes.get(top["kernel"], 0) / (top["ms"] * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": top["kernel"], "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak,
            "traffic": None, "ms_per_pass": top["ms"], "launches_per_pass": top["launches"],
            "algorithmic_bytes_per_pass": algo_bytes.get(top["kernel"], 0),
            "kernels": [dict(p, achieved_gbs=(algo_bytes.get(p["kernel"], 0) / (p["ms"] * 1e-3) / 1e9)) for p in prof]}


def timed_cpu(fn, min_s=1.0, max_reps=5):
    fn()
    t0, reps = time.perf_counter(), 0
    while reps < 1 or (time.perf_counter() - t0 < min_s and reps < max_reps):
        fn(); reps += 1
    return (time.perf_counter() - t0) / reps


def pin(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()


# ---------------------------------------------------------------------------------------------------------------------
def run_c1(torch, mb, orc, ctx, stream, dev, peak, threads):
    n, nb, N = 65536, 16, 1 << 20
    g = torch.Generator(device=dev); g.manual_seed(20261018)
    ids = (torch.randperm(N, generator=g, device=dev) + 10 ** 9).to(torch.int64)
    pos = [torch.rand(N, generator=g, device=dev, dtype=torch.float32) * 125.0 for _ in range(3)]
    px = mb.float_group_pixels(0.0, 125.0, 0.001)
    d = mb.FloatDesc.make(0.0, 125.0, px)
    jit = mb.Jitter.make(mb.JITTER_HASH, 11)
    i64 = dict(dtype=torch.int64, device=dev)
    meta = [[torch.zeros(nb, **i64) for _ in range(3)] for _ in range(4)]
    lens = [torch.zeros(1, **i64) for _ in range(4)]
    outs = [torch.empty(8 * N + 256, dtype=torch.uint8, device=dev) for _ in range(4)]
    dec_i = torch.empty(N, **i64)
    dec_f = [torch.empty(N, dtype=torch.float32, device=dev) for _ in range(3)]

    def enc():
        ctx.encode_int_group_dev(ids, n, nb, *meta[0], outs[0], outs[0].numel(), lens[0])
        for k in range(3):
            ctx.encode_float_group_dev(d, pos[k], n, nb, *meta[k + 1], outs[k + 1], outs[k + 1].numel(), lens[k + 1])

    def dec():
        ctx.decode_int_blocks_dev(outs[0], outs[0].numel(), meta[0][2], meta[0][0], meta[0][1], n, nb, None, dec_i)
        for k in range(3):
            ctx.decode_float_blocks_dev(d, outs[k + 1], outs[k + 1].numel(), meta[k + 1][2], meta[k + 1][0], meta[k + 1][1],
                                        n, nb, None, jit, dec_f[k])
    timer = Timer(torch, ctx, stream, dev, flush=True)
    ms_e, ms_d = timer(enc), timer(dec)
    raw = N * (8 + 12)
    packed = sum(int(l.item()) for l in lens)
    prof = profile_of(ctx, stream, torch, lambda: (enc(), dec()))
    algo = {"k_group_fused": raw + packed, "k_decode_f32c": 12 * N + packed - int(lens[0].item()),
            "k_decode_i64c": 8 * N + int(lens[0].item())}
    roof = roofline_of(prof, algo, peak)

    # ---- parity on this input: bytes, (min, bits), decoded values (HASH jitter) against the oracle
    ids_h, pos_h = ids.cpu().numpy(), [p.cpu().numpy() for p in pos]
    ok = bool(torch.equal(dec_i, ids))
    om, ob, onb, opk, ost, _ = orc.bench_group_encode(ids_h, n, nb, None, threads)
    odesc = (0.0, 125.0, px, 0, 0)
    for gi, (hm, hb, hnb, hpk, hst) in enumerate([(om, ob, onb, opk, ost)] + [orc.bench_group_encode(p, n, nb, odesc, threads)[:5] for p in pos_h]):
        gm, gb, go = (t.cpu().numpy() for t in meta[gi])
        gbytes = outs[gi][:int(lens[gi].item())].cpu().numpy()
        want = b"".join(hpk[b * hst:b * hst + hnb[b]].tobytes() for b in range(nb))
        ok = ok and np.array_equal(gm, hm) and np.array_equal(gb, hb) and gbytes.tobytes() == want
        ok = ok and np.array_equal(go, np.concatenate([[0], np.cumsum(hnb)[:-1]]))
        if gi > 0:
            hd = orc.bench_group_decode(hpk, hst, n, hm, hb, odesc, None, 1, 11, threads)
            ok = ok and dec_f[gi - 1].cpu().numpy().tobytes() == hd.tobytes()

    # ---- e2e through the host-pointer ABI, pinned host buffers
    hid, hpos = pin(torch, ids_h), [pin(torch, p) for p in pos_h]
    hout = [pin(torch, np.empty(8 * N + 64, np.uint8)) for _ in range(4)]
    hdi, hdf = pin(torch, np.empty(N, np.int64)), [pin(torch, np.empty(N, np.float32)) for _ in range(3)]
    hm_ = [[np.zeros(nb, np.int64) for _ in range(3)] for _ in range(4)]
    P = lambda a: C.c_void_p(a.ctypes.data)
    lib, h = ctx.lib, ctx.h

    def e2e_once():
        ln = [C.c_int64(0) for _ in range(4)]
        ctx._check(lib.mnw_encode_int_group(h, P(hid), n, nb, None, P(hm_[0][0]), P(hm_[0][1]), P(hm_[0][2]), P(hout[0]), len(hout[0]), C.byref(ln[0])))
        for k in range(3):
            ctx._check(lib.mnw_encode_float_group(h, C.byref(d), P(hpos[k]), n, nb, None, P(hm_[k + 1][0]), P(hm_[k + 1][1]), P(hm_[k + 1][2]),
                                                  P(hout[k + 1]), len(hout[k + 1]), C.byref(ln[k + 1])))
        ctx._check(lib.mnw_decode_int_blocks(h, P(hout[0]), ln[0].value, P(hm_[0][2]), P(hm_[0][0]), P(hm_[0][1]), n, nb, None, P(hdi)))
        for k in range(3):
            ctx._check(lib.mnw_decode_float_blocks(h, C.byref(d), P(hout[k + 1]), ln[k + 1].value, P(hm_[k + 1][2]), P(hm_[k + 1][0]),
                                                   P(hm_[k + 1][1]), n, nb, None, C.byref(jit), P(hdf[k])))
        return sum(l.value for l in ln)
    pk = e2e_once()
    t_e2e = timed_cpu(e2e_once, 0.5, 20)
    e2e = {"value": 2 * raw / t_e2e / 1e9, "unit": "GB/s", "h2d_bytes_per_step": raw + pk + 8 * 3 * 4 * nb,
           "d2h_bytes_per_step": raw + pk + 8 * 3 * 4 * nb, "api": "mnw_encode_int_group + 3 x mnw_encode_float_group + the 4 decodes, one host thread"}
    ok = ok and np.array_equal(hdi, ids_h)

    # ---- CPU baseline (oracle port), 1 core and all cores
    def cpu(th):
        def once():
            m, b, nbt, pkd, st, _ = orc.bench_group_encode(ids_h, n, nb, None, th)
            orc.bench_group_decode(pkd, st, n, m, b, None, None, 1, 11, th)
            for p in pos_h:
                m, b, nbt, pkd, st, _ = orc.bench_group_encode(p, n, nb, odesc, th)
                orc.bench_group_decode(pkd, st, n, m, b, odesc, None, 1, 11, th)
        return 2 * raw / timed_cpu(once, 1.0, 3) / 1e9
    c1, call = cpu(1), cpu(threads)
    return {"workload": "C1 (BASELINE configs[0]): 2^20 halos, int64 ID IntGroup + 3 float32 position FloatGroups ([0,125) at 0.001 -> 125000 px), "
                        "16 blocks x 65536 per group; encode + decode; L2 flushed before every repetition (21 MB of input)",
            "value": 2 * raw / ((ms_e + ms_d) * 1e-3) / 1e9, "unit": "GB/s", "ms": ms_e + ms_d, "encode_ms": ms_e, "decode_ms": ms_d,
            "mean_bits": 8.0 * packed / (4 * N), "note": "4 MB per column: launch-latency bound on the device",
            "roofline": roof, "e2e": e2e,
            "cpu_baseline": {"value": call, "unit": "GB/s", "cores": threads, "kind": "port", "one_core_value": c1, "cpu": cpu_model(),
                             "sample": "the whole configuration (2^20 halos), encode + decode, OpenMP over blocks"},
            "verified": {"bytes_equal_oracle": bool(ok)}}


# ---------------------------------------------------------------------------------------------------------------------
def run_c3(torch, mb, orc, ctx, stream, dev, peak, threads):
    n = 1 << 22
    g = torch.Generator(device=dev); g.manual_seed(3)
    px = mb.float_group_pixels(0.0, 125.0, 0.001)
    dpos = mb.FloatDesc.make(0.0, 125.0, px, 1, 0, 1)
    lpx = mb.float_group_pixels(10.0, 15.0, 0.01)
    dlog = mb.FloatDesc.make(10.0, 15.0, lpx, 1, 1, 1)
    ids0 = torch.arange(n, dtype=torch.int64, device=dev) * 3 + torch.randint(0, 3, (n,), generator=g, device=dev) + 10 ** 9
    cols = [(ids0 + 7 * k, None) for k in range(6)]
    cols += [(torch.rand(n, generator=g, device=dev, dtype=torch.float32) * 125.0, dpos) for _ in range(12)]
    cols += [(torch.pow(10.0, 10.0 + 5.0 * torch.rand(n, generator=g, device=dev, dtype=torch.float32)), dlog) for _ in range(12)]
    nc = len(cols)
    i64 = dict(dtype=torch.int64, device=dev)
    dec_i, dec_f = torch.empty(n, **i64), torch.empty(n, dtype=torch.float32, device=dev)
    jit = mb.Jitter.make(mb.JITTER_HASH, 3)

    stride_d = 8 * n + 256
    out_all = torch.empty(nc * stride_d, dtype=torch.uint8, device=dev)
    outs = [out_all[c * stride_d:(c + 1) * stride_d] for c in range(nc)]
    mins_all, bits_all, lens_all = (torch.zeros(nc, **i64) for _ in range(3))
    offs0 = torch.zeros(1, **i64)
    meta = [[mins_all[c:c + 1], bits_all[c:c + 1], offs0] for c in range(nc)]
    lens = [lens_all[c:c + 1] for c in range(nc)]

    def enc():   # every column of the block in ONE call (minh.Writer.Block, go/minh/minh.go:99-139)
        ctx.encode_columns_dev(cols, n, mins_all, bits_all, lens_all, out_all, stride_d)

    col_offs = torch.arange(nc, **i64) * stride_d
    dec_outs = [torch.empty(n, **i64) if d is None else torch.empty(n, dtype=torch.float32, device=dev) for _, d in cols]

    def dec():   # every column of the block in ONE call, two launches (minh.Reader.Block, go/minh/minh.go:296-323)
        ctx.decode_columns_dev([d for _, d in cols], out_all, col_offs, mins_all, bits_all, n, jit, dec_outs)
    timer = Timer(torch, ctx, stream, dev, flush=False)   # 604 MB of input per pass: far larger than L2
    ms_e, ms_d = timer(enc, 5), timer(dec, 5)
    raw = n * (6 * 8 + 24 * 4)
    packed = sum(int(l.item()) for l in lens)
    pk_i = sum(int(lens[c].item()) for c in range(6))
    prof = profile_of(ctx, stream, torch, lambda: (enc(), dec()))
    algo = {"k_group_fused": raw + packed, "k_decode_f32c": 24 * 4 * n + packed - pk_i, "k_decode_i64c": 6 * 8 * n + pk_i}
    roof = roofline_of(prof, algo, peak)

    host_cols = [x.cpu().numpy() for x, _ in cols]

    # ---- e2e: mnw_encode_columns (one call, host pointers) + the 30 decodes from host memory
    from minnow_b200.capi import Column
    hc = [pin(torch, a) for a in host_cols]
    stride = 8 * n + 16
    hout = pin(torch, np.empty(nc * stride, np.uint8))
    hdi, hdf = pin(torch, np.empty(n, np.int64)), pin(torch, np.empty(n, np.float32))
    m3, b3, l3 = (np.zeros(nc, np.int64) for _ in range(3))
    zero = np.zeros(1, np.int64)
    carr, parr = (Column * nc)(), (C.c_void_p * nc)()
    for i, (x, d) in enumerate(cols):
        carr[i].is_float = 0 if d is None else 1
        if d is not None:
            carr[i].desc = d
        parr[i] = hc[i].ctypes.data
    P = lambda a: C.c_void_p(a.ctypes.data)
    lib, h = ctx.lib, ctx.h

    def e2e_once():
        ctx._check(lib.mnw_encode_columns(h, nc, carr, parr, n, P(m3), P(b3), P(l3), P(hout), stride))
        for i, (x, d) in enumerate(cols):
            o = hout[i * stride:]
            if d is None:
                ctx._check(lib.mnw_decode_int_blocks(h, P(o), int(l3[i]), P(zero), P(m3[i:]), P(b3[i:]), n, 1, None, P(hdi)))
            else:
                ctx._check(lib.mnw_decode_float_blocks(h, C.byref(d), P(o), int(l3[i]), P(zero), P(m3[i:]), P(b3[i:]), n, 1, None,
                                                       C.byref(jit), P(hdf)))
    e2e_once()
    t_e2e = timed_cpu(e2e_once, 1.0, 3)
    pk = int(l3.sum())
    e2e = {"value": 2 * raw / t_e2e / 1e9, "unit": "GB/s", "h2d_bytes_per_step": raw + pk, "d2h_bytes_per_step": raw + pk,
           "api": "mnw_encode_columns (all 30 columns, one call) + 30 x mnw_decode_{int,float}_blocks, one host thread; the device-resident figure uses mnw_decode_columns_dev (one call)"}

    # ---- CPU baseline: one column per thread (the reference is single-threaded; columns are independent)
    results = {}

    def cpu(th, keep):
        def col(c):
            x, d = cols[c]
            od = None if d is None else (d.low, d.high, d.pixels, d.log10, d.clamp)
            m, b, nbt, pkd, st, _ = orc.bench_group_encode(host_cols[c], n, 1, od, 1)
            hd = orc.bench_group_decode(pkd, st, n, m, b, od, None, 1, 3, 1)
            if keep:
                results[c] = (m, b, nbt, pkd, hd if c in (0, 6) else None)

        def once():
            if th == 1:
                for c in range(nc):
                    col(c)
                return
            todo, lock = list(range(nc)), threading.Lock()

            def work():
                while True:
                    with lock:
                        if not todo:
                            return
                        c = todo.pop()
                    col(c)
            ths = [threading.Thread(target=work) for _ in range(min(th, nc))]
            [t.start() for t in ths]; [t.join() for t in ths]
        t0 = time.perf_counter(); once()
        return 2 * raw / (time.perf_counter() - t0) / 1e9
    call, c1 = cpu(threads, True), cpu(1, False)

    # ---- parity: every column's (min, bits, bytes) against the oracle's, and the decoded values (HASH jitter) of an
    # IntGroup and a FloatGroup column (the Log columns' 10^x read side is a tolerance matter)
    ok = True
    for c, (x, d) in enumerate(cols):
        hm, hb, hnb, hpk, hd = results[c]
        ok = ok and int(meta[c][0].item()) == int(hm[0]) and int(meta[c][1].item()) == int(hb[0])
        ok = ok and outs[c][:int(lens[c].item())].cpu().numpy().tobytes() == hpk[:hnb[0]].tobytes()
        if hd is not None:
            with torch.cuda.stream(stream):
                if d is None:
                    ctx.decode_int_blocks_dev(outs[c], outs[c].numel(), meta[c][2], meta[c][0], meta[c][1], n, 1, None, dec_i)
                else:
                    ctx.decode_float_blocks_dev(d, outs[c], outs[c].numel(), meta[c][2], meta[c][0], meta[c][1], n, 1, None, jit, dec_f)
            ctx.sync()
            ok = ok and (dec_i if d is None else dec_f).cpu().numpy().tobytes() == hd.tobytes()
    results.clear()

    return {"workload": "C3 (BASELINE configs[2]): minh catalogue, one block of 2^22 rows (the 10^8-row catalogue = 24 such blocks) x 40 columns = "
                        "6 IntGroup + 12 linear FloatGroups ([0,125) at 0.001) + 12 log10 FloatGroups ([10,15) at 0.01, clamp) on the GPU; the 10 "
                        "fixed-size columns are a plain copy in the format and not timed; encode + decode; 604 MB of input per pass (>> L2)",
            "value": 2 * raw / ((ms_e + ms_d) * 1e-3) / 1e9, "unit": "GB/s", "ms": ms_e + ms_d, "encode_ms": ms_e, "decode_ms": ms_d,
            "mean_bits": 8.0 * packed / (30 * n), "roofline": roof, "e2e": e2e,
            "cpu_baseline": {"value": call, "unit": "GB/s", "cores": min(threads, nc), "kind": "port", "one_core_value": c1, "cpu": cpu_model(),
                             "sample": "this block (2^22 rows x 30 quantised columns), encode + decode, one pass, one column per thread"},
            "verified": {"bytes_equal_oracle": bool(ok)}}


# ---------------------------------------------------------------------------------------------------------------------
def run_c4(torch, mb, orc, ctx, stream, dev, peak, threads):
    nfile, sub, L, dx = 512, 32, 1000.0, 0.005
    nsub, sc3 = nfile // sub, sub ** 3
    n, nsel = nsub ** 3, 10000
    g = torch.Generator(device=dev); g.manual_seed(4)
    j = torch.arange(nfile, device=dev, dtype=torch.float32) * (L / nfile)
    grid = torch.stack([j.view(1, 1, -1).expand(nfile, nfile, nfile), j.view(1, -1, 1).expand(nfile, nfile, nfile),
                        j.view(-1, 1, 1).expand(nfile, nfile, nfile)], -1).reshape(-1, 3)
    disp = (torch.rand((nfile ** 3, 3), generator=g, device=dev, dtype=torch.float32) +
            torch.rand((nfile ** 3, 3), generator=g, device=dev, dtype=torch.float32) - 1.0) * 4.0
    pos = torch.remainder(grid + disp, L)
    pos = torch.where(pos >= L, torch.zeros_like(pos), pos).contiguous()
    del grid, disp
    px = mb.float_group_pixels(0.0, L, dx)
    descs = [mb.FloatDesc.make(0.0, L, px) for _ in range(3)]
    nb = 3 * sc3
    i64 = dict(dtype=torch.int64, device=dev)
    mins, bits, offs = (torch.zeros(nb, **i64) for _ in range(3))
    lens = torch.zeros(3, **i64)
    stride = 4 * nfile ** 3 + 256
    packed = torch.empty(3 * stride, dtype=torch.uint8, device=dev)
    with torch.cuda.stream(stream):
        ctx.encode_vec3_subcells_dev(descs, pos, nfile, sub, 1, mins, bits, offs, packed, stride, lens)
    ctx.sync()
    # axis 0's group: 32768 blocks of 4096; select 10^4 of them by a seeded shuffle
    sel = torch.randperm(sc3, generator=g, device=dev)[:nsel].contiguous()
    dec = torch.empty(nsel * n, dtype=torch.float32, device=dev)
    jit = mb.Jitter.make(mb.JITTER_HASH, 4)
    m0, b0, o0 = mins[:sc3].contiguous(), bits[:sc3].contiguous(), offs[:sc3].contiguous()
    glen = int(lens[0].item())
    fn = lambda: ctx.decode_float_blocks_dev(descs[0], packed, glen, o0, m0, b0, n, nsel, sel, jit, dec)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lat = []
    for r in range(103):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            flush.fill_(r)
            a.record(stream); fn(); b.record(stream)
        ctx.sync()
        if r >= 3:
            lat.append(a.elapsed_time(b) * 1e3)
    lat = np.array(lat)
    p50 = float(np.percentile(lat, 50))
    sel_h = sel.cpu().numpy()
    bsel = b0.cpu().numpy()[sel_h]
    pk_sel = int(((bsel * n + 7) // 8).sum())
    prof = profile_of(ctx, stream, torch, fn)
    roof = roofline_of(prof, {"k_decode_f32c": 4 * n * nsel + pk_sel}, peak)

    # ---- parity: the selected blocks' bytes and decoded values against the oracle's own encode of the same sub-cells
    cube = pos.view(nfile, nfile, nfile, 3)
    blocks = np.empty((nsel, n), np.float32)
    for t, sc in enumerate(sel_h):   # getSubCell, go/minp/minp.go:246-264 (x fastest)
        ix, iy, iz = int(sc) % sub, (int(sc) // sub) % sub, int(sc) // (sub * sub)
        blocks[t] = cube[iz * nsub:(iz + 1) * nsub, iy * nsub:(iy + 1) * nsub, ix * nsub:(ix + 1) * nsub, 0].reshape(-1).cpu().numpy()
    od = (0.0, L, px, 0, 0)
    hm, hb, hnb, hpk, hst, _ = orc.bench_group_encode(blocks.reshape(-1), n, nsel, od, threads)
    gm, gb, go = m0.cpu().numpy()[sel_h], bsel, o0.cpu().numpy()[sel_h]
    ok = np.array_equal(gm, hm) and np.array_equal(gb, hb)
    pk_h = packed[:glen].cpu().numpy()
    for t in range(0, nsel, 37):
        ok = ok and pk_h[go[t]:go[t] + hnb[t]].tobytes() == hpk[t * hst:t * hst + hnb[t]].tobytes()
    # decoded values: the oracle decodes ITS blocks with the jitter ids of the selected sub-cells
    hd = np.empty((nsel, n), np.float32)
    for t in range(0, nsel, 101):
        hd[t] = orc.float_block_decode(hpk[t * hst:t * hst + hnb[t]], n, int(hm[t]), int(hb[t]), 0.0, L, px, 1, 1, 4, int(sel_h[t]))
        ok = ok and dec.view(nsel, n)[t].cpu().numpy().tobytes() == hd[t].tobytes()

    # ---- e2e: the same 10^4 blocks from a pinned host copy of the group's bytes (only the selected blocks are uploaded)
    hpk_pin = pin(torch, pk_h)
    hdec = pin(torch, np.empty(nsel * n, np.float32))
    o_h, m_h, b_h = o0.cpu().numpy(), m0.cpu().numpy(), b0.cpu().numpy()
    P = lambda a: C.c_void_p(a.ctypes.data)

    def e2e_once():
        ctx._check(ctx.lib.mnw_decode_float_blocks(ctx.h, C.byref(descs[0]), P(hpk_pin), glen, P(o_h), P(m_h), P(b_h), n, nsel, P(sel_h),
                                                   C.byref(jit), P(hdec)))
    e2e_once()
    t_e2e = timed_cpu(e2e_once, 0.5, 20)
    e2e = {"value": 4 * n * nsel / t_e2e / 1e9, "unit": "GB/s", "latency_ms": t_e2e * 1e3, "h2d_bytes_per_step": pk_sel + 24 * nsel,
           "d2h_bytes_per_step": 4 * n * nsel, "api": "mnw_decode_float_blocks with a block selection, one host thread"}

    def cpu(th):
        return 4 * n * nsel / timed_cpu(lambda: orc.bench_group_decode(hpk, hst, n, hm, hb, od, None, 1, 4, th), 1.0, 3) / 1e9
    c1, call = cpu(1), cpu(threads)
    return {"workload": "C4 (BASELINE configs[3]): random access, 10^4 selected blocks of 16^3 = 4096 values out of the 32768 blocks of one axis group "
                        "of a 512^3 snapshot file (SubCells = 32), decode only; L2 flushed before every batch",
            "value": 4.0 * n * nsel / p50 / 1e3, "unit": "GB/s (decoded float32)", "batch_latency_us_p50": p50,
            "batch_latency_us_p99": float(np.percentile(lat, 99)), "mean_bits": float(bsel.mean()), "roofline": roof, "e2e": e2e,
            "cpu_baseline": {"value": call, "unit": "GB/s", "cores": threads, "kind": "port", "one_core_value": c1, "cpu": cpu_model(),
                             "sample": "the same 10^4 blocks, decode only, OpenMP over blocks"},
            "verified": {"bytes_equal_oracle": bool(ok)}}


def run_all(torch, mb, orc, ctx, stream, dev, peak, threads):
    out = {}
    for name, fn in (("C1", run_c1), ("C3", run_c3), ("C4", run_c4)):
        try:
            out[name] = fn(torch, mb, orc, ctx, stream, dev, peak, threads)
        except Exception as exc:   # a failed extra configuration is reported, it never costs the run its headline number
            out[name] = {"error": "%s: %s" % (type(exc).__name__, exc)}
        torch.cuda.empty_cache()
    return out
