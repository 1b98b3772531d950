#!/usr/bin/env python
"""Throughput of the group-granular entry points on device-resident columns (configs C1 / C3 /
C4 of BASELINE.json): FloatGroup and IntGroup encode + decode of contiguous blocks, and random
access decode of selected blocks.  Prints one JSON line per case (CUDA events on the library's
stream; inputs larger than L2 except where noted)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch
    import minnow_b200 as mb
    dev = torch.device("cuda", 0)
    ctx = mb.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    g = torch.Generator(device=dev); g.manual_seed(3)
    i64 = dict(dtype=torch.int64, device=dev)

    def timed(fn):
        best = 1e9
        for r in range(args.reps):
            a, b = ev(), ev()
            with torch.cuda.stream(stream):
                a.record(stream); fn(); b.record(stream)
            ctx.sync()
            if r:
                best = min(best, a.elapsed_time(b))
        return best

    def report(name, elems, esz, pk, ms_e, ms_d, extra=None):
        out = {"case": name, "elements": elems, "mean_bits": 8.0 * pk / elems,
               "encode_ms": ms_e, "decode_ms": ms_d,
               "encode_GBs_uncompressed": elems * esz / ms_e / 1e6, "decode_GBs_uncompressed": elems * esz / ms_d / 1e6,
               "encode_GBs_algorithmic_single_read": (elems * esz + pk) / ms_e / 1e6,
               "decode_GBs_algorithmic": (elems * esz + pk) / ms_d / 1e6}
        if extra:
            out.update(extra)
        print(json.dumps(out))

    # ---- C3-like: one minh column block = one group; here 16 blocks of 2^22 rows per call ----
    n, nb = 1 << 22, 16
    x = torch.rand(n * nb, generator=g, device=dev, dtype=torch.float32) * 125.0
    mins, bits, offs = (torch.zeros(nb, **i64) for _ in range(3))
    out_len = torch.zeros(1, **i64)
    out = torch.empty(4 * n * nb + 256, dtype=torch.uint8, device=dev)
    dec = torch.empty_like(x)
    d = mb.FloatDesc.make(0.0, 125.0, mb.float_group_pixels(0.0, 125.0, 0.001))
    jit = mb.Jitter.make(mb.JITTER_HASH, 3)
    enc = lambda: ctx.encode_float_group_dev(d, x, n, nb, mins, bits, offs, out, out.numel(), out_len)
    de = lambda: ctx.decode_float_blocks_dev(d, out, out.numel(), offs, mins, bits, n, nb, None, jit, dec)
    ms_e = timed(enc); ms_d = timed(de)
    ctx.profile(True)
    with torch.cuda.stream(stream):
        enc(); de()
    ctx.sync()
    ctx.profile(False)
    report("FloatGroup position column, 16 blocks x 2^22 float32 (125000 px)", n * nb, 4, int(out_len.item()), ms_e, ms_d,
           {"kernels": ctx.profile_summary()})

    dl = mb.FloatDesc.make(10.0, 15.0, mb.float_group_pixels(10.0, 15.0, 0.01), log10=1, clamp=1)
    m = torch.pow(10.0, 10.0 + 5.0 * torch.rand(n * nb, generator=g, device=dev, dtype=torch.float32))
    enc = lambda: ctx.encode_float_group_dev(dl, m, n, nb, mins, bits, offs, out, out.numel(), out_len)
    ms_e = timed(enc)
    report("FloatGroup log10 mass column (minh Log, clamp), 16 x 2^22", n * nb, 4, int(out_len.item()), ms_e, float("nan"))

    ids = (torch.arange(n * nb, **i64) * 3 + torch.randint(0, 3, (n * nb,), generator=g, device=dev)) + 10 ** 9
    deci = torch.empty_like(ids)
    outi = torch.empty(8 * n * nb + 256, dtype=torch.uint8, device=dev)
    enc = lambda: ctx.encode_int_group_dev(ids, n, nb, mins, bits, offs, outi, outi.numel(), out_len)
    de = lambda: ctx.decode_int_blocks_dev(outi, outi.numel(), offs, mins, bits, n, nb, None, deci)
    ms_e = timed(enc); ms_d = timed(de)
    assert torch.equal(deci, ids)
    report("IntGroup id column, 16 blocks x 2^22 int64", n * nb, 8, int(out_len.item()), ms_e, ms_d)

    # ---- C1: 2^20 halos, 16 blocks of 65536 (small: 4 MB per column, L2 resident) ----
    n1, nb1 = 65536, 16
    x1 = x[:n1 * nb1].contiguous()
    enc = lambda: ctx.encode_float_group_dev(d, x1, n1, nb1, mins, bits, offs, out, out.numel(), out_len)
    de = lambda: ctx.decode_float_blocks_dev(d, out, out.numel(), offs, mins, bits, n1, nb1, None, jit, dec)
    ms_e = timed(enc); ms_d = timed(de)
    report("C1 FloatGroup, 2^20 halos as 16 blocks x 65536 (L2 resident, launch bound)", n1 * nb1, 4, int(out_len.item()), ms_e, ms_d)

    # ---- C3: one minh block, the 30 quantised columns of the text_to_minh type menu (6 IntGroup, 12 position
    # FloatGroups, 12 log10 FloatGroups) x 2^22 rows, pinned host buffers: one mnw_encode_columns call vs 30 group calls ----
    import time
    n3 = 1 << 22
    rng = np.random.default_rng(3)
    px = mb.float_group_pixels(0.0, 125.0, 0.001)
    dpos = mb.FloatDesc.make(0.0, 125.0, px, 1, 0, 1)
    dlog = mb.FloatDesc.make(10.0, 15.0, mb.float_group_pixels(10.0, 15.0, 0.01), 1, 1, 1)
    ids0 = (np.arange(n3, dtype=np.int64) * 3 + rng.integers(0, 3, n3)) + 10 ** 9
    pos0 = (rng.random(n3) * 125.0).astype(np.float32)
    mass0 = np.power(10.0, 10.0 + 5.0 * rng.random(n3)).astype(np.float32)
    def pinned(a):   # page-locked copy: the H2D copies then run at PCIe speed
        t = torch.from_numpy(a).pin_memory()
        return t.numpy()
    cols3 = [(pinned(ids0 + k), None) for k in range(6)] + [(pinned(np.roll(pos0, k)), dpos) for k in range(12)] + \
            [(pinned(np.roll(mass0, k)), dlog) for k in range(12)]
    raw_bytes = sum(a.nbytes for a, _ in cols3)
    # straight through the C ABI with page-locked output buffers too (what a cgo caller would hold)
    import ctypes as C
    from minnow_b200.capi import Column, _ptr
    nc = len(cols3)
    stride = 8 * n3 + 16
    outp = torch.empty(nc * stride, dtype=torch.uint8).pin_memory().numpy()
    m3, b3, l3, o3 = (np.zeros(nc, np.int64) for _ in range(4))
    carr = (Column * nc)()
    parr = (C.c_void_p * nc)()
    for i, (a, d) in enumerate(cols3):
        carr[i].is_float = 0 if d is None else 1
        if d is not None: carr[i].desc = d
        parr[i] = a.ctypes.data
    def per_column():
        for i, (a, d) in enumerate(cols3):
            ln = C.c_int64(0)
            o = outp[i * stride:(i + 1) * stride]
            if d is None:
                rc = ctx.lib.mnw_encode_int_group(ctx.h, _ptr(a), n3, 1, None, _ptr(m3[i:]), _ptr(b3[i:]), _ptr(o3[i:]), _ptr(o), stride, C.byref(ln))
            else:
                rc = ctx.lib.mnw_encode_float_group(ctx.h, C.byref(d), _ptr(a), n3, 1, None, _ptr(m3[i:]), _ptr(b3[i:]), _ptr(o3[i:]), _ptr(o), stride, C.byref(ln))
            assert rc == 0
    def batched():
        assert ctx.lib.mnw_encode_columns(ctx.h, nc, carr, parr, n3, _ptr(m3), _ptr(b3), _ptr(l3), _ptr(outp), stride) == 0
    for name, fn in (("30 group calls", per_column), ("one mnw_encode_columns call", batched)):
        fn()
        t0 = time.perf_counter()
        for _ in range(5): fn()
        dt = (time.perf_counter() - t0) / 5
        print(json.dumps({"case": "C3 minh block, 30 quantised columns x 2^22 rows, pinned host buffers, C ABI: " + name,
                          "ms": dt * 1e3, "GBs_uncompressed_e2e": raw_bytes / dt / 1e9, "packed_MB": float(l3.sum()) / 1e6}))

    # ---- C4: random access, 10^4 selected blocks of 4096 values out of 98304 ----
    n4, nb4, nsel = 4096, 3 * 32768, 10000
    x4 = (torch.rand(n4 * nb4, generator=g, device=dev, dtype=torch.float32) * 0.3 +
          torch.arange(nb4, device=dev, dtype=torch.float32).repeat_interleave(n4) * (1000.0 / nb4)) % 1000.0
    mins4, bits4, offs4 = (torch.zeros(nb4, **i64) for _ in range(3))
    d4 = mb.FloatDesc.make(0.0, 1000.0, mb.float_group_pixels(0.0, 1000.0, 0.005))
    out4 = torch.empty(4 * n4 * nb4 + 256, dtype=torch.uint8, device=dev)
    with torch.cuda.stream(stream):
        ctx.encode_float_group_dev(d4, x4, n4, nb4, mins4, bits4, offs4, out4, out4.numel(), out_len)
    ctx.sync()
    sel = torch.randperm(nb4, generator=g, device=dev)[:nsel].contiguous()
    dec4 = torch.empty(n4 * nsel, dtype=torch.float32, device=dev)
    lat = []
    for r in range(101):
        a, b = ev(), ev()
        with torch.cuda.stream(stream):
            a.record(stream)
            ctx.decode_float_blocks_dev(d4, out4, out4.numel(), offs4, mins4, bits4, n4, nsel, sel, jit, dec4)
            b.record(stream)
        ctx.sync()
        if r:
            lat.append(a.elapsed_time(b) * 1e3)
    lat = np.array(lat)
    print(json.dumps({"case": "C4 random access: 10^4 blocks of 4096 out of 98304 (512^3, SubCells 32)", "batch_latency_us_p50": float(np.percentile(lat, 50)),
                      "batch_latency_us_p99": float(np.percentile(lat, 99)), "decoded_GBs": 4.0 * n4 * nsel / np.percentile(lat, 50) / 1e3,
                      "mean_bits": float(bits4.double().mean().item())}))
    ctx.close()


if __name__ == "__main__":
    main()
