#!/usr/bin/env python
"""Why is k_decode_vec3 slower inside bench.py than alone?  The headline shape (64 files of 256^3, bench.py's positions)
decoded under several conditions; per-launch kernel times from the library's own CUDA events (ctx.profile).

python tools/probe_decode.py [variants ...]   variants: base extra sampler devapi encfirst
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    variants = sys.argv[1:] or ["base", "devapi", "encfirst", "sampler", "extra"]
    import torch
    import bench
    import minnow_b200 as mb
    dev = torch.device("cuda", 0)
    ctx = mb.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    nfile, subcells, nfiles = 256, 4, 64
    n3 = nfile ** 3
    L, dx = 1000.0, 0.005
    aos = torch.empty((nfiles, n3, 3), dtype=torch.float32, device=dev)
    for f in range(nfiles):
        aos[f].copy_(bench.gen_file(torch, f, 2, dev)[0])
    px = mb.float_group_pixels(0.0, L, dx)
    descs = [mb.FloatDesc.make(0.0, L, px) for _ in range(3)]
    nb = nfiles * 3 * subcells ** 3
    stride = 4 * n3 + 256
    i64 = dict(dtype=torch.int64, device=dev)
    mins, bits, offs = (torch.zeros(nb, **i64) for _ in range(3))
    out_len = torch.zeros(3 * nfiles, **i64)
    out = torch.empty(3 * nfiles * stride, dtype=torch.uint8, device=dev)
    dec = torch.empty_like(aos)
    desc_dev = torch.zeros(24 * 3 * nfiles, dtype=torch.uint8, device=dev)
    jit = mb.Jitter.make(mb.JITTER_HASH, 7)

    def enc():
        ctx.minp_encode_vectors_dev(aos, nfile, subcells, nfiles, True, L, dx, desc_dev, mins, bits, offs, out, stride, out_len)

    def dec_host():
        ctx.decode_vec3_subcells_dev(descs, out, stride, offs, mins, bits, nfile, subcells, nfiles, L, jit, dec)

    def dec_dev():
        ctx.minp_decode_vectors_dev(desc_dev, out, stride, offs, mins, bits, nfile, subcells, nfiles, True, L, jit, dec)

    with torch.cuda.stream(stream):
        enc()
    ctx.sync()

    def run(name, fn, reps=10, before=None):
        with torch.cuda.stream(stream):
            for _ in range(2):
                fn()
        ctx.sync()
        ctx.profile(True)
        with torch.cuda.stream(stream):
            for _ in range(reps):
                if before:
                    before()
                fn()
        ctx.sync()
        ks = {k["kernel"]: k for k in ctx.profile_summary()}
        ctx.profile(False)
        k = ks["k_decode_vec3"]
        print("%-10s k_decode_vec3 %d launches, mean %.3f ms" % (name, k["launches"], k["ms"] / k["launches"]), flush=True)

    extra = None
    for v in variants:
        if v == "base":
            run("base", dec_host)
        elif v == "devapi":
            run("devapi", dec_dev)
        elif v == "encfirst":
            run("encfirst", dec_dev, before=enc)
        elif v == "sampler":
            s = bench.ClockSampler(bench.physical_gpu_index(0))
            s.start()
            time.sleep(0.05)
            run("sampler", dec_dev)
            s.stop_flag.set()
            s.join()
        elif v == "extra":
            extra = [torch.zeros(13 * 2 ** 30, dtype=torch.uint8, device=dev) for _ in range(2)]
            torch.cuda.synchronize()
            run("extra26GB", dec_dev)
            run("extra+enc", dec_dev, before=enc)
    del extra


if __name__ == "__main__":
    main()
