#!/usr/bin/env python
"""Small driver for ncu: the bench's minp workload on a few files only.
  python tools/prof_vec3.py [--nfiles 8] [--reps 3]
Prints per-phase CUDA-event times (not a benchmark number)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nfiles", type=int, default=8)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import minnow_b200 as mb
    dev = torch.device("cuda", 0)
    ctx = mb.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    nf = args.nfiles
    pos = torch.empty((nf, B.NP_FILE, 3), dtype=torch.float32, device=dev)
    vel = torch.empty((nf, B.NP_FILE, 3), dtype=torch.float32, device=dev)
    for f in range(nf):
        p, v = B.gen_file(torch, f, 2, dev)
        pos[f].copy_(p); vel[f].copy_(v)
    nb = nf * 3 * B.SC3
    stride = 4 * B.NFILE ** 3 + 256
    i64 = dict(dtype=torch.int64, device=dev)
    meta = {k: [torch.zeros(nb, **i64) for _ in range(3)] for k in "xv"}
    out_len = {k: torch.zeros(3 * nf, **i64) for k in "xv"}
    packed = {k: torch.empty(3 * nf * stride, dtype=torch.uint8, device=dev) for k in "xv"}
    decoded = torch.empty((nf, B.NP_FILE, 3), dtype=torch.float32, device=dev)
    ppx = mb.float_group_pixels(0.0, B.L_BOX, B.DX_POS)
    pdescs = [mb.FloatDesc.make(0.0, B.L_BOX, ppx) for _ in range(3)]
    lo, hi = ctx.vec3_limits(vel, nf, dev=True)
    vd = [mb.FloatDesc.make(lo[f, k], hi[f, k], mb.float_group_pixels(lo[f, k], hi[f, k], B.DV)) for f in range(nf) for k in range(3)]
    jit = mb.Jitter.make(mb.JITTER_HASH, 7)
    torch.cuda.synchronize()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    for rep in range(args.reps):
        with torch.cuda.stream(stream):
            e = [ev() for _ in range(5)]
            e[0].record(stream)
            ctx.encode_vec3_subcells_dev(pdescs, pos, B.NFILE, B.SUB_CELLS, nf, *meta["x"], packed["x"], stride, out_len["x"])
            e[1].record(stream)
            ctx.encode_vec3_subcells_dev(vd, vel, B.NFILE, B.SUB_CELLS, nf, *meta["v"], packed["v"], stride, out_len["v"])
            e[2].record(stream)
            ctx.decode_vec3_subcells_dev(pdescs, packed["x"], stride, meta["x"][2], meta["x"][0], meta["x"][1], B.NFILE, B.SUB_CELLS, nf, B.L_BOX, jit, decoded)
            e[3].record(stream)
            ctx.decode_vec3_subcells_dev(vd, packed["v"], stride, meta["v"][2], meta["v"][0], meta["v"][1], B.NFILE, B.SUB_CELLS, nf, 0.0, jit, decoded)
            e[4].record(stream)
        ctx.sync()
        t = [e[i].elapsed_time(e[i + 1]) for i in range(4)]
        gb = 12 * B.NP_FILE * nf / 1e9
        print("rep %d: enc_x %.3f ms (%.0f GB/s)  enc_v %.3f ms (%.0f GB/s)  dec_x %.3f ms (%.0f GB/s)  dec_v %.3f ms (%.0f GB/s)  path=%d"
              % (rep, t[0], gb / t[0] * 1e3, t[1], gb / t[1] * 1e3, t[2], gb / t[2] * 1e3, t[3], gb / t[3] * 1e3, ctx.last_path))
    print("mean bits x %.2f v %.2f" % (meta["x"][1].double().mean().item(), meta["v"][1].double().mean().item()))
    ctx.close()


if __name__ == "__main__":
    main()
