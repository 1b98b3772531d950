#!/usr/bin/env python
"""One group-encode case for ncu (k_group_fused): python tools/prof_groups.py f32|log|i64 [reps] [nblocks] [log2 n]."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    case = sys.argv[1] if len(sys.argv) > 1 else "f32"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    nb = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    n = 1 << (int(sys.argv[4]) if len(sys.argv) > 4 else 22)
    import torch
    import minnow_b200 as mb
    dev = torch.device("cuda", 0)
    ctx = mb.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    g = torch.Generator(device=dev); g.manual_seed(3)
    i64 = dict(dtype=torch.int64, device=dev)
    mins, bits, offs = (torch.zeros(nb, **i64) for _ in range(3))
    out_len = torch.zeros(1, **i64)
    out = torch.empty(4 * n * nb + 256, dtype=torch.uint8, device=dev)
    if case == "i64":
        x = torch.randint(10 ** 9, 10 ** 9 + (1 << 24), (n * nb,), generator=g, device=dev, dtype=torch.int64)
        enc = lambda: ctx.encode_int_group_dev(x, n, nb, mins, bits, offs, out, out.numel(), out_len)
        esz = 8
    elif case == "log":
        x = torch.pow(10.0, 10.0 + 5.0 * torch.rand(n * nb, generator=g, device=dev, dtype=torch.float32))
        d = mb.FloatDesc.make(10.0, 15.0, mb.float_group_pixels(10.0, 15.0, 0.01), 1, 1, 1)
        enc = lambda: ctx.encode_float_group_dev(d, x, n, nb, mins, bits, offs, out, out.numel(), out_len)
        esz = 4
    else:
        x = torch.rand(n * nb, generator=g, device=dev, dtype=torch.float32) * 125.0
        d = mb.FloatDesc.make(0.0, 125.0, mb.float_group_pixels(0.0, 125.0, 0.001))
        enc = lambda: ctx.encode_float_group_dev(d, x, n, nb, mins, bits, offs, out, out.numel(), out_len)
        esz = 4
    best = 1e9
    ctx.profile(True)
    for r in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            a.record(stream); enc(); b.record(stream)
        ctx.sync()
        if r:
            best = min(best, a.elapsed_time(b))
    pk = int(out_len.item())
    print(json.dumps({"case": case, "n": n, "nblocks": nb, "ms": best, "packed": pk, "mean_bits": 8.0 * pk / (n * nb),
                      "GBs_uncompressed": n * nb * esz / best / 1e6, "GBs_algorithmic": (n * nb * esz + pk) / best / 1e6,
                      "frac_of_6548": (n * nb * esz + pk) / best / 1e6 / 6548.2, "kernels": ctx.profile_summary()}))


if __name__ == "__main__":
    main()
