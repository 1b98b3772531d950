#!/usr/bin/env python
"""The C1 / C3 / C4 configurations of bench.py alone (tuning runs): python tools/run_configs.py [C1 C3 C4]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    import bench_configs
    import minnow_b200 as mb
    from oracle import oracle as orc
    orc.lib()
    dev = torch.device("cuda", 0)
    ctx = mb.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    peak = bench.measured_peaks()["hbm_gbs"] if hasattr(bench, "measured_peaks") else 6548.2
    want = sys.argv[1:] or ["C1", "C3", "C4"]
    for name, fn in (("C1", bench_configs.run_c1), ("C3", bench_configs.run_c3), ("C4", bench_configs.run_c4)):
        if name not in want:
            continue
        r = fn(torch, mb, orc, ctx, stream, dev, peak, bench.host_threads())
        print(name, json.dumps({k: r[k] for k in r if k in ("value", "ms", "encode_ms", "decode_ms", "verified")}),
              r["roofline"]["kernel"], round(r["roofline"]["frac"], 3), "e2e", round(r["e2e"]["value"], 1), flush=True)


if __name__ == "__main__":
    main()
