import sys, os
sys.path.insert(0, "/root/repo")
import torch, numpy as np
import minnow_b200 as mb
dev = torch.device("cuda", 0)
ctx = mb.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
g = torch.Generator(device=dev); g.manual_seed(1)
CASES = ((256, 8, 16), (256, 16, 16), (256, 4, 16), (256, 2, 16))
if os.environ.get("NFILES"):   # e.g. NFILES=64 SUBCELLS=4: the headline shape
    CASES = ((256, int(os.environ.get("SUBCELLS", "4")), int(os.environ["NFILES"])),)
for nfile, subcells, nfiles in CASES:
    n3 = nfile ** 3
    L, dx = (1000.0 if nfile // subcells < 128 else 250.0), 0.005   # 128^3 sub-cells: a box in which they stay below 17 bits
    j = torch.arange(nfile, device=dev, dtype=torch.float32) * (L / nfile)
    grid = torch.stack(torch.meshgrid(j, j, j, indexing="ij")[::-1], dim=-1).reshape(n3, 3)
    if os.environ.get("BENCHLIKE"):   # bench.py's files (14-bit blocks)
        import bench
        aos = torch.empty((nfiles, n3, 3), dtype=torch.float32, device=dev)
        for f in range(nfiles):
            aos[f].copy_(bench.gen_file(torch, f, 2, dev)[0])
    else:
        aos = torch.remainder(torch.randn((nfiles, n3, 3), generator=g, device=dev) * 2.0 + grid[None], L).contiguous()
        aos[aos >= L] = 0
    px = mb.float_group_pixels(0.0, L, dx)
    descs = [mb.FloatDesc.make(0.0, L, px) for _ in range(3)]
    sc3 = subcells ** 3
    nb = nfiles * 3 * sc3
    stride = 4 * n3 + 256
    i64 = dict(dtype=torch.int64, device=dev)
    mins, bits, offs = (torch.zeros(nb, **i64) for _ in range(3))
    out_len = torch.zeros(3 * nfiles, **i64)
    out = torch.empty(3 * nfiles * stride, dtype=torch.uint8, device=dev)
    dec = torch.empty_like(aos)
    jit = mb.Jitter.make(mb.JITTER_HASH, 7)
    best = [1e9, 1e9]
    for rep in range(4):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        with torch.cuda.stream(stream):
            e[0].record(stream)
            ctx.encode_vec3_subcells_dev(descs, aos, nfile, subcells, nfiles, mins, bits, offs, out, stride, out_len)
            e[1].record(stream)
            ctx.decode_vec3_subcells_dev(descs, out, stride, offs, mins, bits, nfile, subcells, nfiles, L, jit, dec)
            e[2].record(stream)
        ctx.sync()
        if rep:
            best[0] = min(best[0], e[0].elapsed_time(e[1])); best[1] = min(best[1], e[1].elapsed_time(e[2]))
    gb = 12 * n3 * nfiles / 1e9
    mb_ = bits.double().mean().item()
    alg = gb * (1 + mb_ / 32)
    print("nsub %d x %d files: encode %.3f ms (%.0f GB/s of input, %.0f algorithmic)  decode %.3f ms (%.0f GB/s of output, %.0f algorithmic)  mean bits %.2f path %d" %
          (nfile // subcells, nfiles, best[0], gb / best[0] * 1e3, alg / best[0] * 1e3, best[1], gb / best[1] * 1e3, alg / best[1] * 1e3, mb_, ctx.last_path))
    del aos, dec, out, grid
