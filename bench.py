#!/usr/bin/env python
"""bench.py -- encode & decode GB/s (uncompressed) of the minnow block hot path.

Workload (BASELINE.json configs[1]): a minp particle snapshot of 1024^3
particles, positions + velocities, as 64 files of 256^3 (FileCells = 4), each
file split into 4^3 sub-cells of 64^3 (SubCells = 4): 3 FloatGroups x 64 blocks
per file and field.  Positions: periodic [0, 1000) Mpc/h, dx = 0.005 (200 000
pixels).  Velocities: per-file limits [min, nextafter(max)], dv = 1 km/s.
One step = encode(x) + encode(v) + decode(x) + decode(v) of the whole snapshot
on every rank (weak scaling: each rank holds its own 1024^3 block range).

  python bench.py [--gpus N] [--steps K] [--warmup W]      (N>1: under torchrun)
  python bench.py --impl reference ...                     CPU oracle on host cores

Prints ONE JSON line (see the contract in DESIGN.md, "Measurement").
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "encode+decode GB/s (uncompressed)"
L_BOX, DX_POS, DV = 1000.0, 0.005, 1.0
NSIDE, FILE_CELLS, SUB_CELLS = 1024, 4, 4
NFILE = NSIDE // FILE_CELLS            # 256
NFILES = FILE_CELLS ** 3               # 64
SC3 = SUB_CELLS ** 3                   # 64
NSUB3 = (NFILE // SUB_CELLS) ** 3      # 262144
NP_FILE = NFILE ** 3


def workload_config(extra=None):
    cfg = {"workload": "minp 1024^3 positions+velocities: 64 files x 256^3 (FileCells=4), SubCells=4 -> "
                       "3x64 blocks of 64^3 per file and field; x periodic [0,1000) dx=0.005 (200000 px), "
                       "v per-file limits dv=1; one step = encode(x,v) + decode(x,v)",
           "particles_per_gpu": NSIDE ** 3, "uncompressed_bytes_per_step_per_gpu": 4 * 12 * NSIDE ** 3,
           "cache": "inputs (25.8 GB per GPU) are far larger than L2 (126 MB); no flush needed",
           "jitter": "hash (one random sub-pixel offset per decoded value, as the reference draws rand.Float64)"}
    if extra:
        cfg.update(extra)
    return cfg


def gen_file(torch, f, seed, device):
    """Synthetic Lagrangian file f: positions = grid + Irwin-Hall(4) displacement
    (sigma 2 Mpc/h) wrapped into [0, L); velocities = 300 km/s x Irwin-Hall(4)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed * 100003 + f)
    fx, fy, fz = f % FILE_CELLS, (f // FILE_CELLS) % FILE_CELLS, f // (FILE_CELLS * FILE_CELLS)
    j = torch.arange(NFILE, device=device, dtype=torch.float32)
    cell = L_BOX / NSIDE
    gx = ((fx * NFILE + j) * cell).view(1, 1, NFILE).expand(NFILE, NFILE, NFILE)
    gy = ((fy * NFILE + j) * cell).view(1, NFILE, 1).expand(NFILE, NFILE, NFILE)
    gz = ((fz * NFILE + j) * cell).view(NFILE, 1, 1).expand(NFILE, NFILE, NFILE)
    grid = torch.stack([gx, gy, gz], dim=-1).reshape(NP_FILE, 3)

    def irwin_hall():
        u = torch.rand((4, NP_FILE, 3), generator=g, device=device, dtype=torch.float32)
        return (u.sum(0) - 2.0) * (1.0 / 0.5773502691896257)   # unit variance

    pos = torch.remainder(grid + 2.0 * irwin_hall(), L_BOX)
    pos = torch.where(pos >= L_BOX, torch.zeros_like(pos), pos).contiguous()
    vel = (300.0 * irwin_hall()).contiguous()
    return pos, vel


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region."""

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period, self.samples, self.reasons = index, period, [], set()
        self.stop_flag = threading.Event()
        self.max_mhz, self.ok = None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = str(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
                 "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80)}
        while not self.stop_flag.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def result(self):
        self.stop_flag.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ---------------------------------------------------------------------------------------
# CPU legs (oracle = port of the reference; test infrastructure, only ever the thing timed
# here as the BASELINE, never as the product)
# ---------------------------------------------------------------------------------------
def host_threads():
    """All the cores this process may run on (torchrun exports OMP_NUM_THREADS=1, which is not
    what the CPU legs should be limited to)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_descs(orc, vel_np):
    lo, hi = orc.minp_limits(vel_np, False, L_BOX)
    vpx = [orc.float_group_pixels(float(lo[k]), float(hi[k]), np.float32(DV)) for k in range(3)]
    ppx = orc.float_group_pixels(0.0, np.float32(L_BOX), np.float32(DX_POS))
    return ([0.0] * 3, [L_BOX] * 3, [ppx] * 3), (lo.tolist(), hi.tolist(), vpx)


def cpu_step(orc, pos_np, vel_np, threads):
    """encode + decode of ONE file (positions and velocities) with the oracle."""
    pd, vd = cpu_descs(orc, vel_np)
    for arr, (lo, hi, px), per in ((pos_np, pd, True), (vel_np, vd, False)):
        mins, bits, nbytes, packed, stride, total = orc.bench_minp_encode(arr, NFILE, SUB_CELLS, lo, hi, px, threads)
        orc.bench_minp_decode(packed, stride, NFILE, SUB_CELLS, lo, hi, px, mins, bits, per, L_BOX, 1, 7, threads)
    return 4 * 12 * NP_FILE   # uncompressed bytes through encode + decode


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; Go is not in
    this image) on all host cores, bounded sample = one file per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import oracle as orc
    orc.lib()
    threads = host_threads()
    pos, vel = gen_file(torch, 0, 2, "cpu")
    pos_np, vel_np = pos.numpy(), vel.numpy()
    for _ in range(args.warmup):
        cpu_step(orc, pos_np, vel_np, threads)
    t0 = time.perf_counter()
    nbytes = 0
    for _ in range(args.steps):
        nbytes += cpu_step(orc, pos_np, vel_np, threads)
    dt = time.perf_counter() - t0
    val = nbytes / dt / 1e9
    sample = "1 of 64 files (256^3 particles, x and v) per step, encode+decode, OpenMP over blocks"
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": "GB/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32->i64 (f64 floor)",
           "data": "synthetic", "config": workload_config(),
           "cpu_baseline": {"value": val, "unit": "GB/s", "cores": threads, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


# ---------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    import minnow_b200 as mb
    from minnow_b200 import shard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    ctx = mb.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    if world > 1:   # the library's own communicator (mnw_comm_init): the id travels over the launcher's process group
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(mb.Context.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.comm_init(bytes(uid.cpu().numpy().tobytes()), world, rank)

    # ---- synthetic snapshot, resident in HBM -------------------------------------------
    pos = torch.empty((NFILES, NP_FILE, 3), dtype=torch.float32, device=dev)
    vel = torch.empty((NFILES, NP_FILE, 3), dtype=torch.float32, device=dev)
    for f in range(NFILES):
        p, v = gen_file(torch, f, 2 + rank, dev)
        pos[f].copy_(p)
        vel[f].copy_(v)
        del p, v
    torch.cuda.synchronize()

    nb = NFILES * 3 * SC3                  # blocks per field
    stride = 4 * NFILE ** 3 + 256          # bytes reserved per (file, axis) stream (<= 32 bits/value)
    i64 = dict(dtype=torch.int64, device=dev)
    meta = {k: [torch.zeros(nb, **i64) for _ in range(3)] for k in ("x", "v")}     # mins, bits, offsets
    out_len = {k: torch.zeros(3 * NFILES, **i64) for k in ("x", "v")}
    packed = {k: torch.empty(3 * NFILES * stride, dtype=torch.uint8, device=dev) for k in ("x", "v")}
    decoded = torch.empty((NFILES, NP_FILE, 3), dtype=torch.float32, device=dev)

    ppx = mb.float_group_pixels(0.0, L_BOX, DX_POS)
    pdescs = [mb.FloatDesc.make(0.0, L_BOX, ppx) for _ in range(3)]
    jit = mb.Jitter.make(mb.JITTER_HASH, 7)
    state = {}

    def vel_descs(lo, hi):
        return [mb.FloatDesc.make(lo[f, k], hi[f, k], mb.float_group_pixels(lo[f, k], hi[f, k], DV))
                for f in range(NFILES) for k in range(3)]

    ev = lambda: torch.cuda.Event(enable_timing=True)
    marks = []

    desc_dev = {k: torch.zeros(24 * 3 * NFILES, dtype=torch.uint8, device=dev) for k in ("x", "v")}   # mnw_float_desc [3 * NFILES]
    all_sizes = torch.zeros(world * 2 * nb, **i64)
    all_offs = torch.zeros(world * 2 * nb, **i64)
    all_total = torch.zeros(1, **i64)

    def step(record=False):
        """One step, enqueued without a host round trip: limits, pixel counts and group constants of the velocity field
        are derived on the device (mnw_minp_encode_vectors_dev)."""
        with torch.cuda.stream(stream):
            e = [ev() for _ in range(5)] if record else None
            if record: e[0].record(stream)
            ctx.minp_encode_vectors_dev(pos, NFILE, SUB_CELLS, NFILES, True, L_BOX, DX_POS, desc_dev["x"], *meta["x"], packed["x"], stride,
                                        out_len["x"])
            if record: e[1].record(stream)
            ctx.minp_encode_vectors_dev(vel, NFILE, SUB_CELLS, NFILES, False, 0.0, DV, desc_dev["v"], *meta["v"], packed["v"], stride,
                                        out_len["v"])
            if record: e[2].record(stream)
            if world > 1:
                # the one exchange of the sharded path, through the C ABI: NCCL all-gather of the per-block packed sizes,
                # then the same scan on every rank -> global byte offsets of the whole snapshot
                local = torch.cat([shard.packed_sizes(meta[k][1], NSUB3) for k in ("x", "v")])
                ctx.sharded_offsets_dev(local, 2 * nb, all_sizes, all_offs, all_total)
            ctx.minp_decode_vectors_dev(desc_dev["x"], packed["x"], stride, meta["x"][2], meta["x"][0], meta["x"][1],
                                        NFILE, SUB_CELLS, NFILES, True, L_BOX, jit, decoded)
            if record: e[3].record(stream)
            ctx.minp_decode_vectors_dev(desc_dev["v"], packed["v"], stride, meta["v"][2], meta["v"][0], meta["v"][1],
                                        NFILE, SUB_CELLS, NFILES, False, 0.0, jit, decoded)
            if record: e[4].record(stream)
            if record: marks.append(e)

    def host_descs(key):
        a = np.frombuffer(desc_dev[key].cpu().numpy().tobytes(), np.dtype([("low", "<f4"), ("high", "<f4"), ("pixels", "<i8"), ("fl", "u1", 8)]))
        return [mb.FloatDesc.make(float(r["low"]), float(r["high"]), int(r["pixels"])) for r in a]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()

    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    launches0 = ctx.launch_count
    ctx.profile(True)
    start, stop = ev(), ev()
    start.record(stream)
    for _ in range(args.steps):
        step(record=True)
    stop.record(stream)
    barrier()
    ctx.profile(False)
    clocks = sampler.result()
    launches = ctx.launch_count - launches0
    ms_total = start.elapsed_time(stop)
    prof = ctx.profile_summary()

    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    bytes_step = 4 * 12 * NSIDE ** 3                      # per GPU: encode in (x, v) + decode out (x, v)
    value = world * bytes_step / (ms_step * 1e-3) / 1e9

    # per-phase times (rank 0), device events on the launching stream
    ph = np.array([[m[i].elapsed_time(m[i + 1]) for i in range(4)] for m in marks]).mean(0)
    field_bytes = 12 * NSIDE ** 3
    mean_bits = {k: float(meta[k][1].double().mean().item()) for k in ("x", "v")}
    packed_bytes = {k: int(out_len[k].sum().item()) for k in ("x", "v")}
    # encode_x_ms includes the velocity limits kernel (bounds() runs first, see step())
    phases = {"encode_x_ms": ph[0], "encode_v_ms": ph[1], "decode_x_ms": ph[2], "decode_v_ms": ph[3],
              "encode_gbs": 2 * field_bytes / ((ph[0] + ph[1]) * 1e-3) / 1e9,
              "decode_gbs": 2 * field_bytes / ((ph[2] + ph[3]) * 1e-3) / 1e9,
              "mean_bits_x": mean_bits["x"], "mean_bits_v": mean_bits["v"],
              "packed_bytes_x": packed_bytes["x"], "packed_bytes_v": packed_bytes["v"],
              "encode_path": "fused" if ctx.last_path == 1 else "generic-two-pass"}

    # ---- roofline of the dominant kernel -------------------------------------------------
    peaks = {"hbm_gbs": 6650.0, "src": "fallback"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = {"hbm_gbs": float(json.load(f)["hbm_gbs"]), "src": "measured"}
    except Exception:
        pass
    pk_total = packed_bytes["x"] + packed_bytes["v"]
    algo = {  # algorithmic bytes per step, summed over both fields (DESIGN.md "Kernels")
        "k_stats": 2 * field_bytes, "k_pack": 2 * field_bytes + pk_total,
        "k_fused_vec3": 2 * field_bytes + pk_total, "k_pipe_vec3": 2 * field_bytes + pk_total, "k_decode": pk_total + 2 * field_bytes,
        "k_decode_vec3": pk_total + 2 * field_bytes, "k_vec3_limits4": field_bytes}
    traffic = {}
    try:   # DRAM bytes per particle and launch from the committed ncu --set full capture (profiles/)
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f)
    except Exception:
        pass
    roof = None
    if prof:
        top = max(prof, key=lambda r: r["ms"])
        per_launch_ms = top["ms"] / top["launches"]
        launches_per_step = top["launches"] / args.steps
        a = algo.get(top["kernel"], 0) / launches_per_step / (per_launch_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": top["kernel"], "achieved": a, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": a / peaks["hbm_gbs"], "peak_source": peaks["src"] + " (MEASURED_PEAKS.json copy bandwidth)",
                "traffic": (traffic[top["kernel"]]["dram_bytes_per_particle"] * NSIDE ** 3
                            if top["kernel"] in traffic else None),
                "traffic_source": traffic.get(top["kernel"], {}).get("source"),
                "ms_per_launch": per_launch_ms,
                "algorithmic_bytes_per_launch": algo.get(top["kernel"], 0) / launches_per_step,
                "kernels": [dict(r, share=r["ms"] / (ms_total),
                                 achieved_gbs=(algo.get(r["kernel"], 0) * args.steps / (r["ms"] * 1e-3) / 1e9 if r["ms"] else None))
                            for r in prof]}

    # ---- size-independent check of what the timed steps left behind (not timed): every block's byte offset is the
    # running sum of ArrayBytes(bits, n) within its (file, axis) stream, and a decode with the CENTER jitter lands
    # within half a pixel (plus float32 rounding) of every input value
    verified = None
    if rank == 0:
        try:
            verified = verify_roundtrip(torch, mb, ctx, stream, dev, pos, vel, pdescs, host_descs("v"), packed, meta, out_len,
                                        decoded, stride)
        except Exception as exc:   # a failed check is reported, it never costs the run its number
            verified = {"ok": False, "error": "%s: %s" % (type(exc).__name__, exc)}

    # ---- C5: the ranks' files as ONE sharded snapshot (not timed): the global offsets are consumed ---------------------
    c5 = None
    if world > 1:
        try:
            c5 = c5_check(torch, dist, mb, ctx, stream, dev, rank, world, nb, stride, meta, out_len, packed, host_descs("x"),
                          all_sizes, all_offs, all_total)
        except Exception as exc:
            c5 = {"ok": False, "error": "%s: %s" % (type(exc).__name__, exc)}
        if verified is not None:
            verified["c5_sharded_snapshot"] = c5

    # ---- end to end through the host-pointer C ABI (pinned host buffers, copies timed) ------
    e2e = None if args.no_e2e else run_e2e(torch, mb, ctx, pos, vel, pdescs, world, args, dev)

    # ---- CPU baseline beside it (rank 0, N = 1): the oracle port on the host cores ----------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as orc
        orc.lib()
        threads = host_threads()
        p0, v0 = pos[0].cpu().numpy(), vel[0].cpu().numpy()
        cpu_step(orc, p0, v0, threads)
        t0 = time.perf_counter()
        reps, nby = 0, 0
        while reps < 2 or time.perf_counter() - t0 < 4.0:
            nby += cpu_step(orc, p0, v0, threads)
            reps += 1
        dt = time.perf_counter() - t0
        t1 = time.perf_counter()
        nby1 = cpu_step(orc, p0, v0, 1)          # the reference is single-threaded: its own figure is the 1-core one
        dt1 = time.perf_counter() - t1
        import bench_configs
        cpu = {"value": nby / dt / 1e9, "unit": "GB/s", "cores": threads, "kind": "port",
               "one_core_value": nby1 / dt1 / 1e9, "cpu": bench_configs.cpu_model(),
               "sample": "file 0 of 64 (256^3 particles, x and v), encode+decode, %d repetitions, OpenMP over blocks "
                         "(+ one repetition on 1 core)" % reps}

    configs = None
    if rank == 0 and world == 1 and not args.no_configs:
        import bench_configs
        from oracle import oracle as orc
        orc.lib()
        del decoded
        torch.cuda.empty_cache()
        configs = bench_configs.run_all(torch, mb, orc, ctx, stream, dev, peaks["hbm_gbs"], host_threads())

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f32->i64 (f64 floor, as the reference)", "data": "synthetic",
               "config": workload_config({"sharding": "block ranges per GPU; NCCL all-gather of per-block sizes + "
                                          "offset scan" if world > 1 else "single GPU"}),
               "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "phases": phases,
               "roofline": roof, "cpu_baseline": cpu, "verified": verified, "configs": configs}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


def c5_check(torch, dist, mb, ctx, stream, dev, rank, world, nb, stride, meta, out_len, packed, xdescs, all_sizes, all_offs, all_total):
    """The snapshot-level block index of the sharded path, consumed (go/block_index.go:16-35, go/writer.go:84-86):
    1. the global offsets every rank derived (mnw_sharded_offsets_dev: NCCL all-gather + scan) equal ONE scan of the
       gathered sizes, and the gathered sizes equal what torch.distributed gathers;
    2. every rank writes the packed x bytes of its file 0 into one shared snapshot file at groupOffset + base_r
       (base_r = offset of its first block in the global index), each (file, axis) group at its global offset;
    3. every rank reads the NEXT rank's file 0 back from those offsets, assembles a minp file around the bytes with the
       host mirror (minp.Writer.EncodedVectors) and decodes it with minp.Reader: the particles are that rank's."""
    import io
    from minnow_b200 import minp, shard
    ctx.sync()
    sizes, offs = all_sizes.clone(), all_offs.clone()
    local = torch.cat([shard.packed_sizes(meta[k][1], NSUB3) for k in ("x", "v")])
    ref = torch.empty(world * 2 * nb, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(ref, local.contiguous())
    ok_gather = bool(torch.equal(ref, sizes))
    inc = torch.cumsum(sizes, 0)
    ok_scan = bool(torch.equal(offs, inc - sizes)) and int(all_total.item()) == int(inc[-1].item())
    # the metadata of file 0 (x) of every rank: 3 * SC3 (min, bits) pairs -- gathered with the library's collective too
    m0 = torch.stack([meta["x"][0][:3 * SC3], meta["x"][1][:3 * SC3]]).reshape(-1).contiguous()
    mall = torch.empty(world * m0.numel(), dtype=torch.int64, device=dev)
    with torch.cuda.stream(stream):
        ctx.allgather_sizes(m0, m0.numel(), mall)
    ctx.sync()
    path = "/dev/shm/minnow_b200_c5_snapshot.bin"
    base_r = int(offs[rank * 2 * nb].item())
    fd = os.open(path, os.O_RDWR | os.O_CREAT, 0o600)
    for k in range(3):                      # file 0, field x: block (k, sc) is global block rank * 2nb + k * SC3 + sc
        g0 = rank * 2 * nb + k * SC3
        ln = int(out_len["x"][k].item())
        assert ln == int(sizes[g0:g0 + SC3].sum().item())
        os.pwrite(fd, packed["x"][k * stride:k * stride + ln].cpu().numpy().tobytes(), int(offs[g0].item()))
    os.fsync(fd)
    dist.barrier()
    nbr = (rank + 1) % world
    mn = mall.view(world, 2, 3 * SC3)[nbr]
    streams = []
    for k in range(3):
        g0 = nbr * 2 * nb + k * SC3
        streams.append(os.pread(fd, int(sizes[g0:g0 + SC3].sum().item()), int(offs[g0].item())))
    os.close(fd)
    hd = np.zeros(1, minp.Header)
    hd["NSide"], hd["L"] = NSIDE, L_BOX
    cell = np.zeros(1, minp.Cell)
    cell["FileCells"], cell["SubCells"] = FILE_CELLS, SUB_CELLS
    buf = io.BytesIO()
    w = minp.Create(buf, ctx)
    w.Header(hd, b"", cell, DX_POS, True)
    w.EncodedVectors(xdescs[:3], mn[0].cpu().numpy(), mn[1].cpu().numpy(), streams)
    w.Close()
    got = torch.from_numpy(minp.Open(buf.getvalue(), ctx).Vectors()).to(dev)           # CENTER jitter
    want, _ = gen_file(torch, 0, 2 + nbr, dev)
    d = (got - want).abs_()
    d = torch.minimum(d, (L_BOX - d).abs_())
    err_px = float(d.max().item()) / DX_POS
    dist.barrier()
    if rank == 0:
        try:
            os.unlink(path)
        except OSError:
            pass
    res = torch.tensor([int(ok_gather), int(ok_scan), int(err_px <= 0.53)], dtype=torch.int64, device=dev)
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    return {"ok": bool(res.min().item() == 1), "sizes_equal_torch_all_gather": bool(res[0].item()), "offsets_equal_single_scan": bool(res[1].item()),
            "neighbour_file_read_back_through_minp_Reader": bool(res[2].item()), "max_error_pixels_rank0": err_px, "base_rank0": base_r,
            "blocks_in_index": int(world * 2 * nb), "snapshot_bytes": int(all_total.item()),
            "scope": "global index over all ranks' x and v blocks; file 0 (x) of every rank written at its global offsets into one "
                     "tmpfs file and read back by the next rank"}


def verify_roundtrip(torch, mb, ctx, stream, dev, pos, vel, pdescs, vdescs, packed, meta, out_len, decoded, stride):
    """Block offsets = running sums of ArrayBytes(bits, n) per (file, axis) stream; a CENTER-jitter decode lies within half
    a pixel (+ float32 rounding) of every input value, modulo the group's range (every group is periodic in the format,
    go/writer.go:74: a value that quantises to index == pixels, the field maximum, comes back at the other end of the
    range, in the reference as here)."""
    worst, ok = {}, True
    with torch.cuda.stream(stream):
        for key, field, descs_k, wrap in (("x", pos, pdescs, L_BOX), ("v", vel, vdescs, 0.0)):
            ctx.decode_vec3_subcells_dev(descs_k, packed[key], stride, meta[key][2], meta[key][0], meta[key][1],
                                         NFILE, SUB_CELLS, NFILES, wrap, mb.Jitter.make(mb.JITTER_CENTER, 0), decoded)
            ctx.sync()
            w = 0.0
            for f in range(NFILES):
                dk = [descs_k[(3 * f if len(descs_k) > 3 else 0) + k] for k in range(3)]
                span = torch.tensor([q.high - q.low for q in dk], dtype=torch.float32, device=dev)
                dxs = span / torch.tensor([float(q.pixels) for q in dk], dtype=torch.float32, device=dev)
                d = (decoded[f] - field[f]).abs_()
                d = torch.minimum(d, (span - d).abs_())
                w = max(w, float((d / dxs).max().item()))
            worst[key] = w
            nbs = ((meta[key][1] * NSUB3 + 7) // 8).reshape(3 * NFILES, SC3)
            o = meta[key][2].reshape(3 * NFILES, SC3)
            ok = ok and bool(torch.equal(o, torch.cumsum(nbs, 1) - nbs)) and bool(torch.equal(out_len[key], o[:, -1] + nbs[:, -1]))
    ok = ok and worst["x"] <= 0.53 and worst["v"] <= 0.53
    # ---- what the TIMED steps (the cooperative full-batch k_pipe_vec3 run) left behind for file 0 of x and of v, against
    # the oracle on the same particles: (min, bits) of all 3 x 64 blocks, every packed byte, and the HASH-jitter decode
    from oracle import oracle as orc
    orc.lib()
    threads = host_threads()
    bytes_equal, detail = True, {}
    with torch.cuda.stream(stream):
        for key, field, descs_k, wrap in (("x", pos, pdescs, L_BOX), ("v", vel, vdescs, 0.0)):
            dk = [descs_k[k] for k in range(3)]   # file 0: the first three descriptors (shared ones or its own)
            lo, hi, px = [d.low for d in dk], [d.high for d in dk], [d.pixels for d in dk]
            host = field[0].cpu().numpy()
            om, ob, onb, opk, ostride, _ = orc.bench_minp_encode(host, NFILE, SUB_CELLS, lo, hi, px, threads)
            gm, gb, go = (meta[key][i][:3 * SC3].cpu().numpy() for i in range(3))
            e = bool(np.array_equal(gm, om) and np.array_equal(gb, ob))
            for k in range(3):
                want = b"".join(opk[t * ostride:t * ostride + onb[t]].tobytes() for t in range(k * SC3, (k + 1) * SC3))
                ln = int(out_len[key][k].item())
                got = packed[key][k * stride:k * stride + ln].cpu().numpy().tobytes()
                e = e and ln == len(want) and got == want
                e = e and bool(np.array_equal(go[k * SC3:(k + 1) * SC3], np.concatenate([[0], np.cumsum(onb[k * SC3:(k + 1) * SC3])[:-1]])))
            ctx.decode_vec3_subcells_dev(descs_k, packed[key], stride, meta[key][2], meta[key][0], meta[key][1],
                                         NFILE, SUB_CELLS, NFILES, wrap, mb.Jitter.make(mb.JITTER_HASH, 7), decoded)
            ctx.sync()
            want = orc.bench_minp_decode(opk, ostride, NFILE, SUB_CELLS, lo, hi, px, om, ob, wrap > 0, L_BOX, 1, 7, threads)
            e = e and decoded[0].cpu().numpy().tobytes() == want.tobytes()
            detail[key] = bool(e)
            bytes_equal = bytes_equal and e
    return {"ok": bool(ok and bytes_equal), "bytes_equal_oracle": bool(bytes_equal), "bytes_equal_detail": detail,
            "bytes_equal_scope": "file 0 of x and of v from the timed cooperative full-batch run: (min, bits, offsets) of 3 x 64 blocks, all "
                                 "packed bytes, HASH-jitter decoded floats, against the oracle",
            "max_roundtrip_error_pixels": worst,
            "offsets": "running sums of ArrayBytes(bits, n) per stream" if ok else "see ok"}


def pcie_ceiling(torch, dev, mib=512, reps=6):
    """The duplex pinned-copy bandwidth of this GPU in this process, measured now (tools/pcie_probe.cu does the same
    standalone): both directions at once on two streams, CUDA events.  The e2e step moves (nearly) the same number of
    bytes each way, so its ceiling is what the slower direction sustains while the other one is busy too."""
    n = mib << 20
    ha, hb = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    da, db = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
    s0, s1 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s0):
        da.copy_(ha, non_blocking=True)
    with torch.cuda.stream(s1):
        hb.copy_(db, non_blocking=True)
    torch.cuda.synchronize()
    a0, b0, a1, b1 = ev(), ev(), ev(), ev()
    with torch.cuda.stream(s0):
        a0.record(s0)
        for _ in range(reps):
            da.copy_(ha, non_blocking=True)
        b0.record(s0)
    with torch.cuda.stream(s1):
        a1.record(s1)
        for _ in range(reps):
            hb.copy_(db, non_blocking=True)
        b1.record(s1)
    torch.cuda.synchronize()
    return {"duplex_h2d_gbs": n * reps / a0.elapsed_time(b0) / 1e6, "duplex_d2h_gbs": n * reps / a1.elapsed_time(b1) / 1e6,
            "how": "%d x %d MiB pinned copies in each direction at once, CUDA events" % (reps, mib)}


def run_e2e(torch, mb, ctx, pos, vel, pdescs, world, args, dev):
    """The same step through the host-pointer C ABI with ONE host thread: every file's particles come from pinned host
    memory (mnw_pipe_minp_encode_vectors: upload, limits, parameters, encode, packed bytes and metadata back to the host),
    are sent down again and decoded to host memory (mnw_pipe_minp_decode_vectors).  The pipe keeps several files in flight
    on separate streams, so uploads, kernels and downloads of different files overlap; the thread only submits and waits."""
    import ctypes as C
    import psutil
    budget = 0.35 * psutil.virtual_memory().available / max(world, 1)
    per_file = 2 * 12 * NP_FILE
    nfiles = int(max(1, min(NFILES, budget // per_file)))
    hpos = torch.empty((nfiles, NP_FILE, 3), dtype=torch.float32).pin_memory()
    hvel = torch.empty((nfiles, NP_FILE, 3), dtype=torch.float32).pin_memory()
    hpos.copy_(pos[:nfiles])
    hvel.copy_(vel[:nfiles])
    stride = 4 * NP_FILE + 256
    nbk = 3 * SC3
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    K, LAG = 4, 2                                  # host buffer sets; encode -> decode distance in jobs
    pipe = mb.Pipe(local_rank, depth=2 * K)
    P = lambda a: C.c_void_p(a.ctypes.data) if isinstance(a, np.ndarray) else C.c_void_p(a.data_ptr())

    class Buf:
        def __init__(self):
            self.hout = torch.empty(3 * stride, dtype=torch.uint8).pin_memory()
            self.hdec = torch.empty((NP_FILE, 3), dtype=torch.float32).pin_memory()
            self.mins, self.bits, self.offs = (np.zeros(nbk, np.int64) for _ in range(3))
            self.lens = np.zeros(3, np.int64)
            self.d3 = (mb.FloatDesc * 3)()
            self.ptrs = (C.c_void_p * 3)(*[self.hout.data_ptr() + k * stride for k in range(3)])
            self.enc = self.dec = -1
    bufs = [Buf() for _ in range(K)]
    jit = mb.Jitter.make(mb.JITTER_HASH, 7)
    jobs = [(field, f) for f in range(nfiles) for field in ("x", "v")]
    counts = {"h2d": 0, "d2h": 0}

    def one_pass(count):
        nj = len(jobs)
        for i in range(nj + LAG):
            if i < nj:
                b = bufs[i % K]
                if b.dec >= 0:
                    pipe.wait(b.dec)               # the decode that last used this buffer set
                field, f = jobs[i]
                per = field == "x"
                b.enc = pipe.encode(hpos[f] if per else hvel[f], NFILE, SUB_CELLS, per, L_BOX if per else 0.0, DX_POS if per else DV,
                                    b.d3, b.mins, b.bits, b.offs, b.hout, stride, b.lens)
            j = i - LAG
            if j >= 0:
                b = bufs[j % K]
                pipe.wait(b.enc)
                per = jobs[j][0] == "x"
                b.dec = pipe.decode(b.d3, b.ptrs, b.lens, b.offs, b.mins, b.bits, NFILE, SUB_CELLS, L_BOX if per else 0.0, jit, b.hdec)
                if count:
                    pk = int(b.lens.sum())
                    counts["h2d"] += 12 * NP_FILE + pk + 3 * 8 * nbk
                    counts["d2h"] += pk + (3 * 8 * nbk + 24 + 72 + 8) + 12 * NP_FILE
        pipe.drain()
        for b in bufs:
            b.dec = -1
    one_pass(False)                                # warm-up (buffers grow once)
    torch.cuda.synchronize()
    steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(steps):
        one_pass(True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    # what the last pass left in the last buffer set, against the device-resident arm's result for the same file
    pipe.close()
    nbytes = steps * nfiles * 4 * 12 * NP_FILE
    scale = NFILES / nfiles                       # bytes per full step, as counted from the copies made
    h2d, d2h = counts["h2d"] / steps * scale, counts["d2h"] / steps * scale
    del hpos, hvel, bufs
    ceil = pcie_ceiling(torch, dev)
    if world > 1:                                 # every rank probes at the same time: the NODE's ceiling
        import torch.distributed as dist
        c = torch.tensor([ceil["duplex_h2d_gbs"], ceil["duplex_d2h_gbs"]], dtype=torch.float64, device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.MIN)
        ceil["duplex_h2d_gbs"], ceil["duplex_d2h_gbs"] = float(c[0]), float(c[1])
        ceil["how"] += "; all ranks at once, minimum over ranks"
    t_min = max(h2d / (ceil["duplex_h2d_gbs"] * 1e9), d2h / (ceil["duplex_d2h_gbs"] * 1e9))
    value = world * nbytes / dt / 1e9
    ceiling = world * 4 * 12 * NSIDE ** 3 / t_min / 1e9
    return {"value": value, "unit": "GB/s",
            "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "files_timed_per_step": nfiles, "steps": steps, "host_threads": 1,
            "pcie": ceil, "ceiling_gbs": ceiling, "frac_of_ceiling": value / ceiling,
            "api": "mnw_pipe_minp_encode_vectors + mnw_pipe_minp_decode_vectors, one file per call, pinned host buffers, ONE host "
                   "thread; %d host buffer sets, pipe depth %d" % (K, 2 * K)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer e2e leg (profiling runs)")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1 / C3 / C4 configurations (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    # stdout carries exactly ONE JSON line: anything a library prints there (NCCL's version banner at
    # communicator creation) is sent to stderr instead
    _stdout_fd = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(_stdout_fd, "w")
    main()
    sys.stdout.flush()
