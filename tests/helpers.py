"""Shared helpers for the parity tests: the oracle applied block by block the
way the reference's Writer.Data loop would."""
import numpy as np


def oracle_int_group(orc, x, starts):
    mins, bits, offs, chunks, pos = [], [], [], [], 0
    for b in range(len(starts) - 1):
        mn, bt, data = orc.int_block_encode(x[starts[b]:starts[b + 1]])
        mins.append(mn); bits.append(bt); offs.append(pos); chunks.append(data); pos += len(data)
    cat = np.concatenate(chunks) if chunks else np.zeros(0, np.uint8)
    return np.array(mins, np.int64), np.array(bits, np.int64), np.array(offs, np.int64), cat


def oracle_float_group(orc, x, starts, low, high, pixels, periodic=1, log10=0, clamp=0):
    mins, bits, offs, chunks, pos = [], [], [], [], 0
    for b in range(len(starts) - 1):
        blk = np.asarray(x[starts[b]:starts[b + 1]], np.float32)
        if log10 or clamp:
            # minh.Writer.Block: processFloatGroup before Data (go/minh/minh.go:131-136);
            # the reference has no clamp-without-log switch, both come from the Column
            if clamp:
                blk = orc.minh_process_float(blk, log10, low, high)
            else:
                blk = orc.minh_process_float(blk, log10, -np.inf, np.inf)
        mn, bt, data = orc.float_block_encode(blk, low, high, pixels, periodic)
        mins.append(mn); bits.append(bt); offs.append(pos); chunks.append(data); pos += len(data)
    cat = np.concatenate(chunks) if chunks else np.zeros(0, np.uint8)
    return np.array(mins, np.int64), np.array(bits, np.int64), np.array(offs, np.int64), cat


def uniform_starts(n, nblocks):
    return np.arange(nblocks + 1, dtype=np.int64) * n


def splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))
