#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the REFERENCE's own
Python/Cython implementation (/root/reference/python), which cannot travel to
the GPU box.  Run it in the build container only:

    python tests/golden/make_golden.py

It copies /root/reference/python to a temp dir, cythonizes cy_bit.pyx there
(the reference tree is read-only), imports the reference modules and writes:

  bit_arrays.npz        packed bytes of reference bit.array for bits 1..64
  periodic_min.npz      inputs/outputs of the reference cy_bit.periodic_min
                        (the KAT of python/minnow_test.py:216-229 + random cases)
  int_record.minnow     python/minnow_test.py test_int_record file
  group_record.minnow   ... test_group_record file
  bit_int_record.minnow ... test_bit_int_record file (go/minnow_test.go:242-268 vectors)
  q_float_record.minnow ... test_q_float_record file (go/minnow_test.go:270-310 vectors)
  minh_reader_writer.minh  test_minh_reader_writer file (go/minh/minh_test.go:10-117 vectors)
  int_groups_random.npz + int_groups_random.minnow   seeded random IntGroup file

Only inputs on which the Python twin and the Go code provably agree are used:
integers, and floats that are exactly representable with exact dx (the twin
quantises floats with numpy promotion rules, Go in float32 -- SURVEY.md 8c).
"""
import os
import shutil
import struct
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/python"


def load_reference():
    tmp = tempfile.mkdtemp(prefix="minnow_ref_")
    for f in os.listdir(REF):
        if f.endswith((".py", ".pyx")):
            shutil.copy(os.path.join(REF, f), tmp)
    env = dict(os.environ, CFLAGS="-I" + np.get_include())
    subprocess.check_call(["cythonize", "-i", "-3", "cy_bit.pyx"], cwd=tmp, env=env,
                          stdout=subprocess.DEVNULL)
    sys.path.insert(0, tmp)
    import bit      # noqa: E402  (reference module)
    import minnow   # noqa: E402  (reference module)
    import minh     # noqa: E402  (reference module)
    return tmp, bit, minnow, minh


def main():
    tmp, bit, minnow, minh = load_reference()
    rng = np.random.default_rng(20261018)

    # ---- bit arrays: go/bit/bit_test.go:9-31 shape (123 random 63-bit values, bits 1..64)
    data = rng.integers(0, 2 ** 63, size=123, dtype=np.uint64)
    out = {"data": data}
    for bits in range(1, 65):
        mask = np.uint64((1 << bits) - 1)
        # The twin does not apply the mask itself (python/cy_bit.pyx:27-30), Go does
        # (go/bit/bit.go:107): feed pre-masked values, on which both agree.
        out["bits%02d" % bits] = bit.array(bits, data & mask)
    # ragged lengths, including the ArrayBuffer test lengths (go/bit/bit_test.go:38)
    for n in (1, 5, 7, 10, 20, 33):
        x = np.arange(n, dtype=np.uint64)
        b = bit.precision_needed(int(x.max()))
        if b:
            out["arange%02d_bits%02d" % (n, b)] = bit.array(b, x)
    np.savez_compressed(os.path.join(HERE, "bit_arrays.npz"), **out)

    # ---- periodic_min
    cases_x, cases_p, cases_m = [], [], []
    kat = [[0, 1, 2, 3], [10, 11, 12, 13], [18, 19, 0, 1], [1, 0, 19, 18], [1, 19, 18, 0]]
    for x in kat:
        cases_x.append(np.array(x, np.int64)); cases_p.append(20)
    for _ in range(3000):
        pixels = int(rng.integers(2, 400))
        n = int(rng.integers(1, 40))
        kind = rng.integers(0, 4)
        if kind == 0:      # anywhere in range
            x = rng.integers(0, pixels, n)
        elif kind == 1:    # clustered arc, wraps
            c = int(rng.integers(0, pixels)); w = int(rng.integers(1, pixels))
            x = (c + rng.integers(0, w, n)) % pixels
        elif kind == 2:    # includes the out-of-range index == pixels and a few beyond
            x = rng.integers(0, pixels + 1, n)
        else:              # slightly out of range on both sides
            x = rng.integers(-2, pixels + 3, n)
        cases_x.append(np.asarray(x, np.int64)); cases_p.append(pixels)
    for x, p in zip(cases_x, cases_p):
        cases_m.append(int(bit.periodic_min(x, p)))
    lens = np.array([len(x) for x in cases_x], np.int64)
    np.savez_compressed(os.path.join(HERE, "periodic_min.npz"), x=np.concatenate(cases_x), lens=lens,
                        pixels=np.array(cases_p, np.int64), mins=np.array(cases_m, np.int64))

    # ---- container files, written by the reference's python/minnow.py
    def path(name):
        return os.path.join(HERE, name)

    # test_int_record (python/minnow_test.py:67-80; go/minnow_test.go:191-218)
    f = minnow.create(path("int_record.minnow"))
    xs = [np.array([1, 2, 3, 4], np.int64), np.array([5], np.int64),
          np.array([6, 7, 8, 9], np.int64), np.array([10, 11, 12], np.int64)]
    text = b"I am a cat and I like to meow."
    f.header(struct.pack("<qq", 0xdeadbeef, 4))
    f.header(text)
    for x in xs:
        f.fixed_size_group(np.int64, len(x)); f.data(x)
    f.header(np.array([len(x) for x in xs], np.int64))
    f.close()

    # test_group_record (python/minnow_test.py:82-93; go/minnow_test.go:221-240)
    f = minnow.create(path("group_record.minnow"))
    ix = np.arange(20, dtype=np.int32); fx = np.arange(10) / 10.0
    f.header(struct.pack("<qq", 4, 5))
    f.fixed_size_group(np.int32, 5)
    for i in range(4): f.data(ix[5 * i: 5 * (i + 1)])
    f.header(struct.pack("<qq", 2, 5))
    f.fixed_size_group(np.float64, 5)
    for i in range(2): f.data(fx[5 * i: 5 * (i + 1)])
    f.header(b"I'm a caaaat")
    f.close()

    # test_bit_int_record
    f = minnow.create(path("bit_int_record.minnow"))
    x1 = np.array([100, 101, 102, 104], dtype=np.int64)
    x2 = [np.array([1024, 1024, 1024]), np.array([0, 1023, 500])]
    x3 = np.array([-1000000, -500000])
    f.int_group(len(x1)); f.data(x1)
    f.header(struct.pack("<q", len(x2)))
    f.int_group(len(x2[0]))
    for x in x2: f.data(x)
    f.int_group(len(x3)); f.data(x3)
    f.close()

    # test_q_float_record
    f = minnow.create(path("q_float_record.minnow"))
    limit = (-50, 100); dx1, dx2 = 1.0, 10.0
    q1 = [np.array([-50, 0, 50, 49]), np.array([25, 25, 25, 25])]
    q2 = [np.array([-50, 0, 50, 49, 0]), np.array([1, 2, 3, 4, 5]), np.array([0, 20, 0, 20, 0])]
    f.header(struct.pack("<ffffqq", dx1, dx2, limit[0], limit[1], len(q1), len(q2)))
    f.float_group(len(q1[0]), limit, dx1)
    for x in q1: f.data(x)
    f.float_group(len(q2[0]), limit, dx2)
    for x in q2: f.data(x)
    f.close()

    # test_minh_reader_writer
    names = ["int64", "float32", "int", "float", "log"]
    text = ("Cats are the best. Don't we love them?!@#$%^&*(),.." +
            "..[]{};':\"|\\/-=_+`~meow meow meow")
    columns = [minh.Column(minnow.int64_group), minh.Column(minnow.float32_group),
               minh.Column(minnow.int_group), minh.Column(minnow.float_group, 0, 100, 200, 1),
               minh.Column(minnow.float_group, 1, 10, 14, 0.01)]
    block1 = [np.array([100, 200, 300, 400, 500], np.int64), np.array([150, 250, 350, 450, 550], np.float32),
              np.array([-30, -35, -25, -10, -20], np.int64), np.array([100, 200, 125, 150, 100], np.float32),
              np.array([1e10, 1e11, 1e11, 1e14, 3e13], np.float32)]
    block2 = [np.array([125, 225, 325], np.int64), np.array([1750, 2750, 3750], np.float32),
              np.array([1000, 1000, 1000], np.int64), np.array([100, 100, 100], np.float32),
              np.array([1e14, 1e14, 1e14], np.float32)]
    wr = minh.create(path("minh_reader_writer.minh"))
    wr.header(names, text, columns)
    wr.geometry(100.0, 10.0, 4)
    for blk in (block1, block2): wr.block(blk)
    wr.close()

    # seeded random IntGroup file: several groups, many blocks, negative values,
    # zero-width blocks, widths up to 40 bits (all below the 2^48 log2 caveat)
    groups = []
    f = minnow.create(path("int_groups_random.minnow"))
    meta = []
    for g in range(6):
        N = int(rng.integers(1, 70)); nb = int(rng.integers(1, 9))
        f.int_group(N)
        for b in range(nb):
            width = int(rng.integers(0, 41))
            base = int(rng.integers(-2 ** 45, 2 ** 45))
            x = base + rng.integers(0, 2 ** width, N, dtype=np.int64) if width else np.full(N, base, np.int64)
            f.data(np.asarray(x, np.int64))
            groups.append(np.asarray(x, np.int64)); meta.append((g, N))
        if g % 2 == 0:
            f.header(struct.pack("<q", g))
    f.close()
    np.savez_compressed(path("int_groups_random.npz"), x=np.concatenate(groups),
                        group=np.array([m[0] for m in meta], np.int64), n=np.array([m[1] for m in meta], np.int64))

    shutil.rmtree(tmp, ignore_errors=True)
    for fn in sorted(os.listdir(HERE)):
        print("%-28s %8d bytes" % (fn, os.path.getsize(os.path.join(HERE, fn))))


if __name__ == "__main__":
    main()
