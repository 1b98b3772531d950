"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports
every symbol include/minnow_cuda.h declares; scalar helpers agree with the
oracle; and the product refuses to run without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import minnow_b200
from minnow_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols(header):
    with open(os.path.join(ROOT, "include", header)) as f:
        text = f.read()
    return sorted(set(re.findall(r"MNW_API[^;(]*?\b(mnw_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = minnow_b200.load_library()
    for header in sorted(os.listdir(os.path.join(ROOT, "include"))):
        syms = _declared_symbols(header)
        assert syms, header
        for s in syms:
            assert hasattr(lib, s), "%s declared in %s but not exported" % (s, header)


def test_python_binding_covers_cuda_header():
    assert set(_declared_symbols("minnow_cuda.h")) == set(capi.SIGNATURES)


def test_scalar_helpers_match_oracle(orc):
    rng = np.random.default_rng(5)
    vals = [0, 1, 2, 3, 4, 255, 256, 2 ** 31 - 1, 2 ** 31, 2 ** 48 - 1, 2 ** 48, 2 ** 49 - 1, 2 ** 49, 2 ** 49 + 1,
            2 ** 52, 2 ** 53, 2 ** 53 + 1, 2 ** 62, 2 ** 63 - 1, 2 ** 63, 2 ** 64 - 2]
    vals += [int(v) for v in rng.integers(0, 2 ** 63, 2000)]
    vals += [(1 << k) + d for k in range(1, 64) for d in (-2, -1, 0, 1, 2) if 0 <= (1 << k) + d < 2 ** 64 - 1]
    for v in vals:
        assert minnow_b200.precision_needed(v) == orc.precision_needed(v), v
    for bits, n in [(0, 5), (1, 1), (3, 4), (10, 3), (17, 262144), (64, 123), (19, 2)]:
        assert minnow_b200.array_bytes(bits, n) == orc.array_bytes(bits, n)
    for lo, hi, dx in [(0, 125, 0.001), (-50, 100, 1), (-50, 100, 10), (10, 14, 0.01), (0, 1000, 0.005), (0, 250, 1)]:
        assert minnow_b200.float_group_pixels(lo, hi, dx) == orc.float_group_pixels(lo, hi, dx)
    assert minnow_b200.float_group_pixels(0, 125, 0.001) == 125000
    for s, b, i in rng.integers(0, 2 ** 62, (200, 3)):
        assert minnow_b200.jitter_hash32(int(s), int(b), int(i)) == orc.jitter_hash32(int(s), int(b), int(i))


def test_undefined_precision_is_an_error():
    with pytest.raises(minnow_b200.MinnowError):
        minnow_b200.precision_needed(2 ** 64 - 1)


def test_no_cpu_fallback():
    """Without a GPU the product must fail loudly, never compute on the CPU."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(minnow_b200.MinnowError) as e:
        minnow_b200.Context()
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "minnow_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                with open(os.path.join(dirpath, fn)) as f:
                    text = f.read()
                assert "oracle" not in text.replace("CPU oracle under\n// oracle/", "").lower() or \
                    all("import" not in ln and "#include" not in ln and "dlopen" not in ln and "CDLL" not in ln
                        for ln in text.splitlines() if "oracle" in ln.lower()), fn
