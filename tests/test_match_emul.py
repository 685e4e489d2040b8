"""K7 phase 1 as the device runs it (csrc/png_match_core.h: comparison bit arrays per candidate distance, run lengths by funnel shift
and count-trailing-zeros) equals its definition (a byte-compare loop per position and candidate) on streams that exercise long runs,
ties between candidates, chunk ends, stream ends, tiny strides and every pixel size."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL_DIR = os.path.join(ROOT, "tests", "emul")


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(EMUL_DIR, "libmatch_emul.so")
    srcs = [os.path.join(EMUL_DIR, "match_emul.cpp"), os.path.join(ROOT, "caesium-clt_b200", "csrc", "png_match_core.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, srcs[0]])
    lib = C.CDLL(so)
    lib.emul_match_compare.restype = C.c_longlong
    return lib


def _streams():
    rng = np.random.default_rng(3)
    yield "zeros", np.zeros(9000, np.uint8), 4, 257
    yield "noise", rng.integers(0, 256, 20000, dtype=np.uint8), 3, 301
    yield "few", rng.integers(0, 3, 30000, dtype=np.uint8), 1, 120
    yield "rows", np.tile(rng.integers(0, 256, 401, dtype=np.uint8), 60), 4, 401                 # every row repeats the one above
    yield "rows2", np.tile(np.concatenate([rng.integers(0, 256, 200, dtype=np.uint8), rng.integers(0, 256, 200, dtype=np.uint8)]), 40), 2, 200
    yield "pixels", np.tile(rng.integers(0, 256, 8, dtype=np.uint8), 3000), 8, 1601
    px = np.repeat(rng.integers(0, 256, 3000, dtype=np.uint8), rng.integers(1, 40, 3000)); yield "runs", px[:40000], 1, 513
    yield "tiny_stride", rng.integers(0, 2, 5000, dtype=np.uint8), 4, 5                          # stride - bpp = 1, 2 * stride < 3 * bpp
    yield "stride_lt_bpp", rng.integers(0, 2, 3000, dtype=np.uint8), 8, 3                        # stride - bpp < 1: candidate unusable
    yield "short", rng.integers(0, 2, 7, dtype=np.uint8), 1, 3
    yield "one", np.zeros(1, np.uint8), 1, 1
    yield "long_row", np.tile(rng.integers(0, 4, 16385, dtype=np.uint8), 3), 4, 16385            # 2 * stride > 32768: unusable
    g = (np.arange(70000) // 7 % 256).astype(np.uint8); yield "ramp", g, 4, 4097
    m = rng.integers(0, 256, 50000, dtype=np.uint8); m[10000:20000] = 7; m[30000:30300] = m[30000 - 1025:30300 - 1025]; yield "mixed", m, 4, 1025


@pytest.mark.parametrize("case", list(_streams()), ids=lambda c: c[0])
def test_bit_array_match_lengths_equal_the_byte_loops(emul, case):
    _, s, bpp, stride = case
    s = np.ascontiguousarray(s)
    bad = C.c_longlong(-1)
    n = emul.emul_match_compare(s.ctypes.data_as(C.c_void_p), C.c_ulonglong(s.size), bpp, stride, 4096, C.byref(bad))
    assert n == 0, (case[0], "first mismatch at", bad.value)
