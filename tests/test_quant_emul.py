"""The transform kernels' quantiser (biased, sign-free reciprocal division; jpeg_kernels.h) against the jcdctmgr.c rule the
oracle restates (round half away from zero, divisor quantval << 3): every int16 input, quantiser values from 1 to 65535 --
the Robidoux table at q = 1..100 included -- and both kernel variants (with and without the post-shift)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL_DIR = os.path.join(ROOT, "tests", "emul")


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(EMUL_DIR, "libquant_emul.so")
    srcs = [os.path.join(EMUL_DIR, "quant_emul.cpp"), os.path.join(ROOT, "caesium-clt_b200", "csrc", "jpeg_kernels.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, srcs[0]])
    lib = C.CDLL(so)
    lib.emul_quant_check.restype = C.c_longlong
    return lib


def _check(emul, qv):
    qv = np.ascontiguousarray(qv, dtype=np.uint16)
    assert qv.size == 64
    return emul.emul_quant_check(qv.ctypes.data_as(C.c_void_p))


def test_every_small_quantiser_value(emul):
    for base in range(1, 321, 64):
        assert _check(emul, np.arange(base, base + 64)) == 0


def test_large_and_boundary_quantiser_values(emul):
    vals = [1, 2, 3, 255, 256, 257, 511, 512, 513, 1023, 1024, 1025, 2047, 2048, 4095, 4096, 4097, 8191, 8192, 16383, 16384,
            20900, 32766, 32767, 32768, 40000, 65534, 65535, 418, 419, 7, 100]
    qv = np.array(vals + vals, dtype=np.uint16)
    assert _check(emul, qv) == 0
    assert emul.emul_quant_any_shift(qv.ctypes.data_as(C.c_void_p)) == 1
    assert emul.emul_quant_any_shift(np.full(64, 512, dtype=np.uint16).ctypes.data_as(C.c_void_p)) == 0


def test_output_tables_of_every_quality(L, emul):
    """the tables the path actually uses: jpeg_quant_table(q) for q = 1..100 (luma and chroma)"""
    for q in list(range(1, 101, 9)) + [100]:
        for comp in (0, 1):
            assert _check(emul, L.jpeg_quant_table(q, comp)) == 0, (q, comp)
