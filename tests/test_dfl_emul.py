"""The shared DEFLATE block coder (csrc/dfl_core.h, run by the host writer and by the device kernels alike): its code-length
construction equals the heap formulation of the round-1 host writer on arbitrary histograms, its closed-form length / distance
code tables equal RFC 1951's, and deflate_tokens() still produces streams zlib inflates to the right bytes."""
import ctypes as C
import os
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL_DIR = os.path.join(ROOT, "tests", "emul")


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(EMUL_DIR, "libdfl_emul.so")
    srcs = [os.path.join(EMUL_DIR, "dfl_emul.cpp"), os.path.join(ROOT, "caesium-clt_b200", "csrc", "dfl_core.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, srcs[0]])
    return C.CDLL(so)


def test_closed_form_symbol_tables(emul):
    assert emul.emul_len_dist_tables() == 0


def test_two_queue_code_lengths_equal_the_heap_formulation(emul):
    rng = np.random.default_rng(11)
    cases = []
    for n, limit in ((286, 15), (30, 15), (19, 7)):
        for _ in range(400):
            kind = rng.integers(0, 6)
            if kind == 0:
                f = rng.integers(0, 4, n)                               # many ties, many zeros
            elif kind == 1:
                f = rng.integers(1, 70000, n)
            elif kind == 2:
                f = (rng.pareto(0.7, n) * 3).astype(np.int64)           # heavy tail: deep trees, overflow repair
            elif kind == 3:
                f = np.where(rng.random(n) < 0.1, rng.integers(1, 5, n), 0)
            elif kind == 4:
                f = np.array([int(1.6 ** (i % 40)) for i in range(n)])  # Fibonacci-like: maximal depth
            else:
                f = np.full(n, int(rng.integers(1, 9)))
            cases.append((np.minimum(f, 2**31 - 1).astype(np.uint32), n, limit))
        cases.append((np.zeros(n, np.uint32), n, limit))
        one = np.zeros(n, np.uint32); one[n // 2] = 5
        cases.append((one, n, limit))
    for f, n, limit in cases:
        f = np.ascontiguousarray(f)
        assert emul.emul_huff_lengths_compare(f.ctypes.data_as(C.c_void_p), n, limit) == 0


def test_deflate_tokens_round_trips(L):
    rng = np.random.default_rng(5)
    for size, block in ((0, 65536), (1, 65536), (70000, 65536), (200000, 4096), (5000, 100)):
        data = bytearray()
        toks = []
        while len(data) < size:
            if len(data) > 300 and rng.random() < 0.3:
                ln = int(rng.integers(3, 259)); d = int(rng.integers(1, min(len(data), 32768) + 1))
                for _ in range(ln):
                    data.append(data[-d])
                toks.append(0x80000000 | ((ln - 3) << 16) | (d - 1))
            else:
                v = int(rng.integers(0, 256)) if rng.random() < 0.5 else 7
                data.append(v); toks.append(v)
        z = L.png_deflate_tokens(np.array(toks, np.uint32), zlib.adler32(bytes(data)))
        assert zlib.decompress(z) == bytes(data)
