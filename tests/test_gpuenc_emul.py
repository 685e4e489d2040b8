"""CPU validation of the block-parallel entropy-encoder formulation (caesium-clt_b200/csrc/jpeg_gpuenc_core.h): the
bodies of the GPU kernels are run serially by tests/emul/gpuenc_emul.cpp and must reproduce the sequential host writer
(and hence the oracle) byte for byte -- including the rare jcphuff.c flush rules (EOBRUN == 0x7FFF, > 937 pending
correction bits) that the parallel formulation handles with a sequential replay of the affected run."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL_DIR = os.path.join(ROOT, "tests", "emul")
INPUTS = ["in_420_base_355x237.jpg", "in_420_prog_355x237.jpg", "in_444_base_355x237.jpg", "in_422_base_355x237.jpg",
          "in_gray_base_355x237.jpg", "in_420_base_640x480.jpg", "in_420_tiny_17x9.jpg", "in_420_tiny_3x3.jpg"]


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(EMUL_DIR, "libgpuenc_emul.so")
    srcs = [os.path.join(EMUL_DIR, "gpuenc_emul.cpp"), os.path.join(ROOT, "caesium-clt_b200", "csrc", "jpeg_host.cpp"),
            os.path.join(ROOT, "caesium-clt_b200", "csrc", "jpeg_gpuenc_core.h"), os.path.join(ROOT, "caesium-clt_b200", "csrc", "jpeg_gpuenc_plan.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-msse2", "-o", so, srcs[0], srcs[1]])
    return C.CDLL(so)


def transcode(emul, data, prog):
    outp, outl = C.c_void_p(), C.c_size_t()
    rc = emul.emul_gpu_transcode(data, C.c_size_t(len(data)), prog, C.byref(outp), C.byref(outl))
    assert rc == 0, rc
    return C.string_at(outp, outl.value)


def host_transcode(L, data, prog):
    p = L.default_params()
    p.jpeg_optimize, p.jpeg_progressive = 1, prog
    return L.compress_in_memory(data, p)


@pytest.mark.parametrize("name", INPUTS)
@pytest.mark.parametrize("prog", [0, 1])
def test_block_parallel_encoder_matches_sequential_writer(L, O, emul, golden, name, prog):
    data = golden(name)
    got = transcode(emul, data, prog)
    assert got == host_transcode(L, data, prog)
    assert got == O.jpeg_lossless(data, O.params(80, 0, bool(prog)))


def test_swar_threshold_masks_equal_the_definition(emul):
    """make_masks3 (the word-parallel form the kernels run) against |c| >= 1 / 2 / 4 per coefficient: 200k random blocks with
    zeros, small values and the int16 extremes."""
    assert emul.emul_masks_check(200000, 12345) == 0


def _layout(L, w, h, ncomp=1):
    lay = L.JpegLayout()
    lay.width, lay.height, lay.ncomp = w, h, ncomp
    off = 0
    for c in range(ncomp):
        lay.hs[c] = lay.vs[c] = 1
        lay.bw[c] = lay.rbw[c] = -(-w // 8)
        lay.bh[c] = lay.rbh[c] = -(-h // 8)
        lay.comp_offset[c] = off
        off += lay.bw[c] * lay.bh[c] * 64
        for k in range(64):
            lay.qt[c][k] = 1
    lay.total_coefs = off
    return lay


def test_eobrun_counter_overflow(L, emul):
    """> 0x7FFF consecutive blocks with empty AC bands: the EOB run must be split exactly like jcphuff.c does."""
    lay = _layout(L, 2048, 2048)                       # 65536 blocks
    co = np.zeros(lay.total_coefs, dtype=np.int16)
    co[::64] = 5
    co[64 * 40000 + 3] = 7                             # one event block in the middle of the second run
    src = L.jpeg_encode_coefficients(lay, co, 1)
    for prog in (0, 1):
        assert transcode(emul, src, prog) == host_transcode(L, src, prog)


def test_correction_bit_buffer_overflow(L, emul):
    """Refinement scans: long runs of blocks that only carry correction bits (|c| >= 2, no |c| == 1) overflow the 937-bit
    buffer and force mid-run flushes; mixed with event blocks and empty blocks."""
    rng = np.random.default_rng(13)
    lay = _layout(L, 640, 480, 3)
    co = np.zeros(lay.total_coefs, dtype=np.int16)
    blocks = co.reshape(-1, 64)
    blocks[:, 0] = rng.integers(-50, 50, size=len(blocks))
    for b in range(len(blocks)):
        kind = rng.random()
        if kind < 0.80:                                # correction-only block: 10..40 coefficients with |c| >= 2
            idx = rng.choice(np.arange(1, 64), size=int(rng.integers(10, 41)), replace=False)
            blocks[b, idx] = rng.choice([-6, -4, -3, -2, 2, 3, 4, 6], size=len(idx))
        elif kind < 0.85:                              # event block with everything: +-1, bigger values, long zero runs
            idx = rng.choice(np.arange(1, 64), size=12, replace=False)
            blocks[b, idx] = rng.choice([-1, 1, -2, 2, 5], size=12)
        # else: empty AC band
    src = L.jpeg_encode_coefficients(lay, co, 1)
    for prog in (0, 1):
        assert transcode(emul, src, prog) == host_transcode(L, src, prog)


def test_random_sparse_and_dense_blocks(L, emul):
    rng = np.random.default_rng(17)
    for density in (0.02, 0.3, 0.9):
        lay = _layout(L, 256, 128, 3)
        co = np.zeros(lay.total_coefs, dtype=np.int16)
        nz = rng.random(lay.total_coefs) < density
        co[nz] = rng.integers(-40, 41, size=int(nz.sum()))
        co[::64] = rng.integers(-900, 900, size=lay.total_coefs // 64)
        src = L.jpeg_encode_coefficients(lay, co, 0)
        for prog in (0, 1):
            assert transcode(emul, src, prog) == host_transcode(L, src, prog)
