"""GPU tests of the device entropy encoder (jpeg_gpuenc.cu): byte-identical to the host writer / oracle, in both
sequential and progressive mode, including the rare jcphuff.c flush rules, partial MCUs (device dummy-block fill) and
the full 4K configuration; and the two entropy modes of b200_compress_in_memory agree."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

INPUTS = ["in_420_base_355x237.jpg", "in_420_prog_355x237.jpg", "in_444_base_355x237.jpg", "in_422_base_355x237.jpg",
          "in_gray_base_355x237.jpg", "in_420_base_640x480.jpg", "in_420_tiny_17x9.jpg", "in_420_tiny_3x3.jpg"]


@pytest.mark.parametrize("name", INPUTS)
@pytest.mark.parametrize("prog", [0, 1])
def test_device_encoder_matches_host_encoder(L, O, golden, name, prog):
    data = golden(name)
    lay, co = L.jpeg_decode_coefficients(data)
    ref = L.jpeg_encode_coefficients(lay, co, prog)
    # zero the dummy blocks first: the device fills them itself (k_ge_fill_dummy) like jpeg_fill_dummy_blocks does on the host
    got = L.jpeg_encode_coefficients_device(lay, co, prog)
    assert got == ref
    # and the result is a valid file carrying the same coefficients
    l2, c2 = L.jpeg_decode_coefficients(got)
    for c in range(lay.ncomp):
        assert np.array_equal(L.component_view(lay, co, c)[:lay.rbh[c], :lay.rbw[c]], L.component_view(l2, c2, c)[:lay.rbh[c], :lay.rbw[c]])


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


def test_device_encoder_flush_rules(L):
    rng = np.random.default_rng(13)
    # EOBRUN counter overflow
    lay = _layout(L, 2048, 2048)
    co = np.zeros(lay.total_coefs, dtype=np.int16)
    co[::64] = 5
    co[64 * 40000 + 3] = 7
    for prog in (0, 1):
        assert L.jpeg_encode_coefficients_device(lay, co, prog) == L.jpeg_encode_coefficients(lay, co, prog)
    # correction-bit buffer overflow in refinement scans
    lay = _layout(L, 640, 480, 3)
    co = np.zeros(lay.total_coefs, dtype=np.int16)
    blocks = co.reshape(-1, 64)
    blocks[:, 0] = rng.integers(-50, 50, size=len(blocks))
    for b in range(len(blocks)):
        kind = rng.random()
        if kind < 0.80:
            idx = rng.choice(np.arange(1, 64), size=int(rng.integers(10, 41)), replace=False)
            blocks[b, idx] = rng.choice([-6, -4, -3, -2, 2, 3, 4, 6], size=len(idx))
        elif kind < 0.85:
            idx = rng.choice(np.arange(1, 64), size=12, replace=False)
            blocks[b, idx] = rng.choice([-1, 1, -2, 2, 5], size=12)
    for prog in (0, 1):
        assert L.jpeg_encode_coefficients_device(lay, co, prog) == L.jpeg_encode_coefficients(lay, co, prog)
    # dense / sparse random content, big magnitudes (long codes, many 0xFF bytes to stuff)
    for density in (0.02, 0.3, 0.9):
        lay = _layout(L, 256, 128, 3)
        co = np.zeros(lay.total_coefs, dtype=np.int16)
        nz = rng.random(lay.total_coefs) < density
        co[nz] = rng.integers(-1000, 1001, size=int(nz.sum()))
        co[::64] = rng.integers(-1000, 1000, size=lay.total_coefs // 64)
        for prog in (0, 1):
            assert L.jpeg_encode_coefficients_device(lay, co, prog) == L.jpeg_encode_coefficients(lay, co, prog)


def test_entropy_modes_agree_and_match_oracle_4k(L, O):
    from tools.synth import synth_jpeg
    data = synth_jpeg(3840, 2160, 1)
    for prog in (1, 0):
        p = L.default_params()
        p.jpeg_quality, p.jpeg_chroma_subsampling, p.jpeg_progressive = 80, 420, prog
        L.set_entropy_mode(1)
        gpu = L.compress_in_memory(data, p)
        L.set_entropy_mode(0)
        host = L.compress_in_memory(data, p)
        L.set_entropy_mode(1)
        assert gpu == host
        assert gpu == O.jpeg_lossy(data, O.params(80, 420, bool(prog)))


def test_concurrent_callers_device_entropy(L, O, golden):
    datas = [golden(n) for n in INPUTS] * 6
    p = L.default_params()
    p.jpeg_quality, p.jpeg_chroma_subsampling, p.jpeg_progressive = 75, 0, 1
    L.set_entropy_mode(1)
    res = L.compress_batch(datas, p, n_threads=12)
    for d, (out, code, msg) in zip(datas, res):
        assert code == 0, msg
        assert out == O.jpeg_lossy(d, O.params(75, 0, True))
