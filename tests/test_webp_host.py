"""CPU tests of the WebP (lossy VP8) leg, SURVEY.md §8 row a10.

1. The oracle (oracle/webp_oracle.c) is pinned by DECODE: libwebp (through Pillow) must decode the oracle's files to exactly
   the oracle's own reconstruction -- RFC 6386 fixes every decoder step, so equality proves header, bool coder, token
   trees, dequantisation, inverse transforms and intra predictors of the restatement.  The RGB the decoder emits is
   compared through a numpy restatement of libwebp's fancy upsampler + fixed-point YUV->RGB (tests/webputil.py).
2. The product's HOST half (vp8_host.cpp: quality curve, bool coder, partitions, RIFF) must turn the oracle's stage output
   into byte-identical files.  No device work is called here."""
import numpy as np
import pytest

from pngutil import synth
from webputil import decode_like_libwebp, pil_decode

CASES = [(64, 64, 75, "photo"), (37, 53, 50, "photo"), (100, 130, 90, "flat"), (16, 16, 10, "noise"), (200, 300, 100, "photo"),
         (33, 17, 0, "photo"), (1, 1, 80, "flat"), (15, 300, 60, "flat"), (129, 65, 85, "noise"), (48, 48, 30, "flat")]


def planar(img):
    return np.ascontiguousarray(img.transpose(2, 0, 1))


@pytest.mark.parametrize("h,w,q,kind", CASES)
def test_oracle_files_decode_in_libwebp_to_the_oracle_reconstruction(O, h, w, q, kind):
    img = synth(h, w, 3, seed=h + w + q, kind=kind)
    data, (Y, U, V) = O.webp_encode(planar(img), q)
    assert data[:4] == b"RIFF" and data[8:16] == b"WEBPVP8 " and len(data) % 2 == 0
    assert int.from_bytes(data[4:8], "little") == len(data) - 8
    dec = pil_decode(data)
    assert dec.shape == (h, w, 3)
    assert np.array_equal(dec, decode_like_libwebp(Y, U, V, w, h))


def test_oracle_quality_monotonic_and_close_to_source(O):
    img = synth(96, 128, 3, seed=5, kind="photo")
    sizes, errs = [], []
    for q in (10, 40, 75, 95):
        data, _ = O.webp_encode(planar(img), q)
        sizes.append(len(data)); errs.append(np.abs(pil_decode(data).astype(int) - img.astype(int)).mean())
    assert sizes == sorted(sizes) and errs == sorted(errs, reverse=True)
    assert errs[-1] < 3.0 and errs[0] < 12.0


def test_oracle_quality_curve_and_quant_factors(O, L):
    # libwebp's curve: q=100 -> index 0, q=0 -> 127, q=75 -> 127 * (1 - cbrt(0.5))
    assert O.vp8_qindex(100) == 0 and O.vp8_qindex(0) == 127 and O.vp8_qindex(75) == int(127 * (1 - 0.5 ** (1 / 3)))
    prev = 128
    for q in range(0, 101):
        qi, f = L.webp_qindex(q)
        assert qi == O.vp8_qindex(q) and f == O.vp8_quant_factors(qi)
        assert qi <= prev
        prev = qi
    assert O.vp8_quant_factors(0) == [4, 4, 8, 8, 4, 4] and O.vp8_quant_factors(127) == [157, 284, 314, 440, 132, 284]


def test_oracle_rgb_to_yuv_ranges_and_grey(O):
    g = np.repeat(np.arange(256, dtype=np.uint8).reshape(16, 16)[None], 3, axis=0)
    Y, U, V = O.webp_rgb_to_yuv(g)
    assert Y.min() == 16 and Y.max() == 235 and (U == 128).all() and (V == 128).all()
    rgb = np.zeros((3, 20, 20), np.uint8); rgb[0] = 255
    Y, U, V = O.webp_rgb_to_yuv(rgb)
    assert Y.shape == (32, 32) and (Y == Y[0, 0]).all() and (V[:, :] == 240).all()          # pure red, edges replicated into the padding


@pytest.mark.parametrize("h,w,q,kind", CASES)
def test_host_writer_reproduces_oracle_file_from_oracle_stage_output(L, O, h, w, q, kind):
    img = planar(synth(h, w, 3, seed=h + w + q, kind=kind))
    levels, modes = O.webp_analyze(img, q)
    want, _ = O.webp_encode(img, q)
    assert L.webp_write_levels(w, h, q, levels, modes) == want


def test_transparent_png_is_not_silently_flattened(L):
    """A PNG with real transparency cannot become a simple-format WebP without losing it: the call hands it back (code 3)
    before touching the device; the same goes for lossless WebP."""
    from pngutil import pil_png
    rgba = np.concatenate([synth(20, 30, 3, seed=1), synth(20, 30, 1, seed=2)], axis=2)
    p = L.default_params()
    with pytest.raises(L.B200Error) as e:
        L.convert_in_memory(pil_png(rgba), p, 3)
    assert e.value.code == 3 and "alpha" in str(e.value)
    p.webp_lossless = 1
    with pytest.raises(L.B200Error) as e:
        L.convert_in_memory(pil_png(rgba[:, :, :3].copy()), p, 3)
    assert e.value.code == 3


def test_host_writer_extreme_levels(L):
    # every token class incl. DCT_CAT6 (|v| up to 2047), random signs, all four modes, some skipped macroblocks
    rng = np.random.default_rng(3)
    w, h = 64, 48
    nmb = 4 * 3
    levels = np.zeros((nmb, 25, 16), np.int16)
    mags = np.array([0, 0, 0, 1, 1, 2, 3, 4, 5, 6, 7, 10, 11, 18, 19, 34, 35, 66, 67, 500, 2047])
    levels[:] = rng.choice(mags, size=levels.shape) * rng.choice([-1, 1], size=levels.shape)
    levels[:, 1:17, 0] = 0                      # luma DCs travel in Y2
    levels[5] = 0; levels[9] = 0
    modes = np.zeros((nmb, 4), np.uint8)
    modes[:, 0] = rng.integers(0, 4, nmb); modes[:, 1] = rng.integers(0, 4, nmb)
    modes[5, 2] = 1; modes[9, 2] = 1
    data = L.webp_write_levels(w, h, 50, levels, modes)
    dec = pil_decode(data)                      # libwebp must parse it to the last macroblock without error
    assert dec.shape == (h, w, 3)


# ---- round 2: WebP INPUT -- the host VP8 decoder in front of the device encoder -------------------------------------------------
import io
import os

SAMPLES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_samples")


def _pil_rgb(data):
    from PIL import Image
    return np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))


@pytest.mark.parametrize("name", ["w0.webp", "w1.webp"])
def test_vp8_decoder_equals_libwebp_on_the_reference_samples(L, name):
    """compress_in_memory on samples/w0.webp must succeed (compressor.rs:769-787); its first step is this decode: RIFF / VP8X
    container, RFC 6386 key frame (segments, B_PRED sub-block modes, token partitions, normal / simple loop filter), libwebp's fancy
    upsampling and fixed-point YUV -> RGB.  Bit-exact with libwebp (through Pillow)."""
    with open(os.path.join(SAMPLES, name), "rb") as f:
        data = f.read()
    got = L.webp_decode(data)
    assert np.array_equal(got, _pil_rgb(data))


@pytest.mark.parametrize("h,w", [(1, 1), (16, 16), (17, 33), (237, 355), (301, 77)])
def test_vp8_decoder_equals_libwebp_on_libwebp_encodings(L, h, w):
    from PIL import Image
    for kind in ("photo", "flat", "noise"):
        img = synth(h, w, 3, seed=h * 7 + w, kind=kind)
        for q, m in ((0, 0), (20, 2), (50, 4), (75, 6), (90, 4), (100, 3)):
            b = io.BytesIO(); Image.fromarray(img).save(b, "WEBP", quality=q, method=m)
            data = b.getvalue()
            assert np.array_equal(L.webp_decode(data), _pil_rgb(data)), (kind, q, m)


def test_vp8_decoder_refuses_what_it_does_not_decode(L):
    from PIL import Image
    img = synth(20, 30, 4, seed=1)
    b = io.BytesIO(); Image.fromarray(img).save(b, "WEBP", lossless=True)
    with pytest.raises(L.B200Error) as e:
        L.webp_decode(b.getvalue())
    assert e.value.code == 3
    b = io.BytesIO(); Image.fromarray(img).save(b, "WEBP", quality=80)      # lossy with an alpha plane (VP8X + ALPH)
    with pytest.raises(L.B200Error) as e:
        L.webp_decode(b.getvalue())
    assert e.value.code == 3 and "alpha" in str(e.value)
    ok = io.BytesIO(); Image.fromarray(img[:, :, :3]).save(ok, "WEBP", quality=80)
    data = ok.getvalue()
    for cut in (len(data) // 2, 40, 25):
        try:
            L.webp_decode(data[:cut])                   # truncated token partitions decode as zeros (libwebp would report an error)
        except L.B200Error as e2:
            assert e2.code == 4
    with pytest.raises(L.B200Error) as e:
        L.webp_decode(b"RIFF\x10\x00\x00\x00WEBPVP8 \x04\x00\x00\x00abcd")
    assert e.value.code == 4
