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


def test_lossless_webp_is_handed_back(L):
    """webp.lossless selects libwebp's VP8L encoder for the colour planes, which is outside this path: code 3 before the device is touched."""
    from pngutil import pil_png
    p = L.default_params()
    p.webp_lossless = 1
    with pytest.raises(L.B200Error) as e:
        L.convert_in_memory(pil_png(synth(20, 30, 3, seed=1)), p, 3)
    assert e.value.code == 3


def _alpha_planes():
    rng = np.random.default_rng(1)
    yy, xx = np.mgrid[:200, :317]
    boxes = np.zeros((512, 640), np.uint8); boxes[100:400, 50:600] = 255; boxes[200:300, 200:300] = 128
    return {
        "flat": np.full((37, 53), 255, np.uint8), "one": np.full((1, 1), 7, np.uint8),
        "row": (np.arange(300) % 256).astype(np.uint8).reshape(1, 300), "col": (np.arange(300) % 256).astype(np.uint8).reshape(300, 1),
        "disc": np.clip(255 - np.hypot(yy - 100, xx - 150) * 2, 0, 255).astype(np.uint8), "noise": rng.integers(0, 256, (129, 257), dtype=np.uint8),
        "boxes": boxes, "tile": np.tile(rng.integers(0, 256, (8, 8), dtype=np.uint8), (40, 50)), "narrow": (rng.integers(0, 3, (400, 5)) * 100).astype(np.uint8),
        "ramp": (np.mgrid[:600, :1024][1] // 4).astype(np.uint8), "two": rng.integers(0, 2, (64, 64), dtype=np.uint8) * 255,
    }


def _alpha_residual(a, f):
    """numpy restatement of the ALPH prediction filters (row 0 from the left, column 0 from above)"""
    a = a.astype(np.int32); h, w = a.shape
    if f == 0: return a.astype(np.uint8)
    out = np.zeros_like(a)
    out[0, 0] = a[0, 0]; out[0, 1:] = a[0, 1:] - a[0, :-1]
    out[1:, 0] = a[1:, 0] - a[:-1, 0]
    if f == 1: out[1:, 1:] = a[1:, 1:] - a[1:, :-1]
    elif f == 2: out[1:, 1:] = a[1:, 1:] - a[:-1, 1:]
    else: out[1:, 1:] = a[1:, 1:] - np.clip(a[1:, :-1] + a[:-1, 1:] - a[:-1, :-1], 0, 255)
    return (out & 255).astype(np.uint8)


@pytest.mark.parametrize("name", sorted(_alpha_planes()))
def test_alpha_chunk_is_decoded_by_libwebp_to_the_plane(L, O, name):
    """The ALPH chunk writer (VP8L image stream from LZ77 tokens) and the VP8X container, pinned by libwebp's decoder: the alpha
    plane comes back exactly, the colour frame is untouched.  Tokens come from the oracle's K7 twin (the GPU test feeds K7's own)."""
    import io
    from PIL import Image
    a = _alpha_planes()[name]
    h, w = a.shape
    k, res = L.webp_alpha_filter(a)                                  # the prediction filter with the cheapest residuals, and the residuals
    tok, _ = O.png_lz77(res.reshape(-1), 1, w)
    alph = L.webp_alpha_chunk(tok, w, h, k)
    assert alph[0] == (1 | (k << 2))                                 # lossless compression, that filter, no pre-processing
    if name in ("disc", "ramp"): assert k != 0                       # smooth planes are predicted
    if name in ("flat", "noise"): assert k == 0
    rgb = synth(h, w, 3, seed=h + w)
    b = io.BytesIO(); Image.fromarray(rgb).save(b, "WEBP", quality=80); simple = b.getvalue()
    f = L.webp_wrap_alpha(simple, alph, w, h)
    assert f[12:16] == b"VP8X" and f[20] == 0x10 and len(f) % 2 == 0 and int.from_bytes(f[4:8], "little") == len(f) - 8
    im = Image.open(io.BytesIO(f)); im.load()
    got = np.asarray(im.convert("RGBA"))
    assert np.array_equal(got[:, :, 3], a)
    assert np.array_equal(got[:, :, :3], np.asarray(Image.open(io.BytesIO(simple)).convert("RGB")))
    if name in ("flat", "boxes", "ramp", "tile"):
        assert len(alph) < a.size // 20                               # copies are merged across K7's 258-byte / chunk limits
    with pytest.raises(L.B200Error):
        L.webp_alpha_chunk(tok[:-1], w, h, k)                         # tokens that do not cover the plane
    for f in (0, 1, 2, 3):                                            # every filter, whatever the chooser says: libwebp undoes it
        if f == k: continue
        resf = _alpha_residual(a, f)
        tokf, _ = O.png_lz77(resf.reshape(-1), 1, w)
        g = np.asarray(Image.open(io.BytesIO(L.webp_wrap_alpha(simple, L.webp_alpha_chunk(tokf, w, h, f), w, h))).convert("RGBA"))
        assert np.array_equal(g[:, :, 3], a), f


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


def _pil_rgba(data):
    from PIL import Image
    return np.asarray(Image.open(io.BytesIO(data)).convert("RGBA"))


def _lossless_sources():
    rng = np.random.default_rng(11)
    yy, xx = np.mgrid[:97, :131]
    pal = lambda n: rng.integers(0, 256, (n, 3), dtype=np.uint8)[rng.integers(0, n, (97, 131))]
    ramp = np.stack([xx * 2 % 256, yy * 2 % 256, (xx + yy) % 256], axis=2).astype(np.uint8)
    return {
        "photo": synth(97, 131, 3, seed=1, kind="photo"), "flat": synth(64, 64, 3, seed=2, kind="flat"), "noise": synth(33, 47, 3, seed=3, kind="noise"),
        "pal2": pal(2), "pal4": pal(4), "pal16": pal(16), "pal200": pal(200), "ramp": ramp, "one": synth(1, 1, 3, seed=4), "wide": synth(3, 700, 3, seed=5, kind="photo"),
        "rgba": np.concatenate([synth(97, 131, 3, seed=6, kind="photo"), (np.hypot(yy - 48, xx - 65) * 4).clip(0, 255).astype(np.uint8)[:, :, None]], axis=2),
        "rgba_noise": synth(40, 50, 4, seed=7, kind="noise"), "big": synth(400, 600, 3, seed=8, kind="photo"),
    }


@pytest.mark.parametrize("name", sorted(_lossless_sources()))
def test_vp8l_decoder_equals_libwebp_on_lossless_files(L, name):
    """Lossless WebP input (VP8L): prefix codes, LZ77 with the neighbourhood distance codes, colour cache, meta prefix image and the
    four transforms (predictor, cross-colour, subtract-green, colour indexing with pixel bundling), against libwebp on files libwebp
    wrote at every effort level (the level decides which of those tools a file uses)."""
    from PIL import Image
    img = _lossless_sources()[name]
    for method, quality in ((0, 0), (1, 25), (3, 50), (4, 75), (6, 100)):
        b = io.BytesIO(); Image.fromarray(img).save(b, "WEBP", lossless=True, method=method, quality=quality, exact=True)
        data = b.getvalue()
        rgb, alpha = L.webp_decode_rgba(data)
        want = _pil_rgba(data)
        assert np.array_equal(rgb, want[:, :, :3]), (name, method)
        if (want[:, :, 3] != 255).any(): assert np.array_equal(alpha, want[:, :, 3]), (name, method)
        else: assert alpha is None
        assert np.array_equal(L.webp_decode(data), want[:, :, :3])                 # the RGB-only entry point drops the alpha plane


@pytest.mark.parametrize("kind", ["disc", "noise", "steps", "text"])
def test_alpha_plane_of_lossy_files_equals_libwebp(L, kind):
    """VP8X + ALPH input: raw / VP8L-coded planes, level-quantised (alpha_quality < 100) or exact, with whichever prediction filter
    libwebp picked (none / horizontal / vertical / gradient)."""
    from PIL import Image
    rng = np.random.default_rng(5)
    h, w = 120, 167
    yy, xx = np.mgrid[:h, :w]
    a = {"disc": (np.hypot(yy - 60, xx - 80) * 3).clip(0, 255), "noise": rng.integers(0, 256, (h, w)), "steps": (xx // 8 * 16) % 256,
         "text": np.where(rng.random((h, w)) < 0.1, 0, 255)}[kind].astype(np.uint8)
    img = np.concatenate([synth(h, w, 3, seed=2, kind="photo"), a[:, :, None]], axis=2)
    seen_filters = set()
    for aq, method in ((100, 4), (100, 6), (60, 4), (20, 2), (0, 0)):
        b = io.BytesIO(); Image.fromarray(img).save(b, "WEBP", quality=75, alpha_quality=aq, method=method)
        data = b.getvalue()
        i = data.find(b"ALPH"); assert i > 0
        seen_filters.add((data[i + 8] >> 2) & 3)
        rgb, alpha = L.webp_decode_rgba(data)
        want = _pil_rgba(data)
        assert np.array_equal(alpha, want[:, :, 3]), (kind, aq)
        assert np.array_equal(rgb, want[:, :, :3]), (kind, aq)
    assert seen_filters                                                             # (which filters occur depends on the content)


def test_vp8_decoder_refuses_what_it_does_not_decode(L):
    from PIL import Image
    img = synth(20, 30, 4, seed=1)
    frames = [Image.fromarray(synth(20, 30, 3, seed=s)) for s in (1, 2)]
    b = io.BytesIO(); frames[0].save(b, "WEBP", save_all=True, append_images=frames[1:], duration=100)
    with pytest.raises(L.B200Error) as e:
        L.webp_decode(b.getvalue())
    assert e.value.code == 3 and "animated" in str(e.value)
    b = io.BytesIO(); Image.fromarray(img).save(b, "WEBP", lossless=True)
    data = b.getvalue()
    for cut in (len(data) // 2, len(data) - 3, 30):
        with pytest.raises(L.B200Error) as e:
            L.webp_decode(data[:cut])                     # a truncated lossless stream is corrupt, not zero-filled
        assert e.value.code == 4
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
