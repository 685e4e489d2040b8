"""GPU parity tests of the lossless PNG leg: K6 row filtering and K7 LZ77 against the oracle
(bit-exact filter choices, filtered bytes, tokens and histograms), and the whole b200_compress_in_memory PNG path against
the one property that defines it -- the output decodes (libpng via Pillow) to exactly the input's pixels."""
import io
import zlib

import numpy as np
import pytest

from pngutil import idat_stream, pil_pixels, pil_png, synth

pytestmark = pytest.mark.gpu


def lossless_params(L, level=3, keep_metadata=False):
    p = L.default_params()
    p.png_optimize = 1
    p.png_optimization_level = level
    p.keep_metadata = 1 if keep_metadata else 0
    return p


@pytest.mark.parametrize("shape", [(37, 53), (1, 1), (3, 700), (130, 9), (64, 1024)])
@pytest.mark.parametrize("channels", [1, 2, 3, 4])
def test_filter_kernel_matches_oracle(L, O, shape, channels):
    h, w = shape
    for kind in ("photo", "flat"):
        raw = synth(h, w, channels, seed=h + w + channels, kind=kind).reshape(h, w * channels)
        for s in range(10):
            got = L.png_filter(raw, channels, s)
            want = O.png_filter(raw, channels, s)
            assert np.array_equal(got[:, 0], want[:, 0]), (kind, s, "filter choice")
            assert np.array_equal(got, want), (kind, s)


def test_filter_kernel_16bit_and_subbyte_rows(L, O):
    rng = np.random.default_rng(1)
    raw = rng.integers(0, 256, (40, 2 * 3 * 31)).astype(np.uint8)         # 16-bit RGB: bpp 6
    for s in range(10):
        assert np.array_equal(L.png_filter(raw, 6, s), O.png_filter(raw, 6, s))
    raw = rng.integers(0, 256, (40, 13)).astype(np.uint8)                 # packed 1/2/4-bit rows: bpp 1
    for s in range(10):
        assert np.array_equal(L.png_filter(raw, 1, s), O.png_filter(raw, 1, s))


@pytest.mark.parametrize("kind,channels,shape", [("photo", 3, (64, 97)), ("flat", 3, (200, 333)), ("flat", 1, (100, 4100)), ("noise", 4, (50, 50)),
                                                  ("photo", 2, (31, 17)), ("flat", 4, (1, 5)), ("flat", 1, (1, 2))])
def test_lz77_kernels_match_oracle(L, O, kind, channels, shape):
    h, w = shape
    img = synth(h, w, channels, seed=3, kind=kind)
    filt = O.png_filter(img.reshape(h, -1), channels, 4 if kind == "photo" else 0)
    stream = filt.reshape(-1)
    tok, hist = L.png_lz77(stream, channels, filt.shape[1])
    wtok, whist = O.png_lz77(stream, channels, filt.shape[1])
    assert np.array_equal(hist, whist)
    assert np.array_equal(tok, wtok)
    assert np.array_equal(O.png_expand(tok, stream.size), stream)


def _check_lossless(L, png_in, level=3, expect_mode=None):
    out = L.compress_in_memory(png_in, lossless_params(L, level))
    a, b = pil_pixels(png_in), pil_pixels(out)
    assert a.size == b.size
    if expect_mode:
        assert b.mode == expect_mode
    if a.mode != b.mode:                      # a colour-type reduction: compare in the richer mode
        b = b.convert(a.mode)
    assert np.array_equal(np.asarray(a), np.asarray(b))
    ihdr, idat, order = idat_stream(out)
    assert order[0] == b"IHDR" and order[-1] == b"IEND"
    filt = zlib.decompress(idat)              # a complete, valid zlib stream with a correct Adler-32
    return out, ihdr, filt


@pytest.mark.parametrize("level", [0, 1, 2, 3, 5, 6])
def test_compress_png_is_lossless_at_every_level(L, level):
    img = synth(120, 160, 3, seed=level, kind="photo")
    out, ihdr, filt = _check_lossless(L, pil_png(img), level)
    assert len(filt) == 120 * (160 * 3 + 1)
    if level == 0:
        assert set(filt[::160 * 3 + 1]) == {0}


@pytest.mark.parametrize("channels,kind", [(1, "photo"), (2, "photo"), (3, "flat"), (4, "photo"), (4, "flat")])
def test_compress_png_colour_types(L, channels, kind):
    _check_lossless(L, pil_png(synth(77, 131, channels, seed=channels, kind=kind)))


def test_compress_png_smaller_than_a_naive_encoder(L):
    img = synth(256, 256, 3, seed=8, kind="photo")
    src = pil_png(img, compress_level=1)
    out, _, _ = _check_lossless(L, src)
    assert len(out) < len(src)
    flat = synth(256, 256, 3, seed=8, kind="flat")
    out2, _, _ = _check_lossless(L, pil_png(flat, compress_level=1))
    assert len(out2) < 256 * 256 * 3 // 20


def test_compress_png_reductions(L):
    rgb = synth(60, 80, 3, seed=1)
    opaque = np.concatenate([rgb, np.full((60, 80, 1), 255, np.uint8)], axis=2)
    _check_lossless(L, pil_png(opaque), expect_mode="RGB")                   # opaque alpha dropped
    grey3 = np.repeat(synth(60, 80, 1, seed=2), 3, axis=2)
    _check_lossless(L, pil_png(grey3), expect_mode="L")                      # r == g == b -> greyscale
    grey4 = np.concatenate([grey3, synth(60, 80, 1, seed=5)], axis=2)
    _check_lossless(L, pil_png(grey4), expect_mode="LA")                     # grey + real alpha
    both = np.concatenate([grey3, np.full((60, 80, 1), 255, np.uint8)], axis=2)
    _check_lossless(L, pil_png(both), expect_mode="L")
    real = np.concatenate([rgb, synth(60, 80, 1, seed=6)], axis=2)
    _check_lossless(L, pil_png(real), expect_mode="RGBA")                    # nothing to reduce


def test_compress_png_palette_reduction(L):
    """<= 256 colours: the file comes back indexed (PLTE + tRNS), pixel-exact, and far smaller than one byte per sample"""
    for ch in (3, 4):
        img = synth(150, 220, ch, seed=40 + ch, kind="flat")
        if ch == 4:
            img[20:60, 30:90, 3] = 0; img[100:110, :, 3] = 77
        out, ihdr, filt = _check_lossless(L, pil_png(img), expect_mode="P")
        n = len(np.unique(img.reshape(-1, ch), axis=0))
        depth = 1 if n <= 2 else 2 if n <= 4 else 4 if n <= 16 else 8
        assert ihdr[2] == depth and ihdr[3] == 3 and len(filt) == 150 * ((220 * depth + 7) // 8 + 1)
        assert len(out) < 150 * 220 // 8


def test_compress_png_palette_16bit_and_bilevel(L):
    from PIL import Image
    rng = np.random.default_rng(3)
    idx = (synth(50, 70, 1, seed=3, kind="flat")[:, :, 0] % 16).astype(np.uint8)
    im = Image.fromarray(idx, mode="P"); im.putpalette([int(v) for v in rng.integers(0, 256, 48)])
    src = pil_png(im)
    out = L.compress_in_memory(src, lossless_params(L))
    assert np.array_equal(np.asarray(pil_pixels(src).convert("RGB")), np.asarray(pil_pixels(out).convert("RGB")))
    a16 = (synth(40, 40, 1, seed=4)[:, :, 0].astype(np.uint16) * 257) ^ 0x0103
    src = pil_png(Image.fromarray(a16))
    out = L.compress_in_memory(src, lossless_params(L))
    assert np.array_equal(np.asarray(pil_pixels(src)), np.asarray(pil_pixels(out)))
    bw = rng.integers(0, 2, (33, 47)).astype(bool)
    src = pil_png(Image.fromarray(bw))
    out = L.compress_in_memory(src, lossless_params(L))
    assert np.array_equal(np.asarray(pil_pixels(src)), np.asarray(pil_pixels(out)))


def test_compress_png_transparency_chunk_survives(L):
    from PIL import Image
    idx = (synth(20, 20, 1, seed=3, kind="flat")[:, :, 0] % 4).astype(np.uint8)
    im = Image.fromarray(idx, mode="P"); im.putpalette([0, 0, 0, 255, 0, 0, 0, 255, 0, 0, 0, 255])
    b = io.BytesIO(); im.save(b, format="PNG", transparency=bytes([0, 128, 255, 255]))
    out = L.compress_in_memory(b.getvalue(), lossless_params(L))
    assert np.array_equal(np.asarray(pil_pixels(b.getvalue()).convert("RGBA")), np.asarray(pil_pixels(out).convert("RGBA")))


def test_compress_png_metadata_policy(L):
    from PIL import Image
    from PIL.PngImagePlugin import PngInfo
    meta = PngInfo(); meta.add_text("Comment", "hello from the test")
    b = io.BytesIO(); Image.fromarray(synth(16, 16, 3)).save(b, format="PNG", pnginfo=meta, dpi=(300, 300))
    _, _, kept = idat_stream(L.compress_in_memory(b.getvalue(), lossless_params(L, keep_metadata=True)))
    _, _, stripped = idat_stream(L.compress_in_memory(b.getvalue(), lossless_params(L, keep_metadata=False)))
    assert b"tEXt" in kept and b"tEXt" not in stripped
    assert b"pHYs" in kept and b"pHYs" in stripped          # oxipng StripChunks::Safe keeps pHYs


def test_compress_png_large_image(L):
    img = synth(1080, 1920, 3, seed=2, kind="photo")
    src = pil_png(img, compress_level=1)
    out, _, filt = _check_lossless(L, src)
    assert len(filt) == 1080 * (1920 * 3 + 1) and len(out) < len(src)


def test_png_paths_outside_the_gpu_build_are_refused(L):
    src = pil_png(synth(16, 16, 3))
    p = L.default_params(); p.png_optimize = 0                      # lossy = imagequant
    with pytest.raises(L.B200Error) as e:
        L.compress_in_memory(src, p)
    assert e.value.code == 3
    p = lossless_params(L); p.width = 8
    with pytest.raises(L.B200Error) as e:
        L.compress_in_memory(src, p)
    assert e.value.code == 3
    with pytest.raises(L.B200Error) as e:
        L.compress_in_memory(src[:50], lossless_params(L))
    assert e.value.code == 4


def test_compress_batch_mixes_png_and_jpeg(L, golden):
    items = [pil_png(synth(64, 64, 3, seed=i)) for i in range(3)] + [golden("in_420_base_355x237.jpg")]
    p = lossless_params(L)
    outs = L.compress_batch(items, p)
    for src, (data, code, msg) in zip(items, outs):
        assert code == 0, msg
        if src[:2] == b"\xff\xd8":
            assert data[:2] == b"\xff\xd8"
        else:
            assert np.array_equal(np.asarray(pil_pixels(src)), np.asarray(pil_pixels(data)))
