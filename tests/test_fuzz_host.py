"""Mutation fuzzing of the host parsers behind the C ABI (JPEG marker/entropy reader, PNG container + inflate + unfilter):
corrupt, truncated and spliced inputs must come back as a status (usually B200_ERR_CORRUPT_INPUT) or decode -- never crash,
hang or unwind across the boundary.  The reference's rule: one bad file fails that file only (compressor.rs:81-101)."""
import random
import zlib

import numpy as np
import pytest

from pngutil import frame_png, pil_png, synth


def _mutate(rng, src):
    b = bytearray(src)
    mode = rng.random()
    if mode < 0.5:
        for _ in range(rng.randint(1, 6)):
            b[rng.randrange(len(b))] = rng.randrange(256)
    elif mode < 0.75:
        b = b[:rng.randrange(1, len(b))]
    else:
        i = rng.randrange(len(b))
        b[i:i] = bytes(rng.randrange(256) for _ in range(rng.randint(1, 40)))
    return bytes(b)


@pytest.mark.parametrize("seed", [1, 2])
def test_mutated_jpeg_and_png_files_never_crash(L, golden, seed):
    rng = random.Random(seed)
    sources = [golden("in_420_base_355x237.jpg"), golden("in_420_prog_355x237.jpg"), golden("in_gray_base_355x237.jpg"),
               pil_png(synth(40, 50, 3, seed=1)), pil_png(synth(40, 50, 4, seed=2, kind="flat"))]
    p = L.default_params()
    p.jpeg_optimize = 1                     # the transcode runs on the host when no device is present
    seen = set()
    for _ in range(400):
        b = _mutate(rng, rng.choice(sources))
        calls = [lambda: L.png_decode(b)] if b[:4] == b"\x89PNG" else [lambda: L.jpeg_decode_coefficients(b), lambda: L.compress_in_memory(b, p)] if b[:2] == b"\xff\xd8" else [lambda: L.compress_in_memory(b, p)]
        for fn in calls:
            try:
                fn()
                seen.add(0)
            except L.B200Error as e:
                assert e.code in (2, 3, 4), (e.code, str(e))
                seen.add(e.code)
    assert 4 in seen and 0 in seen


def test_mutated_deflate_streams_with_valid_crcs_never_crash(L):
    """Chunk CRCs recomputed after the mutation, so the damage reaches inflate, the Adler-32 check and the unfilter loops."""
    rng = random.Random(7)
    nprng = np.random.default_rng(7)
    ok = bad = 0
    for it in range(500):
        ch = rng.choice([1, 2, 3, 4]); h = rng.randint(1, 40); w = rng.randint(1, 60)
        img = synth(h, w, ch, seed=it, kind=rng.choice(["photo", "flat", "noise"]))
        rows = np.concatenate([nprng.integers(0, 5, (h, 1)).astype(np.uint8), img.reshape(h, -1)], axis=1)
        z = bytearray(zlib.compress(rows.tobytes(), rng.choice([0, 1, 6, 9])))
        m = rng.random()
        if m < 0.6:
            for _ in range(rng.randint(1, 4)):
                z[rng.randrange(len(z))] = rng.randrange(256)
        elif m < 0.8:
            z = z[:rng.randrange(1, len(z))]
        png = frame_png(w + rng.choice([0, 0, 0, 1]), h, 8, {1: 0, 2: 4, 3: 2, 4: 6}[ch], bytes(z))
        try:
            info, raw = L.png_decode(png)
            ok += 1
            if m >= 0.8 and info.width == w:
                assert np.array_equal(raw.reshape(h, w, ch), img) or rows[:, 0].any()      # untouched stream: filter type 0 rows decode to the source
        except L.B200Error as e:
            assert e.code == 4, (e.code, str(e))
            bad += 1
    assert ok > 20 and bad > 100


# ---- round-1 ADVICE items: each was a crash or a silent wrong answer, now a status -------------------------------------------
def test_png_idat_longer_than_ihdr_implies_is_refused(L):
    """A 1x1 image whose IDAT inflates to far more than (row_bytes + 1) * height used to run the inflate's checked path past its
    buffer (heap overflow) and doubled as a decompression bomb; it must come back as corrupt input."""
    for payload in (bytes(range(256)) * 40, b"\x00" * 300000, np.random.default_rng(3).integers(0, 256, 70000, dtype=np.uint8).tobytes()):
        for lvl in (1, 6, 9):
            png = frame_png(1, 1, 8, 0, zlib.compress(payload, lvl))
            with pytest.raises(L.B200Error) as e:
                L.png_decode(png)
            assert e.value.code == 4 and "too long" in str(e.value)
    # the exact size still decodes; a stored (type 0) block that overshoots is refused as well
    info, raw = L.png_decode(frame_png(3, 2, 8, 0, zlib.compress(b"\x00abc\x00def", 6)))
    assert raw.tobytes() == b"abcdef"
    with pytest.raises(L.B200Error):
        L.png_decode(frame_png(1, 1, 8, 0, zlib.compress(b"\x00" * 5000, 0)))


def test_png_illegal_bit_depths_are_refused(L):
    """IHDR (colour type, bit depth) pairs outside PNG 11.2.2 -- depth 0 divided by zero in the grey expansion, RGB at depth 4
    indexed samples with bd / 8 == 0 -- are corrupt input, through both the stage entry point and the conversion path."""
    legal = {0: (1, 2, 4, 8, 16), 2: (8, 16), 3: (1, 2, 4, 8), 4: (8, 16), 6: (8, 16)}
    p = L.default_params()
    for ct in (0, 2, 3, 4, 6):
        for bd in (0, 1, 2, 3, 4, 5, 8, 12, 16, 32, 255):
            png = frame_png(2, 2, bd, ct, zlib.compress(b"\x00" * 64))
            if bd in legal[ct]:
                continue
            with pytest.raises(L.B200Error) as e:
                L.png_decode(png)
            assert e.value.code == 4 and "bit depth" in str(e.value)
            with pytest.raises(L.B200Error) as e:
                L.convert_in_memory(png, p, 0)          # PNG -> JPEG: the expansion that used to SIGFPE
            assert e.value.code == 4
    with pytest.raises(L.B200Error):
        L.png_decode(frame_png(0x80000000, 1, 8, 0, zlib.compress(b"\x00")))


def test_baseline_components_spread_over_scans_keep_their_data(L, golden):
    """Non-interleaved baseline file (one scan per component): every scan must keep what the earlier ones decoded."""
    import io
    from PIL import Image
    src = golden("in_444_base_355x237.jpg")
    lay, coefs = L.jpeg_decode_coefficients(src)
    # host encoder, sequential: one interleaved scan; split it by re-encoding each component as its own greyscale file and splicing
    # the three scans behind one 3-component frame header
    im = Image.open(io.BytesIO(src)); im.draft("YCbCr", im.size)
    planes = np.asarray(im.convert("YCbCr") if im.mode != "YCbCr" else im)
    qt = {0: list(np.asarray(lay.qt[0], dtype=int)), 1: list(np.asarray(lay.qt[1], dtype=int))}
    parts = []
    for c in range(3):
        b = io.BytesIO()
        Image.fromarray(np.ascontiguousarray(planes[:, :, c])).save(b, "JPEG", qtables=[_natural(qt[0 if c == 0 else 1])], optimize=False)
        parts.append(b.getvalue())
    spliced = _splice_scans(parts, lay)
    lay2, coefs2 = L.jpeg_decode_coefficients(spliced)
    assert lay2.ncomp == 3 and lay2.total_coefs == lay.total_coefs
    for c in range(3):
        g = _grey_coefs(L, parts[c])
        off = lay2.comp_offset[c]
        assert np.array_equal(coefs2[off:off + g.size], g), f"component {c} was wiped or not decoded"
    # a file whose last component is never coded: that component reads as zeros, not as stale memory
    lay3, coefs3 = L.jpeg_decode_coefficients(_splice_scans(parts[:2], lay, declare=3))
    assert not coefs3[lay3.comp_offset[2]:].any()
    assert np.array_equal(coefs3[:lay3.comp_offset[2]], coefs2[:lay2.comp_offset[2]])


_ZZ = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36,
       29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63]


def _natural(zz_table):
    out = [0] * 64
    for z, k in enumerate(_ZZ):
        out[k] = int(zz_table[z])
    return out


def _segments(jpg):
    """[(marker, payload-with-length)] up to SOS, then the entropy-coded bytes of the first scan (to EOI)."""
    i, segs = 2, []
    while True:
        assert jpg[i] == 0xFF
        m = jpg[i + 1]; n = (jpg[i + 2] << 8) | jpg[i + 3]
        segs.append((m, jpg[i + 2:i + 2 + n]))
        i += 2 + n
        if m == 0xDA:
            end = jpg.rindex(b"\xff\xd9")
            return segs, jpg[i:end]


def _splice_scans(grey_files, lay, declare=None):
    """One 3-component 4:4:4 baseline frame whose scans are the single scans of the given greyscale files (component ids 1..3,
    table ids as in each file: each scan is preceded by its own DHT segments)."""
    ncomp = declare or len(grey_files)
    out = bytearray(b"\xff\xd8")
    segs0, _ = _segments(grey_files[0])
    for m, body in segs0:
        if m == 0xDB:
            out += b"\xff\xdb" + body
    if len(grey_files) > 1:
        for m, body in _segments(grey_files[1])[0]:
            if m == 0xDB:
                b2 = bytearray(body); b2[2] = (b2[2] & 0xF0) | 1          # chroma table -> slot 1
                out += b"\xff\xdb" + bytes(b2)
    sof = bytearray([0, 0, 8, lay.height >> 8, lay.height & 255, lay.width >> 8, lay.width & 255, ncomp])
    for c in range(ncomp):
        sof += bytes([c + 1, 0x11, 0 if c == 0 else 1])
    sof[0:2] = len(sof).to_bytes(2, "big")
    out += b"\xff\xc0" + sof
    for c, f in enumerate(grey_files):
        segs, ecs = _segments(f)
        for m, body in segs:
            if m == 0xC4:
                out += b"\xff\xc4" + body
        out += b"\xff\xda" + bytes([0, 8, 1, c + 1, 0x00, 0, 63, 0]) + ecs
    return bytes(out + b"\xff\xd9")


def _grey_coefs(L, grey_file):
    lay, coefs = L.jpeg_decode_coefficients(grey_file)
    return coefs[:lay.total_coefs]


def test_adobe_rgb_jpeg_is_handed_back_not_retagged(L, golden):
    """A 3-component file with an Adobe APP14 marker saying transform 0 carries RGB, not YCbCr: re-writing it with a JFIF header
    would change its colours, so the path answers code 3 (the host routes it to libcaesium); transform 1 is ordinary YCbCr."""
    src = golden("in_444_base_355x237.jpg")
    assert src[2:4] == b"\xff\xe0"
    jfif_len = (src[4] << 8) | src[5]
    rest = src[4 + jfif_len:]

    def adobe(transform):
        return b"\xff\xd8" + b"\xff\xee\x00\x0eAdobe\x00\x64\x00\x00\x00\x00" + bytes([transform]) + rest
    p = L.default_params(); p.jpeg_optimize = 1
    with pytest.raises(L.B200Error) as e:
        L.compress_in_memory(adobe(0), p)
    assert e.value.code == 3 and "RGB" in str(e.value)
    with pytest.raises(L.B200Error) as e:
        L.jpeg_decode_coefficients(adobe(0))
    assert e.value.code == 3
    out = L.compress_in_memory(adobe(1), p)            # YCbCr with an Adobe marker: transcoded as usual
    assert out[:2] == b"\xff\xd8"
    # component ids 'R','G','B' without JFIF / Adobe: libjpeg's heuristic says RGB as well
    i = rest.index(b"\xff\xc0")
    rgb_ids = bytearray(rest); rgb_ids[i + 10] = ord("R"); rgb_ids[i + 13] = ord("G"); rgb_ids[i + 16] = ord("B")
    j = rgb_ids.index(b"\xff\xda"); rgb_ids[j + 5] = ord("R"); rgb_ids[j + 7] = ord("G"); rgb_ids[j + 9] = ord("B")
    with pytest.raises(L.B200Error) as e:
        L.compress_in_memory(b"\xff\xd8" + bytes(rgb_ids), p)
    assert e.value.code == 3


@pytest.mark.parametrize("seed", [3, 4])
def test_mutated_webp_files_never_crash(L, seed):
    """The WebP front end (container, VP8 key-frame decoder, VP8L decoder, ALPH chunk): mutated lossless files, lossy files with an alpha
    plane and plain lossy files come back as pixels or as a status."""
    import io
    from PIL import Image
    rng = random.Random(seed)
    rgba = synth(40, 56, 4, seed=seed, kind="photo")
    pal = np.random.default_rng(seed).integers(0, 256, (4, 3), dtype=np.uint8)[np.random.default_rng(seed + 1).integers(0, 4, (40, 56))]
    srcs = []
    for img, kw in ((rgba, dict(lossless=True, method=4)), (pal, dict(lossless=True, method=6)), (rgba, dict(quality=70, alpha_quality=50)), (rgba[:, :, :3].copy(), dict(quality=60))):
        b = io.BytesIO(); Image.fromarray(img).save(b, "WEBP", **kw); srcs.append(b.getvalue())
    done = 0
    for k in range(600):
        src = srcs[k % len(srcs)]
        data = _mutate(rng, src)
        if k % 3 == 0 and len(data) > 40:        # keep the container intact, damage the payload only
            data = src[:30] + data[30:]
        try:
            rgb, alpha = L.webp_decode_rgba(data)
            assert rgb.ndim == 3 and (alpha is None or alpha.shape == rgb.shape[:2])
        except L.B200Error as e:
            assert e.code in (3, 4, 5), e
        done += 1
    assert done == 600
