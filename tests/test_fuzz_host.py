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
