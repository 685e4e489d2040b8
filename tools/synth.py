"""Seeded synthetic natural-image-like inputs: smooth gradient + 1/f^2 noise with correlated
channels + hard-edged shapes + sigma=2 sensor noise.  Data preparation for tests/bench, not product code."""
import io

import numpy as np

SEED0 = 0xCAE51


def synth_rgb(width, height, index=0):
    rng = np.random.default_rng(SEED0 + index)
    x = (np.arange(width, dtype=np.float32) / max(width - 1, 1))[None, :]
    y = (np.arange(height, dtype=np.float32) / max(height - 1, 1))[:, None]
    img = np.empty((height, width, 3), dtype=np.float32)
    img[:, :, 0] = 255.0 * (0.2 + 0.6 * x)
    img[:, :, 1] = 255.0 * (0.8 - 0.5 * y)
    img[:, :, 2] = 255.0 * (0.3 + 0.4 * (x * (width - 1) + y * (height - 1)) / max(width + height - 2, 1))
    # 1/f^2 (power) == 1/f (amplitude) noise, generated at 1/4 resolution then replicated 4x4, for speed
    h4, w4 = (height + 3) // 4, (width + 3) // 4
    fy = np.fft.fftfreq(h4)[:, None]
    fx = np.fft.rfftfreq(w4)[None, :]
    f = np.sqrt(fx * fx + fy * fy)
    f[0, 0] = 1.0
    common = None
    for c in range(3):
        spec = (rng.standard_normal((h4, w4 // 2 + 1)) + 1j * rng.standard_normal((h4, w4 // 2 + 1))) / f
        spec[0, 0] = 0
        n = np.fft.irfft2(spec, s=(h4, w4)).astype(np.float32)
        n /= n.std() + 1e-6
        if common is None:
            common = n
        n = 40.0 * (0.8 * common + 0.6 * n)          # ~0.8 inter-channel correlation
        img[:, :, c] += np.repeat(np.repeat(n, 4, axis=0), 4, axis=1)[:height, :width]
    for _ in range(int(rng.integers(20, 61))):
        x0, y0 = int(rng.integers(0, width)), int(rng.integers(0, height))
        w, h = int(rng.integers(8, max(9, width // 6))), int(rng.integers(8, max(9, height // 6)))
        col = rng.integers(0, 256, size=3).astype(np.float32)
        if rng.random() < 0.5:
            img[y0:y0 + h, x0:x0 + w] = col
        else:
            ya, yb, xa, xb = max(0, y0 - h), min(height, y0 + h), max(0, x0 - w), min(width, x0 + w)
            yy, xx = np.ogrid[ya:yb, xa:xb]
            m = ((yy - y0) / float(h)) ** 2 + ((xx - x0) / float(w)) ** 2 <= 1.0
            img[ya:yb, xa:xb][m] = col
    img += 2.0 * rng.standard_normal(img.shape, dtype=np.float32)
    np.clip(img + 0.5, 0, 255, out=img)
    return img.astype(np.uint8)


def synth_jpeg(width, height, index=0, quality=90, subsampling="4:2:0", progressive=False):
    """Source JPEG as BASELINE config 2 specifies: q=90, 4:2:0, baseline, Annex-K tables (libjpeg-turbo via Pillow)."""
    from PIL import Image
    b = io.BytesIO()
    Image.fromarray(synth_rgb(width, height, index), "RGB").save(b, "JPEG", quality=quality, subsampling=subsampling, progressive=progressive)
    return b.getvalue()


def synth_png_rgba(width, height, index=0, level=6):
    """Source PNG as BASELINE config 4 specifies: the same generator + a smooth-gradient alpha channel (every
    tenth image fully opaque), 8-bit RGBA, every row Paeth-filtered, zlib level 6, one IDAT."""
    import struct
    import zlib
    rgb = synth_rgb(width, height, index)
    rng = np.random.default_rng(SEED0 + 7919 * (index + 1))
    if index % 10 == 9:
        alpha = np.full((height, width), 255, np.uint8)
    else:
        gx, gy = rng.random() * 2 - 1, rng.random() * 2 - 1
        x = np.arange(width, dtype=np.float32)[None, :] / max(width - 1, 1)
        y = np.arange(height, dtype=np.float32)[:, None] / max(height - 1, 1)
        a = 160.0 + 90.0 * (gx * (x - 0.5) + gy * (y - 0.5))
        alpha = np.clip(a + 0.5, 0, 255).astype(np.uint8)
    img = np.concatenate([rgb, alpha[:, :, None]], axis=2)
    co = zlib.compressobj(level)
    parts = []
    strip = 128
    prev = np.zeros((1, width, 4), np.int16)
    for y0 in range(0, height, strip):
        cur = img[y0:y0 + strip].astype(np.int16)
        up = np.concatenate([prev, cur[:-1]], axis=0)
        left = np.concatenate([np.zeros((cur.shape[0], 1, 4), np.int16), cur[:, :-1]], axis=1)
        upleft = np.concatenate([np.zeros((cur.shape[0], 1, 4), np.int16), up[:, :-1]], axis=1)
        p = left + up - upleft
        pa, pb, pc = np.abs(p - left), np.abs(p - up), np.abs(p - upleft)
        pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, up, upleft))
        f = ((cur - pred) & 0xFF).astype(np.uint8).reshape(cur.shape[0], width * 4)
        rows = np.concatenate([np.full((cur.shape[0], 1), 4, np.uint8), f], axis=1)
        parts.append(co.compress(rows.tobytes()))
        prev = cur[-1:]
    parts.append(co.flush())
    z = b"".join(parts)

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)
    return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", width, height, 8, 6, 0, 0, 0)) + chunk(b"IDAT", z) + chunk(b"IEND", b"")
