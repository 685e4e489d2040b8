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
