"""Test helper: what libwebp's simple-API decoder (WebPDecodeRGB, used by Pillow) does AFTER VP8 reconstruction -- the
'fancy' (bilinear 9-3-3-1) chroma upsampler and the 14-bit fixed-point YUV->RGB conversion -- restated in numpy so that a
decoded RGB image can be compared bit for bit with an encoder's own YUV reconstruction."""
import io

import numpy as np


def _clip8(v):
    return np.where((v & ~16383) == 0, v >> 6, np.where(v < 0, 0, 255)).astype(np.uint8)


def yuv_to_rgb(y, u, v):
    y = y.astype(np.int32); u = u.astype(np.int32); v = v.astype(np.int32)
    r = _clip8(((y * 19077) >> 8) + ((v * 26149) >> 8) - 14234)
    g = _clip8(((y * 19077) >> 8) - ((u * 6419) >> 8) - ((v * 13320) >> 8) + 8708)
    b = _clip8(((y * 19077) >> 8) + ((u * 33050) >> 8) - 17685)
    return np.stack([r, g, b], -1)


def _upsample_pair(top, cur, w):
    """One pair of chroma rows (nearer row `a`, farther row `b`) -> full-width chroma for the luma row nearer to `a`
    (UpsampleRgbLinePair: per sample (9a + 3a' + 3b + b' + 8) >> 4 computed as two staged averages)."""
    a = top.astype(np.int32); b = cur.astype(np.int32)
    out = np.zeros(w, np.int32)
    out[0] = (3 * a[0] + b[0] + 2) >> 2
    n = (w - 1) >> 1                                   # last_pixel_pair
    if n > 0:
        tl, t, l, c = a[:n], a[1:n + 1], b[:n], b[1:n + 1]
        avg = tl + t + l + c + 8
        d12 = (avg + 2 * (t + l)) >> 3
        d03 = (avg + 2 * (tl + c)) >> 3
        out[1:2 * n:2] = (d12 + tl) >> 1
        out[2:2 * n + 1:2] = (d03 + t) >> 1
    if not (w & 1):
        out[w - 1] = (3 * a[(w - 1) >> 1] + b[(w - 1) >> 1] + 2) >> 2
    return out


def fancy_upsample(c, w, h):
    """chroma plane [ceil(h/2), ceil(w/2)] -> [h, w] the way EmitFancyRGB walks the rows."""
    out = np.zeros((h, w), np.int32)
    ch = c.shape[0]
    for y in range(h):
        near = y >> 1
        far = near - 1 if (y & 1) == 0 else near + 1
        far = min(max(far, 0), ch - 1)
        out[y] = _upsample_pair(c[near, :(w + 1) // 2], c[far, :(w + 1) // 2], w)
    return out


def decode_like_libwebp(Y, U, V, w, h):
    """Encoder reconstruction (macroblock-padded planes) -> the RGB image libwebp's decoder would hand to Pillow."""
    cu = fancy_upsample(U[:(h + 1) // 2], w, h)
    cv = fancy_upsample(V[:(h + 1) // 2], w, h)
    return yuv_to_rgb(Y[:h, :w], cu, cv)


def pil_decode(data):
    from PIL import Image
    im = Image.open(io.BytesIO(data))
    assert im.format == "WEBP"
    return np.asarray(im.convert("RGB"))
