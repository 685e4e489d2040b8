"""Helpers for the PNG parity tests: synthetic images, a minimal PNG framer around zlib, and decoding through Pillow."""
import io
import struct
import zlib

import numpy as np


def synth(h, w, channels, seed=0, kind="photo"):
    """uint8 [h, w, channels] test picture: 'photo' = smooth gradients + mild noise, 'flat' = few-colour blocks + lines."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    if kind == "photo":
        base = [128 + 100 * np.sin(xx / (17.0 + 3 * c) + c) * np.cos(yy / (23.0 - 2 * c)) for c in range(channels)]
        img = np.stack(base, -1) + rng.normal(0, 3.0, (h, w, channels))
    elif kind == "flat":
        img = np.zeros((h, w, channels))
        for _ in range(12):
            x0, y0 = int(rng.integers(0, max(w - 1, 1))), int(rng.integers(0, max(h - 1, 1)))
            x1, y1 = int(rng.integers(x0, w)) + 1, int(rng.integers(y0, h)) + 1
            img[y0:y1, x0:x1, :] = rng.integers(0, 256, channels)
        img[::7, :, :] = 255
    else:
        img = rng.integers(0, 256, (h, w, channels)).astype(np.float64)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def chunk(tag, data):
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def frame_png(width, height, bit_depth, color_type, zstream, extra=b""):
    ihdr = struct.pack(">IIBBBBB", width, height, bit_depth, color_type, 0, 0, 0)
    return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", ihdr) + extra + chunk(b"IDAT", zstream) + chunk(b"IEND", b"")


def pil_png(arr, **kw):
    """PNG bytes of an ndarray ([h,w] / [h,w,1] grey, [h,w,2] grey+alpha, [h,w,3], [h,w,4]) or of a PIL image."""
    from PIL import Image
    if hasattr(arr, "save"):
        im = arr
    elif arr.ndim == 2 or arr.shape[2] == 1:
        im = Image.fromarray(arr.reshape(arr.shape[0], arr.shape[1]))
    elif arr.shape[2] == 2:
        im = Image.merge("LA", [Image.fromarray(np.ascontiguousarray(arr[:, :, c])) for c in range(2)])
    else:
        im = Image.fromarray(arr)
    b = io.BytesIO()
    im.save(b, format="PNG", **kw)
    return b.getvalue()


def pil_pixels(data):
    from PIL import Image
    im = Image.open(io.BytesIO(data))
    im.load()
    return im


def idat_stream(png):
    """Concatenated IDAT payload and the IHDR fields of a PNG file."""
    pos, idat, ihdr, order = 8, b"", None, []
    while pos < len(png):
        n, tag = struct.unpack(">I4s", png[pos:pos + 8])
        body = png[pos + 8:pos + 8 + n]
        assert zlib.crc32(tag + body) & 0xFFFFFFFF == struct.unpack(">I", png[pos + 8 + n:pos + 12 + n])[0], "chunk CRC"
        order.append(tag)
        if tag == b"IHDR":
            ihdr = struct.unpack(">IIBBBBB", body)
        elif tag == b"IDAT":
            idat += body
        pos += 12 + n
    return ihdr, idat, order
