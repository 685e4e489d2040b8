"""One small call through every device leg, for `compute-sanitizer --tool memcheck|racecheck python tools/sanitize_all.py`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
from conftest import _import_pkg  # noqa: E402

_import_pkg()
import numpy as np  # noqa: E402

import caesium_clt_b200._lib as L  # noqa: E402
from pngutil import pil_png, synth  # noqa: E402

rgb = np.ascontiguousarray(synth(150, 210, 3, seed=1).transpose(2, 0, 1))
print("webp", len(L.webp_encode_rgb(rgb, 70)))
p = L.default_params(); p.png_optimize = 1; p.png_optimization_level = 3
print("png lossless", len(L.compress_in_memory(pil_png(synth(90, 130, 4, seed=2)), p)))
d = open(os.path.join(ROOT, "tests", "golden", "in_420_base_640x480.jpg"), "rb").read()
p = L.default_params(); p.jpeg_quality = 80
print("jpeg lossy", len(L.compress_in_memory(d, p)))
print("jpeg megabatch", [r[1] for r in L.compress_batch([d] * 8, p, 2)])
p.jpeg_optimize = 1
print("jpeg lossless", len(L.compress_in_memory(d, p)))
p = L.default_params(); p.jpeg_quality = 80; p.width = 200
print("jpeg resize", len(L.compress_in_memory(d, p)))
p.webp_quality = 80
print("jpeg -> webp", len(L.convert_in_memory(d, p, 3)))
rgba = synth(120, 170, 4, seed=3); rgba[:, :40, 3] = 255; rgba[:30, :, 3] = 0
p = L.default_params(); p.webp_quality = 75
out = L.convert_in_memory(pil_png(rgba), p, 3)
print("png with alpha -> webp (K7 over the alpha plane)", len(out), out[12:16])
print("webp with alpha -> webp", len(L.compress_in_memory(out, p)))
p = L.default_params(); p.png_optimize = 1; p.png_optimization_level = 2
print("png lossless, wider rows", len(L.compress_in_memory(pil_png(synth(70, 1100, 3, seed=4, kind="photo")), p)))
p = L.default_params(); p.jpeg_quality = 80
res = L.compress_to_size_in_memory(d, p, len(d) // 3, True) if hasattr(L, "compress_to_size_in_memory") else None
print("jpeg compress_to_size", None if res is None else len(res[0]) if isinstance(res, tuple) else len(res))
L.lib().b200_shutdown()
