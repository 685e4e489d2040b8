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
L.lib().b200_shutdown()
