"""Time the lossless PNG path on one large synthetic image per optimisation level (GPU box only).
usage: python tools/png_probe.py [width height]"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from conftest import _import_pkg  # noqa: E402
from pngutil import pil_png, synth  # noqa: E402

_import_pkg()
import caesium_clt_b200._lib as L  # noqa: E402

w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3840, 2160)
for kind in ("photo", "flat"):
    img = synth(h, w, 3, seed=1, kind=kind)
    srcs = {lvl: pil_png(img, compress_level=lvl) for lvl in (1, 6, 9)}
    print(f"{kind} {w}x{h}: pillow/zlib sizes " + ", ".join(f"L{k}={len(v)}" for k, v in srcs.items()), flush=True)
    src = srcs[6]
    for level in (0, 1, 2, 3, 6):
        p = L.default_params(); p.png_optimize = 1; p.png_optimization_level = level
        L.compress_in_memory(src, p)
        t0 = time.perf_counter()
        n = 3
        for _ in range(n):
            out = L.compress_in_memory(src, p)
        dt = (time.perf_counter() - t0) / n
        print(f"  level {level}: {len(out)} bytes ({len(out) / len(src):.3f} of source), {dt * 1e3:.1f} ms/image, {w * h / dt / 1e6:.1f} MP/s", flush=True)
