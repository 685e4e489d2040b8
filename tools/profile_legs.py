"""Profiling helper: one call through the PNG leg (4096x4096 RGBA, --png-opt-level 3) and one through the resize -> WebP leg
(6000x4000 JPEG -> 1920 wide, q85), each run twice so that the second pass has warm buffers (run under ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "png,webp"
    png = bench.make_inputs(1, 0, "png4096")[0] if "png" in which else None
    jpg = bench.make_inputs(1, 0, "jpeg24mp")[0] if "webp" in which else None
    L = bench.load_pkg()
    L.lib().b200_init_device(0)
    for it in range(2):
        if png is not None:
            p = L.default_params(); p.png_optimize = 1; p.png_optimization_level = 3
            print("png", len(png), "->", len(L.compress_in_memory(png, p)), flush=True)
        if jpg is not None:
            p = L.default_params(); p.webp_quality = 85; p.width = 1920
            print("webp", len(jpg), "->", len(L.convert_in_memory(jpg, p, 3)), flush=True)
    L.lib().b200_shutdown()
