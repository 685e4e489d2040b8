"""Profiling helper: push one synthetic 4K JPEG through b200_compress_in_memory a few times (for ncu launch lists)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from tools.synth import synth_jpeg
L = bench.load_pkg()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
data = synth_jpeg(3840, 2160, 0)
p = L.default_params(); p.jpeg_quality, p.jpeg_chroma_subsampling, p.jpeg_progressive = 80, 420, 1
L.lib().b200_init_device(0)
for i in range(n):
    t = time.perf_counter(); out = L.compress_in_memory(data, p); dt = time.perf_counter() - t
    print(f"iter {i}: {dt*1e3:.2f} ms, {len(out)} bytes")
