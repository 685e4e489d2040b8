"""Profiling helper: two megabatches of K synthetic 4K JPEGs through b200_compress_batch on one worker thread (run under
ncu --metrics gpu__time_duration.sum to list every kernel of a group with its isolated duration)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

if __name__ == "__main__":
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    os.environ["B200_MEGABATCH"] = str(K)
    datas = bench.make_inputs(min(K, 8), 0)
    L = bench.load_pkg()
    L.lib().b200_init_device(0)
    p = L.default_params()
    p.jpeg_quality, p.jpeg_chroma_subsampling, p.jpeg_progressive = 80, 420, 1
    work = [datas[i % len(datas)] for i in range(K)]
    for it in range(2):
        t0 = time.perf_counter()
        res = L.compress_batch(work, p, 1)
        print(f"group {it}: {1e3 * (time.perf_counter() - t0):.2f} ms for {K} images", flush=True)
        assert all(r[1] == 0 for r in res)
