"""Per-image C-ABI calls (the literal drop-in shape: one blocking b200_compress_in_memory per rayon worker, compressor.rs:81-83,
:305) from N caller threads, with and without B200_COALESCE, against one b200_compress_batch over the same images.
usage: python tools/coalesce_probe.py [n_images] [callers,callers,...]   (set B200_COALESCE in the environment)"""
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    callers = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "16,32,64").split(",")]
    datas = bench.make_inputs(32, 0)
    work = [datas[i % len(datas)] for i in range(n)]
    L = bench.load_pkg()
    L.lib().b200_init_device(0)
    p = L.default_params(); p.jpeg_quality, p.jpeg_chroma_subsampling, p.jpeg_progressive = 80, 420, 1
    import ctypes as C

    def one(d):
        outp, outl = C.c_void_p(), C.c_size_t()
        st = L.lib().b200_compress_in_memory(d, C.c_size_t(len(d)), C.byref(p), C.byref(outp), C.byref(outl))
        assert st.code == 0
        L.lib().b200_free(outp)
        return outl.value
    L.compress_batch(work[:256], p, 16, copy=False)
    rec = {"tool": "coalesce_probe", "coalesce": os.environ.get("B200_COALESCE", "0"), "images": n}
    for c in callers:
        with ThreadPoolExecutor(c) as ex:
            list(ex.map(one, work[:4 * c]))
            t0 = time.perf_counter(); list(ex.map(one, work)); dt = time.perf_counter() - t0
        rec[f"per_image_calls_{c}_threads_img_s"] = round(n / dt, 1)
    bi = L.BatchInputs(work)
    L.compress_batch(bi, p, 16, copy=False)
    t0 = time.perf_counter(); L.compress_batch(bi, p, 16, copy=False); dt = time.perf_counter() - t0
    rec["batch_call_img_s"] = round(n / dt, 1)
    print(json.dumps(rec), flush=True)
    L.lib().b200_shutdown()
