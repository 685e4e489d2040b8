"""Throughput of the BASELINE configs that are not the bench line (GPU box only), next to the oracle on host cores:
  configs[2]  3840x2160 JPEGs, --lossless re-encode          (device entropy decode -> device entropy encode)
  configs[3]  RGBA PNGs, --lossless --png-opt-level 3         (K6 filter selection + K7 LZ77; 2048x2048 here to bound time)
  configs[4]  6000x4000 JPEGs, -q 85 --width 1920 --format webp
Prints one JSON line per config; copy into profiles/.  Batch calls are timed with their outputs left in the library's buffers
(copy=False: what a C / Rust host sees, see tools/throughput.py); a separate small call fetches bytes for the oracle check.
usage: python tools/other_configs.py [2] [3] [4]"""
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from pngutil import pil_png, synth  # noqa: E402

import numpy as np  # noqa: E402


def timed(fn, reps=1):
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    return (time.perf_counter() - t0) / reps, r


def config2(L, O, cores):
    # ---- configs[2]: lossless JPEG transcode
    datas = bench.make_inputs(8, 0)
    n = 512
    work = [datas[i % len(datas)] for i in range(n)]
    p = L.default_params(); p.jpeg_optimize = 1; p.jpeg_progressive = 1
    bi = L.BatchInputs(work)
    L.compress_batch(work[:128], p, 16, copy=False)
    L.compress_batch(bi, p, 16, copy=False)
    dt, res = timed(lambda: L.compress_batch(bi, p, 16, copy=False))
    assert all(r[1] == 0 for r in res)
    ref = O.jpeg_lossless(datas[0], O.params(80, 0, True))
    assert L.compress_batch(work[:8], p, 8)[0][0] == ref, "device lossless transcode differs from the oracle"
    m = min(len(datas), 8) * 4
    sample = [datas[i % len(datas)] for i in range(m)]
    with ThreadPoolExecutor(cores) as ex:
        cdt, _ = timed(lambda: list(ex.map(lambda d: O.jpeg_lossless(d, O.params(80, 0, True)), sample)))
    print(json.dumps({"config": "configs[2] 3840x2160 JPEG --lossless", "images": n, "images_per_s": round(n / dt, 1), "mp_per_s": round(n * bench.MP_PER_IMAGE / dt, 1),
                      "cpu_oracle_mp_per_s": round(m * bench.MP_PER_IMAGE / cdt, 1), "cpu_cores": cores, "bytes_identical_to_oracle": True}), flush=True)


def config3(L, O, cores):
    # ---- configs[3]: lossless PNG, level 3
    w = h = 2048
    imgs = []
    for s in range(4):
        rgb = synth(h, w, 3, seed=s, kind="photo" if s % 2 == 0 else "flat")
        alpha = synth(h, w, 1, seed=100 + s, kind="flat")
        imgs.append(pil_png(np.concatenate([rgb, alpha], axis=2), compress_level=6))
    n = 32
    work = [imgs[i % len(imgs)] for i in range(n)]
    p = L.default_params(); p.png_optimize = 1; p.png_optimization_level = 3
    L.compress_batch(work[:16], p, 16, copy=False)
    bi = L.BatchInputs(work)
    dt, res = timed(lambda: L.compress_batch(bi, p, 16, copy=False))
    assert all(r[1] == 0 for r in res), [r[2] for r in res if r[1]]
    out_bytes = sum(r[0] for r in res[:4]); in_bytes = sum(len(d) for d in imgs)

    def oracle_png(d):
        info, raw = L.png_decode(d)          # host-side container parse of the product; the oracle restates filter + LZ77
        best = None
        for s in L.png_level_strategies(3):
            f = O.png_filter(raw, info.bpp, s)
            tok, hist = O.png_lz77(f.reshape(-1), info.bpp, f.shape[1])
            if best is None or tok.size < best:
                best = tok.size
        return best
    with ThreadPoolExecutor(cores) as ex:
        cdt, _ = timed(lambda: list(ex.map(oracle_png, imgs)))
    mp = w * h / 1e6
    print(json.dumps({"config": "configs[3] %dx%d RGBA PNG --lossless --png-opt-level 3" % (w, h), "images": n, "images_per_s": round(n / dt, 2), "mp_per_s": round(n * mp / dt, 1),
                      "cpu_oracle_mp_per_s": round(len(imgs) * mp / cdt, 1), "cpu_cores": cores, "out_over_in_bytes": round(out_bytes / in_bytes, 3)}), flush=True)

def config4(L, O, cores):
    # ---- configs[4]: 6000x4000 JPEG -> --width 1920 --format webp -q 85
    import io
    from PIL import Image
    big = []
    for s in range(4):
        b = io.BytesIO(); Image.fromarray(synth(4000, 6000, 3, seed=s, kind="photo")).save(b, format="JPEG", quality=90, subsampling=2)
        big.append(b.getvalue())
    n = 64
    work = [big[i % len(big)] for i in range(n)]
    p = L.default_params(); p.webp_quality = 85; p.width = 1920

    def conv(d):
        return L.convert_in_memory(d, p, 3)
    with ThreadPoolExecutor(16) as ex:
        list(ex.map(conv, work[:32]))
        dt, res = timed(lambda: list(ex.map(conv, work)))

    def oracle_conv(d):
        ycc = O.Jpeg(d).decode_native()
        rgb = O.ycc_to_rgb(ycc)
        nw, nh = O.compute_dimensions(6000, 4000, 1920, 0)
        rgb = np.stack([O.resize_plane(rgb[c], nw, nh) for c in range(3)])
        return O.webp_encode(rgb, 85)[0]
    with ThreadPoolExecutor(cores) as ex:
        cdt, cres = timed(lambda: list(ex.map(oracle_conv, big)))
    assert res[0] == cres[0], "device WebP conversion differs from the oracle"
    mp = 24.0
    print(json.dumps({"config": "configs[4] 6000x4000 JPEG -q 85 --width 1920 --format webp", "images": n, "images_per_s": round(n / dt, 1), "input_mp_per_s": round(n * mp / dt, 1),
                      "cpu_oracle_input_mp_per_s": round(len(big) * mp / cdt, 1), "cpu_cores": cores, "bytes_identical_to_oracle": True, "out_bytes": len(res[0])}), flush=True)


if __name__ == "__main__":
    L = bench.load_pkg()
    from oracle import oracle as O
    O.lib()
    L.lib().b200_init_device(0)
    cores = bench.usable_cores() if hasattr(bench, "usable_cores") else os.cpu_count()
    which = set(int(a) for a in sys.argv[1:]) or {2, 3, 4}
    if 2 in which:
        config2(L, O, cores)
    if 3 in which:
        config3(L, O, cores)
    if 4 in which:
        config4(L, O, cores)
    L.lib().b200_shutdown()
