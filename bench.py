#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: megapixels/s, JPEG q=80 4:2:0 re-encode of 3840x2160 inputs.

One "step" = one pass of the hot path over one megabatch of synthetic 4K JPEGs (BASELINE configs[1]).
  value : device pipeline (dequant+IDCT -> chroma resample -> FDCT+quant+zigzag) with the batch's coefficients
          already resident in HBM, timed with CUDA events on the launching stream, max over ranks.
  e2e   : the same metric through the reference-facing C-ABI (b200_compress_batch: JPEG bytes in host memory ->
          JPEG bytes in host memory: host Huffman decode, pinned H2D, kernels, D2H, host Huffman encode).
  roofline : the dominant kernel (fused luma IDCT->FDCT) against MEASURED_PEAKS.json's HBM copy bandwidth.
  cpu_baseline / --impl reference : the CPU restatement of the reference path (oracle/, "port": the Rust reference
          cannot be built in this image) on the box's usable host cores.
Launch: python bench.py [--gpus N --steps K --warmup W] or torchrun --nproc-per-node N bench.py --gpus N ...
"""
import argparse
import importlib.util
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# the library drives one CUDA stream per in-flight image; give them separate hardware queues (must precede CUDA init,
# and torch may create the context before the library does)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("NCCL_DEBUG", "WARN")      # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line

W4K, H4K = 3840, 2160
MP_PER_IMAGE = W4K * H4K / 1e6
QUALITY, SUBSAMPLING = 80, 420


def usable_cores():
    n = os.cpu_count() or 1
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = max(1, min(n, -(-int(q) // int(p))))
    except Exception:
        pass
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    return n


def _gen_one(idx):
    from tools.synth import synth_jpeg
    return synth_jpeg(W4K, H4K, idx)


def make_inputs(n_unique, first_index):
    """n_unique seeded 4K source JPEGs (q=90, 4:2:0, baseline, Annex-K tables via Pillow/libjpeg-turbo)."""
    import multiprocessing as mp
    procs = min(n_unique, usable_cores())
    if procs <= 1:
        return [_gen_one(first_index + i) for i in range(n_unique)]
    with mp.get_context("fork").Pool(procs) as pool:
        return pool.map(_gen_one, [first_index + i for i in range(n_unique)])


def load_pkg():
    pkg_dir = os.path.join(ROOT, "caesium-clt_b200")
    if "caesium_clt_b200" not in sys.modules:
        spec = importlib.util.spec_from_file_location("caesium_clt_b200", os.path.join(pkg_dir, "__init__.py"), submodule_search_locations=[pkg_dir])
        mod = importlib.util.module_from_spec(spec)
        sys.modules["caesium_clt_b200"] = mod
        spec.loader.exec_module(mod)
    import caesium_clt_b200._lib as lib
    lib.lib()
    return lib


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed regions run."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_rate(datas, cores, n_images):
    """Oracle (CPU restatement of libcaesium jpeg::lossy) on `cores` threads over n_images inputs -> MP/s."""
    from oracle import oracle as O
    O.lib()
    p = O.params(QUALITY, SUBSAMPLING, True)
    work = [datas[i % len(datas)] for i in range(n_images)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        outs = list(ex.map(lambda d: len(O.jpeg_lossy(d, p)), work))
    dt = time.perf_counter() - t0
    return n_images * MP_PER_IMAGE / dt, dt, sum(outs)


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = usable_cores()
    datas = make_inputs(min(8, 2 * cores), 0)
    n = 2 * cores
    for _ in range(args.warmup):
        cpu_reference_rate(datas, cores, min(n, cores))
    t_total, imgs = 0.0, 0
    for _ in range(args.steps):
        _, dt, _ = cpu_reference_rate(datas, cores, n)
        t_total += dt; imgs += n
    v = imgs * MP_PER_IMAGE / t_total
    sample = f"{n} images/step of the 3840x2160 q90 4:2:0 synthetic set, {cores} threads, oracle jpeg_lossy (progressive, optimised Huffman; no trellis / scan search)"
    _emit(({
        "impl": "reference", "metric": "megapixels/sec JPEG q=80 4K re-encode", "value": round(v, 2), "unit": "MP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * t_total / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i32", "data": "synthetic",
        "config": {"workload": "configs[1]: 3840x2160 RGB JPEG q90 4:2:0 -> -q 80 --jpeg-chroma-subsampling 4:2:0", "images_per_step": n, "l2": "n/a (CPU)"},
        "images_per_sec": round(v / MP_PER_IMAGE, 2),
        "cpu_baseline": {"value": round(v, 2), "unit": "MP/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(v, 2), "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


_RESULT_OUT = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too at the C level (NCCL's version banner under
    torchrun), so file descriptor 1 is pointed at stderr for the whole run and the result line goes to a private duplicate
    of the original stdout."""
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def _emit(obj):
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="images per megabatch per GPU (device-resident leg)")
    ap.add_argument("--e2e-batch", type=int, default=256, help="images per step per GPU (C-ABI leg)")
    ap.add_argument("--unique", type=int, default=8, help="unique synthetic sources per rank, cycled to fill a batch")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank, world, local_rank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    # ---- inputs first (fork pool must run before CUDA is initialised in this process)
    datas = make_inputs(args.unique, rank * args.unique)
    cores = usable_cores()
    threads = max(1, cores // max(1, world))
    # batch workers mostly wait for their stream (stream_wait.h: brief poll, then sleeps): on a box with few cores per GPU a
    # rank still keeps eight megabatches in flight
    e2e_threads = max(threads, 8)

    import torch
    import torch.distributed as dist
    L = load_pkg()
    torch.cuda.set_device(local_rank)
    if L.lib().b200_init_device(local_rank) != 0:
        raise SystemExit("bench.py: no B200 visible -- the product has no CPU fallback")
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # quant tables: computed on rank 0, broadcast over NCCL (the only collective of this path), checked locally
    qt = torch.tensor(L.jpeg_quant_table(QUALITY, 0).astype(np.int32), device="cuda")
    if world > 1:
        dist.broadcast(qt, 0)
    assert np.array_equal(qt.cpu().numpy().astype(np.uint16), L.jpeg_quant_table(QUALITY, 0)), "quant-table handshake failed"

    p = L.default_params()
    p.jpeg_quality, p.jpeg_chroma_subsampling, p.jpeg_progressive = QUALITY, SUBSAMPLING, 1

    # ---- device-resident megabatch
    decoded = [L.jpeg_decode_coefficients(d) for d in datas]
    lay = decoded[0][0]
    olay = L.jpeg_output_layout(lay, p)
    B = args.batch
    batch = L.JpegBatch(lay, olay, B)
    for i in range(B):
        batch.upload(i, decoded[i % len(decoded)][1])
    stream = torch.cuda.Stream()            # a real (non-NULL) stream: the kernels are launched on it and the events recorded on it
    sh = stream.cuda_stream
    assert sh != 0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local_rank)
    clocks.start()
    launches_per_step = 0
    for _ in range(args.warmup):
        launches_per_step = batch.run(sh)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        batch.run(sh)
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * B * MP_PER_IMAGE * args.steps / (ms_total / 1e3)

    # ---- per-kernel timing for the roofline (CUDA events on the library's own launching stream)
    kern = {}
    for which, name in ((1, "k_fused_same(luma IDCT->FDCT+quant)"), (2, "k_idct_plane(chroma)"), (3, "k_chroma420_refdct")):
        batch.time(which, 2)
        kern[name] = batch.time(which, max(5, args.steps))
    n_px = B * W4K * H4K
    alg_bytes = {  # algorithmic bytes per launch: int16 coefficients in/out, u8 planes
        "k_fused_same(luma IDCT->FDCT+quant)": n_px * 4.0,
        "k_idct_plane(chroma)": n_px * (1.0 + 0.5),
        "k_chroma420_refdct": n_px * (0.5 + 1.0),
    }
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    dom = "k_fused_same(luma IDCT->FDCT+quant)"
    achieved = alg_bytes[dom] / (kern[dom] / 1e3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": None, "peak_source": peak_src, "ms_per_launch": round(kern[dom], 4),
                "all_kernels": {k: {"ms": round(v, 4), "GBps": round(alg_bytes[k] / (v / 1e3) / 1e9, 1), "frac": round(alg_bytes[k] / (v / 1e3) / 1e9 / peak, 4)} for k, v in kern.items()}}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            tr = json.load(open(traffic_file))
            roofline["traffic"] = tr.get("k_fused_same_bytes_per_image", 0) * B or None
            # what actually bounds the kernel (ncu --set full, profiles/): exact ISLOW butterflies are integer multiply-adds, which
            # issue on the fmaheavy half of the FMA pipe only; reported next to the HBM fraction so the two are not confused
            if "k_fused_same_pipes" in tr:
                roofline["pipes"] = dict(tr["k_fused_same_pipes"], source=tr.get("source", "profiles/traffic.json"))
        except Exception:
            pass
    batch.close()

    # ---- end to end through the C-ABI with host buffers
    Be = args.e2e_batch
    work = [datas[i % len(datas)] for i in range(Be)]
    bi = L.BatchInputs(work)                                   # pointer/length arrays built once: the timed call is the C-ABI call
    L.compress_batch(work[:max(threads, 8)], p, e2e_threads, copy=False)      # warm slot pools / pinned buffers
    for _ in range(max(2, args.warmup)):
        L.compress_batch(bi, p, e2e_threads, copy=False)
    barrier()
    t0 = time.perf_counter()
    out_bytes = 0
    for _ in range(args.steps):
        # copy=False: outputs are read where the library malloc'ed them (length + SOI marker) and freed; duplicating
        # every file into a Python bytes object is ctypes overhead, not part of the C-ABI a host program calls
        res = L.compress_batch(bi, p, e2e_threads, copy=False)
        assert all(r[1] == 0 and r[3] == b"\xff\xd8" for r in res), [r[2] for r in res if r[1]][:1]
        out_bytes = sum(r[0] for r in res)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    e2e_val = world * Be * MP_PER_IMAGE * args.steps / dt
    clk = clocks.stop()
    coef_bytes = int(lay.total_coefs) * 2
    ent = os.environ.get("B200_ENTROPY", "gpu")
    in_bytes = sum(len(w) for w in work)
    # what actually crosses PCIe per step: with the device entropy decoder the entropy-coded scan goes up (not the
    # coefficients), with the device encoder the stuffed scans come back (not the coefficients)
    h2d = in_bytes if ent in ("gpu", "gpudec") else Be * coef_bytes
    d2h = out_bytes if ent in ("gpu", "gpuenc") else Be * int(olay.total_coefs) * 2
    e2e = {"value": round(e2e_val, 2), "unit": "MP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "images_per_sec": round(e2e_val / MP_PER_IMAGE, 2), "host_threads": e2e_threads, "host_cores": threads, "in_bytes_per_step": in_bytes, "out_bytes_per_step": out_bytes,
           "entropy": ent, "megabatch": int(os.environ.get("B200_MEGABATCH", "8")), "group_workers": min(e2e_threads, int(os.environ.get("B200_GROUP_WORKERS", "16"))),
           "note": "JPEG bytes in host memory -> JPEG bytes in host memory via b200_compress_batch (the batch form of start_compression's par_iter), all inside the timed region: marker parsing, pinned H2D of the entropy-coded scans, device Huffman decode, transform kernels, device Huffman encode (statistics, optimal tables, bit packing, stuffing), D2H of the scans, file assembly. Output bytes are identical to the oracle's. B200_ENTROPY=host keeps both entropy stages on host threads."}

    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        n = 2 * cores
        cpu_reference_rate(datas, cores, cores)
        v, cdt, _ = cpu_reference_rate(datas, cores, n)
        cpu = {"value": round(v, 2), "unit": "MP/s", "cores": cores, "kind": "port",
               "sample": f"{n} of the same 4K inputs, {cores} threads, oracle jpeg_lossy (restated reference: progressive + optimised Huffman, no trellis/scan search), {cdt:.1f} s"}

    if rank == 0:
        _emit(({
            "metric": "megapixels/sec JPEG q=80 4K re-encode", "value": round(value, 1), "unit": "MP/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_total / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "i32", "data": "synthetic",
            "config": {"workload": "configs[1]: 3840x2160 RGB JPEG q90 4:2:0 -> -q 80 --jpeg-chroma-subsampling 4:2:0 (progressive, optimised Huffman)",
                       "value_scope": "north_star's named transform kernels (K1-K5) over HBM-resident coefficients; the device Huffman decode / encode passes run inside e2e, which is GPU-bound (profiles/r1c_group_kernels.txt)",
                       "images_per_step_per_gpu": B, "unique_sources_per_gpu": len(datas), "parallelism": f"dp{world} (images sharded, no collective on the path)",
                       "l2": f"inputs larger than L2 ({B * coef_bytes / 1e6:.0f} MB of coefficients per step per GPU vs 126 MB)"},
            "images_per_sec": round(value / MP_PER_IMAGE, 1),
            "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "roofline": roofline, "cpu_baseline": cpu, "clocks": clk,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
