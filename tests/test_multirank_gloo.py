"""N > 1 path on CPU: two gloo ranks shard an image list, each runs the (host-only) lossless JPEG transcode on its shard
through the C-ABI, the quantisation-table handshake goes over the process group, and the gathered results equal the
single-process run in input order.  bench.py takes its per-rank shard of the synthetic set and does the quantisation-table
handshake through the same module (caesium-clt_b200/sharding.py) under torchrun with NCCL; inside one process the library shards
megabatches over its devices itself (api.cpp, LPT by bytes)."""
import hashlib
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INPUTS = ["in_420_base_355x237.jpg", "in_420_prog_355x237.jpg", "in_444_base_355x237.jpg", "in_422_base_355x237.jpg",
          "in_gray_base_355x237.jpg", "in_420_base_640x480.jpg", "in_420_tiny_17x9.jpg", "in_420_tiny_3x3.jpg"]


def _pkg():
    import importlib.util
    sys.path.insert(0, ROOT)
    if "caesium_clt_b200" not in sys.modules:
        pkg_dir = os.path.join(ROOT, "caesium-clt_b200")
        spec = importlib.util.spec_from_file_location("caesium_clt_b200", os.path.join(pkg_dir, "__init__.py"), submodule_search_locations=[pkg_dir])
        mod = importlib.util.module_from_spec(spec)
        sys.modules["caesium_clt_b200"] = mod
        spec.loader.exec_module(mod)
    import caesium_clt_b200._lib as L
    import caesium_clt_b200.sharding as S
    return L, S


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L, S = _pkg()
    files = [open(os.path.join(ROOT, "tests", "golden", n), "rb").read() for n in INPUTS] * 2
    # quant-table handshake: rank 0's table wins; every rank checks it equals its own computation
    mine = L.jpeg_quant_table(80, 0)
    got = S.broadcast_quant_table(mine if rank == 0 else mine * 0, dist, src=0)
    assert (got == mine).all()
    idx = S.shard_indices([len(f) for f in files], world, rank, "lpt")
    p = L.default_params()
    p.jpeg_optimize = 1
    outs = [hashlib.sha256(L.compress_in_memory(files[i], p)).hexdigest() for i in idx]
    gathered = [None] * world
    dist.all_gather_object(gathered, (idx, outs))
    if rank == 0:
        merged = S.merge_in_input_order(len(files), [g[0] for g in gathered], [g[1] for g in gathered])
        q.put(merged)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_and_merge():
    L, S = _pkg()
    files = [open(os.path.join(ROOT, "tests", "golden", n), "rb").read() for n in INPUTS] * 2
    p = L.default_params()
    p.jpeg_optimize = 1
    expect = [hashlib.sha256(L.compress_in_memory(f, p)).hexdigest() for f in files]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    merged = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert merged == expect


def test_shard_policies():
    _, S = _pkg()
    sizes = [10, 200, 30, 400, 50, 60, 700, 80, 90]
    for world in (1, 2, 3, 4, 8):
        for pol in ("lpt", "rr"):
            shards = [S.shard_indices(sizes, world, r, pol) for r in range(world)]
            flat = sorted(i for s in shards for i in s)
            assert flat == list(range(len(sizes)))                       # disjoint cover
            assert all(s == sorted(s) for s in shards)                   # input order inside a shard
    lpt = [sum(sizes[i] for i in S.shard_indices(sizes, 2, r, "lpt")) for r in range(2)]
    assert abs(lpt[0] - lpt[1]) <= max(sizes)                            # balanced by bytes
    assert S.merge_in_input_order(4, [[0, 2], [1, 3]], [["a", "c"], ["b", "d"]]) == ["a", "b", "c", "d"]
