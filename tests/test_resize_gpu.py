"""GPU parity tests for the resize leg (K2 colour, K3 Lanczos3, K4 colour) and the CLI
mirror's lossy flows.  Bar: byte-identical files versus the oracle (integer colour math; f32 Lanczos with host-made
weights and no FMA contraction is bit-exact by construction)."""
import io
import json
import os
import subprocess

import numpy as np
import pytest
from PIL import Image

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "caesium-clt_b200")


def _params(L, q=80, ss=420, prog=True, w=0, h=0):
    p = L.default_params()
    p.jpeg_quality, p.jpeg_chroma_subsampling, p.jpeg_progressive, p.width, p.height = q, ss, int(prog), w, h
    return p


@pytest.mark.parametrize("name", ["in_420_base_355x237.jpg", "in_444_base_355x237.jpg", "in_422_base_355x237.jpg", "in_gray_base_355x237.jpg",
                                  "in_420_prog_355x237.jpg", "in_420_tiny_17x9.jpg"])
@pytest.mark.parametrize("w,h", [(100, 0), (0, 50), (120, 90), (500, 0), (355, 237)])
def test_resized_file_matches_oracle(L, O, golden, name, w, h):
    data = golden(name)
    if name.endswith("17x9.jpg"):
        w, h = (min(w, 40), min(h, 30))
        if w == 0 and h == 0:
            w = 8
    for q, ss, prog in [(80, 420, True), (60, 444, False)]:
        got = L.compress_in_memory(data, _params(L, q, ss, prog, w, h))
        ref = O.jpeg_lossy_resized(data, O.params(q, ss, prog), w, h)
        assert got == ref
        im = Image.open(io.BytesIO(got))
        ow, oh = Image.open(io.BytesIO(data)).size
        assert im.size == O.compute_dimensions(ow, oh, w, h)


def test_resize_baseline_config5_shape(L, O):
    """BASELINE config 5's resize leg at full size: 6000x4000 -> --width 1920 (JPEG output here; WebP encode is a later row)."""
    from tools.synth import synth_jpeg
    data = synth_jpeg(6000, 4000, 3)
    got = L.compress_in_memory(data, _params(L, 85, 420, True, 1920, 0))
    im = Image.open(io.BytesIO(got))
    assert im.size == (1920, 1280)
    assert got == O.jpeg_lossy_resized(data, O.params(85, 420, True), 1920, 0)


def _cli(*args):
    r = subprocess.run([os.path.join(PKG, "b200clt"), *args], capture_output=True, text=True)
    return r.returncode, r.stdout, r.stderr


def test_cli_lossy_flows(L, O, golden, tmp_path):
    """test_perform_compression (compressor.rs:769-896) through the GPU: q80 all Success with exact paths, q100 + Bigger
    policy all Skipped, --max-size, --long-edge with --no-upscale, output identical to the oracle."""
    src = tmp_path / "in"
    (src / "sub").mkdir(parents=True)
    (src / "a.jpg").write_bytes(golden("in_420_base_355x237.jpg"))
    (src / "sub" / "b.jpg").write_bytes(golden("in_420_base_640x480.jpg"))
    out = tmp_path / "out"
    rc, so, _ = _cli("-q", "80", "-o", str(out), "-R", "-S", "--json", "--jpeg-chroma-subsampling", "4:2:0", str(src))
    d = json.loads(so)
    assert rc == 0 and d["summary"]["success"] == 2, d
    assert (out / "a.jpg").read_bytes() == O.jpeg_lossy(golden("in_420_base_355x237.jpg"), O.params(80, 420, True))
    assert (out / "sub" / "b.jpg").read_bytes() == O.jpeg_lossy(golden("in_420_base_640x480.jpg"), O.params(80, 420, True))
    # q100 with policy Bigger: outputs are larger than the q80 files already there -> all skipped (compressor.rs:840-844)
    rc, so, _ = _cli("-q", "100", "-o", str(out), "-R", "-S", "--json", "-O", "bigger", str(src))
    assert json.loads(so)["summary"]["skipped"] == 2
    # --max-size
    rc, so, _ = _cli("--max-size", "20KB", "-o", str(tmp_path / "ms"), "-R", "--json", str(src))
    d = json.loads(so)
    assert d["summary"]["success"] == 2 and all(f["compressed_size"] <= 20000 for f in d["files"])
    # resize flags: long edge 200 (landscape -> width 200), and --no-upscale suppressing an enlarging request
    rc, so, _ = _cli("-q", "80", "--long-edge", "200", "-o", str(tmp_path / "le"), "--json", str(src / "sub" / "b.jpg"))
    assert Image.open(tmp_path / "le" / "b.jpg").size == (200, 150)
    rc, so, _ = _cli("-q", "80", "--width", "5000", "--no-upscale", "-o", str(tmp_path / "nu"), "--json", str(src / "sub" / "b.jpg"))
    assert Image.open(tmp_path / "nu" / "b.jpg").size == (640, 480)
    # baseline + 4:4:4 flags reach the codec
    rc, so, _ = _cli("-q", "70", "--jpeg-baseline", "--jpeg-chroma-subsampling", "4:4:4", "-o", str(tmp_path / "bl"), "--json", str(src / "a.jpg"))
    assert (tmp_path / "bl" / "a.jpg").read_bytes() == O.jpeg_lossy(golden("in_420_base_355x237.jpg"), O.params(70, 444, False))


def test_cli_batches_same_shaped_files(L, O, tmp_path):
    """start_compression hands the codec calls to b200_compress_batch: a folder of same-shaped baseline JPEGs goes through the
    megabatch path (device Huffman decode -> transform -> device Huffman encode, several images per launch), one file of another
    shape and sampling rides along; every output must still be the oracle's file."""
    import io
    from tools.synth import synth_rgb
    src = tmp_path / "in"; src.mkdir()
    datas = {}
    for i in range(9):
        b = io.BytesIO(); Image.fromarray(synth_rgb(322, 199, 300 + i), "RGB").save(b, "JPEG", quality=72 + 2 * i, subsampling="4:2:0")
        datas[f"s{i}.jpg"] = b.getvalue()
    b = io.BytesIO(); Image.fromarray(synth_rgb(200, 120, 7), "RGB").save(b, "JPEG", quality=85, subsampling="4:4:4")
    datas["odd.jpg"] = b.getvalue()
    for name, d in datas.items():
        (src / name).write_bytes(d)
    out = tmp_path / "out"
    rc, so, _ = _cli("-q", "80", "-o", str(out), "--json", "--jpeg-chroma-subsampling", "4:2:0", str(src))
    d = json.loads(so)
    assert rc == 0 and d["summary"]["success"] == len(datas), d
    for name, data in datas.items():
        assert (out / name).read_bytes() == O.jpeg_lossy(data, O.params(80, 420, True)), name
