"""Regenerates tests/golden/: small seeded JPEG inputs (encoded with Pillow/libjpeg-turbo) and expected.json,
the oracle's outputs for them (sha256 of output bytes and of the quantised coefficients).  The reference itself
cannot be built here (no cargo/rustc), so these vectors pin the ORACLE; the oracle in turn is
pinned against libjpeg-turbo and the reference's fixture KATs in tests/test_oracle_jpeg.py.

Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import io
import json
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tools.synth import synth_rgb  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")

INPUTS = {
    "in_420_base_355x237.jpg": dict(size=(355, 237), idx=1, quality=90, subsampling="4:2:0"),
    "in_420_prog_355x237.jpg": dict(size=(355, 237), idx=1, quality=85, subsampling="4:2:0", progressive=True),
    "in_444_base_355x237.jpg": dict(size=(355, 237), idx=1, quality=92, subsampling="4:4:4"),
    "in_422_base_355x237.jpg": dict(size=(355, 237), idx=1, quality=75, subsampling="4:2:2"),
    "in_gray_base_355x237.jpg": dict(size=(355, 237), idx=1, quality=88, gray=True),
    "in_420_base_640x480.jpg": dict(size=(640, 480), idx=2, quality=90, subsampling="4:2:0"),
    "in_420_tiny_17x9.jpg": dict(size=(17, 9), idx=4, quality=90, subsampling="4:2:0"),
    "in_420_tiny_3x3.jpg": dict(size=(3, 3), idx=5, quality=90, subsampling="4:2:0"),
}
# (quality, subsampling, progressive)
CASES = [(80, 420, True), (80, 420, False), (80, 0, True), (80, 444, True), (80, 422, False), (80, 411, True),
         (50, 420, True), (95, 420, True), (100, 444, False), (5, 420, True), (1, 420, False)]


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


def main():
    for name, k in INPUTS.items():
        rgb = synth_rgb(k["size"][0], k["size"][1], k["idx"])
        b = io.BytesIO()
        if k.get("gray"):
            Image.fromarray(rgb[:, :, 1].copy(), "L").save(b, "JPEG", quality=k["quality"])
        else:
            Image.fromarray(rgb, "RGB").save(b, "JPEG", quality=k["quality"], subsampling=k["subsampling"], progressive=k.get("progressive", False))
        with open(os.path.join(G, name), "wb") as f:
            f.write(b.getvalue())
    exp = {}
    for name in INPUTS:
        data = open(os.path.join(G, name), "rb").read()
        e = {"input_sha256": sha(data), "lossy": {}, "lossless": {}}
        for q, ss, prog in CASES:
            out = O.jpeg_lossy(data, O.params(q, ss, prog))
            j = O.Jpeg(out)
            e["lossy"][f"q{q}_s{ss}_p{int(prog)}"] = {
                "size": len(out), "sha256": sha(out),
                "coef_sha256": [sha(np.ascontiguousarray(j.coef(c)).tobytes()) for c in range(j.ncomp)],
            }
        for prog in (True, False):
            out = O.jpeg_lossless(data, O.params(80, 0, prog))
            e["lossless"][f"p{int(prog)}"] = {"size": len(out), "sha256": sha(out)}
        exp[name] = e
    with open(os.path.join(G, "expected.json"), "w") as f:
        json.dump(exp, f, indent=1, sort_keys=True)
    print("wrote", len(exp), "inputs x", len(CASES), "cases")


if __name__ == "__main__":
    main()
