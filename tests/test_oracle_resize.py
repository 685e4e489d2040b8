"""CPU tests pinning oracle/resize_oracle.c: libjpeg colour conversion against libjpeg-turbo (bit-exact), the Lanczos3
resampler against a float64 restatement of image-crate's algorithm (<= 1 LSB, almost always 0) plus invariants, and
libcaesium's compute_dimensions cases.  The true `image` crate cannot be run here (no Rust toolchain)."""
import io

import numpy as np
import pytest
from PIL import Image


def test_ycc_to_rgb_matches_libjpeg_turbo(O, golden):
    """jdcolor.c ycc_rgb_convert: Pillow decodes the same file once to YCbCr (draft) and once to RGB."""
    for name in ("in_444_base_355x237.jpg", "in_420_base_640x480.jpg"):
        data = golden(name)
        im = Image.open(io.BytesIO(data)); im.draft("YCbCr", im.size)
        ycc = np.asarray(im).transpose(2, 0, 1)
        rgb = np.asarray(Image.open(io.BytesIO(data)).convert("RGB")).transpose(2, 0, 1)
        assert np.array_equal(O.ycc_to_rgb(ycc), rgb)


def test_rgb_to_ycc_matches_libjpeg_turbo(O):
    """jccolor.c rgb_ycc_convert: encode RGB at 4:4:4 with all-ones tables, read the DC/AC back through the oracle IDCT."""
    from tools.synth import synth_rgb
    rgb = synth_rgb(64, 48, 21)
    b = io.BytesIO()
    Image.fromarray(rgb, "RGB").save(b, "JPEG", qtables=[[1] * 64] * 2, subsampling=0)
    turbo = O.Jpeg(b.getvalue())
    ycc = O.rgb_to_ycc(rgb.transpose(2, 0, 1))
    mine = O.forward(ycc, O.params(100, 444, False))          # quality 100 -> all-ones tables
    assert int(O.quant_table(100).max()) == 1
    for c in range(3):
        assert np.array_equal(turbo.coef(c), mine.coef(c))


def _lanczos_ref(plane, nw, nh):
    def axis(a, n_out, ax):
        a = np.moveaxis(a.astype(np.float64), ax, 0)
        n_in = a.shape[0]
        ratio = n_in / n_out
        sr = max(ratio, 1.0)
        out = np.zeros((n_out,) + a.shape[1:])
        for o in range(n_out):
            c = (o + 0.5) * ratio
            left = int(min(max(np.floor(c - 3 * sr), 0), n_in - 1))
            right = int(min(max(np.ceil(c + 3 * sr), left + 1), n_in))
            x = (np.arange(left, right) - (c - 0.5)) / sr
            w = np.where(np.abs(x) < 3, np.sinc(x) * np.sinc(x / 3), 0.0)
            w /= w.sum()
            out[o] = np.tensordot(w, a[left:right], axes=(0, 0))
        return np.moveaxis(out, 0, ax)
    t = axis(axis(plane, nh, 0), nw, 1)
    return np.clip(t, 0, 255)


@pytest.mark.parametrize("w,h,nw,nh", [(200, 120, 100, 60), (200, 120, 63, 37), (64, 64, 128, 96), (97, 31, 20, 31), (33, 77, 33, 20), (50, 50, 1, 1)])
def test_lanczos3_against_float64_reference(O, w, h, nw, nh):
    from tools.synth import synth_rgb
    plane = synth_rgb(w, h, 31)[:, :, 1].copy()
    got = O.resize_plane(plane, nw, nh).astype(np.int32)
    ref = _lanczos_ref(plane, nw, nh)
    assert np.abs(got - ref).max() <= 0.5 + 2e-3          # f32 accumulation vs f64: rounding can only flip at the .5 boundary
    assert (got == np.floor(ref + 0.5).astype(np.int32)).mean() > 0.995


def test_lanczos3_invariants(O):
    flat = np.full((40, 60), 173, dtype=np.uint8)
    assert (O.resize_plane(flat, 17, 9) == 173).all()                 # weights are normalised
    rng = np.random.default_rng(5)
    p = rng.integers(0, 256, size=(32, 48)).astype(np.uint8)
    assert np.array_equal(O.resize_plane(p, 48, 32), p)               # same size: copy (imageops::resize short-circuit)
    # resizing commutes with transposition (separable, same kernel on both axes, per-axis windows)
    a = O.resize_plane(p, 20, 11)
    # vertical-then-horizontal is not bitwise symmetric under transpose (f32 intermediate of the first pass only), allow 1 LSB
    assert np.abs(a.astype(int) - O.resize_plane(p.T.copy(), 11, 20).T.astype(int)).max() <= 1


def test_compute_dimensions(O):
    """libcaesium resize.rs compute_dimensions: both given -> exact; one given -> the other follows the f32 ratio, rounded."""
    assert O.compute_dimensions(2000, 3000, 100, 100) == (100, 100)
    assert O.compute_dimensions(2000, 3000, 100, 0) == (100, 150)
    assert O.compute_dimensions(2000, 3000, 0, 100) == (67, 100)
    assert O.compute_dimensions(6000, 4000, 1920, 0) == (1920, 1280)          # BASELINE config 5
    assert O.compute_dimensions(355, 237, 0, 50) == (75, 50)
    assert O.compute_dimensions(3, 1000, 1, 0) == (1, 333)


def test_resized_lossy_decodes_with_independent_decoder(O, golden):
    data = golden("in_420_base_640x480.jpg")
    out = O.jpeg_lossy_resized(data, O.params(80, 420, True), 200, 0)
    im = Image.open(io.BytesIO(out)); im.load()
    assert im.size == (200, 150)
    # content sanity: close to Pillow's own Lanczos downscale of the source
    ref = np.asarray(Image.open(io.BytesIO(data)).convert("RGB").resize((200, 150), Image.LANCZOS)).astype(np.float64)
    psnr = 10 * np.log10(255 ** 2 / np.mean((np.asarray(im.convert("RGB")).astype(np.float64) - ref) ** 2))
    assert psnr > 25          # q80 quantisation of noisy synthetic content dominates the difference
    gray = O.jpeg_lossy_resized(golden("in_gray_base_355x237.jpg"), O.params(80, 0, False), 0, 100)
    assert Image.open(io.BytesIO(gray)).size == (150, 100)
