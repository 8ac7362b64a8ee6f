"""Input pre-processing: oracle vs the reference's own functions (golden fixture, Pillow itself),
host coefficient tables vs the oracle (CPU), CUDA kernels vs oracle + golden through the C ABI (GPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import preprocess as OP

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "preprocess.npz")

# name -> (ih, iw, h, w, letterbox, seed)
IMAGE_CASES = {
    "wide_down": (120, 200, 96, 96, True, 1),          # shrink, bars top/bottom
    "tall_down": (300, 150, 96, 96, True, 2),          # shrink, bars left/right
    "same": (96, 96, 96, 96, True, 3),                 # no resampling at all
    "up": (50, 40, 96, 96, True, 4),                   # enlarge (support 2, no antialiasing)
    "stretch": (77, 133, 96, 64, False, 5),            # letterbox_image=False: anisotropic
    "hd_320": (360, 640, 320, 320, True, 6),           # the benchmark geometry (16:9 camera -> 320x320)
    "h_only": (96, 200, 96, 96, False, 7),             # horizontal pass only
    "v_only": (200, 96, 96, 96, False, 8),             # vertical pass only
}


def make_image(ih, iw, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:ih, 0:iw]
    base = (127 + 100 * np.sin(xx / 7.0 + seed) * np.cos(yy / 5.0))[..., None] + rng.integers(-60, 60, (ih, iw, 3))
    img = np.clip(base, 0, 255).astype(np.uint8)
    img[::17, ::13] = 255    # isolated extremes exercise the negative bicubic lobes / saturation
    img[5::19, 3::11] = 0
    return img


def make_radar(dtype, B=None):
    rng = np.random.default_rng(42)
    shape = (3, 80, 96) if B is None else (B, 3, 80, 96)
    x = np.zeros(shape, dtype)
    m = rng.random(shape) < 0.02
    x[m] = (rng.normal(size=m.sum()) * 30 + 5).astype(dtype)
    return x


def make_points(n=137, C=5, N=512):
    rng = np.random.default_rng(7)
    feats = rng.normal(size=(n, C)) * np.array([30.0, 5.0, 0.0, 2.0, 100.0])   # one all-zero column: zero norm -> 1
    idx = rng.integers(0, n, N)
    return feats, idx


# ------------------------------------------------------------------ CPU
def test_oracle_matches_reference_golden():
    g = np.load(GOLDEN)
    for name, (ih, iw, h, w, letterbox, seed) in IMAGE_CASES.items():
        mine = OP.image_to_input(make_image(ih, iw, seed), (h, w), letterbox)
        assert mine.dtype == np.float32 and np.array_equal(mine, g["image." + name]), name
    assert np.array_equal(OP.preprocess_input_radar(make_radar(np.float64)), g["radar.f64"])
    assert np.array_equal(OP.preprocess_input_radar(make_radar(np.float32)), g["radar.f32"])
    feats, idx = make_points()
    # sklearn sums the squares with einsum; order differs from numpy's, so allow one float32 ulp
    np.testing.assert_allclose(OP.sample_points(feats, idx), g["points"], rtol=1.2e-7, atol=0)


def test_oracle_resize_matches_pillow_bit_for_bit():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(0)
    for ih, iw, nw, nh in [(1080, 1920, 320, 180), (333, 517, 320, 206), (100, 120, 320, 266), (64, 64, 64, 17), (31, 200, 7, 200)]:
        img = rng.integers(0, 256, (ih, iw, 3), dtype=np.uint8)
        ref = np.array(Image.fromarray(img).resize((nw, nh), Image.BICUBIC))
        assert np.array_equal(OP.pil_resize_bicubic(img, (nw, nh)), ref), (ih, iw, nw, nh)


def test_host_tables_match_oracle():
    from achelous_b200.utils import preprocess as P
    for i, o in [(1920, 320), (1080, 180), (640, 320), (360, 180), (50, 96), (40, 76), (200, 96), (97, 96), (96, 97), (5, 64), (4000, 31)]:
        ks, b, k = P.bicubic_tables(i, o)
        ks2, b2, k2 = OP.pil_bicubic_coeffs(i, o)
        assert ks == ks2 and np.array_equal(b, b2) and np.array_equal(k, k2), (i, o)
    assert P.letterbox_geometry((360, 640), (320, 320)) == OP.letterbox_geometry((360, 640), (320, 320))
    assert P.letterbox_geometry((300, 150), (96, 96)) == OP.letterbox_geometry((300, 150), (96, 96))


def test_product_refuses_cpu_tensors():
    from achelous_b200.utils import preprocess as P
    with pytest.raises(RuntimeError):
        P.preprocess_image(torch.zeros(4, 4, 3, dtype=torch.uint8), (8, 8))
    with pytest.raises(RuntimeError):
        P.preprocess_radar(torch.zeros(1, 3, 4, 4))
    with pytest.raises(RuntimeError):
        P.preprocess_points(torch.zeros(4, 5, dtype=torch.float64), torch.zeros(8, dtype=torch.int64))


# ------------------------------------------------------------------ GPU (through the C ABI)
@pytest.mark.gpu
@pytest.mark.parametrize("name", list(IMAGE_CASES))
def test_image_kernels_bit_exact(name):
    from achelous_b200.utils.preprocess import preprocess_image
    ih, iw, h, w, letterbox, seed = IMAGE_CASES[name]
    img = make_image(ih, iw, seed)
    out = preprocess_image(torch.from_numpy(img).cuda(), (h, w), letterbox).cpu().numpy()
    assert np.array_equal(out, OP.image_to_input(img, (h, w), letterbox)), name          # vs oracle
    assert np.array_equal(out, np.load(GOLDEN)["image." + name]), name                    # vs the reference's own output


@pytest.mark.gpu
def test_image_batch_and_full_hd():
    from achelous_b200.utils.preprocess import preprocess_image
    rng = np.random.default_rng(3)
    frames = rng.integers(0, 256, (3, 1080, 1920, 3), dtype=np.uint8)     # camera frames, BASELINE geometry 320x320
    out = preprocess_image(torch.from_numpy(frames).cuda(), (320, 320)).cpu().numpy()
    for b in range(3):
        assert np.array_equal(out[b:b + 1], OP.image_to_input(frames[b], (320, 320)))
    # size-independent properties: bars are exactly the normalised grey 128, a constant image stays constant
    grey = OP.preprocess_input(np.full((1, 1, 3), 128, np.uint8))[0, 0]
    assert np.array_equal(out[0, :, 0, 0], grey) and np.array_equal(out[0, :, -1, -1], grey)
    const = preprocess_image(torch.full((720, 1280, 3), 77, dtype=torch.uint8).cuda(), (320, 320)).cpu().numpy()
    inner = const[0, :, 70:250, :]
    assert np.array_equal(inner, np.broadcast_to(OP.preprocess_input(np.full((1, 1, 3), 77, np.uint8))[0, 0][:, None, None], inner.shape))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_radar_kernel_bit_exact(dtype):
    from achelous_b200.utils.preprocess import preprocess_radar
    x = make_radar(dtype, B=3)
    x[1] *= 3.5
    x[2, 0, 0, 0] = -200.0
    out = preprocess_radar(torch.from_numpy(x).cuda()).cpu().numpy()
    for b in range(3):
        assert np.array_equal(out[b], OP.preprocess_input_radar(x[b]))
    g = np.load(GOLDEN)["radar.f64" if dtype == np.float64 else "radar.f32"]
    assert np.array_equal(preprocess_radar(torch.from_numpy(make_radar(dtype)[None]).cuda()).cpu().numpy()[0], g)


@pytest.mark.gpu
def test_points_kernel():
    from achelous_b200.utils.preprocess import preprocess_points
    feats, idx = make_points()
    idx2 = np.stack([idx, idx[::-1].copy()])
    out = preprocess_points(torch.from_numpy(feats).cuda(), torch.from_numpy(idx2).cuda()).cpu().numpy()
    # column norms are fp64 sums in a different (tree) order: identical after rounding to fp32 up to one ulp
    np.testing.assert_allclose(out[0:1], OP.sample_points(feats, idx), rtol=1.2e-7, atol=0)
    np.testing.assert_allclose(out[1:2], OP.sample_points(feats, idx[::-1]), rtol=1.2e-7, atol=0)
    np.testing.assert_allclose(out[0:1], np.load(GOLDEN)["points"], rtol=1.2e-7, atol=0)
    assert np.all(out[:, 2] == 0.0)     # the all-zero column stays zero (norm 0 -> 1)
    with pytest.raises(IndexError):
        preprocess_points(torch.from_numpy(feats).cuda(), torch.tensor([0, 137]).cuda())
