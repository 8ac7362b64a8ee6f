"""CPU restatement of the reference's input pre-processing.

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.

Restates, in numpy:
  * ``utils/utils.py:20-34`` ``resize_image`` (letterbox: ``Image.resize((nw, nh), Image.BICUBIC)`` pasted on a
    (128,128,128) canvas) and ``:44-48`` ``preprocess_input`` (/255, -mean, /std with numpy's in-place dtype rules),
    called at ``achelous.py:200-205``;
  * ``utils/utils.py:51-54`` ``preprocess_input_radar`` (min-max + 1e-13), ``achelous.py:212``;
  * ``achelous.py:224-246`` point sampling by index + ``sklearn.preprocessing.normalize(X, axis=0)`` + permute.

The bicubic resize lives in a third-party dependency that is not under /root/reference: Pillow
(``requirements.txt`` pins Pillow==8.2.0; 12.2.0 is installed in this image).  Its published algorithm
(``src/libImaging/Resample.c``: ``precompute_coeffs``, ``normalize_coeffs_8bpc``,
``ImagingResampleHorizontal_8bpc`` / ``Vertical_8bpc``, ``ImagingResample``) is restated here:
separable two-pass convolution, horizontal first, 8-bit intermediate; per output sample the filter support is
``2 * max(1, in/out)`` (antialiasing when shrinking), the double-precision cubic (a = -0.5) weights are normalised to
sum 1, converted to 22-bit fixed point with round-half-away, accumulated in int32 from ``1 << 21`` and shifted back
with saturation to [0, 255].  Pinned: ``tests/test_preprocess.py`` checks this restatement bit for bit against
Pillow itself (when importable) and against the committed ``tests/golden/preprocess.npz`` produced by
``tests/golden/make_preprocess_golden.py`` with Pillow / numpy / sklearn.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x):
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_bicubic_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for box (0, in_size): returns
    (ksize, bounds int32 [out][2] = (xmin, count), kk int32 [out][ksize])."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _clip8(acc):
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def _resample_axis0(img, out_size):
    """One 8bpc pass along axis 0 of a (N, ...) uint8 array."""
    _, bounds, kk = pil_bicubic_coeffs(img.shape[0], out_size)
    out = np.empty((out_size,) + img.shape[1:], np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_size):
        xmin, cnt = bounds[xx]
        k = kk[xx, :cnt].astype(np.int64).reshape((cnt,) + (1,) * (img.ndim - 1))
        acc = (src[xmin:xmin + cnt] * k).sum(0) + (1 << (PRECISION_BITS - 1))
        out[xx] = _clip8(acc)
    return out


def pil_resize_bicubic(img, size):
    """img (H, W, 3) uint8, size = (nw, nh) as PIL takes it -> (nh, nw, 3) uint8 (Image.resize(size, Image.BICUBIC))."""
    nw, nh = size
    ih, iw = img.shape[:2]
    if (nw, nh) == (iw, ih):
        return img.copy()
    out = img
    if nw != iw:  # horizontal pass first (only the rows the vertical pass reads; row-wise independent, so all rows here)
        out = np.ascontiguousarray(_resample_axis0(np.ascontiguousarray(out.transpose(1, 0, 2)), nw).transpose(1, 0, 2))
    if nh != ih:
        out = _resample_axis0(out, nh)
    return out


def letterbox_geometry(image_hw, input_hw):
    """(nw, nh, x_off, y_off) of utils.py:20-31 for an (ih, iw) image and an (h, w) network input."""
    ih, iw = image_hw
    h, w = input_hw
    scale = min(w / iw, h / ih)
    nw, nh = int(iw * scale), int(ih * scale)
    return nw, nh, (w - nw) // 2, (h - nh) // 2


def resize_image(img, input_hw, letterbox_image=True):
    """utils.py:20-34 on a (ih, iw, 3) uint8 array -> (h, w, 3) uint8."""
    h, w = input_hw
    if not letterbox_image:
        return pil_resize_bicubic(img, (w, h))
    nw, nh, xo, yo = letterbox_geometry(img.shape[:2], input_hw)
    canvas = np.full((h, w, 3), 128, np.uint8)
    canvas[yo:yo + nh, xo:xo + nw] = pil_resize_bicubic(img, (nw, nh))
    return canvas


def preprocess_input(image_u8):
    """utils.py:44-48 applied to np.array(image, dtype='float32') (achelous.py:205): float32 /255 (weak Python scalar),
    then in-place -= / /= with float64 arrays -> each step computed in float64 and rounded to float32."""
    image = np.array(image_u8, dtype="float32")
    image /= 255.0
    image -= np.array([0.485, 0.456, 0.406])
    image /= np.array([0.229, 0.224, 0.225])
    return image


def image_to_input(img, input_hw, letterbox_image=True):
    """achelous.py:200-205: (ih, iw, 3) uint8 -> (1, 3, h, w) float32."""
    x = preprocess_input(resize_image(img, input_hw, letterbox_image))
    return np.expand_dims(np.transpose(x, (2, 0, 1)), 0)


def preprocess_input_radar(data):
    """utils.py:51-54; the result is then cast to float32 (achelous.py:212 `.type(torch.FloatTensor)`)."""
    _range = np.max(data) - np.min(data)
    data = (data - np.min(data)) / _range + 0.0000000000001
    return data.astype(np.float32)


def sample_points(features, indexes):
    """achelous.py:224-246: rows `indexes` of the (n, C) float64 feature table, every column divided by its L2 norm over
    the sampled rows (sklearn normalize(axis=0): zero norms are replaced by 1), float32, (1, C, N)."""
    x = np.asarray(features, dtype=np.float64)[indexes]
    norms = np.sqrt((x * x).sum(0))
    norms[norms == 0.0] = 1.0
    x = x / norms
    return np.ascontiguousarray(x.astype(np.float32).T)[None]
