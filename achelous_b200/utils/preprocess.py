"""On-device input pre-processing: what the reference does per frame on the host with PIL,
numpy and sklearn (achelous.py:200-246, utils/utils.py:20-54) as four CUDA kernels behind the C ABI.

    from achelous_b200.utils.preprocess import preprocess_image, preprocess_radar, preprocess_points
    x       = preprocess_image(frames_u8, (320, 320))        # (B, ih, iw, 3) uint8 -> (B, 3, 320, 320) fp32, cuda
    x_radar = preprocess_radar(radar_maps)                   # (B, C, H, W) fp32/fp64 -> fp32, cuda
    x_pc    = preprocess_points(feature_table, indexes)      # (n, C) fp64, (B, N) int -> (B, C, N) fp32, cuda

Host side: letterbox geometry (utils.py:23-26) and Pillow's resampling coefficient tables (Resample.c
precompute_coeffs + normalize_coeffs_8bpc: per output sample the first source index, the tap count and 22-bit
fixed-point weights) - a few hundred numbers per image size, computed once per (in, out) pair in float64 exactly as
Pillow does and cached.  Device side: every pixel (integer convolution passes, letterbox, normalisation).  The host
copy shrinks from 3 x 4 bytes per network-input pixel (fp32 CHW) to the camera's own uint8 frame.
"""
import functools
import math

import numpy as np
import torch

from .. import _lib

PRECISION_BITS = 32 - 8 - 2   # Pillow: 8-bit pixels, 2 guard bits


def letterbox_geometry(image_hw, input_hw, letterbox_image=True):
    """(nw, nh, x_off, y_off): size of the resized image and where it is pasted (utils.py:20-34)."""
    ih, iw = int(image_hw[0]), int(image_hw[1])
    h, w = int(input_hw[0]), int(input_hw[1])
    if not letterbox_image:
        return w, h, 0, 0
    scale = min(w / iw, h / ih)
    nw, nh = int(iw * scale), int(ih * scale)
    return nw, nh, (w - nw) // 2, (h - nh) // 2


@functools.lru_cache(maxsize=64)
def bicubic_tables(in_size, out_size):
    """Pillow Resample.c coefficient tables for Image.BICUBIC over the full axis: (ksize, bounds int32 [out, 2],
    kk int32 [out, ksize]).  Vectorised over taps; the running sums follow Pillow's left-to-right order."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    inv = 1.0 / filterscale
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    a = -0.5
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        cnt = min(int(center + support + 0.5), in_size) - xmin
        x = np.abs((np.arange(cnt, dtype=np.float64) + xmin - center + 0.5) * inv)
        w = np.where(x < 1.0, ((a + 2.0) * x - (a + 3.0)) * x * x + 1, np.where(x < 2.0, (((x - 5) * x + 8) * x - 4) * a, 0.0))
        ww = 0.0
        for v in w.tolist():       # sequential double sum, as the C loop
            ww += v
        if ww != 0.0:
            w = w / ww
        fixed = w * float(1 << PRECISION_BITS)
        kk[xx, :cnt] = np.where(w < 0, np.trunc(-0.5 + fixed), np.trunc(0.5 + fixed)).astype(np.int64)
        bounds[xx] = (xmin, cnt)
    return ksize, bounds, kk


@functools.lru_cache(maxsize=64)
def _device_tables(in_size, out_size, device_index):
    ksize, bounds, kk = bicubic_tables(in_size, out_size)
    dev = torch.device("cuda", device_index)
    return ksize, bounds, torch.from_numpy(bounds).to(dev), torch.from_numpy(kk).to(dev)


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor (achelous_b200 has no CPU path); call .cuda() on the raw frame first")


def preprocess_image(frames, input_shape, letterbox_image=True):
    """frames (B, ih, iw, 3) or (ih, iw, 3) uint8 RGB on the GPU -> (B, 3, h, w) float32: resize_image + preprocess_input
    (utils.py:20-34,44-48) as called at achelous.py:200-205."""
    lib = _lib.load()
    _require_cuda(frames, "preprocess_image")
    if frames.dtype != torch.uint8:
        raise RuntimeError("preprocess_image: expected uint8 RGB frames")
    if frames.dim() == 3:
        frames = frames[None]
    if frames.dim() != 4 or frames.shape[-1] != 3:
        raise RuntimeError(f"preprocess_image: expected (B, ih, iw, 3), got {tuple(frames.shape)}")
    frames = frames.contiguous()
    B, ih, iw, _ = frames.shape
    h, w = int(input_shape[0]), int(input_shape[1])
    nw, nh, xo, yo = letterbox_geometry((ih, iw), (h, w), letterbox_image)
    dev = frames.device
    out = torch.empty(B, 3, h, w, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        need_h, need_v = nw != iw, nh != ih
        r0, rows = 0, ih
        vb = vk = None
        vks = 0
        if need_v:
            vks, vbounds_np, vb, vk = _device_tables(ih, nh, dev.index)
            # Pillow resamples only the source rows the vertical pass reads
            r0 = int(vbounds_np[0, 0])
            rows = int(vbounds_np[-1, 0] + vbounds_np[-1, 1]) - r0
        tmp, tmp_bs = frames, frames.stride(0)
        if need_h:
            hks, _, hb, hk = _device_tables(iw, nw, dev.index)
            tmp = torch.empty(B, rows, nw, 3, device=dev, dtype=torch.uint8)
            tmp_bs = tmp.stride(0)
            _lib.check(lib.ach_pre_resize_h(frames.data_ptr(), frames.stride(0), B, iw, r0, rows, nw, hb.data_ptr(), hk.data_ptr(), hks,
                                            tmp.data_ptr(), tmp_bs, st), "ach_pre_resize_h")
            if need_v and r0:
                vb = vb.clone()
                vb[:, 0] -= r0          # bounds relative to the first row of tmp
        _lib.check(lib.ach_pre_resize_v_norm(tmp.data_ptr(), tmp_bs, B, nw, nh, vb.data_ptr() if need_v else None,
                                             vk.data_ptr() if need_v else None, vks, 0 if need_v else 1, out.data_ptr(), out.stride(0),
                                             h, w, xo, yo, st), "ach_pre_resize_v_norm")
    return out


def preprocess_radar(data):
    """(B, ...) or one (...) radar feature map, float32 or float64, on the GPU -> float32, per-sample
    (x - min) / (max - min) + 1e-13 (utils.py:51-54 + the FloatTensor cast of achelous.py:212)."""
    lib = _lib.load()
    _require_cuda(data, "preprocess_radar")
    if data.dtype not in (torch.float32, torch.float64):
        raise RuntimeError("preprocess_radar: expected float32 or float64")
    x = data.contiguous()
    B = x.shape[0]
    n = x[0].numel()
    out = torch.empty(x.shape, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        st = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(lib.ach_pre_radar(x.data_ptr(), n, 1 if x.dtype == torch.float64 else 0, B, n, out.data_ptr(), n, st), "ach_pre_radar")
    return out


def preprocess_points(features, indexes):
    """features (n, C) float64 (the frame's radar point table), indexes (B, N) or (N,) integer sample indices (the
    reference draws them with np.random.choice(..., replace=True)) -> (B, C, N) float32 (achelous.py:224-246)."""
    lib = _lib.load()
    _require_cuda(features, "preprocess_points")
    if features.dtype != torch.float64 or features.dim() != 2:
        raise RuntimeError("preprocess_points: expected an (n, C) float64 feature table")
    idx = indexes.to(device=features.device, dtype=torch.int32)
    if idx.dim() == 1:
        idx = idx[None]
    idx = idx.contiguous()
    if idx.numel() and (int(idx.min()) < 0 or int(idx.max()) >= features.shape[0]):
        raise IndexError("preprocess_points: index out of range")
    f = features.contiguous()
    B, N = idx.shape
    out = torch.empty(B, f.shape[1], N, device=f.device, dtype=torch.float32)
    with torch.cuda.device(f.device):
        st = torch.cuda.current_stream(f.device).cuda_stream
        _lib.check(lib.ach_pre_points(f.data_ptr(), f.shape[0], f.shape[1], idx.data_ptr(), B, N, out.data_ptr(), st), "ach_pre_points")
    return out
