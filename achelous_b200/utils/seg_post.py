"""On-device segmentation post-process.

The reference does this on the host for every frame (achelous.py:283-318): softmax -> ``.cpu()`` of the full
(H, W, K) fp32 map -> letterbox crop -> ``cv2.resize(..., INTER_LINEAR)`` -> argmax.  Here the same sequence runs in
two kernels on the GPU and only the uint8 class map (1 byte per ORIGINAL-image pixel) is left to copy out.

    from achelous_b200.utils.seg_post import seg_argmax
    mask = seg_argmax(se_seg, image_shape=(orig_h, orig_w), letterbox_image=True)   # (B, orig_h, orig_w) uint8, cuda
"""
import torch

from .. import _lib


def letterbox_window(input_shape, image_shape, letterbox_image=True):
    """(y_off, x_off, nh, nw) of the un-padded image inside the network input (achelous.py:196-199,287-288)."""
    ih, iw = int(input_shape[0]), int(input_shape[1])
    oh, ow = int(image_shape[0]), int(image_shape[1])
    if not letterbox_image:
        return 0, 0, ih, iw
    scale = min(iw / ow, ih / oh)
    nw, nh = int(ow * scale), int(oh * scale)
    return int((ih - nh) // 2), int((iw - nw) // 2), nh, nw


def seg_argmax(seg_logits, image_shape, letterbox_image=True):
    """seg_logits (B, K, H, W) fp32 cuda -> (B, orig_h, orig_w) uint8 class map on the same device."""
    lib = _lib.load()
    if not seg_logits.is_cuda or seg_logits.dtype != torch.float32:
        raise RuntimeError("seg_argmax: expected a float32 CUDA tensor (achelous_b200 has no CPU path)")
    x = seg_logits.contiguous()
    B, K, H, W = x.shape
    y_off, x_off, nh, nw = letterbox_window((H, W), image_shape, letterbox_image)
    OH, OW = int(image_shape[0]), int(image_shape[1])
    prob = torch.empty_like(x)
    out = torch.empty(B, OH, OW, device=x.device, dtype=torch.uint8)
    with torch.cuda.device(x.device):
        st = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(lib.ach_seg_softmax(x.data_ptr(), x.stride(0), prob.data_ptr(), prob.stride(0), B, K, H * W, st), "ach_seg_softmax")
        _lib.check(lib.ach_seg_resize_argmax(prob.data_ptr(), prob.stride(0), B, K, H, W, y_off, x_off, nh, nw, out.data_ptr(), OH, OW, st),
                   "ach_seg_resize_argmax")
    return out
