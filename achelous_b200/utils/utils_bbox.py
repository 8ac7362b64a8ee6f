"""Drop-in replacements of the reference's detection post-process free functions
(``utils/utils_bbox.py``: ``decode_outputs`` :33-85, ``non_max_suppression`` :87-181), same names,
argument meaning and return types, executed by the sm_100a kernels in ``csrc/postprocess.cu``.

    from achelous_b200.utils.utils_bbox import decode_outputs, non_max_suppression

Differences by design: the per-call CPU construction + ``.cuda()`` upload of grids/strides
(:54-74) and the Python loop over images with one torchvision call each (:103-130) are gone - decode is
one launch, NMS is one launch for the whole batch.  The final letterbox un-warp stays on the host in
numpy float64 exactly like the reference (:5-30, :177-180)."""
import ctypes as C

import numpy as np
import torch

from .. import _lib


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


def decode_outputs(outputs, input_shape, local_rank=0):
    """outputs: list of (B, 5+K, H_l, W_l) raw head maps -> (B, sum H_l*W_l, 5+K) decoded boxes."""
    lib = _lib.load()
    outs = [o if o.stride(-1) == 1 and o.stride(-2) == o.shape[-1] and o.stride(1) == o.shape[-1] * o.shape[-2] else o.contiguous()
            for o in outputs]
    x0 = outs[0]
    if not x0.is_cuda:
        raise RuntimeError("decode_outputs: achelous_b200 has no CPU path")
    if any(o.dtype != torch.float32 for o in outs):
        raise RuntimeError("decode_outputs: expected float32 head outputs")
    B, CH = x0.shape[0], x0.shape[1]
    n = len(outs)
    A = sum(o.shape[-2] * o.shape[-1] for o in outs)
    out = torch.empty(B, A, CH, device=x0.device, dtype=torch.float32)
    ptrs = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
    bss = (C.c_longlong * n)(*[o.stride(0) for o in outs])
    hs = (C.c_int * n)(*[o.shape[-2] for o in outs])
    ws = (C.c_int * n)(*[o.shape[-1] for o in outs])
    with torch.cuda.device(x0.device):
        _lib.check(lib.ach_decode_outputs(ptrs, bss, hs, ws, n, out.data_ptr(), B, CH - 5, float(input_shape[0]),
                                          float(input_shape[1]), _stream(x0.device)), "ach_decode_outputs")
    return out


def nms_device(prediction, num_classes, conf_thres, nms_thres):
    """Device part of non_max_suppression: returns (kept (B, A, 7), kept_idx (B, A) int32, counts (B) int32)."""
    lib = _lib.load()
    if not prediction.is_cuda:
        raise RuntimeError("non_max_suppression: achelous_b200 has no CPU path")
    pred = prediction.contiguous().float()
    B, A, CH = pred.shape
    if CH != 5 + num_classes:
        raise RuntimeError(f"non_max_suppression: prediction has {CH} channels, expected {5 + num_classes}")
    kept = torch.empty(B, A, 7, device=pred.device, dtype=torch.float32)
    kept_idx = torch.empty(B, A, device=pred.device, dtype=torch.int32)
    counts = torch.zeros(B, device=pred.device, dtype=torch.int32)
    ws_bytes = lib.ach_nms_workspace_bytes(B, A)
    ws = torch.empty(ws_bytes, device=pred.device, dtype=torch.uint8)
    with torch.cuda.device(pred.device):
        _lib.check(lib.ach_nms(pred.data_ptr(), B, A, num_classes, float(conf_thres), float(nms_thres), kept.data_ptr(),
                               kept_idx.data_ptr(), counts.data_ptr(), ws.data_ptr(), ws_bytes, _stream(pred.device)), "ach_nms")
    return kept, kept_idx, counts


def yolo_correct_boxes(box_xy, box_wh, input_shape, image_shape, letterbox_image):
    """Host letterbox un-warp, y-first output (utils_bbox.py:5-30)."""
    box_yx = box_xy[..., ::-1]
    box_hw = box_wh[..., ::-1]
    input_shape = np.array(input_shape)
    image_shape = np.array(image_shape)
    if letterbox_image:
        new_shape = np.round(image_shape * np.min(input_shape / image_shape))
        offset = (input_shape - new_shape) / 2. / input_shape
        scale = input_shape / new_shape
        box_yx = (box_yx - offset) * scale
        box_hw = (box_hw * scale).astype(box_hw.dtype)
    box_mins = box_yx - (box_hw / 2.)
    box_maxes = box_yx + (box_hw / 2.)
    boxes = np.concatenate([box_mins[..., 0:1], box_mins[..., 1:2], box_maxes[..., 0:1], box_maxes[..., 1:2]], axis=-1)
    boxes *= np.concatenate([image_shape, image_shape], axis=-1)
    return boxes


def non_max_suppression(prediction, num_classes, input_shape, image_shape, letterbox_image, conf_thres=0.5, nms_thres=0.4):
    """Same contract as the reference: list (len B) of ndarray (n_i, 7)
    [top, left, bottom, right, obj_conf, class_conf, class_pred] in original-image pixels.
    Like the reference, `prediction[:, :, :4]` is converted to corner form in place."""
    kept, _, counts = nms_device(prediction, num_classes, conf_thres, nms_thres)
    xy, wh = prediction[:, :, 0:2].clone(), prediction[:, :, 2:4].clone()
    prediction[:, :, 0:2] = xy - wh / 2
    prediction[:, :, 2:4] = xy + wh / 2
    counts_h = counts.cpu().numpy()
    kept_h = kept.cpu().numpy()
    output = []
    for i in range(prediction.shape[0]):
        det = kept_h[i, : counts_h[i]].copy()
        if len(det):
            box_xy, box_wh = (det[:, 0:2] + det[:, 2:4]) / 2, det[:, 2:4] - det[:, 0:2]
            det[:, :4] = yolo_correct_boxes(box_xy, box_wh, input_shape, image_shape, letterbox_image)
        output.append(det)
    return output
