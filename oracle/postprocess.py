"""CPU restatement of the detection post-process: box decode + class-aware NMS.

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.

Restates ``utils/utils_bbox.py:33-85`` (decode_outputs) and ``:87-181``
(non_max_suppression, yolo_correct_boxes ``:5-30``).  The greedy NMS itself lives
in torchvision (third party; ``torchvision.ops.boxes.batched_nms`` called at
utils_bbox.py:125, pinned torchvision==0.12.0 / 0.26.0 in this image); its
published algorithm is restated in :func:`batched_nms_coordinate_trick`:
boxes are shifted by ``class_idx * (max_coord + 1)`` so different classes never
overlap, candidates are visited in stable descending-score order and a
candidate is dropped when its IoU with an already-kept box is ``> thr``.
All arithmetic is fp32 to match the reference bit for bit.
"""
import numpy as np
import torch


def decode_outputs(outputs, input_shape):
    """outputs: 3 tensors (B, 5+K, H, W) -> (B, sum(HW), 5+K), xywh normalised to [0,1].
    utils_bbox.py:33-85 (strides are input_shape[0] / h as Python floats, grid is (x, y))."""
    hw = [o.shape[-2:] for o in outputs]
    out = torch.cat([o.flatten(start_dim=2) for o in outputs], dim=2).permute(0, 2, 1).clone()
    out[:, :, 4:] = torch.sigmoid(out[:, :, 4:])
    grids, strides = [], []
    for h, w in hw:
        gy, gx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
        grids.append(torch.stack((gx, gy), 2).view(1, -1, 2).to(out.dtype))
        strides.append(torch.full((1, h * w, 1), input_shape[0] / h, dtype=out.dtype))
    grids, strides = torch.cat(grids, 1), torch.cat(strides, 1)
    out[..., :2] = (out[..., :2] + grids) * strides
    out[..., 2:4] = torch.exp(out[..., 2:4]) * strides
    out[..., [0, 2]] = out[..., [0, 2]] / input_shape[1]
    out[..., [1, 3]] = out[..., [1, 3]] / input_shape[0]
    return out


def nms_greedy(boxes, scores, thr):
    """torchvision::nms semantics on fp32 numpy arrays; returns kept indices, score-descending."""
    boxes = boxes.astype(np.float32)
    order = np.argsort(-scores.astype(np.float32), kind="stable")
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    areas = ((x2 - x1) * (y2 - y1)).astype(np.float32)
    suppressed = np.zeros(len(boxes), dtype=bool)
    keep = []
    thr = np.float32(thr)
    for _i, i in enumerate(order):
        if suppressed[i]:
            continue
        keep.append(i)
        rest = order[_i + 1:]
        xx1 = np.maximum(x1[i], x1[rest])
        yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest])
        yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(np.float32(0), xx2 - xx1)
        h = np.maximum(np.float32(0), yy2 - yy1)
        inter = (w * h).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (areas[i] + areas[rest] - inter)
        suppressed[rest[ovr > thr]] = True
    return np.asarray(keep, dtype=np.int64)


def batched_nms_coordinate_trick(boxes, scores, idxs, thr):
    if len(boxes) == 0:
        return np.zeros((0,), dtype=np.int64)
    boxes = boxes.astype(np.float32)
    max_coordinate = boxes.max()
    offsets = idxs.astype(np.float32) * (max_coordinate + np.float32(1))
    return nms_greedy(boxes + offsets[:, None], scores, thr)


def yolo_correct_boxes(box_xy, box_wh, input_shape, image_shape, letterbox_image):
    """utils_bbox.py:5-30 (numpy, y-first output, float64 shapes exactly like the reference)."""
    box_yx = box_xy[..., ::-1]
    box_hw = box_wh[..., ::-1]
    input_shape = np.array(input_shape)
    image_shape = np.array(image_shape)
    if letterbox_image:
        new_shape = np.round(image_shape * np.min(input_shape / image_shape))
        offset = (input_shape - new_shape) / 2.0 / input_shape
        scale = input_shape / new_shape
        box_yx = (box_yx - offset) * scale
        box_hw = (box_hw * scale).astype(box_hw.dtype)  # the reference multiplies in place on its fp32 view
    mins = box_yx - (box_hw / 2.0)
    maxes = box_yx + (box_hw / 2.0)
    boxes = np.concatenate([mins[..., 0:1], mins[..., 1:2], maxes[..., 0:1], maxes[..., 1:2]], axis=-1)
    boxes = boxes * np.concatenate([image_shape, image_shape], axis=-1)
    return boxes


def non_max_suppression(prediction, num_classes, input_shape, image_shape, letterbox_image,
                        conf_thres=0.5, nms_thres=0.4, return_indices=False):
    """prediction (B, A, 5+K) decoded -> list of None | ndarray (n, 7)
    [top, left, bottom, right, obj, cls_conf, cls_idx].  utils_bbox.py:87-181."""
    pred = prediction.detach().cpu().float().clone()
    xy, wh = pred[:, :, 0:2].clone(), pred[:, :, 2:4].clone()
    pred[:, :, 0:2] = xy - wh / 2
    pred[:, :, 2:4] = xy + wh / 2
    output, kept = [], []
    for image_pred in pred:
        class_conf, class_pred = torch.max(image_pred[:, 5:5 + num_classes], 1, keepdim=True)
        score = image_pred[:, 4] * class_conf[:, 0]
        cand = torch.nonzero(score >= conf_thres).flatten()
        det = torch.cat((image_pred[:, :5], class_conf, class_pred.float()), 1)[cand].numpy()
        keep = batched_nms_coordinate_trick(det[:, :4], (det[:, 4] * det[:, 5]).astype(np.float32), det[:, 6], nms_thres)
        det = det[keep]
        kept.append(cand.numpy()[keep])
        if len(cand) == 0:
            # the reference indexes an empty tensor and returns an empty (0, 7) array
            output.append(det.reshape(0, 7))
            continue
        box_xy, box_wh = (det[:, 0:2] + det[:, 2:4]) / 2, det[:, 2:4] - det[:, 0:2]
        det[:, :4] = yolo_correct_boxes(box_xy, box_wh, input_shape, image_shape, letterbox_image)
        output.append(det)
    return (output, kept) if return_indices else output


def seg_postprocess(seg_logits, image_shape, letterbox_image=True):
    """CPU oracle of the caller-side segmentation post-process, written with the reference's own calls
    (achelous.py:283-296): softmax -> crop of the letterbox window -> cv2.resize(INTER_LINEAR) -> argmax.
    seg_logits (K, H, W) torch tensor -> (prob (oh, ow, K) float32, argmax (oh, ow) int64)."""
    import cv2
    import torch.nn.functional as F
    K, H, W = seg_logits.shape
    oh, ow = int(image_shape[0]), int(image_shape[1])
    if letterbox_image:
        scale = min(W / ow, H / oh)
        nw, nh = int(ow * scale), int(oh * scale)
    else:
        nw, nh = W, H
    pr = F.softmax(seg_logits.permute(1, 2, 0), dim=-1).cpu().numpy()
    pr = pr[int((H - nh) // 2): int((H - nh) // 2 + nh), int((W - nw) // 2): int((W - nw) // 2 + nw)]
    pr = cv2.resize(pr, (ow, oh), interpolation=cv2.INTER_LINEAR)
    return pr, pr.argmax(axis=-1)
