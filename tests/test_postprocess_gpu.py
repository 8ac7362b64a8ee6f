"""GPU: decode + NMS kernels (achelous_b200.utils.utils_bbox) against the CPU oracle and the golden
fixtures of the reference's own decode_outputs / non_max_suppression.  Kept indices: bit-exact."""
import numpy as np
import pytest
import torch

from achelous_b200.utils import utils_bbox as UB
from oracle import postprocess as OP
from tests.common import GOLDEN_CONFIGS, WH_BIAS, load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(GOLDEN_CONFIGS))
def test_decode_and_nms_vs_golden(name):
    g = load_golden(name)
    det = [torch.from_numpy(g[f"det{i}"]).clone() for i in range(3)]
    for d in det:
        d[:, 4] += float(g["obj_bias"])
        d[:, 2:4] += WH_BIAS
    decoded = UB.decode_outputs([d.cuda() for d in det], (320, 320), 0)
    # exp/sigmoid come from different libm implementations on host and device: 2 ulp
    np.testing.assert_allclose(decoded.cpu().numpy(), g["decoded"], rtol=3e-7, atol=1e-7)
    gold_dec = torch.from_numpy(g["decoded"]).cuda()
    for tag, conf, iou, shape, lb in (("a", 0.35, 0.35, (320, 320), False), ("b", 0.25, 0.5, (1080, 1920), True)):
        kept, kept_idx, counts = UB.nms_device(gold_dec.clone(), 7, conf, iou)
        o_res, o_idx = OP.non_max_suppression(gold_dec.cpu(), 7, (320, 320), np.array(shape), lb, conf, iou, return_indices=True)
        res = UB.non_max_suppression(gold_dec.clone(), 7, (320, 320), np.array(shape), lb, conf, iou)
        for b in range(gold_dec.shape[0]):
            n = int(counts[b])
            assert np.array_equal(kept_idx[b, :n].cpu().numpy(), o_idx[b]), (tag, b)   # bit-exact kept indices vs oracle
            np.testing.assert_allclose(res[b], o_res[b], rtol=1e-6, atol=1e-6)
            gold = g[f"nms_{tag}_{b}"]
            n_cand = int(((g["decoded"][b, :, 4] * g["decoded"][b, :, 5:].max(-1)) >= conf).sum())
            if n_cand <= 1000:  # above that the reference's CPU torchvision takes the per-class branch
                assert res[b].shape == gold.shape
                np.testing.assert_allclose(res[b], gold, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("n_boxes,seed", [(0, 0), (1, 1), (37, 2), (700, 3), (2100, 4)])
def test_nms_random_dense(n_boxes, seed):
    """Heavily overlapping random boxes incl. exact score ties; compares kept anchor indices bit-exactly."""
    rng = np.random.default_rng(seed)
    B, A, K = 3, 2100, 7
    pred = np.zeros((B, A, 5 + K), np.float32)
    pred[..., 0:2] = rng.uniform(0.2, 0.8, (B, A, 2))
    pred[..., 2:4] = rng.uniform(0.05, 0.4, (B, A, 2))
    pred[..., 4] = 0.01
    for b in range(B):
        idx = rng.choice(A, n_boxes, replace=False)
        pred[b, idx, 4] = np.round(rng.uniform(0.5, 1.0, n_boxes), 2)  # rounded -> many exact ties
    pred[..., 5:] = np.round(rng.uniform(0.5, 1.0, (B, A, K)), 1)
    t = torch.from_numpy(pred)
    kept, kept_idx, counts = UB.nms_device(t.cuda(), K, 0.2, 0.45)
    _, o_idx = OP.non_max_suppression(t, K, (320, 320), np.array((320, 320)), False, 0.2, 0.45, return_indices=True)
    for b in range(B):
        n = int(counts[b])
        assert np.array_equal(kept_idx[b, :n].cpu().numpy(), o_idx[b])


@pytest.mark.parametrize("shape,lb", [((1080, 1920), True), ((480, 640), True), ((320, 320), False), ((97, 211), True)])
def test_seg_postprocess_on_device(shape, lb):
    """softmax -> letterbox crop -> cv2-style bilinear resize -> argmax on the GPU vs the reference's host sequence
    (oracle/postprocess.py::seg_postprocess = torch softmax + cv2.resize).  Exact wherever the interpolated top-2
    probabilities differ by more than 1e-5 (OpenCV's SIMD build may contract a*b+c differently by one ulp)."""
    from achelous_b200.utils.seg_post import seg_argmax
    torch.manual_seed(0)
    B, K, H, W = 2, 9, 320, 320
    logits = torch.relu(torch.randn(B, K, H, W) * 2)
    # smooth the logits so that class regions exist (and ReLU zeros tie exactly, as in the real head)
    logits = torch.nn.functional.avg_pool2d(logits, 9, 1, 4)
    mine = seg_argmax(logits.cuda(), shape, lb).cpu().numpy()
    for b in range(B):
        pr, am = OP.seg_postprocess(logits[b], shape, lb)
        top2 = np.sort(pr, axis=-1)[..., -2:]
        safe = (top2[..., 1] - top2[..., 0]) > 1e-5
        diff = mine[b] != am
        assert (diff & safe).sum() == 0
        assert diff.mean() < 2e-3


@pytest.mark.parametrize("n_boxes,seed", [(37, 2), (700, 3), (1500, 5), (2100, 4)])
def test_nms_vs_torchvision_batched_nms_on_cuda(n_boxes, seed):
    """utils/utils_bbox.py:125 calls torchvision.ops.boxes.batched_nms on the device the predictions live on.  Its CPU build
    switches to the per-class branch above 1000 candidates (boxes.numel() > 4000), the CUDA build keeps the coordinate trick up
    to 100 000 - so the reference ON A GPU is pinned here, against torchvision's own CUDA kernel, incl. > 1000 candidates."""
    tv = pytest.importorskip("torchvision")
    rng = np.random.default_rng(seed)
    B, A, K = 3, 2100, 7
    pred = np.zeros((B, A, 5 + K), np.float32)
    pred[..., 0:2] = rng.uniform(0.2, 0.8, (B, A, 2))
    pred[..., 2:4] = rng.uniform(0.05, 0.4, (B, A, 2))
    pred[..., 4] = 0.01
    for b in range(B):
        idx = rng.choice(A, n_boxes, replace=False)
        pred[b, idx, 4] = np.round(rng.uniform(0.5, 1.0, n_boxes), 2)
    pred[..., 5:] = np.round(rng.uniform(0.5, 1.0, (B, A, K)), 1)
    t = torch.from_numpy(pred).cuda()
    kept, kept_idx, counts = UB.nms_device(t.clone(), K, 0.2, 0.45)
    # the reference's per-image sequence (utils_bbox.py:95-130), on CUDA tensors
    box = t.clone()
    box[:, :, 0:2] = t[:, :, 0:2] - t[:, :, 2:4] / 2
    box[:, :, 2:4] = t[:, :, 0:2] + t[:, :, 2:4] / 2
    for b in range(B):
        conf, cls = torch.max(box[b, :, 5:5 + K], 1, keepdim=True)
        mask = (box[b, :, 4] * conf[:, 0] >= 0.2)
        det = torch.cat((box[b, :, :5], conf, cls.float()), 1)[mask]
        keep = tv.ops.boxes.batched_nms(det[:, :4], det[:, 4] * det[:, 5], det[:, 6], 0.45)
        ref_rows = det[keep]
        ref_idx = torch.nonzero(mask).flatten()[keep]
        n = int(counts[b])
        assert n == ref_rows.shape[0], (b, n, ref_rows.shape[0])
        assert torch.equal(kept_idx[b, :n].long(), ref_idx)
        assert torch.equal(kept[b, :n], ref_rows)


def test_nms_rows_compact_record():
    """ach_nms_rows: first max_keep rows + true count at caller strides, zeros behind the survivors"""
    import ctypes as C
    from achelous_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(9)
    B, A, K, MK = 2, 2100, 7, 16
    pred = np.zeros((B, A, 5 + K), np.float32)
    pred[..., 0:2] = rng.uniform(0.1, 0.9, (B, A, 2))
    pred[..., 2:4] = rng.uniform(0.02, 0.1, (B, A, 2))
    pred[..., 4] = rng.uniform(0, 1, (B, A))
    pred[1, :, 4] = 0.0                                   # image 1: no candidates
    pred[..., 5:] = rng.uniform(0.3, 1.0, (B, A, K))
    t = torch.from_numpy(pred).cuda()
    kept, kept_idx, counts = UB.nms_device(t.clone(), K, 0.5, 0.4)
    rec = torch.full((B, 64 + MK * 7), 7.0, device="cuda")
    cnt = rec.view(torch.int32)
    ws_bytes = lib.ach_nms_workspace_bytes(B, A)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    _lib.check(lib.ach_nms_rows(t.data_ptr(), B, A, K, 0.5, 0.4, rec.data_ptr() + 64 * 4, rec.stride(0), MK, cnt.data_ptr(), rec.stride(0),
                                ws.data_ptr(), ws_bytes, torch.cuda.current_stream().cuda_stream), "ach_nms_rows")
    torch.cuda.synchronize()
    assert int(counts[0]) > MK and int(counts[1]) == 0
    assert cnt[:, 0].tolist() == counts.tolist()
    assert torch.equal(rec[0, 64:].view(MK, 7), kept[0, :MK]) and not rec[1, 64:].any()
    assert (rec[:, 1:64] == 7.0).all()
