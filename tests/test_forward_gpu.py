"""GPU parity tests proper: the CUDA forward (through the nn.Module facade -> C ABI) against
(a) the CPU oracle on the same seeded weights/inputs, incl. per-block taps, and
(b) the committed golden fixtures produced by the unmodified reference.

Tolerance (BASELINE.json north_star): 1e-3 relative -> max|a-b| / max|b| <= 1e-3 per tensor; the
kernels are fp32 end to end so the observed error is ~1e-5, asserted at 2e-4 to catch regressions.
Seg argmax: exact on every pixel whose top-2 margin exceeds 1e-4 of the logit range (ReLU-ed logits tie
exactly at 0 - ties are resolved lowest-index-first like torch.argmax - and near-ties flip with any change
of summation order); total mismatch additionally bounded at 0.1 % of pixels."""
import os

import numpy as np
import pytest
import torch

from achelous_b200.nets.Achelous import Achelous, Achelous3T
from achelous_b200.synthetic import make_inputs
from achelous_b200.weights import fill_state_dict
from oracle import functional as OF
from tests.common import (GOLDEN_CONFIGS, neck_of, MODEL_KW, REL_TOL, argmax_mismatch, load_golden, rel_err, summarize,
                          summary_rel_err)

pytestmark = pytest.mark.gpu

TIGHT = 2e-4
SUPPORTED = list(GOLDEN_CONFIGS)


def build(phi, bb, wseed, graph=True, fuse="chain", tc=True, neck="gdf"):
    model = Achelous(phi=phi, backbone=bb, **dict(MODEL_KW, neck=neck)).eval()
    model.fuse_seg_decoder = bool(fuse)
    model.fuse_seg_chain = fuse == "chain"
    model.use_tensor_cores = tc
    model.fuse_mlp = bool(fuse)     # the blockwise mode also keeps the MLP blocks as two GEMM launches
    sd = fill_state_dict(model.state_dict(), seed=wseed)
    model.load_state_dict(sd, strict=True)
    model.use_cuda_graph = graph
    return model.cuda(), sd


@pytest.mark.parametrize("fuse,tc", [("chain", True), (True, True), (False, "all"), (True, False)], ids=["chained_seg-tcgen05", "fused_seg-tcgen05", "blockwise_seg-tcgen05_everywhere", "fused_seg-simt"])
@pytest.mark.parametrize("name", SUPPORTED)
def test_forward_vs_golden_and_oracle(name, fuse, tc):
    phi, bb, wseed, iseed = GOLDEN_CONFIGS[name]
    model, sd = build(phi, bb, wseed, fuse=fuse, tc=tc, neck=neck_of(name))
    x, xr, pc = make_inputs(2, seed=iseed)
    det, se, lane, pcs = model(x.cuda(), xr.cuda(), pc.cuda())
    torch.cuda.synchronize()
    g = load_golden(name)
    # (b) golden fixtures from the reference
    for i in range(3):
        assert rel_err(det[i], g[f"det{i}"]) < TIGHT
    assert rel_err(pcs, g["pc"]) < TIGHT
    assert rel_err(se[:, :, ::4, ::4], g["se_sub"]) < TIGHT
    assert rel_err(lane[:, :, ::4, ::4], g["lane_sub"]) < TIGHT
    assert summary_rel_err(summarize(se), g["se_sum"]) < TIGHT
    for logits, key in ((se, "se_argmax"), (lane, "lane_argmax")):
        frac_all, n_safe_diff, frac_safe = argmax_mismatch(logits, g[key])
        assert n_safe_diff == 0 and frac_all < 1e-3, (key, frac_all, n_safe_diff, frac_safe)
    # exact zeros where the reference has them (ReLU-ed logits): compare zero masks on the sub-sampled maps
    z_ref = g["se_sub"] == 0
    z_mine = se[:, :, ::4, ::4].cpu().numpy() == 0
    assert (z_ref != z_mine).mean() < 1e-3
    # (a) per-block taps against the oracle
    taps = {}
    OF.achelous_forward(sd, x, xr, pc, phi=phi, backbone=bb, taps=taps, neck=neck_of(name))
    eng = next(iter(model._engines.values()))
    worst = {}
    for tname in eng.taps:
        worst[tname] = rel_err(eng.tap(tname), taps[tname])
    bad = {k: v for k, v in worst.items() if v >= TIGHT}
    assert not bad, bad
    for k in g.files:
        if k.startswith("tap/") and k[4:] in eng.taps:
            assert summary_rel_err(summarize(eng.tap(k[4:])), g[k]) < TIGHT, k


def test_graph_equals_eager_and_batch_invariance():
    """CUDA-graph replay == eager launches bit for bit; frame b of a batch == the same frame alone
    (no reduction crosses the batch dimension - the property the multi-GPU sharding relies on)."""
    phi, bb, wseed, iseed = GOLDEN_CONFIGS["en_gdf_pn_s0"]
    m_graph, _ = build(phi, bb, wseed, graph=True)
    m_eager, _ = build(phi, bb, wseed, graph=False)
    x, xr, pc = [t.cuda() for t in make_inputs(3, seed=5)]
    a = m_graph(x, xr, pc)
    a2 = m_graph(x, xr, pc)  # replay
    b = m_eager(x, xr, pc)
    single = m_eager(x[1:2], xr[1:2], pc[1:2])
    flat = lambda o: list(o[0]) + [o[1], o[2], o[3]]
    for ta, ta2, tb, ts in zip(flat(a), flat(a2), flat(b), flat(single)):
        assert torch.equal(ta, tb) and torch.equal(ta, ta2)
        assert torch.equal(ta[1:2], ts)


def test_three_task_variant_and_host_radar():
    phi, bb, wseed, iseed = GOLDEN_CONFIGS["en_gdf_pn_s0"]
    kw = dict(MODEL_KW)
    m3 = Achelous3T(phi=phi, backbone=bb, **kw).eval()
    full, sd = build(phi, bb, wseed)
    m3.load_state_dict({k: v for k, v in sd.items() if not k.startswith("pc_seg_model.")}, strict=True)
    m3 = m3.cuda()
    x, xr, pc = make_inputs(1, seed=9)
    det3, se3, lane3 = m3(x.cuda(), xr)  # x_radar left on the host, as achelous.py:212 does
    det, se, lane, _ = full(x.cuda(), xr.cuda(), pc.cuda())
    assert all(torch.equal(a, b) for a, b in zip(det3, det)) and torch.equal(se3, se) and torch.equal(lane3, lane)


def test_errors_are_loud():
    model, _ = build("S0", "en", 0)
    x, xr, pc = [t.cuda() for t in make_inputs(1, seed=1)]
    with pytest.raises(RuntimeError):
        model(x.double(), xr, pc)
    with pytest.raises(RuntimeError):
        model(x[:, :, :300], xr, pc)
    with pytest.raises(NotImplementedError):
        model.train()(x, xr, pc)
    with pytest.raises(NotImplementedError):
        Achelous(phi="S0", backbone="rv", **MODEL_KW)      # RepViT / PoolFormer are outside SURVEY.md §8 ('en', 'mv', 'ev', 'ef' are built)
    with pytest.raises(NotImplementedError):
        Achelous(phi="S0", backbone="en", **dict(MODEL_KW, neck="rdf"))


def test_weight_update_invalidates_packs():
    model, sd = build("S0", "en", 0)
    x, xr, pc = [t.cuda() for t in make_inputs(1, seed=3)]
    out0 = model(x, xr, pc)[1].clone()
    sd2 = fill_state_dict(model.state_dict(), seed=5)
    model.load_state_dict(sd2)
    out1 = model(x, xr, pc)[1]
    ref = OF.achelous_forward(sd2, x.cpu(), xr.cpu(), pc.cpu())[1]
    assert rel_err(out1, ref) < TIGHT and not torch.equal(out0, out1)


def test_config4_pn2_s2_vs_builder_oracle():
    """BASELINE config 4 (EN-GDF-PN2-S2).  The point-cloud branch is compared with the BUILDER-DEFINED oracle
    (oracle/pn2.py; the reference has no PN2 implementation - parity unpinned); the image/radar branches are
    the reference-pinned EN-S2 network.  FPS / ball-query indices must be bit-identical."""
    kw = dict(MODEL_KW, pc_seg="pn2")
    model = Achelous(phi="S2", backbone="en", **kw).eval()
    sd = fill_state_dict(model.state_dict(), seed=3)
    model.load_state_dict(sd, strict=True)
    model = model.cuda()
    x, xr, pc = make_inputs(3, seed=22)
    det, se, lane, pcs = model(x.cuda(), xr.cuda(), pc.cuda())
    taps = {}
    o_det, o_se, o_lane, o_pc = OF.achelous_forward(sd, x, xr, pc, phi="S2", backbone="en", pc_seg="pn2", taps=taps)
    eng = next(iter(model._engines.values()))
    for k, v in eng.taps_int.items():
        assert torch.equal(v.cpu().long(), taps[k]), k
    for name in ("pc.sa1.out", "pc.sa2.out", "pc.sa3.out", "pc.fp3", "pc.fp2", "pc.fp1"):
        assert rel_err(eng.tap(name), taps[name]) < TIGHT, name
    assert rel_err(pcs, o_pc) < TIGHT
    for i in range(3):
        assert rel_err(det[i], o_det[i]) < TIGHT
    assert rel_err(se, o_se) < TIGHT and rel_err(lane, o_lane) < TIGHT
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pn2_builder_defined.npz"))
    assert rel_err(pcs[:2], g["pc"]) < TIGHT


def test_stream_forward_matches_forward():
    """The pipelined throughput API returns exactly what forward() returns, batch by batch, in order."""
    model, _ = build("S0", "en", 2)
    batches = [make_inputs(2, seed=100 + i) for i in range(5)]
    ref = [model(*[t.cuda() for t in b]) for b in batches]
    ref = [([d.cpu() for d in r[0]], r[1].cpu(), r[2].cpu(), r[3].cpu()) for r in ref]
    pinned = [tuple(t.pin_memory() for t in b) for b in batches]
    n = 0
    for out, r in zip(model.stream_forward(iter(pinned)), ref):
        for a, b_ in zip(list(out[0]) + [out[1], out[2], out[3]], list(r[0]) + [r[1], r[2], r[3]]):
            assert torch.equal(a, b_)
        n += 1
    assert n == 5


@pytest.mark.parametrize("sizes", [[2], [2, 2], [2, 2, 3, 3, 2, 1, 1, 1]])
def test_stream_forward_lookahead_and_batch_size_changes(sizes):
    """Inputs are staged one batch ahead; a change of batch size drains the pipeline and rebinds the slot's buffers.
    Every yielded result must equal forward() on the same batch, in order, for any length / size pattern."""
    model, _ = build("S0", "en", 2)
    batches = [make_inputs(b, seed=200 + i) for i, b in enumerate(sizes)]
    ref = []
    for b in batches:
        r = model(*[t.cuda() for t in b])
        ref.append([d.cpu().clone() for d in r[0]] + [r[1].cpu().clone(), r[2].cpu().clone(), r[3].cpu().clone()])
    pinned = [tuple(t.pin_memory() for t in b) for b in batches]
    n = 0
    for out, r in zip(model.stream_forward(iter(pinned)), ref):
        got = list(out[0]) + [out[1], out[2], out[3]]
        for a, b_ in zip(got, r):
            assert a.shape == b_.shape and torch.equal(a, b_), n
        n += 1
    assert n == len(sizes)


# ------------------------------------------------------------------ round-2 parity holes
def _flat(o):
    return list(o[0]) + [o[1], o[2], o[3]]


def test_b1_outputs_are_fresh_tensors():
    """with B == 1 the slices of the static output buffer are already contiguous; forward() must still hand out
    fresh storage - predict.py's B=1 path keeps results across calls"""
    model, _ = build("S0", "en", 2)
    x1, r1, p1 = [t.cuda() for t in make_inputs(1, seed=31)]
    x2, r2, p2 = [t.cuda() for t in make_inputs(1, seed=32)]
    a = model(x1, r1, p1)
    snap = [t.clone() for t in _flat(a)]
    b = model(x2, r2, p2)
    eng = next(iter(model._engines.values()))
    for ta, ts, tb in zip(_flat(a), snap, _flat(b)):
        assert torch.equal(ta, ts) and not torch.equal(ta, tb)
        assert ta.data_ptr() != tb.data_ptr()
        assert not (eng.packed_out.data_ptr() <= ta.data_ptr() < eng.packed_out.data_ptr() + eng.packed_out.numel() * 4)


@pytest.mark.parametrize("name", ["en_gdf_pn_s0", "mv_gdf_pn_s0"])
def test_benchmark_batch_64_equals_small_batches_and_golden(name):
    """The benchmarked batch (BASELINE configs 2/3: B = 64) runs tile schedules (persistent CTAs sized from B) that the B <= 3
    tests never see: forward(B=64) must equal 32 x forward(B=2) BITWISE, and the two golden frames embedded in the batch must
    match the reference-generated fixtures."""
    phi, bb, wseed, iseed = GOLDEN_CONFIGS[name]
    model, _ = build(phi, bb, wseed)
    parts = [make_inputs(2, seed=500 + i) for i in range(32)]
    parts[10] = make_inputs(2, seed=iseed)                      # golden frames at 20, 21
    x, xr, pc = [torch.cat([p[j] for p in parts]).cuda() for j in range(3)]
    big = _flat(model(x, xr, pc))
    for i in (0, 10, 17, 31):
        small = _flat(model(x[2 * i:2 * i + 2], xr[2 * i:2 * i + 2], pc[2 * i:2 * i + 2]))
        for tb, ts in zip(big, small):
            assert torch.equal(tb[2 * i:2 * i + 2], ts), (name, i)
    g = load_golden(name)
    det, se, lane, pcs = big[:3], big[3][20:22], big[4][20:22], big[5][20:22]
    for i in range(3):
        assert rel_err(det[i][20:22], g[f"det{i}"]) < TIGHT
    assert rel_err(pcs, g["pc"]) < TIGHT
    assert rel_err(se[:, :, ::4, ::4], g["se_sub"]) < TIGHT and rel_err(lane[:, :, ::4, ::4], g["lane_sub"]) < TIGHT
    for logits, key in ((se, "se_argmax"), (lane, "lane_argmax")):
        frac_all, n_safe_diff, _ = argmax_mismatch(logits, g[key])
        assert n_safe_diff == 0 and frac_all < 1e-3


def test_elementwise_relative_error():
    """North-star wording is '1e-3 relative'; the other tests use the max-norm max|a-b| / max|b| (README).  Second, ELEMENT-WISE
    check against the oracle evaluated in float64 ("truth"): |a - t| / |t| on every element with |t| > 1e-2 * max|t| (below that,
    ReLU-ed maps sit at cancellation level).  The network itself amplifies fp32 rounding on such small elements: the reference's
    OWN fp32 arithmetic (oracle in fp32) deviates from truth by up to 8e-4 element-wise on the lane map (1e-5 in max-norm).  Bound
    asserted: 1e-3, or 8x the reference's own fp32 deviation where that is larger; the observed figures are printed."""
    name = "en_gdf_pn_s0"
    phi, bb, wseed, iseed = GOLDEN_CONFIGS[name]
    model, sd = build(phi, bb, wseed)
    x, xr, pc = make_inputs(2, seed=iseed)
    mine = _flat(model(x.cuda(), xr.cuda(), pc.cuda()))
    ref32 = _flat(OF.achelous_forward(sd, x, xr, pc, phi=phi, backbone=bb))
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    truth = _flat(OF.achelous_forward(sd64, x.double(), xr.double(), pc.double(), phi=phi, backbone=bb))
    report, bad = [], []
    for nm, a, r, t in zip(("det40", "det20", "det10", "se", "lane", "pc"), mine, ref32, truth):
        a, r = a.cpu().double(), r.double()
        big = t.abs() > 1e-2 * t.abs().max()
        e_mine = ((a - t).abs()[big] / t.abs()[big]).max().item()
        e_ref = ((r - t).abs()[big] / t.abs()[big]).max().item()
        m_mine = ((a - t).abs().max() / t.abs().max()).item()
        report.append(f"{nm}: ours {e_mine:.2e} (max-norm {m_mine:.2e}), reference fp32 {e_ref:.2e}")
        if e_mine > max(REL_TOL, 8 * e_ref) or m_mine > 1e-4:
            bad.append(nm)
    print("element-wise relative error vs float64 truth, |t| > 1e-2 max|t|:\n  " + "\n  ".join(report))
    assert not bad, report


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_dataparallel_two_gpus_bitwise():
    """achelous.py:176-177 wraps the net in nn.DataParallel unconditionally: replicas own no parameters and are called from one
    thread per GPU.  x_radar stays on the host as in achelous.py:212."""
    model, _ = build("S0", "en", 2)
    x, xr, pc = make_inputs(4, seed=77)
    single = _flat(model(x.cuda(), xr.cuda(), pc.cuda()))
    dp = torch.nn.DataParallel(model).cuda()
    for _ in range(2):      # second call: replicas are new objects, plans and packed weights are reused
        multi = _flat(dp(x.cuda(), xr, pc.cuda()))
        for a, b in zip(single, multi):
            assert a.device == b.device and torch.equal(a, b)
    devs = {k[0] for k in model._engines}
    assert devs == {0, 1}, devs
    n_eng = len(model._engines)
    dp(x.cuda(), xr, pc.cuda())
    assert len(model._engines) == n_eng


def _compact_reference(model, x, xr, pc, conf, iou, keep=None):
    from achelous_b200.utils.utils_bbox import decode_outputs, nms_device
    det, se, lane, pcs = model(x, xr, pc)
    kept, _, counts = nms_device(decode_outputs(det, (320, 320), 0), 7, conf, iou)
    am = se.argmax(1)
    if keep is not None:
        lut = torch.tensor([k in keep for k in range(se.shape[1])], device=am.device)
        am = torch.where(lut[am], am, torch.zeros_like(am))
    return kept, counts, am, lane.argmax(1), pcs.argmax(-1)


@pytest.mark.parametrize("name,keep", [("en_gdf_pn_s0", None), ("en_gdf_pn_s0", (0, 8)), ("mv_gdf_pn_s0", None), ("en_cdf_pn_s0", None)])
def test_compact_outputs_equal_postprocessed_raw_outputs(name, keep):
    """forward(outputs="compact") == argmax / decode+NMS of forward() on the same inputs, BITWISE: the class maps are taken over
    the very same fp32 logits in registers, the rows are the NMS kernel's"""
    phi, bb, wseed, iseed = GOLDEN_CONFIGS[name]
    model, _ = build(phi, bb, wseed, neck=neck_of(name))
    with torch.no_grad():
        for k in range(3):
            model.det_head.obj_preds[k].bias += 2.0
            model.det_head.reg_preds[k].bias[2:] += 1.3
    x, xr, pc = [t.cuda() for t in make_inputs(3, seed=iseed)]
    kept, counts, se_am, lane_am, pc_am = _compact_reference(model, x, xr, pc, 0.3, 0.45, keep)
    c = model(x, xr, pc, outputs="compact", conf_thres=0.3, nms_thres=0.45, max_det=64, keep_classes=keep)
    assert torch.equal(c.se_mask.long(), se_am) and torch.equal(c.lane_mask.long(), lane_am) and torch.equal(c.pc_cls.long(), pc_am)
    assert torch.equal(c.det_count, counts) and int(counts.max()) > 0
    for b in range(3):
        m = min(int(counts[b]), 64)
        assert torch.equal(c.det_rows[b, :m], kept[b, :m]) and not c.det_rows[b, m:].any()
    if name == "en_gdf_pn_s0" and keep is None:
        g = load_golden(name)
        c2 = model(*[t.cuda() for t in make_inputs(2, seed=iseed)], outputs="compact")
        for mask, key in ((c2.se_mask, "se_argmax"), (c2.lane_mask, "lane_argmax")):
            assert (mask.cpu().numpy() != g[key]).mean() < 1e-3       # vs the reference's own argmax (near-ties excepted)


@pytest.mark.parametrize("shape", [(360, 640), (97, 211)])
def test_compact_masks_at_original_image_size(shape):
    """image_shape=(h, w): softmax -> letterbox crop -> cv2-style bilinear resize -> argmax (achelous.py:283-297) fused into one
    kernel inside the plan; equal to the standalone two-kernel post-process (itself pinned to torch + cv2)"""
    from achelous_b200.utils.seg_post import seg_argmax
    model, _ = build("S0", "en", 2)
    x, xr, pc = [t.cuda() for t in make_inputs(2, seed=41)]
    det, se, lane, pcs = model(x, xr, pc)
    c = model(x, xr, pc, outputs="compact", image_shape=shape, keep_classes=(0, 8))
    se_ref = seg_argmax(se, shape, True)
    lut = torch.tensor([k in (0, 8) for k in range(9)], device=se_ref.device)
    se_ref = torch.where(lut[se_ref.long()], se_ref, torch.zeros_like(se_ref))
    assert c.se_mask.shape == (2,) + tuple(shape) and torch.equal(c.se_mask, se_ref)
    assert torch.equal(c.lane_mask, seg_argmax(lane, shape, True))


def test_stream_forward_compact_and_buffer_lifetime():
    """stream_forward(compact=True) yields forward(outputs="compact") batch by batch; a yielded batch stays intact while the next
    one is requested and consumed (it used to be overwritten one request later)"""
    model, _ = build("S0", "en", 2)
    batches = [make_inputs(2, seed=300 + i) for i in range(6)]
    ref = [model(*[t.cuda() for t in b], outputs="compact") for b in batches]
    pinned = [tuple(t.pin_memory() for t in b) for b in batches]
    prev = None
    for i, out in enumerate(model.stream_forward(iter(pinned), compact=True)):
        for a, b_ in zip(out, ref[i]):
            assert torch.equal(a, b_.cpu()), i
        if prev is not None:            # batch i-1 is still intact after batch i has been produced
            torch.cuda.synchronize()
            for a, b_ in zip(prev, ref[i - 1]):
                assert torch.equal(a, b_.cpu()), ("lifetime", i)
        prev = out
    # raw mode, same property
    ref = [model(*[t.cuda() for t in b]) for b in batches]
    prev = None
    for i, out in enumerate(model.stream_forward(iter(pinned))):
        if prev is not None:
            torch.cuda.synchronize()
            for a, b_ in zip(_flat(prev), _flat(ref[i - 1])):
                assert torch.equal(a, b_.cpu()), ("lifetime raw", i)
        prev = out
    with pytest.raises(RuntimeError):   # every batch is validated (it used to be silently cast / broadcast)
        list(model.stream_forward(iter([(pinned[0][0].double(), pinned[0][1], pinned[0][2])])))


def test_radar_blocks_vs_torchvision_deform_conv2d_cuda():
    """RCBlock (RadarEncoder.py:65-74 -> dcn.py:56 torchvision.ops.deform_conv2d) against torchvision's own CUDA kernel on this
    device: the oracle's block composition with its deformable-conv restatement swapped for the library call"""
    tv = pytest.importorskip("torchvision")
    model, sd = build("S0", "en", 2)
    x, xr, pc = [t.cuda() for t in make_inputs(2, seed=51)]
    model(x, xr, pc)
    eng = next(iter(model._engines.values()))
    old = OF.deform_conv2d_3x3
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    OF.deform_conv2d_3x3 = lambda x_, off, mask, w: tv.ops.deform_conv2d(x_, off, w, None, stride=1, padding=1, mask=mask)
    try:
        taps = {}
        sd_cuda = {k: v.cuda() for k, v in sd.items() if k.startswith("image_radar_encoder.radar_encoder.")}
        OF.rcnet(xr, OF.SD(sd_cuda, "image_radar_encoder.radar_encoder."), taps)
    finally:
        OF.deform_conv2d_3x3 = old
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    for i in range(8):
        assert rel_err(eng.tap(f"radar.block{i}"), taps[f"radar.block{i}"]) < TIGHT, i


@pytest.mark.parametrize("bb", ["en", "mv"])
def test_graph_replays_are_reproducible(bb):
    """Run-to-run determinism of the multi-lane CUDA graph: 200 replays of a B = 32 plan on the same inputs must give bit-identical
    outputs AND intermediate buffers.  Round 2 finding: with a kernel-parameter tensor-map load in the GEMM kernel 10 %
    of the replays differed in one accumulator tile once several lanes shared the GPU; tools/determinism_probe.py is the long version
    of this test (B = 64, copy traffic on a side stream, per-buffer report)."""
    model, _ = build("S0", bb, 0)
    B = 32
    x, xr, pc = [t.cuda() for t in make_inputs(B, seed=11)]
    model(x, xr, pc)
    eng = next(e for e in model._engines.values() if e.B == B)
    bufs = [t for n, t in eng._bufs.items() if not n.startswith("in.")] + [eng.packed_out]

    def sums():
        return torch.stack([b.contiguous().view(torch.int32).sum(dtype=torch.int64) for b in bufs])

    eng.forward_static()
    torch.cuda.synchronize()
    ref = sums()
    bad = 0
    for _ in range(200):
        eng.forward_static()
        bad += int((sums() != ref).any().item())
    torch.cuda.synchronize()
    assert bad == 0, f"{bad} of 200 replays differ from the first one"
