"""CPU: the engine's host logic (weight folding/packing, views, wiring, output packing) executed through
the C-ABI emulator (tests/abi_emulator.py) on a dry-run plan, compared with the oracle."""
import pytest
import torch

from achelous_b200.engine import Engine
from achelous_b200.nets.Achelous import Achelous
from achelous_b200.synthetic import make_inputs
from achelous_b200.weights import fill_state_dict
from oracle import functional as OF
from achelous_b200 import _lib
from achelous_b200._lib import AchRcDeform
from tests import abi_emulator as emu
from tests.abi_emulator import emulate_engine
from tests.common import MODEL_KW, argmax_mismatch, rel_err

TOL = 5e-5


@pytest.mark.parametrize("phi,backbone,seed,fuse,tc,neck", [("S0", "en", 2, True, "all", "gdf"), ("S0", "en", 2, False, False, "gdf"), ("S2", "en", 0, True, True, "gdf"),
                                                             ("S0", "mv", 0, True, True, "gdf"), ("S0", "en", 3, True, True, "cdf"), ("S0", "ev", 4, True, True, "gdf"), ("S0", "ef", 5, True, True, "gdf"), ("S1", "ev", 7, True, True, "gdf"), ("S2", "ef", 7, True, True, "cdf")])
def test_plan_matches_oracle(phi, backbone, seed, fuse, tc, neck):
    torch.set_num_threads(4)
    model = Achelous(phi=phi, backbone=backbone, **dict(MODEL_KW, neck=neck)).eval()
    model.fuse_seg_decoder = fuse
    model.fuse_seg_chain = fuse and phi == "S0" and backbone == "en" and tc == "all" or (phi == "S2")
    model.use_tensor_cores = tc
    sd = fill_state_dict(model.state_dict(), seed=seed)
    model.load_state_dict(sd, strict=True)
    B = 2
    x, xr, pc = make_inputs(B, seed=21)
    eng = Engine(model, B, "cpu", dry_run=True)
    ins = eng.input_tensors()
    ins[0].copy_(x), ins[1].copy_(xr), ins[2].copy_(pc)
    emulate_engine(eng)
    det, se, lane, pcs = eng.output_views()
    taps = {}
    o_det, o_se, o_lane, o_pc = OF.achelous_forward(sd, x, xr, pc, phi=phi, backbone=backbone, taps=taps, neck=neck)
    for name in eng.taps:
        assert rel_err(eng.tap(name), taps[name]) < TOL, name
    for i in range(3):
        assert rel_err(det[i], o_det[i]) < TOL
    assert rel_err(se, o_se) < TOL and rel_err(lane, o_lane) < TOL and rel_err(pcs, o_pc) < TOL
    for logits, ref in ((se, o_se), (lane, o_lane)):
        _, n_safe_diff, _ = argmax_mismatch(logits, ref.argmax(1))
        assert n_safe_diff == 0


def test_pn2_plan_matches_builder_oracle():
    """Config 4 (EN-GDF-PN2-S2): PointNet++ is builder-defined (oracle/pn2.py) - parity UNPINNED against the reference,
    which ships no PN2 code.  FPS / ball-query indices must be bit-identical, features within tolerance."""
    torch.set_num_threads(4)
    kw = dict(MODEL_KW, pc_seg="pn2")
    model = Achelous(phi="S2", backbone="en", **kw).eval()
    sd = fill_state_dict(model.state_dict(), seed=3)
    model.load_state_dict(sd, strict=True)
    B = 2
    x, xr, pc = make_inputs(B, seed=22)
    eng = Engine(model, B, "cpu", dry_run=True)
    ins = eng.input_tensors()
    ins[0].copy_(x), ins[1].copy_(xr), ins[2].copy_(pc)
    emulate_engine(eng)
    taps = {}
    o_det, o_se, o_lane, o_pc = OF.achelous_forward(sd, x, xr, pc, phi="S2", backbone="en", pc_seg="pn2", taps=taps)
    for k, v in eng.taps_int.items():
        assert torch.equal(v.long(), taps[k]), k
    for name in ("pc.sa1.out", "pc.sa2.out", "pc.sa3.out", "pc.fp3", "pc.fp2", "pc.fp1"):
        assert rel_err(eng.tap(name), taps[name]) < TOL, name
    assert rel_err(eng.output_views()[3], o_pc) < TOL


def test_repack_tracks_parameter_updates():
    model = Achelous(phi="S0", backbone="en", **MODEL_KW).eval()
    model.load_state_dict(fill_state_dict(model.state_dict(), seed=1))
    eng = Engine(model, 1, "cpu", dry_run=True)
    sig = eng._sig
    key = "det0.stem.wt"
    before = eng._weights[key][0].clone()
    with torch.no_grad():
        model.det_head.stems[0].conv.weight.mul_(2.0)
    assert eng._signature() != sig
    eng.repack()
    assert torch.allclose(eng._weights[key][0], before * 2.0)


@pytest.mark.parametrize("image_shape,keep,fuse", [(None, None, True), (None, (0, 8), False), ((97, 211), (0, 8), True)])
def test_compact_plan_matches_raw_plan(image_shape, keep, fuse):
    """outputs="compact": decode + NMS rows, uint8 class maps and point classes produced INSIDE the plan equal what the
    reference's caller derives from the raw outputs (achelous.py:259-297: decode_outputs -> non_max_suppression, softmax ->
    [crop -> cv2.resize] -> argmax -> class-mask rule, argmax over point classes)."""
    import numpy as np
    from oracle import postprocess as OP
    torch.set_num_threads(4)
    model = Achelous(phi="S0", backbone="en", **MODEL_KW).eval()
    model.fuse_seg_decoder = fuse
    model.load_state_dict(fill_state_dict(model.state_dict(), seed=2), strict=True)
    with torch.no_grad():                                  # bias the obj logits so that a useful number of boxes survive
        for k in range(3):
            model.det_head.obj_preds[k].bias += 2.0
            model.det_head.reg_preds[k].bias[2:] += 1.3
    B = 2
    x, xr, pc = make_inputs(B, seed=21)
    outs = {}
    spec = dict(conf_thres=0.3, nms_thres=0.45, max_det=40, keep_classes=keep, image_shape=image_shape)
    for mode, compact in (("raw", None), ("compact", spec)):
        eng = Engine(model, B, "cpu", dry_run=True, compact=compact)
        ins = eng.input_tensors()
        ins[0].copy_(x), ins[1].copy_(xr), ins[2].copy_(pc)
        emulate_engine(eng)
        outs[mode] = eng.output_views()
        if mode == "compact":
            names = " ".join(fn.__name__ for fn, _ in eng.ops)
            assert ("ach_up_ghost_head_argmax" in names) == (fuse and image_shape is None)     # fused argmax head only at network resolution
            assert ("out.se_logits" in eng._bufs) == (not fuse or image_shape is not None)
    det, se, lane, pcs = outs["raw"]
    c = outs["compact"]
    assert torch.equal(c.pc_cls.long(), pcs.argmax(-1))
    keep_lut = torch.ones(9, dtype=torch.bool) if keep is None else torch.tensor([k in keep for k in range(9)])
    if image_shape is None:
        am = se.argmax(1)
        assert torch.equal(c.se_mask.long(), torch.where(keep_lut[am], am, torch.zeros_like(am)))
        assert torch.equal(c.lane_mask.long(), lane.argmax(1))
    else:
        for b in range(B):
            _, am = OP.seg_postprocess(se[b], image_shape, True)
            am = torch.from_numpy(am)
            assert torch.equal(c.se_mask[b].long(), torch.where(keep_lut[am], am, torch.zeros_like(am)))
            assert torch.equal(c.lane_mask[b].long(), torch.from_numpy(OP.seg_postprocess(lane[b], image_shape, True)[1]))
    decoded = OP.decode_outputs([d.clone() for d in det], (320, 320))
    ref_rows, _ = OP.non_max_suppression(decoded, 7, (320, 320), np.array([320, 320]), False, conf_thres=0.3, nms_thres=0.45,
                                         return_indices=True)
    n_trunc = 0
    for b in range(B):
        n = ref_rows[b].shape[0]
        assert int(c.det_count[b]) == n and n > 0
        m = min(n, 40)
        n_trunc += n > 40
        # rows before the letterbox un-warp: (x1, y1, x2, y2) normalised; the oracle returns them un-warped to a 320x320 image, y first
        mine = c.det_rows[b, :m].numpy()
        ref = ref_rows[b][:m]
        assert np.array_equal(mine[:, 4:], ref[:, 4:])
        assert np.allclose(mine[:, [1, 0, 3, 2]] * 320.0, ref[:, :4], rtol=0, atol=1e-3)
        assert not c.det_rows[b, m:].any()
    assert n_trunc > 0      # the cap was exercised


def test_grouped_mlp_plan_equals_whole_batch_plan():
    """The LN -> Linear(4C) -> GELU -> Linear pairs run per group of frames through one small (L2-resident) hidden buffer when the
    whole-batch hidden tensor is large; same arithmetic per frame, so the outputs must be IDENTICAL to the ungrouped plan."""
    torch.set_num_threads(4)
    model = Achelous(phi="S0", backbone="en", **MODEL_KW).eval()
    model.load_state_dict(fill_state_dict(model.state_dict(), seed=2), strict=True)
    model.fuse_mlp = False      # the two-launch path (the fused kernel needs no hidden buffer at all)
    B = 3
    x, xr, pc = make_inputs(B, seed=23)
    outs = []
    for group_bytes in (0, 1):
        model.mlp_group_bytes, model.mlp_group_min_tiles = group_bytes, 1       # 1 byte: one frame per group
        eng = Engine(model, B, "cpu", dry_run=True)
        if group_bytes:
            assert sum(n.startswith("bb.s0.0.pw1") for n in eng.op_names) == B and eng._bufs["bb.s0.0.h"].shape[0] == 1
        ins = eng.input_tensors()
        ins[0].copy_(x), ins[1].copy_(xr), ins[2].copy_(pc)
        emulate_engine(eng)
        outs.append(eng.packed_out.clone())
    assert torch.equal(outs[0], outs[1])


def test_fused_mlp_plan_is_used_and_matches_two_launch_plan():
    """ach_mlp_tc replaces the pw1 / pw2 pair wherever the width is instantiated (EN-S0: stages 0-2); through the CPU emulator both
    plans evaluate the same fp64 arithmetic, so the packed weights (32-column W1 tiles, one C-column W2 tile) are what is pinned here."""
    torch.set_num_threads(4)
    model = Achelous(phi="S0", backbone="en", **MODEL_KW).eval()
    model.load_state_dict(fill_state_dict(model.state_dict(), seed=5), strict=True)
    B = 2
    x, xr, pc = make_inputs(B, seed=29)
    outs = []
    for fuse in (True, False):
        model.fuse_mlp = fuse
        eng = Engine(model, B, "cpu", dry_run=True)
        n_mlp = sum(n.endswith(".mlp") for n in eng.op_names)
        n_pw1 = sum(n.endswith(".pw1") for n in eng.op_names)
        assert (n_mlp, n_pw1) == ((10, 2) if fuse else (0, 12))     # C = 176 (stage 3) keeps the two GEMM launches
        ins = eng.input_tensors()
        ins[0].copy_(x), ins[1].copy_(xr), ins[2].copy_(pc)
        emulate_engine(eng)
        outs.append(eng.packed_out.clone())
    assert rel_err(outs[0], outs[1]) < 1e-6


# ------------------------------------------------------------------ folded RCBlock operands (rcblock_tc.cu v5)
class _Arena:
    def __init__(self):
        self.t = {}

    def new(self, name, tensor):
        self.t[name] = tensor.detach().to(torch.float32).clone().contiguous()
        return self.t[name]

    def ptr(self, name):
        return self.t[name].data_ptr()


R = torch.randn


@pytest.mark.parametrize("Cc,H,W", [(3, 20, 24), (16, 12, 12)])
def test_rc_deform_tc_emulator_matches_unfolded_block(Cc, H, W):
    """the host fold (engine.rc_tc_fold) + the emulator's folded ach_rc_deform_tc == the unfolded RCBlock contract (ach_rc_deform)"""
    from achelous_b200.engine import rc_tc_fold
    torch.manual_seed(3)
    B, CP = 2, (Cc + 3) // 4 * 4
    A = _Arena()
    A.new("x", R(B, Cc, H, W))
    pc = torch.zeros(B, H * W, CP)
    pc[:, :, :Cc] = R(B, H * W, Cc)
    A.new("pooled", pc)
    w_offmod, b_offmod = R(27, Cc, 3, 3) / (Cc * 9) ** 0.5, torch.rand(27) * 2 - 1
    w_reg, w1, scale, bias = R(Cc, Cc, 3, 3) / (Cc * 9) ** 0.5, R(Cc, Cc) / Cc ** 0.5, torch.rand(Cc) + 0.5, R(Cc) * 0.1
    outs = []
    for folded in (False, True):
        A.new("out", torch.zeros(B, Cc, H, W))
        s = AchRcDeform()
        s.x, s.pooled, s.out = A.ptr("x"), A.ptr("pooled"), A.ptr("out")
        s.x_bs = s.out_bs = Cc * H * W
        s.pooled_cl, s.pooled_bs = 1, CP * H * W
        s.B, s.C, s.H, s.W = B, Cc, H, W
        if not folded:
            wom = torch.zeros(Cc * 9, 28)
            wom[:, :27] = w_offmod.reshape(27, Cc * 9).t()
            A.new("w_om", wom), A.new("b_om", b_offmod), A.new("w_reg", w_reg.reshape(Cc, Cc * 9).t()), A.new("w1", w1.t())
            A.new("scale", scale), A.new("bias", bias)
            for n in ("w_om", "b_om", "w_reg", "w1", "scale", "bias"):
                setattr(s, n, A.ptr(n))
            emu.ach_rc_deform(s)
        else:
            om, consts, reg, bias_f, K = rc_tc_fold(w_offmod.double(), b_offmod.double(), w_reg.double(), w1.double(), scale.double(), bias.double())
            n_om, n_reg = _lib.load().ach_pack_pw_tc_elems(K, 27), _lib.load().ach_pack_pw_tc_elems(K, Cc)
            A.new("w_om_tap", om), A.new("w_reg_tap", reg), A.new("consts", consts), A.new("bias_f", bias_f)
            for n_, sz in (("omh", n_om), ("oml", n_om), ("rgh", n_reg), ("rgl", n_reg)):
                A.new(n_, torch.zeros(sz))
            s.b_om, s.bias = A.ptr("consts"), A.ptr("bias_f")
            emu.ach_pack_pw_tc(A.ptr("w_om_tap"), K, 27, 28, A.ptr("omh"), A.ptr("oml"))
            emu.ach_pack_pw_tc(A.ptr("w_reg_tap"), K, Cc, CP, A.ptr("rgh"), A.ptr("rgl"))
            emu.ach_rc_deform_tc(s, A.ptr("omh"), A.ptr("oml"), A.ptr("rgh"), A.ptr("rgl"))
        outs.append(A.t["out"].clone())
    assert (outs[0] - outs[1]).abs().max().item() <= 2e-5 * outs[0].abs().max().item()
