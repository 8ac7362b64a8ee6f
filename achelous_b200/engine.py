"""Execution engine: turns an :class:`achelous_b200.nets.Achelous.Achelous` parameter tree into
(a) packed device weights (BN / LayerNorm folding, K-major GEMM layouts) and (b) a static launch
plan of C-ABI kernel calls over pre-allocated HBM buffers, replayed per forward (optionally as one
CUDA graph).  PyTorch is used only for device memory, streams and the weight folding arithmetic;
every activation of the forward pass is produced by the sm_100a kernels in ``csrc/``.

Data layout in HBM: fp32, channel-major planes exactly as torch NCHW / (B, C, N); a "view" is
(pointer, batch stride), so channel slices and concatenations cost nothing.  All six outputs of a
frame live in ONE packed buffer ``[det40 | det20 | det10 | se_seg | lane_seg | pc_seg]``
(4 622 784 B / frame) so that the multi-GPU path is a single all-gather.
"""
import ctypes as C
import math
import os
from collections import namedtuple

import torch

from . import _lib
from ._lib import (ACT_GELU, ACT_NONE, ACT_RELU, ACT_SILU, AchConv3x3Tc, AchConvDense, AchDwConv, AchMlp, AchPwConv, AchRcDeform, AchUpGhost,
                   AchUpGhostHead, AchUpGhostPw2)
from .nets import holders as Hd

View = namedtuple("View", "ptr bs C H W")

# what the compact output record holds per frame (forward(..., outputs="compact"), stream_forward(compact=True)):
#   det_rows (B, max_det, 7) fp32 [x1, y1, x2, y2, obj, cls_conf, cls] in network-input-normalised corner form = the rows
#            utils_bbox.non_max_suppression keeps (score-descending) BEFORE its host-side letterbox un-warp; rows >= det_count are 0
#   det_count (B,) int32  survivors of the NMS (if > max_det the record was truncated to the max_det best)
#   se_mask / lane_mask (B, H, W) uint8  argmax class maps (network resolution, or the original image size when image_shape is set)
#   pc_cls (B, N) uint8  argmax class per radar point
CompactOutputs = namedtuple("CompactOutputs", "det_rows det_count se_mask lane_mask pc_cls")
COMPACT_DEFAULTS = dict(conf_thres=0.35, nms_thres=0.35,   # achelous.py:52,56
                        max_det=256, keep_classes=None, image_shape=None, letterbox_image=True)


def compact_spec(spec=None, **kw):
    """Normalised, hashable compact-output specification"""
    d = dict(COMPACT_DEFAULTS)
    d.update(spec or {})
    d.update(kw)
    unknown = set(d) - set(COMPACT_DEFAULTS)
    if unknown:
        raise TypeError(f"unknown compact-output options: {sorted(unknown)}")
    if d["keep_classes"] is not None:
        d["keep_classes"] = tuple(sorted(int(c) for c in d["keep_classes"]))
    if d["image_shape"] is not None:
        d["image_shape"] = (int(d["image_shape"][0]), int(d["image_shape"][1]))
    d["conf_thres"], d["nms_thres"], d["max_det"] = float(d["conf_thres"]), float(d["nms_thres"]), int(d["max_det"])
    d["letterbox_image"] = bool(d["letterbox_image"])
    return tuple(sorted(d.items()))


def _align16(n):
    return (n + 15) // 16 * 16


class _LazyOut:
    """Full-resolution logit map that is only allocated when a kernel really has to write it (compact mode: the fused head
    writes class indices straight away and the fp32 planes never exist)."""

    def __init__(self, eng, name, C_, H, W):
        self.eng, self.name, self.C, self.H, self.W, self._v = eng, name, C_, H, W, None

    def view(self):
        if self._v is None:
            self._v = self.eng.buf(self.name, self.C, self.H, self.W)
        return self._v

    @property
    def allocated(self):
        return self._v is not None


def _ceil4(n):
    return (n + 3) // 4 * 4


def rc_tc_push_k(C_):
    """k handed to the tensor core per push by ach_rc_deform_tc (rcblock_tc.cu): 32 when a whole GEMM fits one push"""
    return 32 if 9 * C_ <= 32 else 16


def rc_tc_fold(w_offmod, b_offmod, w_reg, w1, scale, bias):
    """Operands of ach_rc_deform_tc (include/achelous_b200.h) from the RCBlock's parameters, host float64.
    Everything after the bilinear sampling of RadarEncoder.py:65-72 / dcn.py:49-63 is linear up to the ReLU -
    deform_conv2d (no bias) -> weight_conv1 -> eval BatchNorm - so it collapses into ONE (9C x C) matrix
    2 * diag(scale) . W1 . Wreg (the 2 of `2 * sigmoid`) and the bias; the tap coordinate (t/3 - 1, t%3 - 1) joins the offset
    bias, and the modulator rows are pre-scaled by -log2(e) (sigmoid on ex2).  Where the push has a spare k column
    (9C % push_k != 0) the constants ride in the GEMMs as an extra weight row against a constant-1 operand column.
      w_offmod (27, C, 3, 3) = cat(offset_conv.weight, modulator_conv.weight), b_offmod (27,), w_reg (C, C, 3, 3), w1 (C, C),
      scale / bias: folded BatchNorm(weight_conv1 + its bias)
    -> w_om_tap (K, 28), consts (32,), w_reg_tap (K, ceil4(C)), bias (C,), K (= 9C + 1 with the constants row, else 9C)"""
    C_ = w_reg.shape[0]
    log2e = 1.4426950408889634
    consts = torch.zeros(32, dtype=torch.float64)
    t = torch.arange(9)
    consts[0:18:2] = b_offmod[0:18:2] + (t // 3 - 1)
    consts[1:18:2] = b_offmod[1:18:2] + (t % 3 - 1)
    consts[18:27] = -log2e * b_offmod[18:27]
    om = w_offmod.reshape(27, C_, 9).permute(2, 1, 0).reshape(9 * C_, 27).clone()      # rows k = tap * C + ch
    om[:, 18:] *= -log2e
    m = 2.0 * scale[:, None] * (w1.reshape(C_, C_) @ w_reg.reshape(C_, C_ * 9))            # (o, ch * 9 + tap)
    reg = m.reshape(C_, C_, 9).permute(2, 1, 0).reshape(9 * C_, C_)
    K = 9 * C_
    if K % rc_tc_push_k(C_) != 0:
        om = torch.cat([om, consts[None, :27]], 0)
        reg = torch.cat([reg, bias[None, :].to(reg.dtype)], 0)
        K += 1
    om = torch.nn.functional.pad(om, (0, 1))
    reg = torch.nn.functional.pad(reg, (0, _ceil4(C_) - C_))
    return om, consts, reg, bias, K


class Engine:
    def __init__(self, model, batch, device, use_graph=True, dry_run=False, n_points=None, compact=None, out=None):
        """dry_run=True builds the plan and packs the weights without a GPU (host-logic tests only):
        nothing can be launched from such an engine.  `model` is the module that OWNS the parameters (under
        nn.DataParallel: the wrapped module, not a replica); `device` may differ from the parameters' device -
        the folded weights are produced on the host and copied to `device`."""
        self.lib = _lib.load()
        self.model = model
        self.B = int(batch)
        self.device = torch.device(device)
        self.dry_run = dry_run
        if self.device.type != "cuda" and not dry_run:
            raise _lib.AchelousKernelError("achelous_b200 runs on CUDA devices only (no CPU fallback)")
        self.use_graph = use_graph
        self.res = model.resolution
        self.n_points = int(model.n_points if n_points is None else n_points)
        self.compact = dict(compact_spec(compact) if isinstance(compact, dict) else compact) if compact else None
        self._out = out        # caller-owned packed output rows (e.g. this rank's slice of an all-gather buffer: in-place gather)
        self.ops = []          # (cfunc, args) - stream appended at call time
        self.op_names = []
        self.op_bytes = []
        self.op_lane = []      # stream lane of every launch (0 = caller's stream)
        self.sync_before = {}  # op index -> [(waiter_lane, waited_lane)] cross-lane dependencies
        self.sync_end = []
        self.cur_lane = 0
        self.in_pack = False
        self.multi_stream = True
        self._side_streams = {}
        self._keep = []        # ctypes structs / tensors that must outlive the plan
        self._bufs = {}
        self._weights = {}
        self._host = {}        # host-side weight arrays that travel as kernel parameters
        self.graph = None
        self._sig = None
        if dry_run:
            self._build()
        else:
            with torch.cuda.device(self.device):
                self._build()

    # ------------------------------------------------------------------ parameter access / packing
    def _resolve_params(self):
        """Looks every parameter / buffer up BY NAME on the owning module (load_state_dict(assign=True) and
        `module.weight = nn.Parameter(...)` replace the objects) and remembers the module's weight epoch."""
        m = self.model
        self._params = {k: v for k, v in list(m.named_parameters()) + list(m.named_buffers())}
        self._epoch = getattr(m, "_weights_epoch", 0)
        self._cpu64 = {}

    def _signature(self):
        return tuple((p.data_ptr(), p._version) for p in self._params.values())

    def _p(self, name):
        """Parameter as float64 ON THE HOST: BN / LayerNorm folding and layout packing are host arithmetic (one
        device->host copy per parameter and (re)pack), so no ATen kernel runs on the GPU for them."""
        t = self._cpu64.get(name)
        if t is None:
            t = self._params[name].detach().to(device="cpu", dtype=torch.float64)
            self._cpu64[name] = t
        return t

    def _pi(self, name):
        """Integer buffer (index tables) on the host"""
        return self._params[name].detach().cpu()

    def _zeros(self, n):
        """Zero-filled fp32 device slot carved out of a few large arenas (one fill launch per arena, not per weight tile)"""
        n = (int(n) + 63) // 64 * 64
        a = getattr(self, "_arena", None)
        if a is None or a[1] + n > a[0].numel():
            a = [torch.zeros(max(n, 1 << 22), device=self.device, dtype=torch.float32), 0]
            self._arena = a
            self._keep.append(a[0])
        t = a[0][a[1]:a[1] + n]
        a[1] += n
        return t

    def _dev(self, t):
        t = t.to(device=self.device, dtype=torch.float32).contiguous()
        self._keep.append(t)
        return t

    def _w(self, key, fn):
        """Packed-weight slot: allocated once, refilled in place by repack()."""
        val = fn()
        val = val.to(device=self.device, dtype=torch.float32).contiguous()
        if key in self._weights:
            self._weights[key][0].copy_(val)
        else:
            self._weights[key] = (val, fn)
        return self._weights[key][0]

    def _h(self, key, fn):
        """Host float array slot (for weights passed as kernel parameters); refreshed by repack()."""
        val = fn().to(dtype=torch.float32, device="cpu").contiguous().flatten()
        arr = (C.c_float * val.numel())()
        C.memmove(arr, val.data_ptr(), val.numel() * 4)
        self._host[key] = (arr, fn)
        return C.addressof(arr)

    def repack(self):
        self._cpu64 = {}
        for key, (t, fn) in self._weights.items():
            t.copy_(fn().to(device=self.device, dtype=torch.float32))
        for key, (arr, fn) in self._host.items():
            val = fn().to(dtype=torch.float32, device="cpu").contiguous().flatten()
            C.memmove(arr, val.data_ptr(), val.numel() * 4)
        if self._host:
            self.graph = None  # kernel-parameter weights are baked into a captured graph: re-capture
        self._cpu64 = {}
        self._sig = self._signature()

    def _bn_fold(self, prefix, eps, conv_bias=None):
        """eval BatchNorm(conv + conv_bias) == scale * conv + bias"""
        s = self._p(prefix + ".weight") / torch.sqrt(self._p(prefix + ".running_var") + eps)
        b = self._p(prefix + ".bias") - self._p(prefix + ".running_mean") * s
        if conv_bias is not None:
            b = b + conv_bias * s
        return s, b

    @staticmethod
    def _kmajor(w2d):
        """(O, K) -> K-major (K, ceil4(O)) zero padded"""
        O, K = w2d.shape
        out = torch.zeros(K, _ceil4(O), dtype=w2d.dtype, device=w2d.device)
        out[:, :O] = w2d.t()
        return out

    # ------------------------------------------------------------------ buffers / views
    def buf(self, name, C_, H, W=1, fill=None):
        t = torch.empty(self.B, C_, H, W, device=self.device, dtype=torch.float32)
        if fill is not None:
            t.fill_(fill)
        assert name not in self._bufs, name
        self._bufs[name] = t
        return View(t.data_ptr(), t.stride(0), C_, H, W)

    @staticmethod
    def sl(v, c0, c1):
        return View(v.ptr + c0 * v.H * v.W * 4, v.bs, c1 - c0, v.H, v.W)

    def _ptr(self, t):
        return None if t is None else t.data_ptr()

    def _add(self, name, fn, *args, nbytes=None):
        self.ops.append((fn, args))
        self.op_names.append(name)
        self.op_bytes.append(nbytes)
        self.op_lane.append(self.cur_lane)

    def wait(self, waiter, waited):
        """Lane `waiter` waits for everything recorded so far on lane `waited` (before the next launch)."""
        self.sync_before.setdefault(len(self.ops), []).append((waiter, waited))

    def algorithmic_bytes(self, i):
        """Unique input + output + weight bytes of launch i (fp32 storage), or None when not modelled."""
        return self.op_bytes[i]

    # ------------------------------------------------------------------ op recorders
    def pw(self, name, x0, out, wt, O, x1=None, scale=None, bias=None, pbias=None, res=None, gamma=None, ln=False,
           ln_eps=1e-6, act=ACT_NONE, wt_bs=0, ldw=None, reduce_max=False, out_bs=None, frames=None):
        """frames = (b0, nb): the launch covers frames [b0, b0 + nb) of every view only (shared weights, no reduce_max / pbias)"""
        P = x0.H * x0.W
        B_ = self.B
        if frames is not None:
            b0, B_ = frames
            assert wt_bs == 0 and not reduce_max and pbias is None and x1 is None
            shift = lambda v: None if v is None else v._replace(ptr=v.ptr + b0 * v.bs * 4)
            x0, out, res = shift(x0), shift(out), shift(res)
        s = AchPwConv()
        s.x0, s.x0_bs, s.c0 = x0.ptr, x0.bs, x0.C
        s.x1, s.x1_bs, s.c1 = (x1.ptr, x1.bs, x1.C) if x1 is not None else (None, 0, 0)
        if isinstance(wt, torch.Tensor):
            s.ldw = wt.shape[-1] if ldw is None else ldw
            s.wt = wt.data_ptr()
        else:
            s.wt, s.ldw = wt, ldw
        s.wt_bs = wt_bs
        s.scale, s.bias, s.pbias, s.gamma = self._ptr(scale), self._ptr(bias), self._ptr(pbias), self._ptr(gamma)
        s.res, s.res_bs = (res.ptr, res.bs) if res is not None else (None, 0)
        if reduce_max:
            s.out, s.out_bs = out.data_ptr(), out.stride(0)
        else:
            assert out.C == O and out.H * out.W == P, (name, out, O, P)
            s.out, s.out_bs = out.ptr, out.bs
        s.B, s.O, s.P = B_, O, P
        s.ln, s.ln_eps, s.act, s.reduce_max = int(ln), ln_eps, act, int(reduce_max)
        self._keep.append(s)
        K_ = s.c0 + s.c1
        nb = 4 * (B_ * K_ * P + (B_ * O if reduce_max else B_ * O * P) + (B_ * O * P if res is not None else 0)
                  + K_ * s.ldw * (B_ if wt_bs else 1))
        tc_mode = self.model.use_tensor_cores
        # measured on B200: the warp-specialised tcgen05 kernel beats (or ties) the SIMT GEMM on every shared-weight layer with
        # >= 16 outputs (32 -> 16 at 320^2: 0.51 -> 0.47 ms); below that the two are within noise of each other
        tc_ok = tc_mode in (True, "all") and O >= 16
        if tc_ok and wt_bs == 0 and isinstance(wt, torch.Tensor) and not self.in_pack:
            # tcgen05 path: weights re-packed on the device into hi/lo UMMA tile images whenever they change
            ck = (wt.data_ptr(), K_, O, s.ldw, bool(ln))          # launches that share a weight tensor share its packed tiles
            cache = self.__dict__.setdefault("_tc_tiles", {})
            if ck not in cache:
                n = self.lib.ach_pack_pw_tc_elems(K_, O)
                hi = self._zeros(n)
                lo = self._zeros(n)
                self._keep += [hi, lo]
                self.pack_ops.append((self.lib.ach_pack_pw_tc, (wt.data_ptr(), K_, O, s.ldw, hi.data_ptr(), lo.data_ptr())))
                wsum = self._w(name + ".wsum", lambda wt=wt, O=O: wt.detach().cpu()[:, :O].double().sum(0)) if ln else None
                cache[ck] = (hi, lo, wsum)
            hi, lo, wsum = cache[ck]
            self._add(name, self.lib.ach_pw_conv_tc, C.byref(s), hi.data_ptr(), lo.data_ptr(), self._ptr(wsum), nbytes=nb)
            return
        self._add(name, self.lib.ach_pw_conv, C.byref(s), nbytes=nb)

    def dw(self, name, x, out, w, k, stride=1, scale=None, bias=None, act=ACT_NONE, xadd=None, post=None):
        s = AchDwConv()
        s.x, s.x_bs = x.ptr, x.bs
        s.xadd, s.xadd_bs = (xadd.ptr, xadd.bs) if xadd is not None else (None, 0)
        s.w, s.scale, s.bias, s.post = w.data_ptr(), self._ptr(scale), self._ptr(bias), post
        s.out, s.out_bs = out.ptr, out.bs
        s.B, s.C, s.H, s.W, s.Ho, s.Wo, s.k, s.stride, s.act = self.B, x.C, x.H, x.W, out.H, out.W, k, stride, act
        assert out.C == x.C, (name, x, out)
        self._keep.append(s)
        nb = 4 * self.B * x.C * (x.H * x.W * (2 if xadd is not None else 1) + out.H * out.W)
        self._add(name, self.lib.ach_dw_conv, C.byref(s), nbytes=nb)

    def conv(self, name, x, out, w, k, stride, pad, scale=None, bias=None, act=ACT_NONE, ln_w=None, ln_b=None, ln_eps=1e-6):
        s = AchConvDense()
        s.x, s.x_bs, s.w = x.ptr, x.bs, w.data_ptr()
        s.scale, s.bias, s.ln_w, s.ln_b = self._ptr(scale), self._ptr(bias), self._ptr(ln_w), self._ptr(ln_b)
        s.out, s.out_bs = out.ptr, out.bs
        s.B, s.Cin, s.H, s.W, s.O, s.ldo = self.B, x.C, x.H, x.W, out.C, w.shape[-1]
        s.Ho, s.Wo, s.k, s.stride, s.pad, s.act = out.H, out.W, k, stride, pad, act
        s.ln_out, s.ln_eps = int(ln_w is not None), ln_eps
        self._keep.append(s)
        nb = 4 * (self.B * (x.C * x.H * x.W + out.C * out.H * out.W) + w.numel())
        self._add(name, self.lib.ach_conv_dense, C.byref(s), nbytes=nb)

    def conv3_tc(self, name, x, out, wname, scale=None, bias=None, act=ACT_NONE, res=None):
        """3x3 / stride 1 / pad 1 dense conv as an implicit GEMM on tcgen05 (conv3x3_tc.cu); `res` is added after the activation"""
        cin, O = x.C, out.C
        kpad = self.lib.ach_conv3x3_tc_k(cin)

        def wt(wname=wname, cin=cin, O=O, kpad=kpad):   # rows k = (g*9 + tap)*16 + c for channel 16g + c
            w = self._p(wname)
            wp = torch.zeros(O, kpad // 9, 9, dtype=w.dtype, device=w.device)
            wp[:, :cin] = w.reshape(O, cin, 9)
            m = wp.reshape(O, kpad // 144, 16, 9).permute(1, 3, 2, 0).reshape(kpad, O)
            return torch.nn.functional.pad(m, (0, _ceil4(O) - O))
        w_t = self._w(name + ".wt3", wt)
        n_ = self.lib.ach_pack_pw_tc_elems(kpad, O)
        hi = self._zeros(n_)
        lo = self._zeros(n_)
        self._keep += [hi, lo]
        self.pack_ops.append((self.lib.ach_pack_pw_tc, (w_t.data_ptr(), kpad, O, w_t.shape[-1], hi.data_ptr(), lo.data_ptr())))
        s = AchConv3x3Tc()
        s.x, s.x_bs, s.out, s.out_bs = x.ptr, x.bs, out.ptr, out.bs
        s.scale, s.bias = self._ptr(scale), self._ptr(bias)
        if res is not None:
            s.res, s.res_bs = res.ptr, res.bs
        s.B, s.Cin, s.H, s.W, s.O, s.act = self.B, cin, x.H, x.W, O, act
        self._keep.append(s)
        nb = 4 * (self.B * x.H * x.W * (cin + O * (2 if res is not None else 1)) + kpad * O)
        self._add(name, self.lib.ach_conv3x3_tc, C.byref(s), hi.data_ptr(), lo.data_ptr(), nbytes=nb)

    def _pack_conv(self, key, wname):
        """(O, Cin, k, k) -> [Cin][k*k][ceil4(O)]"""
        def f():
            w = self._p(wname)
            O, Cin, kh, kw = w.shape
            out = torch.zeros(Cin, kh * kw, _ceil4(O), dtype=w.dtype, device=w.device)
            out[:, :, :O] = w.permute(1, 2, 3, 0).reshape(Cin, kh * kw, O)
            return out
        return self._w(key, f)

    def _vec(self, key, fn):
        return self._w(key, fn)

    # ------------------------------------------------------------------ composite blocks
    def pw_bn_act(self, name, prefix_conv, prefix_bn, eps, x0, out, act, x1=None, conv_bias=False, res=None, O=None):
        """1x1 conv (+bias) + BN + act"""
        wt = self._w(name + ".wt", lambda: self._kmajor(self._p(prefix_conv + ".weight").flatten(1)))
        sc = self._vec(name + ".s", lambda: self._bn_fold(prefix_bn, eps, self._p(prefix_conv + ".bias") if conv_bias else None)[0])
        bi = self._vec(name + ".b", lambda: self._bn_fold(prefix_bn, eps, self._p(prefix_conv + ".bias") if conv_bias else None)[1])
        self.pw(name, x0, out, wt, out.C if O is None else O, x1=x1, scale=sc, bias=bi, act=act, res=res)

    def pw_bias(self, name, prefix, x0, out, act=ACT_NONE):
        wt = self._w(name + ".wt", lambda: self._kmajor(self._p(prefix + ".weight").flatten(1)))
        bi = self._vec(name + ".b", lambda: self._p(prefix + ".bias"))
        self.pw(name, x0, out, wt, out.C, bias=bi, act=act)

    def dw_bn_act(self, name, prefix_conv, prefix_bn, eps, x, out, k, act, c0=0, c1=None, stride=1):
        """depthwise conv (channel slice [c0, c1) of the parameter tensors) + BN + act"""
        c1 = x.C + c0 if c1 is None else c1
        w = self._w(name + ".w", lambda: self._p(prefix_conv + ".weight")[c0:c1].flatten(1))
        sc = self._vec(name + ".s", lambda: self._bn_fold(prefix_bn, eps)[0][c0:c1])
        bi = self._vec(name + ".b", lambda: self._bn_fold(prefix_bn, eps)[1][c0:c1])
        self.dw(name, x, out, w, k, stride=stride, scale=sc, bias=bi, act=act)

    def ghost(self, name, prefix, x0, out, relu, x1=None):
        """GhostModule writing straight into `out` (oup = out.C channels)"""
        oup = out.C
        init = math.ceil(oup / 2)
        act = ACT_RELU if relu else ACT_NONE
        prim = self.sl(out, 0, init)
        self.pw_bn_act(name + ".primary", prefix + ".primary_conv.0", prefix + ".primary_conv.1", 1e-5, x0, prim, act, x1=x1)
        if oup > init:
            src = self.sl(out, 0, oup - init)
            self.dw_bn_act(name + ".cheap", prefix + ".cheap_operation.0", prefix + ".cheap_operation.1", 1e-5, src,
                           self.sl(out, init, oup), 3, act, c0=0, c1=oup - init)

    def upsample_block(self, name, prefix, x, out):
        """Upsample: BaseConv 1x1 relu -> bilinear x2 into `out`"""
        t = self.buf(name + ".pw", out.C, x.H, x.W)
        self.pw_bn_act(name + ".conv", prefix + ".upsample.0.conv", prefix + ".upsample.0.bn", 1e-3, x, t, ACT_RELU)
        self._add(name + ".up", self.lib.ach_upsample2x, t.ptr, t.bs, out.ptr, out.bs, self.B, t.C, t.H, t.W,
                  nbytes=4 * self.B * t.C * t.H * t.W * 5)

    def ghost_bottleneck(self, name, prefix, xa, xb, mid, outc):
        g1 = self.buf(name + ".g1", mid, xa.H, xa.W)
        self.ghost(name + ".ghost1", prefix + ".ghost1", xa, g1, True, x1=xb)
        g2 = self.buf(name + ".g2", outc, xa.H, xa.W)
        self.ghost(name + ".ghost2", prefix + ".ghost2", g1, g2, False)
        cin = xa.C + xb.C
        sc = self.buf(name + ".sc", cin, xa.H, xa.W)
        self.dw_bn_act(name + ".sc_dw_a", prefix + ".shortcut.0", prefix + ".shortcut.1", 1e-5, xa, self.sl(sc, 0, xa.C), 3,
                       ACT_NONE, c0=0, c1=xa.C)
        self.dw_bn_act(name + ".sc_dw_b", prefix + ".shortcut.0", prefix + ".shortcut.1", 1e-5, xb, self.sl(sc, xa.C, cin), 3,
                       ACT_NONE, c0=xa.C, c1=cin)
        out = self.buf(name + ".out", outc, xa.H, xa.W)
        self.pw_bn_act(name + ".sc_pw", prefix + ".shortcut.2", prefix + ".shortcut.3", 1e-5, sc, out, ACT_NONE, res=g2)
        return out

    # ---- EdgeNeXt
    def _ln_fold_linear(self, name, lin, ln):
        """Linear(LayerNorm_affine(x)) == (W * ln_w) x_hat + (b + W ln_b)"""
        wt = self._w(name + ".wt", lambda: self._kmajor(self._p(lin + ".weight") * self._p(ln + ".weight")[None, :]))
        bi = self._vec(name + ".b", lambda: self._p(lin + ".bias") + self._p(lin + ".weight") @ self._p(ln + ".bias"))
        return wt, bi

    def mlp_res(self, name, prefix, x_ln_src, res, out):
        """LN -> Linear(4C) -> GELU -> Linear(C) -> gamma -> + res   (conv_encoder.py:23-31, sdta_encoder.py:64-73)"""
        Cc = x_ln_src.C
        P = x_ln_src.H * x_ln_src.W
        wt1, b1 = self._ln_fold_linear(name + ".pw1", prefix + ".pwconv1", prefix + ".norm")
        wt2 = self._w(name + ".pw2.wt", lambda: self._kmajor(self._p(prefix + ".pwconv2.weight")))
        b2 = self._vec(name + ".pw2.b", lambda: self._p(prefix + ".pwconv2.bias"))
        gm = self._vec(name + ".gamma", lambda: self._p(prefix + ".gamma"))
        # The 4C-wide hidden tensor is written by pw1 and read once by pw2.  For the whole batch it is far larger than the L2
        # (B=64, C=32 at 80^2: 210 MB); run the pair per group of frames through ONE small hidden buffer that stays L2-resident
        # (dirty lines are overwritten by the next group instead of being written back), as long as a group still fills the GPU.
        tc_mode = self.model.use_tensor_cores
        if (tc_mode in (True, "all") and getattr(self.model, "fuse_mlp", True) and not self.in_pack
                and self.lib.ach_mlp_tc_supported(Cc) and P % 4 == 0):
            # ONE launch: the hidden tile lives in tensor memory (csrc/mlp_tc.cu); W1 in 32-column tiles, W2 as one C-column tile
            def tiles(wt, K_, O, NT):
                n = self.lib.ach_pack_pw_tc_nt_elems(K_, O, NT)
                hi, lo = self._zeros(n), self._zeros(n)
                self._keep += [hi, lo]
                self.pack_ops.append((self.lib.ach_pack_pw_tc_nt, (wt.data_ptr(), K_, O, wt.shape[-1], NT, hi.data_ptr(), lo.data_ptr())))
                return hi, lo
            h1, l1 = tiles(wt1, Cc, 4 * Cc, 32)
            h2, l2 = tiles(wt2, 4 * Cc, Cc, Cc)
            wsum = self._w(name + ".pw1.wsum", lambda wt=wt1, O=4 * Cc: wt.detach().cpu()[:, :O].double().sum(0))
            s = AchMlp()
            s.x, s.x_bs, s.res, s.res_bs, s.out, s.out_bs = x_ln_src.ptr, x_ln_src.bs, res.ptr, res.bs, out.ptr, out.bs
            s.b1, s.b2, s.gamma = b1.data_ptr(), b2.data_ptr(), gm.data_ptr()
            s.B, s.C, s.P, s.ln_eps = self.B, Cc, P, 1e-6
            assert x_ln_src.C == res.C == out.C == Cc
            self._keep.append(s)
            nbytes = 4 * (3 * self.B * Cc * P + 2 * 4 * Cc * Cc)
            self._add(name + ".mlp", self.lib.ach_mlp_tc, C.byref(s), h1.data_ptr(), l1.data_ptr(), h2.data_ptr(), l2.data_ptr(),
                      wsum.data_ptr(), nbytes=nbytes)
            return
        hid_bytes = 4 * 4 * Cc * P
        nb = self.B
        g_max = getattr(self.model, "mlp_group_bytes", 0)
        if g_max and hid_bytes * self.B > g_max and self.model.use_tensor_cores:
            tiles = (P + 127) // 128 * ((4 * Cc + 127) // 128)
            nb = max(1, min(self.B, g_max // hid_bytes))
            while nb < self.B and nb * tiles < getattr(self.model, "mlp_group_min_tiles", 444):   # a group still fills a wave of CTAs
                nb += 1
        if nb >= self.B:
            h = self.buf(name + ".h", 4 * Cc, x_ln_src.H, x_ln_src.W)
            self.pw(name + ".pw1", x_ln_src, h, wt1, 4 * Cc, bias=b1, ln=True, ln_eps=1e-6, act=ACT_GELU)
            self.pw(name + ".pw2", h, out, wt2, Cc, bias=b2, res=res, gamma=gm)
            return
        ht = torch.empty(nb, 4 * Cc, x_ln_src.H, x_ln_src.W, device=self.device, dtype=torch.float32)
        self._bufs[name + ".h"] = ht
        hv = View(ht.data_ptr(), ht.stride(0), 4 * Cc, x_ln_src.H, x_ln_src.W)
        for gi, b0 in enumerate(range(0, self.B, nb)):
            n = min(nb, self.B - b0)
            sfx = "" if gi == 0 else f"@{b0}"
            # the hidden view always starts at the buffer's first frame: shift it back by b0
            hshift = hv._replace(ptr=hv.ptr - b0 * hv.bs * 4)
            self.pw(name + ".pw1" + sfx, x_ln_src, hshift, wt1, 4 * Cc, bias=b1, ln=True, ln_eps=1e-6, act=ACT_GELU, frames=(b0, n))
            self.pw(name + ".pw2" + sfx, hshift, out, wt2, Cc, bias=b2, res=res, gamma=gm, frames=(b0, n))

    def conv_encoder(self, name, prefix, x, k):
        d = self.buf(name + ".dw", x.C, x.H, x.W)
        w = self._w(name + ".dw.w", lambda: self._p(prefix + ".dwconv.weight").flatten(1))
        b = self._vec(name + ".dw.b", lambda: self._p(prefix + ".dwconv.bias"))
        self.dw(name + ".dw", x, d, w, k, bias=b)
        out = self.buf(name + ".out", x.C, x.H, x.W)
        self.mlp_res(name, prefix, d, x, out)
        return out

    def _fourier_pos(self, prefix, dim, H, W, hidden=32, temperature=10000.0):
        """PositionalEncodingFourier features (layers.py:47-58) -> constant (64, H*W); the learned 1x1
        token_projection is applied by the pw kernel at pack time."""
        scale = 2 * math.pi
        y = torch.arange(1, H + 1, dtype=torch.float32).view(H, 1).expand(H, W) / (float(H) + 1e-6) * scale
        x = torch.arange(1, W + 1, dtype=torch.float32).view(1, W).expand(H, W) / (float(W) + 1e-6) * scale
        dim_t = torch.arange(hidden, dtype=torch.float32)
        dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / hidden)
        px, py = x[:, :, None] / dim_t, y[:, :, None] / dim_t
        px = torch.stack((px[:, :, 0::2].sin(), px[:, :, 1::2].cos()), dim=3).flatten(2)
        py = torch.stack((py[:, :, 0::2].sin(), py[:, :, 1::2].cos()), dim=3).flatten(2)
        return torch.cat((py, px), dim=2).permute(2, 0, 1).reshape(2 * hidden, H * W).contiguous()

    def sdta(self, name, prefix, x, scales, heads, use_pos):
        Cc, H, W = x.C, x.H, x.W
        N = H * W
        width = max(int(math.ceil(Cc / scales)), int(math.floor(Cc // scales)))
        nums = 1 if scales == 1 else scales - 1
        y = self.buf(name + ".y", Cc, H, W)
        pos_ptr = None
        if use_pos:
            # pos = token_projection(fourier(H, W)) is input independent: evaluated by the pw kernel when
            # weights are (re)packed, into a (C, N) constant that is added to every frame.
            feat = self._dev(self._fourier_pos(prefix, Cc, H, W))
            pos = torch.empty(1, Cc, H, W, device=self.device, dtype=torch.float32)
            self._keep.append(pos)
            wt = self._w(name + ".pos.wt", lambda: self._kmajor(self._p(prefix + ".pos_embd.token_projection.weight").flatten(1)))
            bi = self._vec(name + ".pos.b", lambda: self._p(prefix + ".pos_embd.token_projection.bias"))
            s = AchPwConv()
            s.x0, s.x0_bs, s.c0, s.c1 = feat.data_ptr(), 0, feat.shape[0], 0
            s.wt, s.ldw, s.bias = wt.data_ptr(), wt.shape[-1], bi.data_ptr()
            s.out, s.out_bs, s.B, s.O, s.P = pos.data_ptr(), 0, 1, Cc, N
            self._keep.append(s)
            self.pack_ops.append((self.lib.ach_pw_conv, (C.byref(s),)))
            pos_ptr = pos.data_ptr()
        fuse_pos = use_pos and nums == 1
        for i in range(nums):
            w = self._w(f"{name}.convs{i}.w", (lambda i=i: self._p(f"{prefix}.convs.{i}.weight").flatten(1)))
            b = self._vec(f"{name}.convs{i}.b", (lambda i=i: self._p(f"{prefix}.convs.{i}.bias")))
            self.dw(f"{name}.convs{i}", self.sl(x, i * width, (i + 1) * width), self.sl(y, i * width, (i + 1) * width), w, 3, bias=b,
                    xadd=self.sl(y, (i - 1) * width, i * width) if i > 0 else None,
                    post=(pos_ptr + i * width * N * 4) if fuse_pos else None)
        rest0 = nums * width
        self._add(name + ".copy", self.lib.ach_copy_add, x.ptr + rest0 * N * 4, x.bs, (pos_ptr + rest0 * N * 4) if fuse_pos else None,
                  y.ptr + rest0 * N * 4, y.bs, self.B, Cc - rest0, N)
        if use_pos and not fuse_pos:
            self._add(name + ".pos", self.lib.ach_copy_add, y.ptr, y.bs, pos_ptr, y.ptr, y.bs, self.B, Cc, N)
        # XCA
        qkv = self.buf(name + ".qkv", 3 * Cc, H, W)
        wtq, bq = self._ln_fold_linear(name + ".qkv", prefix + ".xca.qkv", prefix + ".norm_xca")
        self.pw(name + ".qkv", y, qkv, wtq, 3 * Cc, bias=bq, ln=True, ln_eps=1e-6)
        ldw = _ceil4(Cc)
        pwt = self._w(name + ".proj.wt", lambda: self._kmajor(self._p(prefix + ".xca.proj.weight")))
        temp = self._vec(name + ".temp", lambda: self._p(prefix + ".xca.temperature").flatten())
        weff = torch.zeros(self.B, Cc, ldw, device=self.device, dtype=torch.float32)
        self._keep.append(weff)
        self._add(name + ".xca_fold", self.lib.ach_xca_fold, qkv.ptr, qkv.bs, temp.data_ptr(), pwt.data_ptr(), ldw, weff.data_ptr(),
                  Cc * ldw, self.B, Cc, heads, N)
        t = self.buf(name + ".t", Cc, H, W)
        pb = self._vec(name + ".proj.b", lambda: self._p(prefix + ".xca.proj.bias"))
        gx = self._vec(name + ".gamma_xca", lambda: self._p(prefix + ".gamma_xca"))
        self.pw(name + ".xca_apply", self.sl(qkv, 2 * Cc, 3 * Cc), t, weff.data_ptr(), Cc, bias=pb, res=y, gamma=gx, wt_bs=Cc * ldw, ldw=ldw)
        out = self.buf(name + ".out", Cc, H, W)
        self.mlp_res(name, prefix, t, x, out)
        return out

    def edgenext(self, x, prefix, phi):
        cfg = Hd.EDGENEXT_CFG[phi]
        dims, depths, heads = cfg["dims"], cfg["depths"], cfg["heads"]
        H = self.res // 4
        cur = self.buf("bb.stem", dims[0], H, H)
        ds0 = prefix + ".downsample_layers.0"
        self.conv("bb.stem", x, cur, self._pack_conv("bb.stem.w", ds0 + ".0.weight"), 4, 4, 0,
                  bias=self._vec("bb.stem.b", lambda: self._p(ds0 + ".0.bias")),
                  ln_w=self._vec("bb.stem.lnw", lambda: self._p(ds0 + ".1.weight")),
                  ln_b=self._vec("bb.stem.lnb", lambda: self._p(ds0 + ".1.bias")), ln_eps=1e-6)
        feats = []
        for i in range(4):
            if i > 0:
                ds = f"{prefix}.downsample_layers.{i}"
                # LayerNorm + 2x2 space-to-depth, then the k=2,s=2 conv as a pointwise GEMM over K = 4C
                s2d = self.buf(f"bb.ds{i}.s2d", 4 * dims[i - 1], H // 2, H // 2)
                lw = self._vec(f"bb.ds{i}.lnw", (lambda ds=ds: self._p(ds + ".0.weight")))
                lb = self._vec(f"bb.ds{i}.lnb", (lambda ds=ds: self._p(ds + ".0.bias")))
                self._add(f"bb.ds{i}.ln_s2d", self.lib.ach_ln_s2d, cur.ptr, cur.bs, lw.data_ptr(), lb.data_ptr(), s2d.ptr, s2d.bs,
                          self.B, dims[i - 1], H, H, 1e-6, nbytes=4 * self.B * dims[i - 1] * H * H * 2)
                H //= 2
                nxt = self.buf(f"bb.ds{i}", dims[i], H, H)
                self.pw_bias(f"bb.ds{i}.conv", ds + ".1", s2d, nxt)
                cur = nxt
            for j in range(depths[i]):
                bp = f"{prefix}.stages.{i}.{j}"
                if i > 0 and j == depths[i] - 1:
                    cur = self.sdta(f"bb.s{i}.{j}", bp, cur, Hd.EN_SCALES[i], heads[i], Hd.EN_POS[i])
                else:
                    cur = self.conv_encoder(f"bb.s{i}.{j}", bp, cur, Hd.EN_KERNELS[i])
                self.taps[f"backbone.stage{i}.{j}"] = cur
            feats.append(cur)
        return feats

    # ---- EdgeViT (edgevit_modules/edgevit.py)
    def _ev_pos_embed(self, name, prefix, x):
        """x + dwconv3x3(x): the identity is folded into the centre tap"""
        def w(prefix=prefix):
            wt = self._p(prefix + ".weight").flatten(1).clone()
            wt[:, 4] += 1.0
            return wt
        out = self.buf(name, x.C, x.H, x.W)
        self.dw(name, x, out, self._w(name + ".w", w), 3, bias=self._vec(name + ".b", lambda: self._p(prefix + ".bias")))
        return out

    def _ev_lin(self, name, lin, x, out, act=ACT_NONE, res=None, ln=None, ln_eps=1e-6, bn=None):
        """Linear / 1x1 conv (+bias) over channels, optionally behind a LayerNorm (affine folded, statistics in the GEMM) or
        an eval BatchNorm (folded into weight and bias: W' = W diag(s), b' = b + W t)"""
        def w2d():
            return self._p(lin + ".weight").flatten(1)
        if ln is not None:
            wt = self._w(name + ".wt", lambda: self._kmajor(w2d() * self._p(ln + ".weight")[None, :]))
            bi = self._vec(name + ".b", lambda: self._p(lin + ".bias") + w2d() @ self._p(ln + ".bias"))
            self.pw(name, x, out, wt, out.C, bias=bi, ln=True, ln_eps=ln_eps, act=act, res=res)
        elif bn is not None:
            wt = self._w(name + ".wt", lambda: self._kmajor(w2d() * self._bn_fold(bn, 1e-5)[0][None, :]))
            bi = self._vec(name + ".b", lambda: self._p(lin + ".bias") + w2d() @ self._bn_fold(bn, 1e-5)[1])
            self.pw(name, x, out, wt, out.C, bias=bi, act=act, res=res)
        else:
            wt = self._w(name + ".wt", lambda: self._kmajor(w2d()))
            self.pw(name, x, out, wt, out.C, bias=self._vec(name + ".b", lambda: self._p(lin + ".bias")), act=act, res=res)

    def ev_local_agg(self, name, prefix, x):
        """LocalAgg (edgevit.py:101-119)"""
        C_, H, W = x.C, x.H, x.W
        xa = self._ev_pos_embed(name + ".pos", prefix + ".pos_embed", x)
        c1 = self.buf(name + ".c1", C_, H, W)
        self._ev_lin(name + ".conv1", prefix + ".conv1", xa, c1, bn=prefix + ".norm1")
        a5 = self.buf(name + ".a5", C_, H, W)
        self.dw(name + ".attn", c1, a5, self._w(name + ".attn.w", lambda: self._p(prefix + ".attn.weight").flatten(1)), 5,
                bias=self._vec(name + ".attn.b", lambda: self._p(prefix + ".attn.bias")))
        xb = self.buf(name + ".xb", C_, H, W)
        self._ev_lin(name + ".conv2", prefix + ".conv2", a5, xb, res=xa)
        h = self.buf(name + ".h", 4 * C_, H, W)
        self._ev_lin(name + ".fc1", prefix + ".mlp.fc1", xb, h, act=ACT_GELU, bn=prefix + ".norm2")
        out = self.buf(name + ".out", C_, H, W)
        self._ev_lin(name + ".fc2", prefix + ".mlp.fc2", h, out, res=xb)
        return out

    def ev_self_attn(self, name, prefix, x, heads, sr):
        """SelfAttn with GlobalSparseAttn (edgevit.py:50-98,122-148): tokens stay channel-major (C, H*W)"""
        C_, H, W = x.C, x.H, x.W
        a = prefix + ".attn"
        xa = self._ev_pos_embed(name + ".pos", prefix + ".pos_embed", x)
        xs = xa
        if sr > 1:
            xs = self.buf(name + ".sub", C_, H // sr, W // sr)
            self._add(name + ".sub", self.lib.ach_subsample, xa.ptr, xa.bs, xs.ptr, xs.bs, self.B, C_, H, W, sr)
        n = xs.H * xs.W
        qkv = self.buf(name + ".qkv", 3 * C_, xs.H, xs.W)
        self._ev_lin(name + ".qkv", a + ".qkv", xs, qkv, ln=prefix + ".norm1", ln_eps=1e-6)   # LN per token commutes with the sampling
        ao = self.buf(name + ".ao", C_, xs.H, xs.W)
        d = C_ // heads
        self._add(name + ".mhsa", self.lib.ach_mhsa, qkv.ptr, qkv.bs, ao.ptr, ao.bs, self.B, heads, d, n, float(d) ** -0.5)
        t = self.buf(name + ".t", C_, H, W)
        if sr > 1:
            up = self.buf(name + ".up", C_, H, W)
            wT = self._w(name + ".lp.w", lambda: self._p(a + ".LocalProp.weight").flatten(1))
            bT = self._vec(name + ".lp.b", lambda: self._p(a + ".LocalProp.bias"))
            self._add(name + ".localprop", self.lib.ach_dw_convT, ao.ptr, ao.bs, wT.data_ptr(), bT.data_ptr(), up.ptr, up.bs, self.B, C_,
                      xs.H, xs.W, sr)
            self._ev_lin(name + ".proj", a + ".proj", up, t, res=xa, ln=a + ".norm", ln_eps=1e-5)
        else:
            self._ev_lin(name + ".proj", a + ".proj", ao, t, res=xa)
        h = self.buf(name + ".h", 4 * C_, H, W)
        self._ev_lin(name + ".fc1", prefix + ".mlp.fc1", t, h, act=ACT_GELU, ln=prefix + ".norm2", ln_eps=1e-6)
        out = self.buf(name + ".out", C_, H, W)
        self._ev_lin(name + ".fc2", prefix + ".mlp.fc2", h, out, res=t)
        return out

    def edgevit(self, x, prefix, phi):
        cfg = Hd.EDGEVIT_CFG[phi]
        dims, depth = cfg["dims"], cfg["depth"]
        heads = [d // cfg["head_dim"] for d in dims]
        feats, cur, H = [], x, self.res
        for i in range(4):
            patch = 4 if i == 0 else 2
            H //= patch
            pe = f"{prefix}.patch_embed{i + 1}"
            emb = self.buf(f"bb.pe{i}", dims[i], H, H)
            lw = self._vec(f"bb.pe{i}.lnw", (lambda pe=pe: self._p(pe + ".norm.weight")))
            lb = self._vec(f"bb.pe{i}.lnb", (lambda pe=pe: self._p(pe + ".norm.bias")))
            if dims[i] <= 32:      # conv_dense applies the channel LayerNorm in its epilogue
                w = self._pack_conv(f"bb.pe{i}.w", pe + ".proj.weight")
                b = self._vec(f"bb.pe{i}.b", (lambda pe=pe: self._p(pe + ".proj.bias")))
                self.conv(f"bb.pe{i}", cur, emb, w, patch, patch, 0, bias=b, ln_w=lw, ln_b=lb, ln_eps=1e-5)
            else:
                # kernel = stride conv == space-to-depth + pointwise GEMM over K = C p^2 (tcgen05); the direct conv took 0.2-0.3 ms
                tmp = self.buf(f"bb.pe{i}.conv", dims[i], H, H)
                s2d = self.buf(f"bb.pe{i}.s2d", cur.C * patch * patch, H, H)
                self._add(f"bb.pe{i}.s2d", self.lib.ach_s2d, cur.ptr, cur.bs, s2d.ptr, s2d.bs, self.B, cur.C, cur.H, cur.W, patch,
                          nbytes=8 * self.B * cur.C * cur.H * cur.W)
                self.pw_bias(f"bb.pe{i}.conv", pe + ".proj", s2d, tmp)
                self._add(f"bb.pe{i}.ln", self.lib.ach_layernorm_cf, tmp.ptr, tmp.bs, lw.data_ptr(), lb.data_ptr(), emb.ptr, emb.bs,
                          self.B, dims[i], H * H, 1e-5)
            cur = emb
            if i < 3:
                feats.append(cur)
            for j in range(depth[i]):
                bp = f"{prefix}.blocks{i + 1}.{j}"
                if Hd.EDGEVIT_SR[i] > 1:
                    cur = self.ev_local_agg(f"bb.s{i}.{j}.la", bp + ".LocalAgg", cur)
                cur = self.ev_self_attn(f"bb.s{i}.{j}.sa", bp + ".SelfAttn", cur, heads[i], Hd.EDGEVIT_SR[i])
                self.taps[f"backbone.stage{i}.{j}"] = cur
        # final BatchNorm2d: a per-channel affine, run as a GEMM with the identity matrix
        f5 = self.buf("bb.f5", dims[3], H, H)
        eye = self._w("bb.norm.eye", lambda: self._kmajor(torch.eye(dims[3])))
        sc = self._vec("bb.norm.s", lambda: self._bn_fold(prefix + ".norm", 1e-5)[0])
        bi = self._vec("bb.norm.b", lambda: self._bn_fold(prefix + ".norm", 1e-5)[1])
        self.pw("bb.norm", cur, f5, eye, dims[3], scale=sc, bias=bi)
        feats.append(f5)
        for n_, f in zip("2345", feats):
            self.taps[f"backbone.feat{n_}"] = f
        return feats

    # ---- EfficientFormerV2 "ImageEncoder" (backbone/vision/ImageEncoder.py)
    def _ef_cbn(self, name, prefixes, x, out, act=ACT_NONE, res=None, gamma=None, conv_bn=None):
        """1x1 Conv2d (bias) + BatchNorm2d as one GEMM; several (conv, bn) pairs are concatenated along the outputs"""
        pairs = conv_bn if conv_bn is not None else [(p_ + ".0", p_ + ".1") for p_ in prefixes]
        wt = self._w(name + ".wt", lambda: self._kmajor(torch.cat([self._p(c + ".weight").flatten(1) for c, _ in pairs], 0)))
        sc = self._vec(name + ".s", lambda: torch.cat([self._bn_fold(b_, 1e-5, self._p(c + ".bias"))[0] for c, b_ in pairs]))
        bi = self._vec(name + ".b", lambda: torch.cat([self._bn_fold(b_, 1e-5, self._p(c + ".bias"))[1] for c, b_ in pairs]))
        self.pw(name, x, out, wt, out.C, scale=sc, bias=bi, act=act, res=res, gamma=gamma)

    def _ef_dwbn(self, name, conv, bnp, x, out, stride=1, act=ACT_NONE, center_identity=False):
        """depthwise 3x3 (bias) [+ BatchNorm]; center_identity adds the input pixel under the centre tap (LGQuery's pool branch)"""
        def w(conv=conv):
            wt = self._p(conv + ".weight").flatten(1).clone()
            if center_identity:
                wt[:, 4] += 1.0
            return wt
        if bnp is None:
            self.dw(name, x, out, self._w(name + ".w", w), 3, stride=stride, bias=self._vec(name + ".b", lambda: self._p(conv + ".bias")), act=act)
        else:
            sc = self._vec(name + ".s", lambda: self._bn_fold(bnp, 1e-5, self._p(conv + ".bias"))[0])
            bi = self._vec(name + ".b", lambda: self._bn_fold(bnp, 1e-5, self._p(conv + ".bias"))[1])
            self.dw(name, x, out, self._w(name + ".w", w), 3, stride=stride, scale=sc, bias=bi, act=act)

    def _ef_conv3s2(self, name, conv, bnp, x, out, act=ACT_NONE):
        sc = self._vec(name + ".s", lambda: self._bn_fold(bnp, 1e-5, self._p(conv + ".bias"))[0])
        bi = self._vec(name + ".b", lambda: self._bn_fold(bnp, 1e-5, self._p(conv + ".bias"))[1])
        self.conv(name, x, out, self._pack_conv(name + ".w", conv + ".weight"), 3, 2, 1, scale=sc, bias=bi, act=act)

    def ef_mlp(self, name, prefix, x, ls_key):
        """x + layer_scale_2 * Mlp(x)  (ImageEncoder.py:374-389,444-449)"""
        hidden = self._params[prefix + ".fc1.weight"].shape[0]
        h1 = self.buf(name + ".h1", hidden, x.H, x.W)
        self._ef_cbn(name + ".fc1", None, x, h1, act=ACT_GELU, conv_bn=[(prefix + ".fc1", prefix + ".norm1")])
        h2 = self.buf(name + ".h2", hidden, x.H, x.W)
        self._ef_dwbn(name + ".mid", prefix + ".mid", prefix + ".mid_norm", h1, h2, act=ACT_GELU)
        out = self.buf(name + ".out", x.C, x.H, x.W)
        ls = self._vec(name + ".ls", lambda: self._p(ls_key).flatten())
        self._ef_cbn(name + ".fc2", None, h2, out, res=x, gamma=ls, conv_bn=[(prefix + ".fc2", prefix + ".norm2")])
        return out

    def _ef_ab(self, name, prefix):
        return self._w(name + ".ab", lambda: self._p(prefix + ".attention_biases")[:, self._pi(prefix + ".attention_bias_idxs")])

    def ef_attention4d(self, name, prefix, x, ls_key, stride):
        """x + layer_scale_1 * Attention4D(x)  (ImageEncoder.py:63-160,415-418)"""
        heads, kd, d = 8, 32, 128
        C_, H, W = x.C, x.H, x.W
        xs = x
        if stride is not None:
            xs = self.buf(name + ".sc", C_, H // stride, W // stride)
            self._ef_dwbn(name + ".stride_conv", prefix + ".stride_conv.0", prefix + ".stride_conv.1", x, xs, stride=stride)
        n = xs.H * xs.W
        qkv = self.buf(name + ".qkv", heads * (2 * kd + d), xs.H, xs.W)
        self._ef_cbn(name + ".qkv", [prefix + ".q", prefix + ".k", prefix + ".v"], xs, qkv)
        q, k, v = self.sl(qkv, 0, heads * kd), self.sl(qkv, heads * kd, 2 * heads * kd), self.sl(qkv, 2 * heads * kd, heads * (2 * kd + d))
        vl = self.buf(name + ".vlocal", heads * d, xs.H, xs.W)
        self._ef_dwbn(name + ".v_local", prefix + ".v_local.0", prefix + ".v_local.1", v, vl)
        ab = self._ef_ab(name, prefix)
        th = [self._w(f"{name}.th{i}", (lambda i=i: torch.cat([self._p(f"{prefix}.talking_head{i}.weight").flatten(),
                                                              self._p(f"{prefix}.talking_head{i}.bias")]))) for i in (1, 2)]
        o = self.buf(name + ".o", heads * d, xs.H, xs.W)
        self._add(name + ".attn", self.lib.ach_ef_attention, q.ptr, q.bs, k.ptr, k.bs, v.ptr, v.bs, ab.data_ptr(), th[0].data_ptr(),
                  th[1].data_ptr(), vl.ptr, vl.bs, o.ptr, o.bs, self.B, heads, kd, d, n, n, float(kd) ** -0.5, 0 if stride is not None else 1)
        src = o
        if stride is not None:
            src = self.buf(name + ".up", heads * d, H, W)
            self._add(name + ".up", self.lib.ach_upsample2x_hp, o.ptr, o.bs, src.ptr, src.bs, self.B, heads * d, xs.H, xs.W, 1)
        out = self.buf(name + ".out", C_, H, W)
        ls = self._vec(name + ".ls", lambda: self._p(ls_key).flatten())
        self._ef_cbn(name + ".proj", None, src, out, res=x, gamma=ls, conv_bn=[(prefix + ".proj.1", prefix + ".proj.2")])
        return out

    def ef_embedding_asub(self, name, prefix, x, cout):
        """Embedding(asub=True): Attention4DDownsample(x) + BN(conv3x3 s2(x))  (ImageEncoder.py:193-289,329-336)"""
        heads, kd, d = 8, 16, 64
        a = prefix + ".attn"
        C_, H, W = x.C, x.H, x.W
        h2, w2 = H // 2, W // 2
        qd = self.buf(name + ".qd", C_, h2, w2)
        self._ef_dwbn(name + ".q.local", a + ".q.local.0", None, x, qd, stride=2, center_identity=True)   # local(x) + AvgPool2d(1, 2)(x)
        q = self.buf(name + ".q", heads * kd, h2, w2)
        self._ef_cbn(name + ".q.proj", [a + ".q.proj"], qd, q)
        kv = self.buf(name + ".kv", heads * (kd + d), H, W)
        self._ef_cbn(name + ".kv", [a + ".k", a + ".v"], x, kv)
        k, v = self.sl(kv, 0, heads * kd), self.sl(kv, heads * kd, heads * (kd + d))
        vl = self.buf(name + ".vlocal", heads * d, h2, w2)
        self._ef_dwbn(name + ".v_local", a + ".v_local.0", a + ".v_local.1", v, vl, stride=2)
        ab = self._ef_ab(name, a)
        o = self.buf(name + ".o", heads * d, h2, w2)
        self._add(name + ".attn", self.lib.ach_ef_attention, q.ptr, q.bs, k.ptr, k.bs, v.ptr, v.bs, ab.data_ptr(), None, None, vl.ptr, vl.bs,
                  o.ptr, o.bs, self.B, heads, kd, d, h2 * w2, H * W, float(kd) ** -0.5, 1)
        cb = self.buf(name + ".conv", cout, h2, w2)
        self._ef_conv3s2(name + ".conv", prefix + ".conv", prefix + ".bn", x, cb)
        out = self.buf(name + ".out", cout, h2, w2)
        self._ef_cbn(name + ".proj", None, o, out, res=cb, conv_bn=[(a + ".proj.1", a + ".proj.2")])
        return out

    def efficientformer(self, x, prefix, phi):
        dims, depth, vit = Hd.WIDTHS[phi], Hd.EF_DEPTH[phi], Hd.EF_VIT_NUM[phi]
        H = self.res // 2
        s1 = self.buf("bb.stem1", dims[0] // 2, H, H)
        self._ef_conv3s2("bb.stem1", prefix + ".patch_embed.0", prefix + ".patch_embed.1", x, s1, act=ACT_GELU)
        H //= 2
        cur = self.buf("bb.stem2", dims[0], H, H)
        self._ef_conv3s2("bb.stem2", prefix + ".patch_embed.3", prefix + ".patch_embed.4", s1, cur, act=ACT_GELU)
        feats, idx = [], 0
        for i in range(4):
            for j in range(depth[i]):
                bp = f"{prefix}.network.{idx}.{j}"
                if i >= 2 and j > depth[i] - 1 - vit:
                    cur = self.ef_attention4d(f"bb.s{i}.{j}.tm", bp + ".token_mixer", cur, bp + ".layer_scale_1", 2 if i == 2 else None)
                cur = self.ef_mlp(f"bb.s{i}.{j}.mlp", bp + ".mlp", cur, bp + ".layer_scale_2")
                self.taps[f"backbone.stage{i}.{j}"] = cur
            # forked output: BatchNorm2d as a per-channel affine (identity GEMM)
            f = self.buf(f"bb.f{i}", dims[i], cur.H, cur.W)
            eye = self._w(f"bb.norm{idx}.eye", (lambda c=dims[i]: self._kmajor(torch.eye(c))))
            sc = self._vec(f"bb.norm{idx}.s", (lambda idx=idx: self._bn_fold(f"{prefix}.norm{idx}", 1e-5)[0]))
            bi = self._vec(f"bb.norm{idx}.b", (lambda idx=idx: self._bn_fold(f"{prefix}.norm{idx}", 1e-5)[1]))
            self.pw(f"bb.norm{idx}", cur, f, eye, dims[i], scale=sc, bias=bi)
            feats.append(f)
            idx += 1
            if i < 3:
                ep = f"{prefix}.network.{idx}"
                if i >= 2:
                    cur = self.ef_embedding_asub(f"bb.emb{i}", ep, cur, dims[i + 1])
                else:
                    nxt = self.buf(f"bb.emb{i}", dims[i + 1], cur.H // 2, cur.W // 2)
                    self._ef_conv3s2(f"bb.emb{i}", ep + ".proj", ep + ".norm", cur, nxt)
                    cur = nxt
                idx += 1
        for n_, f in zip("2345", feats):
            self.taps[f"backbone.feat{n_}"] = f
        return feats

    # ---- MobileViT (mobilevit_modules/mobilevit.py)
    def conv_bn_silu(self, name, prefix, x, out, k, stride=1):
        """conv_nxn_bn / conv_1x1_bn: conv (no bias) + BN(1e-5) + SiLU  (mobilevit.py:7-21)"""
        if k == 1:
            self.pw_bn_act(name, prefix + ".0", prefix + ".1", 1e-5, x, out, ACT_SILU)
        else:
            sc = self._vec(name + ".s", lambda: self._bn_fold(prefix + ".1", 1e-5)[0])
            bi = self._vec(name + ".b", lambda: self._bn_fold(prefix + ".1", 1e-5)[1])
            if k == 3 and stride == 1 and self.model.use_tensor_cores:
                self.conv3_tc(name, x, out, prefix + ".0.weight", scale=sc, bias=bi, act=ACT_SILU)
            else:
                self.conv(name, x, out, self._pack_conv(name + ".w", prefix + ".0.weight"), k, stride, 1, scale=sc, bias=bi, act=ACT_SILU)

    def mv2(self, name, prefix, x, out, stride, expansion):
        """MV2Block: pw+BN+SiLU -> dw3x3(stride)+BN+SiLU -> pw+BN (+x)  (mobilevit.py:93-131)"""
        c = prefix + ".conv"
        inp, oup = x.C, out.C
        hidden = int(inp * expansion)
        assert expansion != 1
        h1 = self.buf(name + ".h1", hidden, x.H, x.W)
        self.pw_bn_act(name + ".pw1", c + ".0", c + ".1", 1e-5, x, h1, ACT_SILU)
        h2 = self.buf(name + ".h2", hidden, out.H, out.W)
        self.dw_bn_act(name + ".dw", c + ".3", c + ".4", 1e-5, h1, h2, 3, ACT_SILU, stride=stride)
        res = x if (stride == 1 and inp == oup) else None
        self.pw_bn_act(name + ".pw2", c + ".6", c + ".7", 1e-5, h2, out, ACT_NONE, res=res)

    def mvit_block(self, name, prefix, cat, depth, dim, mlp_dim):
        """MobileViTBlock (mobilevit.py:134-165).  `cat` is the (B, 2*channel, H, W) buffer whose second half
        already holds the block input y (written there by the producer), so torch.cat((x, y)) is free."""
        ch = cat.C // 2
        H, W = cat.H, cat.W
        y = self.sl(cat, ch, 2 * ch)
        c1 = self.buf(name + ".c1", ch, H, W)
        self.conv_bn_silu(name + ".conv1", prefix + ".conv1", y, c1, 3)
        t = self.buf(name + ".t0", dim, H, W)
        self.conv_bn_silu(name + ".conv2", prefix + ".conv2", c1, t, 1)
        heads, dh = 4, 8
        inner = heads * dh
        for l in range(depth):
            lp = f"{prefix}.transformer.layers.{l}"
            n = f"{name}.l{l}"
            # attention: LN -> qkv (no bias) -> softmax(QK^T/sqrt(d))V per patch-position group -> out proj + residual
            qkv = self.buf(n + ".qkv", 3 * inner, H, W)
            wq = self._w(n + ".qkv.wt", (lambda lp=lp: self._kmajor(self._p(lp + ".0.fn.to_qkv.weight") * self._p(lp + ".0.norm.weight")[None, :])))
            bq = self._vec(n + ".qkv.b", (lambda lp=lp: self._p(lp + ".0.fn.to_qkv.weight") @ self._p(lp + ".0.norm.bias")))
            self.pw(n + ".qkv", t, qkv, wq, 3 * inner, bias=bq, ln=True, ln_eps=1e-5)
            ao = self.buf(n + ".attn", inner, H, W)
            self._add(n + ".attn", self.lib.ach_mvit_attention, qkv.ptr, qkv.bs, ao.ptr, ao.bs, self.B, heads, dh, H, W)
            t1 = self.buf(n + ".t1", dim, H, W)
            wo = self._w(n + ".out.wt", (lambda lp=lp: self._kmajor(self._p(lp + ".0.fn.to_out.0.weight"))))
            bo = self._vec(n + ".out.b", (lambda lp=lp: self._p(lp + ".0.fn.to_out.0.bias")))
            self.pw(n + ".out", ao, t1, wo, dim, bias=bo, res=t)
            # feed-forward: LN -> Linear -> SiLU -> Linear + residual
            hbuf = self.buf(n + ".ffh", mlp_dim, H, W)
            w1 = self._w(n + ".ff1.wt", (lambda lp=lp: self._kmajor(self._p(lp + ".1.fn.net.0.weight") * self._p(lp + ".1.norm.weight")[None, :])))
            b1 = self._vec(n + ".ff1.b", (lambda lp=lp: self._p(lp + ".1.fn.net.0.bias") + self._p(lp + ".1.fn.net.0.weight") @ self._p(lp + ".1.norm.bias")))
            self.pw(n + ".ff1", t1, hbuf, w1, mlp_dim, bias=b1, ln=True, ln_eps=1e-5, act=ACT_SILU)
            t2 = self.buf(n + ".t2", dim, H, W)
            w2 = self._w(n + ".ff2.wt", (lambda lp=lp: self._kmajor(self._p(lp + ".1.fn.net.3.weight"))))
            b2 = self._vec(n + ".ff2.b", (lambda lp=lp: self._p(lp + ".1.fn.net.3.bias")))
            self.pw(n + ".ff2", hbuf, t2, w2, dim, bias=b2, res=t1)
            t = t2
        self.conv_bn_silu(name + ".conv3", prefix + ".conv3", t, self.sl(cat, 0, ch), 1)
        out = self.buf(name + ".out", ch, H, W)
        self.conv_bn_silu(name + ".conv4", prefix + ".conv4", cat, out, 3)
        return out

    def mobilevit(self, x, prefix, phi):
        cfg = Hd.MOBILEVIT_CFG[phi]
        dims, ch, e = cfg["dims"], cfg["channels"], cfg["expansion"]
        L = [2, 4, 3]
        R = self.res
        c0 = self.buf("mv.conv1", ch[0], R // 2, R // 2)
        self.conv_bn_silu("mv.conv1", prefix + ".conv1", x, c0, 3, stride=2)
        m0 = self.buf("mv.mv2_0", ch[1], R // 2, R // 2)
        self.mv2("mv.mv2_0", prefix + ".mv2.0", c0, m0, 1, e)
        m1 = self.buf("mv.mv2_1", ch[2], R // 4, R // 4)
        self.mv2("mv.mv2_1", prefix + ".mv2.1", m0, m1, 2, e)
        m2 = self.buf("mv.mv2_2", ch[3], R // 4, R // 4)
        self.mv2("mv.mv2_2", prefix + ".mv2.2", m1, m2, 1, e)
        f2 = self.buf("mv.mv2_3", ch[3], R // 4, R // 4)
        self.mv2("mv.mv2_3", prefix + ".mv2.3", m2, f2, 1, e)
        feats = [f2]
        cur = f2
        for i, (mi, cidx) in enumerate(((4, 5), (5, 7), (6, 9))):
            Hs = R // (8 << i)
            cat = self.buf(f"mv.cat{i}", 2 * ch[cidx], Hs, Hs)
            self.mv2(f"mv.mv2_{mi}", f"{prefix}.mv2.{mi}", cur, self.sl(cat, ch[cidx], 2 * ch[cidx]), 2, e)
            mlp = int(dims[i] * 2) if i == 0 else int(dims[i] * 4)
            cur = self.mvit_block(f"mv.mvit{i}", f"{prefix}.mvit.{i}", cat, L[i], dims[i], mlp)
            if i < 2:
                feats.append(cur)
        f5 = self.buf("mv.conv2", ch[-1], R // 32, R // 32)
        self.conv_bn_silu("mv.conv2", prefix + ".conv2", cur, f5, 1)
        feats.append(f5)
        for n_, f in zip("2345", feats):
            self.taps[f"backbone.feat{n_}"] = f
        return feats

    # ---- neck
    def spp(self, x, prefix):
        c_ = x.C // 2
        cat = self.buf("spp.cat", 4 * c_, x.H, x.W)
        self.pw_bn_act("spp.cv1", prefix + ".cv1.conv", prefix + ".cv1.bn", 1e-3, x, self.sl(cat, 0, c_), ACT_SILU)
        P4 = c_ * x.H * x.W * 4
        self._add("spp.pool", self.lib.ach_spp_maxpool, cat.ptr, cat.bs, cat.ptr + P4, cat.ptr + 2 * P4, cat.ptr + 3 * P4, cat.bs,
                  self.B, c_, x.H, x.W)
        out = self.buf("spp.out", x.C, x.H, x.W)
        self.pw_bn_act("spp.cv2", prefix + ".cv2.conv", prefix + ".cv2.bn", 1e-3, cat, out, ACT_SILU)
        return out

    def shuffle_attention(self, name, prefix, x):
        out = self.buf(name, x.C, x.H, x.W)
        ps = [self._vec(f"{name}.{n}", (lambda n=n: self._p(f"{prefix}.{n}").flatten()))
              for n in ("cweight", "cbias", "sweight", "sbias", "gn.weight", "gn.bias")]
        self._add(name, self.lib.ach_shuffle_attention, x.ptr, x.bs, out.ptr, out.bs, *[p.data_ptr() for p in ps], self.B, x.C,
                  x.H * x.W, 4, 1e-5)
        return out

    def seg_decoder(self, name, prefix, x, widths, out, mask=None):
        out = out.view() if isinstance(out, _LazyOut) else out
        chans = [widths[1], widths[0], widths[0]]
        cur = x
        for stage, c in zip(("3_to_2", "2_to_1", "1_to_0"), chans):
            up = self.buf(f"{name}.{stage}.up", c, cur.H * 2, cur.W * 2)
            self.upsample_block(f"{name}.{stage}", f"{prefix}.{name}_seg_{stage}", cur, up)
            g = self.buf(f"{name}.{stage}.ghost", c, up.H, up.W)
            self.ghost(f"{name}.{stage}.ghost", f"{prefix}.{name}_seg_ghost_{stage}", up, g, True)
            self.taps[f"neck.{name}_{stage}"] = g
            cur = g
        self.ghost(f"{name}.head", f"{prefix}.{name}_seg_head", cur, out, True)

    def seg_decoder_fused(self, name, prefix, x, widths, out, mask=None):
        """Same decoder with the Ghost primary conv hoisted below the upsampling (it commutes with the
        bilinear interpolation) and the full-resolution work in the fused ach_up_ghost[_head] kernels.
        mask = (uint8 pointer, batch stride in bytes, keep_mask): compact mode - the fused head writes class indices
        instead of logit planes (`out` is then a _LazyOut that stays unallocated)."""
        chans = [widths[1], widths[0], widths[0]]
        stages = ("3_to_2", "2_to_1", "1_to_0")
        cur = x
        chained_v = None   # low-res Ghost-primary output handed over by the previous stage's ach_up_ghost_pw2
        for si, (stage, c) in enumerate(zip(stages, chans)):
            up_p, g_p = f"{prefix}.{name}_seg_{stage}", f"{prefix}.{name}_seg_ghost_{stage}"
            n = f"{name}.{stage}"
            init = math.ceil(c / 2)
            cn = c - init
            if chained_v is not None:
                v = chained_v
            else:
                t = self.buf(n + ".t", c, cur.H, cur.W)
                self.pw_bn_act(n + ".conv", up_p + ".upsample.0.conv", up_p + ".upsample.0.bn", 1e-3, cur, t, ACT_RELU)
                v = self.buf(n + ".v", init, cur.H, cur.W)
                wt = self._w(n + ".prim.wt", (lambda g_p=g_p: self._kmajor(
                    self._bn_fold(g_p + ".primary_conv.1", 1e-5)[0][:, None] * self._p(g_p + ".primary_conv.0.weight").flatten(1))))
                self.pw(n + ".prim", t, v, wt, init)
            b1f = (lambda g_p=g_p: self._bn_fold(g_p + ".primary_conv.1", 1e-5)[1])
            w2f = (lambda g_p=g_p, cn=cn: self._p(g_p + ".cheap_operation.0.weight")[:cn].flatten(1))
            s2f = (lambda g_p=g_p, cn=cn: self._bn_fold(g_p + ".cheap_operation.1", 1e-5)[0][:cn])
            b2f = (lambda g_p=g_p, cn=cn: self._bn_fold(g_p + ".cheap_operation.1", 1e-5)[1][:cn])
            if si < 2 and self.model.fuse_seg_chain and cn == init:
                c_next = chans[si + 1]
                init_next = math.ceil(c_next / 2)
                if self.lib.ach_up_ghost_pw2_supported(init, c_next, init_next):
                    # whole stage + next stage's 1x1 conv + next Ghost primary in one kernel: 2c-channel map never hits HBM
                    up_n, g_n = f"{prefix}.{name}_seg_{stages[si + 1]}", f"{prefix}.{name}_seg_ghost_{stages[si + 1]}"
                    vn = self.buf(n + ".vnext", init_next, v.H * 2, v.W * 2)
                    u = AchUpGhostPw2()
                    u.v, u.v_bs, u.out, u.out_bs = v.ptr, v.bs, vn.ptr, vn.bs
                    u.b1 = self._vec(n + ".b1", b1f).data_ptr()
                    u.w2, u.s2, u.b2 = (self._w(n + ".w2", w2f).data_ptr(), self._vec(n + ".s2", s2f).data_ptr(),
                                        self._vec(n + ".b2", b2f).data_ptr())
                    u.w1t = self._w(n + ".chain.w1t", (lambda up_n=up_n: (self._bn_fold(up_n + ".upsample.0.bn", 1e-3)[0][:, None]
                                                                         * self._p(up_n + ".upsample.0.conv.weight").flatten(1)).t())).data_ptr()
                    u.c1 = self._vec(n + ".chain.c1", (lambda up_n=up_n: self._bn_fold(up_n + ".upsample.0.bn", 1e-3)[1])).data_ptr()
                    u.w2t = self._w(n + ".chain.w2t", (lambda g_n=g_n: (self._bn_fold(g_n + ".primary_conv.1", 1e-5)[0][:, None]
                                                                       * self._p(g_n + ".primary_conv.0.weight").flatten(1)).t())).data_ptr()
                    u.B, u.Ci, u.C1, u.N2, u.h, u.w = self.B, init, c_next, init_next, v.H, v.W
                    self._keep.append(u)
                    nb_u = 4 * self.B * v.H * v.W * (init + 4 * init_next)
                    if self.model.use_tensor_cores and self.model.seg_tensor_cores and self.lib.ach_up_ghost_pw2_tc_supported(init, c_next, init_next):
                        # the two 1x1 convs of the chain on tcgen05: weight tiles packed on the device at (re)pack time
                        tiles = []
                        for key, K_, O_ in ((n + ".chain.w1t", 2 * init, c_next), (n + ".chain.w2t", c_next, init_next)):
                            wt_ = self._weights[key][0]
                            n_ = self.lib.ach_pack_pw_tc_elems(K_, O_)
                            hi = self._zeros(n_)
                            lo = self._zeros(n_)
                            self._keep += [hi, lo]
                            self.pack_ops.append((self.lib.ach_pack_pw_tc, (wt_.data_ptr(), K_, O_, wt_.shape[-1], hi.data_ptr(), lo.data_ptr())))
                            tiles += [hi.data_ptr(), lo.data_ptr()]
                        # depthwise taps + affines as kernel parameters (host array [Ci][12]: 9 taps, s2, b2, b1)
                        dwh = self._h(n + ".chain.dw", (lambda w2f=w2f, s2f=s2f, b2f=b2f, b1f=b1f: torch.cat(
                            [w2f(), s2f()[:, None], b2f()[:, None], b1f()[:, None]], 1)))
                        self._add(n + ".up_ghost_pw2", self.lib.ach_up_ghost_pw2_tc, C.byref(u), *tiles, dwh, nbytes=nb_u)
                    else:
                        self._add(n + ".up_ghost_pw2", self.lib.ach_up_ghost_pw2, C.byref(u), nbytes=nb_u)
                    chained_v = vn
                    cur = vn   # only its spatial size is used below
                    continue
            chained_v = None
            K = out.C
            hinit = math.ceil(K / 2)
            if si == 2 and self.lib.ach_up_ghost_head_supported(init, hinit, K) and cn == init:
                h_p = f"{prefix}.{name}_seg_head"
                a = AchUpGhostHead()
                if mask is None:
                    out = out.view() if isinstance(out, _LazyOut) else out
                    a.out, a.out_bs = out.ptr, out.bs
                a.v, a.v_bs = v.ptr, v.bs
                a.b1, a.w2, a.s2, a.b2 = self._h(n + ".b1", b1f), self._h(n + ".w2", w2f), self._h(n + ".s2", s2f), self._h(n + ".b2", b2f)
                a.w3 = self._h(n + ".w3", lambda: (self._bn_fold(h_p + ".primary_conv.1", 1e-5)[0][:, None]
                                                   * self._p(h_p + ".primary_conv.0.weight").flatten(1)).t())
                a.b3 = self._h(n + ".b3", lambda: self._bn_fold(h_p + ".primary_conv.1", 1e-5)[1])
                a.w4 = self._h(n + ".w4", lambda: self._p(h_p + ".cheap_operation.0.weight")[:K - hinit].flatten(1))
                a.s4 = self._h(n + ".s4", lambda: self._bn_fold(h_p + ".cheap_operation.1", 1e-5)[0][:K - hinit])
                a.b4 = self._h(n + ".b4", lambda: self._bn_fold(h_p + ".cheap_operation.1", 1e-5)[1][:K - hinit])
                a.B, a.C, a.init, a.K, a.h, a.w = self.B, init, hinit, K, cur.H, cur.W
                self._keep.append(a)
                if mask is not None:
                    self._add(n + ".up_ghost_head", self.lib.ach_up_ghost_head_argmax, C.byref(a), mask[0], mask[1], mask[2],
                              nbytes=self.B * cur.H * cur.W * (4 * init + 4))
                else:
                    self._add(n + ".up_ghost_head", self.lib.ach_up_ghost_head, C.byref(a),
                              nbytes=4 * self.B * (init * cur.H * cur.W + K * 4 * cur.H * cur.W))
                return
            g = self.buf(n + ".ghost", c, cur.H * 2, cur.W * 2)
            u = AchUpGhost()
            u.v, u.v_bs, u.out, u.out_bs = v.ptr, v.bs, g.ptr, g.bs
            u.b1 = self._vec(n + ".b1", b1f).data_ptr()
            if cn:
                u.w2, u.s2, u.b2 = (self._w(n + ".w2", w2f).data_ptr(), self._vec(n + ".s2", s2f).data_ptr(),
                                    self._vec(n + ".b2", b2f).data_ptr())
            u.B, u.Ci, u.Cn, u.h, u.w = self.B, init, cn, cur.H, cur.W
            self._keep.append(u)
            self._add(n + ".up_ghost", self.lib.ach_up_ghost, C.byref(u), nbytes=4 * self.B * cur.H * cur.W * (init + 4 * c))
            self.taps[f"neck.{name}_{stage}"] = g
            cur = g
        out = out.view() if isinstance(out, _LazyOut) else out
        self.ghost(f"{name}.head", f"{prefix}.{name}_seg_head", cur, out, True)

    def gdf_neck(self, feats, prefix, phi, out_se, out_lane, masks=None):
        w = Hd.WIDTHS[phi]
        m2, m3, m4, m5 = feats
        f5 = self.spp(m5, prefix + ".spp")
        up4 = self.buf("fpn.up4", w[2], m4.H, m4.W)
        self.upsample_block("fpn.up54", prefix + ".upsample_5_to_4", f5, up4)
        f4 = self.ghost_bottleneck("fpn.g54", prefix + ".ghost_5_to_4", up4, m4, w[2] * 2, w[2])
        up3 = self.buf("fpn.up3", w[1], m3.H, m3.W)
        self.upsample_block("fpn.up43", prefix + ".upsample_4_to_3", f4, up3)
        f3 = self.ghost_bottleneck("fpn.g43", prefix + ".ghost_4_to_3", up3, m3, w[1] * 2, w[1])
        sa_lane = self.shuffle_attention("fpn.sa_lane", prefix + ".stage_3_lane_seg", f3)
        sa_se = self.shuffle_attention("fpn.sa_se", prefix + ".stage_3_semantic_seg", f3)
        self.taps.update({"neck.spp": f5, "neck.fpn4": f4, "neck.fpn3": f3, "neck.sa_lane": sa_lane, "neck.sa_se": sa_se})
        dec = self.seg_decoder_fused if self.model.fuse_seg_decoder else self.seg_decoder
        masks = masks or {}
        self.wait(6, 0)                    # the fusion + detection lane forks here: it needs the FPN maps, not the decoders
        self.cur_lane = 3                  # the two decoders are independent of each other
        self.wait(3, 0)
        dec("lane", prefix, sa_lane, w, out_lane, mask=masks.get("lane"))
        self._seg_finish("lane", out_lane)
        self.cur_lane = 0
        dec("se", prefix, sa_se, w, out_se, mask=masks.get("se"))
        self._seg_finish("se", out_se)
        return (f5, m5), (f4, m4), (f3, m3)

    # ---- CSP-Dual-FPN neck (neck/cspdualfpn.py)
    def conv3_bn_act(self, name, prefix, x, out, act, res=None):
        """BaseConv with a 3x3 kernel: conv (no bias) + BN(1e-3) + act (normal_conv.py:36-49); returns True when `res` was
        added by the kernel (tensor-core path), False when the caller still has to add it"""
        sc = self._vec(name + ".s", lambda: self._bn_fold(prefix + ".bn", 1e-3)[0])
        bi = self._vec(name + ".b", lambda: self._bn_fold(prefix + ".bn", 1e-3)[1])
        if self.model.use_tensor_cores and getattr(self.model, "conv3_tensor_cores", True) and x.C >= 12:   # below ~12 input channels the 16-wide K chunks are mostly padding
            self.conv3_tc(name, x, out, prefix + ".conv.weight", scale=sc, bias=bi, act=act, res=res)
            return True
        self.conv(name, x, out, self._pack_conv(name + ".w", prefix + ".conv.weight"), 3, 1, 1, scale=sc, bias=bi, act=act)
        return False

    def bottleneck(self, name, prefix, x, out, x1=None):
        """Bottleneck (cspdualfpn.py:42-56): 1x1+BN+SiLU -> 3x3+BN+ReLU, + x when in == out"""
        hidden = self._p(prefix + ".conv1.conv.weight").shape[0]
        h = self.buf(name + ".h", hidden, x.H, x.W)
        self.pw_bn_act(name + ".conv1", prefix + ".conv1.conv", prefix + ".conv1.bn", 1e-3, x, h, ACT_SILU, x1=x1)
        if x1 is None and x.C == out.C:
            if self.model.use_tensor_cores and h.C >= 12:      # the "+ x" shortcut rides in the implicit-GEMM epilogue
                self.conv3_bn_act(name + ".conv2", prefix + ".conv2", h, out, ACT_RELU, res=x)
                return
            t = self.buf(name + ".t", out.C, x.H, x.W)
            self.conv3_bn_act(name + ".conv2", prefix + ".conv2", h, t, ACT_RELU)
            self._add(name + ".add", self.lib.ach_add, t.ptr, t.bs, x.ptr, x.bs, out.ptr, out.bs, self.B, out.C, x.H * x.W,
                      nbytes=4 * self.B * out.C * x.H * x.W * 3)
        else:
            self.conv3_bn_act(name + ".conv2", prefix + ".conv2", h, out, ACT_RELU)

    def csp_layer(self, name, prefix, xa, xb, outc):
        """CSPLayer over cat(xa, xb) (cspdualfpn.py:59-78); both concatenations are channel slices / two-source reads"""
        hidden = outc // 2
        x1 = self.buf(name + ".x1", hidden, xa.H, xa.W)
        cat = self.buf(name + ".cat", 2 * hidden, xa.H, xa.W)
        self.pw_bn_act(name + ".conv1", prefix + ".conv1.conv", prefix + ".conv1.bn", 1e-3, xa, x1, ACT_SILU, x1=xb)
        self.pw_bn_act(name + ".conv2", prefix + ".conv2.conv", prefix + ".conv2.bn", 1e-3, xa, self.sl(cat, hidden, 2 * hidden), ACT_SILU, x1=xb)
        self.bottleneck(name + ".m0", prefix + ".m.0", x1, self.sl(cat, 0, hidden))
        out = self.buf(name + ".out", outc, xa.H, xa.W)
        self.pw_bn_act(name + ".conv3", prefix + ".conv3.conv", prefix + ".conv3.bn", 1e-3, cat, out, ACT_SILU)
        return out

    def seg_decoder_csp(self, name, prefix, x, widths, out):
        out = out.view() if isinstance(out, _LazyOut) else out
        chans = [widths[1], widths[0], widths[0]]
        cur = x
        for stage, c in zip(("3_to_2", "2_to_1", "1_to_0"), chans):
            up = self.buf(f"{name}.{stage}.up", c, cur.H * 2, cur.W * 2)
            self.upsample_block(f"{name}.{stage}", f"{prefix}.{name}_seg_{stage}", cur, up)
            g = self.buf(f"{name}.{stage}.bneck", c, up.H, up.W)
            self.bottleneck(f"{name}.{stage}.bneck", f"{prefix}.{name}_seg_ghost_{stage}", up, g)
            self.taps[f"neck.{name}_{stage}"] = g
            cur = g
        self.bottleneck(f"{name}.head", f"{prefix}.{name}_seg_head", cur, out)

    def cdf_neck(self, feats, prefix, phi, out_se, out_lane, masks=None):
        w = Hd.WIDTHS[phi]
        m2, m3, m4, m5 = feats
        f5 = self.spp(m5, prefix + ".spp")
        up4 = self.buf("fpn.up4", w[2], m4.H, m4.W)
        self.upsample_block("fpn.up54", prefix + ".upsample_5_to_4", f5, up4)
        f4 = self.csp_layer("fpn.c54", prefix + ".ghost_5_to_4", up4, m4, w[2])
        up3 = self.buf("fpn.up3", w[1], m3.H, m3.W)
        self.upsample_block("fpn.up43", prefix + ".upsample_4_to_3", f4, up3)
        f3 = self.csp_layer("fpn.c43", prefix + ".ghost_4_to_3", up3, m3, w[1])
        sa_lane = self.shuffle_attention("fpn.sa_lane", prefix + ".stage_3_lane_seg", f3)
        sa_se = self.shuffle_attention("fpn.sa_se", prefix + ".stage_3_semantic_seg", f3)
        self.taps.update({"neck.spp": f5, "neck.fpn4": f4, "neck.fpn3": f3, "neck.sa_lane": sa_lane, "neck.sa_se": sa_se})
        self.wait(6, 0)                    # the fusion + detection lane forks here: it needs the FPN maps, not the decoders
        # The two decoders share lane 0 here (the Ghost neck runs them on two lanes): with the lane decoder on its own lane 2 of 1 500
        # graph replays under HBM-saturating copy traffic differed in a 1x1-conv tile of one decoder while the other decoder's
        # conv3x3_tc launches shared the SMs; on one lane - or with the 3x3 convs on CUDA cores - 1 500 of 1 500 were identical
        # (tools/determinism_probe.py --neck cdf).  Reproducibility first: the interaction is not understood yet.
        self.seg_decoder_csp("lane", prefix, sa_lane, w, out_lane)
        self._seg_finish("lane", out_lane)
        self.seg_decoder_csp("se", prefix, sa_se, w, out_se)
        self._seg_finish("se", out_se)
        return (f5, m5), (f4, m4), (f3, m3)

    # ---- compact output record (achelous.py:259-318 on the device, inside the launch plan)
    def _keep_mask(self, K):
        kc = self.compact["keep_classes"]
        if kc is None:
            return 0xFFFFFFFF
        m = 0
        for c in kc:
            if not 0 <= c < min(K, 32):
                raise ValueError(f"keep_classes: class {c} outside [0, {K})")
            m |= 1 << c
        return m

    def _seg_finish(self, name, out):
        """compact mode: logits that were materialised (no fused argmax head, or masks at the original image size) -> uint8 map"""
        if self.compact is None or not isinstance(out, _LazyOut) or not out.allocated:
            return
        from .utils.seg_post import letterbox_window
        v = out.view()
        ptr, nbytes = self._mask_slots[name]
        bs = self.packed_out.stride(0)
        shape = self.compact["image_shape"]
        keep_mask = self._keep_mask(v.C) if name == "se" else 0xFFFFFFFF   # achelous.py:297 masks the semantic map only
        if shape is None:
            self._add(name + ".argmax", self.lib.ach_seg_argmax_u8, v.ptr, v.bs, self.B, v.C, v.H * v.W, keep_mask, ptr, bs,
                      nbytes=self.B * v.H * v.W * (4 * v.C + 1))
        else:
            y_off, x_off, nh, nw = letterbox_window((v.H, v.W), shape, self.compact["letterbox_image"])
            self._add(name + ".softmax_resize_argmax", self.lib.ach_seg_softmax_resize_argmax, v.ptr, v.bs, self.B, v.C, v.H, v.W,
                      y_off, x_off, nh, nw, ptr, bs, shape[0], shape[1], keep_mask,
                      nbytes=self.B * (4 * v.C * nh * nw + shape[0] * shape[1]))

    def _pc_out(self, logits, num_class, N):
        """log_softmax over the classes into the output record: fp32 (B, N, K) rows, or (compact) the argmax class byte"""
        if self.compact is None:
            self._add("pc.logsoftmax", self.lib.ach_logsoftmax_t, logits.ptr, logits.bs, self._pc_slot, self.packed_out.stride(0), self.B,
                      num_class, N)
        else:
            self._add("pc.logsoftmax_argmax", self.lib.ach_logsoftmax_argmax_t, logits.ptr, logits.bs, self._pc_slot,
                      self.packed_out.stride(0), self.B, num_class, N)

    def _det_finish(self, det_views, K):
        """compact mode: decode_outputs + non_max_suppression (utils_bbox.py:33-130) as two launches of the plan"""
        n = len(det_views)
        A = sum(v.H * v.W for v in det_views)
        ptrs = (C.c_void_p * n)(*[v.ptr for v in det_views])
        bss = (C.c_longlong * n)(*[v.bs for v in det_views])
        hs = (C.c_int * n)(*[v.H for v in det_views])
        ws = (C.c_int * n)(*[v.W for v in det_views])
        decoded = torch.empty(self.B, A, 5 + K, device=self.device, dtype=torch.float32)
        ws_bytes = self.lib.ach_nms_workspace_bytes(self.B, A)
        work = torch.empty(ws_bytes, device=self.device, dtype=torch.uint8)
        self._keep += [ptrs, bss, hs, ws, decoded, work]
        self._bufs["out.decoded"] = decoded
        R = float(self.res)
        self._add("det.decode", self.lib.ach_decode_outputs, ptrs, bss, hs, ws, n, decoded.data_ptr(), self.B, K, R, R,
                  nbytes=4 * self.B * A * (5 + K) * 2)
        cp = self.compact
        base, bs = self.packed_out.data_ptr(), self.packed_out.stride(0)
        md = min(cp["max_det"], A)
        self._add("det.nms", self.lib.ach_nms_rows, decoded.data_ptr(), self.B, A, K, cp["conf_thres"], cp["nms_thres"],
                  base + self.rec["det"][0], bs // 4, md, base + self.rec["count"][0], bs // 4, work.data_ptr(), ws_bytes,
                  nbytes=4 * self.B * (A * (5 + K) + md * 7))

    # ---- radar encoder
    def rcnet(self, x, prefix, phi):
        w = [c // 4 for c in Hd.WIDTHS[phi]]
        spec = [(x.C, w[0], True), (w[0], w[0], True)]
        for i in range(1, 4):
            spec += [(w[i - 1], w[i - 1], False), (w[i - 1], w[i], True)]
        feats = []
        cur = x
        for i, (cin, cout, down) in enumerate(spec):
            bp = f"{prefix}.rc_blocks.{i}"
            d = bp + ".radar_conv.deformable_conv"
            # tensor-core RCBlock (rcblock_tc.cu v2: channel-last gathers, resident weight tiles, chunk ring) for the
            # high-resolution blocks; use_tensor_cores = True / "all" -> on, False -> SIMT (profiles/r1_tc_vs_simt.md)
            use_tc = bool(self.model.use_tensor_cores) and self.model.rc_tensor_cores and bool(self.lib.ach_rc_deform_tc_supported(cin))
            # channel-last pooled map [P][ceil4(C)]: the 3x3 window and every bilinear corner are 16-byte loads
            pooled = self.buf(f"rc{i}.pool", _ceil4(cin), cur.H, cur.W)
            self._add(f"rc{i}.pool", self.lib.ach_avgpool3_cl, cur.ptr, cur.bs, pooled.ptr,
                      pooled.bs, self.B, cin, cur.H, cur.W, nbytes=4 * self.B * cur.H * cur.W * (cin + _ceil4(cin)))

            def w_om(d=d, cin=cin):
                wo = torch.cat([self._p(d + ".offset_conv.weight"), self._p(d + ".modulator_conv.weight")], 0)  # (27, C, 3, 3)
                out = torch.zeros(cin * 9, 28, dtype=wo.dtype, device=wo.device)
                out[:, :27] = wo.reshape(27, cin * 9).t()
                return out
            s = AchRcDeform()
            s.x, s.x_bs, s.pooled, s.pooled_bs = cur.ptr, cur.bs, pooled.ptr, pooled.bs
            s.w_om = self._w(f"rc{i}.w_om", w_om).data_ptr()
            s.b_om = self._vec(f"rc{i}.b_om", (lambda d=d: torch.cat([self._p(d + ".offset_conv.bias"), self._p(d + ".modulator_conv.bias")]))).data_ptr()
            s.w_reg = self._w(f"rc{i}.w_reg", (lambda d=d, cin=cin: self._p(d + ".regular_conv.weight").reshape(cin, cin * 9).t())).data_ptr()
            s.w1 = self._w(f"rc{i}.w1", (lambda bp=bp, cin=cin: self._p(bp + ".weight_conv1.weight").reshape(cin, cin).t())).data_ptr()
            s.scale = self._vec(f"rc{i}.s", (lambda bp=bp: self._bn_fold(bp + ".norm", 1e-5, self._p(bp + ".weight_conv1.bias"))[0])).data_ptr()
            s.bias = self._vec(f"rc{i}.b", (lambda bp=bp: self._bn_fold(bp + ".norm", 1e-5, self._p(bp + ".weight_conv1.bias"))[1])).data_ptr()
            y = self.buf(f"rc{i}.y", cin, cur.H, cur.W)
            s.out, s.out_bs, s.B, s.C, s.H, s.W = y.ptr, y.bs, self.B, cin, cur.H, cur.W
            s.pooled_cl = 1
            self._keep.append(s)
            nb_rc = 4 * self.B * cin * cur.H * cur.W * 3
            if use_tc:
                # both contractions of the block as implicit GEMMs on tcgen05, with everything linear folded into them on the host
                # (rc_tc_fold); weights packed on the device at (re)pack time
                def fold(d=d, bp=bp, cin=cin):
                    sc, bi = self._bn_fold(bp + ".norm", 1e-5, self._p(bp + ".weight_conv1.bias"))
                    return rc_tc_fold(torch.cat([self._p(d + ".offset_conv.weight"), self._p(d + ".modulator_conv.weight")], 0),
                                      torch.cat([self._p(d + ".offset_conv.bias"), self._p(d + ".modulator_conv.bias")]),
                                      self._p(d + ".regular_conv.weight"), self._p(bp + ".weight_conv1.weight"), sc, bi)
                K_ = fold()[4]
                w_om_t = self._w(f"rc{i}.w_om_tap", lambda fold=fold: fold()[0])
                w_reg_t = self._w(f"rc{i}.w_reg_tap", lambda fold=fold: fold()[2])
                s.b_om = self._vec(f"rc{i}.om_consts", lambda fold=fold: fold()[1]).data_ptr()
                tiles = []
                for wt_, O_ in ((w_om_t, 27), (w_reg_t, cin)):
                    n_ = self.lib.ach_pack_pw_tc_elems(K_, O_)
                    hi = self._zeros(n_)
                    lo = self._zeros(n_)
                    self._keep += [hi, lo]
                    self.pack_ops.append((self.lib.ach_pack_pw_tc, (wt_.data_ptr(), K_, O_, wt_.shape[-1], hi.data_ptr(), lo.data_ptr())))
                    tiles += [hi.data_ptr(), lo.data_ptr()]
                self._add(f"rc{i}.deform", self.lib.ach_rc_deform_tc, C.byref(s), *tiles, nbytes=nb_rc)
            else:
                self._add(f"rc{i}.deform", self.lib.ach_rc_deform, C.byref(s), nbytes=nb_rc)
            if down:
                out = self.buf(f"rc{i}.out", cout, cur.H // 2, cur.W // 2)
                self.conv(f"rc{i}.down", y, out, self._pack_conv(f"rc{i}.w2", bp + ".weight_conv2.weight"), 3, 2, 1,
                          bias=self._vec(f"rc{i}.b2", (lambda bp=bp: self._p(bp + ".weight_conv2.bias"))))
            else:
                out = self.buf(f"rc{i}.out", cout, cur.H, cur.W)
                self.pw_bias(f"rc{i}.pw2", bp + ".weight_conv2", y, out)
            self.taps[f"radar.block{i}"] = out
            cur = out
            if i > 1 and i % 2 == 1:
                feats.append(out)
        return feats

    # ---- image-radar fusion (IREncoder.py:79-89)
    def fuse_stage(self, s, prefix, fpn_map, radar):
        (fa, fb) = fpn_map
        Ci, Cr = fa.C, radar.C
        P = fa.H * fa.W
        out = self.buf(f"fuse{s}", Ci + Cr, fa.H, fa.W)
        sc = self._vec(f"fuse{s}.s", lambda: self._bn_fold(f"{prefix}.norm_stage{s}", 1e-5)[0])
        bi = self._vec(f"fuse{s}.b", lambda: self._bn_fold(f"{prefix}.norm_stage{s}", 1e-5)[1])
        for j, (xa, xb, Cc, off) in enumerate(((fa, fb, Ci, 0), (radar, None, Cr, Ci))):
            mean = torch.empty(self.B, Cc, device=self.device, dtype=torch.float32)
            self._keep.append(mean)
            w1d = self._vec(f"fuse{s}.eca{j}", (lambda j=j: self._p(f"{prefix}.channel_attn_stage{s}.{j}.conv.weight").flatten()))
            self._add(f"fuse{s}.mean{j}", self.lib.ach_plane_mean, xa.ptr, xa.bs, xb.ptr if xb else None, xb.bs if xb else 0,
                      mean.data_ptr(), self.B, Cc, P)
            self._add(f"fuse{s}.eca{j}", self.lib.ach_eca_fuse, xa.ptr, xa.bs, xb.ptr if xb else None, xb.bs if xb else 0,
                      mean.data_ptr(), w1d.data_ptr(), w1d.numel(), sc.data_ptr() + off * 4, bi.data_ptr() + off * 4,
                      out.ptr + off * P * 4, out.bs, self.B, Cc, P)
        self.taps[f"fuse.p{s}"] = out
        return out

    # ---- detection head
    def det_level(self, k, prefix, x, out, num_det):
        stem = self.buf(f"det{k}.stem", 64 if self.model.nano_head else 256, x.H, x.W)
        self.pw_bn_act(f"det{k}.stem", f"{prefix}.stems.{k}.conv", f"{prefix}.stems.{k}.bn", 1e-3, x, stem, ACT_RELU)
        feats = {}
        for br in ("cls", "reg"):
            cur = stem
            for j in range(2):
                bp = f"{prefix}.{br}_convs.{k}.{j}"
                d = self.buf(f"det{k}.{br}{j}.dw", cur.C, x.H, x.W)
                w = self._w(f"det{k}.{br}{j}.dw.w", (lambda bp=bp: self._p(bp + ".conv.dconv.weight").flatten(1)))
                has_bias = (bp + ".conv.dconv.bias") in self._params
                b = self._vec(f"det{k}.{br}{j}.dw.b", (lambda bp=bp: self._p(bp + ".conv.dconv.bias"))) if has_bias else None
                self.dw(f"det{k}.{br}{j}.dw", cur, d, w, 5, bias=b)
                nxt = self.buf(f"det{k}.{br}{j}.pw", cur.C, x.H, x.W)
                self.pw_bn_act(f"det{k}.{br}{j}.pw", bp + ".conv.pconv", bp + ".bn", 1e-3, d, nxt, ACT_RELU, conv_bias=has_bias)
                cur = nxt
            feats[br] = cur
        # preds: [reg(4), obj(1)] from reg_feat, cls(K) from cls_feat, written into the level's output planes
        wt_ro = self._w(f"det{k}.ro.wt", lambda: self._kmajor(torch.cat([self._p(f"{prefix}.reg_preds.{k}.weight").flatten(1),
                                                                        self._p(f"{prefix}.obj_preds.{k}.weight").flatten(1)], 0)))
        b_ro = self._vec(f"det{k}.ro.b", lambda: torch.cat([self._p(f"{prefix}.reg_preds.{k}.bias"), self._p(f"{prefix}.obj_preds.{k}.bias")]))
        self.pw(f"det{k}.regobj", feats["reg"], self.sl(out, 0, 5), wt_ro, 5, bias=b_ro)
        self.pw_bias(f"det{k}.cls", f"{prefix}.cls_preds.{k}", feats["cls"], self.sl(out, 5, 5 + num_det))

    # ---- PointNet
    def _fc(self, name, x, xk, w, O, scale=None, bias=None, act=ACT_NONE):
        out = torch.empty(self.B, O, device=self.device, dtype=torch.float32)
        self._keep.append(out)
        self._add(name, self.lib.ach_fc, x.data_ptr(), x.stride(0), w.data_ptr(), self._ptr(scale), self._ptr(bias), out.data_ptr(),
                  out.stride(0), self.B, xk, O, act)
        return out

    def _c1_bn(self, name, conv, bnp, x, O, act, **kw):
        out = self.buf(name, O, x.H, x.W)
        self.pw_bn_act(name, conv, bnp, 1e-5, x, out, act, conv_bias=True)
        return out

    def stn(self, name, prefix, x, k, kpad, ldw):
        """STN3d / STNkd -> per-frame K-major transform (B, kpad, ldw): rows k < `k` hold T (+I), the
        remaining rows are identity (feature channels pass through, pointnet_utils.py:107-112)."""
        a1 = self._c1_bn(name + ".c1", prefix + ".conv1", prefix + ".bn1", x, 64, ACT_RELU)
        a2 = self._c1_bn(name + ".c2", prefix + ".conv2", prefix + ".bn2", a1, 128, ACT_RELU)
        g = torch.empty(self.B, 1024, device=self.device, dtype=torch.float32)
        self._keep.append(g)
        self._add(name + ".fill", self.lib.ach_fill, g.data_ptr(), g.numel(), float("-inf"))
        wt = self._w(name + ".c3.wt", lambda: self._kmajor(self._p(prefix + ".conv3.weight").flatten(1)))
        sc = self._vec(name + ".c3.s", lambda: self._bn_fold(prefix + ".bn3", 1e-5, self._p(prefix + ".conv3.bias"))[0])
        bi = self._vec(name + ".c3.b", lambda: self._bn_fold(prefix + ".bn3", 1e-5, self._p(prefix + ".conv3.bias"))[1])
        self.pw(name + ".c3max", a2, g, wt, 1024, scale=sc, bias=bi, act=ACT_RELU, reduce_max=True)
        f = g
        for j, (fc, bnp, kin, kout) in enumerate(((".fc1", ".bn4", 1024, 512), (".fc2", ".bn5", 512, 256))):
            w = self._w(f"{name}{fc}.w", (lambda fc=fc: self._p(prefix + fc + ".weight")))
            s_ = self._vec(f"{name}{fc}.s", (lambda fc=fc, bnp=bnp: self._bn_fold(prefix + bnp, 1e-5, self._p(prefix + fc + ".bias"))[0]))
            b_ = self._vec(f"{name}{fc}.b", (lambda fc=fc, bnp=bnp: self._bn_fold(prefix + bnp, 1e-5, self._p(prefix + fc + ".bias"))[1]))
            f = self._fc(f"{name}{fc}", f, kin, w, kout, s_, b_, ACT_RELU)

        def w3():
            w = self._p(prefix + ".fc3.weight")  # (k*k, 256)
            out = torch.zeros(kpad, ldw, 256, dtype=w.dtype, device=w.device)
            out[:k, :k] = w.reshape(k, k, 256)
            return out.reshape(kpad * ldw, 256)

        def b3():
            b = self._p(prefix + ".fc3.bias")
            out = torch.zeros(kpad, ldw, dtype=b.dtype, device=b.device)
            out[:k, :k] = b.reshape(k, k)
            out[:kpad, :kpad] += torch.eye(kpad, dtype=b.dtype, device=b.device)
            return out.flatten()
        return self._fc(name + ".fc3", f, 256, self._w(name + ".fc3.w", w3), kpad * ldw, None, self._vec(name + ".fc3.b", b3))

    def pointnet(self, x, prefix, num_class):
        D, N = x.C, x.H * x.W
        f = prefix + ".feat"
        ldw5 = _ceil4(D)
        t3 = self.stn("pn.stn", f + ".stn", x, 3, D, ldw5)
        xt = self.buf("pn.xt", D, N)
        self.pw("pn.apply_t3", x, xt, t3.data_ptr(), D, wt_bs=D * ldw5, ldw=ldw5)
        p1 = self._c1_bn("pn.c1", f + ".conv1", f + ".bn1", xt, 32, ACT_RELU)
        tf = self.stn("pn.fstn", f + ".fstn", p1, 32, 32, 32)
        pointfeat = self.buf("pn.pointfeat", 32, N)
        self.pw("pn.apply_tf", p1, pointfeat, tf.data_ptr(), 32, wt_bs=32 * 32, ldw=32)
        p2 = self._c1_bn("pn.c2", f + ".conv2", f + ".bn2", pointfeat, 64, ACT_RELU)
        g = torch.empty(self.B, 128, device=self.device, dtype=torch.float32)
        self._keep.append(g)
        self._add("pn.gfill", self.lib.ach_fill, g.data_ptr(), g.numel(), float("-inf"))
        wt = self._w("pn.c3.wt", lambda: self._kmajor(self._p(f + ".conv3.weight").flatten(1)))
        sc = self._vec("pn.c3.s", lambda: self._bn_fold(f + ".bn3", 1e-5, self._p(f + ".conv3.bias"))[0])
        bi = self._vec("pn.c3.b", lambda: self._bn_fold(f + ".bn3", 1e-5, self._p(f + ".conv3.bias"))[1])
        self.pw("pn.c3max", p2, g, wt, 128, scale=sc, bias=bi, reduce_max=True)
        # head conv1 over cat(global(128) repeated, pointfeat(32)): the global half is a per-frame bias
        wg = self._w("pn.h1.wg", lambda: self._p(prefix + ".conv1.weight")[:, :128, 0])
        pb = self._fc("pn.h1.gbias", g, 128, wg, 128)
        wp = self._w("pn.h1.wt", lambda: self._kmajor(self._p(prefix + ".conv1.weight")[:, 128:, 0]))
        s1 = self._vec("pn.h1.s", lambda: self._bn_fold(prefix + ".bn1", 1e-5, self._p(prefix + ".conv1.bias"))[0])
        b1 = self._vec("pn.h1.b", lambda: self._bn_fold(prefix + ".bn1", 1e-5, self._p(prefix + ".conv1.bias"))[1])
        h1 = self.buf("pn.h1", 128, N)
        self.pw("pn.h1", pointfeat, h1, wp, 128, scale=s1, bias=b1, pbias=pb, act=ACT_RELU)
        h2 = self._c1_bn("pn.h2", prefix + ".conv2", prefix + ".bn2", h1, 100, ACT_RELU)
        h3 = self._c1_bn("pn.h3", prefix + ".conv3", prefix + ".bn3", h2, 64, ACT_RELU)
        h4 = self.buf("pn.h4", num_class, N)
        self.pw_bias("pn.h4", prefix + ".conv4", h3, h4)
        self._pc_out(h4, num_class, N)
        self.taps_raw = {"pc.trans": t3, "pc.trans_feat": tf, "pc.global": g}
        self.taps["pc.pointfeat"] = pointfeat

    # ---- PointNet++ (builder-defined, oracle/pn2.py)
    def pointnet2(self, x, prefix, num_class):
        D, N = x.C, x.H * x.W
        ibuf = lambda name, *shape: self._keep.append(torch.zeros(*shape, device=self.device, dtype=torch.int32)) or self._keep[-1]
        xyz = [self.sl(x, 0, 3)]
        pts = [x]
        n_prev = N
        for li, cfg in enumerate(Hd.PN2_SA, 1):
            S, ns = cfg["npoint"], cfg["nsample"]
            sp = f"{prefix}.sa{li}"
            fps_idx = ibuf(f"pn2.sa{li}.fps", self.B, S)
            nxyz = self.buf(f"pn2.sa{li}.xyz", 3, S)
            self._add(f"pn2.sa{li}.fps", self.lib.ach_pn2_fps, xyz[-1].ptr, xyz[-1].bs, self.B, n_prev, S, fps_idx.data_ptr(), nxyz.ptr, nxyz.bs)
            Cin = pts[-1].C
            g = self.buf(f"pn2.sa{li}.group", 3 + Cin, S * ns)
            gidx = ibuf(f"pn2.sa{li}.idx", self.B, S, ns)
            self._add(f"pn2.sa{li}.group", self.lib.ach_pn2_group, xyz[-1].ptr, xyz[-1].bs, pts[-1].ptr, pts[-1].bs, Cin, nxyz.ptr, nxyz.bs,
                      self.B, n_prev, S, ns, float(cfg["radius"]), g.ptr, g.bs, gidx.data_ptr())
            cur = g
            for i, c in enumerate(cfg["mlp"]):
                nxt = self.buf(f"pn2.sa{li}.mlp{i}", c, S * ns)
                self.pw_bn_act(f"pn2.sa{li}.mlp{i}", f"{sp}.mlp_convs.{i}", f"{sp}.mlp_bns.{i}", 1e-5, cur, nxt, ACT_RELU, conv_bias=True)
                cur = nxt
            o = self.buf(f"pn2.sa{li}.out", cfg["mlp"][-1], S)
            self._add(f"pn2.sa{li}.max", self.lib.ach_pn2_group_max, cur.ptr, cur.bs, o.ptr, o.bs, self.B, cur.C, S, ns)
            self.taps[f"pc.sa{li}.out"] = o
            self.taps_int = getattr(self, "taps_int", {})
            self.taps_int[f"pc.sa{li}.fps"], self.taps_int[f"pc.sa{li}.idx"] = fps_idx, gidx
            xyz.append(nxyz)
            pts.append(o)
            n_prev = S

        def fp(name, pfx, xyz1, xyz2, p1, p2, mlp):
            n1, S = xyz1.H * xyz1.W, xyz2.H * xyz2.W
            it = self.buf(name + ".interp", p2.C, n1)
            self._add(name + ".interp", self.lib.ach_pn2_interp3, xyz1.ptr, xyz1.bs, xyz2.ptr, xyz2.bs, p2.ptr, p2.bs, self.B, p2.C, n1, S,
                      it.ptr, it.bs)
            cur, x1 = p1, it
            for i, c in enumerate(mlp):
                nxt = self.buf(f"{name}.mlp{i}", c, n1)
                self.pw_bn_act(f"{name}.mlp{i}", f"{pfx}.mlp_convs.{i}", f"{pfx}.mlp_bns.{i}", 1e-5, cur, nxt, ACT_RELU, x1=x1, conv_bias=True)
                cur, x1 = nxt, None
            return cur
        f2 = fp("pn2.fp3", prefix + ".fp3", xyz[2], xyz[3], pts[2], pts[3], Hd.PN2_FP[3])
        f1 = fp("pn2.fp2", prefix + ".fp2", xyz[1], xyz[2], pts[1], f2, Hd.PN2_FP[2])
        f0 = fp("pn2.fp1", prefix + ".fp1", xyz[0], xyz[1], pts[0], f1, Hd.PN2_FP[1])
        self.taps.update({"pc.fp3": f2, "pc.fp2": f1, "pc.fp1": f0})
        h = self.buf("pn2.h1", 128, N)
        self.pw_bn_act("pn2.h1", prefix + ".conv1", prefix + ".bn1", 1e-5, f0, h, ACT_RELU, conv_bias=True)
        h2 = self.buf("pn2.h2", num_class, N)
        self.pw_bias("pn2.h2", prefix + ".conv2", h, h2)
        self._pc_out(h2, num_class, N)

    def _adopt_out(self, dtype, width):
        """The packed output buffer: allocated here, or the caller's rows (B, width) - e.g. this rank's slice of the all-gather
        destination, so that the gather is in place and no staging copy exists"""
        if self._out is None:
            return torch.zeros(self.B, width, device=self.device, dtype=dtype)
        o = self._out
        if (o.dtype != dtype or tuple(o.shape) != (self.B, width) or o.stride(1) != 1 or o.device != self.device or o.data_ptr() % 16
                or (o.stride(0) * o.element_size()) % 16):
            raise _lib.AchelousKernelError(f"out buffer must be a 16-byte aligned ({self.B}, {width}) {dtype} view on {self.device}")
        return o

    # ------------------------------------------------------------------ plan
    def _build(self):
        m = self.model
        B, R, N = self.B, self.res, self.n_points
        self._resolve_params()
        self.taps = {}
        self.pack_ops = []
        K, S, PC = m.num_det, m.num_seg, m.pc_classes
        h3, h4, h5 = R // 8, R // 16, R // 32
        sizes = [(5 + K) * h3 * h3, (5 + K) * h4 * h4, (5 + K) * h5 * h5, S * R * R, 2 * R * R]
        if m.has_pc:
            sizes.append(N * PC)
        self.out_offsets = [0]
        for s_ in sizes:
            self.out_offsets.append(self.out_offsets[-1] + s_)
        self.frame_elems = self.out_offsets[-1]
        self.x_in = self.buf("in.x", m.image_channels, R, R)
        self.r_in = self.buf("in.radar", m.radar_channels, R, R)
        masks = None
        if self.compact is None:
            self.packed_out = self._adopt_out(torch.float32, self.frame_elems)
            base, bs = self.packed_out.data_ptr(), self.packed_out.stride(0)

            def oview(i, C_, H, W):
                return View(base + self.out_offsets[i] * 4, bs, C_, H, W)
            det_views = [oview(0, 5 + K, h3, h3), oview(1, 5 + K, h4, h4), oview(2, 5 + K, h5, h5)]
            out_se, out_lane = oview(3, S, R, R), oview(4, 2, R, R)
            self._pc_slot = base + self.out_offsets[5] * 4 if m.has_pc else None
        else:
            # compact record, bytes per frame: [count i32 | pad | det rows fp32 (max_det, 7) | se mask u8 | lane mask u8 | pc class u8]
            cp = self.compact
            md = min(cp["max_det"], h3 * h3 + h4 * h4 + h5 * h5)
            mh, mw = cp["image_shape"] if cp["image_shape"] is not None else (R, R)
            self.rec, off = {}, 0
            for key, nbytes in (("count", 16), ("det", md * 28), ("se", mh * mw), ("lane", mh * mw), ("pc", N if m.has_pc else 0)):
                self.rec[key] = (off, nbytes)
                off = _align16(off + nbytes)
            self.rec["shape"] = (md, mh, mw)
            self.frame_bytes = off
            self.packed_out = self._adopt_out(torch.uint8, off)
            base, bs = self.packed_out.data_ptr(), self.packed_out.stride(0)
            raw_det = torch.empty(B, sum(sizes[:3]), device=self.device, dtype=torch.float32)
            self._bufs["out.det_raw"] = raw_det
            det_views = [View(raw_det.data_ptr() + self.out_offsets[i] * 4, raw_det.stride(0), 5 + K, h_, h_) for i, h_ in enumerate((h3, h4, h5))]
            out_se, out_lane = _LazyOut(self, "out.se_logits", S, R, R), _LazyOut(self, "out.lane_logits", 2, R, R)
            self._mask_slots = {"se": (base + self.rec["se"][0], mh * mw), "lane": (base + self.rec["lane"][0], mh * mw)}
            if cp["image_shape"] is None:   # network-resolution masks: the fused head writes the class byte itself
                masks = {"se": (base + self.rec["se"][0], bs, self._keep_mask(S)), "lane": (base + self.rec["lane"][0], bs, 0xFFFFFFFF)}
            self._pc_slot = base + self.rec["pc"][0] if m.has_pc else None

        if m.has_pc:
            self.pc_in = self.buf("in.pc", m.pc_channels, N, 1)
            if m.pc_seg == "pn":
                self.cur_lane = 1          # the point-cloud branch is independent of the image/radar graph
                self.wait(1, 0)
                self.pointnet(self.pc_in, "pc_seg_model", PC)
                self.cur_lane = 0
            elif m.pc_seg == "pn2":
                self.cur_lane = 1
                self.wait(1, 0)
                self.pointnet2(self.pc_in, "pc_seg_model", PC)
                self.cur_lane = 0
            else:
                raise NotImplementedError(f"pc_seg={m.pc_seg!r}")
        ire = "image_radar_encoder"
        # Stream lanes (parallel branches of the captured graph).  A lane forks from lane 0 at the point its first wait() is
        # recorded, so the order of the blocks below is the dependency structure: the radar encoder depends on nothing but its
        # input and forks at the very start (recorded after the neck it used to wait for backbone, neck AND the semantic decoder);
        # fusion + detection (+ decode + NMS) fork when the FPN maps exist and run beside the two decoders.
        self.cur_lane = 2
        self.wait(2, 0)
        radar = self.rcnet(self.r_in, ire + ".radar_encoder", m.phi)
        self.cur_lane = 0
        if m.backbone == "en":
            feats = self.edgenext(self.x_in, ire + ".fpn.backbone", m.phi)
        elif m.backbone == "ev":
            feats = self.edgevit(self.x_in, ire + ".fpn.backbone", m.phi)
        elif m.backbone == "ef":
            feats = self.efficientformer(self.x_in, ire + ".fpn.backbone", m.phi)
        else:
            feats = self.mobilevit(self.x_in, ire + ".fpn.backbone", m.phi)
        maps = (self.gdf_neck if m.neck == "gdf" else self.cdf_neck)(feats, ire + ".fpn", m.phi, out_se, out_lane, masks=masks)
        self.cur_lane = 6                  # forked inside the neck (wait(6, 0) before the decoders); joins the radar lane here
        self.wait(6, 2)
        fused = [self.fuse_stage(s, ire, maps[2 - i], radar[i]) for i, s in enumerate((3, 4, 5))]
        for k, lane in enumerate((6, 4, 5)):   # the three detection levels are independent
            if lane != 6:
                self.cur_lane = lane
                self.wait(lane, 6)
            self.det_level(k, "det_head", fused[k], det_views[k], K)
        self.cur_lane = 6
        if self.compact is not None:
            self.wait(6, 4)
            self.wait(6, 5)
            self._det_finish(det_views, K)
        self.cur_lane = 0
        self.sync_end = [(0, l) for l in sorted(set(self.op_lane)) if l]
        self._cpu64 = {}
        self._sig = self._signature()

    # ------------------------------------------------------------------ execution
    def _launch_all(self, stream=None):
        """Issues the whole plan.  Lane 0 is the caller's current stream; with multi_stream the independent
        sub-graphs (point cloud, radar, lane decoder, detection levels) go to side streams joined by events,
        which a CUDA-graph capture turns into parallel branches of the graph."""
        if self.dry_run:
            raise _lib.AchelousKernelError("dry-run engine cannot launch kernels")
        check = _lib.check
        main = torch.cuda.current_stream(self.device)
        if not self.multi_stream:
            sp = main.cuda_stream
            for (fn, args), name in zip(self.ops, self.op_names):
                st = fn(*args, sp)
                if st:
                    check(st, name)
            return
        streams = {0: main}
        for l in set(self.op_lane):
            if l:
                if l not in self._side_streams:
                    self._side_streams[l] = torch.cuda.Stream(self.device)
                streams[l] = self._side_streams[l]
        ptrs = {l: s_.cuda_stream for l, s_ in streams.items()}
        sync = self.sync_before
        for i, (fn, args) in enumerate(self.ops):
            if i in sync:
                for waiter, waited in sync[i]:
                    streams[waiter].wait_stream(streams[waited])
            st = fn(*args, ptrs[self.op_lane[i]])
            if st:
                check(st, self.op_names[i])
        for waiter, waited in self.sync_end:
            streams[waiter].wait_stream(streams[waited])

    def run_packs(self, stream):
        for fn, args in self.pack_ops:
            _lib.check(fn(*args, stream), "pack op")

    def refresh_weights(self):
        """Host part of the stale-weight check: True when the packed weights were re-derived"""
        if getattr(self.model, "_weights_epoch", 0) != self._epoch:
            self._resolve_params()      # load_state_dict / invalidate(): parameter objects may have been replaced
            self._sig = None
        if self._signature() != self._sig:
            self.repack()
            return True
        return False

    def ensure_packed(self):
        if self.refresh_weights():
            self._packs_done = False
        if not getattr(self, "_packs_done", False):
            self.run_packs(torch.cuda.current_stream(self.device).cuda_stream)
            self._packs_done = True

    def forward_static(self):
        """Runs the plan on the static input buffers; results land in self.packed_out."""
        self.ensure_packed()
        if self.use_graph:
            if self.graph is None:
                self._launch_all()  # warm-up: module loading, function attributes
                torch.cuda.synchronize(self.device)
                g = torch.cuda.CUDAGraph()
                # nn.DataParallel runs one Python thread per GPU, each may be capturing its own plan: thread_local error mode, and
                # a capture stream of THIS device (torch.cuda.graph's default capture stream is one process-wide stream on
                # whichever device captured first - a replica on another GPU then records an empty graph)
                cap = torch.cuda.Stream(self.device)
                with torch.cuda.graph(g, stream=cap, capture_error_mode="thread_local"):
                    self._launch_all()
                self.graph = g
            self.graph.replay()
        else:
            self._launch_all()

    def input_tensors(self):
        t = [self._bufs["in.x"], self._bufs["in.radar"]]
        if self.model.has_pc:
            t.append(self._bufs["in.pc"].view(self.B, self.model.pc_channels, self.n_points))
        return t

    def unpack(self, po):
        """Views of one packed output buffer - this engine's `packed_out` or a host / gathered copy with the same row layout
        (any number of rows).  Raw mode: (det[3], se_seg, lane_seg, pc_seg) fp32; compact mode: CompactOutputs."""
        m = self.model
        K, S, R = m.num_det, m.num_seg, self.res
        if self.compact is not None:
            rec = self.rec
            md, mh, mw = rec["shape"]
            sl = lambda key: po[:, rec[key][0]:rec[key][0] + rec[key][1]]
            count = sl("count")[:, :4].view(torch.int32)[:, 0]
            rows = sl("det").view(torch.float32).unflatten(1, (md, 7))
            pc = sl("pc") if m.has_pc else None
            return CompactOutputs(rows, count, sl("se").unflatten(1, (mh, mw)), sl("lane").unflatten(1, (mh, mw)), pc)
        o = self.out_offsets
        det = [po[:, o[i]:o[i + 1]].unflatten(1, (5 + K, R // s, R // s)) for i, s in enumerate((8, 16, 32))]
        se = po[:, o[3]:o[4]].unflatten(1, (S, R, R))
        lane = po[:, o[4]:o[5]].unflatten(1, (2, R, R))
        pc = po[:, o[5]:o[6]].unflatten(1, (self.n_points, m.pc_classes)) if m.has_pc else None
        return det, se, lane, pc

    def output_views(self):
        return self.unpack(self.packed_out)

    def tap(self, name):
        """Intermediate activation by oracle tap name (tests only)."""
        v = self.taps[name]
        for t in self._bufs.values():
            if t.data_ptr() == v.ptr:
                return t
        raise KeyError(name)
