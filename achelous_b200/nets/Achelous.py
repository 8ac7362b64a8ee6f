"""Drop-in replacement of the reference's ``nets/Achelous.py`` module surface
(``Achelous`` :26-53, ``Achelous3T`` :56-76): same constructor keywords, same parameter / buffer
names (strict ``load_state_dict`` with reference checkpoints), same ``forward`` signature and return
nesting - with the forward pass executed by the hand-written sm_100a kernels (no ATen compute, no CPU
or eager fallback: a missing CUDA library or a non-CUDA input raises).

    from achelous_b200.nets.Achelous import Achelous       # instead of `from nets.Achelous import *`
    net = Achelous(num_det=7, num_seg=9, phi='S0', resolution=320, backbone='en', neck='gdf',
                   pc_seg='pn', pc_channels=5, pc_classes=8, nano_head=True, spp=True).eval().cuda()
    det, se_seg, lane_seg, pc_seg = net(x, x_radar, x_point_clouds)
"""
import weakref

import torch
import torch.nn as nn

from . import holders as Hd

image_encoder_width = Hd.WIDTHS  # nets/Achelous.py:18-23


class _AchelousBase(nn.Module):
    has_pc = True

    def _init_common(self, num_det, num_seg, phi, image_channels, radar_channels, resolution, backbone, neck, pc_seg,
                     pc_channels, pc_classes, nano_head, spp):
        if phi not in ("S0", "S1", "S2"):
            raise NotImplementedError(f"phi={phi!r}: achelous_b200 implements S0/S1/S2")
        if not spp:
            raise NotImplementedError("spp=False (SPPF) is outside the accelerated path")
        if resolution % 64 != 0:
            # the stride-32 map must hold a multiple of 4 positions ((R/32)^2 % 4 == 0: 16-byte vector rows in every kernel);
            # the reference's constructor default (416 = 13 * 32) is therefore not runnable here - predict.py / train.py use 320
            raise ValueError(f"resolution={resolution}: achelous_b200 needs a multiple of 64 (e.g. 320, 384, 448)")
        self.num_det, self.num_seg, self.resolution = num_det, num_seg, resolution
        self.phi, self.image_channels, self.radar_channels = phi, image_channels, radar_channels
        self.backbone, self.neck, self.pc_seg = backbone, neck, pc_seg
        self.pc_channels, self.pc_classes, self.nano_head = pc_channels, pc_classes, nano_head
        self.n_points = 512
        self.use_cuda_graph = True
        self.use_tensor_cores = True   # 1x1 convs / Linears on tcgen05 (3xTF32, fp32-accurate); False: fp32 CUDA-core GEMM
        self.rc_tensor_cores = True    # RCBlock contractions on tcgen05 for C in {3, 8, 12, 16} (needs use_tensor_cores); False: SIMT kernel
        self.seg_tensor_cores = True   # chained decoder stages: the two 1x1 convs of ach_up_ghost_pw2 on tcgen05 (needs use_tensor_cores)
        self.conv3_tensor_cores = True  # dense 3x3 convs (CSP neck, MobileViT) as implicit GEMMs on tcgen05; False: conv_dense (CUDA cores)
        self.fuse_seg_decoder = True   # False: block-by-block decoder (keeps every reference intermediate)
        self.fuse_seg_chain = True     # decoder stages chained through ach_up_ghost_pw2 (no full-width maps in HBM)
        self.fuse_mlp = True           # LN -> Linear(4C) -> GELU -> Linear -> gamma -> + res as ONE tcgen05 launch (hidden tile in tensor memory)
        self.mlp_group_bytes = 0          # > 0: LN->Linear->GELU->Linear pairs run per group of frames whose 4C-wide hidden tensor fits this
        #                                   many bytes (L2-resident between the two GEMMs).  Measured on B200 at B=64: 5.12 ms whole batch,
        #                                   5.56 / 5.30 / 5.20 ms at 24 / 48 / 96 MB groups (more, smaller launches lose more than L2 hits win)
        self.mlp_group_min_tiles = 444    # ... unless a group would no longer fill one wave of resident CTAs
        self._engines = {}
        self._host_bufs = {}
        self._weights_epoch = 0
        self._origin_ref = weakref.ref(self)
        self.register_load_state_dict_post_hook(_bump_weights_epoch)

    # ---- weight ownership
    def _owner(self):
        """The module that owns the parameters.  nn.DataParallel replicas (achelous.py:176) are shallow copies whose
        `_parameters` are empty (torch/nn/parallel/replicate.py): they run on the wrapped module's packed weights, copied
        once per device, and share its engine cache (the replica's `__dict__` is a shallow copy, so `_engines` is the
        same dict object)."""
        if getattr(self, "_is_replica", False):
            owner = self._origin_ref()
            if owner is None:
                raise RuntimeError("nn.DataParallel replica outlived the module it was replicated from")
            return owner
        return self

    def invalidate(self):
        """Call after writing parameters in a way autograd's version counter cannot see (`p.data.copy_()`, `nn.init.*_(p.data)`,
        the reference's `weights_init`): the packed device weights are re-derived on the next forward.  In-place ops on the
        parameters themselves, `load_state_dict` (also `assign=True`) and `.to()/.cuda()` are detected automatically."""
        self._owner()._weights_epoch += 1

    refresh_weights = invalidate

    # ---- engine cache: one plan per (device, batch, output mode); weights are re-packed when parameters change
    def _engine(self, device, batch, n_points, slot=0, compact=None):
        from ..engine import Engine
        key = (device.index if device.index is not None else torch.cuda.current_device(), batch, n_points, slot, compact)
        eng = self._engines.get(key)
        if eng is None:
            owner = self._owner()
            eng = Engine(owner, batch, torch.device("cuda", key[0]), use_graph=owner.use_cuda_graph, n_points=n_points, compact=compact)
            self._engines[key] = eng
        return eng

    def _compact_key(self, outputs, options):
        """None for the raw fp32 outputs, else the normalised compact-output specification (engine cache key)"""
        from ..engine import compact_spec
        if outputs in (None, "raw"):
            if options:
                raise TypeError("compact-output options given with outputs='raw'")
            return None
        if outputs != "compact":
            raise ValueError(f"outputs={outputs!r}: expected 'raw' or 'compact'")
        return compact_spec(getattr(self._owner(), "compact_options", None), **options)

    def _apply(self, fn, *a, **k):
        self._engines = {}  # .to() / .cuda() / .float(): parameter storage moves, plans are rebuilt lazily
        self._weights_epoch += 1
        return super()._apply(fn, *a, **k)

    def __deepcopy__(self, memo):  # ModelEMA deep-copies the model (detection_loss.py:441)
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k == "_origin_ref":
                continue
            new.__dict__[k] = {} if k in ("_engines", "_host_bufs") else copy.deepcopy(v, memo)
        new.__dict__["_origin_ref"] = weakref.ref(new)
        return new

    def __getstate__(self):  # torch.save(model): plans, pinned buffers and the weak self-reference are per-process state
        st = self.__dict__.copy()
        st["_engines"], st["_host_bufs"] = {}, {}
        st.pop("_origin_ref", None)
        return st

    def __setstate__(self, st):
        super().__setstate__(st)
        self._origin_ref = weakref.ref(self)

    def _device_for(self, x):
        """Device the forward runs on: the parameters' device; for a DataParallel replica the device its inputs were
        scattered to."""
        owner = self._owner()
        if owner.training:
            raise NotImplementedError("achelous_b200 is inference-only: call .eval() first (no autograd kernels)")
        p = next(owner.parameters(), None)
        pdev = p.device if p is not None else torch.device("cpu")
        if pdev.type != "cuda":
            raise RuntimeError("achelous_b200 has no CPU path: move the module to a CUDA device")
        if owner is not self:
            return x.device if isinstance(x, torch.Tensor) and x.is_cuda else pdev
        if isinstance(x, torch.Tensor) and x.is_cuda and x.device != pdev:
            raise RuntimeError(f"x is on {x.device} but the module's parameters are on {pdev}")
        return pdev

    def _check(self, t, name, shape_tail):
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name} must be a torch.Tensor")
        if t.dtype != torch.float32:
            raise RuntimeError(f"{name}: expected float32, got {t.dtype}")
        if tuple(t.shape[1:]) != tuple(shape_tail):
            raise RuntimeError(f"{name}: expected shape (B, {', '.join(map(str, shape_tail))}), got {tuple(t.shape)}")

    def _check_batch(self, x, x_radar, x_pc):
        """dtype / shape validation of one batch; returns (B, n_points)"""
        R = self.resolution
        self._check(x, "x", (self.image_channels, R, R))
        self._check(x_radar, "x_radar", (self.radar_channels, R, R))
        B = x.shape[0]
        n_points = self.n_points
        if self.has_pc:
            if not isinstance(x_pc, torch.Tensor) or x_pc.dim() != 3 or x_pc.shape[1] != self.pc_channels:
                raise RuntimeError(f"x_point_clouds: expected (B, {self.pc_channels}, N), got "
                                   f"{tuple(x_pc.shape) if isinstance(x_pc, torch.Tensor) else type(x_pc)}")
            n_points = x_pc.shape[2]
            self._check(x_pc, "x_point_clouds", (self.pc_channels, n_points))
            if n_points % 4:
                raise RuntimeError("x_point_clouds: N must be a multiple of 4")
        if x_radar.shape[0] != B or (self.has_pc and x_pc.shape[0] != B):
            raise RuntimeError("batch sizes of x / x_radar / x_point_clouds differ")
        return B, n_points

    def _run(self, x, x_radar, x_pc, outputs="raw", **options):
        dev = self._device_for(x)
        B, n_points = self._check_batch(x, x_radar, x_pc)
        ckey = self._compact_key(outputs, options)
        with torch.cuda.device(dev), torch.no_grad():
            eng = self._engine(dev, B, n_points, compact=ckey)
            ins = eng.input_tensors()
            # inputs may still be on the host (achelous.py:212 never moves x_radar): copy_ handles both
            ins[0].copy_(x, non_blocking=True)
            ins[1].copy_(x_radar, non_blocking=True)
            if self.has_pc:
                ins[2].copy_(x_pc, non_blocking=True)
            eng.forward_static()
            # fresh tensors owned by the caller.  clone(), not contiguous(): with B == 1 the slices of the static output
            # buffer already count as contiguous and would be handed out aliased (the next forward would overwrite them)
            fresh = lambda t: None if t is None else t.clone(memory_format=torch.contiguous_format)
            out = eng.output_views()
            if ckey is not None:
                return type(out)(*[fresh(t) for t in out])
            det, se, lane, pc = out
            return [fresh(d) for d in det], fresh(se), fresh(lane), fresh(pc)

    # ------------------------------------------------------------------ throughput API (host batches in, host results out)
    def stream_forward(self, batches, outputs="raw", compact=False, **options):
        """Pipelined inference over an iterable of HOST batches ``(x, x_radar[, x_point_clouds])`` (pinned memory
        recommended).  Yields, per batch and in order, what ``forward`` returns for it - ``(det[3], se_seg, lane_seg,
        pc_seg)``, or a ``CompactOutputs`` record with ``compact=True`` / ``outputs="compact"`` (decode + NMS rows, uint8 class
        maps: ~0.23 MB instead of 4.6 MB per frame cross PCIe) - as views of a pinned host buffer.

        Lifetime of a yielded batch: it stays intact while the NEXT batch is requested and consumed; requesting the batch after
        that recycles its buffer (three host buffers rotate).  ``list(model.stream_forward(...))`` therefore needs a ``.clone()``
        per batch.

        Same results as ``forward`` (same kernels, same plan); what changes is the schedule: two plans (double
        buffering) and three streams overlap the host->device copy of batch i+1 and the device->host copy of
        batch i-1 with the kernels of batch i, so a PCIe-bound caller sees max(copy, compute) per batch instead of
        their sum.  The reference has no equivalent (achelous.py:244-266 copies, runs and reads back serially)."""
        if compact:
            outputs = "compact"
        dev = self._device_for(None)
        ckey = self._compact_key(outputs, options)
        NH = 3   # host buffers in rotation
        with torch.cuda.device(dev), torch.no_grad():
            main = torch.cuda.current_stream(dev)
            h2d, d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            engines, ev = [None, None], [dict(), dict()]
            pending = []   # (engine, host buffer, copied event) of batches not handed out yet, oldest first
            host_free = {}  # host-buffer id -> event of the last D2H into it (a buffer is rewritten two requests after its yield)

            def host_buffer(eng, i):
                hk = (dev.index, eng.B, eng.n_points, ckey, i % NH)
                hb = self._host_bufs.get(hk)
                if hb is None:      # pinning 300 MB costs ~100 ms: once per (device, batch, mode, rotation slot)
                    hb = torch.empty(eng.packed_out.shape, dtype=eng.packed_out.dtype).pin_memory()
                    self._host_bufs[hk] = hb
                return hb

            def stage_in(slot, batch):
                """Binds the slot's engine and enqueues the host->device copies of `batch`.  Called one batch AHEAD of its
                compute: issued only when the consumer asks for the next result, the copy of batch i+1 started one D2H after
                compute(i-1) ended and compute(i+1) then waited ~2.6 ms for its inputs."""
                x, xr = batch[0], batch[1]
                xp = batch[2] if self.has_pc else None
                B, n_points = self._check_batch(x, xr, xp)
                if engines[slot] is None or engines[slot].B != B or engines[slot].n_points != n_points:
                    engines[slot] = self._engine(dev, B, n_points, slot=slot, compact=ckey)
                    engines[slot].ensure_packed()
                ins = engines[slot].input_tensors()
                if "done" in ev[slot]:
                    h2d.wait_event(ev[slot]["done"])      # the previous compute on this slot no longer reads its inputs
                with torch.cuda.stream(h2d):
                    ins[0].copy_(x, non_blocking=True)
                    ins[1].copy_(xr, non_blocking=True)
                    if self.has_pc:
                        ins[2].copy_(xp, non_blocking=True)
                    ev[slot]["in"] = h2d.record_event()

            def hand_out():
                eng, hb, copied = pending.pop(0)
                copied.synchronize()
                return eng.unpack(hb)

            it = iter(batches)
            nxt = next(it, None)
            i = 0
            staged = False
            while nxt is not None:
                slot = i & 1
                if not staged:
                    stage_in(slot, nxt)
                eng = engines[slot]
                main.wait_event(ev[slot]["in"])
                if "copied" in ev[slot]:
                    main.wait_event(ev[slot]["copied"])   # the previous D2H out of this slot's device buffer has finished
                eng.forward_static()
                ev[slot]["done"] = main.record_event()
                hb = host_buffer(eng, i)
                # the host buffer's previous occupant (batch i - NH) was handed out at least two requests ago
                assert all(p[1] is not hb for p in pending)
                d2h.wait_event(ev[slot]["done"])
                with torch.cuda.stream(d2h):
                    hb.copy_(eng.packed_out, non_blocking=True)
                    ev[slot]["copied"] = d2h.record_event()
                pending.append((eng, hb, ev[slot]["copied"]))
                # stage the NEXT batch's inputs now, while this batch computes
                nxt = next(it, None)
                staged = False
                if nxt is not None:
                    stage_in((i + 1) & 1, nxt)
                    staged = True
                # hand out the previous batch while this one is in flight
                if len(pending) == 2:
                    yield hand_out()
                i += 1
            while pending:
                yield hand_out()


def _bump_weights_epoch(module, incompatible_keys):
    """load_state_dict post-hook: parameter contents (or, with assign=True, the Parameter objects) changed"""
    module._weights_epoch += 1


class Achelous(_AchelousBase):
    """nets/Achelous.py:26-53"""

    def __init__(self, num_det, num_seg, phi='S0', image_channels=3, radar_channels=3, resolution=416,
                 backbone='ef', neck='gdf', pc_seg='pn', pc_channels=6, pc_classes=9, nano_head=False, spp=True):
        super().__init__()
        self._init_common(num_det, num_seg, phi, image_channels, radar_channels, resolution, backbone, neck, pc_seg,
                          pc_channels, pc_classes, nano_head, spp)
        if pc_seg == 'pn':
            self.pc_seg_model = Hd.PointNet_SEG(num_class=pc_classes, point_cloud_channels=pc_channels)
        elif pc_seg == 'pn2':
            # the reference advertises pn2 but ships no code for it: builder-defined network
            self.pc_seg_model = Hd.PointNet2_SEG(num_class=pc_classes, point_cloud_channels=pc_channels)
        else:
            raise NotImplementedError(f"pc_seg={pc_seg!r}: implemented: 'pn', 'pn2'")
        self.image_radar_encoder = Hd.IREncoder(num_class_seg=num_seg, phi=phi, backbone=backbone, neck=neck,
                                                radar_channels=radar_channels, resolution=resolution)
        self.det_head = Hd.DecoupleHead(num_classes=num_det, phi=phi, nano_head=nano_head)

    def forward(self, x, x_radar, x_point_clouds, outputs="raw", **compact_options):
        """outputs="raw" (default): the reference's return value.  outputs="compact": a CompactOutputs record (engine.py) -
        NMS rows, uint8 argmax maps and point classes computed inside the same launch plan; options: conf_thres, nms_thres,
        max_det, keep_classes (achelous.py:297 keeps (0, 8)), image_shape=(h, w) for masks at the original image size."""
        return self._run(x, x_radar, x_point_clouds, outputs, **compact_options)


class Achelous3T(_AchelousBase):
    """nets/Achelous.py:56-76 (no point-cloud branch)"""
    has_pc = False

    def __init__(self, num_det, num_seg, phi='S0', image_channels=3, radar_channels=3, resolution=320,
                 backbone='en', neck='gdf', pc_seg='pn', pc_channels=6, pc_classes=9, nano_head=True, spp=True):
        super().__init__()
        self._init_common(num_det, num_seg, phi, image_channels, radar_channels, resolution, backbone, neck, pc_seg,
                          pc_channels, pc_classes, nano_head, spp)
        self.image_radar_encoder = Hd.IREncoder(num_class_seg=num_seg, phi=phi, backbone=backbone, neck=neck,
                                                radar_channels=radar_channels, resolution=resolution)
        self.det_head = Hd.DecoupleHead(num_classes=num_det, phi=phi, nano_head=nano_head)

    def forward(self, x, x_radar, outputs="raw", **compact_options):
        out = self._run(x, x_radar, None, outputs, **compact_options)
        if outputs == "compact":
            return out
        det, se, lane, _ = out
        return det, se, lane
