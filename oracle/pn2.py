"""BUILDER-DEFINED PointNet++ (PN2) point-cloud segmentation oracle - PARITY UNPINNED.

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.

The reference advertises ``pc_seg='pn2'`` (README.md:63,82-83) but contains no PointNet++ code
(``nets/Achelous.py:31-32`` only builds ``PointNet_SEG``).  BASELINE config 4 names
EN-GDF-PN2-S2, so this file DEFINES the network the CUDA path is checked against.  It follows the public
PointNet++ semantic-segmentation recipe (Qi et al. 2017; the yanx27/Pointnet_Pointnet2_pytorch project the
reference's PointNet files derive from): set abstraction = farthest-point sampling -> ball query -> group
(relative xyz first, then features) -> shared MLP (Conv2d 1x1 + BN + ReLU) -> max over the group;
feature propagation = 3-NN inverse-squared-distance interpolation -> concat(skip, interpolated) ->
Conv1d + BN + ReLU; head Conv1d+BN+ReLU -> Conv1d -> log_softmax.  It keeps PointNet_SEG's I/O contract
(B, D, N) -> (B, N, K) log-probabilities (pointnet_sem_seg.py:26-37), narrowed for N = 512:

    SA1 128 pts, r 0.04, 16 nbrs, [32, 32, 64]      FP3 256+128 -> [256, 128]
    SA2  32 pts, r 0.08, 16 nbrs, [64, 64, 128]     FP2 128+64  -> [128, 128]
    SA3   8 pts, r 0.16,  8 nbrs, [128, 128, 256]   FP1 128+D   -> [128, 128]    head 128 -> 128 -> K

Determinism (needed for bit-comparable indices): FPS starts at index 0 (the public code draws a random
start), ties in arg-max / nearest-neighbour selection go to the lowest index, and every squared distance is
evaluated as ((dx*dx + dy*dy) + dz*dz) in fp32 without fused multiply-add.
"""
import torch
import torch.nn.functional as F

PN2_SA = [dict(npoint=128, radius=0.04, nsample=16, mlp=[32, 32, 64]),
          dict(npoint=32, radius=0.08, nsample=16, mlp=[64, 64, 128]),
          dict(npoint=8, radius=0.16, nsample=8, mlp=[128, 128, 256])]
PN2_FP = {3: [256, 128], 2: [128, 128], 1: [128, 128]}


def sqdist(a, b):
    """a (B, S, 3), b (B, N, 3) -> (B, S, N), ((dx*dx + dy*dy) + dz*dz) in fp32."""
    d = a[:, :, None, :] - b[:, None, :, :]
    return (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]


def farthest_point_sample(xyz, npoint):
    """xyz (B, N, 3) -> (B, npoint) indices; start index 0; first max on ties."""
    B, N, _ = xyz.shape
    idx = torch.zeros(B, npoint, dtype=torch.long)
    dist = torch.full((B, N), 1e10)
    far = torch.zeros(B, dtype=torch.long)
    ar = torch.arange(B)
    for i in range(npoint):
        idx[:, i] = far
        c = xyz[ar, far][:, None, :]
        d = sqdist(c, xyz)[:, 0]
        dist = torch.minimum(dist, d)
        far = dist.argmax(-1)
    return idx


def ball_query(radius, nsample, xyz, new_xyz):
    """first `nsample` indices (ascending) with d2 <= r2, padded with the first hit -> (B, S, nsample)."""
    B, N, _ = xyz.shape
    d2 = sqdist(new_xyz, xyz)
    gi = torch.arange(N).view(1, 1, N).repeat(B, new_xyz.shape[1], 1)
    gi[d2 > torch.tensor(radius, dtype=torch.float32) ** 2] = N
    gi = gi.sort(-1)[0][:, :, :nsample]
    first = gi[:, :, :1].expand(-1, -1, nsample)
    return torch.where(gi == N, first, gi)


def _gather(points, idx):
    """points (B, N, C), idx (B, ...) -> (B, ..., C)"""
    B = points.shape[0]
    flat = idx.reshape(B, -1)
    out = torch.gather(points, 1, flat[:, :, None].expand(-1, -1, points.shape[-1]))
    return out.reshape(*idx.shape, points.shape[-1])


def set_abstraction(xyz, points, p, cfg):
    """xyz (B, 3, N), points (B, C, N) -> new_xyz (B, 3, S), new_points (B, C', S)"""
    xyz_t, pts_t = xyz.permute(0, 2, 1), points.permute(0, 2, 1)
    fps = farthest_point_sample(xyz_t, cfg["npoint"])
    new_xyz = _gather(xyz_t, fps)                                   # (B, S, 3)
    idx = ball_query(cfg["radius"], cfg["nsample"], xyz_t, new_xyz)  # (B, S, ns)
    g_xyz = _gather(xyz_t, idx) - new_xyz[:, :, None, :]
    g = torch.cat([g_xyz, _gather(pts_t, idx)], -1).permute(0, 3, 2, 1)  # (B, 3+C, ns, S)
    for i in range(len(cfg["mlp"])):
        g = F.conv2d(g, p(f"mlp_convs.{i}.weight"), p(f"mlp_convs.{i}.bias"))
        g = F.relu(F.batch_norm(g, p(f"mlp_bns.{i}.running_mean"), p(f"mlp_bns.{i}.running_var"), p(f"mlp_bns.{i}.weight"),
                                p(f"mlp_bns.{i}.bias"), False, 0.0, 1e-5))
    return new_xyz.permute(0, 2, 1), g.max(2)[0], fps, idx


def feature_propagation(xyz1, xyz2, points1, points2, p, n_layers):
    """interpolate points2 (B, C2, S) at xyz1 (B, 3, N) -> concat(points1, interp) -> MLP -> (B, C', N)"""
    x1, x2 = xyz1.permute(0, 2, 1), xyz2.permute(0, 2, 1)
    d2 = sqdist(x1, x2)                                    # (B, N, S)
    d3, i3 = torch.sort(d2, dim=-1, stable=True)
    d3, i3 = d3[:, :, :3], i3[:, :, :3]
    rec = 1.0 / (d3 + 1e-8)
    w = rec / ((rec[..., 0] + rec[..., 1]) + rec[..., 2])[..., None]
    nb = _gather(points2.permute(0, 2, 1), i3)             # (B, N, 3, C2)
    interp = (nb[:, :, 0] * w[:, :, 0:1] + nb[:, :, 1] * w[:, :, 1:2]) + nb[:, :, 2] * w[:, :, 2:3]
    x = torch.cat([points1, interp.permute(0, 2, 1)], 1) if points1 is not None else interp.permute(0, 2, 1)
    for i in range(n_layers):
        x = F.conv1d(x, p(f"mlp_convs.{i}.weight"), p(f"mlp_convs.{i}.bias"))
        x = F.relu(F.batch_norm(x, p(f"mlp_bns.{i}.running_mean"), p(f"mlp_bns.{i}.running_var"), p(f"mlp_bns.{i}.weight"),
                                p(f"mlp_bns.{i}.bias"), False, 0.0, 1e-5))
    return x


def pointnet2_seg(x, p, taps=None):
    """(B, D, N) -> (B, N, K) log-probabilities."""
    xyz = [x[:, :3].contiguous()]
    pts = [x]
    for li, cfg in enumerate(PN2_SA, 1):
        nx, npts, fps, idx = set_abstraction(xyz[-1], pts[-1], p.sub(f"sa{li}"), cfg)
        xyz.append(nx)
        pts.append(npts)
        if taps is not None:
            taps[f"pc.sa{li}.fps"], taps[f"pc.sa{li}.idx"], taps[f"pc.sa{li}.out"] = fps, idx, npts
    f2 = feature_propagation(xyz[2], xyz[3], pts[2], pts[3], p.sub("fp3"), 2)
    f1 = feature_propagation(xyz[1], xyz[2], pts[1], f2, p.sub("fp2"), 2)
    f0 = feature_propagation(xyz[0], xyz[1], pts[0], f1, p.sub("fp1"), 2)
    if taps is not None:
        taps.update({"pc.fp3": f2, "pc.fp2": f1, "pc.fp1": f0})
    h = F.conv1d(f0, p("conv1.weight"), p("conv1.bias"))
    h = F.relu(F.batch_norm(h, p("bn1.running_mean"), p("bn1.running_var"), p("bn1.weight"), p("bn1.bias"), False, 0.0, 1e-5))
    h = F.conv1d(h, p("conv2.weight"), p("conv2.bias")).transpose(2, 1)
    return F.log_softmax(h, dim=-1)
