"""CPU restatement (plain PyTorch fp32 functional ops) of the reference's 5-task
forward, driven directly by a reference-keyed state-dict.

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  Every function cites the
reference file:line it restates (paths relative to /root/reference).  Pinned by
``tests/test_oracle_vs_golden.py`` against fixtures generated from the
unmodified reference (``tests/golden/make_golden.py``) and, when
the reference is importable (``/root/reference`` or the staged copy ``oracle/_ref``), against the reference itself
(``tests/test_oracle_vs_reference.py``).

``torchvision.ops.deform_conv2d`` (third-party, pinned torchvision==0.12.0 in
requirements.txt:17; 0.26.0 in this image) is restated from its published
algorithm in :func:`deform_conv2d_3x3`; no torchvision call is made here.
"""
import math

import torch
import torch.nn.functional as F

WIDTHS = {  # nets/Achelous.py:18-23
    "L": [40, 80, 192, 384],
    "S2": [32, 64, 144, 288],
    "S1": [32, 48, 120, 224],
    "S0": [32, 48, 96, 176],
}

EDGENEXT_CFG = {  # edgenext_modules/model.py:14-66 (heads default 8: edgenext.py:14)
    "S0": dict(depths=[2, 2, 6, 2], dims=[32, 48, 96, 176], heads=[4, 4, 4, 4]),
    "S1": dict(depths=[3, 3, 9, 3], dims=[32, 48, 120, 224], heads=[4, 4, 4, 4]),
    "S2": dict(depths=[3, 3, 9, 3], dims=[32, 64, 144, 288], heads=[8, 8, 8, 8]),
}
EDGENEXT_KERNELS = [3, 5, 7, 9]
EDGENEXT_SCALES = [2, 2, 3, 4]
EDGENEXT_POS = [False, True, False, False]

MOBILEVIT_CFG = {  # mobilevit_modules/mobilevit.py:225-240
    "S0": dict(dims=[64, 80, 96], channels=[16, 16, 32, 32, 48, 48, 96, 96, 96, 96, 176], expansion=2),
    "S1": dict(dims=[96, 120, 144], channels=[16, 32, 32, 32, 48, 48, 120, 120, 120, 120, 224], expansion=4),
    "S2": dict(dims=[144, 192, 240], channels=[16, 32, 32, 32, 64, 64, 144, 144, 144, 144, 288], expansion=4),
}


class SD:
    """Prefix view over a flat state-dict."""

    def __init__(self, sd, prefix=""):
        self.sd, self.prefix = sd, prefix

    def __call__(self, name):
        return self.sd[self.prefix + name]

    def sub(self, name):
        return SD(self.sd, self.prefix + name + ".")

    def has(self, name):
        return (self.prefix + name) in self.sd


# ----------------------------------------------------------------------------- primitives
def bn(x, p, eps):
    """Eval-mode BatchNorm (running stats)."""
    return F.batch_norm(x, p("running_mean"), p("running_var"), p("weight"), p("bias"), False, 0.0, eps)


def conv(x, p, stride=1, padding=0, groups=1):
    return F.conv2d(x, p("weight"), p("bias") if p.has("bias") else None, stride, padding, 1, groups)


def silu(x):
    return x * torch.sigmoid(x)


def base_conv(x, p, ksize, act="relu", ds_conv=False):
    """BaseConv: conv (or dw+pw) -> BN(eps 1e-3) -> act.  normal_conv.py:23-49"""
    pad = (ksize - 1) // 2
    if ds_conv:
        c = p.sub("conv")
        x = conv(x, c.sub("dconv"), 1, pad, groups=x.shape[1])
        x = conv(x, c.sub("pconv"))
    else:
        x = conv(x, p.sub("conv"), 1, pad)
    x = bn(x, p.sub("bn"), 1e-3)
    return F.relu(x) if act == "relu" else silu(x)


def ghost_module(x, p, oup, relu=True):
    """GhostModule: 1x1 primary + dw3x3 cheap, cat, slice.  ghost_conv.py:6-29"""
    x1 = bn(conv(x, p.sub("primary_conv.0")), p.sub("primary_conv.1"), 1e-5)
    if relu:
        x1 = F.relu(x1)
    x2 = bn(conv(x1, p.sub("cheap_operation.0"), 1, 1, groups=x1.shape[1]), p.sub("cheap_operation.1"), 1e-5)
    if relu:
        x2 = F.relu(x2)
    return torch.cat([x1, x2], 1)[:, :oup]


def ghost_bottleneck(x, p, mid, out):
    """GhostBottleneck (stride 1; shortcut dw3x3+BN+1x1+BN when in != out).  ghost_conv.py:32-70"""
    cin = x.shape[1]
    y = ghost_module(x, p.sub("ghost1"), mid, relu=True)
    y = ghost_module(y, p.sub("ghost2"), out, relu=False)
    if cin == out:
        s = x
    else:
        s = bn(conv(x, p.sub("shortcut.0"), 1, 1, groups=cin), p.sub("shortcut.1"), 1e-5)
        s = bn(conv(s, p.sub("shortcut.2")), p.sub("shortcut.3"), 1e-5)
    return y + s


def upsample_block(x, p):
    """Upsample: BaseConv 1x1 relu -> bilinear x2 align_corners.  ghostdualfpn.py:28-39"""
    x = base_conv(x, p.sub("upsample.0"), 1, "relu")
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)


def eca(x, p):
    """eca_block: GAP -> conv1d over channels -> sigmoid -> scale.  eca.py:5-22"""
    w = p("conv.weight")
    k = w.shape[-1]
    y = x.mean((2, 3))  # (B, C)
    y = F.conv1d(y.unsqueeze(1), w, None, 1, (k - 1) // 2).squeeze(1)
    return x * torch.sigmoid(y)[:, :, None, None]


def shuffle_attention(x, p, G=4):
    """ShuffleAttention.  shuffle_attention.py:48-72"""
    b, c, h, w = x.shape
    x = x.view(b * G, -1, h, w)
    x0, x1 = x.chunk(2, dim=1)
    xc = x0 * torch.sigmoid(p("cweight") * x0.mean((2, 3), keepdim=True) + p("cbias"))
    cg = c // (2 * G)
    xs = F.group_norm(x1, cg, p("gn.weight"), p("gn.bias"), 1e-5)
    xs = x1 * torch.sigmoid(p("sweight") * xs + p("sbias"))
    out = torch.cat([xc, xs], 1).contiguous().view(b, -1, h, w)
    out = out.reshape(b, 2, -1, h, w).permute(0, 2, 1, 3, 4).reshape(b, -1, h, w)  # channel_shuffle(out, 2)
    return out


def layer_norm_cf(x, p, eps=1e-6):
    """Channels-first LayerNorm, biased variance.  edgenext_modules/layers.py:21-26"""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return p("weight")[:, None, None] * x + p("bias")[:, None, None]


def layer_norm_cl(x, p, eps=1e-6):
    return F.layer_norm(x, (x.shape[-1],), p("weight"), p("bias"), eps)


def linear(x, p):
    return F.linear(x, p("weight"), p("bias") if p.has("bias") else None)


# ----------------------------------------------------------------------------- deformable conv
def _bilinear_zero(x, py, px):
    """x (B,C,H,W); py/px (B,H,W) float sample coords.  Each of the 4 corners contributes 0
    when it lies outside the image (torchvision deform_conv2d bilinear_interpolate)."""
    B, C, H, W = x.shape
    y0 = torch.floor(py)
    x0 = torch.floor(px)
    ly, lx = py - y0, px - x0
    hy, hx = 1 - ly, 1 - lx
    y0, x0 = y0.long(), x0.long()
    flat = x.reshape(B, C, H * W)

    def corner(yy, xx, wgt):
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        idx = (yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1)).reshape(B, 1, H_out * W_out).expand(-1, C, -1)
        v = torch.gather(flat, 2, idx).reshape(B, C, H_out, W_out)
        return v * (wgt * ok.to(x.dtype)).unsqueeze(1)

    H_out, W_out = py.shape[1], py.shape[2]
    return (corner(y0, x0, hy * hx) + corner(y0, x0 + 1, hy * lx)
            + corner(y0 + 1, x0, ly * hx) + corner(y0 + 1, x0 + 1, ly * lx))


def deform_conv2d_3x3(x, offset, mask, weight):
    """Modulated deformable conv (DCNv2), 3x3, stride 1, pad 1, dil 1, one offset group, no bias.
    Restates torchvision.ops.deform_conv2d as called from dcn.py:56-63: tap k = i*3 + j;
    offset channel 2k = dy, 2k+1 = dx; sample at (y - 1 + i + dy, x - 1 + j + dx);
    sampled value * mask[k]; out[o] = sum_c sum_k W[o,c,i,j] * sampled[c,k]."""
    B, C, H, W = x.shape
    O = weight.shape[0]
    ys = torch.arange(H, dtype=x.dtype).view(1, H, 1)
    xs = torch.arange(W, dtype=x.dtype).view(1, 1, W)
    cols = []
    for k in range(9):
        i, j = divmod(k, 3)
        py = ys - 1 + i + offset[:, 2 * k]
        px = xs - 1 + j + offset[:, 2 * k + 1]
        cols.append(_bilinear_zero(x, py, px) * mask[:, k].unsqueeze(1))
    col = torch.stack(cols, 2)  # (B, C, 9, H, W)
    return torch.einsum("ock,bckhw->bohw", weight.reshape(O, C, 9), col)


def rc_block(x, p, down):
    """RCBlock.  RadarEncoder.py:44-74, dcn.py:49-63"""
    pooled = F.avg_pool2d(x, 3, 1, 1)  # count_include_pad=True
    d = p.sub("radar_conv.deformable_conv")
    offset = conv(pooled, d.sub("offset_conv"), 1, 1)
    modulator = 2.0 * torch.sigmoid(conv(pooled, d.sub("modulator_conv"), 1, 1))
    y = deform_conv2d_3x3(pooled, offset, modulator, d("regular_conv.weight"))
    y = F.relu(bn(conv(y, p.sub("weight_conv1")), p.sub("norm"), 1e-5))
    y = x + y
    return conv(y, p.sub("weight_conv2"), 2, 1) if down else conv(y, p.sub("weight_conv2"))


def rcnet(x, p, taps=None):
    """RCNet: 8 RCBlocks; taps after blocks 3, 5, 7.  RadarEncoder.py:77-109"""
    downs = [True, True, False, True, False, True, False, True]
    feats = []
    for i, dn in enumerate(downs):
        x = rc_block(x, p.sub(f"rc_blocks.{i}"), dn)
        if taps is not None:
            taps[f"radar.block{i}"] = x
        if i > 1 and i % 2 == 1:
            feats.append(x)
    return feats


# ----------------------------------------------------------------------------- EdgeNeXt
def fourier_pos(p, H, W, hidden=32, temperature=10000.0):
    """PositionalEncodingFourier(B,H,W) for one frame -> (1, dim, H, W).  layers.py:38-59"""
    scale = 2 * math.pi
    y_embed = torch.arange(1, H + 1, dtype=torch.float32).view(H, 1).expand(H, W)
    x_embed = torch.arange(1, W + 1, dtype=torch.float32).view(1, W).expand(H, W)
    y_embed = y_embed / (float(H) + 1e-6) * scale
    x_embed = x_embed / (float(W) + 1e-6) * scale
    dim_t = torch.arange(hidden, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / hidden)
    pos_x = x_embed[:, :, None] / dim_t
    pos_y = y_embed[:, :, None] / dim_t
    pos_x = torch.stack((pos_x[:, :, 0::2].sin(), pos_x[:, :, 1::2].cos()), dim=3).flatten(2)
    pos_y = torch.stack((pos_y[:, :, 0::2].sin(), pos_y[:, :, 1::2].cos()), dim=3).flatten(2)
    pos = torch.cat((pos_y, pos_x), dim=2).permute(2, 0, 1).unsqueeze(0)
    pos = pos.to(p.sub("token_projection")("weight").dtype)   # no-op in fp32; lets the tests run this oracle in float64 as "truth"
    return conv(pos, p.sub("token_projection"))


def conv_encoder(x, p, k):
    """ConvEncoder.  conv_encoder.py:19-32"""
    y = conv(x, p.sub("dwconv"), 1, k // 2, groups=x.shape[1])
    y = y.permute(0, 2, 3, 1)
    y = layer_norm_cl(y, p.sub("norm"))
    y = linear(F.gelu(linear(y, p.sub("pwconv1"))), p.sub("pwconv2"))
    y = p("gamma") * y
    return x + y.permute(0, 3, 1, 2)


def xca(x, p, heads):
    """XCA cross-covariance attention on tokens (B,N,C).  sdta_encoder.py:162-185"""
    B, N, C = x.shape
    qkv = linear(x, p.sub("qkv")).reshape(B, N, 3, heads, C // heads).permute(2, 0, 3, 4, 1)  # (3,B,h,d,N)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = F.normalize(q, dim=-1)
    k = F.normalize(k, dim=-1)
    attn = (q @ k.transpose(-2, -1)) * p("temperature")
    attn = attn.softmax(dim=-1)
    y = (attn @ v).permute(0, 3, 1, 2).reshape(B, N, C)
    return linear(y, p.sub("proj"))


def sdta_encoder(x, p, scales, heads, use_pos):
    """SDTAEncoder.  sdta_encoder.py:39-74"""
    B, C, H, W = x.shape
    width = max(int(math.ceil(C / scales)), int(math.floor(C // scales)))
    nums = 1 if scales == 1 else scales - 1
    spx = torch.split(x, width, 1)
    outs = []
    sp = None
    for i in range(nums):
        sp = spx[i] if i == 0 else sp + spx[i]
        sp = conv(sp, p.sub(f"convs.{i}"), 1, 1, groups=width)
        outs.append(sp)
    y = torch.cat(outs + [spx[nums]], 1)
    t = y.reshape(B, C, H * W).permute(0, 2, 1)
    if use_pos:
        t = t + fourier_pos(p.sub("pos_embd"), H, W).reshape(1, C, H * W).permute(0, 2, 1)
    t = t + p("gamma_xca") * xca(layer_norm_cl(t, p.sub("norm_xca")), p.sub("xca"), heads)
    t = t.reshape(B, H, W, C)
    z = layer_norm_cl(t, p.sub("norm"))
    z = linear(F.gelu(linear(z, p.sub("pwconv1"))), p.sub("pwconv2"))
    z = p("gamma") * z
    return x + z.permute(0, 3, 1, 2)


def edgenext(x, p, phi, taps=None):
    """EdgeNeXt.forward_features -> 4 maps.  edgenext.py:73-86"""
    cfg = EDGENEXT_CFG[phi]
    depths, heads = cfg["depths"], cfg["heads"]
    feats = []
    for i in range(4):
        ds = p.sub(f"downsample_layers.{i}")
        if i == 0:
            x = layer_norm_cf(conv(x, ds.sub("0"), 4), ds.sub("1"))
        else:
            x = conv(layer_norm_cf(x, ds.sub("0")), ds.sub("1"), 2)
        if taps is not None:
            taps[f"backbone.down{i}"] = x
        for j in range(depths[i]):
            bp = p.sub(f"stages.{i}.{j}")
            is_sdta = i > 0 and j == depths[i] - 1  # global_block=[0,1,1,1]: last block of stages 1..3
            if is_sdta:
                x = sdta_encoder(x, bp, EDGENEXT_SCALES[i], heads[i], EDGENEXT_POS[i])
            else:
                x = conv_encoder(x, bp, EDGENEXT_KERNELS[i])
            if taps is not None:
                taps[f"backbone.stage{i}.{j}"] = x
        feats.append(x)
    return feats


# ----------------------------------------------------------------------------- MobileViT
def _conv_bn_silu(x, p, stride=1, padding=0):
    return silu(bn(conv(x, p.sub("0"), stride, padding), p.sub("1"), 1e-5))


def mv2_block(x, p, stride, expansion):
    """MV2Block.  mobilevit.py:93-131"""
    c = p.sub("conv")
    inp = x.shape[1]
    if expansion == 1:
        y = silu(bn(conv(x, c.sub("0"), stride, 1, groups=inp), c.sub("1"), 1e-5))
        y = bn(conv(y, c.sub("3")), c.sub("4"), 1e-5)
    else:
        y = silu(bn(conv(x, c.sub("0")), c.sub("1"), 1e-5))
        y = silu(bn(conv(y, c.sub("3"), stride, 1, groups=y.shape[1]), c.sub("4"), 1e-5))
        y = bn(conv(y, c.sub("6")), c.sub("7"), 1e-5)
    return x + y if (stride == 1 and inp == y.shape[1]) else y


def mvit_attention(x, p, heads=4, dim_head=8):
    """Attention on (B,P,N,D) tokens: softmax(QK^T * d^-0.5) V, out-proj.  mobilevit.py:48-73"""
    B, P, N, D = x.shape
    qkv = linear(x, p.sub("to_qkv")).reshape(B, P, N, 3, heads, dim_head).permute(3, 0, 1, 4, 2, 5)
    q, k, v = qkv[0], qkv[1], qkv[2]  # (B,P,h,N,d)
    attn = torch.softmax((q @ k.transpose(-1, -2)) * dim_head ** -0.5, dim=-1)
    out = (attn @ v).permute(0, 1, 3, 2, 4).reshape(B, P, N, heads * dim_head)
    return linear(out, p.sub("to_out.0"))


def mvit_block(x, p, depth):
    """MobileViTBlock (patch 2x2).  mobilevit.py:134-165"""
    y = x
    x = _conv_bn_silu(x, p.sub("conv1"), 1, 1)
    x = _conv_bn_silu(x, p.sub("conv2"))
    B, D, H, W = x.shape
    h, w = H // 2, W // 2
    # 'b d (h ph) (w pw) -> b (ph pw) (h w) d'
    t = x.reshape(B, D, h, 2, w, 2).permute(0, 3, 5, 2, 4, 1).reshape(B, 4, h * w, D)
    for l in range(depth):
        lp = p.sub(f"transformer.layers.{l}")
        t = mvit_attention(F.layer_norm(t, (D,), lp("0.norm.weight"), lp("0.norm.bias"), 1e-5), lp.sub("0.fn")) + t
        f = F.layer_norm(t, (D,), lp("1.norm.weight"), lp("1.norm.bias"), 1e-5)
        f = linear(silu(linear(f, lp.sub("1.fn.net.0"))), lp.sub("1.fn.net.3"))
        t = f + t
    x = t.reshape(B, 2, 2, h, w, D).permute(0, 5, 3, 1, 4, 2).reshape(B, D, H, W)
    x = _conv_bn_silu(x, p.sub("conv3"))
    x = torch.cat((x, y), 1)
    return _conv_bn_silu(x, p.sub("conv4"), 1, 1)


def mobilevit(x, p, phi, taps=None):
    """MobileViT.forward -> 4 maps.  mobilevit.py:198-222"""
    e = MOBILEVIT_CFG[phi]["expansion"]
    L = [2, 4, 3]
    x = _conv_bn_silu(x, p.sub("conv1"), 2, 1)
    for i, s in enumerate([1, 2, 1, 1]):
        x = mv2_block(x, p.sub(f"mv2.{i}"), s, e)
    f2 = x
    x = mvit_block(mv2_block(x, p.sub("mv2.4"), 2, e), p.sub("mvit.0"), L[0])
    f3 = x
    x = mvit_block(mv2_block(x, p.sub("mv2.5"), 2, e), p.sub("mvit.1"), L[1])
    f4 = x
    x = mvit_block(mv2_block(x, p.sub("mv2.6"), 2, e), p.sub("mvit.2"), L[2])
    f5 = _conv_bn_silu(x, p.sub("conv2"))
    if taps is not None:
        for n, f in zip("2345", (f2, f3, f4, f5)):
            taps[f"backbone.feat{n}"] = f
    return [f2, f3, f4, f5]


# ----------------------------------------------------------------------------- EdgeViT
EDGEVIT_CFG = {  # edgevit_modules/edgevit.py:320-349
    "S0": dict(depth=[1, 1, 3, 2], dims=[32, 48, 96, 176], head_dim=20),
    "S1": dict(depth=[1, 1, 3, 1], dims=[32, 48, 120, 224], head_dim=32),
    "S2": dict(depth=[1, 2, 5, 3], dims=[32, 64, 144, 288], head_dim=32),
}
EDGEVIT_SR = [4, 2, 2, 1]


def ev_patch_embed(x, p, patch):
    """PatchEmbed: conv k = s = patch, then nn.LayerNorm (eps 1e-5) over the channels of every token.  edgevit.py:171-193"""
    return layer_norm_cf(conv(x, p.sub("proj"), patch), p.sub("norm"), 1e-5)


def ev_local_agg(x, p):
    """LocalAgg.  edgevit.py:101-119"""
    C = x.shape[1]
    x = x + conv(x, p.sub("pos_embed"), 1, 1, groups=C)
    y = conv(bn(x, p.sub("norm1"), 1e-5), p.sub("conv1"))
    y = conv(conv(y, p.sub("attn"), 1, 2, groups=C), p.sub("conv2"))
    x = x + y
    y = conv(F.gelu(conv(bn(x, p.sub("norm2"), 1e-5), p.sub("mlp.fc1"))), p.sub("mlp.fc2"))
    return x + y


def ev_global_sparse_attn(t, p, H, W, heads, sr):
    """GlobalSparseAttn on tokens t (B, N, C).  edgevit.py:50-98"""
    B, N, C = t.shape
    h, w = H, W
    if sr > 1:
        t = t.transpose(1, 2).reshape(B, C, H, W)[:, :, ::sr, ::sr]        # AvgPool2d(1, sr): kernel 1, stride sr
        h, w = t.shape[2], t.shape[3]
        t = t.flatten(2).transpose(1, 2)
    d = C // heads
    qkv = F.linear(t, p("qkv.weight"), p("qkv.bias")).reshape(B, -1, 3, heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = ((q @ k.transpose(-2, -1)) * d ** -0.5).softmax(dim=-1)
    o = (attn @ v).transpose(1, 2).reshape(B, -1, C)
    if sr > 1:
        o = o.permute(0, 2, 1).reshape(B, C, h, w)
        o = F.conv_transpose2d(o, p("LocalProp.weight"), p("LocalProp.bias"), stride=sr, groups=C)
        o = o.reshape(B, C, -1).permute(0, 2, 1)
        o = F.layer_norm(o, (C,), p("norm.weight"), p("norm.bias"), 1e-5)
    return F.linear(o, p("proj.weight"), p("proj.bias"))


def ev_self_attn(x, p, heads, sr):
    """SelfAttn.  edgevit.py:122-148 (norm_layer = LayerNorm eps 1e-6, :334)"""
    B, C, H, W = x.shape
    x = x + conv(x, p.sub("pos_embed"), 1, 1, groups=C)
    t = x.flatten(2).transpose(1, 2)
    t = t + ev_global_sparse_attn(F.layer_norm(t, (C,), p("norm1.weight"), p("norm1.bias"), 1e-6), p.sub("attn"), H, W, heads, sr)
    m = F.linear(F.gelu(F.linear(F.layer_norm(t, (C,), p("norm2.weight"), p("norm2.bias"), 1e-6), p("mlp.fc1.weight"), p("mlp.fc1.bias"))),
                 p("mlp.fc2.weight"), p("mlp.fc2.bias"))
    t = t + m
    return t.transpose(1, 2).reshape(B, C, H, W)


def edgevit(x, p, phi, taps=None):
    """EdgeVit.forward_features -> 4 maps (feat2..4 are the patch embeddings, taken BEFORE their stage's blocks).
    edgevit.py:288-311"""
    cfg = EDGEVIT_CFG[phi]
    feats = []
    for i in range(4):
        dim = cfg["dims"][i]
        heads, sr = dim // cfg["head_dim"], EDGEVIT_SR[i]
        x = ev_patch_embed(x, p.sub(f"patch_embed{i + 1}"), 4 if i == 0 else 2)
        if i < 3:
            feats.append(x)
        for j in range(cfg["depth"][i]):
            bp = p.sub(f"blocks{i + 1}.{j}")
            if sr > 1:
                x = ev_local_agg(x, bp.sub("LocalAgg"))
            x = ev_self_attn(x, bp.sub("SelfAttn"), heads, sr)
            if taps is not None:
                taps[f"backbone.stage{i}.{j}"] = x
    feats.append(bn(x, p.sub("norm"), 1e-5))
    if taps is not None:
        for n, f in zip("2345", feats):
            taps[f"backbone.feat{n}"] = f
    return feats


# ----------------------------------------------------------------------------- EfficientFormerV2 "ImageEncoder" (backbone='ef')
EF_DEPTH = {"S0": [2, 2, 6, 4], "S1": [3, 3, 9, 6], "S2": [4, 4, 12, 8]}          # ImageEncoder.py:23-28
EF_VIT_NUM = {"S0": 2, "S1": 2, "S2": 4}                                           # :627-670
EF_RATIOS = {                                                                      # :30-60
    "S0": [[4, 4], [4, 4], [4, 3, 3, 3, 4, 4], [4, 3, 3, 4]],
    "S1": [[4, 4, 4], [4, 4, 4], [4, 4, 3, 3, 3, 3, 4, 4, 4], [4, 4, 3, 3, 4, 4]],
    "S2": [[4, 4, 4, 4], [4, 4, 4, 4], [4, 4, 3, 3, 3, 3, 3, 3, 4, 4, 4, 4], [4, 4, 3, 3, 3, 3, 4, 4]],
}


def _cbn(x, p, stride=1, padding=0, groups=1):
    """nn.Sequential(Conv2d, BatchNorm2d) with keys 0 / 1"""
    return bn(conv(x, p.sub("0"), stride, padding, groups), p.sub("1"), 1e-5)


def ef_mlp(x, p):
    """Mlp with mid depthwise conv.  ImageEncoder.py:342-389"""
    x = F.gelu(bn(conv(x, p.sub("fc1")), p.sub("norm1"), 1e-5))
    x = F.gelu(bn(conv(x, p.sub("mid"), 1, 1, groups=x.shape[1]), p.sub("mid_norm"), 1e-5))
    return bn(conv(x, p.sub("fc2")), p.sub("norm2"), 1e-5)


def ef_attention4d(x, p, resolution, stride, heads=8, key_dim=32, attn_ratio=4):
    """Attention4D with talking heads and relative-position biases.  ImageEncoder.py:63-160"""
    B, C, H, W = x.shape
    d = attn_ratio * key_dim
    if stride is not None:
        resolution = math.ceil(resolution / stride)
        x = _cbn(x, p.sub("stride_conv"), stride, 1, groups=C)
    N = resolution * resolution
    q = _cbn(x, p.sub("q")).flatten(2).reshape(B, heads, -1, N).permute(0, 1, 3, 2)
    k = _cbn(x, p.sub("k")).flatten(2).reshape(B, heads, -1, N)
    v = _cbn(x, p.sub("v"))
    v_local = _cbn(v, p.sub("v_local"), 1, 1, groups=v.shape[1])
    v = v.flatten(2).reshape(B, heads, -1, N).permute(0, 1, 3, 2)
    attn = (q @ k) * key_dim ** -0.5 + p("attention_biases")[:, p("attention_bias_idxs")]
    attn = conv(attn, p.sub("talking_head1")).softmax(dim=-1)
    attn = conv(attn, p.sub("talking_head2"))
    out = (attn @ v).transpose(2, 3).reshape(B, heads * d, resolution, resolution) + v_local
    if stride is not None:
        out = F.interpolate(out, scale_factor=stride, mode="bilinear")
    return bn(conv(F.gelu(out), p.sub("proj.1")), p.sub("proj.2"), 1e-5)


def ef_attention4d_down(x, p, resolution, heads=8, key_dim=16, attn_ratio=4):
    """Attention4DDownsample with the LGQuery.  ImageEncoder.py:174-289"""
    B, C, H, W = x.shape
    d = attn_ratio * key_dim
    res2 = math.ceil(resolution / 2)
    N, N2 = resolution * resolution, res2 * res2
    qp = p.sub("q")
    q = conv(x, qp.sub("local.0"), 2, 1, groups=C) + x[:, :, ::2, ::2]          # local dw conv + AvgPool2d(1, 2, 0)
    q = _cbn(q, qp.sub("proj")).flatten(2).reshape(B, heads, -1, N2).permute(0, 1, 3, 2)
    k = _cbn(x, p.sub("k")).flatten(2).reshape(B, heads, -1, N)
    v = _cbn(x, p.sub("v"))
    v_local = _cbn(v, p.sub("v_local"), 2, 1, groups=v.shape[1])
    v = v.flatten(2).reshape(B, heads, -1, N).permute(0, 1, 3, 2)
    attn = ((q @ k) * key_dim ** -0.5 + p("attention_biases")[:, p("attention_bias_idxs")]).softmax(dim=-1)
    out = (attn @ v).transpose(2, 3).reshape(B, heads * d, res2, res2) + v_local
    return bn(conv(F.gelu(out), p.sub("proj.1")), p.sub("proj.2"), 1e-5)


def efficientformer(x, p, phi, resolution=320, taps=None):
    """ImageEncoder.forward (fork_feat) -> 4 normed maps.  ImageEncoder.py:488-612"""
    depth, ratios, vit_num = EF_DEPTH[phi], EF_RATIOS[phi], EF_VIT_NUM[phi]
    pe = p.sub("patch_embed")
    x = F.gelu(bn(conv(x, pe.sub("0"), 2, 1), pe.sub("1"), 1e-5))
    x = F.gelu(bn(conv(x, pe.sub("3"), 2, 1), pe.sub("4"), 1e-5))
    feats, idx = [], 0
    for i in range(4):
        res = math.ceil(resolution / (2 ** (i + 2)))
        for j in range(depth[i]):
            bp = p.sub(f"network.{idx}.{j}")
            if i >= 2 and j > depth[i] - 1 - vit_num:
                x = x + bp("layer_scale_1") * ef_attention4d(x, bp.sub("token_mixer"), res, 2 if i == 2 else None)
            x = x + bp("layer_scale_2") * ef_mlp(x, bp.sub("mlp"))
            if taps is not None:
                taps[f"backbone.stage{i}.{j}"] = x
        feats.append(bn(x, p.sub(f"norm{idx}"), 1e-5))
        idx += 1
        if i < 3:
            ep = p.sub(f"network.{idx}")
            if i >= 2:   # asub: attention downsample + strided conv
                x = ef_attention4d_down(x, ep.sub("attn"), res) + bn(conv(x, ep.sub("conv"), 2, 1), ep.sub("bn"), 1e-5)
            else:
                x = bn(conv(x, ep.sub("proj"), 2, 1), ep.sub("norm"), 1e-5)
            idx += 1
    if taps is not None:
        for n, f in zip("2345", feats):
            taps[f"backbone.feat{n}"] = f
    return feats


# ----------------------------------------------------------------------------- neck / fusion / heads
def spp(x, p):
    """SPP(5,9,13).  neck/spp.py:41-52"""
    def cv(x, q):
        return silu(bn(conv(x, q.sub("conv")), q.sub("bn"), 1e-3))
    x = cv(x, p.sub("cv1"))
    return cv(torch.cat([x] + [F.max_pool2d(x, k, 1, k // 2) for k in (5, 9, 13)], 1), p.sub("cv2"))


def seg_decoder(x, p, name, widths, num_out, taps=None):
    """One segmentation decoder (lane or se).  ghostdualfpn.py:175-197"""
    chans = [widths[1], widths[0], widths[0]]
    for stage, c in zip(("3_to_2", "2_to_1", "1_to_0"), chans):
        x = upsample_block(x, p.sub(f"{name}_seg_{stage}"))
        x = ghost_module(x, p.sub(f"{name}_seg_ghost_{stage}"), c)
        if taps is not None:
            taps[f"neck.{name}_{stage}"] = x
    return ghost_module(x, p.sub(f"{name}_seg_head"), num_out)


def ghost_dual_fpn(x, p, phi, backbone, num_seg, taps=None):
    """GhostDualFPN.forward.  ghostdualfpn.py:156-200"""
    w = WIDTHS[phi]
    bb = p.sub("backbone")
    m2, m3, m4, m5 = {"en": edgenext, "mv": mobilevit, "ev": edgevit, "ef": lambda x_, b_, p_, t_: efficientformer(x_, b_, p_, x_.shape[-1], t_)}[backbone](x, bb, phi, taps)
    f5 = spp(m5, p.sub("spp"))
    f4 = ghost_bottleneck(torch.cat([upsample_block(f5, p.sub("upsample_5_to_4")), m4], 1),
                          p.sub("ghost_5_to_4"), w[2] * 2, w[2])
    f3 = ghost_bottleneck(torch.cat([upsample_block(f4, p.sub("upsample_4_to_3")), m3], 1),
                          p.sub("ghost_4_to_3"), w[1] * 2, w[1])
    f3_lane = shuffle_attention(f3, p.sub("stage_3_lane_seg"))
    f3_se = shuffle_attention(f3, p.sub("stage_3_semantic_seg"))
    if taps is not None:
        taps.update({"neck.spp": f5, "neck.fpn4": f4, "neck.fpn3": f3, "neck.sa_lane": f3_lane, "neck.sa_se": f3_se})
    lane = seg_decoder(f3_lane, p, "lane", w, 2, taps)
    se = seg_decoder(f3_se, p, "se", w, num_seg, taps)
    return se, lane, (f5 + m5, f4 + m4, f3 + m3)


def bottleneck(x, p, add):
    """Bottleneck: BaseConv 1x1 (SiLU) -> BaseConv 3x3 (BaseConv's default ReLU), + x when in == out.  cspdualfpn.py:42-56"""
    y = base_conv(base_conv(x, p.sub("conv1"), 1, "silu"), p.sub("conv2"), 3, "relu")
    return y + x if add else y


def csp_layer(x, p):
    """CSPLayer (n = 1, expansion 0.5, SiLU).  cspdualfpn.py:59-78"""
    x1 = base_conv(x, p.sub("conv1"), 1, "silu")
    x2 = base_conv(x, p.sub("conv2"), 1, "silu")
    x1 = bottleneck(x1, p.sub("m.0"), True)
    return base_conv(torch.cat((x1, x2), 1), p.sub("conv3"), 1, "silu")


def seg_decoder_csp(x, p, name, num_out, taps=None):
    """One segmentation decoder of the CSP neck: Upsample -> Bottleneck x3, Bottleneck head.  cspdualfpn.py:213-237"""
    for stage in ("3_to_2", "2_to_1", "1_to_0"):
        x = upsample_block(x, p.sub(f"{name}_seg_{stage}"))
        x = bottleneck(x, p.sub(f"{name}_seg_ghost_{stage}"), True)
        if taps is not None:
            taps[f"neck.{name}_{stage}"] = x
    return bottleneck(x, p.sub(f"{name}_seg_head"), x.shape[1] == num_out)


def csp_dual_fpn(x, p, phi, backbone, num_seg, taps=None):
    """CSPDualFPN.forward.  cspdualfpn.py:193-239"""
    bb = p.sub("backbone")
    m2, m3, m4, m5 = {"en": edgenext, "mv": mobilevit, "ev": edgevit, "ef": lambda x_, b_, p_, t_: efficientformer(x_, b_, p_, x_.shape[-1], t_)}[backbone](x, bb, phi, taps)
    f5 = spp(m5, p.sub("spp"))
    f4 = csp_layer(torch.cat([upsample_block(f5, p.sub("upsample_5_to_4")), m4], 1), p.sub("ghost_5_to_4"))
    f3 = csp_layer(torch.cat([upsample_block(f4, p.sub("upsample_4_to_3")), m3], 1), p.sub("ghost_4_to_3"))
    f3_lane = shuffle_attention(f3, p.sub("stage_3_lane_seg"))
    f3_se = shuffle_attention(f3, p.sub("stage_3_semantic_seg"))
    if taps is not None:
        taps.update({"neck.spp": f5, "neck.fpn4": f4, "neck.fpn3": f3, "neck.sa_lane": f3_lane, "neck.sa_se": f3_se})
    lane = seg_decoder_csp(f3_lane, p, "lane", 2, taps)
    se = seg_decoder_csp(f3_se, p, "se", num_seg, taps)
    return se, lane, (f5 + m5, f4 + m4, f3 + m3)


def ir_encoder(x, x_radar, p, phi, backbone, num_seg, taps=None, neck="gdf"):
    """IREncoder.forward.  backbone/IREncoder.py:72-91 (neck selection :33-41)"""
    fpn = {"gdf": ghost_dual_fpn, "cdf": csp_dual_fpn}[neck]
    se, lane, (map5, map4, map3) = fpn(x, p.sub("fpn"), phi, backbone, num_seg, taps)
    r3, r4, r5 = rcnet(x_radar, p.sub("radar_encoder"), taps)
    outs = []
    for s, m, r in ((3, map3, r3), (4, map4, r4), (5, map5, r5)):
        fuse = torch.cat([eca(m, p.sub(f"channel_attn_stage{s}.0")), eca(r, p.sub(f"channel_attn_stage{s}.1"))], 1)
        outs.append(F.relu(bn(fuse, p.sub(f"norm_stage{s}"), 1e-5)))
    if taps is not None:
        taps.update({"fuse.p3": outs[0], "fuse.p4": outs[1], "fuse.p5": outs[2]})
    return outs, se, lane


def decouple_head(feats, p):
    """DecoupleHead (nano, depthwise 5x5).  head/decouplehead.py:58-104"""
    outs = []
    for k, x in enumerate(feats):
        x = base_conv(x, p.sub(f"stems.{k}"), 1)
        c = base_conv(base_conv(x, p.sub(f"cls_convs.{k}.0"), 5, ds_conv=True), p.sub(f"cls_convs.{k}.1"), 5, ds_conv=True)
        r = base_conv(base_conv(x, p.sub(f"reg_convs.{k}.0"), 5, ds_conv=True), p.sub(f"reg_convs.{k}.1"), 5, ds_conv=True)
        outs.append(torch.cat([conv(r, p.sub(f"reg_preds.{k}")), conv(r, p.sub(f"obj_preds.{k}")),
                               conv(c, p.sub(f"cls_preds.{k}"))], 1))
    return outs


# ----------------------------------------------------------------------------- PointNet
def _c1(x, p):
    return F.conv1d(x, p("weight"), p("bias"))


def _bn1(x, p):
    return F.batch_norm(x, p("running_mean"), p("running_var"), p("weight"), p("bias"), False, 0.0, 1e-5)


def stn(x, p, k):
    """STN3d / STNkd trunk -> (B,k,k) transform (+identity).  pointnet_utils.py:27-45, 67-85"""
    B = x.shape[0]
    x = F.relu(_bn1(_c1(x, p.sub("conv1")), p.sub("bn1")))
    x = F.relu(_bn1(_c1(x, p.sub("conv2")), p.sub("bn2")))
    x = F.relu(_bn1(_c1(x, p.sub("conv3")), p.sub("bn3")))
    x = x.max(2)[0]
    x = F.relu(_bn1(linear(x, p.sub("fc1")), p.sub("bn4")))
    x = F.relu(_bn1(linear(x, p.sub("fc2")), p.sub("bn5")))
    x = linear(x, p.sub("fc3")) + torch.eye(k).reshape(1, k * k)
    return x.view(B, k, k)


def pointnet_seg(x, p, taps=None):
    """PointNet_SEG.forward: (B,D,N) -> (B,N,K) log-probs.
    pointnet_sem_seg.py:26-37, pointnet_utils.py:103-133"""
    B, D, N = x.shape
    f = p.sub("feat")
    trans = stn(x, f.sub("stn"), 3)
    xt = x.transpose(2, 1)
    xyz = torch.bmm(xt[:, :, :3], trans)
    xt = torch.cat([xyz, xt[:, :, 3:]], 2) if D > 3 else xyz
    x = F.relu(_bn1(_c1(xt.transpose(2, 1), f.sub("conv1")), f.sub("bn1")))
    trans_feat = stn(x, f.sub("fstn"), 32)
    x = torch.bmm(x.transpose(2, 1), trans_feat).transpose(2, 1)
    pointfeat = x
    x = F.relu(_bn1(_c1(x, f.sub("conv2")), f.sub("bn2")))
    x = _bn1(_c1(x, f.sub("conv3")), f.sub("bn3"))
    g = x.max(2, keepdim=True)[0]
    x = torch.cat([g.expand(-1, -1, N), pointfeat], 1)
    if taps is not None:
        taps.update({"pc.trans": trans, "pc.trans_feat": trans_feat, "pc.pointfeat": pointfeat, "pc.global": g})
    x = F.relu(_bn1(_c1(x, p.sub("conv1")), p.sub("bn1")))
    x = F.relu(_bn1(_c1(x, p.sub("conv2")), p.sub("bn2")))
    x = F.relu(_bn1(_c1(x, p.sub("conv3")), p.sub("bn3")))
    x = _c1(x, p.sub("conv4")).transpose(2, 1)
    return F.log_softmax(x, dim=-1)


# ----------------------------------------------------------------------------- facade
def achelous_forward(sd, x, x_radar, x_pc, phi="S0", backbone="en", num_seg=9, pc_seg="pn", taps=None, neck="gdf"):
    """Achelous.forward -> (det[3], se_seg, lane_seg, pc_seg).  nets/Achelous.py:49-53"""
    p = SD(sd)
    with torch.no_grad():
        if pc_seg == "pn":
            pc = pointnet_seg(x_pc, p.sub("pc_seg_model"), taps)
        elif pc_seg == "pn2":
            from oracle.pn2 import pointnet2_seg
            pc = pointnet2_seg(x_pc, p.sub("pc_seg_model"), taps)
        else:
            pc = None
        fpn_out, se, lane = ir_encoder(x, x_radar, p.sub("image_radar_encoder"), phi, backbone, num_seg, taps, neck)
        det = decouple_head(fpn_out, p.sub("det_head"))
    return det, se, lane, pc
