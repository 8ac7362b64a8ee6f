"""Parameter containers that reproduce the reference's module tree - and therefore its state-dict
keys, shapes and dtypes - exactly (strict load at achelous.py:171).

They are real torch.nn layers so that everything the reference's callers do to the model keeps
working (``weights_init`` dispatch on class names, ``named_modules`` optimizer groups, EMA deep-copies,
``.to()``/``.cuda()``), but their ``forward`` is never used: the forward pass is executed by
:mod:`achelous_b200.engine` on the sm_100a kernels.  Structure citations are next to each class.
"""
import math

import torch
import torch.nn as nn

WIDTHS = {"L": [40, 80, 192, 384], "S2": [32, 64, 144, 288], "S1": [32, 48, 120, 224], "S0": [32, 48, 96, 176]}


class Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("achelous_b200 parameter holders are not executable; call the top-level module")


# ------------------------------------------------------------------ conv utils (normal_conv.py, ghost_conv.py, dcn.py)
class DWConv(Holder):
    def __init__(self, cin, cout, k, bias=True):
        super().__init__()
        self.dconv = nn.Conv2d(cin, cin, k, 1, (k - 1) // 2, groups=cin, bias=bias)
        self.pconv = nn.Conv2d(cin, cout, 1, 1, bias=bias)


class BaseConv(Holder):
    """normal_conv.py:36-49 (BN eps 1e-3, momentum 0.03)"""

    def __init__(self, cin, cout, ksize, ds_conv=False):
        super().__init__()
        self.conv = DWConv(cin, cout, ksize, bias=False) if ds_conv else nn.Conv2d(cin, cout, ksize, 1, (ksize - 1) // 2, bias=False)
        self.bn = nn.BatchNorm2d(cout, eps=0.001, momentum=0.03)


class GhostModule(Holder):
    """ghost_conv.py:6-29"""

    def __init__(self, inp, oup):
        super().__init__()
        self.oup = oup
        init = math.ceil(oup / 2)
        self.primary_conv = nn.Sequential(nn.Conv2d(inp, init, 1, 1, 0, bias=False), nn.BatchNorm2d(init))
        self.cheap_operation = nn.Sequential(nn.Conv2d(init, init, 3, 1, 1, groups=init, bias=False), nn.BatchNorm2d(init))


class GhostBottleneck(Holder):
    """ghost_conv.py:32-70 (stride 1)"""

    def __init__(self, in_chs, mid_chs, out_chs):
        super().__init__()
        self.ghost1 = GhostModule(in_chs, mid_chs)
        self.ghost2 = GhostModule(mid_chs, out_chs)
        if in_chs == out_chs:
            self.shortcut = nn.Sequential()
        else:
            self.shortcut = nn.Sequential(nn.Conv2d(in_chs, in_chs, 3, 1, 1, groups=in_chs, bias=False), nn.BatchNorm2d(in_chs),
                                          nn.Conv2d(in_chs, out_chs, 1, 1, 0, bias=False), nn.BatchNorm2d(out_chs))


class DeformableConv2d(Holder):
    """dcn.py:6-47"""

    def __init__(self, c):
        super().__init__()
        self.offset_conv = nn.Conv2d(c, 18, 3, 1, 1, bias=True)
        self.modulator_conv = nn.Conv2d(c, 9, 3, 1, 1, bias=True)
        self.regular_conv = nn.Conv2d(c, c, 3, 1, 1, bias=False)
        for m in (self.offset_conv, self.modulator_conv):
            nn.init.constant_(m.weight, 0.0)
            nn.init.constant_(m.bias, 0.0)


# ------------------------------------------------------------------ radar encoder (RadarEncoder.py)
class RadarConv(Holder):
    def __init__(self, c):
        super().__init__()
        self.deformable_conv = DeformableConv2d(c)


class RCBlock(Holder):
    """RadarEncoder.py:44-63"""

    def __init__(self, cin, cout, down):
        super().__init__()
        self.down = down
        self.radar_conv = RadarConv(cin)
        self.weight_conv1 = nn.Conv2d(cin, cin, 1)
        self.norm = nn.BatchNorm2d(cin)
        self.weight_conv2 = nn.Conv2d(cin, cout, 3, 2, 1) if down else nn.Conv2d(cin, cout, 1)


class RCNet(Holder):
    """RadarEncoder.py:77-94"""

    def __init__(self, in_channels, phi):
        super().__init__()
        w = [c // 4 for c in WIDTHS[phi]]
        blocks = [RCBlock(in_channels, w[0], True), RCBlock(w[0], w[0], True)]
        for i in range(1, 4):
            blocks += [RCBlock(w[i - 1], w[i - 1], False), RCBlock(w[i - 1], w[i], True)]
        self.rc_blocks = nn.ModuleList(blocks)


# ------------------------------------------------------------------ EdgeNeXt (edgenext_modules/*.py)
class LayerNorm(Holder):
    """layers.py:7-18"""

    def __init__(self, dim, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.bias = nn.Parameter(torch.zeros(dim))
        self.eps = eps


class PositionalEncodingFourier(Holder):
    """layers.py:38-46"""

    def __init__(self, dim, hidden_dim=32):
        super().__init__()
        self.token_projection = nn.Conv2d(hidden_dim * 2, dim, 1)
        self.hidden_dim = hidden_dim


class ConvEncoder(Holder):
    """conv_encoder.py:8-17"""

    def __init__(self, dim, k, expan=4, ls=1e-6):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, k, padding=k // 2, groups=dim)
        self.norm = LayerNorm(dim)
        self.pwconv1 = nn.Linear(dim, expan * dim)
        self.pwconv2 = nn.Linear(expan * dim, dim)
        self.gamma = nn.Parameter(ls * torch.ones(dim))


class XCA(Holder):
    """sdta_encoder.py:151-160"""

    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads = heads
        self.temperature = nn.Parameter(torch.ones(heads, 1, 1))
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)


class SDTAEncoder(Holder):
    """sdta_encoder.py:9-37"""

    def __init__(self, dim, scales, heads, use_pos, expan=4, ls=1e-6):
        super().__init__()
        self.width = max(int(math.ceil(dim / scales)), int(math.floor(dim // scales)))
        self.nums = 1 if scales == 1 else scales - 1
        self.convs = nn.ModuleList([nn.Conv2d(self.width, self.width, 3, padding=1, groups=self.width) for _ in range(self.nums)])
        self.pos_embd = PositionalEncodingFourier(dim) if use_pos else None
        self.norm_xca = LayerNorm(dim)
        self.gamma_xca = nn.Parameter(ls * torch.ones(dim))
        self.xca = XCA(dim, heads)
        self.norm = LayerNorm(dim)
        self.pwconv1 = nn.Linear(dim, expan * dim)
        self.pwconv2 = nn.Linear(expan * dim, dim)
        self.gamma = nn.Parameter(ls * torch.ones(dim))


EDGENEXT_CFG = {  # model.py:14-66; heads default 8 (edgenext.py:14)
    "S0": dict(depths=[2, 2, 6, 2], dims=[32, 48, 96, 176], heads=[4, 4, 4, 4]),
    "S1": dict(depths=[3, 3, 9, 3], dims=[32, 48, 120, 224], heads=[4, 4, 4, 4]),
    "S2": dict(depths=[3, 3, 9, 3], dims=[32, 64, 144, 288], heads=[8, 8, 8, 8]),
}
EN_KERNELS, EN_SCALES, EN_POS = [3, 5, 7, 9], [2, 2, 3, 4], [False, True, False, False]


class EdgeNeXt(Holder):
    """edgenext.py:9-57 (classifier norm/head kept: dead parameters that must round-trip)"""

    def __init__(self, phi):
        super().__init__()
        cfg = EDGENEXT_CFG[phi]
        dims, depths, heads = cfg["dims"], cfg["depths"], cfg["heads"]
        self.cfg = cfg
        self.downsample_layers = nn.ModuleList([nn.Sequential(nn.Conv2d(3, dims[0], 4, 4), LayerNorm(dims[0]))])
        for i in range(3):
            self.downsample_layers.append(nn.Sequential(LayerNorm(dims[i]), nn.Conv2d(dims[i], dims[i + 1], 2, 2)))
        self.stages = nn.ModuleList()
        for i in range(4):
            blocks = []
            for j in range(depths[i]):
                if i > 0 and j == depths[i] - 1:
                    blocks.append(SDTAEncoder(dims[i], EN_SCALES[i], heads[i], EN_POS[i]))
                else:
                    blocks.append(ConvEncoder(dims[i], EN_KERNELS[i]))
            self.stages.append(nn.Sequential(*blocks))
        self.norm = nn.LayerNorm(dims[-1], eps=1e-6)
        self.head = nn.Linear(dims[-1], 1000)


# ------------------------------------------------------------------ MobileViT (mobilevit_modules/mobilevit.py)
def _conv_bn(cin, cout, k, stride=1):
    return nn.Sequential(nn.Conv2d(cin, cout, k, stride, 1 if k == 3 else 0, bias=False), nn.BatchNorm2d(cout), nn.SiLU())


class PreNorm(Holder):
    def __init__(self, dim, fn):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.fn = fn


class FeedForward(Holder):
    def __init__(self, dim, hidden):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, hidden), nn.SiLU(), nn.Dropout(0.0), nn.Linear(hidden, dim), nn.Dropout(0.0))


class MVAttention(Holder):
    def __init__(self, dim, heads=4, dim_head=8):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.dim_head = heads, dim_head
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Dropout(0.0))


class Transformer(Holder):
    def __init__(self, dim, depth, mlp_dim):
        super().__init__()
        self.layers = nn.ModuleList([nn.ModuleList([PreNorm(dim, MVAttention(dim)), PreNorm(dim, FeedForward(dim, mlp_dim))])
                                     for _ in range(depth)])


class MV2Block(Holder):
    """mobilevit.py:93-125"""

    def __init__(self, inp, oup, stride, expansion):
        super().__init__()
        self.stride, self.inp, self.oup = stride, inp, oup
        hidden = int(inp * expansion)
        self.hidden = hidden
        self.conv = nn.Sequential(
            nn.Conv2d(inp, hidden, 1, 1, 0, bias=False), nn.BatchNorm2d(hidden), nn.SiLU(),
            nn.Conv2d(hidden, hidden, 3, stride, 1, groups=hidden, bias=False), nn.BatchNorm2d(hidden), nn.SiLU(),
            nn.Conv2d(hidden, oup, 1, 1, 0, bias=False), nn.BatchNorm2d(oup))


class MobileViTBlock(Holder):
    """mobilevit.py:134-146"""

    def __init__(self, dim, depth, channel, mlp_dim):
        super().__init__()
        self.dim, self.depth, self.channel, self.mlp_dim = dim, depth, channel, mlp_dim
        self.conv1 = _conv_bn(channel, channel, 3)
        self.conv2 = _conv_bn(channel, dim, 1)
        self.transformer = Transformer(dim, depth, mlp_dim)
        self.conv3 = _conv_bn(dim, channel, 1)
        self.conv4 = _conv_bn(2 * channel, channel, 3)


MOBILEVIT_CFG = {  # mobilevit.py:225-240
    "S0": dict(dims=[64, 80, 96], channels=[16, 16, 32, 32, 48, 48, 96, 96, 96, 96, 176], expansion=2),
    "S1": dict(dims=[96, 120, 144], channels=[16, 32, 32, 32, 48, 48, 120, 120, 120, 120, 224], expansion=4),
    "S2": dict(dims=[144, 192, 240], channels=[16, 32, 32, 32, 64, 64, 144, 144, 144, 144, 288], expansion=4),
}


class MobileViT(Holder):
    """mobilevit.py:168-196"""

    def __init__(self, phi):
        super().__init__()
        cfg = MOBILEVIT_CFG[phi]
        dims, ch, e = cfg["dims"], cfg["channels"], cfg["expansion"]
        self.cfg = cfg
        L = [2, 4, 3]
        self.conv1 = _conv_bn(3, ch[0], 3, 2)
        self.mv2 = nn.ModuleList([MV2Block(ch[0], ch[1], 1, e), MV2Block(ch[1], ch[2], 2, e), MV2Block(ch[2], ch[3], 1, e),
                                  MV2Block(ch[2], ch[3], 1, e), MV2Block(ch[3], ch[4], 2, e), MV2Block(ch[5], ch[6], 2, e),
                                  MV2Block(ch[7], ch[8], 2, e)])
        self.mvit = nn.ModuleList([MobileViTBlock(dims[0], L[0], ch[5], int(dims[0] * 2)),
                                   MobileViTBlock(dims[1], L[1], ch[7], int(dims[1] * 4)),
                                   MobileViTBlock(dims[2], L[2], ch[9], int(dims[2] * 4))])
        self.conv2 = _conv_bn(ch[-2], ch[-1], 1)


# ------------------------------------------------------------------ neck (ghostdualfpn.py, spp.py, attention_modules)
# ------------------------------------------------------------------ EdgeViT (edgevit_modules/edgevit.py)
EDGEVIT_CFG = {
    "S0": dict(depth=[1, 1, 3, 2], dims=[32, 48, 96, 176], head_dim=20),
    "S1": dict(depth=[1, 1, 3, 1], dims=[32, 48, 120, 224], head_dim=32),
    "S2": dict(depth=[1, 2, 5, 3], dims=[32, 64, 144, 288], head_dim=32),
}
EDGEVIT_SR = [4, 2, 2, 1]


class EVMlp(Holder):
    """edgevit.py:13-29 (Linear) / :32-47 (CMlp, 1x1 convs)"""

    def __init__(self, dim, hidden, conv):
        super().__init__()
        self.fc1 = nn.Conv2d(dim, hidden, 1) if conv else nn.Linear(dim, hidden)
        self.fc2 = nn.Conv2d(hidden, dim, 1) if conv else nn.Linear(hidden, dim)


class EVGlobalSparseAttn(Holder):
    """edgevit.py:50-73"""

    def __init__(self, dim, num_heads, sr_ratio):
        super().__init__()
        self.num_heads, self.sr = num_heads, sr_ratio
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)
        if sr_ratio > 1:
            self.LocalProp = nn.ConvTranspose2d(dim, dim, sr_ratio, stride=sr_ratio, groups=dim)
            self.norm = nn.LayerNorm(dim)


class EVLocalAgg(Holder):
    """edgevit.py:101-113"""

    def __init__(self, dim):
        super().__init__()
        self.pos_embed = nn.Conv2d(dim, dim, 3, padding=1, groups=dim)
        self.norm1 = nn.BatchNorm2d(dim)
        self.conv1 = nn.Conv2d(dim, dim, 1)
        self.conv2 = nn.Conv2d(dim, dim, 1)
        self.attn = nn.Conv2d(dim, dim, 5, padding=2, groups=dim)
        self.norm2 = nn.BatchNorm2d(dim)
        self.mlp = EVMlp(dim, int(dim * 4.0), conv=True)


class EVSelfAttn(Holder):
    """edgevit.py:122-138 (norm_layer = LayerNorm eps 1e-6)"""

    def __init__(self, dim, num_heads, sr_ratio):
        super().__init__()
        self.pos_embed = nn.Conv2d(dim, dim, 3, padding=1, groups=dim)
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = EVGlobalSparseAttn(dim, num_heads, sr_ratio)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = EVMlp(dim, int(dim * 4.0), conv=False)


class EVLGLBlock(Holder):
    """edgevit.py:151-168"""

    def __init__(self, dim, num_heads, sr_ratio):
        super().__init__()
        self.LocalAgg = EVLocalAgg(dim) if sr_ratio > 1 else nn.Identity()
        self.SelfAttn = EVSelfAttn(dim, num_heads, sr_ratio)


class EVPatchEmbed(Holder):
    """edgevit.py:171-184"""

    def __init__(self, patch_size, in_chans, embed_dim):
        super().__init__()
        self.norm = nn.LayerNorm(embed_dim)
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


class EdgeViT(Holder):
    """edgevit.py:196-263"""

    def __init__(self, phi):
        super().__init__()
        cfg = EDGEVIT_CFG[phi]
        dims, depth = cfg["dims"], cfg["depth"]
        heads = [d // cfg["head_dim"] for d in dims]
        cin = [3] + dims[:3]
        for i in range(4):
            setattr(self, f"patch_embed{i + 1}", EVPatchEmbed(4 if i == 0 else 2, cin[i], dims[i]))
        for i in range(4):
            setattr(self, f"blocks{i + 1}", nn.ModuleList([EVLGLBlock(dims[i], heads[i], EDGEVIT_SR[i]) for _ in range(depth[i])]))
        self.norm = nn.BatchNorm2d(dims[-1])


# ------------------------------------------------------------------ EfficientFormerV2 "ImageEncoder" (backbone/vision/ImageEncoder.py)
EF_DEPTH = {"S0": [2, 2, 6, 4], "S1": [3, 3, 9, 6], "S2": [4, 4, 12, 8]}
EF_VIT_NUM = {"S0": 2, "S1": 2, "S2": 4}
EF_RATIOS = {
    "S0": [[4, 4], [4, 4], [4, 3, 3, 3, 4, 4], [4, 3, 3, 4]],
    "S1": [[4, 4, 4], [4, 4, 4], [4, 4, 3, 3, 3, 3, 4, 4, 4], [4, 4, 3, 3, 4, 4]],
    "S2": [[4, 4, 4, 4], [4, 4, 4, 4], [4, 4, 3, 3, 3, 3, 3, 3, 4, 4, 4, 4], [4, 4, 3, 3, 3, 3, 4, 4]],
}


def _ef_bias_table(res_q, res_k):
    """(number of distinct offsets, index table (res_q^2, res_k^2)): ImageEncoder.py:104-120 / :245-264"""
    import itertools
    step = math.ceil(res_k / res_q)
    offsets, idxs = {}, []
    for p1 in itertools.product(range(res_q), range(res_q)):
        for p2 in itertools.product(range(res_k), range(res_k)):
            off = (abs(p1[0] * step - p2[0]), abs(p1[1] * step - p2[1]))
            if off not in offsets:
                offsets[off] = len(offsets)
            idxs.append(offsets[off])
    return len(offsets), torch.LongTensor(idxs).view(res_q * res_q, res_k * res_k)


def _cbn_seq(cin, cout, k=1, stride=1, groups=1):
    return nn.Sequential(nn.Conv2d(cin, cout, k, stride, k // 2, groups=groups), nn.BatchNorm2d(cout))


class EFMlp(Holder):
    """ImageEncoder.py:342-366 (mid_conv=True)"""

    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Conv2d(dim, hidden, 1)
        self.fc2 = nn.Conv2d(hidden, dim, 1)
        self.mid = nn.Conv2d(hidden, hidden, 3, 1, 1, groups=hidden)
        self.mid_norm = nn.BatchNorm2d(hidden)
        self.norm1 = nn.BatchNorm2d(hidden)
        self.norm2 = nn.BatchNorm2d(dim)


class EFAttention4D(Holder):
    """ImageEncoder.py:63-121"""

    def __init__(self, dim, resolution, stride, key_dim=32, num_heads=8, attn_ratio=4):
        super().__init__()
        self.num_heads, self.key_dim, self.stride = num_heads, key_dim, stride
        if stride is not None:
            resolution = math.ceil(resolution / stride)
            self.stride_conv = _cbn_seq(dim, dim, 3, stride, groups=dim)
        self.resolution = resolution
        self.d = attn_ratio * key_dim
        dh = self.d * num_heads
        self.q = _cbn_seq(dim, num_heads * key_dim)
        self.k = _cbn_seq(dim, num_heads * key_dim)
        self.v = _cbn_seq(dim, dh)
        self.v_local = _cbn_seq(dh, dh, 3, 1, groups=dh)
        self.talking_head1 = nn.Conv2d(num_heads, num_heads, 1)
        self.talking_head2 = nn.Conv2d(num_heads, num_heads, 1)
        self.proj = nn.Sequential(nn.GELU(), nn.Conv2d(dh, dim, 1), nn.BatchNorm2d(dim))
        n_off, idx = _ef_bias_table(resolution, resolution)
        self.attention_biases = nn.Parameter(torch.zeros(num_heads, n_off))
        self.register_buffer("attention_bias_idxs", idx)


class EFLGQuery(Holder):
    """ImageEncoder.py:174-183"""

    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.local = nn.Sequential(nn.Conv2d(in_dim, in_dim, 3, 2, 1, groups=in_dim))
        self.proj = _cbn_seq(in_dim, out_dim)


class EFAttention4DDownsample(Holder):
    """ImageEncoder.py:193-264"""

    def __init__(self, dim, out_dim, resolution, key_dim=16, num_heads=8, attn_ratio=4):
        super().__init__()
        self.num_heads, self.key_dim, self.resolution = num_heads, key_dim, resolution
        self.d = attn_ratio * key_dim
        dh = self.d * num_heads
        self.resolution2 = math.ceil(resolution / 2)
        self.q = EFLGQuery(dim, num_heads * key_dim)
        self.k = _cbn_seq(dim, num_heads * key_dim)
        self.v = _cbn_seq(dim, dh)
        self.v_local = _cbn_seq(dh, dh, 3, 2, groups=dh)
        self.proj = nn.Sequential(nn.GELU(), nn.Conv2d(dh, out_dim, 1), nn.BatchNorm2d(out_dim))
        n_off, idx = _ef_bias_table(self.resolution2, resolution)
        self.attention_biases = nn.Parameter(torch.zeros(num_heads, n_off))
        self.register_buffer("attention_bias_idxs", idx)


class EFEmbedding(Holder):
    """ImageEncoder.py:292-339 (light=False)"""

    def __init__(self, cin, cout, asub, resolution):
        super().__init__()
        self.asub = asub
        if asub:
            self.attn = EFAttention4DDownsample(cin, cout, resolution)
            self.conv = nn.Conv2d(cin, cout, 3, 2, 1)
            self.bn = nn.BatchNorm2d(cout)
        else:
            self.proj = nn.Conv2d(cin, cout, 3, 2, 1)
            self.norm = nn.BatchNorm2d(cout)


class EFBlock(Holder):
    """FFN (ImageEncoder.py:426-449) or AttnFFN (:392-423)"""

    def __init__(self, dim, ratio, attn, resolution=None, stride=None):
        super().__init__()
        if attn:
            self.layer_scale_1 = nn.Parameter(1e-5 * torch.ones(dim, 1, 1))
        self.layer_scale_2 = nn.Parameter(1e-5 * torch.ones(dim, 1, 1))
        if attn:
            self.token_mixer = EFAttention4D(dim, resolution, stride)
        self.mlp = EFMlp(dim, int(dim * ratio))


class EfficientFormer(Holder):
    """ImageEncoder.py:488-568 (fork_feat=True)"""

    def __init__(self, phi, resolution):
        super().__init__()
        dims, depth, ratios, vit = WIDTHS[phi], EF_DEPTH[phi], EF_RATIOS[phi], EF_VIT_NUM[phi]
        self.patch_embed = nn.Sequential(nn.Conv2d(3, dims[0] // 2, 3, 2, 1), nn.BatchNorm2d(dims[0] // 2), nn.GELU(),
                                         nn.Conv2d(dims[0] // 2, dims[0], 3, 2, 1), nn.BatchNorm2d(dims[0]), nn.GELU())
        net = []
        for i in range(4):
            res = math.ceil(resolution / (2 ** (i + 2)))
            net.append(nn.Sequential(*[EFBlock(dims[i], ratios[i][j], i >= 2 and j > depth[i] - 1 - vit, res, 2 if i == 2 else None)
                                       for j in range(depth[i])]))
            if i < 3:
                net.append(EFEmbedding(dims[i], dims[i + 1], i >= 2, res))
        self.network = nn.ModuleList(net)
        for i_emb, i_layer in enumerate((0, 2, 4, 6)):
            self.add_module(f"norm{i_layer}", nn.BatchNorm2d(dims[i_emb]))


class SPPConv(Holder):
    """spp.py:27-31"""

    def __init__(self, c1, c2):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, 1, 1, 0, bias=False)
        self.bn = nn.BatchNorm2d(c2, eps=0.001, momentum=0.03)


class SPP(Holder):
    """spp.py:41-48"""

    def __init__(self, c1, c2, k=(5, 9, 13)):
        super().__init__()
        c_ = c1 // 2
        self.cv1 = SPPConv(c1, c_)
        self.cv2 = SPPConv(c_ * (len(k) + 1), c2)
        self.m = nn.ModuleList([nn.MaxPool2d(kernel_size=x, stride=1, padding=x // 2) for x in k])


class Upsample(Holder):
    """ghostdualfpn.py:28-35"""

    def __init__(self, cin, cout):
        super().__init__()
        self.upsample = nn.Sequential(BaseConv(cin, cout, 1), nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True))


class ShuffleAttention(Holder):
    """shuffle_attention.py:9-21"""

    def __init__(self, channel, G):
        super().__init__()
        self.G, self.channel = G, channel
        c = channel // (2 * G)
        self.gn = nn.GroupNorm(c, c)
        self.cweight = nn.Parameter(torch.zeros(1, c, 1, 1))
        self.cbias = nn.Parameter(torch.ones(1, c, 1, 1))
        self.sweight = nn.Parameter(torch.zeros(1, c, 1, 1))
        self.sbias = nn.Parameter(torch.ones(1, c, 1, 1))


class GhostDualFPN(Holder):
    """ghostdualfpn.py:42-151 (EN / MV backbones only)"""

    def __init__(self, num_class_seg, phi, backbone, resolution=320):
        super().__init__()
        w = WIDTHS[phi]
        if backbone == "en":
            self.backbone = EdgeNeXt(phi)
        elif backbone == "mv":
            self.backbone = MobileViT(phi)
        elif backbone == "ev":
            self.backbone = EdgeViT(phi)
        elif backbone == "ef":
            self.backbone = EfficientFormer(phi, resolution)
        else:
            raise NotImplementedError(f"backbone={backbone!r}: achelous_b200 implements 'en', 'mv', 'ev' and 'ef' (SURVEY.md §8b, §8f)")
        self.spp = SPP(w[3], w[3])
        self.upsample_5_to_4 = Upsample(w[3], w[2])
        self.ghost_5_to_4 = GhostBottleneck(w[2] * 2, w[2] * 2, w[2])
        self.upsample_4_to_3 = Upsample(w[2], w[1])
        self.ghost_4_to_3 = GhostBottleneck(w[1] * 2, w[1] * 2, w[1])
        self.stage_3_lane_seg = ShuffleAttention(w[1], 4)
        self.stage_3_semantic_seg = ShuffleAttention(w[1], 4)
        for name, k in (("lane", 2), ("se", num_class_seg)):
            setattr(self, f"{name}_seg_3_to_2", Upsample(w[1], w[1]))
            setattr(self, f"{name}_seg_ghost_3_to_2", GhostModule(w[1], w[1]))
            setattr(self, f"{name}_seg_2_to_1", Upsample(w[1], w[0]))
            setattr(self, f"{name}_seg_ghost_2_to_1", GhostModule(w[0], w[0]))
            setattr(self, f"{name}_seg_1_to_0", Upsample(w[0], w[0]))
            setattr(self, f"{name}_seg_ghost_1_to_0", GhostModule(w[0], w[0]))
            setattr(self, f"{name}_seg_head", GhostModule(w[0], k))


class Bottleneck(Holder):
    """cspdualfpn.py:42-56 (conv1 SiLU, conv2 3x3 with BaseConv's default ReLU)"""

    def __init__(self, cin, cout, shortcut=True, expansion=0.5):
        super().__init__()
        hidden = int(cout * expansion)
        self.conv1 = BaseConv(cin, hidden, 1)
        self.conv2 = BaseConv(hidden, cout, 3)
        self.use_add = shortcut and cin == cout


class CSPLayer(Holder):
    """cspdualfpn.py:59-78 (n = 1)"""

    def __init__(self, cin, cout, expansion=0.5):
        super().__init__()
        hidden = int(cout * expansion)
        self.conv1 = BaseConv(cin, hidden, 1)
        self.conv2 = BaseConv(cin, hidden, 1)
        self.conv3 = BaseConv(2 * hidden, cout, 1)
        self.m = nn.Sequential(Bottleneck(hidden, hidden, True, 1.0))


class CSPDualFPN(Holder):
    """cspdualfpn.py:81-191 (EN / MV backbones only)"""

    def __init__(self, num_class_seg, phi, backbone, resolution=320):
        super().__init__()
        w = WIDTHS[phi]
        if backbone == "en":
            self.backbone = EdgeNeXt(phi)
        elif backbone == "mv":
            self.backbone = MobileViT(phi)
        elif backbone == "ev":
            self.backbone = EdgeViT(phi)
        elif backbone == "ef":
            self.backbone = EfficientFormer(phi, resolution)
        else:
            raise NotImplementedError(f"backbone={backbone!r}: achelous_b200 implements 'en', 'mv', 'ev' and 'ef' (SURVEY.md §8b, §8f)")
        self.spp = SPP(w[3], w[3])
        self.upsample_5_to_4 = Upsample(w[3], w[2])
        self.ghost_5_to_4 = CSPLayer(w[2] * 2, w[2])
        self.upsample_4_to_3 = Upsample(w[2], w[1])
        self.ghost_4_to_3 = CSPLayer(w[1] * 2, w[1])
        self.stage_3_lane_seg = ShuffleAttention(w[1], 4)
        self.stage_3_semantic_seg = ShuffleAttention(w[1], 4)
        for name, k in (("lane", 2), ("se", num_class_seg)):
            setattr(self, f"{name}_seg_3_to_2", Upsample(w[1], w[1]))
            setattr(self, f"{name}_seg_ghost_3_to_2", Bottleneck(w[1], w[1]))
            setattr(self, f"{name}_seg_2_to_1", Upsample(w[1], w[0]))
            setattr(self, f"{name}_seg_ghost_2_to_1", Bottleneck(w[0], w[0]))
            setattr(self, f"{name}_seg_1_to_0", Upsample(w[0], w[0]))
            setattr(self, f"{name}_seg_ghost_1_to_0", Bottleneck(w[0], w[0]))
            setattr(self, f"{name}_seg_head", Bottleneck(w[0], k))


class ECA(Holder):
    """eca.py:5-14"""

    def __init__(self, channel, b=1, gamma=2):
        super().__init__()
        k = int(abs((math.log(channel, 2) + b) / gamma))
        k = k if k % 2 else k + 1
        self.conv = nn.Conv1d(1, 1, kernel_size=k, padding=(k - 1) // 2, bias=False)


class IREncoder(Holder):
    """IREncoder.py:26-70"""

    def __init__(self, num_class_seg, phi, backbone, neck, radar_channels=3, resolution=320):
        super().__init__()
        if neck not in ("gdf", "cdf"):
            raise NotImplementedError(f"neck={neck!r}: achelous_b200 implements 'gdf' and 'cdf' (SURVEY.md §8b, §8f rank 3)")
        w = WIDTHS[phi]
        self.fpn = (GhostDualFPN if neck == "gdf" else CSPDualFPN)(num_class_seg, phi, backbone, resolution)
        self.radar_encoder = RCNet(radar_channels, phi)
        for s, c in zip((3, 4, 5), w[1:]):
            setattr(self, f"channel_attn_stage{s}", nn.ModuleList([ECA(c), ECA(c // 4)]))
            setattr(self, f"norm_stage{s}", nn.BatchNorm2d(c * 5 // 4))
            setattr(self, f"act_stage{s}", nn.ReLU(inplace=True))


# ------------------------------------------------------------------ head (decouplehead.py)
class DecoupleHead(Holder):
    """decouplehead.py:16-56"""

    def __init__(self, num_classes, phi, nano_head=True):
        super().__init__()
        cin = [c * 5 // 4 for c in WIDTHS[phi][1:]]
        base = 64 if nano_head else 256
        self.cls_convs, self.reg_convs = nn.ModuleList(), nn.ModuleList()
        self.cls_preds, self.reg_preds, self.obj_preds, self.stems = nn.ModuleList(), nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        for c in cin:
            self.stems.append(BaseConv(c, base, 1))
            self.cls_convs.append(nn.Sequential(BaseConv(base, base, 5, ds_conv=True), BaseConv(base, base, 5, ds_conv=True)))
            self.cls_preds.append(nn.Conv2d(base, num_classes, 1))
            self.reg_convs.append(nn.Sequential(BaseConv(base, base, 5, ds_conv=True), BaseConv(base, base, 5, ds_conv=True)))
            self.reg_preds.append(nn.Conv2d(base, 4, 1))
            self.obj_preds.append(nn.Conv2d(base, 1, 1))


# ------------------------------------------------------------------ PointNet (pointnet_utils.py, pointnet_sem_seg.py)
class STN(Holder):
    """STN3d / STNkd: pointnet_utils.py:10-25, 48-65"""

    def __init__(self, cin, k):
        super().__init__()
        self.k = k
        self.conv1 = nn.Conv1d(cin, 64, 1)
        self.conv2 = nn.Conv1d(64, 128, 1)
        self.conv3 = nn.Conv1d(128, 1024, 1)
        self.fc1 = nn.Linear(1024, 512)
        self.fc2 = nn.Linear(512, 256)
        self.fc3 = nn.Linear(256, k * k)
        self.bn1, self.bn2, self.bn3 = nn.BatchNorm1d(64), nn.BatchNorm1d(128), nn.BatchNorm1d(1024)
        self.bn4, self.bn5 = nn.BatchNorm1d(512), nn.BatchNorm1d(256)


class PointNetEncoder(Holder):
    """pointnet_utils.py:88-101"""

    def __init__(self, channel):
        super().__init__()
        self.stn = STN(channel, 3)
        self.conv1 = nn.Conv1d(channel, 32, 1)
        self.conv2 = nn.Conv1d(32, 64, 1)
        self.conv3 = nn.Conv1d(64, 128, 1)
        self.bn1, self.bn2, self.bn3 = nn.BatchNorm1d(32), nn.BatchNorm1d(64), nn.BatchNorm1d(128)
        self.fstn = STN(32, 32)


class PointNet_SEG(Holder):
    """pointnet_sem_seg.py:13-24"""

    def __init__(self, num_class, point_cloud_channels):
        super().__init__()
        self.k = num_class
        self.feat = PointNetEncoder(point_cloud_channels)
        self.conv1 = nn.Conv1d(160, 128, 1)
        self.conv2 = nn.Conv1d(128, 100, 1)
        self.conv3 = nn.Conv1d(100, 64, 1)
        self.conv4 = nn.Conv1d(64, self.k, 1)
        self.bn1, self.bn2, self.bn3 = nn.BatchNorm1d(128), nn.BatchNorm1d(100), nn.BatchNorm1d(64)


# ------------------------------------------------------------------ PointNet++ (builder-defined: the reference has no PN2 code)
PN2_SA = [dict(npoint=128, radius=0.04, nsample=16, mlp=[32, 32, 64]),
          dict(npoint=32, radius=0.08, nsample=16, mlp=[64, 64, 128]),
          dict(npoint=8, radius=0.16, nsample=8, mlp=[128, 128, 256])]
PN2_FP = {3: [256, 128], 2: [128, 128], 1: [128, 128]}


class PN2SetAbstraction(Holder):
    def __init__(self, in_channel, mlp):
        super().__init__()
        self.mlp_convs, self.mlp_bns = nn.ModuleList(), nn.ModuleList()
        last = in_channel
        for c in mlp:
            self.mlp_convs.append(nn.Conv2d(last, c, 1))
            self.mlp_bns.append(nn.BatchNorm2d(c))
            last = c


class PN2FeaturePropagation(Holder):
    def __init__(self, in_channel, mlp):
        super().__init__()
        self.mlp_convs, self.mlp_bns = nn.ModuleList(), nn.ModuleList()
        last = in_channel
        for c in mlp:
            self.mlp_convs.append(nn.Conv1d(last, c, 1))
            self.mlp_bns.append(nn.BatchNorm1d(c))
            last = c


class PointNet2_SEG(Holder):
    """Same I/O contract as PointNet_SEG; architecture defined in oracle/pn2.py (parity unpinned)."""

    def __init__(self, num_class, point_cloud_channels):
        super().__init__()
        self.k, D = num_class, point_cloud_channels
        cin = D
        outs = []
        for i, cfg in enumerate(PN2_SA, 1):
            setattr(self, f"sa{i}", PN2SetAbstraction(cin + 3, cfg["mlp"]))
            cin = cfg["mlp"][-1]
            outs.append(cin)
        self.fp3 = PN2FeaturePropagation(outs[2] + outs[1], PN2_FP[3])
        self.fp2 = PN2FeaturePropagation(PN2_FP[3][-1] + outs[0], PN2_FP[2])
        self.fp1 = PN2FeaturePropagation(PN2_FP[2][-1] + D, PN2_FP[1])
        self.conv1 = nn.Conv1d(PN2_FP[1][-1], 128, 1)
        self.bn1 = nn.BatchNorm1d(128)
        self.conv2 = nn.Conv1d(128, num_class, 1)
