"""Seeded synthetic inputs with the value distribution of the reference's
pre-processing: ImageNet-normalised uint8 image
(utils/utils.py:44-48), sparse min-max radar map with the +1e-13 floor
(utils/utils.py:51-54; radar_feature_map_generate.ipynb cell 6) and 512 points
sampled with replacement, columns L2-normalised (achelous.py:224,240)."""
import torch

_MEAN = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
_STD = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)


def make_inputs(batch: int, seed: int = 1234, resolution: int = 320, n_points: int = 512, pc_channels: int = 5):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    H = W = resolution
    img = torch.randint(0, 256, (batch, 3, H, W), generator=g, dtype=torch.int32).float() / 255.0
    x = ((img - _MEAN) / _STD).contiguous()

    x_radar = torch.zeros(batch, 3, H, W)
    for b in range(batch):
        n = int(torch.randint(50, 501, (1,), generator=g))
        rows = torch.randint(0, H, (n,), generator=g)
        cols = torch.randint(0, W, (n,), generator=g)
        vals = torch.rand(3, n, generator=g)
        x_radar[b, :, rows, cols] = vals
    x_radar = (x_radar + 1e-13).contiguous()

    base = torch.randn(batch, n_points, pc_channels, generator=g)
    idx = torch.randint(0, n_points, (batch, n_points), generator=g)  # sampling with replacement
    pts = torch.gather(base, 1, idx.unsqueeze(-1).expand(-1, -1, pc_channels))
    pts = pts / pts.norm(dim=1, keepdim=True).clamp_min(1e-12)
    x_pc = pts.permute(0, 2, 1).contiguous()
    return x, x_radar, x_pc
