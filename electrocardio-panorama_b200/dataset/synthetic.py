"""Synthetic Tianchi-shaped batches for benchmarks and profiling: ECG-like segments in [0, 1], the twelve
Tianchi lead angles with 2.5 degree jitter, int64 ROIs tiling [0, L] on multiples of 4, a U(0, 1) target lead and,
optionally, V query views.  Host tensors; the caller moves them to the device."""
import math

import torch

from .tianchi import LEAD_THETA


def make_inputs(B, G, L, seed=0, V=0):
    assert L % 4 == 0
    gen = torch.Generator().manual_seed(104729 * seed + 31 * B + 7 * G + L)
    theta = torch.tensor(LEAD_THETA, dtype=torch.float32)
    x = 0.4 + 0.1 * torch.rand(B, G, L, generator=gen)
    n_beats = max(1, L // 400)                       # narrow QRS-like spikes every ~400 samples
    centres = torch.randint(0, L, (B, n_beats), generator=gen)
    t = torch.arange(L, dtype=torch.float32)
    for k in range(n_beats):
        bump = 0.4 * torch.exp(-0.5 * ((t[None] - centres[:, k:k + 1].float()) / 2.5) ** 2)      # (B, L)
        x += bump[:, None, :] * (0.5 + 0.5 * torch.rand(B, G, 1, generator=gen))
    x.clamp_(0.0, 1.0)
    jitter = 2.5 / 180.0 * math.pi                   # nef_net.yml:5, tianchi.py:77-82
    input_thetas = theta[torch.arange(G) % 12][None].repeat(B, 1, 1) + jitter * torch.randn(B, G, 2, generator=gen)
    query_theta = theta[torch.randint(0, 12, (B,), generator=gen)].clone()
    cuts = torch.sort(torch.randint(1, L // 4, (B, 6), generator=gen), dim=1).values * 4
    edges = torch.cat([torch.zeros(B, 1, dtype=torch.long), cuts, torch.full((B, 1), L, dtype=torch.long)], dim=1)
    rois = torch.stack([edges[:, :-1], edges[:, 1:]], dim=-1).contiguous()                       # tianchi.py:103-106
    out = dict(x=x, input_thetas=input_thetas, query_theta=query_theta, rois=rois,
               target=torch.rand(B, 1, L, generator=gen))
    if V > 0:
        th = torch.tensor([math.pi / 6, math.pi / 3, math.pi / 2, 2 * math.pi / 3])
        ph = torch.tensor([-math.pi, -2 * math.pi / 3, -math.pi / 3, 0.0, math.pi / 3, 2 * math.pi / 3])
        grid = torch.stack(torch.meshgrid(th, ph, indexing="ij"), dim=-1).reshape(-1, 2)
        out["rest_theta"] = grid[torch.arange(V) % grid.shape[0]][None].repeat(B, 1, 1).contiguous()
        out["rest_view"] = torch.rand(B, V, L, generator=gen)
    return out
