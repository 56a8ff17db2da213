"""Synthetic network outputs generated on the device (benchmark / demo inputs).

Follows SURVEY.md 8(d): one ellipse per jittered grid cell (never touching), ground-truth
flows from the product's own `masks_to_flows` kernel x5 (network scale), cellprob = +-6,
logits +4 on the cell's class, Gaussian noise on all three.  torch is used for the random
numbers and the per-pixel ellipse test; there is no checkpoint or dataset behind it.
"""
from __future__ import annotations

import math

import torch

from .engine import get_engine


def make_batch(B, H=256, W=256, C=7, n_grid=10, axes=(5.0, 9.0), drop=0.1, sigma_flow=0.5, sigma_prob=1.0,
               sigma_logit=1.0, seed=1234, device=None, chunk=128):
    """Returns dict(dP [B,2,H,W] f32, cellprob [B,H,W] f32, logits [B,C,H,W] f32, labels [B,H,W] i32) on device."""
    eng = get_engine(device)
    dev = eng.device
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    dP = torch.empty((B, 2, H, W), dtype=torch.float32, device=dev)
    cellprob = torch.empty((B, H, W), dtype=torch.float32, device=dev)
    logits = torch.empty((B, C, H, W), dtype=torch.float32, device=dev)
    labels = torch.empty((B, H, W), dtype=torch.int32, device=dev)
    gy, gx = H / n_grid, W / n_grid
    yy = torch.arange(H, device=dev, dtype=torch.float32).view(1, H, 1)
    xx = torch.arange(W, device=dev, dtype=torch.float32).view(1, 1, W)
    cj = torch.clamp((yy / gy).floor().long(), max=n_grid - 1)       # grid row of every pixel
    ci = torch.clamp((xx / gx).floor().long(), max=n_grid - 1)
    cell_of_pixel = (cj * n_grid + ci).expand(1, H, W)               # [1,H,W]
    ncell = n_grid * n_grid
    for b0 in range(0, B, chunk):
        nb = min(chunk, B - b0)
        u = lambda lo, hi: lo + (hi - lo) * torch.rand((nb, ncell), generator=g, device=dev)
        a, bb = u(*axes), u(*axes)
        th = u(0.0, math.pi)
        big = torch.maximum(a, bb)
        jy = torch.clamp(gy / 2 - big - 1.0, min=0.0)
        jx = torch.clamp(gx / 2 - big - 1.0, min=0.0)
        jj = torch.arange(ncell, device=dev) // n_grid
        ii = torch.arange(ncell, device=dev) % n_grid
        cy = (jj + 0.5) * gy + (2 * torch.rand((nb, ncell), generator=g, device=dev) - 1) * jy
        cx = (ii + 0.5) * gx + (2 * torch.rand((nb, ncell), generator=g, device=dev) - 1) * jx
        keep = torch.rand((nb, ncell), generator=g, device=dev) >= drop
        cls = torch.randint(1, max(C, 2), (nb, ncell), generator=g, device=dev)
        idx = cell_of_pixel.expand(nb, H, W).reshape(nb, -1)

        def per_pixel(t):
            return torch.gather(t, 1, idx).view(nb, H, W)
        dy = yy - per_pixel(cy)
        dx = xx - per_pixel(cx)
        ct, sn = per_pixel(torch.cos(th)), per_pixel(torch.sin(th))
        uu = (dx * ct + dy * sn) / per_pixel(a)
        vv = (-dx * sn + dy * ct) / per_pixel(bb)
        inside = ((uu * uu + vv * vv) <= 1.0) & per_pixel(keep.float()).bool()
        lab = torch.where(inside, idx.view(nb, H, W) + 1, torch.zeros((), dtype=torch.long, device=dev)).to(torch.int32)
        labels[b0:b0 + nb] = lab
        mu = eng.masks_to_flows(lab.contiguous(), ncell + 2)               # float64 [nb,2,H,W]
        dP[b0:b0 + nb] = (5.0 * mu).float() + sigma_flow * torch.randn((nb, 2, H, W), generator=g, device=dev)
        cellprob[b0:b0 + nb] = torch.where(inside, 6.0, -6.0) + sigma_prob * torch.randn((nb, H, W), generator=g, device=dev)
        lg = sigma_logit * torch.randn((nb, C, H, W), generator=g, device=dev)
        cls_img = torch.where(inside, per_pixel(cls.float()).long(), torch.zeros((), dtype=torch.long, device=dev))
        lg.scatter_add_(1, cls_img.unsqueeze(1), torch.full((nb, 1, H, W), 4.0, device=dev))
        logits[b0:b0 + nb] = lg
        del mu, lg
    return dict(dP=dP, cellprob=cellprob, logits=logits, labels=labels)
