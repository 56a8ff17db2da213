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
               sigma_logit=1.0, seed=1234, device=None, chunk=128, style="isolated", corrupt=0.15):
    """Returns dict(dP [B,2,H,W] f32, cellprob [B,H,W] f32, logits [B,C,H,W] f32, labels [B,H,W] i32) on device.
    style "isolated": the SURVEY 8(d) generator (cells never touch).  style "touching": the hostile variant --
    Voronoi-clipped discs around freely jittered seeds (neighbours share flat borders), `corrupt` of the cells get
    noise for flows (the flow check must remove them), every 8th tile holds a cell wider than the 30 x 32 warp
    path, every 16th a ring with background inside (hole fill)."""
    if style == "touching":
        return _make_touching(B, H, W, C, n_grid, drop, sigma_flow, sigma_prob, sigma_logit, seed, device, chunk, corrupt)
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


def _make_touching(B, H, W, C, n_grid, drop, sigma_flow, sigma_prob, sigma_logit, seed, device, chunk, corrupt):
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
    cj = torch.clamp((yy / gy).floor().long(), max=n_grid - 1)
    ci = torch.clamp((xx / gx).floor().long(), max=n_grid - 1)
    ncell = n_grid * n_grid
    jj = torch.arange(ncell, device=dev) // n_grid
    ii = torch.arange(ncell, device=dev) % n_grid
    for b0 in range(0, B, chunk):
        nb = min(chunk, B - b0)
        rnd = lambda: torch.rand((nb, ncell), generator=g, device=dev)
        # seeds jittered over the whole grid cell: neighbours come closer than their radii
        cy = (jj + 0.15 + 0.7 * rnd()) * gy
        cx = (ii + 0.15 + 0.7 * rnd()) * gx
        rad = 0.34 * min(gy, gx) + 0.22 * min(gy, gx) * rnd()                 # ~ 8.7 .. 14.3 px on the conic grid
        keep = rnd() >= drop
        tile_id = torch.arange(b0, b0 + nb, device=dev)
        bigcell = (torch.arange(ncell, device=dev) == (n_grid // 2) * n_grid + n_grid // 2).view(1, ncell)
        rad = torch.where(bigcell & (tile_id % 8 == 3).view(nb, 1), torch.full_like(rad, 21.0), rad)   # 43 px wide
        ring = (torch.arange(ncell, device=dev) == n_grid + 1).view(1, ncell) & (tile_id % 16 == 5).view(nb, 1)
        rad = torch.where(ring, torch.full_like(rad, 12.0), rad)
        cls = torch.randint(1, max(C, 2), (nb, ncell), generator=g, device=dev)
        bad = rnd() < corrupt
        best_d = torch.full((nb, H, W), 1e9, device=dev)
        best_i = torch.zeros((nb, H, W), dtype=torch.long, device=dev)
        for oj in (-1, 0, 1):
            for oi in (-1, 0, 1):                      # nearest seed among the 3 x 3 neighbouring grid cells
                nj = torch.clamp(cj + oj, 0, n_grid - 1); ni = torch.clamp(ci + oi, 0, n_grid - 1)
                idx = (nj * n_grid + ni).expand(nb, H, W).reshape(nb, -1)
                sy = torch.gather(cy, 1, idx).view(nb, H, W); sx = torch.gather(cx, 1, idx).view(nb, H, W)
                d = (yy - sy) ** 2 + (xx - sx) ** 2
                closer = d < best_d
                best_d = torch.where(closer, d, best_d); best_i = torch.where(closer, idx.view(nb, H, W), best_i)
        flat = best_i.view(nb, -1)
        pp = lambda t: torch.gather(t, 1, flat).view(nb, H, W)
        r_px = pp(rad)
        inside = (best_d <= r_px * r_px) & pp(keep.float()).bool()
        inside &= ~(pp(ring.float()).bool() & (best_d < 36.0))                # the ring's hole (radius 6)
        lab = torch.where(inside, best_i + 1, torch.zeros((), dtype=torch.long, device=dev)).to(torch.int32)
        labels[b0:b0 + nb] = lab
        mu = eng.masks_to_flows(lab.contiguous(), ncell + 2)
        flow = (5.0 * mu).float() + sigma_flow * torch.randn((nb, 2, H, W), generator=g, device=dev)
        noise = 3.0 * torch.randn((nb, 2, H, W), generator=g, device=dev)
        dP[b0:b0 + nb] = torch.where((pp(bad.float()).bool() & inside).unsqueeze(1), noise, flow)
        cellprob[b0:b0 + nb] = torch.where(inside, 6.0, -6.0) + sigma_prob * torch.randn((nb, H, W), generator=g, device=dev)
        lg = sigma_logit * torch.randn((nb, C, H, W), generator=g, device=dev)
        cls_img = torch.where(inside, pp(cls.float()).long(), torch.zeros((), dtype=torch.long, device=dev))
        lg.scatter_add_(1, cls_img.unsqueeze(1), torch.full((nb, 1, H, W), 4.0, device=dev))
        logits[b0:b0 + nb] = lg
        del mu, lg, flow, noise
    return dict(dP=dP, cellprob=cellprob, logits=logits, labels=labels)
