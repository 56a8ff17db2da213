"""Study (test infrastructure, CPU): how many DISTINCT float32 positions remain among the trajectories of a tile
after t Euler steps, and how much Euler work a given pool size / merge schedule leaves.  These numbers size
k_follow_pool.  Uses the oracle's follow_flows arithmetic (torch-CPU grid_sample).

    python tests/studies/trajectory_merging.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import synth  # noqa: E402


def trajectories(seed, niter=200):
    tile = synth.make_tile(seed)
    dP, cp = tile["dP"], tile["cellprob"]
    H, W = cp.shape
    fg = cp > 0
    d = dP * fg / 5.0
    ys, xs = np.nonzero(fg)
    n = len(ys)
    pt = torch.zeros((1, 1, n, 2)); im = torch.zeros((1, 2, H, W))
    pt[0, 0, :, 0] = torch.from_numpy(xs).float(); pt[0, 0, :, 1] = torch.from_numpy(ys).float()
    im[0, 0] = torch.from_numpy(d[1]); im[0, 1] = torch.from_numpy(d[0])
    shape = np.array([W, H]).astype(float) - 1
    for k in range(2):
        im[:, k] *= 2.0 / shape[k]; pt[..., k] /= shape[k]
    pt *= 2; pt -= 1

    def key(p):
        k = p[0, 0].numpy().view(np.uint32).astype(np.uint64)
        return k[:, 0] | (k[:, 1] << np.uint64(32))
    hist = np.zeros((niter + 1, n), np.uint64)
    hist[0] = key(pt)
    for t in range(niter):
        dd = torch.nn.functional.grid_sample(im, pt, align_corners=False)
        for k in range(2):
            pt[..., k] = torch.clamp(pt[..., k] + dd[:, k], -1, 1)
        hist[t + 1] = key(pt)
    return ys, xs, hist


def euler_fraction(ys, xs, hist, schedule, pool):
    """Euler steps executed (whole warps) / steps of the plain kernel, for chunks of `pool` list entries in patch
    order (16 x 64 patches) merged at the steps in `schedule`."""
    order = np.lexsort((xs, ys, xs // 64, ys // 16))
    chunk = np.empty(len(ys), np.int64); chunk[order] = np.arange(len(ys)) // pool
    total = 0
    pts = [0] + list(schedule) + [hist.shape[0] - 1]
    for c in np.unique(chunk):
        alive = np.nonzero(chunk == c)[0]
        for a, b in zip(pts[:-1], pts[1:]):
            if a > 0:
                _, first = np.unique(hist[a, alive], return_index=True)
                alive = alive[first]
            total += -(-len(alive) // 32) * 32 * (b - a)
    return total / (len(ys) * (hist.shape[0] - 1))


if __name__ == "__main__":
    data = [trajectories(s) for s in (3, 4)]
    ys, xs, hist = data[0]
    print("distinct positions per tile (seed 3, %d foreground pixels):" % len(ys))
    for t in (10, 20, 30, 40, 50, 60, 80, 100, 120, 160, 200):
        print("  step %3d: %5.1f %%" % (t, 100.0 * len(np.unique(hist[t])) / len(ys)))
    for name, sch, pool in (("two points, 256-entry chunks (k_follow_merge)", [48, 96], 256),
                            ("11 points, 1024-entry pool", [24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160], 1024),
                            ("6 points, 1024-entry pool", [28, 40, 56, 72, 96, 128], 1024),
                            ("4 points, 1024-entry pool (default)", [36, 56, 88, 136], 1024),
                            ("3 points, 1024-entry pool", [32, 56, 96], 1024)):
        f = np.mean([euler_fraction(*d, sch, pool) for d in data])
        print("  %-48s Euler work %.3f of the plain kernel" % (name, f))
