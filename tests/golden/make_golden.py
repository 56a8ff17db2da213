"""Generate the golden vectors under tests/golden/ from the reference's own source.

Run in the build container only (needs /root/reference, which does not exist on the
GPU box):   python tests/golden/make_golden.py

The reference package cannot be imported here (its `cellpose` dependency is absent),
so the three hot-path functions it owns are lifted out of its source files by AST and
executed as-is -- the bytes that run are the reference's, not a copy kept in this repo:

  compute_class_masks      /root/reference/src/classpose/models.py:191-230
  remove_border_instances  /root/reference/src/classpose/metrics/pq.py:65-92
  unaugment_class_tiles    /root/reference/src/classpose/transforms/transforms.py:4-21

Outputs (committed): ref_class_vote.npz, ref_border.npz, ref_unaugment.npz
"""
import ast
import os
import sys

import numpy as np
import torch

REF = "/root/reference/src/classpose"
HERE = os.path.dirname(os.path.abspath(__file__))


def lift(path, name):
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            mod = ast.Module(body=[node], type_ignores=[])
            ns = {"np": np, "torch": torch}
            exec(compile(mod, path, "exec"), ns)
            return ns[name]
    raise KeyError(name)


def main():
    ccm = lift(f"{REF}/models.py", "compute_class_masks")
    rbi = lift(f"{REF}/metrics/pq.py", "remove_border_instances")
    uct = lift(f"{REF}/transforms/transforms.py", "unaugment_class_tiles")
    rng = np.random.default_rng(20261017)

    # ---- class vote: blobs of labels, several class counts, ties, bg-majority, gaps in ids
    vote = {}
    for case, (H, W, C, nlab) in enumerate([(48, 64, 7, 12), (40, 40, 10, 30), (33, 57, 5, 6), (24, 24, 2, 3)]):
        yy, xx = np.mgrid[0:H, 0:W]
        masks = np.zeros((H, W), np.int64)
        ids = rng.choice(np.arange(1, 3 * nlab), size=nlab, replace=False)  # non-contiguous ids are legal
        for k in ids:
            cy, cx, r = rng.uniform(0, H), rng.uniform(0, W), rng.uniform(2, 7)
            masks[(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = k
        logits = rng.normal(size=(C, 1, H, W)).astype(np.float32)
        # exact ties between two channels on one instance; background wins on another
        present = np.unique(masks)[1:]
        if len(present) >= 2 and C > 2:
            m = masks == present[0]
            logits[1, 0][m] = 9.0
            logits[2, 0][m] = 9.0
            m = masks == present[1]
            logits[0, 0][m] = 9.0
        cm, uniq = ccm(masks.copy(), logits.copy())
        vote[f"masks{case}"] = masks
        vote[f"logits{case}"] = logits
        vote[f"class_masks{case}"] = cm
        vote[f"unique{case}"] = uniq
    vote["ncases"] = np.array(4)
    np.savez_compressed(f"{HERE}/ref_class_vote.npz", **vote)

    # ---- border removal: random label images, 2-D and (H,W,2)
    border = {}
    for case, (H, W) in enumerate([(17, 23), (32, 32), (5, 5), (64, 40)]):
        yy, xx = np.mgrid[0:H, 0:W]
        inst = np.zeros((H, W), np.int64)
        for k in range(1, 9):
            cy, cx, r = rng.uniform(0, H), rng.uniform(0, W), rng.uniform(1, 5)
            inst[(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = k
        border[f"in2d_{case}"] = inst
        border[f"out2d_{case}"] = rbi(inst.copy())
        both = np.stack([inst, (inst % 3) + (inst > 0)], axis=-1)
        border[f"in3d_{case}"] = both
        border[f"out3d_{case}"] = rbi(both.copy())
    border["ncases"] = np.array(4)
    np.savez_compressed(f"{HERE}/ref_border.npz", **border)

    # ---- class un-augment on numpy arrays (what run_net passes)
    un = {}
    for case, (ny, nx, C, ly, lx) in enumerate([(3, 3, 4, 6, 8), (2, 3, 2, 5, 5)]):
        y = rng.normal(size=(ny, nx, C, ly, lx)).astype(np.float32)
        un[f"in{case}"] = y
        un[f"out{case}"] = uct(y.copy())
    un["ncases"] = np.array(2)
    np.savez_compressed(f"{HERE}/ref_unaugment.npz", **un)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    sys.exit(main())
