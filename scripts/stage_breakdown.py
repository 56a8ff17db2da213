"""Per-stage CUDA-event times of the fused path for the extra BASELINE workloads."""
import json, sys
import torch
sys.path.insert(0, ".")
from classpose_b200 import synth
from classpose_b200.engine import get_engine
eng = get_engine()
P = dict(niter=200, cellprob_threshold=0.0, flow_threshold=0.4, min_size=15, max_size_fraction=0.4)
for name, kw, B in (("dense512", dict(H=512, W=512, C=7, n_grid=45, axes=(3.5, 5.0)), 128),
                    ("puma256", dict(H=256, W=256, C=10, n_grid=10, axes=(5.0, 9.0)), 1024)):
    d = synth.make_batch(B, seed=5, chunk=64, **kw)
    for _ in range(2): eng.compute_masks_batch(d["dP"], d["cellprob"], d["logits"], **P)
    st = eng.profile_stages(d["dP"], d["cellprob"], d["logits"], **P)
    st = eng.profile_stages(d["dP"], d["cellprob"], d["logits"], **P)
    print(name, "total %.3f ms" % sum(st.values()), json.dumps({k: round(v, 3) for k, v in st.items() if v > 0}))
    del d
