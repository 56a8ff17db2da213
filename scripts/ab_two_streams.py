"""A/B: the fused path over the whole batch on one stream vs. the batch cut into k parts on k streams (run on the GPU box).
Kernels of different parts can then share the SMs: the latency-bound stages (lookup, label scan, seeds) of one part fill
issue slots the issue-bound Euler kernel of another leaves, and no kernel's tail wave leaves the GPU idle."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from classpose_b200 import synth
from classpose_b200.engine import get_engine

eng = get_engine()
dev = eng.device
B = 1024
d = synth.make_batch(B, 256, 256, 7, seed=1234, device=dev)
dP, cp, lg = d["dP"], d["cellprob"], d["logits"]
P = dict(niter=200, cellprob_threshold=0.0, flow_threshold=0.4, min_size=15, max_size_fraction=0.4)


def run(parts):
    if parts == 1:
        return [eng.compute_masks_batch(dP, cp, lg, **P)]
    main = torch.cuda.current_stream()
    ready = torch.cuda.Event(); ready.record(main)
    outs, evs = [], []
    n = B // parts
    for k in range(parts):
        s = streams[k]
        s.wait_event(ready)
        with torch.cuda.stream(s):
            outs.append(eng.compute_masks_batch(dP[k * n:(k + 1) * n], cp[k * n:(k + 1) * n], lg[k * n:(k + 1) * n], **P))
            e = torch.cuda.Event(); e.record(s); evs.append(e)
    for e in evs:
        main.wait_event(e)
    return outs


streams = [torch.cuda.Stream() for _ in range(8)]
ref = run(1)[0]
for parts in (1, 2, 4, 8, 1):
    for _ in range(3):
        o = run(parts)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        o = run(parts)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    same = torch.equal(torch.cat([x[0] for x in o]), ref[0])
    print(f"parts={parts}: {ms:.3f} ms per 1024 tiles = {B / ms * 1e3:,.0f} tiles/s, identical masks: {same}")
