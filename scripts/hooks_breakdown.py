"""Where a single-tile hook call spends its time (run on the GPU box)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from classpose_b200 import fastpath, models
from classpose_b200.engine import get_engine
from oracle import synth

eng = get_engine()
t = synth.make_tile(1)
dP4, cp3, lg = np.ascontiguousarray(t["dP"][:, None]), t["cellprob"][None], t["logits"][:, None]
plan = fastpath.tile_plan(eng, 256, 256, 200, 0.0, 0.4, 15, 0.4, True)


def bench(f, n=300):
    for _ in range(10):
        f()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    return 1e6 * (time.perf_counter() - t0) / n


d3, c2 = np.ascontiguousarray(dP4[:, 0]), cp3[0]
m32, n = plan.run(d3, c2)
print("copy dP+cellprob into pinned     %7.1f us" % bench(lambda: (plan.h_dP.__setitem__(Ellipsis, d3), plan.h_cp.__setitem__(Ellipsis, c2))))
import ctypes as C
cnt = C.c_int32(0)
print("graph launch + sync (masks)      %7.1f us" % bench(lambda: plan.lib.cpb_tile_plan_run(plan.handle, C.byref(cnt))))
print("masks int32 -> uint16 copy       %7.1f us" % bench(lambda: m32.astype(np.uint16)))
print("plan.run total                   %7.1f us" % bench(lambda: plan.run(d3, c2)))
m = models.compute_masks(dP4, cp3, (1, 256, 256), False, 200, 0.0, 0.4, 15, 0.4, 0.0, None)
print("hook A total                     %7.1f us" % bench(lambda: models.compute_masks(dP4, cp3, (1, 256, 256), False, 200, 0.0, 0.4, 15, 0.4, 0.0, None)))
m = models.compute_masks(dP4, cp3, (1, 256, 256), False, 200, 0.0, 0.4, 15, 0.4, 0.0, None)
l3 = lg.reshape(7, 256, 256)
plan.vote(l3)
print("copy logits into pinned          %7.1f us" % bench(lambda: plan.h_lg.__setitem__(Ellipsis, l3)))
cm_, cc_ = C.c_void_p(), C.c_void_p()
print("graph launch + sync (vote)       %7.1f us" % bench(lambda: plan.lib.cpb_tile_plan_vote(plan.handle, C.byref(cm_), C.byref(cc_))))
cm8 = plan.vote(l3)
print("class image uint8 -> int64       %7.1f us" % bench(lambda: cm8.astype(np.int64)))
print("hook C total (cached labels)     %7.1f us" % bench(lambda: models.compute_class_masks(m, lg)))
print("hook C total (general path)      %7.1f us" % bench(lambda: models.compute_class_masks(m.copy(), lg), n=50))
import torch
st = eng.profile_stages(torch.from_numpy(t["dP"][None]).cuda(), torch.from_numpy(t["cellprob"][None]).cuda(), None)
print("single-tile stage times (us):", {k: round(1e3 * v, 1) for k, v in st.items() if v > 0}, "sum", round(1e3 * sum(st.values()), 1))
