#!/bin/bash
cd "$(dirname "$0")/.."
build() { nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared "$@" -I include -I classpose_b200/csrc -o classpose_b200/libclasspose_b200.so classpose_b200/csrc/cpb_api.cu; }
run() { python - "$1" <<'PY'
import sys, torch, numpy as np
sys.path.insert(0, ".")
from classpose_b200 import transforms as btf
from classpose_b200.engine import get_engine
eng = get_engine()
B = 256
pad = btf.get_pad_yx(256, 256, min_size=(256, 256)); Ly = Lx = 272
geo = btf.tile_geometry(Ly, Lx, 256, augment=True)
g = {k: torch.from_numpy(geo[k]).cuda() for k in ("y0", "x0", "flip")}
ty, tx = btf.taper_1d(256, 256); tyd, txd = torch.from_numpy(ty).cuda(), torch.from_numpy(tx).cuda()
x4, cover = btf.tile_cover(geo["y0"], geo["x0"], 256, 256, Ly, Lx)
for nch in (3, 5):
    y = torch.randn((B, 9, nch, 256, 256), device="cuda")
    for _ in range(3): o = eng.calls.average_tiles(y, g["y0"], g["x0"], g["flip"], nch == 3, tyd, txd, Ly, Lx, pad, x4, cover)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): o = eng.calls.average_tiles(y, g["y0"], g["x0"], g["flip"], nch == 3, tyd, txd, Ly, Lx, pad, x4, cover)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gb = (y.numel() * 4 * (256 * 256) / (256 * 256) + o.numel() * 4) / 1e9
    print(sys.argv[1], "nch", nch, "ms %.3f" % ms, "GB/s %.0f" % (gb / (ms * 1e-3)))
    del y, o
PY
}
for mb in 2 3 4; do build -DCPB_BLEND_MINB=$mb; run "minb=$mb"; done
build
