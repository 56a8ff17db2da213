#!/usr/bin/env python
"""Benchmark of the Classpose post-network path (BASELINE.json metric: tiles/s & cells/s on
synthetic 256x256 tiles; % of the HBM roofline).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--tiles B] [--impl reference]

Our arm: one "step" = one pass of the fused path (dP, cellprob, logits -> masks, per-cell class) over
a batch of B=1024 synthetic conic tiles (256x256, C=7) per GPU -- BASELINE.json configs[1].  Inputs
are resident in HBM for `value`; `e2e` times the host-buffer C-ABI call (pinned host inputs, H2D and
D2H inside the timed region).  With N>1 (torchrun, one rank per GPU) every rank processes its own B
tiles (weak scaling) and the one exchange -- the all-gather of per-rank instance totals for global
label offsets -- is inside the step.

Reference arm (`--impl reference`): the reference's own implementation of this path lives in
cellpose==4.0.8, which cannot be installed here (absent from the image and the wheelhouse, no
network), so the arm times the oracle port of it (oracle/, op-for-op restatement using torch-CPU
grid_sample / scipy) on all host cores, one process per core, on a bounded sample of the same
workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = W = 256
C = 7
WORKLOAD = "conic config: batch of 1024 synthetic 256x256 tiles (C=7), compute_masks + class vote (BASELINE configs[1])"
PARAMS = dict(niter=200, cellprob_threshold=0.0, flow_threshold=0.4, min_size=15, max_size_fraction=0.4)

# Additional BASELINE.json configs (not the driver's bench line; run with --workload):
#   wsi   configs[2]: synthetic 40x WSI 100k x 100k px, tile 256 / overlap 0.1 -> 494^2 = 244,036 tiles, puma (C=10),
#                     tiles sharded over the ranks, cyclic pool of 1024 distinct resident tiles per rank
#   tta   configs[3]: monusac (C=5) with --tta: 9 flipped sub-tile maps per padded 272^2 tile blended before dynamics
#   dense configs[4]: 512^2 tiles, ~2k cells per tile, flow check on
EXTRA = {
    "wsi": dict(H=256, W=256, C=10, tiles=1024, n_grid=10, axes=(5.0, 9.0), total_tiles=244036,
                name="synthetic 40x WSI 100k x 100k px, tile 256 overlap 0.1 (244,036 tiles), puma C=10 (BASELINE configs[2])"),
    "tta": dict(H=256, W=256, C=5, tiles=256, n_grid=10, axes=(5.0, 9.0),
                name="monusac C=5 with TTA: 9 sub-tile maps per padded 272^2 tile blended, then dynamics (BASELINE configs[3])"),
    "dense": dict(H=512, W=512, C=7, tiles=256, n_grid=45, axes=(3.5, 5.0),
                  name="dense nuclei stress: 512x512 tiles, ~2k cells per tile, flow check on (BASELINE configs[4])"),
    "touching": dict(H=256, W=256, C=7, tiles=512, n_grid=10, axes=(5.0, 9.0), style="touching",
                     name="hostile conic variant: Voronoi-clipped touching cells, 15 % of the cells with noise for flows, "
                          "every 8th tile a 43-px cell (block kernels), every 16th a ring (hole fill)"),
}


def cpu_sample_size(B):
    """Tiles the CPU arms time per step (about 10-30 s of host work on the box's cores)."""
    cores = os.cpu_count() or 1
    return min(B, 256, max(64, 4 * cores))


def shared_config(B):
    """The workload description BOTH arms print, so that their lines describe the same configuration."""
    S = cpu_sample_size(B)
    return {"workload": WORKLOAD, "tiles_per_gpu": B, "tile": [H, W], "classes": C, **PARAMS,
            "inputs": f"tiles 0..{S - 1} of every batch come from oracle.synth.make_tile(seed = tile index) on the host and are "
                      f"the tiles the CPU arms time; the remaining tiles come from the device generator of the same "
                      f"specification (SURVEY 8d)"}


_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


# ----------------------------------------------------------------------------------------------
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thread = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def wait_first(self, timeout=5.0):
        t0 = time.time()
        while self.proc is not None and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.05)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); smax.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------
def _real():
    """The real cellpose, if this box has it (scripts/probe_reference.py); None -> the oracle port is timed."""
    from oracle import real
    return real.find()


def _oracle_tile(args):
    """One tile through the reference path on the host (worker process): the real cellpose when importable,
    otherwise the oracle port of it; the class vote is the reference's own arithmetic either way."""
    import numpy as np
    import torch
    torch.set_num_threads(1)
    from oracle import classpose_ref, dynamics, real
    dP, cellprob, logits = args
    r = real.find()
    if r is not None:
        m = real.resize_and_compute_masks(r, dP, cellprob, **PARAMS)
    else:
        m = dynamics.resize_and_compute_masks(dP, cellprob, **PARAMS)
    cm, _ = classpose_ref.compute_class_masks(m, logits[:, None])
    return int(m.max())


def _oracle_make(seed):
    import torch
    torch.set_num_threads(1)
    from oracle import synth
    t = synth.make_tile(seed, H=H, W=W, C=C)
    return t["dP"], t["cellprob"], t["logits"]


def cpu_reference_throughput(tiles, cores, repeats=1):
    """tiles: list of (dP, cellprob, logits) numpy triples.  One process per core, like the reference's
    one-worker-per-device model.  Returns (tiles/s, cells/s)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(_oracle_tile, tiles[:cores])          # warm-up: imports, torch init
        best = None
        cells = 0
        for _ in range(repeats):
            t0 = time.perf_counter()
            res = pool.map(_oracle_tile, tiles, chunksize=1)
            dt = time.perf_counter() - t0
            cells = sum(res)
            best = dt if best is None else min(best, dt)
    return len(tiles) / best, cells / best, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    REAL = _real()
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    sample = cpu_sample_size(args.tiles)
    ctx = mp.get_context("spawn")
    with ctx.Pool(min(cores, sample)) as pool:
        tiles = pool.map(_oracle_make, range(sample))
    steps, warm = max(1, args.steps), max(0, args.warmup)
    import multiprocessing
    ctx = multiprocessing.get_context("spawn")
    times, cells = [], 0
    with ctx.Pool(cores) as pool:
        for _ in range(max(1, warm)):
            pool.map(_oracle_tile, tiles[:cores], chunksize=1)
        for _ in range(steps):
            t0 = time.perf_counter()
            res = pool.map(_oracle_tile, tiles, chunksize=1)
            times.append(time.perf_counter() - t0)
            cells = sum(res)
    total = sum(times)
    value = sample * steps / total
    line = {
        "impl": "reference", "metric": "tiles_per_sec", "value": value, "unit": "tiles/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 * total / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "cells_per_sec": cells * steps / total,
        "config": shared_config(args.tiles),
        "note": ("times the real cellpose " + REAL["version"] if REAL else
                 "reference arithmetic lives in cellpose==4.0.8 (absent, not installable offline: scripts/probe_reference.py); "
                 "this arm times the oracle port of it") + " on the host cores, one process per core",
        "cpu_baseline": {"value": value, "unit": "tiles/s", "cores": cores, "kind": "reference" if REAL else "port",
                         "sample": f"tiles 0..{sample - 1} of the workload per step, one process per core"},
        "e2e": {"value": value, "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from classpose_b200 import distributed as cdist
    from classpose_b200 import synth
    from classpose_b200.engine import get_engine

    eng = get_engine(dev)
    B = args.tiles
    N = H * W
    data = synth.make_batch(B, H, W, C, seed=1234 + 7919 * rank, device=dev)
    dP, cellprob, logits = data["dP"], data["cellprob"], data["logits"]
    del data
    # the first S tiles are the host-generated ones the CPU arms time (same generator, same seeds in both arms)
    S = cpu_sample_size(B)
    import multiprocessing as mp
    with mp.get_context("spawn").Pool(min(os.cpu_count() or 1, S)) as pool:
        host_tiles = pool.map(_oracle_make, range(S))
    for i, (a_, b_, c_) in enumerate(host_tiles):
        dP[i].copy_(torch.from_numpy(a_)); cellprob[i].copy_(torch.from_numpy(b_)); logits[i].copy_(torch.from_numpy(c_))
    torch.cuda.synchronize()

    def step():
        masks, counts, cell_class, _ = eng.compute_masks_batch(dP, cellprob, logits, **PARAMS)
        offs, total, base = cdist.global_label_offsets(counts, eng) if world > 1 else (None, None, 0)
        return masks, counts, cell_class

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    for _ in range(max(3, args.warmup)):
        out = step()
    barrier()
    n_cells_step = int(out[1].sum().item())
    if rank == 0:
        sampler.rows.clear()          # keep only samples taken during the timed region
    launches0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        out = step()
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        c = torch.tensor([n_cells_step], device=dev, dtype=torch.int64)
        dist.all_reduce(c)
        n_cells_all = int(c.item())
    else:
        n_cells_all = n_cells_step
    ms_step = ms_total / args.steps
    tiles_per_s = world * B / (ms_step * 1e-3)
    cells_per_s = n_cells_all / (ms_step * 1e-3)

    if args.profile_only:
        if rank == 0:
            emit({"metric": "tiles_per_sec", "value": tiles_per_s, "unit": "tiles/s", "n_gpus": world, "steps": args.steps,
                  "ms_per_step": ms_step, "note": "profile-only run (under a profiler: not a bench value)"})
        if world > 1:
            dist.destroy_process_group()
        return 0
    # ---- end to end through the host-buffer C-ABI call (pinned host inputs, copies inside the timed region)
    hdP = torch.empty(dP.shape, dtype=torch.float32, pin_memory=True); hdP.copy_(dP)
    hcp = torch.empty(cellprob.shape, dtype=torch.float32, pin_memory=True); hcp.copy_(cellprob)
    hlg = torch.empty(logits.shape, dtype=torch.float32, pin_memory=True); hlg.copy_(logits)
    LC = eng.label_capacity(H, W)
    # label images come back as uint16, the dtype Cellpose (and so the reference) returns below 65,536 labels
    outbuf = {"masks": torch.empty((B, H, W), dtype=torch.uint16, pin_memory=True),
              "counts": torch.empty((B,), dtype=torch.int32, pin_memory=True),
              "cell_class": torch.zeros((B, LC), dtype=torch.int32, pin_memory=True)}
    torch.cuda.synchronize()
    # pinned-copy peak of this box, measured live (the e2e path is PCIe-bound: this is its roofline)
    big = hlg.reshape(-1)[: min(hlg.numel(), 256 * 1024 * 1024)]
    dbig = torch.empty_like(big, device=dev)
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pcie_peak = 0.0
    for _ in range(3):
        pe0.record(); dbig.copy_(big, non_blocking=True); pe1.record(); torch.cuda.synchronize()
        pcie_peak = max(pcie_peak, big.numel() * 4 / (pe0.elapsed_time(pe1) * 1e-3) / 1e9)
    del dbig
    e2e_steps = max(2, min(args.steps, 5))

    def time_e2e(logits_mode, flows_mode, threads=1):
        """Host-buffer call over the whole batch; with threads=2 two host threads (the reference's own model:
        predict_wsi.py:728-797 runs two inference threads per process) each push half of the batch, so the ramp-up of
        one call overlaps the steady state of the other."""
        def call(lo, hi, ob):
            eng.compute_masks_host(hdP[lo:hi], hcp[lo:hi], hlg[lo:hi], out=ob, tiles_per_chunk=args.chunk,
                                   logits_mode=logits_mode, flows_mode=flows_mode, masks_u16=True, **PARAMS)
        if threads == 1:
            parts = [(0, B, outbuf)]
        else:
            cut = [B * i // threads for i in range(threads + 1)]
            parts = [(cut[i], cut[i + 1], {k: v[cut[i]:cut[i + 1]] for k, v in outbuf.items()}) for i in range(threads)]

        def once():
            if len(parts) == 1:
                call(*parts[0])
            else:
                th = [threading.Thread(target=call, args=p_) for p_ in parts]
                [t.start() for t in th]; [t.join() for t in th]
        for _ in range(2):
            once()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            once()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / e2e_steps
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt
    e2e_upload_s = time_e2e("upload", "upload")
    e2e_lg_s = time_e2e("auto", "upload")
    e2e_s = time_e2e("auto", "auto")    # the API default: pinned logits / flows are read in place where they are needed
    same = bool((outbuf["masks"][:8].to(torch.int32).to(dev) == out[0][:8]).all().item())
    same = same and bool((outbuf["cell_class"][:8, :64].to(dev) == out[2][:8, :64]).all().item())
    fg4 = float((out[0].reshape(-1, 4) > 0).any(dim=1).float().mean().item())
    fgrp = float((cellprob.reshape(-1, 4) > PARAMS["cellprob_threshold"]).any(dim=1).float().mean().item())
    copied_up = B * N * 4                                      # cellprob: every pixel
    # in place: 16 bytes per class for every 4-pixel group that holds a cell, 32 bytes of flow per group with foreground
    mapped_up = int(B * N * C * 4 * fg4) + int(B * N * 8 * fgrp)
    h2d = copied_up + mapped_up
    # what the link actually carries: kernel reads of host memory move 64-byte half lines (tests/studies/zerocopy_bw.cu),
    # i.e. every 16-pixel aligned run that holds a cell (logits, per class) / foreground (flows, per component)
    moved_up = None
    if W % 16 == 0:
        cell16 = float((out[0].reshape(-1, 16) > 0).any(dim=1).float().mean().item())
        fg16 = float((cellprob.reshape(-1, 16) > PARAMS["cellprob_threshold"]).any(dim=1).float().mean().item())
        moved_up = copied_up + int(B * N * C * 4 * cell16) + int(B * N * 8 * fg16)
    d2h = B * N * 2 + B * 4 + B * min(LC, 512) * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- per-stage device times (CUDA events around every stage, separate pass) and the roofline of the
    #      dominant kernel
    stage_runs = [eng.profile_stages(dP, cellprob, logits, with_qc=True, **PARAMS) for _ in range(3)]
    qc_stats = stage_runs[-1][1]
    stage_runs = [r[0] for r in stage_runs]
    stages = {k: statistics.median(r[k] for r in stage_runs) for k in stage_runs[0]}
    fg_frac = float((cellprob > PARAMS["cellprob_threshold"]).float().mean().item())
    fg4_frac = float((out[0].reshape(-1, 4) > 0).any(dim=1).float().mean().item()) if (H * W) % 4 == 0 else 1.0
    stage_bytes = {   # algorithmic bytes per tile (SURVEY.md 8d / DESIGN.md)
        "follow_flows": 12 * N + 4 * fg_frac * N,
        "diffuse": 4 * N + 8 * fg_frac * N,
        "vote": 4 * C * N + 4 * N,
        # final ids + class vote in one pass: labels in / out, logits of the 4-pixel groups that hold a cell pixel
        "final_map": 8 * N + 4 * C * N * fg4_frac,
        "prep_flow": 12 * N + 8 * N + 4 * N + 4 * fg_frac * N,   # dP, cellprob in; scaled flow, zeroed labels, fg list out
    }
    dom = max(stages, key=stages.get)
    peak, peak_src = load_peaks()
    # DRAM traffic of the dominant kernel: from the committed ncu capture of this same workload (per launch)
    traffic, traffic_src = None, None
    kern = {"follow_flows": "k_follow_pool", "diffuse": "k_diffuse32", "vote": "k_vote", "prep_flow": "k_prep_flow_v4",
            "final_map": "k_final_vote_v4"}.get(dom)
    try:
        import glob
        tj = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*", "traffic.json")))[-1]
        tdata = json.load(open(tj))
        if kern in tdata["kernels"] and tdata.get("tiles_per_launch") == B:
            k_ = tdata["kernels"][kern]
            traffic = k_["dram_read_bytes"] + k_["dram_write_bytes"]
            traffic_src = os.path.relpath(tj, ROOT)
    except Exception:
        pass
    dom_bytes = stage_bytes.get(dom, (16 + 4 * C) * N) * B
    achieved = dom_bytes / (stages[dom] * 1e-3) / 1e9
    whole_bytes = (16 + 4 * C) * N * B
    # the streaming stages against the same HBM peak (their design traffic / CUDA-event time), for orientation
    stage_hbm = {k: {"GBs": stage_bytes[k] * B / (stages[k] * 1e-3) / 1e9, "frac": stage_bytes[k] * B / (stages[k] * 1e-3) / 1e9 / peak}
                 for k in ("prep_flow", "final_map") if stages.get(k, 0) > 0}
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes": dom_bytes, "peak_source": peak_src,
                "kernel_ms": stages[dom], "kernel_share_of_step": stages[dom] / sum(stages.values()),
                "whole_path_achieved_GBs": whole_bytes / (ms_step * 1e-3) / 1e9,
                "whole_path_frac": whole_bytes / (ms_step * 1e-3) / 1e9 / peak,
                "streaming_stages": stage_hbm,
                "note": "follow_flows (200 dependent Euler steps per foreground pixel, ~300 flop per byte) is bound by its instruction "
                        "mix (73 % of the issue slots, FP32 pipe 63 % of cycles, DRAM at 3 % of peak); the flow check (float32 "
                        "register-resident screen) is shuffle / issue bound.  The HBM fraction is reported as required; the pipe "
                        "utilisations and the kernels that ARE HBM streams (prep 97 %, final+vote 78 %, blend 71-86 % of the "
                        "measured copy peak, by ncu DRAM bytes) are in DESIGN.md section 4"}

    # ---- CPU baseline: oracle port on the host cores, bounded sample of the same workload
    cores = os.cpu_count() or 1
    sample = S
    cpu = None
    if not args.no_cpu_baseline:
        REAL = _real()
        tps, cps, dt = cpu_reference_throughput(host_tiles, min(cores, sample))
        cpu = {"value": tps, "unit": "tiles/s", "cores": min(cores, sample), "kind": "reference" if REAL else "port",
               "sample": f"tiles 0..{sample - 1} of the batch (host-generated), " +
                         (f"real cellpose {REAL['version']}" if REAL else "oracle port") + f", one process per core, {dt:.1f} s",
               "cells_per_sec": cps}

    line = {
        "metric": "tiles_per_sec", "value": tiles_per_s, "unit": "tiles/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "cells_per_sec": cells_per_s, "cells_per_step": n_cells_all,
        "config": shared_config(B),
        "run": {"foreground_fraction": fg_frac, "parallelism": f"tiles sharded over {world} GPU(s), no data-path collective",
                "cache": "inputs 2.5 GiB per step >> 126 MB L2 (no flush needed)"},
        "e2e": {"value": world * B / e2e_s, "unit": "tiles/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s * 1e3, "matches_device_path": same,
                "inputs": "cellprob copied; dP and logits read in place from the pinned host buffers (dP where a 4-pixel group holds "
                          "foreground, logits where it holds a cell); one host thread, one call per step",
                "h2d_copied_bytes": copied_up, "h2d_mapped_bytes_min": mapped_up,
                "pcie_gbs": h2d / e2e_s / 1e9, "pcie_peak_gbs": pcie_peak, "pcie_frac": h2d / e2e_s / 1e9 / pcie_peak,
                "h2d_moved_bytes_64B_lines": moved_up,
                "pcie_link_gbs": (moved_up / e2e_s / 1e9) if moved_up else None,
                "pcie_link_frac": (moved_up / e2e_s / 1e9 / pcie_peak) if moved_up else None,
                "variants": {
                    "upload_everything": {"value": world * B / e2e_upload_s, "ms_per_step": e2e_upload_s * 1e3,
                                          "h2d_bytes_per_step": B * (3 + C) * N * 4,
                                          "pcie_gbs": B * (3 + C) * N * 4 / e2e_upload_s / 1e9},
                    "logits_in_place_flows_uploaded": {"value": world * B / e2e_lg_s, "ms_per_step": e2e_lg_s * 1e3}}},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "stages_ms": stages,
        "flow_check": qc_stats,
    }
    if world == 1 and not args.no_extras:
        try:
            line["hooks_e2e"] = hooks_e2e(host_tiles)
        except Exception as e:
            line["hooks_e2e"] = {"error": f"{type(e).__name__}: {e}"}
        del dP, cellprob, logits, hdP, hcp, hlg, outbuf, out
        torch.cuda.empty_cache()
        line["extra_configs"] = brief_extras(eng, dev, peak)
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def build_extra(eng, dev, workload, B, seed):
    """Inputs and the step function of one of the additional BASELINE configs (device-resident)."""
    import torch
    from classpose_b200 import synth, transforms as btf
    cfg = EXTRA[workload]
    Hh, Ww, Cc = cfg["H"], cfg["W"], cfg["C"]
    data = synth.make_batch(B, Hh, Ww, Cc, n_grid=cfg["n_grid"], axes=cfg["axes"], seed=seed, device=dev,
                            chunk=min(64, B), style=cfg.get("style", "isolated"))
    dP, cellprob, logits = data["dP"], data["cellprob"], data["logits"]
    extra = {}
    if workload == "tta":
        # what run_net holds after the network when augment=True: 9 flipped sub-tile maps of the padded tile
        pad = btf.get_pad_yx(Hh, Ww, min_size=(256, 256))
        Ly, Lx = Hh + pad[0] + pad[1], Ww + pad[2] + pad[3]
        geo = btf.tile_geometry(Ly, Lx, 256, augment=True)
        full = torch.zeros((B, 3 + Cc, Ly, Lx), device=dev)
        full[:, 0:2, pad[0]:pad[0] + Hh, pad[2]:pad[2] + Ww] = dP
        full[:, 2, pad[0]:pad[0] + Hh, pad[2]:pad[2] + Ww] = cellprob
        full[:, 3:, pad[0]:pad[0] + Hh, pad[2]:pad[2] + Ww] = logits
        nt = len(geo["y0"])
        sub = torch.empty((B, nt, 3 + Cc, 256, 256), device=dev)
        for j in range(nt):
            t = full[:, :, geo["y0"][j]:geo["y0"][j] + 256, geo["x0"][j]:geo["x0"][j] + 256].clone()
            f = int(geo["flip"][j])
            if f & 1:
                t = t.flip(2); t[:, 0] *= -1          # the network sees a Y-flipped tile: dY changes sign
            if f & 2:
                t = t.flip(3); t[:, 1] *= -1
            sub[:, j] = t
        y_flow = sub[:, :, :3].contiguous()
        y_cls = sub[:, :, 3:].contiguous()
        del sub, full
        ty, tx = btf.taper_1d(256, 256)
        g = {k: torch.from_numpy(geo[k]).to(dev) for k in ("y0", "x0", "flip")}
        tyd, txd = torch.from_numpy(ty).to(dev), torch.from_numpy(tx).to(dev)
        x4, cover = btf.tile_cover(geo["y0"], geo["x0"], 256, 256, Ly, Lx)

        from classpose_b200 import make_params

        def step():
            # one library call: blend of the logits, blend of the flow map fused with the cellprob threshold (foreground
            # list / scaled flow field / zeroed labels come out of the blend), mask path, class vote
            o = eng.calls.eval_tail(y_flow, y_cls, g["y0"], g["x0"], g["flip"], True, tyd, txd, Ly, Lx,
                                    tuple(int(p_) for p_ in pad), make_params(**PARAMS))
            return o[0], o[1], o[2]
        out = step()
        ref = eng.compute_masks_batch(dP, cellprob, logits, **PARAMS)
        extra["blend_max_abs_err_vs_unblended"] = float((eng.calls.average_tiles(
            y_flow, g["y0"], g["x0"], g["flip"], True, tyd, txd, Ly, Lx, pad)[:, :2] - dP).abs().max().item())
        extra["cells_vs_unblended"] = [int(out[1].sum().item()), int(ref[1].sum().item())]
        for _ in range(3):      # warm-up (stream-ordered allocator pool, clocks)
            eng.calls.average_tiles(y_flow, g["y0"], g["x0"], g["flip"], True, tyd, txd, Ly, Lx, pad, x4, cover)
            eng.calls.average_tiles(y_cls, g["y0"], g["x0"], g["flip"], False, tyd, txd, Ly, Lx, pad, x4, cover)
        torch.cuda.synchronize()
        reps = []
        for _ in range(7):      # median of single repetitions: the pair is ~1 ms of GPU time behind six python calls
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            eng.calls.average_tiles(y_flow, g["y0"], g["x0"], g["flip"], True, tyd, txd, Ly, Lx, pad, x4, cover)
            eng.calls.average_tiles(y_cls, g["y0"], g["x0"], g["flip"], False, tyd, txd, Ly, Lx, pad, x4, cover)
            ev[1].record(); torch.cuda.synchronize()
            reps.append(ev[0].elapsed_time(ev[1]))
        blend_ms = statistics.median(reps)
        blend_bytes = B * (3 + Cc) * 4 * (nt * 256 * 256 + Hh * Ww)
        extra["blend_ms"] = blend_ms
        extra["blend_GBs"] = blend_bytes / (blend_ms * 1e-3) / 1e9
    else:
        def step():
            return eng.compute_masks_batch(dP, cellprob, logits, **PARAMS)
    return cfg, step, (dP, cellprob, logits), extra


def hooks_e2e(host_tiles, threads=2, n=400):
    """Numpy in / numpy out through the callables install() puts in place of the reference's (hooks A and C), one tile per
    call as predict_wsi.py:749-756 does: single-thread latency and the throughput of `threads` host threads."""
    import numpy as np
    from classpose_b200 import models
    tiles = [(np.ascontiguousarray(a[:, None]), b[None], c[:, None]) for a, b, c in host_tiles[:16]]

    def one(i):
        dP, cp, lg = tiles[i % len(tiles)]
        m = models.compute_masks(dP, cp, (1, H, W), False, PARAMS["niter"], PARAMS["cellprob_threshold"],
                                 PARAMS["flow_threshold"], PARAMS["min_size"], PARAMS["max_size_fraction"], 0.0, None)
        cm, _ = models.compute_class_masks(m, lg)
        return m
    for i in range(40):
        one(i)
    lat = []
    for i in range(n):
        t0 = time.perf_counter(); one(i); lat.append(time.perf_counter() - t0)
    lat.sort()

    errors = []

    def worker(k):
        try:
            for i in range(40):           # plan creation (graph capture, pinned allocations) and warm-up
                one(i)
            bar.wait(timeout=120)
            for i in range(n):
                one(i + k)
        except Exception as e:            # a failing thread must not leave the others waiting
            errors.append(f"{type(e).__name__}: {e}")
            bar.abort()
    bar = threading.Barrier(threads + 1)
    th = [threading.Thread(target=worker, args=(k,), daemon=True) for k in range(threads)]
    [t.start() for t in th]
    try:
        bar.wait(timeout=120)
    except threading.BrokenBarrierError:
        pass
    t0 = time.perf_counter()
    [t.join(timeout=300) for t in th]
    dt = time.perf_counter() - t0
    if errors or any(t.is_alive() for t in th):
        return {"single_tile_ms_median": 1e3 * lat[len(lat) // 2], "single_tile_ms_p90": 1e3 * lat[int(0.9 * len(lat))],
                "threads": threads, "tiles_per_sec": None, "error": errors or ["a hook thread did not finish"]}
    return {"single_tile_ms_median": 1e3 * lat[len(lat) // 2], "single_tile_ms_p90": 1e3 * lat[int(0.9 * len(lat))],
            "threads": threads, "tiles_per_sec": threads * n / dt,
            "path": "models.compute_masks -> models.compute_class_masks (hooks A, C), host numpy in / out, one 256x256 "
                    "tile per call: CUDA-graph plan + labels kept on the device between the two hooks"}


def brief_extras(eng, dev, peak):
    """Short device-resident passes over the other BASELINE configs and the hostile workload (N = 1 only), so that the
    driver's line carries them: tiles/s, cells/s, per-stage device times, flow-check statistics."""
    import torch
    out = {}
    for name, B in (("wsi", 512), ("tta", 256), ("dense", 64), ("touching", 512)):
        try:
            cfg, step, (dP, cellprob, logits), extra = build_extra(eng, dev, name, B, seed=99)
            for _ in range(3):
                res = step()
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 5
            ev0.record()
            for _ in range(n):
                res = step()
            ev1.record(); torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1) / n
            cells = int(res[1].sum().item())
            stages, qc = eng.profile_stages(dP, cellprob, logits, with_qc=True, **PARAMS)
            N_ = cfg["H"] * cfg["W"]
            line = {"workload": cfg["name"], "tiles": B, "tile": [cfg["H"], cfg["W"]], "classes": cfg["C"],
                    "tiles_per_sec": B / (ms * 1e-3), "cells_per_sec": cells / (ms * 1e-3), "cells_per_tile": cells / B,
                    "ms_per_step": ms, "whole_path_frac_of_hbm_peak": (16 + 4 * cfg["C"]) * N_ * B / (ms * 1e-3) / 1e9 / peak,
                    "stages_ms": {k: round(v, 4) for k, v in stages.items() if v > 0}, "flow_check": qc, **extra}
            if name == "wsi":
                line["slide_seconds_1gpu_at_this_rate"] = cfg["total_tiles"] / line["tiles_per_sec"]
            if name == "tta":
                line["note"] = "stages_ms is the dynamics part on unblended inputs; ms_per_step includes the two 9-way blends"
            out[name] = line
            del dP, cellprob, logits, res, step
            torch.cuda.empty_cache()
        except Exception as e:          # an extra must never cost the main line
            out[name] = {"error": f"{type(e).__name__}: {e}"}
    return out


def run_extra(args):
    """BASELINE configs[2..4]; same timing rules as the main arm (CUDA events, max over ranks)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from classpose_b200 import distributed as cdist
    from classpose_b200.engine import get_engine
    eng = get_engine(dev)
    cfg = EXTRA[args.workload]
    Hh, Ww, Cc, B = cfg["H"], cfg["W"], cfg["C"], (args.tiles if args.tiles != 1024 or args.workload == "wsi" else cfg["tiles"])
    cfg, step, (dP, cellprob, logits), extra = build_extra(eng, dev, args.workload, B, seed=99 + 7919 * rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    nbatches = 1
    if args.workload == "wsi":
        a, b_ = cdist.shard_range(cfg["total_tiles"], rank, world)
        nbatches = -(-(b_ - a) // B)
        extra["tiles_this_rank"] = b_ - a

    for _ in range(max(3, args.warmup)):
        out = step()
    barrier()
    cells_step = int(out[1].sum().item())
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = args.steps if args.workload != "wsi" else nbatches
    barrier()
    ev0.record()
    total_cells = torch.zeros((), dtype=torch.int64, device=dev)
    for _ in range(steps):
        out = step()
        total_cells += out[1].sum()
    if world > 1:      # the one exchange: per-rank instance totals -> global label offsets
        base = cdist.rank_base_offset(total_cells.reshape(1))      # stays on the device
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        c = total_cells.clone()
        dist.all_reduce(c)
        total_cells = c
    if rank == 0:
        if args.workload == "wsi":
            tiles_done = cfg["total_tiles"]
            extra["slide_seconds"] = ms * 1e-3
            extra["note"] = "each rank loops over its shard in batches of %d tiles drawn cyclically from %d resident tiles" % (B, B)
        else:
            tiles_done = world * B * steps
        emit({"metric": "tiles_per_sec", "value": tiles_done / (ms * 1e-3), "unit": "tiles/s", "n_gpus": world,
              "steps": steps, "warmup": max(3, args.warmup), "ms_per_step": ms / steps, "higher_is_better": True,
              "scaling": "strong" if args.workload == "wsi" else "weak", "vs_baseline": None, "dtype": "f32",
              "data": "synthetic", "cells_per_sec": float(total_cells.item()) / (ms * 1e-3),
              "cells_per_tile": cells_step / B,
              "config": {"workload": cfg["name"], "tiles_per_batch": B, "tile": [Hh, Ww], "classes": Cc, **PARAMS}, **extra})
    if world > 1:
        dist.destroy_process_group()
    return 0


def _protect_stdout():
    """Libraries (NCCL prints its version) may write to fd 1; the contract is ONE JSON line on stdout.
    Point fd 1 at stderr for the duration of the run and return a file object on the real stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--tiles", type=int, default=1024, help="tiles per GPU per step")
    ap.add_argument("--chunk", type=int, default=128, help="tiles per chunk of the host-buffer call")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="conic1024", choices=["conic1024", "wsi", "tta", "dense", "touching"])
    ap.add_argument("--no-extras", action="store_true", help="skip the short passes over the other configs")
    ap.add_argument("--profile-only", action="store_true",
                    help="device-resident steps only (no e2e, no stage pass, no CPU baseline): what ncu launch lists wrap")
    args = ap.parse_args()
    global _REAL_STDOUT
    _REAL_STDOUT = _protect_stdout()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload != "conic1024":
        return run_extra(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
