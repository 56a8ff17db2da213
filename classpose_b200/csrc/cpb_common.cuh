// Shared device helpers and the per-tile label-table layout.
#pragma once
#include "cpb_platform.h"

#define CPB_FULL 0xffffffffu
#define CPB_IMAX 0x7fffffff

typedef unsigned long long u64;
#define CPB_SEED_MIN 10      // get_masks: a seed collects more than this many end points

// Per-tile, per-label tables (struct of arrays, each [B][LC]); entry 0 is background.
struct LabelTables {
    int LC;          // entries per tile
    int* cnt;        // pixel count
    int* first;      // smallest raster index (first appearance)
    int* ymin; int* ymax; int* xmin; int* xmax;   // bounding box (inclusive)
    u64* sumy; u64* sumx;                          // coordinate sums
    int* remap;      // label -> new label (0 = dropped)
    int* flag;       // scratch flags (bad flow / removed-by-position / border)
    int* alive;      // fused path: label still exists (NULL = every label with cnt > 0)
    int* cy; int* cx;                              // diffusion centre
    double* err;     // flow error
    int* done;       // flow error already computed by the diffusion warp (label touches no other live label)
    int* lbound;     // [B] highest label value that may be present in the tile
    int* nlab;       // [B] number of instances after the last renumbering
    int* niter;      // [B] diffusion iterations (2 * max ext)
    int* misc;       // [B] stage scratch (seed count, hole flag, ...)
    int* fail;       // [B] != 0: a stage could not process this tile (reported as counts[b] = -1)
};

CPB_DEVICE int cpb_lane() { return threadIdx.x & 31; }

// Work items of the one-block-per-label kernels (centres, block diffusion, flow error, block hole fill).
//   list == NULL: tile = blockIdx.y, labels 1 + blockIdx.x, 1 + blockIdx.x + gridDim.x, ... (stage calls: every label)
//   list != NULL: explicit (tile, label) pairs, grid-stride over entries [0, count[0]) at the front of the list and,
//                 when `back`, [cap - count[1], cap) at its end.  The fused path lists only the labels that need a
//                 block (bbox beyond the warp kernels' 30 x 32, or in contact with another label): nuclei-scale
//                 batches have almost none, and a (24, B) grid whose blocks only discover that costs 50-190 us per
//                 kernel in block launches alone.
struct LabelWork { const int2* list; const int* count; int cap; int back; };

CPB_DEVICE bool cpb_next_label(const LabelWork& wk, const int* CPB_RESTRICT lbound, int& it, int& b, int& l) {
    if (wk.list) {
        const int nf = wk.count[0], nb = wk.back ? wk.count[1] : 0;
        const int i = blockIdx.x + it * gridDim.x;
        if (i >= nf + nb) return false;
        const int2 e = i < nf ? wk.list[i] : wk.list[wk.cap - 1 - (i - nf)];
        b = e.x; l = e.y;
    } else {
        b = blockIdx.y; l = 1 + blockIdx.x + it * gridDim.x;
        if (l > lbound[b]) return false;
    }
    it++;
    return true;
}

CPB_DEVICE bool cpb_label_live(const LabelTables& t, size_t k) {
    return t.cnt[k] > 0 && (t.alive == nullptr || t.alive[k] != 0);
}

// Block-wide inclusive scan of one int per thread (blockDim.x multiple of 32, <= 1024).
// `warp_tot` must point to >= 33 ints of shared memory.  Returns inclusive prefix;
// *block_total receives the sum.  Contains __syncthreads().
CPB_DEVICE int cpb_block_scan_incl(int v, int* warp_tot, int* block_total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int x = v;
    for (int d = 1; d < 32; d <<= 1) {
        int y = __shfl_up_sync(CPB_FULL, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int t = lane < nw ? warp_tot[lane] : 0;
        for (int d = 1; d < 32; d <<= 1) {
            int y = __shfl_up_sync(CPB_FULL, t, d);
            if (lane >= d) t += y;
        }
        if (lane < nw) warp_tot[lane] = t;
        if (lane == 31) warp_tot[32] = t;
    }
    __syncthreads();
    int base = warp > 0 ? warp_tot[warp - 1] : 0;
    *block_total = warp_tot[32];
    return base + x;
}

// Warp-aggregated update of the label tables for one pixel per lane.
// All 32 lanes must call; `lab` <= 0 lanes contribute nothing.  key = tile*LC + lab.
CPB_DEVICE void cpb_stats_accum(const LabelTables& t, int b, int lab, int ridx, int y, int x) {
    const int lane = threadIdx.x & 31;
    const bool act = lab > 0;
    const unsigned amask = __ballot_sync(CPB_FULL, act);
    if (!act) return;
    const int key = b * t.LC + lab;
    const unsigned peers = __match_any_sync(amask, key);
    const int leader = __ffs((int)peers) - 1;
    const int n = __popc(peers);
    const int fmin = __reduce_min_sync(peers, ridx);
    const int y0 = __reduce_min_sync(peers, y), y1 = __reduce_max_sync(peers, y);
    const int x0 = __reduce_min_sync(peers, x), x1 = __reduce_max_sync(peers, x);
    const int sy = __reduce_add_sync(peers, y), sx = __reduce_add_sync(peers, x);
    if (lane == leader) {
        atomicAdd(&t.cnt[key], n);
        atomicAdd(&t.sumy[key], (u64)sy); atomicAdd(&t.sumx[key], (u64)sx);
        // (reading the monotone min / max entries first to skip no-op atomics was slower, 0.40 ms vs 0.26 ms in the
        //  lookup: the fire-and-forget reductions do not stall the warp, the reads do)
        atomicMin(&t.first[key], fmin);
        atomicMin(&t.ymin[key], y0); atomicMax(&t.ymax[key], y1);
        atomicMin(&t.xmin[key], x0); atomicMax(&t.xmax[key], x1);
    }
}

