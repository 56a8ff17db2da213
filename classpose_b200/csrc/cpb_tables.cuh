// Label-table maintenance shared by rows (3)-(7): table init, "map labels + gather
// statistics" pixel pass, the size filter / renumbering table op of row (5).
#pragma once
#include "cpb_common.cuh"
#include "cpb_masks.cuh"

// grid (ceil(LC/256), B): reset entries 0..lbound[b] of the statistics tables.
CPB_KERNEL k_init_tables(LabelTables t) {
    const int b = blockIdx.y;
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l > t.lbound[b] || l >= t.LC) return;
    const size_t k = (size_t)b * t.LC + l;
    t.cnt[k] = 0; t.first[k] = CPB_IMAX;
    t.ymin[k] = CPB_IMAX; t.ymax[k] = -1; t.xmin[k] = CPB_IMAX; t.xmax[k] = -1;
    t.sumy[k] = 0; t.sumx[k] = 0; t.flag[k] = 0; t.done[k] = 0;
}

CPB_KERNEL k_fill_i32(int* p, int n, int v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// k_map_stats: one thread per pixel of a [B,H,W,nch] label image (statistics on channel 0).
//   l = lab[p]
//   if (holekey && holekey[p]) l = low 32 bits of holekey[p]          (hole fill result)
//   else if (map)              l = map[b][l]                            (remap / drop)
//   else if (drop)             l = drop[b][l] ? 0 : l                   (flagged labels -> 0)
//   write back when changed (other channels are zeroed when a label is dropped to 0);
//   accumulate the tables for the new label when `stats`.
CPB_KERNEL CPB_LAUNCH_BOUNDS(256, 4)
k_map_stats(int* CPB_RESTRICT lab, int B, int H, int W, int nch, const int* CPB_RESTRICT map,
            const int* CPB_RESTRICT drop, const u64* CPB_RESTRICT holekey, int stats, LabelTables t) {
    const int N = H * W;
    const long long total = (long long)B * N;
    const long long base = (long long)blockIdx.x * blockDim.x;
    if (base >= total) return;
    const long long gi = base + threadIdx.x;
    int l = 0, b = 0, r = 0, y = 0, x = 0;
    if (gi < total) {
        b = (int)(gi / N);
        r = (int)(gi - (long long)b * N);
        const int l0 = lab[gi * nch];
        l = l0;
        u64 hk = holekey ? holekey[gi] : 0;
        if (hk) l = (int)(hk & 0xffffffffu);
        else if (map) l = map[(size_t)b * t.LC + l0];
        else if (drop) l = drop[(size_t)b * t.LC + l0] ? 0 : l0;
        if (l != l0) {
            lab[gi * nch] = l;
            if (l == 0) for (int c = 1; c < nch; c++) lab[gi * nch + c] = 0;
        }
        y = r / W; x = r - y * W;
    }
    if (stats) cpb_stats_accum(t, b, l, r, y, x);
}

// k_size_renumber: one block per tile; the table half of fill_holes_and_remove_small_masks.
//   mode 1 (min_size > 0):  counts = counts of the sorted unique values, first one dropped;
//            every position k with counts[k] < min_size removes the label VALUE k+1 (upstream
//            indexes by position -- identical to the label only for contiguous labels with
//            background present; kept as is), then renumber by first appearance.
//   mode 0:  compact labels in increasing value order (the `j` counter of the fill loop).
// Produces remap (old -> new), nlab, lbound(new) and -- permuted to the new ids -- the
// bbox tables the hole fill needs (nbbox = 4 arrays [B][LC], may be NULL).
CPB_KERNEL CPB_LAUNCH_BOUNDS(1024, 1)
k_size_renumber(LabelTables t, int H, int W, int min_size, int mode, u64* CPB_RESTRICT scratch_key,
                int* CPB_RESTRICT scratch_idx, int* CPB_RESTRICT counts_out) {
    CPB_SHARED int s_scan[33];
    CPB_SHARED int s_n, s_carry, s_fg;
    CPB_SHARED u64 s_keys[CPB_RANK_CHUNK];
    const int b = blockIdx.x;
    const int LC = t.LC;
    const int lb = t.lbound[b];
    const int* cnt = t.cnt + (size_t)b * LC;
    const int* first = t.first + (size_t)b * LC;
    int* remap = t.remap + (size_t)b * LC;
    int* flag = t.flag + (size_t)b * LC;       // removed-by-position marks (zeroed by k_init_tables)
    u64* keys = scratch_key + (size_t)b * LC;
    int* rank = scratch_idx + (size_t)b * LC;
    if (threadIdx.x == 0) { s_n = 0; s_carry = 0; s_fg = 0; }
    __syncthreads();
    // foreground pixel total -> is background present?
    int part = 0;
    for (int l = 1 + threadIdx.x; l <= lb; l += blockDim.x) part += cnt[l];
    for (int d = 16; d; d >>= 1) part += __shfl_xor_sync(CPB_FULL, part, d);
    if ((threadIdx.x & 31) == 0 && part) atomicAdd(&s_fg, part);
    __syncthreads();
    const int has_bg = (s_fg < H * W) ? 1 : 0;

    if (mode == 1) {
        // position (1-based rank among present labels, ascending value) via a chunked block scan
        for (int l0 = 1; l0 <= lb; l0 += blockDim.x) {
            const int l = l0 + threadIdx.x;
            const int present = (l <= lb && cnt[l] > 0) ? 1 : 0;
            int tot;
            const int incl = cpb_block_scan_incl(present, s_scan, &tot);
            const int pos = s_carry + incl;            // 1-based position among present labels
            if (present && cnt[l] < min_size) {
                const int victim = pos - (has_bg ? 0 : 1);
                if (victim >= 1 && victim < LC) flag[victim] = 1;
            }
            __syncthreads();
            if (threadIdx.x == 0) s_carry += tot;
            __syncthreads();
        }
    }
    __syncthreads();
    for (int l = threadIdx.x; l <= lb; l += blockDim.x) {
        remap[l] = 0;
        if (l >= 1 && cnt[l] > 0 && !(mode == 1 && flag[l])) {
            const int k = atomicAdd(&s_n, 1);
            const unsigned ord = mode == 1 ? (unsigned)first[l] : (unsigned)l;
            keys[k] = ((u64)ord << 32) | (unsigned)l;
        }
    }
    __syncthreads();
    const int n = s_n;
    cpb_block_rank(keys, n, rank, s_keys);
    for (int k = threadIdx.x; k < n; k += blockDim.x) remap[(int)(keys[k] & 0xffffffffu)] = rank[k];
    if (threadIdx.x == 0) {
        t.nlab[b] = n;
        t.lbound[b] = max(n, 0);
        if (counts_out) counts_out[b] = n;
    }
}
