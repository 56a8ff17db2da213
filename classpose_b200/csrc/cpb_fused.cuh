// Fused path only: the label image keeps the RAW labels (seed order + 1) from the lookup until the
// final pass; every renumbering / removal between get_masks and the output is carried by per-tile
// tables (remap = current id of a raw label, alive), so pixels are touched by three passes only:
// lookup (statistics), recount (tiles that had a hole filled) and final.
#pragma once
#include "cpb_common.cuh"
#include "cpb_masks.cuh"

// k_lookup_list: grid-stride over the compacted foreground list (full warps).  raw label = painted
// label at the pixel's end point; statistics of the raw labels go to the tables.  `lab` must be
// zeroed beforehand (background stays 0).
#ifndef CPB_LK_ILP
#define CPB_LK_ILP 4
#endif
CPB_KERNEL CPB_LAUNCH_BOUNDS(256, 4)
k_lookup_list(const unsigned* CPB_RESTRICT list, const unsigned* CPB_RESTRICT list_n,
              const int* CPB_RESTRICT pfinal, const int* CPB_RESTRICT M, int H, int W,
              int* CPB_RESTRICT lab, LabelTables t) {
    const unsigned total = *list_n;
    const int N = H * W;
    // CPB_LK_ILP list entries per thread and round, each level of the dependent chain (list -> end point -> painted
    // label) loaded for all of them before the next: the kernel is a chain of memory latencies otherwise
    for (unsigned i0 = blockIdx.x * blockDim.x * CPB_LK_ILP; i0 < total; i0 += gridDim.x * blockDim.x * CPB_LK_ILP) {
        unsigned gi[CPB_LK_ILP];
        int pf[CPB_LK_ILP], l[CPB_LK_ILP], b[CPB_LK_ILP];
        #pragma unroll
        for (int j = 0; j < CPB_LK_ILP; j++) {
            const unsigned i = i0 + j * blockDim.x + threadIdx.x;
            gi[j] = i < total ? list[i] : 0xffffffffu;
        }
        #pragma unroll
        for (int j = 0; j < CPB_LK_ILP; j++) {
            b[j] = gi[j] != 0xffffffffu ? (int)(gi[j] / (unsigned)N) : 0;
            pf[j] = gi[j] != 0xffffffffu ? pfinal[gi[j]] : 0;
        }
        #pragma unroll
        for (int j = 0; j < CPB_LK_ILP; j++)
            l[j] = gi[j] != 0xffffffffu ? max(-M[(size_t)b[j] * N + (pf[j] >> 16) * W + (pf[j] & 0xffff)], 0) : 0;   // painted histogram, see k_seeds
        #pragma unroll
        for (int j = 0; j < CPB_LK_ILP; j++) {
            if (i0 + j * blockDim.x >= total) break;                      // uniform: the whole block is past the list
            int r = 0, y = 0, x = 0;
            if (gi[j] != 0xffffffffu) {
                r = (int)(gi[j] - (unsigned)b[j] * (unsigned)N);
                lab[gi[j]] = l[j];
                y = r / W; x = r - y * W;
            }
            cpb_stats_accum(t, b[j], l[j], r, y, x);      // (a block-level shared-memory merge before the L2 atomics was slower:
                                                          //  0.355 ms vs 0.261 ms, three block barriers per 256 pixels)
        }
    }
}

// k_fuse_size: one block per tile.  Table form of "drop flagged labels, size filter, renumber":
//   present(l) = alive[l] && !flag[l] && cnt[l] > 0        (flag = bad flow, set by k_flow_err)
//   mode 2: no hole fill / size filter requested: remap = present ? current id : 0, nothing renumbered
//   mode 3 / 4 (hole fill with min_size <= 0, where upstream never renumbers by first appearance): 3 = before the
//           fill, ids compacted in increasing id order (the fill loop's `j` counter); 4 = after it, labels that were
//           overwritten inside a hole vanish and leave their gap, nlab = highest surviving id
//   otherwise: position of a present label = rank of its CURRENT id (remap) among present labels; when
//   min_size > 0 every present label with cnt < min_size removes the label whose current id equals that
//   position (upstream indexes by position, see k_size_renumber); survivors get new ids 1..n in order of
//   first appearance.  Writes remap, alive, nlab.  Tiles with only_if[b] == 0 are skipped when only_if != NULL.
CPB_KERNEL CPB_LAUNCH_BOUNDS(1024, 1)
k_fuse_size(LabelTables t, int H, int W, int min_size, int mode, const int* CPB_RESTRICT only_if,
            u64* CPB_RESTRICT scratch_key, int* CPB_RESTRICT scratch_idx, int* CPB_RESTRICT scratch_inv) {
    CPB_SHARED int s_n, s_fg;
    CPB_SHARED u64 s_keys[CPB_RANK_CHUNK];
    const int b = blockIdx.x;
    if (only_if && only_if[b] == 0) return;
    const int LC = t.LC;
    const int lb = t.lbound[b];
    const int* cnt = t.cnt + (size_t)b * LC;
    const int* first = t.first + (size_t)b * LC;
    int* remap = t.remap + (size_t)b * LC;
    int* flag = t.flag + (size_t)b * LC;
    int* alive = t.alive + (size_t)b * LC;
    u64* keys = scratch_key + (size_t)b * LC;
    int* rank = scratch_idx + (size_t)b * LC;
    int* inv = scratch_inv + (size_t)b * LC;     // current id -> raw label
    const int n_cur = t.nlab[b];
    if (mode == 2 || mode == 4) {
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        int top = 0;
        for (int l = threadIdx.x; l <= lb; l += blockDim.x) {
            const bool pres = l >= 1 && alive[l] && !flag[l] && cnt[l] > 0;
            if (!pres) { remap[l] = 0; alive[l] = 0; }
            else top = max(top, remap[l]);
        }
        if (mode == 4) {
            for (int d = 16; d; d >>= 1) top = max(top, __shfl_xor_sync(CPB_FULL, top, d));
            if ((threadIdx.x & 31) == 0) atomicMax(&s_n, top);
            __syncthreads();
            if (threadIdx.x == 0) t.nlab[b] = s_n;
        }
        return;
    }
    if (threadIdx.x == 0) { s_n = 0; s_fg = 0; }
    __syncthreads();
    int part = 0;
    for (int l = 1 + threadIdx.x; l <= lb; l += blockDim.x) {
        const bool pres = alive[l] && !flag[l] && cnt[l] > 0;
        if (pres) {
            part += cnt[l];
            const int k = atomicAdd(&s_n, 1);
            keys[k] = ((u64)(unsigned)remap[l] << 32) | (unsigned)l;
        }
        if (alive[l]) inv[remap[l]] = l;
    }
    for (int d = 16; d; d >>= 1) part += __shfl_xor_sync(CPB_FULL, part, d);
    if ((threadIdx.x & 31) == 0 && part) atomicAdd(&s_fg, part);
    __syncthreads();
    const int n = s_n;
    const int has_bg = (s_fg < H * W) ? 1 : 0;
    if (min_size > 0) {
        cpb_block_rank(keys, n, rank, s_keys);          // rank[k] = 1-based position among present labels
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            const int l = (int)(keys[k] & 0xffffffffu);
            if (cnt[l] < min_size) {
                const int victim = rank[k] - (has_bg ? 0 : 1);
                if (victim >= 1 && victim <= n_cur) flag[inv[victim]] = 1;
            }
        }
        __syncthreads();
    }
    // survivors, ordered by first appearance
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (int l = threadIdx.x; l <= lb; l += blockDim.x) {
        const bool pres = l >= 1 && alive[l] && !flag[l] && cnt[l] > 0;
        if (pres) {
            const int k = atomicAdd(&s_n, 1);
            keys[k] = ((u64)(unsigned)(mode == 3 ? remap[l] : first[l]) << 32) | (unsigned)l;
        } else {
            remap[l] = 0; alive[l] = 0;
        }
    }
    __syncthreads();
    const int n2 = s_n;
    cpb_block_rank(keys, n2, rank, s_keys);
    for (int k = threadIdx.x; k < n2; k += blockDim.x) remap[(int)(keys[k] & 0xffffffffu)] = rank[k];
    if (threadIdx.x == 0) t.nlab[b] = n2;
}

// k_recount_reset + k_recount: only tiles in which a hole was filled (t.misc[b] != 0) do any work: reset count /
// first appearance of the tile's labels, then recompute them on the image with the hole proposals applied (grid
// (slices, B): a tile is shared by several blocks, so the reset is its own launch).
CPB_KERNEL k_recount_reset(LabelTables t) {
    const int b = blockIdx.x;
    if (t.misc[b] == 0) return;
    const int LC = t.LC, lb = t.lbound[b];
    for (int l = threadIdx.x; l <= lb; l += blockDim.x) { t.cnt[(size_t)b * LC + l] = 0; t.first[(size_t)b * LC + l] = CPB_IMAX; }
}

CPB_KERNEL CPB_LAUNCH_BOUNDS(256, 4)
k_recount(const int* CPB_RESTRICT lab, const u64* CPB_RESTRICT holekey, int H, int W, LabelTables t) {
    const int b = blockIdx.y;
    if (t.misc[b] == 0) return;
    const int N = H * W, LC = t.LC;
    const int* L = lab + (size_t)b * N;
    const u64* HK = holekey + (size_t)b * N;
    const int per = ((N + gridDim.x - 1) / gridDim.x + 31) & ~31;          // pixels per slice, whole warps
    const int r_end = min(N, (int)(blockIdx.x + 1) * per);
    for (int r0 = blockIdx.x * per; r0 < r_end; r0 += blockDim.x) {
        const int r = r0 + threadIdx.x;
        int l = 0, y = 0, x = 0;
        if (r < r_end) {
            const u64 hk = HK[r];
            l = hk ? (int)(hk & 0xffffffffu) : L[r];
            if (l > 0 && t.alive[(size_t)b * LC + l] == 0) l = 0;
            y = r / W; x = r - y * W;
        }
        cpb_stats_accum(t, b, l, r, y, x);
    }
}

// k_final: raw label (or hole proposal) -> final id, in place; also publishes the per-tile counts.
CPB_KERNEL CPB_LAUNCH_BOUNDS(256, 4)
k_final(int* CPB_RESTRICT lab, const u64* CPB_RESTRICT holekey, int B, int H, int W, LabelTables t,
        int* CPB_RESTRICT counts_out) {
    const int N = H * W;
    const long long total = (long long)B * N;
    const long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= total) return;
    const int b = (int)(gi / N);
    const int was = lab[gi];
    int l = was;
    if (holekey && t.misc[b] != 0) {
        const u64 hk = holekey[gi];
        if (hk) l = (int)(hk & 0xffffffffu);
    }
    const int out = l > 0 ? t.remap[(size_t)b * t.LC + l] : 0;
    if (out != was) lab[gi] = out;      // (compared with what the pixel HOLDS: a filled hole whose label keeps its number
                                        //  under the remap used to be skipped -- found by tests/studies/fuzz_sim.py)
    if (gi - (long long)b * N == 0) {
        const int n = t.nlab[b];
        if (counts_out) counts_out[b] = n;
    }
}

// k_final_v4: N % 4 == 0; four pixels per thread with 128-bit accesses (tiles with filled holes take the scalar
// route per pixel, they are rare).
CPB_KERNEL CPB_LAUNCH_BOUNDS(256, 4)
k_final_v4(int4* CPB_RESTRICT lab, const u64* CPB_RESTRICT holekey, int B, int H, int W, LabelTables t,
           int* CPB_RESTRICT counts_out) {
    const int N4 = (H * W) >> 2;
    const long long total = (long long)B * N4;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    const int b = (int)(g / N4);
    const int* remap = t.remap + (size_t)b * t.LC;
    int4 v = lab[g];
    int l[4] = {v.x, v.y, v.z, v.w};
    if (holekey && t.misc[b] != 0) {
        #pragma unroll
        for (int e = 0; e < 4; e++) {
            const u64 hk = holekey[g * 4 + e];
            if (hk) l[e] = (int)(hk & 0xffffffffu);
        }
    }
    int4 o;
    o.x = l[0] > 0 ? remap[l[0]] : 0; o.y = l[1] > 0 ? remap[l[1]] : 0;
    o.z = l[2] > 0 ? remap[l[2]] : 0; o.w = l[3] > 0 ? remap[l[3]] : 0;
    if (o.x != v.x || o.y != v.y || o.z != v.z || o.w != v.w) lab[g] = o;
    if (g - (long long)b * N4 == 0 && counts_out) counts_out[b] = t.nlab[b];
}

// ---- class vote folded into the final pass ------------------------------------------------------------
// The final pass already visits every pixel with its final id in registers, so the per-pixel arg-max over the
// C logits (first maximum wins) and the (instance, class) histogram ride along: the label image is not read a
// second time and the logits are streamed by the whole grid with 128-bit loads instead of one block per tile.
// Histogram in global memory (L2 atomics, aggregated per thread and per warp), rows 0..nlab of each tile zeroed
// by k_vote_zero; k_vote_finish takes the per-instance arg-max (first maximum wins).
CPB_KERNEL k_vote_zero(LabelTables t, int C, int* CPB_RESTRICT vote) {
    const int b = blockIdx.x;
    int* tab = vote + (size_t)b * t.LC * C;
    const int need = (t.nlab[b] + 1) * C;
    for (int i = threadIdx.x; i < need; i += blockDim.x) tab[i] = 0;
}

template <int CT>      // CT > 0: number of classes known at compile time (all logit loads of a thread in flight at once)
CPB_KERNEL CPB_LAUNCH_BOUNDS(256, 4)
k_final_vote_v4(int4* CPB_RESTRICT lab, const u64* CPB_RESTRICT holekey, const float4* CPB_RESTRICT logits, int B, int H,
                int W, int C_rt, LabelTables t, int* CPB_RESTRICT counts_out, int* CPB_RESTRICT vote) {
    const int C = CT > 0 ? CT : C_rt;
    const int N4 = (H * W) >> 2;
    const long long total = (long long)B * N4;
    const long long g0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = g0 < total;
    const long long g = valid ? g0 : total - 1;
    const int b = (int)(g / N4);
    const int q = (int)(g - (long long)b * N4);
    const int* remap = t.remap + (size_t)b * t.LC;
    const int4 v = lab[g];
    int l[4] = {v.x, v.y, v.z, v.w};
    if (holekey && t.misc[b] != 0) {
        #pragma unroll
        for (int e = 0; e < 4; e++) {
            const u64 hk = holekey[g * 4 + e];
            if (hk) l[e] = (int)(hk & 0xffffffffu);
        }
    }
    int o[4];
    #pragma unroll
    for (int e = 0; e < 4; e++) o[e] = l[e] > 0 ? remap[l[e]] : 0;
    if (valid && (o[0] != v.x || o[1] != v.y || o[2] != v.z || o[3] != v.w)) lab[g] = make_int4(o[0], o[1], o[2], o[3]);
    if (valid && q == 0 && counts_out) counts_out[b] = t.nlab[b];
    int key[4] = {-1, -1, -1, -1}, cnt[4] = {1, 1, 1, 1};
    if (valid && (o[0] | o[1] | o[2] | o[3]) != 0) {
        const float4* G = logits + (size_t)b * C * N4 + q;
        float4 best = G[0];
        int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        if (CT > 0) {
            float4 x[CT > 1 ? CT - 1 : 1];
            #pragma unroll
            for (int c = 1; c < CT; c++) x[c - 1] = G[(size_t)c * N4];
            #pragma unroll
            for (int c = 1; c < CT; c++) {
                if (x[c - 1].x > best.x) { best.x = x[c - 1].x; a0 = c; }
                if (x[c - 1].y > best.y) { best.y = x[c - 1].y; a1 = c; }
                if (x[c - 1].z > best.z) { best.z = x[c - 1].z; a2 = c; }
                if (x[c - 1].w > best.w) { best.w = x[c - 1].w; a3 = c; }
            }
        } else {
            #pragma unroll 4
            for (int c = 1; c < C; c++) {
                const float4 x = G[(size_t)c * N4];
                if (x.x > best.x) { best.x = x.x; a0 = c; }
                if (x.y > best.y) { best.y = x.y; a1 = c; }
                if (x.z > best.z) { best.z = x.z; a2 = c; }
                if (x.w > best.w) { best.w = x.w; a3 = c; }
            }
        }
        const int arg[4] = {a0, a1, a2, a3};
        #pragma unroll
        for (int e = 0; e < 4; e++) if (o[e] > 0) key[e] = o[e] * C + arg[e];
        // the four pixels of a thread mostly agree: fold equal keys into the first of them
        #pragma unroll
        for (int e = 1; e < 4; e++) {
            #pragma unroll
            for (int f = 0; f < e; f++)
                if (key[e] >= 0 && key[e] == key[f]) { cnt[f] += cnt[e]; key[e] = -1; }
        }
    }
    int* tab = vote + (size_t)b * t.LC * C;
    // first key of every lane: one atomic per distinct (tile, key) of the warp
    if (__any_sync(CPB_FULL, key[0] >= 0)) {
        const long long wkey = key[0] >= 0 ? (long long)b * ((long long)t.LC * C) + key[0] : -1;
        const unsigned peers = __match_any_sync(CPB_FULL, wkey);
        const int sum = __reduce_add_sync(peers, cnt[0]);
        if (key[0] >= 0 && (threadIdx.x & 31) == __ffs((int)peers) - 1) atomicAdd(&tab[key[0]], sum);
    }
    #pragma unroll
    for (int e = 1; e < 4; e++) if (key[e] >= 0) atomicAdd(&tab[key[e]], cnt[e]);
}

CPB_KERNEL k_vote_finish(LabelTables t, int C, const int* CPB_RESTRICT vote, int* CPB_RESTRICT cell_class) {
    const int b = blockIdx.x;
    const int* tab = vote + (size_t)b * t.LC * C;
    int* cc = cell_class + (size_t)b * t.LC;
    const int nl = min(t.nlab[b], t.LC - 1);
    for (int l = threadIdx.x; l <= nl; l += blockDim.x) {
        int arg = 0;
        if (l > 0) {
            int best = tab[l * C];
            for (int c = 1; c < C; c++) {
                const int x = tab[l * C + c];
                if (x > best) { best = x; arg = c; }
            }
        }
        cc[l] = arg;
    }
}

// ---- stand-alone streaming vote (stage call on finished labels): the same per-pixel arg-max / histogram as the
// fused pass, spread over the whole grid -- one block per tile (k_vote) leaves a single-tile call on one SM.
CPB_KERNEL k_vote_zero_lb(const int* CPB_RESTRICT lbound, int LC, int C, int* CPB_RESTRICT vote) {
    const int b = blockIdx.x;
    int* tab = vote + (size_t)b * LC * C;
    const int need = (min(lbound[b], LC - 1) + 1) * C;
    for (int i = threadIdx.x; i < need; i += blockDim.x) tab[i] = 0;
}

template <int CT>
CPB_KERNEL CPB_LAUNCH_BOUNDS(256, 4)
k_vote_px_v4(const int4* CPB_RESTRICT lab, const float4* CPB_RESTRICT logits, int B, int H, int W, int C_rt, int LC,
             const int* CPB_RESTRICT lbound, int* CPB_RESTRICT vote) {
    const int C = CT > 0 ? CT : C_rt;
    const int N4 = (H * W) >> 2;
    const long long total = (long long)B * N4;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    const int b = (int)(g / N4);
    const int q = (int)(g - (long long)b * N4);
    const int lb = min(lbound[b], LC - 1);
    const int4 v = lab[g];
    int o[4] = {v.x, v.y, v.z, v.w};
    #pragma unroll
    for (int e = 0; e < 4; e++) if (o[e] < 0 || o[e] > lb) o[e] = 0;
    if ((o[0] | o[1] | o[2] | o[3]) == 0) return;
    const float4* G = logits + (size_t)b * C * N4 + q;
    float4 best = G[0];
    int arg[4] = {0, 0, 0, 0};
    #pragma unroll 4
    for (int c = 1; c < C; c++) {
        const float4 x = G[(size_t)c * N4];
        if (x.x > best.x) { best.x = x.x; arg[0] = c; }
        if (x.y > best.y) { best.y = x.y; arg[1] = c; }
        if (x.z > best.z) { best.z = x.z; arg[2] = c; }
        if (x.w > best.w) { best.w = x.w; arg[3] = c; }
    }
    int* tab = vote + (size_t)b * LC * C;
    #pragma unroll
    for (int e = 0; e < 4; e++) if (o[e] > 0) atomicAdd(&tab[o[e] * C + arg[e]], 1);
}

CPB_KERNEL k_vote_finish_lb(const int* CPB_RESTRICT lbound, int LC, int C, const int* CPB_RESTRICT vote,
                            int* CPB_RESTRICT cell_class) {
    const int b = blockIdx.x;
    const int* tab = vote + (size_t)b * LC * C;
    int* cc = cell_class + (size_t)b * LC;
    const int nl = min(lbound[b], LC - 1);
    for (int l = threadIdx.x; l <= nl; l += blockDim.x) {
        int arg = 0;
        if (l > 0) {
            int best = tab[l * C];
            for (int c = 1; c < C; c++) {
                const int x = tab[l * C + c];
                if (x > best) { best = x; arg = c; }
            }
        }
        cc[l] = arg;
    }
}

CPB_KERNEL k_class_image_v4(const int4* CPB_RESTRICT lab, int B, int H, int W, int LC, const int* CPB_RESTRICT lbound,
                            const int* CPB_RESTRICT cell_class, unsigned* CPB_RESTRICT class_masks) {
    const int N4 = (H * W) >> 2;
    const long long total = (long long)B * N4;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    const int b = (int)(g / N4);
    const int lb = min(lbound[b], LC - 1);
    const int* cc = cell_class + (size_t)b * LC;
    const int4 v = lab[g];
    const int l[4] = {v.x, v.y, v.z, v.w};
    unsigned out = 0;
    #pragma unroll
    for (int e = 0; e < 4; e++) if (l[e] > 0 && l[e] <= lb) out |= (unsigned)(cc[l[e]] & 0xff) << (8 * e);
    class_masks[g] = out;
}

// after k_final the image holds ids 1..nlab: shrink the label bound accordingly (vote, border); a tile some
// stage could not process (t.fail, e.g. the hole-fill bitmap pool ran out) reports counts[b] = -1
CPB_KERNEL k_finish_bounds(LabelTables t, int B, int* CPB_RESTRICT counts_out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    t.lbound[b] = t.nlab[b];
    if (counts_out && t.fail[b]) counts_out[b] = -1;
}

CPB_KERNEL k_apply_fail(const int* CPB_RESTRICT fail, int B, int* CPB_RESTRICT counts_out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B && fail[b]) counts_out[b] = -1;
}
