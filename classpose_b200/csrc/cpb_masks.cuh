// Row (3): get_masks -- seeds of the end-point histogram, constrained 11x11 growth,
// label lookup, over-sized label removal, first-appearance renumbering.
// Semantics follow cellpose.dynamics.get_masks_torch.  The reference pads
// the histogram by rpad=20 so that windows never leave the array; end points never fall in
// the padding, so here the histogram lives on the H x W tile and out-of-tile reads are 0.
#pragma once
#include "cpb_common.cuh"

#define CPB_GROW_MIN 2
#define CPB_RANK_CHUNK 2048

// rank[k] = #{ j : key[j] < key[k] } for n distinct 64-bit keys.  s_keys: CPB_RANK_CHUNK u64 of shared memory.
// out[k] = rank + 1 for k < n.
//   n <= CPB_RANK_CHUNK: bitonic sort of a copy in shared memory, then every key finds its position by binary search
//                        (O(n log^2 n / blockDim): a dense 512 x 512 tile ranks ~1,800 labels several times per call);
//   larger n           : block-cooperative O(n^2 / blockDim) counting, chunk by chunk.
CPB_DEVICE void cpb_block_rank(const u64* CPB_RESTRICT keys, int n, int* CPB_RESTRICT out, u64* s_keys) {
    if (n <= CPB_RANK_CHUNK) {
        int m = 32;
        while (m < n) m <<= 1;
        __syncthreads();
        for (int i = threadIdx.x; i < m; i += blockDim.x) s_keys[i] = i < n ? keys[i] : ~0ull;
        __syncthreads();
        for (int k = 2; k <= m; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = threadIdx.x; i < m; i += blockDim.x) {
                    const int p = i ^ j;
                    if (p > i) {
                        const u64 a = s_keys[i], b = s_keys[p];
                        const bool up = (i & k) == 0;
                        if ((a > b) == up) { s_keys[i] = b; s_keys[p] = a; }
                    }
                }
                __syncthreads();
            }
        }
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            const u64 mine = keys[k];
            int lo = 0, hi = n - 1;                   // keys are distinct: exactly one match
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (s_keys[mid] < mine) lo = mid + 1; else hi = mid;
            }
            out[k] = lo + 1;
        }
        __syncthreads();
        return;
    }
    for (int kb = 0; kb < n; kb += blockDim.x) {
        const int k = kb + threadIdx.x;
        const u64 mine = k < n ? keys[k] : 0;
        int rank = 0;
        for (int c0 = 0; c0 < n; c0 += CPB_RANK_CHUNK) {
            const int cn = min(CPB_RANK_CHUNK, n - c0);
            __syncthreads();
            for (int j = threadIdx.x; j < cn; j += blockDim.x) s_keys[j] = keys[c0 + j];
            __syncthreads();
            if (k < n)
                for (int j = 0; j < cn; j++) rank += s_keys[j] < mine ? 1 : 0;
        }
        if (k < n) out[k] = rank + 1;
    }
    __syncthreads();
}

// k_seed_scan + k_seeds (one block per tile).
//  1. seeds = pixels with h > 10 (k_seed_scan collects them) that equal the maximum of their 5x5 neighbourhood (first
//     step of k_seeds)
//  2. order: ascending count, ties by raster position (stable sort of a raster-ordered list)
//  3. each seed grows inside its 11x11 window: 5 x { 3x3 dilation ; &= h > 2 }
//  4. paint label = order+1; later (larger) labels overwrite.  The labels are painted INTO the histogram as
//     negative numbers (atomicMin of -label): only pixels with h > 2 are ever painted and the only reads that
//     follow are "h > 2" tests, which a painted pixel passes by construction -- so no separate label plane has to
//     be zeroed, written and read.  Afterwards hist[p] < 0 means label -hist[p], anything else label 0.
// k_seed_scan: the whole batch as one stream (one thread per 4 pixels, 128-bit loads when vec != 0): every pixel
// with h > 10 is appended to its tile's candidate list (cand_count must be zeroed; at most N / 11 < LC pixels of a tile
// can hold more than 10 end points).  The histogram is almost everywhere 0, so nearly every thread stops after its
// load; the 5 x 5 maximum test of the candidates -- 25 scattered loads for the few lanes of a warp that hold one --
// is left to k_seeds, where the candidates are compact (in this kernel it cost more than the stream itself).
CPB_KERNEL CPB_LAUNCH_BOUNDS(256, 4)
k_seed_scan(const int* CPB_RESTRICT hist, int B, int H, int W, int LC, int vec, u64* CPB_RESTRICT seed_key,
            int* CPB_RESTRICT cand_count) {
    const int N = H * W;
    const long long nq = vec ? (long long)B * N / 4 : (long long)B * N;
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    int vals[4] = {0, 0, 0, 0};
    if (vec) {
        const int4 v4 = *reinterpret_cast<const int4*>(hist + 4 * q);
        if (max(max(v4.x, v4.y), max(v4.z, v4.w)) <= CPB_SEED_MIN) return;
        vals[0] = v4.x; vals[1] = v4.y; vals[2] = v4.z; vals[3] = v4.w;
    } else {
        vals[0] = hist[q];
    }
    const long long g0 = vec ? 4 * q : q;
    const int b = (int)(g0 / N);
    const int p0 = (int)(g0 - (long long)b * N);
    for (int e = 0; e < (vec ? 4 : 1); e++) {
        const int v = vals[e];
        if (v <= CPB_SEED_MIN) continue;
        const int k = atomicAdd(&cand_count[b], 1);
        if (k < LC) seed_key[(size_t)b * LC + k] = (u64)(unsigned)(p0 + e);      // (k_seeds reads the count itself)
    }
}

CPB_KERNEL CPB_LAUNCH_BOUNDS(1024, 1)
k_seeds(int* hist, int H, int W, int LC, u64* CPB_RESTRICT seed_key,
        int* CPB_RESTRICT seed_lab, int* CPB_RESTRICT nseeds, const int* CPB_RESTRICT cand_count) {
    CPB_SHARED u64 s_keys[CPB_RANK_CHUNK];
    const int b = blockIdx.x;
    const int N = H * W;
    u64* keys = seed_key + (size_t)b * LC;
    int* labs = seed_lab + (size_t)b * LC;
    int* Mb = hist + (size_t)b * N;
    // candidates (h > 10) -> seeds: those that equal the maximum of their 5 x 5 neighbourhood, compacted in place a
    // block-width at a time (a round only writes below the positions it has already read; nothing is painted yet)
    CPB_SHARED int s_cscan[33];
    const int nc = min(cand_count[b], LC - 1);
    int n = 0;
    for (int k0 = 0; k0 < nc; k0 += blockDim.x) {
        const int k = k0 + threadIdx.x;
        u64 key = 0;
        bool ismax = false;
        if (k < nc) {
            const int p = (int)(keys[k] & 0xffffffffu);       // candidate: a pixel with h > 10 (from the counting
            const int v = Mb[p];                              // kernel or from k_seed_scan)
            key = ((u64)(unsigned)v << 32) | (unsigned)p;
            const int y = p / W, x = p - y * W;
            ismax = true;
            for (int dy = -2; dy <= 2 && ismax; dy++)
                for (int dx = -2; dx <= 2; dx++) {
                    const int yy = y + dy, xx = x + dx;     // plain loads: this kernel writes the plane further down
                    if (yy >= 0 && yy < H && xx >= 0 && xx < W && Mb[yy * W + xx] > v) { ismax = false; break; }
                }
        }
        int tot;
        const int incl = cpb_block_scan_incl(ismax ? 1 : 0, s_cscan, &tot);      // barriers: every key of the round is read
        if (ismax) keys[n + incl - 1] = key;
        n += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) nseeds[b] = n;
    cpb_block_rank(keys, n, labs, s_keys);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int s = warp; s < n; s += nw) {
        const u64 key = keys[s];
        const int p = (int)(key & 0xffffffffu);
        const int label = labs[s];
        const int sy = p / W, sx = p - sy * W;
        const int y = sy - 5 + lane;
        unsigned allowed = 0;
        if (lane < 11 && y >= 0 && y < H) {
            int v[11];                                 // the row's 11 loads in flight together
            #pragma unroll
            for (int c = 0; c < 11; c++) {
                const int x = sx - 5 + c;
                v[c] = (x >= 0 && x < W) ? Mb[y * W + x] : 0;
            }
            #pragma unroll
            for (int c = 0; c < 11; c++)
                if (v[c] > CPB_GROW_MIN || v[c] < 0) allowed |= 1u << c;
        }
        unsigned m = (lane == 5) ? (1u << 5) : 0u;
        for (int it = 0; it < 5; it++) {
            const unsigned d = m | (m << 1) | (m >> 1);
            unsigned up = __shfl_up_sync(CPB_FULL, d, 1);
            unsigned dn = __shfl_down_sync(CPB_FULL, d, 1);
            if (lane == 0) up = 0;
            if (lane == 31) dn = 0;
            m = (d | up | dn) & allowed;
        }
        while (m) {
            const int c = __ffs((int)m) - 1;
            m &= m - 1;
            atomicMin(&Mb[y * W + sx - 5 + c], -label);
        }
    }
}

// k_lookup: label of every foreground pixel = painted label at its end point; also fills
// the label tables (count / first appearance / bbox / sums) of the raw labels.
CPB_KERNEL CPB_LAUNCH_BOUNDS(256, 4)
k_lookup(const int* CPB_RESTRICT pfinal, const int* CPB_RESTRICT M, int B, int H, int W,
         int* CPB_RESTRICT lab, LabelTables t) {
    const int N = H * W;
    const long long total = (long long)B * N;
    const long long base = (long long)blockIdx.x * blockDim.x;
    if (base >= total) return;
    const long long gi = base + threadIdx.x;
    int l = 0, b = 0, r = 0, y = 0, x = 0;
    if (gi < total) {
        b = (int)(gi / N);
        r = (int)(gi - (long long)b * N);
        const int pf = pfinal[gi];
        if (pf >= 0) l = max(-M[(size_t)b * N + (pf >> 16) * W + (pf & 0xffff)], 0);    // painted histogram, see k_seeds
        lab[gi] = l;
        y = r / W; x = r - y * W;
    }
    cpb_stats_accum(t, b, l, r, y, x);
}

// k_gm_finalize: one block per tile.  Drop labels larger than max_size_fraction of the tile,
// renumber the rest 1..n in order of first appearance.
CPB_KERNEL CPB_LAUNCH_BOUNDS(1024, 1)
k_gm_finalize(LabelTables t, int H, int W, double max_size_fraction, u64* CPB_RESTRICT scratch_key,
              int* CPB_RESTRICT scratch_idx, int* CPB_RESTRICT counts_out, int keep_raw) {
    CPB_SHARED int s_n;
    CPB_SHARED u64 s_keys[CPB_RANK_CHUNK];
    const int b = blockIdx.x;
    const int LC = t.LC;
    const int lb = t.lbound[b];
    const double big = (double)((long long)H * W) * max_size_fraction;
    const int* cnt = t.cnt + (size_t)b * LC;
    const int* first = t.first + (size_t)b * LC;
    int* remap = t.remap + (size_t)b * LC;
    u64* keys = scratch_key + (size_t)b * LC;   // (first << 32) | label of kept labels
    int* rank = scratch_idx + (size_t)b * LC;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (int l = threadIdx.x; l <= lb; l += blockDim.x) {
        remap[l] = 0;
        if (keep_raw) t.alive[(size_t)b * LC + l] = 0;
        if (l >= 1) {
            const int c = cnt[l];
            if (c > 0 && !((double)c > big)) {
                if (keep_raw) t.alive[(size_t)b * LC + l] = 1;
                const int k = atomicAdd(&s_n, 1);
                keys[k] = ((u64)(unsigned)first[l] << 32) | (unsigned)l;
            }
        }
    }
    __syncthreads();
    const int n = s_n;
    cpb_block_rank(keys, n, rank, s_keys);
    for (int k = threadIdx.x; k < n; k += blockDim.x) remap[(int)(keys[k] & 0xffffffffu)] = rank[k];
    if (threadIdx.x == 0) {
        t.nlab[b] = n;
        if (!keep_raw) t.lbound[b] = n;   // labels are 1..n after the remap is applied
        if (counts_out) counts_out[b] = n;
    }
}
