// Rows (5)-(7), (1) and (e): hole filling, class vote, border-instance removal,
// taper blending of sub-tiles, global label offsets.
#pragma once
#include "cpb_common.cuh"

#define CPB_FILL_THREADS 128
#define CPB_FILL_WORDS 8192          // 32-bit words per bitmap in shared memory (2 bitmaps = 64 KB)
// Passes of the hole-fill kernels.  BOTH (stage calls): flag the tile and write the proposals into a zeroed
// `holekey` plane.  The fused path avoids zeroing 8 bytes per pixel of every tile when few tiles have holes:
// DETECT only flags tiles (t.misc), k_zero_hole_tiles clears the plane of the flagged tiles, WRITE redoes the
// flagged tiles and writes the proposals.
#define CPB_FILL_BOTH 0
#define CPB_FILL_DETECT 1
#define CPB_FILL_WRITE 2

CPB_KERNEL k_zero_hole_tiles(u64* CPB_RESTRICT holekey, int H, int W, LabelTables t) {
    const int b = blockIdx.y;
    if (t.misc[b] == 0) return;
    const int N = H * W;
    u64* HK = holekey + (size_t)b * N;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) HK[i] = 0;
}

// k_fill_holes: one block per label (labels strided over gridDim.x, tile = blockIdx.y).
// Holes of label l = pixels of its bbox crop that are not l and are not 4-connected, through
// non-l pixels, to the border of the crop (fill_voids.fill on `crop == l`).
// Every hole pixel proposes (bbox area, l) into holekey with atomicMax: when holes nest, the
// outermost instance -- the one the reference's sequential loop ends with -- wins.
// The two bitmaps of the crop live in shared memory up to CPB_FILL_WORDS words each; a larger crop (a label
// spanning more than ~512 x 512 pixels) takes them from `pool`, a bump allocator over a free global plane
// (pool.cursor must be zeroed per call).  If even that is exhausted the tile is marked failed (t.fail[b]),
// which the final pass reports as counts[b] = -1: never a silently unfilled mask.
struct FillPool { unsigned* words; u64* cursor; u64 cap; };

CPB_KERNEL CPB_LAUNCH_BOUNDS(CPB_FILL_THREADS, 3)
k_fill_holes(const int* CPB_RESTRICT lab, int H, int W, LabelTables t, u64* CPB_RESTRICT holekey,
             FillPool pool, int skip_small, LabelWork wk, int pass) {
    CPB_DYN_SMEM(unsigned, s_bits);     // free[CPB_FILL_WORDS] | reach[CPB_FILL_WORDS]
    CPB_SHARED int s_changed;
    CPB_SHARED long long s_off;
    const int LC = t.LC, N = H * W;
    int it_ = 0, b, l;
    while (cpb_next_label(wk, t.lbound, it_, b, l)) {
        const int* L = lab + (size_t)b * N;
        u64* HK = holekey + (size_t)b * N;
        const size_t k = (size_t)b * LC + l;
        if (pass == CPB_FILL_WRITE && t.misc[b] == 0) continue;   // no hole anywhere in this tile (block-uniform)
        if (!cpb_label_live(t, k)) continue;
        const int y0 = t.ymin[k], x0 = t.xmin[k];
        const int h = t.ymax[k] - y0 + 1, w = t.xmax[k] - x0 + 1;
        if (h < 3 || w < 3) continue;
        if (skip_small && h <= 32 && w <= 32) continue;     // handled by k_fill_holes_warp
        const int wpr = (w + 31) >> 5;
        const int words = h * wpr;
        volatile unsigned* fr = s_bits;
        volatile unsigned* rc = s_bits + CPB_FILL_WORDS;
        __syncthreads();
        if (words > CPB_FILL_WORDS) {                        // block-uniform
            if (threadIdx.x == 0) {
                const u64 off = atomicAdd(pool.cursor, (u64)(2 * words));
                s_off = (pool.words && off + 2ull * words <= pool.cap) ? (long long)off : -1;
                if (s_off < 0) t.fail[b] = 1;
            }
            __syncthreads();
            if (s_off < 0) continue;
            fr = pool.words + s_off;
            rc = fr + words;
        }
        // bitmaps: fr = pixel is not l ; rc = fr on the crop border
        for (int i = threadIdx.x; i < words; i += blockDim.x) {
            const int r = i / wpr, j = i - r * wpr;
            unsigned f = 0, e = 0;
            const int cmax = min(32, w - j * 32);
            for (int c = 0; c < cmax; c++) {
                const int x = j * 32 + c;
                if (L[(y0 + r) * W + x0 + x] != l) {
                    f |= 1u << c;
                    if (r == 0 || r == h - 1 || x == 0 || x == w - 1) e |= 1u << c;
                }
            }
            fr[i] = f; rc[i] = e;
        }
        __syncthreads();
        // flood from the border through free pixels (4-connectivity), monotone in-place sweeps
        for (;;) {
            if (threadIdx.x == 0) s_changed = 0;
            __syncthreads();
            bool ch = false;
            for (int i = threadIdx.x; i < words; i += blockDim.x) {
                const int r = i / wpr, j = i - r * wpr;
                const unsigned f = fr[i];
                unsigned cur = rc[i];
                unsigned n = cur;
                if (r > 0) n |= rc[i - wpr];
                if (r < h - 1) n |= rc[i + wpr];
                if (j > 0) n |= rc[i - 1] >> 31;
                if (j < wpr - 1) n |= rc[i + 1] << 31;
                n &= f;
                for (;;) {   // smear along the row inside the word
                    const unsigned m = (n | (n << 1) | (n >> 1)) & f;
                    if (m == n) break;
                    n = m;
                }
                if (n != cur) { rc[i] = n; ch = true; }
            }
            if (ch) s_changed = 1;
            __syncthreads();
            const int any = s_changed;
            __syncthreads();
            if (!any) break;
        }
        const u64 prio = (u64)((unsigned)(h * w)) << 32;
        for (int i = threadIdx.x; i < words; i += blockDim.x) {
            unsigned hole = fr[i] & ~rc[i];
            const int r = i / wpr, j = i - r * wpr;
            if (hole && pass != CPB_FILL_WRITE) t.misc[b] = 1;     // tile has at least one filled hole
            while (hole && pass != CPB_FILL_DETECT) {
                const int c = __ffs((int)hole) - 1;
                hole &= hole - 1;
                atomicMax(&HK[(y0 + r) * W + x0 + j * 32 + c], prio | (unsigned)l);
            }
        }
    }
}

// k_fill_holes_warp: the same hole detection for bboxes up to 32 x 32, one WARP per label and no shared
// memory: lane r holds row r of the crop as a 32-bit mask; the border flood runs on warp shuffles.
CPB_KERNEL CPB_LAUNCH_BOUNDS(128, 8)
k_fill_holes_warp(const int* CPB_RESTRICT lab, int H, int W, LabelTables t, u64* CPB_RESTRICT holekey, int pass) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int b = blockIdx.y, LC = t.LC, N = H * W;
    if (pass == CPB_FILL_WRITE && t.misc[b] == 0) return;      // no hole anywhere in this tile
    const int lb = t.lbound[b];
    const int* L = lab + (size_t)b * N;
    u64* HK = holekey + (size_t)b * N;
    for (int l = 1 + blockIdx.x * nw + warp; l <= lb; l += gridDim.x * nw) {
        const size_t k = (size_t)b * LC + l;
        if (!cpb_label_live(t, k)) continue;
        const int y0 = t.ymin[k], x0 = t.xmin[k];
        const int h = t.ymax[k] - y0 + 1, w = t.xmax[k] - x0 + 1;
        if (h < 3 || w < 3 || h > 32 || w > 32) continue;
        unsigned fr = 0;                       // row `lane`: bit c = pixel (lane, c) is not l
        #pragma unroll 4
        for (int r = 0; r < h; r++) {
            const bool notl = lane < w && L[(y0 + r) * W + x0 + lane] != l;
            const unsigned m = __ballot_sync(CPB_FULL, notl);
            if (lane == r) fr = m;
        }
        unsigned rc = (lane == 0 || lane == h - 1) ? fr : (fr & (1u | (1u << (w - 1))));
        if (lane >= h) rc = 0;
        for (;;) {
            unsigned up = __shfl_up_sync(CPB_FULL, rc, 1), dn = __shfl_down_sync(CPB_FULL, rc, 1);
            if (lane == 0) up = 0;
            if (lane == 31) dn = 0;
            unsigned n = (rc | up | dn) & fr;
            for (;;) {
                const unsigned m = (n | (n << 1) | (n >> 1)) & fr;
                if (m == n) break;
                n = m;
            }
            const bool ch = n != rc;
            rc = n;
            if (!__any_sync(CPB_FULL, ch)) break;
        }
        unsigned hole = fr & ~rc;
        if (hole && pass != CPB_FILL_WRITE) t.misc[b] = 1;
        const u64 prio = (u64)((unsigned)(h * w)) << 32;
        while (hole && pass != CPB_FILL_DETECT) {
            const int c = __ffs((int)hole) - 1;
            hole &= hole - 1;
            atomicMax(&HK[(y0 + lane) * W + x0 + c], prio | (unsigned)l);
        }
    }
}

// ---- exact hole fill for tangled labels (CPB_FILL_EXACT, default off: validated on the simulator only) ----------
// The proposals above take every label's holes from the input image.  Upstream fills label by label in id order on
// the image as the earlier fills left it; the two differ only when a label loses SOME of its pixels to another
// label's fill and survives (it is then cut before its own turn).  k_fill_cover counts, per label, the pixels it
// loses to the proposals; k_fill_conflict flags the tiles in which a label loses some but not all of them; and
// k_fill_sequential replays upstream's loop on those tiles (one thread per tile: they are rare, and a sequential
// replay is the specification) and rewrites their proposal plane with the result.
CPB_KERNEL k_fill_cover(const int* CPB_RESTRICT lab, const u64* CPB_RESTRICT holekey, int H, int W, LabelTables t,
                        int* CPB_RESTRICT cover) {
    const int b = blockIdx.y;
    if (t.misc[b] == 0) return;                 // no proposal anywhere in this tile
    const int N = H * W, LC = t.LC;
    const int* L = lab + (size_t)b * N;
    const u64* HK = holekey + (size_t)b * N;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < N; r += gridDim.x * blockDim.x) {
        const u64 hk = HK[r];
        if (hk == 0) continue;
        const int o = L[r];
        if (o > 0 && o < LC && o != (int)(hk & 0xffffffffu)) atomicAdd(&cover[(size_t)b * LC + o], 1);
    }
}

CPB_KERNEL k_fill_conflict(LabelTables t, const int* CPB_RESTRICT cover, int* CPB_RESTRICT seq) {
    const int b = blockIdx.x;
    if (t.misc[b] == 0) return;
    const int LC = t.LC, lb = min(t.lbound[b], LC - 1);
    for (int l = 1 + threadIdx.x; l <= lb; l += blockDim.x) {
        const size_t k = (size_t)b * LC + l;
        const int c = cover[k];
        if (c > 0 && c < t.cnt[k] && (t.alive == nullptr || t.alive[k] != 0)) seq[b] = 1;
    }
}

// order_by_remap != 0 (fused path): the image holds raw labels and upstream's id of raw label l at this point is
// t.remap[l]; otherwise (stage call) the image holds those ids themselves.  state / mark / stack: three scratch
// planes of N ints per tile; inv: LC ints per tile.
CPB_KERNEL k_fill_sequential(const int* CPB_RESTRICT lab, int H, int W, LabelTables t, const int* CPB_RESTRICT seq,
                             int order_by_remap, int* CPB_RESTRICT state_, int* CPB_RESTRICT mark_,
                             int* CPB_RESTRICT stack_, int* CPB_RESTRICT inv_, u64* CPB_RESTRICT holekey) {
    const int b = blockIdx.x;
    if (seq[b] == 0 || threadIdx.x != 0) return;
    const int N = H * W, LC = t.LC, lb = min(t.lbound[b], LC - 1);
    const int* L = lab + (size_t)b * N;
    int* state = state_ + (size_t)b * N;
    int* mark = mark_ + (size_t)b * N;
    int* stack = stack_ + (size_t)b * N;
    int* inv = inv_ + (size_t)b * LC;
    u64* HK = holekey + (size_t)b * N;
    // labels on the image before the fill: everything the size filter in front of it left alive
    for (int p = 0; p < N; p++) {
        const int o = L[p];
        const bool on = o > 0 && o <= lb && (t.alive == nullptr || t.alive[(size_t)b * LC + o] != 0);
        state[p] = on ? o : 0;
        mark[p] = 0;
    }
    for (int i = 0; i <= lb; i++) inv[i] = 0;
    for (int l = 1; l <= lb; l++) {
        const size_t k = (size_t)b * LC + l;
        if (t.cnt[k] <= 0 || (t.alive != nullptr && t.alive[k] == 0)) continue;
        const int id = order_by_remap ? t.remap[k] : l;
        if (id >= 1 && id <= lb) inv[id] = l;
    }
    for (int id = 1; id <= lb; id++) {
        const int l = inv[id];
        if (l == 0) continue;
        const size_t k = (size_t)b * LC + l;
        const int y0 = t.ymin[k], y1 = t.ymax[k], x0 = t.xmin[k], x1 = t.xmax[k];
        if (y1 - y0 < 2 || x1 - x0 < 2) continue;             // a crop under 3 x 3 encloses nothing
        // flood the crop from its border through pixels that are not l (4-connectivity)
        int sp = 0;
        for (int y = y0; y <= y1; y++)
            for (int x = x0; x <= x1; x++) {
                if (y != y0 && y != y1 && x != x0 && x != x1) continue;
                const int p = y * W + x;
                if (state[p] != l && mark[p] != id) { mark[p] = id; stack[sp++] = p; }
            }
        while (sp > 0) {
            const int p = stack[--sp];
            const int y = p / W, x = p - y * W;
            if (y > y0) { const int q = p - W; if (state[q] != l && mark[q] != id) { mark[q] = id; stack[sp++] = q; } }
            if (y < y1) { const int q = p + W; if (state[q] != l && mark[q] != id) { mark[q] = id; stack[sp++] = q; } }
            if (x > x0) { const int q = p - 1; if (state[q] != l && mark[q] != id) { mark[q] = id; stack[sp++] = q; } }
            if (x < x1) { const int q = p + 1; if (state[q] != l && mark[q] != id) { mark[q] = id; stack[sp++] = q; } }
        }
        for (int y = y0; y <= y1; y++)
            for (int x = x0; x <= x1; x++) {
                const int p = y * W + x;
                if (state[p] != l && mark[p] != id) state[p] = l;      // enclosed: the label takes it
            }
    }
    // the proposal plane of the tile, rewritten: a pixel whose owner changed carries its new owner
    for (int p = 0; p < N; p++) {
        const int o = L[p];
        const bool on = o > 0 && o <= lb && (t.alive == nullptr || t.alive[(size_t)b * LC + o] != 0);
        const int was = on ? o : 0;
        HK[p] = state[p] != was ? ((1ull << 32) | (u64)(unsigned)state[p]) : 0ull;
    }
}

// k_vote: one block per tile.  Per-pixel arg-max over the C logits (first maximum wins),
// histogram (instance, class) in shared memory, per-instance arg-max (first maximum wins).
// Falls back to a global table when (lbound+1)*C ints exceed the shared-memory budget.
CPB_KERNEL CPB_LAUNCH_BOUNDS(512, 4)
k_vote(const int* CPB_RESTRICT lab, const float* CPB_RESTRICT logits, int H, int W, int C, int LC,
       const int* CPB_RESTRICT lbound, int smem_ints, int* CPB_RESTRICT gtable,
       int* CPB_RESTRICT cell_class, unsigned char* CPB_RESTRICT class_masks) {
    CPB_DYN_SMEM(int, s_tab);
    const int b = blockIdx.x, N = H * W;
    const int lb = min(lbound[b], LC - 1);
    const int* L = lab + (size_t)b * N;
    const float* G = logits + (size_t)b * C * N;
    int* cc = cell_class + (size_t)b * LC;
    const int need = (lb + 1) * C;
    int* tab = need <= smem_ints ? s_tab : gtable + (size_t)b * LC * C;
    for (int i = threadIdx.x; i < need; i += blockDim.x) tab[i] = 0;
    __syncthreads();
    for (int p = threadIdx.x; p < N; p += blockDim.x) {
        const int l = L[p];
        if (l <= 0 || l > lb) continue;
        float best = G[p];
        int arg = 0;
        for (int c = 1; c < C; c++) {
            const float v = G[(size_t)c * N + p];
            if (v > best) { best = v; arg = c; }
        }
        atomicAdd(&tab[l * C + arg], 1);
    }
    __syncthreads();
    for (int l = threadIdx.x; l <= lb; l += blockDim.x) {
        int arg = 0;
        if (l > 0) {
            int best = tab[l * C];
            for (int c = 1; c < C; c++) {
                const int v = tab[l * C + c];
                if (v > best) { best = v; arg = c; }
            }
        }
        cc[l] = arg;
    }
    if (class_masks) {
        __syncthreads();
        unsigned char* CM = class_masks + (size_t)b * N;
        for (int p = threadIdx.x; p < N; p += blockDim.x) {
            const int l = L[p];
            CM[p] = (l > 0 && l <= lb) ? (unsigned char)cc[l] : 0;
        }
    }
}

// k_border_flags: one block per tile; flag every label that owns a pixel on the tile border.
CPB_KERNEL k_border_flags(const int* CPB_RESTRICT lab, int H, int W, int nch, LabelTables t) {
    const int b = blockIdx.x, N = H * W;
    const int* L = lab + (size_t)b * N * nch;
    int* flag = t.flag + (size_t)b * t.LC;
    for (int i = threadIdx.x; i < 2 * (H + W); i += blockDim.x) {
        int y, x;
        if (i < W) { y = 0; x = i; }
        else if (i < 2 * W) { y = H - 1; x = i - W; }
        else if (i < 2 * W + H) { y = i - 2 * W; x = 0; }
        else { y = i - 2 * W - H; x = W - 1; }
        const int l = L[((size_t)y * W + x) * nch];
        if (l > 0 && l < t.LC) flag[l] = 1;
    }
}

// k_average_tiles: one thread per output element (b, ch, Y, X) of the cropped blend.
// Mirrors numpy's arithmetic in cellpose.transforms.average_tiles: the float32 accumulator
// is updated as float32(double(acc) + double(v) * w) per covering tile in tile order, the
// weight sum is float64, the final division is float32(double(acc) / Navg).
CPB_KERNEL CPB_LAUNCH_BOUNDS(256, 4)
k_average_tiles(const float* CPB_RESTRICT y, int B, int ntiles, int nch, int ly, int lx,
                const int* CPB_RESTRICT ty0, const int* CPB_RESTRICT tx0, const int* CPB_RESTRICT flip,
                int negate_flow, const double* CPB_RESTRICT taper_y, const double* CPB_RESTRICT taper_x,
                int cy0, int cx0, int oH, int oW, float* CPB_RESTRICT yf) {
    const long long total = (long long)B * nch * oH * oW;
    const long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= total) return;
    const int X = (int)(gi % oW);
    const int Y = (int)((gi / oW) % oH);
    const int ch = (int)((gi / ((long long)oW * oH)) % nch);
    const int b = (int)(gi / ((long long)oW * oH * nch));
    const int gy = Y + cy0, gx = X + cx0;
    float acc = 0.f;
    double navg = 0.0;
    for (int j = 0; j < ntiles; j++) {
        const int ry = gy - ty0[j], rx = gx - tx0[j];
        if (ry < 0 || ry >= ly || rx < 0 || rx >= lx) continue;
        const int f = flip[j];
        const int sy = (f & 1) ? ly - 1 - ry : ry;
        const int sx = (f & 2) ? lx - 1 - rx : rx;
        float v = y[((((size_t)b * ntiles + j) * nch + ch) * ly + sy) * lx + sx];
        if (negate_flow && ((ch == 0 && (f & 1)) || (ch == 1 && (f & 2)))) v = -v;
        const double wgt = __dmul_rn(taper_y[ry], taper_x[rx]);
        acc = (float)__dadd_rn((double)acc, __dmul_rn((double)v, wgt));
        navg = __dadd_rn(navg, wgt);
    }
    yf[gi] = (float)__ddiv_rn((double)acc, navg);
}

// k_average_tiles_v4: same arithmetic, one thread per 4 consecutive output pixels of one row and ALL channels
// (NCH >= nch accumulators of 4 floats live in registers).  Requires lx, crop offset, output width and every
// tile origin x0 to be multiples of 4, so that the four pixels are covered by exactly the same tiles and map
// to one aligned float4 of each tile (reversed when the tile is X-flipped).  Tiles are visited in tile order,
// the taper weights of a tile are formed once and reused for every channel.
// blocks per SM measured on B200: 4 accumulator channels -> 4 blocks (4.0 TB/s), 8 -> 2 blocks (3.7 TB/s), 16 -> 1
template <int NCH>
CPB_KERNEL CPB_LAUNCH_BOUNDS(256, (NCH > 8 ? 1 : (NCH > 4 ? 2 : 4)))
k_average_tiles_v4(const float* CPB_RESTRICT y, int B, int ntiles, int nch, int ly, int lx,
                   const int* CPB_RESTRICT ty0, const int* CPB_RESTRICT tx0, const int* CPB_RESTRICT flip,
                   int negate_flow, const double* CPB_RESTRICT taper_y, const double* CPB_RESTRICT taper_x,
                   int cy0, int cx0, int oH, int oW, float* CPB_RESTRICT yf) {
    const int oW4 = oW >> 2;
    const long long total = (long long)B * oH * oW4;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    const int X4 = (int)(g % oW4);
    const int Y = (int)((g / oW4) % oH);
    const int b = (int)(g / ((long long)oW4 * oH));
    const int gy = Y + cy0, gx = X4 * 4 + cx0;
    float acc[NCH][4];
    #pragma unroll
    for (int ch = 0; ch < NCH; ch++) { acc[ch][0] = 0.f; acc[ch][1] = 0.f; acc[ch][2] = 0.f; acc[ch][3] = 0.f; }
    double navg[4] = {0.0, 0.0, 0.0, 0.0};
    const size_t plane = (size_t)ly * lx;
    for (int j = 0; j < ntiles; j++) {
        const int ry = gy - ty0[j], rx = gx - tx0[j];
        if (ry < 0 || ry >= ly || rx < 0 || rx >= lx) continue;
        const int f = flip[j];
        const int sy = (f & 1) ? ly - 1 - ry : ry;
        const int sx = (f & 2) ? lx - 4 - rx : rx;           // first of the 4 source pixels (reversed if flipped)
        const float* src = y + ((size_t)b * ntiles + j) * nch * plane + (size_t)sy * lx + sx;
        const double wy = taper_y[ry];
        double wgt[4];
        #pragma unroll
        for (int k = 0; k < 4; k++) {
            wgt[k] = __dmul_rn(wy, taper_x[rx + k]);
            navg[k] = __dadd_rn(navg[k], wgt[k]);
        }
        constexpr int G = NCH < 8 ? NCH : 8;                 // channels loaded together
        #pragma unroll
        for (int c0 = 0; c0 < NCH; c0 += G) {
            float4 v4[G];
            #pragma unroll
            for (int q = 0; q < G; q++)
                if (c0 + q < nch) v4[q] = *reinterpret_cast<const float4*>(src + (c0 + q) * plane);
            #pragma unroll
            for (int q = 0; q < G; q++) {
                const int ch = c0 + q;
                if (ch < nch) {
                    float v[4];
                    if (f & 2) { v[0] = v4[q].w; v[1] = v4[q].z; v[2] = v4[q].y; v[3] = v4[q].x; }
                    else       { v[0] = v4[q].x; v[1] = v4[q].y; v[2] = v4[q].z; v[3] = v4[q].w; }
                    const bool neg = negate_flow && ((ch == 0 && (f & 1)) || (ch == 1 && (f & 2)));
                    #pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const float vv = neg ? -v[k] : v[k];
                        acc[ch][k] = (float)__dadd_rn((double)acc[ch][k], __dmul_rn((double)vv, wgt[k]));
                    }
                }
            }
        }
    }
    #pragma unroll
    for (int ch = 0; ch < NCH; ch++) {
        if (ch < nch) {
            float4 o;
            o.x = (float)__ddiv_rn((double)acc[ch][0], navg[0]);
            o.y = (float)__ddiv_rn((double)acc[ch][1], navg[1]);
            o.z = (float)__ddiv_rn((double)acc[ch][2], navg[2]);
            o.w = (float)__ddiv_rn((double)acc[ch][3], navg[3]);
            *reinterpret_cast<float4*>(yf + (((size_t)b * nch + ch) * oH + Y) * oW + X4 * 4) = o;
        }
    }
}

// ---- blend without float64 on the per-element path ---------------------------------------------------------------
// The reference accumulates acc = float32(double(acc) + double(v) * w) with a float64 weight w and ends with
// float32(double(acc) / Navg).  Converting every element to float64 and back keeps the conversion (XU) pipe at 60-70 %
// and the kernel at 40 % of the HBM roofline (profiles/r02: k_average_tiles_v4).  The same result comes out of float32
// arithmetic with error-free transformations:
//   w = wh + wl (two floats, |w - wh - wl| <= 2^-48 |w|),  v * wh = ph + pl exactly (one FMA),
//   acc + ph = s + e exactly (two-sum),  acc' = RN32(s + (e + pl + v * wl))
// which is the correctly rounded value of acc + v * w up to a perturbation of ~2^-23 ulp; numpy's double-rounded value can
// differ from it only when the exact sum lies that close to a rounding boundary (probability ~1e-6 per operation).  The
// final division uses r = 1 / Navg (float64, split the same way): acc * r = ph + pl exactly, out = RN32(ph + (pl + acc * rl)).
// The weight table (wh, wl)[ly][lx] and the per-pixel (rh, rl)[oH][oW] depend on the geometry only and are built by two
// tiny kernels per call.
CPB_KERNEL k_blend_weights(const double* CPB_RESTRICT taper_y, const double* CPB_RESTRICT taper_x, int ly, int lx,
                           float* CPB_RESTRICT wh, float* CPB_RESTRICT wl) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ly * lx) return;
    const int ry = i / lx, rx = i - ry * lx;
    const double w = __dmul_rn(taper_y[ry], taper_x[rx]);
    const float h = (float)w;
    wh[i] = h; wl[i] = (float)__dsub_rn(w, (double)h);
}

CPB_KERNEL k_blend_rinv(int ntiles, int ly, int lx, const int* CPB_RESTRICT ty0, const int* CPB_RESTRICT tx0,
                        const double* CPB_RESTRICT taper_y, const double* CPB_RESTRICT taper_x, int cy0, int cx0, int oH, int oW,
                        float* CPB_RESTRICT rh, float* CPB_RESTRICT rl) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= oH * oW) return;
    const int Y = i / oW, X = i - Y * oW;
    const int gy = Y + cy0, gx = X + cx0;
    double navg = 0.0;
    for (int j = 0; j < ntiles; j++) {
        const int ry = gy - ty0[j], rx = gx - tx0[j];
        if (ry < 0 || ry >= ly || rx < 0 || rx >= lx) continue;
        navg = __dadd_rn(navg, __dmul_rn(taper_y[ry], taper_x[rx]));
    }
    const double r = __ddiv_rn(1.0, navg);
    const float h = (float)r;
    rh[i] = h; rl[i] = (float)__dsub_rn(r, (double)h);
}

// two pixels per instruction (packed f32x2: every half is an ordinary IEEE round-to-nearest operation).
// nwh = -wh, so that v * nwh = -(v * wh) exactly and pl = fma(v, wh, -ph) needs no packed negation.
// ph is formed by SCALAR multiplies: ptxas contracts a packed mul.rn.f32x2 that feeds a packed add / sub into FFMA2
// (despite the .rn), which would turn s = RN(acc + RN(v * wh)) into RN(acc + v * wh) and break the two-sum that follows
// -- seen on the B200: only 83 % of the elements were bit-identical to numpy with the packed multiply, all of them
// with the scalar one.
CPB_DEVICE void cpb_eft_acc2(pf2& acc, pf2 v, pf2 wh, pf2 nwh, pf2 wl) {
    float v0, v1, h0, h1;
    pf2_get(v, v0, v1); pf2_get(wh, h0, h1);
    const pf2 ph = pf2_make(__fmul_rn(v0, h0), __fmul_rn(v1, h1));
    pf2 pl = pf2_fma(v, wh, pf2_mul(v, nwh));
    pl = pf2_fma(v, wl, pl);
    const pf2 s = pf2_add(acc, ph);
    const pf2 bb = pf2_sub(s, acc);
    const pf2 e = pf2_add(pf2_sub(acc, pf2_sub(s, bb)), pf2_sub(ph, bb));
    acc = pf2_add(s, pf2_add(e, pl));
}

CPB_DEVICE float cpb_eft_scale(float acc, float rh, float rl) {
    const float ph = __fmul_rn(acc, rh);
    float pl = __fmaf_rn(acc, rh, -ph);
    pl = __fmaf_rn(acc, rl, pl);
    return __fadd_rn(ph, pl);
}

// one thread per 4 consecutive output pixels of one row and NCH channels starting at c0 (same geometry requirements as
// k_average_tiles_v4: lx, crop offset, output width and every window origin x0 multiples of 4).  The flip / sign cases
// are warp-uniform branches (every thread of a warp sees the same tile), a sign change is folded into the weights
// ((-v) * w == v * (-w) exactly).
template <int NCH>
CPB_KERNEL CPB_LAUNCH_BOUNDS(256, (NCH > 4 ? 2 : 4))
k_average_tiles_eft(const float* CPB_RESTRICT y, int B, int ntiles, int nch, int c0, int ly, int lx,
                    const int* CPB_RESTRICT ty0, const int* CPB_RESTRICT tx0, const int* CPB_RESTRICT flip, int negate_flow,
                    const float* CPB_RESTRICT wh, const float* CPB_RESTRICT wl, const float* CPB_RESTRICT rh,
                    const float* CPB_RESTRICT rl, int cy0, int cx0, int oH, int oW, float* CPB_RESTRICT yf) {
    const int oW4 = oW >> 2;
    const long long total = (long long)B * oH * oW4;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    const int X4 = (int)(g % oW4);
    const int Y = (int)((g / oW4) % oH);
    const int b = (int)(g / ((long long)oW4 * oH));
    const int gy = Y + cy0, gx = X4 * 4 + cx0;
    pf2 acc[NCH][2];
    #pragma unroll
    for (int q = 0; q < NCH; q++) { acc[q][0] = pf2_make(0.f, 0.f); acc[q][1] = pf2_make(0.f, 0.f); }
    const size_t plane = (size_t)ly * lx;
    for (int j = 0; j < ntiles; j++) {
        const int ry = gy - ty0[j], rx = gx - tx0[j];
        if (ry < 0 || ry >= ly || rx < 0 || rx >= lx) continue;
        const int f = flip[j];
        const int sy = (f & 1) ? ly - 1 - ry : ry;
        const int sx = (f & 2) ? lx - 4 - rx : rx;           // first of the 4 source pixels (reversed if flipped)
        const float* src = y + (((size_t)b * ntiles + j) * nch + c0) * plane + (size_t)sy * lx + sx;
        float4 v4[NCH];
        #pragma unroll
        for (int q = 0; q < NCH; q++) v4[q] = *reinterpret_cast<const float4*>(src + q * plane);
        const float4 h4 = *reinterpret_cast<const float4*>(wh + (size_t)ry * lx + rx);
        const float4 l4 = *reinterpret_cast<const float4*>(wl + (size_t)ry * lx + rx);
        const pf2 h01 = pf2_make(h4.x, h4.y), h23 = pf2_make(h4.z, h4.w), l01 = pf2_make(l4.x, l4.y), l23 = pf2_make(l4.z, l4.w);
        const pf2 n01 = pf2_make(-h4.x, -h4.y), n23 = pf2_make(-h4.z, -h4.w), m01 = pf2_make(-l4.x, -l4.y), m23 = pf2_make(-l4.z, -l4.w);
        #pragma unroll
        for (int q = 0; q < NCH; q++) {
            const int ch = c0 + q;
            const bool neg = negate_flow && ((ch == 0 && (f & 1)) || (ch == 1 && (f & 2)));       // warp-uniform
            pf2 va, vb;
            if (f & 2) { va = pf2_make(v4[q].w, v4[q].z); vb = pf2_make(v4[q].y, v4[q].x); }
            else       { va = pf2_make(v4[q].x, v4[q].y); vb = pf2_make(v4[q].z, v4[q].w); }
            if (neg) { cpb_eft_acc2(acc[q][0], va, n01, h01, m01); cpb_eft_acc2(acc[q][1], vb, n23, h23, m23); }
            else     { cpb_eft_acc2(acc[q][0], va, h01, n01, l01); cpb_eft_acc2(acc[q][1], vb, h23, n23, l23); }
        }
    }
    const float4 r_h = *reinterpret_cast<const float4*>(rh + (size_t)Y * oW + X4 * 4);
    const float4 r_l = *reinterpret_cast<const float4*>(rl + (size_t)Y * oW + X4 * 4);
    #pragma unroll
    for (int q = 0; q < NCH; q++) {
        float a0, a1, a2, a3;
        pf2_get(acc[q][0], a0, a1); pf2_get(acc[q][1], a2, a3);
        float4 o;
        o.x = cpb_eft_scale(a0, r_h.x, r_l.x);
        o.y = cpb_eft_scale(a1, r_h.y, r_l.y);
        o.z = cpb_eft_scale(a2, r_h.z, r_l.z);
        o.w = cpb_eft_scale(a3, r_h.w, r_l.w);
        *reinterpret_cast<float4*>(yf + (((size_t)b * nch + c0 + q) * oH + Y) * oW + X4 * 4) = o;
    }
}

// ---- blend of the flow map FUSED with the cellprob threshold (north_star (1); SURVEY 8b: average_tiles(...,
// cellprob_threshold) -> yf + foreground) ---------------------------------------------------------------------------
// The flow map of run_net has three channels (dY, dX, cellprob: core.py:215-216).  One thread blends all three for its
// 4 pixels (error-free float32 accumulate as above) and, holding them in registers, also does everything the first
// pass of the mask path would re-read them for: foreground = cellprob > threshold, the masked / scaled zero-padded flow
// field of follow_flows, the compacted foreground list (block scan + one atomic per block, blocks in 16 x 64 patch
// order so that a run of the list holds whole cells) and the zeroed label image.  dP / cellprob are written once, for
// the flow check and the caller; k_prep_flow_v4 never runs on this path.
// Requirements (checked by the host): lx, crop offset, every x0 multiples of 4, oW % 64 == 0.
CPB_KERNEL CPB_LAUNCH_BOUNDS(256, 4)
k_blend_prep(const float* CPB_RESTRICT y, int B, int ntiles, int ly, int lx, const int* CPB_RESTRICT ty0,
             const int* CPB_RESTRICT tx0, const int* CPB_RESTRICT flip, int negate_flow, const float* CPB_RESTRICT wh,
             const float* CPB_RESTRICT wl, const float* CPB_RESTRICT rh, const float* CPB_RESTRICT rl, int cy0, int cx0,
             int oH, int oW, float thr, float sxs, float sys, float* CPB_RESTRICT dP, float* CPB_RESTRICT cellprob,
             float4* CPB_RESTRICT flow, int4* CPB_RESTRICT labels, unsigned* CPB_RESTRICT list, unsigned* CPB_RESTRICT list_n) {
    CPB_SHARED int s_scan[33];
    CPB_SHARED unsigned s_base;
    constexpr int NCH = 3;
    const int W4 = oW >> 2, N = oH * oW;
    const int Wp4 = (oW + 2 * CPB_FLOW_PADX) >> 1;           // padded row pitch in float4 (2 pixels each)
    const int pbx = W4 >> 4, pby = (oH + 15) >> 4;
    const int blk = blockIdx.x;
    const int b = blk / (pbx * pby);
    const int rem = blk - b * (pbx * pby);
    const int Y = (rem / pbx) * 16 + (threadIdx.x >> 4);
    const int X4 = (rem % pbx) * 16 + (threadIdx.x & 15);
    const bool in = b < B && Y < oH;
    int nfg = 0;
    unsigned gi0 = 0;
    bool f0 = false, f1 = false, f2 = false, f3 = false;
    if (in) {
        const int gy = Y + cy0, gx = X4 * 4 + cx0;
        pf2 acc[NCH][2];
        #pragma unroll
        for (int q = 0; q < NCH; q++) { acc[q][0] = pf2_make(0.f, 0.f); acc[q][1] = pf2_make(0.f, 0.f); }
        const size_t plane = (size_t)ly * lx;
        for (int j = 0; j < ntiles; j++) {
            const int ry = gy - ty0[j], rx = gx - tx0[j];
            if (ry < 0 || ry >= ly || rx < 0 || rx >= lx) continue;
            const int f = flip[j];
            const int sy = (f & 1) ? ly - 1 - ry : ry;
            const int sx = (f & 2) ? lx - 4 - rx : rx;
            const float* src = y + ((size_t)b * ntiles + j) * NCH * plane + (size_t)sy * lx + sx;
            float4 v4[NCH];
            #pragma unroll
            for (int q = 0; q < NCH; q++) v4[q] = *reinterpret_cast<const float4*>(src + q * plane);
            const float4 h4 = *reinterpret_cast<const float4*>(wh + (size_t)ry * lx + rx);
            const float4 l4 = *reinterpret_cast<const float4*>(wl + (size_t)ry * lx + rx);
            const pf2 h01 = pf2_make(h4.x, h4.y), h23 = pf2_make(h4.z, h4.w), l01 = pf2_make(l4.x, l4.y), l23 = pf2_make(l4.z, l4.w);
            const pf2 n01 = pf2_make(-h4.x, -h4.y), n23 = pf2_make(-h4.z, -h4.w), m01 = pf2_make(-l4.x, -l4.y), m23 = pf2_make(-l4.z, -l4.w);
            #pragma unroll
            for (int q = 0; q < NCH; q++) {
                const bool neg = negate_flow && ((q == 0 && (f & 1)) || (q == 1 && (f & 2)));
                pf2 va, vb;
                if (f & 2) { va = pf2_make(v4[q].w, v4[q].z); vb = pf2_make(v4[q].y, v4[q].x); }
                else       { va = pf2_make(v4[q].x, v4[q].y); vb = pf2_make(v4[q].z, v4[q].w); }
                if (neg) { cpb_eft_acc2(acc[q][0], va, n01, h01, m01); cpb_eft_acc2(acc[q][1], vb, n23, h23, m23); }
                else     { cpb_eft_acc2(acc[q][0], va, h01, n01, l01); cpb_eft_acc2(acc[q][1], vb, h23, n23, l23); }
            }
        }
        const float4 r_h = *reinterpret_cast<const float4*>(rh + (size_t)Y * oW + X4 * 4);
        const float4 r_l = *reinterpret_cast<const float4*>(rl + (size_t)Y * oW + X4 * 4);
        float4 o[NCH];
        #pragma unroll
        for (int q = 0; q < NCH; q++) {
            float a0, a1, a2, a3;
            pf2_get(acc[q][0], a0, a1); pf2_get(acc[q][1], a2, a3);
            o[q].x = cpb_eft_scale(a0, r_h.x, r_l.x); o[q].y = cpb_eft_scale(a1, r_h.y, r_l.y);
            o[q].z = cpb_eft_scale(a2, r_h.z, r_l.z); o[q].w = cpb_eft_scale(a3, r_h.w, r_l.w);
        }
        const size_t pix = (size_t)Y * oW + X4 * 4;
        *reinterpret_cast<float4*>(dP + ((size_t)b * 2 + 0) * N + pix) = o[0];
        *reinterpret_cast<float4*>(dP + ((size_t)b * 2 + 1) * N + pix) = o[1];
        *reinterpret_cast<float4*>(cellprob + (size_t)b * N + pix) = o[2];
        // ---- what k_prep_flow_v4 does with these three values
        f0 = o[2].x > thr; f1 = o[2].y > thr; f2 = o[2].z > thr; f3 = o[2].w > thr;
        const float2 a0 = cpb_scaled_flow(o[0].x, o[1].x, f0, sxs, sys), a1 = cpb_scaled_flow(o[0].y, o[1].y, f1, sxs, sys);
        const float2 a2 = cpb_scaled_flow(o[0].z, o[1].z, f2, sxs, sys), a3 = cpb_scaled_flow(o[0].w, o[1].w, f3, sxs, sys);
        float4* row = flow + ((size_t)b * (oH + 2) + Y + 1) * Wp4;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        row[1 + 2 * X4] = make_float4(a0.x, a0.y, a1.x, a1.y);
        row[2 + 2 * X4] = make_float4(a2.x, a2.y, a3.x, a3.y);
        if (X4 == 0) row[0] = z;
        if (X4 == W4 - 1) row[Wp4 - 1] = z;
        if (Y == 0 || Y == oH - 1) {                                    // the zero rows above / below the tile
            float4* prow = flow + ((size_t)b * (oH + 2) + (Y == 0 ? 0 : oH + 1)) * Wp4;
            prow[1 + 2 * X4] = z; prow[2 + 2 * X4] = z;
            if (X4 == 0) prow[0] = z;
            if (X4 == W4 - 1) prow[Wp4 - 1] = z;
        }
        labels[((size_t)b * N + pix) >> 2] = make_int4(0, 0, 0, 0);
        gi0 = (unsigned)b * (unsigned)N + (unsigned)pix;
        nfg = (int)f0 + (int)f1 + (int)f2 + (int)f3;
    }
    int tot;
    const int incl = cpb_block_scan_incl(nfg, s_scan, &tot);
    if (threadIdx.x == 0 && tot > 0) s_base = atomicAdd(list_n, (unsigned)tot);
    __syncthreads();
    unsigned op = s_base + (unsigned)(incl - nfg);
    if (f0) list[op++] = gi0;
    if (f1) list[op++] = gi0 + 1;
    if (f2) list[op++] = gi0 + 2;
    if (f3) list[op++] = gi0 + 3;
}

// k_label_offsets: single block; offsets[b] = base + sum(counts[0..b)), total = sum(counts).
CPB_KERNEL k_label_offsets(const int* CPB_RESTRICT counts, int B, long long base,
                           long long* CPB_RESTRICT offsets, long long* CPB_RESTRICT total) {
    CPB_SHARED int s_scan[33];
    CPB_SHARED long long s_carry;
    if (threadIdx.x == 0) s_carry = base;
    __syncthreads();
    for (int b0 = 0; b0 < B; b0 += blockDim.x) {
        const int b = b0 + threadIdx.x;
        const int c = b < B ? counts[b] : 0;
        int tot;
        const int incl = cpb_block_scan_incl(c, s_scan, &tot);
        if (b < B) offsets[b] = s_carry + incl - c;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = s_carry - base;
}
