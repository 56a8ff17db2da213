// Row (4): remove_bad_flow_masks -> flow_error -> masks_to_flows.
//   k_centres : per label, the label pixel nearest to the label's mean position, ext
//   k_diffuse : per label, float64 Jacobi heat diffusion from the centre, n_iter = 2*max(ext)
//   k_flow_err: per label, gradient of T (raw, neighbours of other labels leak in), unit
//               vectors, mean squared difference to dP/5, bad = err > threshold
// One block per label, labels of a tile strided over gridDim.x, tile = blockIdx.y.
#pragma once
#include "cpb_common.cuh"

#define CPB_QC_THREADS 128
#define CPB_DIFF_SMEM_CELLS 2304   // (bbox_h+2)*(bbox_w+2) cells that fit the shared-memory path

#define CPB_DW_WARPS 4         // warp path: labels in flight per block (one warp each)
#define CPB_DC_MAXH 30         // warp path: bbox up to 30 rows x 32 columns
#define CPB_DC_MIDH 22         // ... in two size classes (<= 22 rows: 8 blocks of 4 warps per SM; <= 30 rows: 6)
#define CPB_DC_MAXW 32
#define CPB_DC_PITCH 34        // + one halo column on each side

CPB_DEVICE bool cpb_diffuse_is_small(int h, int w) { return h <= CPB_DC_MAXH && w <= CPB_DC_MAXW; }

struct MinKey { double d; int idx; };

CPB_DEVICE bool cpb_minkey_less(double d0, int i0, double d1, int i1) {
    return d0 < d1 || (d0 == d1 && i0 < i1);
}

// k_qc_scan: one block per tile over the label tables.  n_iter of the tile (2 * max ext over live labels) and the
// list of labels whose bbox is beyond the warp kernels (they need the block kernels); count[0] / count[1] must be 0.
CPB_KERNEL k_qc_scan(LabelTables t, int2* CPB_RESTRICT list, int* CPB_RESTRICT count) {
    const int b = blockIdx.x, LC = t.LC;
    const int lb = t.lbound[b];
    int ext = 0;
    for (int l = 1 + threadIdx.x; l <= lb; l += blockDim.x) {
        const size_t k = (size_t)b * LC + l;
        if (!cpb_label_live(t, k)) continue;
        const int h = t.ymax[k] - t.ymin[k] + 1, w = t.xmax[k] - t.xmin[k] + 1;
        ext = max(ext, 2 * (h + w + 2));
        if (!cpb_diffuse_is_small(h, w)) list[atomicAdd(count, 1)] = make_int2(b, l);
    }
    for (int s = 16; s; s >>= 1) ext = max(ext, __shfl_xor_sync(CPB_FULL, ext, s));
    if ((threadIdx.x & 31) == 0 && ext > 0) atomicMax(&t.niter[b], ext);
}

CPB_KERNEL CPB_LAUNCH_BOUNDS(CPB_QC_THREADS, 8)
k_centres(const int* CPB_RESTRICT lab, int H, int W, LabelTables t, int skip_small, LabelWork wk) {
    CPB_SHARED double s_d[CPB_QC_THREADS / 32];
    CPB_SHARED int s_i[CPB_QC_THREADS / 32];
    const int LC = t.LC, N = H * W;
    int it = 0, b, l;
    while (cpb_next_label(wk, t.lbound, it, b, l)) {
        const int* L = lab + (size_t)b * N;
        const size_t k = (size_t)b * LC + l;
        const int c = t.cnt[k];
        if (!cpb_label_live(t, k)) continue;   // block-uniform
        const int y0 = t.ymin[k], x0 = t.xmin[k];
        const int h = t.ymax[k] - y0 + 1, w = t.xmax[k] - x0 + 1;
        if (skip_small && cpb_diffuse_is_small(h, w)) {     // centre computed by the warp kernel itself
            if (threadIdx.x == 0) atomicMax(&t.niter[b], 2 * (h + w + 2));
            continue;
        }
        // means of (coordinate relative to the bbox + 1), as the reference computes them
        const double ymed = __ddiv_rn(__ll2double_rn((long long)t.sumy[k] - (long long)c * y0 + c), __int2double_rn(c));
        const double xmed = __ddiv_rn(__ll2double_rn((long long)t.sumx[k] - (long long)c * x0 + c), __int2double_rn(c));
        double bd = 1e300; int bi = CPB_IMAX;
        for (int i = threadIdx.x; i < h * w; i += blockDim.x) {
            const int ry = i / w, rx = i - ry * w;
            if (L[(y0 + ry) * W + x0 + rx] == l) {
                const double dx = __dsub_rn(__int2double_rn(rx + 1), xmed);
                const double dy = __dsub_rn(__int2double_rn(ry + 1), ymed);
                const double d = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                if (cpb_minkey_less(d, i, bd, bi)) { bd = d; bi = i; }
            }
        }
        for (int s = 16; s; s >>= 1) {
            const double od = __shfl_xor_sync(CPB_FULL, bd, s);
            const int oi = __shfl_xor_sync(CPB_FULL, bi, s);
            if (cpb_minkey_less(od, oi, bd, bi)) { bd = od; bi = oi; }
        }
        __syncthreads();
        if ((threadIdx.x & 31) == 0) { s_d[threadIdx.x >> 5] = bd; s_i[threadIdx.x >> 5] = bi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int q = 1; q < CPB_QC_THREADS / 32; q++)
                if (cpb_minkey_less(s_d[q], s_i[q], bd, bi)) { bd = s_d[q]; bi = s_i[q]; }
            const int ry = bi / w, rx = bi - ry * w;
            t.cy[k] = y0 + ry; t.cx[k] = x0 + rx;
            atomicMax(&t.niter[b], 2 * (h + w + 2));
        }
    }
}

// Exact x/9 in three ops instead of the ~25-instruction IEEE division sequence: with r = RN(1/9) and
// q = RN(x*r), the residual x - 9q is exact in one FMA and RN(q + rem*r) is the correctly rounded
// quotient (Markstein); checked against true division on 6e8 random operands.  Outside
// the safe exponent range fall back to the real division.
CPB_DEVICE double cpb_div9(double x) {
    const double ax = fabs(x);
    if (ax < 1e-280 || ax > 1e280) return __ddiv_rn(x, 9.0);
    const double r = 1.0 / 9.0;
    const double q = __dmul_rn(x, r);
    const double rem = __fma_rn(-9.0, q, x);
    return __fma_rn(rem, r, q);
}

// Branch-free variant for operands known to stay far from the subnormal / overflow range (the warp
// kernel: inside a 32 x 32 bbox T is 0 or within [9^-64, n_iter]).
CPB_DEVICE double cpb_div9_fast(double x) {
    const double r = 1.0 / 9.0;
    const double q = __dmul_rn(x, r);
    const double rem = __fma_rn(-9.0, q, x);
    return __fma_rn(rem, r, q);
}


// k_diffuse_warp: one WARP per label -- or per PAIR of labels -- for labels whose bbox fits 30 x 32 (nuclei-sized).
// The label's T lives in a per-warp shared-memory tile with a zero halo; lane j owns column j and walks down the
// rows with a 3 x 3 sliding window in registers: conflict-free row loads, nine neighbours summed in the
// reference's order, in-place Jacobi update (a row is written only after every lane holds the rows it still
// needs), two rows per step.  The kernel is bound by the float64 pipe and its cost is (rows x 11 warp
// instructions) whatever the number of busy lanes, so two consecutive labels whose widths fit side by side
// (wA + 1 + wB <= 32, a zero column between them) share one pass: same arithmetic per label, half the issue slots.
struct DiffSub { int l; size_t k; int y0, x0, h, w, coff; };

// Fused flow error: a label whose bbox grown by one pixel holds no pixel of another live label ("clean") never
// sees foreign T in its gradient and nobody reads its T, so its error is taken straight from the shared-memory
// tile (same per-pixel arithmetic as k_flow_err) and T is not written to global memory at all.  Labels in
// contact with another live label write T as before and are left to k_flow_err (t.done stays 0).
struct DiffQC { const float* dPy; const float* dPx; const int* alive; double threshold; int H;
                int2* list; int* count; int cap; int b; };    // labels left to k_flow_err go to the END of `list`

CPB_DEVICE bool cpb_foreign_live(int v, int l, const int* CPB_RESTRICT alive) {
    return v > 0 && v != l && (alive == nullptr || alive[v] != 0);
}

// Common head of a warp job: membership bit masks (lane = column, bit = row), contact with other live labels
// (bbox grown by one pixel), diffusion centres.  Sub-label q occupies lanes [coff, coff + w).
struct DiffPro { unsigned member; bool clean[2]; int cr[2], cl[2]; };   // centre row (bbox-relative) and lane

CPB_DEVICE void cpb_diffuse_prologue(const int* CPB_RESTRICT L, int W, const LabelTables& t, const DiffSub& A,
                                     const DiffSub& B, bool has_b, const DiffQC& qc, DiffPro& o) {
    const int lane = threadIdx.x & 31;
    const bool inA = lane >= A.coff && lane < A.coff + A.w;
    const bool inB = has_b && lane >= B.coff && lane < B.coff + B.w;
    const DiffSub& my = inB ? B : A;
    const bool mine = inA || inB;
    const int col = lane - my.coff;
    const bool fuse = qc.dPy != nullptr;
    unsigned member = 0;               // bit r: pixel (y0+r, x0+col) belongs to this lane's label
    bool foreign = false;              // a pixel of another live label inside the bbox grown by one
    if (mine) {
        const int x = my.x0 + col;
        for (int r = 0; r < my.h; r++) {
            const int v = L[(my.y0 + r) * W + x];
            if (v == my.l) member |= 1u << r;
            else if (fuse) foreign |= cpb_foreign_live(v, my.l, qc.alive);
        }
        if (fuse) {
            if (my.y0 > 0) foreign |= cpb_foreign_live(L[(my.y0 - 1) * W + x], my.l, qc.alive);
            if (my.y0 + my.h < qc.H) foreign |= cpb_foreign_live(L[(my.y0 + my.h) * W + x], my.l, qc.alive);
        }
    }
    o.member = member;
    o.clean[0] = o.clean[1] = false;
    if (fuse) {
        for (int q = 0; q < (has_b ? 2 : 1); q++) {
            const DiffSub& sub = q ? B : A;
            bool f = (q ? inB : inA) && foreign;
            const int y = sub.y0 - 1 + lane;                     // halo columns: lane -> row of the grown bbox
            if (lane < sub.h + 2 && y >= 0 && y < qc.H) {
                if (sub.x0 > 0) f |= cpb_foreign_live(L[y * W + sub.x0 - 1], sub.l, qc.alive);
                if (sub.x0 + sub.w < W) f |= cpb_foreign_live(L[y * W + sub.x0 + sub.w], sub.l, qc.alive);
            }
            o.clean[q] = !__any_sync(CPB_FULL, f);
        }
    }
    // centres: member pixel nearest to the mean of (bbox-relative coordinate + 1); first in raster order on ties
    for (int q = 0; q < (has_b ? 2 : 1); q++) {
        const DiffSub& sub = q ? B : A;
        const bool in = q ? inB : inA;
        const int c = t.cnt[sub.k];
        const double ymed = __ddiv_rn(__ll2double_rn((long long)t.sumy[sub.k] - (long long)c * sub.y0 + c), __int2double_rn(c));
        const double xmed = __ddiv_rn(__ll2double_rn((long long)t.sumx[sub.k] - (long long)c * sub.x0 + c), __int2double_rn(c));
        const double dx = __dsub_rn(__int2double_rn(col + 1), xmed);
        const double dx2 = __dmul_rn(dx, dx);
        double bd = 1e300; int bi = CPB_IMAX;
        if (in)
            for (int r = 0; r < sub.h; r++) {
                if (member >> r & 1) {
                    const double dy = __dsub_rn(__int2double_rn(r + 1), ymed);
                    const double d = __dadd_rn(dx2, __dmul_rn(dy, dy));
                    const int idx = r * sub.w + col;
                    if (cpb_minkey_less(d, idx, bd, bi)) { bd = d; bi = idx; }
                }
            }
        for (int sft = 16; sft; sft >>= 1) {
            const double od = __shfl_xor_sync(CPB_FULL, bd, sft);
            const int oi = __shfl_xor_sync(CPB_FULL, bi, sft);
            if (cpb_minkey_less(od, oi, bd, bi)) { bd = od; bi = oi; }
        }
        const int cr = bi / sub.w, cc = bi - cr * sub.w;
        if (lane == 0) { t.cy[sub.k] = sub.y0 + cr; t.cx[sub.k] = sub.x0 + cc; }
        o.cr[q] = cr; o.cl[q] = sub.coff + cc;
    }
}

// flow error of one clean label from per-lane partial sums (lanes outside the sub-label contribute 0)
CPB_DEVICE void cpb_diffuse_publish_err(const LabelTables& t, const DiffSub& sub, bool in, double ey, double ex,
                                        double threshold) {
    double sy = in ? ey : 0.0, sx = in ? ex : 0.0;
    for (int sft = 16; sft; sft >>= 1) {
        sy += __shfl_xor_sync(CPB_FULL, sy, sft);
        sx += __shfl_xor_sync(CPB_FULL, sx, sft);
    }
    if ((threadIdx.x & 31) == 0) {
        const double c = (double)t.cnt[sub.k];
        const double e = sy / c + sx / c;
        t.err[sub.k] = e;
        t.flag[sub.k] = e > threshold ? 1 : 0;
        t.done[sub.k] = 1;
    }
}

// one pixel's contribution to the flow error: gradient of T -> unit vector -> squared difference to dP / 5
CPB_DEVICE void cpb_flow_err_pixel(double dy, double dx, float dpy, float dpx, double& ey, double& ex) {
    const double nrm = __dadd_rn(1e-60, __dsqrt_rn(__dadd_rn(__dmul_rn(dy, dy), __dmul_rn(dx, dx))));
    const double muy = __ddiv_rn(dy, nrm), mux = __ddiv_rn(dx, nrm);
    const double ry = __dsub_rn(muy, (double)__fdiv_rn(dpy, 5.0f));
    const double rx = __dsub_rn(mux, (double)__fdiv_rn(dpx, 5.0f));
    ey += __dmul_rn(ry, ry);
    ex += __dmul_rn(rx, rx);
}

template <int R>      // rows per step of the sliding window (independent add chains in flight per lane)
CPB_DEVICE void cpb_diffuse_job(const int* CPB_RESTRICT L, int W, const LabelTables& t, double* CPB_RESTRICT Tb,
                                double* S, const DiffSub& A, const DiffSub& B, bool has_b, int n_it, const DiffQC& qc) {
    const int lane = threadIdx.x & 31;
    const bool inA = lane < A.w;
    const bool inB = has_b && lane >= B.coff && lane < B.coff + B.w;
    const DiffSub& my = inB ? B : A;
    const bool mine = inA || inB;
    const int col = lane - my.coff;
    const int hj = has_b ? max(A.h, B.h) : A.h;
    const bool fuse = qc.dPy != nullptr;
    __syncwarp();
    for (int i = lane; i < (hj + R + 1) * CPB_DC_PITCH; i += 32) S[i] = 0.0;
    DiffPro pro;
    cpb_diffuse_prologue(L, W, t, A, B, has_b, qc, pro);
    const unsigned member = pro.member;
    const bool clean[2] = {pro.clean[0], pro.clean[1]};
    if (qc.list && lane == 0) {            // whatever this warp does not finish itself is queued for k_flow_err
        if (!clean[0]) qc.list[qc.cap - 1 - atomicAdd(qc.count + 1, 1)] = make_int2(qc.b, A.l);
        if (has_b && !clean[1]) qc.list[qc.cap - 1 - atomicAdd(qc.count + 1, 1)] = make_int2(qc.b, B.l);
    }
    const int ci[2] = {(pro.cr[0] + 1) * CPB_DC_PITCH + pro.cl[0] + 1, (pro.cr[1] + 1) * CPB_DC_PITCH + pro.cl[1] + 1};
    const double* p = S + lane;        // p[0], p[1], p[2] = columns j-1, j, j+1 of the halo row
    double* own = S + CPB_DC_PITCH + lane + 1;
    __syncwarp();
#ifndef CPB_DIFFUSE_FRONT
#define CPB_DIFFUSE_FRONT 0        // measured on B200: 2.52 ms with the restriction, 2.43 ms without (variable loop bounds)
#endif
#if CPB_DIFFUSE_FRONT
    const int c_lo = has_b ? min(pro.cr[0], pro.cr[1]) : pro.cr[0], c_hi = has_b ? max(pro.cr[0], pro.cr[1]) : pro.cr[0];
#endif
    for (int it = 0; it < n_it; it++) {
        if (lane == 0) { S[ci[0]] += 1.0; if (has_b) S[ci[1]] += 1.0; }   // T[centre] += 1 before averaging
        __syncwarp();
        // the heat front moves one row per iteration: rows further than it + 1 from a centre still have an all-zero
        // 3 x 3 neighbourhood and stay exactly 0, so the first iterations only walk the rows the front has reached
#if CPB_DIFFUSE_FRONT
        const int r_lo = max(0, c_lo - it - 1) & ~(R - 1), r_hi = min(hj, c_hi + it + 2);
#else
        const int r_lo = 0, r_hi = hj;
#endif
        // win[k] = row r-1+k of the tile (left, centre, right of this lane's column)
        double win[R + 2][3];
        #pragma unroll
        for (int k = 0; k < 2; k++) {
            const double* q0 = p + (r_lo + k) * CPB_DC_PITCH;
            win[k][0] = q0[0]; win[k][1] = q0[1]; win[k][2] = q0[2];
        }
        for (int r = r_lo; r < r_hi; r += R) {
            const double* q = p + (r + 2) * CPB_DC_PITCH;
            #pragma unroll
            for (int k = 0; k < R; k++) {                       // rows r+1 .. r+R
                win[k + 2][0] = q[k * CPB_DC_PITCH]; win[k + 2][1] = q[k * CPB_DC_PITCH + 1]; win[k + 2][2] = q[k * CPB_DC_PITCH + 2];
            }
            double v[R];
            #pragma unroll
            for (int k = 0; k < R; k++) {
                // self, up, down, left, right, up-left, up-right, down-left, down-right
                double s0 = __dadd_rn(win[k + 1][1], win[k][1]);
                s0 = __dadd_rn(s0, win[k + 2][1]);
                s0 = __dadd_rn(s0, win[k + 1][0]);
                s0 = __dadd_rn(s0, win[k + 1][2]);
                s0 = __dadd_rn(s0, win[k][0]);
                s0 = __dadd_rn(s0, win[k][2]);
                s0 = __dadd_rn(s0, win[k + 2][0]);
                s0 = __dadd_rn(s0, win[k + 2][2]);
                v[k] = cpb_div9_fast(s0);
            }
            __syncwarp();              // every lane holds rows r-1 .. r+R before rows r .. r+R-1 are overwritten
            #pragma unroll
            for (int k = 0; k < R; k++)
                if (member >> (r + k) & 1) own[(r + k) * CPB_DC_PITCH] = v[k];
            #pragma unroll
            for (int k = 0; k < 2; k++) { win[k][0] = win[R + k][0]; win[k][1] = win[R + k][1]; win[k][2] = win[R + k][2]; }
        }
        __syncwarp();
    }
    const bool my_clean = inB ? clean[1] : clean[0];
    if (mine && !my_clean)
        for (int r = 0; r < my.h; r++)
            if (member >> r & 1) Tb[(my.y0 + r) * W + my.x0 + col] = own[r * CPB_DC_PITCH];
    if (!fuse || !(clean[0] || clean[1])) return;
    // flow error of the clean labels from the tile (non-member cells and the halo hold 0, as global T would)
    double ey = 0.0, ex = 0.0;
    if (mine && my_clean) {
        for (int r = 0; r < my.h; r++) {
            if (!(member >> r & 1)) continue;
            const double* c = own + r * CPB_DC_PITCH;
            const int pix = (my.y0 + r) * W + my.x0 + col;
            cpb_flow_err_pixel(__dsub_rn(c[CPB_DC_PITCH], c[-CPB_DC_PITCH]), __dsub_rn(c[1], c[-1]), qc.dPy[pix], qc.dPx[pix], ey, ex);
        }
    }
    for (int q = 0; q < (has_b ? 2 : 1); q++)
        if (clean[q]) cpb_diffuse_publish_err(t, q ? B : A, q ? inB : inA, ey, ex, qc.threshold);   // warp-uniform
}

// a label belongs to the kernel instance whose row capacity is the smallest that holds it
template <int MAXH>
CPB_DEVICE bool cpb_diffuse_load_sub(const LabelTables& t, int b, int l, int lb, DiffSub& s) {
    if (l > lb) return false;
    s.l = l; s.k = (size_t)b * t.LC + l; s.coff = 0;
    if (!cpb_label_live(t, s.k)) return false;
    s.y0 = t.ymin[s.k]; s.x0 = t.xmin[s.k];
    s.h = t.ymax[s.k] - s.y0 + 1; s.w = t.xmax[s.k] - s.x0 + 1;
    if (!cpb_diffuse_is_small(s.h, s.w)) return false;
    return MAXH == CPB_DC_MIDH ? s.h <= CPB_DC_MIDH : s.h > CPB_DC_MIDH;
}

template <int MAXH>
CPB_KERNEL CPB_LAUNCH_BOUNDS(CPB_DW_WARPS * 32, (MAXH == CPB_DC_MIDH ? 8 : 6))
k_diffuse_warp(const int* CPB_RESTRICT lab, int H, int W, LabelTables t, double* CPB_RESTRICT T,
               int niter_override, const float* CPB_RESTRICT dP, double threshold, int2* todo, int* todo_count, int todo_cap) {
    CPB_SHARED double s_T[CPB_DW_WARPS][(MAXH + 3) * CPB_DC_PITCH];
    constexpr int R = 2;
    const int warp = threadIdx.x >> 5;
    const int b = blockIdx.y, N = H * W;
    const int lb = t.lbound[b];
    const int* L = lab + (size_t)b * N;
    double* Tb = T + (size_t)b * N;
    const int n_it = niter_override > 0 ? niter_override : t.niter[b];
    double* S = s_T[warp];
    const DiffQC qc{dP ? dP + ((size_t)b * 2 + 0) * N : nullptr, dP ? dP + ((size_t)b * 2 + 1) * N : nullptr,
                    t.alive ? t.alive + (size_t)b * t.LC : nullptr, threshold, H, todo, todo_count, todo_cap, b};
    // work item = pair of consecutive labels (2i+1, 2i+2); everything below is warp-uniform
    for (int wi = blockIdx.x * CPB_DW_WARPS + warp; 2 * wi + 1 <= lb; wi += gridDim.x * CPB_DW_WARPS) {
        DiffSub A, B;
        const bool okA = cpb_diffuse_load_sub<MAXH>(t, b, 2 * wi + 1, lb, A);
        const bool okB = cpb_diffuse_load_sub<MAXH>(t, b, 2 * wi + 2, lb, B);
        if (okA && okB && A.w + 1 + B.w <= CPB_DC_MAXW) {
            B.coff = A.w + 1;
            cpb_diffuse_job<R>(L, W, t, Tb, S, A, B, true, n_it, qc);
        } else {
            if (okA) cpb_diffuse_job<R>(L, W, t, Tb, S, A, A, false, n_it, qc);
            if (okB) cpb_diffuse_job<R>(L, W, t, Tb, S, B, B, false, n_it, qc);
        }
    }
}

// ---- dynamic job queue for the warp kernel --------------------------------------------------------------
// Labels differ in height and tiles in n_iter, so a static (block, warp) -> label map leaves half of the
// resident warps idle while the slowest warp of each block finishes.  k_diffuse_jobs numbers the label pairs of
// the whole batch (exclusive scan of ceil(lbound/2) over tiles); persistent warps of k_diffuse_warp_q then pull
// pair after pair from one atomic counter, so every resident warp stays busy until the queue is empty.
CPB_KERNEL k_diffuse_jobs(const int* CPB_RESTRICT lbound, int B, int* CPB_RESTRICT joboff, int* CPB_RESTRICT counters) {
    CPB_SHARED int s_scan[33];
    CPB_SHARED int s_base;
    if (threadIdx.x == 0) { s_base = 0; counters[0] = 0; counters[1] = 0; }
    __syncthreads();
    for (int b0 = 0; b0 < B; b0 += blockDim.x) {
        const int b = b0 + threadIdx.x;
        const int n = b < B ? (lbound[b] + 1) / 2 : 0;
        int tot;
        const int incl = cpb_block_scan_incl(n, s_scan, &tot);
        if (b < B) joboff[b] = s_base + incl - n;
        __syncthreads();
        if (threadIdx.x == 0) s_base += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) joboff[B] = s_base;
}

#ifndef CPB_DQ_MINBLOCKS
#define CPB_DQ_MINBLOCKS 8
#endif
template <int MAXH, int R>
CPB_KERNEL CPB_LAUNCH_BOUNDS(CPB_DW_WARPS * 32, (MAXH == CPB_DC_MIDH ? CPB_DQ_MINBLOCKS : 6))
k_diffuse_warp_q(const int* CPB_RESTRICT lab, int B, int H, int W, LabelTables t, double* CPB_RESTRICT T,
                 int niter_override, const int* CPB_RESTRICT joboff, int* CPB_RESTRICT counter,
                 const float* CPB_RESTRICT dP, double threshold, int2* todo, int* todo_count, int todo_cap) {
    CPB_SHARED double s_T[CPB_DW_WARPS][((MAXH + R - 1) / R * R + R + 1) * CPB_DC_PITCH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int N = H * W;
    const int total = joboff[B];
    double* S = s_T[warp];
    for (;;) {
        int j = 0;
        if (lane == 0) j = atomicAdd(counter, 1);
        j = __shfl_sync(CPB_FULL, j, 0);
        if (j >= total) break;
        int lo = 0, hi = B;                       // joboff[lo] <= j < joboff[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (joboff[mid] <= j) lo = mid; else hi = mid;
        }
        const int b = lo, wi = j - joboff[lo];
        const int lb = t.lbound[b];
        const int* L = lab + (size_t)b * N;
        double* Tb = T + (size_t)b * N;
        const int n_it = niter_override > 0 ? niter_override : t.niter[b];
        const DiffQC qc{dP ? dP + ((size_t)b * 2 + 0) * N : nullptr, dP ? dP + ((size_t)b * 2 + 1) * N : nullptr,
                        t.alive ? t.alive + (size_t)b * t.LC : nullptr, threshold, H, todo, todo_count, todo_cap, b};
        DiffSub A, Bs;
        const bool okA = cpb_diffuse_load_sub<MAXH>(t, b, 2 * wi + 1, lb, A);
        const bool okB = cpb_diffuse_load_sub<MAXH>(t, b, 2 * wi + 2, lb, Bs);
        if (okA && okB && A.w + 1 + Bs.w <= CPB_DC_MAXW) {
            Bs.coff = A.w + 1;
            cpb_diffuse_job<R>(L, W, t, Tb, S, A, Bs, true, n_it, qc);
        } else {
            if (okA) cpb_diffuse_job<R>(L, W, t, Tb, S, A, A, false, n_it, qc);
            if (okB) cpb_diffuse_job<R>(L, W, t, Tb, S, Bs, Bs, false, n_it, qc);
        }
    }
}

// Neighbour order of the reference: self, (-1,0), (1,0), (0,-1), (0,1), (-1,-1), (-1,1), (1,-1), (1,1)
CPB_KERNEL CPB_LAUNCH_BOUNDS(CPB_QC_THREADS, 8)
k_diffuse(const int* CPB_RESTRICT lab, int H, int W, LabelTables t, double* CPB_RESTRICT T,
          double* CPB_RESTRICT T2, int niter_override, int skip_small, LabelWork wk) {
    CPB_DYN_SMEM(double, s_buf);   // 2 * CPB_DIFF_SMEM_CELLS doubles + CPB_DIFF_SMEM_CELLS bytes
    const int LC = t.LC, N = H * W;
    int it_ = 0, b, l;
    while (cpb_next_label(wk, t.lbound, it_, b, l)) {
        const int* L = lab + (size_t)b * N;
        double* Tb = T + (size_t)b * N;
        double* T2b = T2 + (size_t)b * N;
        const int n_it = niter_override > 0 ? niter_override : t.niter[b];
        const size_t k = (size_t)b * LC + l;
        if (!cpb_label_live(t, k)) continue;
        const int y0 = t.ymin[k], x0 = t.xmin[k];
        const int h = t.ymax[k] - y0 + 1, w = t.xmax[k] - x0 + 1;
        if (skip_small && cpb_diffuse_is_small(h, w)) continue;
        const int cy = t.cy[k], cx = t.cx[k];
        const int hh = h + 2, ww = w + 2, cells = hh * ww;
        if (cells <= CPB_DIFF_SMEM_CELLS) {
            double* A = s_buf;
            double* Bf = s_buf + CPB_DIFF_SMEM_CELLS;
            unsigned char* mem = reinterpret_cast<unsigned char*>(s_buf + 2 * CPB_DIFF_SMEM_CELLS);
            __syncthreads();
            for (int i = threadIdx.x; i < cells; i += blockDim.x) {
                const int ry = i / ww - 1, rx = i % ww - 1;
                const bool in = ry >= 0 && ry < h && rx >= 0 && rx < w && L[(y0 + ry) * W + x0 + rx] == l;
                mem[i] = in ? 1 : 0;
                A[i] = 0.0; Bf[i] = 0.0;
            }
            const int ci = (cy - y0 + 1) * ww + (cx - x0 + 1);
            __syncthreads();
            double* cur = A; double* nxt = Bf;
            for (int it = 0; it < n_it; it++) {
                for (int i = threadIdx.x; i < cells; i += blockDim.x) {
                    if (!mem[i]) continue;
                    // T[centre] += 1 happens before the averaging step: fold it into the reads
                    #define CPB_TV(j) (cur[(j)] + ((j) == ci ? 1.0 : 0.0))
                    double s = CPB_TV(i);
                    s = __dadd_rn(s, CPB_TV(i - ww));
                    s = __dadd_rn(s, CPB_TV(i + ww));
                    s = __dadd_rn(s, CPB_TV(i - 1));
                    s = __dadd_rn(s, CPB_TV(i + 1));
                    s = __dadd_rn(s, CPB_TV(i - ww - 1));
                    s = __dadd_rn(s, CPB_TV(i - ww + 1));
                    s = __dadd_rn(s, CPB_TV(i + ww - 1));
                    s = __dadd_rn(s, CPB_TV(i + ww + 1));
                    #undef CPB_TV
                    nxt[i] = cpb_div9(s);
                }
                __syncthreads();
                double* tmp = cur; cur = nxt; nxt = tmp;
            }
            for (int i = threadIdx.x; i < cells; i += blockDim.x)
                if (mem[i]) Tb[(y0 + i / ww - 1) * W + x0 + i % ww - 1] = cur[i];
        } else {
            // large instance: iterate in global memory (two full-tile float64 planes)
            __syncthreads();
            for (int i = threadIdx.x; i < h * w; i += blockDim.x) {
                const int p = (y0 + i / w) * W + x0 + i % w;
                if (L[p] == l) { Tb[p] = 0.0; T2b[p] = 0.0; }
            }
            __syncthreads();
            double* cur = Tb; double* nxt = T2b;
            const int cp = cy * W + cx;
            for (int it = 0; it < n_it; it++) {
                for (int i = threadIdx.x; i < h * w; i += blockDim.x) {
                    const int y = y0 + i / w, x = x0 + i % w;
                    const int p = y * W + x;
                    if (L[p] != l) continue;
                    double s = 0.0;
                    const int oy[9] = {0, -1, 1, 0, 0, -1, -1, 1, 1};
                    const int ox[9] = {0, 0, 0, -1, 1, -1, 1, -1, 1};
                    for (int q = 0; q < 9; q++) {
                        const int yy = y + oy[q], xx = x + ox[q];
                        double v = 0.0;
                        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                            const int pq = yy * W + xx;
                            if (L[pq] == l) v = cur[pq] + (pq == cp ? 1.0 : 0.0);
                        }
                        s = q == 0 ? v : __dadd_rn(s, v);
                    }
                    nxt[p] = cpb_div9(s);
                }
                __syncthreads();
                double* tmp = cur; cur = nxt; nxt = tmp;
            }
            if (cur != Tb) {
                for (int i = threadIdx.x; i < h * w; i += blockDim.x) {
                    const int p = (y0 + i / w) * W + x0 + i % w;
                    if (L[p] == l) Tb[p] = cur[p];
                }
            }
            __syncthreads();
        }
    }
}

// T of a neighbouring pixel as the reference's padded array holds it: 0 off-tile and on
// background, the (possibly foreign) instance's T otherwise.
CPB_DEVICE double cpb_T_at(const double* CPB_RESTRICT Tb, const int* CPB_RESTRICT L, const int* CPB_RESTRICT alive,
                           int H, int W, int y, int x) {
    // branch-free: the label and T loads are issued unconditionally at a clamped address so that the four
    // neighbour fetches of a pixel overlap; T of a non-label pixel is never written and is discarded here
    const bool inb = y >= 0 && y < H && x >= 0 && x < W;
    const int p = min(max(y, 0), H - 1) * W + min(max(x, 0), W - 1);
    const int l = L[p];
    const double v = Tb[p];
    const bool live = l > 0 && (alive == nullptr || alive[l] != 0);
    return (inb && live) ? v : 0.0;
}

CPB_KERNEL CPB_LAUNCH_BOUNDS(CPB_QC_THREADS, 8)
k_flow_err(const int* CPB_RESTRICT lab, const float* CPB_RESTRICT dP, int H, int W, LabelTables t,
           const double* CPB_RESTRICT T, double threshold, double* CPB_RESTRICT mu_out, int skip_small, LabelWork wk) {
    CPB_SHARED double s_ey[CPB_QC_THREADS / 32], s_ex[CPB_QC_THREADS / 32];
    const int LC = t.LC, N = H * W;
    int it_ = 0, b, l;
    while (cpb_next_label(wk, t.lbound, it_, b, l)) {
        const int* L = lab + (size_t)b * N;
        const double* Tb = T + (size_t)b * N;
        const int* alive = t.alive ? t.alive + (size_t)b * LC : nullptr;
        const float* dPy = dP ? dP + ((size_t)b * 2 + 0) * N : nullptr;
        const float* dPx = dP ? dP + ((size_t)b * 2 + 1) * N : nullptr;
        const size_t k = (size_t)b * LC + l;
        const int c = t.cnt[k];
        if (!cpb_label_live(t, k)) continue;
        const int y0 = t.ymin[k], x0 = t.xmin[k];
        const int h = t.ymax[k] - y0 + 1, w = t.xmax[k] - x0 + 1;
        if (skip_small && w <= 32) continue;       // handled by k_flow_err_warp
        if (t.done[k]) continue;                    // error already taken from the diffusion tile (block-uniform)
        double ey = 0.0, ex = 0.0;
        for (int i = threadIdx.x; i < h * w; i += blockDim.x) {
            const int y = y0 + i / w, x = x0 + i % w;
            const int p = y * W + x;
            if (L[p] != l) continue;
            const double dy = __dsub_rn(cpb_T_at(Tb, L, alive, H, W, y + 1, x), cpb_T_at(Tb, L, alive, H, W, y - 1, x));
            const double dx = __dsub_rn(cpb_T_at(Tb, L, alive, H, W, y, x + 1), cpb_T_at(Tb, L, alive, H, W, y, x - 1));
            const double nrm = __dadd_rn(1e-60, __dsqrt_rn(__dadd_rn(__dmul_rn(dy, dy), __dmul_rn(dx, dx))));
            const double my = __ddiv_rn(dy, nrm), mx = __ddiv_rn(dx, nrm);
            if (mu_out) {
                mu_out[((size_t)b * 2 + 0) * N + p] = my;
                mu_out[((size_t)b * 2 + 1) * N + p] = mx;
            }
            if (dP) {
                const double ry = __dsub_rn(my, (double)__fdiv_rn(dPy[p], 5.0f));
                const double rx = __dsub_rn(mx, (double)__fdiv_rn(dPx[p], 5.0f));
                ey += __dmul_rn(ry, ry);
                ex += __dmul_rn(rx, rx);
            }
        }
        if (!dP) continue;
        for (int s = 16; s; s >>= 1) {
            ey += __shfl_xor_sync(CPB_FULL, ey, s);
            ex += __shfl_xor_sync(CPB_FULL, ex, s);
        }
        __syncthreads();
        if ((threadIdx.x & 31) == 0) { s_ey[threadIdx.x >> 5] = ey; s_ex[threadIdx.x >> 5] = ex; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double sy = 0.0, sx = 0.0;
            for (int q = 0; q < CPB_QC_THREADS / 32; q++) { sy += s_ey[q]; sx += s_ex[q]; }
            const double e = sy / (double)c + sx / (double)c;
            t.err[k] = e;
            t.flag[k] = e > threshold ? 1 : 0;
        }
    }
}
