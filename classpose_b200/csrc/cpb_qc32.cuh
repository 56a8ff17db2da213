// Row (4), decision-exact: a float32 screen in front of the float64 diffusion of cpb_qc.cuh.
//
// remove_bad_flow_masks only needs the DECISION err > flow_threshold per label (north_star: instance F1, class
// labels and cell counts -- not the bits of T).  The screen runs the same Jacobi iteration in float32 with the
// label's T held in REGISTERS, takes the flow error from it and carries a rigorous bound on
// |err32 - err64| (derivation below).  A label is decided by the screen only when err32 +- bound clears the
// threshold; every other label (and every label the screen cannot hold or whose gradient sees a foreign label)
// is handed to the float64 warp kernel, so the removal set is identical to the float64 path by construction.
//
// Layout of a job (one warp): a strip of 64 columns, lane j owns columns 2j (.x) and 2j+1 (.y) as one packed
// f32x2 register per row; up to CPB_Q32_MAXSUB labels of one tile sit side by side, separated by one empty
// column (the last column of the strip is always empty, so the wrap-around shuffles read zeros).  Every label is
// shifted vertically so that its diffusion centre lies on register row RC = NR/2: the source injection is ONE
// packed add per iteration instead of a dynamically indexed row.  Per row and iteration: vertical 3-sums in
// registers (2 FADD2), the two values a lane lacks by shuffle (2 SHFL), 3 FADD, and one FMUL2 by M = member ? 1/9 : 0
// (non-members stay exactly 0).  No shared memory in the loop.
//
// Error bound.  All T are >= 0 and the iteration only adds non-negative numbers and multiplies by a positive
// constant, so rounding errors stay RELATIVE: with u = 2^-24 and at most 11 roundings between a value and its
// successor (1 source add, 4 adds on the longest path of the separable 9-sum, the constant fl(1/9) and the
// multiply -- 11 is the budget of the plain 9-term order and covers both),
//     T32 = T_exact (1 + e),  |e| <= (1 + u)^(11 k) - 1 =: eta  after k iterations,
// and likewise for the float64 reference with u = 2^-53 (absorbed by the 2 % head-room on eta).  Results in the
// float32 subnormal range carry an absolute error instead, bounded by alpha = 1e-40 in total.  Hence per pixel
//     |dy32 - dy64| <= eta (T_dn + T_up) + u |dy32| + 2 alpha   (same for dx),   e_g := e_y + e_x >= |g32 - g64|,
//     |mu32 - mu64| <= e_g / (|g32| - e_g)        (Dunkl-Williams; used when |g32| > 4 e_g),
//     otherwise mu32 := 0 and |mu32 - mu64| <= 1  (mu64 is a unit vector or 0),
//     | |mu32 - a|^2 - |mu64 - a|^2 | <= 2 |mu32 - a| d + d^2   with d the bound above, a = dP / 5,
// plus generous slack for the float32 evaluation of these expressions and of the sums.
#pragma once
#include "cpb_common.cuh"
#include "cpb_flow.cuh"
#include "cpb_qc.cuh"

#define CPB_Q32_MAXSUB 8
#define CPB_Q32_COLS 64            // columns of a job strip (two per lane); column 63 stays empty
#define CPB_Q32_NCLS 6             // register-row classes NR = 9, 13, 17, 21, 25, 31 (centre row RC = NR / 2 = 4 .. 15)
#define CPB_Q32_MAXNR 31
#define CPB_Q32_CLS64 6            // info class of labels left to the float64 warp kernel
#define CPB_Q32_CLSBIG 7           // info class of labels left to the block kernels

#define CPB_QI_CLEAN 8             // info bit: no pixel of another live label in the bbox grown by one
#define CPB_QI_T32 16              // info bit: the screen wrote this label's float32 T to the global plane
#define CPB_QI_QUEUED64 32         // info bit: the label is already on the float64 list

// counters: [0] jobs appended, [1] jobs pulled, [2] float64 list appended, [3] float64 list pulled,
//           [4] labels decided by the screen, [5] labels the screen left undecided (statistics),
//           [6] contact list appended, [7] contact list pulled,
//           [8..13] jobs per class, [14..19] scatter cursors of k_q32_sort
#ifndef CPB_QCTR_INTS
#define CPB_QCTR_INTS 24
#endif

struct Q32 {
    int* info;        // [B*LC]  class (bits 0..2) | CPB_QI_CLEAN | bbox width << 8
    int* ent;         // [B*LC]  per tile: the screen's labels, grouped by class
    int4* jobs;       // [B*LC]  (tile, first entry, nsub, class) in the order k_qc_pack emitted them
    int4* sorted;     // [B*LC]  the same jobs, tallest class first (k_q32_sort)
    int2* l64;        // [B*LC]  (tile, label) for the float64 warp kernel
    int2* lc;         // [B*LC]  (tile, label): screened labels in contact with another label (k_flow_err32)
    int* cls_cnt;     // [B*6]   screened labels per tile and class
    float* T32;       // [B*N]   float32 T of the screened labels in contact (their neighbours read it)
    int* ctr;         // [CPB_QCTR_INTS]
};

// append (tile, label) to the float64 list once
CPB_DEVICE void cpb_q32_queue64(const Q32& q, int b, int l, size_t k) {
    if ((atomicOr(&q.info[k], CPB_QI_QUEUED64) & CPB_QI_QUEUED64) == 0) q.l64[atomicAdd(&q.ctr[2], 1)] = make_int2(b, l);
}

CPB_DEVICE int cpb_q32_class(int cr, int h) {
    const int rho = max(cr, h - 1 - cr);
    return rho <= 4 ? 0 : rho <= 6 ? 1 : rho <= 8 ? 2 : rho <= 10 ? 3 : rho <= 12 ? 4 : rho <= 15 ? 5 : CPB_Q32_CLS64;
}

// ---- k_qc_scan32: one warp per label, grid (slices, B) ------------------------------------------------------
//  diffusion centre, contact with other live labels, class; labels beyond the warp kernels go to `big_list`, small
//  labels the register classes cannot hold go to q.l64; per-tile n_iter (t.niter, zeroed) and class counts
//  (q.cls_cnt, zeroed).
CPB_KERNEL CPB_LAUNCH_BOUNDS(256, 4)
k_qc_scan32(const int* CPB_RESTRICT lab, int H, int W, LabelTables t, Q32 q, int2* CPB_RESTRICT big_list,
            int* CPB_RESTRICT big_count, int screen) {
    CPB_SHARED unsigned s_bits[8][36];            // per warp: member bits of the grown bbox columns
    const int b = blockIdx.y, LC = t.LC, N = H * W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int lb = t.lbound[b];
    const int* L = lab + (size_t)b * N;
    const int* alive = t.alive ? t.alive + (size_t)b * LC : nullptr;
    int ext = 0;
    for (int l = 1 + blockIdx.x * nw + warp; l <= lb; l += gridDim.x * nw) {
        const size_t k = (size_t)b * LC + l;
        int info = 0;
        if (cpb_label_live(t, k)) {                              // warp-uniform
            const int y0 = t.ymin[k], x0 = t.xmin[k];
            const int h = t.ymax[k] - y0 + 1, w = t.xmax[k] - x0 + 1;
            ext = max(ext, 2 * (h + w + 2));
            if (!cpb_diffuse_is_small(h, w)) {
                if (lane == 0) big_list[atomicAdd(big_count, 1)] = make_int2(b, l);
                info = CPB_Q32_CLSBIG;
            } else {
                // means of (bbox-relative coordinate + 1), as the reference computes them (same ops as k_centres)
                const int c = t.cnt[k];
                const double ymed = __ddiv_rn(__ll2double_rn((long long)t.sumy[k] - (long long)c * y0 + c), __int2double_rn(c));
                const double xmed = __ddiv_rn(__ll2double_rn((long long)t.sumx[k] - (long long)c * x0 + c), __int2double_rn(c));
                // scan of the bbox grown by one pixel (columns x0-1 .. x0+w: two passes when w + 2 > 32; rows y0-1 .. y0+h):
                // integer work only -- member bits of this lane's column (bit r = row y0 + r) and contact with other labels
                unsigned member = 0;               // bbox column `lane` (first pass: grown column lane + 1)
                bool foreign = false;
                for (int c0 = 0; c0 < w + 2; c0 += 32) {
                    const int gc = c0 + lane;                    // grown column index; bbox column gc - 1
                    const int x = x0 - 1 + gc;
                    if (gc < w + 2 && x >= 0 && x < W) {
                        unsigned mb = 0;
                        for (int rb = -1; rb <= h; rb += 8) {      // eight loads in flight, then the tests
                            int vv[8];
                            #pragma unroll
                            for (int q8 = 0; q8 < 8; q8++) {
                                const int y = y0 + rb + q8;
                                vv[q8] = (rb + q8 <= h && y >= 0 && y < H) ? L[y * W + x] : 0;
                            }
                            #pragma unroll
                            for (int q8 = 0; q8 < 8; q8++) {
                                const int ry = rb + q8;
                                if (vv[q8] == l) mb |= 1u << (ry & 31);          // members only occur for 0 <= ry < h <= 30
                                else foreign |= cpb_foreign_live(vv[q8], l, alive);
                            }
                        }
                        // hand the bits to the lane that owns bbox column gc - 1
                        s_bits[warp][gc] = mb;
                    }
                }
                __syncwarp();
                member = (lane < w) ? s_bits[warp][lane + 1] : 0u;
                // centre = member pixel nearest to the mean, first in raster order on ties.  Every pixel outside the
                // 4 x 4 window around the mean is at least 2 away along one axis (squared distance >= 4, exactly, in
                // float64 too), so a member of the window with squared distance < 4 decides -- one candidate per
                // lane; otherwise (a label whose mean falls outside itself) every member is evaluated.
                double bd = 1e300; int bi = CPB_IMAX;
                {
                    const int wy = (int)floor(ymed) - 2 + (lane >> 2), wx = (int)floor(xmed) - 2 + (lane & 3);   // bbox-relative
                    if (lane < 16 && wy >= 0 && wy < h && wx >= 0 && wx < w) {
                        const unsigned colbits = s_bits[warp][wx + 1];
                        if (colbits >> wy & 1) {
                            const double dx = __dsub_rn(__int2double_rn(wx + 1), xmed), dy = __dsub_rn(__int2double_rn(wy + 1), ymed);
                            bd = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)); bi = wy * w + wx;
                        }
                    }
                    double wbd = bd;
                    for (int sft = 16; sft; sft >>= 1) wbd = fmin(wbd, __shfl_xor_sync(CPB_FULL, wbd, sft));
                    if (!(wbd < 4.0)) {                          // warp-uniform: no member that close, look at all of them
                        bd = 1e300; bi = CPB_IMAX;
                        if (lane < w) {
                            const double dx = __dsub_rn(__int2double_rn(lane + 1), xmed);
                            const double dx2 = __dmul_rn(dx, dx);
                            for (int r = 0; r < h; r++) {
                                if (!(member >> r & 1)) continue;
                                const double dy = __dsub_rn(__int2double_rn(r + 1), ymed);
                                const double d = __dadd_rn(dx2, __dmul_rn(dy, dy));
                                const int idx = r * w + lane;
                                if (cpb_minkey_less(d, idx, bd, bi)) { bd = d; bi = idx; }
                            }
                        }
                    }
                }
                __syncwarp();
                for (int sft = 16; sft; sft >>= 1) {
                    const double od = __shfl_xor_sync(CPB_FULL, bd, sft);
                    const int oi = __shfl_xor_sync(CPB_FULL, bi, sft);
                    if (cpb_minkey_less(od, oi, bd, bi)) { bd = od; bi = oi; }
                }
                const bool clean = !__any_sync(CPB_FULL, foreign);
                const int cr = bi / w, cc = bi - cr * w;
                int cls = cpb_q32_class(cr, h);
                if (!screen) cls = CPB_Q32_CLS64;
                info = cls | (clean ? CPB_QI_CLEAN : 0) | (w << 8);
                if (lane == 0) {
                    t.cy[k] = y0 + cr; t.cx[k] = x0 + cc;
                    if (cls == CPB_Q32_CLS64) { info |= CPB_QI_QUEUED64; q.l64[atomicAdd(&q.ctr[2], 1)] = make_int2(b, l); }
                    else atomicAdd(&q.cls_cnt[b * CPB_Q32_NCLS + cls], 1);
                }
            }
        }
        if (lane == 0) q.info[k] = info;
    }
    for (int s = 16; s; s >>= 1) ext = max(ext, __shfl_xor_sync(CPB_FULL, ext, s));
    if (lane == 0 && ext > 0) atomicMax(&t.niter[b], ext);
}

// ---- k_qc_pack: one block of CPB_Q32_NCLS warps per tile -----------------------------------------------------
//  the screen's labels grouped by class into q.ent; warp c packs class c greedily into jobs of up to 8 labels / 63 columns
CPB_KERNEL CPB_LAUNCH_BOUNDS(32 * CPB_Q32_NCLS, 8)
k_qc_pack(LabelTables t, Q32 q) {
    CPB_SHARED int s_base[CPB_Q32_NCLS], s_pos[CPB_Q32_NCLS];
    const int b = blockIdx.x, LC = t.LC;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lb = t.lbound[b];
    const int* cnt = q.cls_cnt + b * CPB_Q32_NCLS;
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int c = 0; c < CPB_Q32_NCLS; c++) { s_base[c] = acc; s_pos[c] = 0; acc += cnt[c]; }
    }
    __syncthreads();
    int* ent = q.ent + (size_t)b * LC;
    for (int l = 1 + threadIdx.x; l <= lb; l += blockDim.x) {
        const int info = q.info[(size_t)b * LC + l];
        const int cls = info & 7;
        if (cls < CPB_Q32_NCLS && (info >> 8) != 0) ent[s_base[cls] + atomicAdd(&s_pos[cls], 1)] = l;
    }
    __syncthreads();
    if (warp < CPB_Q32_NCLS) {
        const int cls = warp, n = cnt[cls];
        const int* e = ent + s_base[cls];
        int coff = 0, nsub = 0, first = 0;
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            const int wi = i < n ? (q.info[(size_t)b * LC + e[i]] >> 8) & 0xff : 0;
            const int m = min(32, n - i0);
            for (int j = 0; j < m; j++) {
                const int wj = __shfl_sync(CPB_FULL, wi, j);
                if (nsub == CPB_Q32_MAXSUB || coff + wj > CPB_Q32_COLS - 1) {
                    if (lane == 0) { q.jobs[atomicAdd(&q.ctr[0], 1)] = make_int4(b, (int)(e - q.ent) + first, nsub, cls); atomicAdd(&q.ctr[8 + cls], 1); }
                    first = i0 + j; nsub = 0; coff = 0;
                }
                nsub++; coff += wj + 1;
            }
        }
        if (nsub > 0 && lane == 0) { q.jobs[atomicAdd(&q.ctr[0], 1)] = make_int4(b, (int)(e - q.ent) + first, nsub, cls); atomicAdd(&q.ctr[8 + cls], 1); }
    }
}

// k_q32_sort: jobs grouped by class, tallest first.  Warps of k_diffuse32 then run the same unrolled loop most of the
// time (with the classes interleaved the instruction cache was the top stall reason) and the longest jobs start first.
CPB_KERNEL k_q32_sort(Q32 q) {
    const int n = q.ctr[0];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int4 jb = q.jobs[i];
        int base = 0;
        for (int c = CPB_Q32_NCLS - 1; c > jb.w; c--) base += q.ctr[8 + c];
        q.sorted[base + atomicAdd(&q.ctr[14 + jb.w], 1)] = jb;
    }
}

// ---- packed float32 pairs ------------------------------------------------------------------------------
#ifdef CPB_SIM
struct pf2 { float x, y; };
CPB_DEVICE pf2 pf2_make(float x, float y) { pf2 r; r.x = x; r.y = y; return r; }
CPB_DEVICE void pf2_get(pf2 a, float& x, float& y) { x = a.x; y = a.y; }
CPB_DEVICE pf2 pf2_add(pf2 a, pf2 b) { return pf2_make(a.x + b.x, a.y + b.y); }
CPB_DEVICE pf2 pf2_mul(pf2 a, pf2 b) { return pf2_make(a.x * b.x, a.y * b.y); }
CPB_DEVICE pf2 pf2_sub(pf2 a, pf2 b) { return pf2_make(a.x - b.x, a.y - b.y); }
CPB_DEVICE pf2 pf2_fma(pf2 a, pf2 b, pf2 c) { return pf2_make(__fmaf_rn(a.x, b.x, c.x), __fmaf_rn(a.y, b.y, c.y)); }
#else
typedef u64 pf2;
CPB_DEVICE pf2 pf2_make(float x, float y) { return cpb_pk(x, y); }
CPB_DEVICE void pf2_get(pf2 a, float& x, float& y) { cpb_upk(a, x, y); }
CPB_DEVICE pf2 pf2_add(pf2 a, pf2 b) { return cpb_add2(a, b); }
CPB_DEVICE pf2 pf2_mul(pf2 a, pf2 b) { return cpb_mul2(a, b); }
CPB_DEVICE pf2 pf2_sub(pf2 a, pf2 b) { return cpb_sub2(a, b); }
CPB_DEVICE pf2 pf2_fma(pf2 a, pf2 b, pf2 c) { return cpb_fma2(a, b, c); }
#endif

// approximate reciprocal / square roots (MUFU, <= 2 ulp): their error is part of the bound's slack below
#ifdef CPB_SIM
CPB_DEVICE float cpb_rsqrt_approx(float x) { return 1.0f / std::sqrt(x); }
CPB_DEVICE float cpb_sqrt_approx(float x) { return std::sqrt(x); }
CPB_DEVICE float cpb_div_approx(float a, float b) { return a / b; }
#else
CPB_DEVICE float cpb_rsqrt_approx(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
CPB_DEVICE float cpb_sqrt_approx(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
CPB_DEVICE float cpb_div_approx(float a, float b) { return __fdividef(a, b); }
#endif

// one pixel of the float32 flow error and of its bound (see the header comment).  Slack: mu32 carries <= 1e-6 of
// absolute error from the approximate rsqrt and the multiplies (added to d), the squared distance c and the
// rounded dP / 5 <= 3e-6 (1 + c) relative-ish (last term), the quotient eg / (gn - eg) 2 ulp (inside the 1.01).
CPB_DEVICE void cpb_q32_pixel(float up, float dn, float lf, float rt, float dpy, float dpx, float eta, float& sc, float& sb) {
    const float U = 5.9604645e-8f, ALPHA4 = 4e-40f;
    const float dy = __fsub_rn(dn, up), dx = __fsub_rn(rt, lf);
    const float eg = __fadd_rn(__fadd_rn(__fmul_rn(eta, __fadd_rn(__fadd_rn(dn, up), __fadd_rn(rt, lf))),
                                         __fmul_rn(U, __fadd_rn(fabsf(dy), fabsf(dx)))), ALPHA4);
    const float g2 = __fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dx, dx));
    const float inv = cpb_rsqrt_approx(fmaxf(g2, 1e-37f));
    const float gn = __fmul_rn(g2, inv);
    float muy = 0.f, mux = 0.f, d = 1.000001f;
    if (gn > 4.f * eg && gn > 1e-18f) {
        muy = __fmul_rn(dy, inv); mux = __fmul_rn(dx, inv);
        d = __fadd_rn(__fmul_rn(1.01f, cpb_div_approx(eg, __fsub_rn(gn, eg))), 2e-6f);
    }
    const float ry = __fsub_rn(muy, __fdiv_rn(dpy, 5.0f)), rx = __fsub_rn(mux, __fdiv_rn(dpx, 5.0f));
    const float c = __fadd_rn(__fmul_rn(ry, ry), __fmul_rn(rx, rx));
    const float rn = __fmul_rn(1.00001f, cpb_sqrt_approx(c));
    sc = __fadd_rn(sc, c);
    sb = __fadd_rn(sb, __fadd_rn(__fadd_rn(__fmul_rn(__fadd_rn(rn, rn), d), __fmul_rn(d, d)), __fmul_rn(3e-6f, __fadd_rn(1.f, c))));
}

// decision of one screened label from the float32 error sum and the bound sum (see the header comment)
CPB_DEVICE bool cpb_q32_decide(const LabelTables& t, const Q32& q, int b, int l, size_t k, float cs, float bd, double threshold,
                               int pack_err) {
    const double cnt = (double)t.cnt[k];
    const double err = (double)cs / cnt;
    const double bound = 1.01 * (double)bd / cnt + 1e-5 * err + 1e-7;
    // pack_err (tests): sign bit set, float32 error in the high word, its bound (rounded) in the low word of t.err
    t.err[k] = pack_err ? __longlong_as_double((long long)((1ull << 63) | ((u64)__float_as_uint((float)err) << 32) | (u64)__float_as_uint((float)bound))) : err;
    if (err - bound > threshold) { t.flag[k] = 1; t.done[k] = 1; atomicAdd(&q.ctr[4], 1); return true; }
    if (err + bound < threshold) { t.flag[k] = 0; t.done[k] = 1; atomicAdd(&q.ctr[4], 1); return true; }
    cpb_q32_queue64(q, b, l, k); atomicAdd(&q.ctr[5], 1);
    return false;
}

// Per-lane description of the two strip columns (2 * lane, 2 * lane + 1) of a job.
struct Q32Cols { int l[2], x[2], yb[2]; float inj[2]; };

// The iteration proper: M (member ? 1/9 : 0) comes from the warp's shared-memory tile, T lives in registers for all
// n_it iterations and goes back to the tile at the end as (member ? T : -0.0f) -- the sign bit carries membership.
// Only this part is instantiated per row class; the set-up and the error pass are rolled loops shared by all classes
// (fully unrolled they are ~5000 instructions per class, which the instruction cache does not forgive).
template <int NR, bool MREG>       // MREG: M in registers (NR <= 21); otherwise M is re-read from the tile every row
CPB_DEVICE void cpb_q32_iterate(float2* S, float inj0, float inj1, int n_it) {
    constexpr int RC = NR / 2;
    constexpr int NM = MREG ? NR : 1;
    const int lane = threadIdx.x & 31;
    pf2 T[NR], M[NM];
    #pragma unroll
    for (int r = 0; r < NR; r++) {
        if (MREG) {
            const float2 m = S[(r + 1) * 32 + lane];
            M[r] = pf2_make(m.x, m.y);
        }
        T[r] = pf2_make(0.f, 0.f);
    }
    const pf2 J = pf2_make(inj0, inj1);
    const int lm = (lane + 31) & 31, lp = (lane + 1) & 31;
    for (int it = 0; it < n_it; it++) {
        T[RC] = pf2_add(T[RC], J);                                  // T[centre] += 1 before averaging
        pf2 vcur = pf2_add(T[0], T[1]);
        #pragma unroll
        for (int r = 0; r < NR; r++) {
            pf2 vnext = vcur;
            if (r + 2 < NR) vnext = pf2_add(pf2_add(T[r], T[r + 1]), T[r + 2]);      // old rows r .. r+2
            else if (r + 1 < NR) vnext = pf2_add(T[r], T[r + 1]);
            float vx, vy;
            pf2_get(vcur, vx, vy);
            const float ly = __shfl_sync(CPB_FULL, vy, lm);          // column 2j-1
            const float rx = __shfl_sync(CPB_FULL, vx, lp);          // column 2j+2
            const float s = __fadd_rn(vx, vy);
            pf2 m;
            if (MREG) m = M[r];
            else { const float2 mm = S[(r + 1) * 32 + lane]; m = pf2_make(mm.x, mm.y); }
            T[r] = pf2_mul(pf2_add(pf2_make(s, s), pf2_make(ly, rx)), m);    // ({s, s} is a broadcast operand)
            vcur = vnext;
        }
    }
    #pragma unroll
    for (int r = 0; r < NR; r++) {
        float tx, ty, mx, my;
        pf2_get(T[r], tx, ty);
        if (MREG) pf2_get(M[r], mx, my);
        else { const float2 mm = S[(r + 1) * 32 + lane]; mx = mm.x; my = mm.y; }
        S[(r + 1) * 32 + lane] = make_float2(mx != 0.f ? tx : -0.0f, my != 0.f ? ty : -0.0f);
    }
}

CPB_DEVICE bool cpb_q32_member(float v) { return (__float_as_uint(v) >> 31) == 0u; }

// One job: set-up (rolled), iteration (per class), error pass (rolled).  S: the warp's tile of (31 + 2) x 32 float2.
CPB_DEVICE void cpb_q32_run(const int* CPB_RESTRICT lab, const float* CPB_RESTRICT dP, int H, int W, const LabelTables& t,
                            const Q32& q, int b, int first, int nsub, int cls, double threshold, float2* S, int pack_err) {
    const int NR = cls < 5 ? 9 + 4 * cls : CPB_Q32_MAXNR, RC = NR / 2;
    const int lane = threadIdx.x & 31;
    const int N = H * W, LC = t.LC;
    const int* L = lab + (size_t)b * N;
    const float* dPy = dP + ((size_t)b * 2 + 0) * N;
    const float* dPx = dP + ((size_t)b * 2 + 1) * N;
    const int n_it = t.niter[b];
    // sub i lives on lane i: label, width, first column of the strip
    int e_l = 0, e_w = 0, e_clean = 0;
    if (lane < nsub) {
        e_l = q.ent[first + lane];
        const int info = q.info[(size_t)b * LC + e_l];
        e_w = (info >> 8) & 0xff; e_clean = (info & CPB_QI_CLEAN) ? 1 : 0;
    }
    int e_off = lane < nsub ? e_w + 1 : 0;
    for (int d = 1; d < CPB_Q32_MAXSUB; d <<= 1) {
        const int v = __shfl_up_sync(CPB_FULL, e_off, d);
        if (lane >= d) e_off += v;
    }
    e_off -= lane < nsub ? e_w + 1 : 0;
    // this lane's two columns
    Q32Cols c;
    int rlo[2] = {0, 0}, rhi[2] = {-1, -1};
    bool col_clean[2] = {true, true};
    c.l[0] = c.l[1] = 0; c.x[0] = c.x[1] = 0; c.yb[0] = c.yb[1] = 0; c.inj[0] = c.inj[1] = 0.f;
    for (int s = 0; s < nsub; s++) {
        const int sl = __shfl_sync(CPB_FULL, e_l, s), sw = __shfl_sync(CPB_FULL, e_w, s), so = __shfl_sync(CPB_FULL, e_off, s);
        const int sc_ = __shfl_sync(CPB_FULL, e_clean, s);
        const size_t k = (size_t)b * LC + sl;
        const int y0 = t.ymin[k], x0 = t.xmin[k], h = t.ymax[k] - y0 + 1, cy = t.cy[k], cx = t.cx[k];
        #pragma unroll
        for (int hf = 0; hf < 2; hf++) {
            const int col = 2 * lane + hf;
            if (col >= so && col < so + sw) {
                c.l[hf] = sl; c.x[hf] = x0 + col - so; c.yb[hf] = cy - RC;
                rlo[hf] = y0 - (cy - RC); rhi[hf] = y0 + h - 1 - (cy - RC);
                c.inj[hf] = (x0 + col - so == cx) ? 1.f : 0.f;
                col_clean[hf] = sc_ != 0;
            }
        }
    }
    const float ninth = 1.f / 9.f;
    __syncwarp();
    // membership: unconditional loads at clamped addresses so that all of a lane's loads are in flight together;
    // the flows the error pass will read (long evicted from L2 by the batch) are prefetched on the way
    #pragma unroll 7
    for (int r = 0; r < NR; r++) {
        float m[2];
        #pragma unroll
        for (int hf = 0; hf < 2; hf++) {
            const bool in = c.l[hf] != 0 && r >= rlo[hf] && r <= rhi[hf];
            const int pix = in ? (c.yb[hf] + r) * W + c.x[hf] : 0;
            const bool mem = L[pix] == c.l[hf] && in;
            m[hf] = mem ? ninth : 0.f;
#ifndef CPB_SIM
            if (mem && col_clean[hf]) {
                asm volatile("prefetch.global.L2 [%0];" :: "l"(dPy + pix));
                asm volatile("prefetch.global.L2 [%0];" :: "l"(dPx + pix));
            }
#endif
        }
        S[(r + 1) * 32 + lane] = make_float2(m[0], m[1]);
    }
    S[lane] = make_float2(-0.0f, -0.0f);
    S[(NR + 1) * 32 + lane] = make_float2(-0.0f, -0.0f);
    switch (cls) {              // each lane reads back only what it wrote: no warp barrier needed here
        case 0: cpb_q32_iterate<9, true>(S, c.inj[0], c.inj[1], n_it); break;
        case 1: cpb_q32_iterate<13, true>(S, c.inj[0], c.inj[1], n_it); break;
        case 2: cpb_q32_iterate<17, true>(S, c.inj[0], c.inj[1], n_it); break;
        case 3: cpb_q32_iterate<21, true>(S, c.inj[0], c.inj[1], n_it); break;
        case 4: cpb_q32_iterate<25, false>(S, c.inj[0], c.inj[1], n_it); break;
        default: cpb_q32_iterate<31, false>(S, c.inj[0], c.inj[1], n_it); break;
    }
    __syncwarp();
    // flow error and its bound from the tile
    // eta >= (1 + u)^(11 k) - 1:  e^x - 1 <= x + x^2 for 0 <= x <= 1, x = 11 k u  (k < 1.5e6 iterations)
    const double xk = 11.0 * (double)n_it * 5.9604644775390625e-08;
    const float eta = (float)(1.02 * (xk + xk * xk) + 1e-9);
    float sc[2] = {0.f, 0.f}, sb[2] = {0.f, 0.f};
    const int lm = (lane + 31) & 31, lp = (lane + 1) & 31;
    // labels in contact with another label: their gradient needs the neighbour's T, so T goes to the global float32
    // plane and k_flow_err32 takes the error from there
    float* T32b = q.T32 + (size_t)b * N;
    for (int r0 = 0; r0 < NR; r0 += 4) {
        // the flows of four rows first (unconditional loads at clamped addresses), then the arithmetic
        float a[4][4];
        #pragma unroll
        for (int q4 = 0; q4 < 4; q4++) {
            const int r = min(r0 + q4, NR - 1);
            const float2 v = S[(r + 1) * 32 + lane];
            const int p0 = (cpb_q32_member(v.x) && col_clean[0]) ? (c.yb[0] + r) * W + c.x[0] : 0;
            const int p1 = (cpb_q32_member(v.y) && col_clean[1]) ? (c.yb[1] + r) * W + c.x[1] : 0;
            a[q4][0] = dPy[p0]; a[q4][1] = dPx[p0]; a[q4][2] = dPy[p1]; a[q4][3] = dPx[p1];
        }
        #pragma unroll
        for (int q4 = 0; q4 < 4; q4++) {
            const int r = r0 + q4;
            if (r >= NR) break;
            const float2 v = S[(r + 1) * 32 + lane];
            if (!cpb_q32_member(v.x) && !cpb_q32_member(v.y)) continue;
            const float2 up = S[r * 32 + lane], dn = S[(r + 2) * 32 + lane];
            const float lf = fabsf(S[(r + 1) * 32 + lm].y), rt = fabsf(S[(r + 1) * 32 + lp].x);
            if (cpb_q32_member(v.x)) {
                if (col_clean[0]) cpb_q32_pixel(fabsf(up.x), fabsf(dn.x), lf, fabsf(v.y), a[q4][0], a[q4][1], eta, sc[0], sb[0]);
                else T32b[(c.yb[0] + r) * W + c.x[0]] = v.x;
            }
            if (cpb_q32_member(v.y)) {
                if (col_clean[1]) cpb_q32_pixel(fabsf(up.y), fabsf(dn.y), fabsf(v.x), rt, a[q4][2], a[q4][3], eta, sc[1], sb[1]);
                else T32b[(c.yb[1] + r) * W + c.x[1]] = v.y;
            }
        }
    }
    __syncwarp();
    float* red = reinterpret_cast<float*>(S);
    red[4 * lane + 0] = sc[0]; red[4 * lane + 1] = sb[0];
    red[4 * lane + 2] = sc[1]; red[4 * lane + 3] = sb[1];
    __syncwarp();
    if (lane < nsub) {
        const size_t k = (size_t)b * LC + e_l;
        if (!e_clean) {
            atomicOr(&q.info[k], CPB_QI_T32);
            q.lc[atomicAdd(&q.ctr[6], 1)] = make_int2(b, e_l);
        } else {
            float cs = 0.f, bd = 0.f;
            for (int i = e_off; i < e_off + e_w; i++) { cs += red[2 * i]; bd += red[2 * i + 1]; }
            cpb_q32_decide(t, q, b, e_l, k, cs, bd, threshold, pack_err);
        }
    }
    __syncwarp();
}

// k_diffuse32: persistent warps pull jobs (tile, first entry, nsub, class) from the queue k_qc_pack filled.
#ifndef CPB_Q32_MINBLOCKS
#define CPB_Q32_MINBLOCKS 4
#endif
CPB_KERNEL CPB_LAUNCH_BOUNDS(128, CPB_Q32_MINBLOCKS)
k_diffuse32(const int* CPB_RESTRICT lab, const float* CPB_RESTRICT dP, int H, int W, LabelTables t, Q32 q, double threshold,
            int pack_err) {
    CPB_SHARED float2 s_tile[4][(CPB_Q32_MAXNR + 2) * 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int njobs = q.ctr[0];
    for (;;) {
        int j = 0;
        if (lane == 0) j = atomicAdd(&q.ctr[1], 1);
        j = __shfl_sync(CPB_FULL, j, 0);
        if (j >= njobs) break;
        const int4 jb = q.sorted[j];
        cpb_q32_run(lab, dP, H, W, t, q, jb.x, jb.y, jb.z, jb.w, threshold, s_tile[warp], pack_err);
    }
}

// k_flow_err32: screened labels in contact with another label, one warp per label (lane = bbox column).  Same per-pixel
// arithmetic and bound as the register path, with T read from the global float32 plane: a neighbouring pixel of another
// live label contributes that label's T (as the reference's padded array does) -- it carries the same relative bound,
// having run the same number of iterations of the same arithmetic.  If a neighbour has no float32 T (it took the
// float64 path) or the bound does not clear the threshold, the label goes to the float64 list together with every
// label it touches (k_flow_err then finds the float64 T of all of them).
CPB_DEVICE float cpb_q32_T_at(const float* CPB_RESTRICT Tb, const int* CPB_RESTRICT L, const int* CPB_RESTRICT alive,
                              const int* CPB_RESTRICT info, int H, int W, int y, int x, int l, bool& missing) {
    if (y < 0 || y >= H || x < 0 || x >= W) return 0.f;
    const int p = y * W + x;
    const int v = L[p];
    if (v == l) return Tb[p];
    if (!cpb_foreign_live(v, l, alive)) return 0.f;
    if (!(info[v] & CPB_QI_T32)) { missing = true; return 0.f; }
    return Tb[p];
}

CPB_KERNEL CPB_LAUNCH_BOUNDS(128, 8)
k_flow_err32(const int* CPB_RESTRICT lab, const float* CPB_RESTRICT dP, int H, int W, LabelTables t, Q32 q, double threshold,
             int pack_err) {
    const int lane = threadIdx.x & 31;
    const int N = H * W, LC = t.LC;
    const int total = q.ctr[6];
    for (;;) {
        int j = 0;
        if (lane == 0) j = atomicAdd(&q.ctr[7], 1);
        j = __shfl_sync(CPB_FULL, j, 0);
        if (j >= total) break;
        const int2 e = q.lc[j];
        const int b = e.x, l = e.y;
        const size_t k = (size_t)b * LC + l;
        const int* L = lab + (size_t)b * N;
        const float* Tb = q.T32 + (size_t)b * N;
        const int* alive = t.alive ? t.alive + (size_t)b * LC : nullptr;
        const int* info = q.info + (size_t)b * LC;
        const float* dPy = dP + ((size_t)b * 2 + 0) * N;
        const float* dPx = dP + ((size_t)b * 2 + 1) * N;
        const int y0 = t.ymin[k], x0 = t.xmin[k];
        const int h = t.ymax[k] - y0 + 1, w = t.xmax[k] - x0 + 1;
        const double xk = 11.0 * (double)t.niter[b] * 5.9604644775390625e-08;
        const float eta = (float)(1.02 * (xk + xk * xk) + 1e-9);
        float sc = 0.f, sb = 0.f;
        bool missing = false;
        if (w <= 30) {
            // lane = column x0 - 1 + lane of the bbox grown by one; three rows (label's effective T, "no float32 T" mark)
            // slide through registers, left / right neighbours come by shuffle: ~4 loads per row and lane, none dependent
            const int xg = x0 - 1 + lane;
            const bool col_ok = lane < w + 2 && xg >= 0 && xg < W;
            float t_prev = 0.f, t_cur = 0.f, t_next = 0.f;
            bool m_prev = false, m_cur = false, m_next = false, mem_cur = false, mem_next = false;
            for (int ry = -2; ry < h; ry++) {
                // load row ry + 1 into "next" (row ry is "cur", row ry - 1 is "prev"); the first two rounds only fill the window
                const int yn = y0 + ry + 1;
                float tn = 0.f; bool mn = false, memn = false;
                if (ry + 1 <= h && col_ok && yn >= 0 && yn < H) {
                    const int p = yn * W + xg;
                    const int v = L[p];
                    if (v == l) { tn = Tb[p]; memn = true; }
                    else if (cpb_foreign_live(v, l, alive)) {
                        if (info[v] & CPB_QI_T32) tn = Tb[p]; else mn = true;
                    }
                }
                t_prev = t_cur; m_prev = m_cur;
                t_cur = t_next; m_cur = m_next; mem_cur = mem_next;
                t_next = tn; m_next = mn; mem_next = memn;
                // neighbours of row ry ("cur"); every lane takes part in the shuffles
                const float lf = __shfl_up_sync(CPB_FULL, t_cur, 1), rt = __shfl_down_sync(CPB_FULL, t_cur, 1);
                const int mlf = __shfl_up_sync(CPB_FULL, (int)m_cur, 1), mrt = __shfl_down_sync(CPB_FULL, (int)m_cur, 1);
                if (ry >= 0 && ry < h && mem_cur) {          // member lanes are 1 .. w: both shuffles have a source
                    missing |= m_prev || m_next || mlf != 0 || mrt != 0;
                    const int p = (y0 + ry) * W + xg;
                    cpb_q32_pixel(t_prev, t_next, lf, rt, dPy[p], dPx[p], eta, sc, sb);
                }
            }
        } else if (lane < w) {
            const int x = x0 + lane;
            for (int r = 0; r < h; r++) {
                const int y = y0 + r, p = y * W + x;
                if (L[p] != l) continue;
                const float up = cpb_q32_T_at(Tb, L, alive, info, H, W, y - 1, x, l, missing);
                const float dn = cpb_q32_T_at(Tb, L, alive, info, H, W, y + 1, x, l, missing);
                const float lf = cpb_q32_T_at(Tb, L, alive, info, H, W, y, x - 1, l, missing);
                const float rt = cpb_q32_T_at(Tb, L, alive, info, H, W, y, x + 1, l, missing);
                cpb_q32_pixel(up, dn, lf, rt, dPy[p], dPx[p], eta, sc, sb);
            }
        }
        for (int sft = 16; sft; sft >>= 1) {
            sc += __shfl_xor_sync(CPB_FULL, sc, sft);
            sb += __shfl_xor_sync(CPB_FULL, sb, sft);
        }
        missing = __any_sync(CPB_FULL, missing);
        int decided = 0;
        if (lane == 0) {
            if (missing) cpb_q32_queue64(q, b, l, k);
            else decided = cpb_q32_decide(t, q, b, l, k, sc, sb, threshold, pack_err) ? 1 : 0;
        }
        decided = __shfl_sync(CPB_FULL, decided, 0);
        if (!decided && lane < w) {
            // the float64 flow error of this label reads the float64 T of every label it touches
            const int x = x0 + lane;
            for (int r = 0; r < h; r++) {
                const int y = y0 + r;
                if (L[y * W + x] != l) continue;
                const int ny[4] = {y - 1, y + 1, y, y}, nx[4] = {x, x, x - 1, x + 1};
                for (int q4 = 0; q4 < 4; q4++) {
                    if (ny[q4] < 0 || ny[q4] >= H || nx[q4] < 0 || nx[q4] >= W) continue;
                    const int v = L[ny[q4] * W + nx[q4]];
                    // (labels beyond the warp kernels always get their float64 T from the block kernel k_diffuse)
                    if (cpb_foreign_live(v, l, alive) && (info[v] & 7) != CPB_Q32_CLSBIG) cpb_q32_queue64(q, b, v, (size_t)b * LC + v);
                }
            }
        }
    }
}

// k_diffuse64_list: the float64 warp kernel of cpb_qc.cuh driven by the explicit list q.l64 (one label per warp):
// labels in contact with another live label (they write T; their error comes from k_flow_err), labels the screen
// cannot hold, and labels the screen left undecided (error taken from the float64 tile, as before).
CPB_KERNEL CPB_LAUNCH_BOUNDS(CPB_DW_WARPS * 32, 6)
k_diffuse64_list(const int* CPB_RESTRICT lab, int H, int W, LabelTables t, double* CPB_RESTRICT T, Q32 q,
                 const float* CPB_RESTRICT dP, double threshold, int2* todo, int* todo_count, int todo_cap) {
    CPB_SHARED double s_T[CPB_DW_WARPS][(CPB_DC_MAXH + 3) * CPB_DC_PITCH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int N = H * W;
    const int total = q.ctr[2];
    for (;;) {
        int j = 0;
        if (lane == 0) j = atomicAdd(&q.ctr[3], 1);
        j = __shfl_sync(CPB_FULL, j, 0);
        if (j >= total) break;
        const int2 e = q.l64[j];
        const int b = e.x;
        DiffSub A;
        A.l = e.y; A.k = (size_t)b * t.LC + e.y; A.coff = 0;
        A.y0 = t.ymin[A.k]; A.x0 = t.xmin[A.k];
        A.h = t.ymax[A.k] - A.y0 + 1; A.w = t.xmax[A.k] - A.x0 + 1;
        const DiffQC qc{dP + ((size_t)b * 2 + 0) * N, dP + ((size_t)b * 2 + 1) * N,
                        t.alive ? t.alive + (size_t)b * t.LC : nullptr, threshold, H, todo, todo_count, todo_cap, b};
        cpb_diffuse_job<2>(lab + (size_t)b * N, W, t, T + (size_t)b * N, s_T[warp], A, A, false, t.niter[b], qc);
    }
}
