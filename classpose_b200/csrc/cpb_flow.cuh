// Row (2): follow_flows -- Euler integration of every foreground pixel through the
// bilinearly sampled flow field, then the end-point histogram of row (3).
// Semantics follow cellpose.dynamics.steps_interp with the arithmetic of
// torch's grid_sample (bilinear, zero padding, align_corners=False) on coordinates that
// were normalised by (L-1).
#pragma once
#include "cpb_common.cuh"

// The masked, scaled flow field is stored zero-padded (one row above / below, CPB_FLOW_PADX columns left /
// right), x component first:
//   flow[b][y+1][x+PADX] = ( dX*fg/5 * 2/(W-1) , dY*fg/5 * 2/(H-1) ),  row pitch Wp = W + 2*PADX.
// Clamped positions sample taps in [-1, L], so all four bilinear taps are always inside the padded array and
// grid_sample's zero padding costs no bounds checks (a zero tap adds v*w = 0).  PADX = 2 keeps the interior
// of every row 16-byte aligned for the vectorised writer.
#define CPB_FLOW_PADX 2

// End-point histogram with seed candidates on the side.  get_masks needs the pixels that collect more than 10 end points;
// the increment that takes a bin from <= 10 to > 10 happens exactly once per such pixel, so the kernel that counts can
// list them (tile-local pixel index; k_seeds reads the final count) and nobody has to stream the whole histogram to
// find them afterwards.  count == NULL: plain counting.
struct SeedCands { u64* key; int* count; int LC; };
CPB_DEVICE void cpb_hist_count(int* CPB_RESTRICT hist, int key, int cnt, int b, int p, const SeedCands& c) {
    if (c.count == nullptr) { atomicAdd(&hist[key], cnt); return; }
    const int old = atomicAdd(&hist[key], cnt);
    if (old <= CPB_SEED_MIN && old + cnt > CPB_SEED_MIN) {
        const int k = atomicAdd(&c.count[b], 1);
        if (k < c.LC) c.key[(size_t)b * c.LC + k] = (u64)(unsigned)p;
    }
}

CPB_DEVICE float2 cpb_scaled_flow(float dy, float dx, bool fg, float sx, float sy) {
    // (dP * fg) / 5.  then  *= 2/(L-1)   -- each a separately rounded float32 op.  Background is 0 whatever dP holds
    // (the reference's dP * 0 keeps the sign of dP on its zero, which no later operation can see), and it must not reach
    // the division: nvcc's division by a constant tests its operands with FCHK and a ZERO numerator takes the slow
    // path (a call, ~60 instructions) -- four pixels in five are background.  Foreground lanes divide dP itself,
    // background lanes a harmless 5, and the select picks.
    // (a three-operation exact x / 5 -- tests/studies/div5_exhaustive.cu, 0 mismatches over all float32 -- is the
    //  alternative without FCHK)
    const float qx = __fmul_rn(__fdiv_rn(fg ? dx : 5.0f, 5.0f), sx), qy = __fmul_rn(__fdiv_rn(fg ? dy : 5.0f, 5.0f), sy);
    return make_float2(fg ? qx : 0.0f, fg ? qy : 0.0f);
}

// k_prep_flow: one thread per padded pixel (any W).  Also writes bg_value on background (-1: the p_final contract;
// 0: the fused path points `pfinal` at the label image, which saves zeroing it separately) and appends
// foreground pixels to `list` (global pixel index in the un-padded layout), block-contiguous.
CPB_KERNEL CPB_LAUNCH_BOUNDS(256, 4)
k_prep_flow(const float* CPB_RESTRICT dP, const float* CPB_RESTRICT cellprob, int B, int H, int W,
            float thr, float sx, float sy, float2* CPB_RESTRICT flow, int* CPB_RESTRICT pfinal,
            unsigned* CPB_RESTRICT list, unsigned* CPB_RESTRICT list_n, int bg_value) {
    CPB_SHARED int s_scan[33];
    CPB_SHARED unsigned s_base;
    const int N = H * W, Wp = W + 2 * CPB_FLOW_PADX, Np = (H + 2) * Wp;
    const long long total = (long long)B * Np;
    const long long gp = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool fg = false;
    unsigned gi = 0;
    if (gp < total) {
        const int b = (int)(gp / Np);
        const int rp = (int)(gp - (long long)b * Np);
        const int yp = rp / Wp, xp = rp - yp * Wp;
        float2 out = make_float2(0.f, 0.f);
        if (yp >= 1 && yp <= H && xp >= CPB_FLOW_PADX && xp < W + CPB_FLOW_PADX) {
            const int r = (yp - 1) * W + (xp - CPB_FLOW_PADX);
            gi = (unsigned)b * (unsigned)N + (unsigned)r;
            fg = cellprob[gi] > thr;
            out = cpb_scaled_flow(dP[((size_t)b * 2 + 0) * N + r], dP[((size_t)b * 2 + 1) * N + r], fg, sx, sy);
            if (!fg || bg_value == 0) pfinal[gi] = bg_value;
        }
        flow[gp] = out;
    }
    int tot;
    const int incl = cpb_block_scan_incl(fg ? 1 : 0, s_scan, &tot);
    if (threadIdx.x == 0 && tot > 0) s_base = atomicAdd(list_n, (unsigned)tot);
    __syncthreads();
    if (fg) list[s_base + incl - 1] = gi;
}

// k_prep_flow_v4: W % 4 == 0.  One thread per group of 4 pixels of a padded row (rows -1 and H are the zero
// rows): three 128-bit loads, two 128-bit flow stores, one 128-bit p_final store.
#ifndef CPB_PREP_MINBLOCKS
#define CPB_PREP_MINBLOCKS 8
#endif
CPB_KERNEL CPB_LAUNCH_BOUNDS(256, CPB_PREP_MINBLOCKS)
k_prep_flow_v4(const float4* CPB_RESTRICT dP, const float4* CPB_RESTRICT cellprob, int B, int H, int W,
               float thr, float sx, float sy, float4* CPB_RESTRICT flow, int4* CPB_RESTRICT pfinal,
               unsigned* CPB_RESTRICT list, unsigned* CPB_RESTRICT list_n, int patch, int bg_value,
               float4* CPB_RESTRICT dP_copy) {
    CPB_SHARED int s_scan[33];
    CPB_SHARED unsigned s_base;
    const int W4 = W >> 2, N4 = (H * W) >> 2;
    const int Wp4 = (W + 2 * CPB_FLOW_PADX) >> 1;           // row pitch in float4 (2 pixels each)
    // thread -> (tile b, padded row yp, group of 4 pixels xg).  patch != 0 (W % 64 == 0): a block covers a
    // 16-row x 64-pixel patch, so that its foreground pixels -- one contiguous run of the list -- are whole
    // cells rather than 4-row slices (k_follow_merge merges trajectories within such a run).
    int b, yp, xg;
    bool in;
    if (patch) {
        const int pbx = W4 >> 4, pby = (H + 2 + 15) >> 4;
        const int blk = blockIdx.x;
        b = blk / (pbx * pby);
        const int rem = blk - b * (pbx * pby);
        yp = (rem / pbx) * 16 + (threadIdx.x >> 4);
        xg = (rem % pbx) * 16 + (threadIdx.x & 15);
        in = b < B && yp < H + 2;
    } else {
        const long long total = (long long)B * (H + 2) * W4;
        const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        in = g < total;
        b = (int)(g / ((long long)(H + 2) * W4));
        const int rem = (int)(g - (long long)b * (H + 2) * W4);
        yp = rem / W4; xg = rem - yp * W4;
    }
    int nfg = 0;
    unsigned gi0 = 0;
    bool f0 = false, f1 = false, f2 = false, f3 = false;
    if (in) {
        float4* row = flow + ((size_t)b * (H + 2) + yp) * Wp4;
        float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
        if (yp >= 1 && yp <= H) {
            const int y = yp - 1;
            const int q = y * W4 + xg;                          // float4 index inside the tile plane
            const float4 cp = cellprob[(size_t)b * N4 + q];
            f0 = cp.x > thr; f1 = cp.y > thr; f2 = cp.z > thr; f3 = cp.w > thr;
            // host path (dP_copy != NULL): dP is a mapped HOST pointer and is only read where it is used -- a group
            // without foreground is all zeros whatever dP holds -- so only these groups cross PCIe; dP_copy keeps
            // them on the device for the flow check
            float4 dy = make_float4(0.f, 0.f, 0.f, 0.f), dx = dy;
            if (dP_copy == nullptr) {          // device-resident flows: unconditional loads, in flight with cellprob's
                dy = dP[((size_t)b * 2 + 0) * N4 + q];
                dx = dP[((size_t)b * 2 + 1) * N4 + q];
            } else if (f0 || f1 || f2 || f3) {
                dy = dP[((size_t)b * 2 + 0) * N4 + q];
                dx = dP[((size_t)b * 2 + 1) * N4 + q];
                dP_copy[((size_t)b * 2 + 0) * N4 + q] = dy; dP_copy[((size_t)b * 2 + 1) * N4 + q] = dx;
            }
            const float2 a0 = cpb_scaled_flow(dy.x, dx.x, f0, sx, sy), a1 = cpb_scaled_flow(dy.y, dx.y, f1, sx, sy);
            const float2 a2 = cpb_scaled_flow(dy.z, dx.z, f2, sx, sy), a3 = cpb_scaled_flow(dy.w, dx.w, f3, sx, sy);
            o0 = make_float4(a0.x, a0.y, a1.x, a1.y);
            o1 = make_float4(a2.x, a2.y, a3.x, a3.y);
            gi0 = (unsigned)b * (unsigned)(H * W) + (unsigned)(q << 2);
            pfinal[(size_t)b * N4 + q] = make_int4(f0 ? 0 : bg_value, f1 ? 0 : bg_value, f2 ? 0 : bg_value, f3 ? 0 : bg_value);
            nfg = (int)f0 + (int)f1 + (int)f2 + (int)f3;
        }
        row[1 + 2 * xg] = o0;                                   // pixels 4xg, 4xg+1  (padded x = 4xg + 2)
        row[2 + 2 * xg] = o1;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (xg == 0) row[0] = z;                                // left pad (2 px)
        if (xg == W4 - 1) row[Wp4 - 1] = z;                     // right pad (2 px)
    }
    int tot;
    const int incl = cpb_block_scan_incl(nfg, s_scan, &tot);
    if (threadIdx.x == 0 && tot > 0) s_base = atomicAdd(list_n, (unsigned)tot);
    __syncthreads();
    unsigned o = s_base + (unsigned)(incl - nfg);
    if (f0) list[o++] = gi0;
    if (f1) list[o++] = gi0 + 1;
    if (f2) list[o++] = gi0 + 2;
    if (f3) list[o++] = gi0 + 3;
}

// One Euler step in normalised coordinates, arithmetic order as ATen's grid_sampler_2d.
// f points at pixel (0,0) of the padded tile; Wp is its row pitch.
// Packed FP32 (Blackwell FADD2 / FMUL2 / FFMA2 through PTX add/mul/fma.rn.f32x2): the x and y components of a
// position take the same arithmetic, so one instruction carries both -- each half is an ordinary IEEE
// round-to-nearest op, bit-identical to the scalar form.  A scalar weight packed as {w, w} is encoded by
// ptxas as a broadcast operand (R.F32), it costs no move.  The kernel is issue-bound: 28 instead of 39
// instructions per Euler step.  No packed mul feeds a packed add anywhere (ptxas would contract the pair).
#ifndef CPB_SIM
CPB_DEVICE u64 cpb_pk(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
CPB_DEVICE void cpb_upk(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
CPB_DEVICE u64 cpb_add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
CPB_DEVICE u64 cpb_add2_rm(u64 a, u64 b) { u64 r; asm("add.rm.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
CPB_DEVICE float cpb_clamp1(float v) { float r; asm("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(r) : "f"(v), "f"(1.0f)); return r; }
CPB_DEVICE u64 cpb_sub2(u64 a, u64 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
CPB_DEVICE u64 cpb_mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
CPB_DEVICE u64 cpb_fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
#endif

// WP > 0: row pitch known at compile time (both tap rows are addressed off one base register).
// The four taps of the previous step: a trajectory that has reached its sink keeps sampling the same 2 x 2 cell
// (sub-pixel steps), so the gathers -- the L1 data pipe is the busiest unit of the packed kernel, ~3 wavefronts per
// scattered 8-byte load -- are skipped while the tap index does not change.  Same values, same arithmetic.
struct EulerTaps { int key; float2 nw, ne, sw, se; u64 pm, mp, one0; };   // pm .. one0: constant pairs (cpb_taps_init)
#define CPB_TAPS_NONE 0x7fffffff

// CPB_EULER_VARIANT (compile time, A/B only; all bit-identical; follow_flows stage per 1024 conic tiles,
// profiles/r02/ab_euler_variants.txt): 0 round-1 form, 29 instructions of which 10 packed, 2.09 ms; 1 (default) the same
// with a one-instruction clamp, 27 / 10, 2.04 ms; 2 also FADD2.RM floor and paired tap distances, 25 / 15, 2.13 ms;
// 3 all scalar, 37 / 0, 2.35 ms.  A packed instruction holds the fmaheavy pipe for two cycles (FFMA2 : FFMA = 1.98 : 1
// in tests/studies/f32x2_issue.cu), so packing pays for operations that would otherwise be two instructions, and
// turning scalar work into extra packed work (variant 2) does not.
#ifndef CPB_EULER_VARIANT
#define CPB_EULER_VARIANT 1
#endif
template <int WP, bool PACKED>
CPB_DEVICE void cpb_euler_step_t(const float2* CPB_RESTRICT f, int Wp_rt, float fH, float fW, float& px, float& py,
                                 EulerTaps& tp) {
    const int Wp = WP > 0 ? WP : Wp_rt;
#ifndef CPB_SIM
  if (PACKED) {
#if CPB_EULER_VARIANT == 3
    // all-scalar form of the same step (tap reuse, FADD-formed tap index, one-instruction clamp): the A/B that
    // separates "fewer instructions" from "fewer issue cycles" for the packed forms
    const float ix = fmaf(__fadd_rn(px, 1.f), fW, -0.5f), iy = fmaf(__fadd_rn(py, 1.f), fH, -0.5f);
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const int key = __float_as_int(fmaf(fy0, (float)Wp, __fadd_rn(fx0, 12582912.f)));
    if (key != tp.key) {
        const float2* r0 = f + key;
        const float2* r1 = r0 + Wp;
        tp.nw = __ldg(r0); tp.ne = __ldg(r0 + 1); tp.sw = __ldg(r1); tp.se = __ldg(r1 + 1);
        tp.key = key;
    }
    const float ax1 = __fsub_rn(__fadd_rn(fx0, 1.f), ix), ay1 = __fsub_rn(__fadd_rn(fy0, 1.f), iy);
    const float ax0 = __fsub_rn(ix, fx0), ay0 = __fsub_rn(iy, fy0);
    const float wnw = __fmul_rn(ax1, ay1), wne = __fmul_rn(ax0, ay1);
    const float wsw = __fmul_rn(ax1, ay0), wse = __fmul_rn(ax0, ay0);
    float ox = __fmul_rn(tp.nw.x, wnw), oy = __fmul_rn(tp.nw.y, wnw);
    ox = fmaf(tp.ne.x, wne, ox); oy = fmaf(tp.ne.y, wne, oy);
    ox = fmaf(tp.sw.x, wsw, ox); oy = fmaf(tp.sw.y, wsw, oy);
    ox = fmaf(tp.se.x, wse, ox); oy = fmaf(tp.se.y, wse, oy);
    px = cpb_clamp1(__fadd_rn(px, ox));
    py = cpb_clamp1(__fadd_rn(py, oy));
#else
    // same operations, same order, two lanes (x, y) per instruction
    const u64 i2 = cpb_fma2(cpb_add2(cpb_pk(px, py), cpb_pk(1.f, 1.f)), cpb_pk(fW, fH), cpb_pk(-0.5f, -0.5f));
#if CPB_EULER_VARIANT == 2
    // floor without the XU pipe: i + 1.5 * 2^23 rounded toward -inf is exactly 1.5 * 2^23 + floor(i) (floats of that
    // binade are the integers, -0.5 <= i < 2^22), and subtracting the constant again is exact -- FADD2.RM + FADD2
    // instead of two FRND, and the x half of t2 is already the biased column the tap index needs.
    const u64 t2 = cpb_add2_rm(i2, cpb_pk(12582912.f, 12582912.f));
    const u64 f0 = cpb_add2(t2, cpb_pk(-12582912.f, -12582912.f));
    float tx, ty, fx0, fy0;
    cpb_upk(t2, tx, ty);
    cpb_upk(f0, fx0, fy0);
    const int key = __float_as_int(fmaf(fy0, (float)Wp, tx));
#else
    float ix, iy;
    cpb_upk(i2, ix, iy);
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const u64 f0 = cpb_pk(fx0, fy0);
    // tap index: fy0 * Wp + fx0 is an exact small integer; adding 1.5 * 2^23 leaves it in the low mantissa bits, so
    // the float -> int conversion (XU pipe) becomes an FADD.  `f` arrives biased by -0x4B400000 elements.
    const int key = __float_as_int(fmaf(fy0, (float)Wp, fx0 + 12582912.f));
#endif
    if (key != tp.key) {
        const float2* r0 = f + key;
        const float2* r1 = r0 + Wp;
        tp.nw = __ldg(r0); tp.ne = __ldg(r0 + 1); tp.sw = __ldg(r1); tp.se = __ldg(r1 + 1);
        tp.key = key;
    }
    float wnw, wne, wsw, wse;
#if CPB_EULER_VARIANT == 2
    // the two tap distances of an axis as ONE pair {f1 - i, i - f0}: {f0 + 1, -f0} is exact, and an fma whose
    // product is +-i is the subtraction itself (one rounding); a pair times a broadcast weight then gives two of
    // the four bilinear weights per FMUL2
    float ix, iy;
    cpb_upk(i2, ix, iy);
    const u64 pm = tp.pm, mp = tp.mp, one0 = tp.one0;      // {1, -1}, {-1, 1}, {1, 0}
    const u64 ax = cpb_fma2(cpb_pk(ix, ix), mp, cpb_fma2(cpb_pk(fx0, fx0), pm, one0));   // {ax1, ax0}
    const u64 ay = cpb_fma2(cpb_pk(iy, iy), mp, cpb_fma2(cpb_pk(fy0, fy0), pm, one0));   // {ay1, ay0}
    float ay1, ay0;
    cpb_upk(ay, ay1, ay0);
    cpb_upk(cpb_mul2(ax, cpb_pk(ay1, ay1)), wnw, wne);
    cpb_upk(cpb_mul2(ax, cpb_pk(ay0, ay0)), wsw, wse);
#else
    const u64 f1 = cpb_add2(f0, cpb_pk(1.f, 1.f));
    float ax1, ay1, ax0, ay0;
    cpb_upk(cpb_sub2(f1, i2), ax1, ay1);          // (fx1 - ix, fy1 - iy)
    cpb_upk(cpb_sub2(i2, f0), ax0, ay0);          // (ix - fx0, iy - fy0)
    wnw = __fmul_rn(ax1, ay1); wne = __fmul_rn(ax0, ay1);
    wsw = __fmul_rn(ax1, ay0); wse = __fmul_rn(ax0, ay0);
#endif
    u64 o = cpb_mul2(cpb_pk(tp.nw.x, tp.nw.y), cpb_pk(wnw, wnw));   // 0 + v*w
    o = cpb_fma2(cpb_pk(tp.ne.x, tp.ne.y), cpb_pk(wne, wne), o);
    o = cpb_fma2(cpb_pk(tp.sw.x, tp.sw.y), cpb_pk(wsw, wsw), o);
    o = cpb_fma2(cpb_pk(tp.se.x, tp.se.y), cpb_pk(wse, wse), o);
    float nx, ny;
    cpb_upk(cpb_add2(cpb_pk(px, py), o), nx, ny);
#if CPB_EULER_VARIANT == 0
    px = fminf(fmaxf(nx, -1.f), 1.f);
    py = fminf(fmaxf(ny, -1.f), 1.f);
#else
    // clamp to [-1, 1] in one instruction per coordinate: min(|n|, 1) with the sign of n (FMNMX.XORSIGN) is
    // fmin(fmax(n, -1), 1) for every n, signed zeros included
    px = cpb_clamp1(nx);
    py = cpb_clamp1(ny);
#endif
#endif
    return;
  }
#endif
    (void)tp;
    // ATen: ix = ((x + 1) * W - 1) / 2, which nvcc contracts to fma(x + 1, W, -1) * 0.5.  Scaling by 0.5 commutes
    // with rounding, so fma(x + 1, W/2, -0.5) is the same float with one instruction less (fH, fW arrive halved).
    const float ix = fmaf(px + 1.f, fW, -0.5f);
    const float iy = fmaf(py + 1.f, fH, -0.5f);
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    // tap index from the (integer-valued, < 2^24) floats: one conversion instead of two
    const int idx = (int)fmaf(fy0, (float)Wp, fx0);
    const float fx1 = fx0 + 1.f, fy1 = fy0 + 1.f;
    const float2* r0 = f + idx;      // f is an opaque per-thread pointer: one IMAD.WIDE
    const float2* r1 = r0 + Wp;
    const float2 vnw = __ldg(r0), vne = __ldg(r0 + 1), vsw = __ldg(r1), vse = __ldg(r1 + 1);
    const float wnw = (fx1 - ix) * (fy1 - iy);
    const float wne = (ix - fx0) * (fy1 - iy);
    const float wsw = (fx1 - ix) * (iy - fy0);
    const float wse = (ix - fx0) * (iy - fy0);
    float ox = vnw.x * wnw, oy = vnw.y * wnw;   // 0 + v*w
    ox = fmaf(vne.x, wne, ox); oy = fmaf(vne.y, wne, oy);
    ox = fmaf(vsw.x, wsw, ox); oy = fmaf(vsw.y, wsw, oy);
    ox = fmaf(vse.x, wse, ox); oy = fmaf(vse.y, wse, oy);
    px = fminf(fmaxf(px + ox, -1.f), 1.f);
    py = fminf(fmaxf(py + oy, -1.f), 1.f);
}
CPB_DEVICE void cpb_euler_step(const float2* CPB_RESTRICT f, int Wp, float fH, float fW, float& px, float& py) {
    EulerTaps unused;
    cpb_euler_step_t<0, false>(f, Wp, fH, fW, px, py, unused);   // scalar form: k_follow / k_follow_merge are the A/B references
}

// k_follow: grid-stride over the compacted foreground list, one pixel per thread.
#ifndef CPB_F_MINBLOCKS
#define CPB_F_MINBLOCKS 4
#endif
CPB_KERNEL CPB_LAUNCH_BOUNDS(256, CPB_F_MINBLOCKS)
k_follow(const float2* CPB_RESTRICT flow, const unsigned* CPB_RESTRICT list,
         const unsigned* CPB_RESTRICT list_n, int H, int W, int niter,
         int* CPB_RESTRICT pfinal, float* CPB_RESTRICT pfloat, int* CPB_RESTRICT hist, SeedCands cands) {
    const unsigned total = *list_n;
    const int N = H * W, Wp = W + 2 * CPB_FLOW_PADX, Np = (H + 2) * Wp;
    const float fW = 0.5f * (float)W, fH = 0.5f * (float)H;      // halved: see cpb_euler_step
    const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
    const int lane = threadIdx.x & 31;
    for (unsigned i0 = blockIdx.x * blockDim.x; i0 < total; i0 += gridDim.x * blockDim.x) {
        const unsigned i = i0 + threadIdx.x;
        const bool act = i < total;
        const unsigned amask = __ballot_sync(CPB_FULL, act);
        if (!act) continue;
        const unsigned gi = list[i];
        const int b = (int)(gi / (unsigned)N);
        const int r = (int)(gi - (unsigned)b * (unsigned)N);
        const int y = r / W, x = r - y * W;
        const float2* f = flow + (size_t)b * Np + Wp + CPB_FLOW_PADX;
#ifndef CPB_SIM
        asm volatile("" : "+l"(f));   // keep the tile base in a register pair (address = base + idx*8)
#endif
        // pt = idx / (L-1) * 2 - 1
        float px = __fsub_rn(__fmul_rn(__fdiv_rn((float)x, wm1), 2.f), 1.f);
        float py = __fsub_rn(__fmul_rn(__fdiv_rn((float)y, hm1), 2.f), 1.f);
        for (int t = 0; t < niter; t++) cpb_euler_step(f, Wp, fH, fW, px, py);
        // undo: (pt + 1) * 0.5 * (L-1)
        const float ex = __fmul_rn(__fmul_rn(__fadd_rn(px, 1.f), 0.5f), wm1);
        const float ey = __fmul_rn(__fmul_rn(__fadd_rn(py, 1.f), 0.5f), hm1);
        int xi = __float2int_rz(ex), yi = __float2int_rz(ey);
        xi = min(max(xi, 0), W - 1);
        yi = min(max(yi, 0), H - 1);
        pfinal[gi] = (yi << 16) | xi;
        if (pfloat) {
            pfloat[((size_t)b * 2 + 0) * N + r] = ey;
            pfloat[((size_t)b * 2 + 1) * N + r] = ex;
        }
        if (hist) {
            const int key = b * N + yi * W + xi;
            const unsigned peers = __match_any_sync(amask, key);
            if (lane == __ffs((int)peers) - 1) cpb_hist_count(hist, key, __popc(peers), b, yi * W + xi, cands);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// k_follow_staged (A/B experiment, CPB_FOLLOW_MERGE=3; north_star (2) names "tiles staged by TMA"): the plain kernel's
// arithmetic, with the flow window of a 32 x 32 pixel patch staged into shared memory by the TMA engine
// (cp.async.bulk global -> shared, one bulk copy per window row, completion on an mbarrier; UBLKCP in the SASS) for
// the first CPB_FS_STEPS Euler steps.  A tap whose 2 x 2 cell lies inside the staged window comes from shared
// memory, any other from global memory, so the result is bit-identical to k_follow whatever a trajectory does.
// Measured against k_follow on the B200: the staging does not pay -- the kernel is bound by issue
// slots, an LDS costs the slot of the L1-hit LDG it replaces, and the window test adds instructions.
#define CPB_FS_PATCH 32
#define CPB_FS_HALO 24
#define CPB_FS_STEPS 24
#define CPB_FS_WIN (CPB_FS_PATCH + 2 * CPB_FS_HALO)            // 80 rows x 80 columns of float2 = 51,200 bytes

#ifndef CPB_SIM
CPB_KERNEL CPB_LAUNCH_BOUNDS(256, 4)
k_follow_staged(const float2* CPB_RESTRICT flow, const float* CPB_RESTRICT cellprob, float thr, int B, int H, int W,
                int niter, int* CPB_RESTRICT pfinal, float* CPB_RESTRICT pfloat, int* CPB_RESTRICT hist, SeedCands cands) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    float2* s_win = reinterpret_cast<float2*>(s_raw);
    __shared__ __align__(8) unsigned long long s_bar;
    const int N = H * W, Wp = W + 2 * CPB_FLOW_PADX, Np = (H + 2) * Wp;
    const int pbx = (W + CPB_FS_PATCH - 1) / CPB_FS_PATCH, pby = (H + CPB_FS_PATCH - 1) / CPB_FS_PATCH;
    const int b = blockIdx.x / (pbx * pby);
    const int rem = blockIdx.x - b * (pbx * pby);
    const int py0 = (rem / pbx) * CPB_FS_PATCH, px0 = (rem % pbx) * CPB_FS_PATCH;
    // window in PADDED coordinates (row yp = y + 1, column xp = x + PADX), clipped to the padded tile, even start column
    const int wy0 = max(py0 + 1 - CPB_FS_HALO, 0), wy1 = min(py0 + 1 + CPB_FS_PATCH + CPB_FS_HALO, H + 2);
    const int wx0 = max(px0 + CPB_FLOW_PADX - CPB_FS_HALO, 0) & ~1;
    const int wx1 = min(wx0 + CPB_FS_WIN, Wp);
    const int wcols = (wx1 - wx0) & ~1, wrows = wy1 - wy0;
    const float2* tile = flow + (size_t)b * Np;
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned row_bytes = (unsigned)wcols * 8u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(row_bytes * (unsigned)wrows) : "memory");
        for (int r = 0; r < wrows; r++) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(s_win + r * CPB_FS_WIN);
            const float2* src = tile + (size_t)(wy0 + r) * Wp + wx0;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(dst), "l"(src), "r"(row_bytes), "r"(bar) : "memory");
        }
    }
    {   // every thread waits for the bytes to land (phase 0)
        unsigned done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(bar) : "memory");
    }
    const float fW = 0.5f * (float)W, fH = 0.5f * (float)H;
    const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
    const float2* f = tile + Wp + CPB_FLOW_PADX;                 // pixel (0, 0) of the padded tile
    const int lane = threadIdx.x & 31;
    // the patch's foreground pixels, compacted (full warps, like the list the other kernels walk)
    __shared__ unsigned short s_idx[CPB_FS_PATCH * CPB_FS_PATCH];
    __shared__ int s_scan[33];
    __shared__ int s_nfg;
    if (threadIdx.x == 0) s_nfg = 0;
    __syncthreads();
    for (int k = 0; k < (CPB_FS_PATCH * CPB_FS_PATCH) / 256; k++) {
        const int i = threadIdx.x + k * 256;
        const int y = py0 + i / CPB_FS_PATCH, x = px0 + (i % CPB_FS_PATCH);
        const bool fg = y < H && x < W && cellprob[(size_t)b * N + y * W + x] > thr;
        int tot;
        const int incl = cpb_block_scan_incl(fg ? 1 : 0, s_scan, &tot);
        if (fg) s_idx[s_nfg + incl - 1] = (unsigned short)i;
        __syncthreads();
        if (threadIdx.x == 0) s_nfg += tot;
        __syncthreads();
    }
    const int nfg = s_nfg;
    for (int e = threadIdx.x; e < ((nfg + 31) & ~31); e += 256) {
        const bool act = e < nfg;
        const unsigned amask = __ballot_sync(CPB_FULL, act);
        if (!act) continue;
        const int i = s_idx[e];
        const int y = py0 + i / CPB_FS_PATCH, x = px0 + (i % CPB_FS_PATCH);
        float px = __fsub_rn(__fmul_rn(__fdiv_rn((float)x, wm1), 2.f), 1.f);
        float py = __fsub_rn(__fmul_rn(__fdiv_rn((float)y, hm1), 2.f), 1.f);
        const int nst = min(CPB_FS_STEPS, niter);
        for (int t = 0; t < nst; t++) {
            // same operations as cpb_euler_step_t<0, false>, taps from the window when the 2 x 2 cell is inside it
            const float ix = fmaf(px + 1.f, fW, -0.5f), iy = fmaf(py + 1.f, fH, -0.5f);
            const float fx0 = floorf(ix), fy0 = floorf(iy);
            const int tx = (int)fx0 + CPB_FLOW_PADX, ty = (int)fy0 + 1;          // padded coordinates of the NW tap
            float2 vnw, vne, vsw, vse;
            if (tx >= wx0 && tx + 1 < wx0 + wcols && ty >= wy0 && ty + 1 < wy1) {
                const float2* r0 = s_win + (ty - wy0) * CPB_FS_WIN + (tx - wx0);
                vnw = r0[0]; vne = r0[1]; vsw = r0[CPB_FS_WIN]; vse = r0[CPB_FS_WIN + 1];
            } else {
                const float2* r0 = f + ((int)fy0 * Wp + (int)fx0);
                vnw = __ldg(r0); vne = __ldg(r0 + 1); vsw = __ldg(r0 + Wp); vse = __ldg(r0 + Wp + 1);
            }
            const float fx1 = fx0 + 1.f, fy1 = fy0 + 1.f;
            const float wnw = (fx1 - ix) * (fy1 - iy), wne = (ix - fx0) * (fy1 - iy);
            const float wsw = (fx1 - ix) * (iy - fy0), wse = (ix - fx0) * (iy - fy0);
            float ox = vnw.x * wnw, oy = vnw.y * wnw;
            ox = fmaf(vne.x, wne, ox); oy = fmaf(vne.y, wne, oy);
            ox = fmaf(vsw.x, wsw, ox); oy = fmaf(vsw.y, wsw, oy);
            ox = fmaf(vse.x, wse, ox); oy = fmaf(vse.y, wse, oy);
            px = fminf(fmaxf(px + ox, -1.f), 1.f);
            py = fminf(fmaxf(py + oy, -1.f), 1.f);
        }
        for (int t = nst; t < niter; t++) cpb_euler_step(f, Wp, fH, fW, px, py);
        const float ex = __fmul_rn(__fmul_rn(__fadd_rn(px, 1.f), 0.5f), wm1);
        const float ey = __fmul_rn(__fmul_rn(__fadd_rn(py, 1.f), 0.5f), hm1);
        int xi = __float2int_rz(ex), yi = __float2int_rz(ey);
        xi = min(max(xi, 0), W - 1);
        yi = min(max(yi, 0), H - 1);
        const int r = y * W + x;
        pfinal[(size_t)b * N + r] = (yi << 16) | xi;
        if (pfloat) {
            pfloat[((size_t)b * 2 + 0) * N + r] = ey;
            pfloat[((size_t)b * 2 + 1) * N + r] = ex;
        }
        if (hist) {
            const int key = b * N + yi * W + xi;
            const unsigned peers = __match_any_sync(amask, key);
            if (lane == __ffs((int)peers) - 1) cpb_hist_count(hist, key, __popc(peers), b, yi * W + xi, cands);
        }
    }
}
#endif

// ---------------------------------------------------------------------------------------------------------
// k_follow_merge: same integration, with exact trajectory merging.  The Euler map is deterministic, so two
// pixels of a tile whose float32 positions are bitwise equal at some step stay equal forever; pixels of a cell
// fall into the same attractor and do become bitwise equal (about a third of the trajectories of a 256-pixel
// chunk are duplicates after ~50 steps, more than half after ~100).  At two merge points the block
// deduplicates its positions through a shared-memory hash table, compacts the distinct trajectories to the
// low threads and lets whole warps go idle for the remaining steps; at the end every pixel reads the end
// point of the trajectory it was merged into.  Results are bit-identical to k_follow.
//
// A chunk of the foreground list can span a few tiles (it is a concatenation of per-block segments); the tile
// "group" (number of tile changes before the entry, at most 3) is folded into the key using bit 30 of each
// coordinate, which is always 0 for |v| <= 1.  Chunks with more than 4 groups are integrated without merging.
#define CPB_FM_THREADS 256
#define CPB_FM_SLOTS 512
#define CPB_FM_EMPTY 0xffffffffffffffffull

// Deduplicate the keys of threads [0, nact) (every thread is its own trajectory when `unique`): returns this
// thread's compact trajectory index (valid for tid < nact); distinct trajectories are numbered in thread order
// of their first inserter, *n_out receives their number.  Contains __syncthreads().
CPB_DEVICE int cpb_block_merge(u64 key, int nact, bool unique, u64* s_keys, int* s_own, int* s_scan, int* n_out) {
    const int t = threadIdx.x;
    for (int i = t; i < CPB_FM_SLOTS; i += CPB_FM_THREADS) s_keys[i] = CPB_FM_EMPTY;
    __syncthreads();
    int slot = 0;
    bool owner = false;
    if (t < nact) {
        if (unique) {
            owner = true;
        } else {
            unsigned h = (unsigned)((key * 0x9E3779B97F4A7C15ull) >> 55) & (CPB_FM_SLOTS - 1);
            for (;;) {
                const u64 old = atomicCAS(&s_keys[h], CPB_FM_EMPTY, key);
                if (old == CPB_FM_EMPTY) { owner = true; break; }
                if (old == key) break;
                h = (h + 1) & (CPB_FM_SLOTS - 1);
            }
            slot = (int)h;
        }
    }
    int tot;
    const int incl = cpb_block_scan_incl(owner ? 1 : 0, s_scan, &tot);      // contains __syncthreads()
    if (owner && !unique) s_own[slot] = incl - 1;
    __syncthreads();
    *n_out = tot;
    if (t >= nact) return 0;
    return unique ? incl - 1 : s_own[slot];
}

#ifndef CPB_FM_MINBLOCKS
#define CPB_FM_MINBLOCKS 6
#endif
CPB_KERNEL CPB_LAUNCH_BOUNDS(CPB_FM_THREADS, CPB_FM_MINBLOCKS)
k_follow_merge(const float2* CPB_RESTRICT flow, const unsigned* CPB_RESTRICT list,
               const unsigned* CPB_RESTRICT list_n, int H, int W, int niter, int m1, int m2,
               int* CPB_RESTRICT pfinal, float* CPB_RESTRICT pfloat, int* CPB_RESTRICT hist, SeedCands cands) {
    CPB_SHARED u64 s_keys[CPB_FM_SLOTS];
    CPB_SHARED int s_own[CPB_FM_SLOTS];
    CPB_SHARED float2 s_pos[CPB_FM_THREADS];
    CPB_SHARED int s_tile[CPB_FM_THREADS];      // tile (bits 0..27) and group (bits 28..29) of a trajectory
    CPB_SHARED unsigned short s_map1[CPB_FM_THREADS], s_map2[CPB_FM_THREADS];
    CPB_SHARED int s_scan[33];
    const unsigned total = *list_n;
    const int N = H * W, Wp = W + 2 * CPB_FLOW_PADX, Np = (H + 2) * Wp;
    const float fW = 0.5f * (float)W, fH = 0.5f * (float)H;      // halved: see cpb_euler_step
    const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
    const int t = threadIdx.x, lane = t & 31;
    for (unsigned i0 = blockIdx.x * blockDim.x; i0 < total; i0 += gridDim.x * blockDim.x) {
        const int nact = (int)min((unsigned)CPB_FM_THREADS, total - i0);
        const bool act = t < nact;
        unsigned gi = 0;
        int b = 0, r = 0;
        float px = 0.f, py = 0.f;
        if (act) {
            gi = list[i0 + t];
            b = (int)(gi / (unsigned)N);
            r = (int)(gi - (unsigned)b * (unsigned)N);
            const int y = r / W, x = r - y * W;
            px = __fsub_rn(__fmul_rn(__fdiv_rn((float)x, wm1), 2.f), 1.f);
            py = __fsub_rn(__fmul_rn(__fdiv_rn((float)y, hm1), 2.f), 1.f);
        }
        // tile group of every entry = number of tile changes before it inside the chunk
        s_tile[t] = b;
        __syncthreads();
        const int change = (act && t > 0 && s_tile[t - 1] != b) ? 1 : 0;
        int nchange;
        const int grp = cpb_block_scan_incl(change, s_scan, &nchange);
        const bool unique = nchange > 3;                                   // block-uniform
        __syncthreads();
        // ---- segment 0: every pixel
        {
            const float2* f = flow + (size_t)b * Np + Wp + CPB_FLOW_PADX;
#ifndef CPB_SIM
            asm volatile("" : "+l"(f));
#endif
            if (act) for (int s = 0; s < m1; s++) cpb_euler_step(f, Wp, fH, fW, px, py);
        }
        const unsigned gbits = unique ? 0u : (unsigned)grp;
        u64 key = ((u64)(__float_as_uint(py) | ((gbits >> 1) << 30)) << 32) | (u64)(__float_as_uint(px) | ((gbits & 1u) << 30));
        int n1;
        const int c1 = cpb_block_merge(key, nact, unique, s_keys, s_own, s_scan, &n1);
        if (act) {
            s_map1[t] = (unsigned short)c1;
            s_pos[c1] = make_float2(px, py);          // all members hold the same position and tile / group
            s_tile[c1] = b | ((int)gbits << 28);
        }
        __syncthreads();
        // ---- segment 1: distinct trajectories only (threads 0 .. n1-1)
        const bool act1 = t < n1;
        float qx = 0.f, qy = 0.f;
        int tg1 = 0;
        if (act1) { qx = s_pos[t].x; qy = s_pos[t].y; tg1 = s_tile[t]; }
        {
            const float2* f = flow + (size_t)(tg1 & 0x0fffffff) * Np + Wp + CPB_FLOW_PADX;
#ifndef CPB_SIM
            asm volatile("" : "+l"(f));
#endif
            if (act1) for (int s = m1; s < m2; s++) cpb_euler_step(f, Wp, fH, fW, qx, qy);
        }
        const unsigned g1 = (unsigned)tg1 >> 28;
        key = ((u64)(__float_as_uint(qy) | ((g1 >> 1) << 30)) << 32) | (u64)(__float_as_uint(qx) | ((g1 & 1u) << 30));
        __syncthreads();                               // everyone has read s_pos / s_tile
        int n2;
        const int c2 = cpb_block_merge(key, n1, unique, s_keys, s_own, s_scan, &n2);
        if (act1) {
            s_map2[t] = (unsigned short)c2;
            s_pos[c2] = make_float2(qx, qy);
            s_tile[c2] = tg1;
        }
        __syncthreads();
        // ---- segment 2
        const bool act2 = t < n2;
        float rx = 0.f, ry = 0.f;
        int tg2 = 0;
        if (act2) { rx = s_pos[t].x; ry = s_pos[t].y; tg2 = s_tile[t]; }
        {
            const float2* f = flow + (size_t)(tg2 & 0x0fffffff) * Np + Wp + CPB_FLOW_PADX;
#ifndef CPB_SIM
            asm volatile("" : "+l"(f));
#endif
            if (act2) for (int s = m2; s < niter; s++) cpb_euler_step(f, Wp, fH, fW, rx, ry);
        }
        __syncthreads();
        if (act2) s_pos[t] = make_float2(rx, ry);
        __syncthreads();
        // ---- every pixel reads the end point of the trajectory it was merged into
        const unsigned amask = __ballot_sync(CPB_FULL, act);
        if (act) {
            const float2 e = s_pos[s_map2[s_map1[t]]];
            const float ex = __fmul_rn(__fmul_rn(__fadd_rn(e.x, 1.f), 0.5f), wm1);
            const float ey = __fmul_rn(__fmul_rn(__fadd_rn(e.y, 1.f), 0.5f), hm1);
            int xi = __float2int_rz(ex), yi = __float2int_rz(ey);
            xi = min(max(xi, 0), W - 1);
            yi = min(max(yi, 0), H - 1);
            pfinal[gi] = (yi << 16) | xi;
            if (pfloat) {
                pfloat[((size_t)b * 2 + 0) * N + r] = ey;
                pfloat[((size_t)b * 2 + 1) * N + r] = ex;
            }
            if (hist) {
                const int hkey = b * N + yi * W + xi;
                const unsigned peers = __match_any_sync(amask, hkey);
                if (lane == __ffs((int)peers) - 1) cpb_hist_count(hist, hkey, __popc(peers), b, yi * W + xi, cands);
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------
// k_follow_pool: exact trajectory merging with a shared-memory trajectory pool and many merge points.
// A block owns a chunk of CPB_FP_POOL consecutive entries of the foreground list (about five 16 x 64 patches,
// i.e. whole cells) and keeps one position per DISTINCT trajectory in shared memory.  Between merge points the
// live trajectories are integrated 256 at a time (warps beyond the live count skip the segment); at a merge point
// bitwise-equal (tile, x, y) are found through a hash of trajectory INDICES (32-bit CAS on the slot, full key
// compared in the pool, so the key may be as wide as it likes), survivors are compacted and every pixel's
// trajectory index is redirected.  On the synthetic conic tiles the live count falls to 72 % at step 30, 46 % at
// step 50, 23 % at step 100 and 16 % at step 200, so six merge points leave ~46 % of the Euler steps of the
// plain kernel (two merge points over 256 pixels: ~60 %).  Which duplicate survives depends on the CAS race, but
// duplicates hold the same bits, so the output is bit-identical to k_follow.
#define CPB_FP_THREADS 256
#ifndef CPB_FP_POOL
#define CPB_FP_POOL 1024        // entries of the foreground list per block (throughput form)
#endif
#define CPB_FP_POOL_SMALL 256   // latency form for a handful of tiles: one trajectory per thread, four times the blocks
#define CPB_FP_MAXMERGE 16

struct FollowSchedule { int n; int at[CPB_FP_MAXMERGE]; };   // merge points (step numbers), ascending, < niter

#ifndef CPB_FP_MINBLOCKS
#define CPB_FP_MINBLOCKS 6
#endif
template <int WP, int POOL>
CPB_KERNEL CPB_LAUNCH_BOUNDS(CPB_FP_THREADS, CPB_FP_MINBLOCKS)
k_follow_pool(const float2* CPB_RESTRICT flow, const unsigned* CPB_RESTRICT list,
              const unsigned* CPB_RESTRICT list_n, int H, int W, int niter, FollowSchedule sch,
              int* CPB_RESTRICT pfinal, float* CPB_RESTRICT pfloat, int* CPB_RESTRICT hist, SeedCands cands) {
    constexpr int PER = POOL / CPB_FP_THREADS, SLOTS = 2 * POOL;
    CPB_SHARED float2 s_pos[POOL];               // position of live trajectory i
    CPB_SHARED int s_tile[POOL];                 // its tile
    CPB_SHARED int s_slot[SLOTS];                // hash slot -> trajectory index (-1 empty)
    CPB_SHARED unsigned short s_cur[POOL];       // pixel -> live trajectory
    CPB_SHARED unsigned short s_new[POOL];       // trajectory -> index after the merge
    CPB_SHARED int s_scan[33];
    const unsigned total = *list_n;
    const int N = H * W, Wp = W + 2 * CPB_FLOW_PADX, Np = (H + 2) * Wp;
    const float fW = 0.5f * (float)W, fH = 0.5f * (float)H;      // halved: see cpb_euler_step
    const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
    const int t = threadIdx.x, lane = t & 31;
    for (unsigned i0 = blockIdx.x * POOL; i0 < total; i0 += gridDim.x * POOL) {
        const int n0 = (int)min((unsigned)POOL, total - i0);
        unsigned gi[PER];
#pragma unroll
        for (int k = 0; k < PER; k++) {
            const int i = t + k * CPB_FP_THREADS;
            gi[k] = 0;
            if (i < n0) {
                gi[k] = list[i0 + i];
                const int b = (int)(gi[k] / (unsigned)N);
                const int r = (int)(gi[k] - (unsigned)b * (unsigned)N);
                const int y = r / W, x = r - y * W;
                // pt = idx / (L-1) * 2 - 1
                s_pos[i] = make_float2(__fsub_rn(__fmul_rn(__fdiv_rn((float)x, wm1), 2.f), 1.f),
                                       __fsub_rn(__fmul_rn(__fdiv_rn((float)y, hm1), 2.f), 1.f));
                s_tile[i] = b;
                s_cur[i] = (unsigned short)i;
            }
        }
        __syncthreads();
        int n = n0, step = 0;
        EulerTaps tp;
#if !defined(CPB_SIM) && CPB_EULER_VARIANT == 2
        // the constant pairs of variant 2 stay in registers only if they are built from a per-thread run-time value
        // (ptxas folds literals back into immediates / uniform registers and rebuilds the pairs with moves every step)
        const bool tv = __float_as_int(s_pos[t].x) != 0x7fc00123, tw = __float_as_int(s_pos[t].y) != 0x7fc00123;   // always true
        const float c1 = tv ? 1.f : 2.f, cm1 = tv ? -1.f : 2.f, c0 = tw ? 0.f : 2.f, d1 = tw ? 1.f : 2.f;
        tp.pm = cpb_pk(c1, cm1); tp.mp = cpb_pk(cm1, c1); tp.one0 = cpb_pk(d1, c0);
#endif
        for (int m = 0; m <= sch.n; m++) {
            const int until = m < sch.n ? sch.at[m] : niter;
            // ---- integrate the live trajectories from `step` to `until`
            for (int i = t; i < n; i += CPB_FP_THREADS) {
                float2 p = s_pos[i];
                const float2* f = flow + (size_t)s_tile[i] * Np + Wp + CPB_FLOW_PADX;
#ifndef CPB_SIM
                f -= 0x4B400000;       // bias of the FADD-based tap index (see cpb_euler_step_t)
                asm volatile("" : "+l"(f));
#endif
                tp.key = CPB_TAPS_NONE;
                for (int s = step; s < until; s++) cpb_euler_step_t<WP, true>(f, Wp, fH, fW, p.x, p.y, tp);
                s_pos[i] = p;
            }
            step = until;
            if (m == sch.n) break;
            // ---- merge: thread t owns trajectories [t*per, t*per + per)
            // table of the smallest power of two >= 2n slots (load factor <= 1/2)
            const int smask = n <= 32 ? 63 : min(SLOTS, 1 << (33 - __clz(n - 1))) - 1;
            for (int i = t; i <= smask; i += CPB_FP_THREADS) s_slot[i] = -1;
            __syncthreads();                                  // positions written, table cleared
            const int per = (n + CPB_FP_THREADS - 1) / CPB_FP_THREADS;   // <= PER
            int rep[PER];                              // -1: owner, else the trajectory it duplicates
            float2 mp[PER];
            int mt[PER];
            int owners = 0;
#pragma unroll
            for (int k = 0; k < PER; k++) {
                const int i = t * per + k;
                rep[k] = -2;
                if (k < per && i < n) {
                    const float2 p = s_pos[i];
                    const int tl = s_tile[i];
                    mp[k] = p; mt[k] = tl;
                    const unsigned ux = __float_as_uint(p.x), uy = __float_as_uint(p.y);
                    unsigned h = (ux * 0x9E3779B1u) ^ (uy * 0x85EBCA77u) ^ ((unsigned)tl * 0xC2B2AE3Du);
                    h = (h ^ (h >> 15)) & smask;
                    for (;;) {
                        const int old = atomicCAS(&s_slot[h], -1, i);
                        if (old == -1) { rep[k] = -1; owners++; break; }
                        const float2 q = s_pos[old];
                        if (__float_as_uint(q.x) == ux && __float_as_uint(q.y) == uy && s_tile[old] == tl) { rep[k] = old; break; }
                        h = (h + 1) & smask;
                    }
                }
            }
            int tot;
            int idx = cpb_block_scan_incl(owners, s_scan, &tot) - owners;     // contains __syncthreads()
#pragma unroll
            for (int k = 0; k < PER; k++)
                if (rep[k] == -1) s_new[t * per + k] = (unsigned short)(idx++);
            __syncthreads();                                  // owners' new indices visible; all keys were read
            idx -= owners;
#pragma unroll
            for (int k = 0; k < PER; k++) {
                if (rep[k] == -1) { s_pos[idx] = mp[k]; s_tile[idx] = mt[k]; idx++; }
            }
            // (duplicates resolve through their representative after the owners' entries are final)
            unsigned short dupnew[PER];
#pragma unroll
            for (int k = 0; k < PER; k++) dupnew[k] = rep[k] >= 0 ? s_new[rep[k]] : (unsigned short)0;
            __syncthreads();
#pragma unroll
            for (int k = 0; k < PER; k++)
                if (rep[k] >= 0) s_new[t * per + k] = dupnew[k];
            __syncthreads();
            for (int i = t; i < n0; i += CPB_FP_THREADS) s_cur[i] = s_new[s_cur[i]];
            n = tot;
            __syncthreads();
        }
        __syncthreads();                                      // final positions visible
        // ---- every pixel reads the end point of the trajectory it was merged into
#pragma unroll
        for (int k = 0; k < PER; k++) {
            const int i = t + k * CPB_FP_THREADS;
            const bool act = i < n0;
            const unsigned amask = __ballot_sync(CPB_FULL, act);
            if (act) {
                const float2 e = s_pos[s_cur[i]];
                const unsigned g = gi[k];
                const int b = (int)(g / (unsigned)N);
                const int r = (int)(g - (unsigned)b * (unsigned)N);
                // undo: (pt + 1) * 0.5 * (L-1)
                const float ex = __fmul_rn(__fmul_rn(__fadd_rn(e.x, 1.f), 0.5f), wm1);
                const float ey = __fmul_rn(__fmul_rn(__fadd_rn(e.y, 1.f), 0.5f), hm1);
                int xi = __float2int_rz(ex), yi = __float2int_rz(ey);
                xi = min(max(xi, 0), W - 1);
                yi = min(max(yi, 0), H - 1);
                pfinal[g] = (yi << 16) | xi;
                if (pfloat) {
                    pfloat[((size_t)b * 2 + 0) * N + r] = ey;
                    pfloat[((size_t)b * 2 + 1) * N + r] = ex;
                }
                if (hist) {
                    const int hkey = b * N + yi * W + xi;
                    const unsigned peers = __match_any_sync(amask, hkey);
                    if (lane == __ffs((int)peers) - 1) cpb_hist_count(hist, hkey, __popc(peers), b, yi * W + xi, cands);
                }
            }
        }
        __syncthreads();
    }
}
