// Next row N4: the tile preparation in front of the network.
//   /root/reference/src/classpose/models.py:641-666  transforms.normalize_img  -> per-channel 1st / 99th percentile
//                                                     normalisation (cellpose normalize99, numpy percentile, linear)
//   /root/reference/src/classpose/core.py:129-178    np.pad to the /16 grid, transforms.make_tiles (+ parity flips)
// k_percentiles finds the exact order statistics with a 4-pass radix select per (image, channel);
// k_make_tiles writes the normalised, zero-padded, flipped sub-tiles in the layout the network consumes.
#pragma once
#include "cpb_common.cuh"

CPB_DEVICE unsigned cpb_float_key(float f) {          // order-preserving map float -> uint
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
CPB_DEVICE float cpb_key_float(unsigned k) {
    const unsigned u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

// value with 0-based rank `rank` among the n values img[i*stride] (block-cooperative, 4 passes of 8 bits)
CPB_DEVICE float cpb_block_select(const float* CPB_RESTRICT img, int n, int stride, int rank, unsigned* s_hist,
                                  unsigned* s_state) {
    unsigned prefix = 0, mask = 0;
    int r = rank;
    for (int pass = 3; pass >= 0; pass--) {
        const int shift = pass * 8;
        for (int i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const unsigned k = cpb_float_key(img[(size_t)i * stride]);
            if ((k & mask) == prefix) atomicAdd(&s_hist[(k >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int acc = 0, bkt = 0;
            for (; bkt < 256; bkt++) {
                if (acc + (int)s_hist[bkt] > r) break;
                acc += (int)s_hist[bkt];
            }
            s_state[0] = (unsigned)bkt; s_state[1] = (unsigned)(r - acc);
        }
        __syncthreads();
        prefix |= s_state[0] << shift;
        mask |= 255u << shift;
        r = (int)s_state[1];
        __syncthreads();
    }
    return cpb_key_float(prefix);
}

// numpy's _lerp(a, b, t) of the 'linear' percentile method.  For a float32 array numpy keeps everything in
// float32: q/100 (np.true_divide(q, float32(100))), the virtual index (n-1)*q, gamma and the interpolation.
CPB_DEVICE float cpb_np_lerp(float a, float b, float t) {
    const float d = __fsub_rn(b, a);
    float v = __fadd_rn(a, __fmul_rn(d, t));
    if (t >= 0.5f) v = __fsub_rn(b, __fmul_rn(d, __fsub_rn(1.0f, t)));
    return v;
}

// one block per (image b, channel c): lowhigh[b][c] = (x01, x99 - x01), code[b][c] = 1 normalise, 2 zero, 0 leave
CPB_KERNEL CPB_LAUNCH_BOUNDS(1024, 1)
k_percentiles(const float* CPB_RESTRICT img, int H, int W, int C, double lower, double upper,
              float* CPB_RESTRICT lowhigh, int* CPB_RESTRICT code) {
    CPB_SHARED unsigned s_hist[256];
    CPB_SHARED unsigned s_state[2];
    const int b = blockIdx.x / C, c = blockIdx.x % C;
    const int n = H * W;
    const float* src = img + (size_t)b * n * C + c;
    float out[2];
    for (int q = 0; q < 2; q++) {
        const float q32 = __fdiv_rn((float)(q == 0 ? lower : upper), 100.0f);
        const float vi = __fmul_rn((float)(n - 1), q32);
        int lo = (int)floorf(vi);
        const float g = __fsub_rn(vi, (float)lo);
        lo = min(max(lo, 0), n - 1);
        const int hi = min(lo + 1, n - 1);
        const float a = cpb_block_select(src, n, C, lo, s_hist, s_state);
        const float bb = cpb_block_select(src, n, C, hi, s_hist, s_state);
        out[q] = cpb_np_lerp(a, bb, g);
    }
    const float mn = cpb_block_select(src, n, C, 0, s_hist, s_state);
    const float mx = cpb_block_select(src, n, C, n - 1, s_hist, s_state);
    if (threadIdx.x == 0) {
        const float rng = __fsub_rn(out[1], out[0]);
        lowhigh[(size_t)blockIdx.x * 2] = out[0];
        lowhigh[(size_t)blockIdx.x * 2 + 1] = rng;
        code[blockIdx.x] = !(mx > mn) ? 0 : (rng > 1e-3f ? 1 : 2);      // np.ptp > 0 ; x99 - x01 > 1e-3
    }
}

// one thread per output element of tiles[b][j][c][sy][sx]
CPB_KERNEL CPB_LAUNCH_BOUNDS(256, 4)
k_make_tiles(const float* CPB_RESTRICT img, int B, int H, int W, int C, int pad_y, int pad_x, int ntiles, int ly, int lx,
             const int* CPB_RESTRICT ty0, const int* CPB_RESTRICT tx0, const int* CPB_RESTRICT flip,
             const float* CPB_RESTRICT lowhigh, const int* CPB_RESTRICT code, float* CPB_RESTRICT tiles) {
    const long long total = (long long)B * ntiles * C * ly * lx;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    const int sx = (int)(g % lx);
    const int sy = (int)((g / lx) % ly);
    const int c = (int)((g / ((long long)lx * ly)) % C);
    const int j = (int)((g / ((long long)lx * ly * C)) % ntiles);
    const int b = (int)(g / ((long long)lx * ly * C * ntiles));
    const int f = flip[j];
    const int ry = (f & 1) ? ly - 1 - sy : sy, rx = (f & 2) ? lx - 1 - sx : sx;     // position inside the un-flipped window
    const int y = ty0[j] + ry - pad_y, x = tx0[j] + rx - pad_x;                    // position in the un-padded image
    float v = 0.f;
    if (y >= 0 && y < H && x >= 0 && x < W) {
        v = img[(((size_t)b * H + y) * W + x) * C + c];
        const int k = b * C + c;
        const int cd = code[k];
        if (cd == 1) v = __fdiv_rn(__fsub_rn(v, lowhigh[2 * k]), lowhigh[2 * k + 1]);
        else if (cd == 2) v = 0.f;
    }
    tiles[g] = v;
}
