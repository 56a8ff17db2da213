// Next row N1: what the reference's PostProcessor computes per cell on the host
// (/root/reference/src/classpose/entrypoints/predict_wsi.py:595-656): ndimage.find_objects (bbox),
// cv2.findContours(cell_mask, RETR_EXTERNAL, CHAIN_APPROX_SIMPLE)[0], and the polygon's area / perimeter /
// centroid / validity (shapely).  Here: one thread per cell follows the outer border exactly as OpenCV's
// Suzuki-Abe implementation does (same start pixel, same search order, same "emit on direction change"
// compression), so the point list is identical to cv2's; polygon measures come from exact integer sums.
#pragma once
#include "cpb_common.cuh"

// neighbour codes of OpenCV: 0 E, 1 NE, 2 N, 3 NW, 4 W, 5 SW, 6 S, 7 SE   (y grows downwards)
CPB_DEVICE int cpb_cdx(int s) { return (s == 0 || s == 1 || s == 7) ? 1 : ((s >= 3 && s <= 5) ? -1 : 0); }
CPB_DEVICE int cpb_cdy(int s) { return (s >= 1 && s <= 3) ? -1 : ((s >= 5 && s <= 7) ? 1 : 0); }

struct ContourAcc {
    int n;                 // points emitted
    long long a2;          // sum of cross products  x_i*y_{i+1} - x_{i+1}*y_i   (twice the signed area)
    long long sx, sy;      // sums of (x_i + x_{i+1}) * cross_i  /  (y_i + y_{i+1}) * cross_i
    double perim;
    int fx, fy, px, py;    // first / previous emitted point
};

CPB_DEVICE void cpb_acc_point(ContourAcc& a, int x, int y, short* out) {
    if (out) { out[2 * a.n] = (short)x; out[2 * a.n + 1] = (short)y; }
    if (a.n == 0) { a.fx = x; a.fy = y; }
    else {
        const long long cr = (long long)a.px * y - (long long)x * a.py;
        a.a2 += cr; a.sx += (long long)(a.px + x) * cr; a.sy += (long long)(a.py + y) * cr;
        const int dx = x - a.px, dy = y - a.py;
        a.perim += sqrt((double)(dx * dx + dy * dy));
    }
    a.px = x; a.py = y; a.n++;
}

CPB_DEVICE void cpb_acc_close(ContourAcc& a) {
    if (a.n >= 2) {
        const int x = a.fx, y = a.fy;
        const long long cr = (long long)a.px * y - (long long)x * a.py;
        a.a2 += cr; a.sx += (long long)(a.px + x) * cr; a.sy += (long long)(a.py + y) * cr;
        const int dx = x - a.px, dy = y - a.py;
        a.perim += sqrt((double)(dx * dx + dy * dy));
    }
}

// Follow the outer border of the 8-connected component of label `l` that starts at (x0, y0) -- its first pixel in
// raster order -- the way OpenCV's icvFetchContour does for an outer border with CHAIN_APPROX_SIMPLE.
// marks (optional): visited border pixels get OpenCV's values, 2 (border pixel) or -126 (border pixel whose right
// side leaves the component) -- the scanner needs them to tell "inside a traced contour" from "outside".
// out (optional): receives the points as (x, y) int16 pairs.
CPB_DEVICE void cpb_trace_outer(const int* CPB_RESTRICT L, int H, int W, int l, int x0, int y0, int* marks,
                                short* out, ContourAcc& acc) {
    acc.n = 0; acc.a2 = 0; acc.sx = 0; acc.sy = 0; acc.perim = 0.0; acc.fx = acc.fy = acc.px = acc.py = 0;
    #define CPB_PIX(xx, yy) ((xx) >= 0 && (xx) < W && (yy) >= 0 && (yy) < H && L[(yy) * W + (xx)] == l)
    int s = 4, s_end = 4;
    int x1 = x0, y1 = y0;
    do {
        s = (s - 1) & 7;
        x1 = x0 + cpb_cdx(s); y1 = y0 + cpb_cdy(s);
    } while (!CPB_PIX(x1, y1) && s != s_end);
    if (s == s_end) {                       // single-pixel component
        if (marks) marks[y0 * W + x0] = -126;
        cpb_acc_point(acc, x0, y0, out);
        return;
    }
    int x3 = x0, y3 = y0, x4 = x0, y4 = y0;
    int prev_s = s ^ 4;
    int ptx = x0, pty = y0;
    for (;;) {
        s_end = s;
        while (s < 15) {
            ++s;
            x4 = x3 + cpb_cdx(s & 7); y4 = y3 + cpb_cdy(s & 7);
            if (CPB_PIX(x4, y4)) break;
        }
        s &= 7;
        if (marks) {                            // "right bound" check of icvFetchContour
            int* m = &marks[y3 * W + x3];
            if ((unsigned)(s - 1) < (unsigned)s_end) *m = -126;
            else if (*m == 0) *m = 2;
        }
        if (s != prev_s) { cpb_acc_point(acc, ptx, pty, out); prev_s = s; }
        ptx += cpb_cdx(s); pty += cpb_cdy(s);
        if (x4 == x0 && y4 == y0 && x3 == x1 && y3 == y1) break;
        x3 = x4; y3 = y4;
        s = (s + 4) & 7;
    }
    #undef CPB_PIX
}

CPB_DEVICE int cpb_orient(int ax, int ay, int bx, int by, int cx, int cy) {
    const long long v = (long long)(bx - ax) * (cy - ay) - (long long)(by - ay) * (cx - ax);
    return v > 0 ? 1 : (v < 0 ? -1 : 0);
}
CPB_DEVICE bool cpb_on_seg(int ax, int ay, int bx, int by, int px, int py) {   // p collinear with ab: inside the box?
    return px >= min(ax, bx) && px <= max(ax, bx) && py >= min(ay, by) && py <= max(ay, by);
}
// closed segments ab and cd share at least one point
CPB_DEVICE bool cpb_seg_touch(int ax, int ay, int bx, int by, int cx, int cy, int dx, int dy) {
    const int o1 = cpb_orient(ax, ay, bx, by, cx, cy), o2 = cpb_orient(ax, ay, bx, by, dx, dy);
    const int o3 = cpb_orient(cx, cy, dx, dy, ax, ay), o4 = cpb_orient(cx, cy, dx, dy, bx, by);
    if (o1 != o2 && o3 != o4) return true;
    if (o1 == 0 && cpb_on_seg(ax, ay, bx, by, cx, cy)) return true;
    if (o2 == 0 && cpb_on_seg(ax, ay, bx, by, dx, dy)) return true;
    if (o3 == 0 && cpb_on_seg(cx, cy, dx, dy, ax, ay)) return true;
    if (o4 == 0 && cpb_on_seg(cx, cy, dx, dy, bx, by)) return true;
    return false;
}

// Ring validity in the sense the reference relies on (shapely Polygon.is_valid on the closed contour): at least
// 4 distinct points' worth of ring, non-zero area, no two non-adjacent edges touching, no spike between
// adjacent edges.  O(n^2) on the emitted points (a nucleus contour has a few dozen).
CPB_DEVICE bool cpb_ring_valid(const short* p, int n, long long a2) {
    if (n < 4 || a2 == 0) return false;
    for (int i = 0; i < n; i++) {
        const int i1 = (i + 1) % n, i2 = (i + 2) % n;
        const int ax = p[2 * i], ay = p[2 * i + 1], bx = p[2 * i1], by = p[2 * i1 + 1];
        const int cx = p[2 * i2], cy = p[2 * i2 + 1];
        // spike: next edge folds back onto this one
        if (cpb_orient(ax, ay, bx, by, cx, cy) == 0 &&
            (long long)(bx - ax) * (cx - bx) + (long long)(by - ay) * (cy - by) < 0) return false;
        for (int j = i + 2; j < n; j++) {
            if (i == 0 && j == n - 1) continue;           // adjacent through the closing edge
            const int j1 = (j + 1) % n;
            if (cpb_seg_touch(ax, ay, bx, by, p[2 * j], p[2 * j + 1], p[2 * j1], p[2 * j1 + 1])) return false;
        }
    }
    return true;
}

// Per-cell outputs (indexed [b][l], l = 1..lbound[b]):
//   npts   int32    points of contours[0] (0 when the label is absent)
//   feat   int64x8  pixel area, ymin, ymax, xmin, xmax, A2, Sx, Sy   (polygon area = |A2|/2, centroid = Sx/(3 A2), Sy/(3 A2))
//   perim  float64  polygon perimeter (tile pixels)
//   start  int32    raster index of the first pixel of the traced component (second pass)
// pass 0 (points == NULL): trace every component of every label in raster order (marks suppress restarts) and keep
//   the LAST one -- cv2 returns contours in reverse scan order, so that is contours[0].
// pass 1: re-trace from `start`, write the points at offsets[b][l], evaluate validity.
CPB_KERNEL CPB_LAUNCH_BOUNDS(128, 8)
k_contours(const int* CPB_RESTRICT lab, int H, int W, LabelTables t, int* CPB_RESTRICT marks,
           int* CPB_RESTRICT npts, long long* CPB_RESTRICT feat, double* CPB_RESTRICT perim, int* CPB_RESTRICT start,
           const long long* CPB_RESTRICT offsets, short* CPB_RESTRICT points, long long points_cap,
           int* CPB_RESTRICT valid) {
    const int b = blockIdx.y, LC = t.LC, N = H * W;
    const int lb = t.lbound[b];
    const int* L = lab + (size_t)b * N;
    for (int l = 1 + blockIdx.x * blockDim.x + threadIdx.x; l <= lb; l += gridDim.x * blockDim.x) {
        const size_t k = (size_t)b * LC + l;
        ContourAcc acc;
        if (points == nullptr) {
            npts[k] = 0; start[k] = -1; perim[k] = 0.0;
            for (int q = 0; q < 8; q++) feat[k * 8 + q] = 0;
            if (t.cnt[k] <= 0) continue;
            const int y0 = t.ymin[k], y1 = t.ymax[k], x0 = t.xmin[k], x1 = t.xmax[k];
            int* mk = marks + (size_t)b * N;
            // OpenCV's scanner in RETR_EXTERNAL mode: a border starts where an unmarked component pixel (value 1)
            // follows a background pixel, unless the last marked pixel met on this row is a positive mark (we are
            // inside an already traced contour, e.g. to the right of a hole).  Values: 0 bg, 1 unmarked, 2 / -126.
            for (int y = y0; y <= y1; y++) {
                int prev = 0, lnbd = 0;
                for (int x = x0; x <= x1; x++) {
                    const int p = y * W + x;
                    int v = (L[p] == l) ? (mk[p] ? mk[p] : 1) : 0;
                    if (v == prev) continue;
                    if (prev == 0 && v == 1 && lnbd <= 0) {
                        cpb_trace_outer(L, H, W, l, x, y, mk, nullptr, acc);
                        cpb_acc_close(acc);
                        npts[k] = acc.n; start[k] = p; perim[k] = acc.perim;
                        feat[k * 8 + 5] = acc.a2; feat[k * 8 + 6] = acc.sx; feat[k * 8 + 7] = acc.sy;
                        v = mk[p];
                    }
                    prev = v;
                    if (v != 0 && v != 1) lnbd = v;
                }
            }
            feat[k * 8 + 0] = t.cnt[k];
            feat[k * 8 + 1] = y0; feat[k * 8 + 2] = y1; feat[k * 8 + 3] = x0; feat[k * 8 + 4] = x1;
        } else {
            valid[k] = 0;
            const int n = npts[k];
            if (n <= 0) continue;
            const long long off = offsets[k];
            if (off + n > points_cap) continue;           // caller's buffer too small: total is reported, nothing written
            const int p = start[k];
            short* out = points + 2 * off;
            cpb_trace_outer(L, H, W, l, p % W, p / W, nullptr, out, acc);
            valid[k] = cpb_ring_valid(out, n, feat[k * 8 + 5]) ? 1 : 0;
        }
    }
}

// offsets[b][l] = exclusive prefix sum of npts over (b, l) in (tile, label) order; total[0] = sum.
// One block per tile computes the tile's sum (phase 0) / the per-label offsets from the tile base (phase 1).
CPB_KERNEL k_contour_tile_sums(const int* CPB_RESTRICT npts, LabelTables t, int* CPB_RESTRICT tile_sum) {
    CPB_SHARED int s_part[32];
    const int b = blockIdx.x, lb = t.lbound[b];
    int part = 0;
    for (int l = 1 + threadIdx.x; l <= lb; l += blockDim.x) part += npts[(size_t)b * t.LC + l];
    for (int d = 16; d; d >>= 1) part += __shfl_xor_sync(CPB_FULL, part, d);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int q = 0; q < (int)(blockDim.x >> 5); q++) tot += s_part[q];
        tile_sum[b] = tot;
    }
}

CPB_KERNEL k_contour_offsets(const int* CPB_RESTRICT npts, LabelTables t, const long long* CPB_RESTRICT tile_base,
                             long long* CPB_RESTRICT offsets) {
    CPB_SHARED int s_scan[33];
    CPB_SHARED long long s_carry;
    const int b = blockIdx.x, lb = t.lbound[b];
    if (threadIdx.x == 0) s_carry = tile_base[b];
    __syncthreads();
    for (int l0 = 1; l0 <= lb; l0 += blockDim.x) {
        const int l = l0 + threadIdx.x;
        const int c = l <= lb ? npts[(size_t)b * t.LC + l] : 0;
        int tot;
        const int incl = cpb_block_scan_incl(c, s_scan, &tot);
        if (l <= lb) offsets[(size_t)b * t.LC + l] = s_carry + incl - c;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += tot;
        __syncthreads();
    }
}
