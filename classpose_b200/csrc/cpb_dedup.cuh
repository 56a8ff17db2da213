// Next row N2: overlap de-duplication of cells across tiles
// (/root/reference/src/classpose/entrypoints/predict_wsi.py:896-965: scipy KDTree.query_pairs(7.5) on the centroids,
// group linked cells, keep the largest of each group).  Here: uniform grid hash of the centroids (cell size =
// max_dist), lock-free union-find over all pairs within max_dist, then per component the largest cell survives
// (ties: lowest index).  For isolated pairs and cliques -- what tile overlaps produce -- this equals the
// reference's greedy grouping; for chains the reference's result depends on Python's set iteration order.
#pragma once
#include "cpb_common.cuh"

CPB_DEVICE unsigned cpb_grid_hash(long long gx, long long gy, unsigned mask) {
    u64 h = (u64)gx * 0x9E3779B97F4A7C15ull ^ ((u64)gy * 0xC2B2AE3D27D4EB4Full + 0x165667B19E3779F9ull);
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    return (unsigned)h & mask;
}

CPB_KERNEL k_dedup_insert(const double* CPB_RESTRICT cx, const double* CPB_RESTRICT cy, long long n, double inv_cell,
                          unsigned mask, int* CPB_RESTRICT head, int* CPB_RESTRICT next, int* CPB_RESTRICT parent,
                          u64* CPB_RESTRICT best_size, int* CPB_RESTRICT best_idx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long gx = (long long)floor(cx[i] * inv_cell), gy = (long long)floor(cy[i] * inv_cell);
    next[i] = atomicExch(&head[cpb_grid_hash(gx, gy, mask)], (int)i);
    parent[i] = (int)i;
    best_size[i] = 0;
    best_idx[i] = CPB_IMAX;
}

CPB_DEVICE int cpb_uf_find(int* parent, int x) {
    for (;;) {
        const int p = parent[x];
        if (p == x) return x;
        const int gp = parent[p];
        if (gp != p) parent[x] = gp;        // path halving (benign race: only ever shortens towards a root)
        x = p;
    }
}

CPB_DEVICE void cpb_uf_union(int* parent, int a, int b) {
    for (;;) {
        a = cpb_uf_find(parent, a);
        b = cpb_uf_find(parent, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }      // link the larger root under the smaller
        if (atomicCAS(&parent[a], a, b) == a) return;
    }
}

CPB_KERNEL k_dedup_link(const double* CPB_RESTRICT cx, const double* CPB_RESTRICT cy, long long n, double inv_cell,
                        double r2, unsigned mask, const int* CPB_RESTRICT head, const int* CPB_RESTRICT next,
                        int* CPB_RESTRICT parent) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = cx[i], y = cy[i];
    const long long gx = (long long)floor(x * inv_cell), gy = (long long)floor(y * inv_cell);
    for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) {
            const long long ngx = gx + dx, ngy = gy + dy;
            for (int j = head[cpb_grid_hash(ngx, ngy, mask)]; j >= 0; j = next[j]) {
                if (j <= (int)i) continue;                       // every pair once
                const double xj = cx[j], yj = cy[j];
                if ((long long)floor(xj * inv_cell) != ngx || (long long)floor(yj * inv_cell) != ngy) continue;
                const double ddx = xj - x, ddy = yj - y;
                if (ddx * ddx + ddy * ddy <= r2) cpb_uf_union(parent, (int)i, j);
            }
        }
}

// phase 0: best size per component; phase 1: lowest index among the cells of that size; phase 2: keep flags
CPB_KERNEL k_dedup_select(const double* CPB_RESTRICT size, long long n, int* CPB_RESTRICT parent,
                          u64* CPB_RESTRICT best_size, int* CPB_RESTRICT best_idx, int phase,
                          int* CPB_RESTRICT keep, int* CPB_RESTRICT group) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int root = cpb_uf_find(parent, (int)i);
    const double s = size[i];
    u64 bits;
    {   // order-preserving map of a double onto u64 (NaN / negative sizes do not occur; handled anyway)
        long long v = *reinterpret_cast<const long long*>(&s);
        bits = v < 0 ? ~(u64)v : ((u64)v | 0x8000000000000000ull);
    }
    if (phase == 0) atomicMax(&best_size[root], bits);
    else if (phase == 1) { if (best_size[root] == bits) atomicMin(&best_idx[root], (int)i); }
    else { keep[i] = best_idx[root] == (int)i ? 1 : 0; if (group) group[i] = root; }
}
