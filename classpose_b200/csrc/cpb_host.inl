// Host-buffer entry points: H2D -> fused path -> D2H, chunked over a ring of streams so the copies of one chunk
// overlap the kernels of the others.  Device buffers are cached per host thread (the reference calls in from
// >= 2 threads per process; nothing is shared between them).
//
// The path is PCIe-bound (40 bytes per pixel up in the plain form: dP 8, cellprob 4, logits 4 C), so the
// bytes are what is optimised:
//  * logits of a PINNED (or registered) host buffer are not uploaded at all: the final label pass reads them
//    through the mapped host pointer, and it only touches the 4-pixel groups that hold a cell (about a quarter
//    of them), so only those cross the bus (CPB_HOST_LOGITS_MAPPED, the default when the buffer allows it);
//  * dP of a pinned buffer is likewise read in place by the prep kernel, only where a 4-pixel group holds
//    foreground; those groups are kept in a device copy for the flow check.  cellprob is needed everywhere and is
//    always uploaded.  (Kernel reads of host memory move 64-byte half lines: tests/studies/zerocopy_bw.cu.)
//  * cell_class rows are copied back `cc_width` entries wide (the table is LC ~ N/11 entries per tile, a tile
//    holds ~100 cells); a chunk whose largest count exceeds the width is copied again in full after the sync;
//  * label masks can be delivered as uint16 (the dtype Cellpose returns below 65536 labels).
#include <algorithm>
#include <vector>

namespace {

constexpr int kHostSlots = 3;

struct HostSlot {
    cudaStream_t stream = nullptr;
    char* blob = nullptr;
    size_t cap = 0;
};

struct HostCtx {
    int device = -1;
    HostSlot slot[kHostSlots];
    int cc_width = 512;            // entries of a cell_class row copied back before the counts are known
    void release() {
        for (auto& s : slot) {
            if (s.blob) { cudaFree(s.blob); s.blob = nullptr; s.cap = 0; }
            if (s.stream) { cudaStreamDestroy(s.stream); s.stream = nullptr; }
        }
    }
    ~HostCtx() { release(); }
};

thread_local HostCtx tl_ctx;

struct ChunkBufs {
    float* dP; float* cellprob; float* logits; int32_t* masks; uint16_t* masks16; int32_t* counts; int32_t* cell_class;
    uint8_t* class_masks; void* ws; size_t ws_bytes; size_t total;
};

ChunkBufs carve_chunk(char* base, int Bc, int H, int W, int C, bool has_logits, bool upload_logits, bool has_cm, bool u16) {
    const size_t N = (size_t)H * W;
    const int LC = cpb_label_capacity(H, W);
    Carver c{base, 0};
    ChunkBufs b{};
    b.dP = c.take<float>(2 * N * Bc);                 // uploaded flows, or the device copy of the groups read in place
    b.cellprob = c.take<float>(N * Bc);
    b.logits = c.take<float>(upload_logits ? (size_t)C * N * Bc : 0);
    b.masks = c.take<int32_t>(N * Bc);
    b.masks16 = c.take<uint16_t>(u16 ? N * Bc : 0);
    b.counts = c.take<int32_t>(Bc);
    b.cell_class = c.take<int32_t>(has_logits ? (size_t)LC * Bc : 0);
    b.class_masks = c.take<uint8_t>(has_cm ? N * Bc : 0);
    b.ws_bytes = cpb_workspace_bytes(Bc, H, W, has_logits ? C : 0, 0);
    b.ws = c.take<char>(b.ws_bytes);
    b.total = c.off;
    return b;
}

CPB_KERNEL k_narrow_u16(const int4* CPB_RESTRICT in, long long n4, uint2* CPB_RESTRICT out) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n4) return;
    const int4 v = in[g];
    out[g] = make_uint2((unsigned)(v.x & 0xffff) | ((unsigned)v.y << 16), (unsigned)(v.z & 0xffff) | ((unsigned)v.w << 16));
}

// device-visible alias of a host pointer when the allocation is pinned / registered, else NULL
const float* mapped_alias(const float* host) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (a.type != cudaMemoryTypeHost || !a.devicePointer) return nullptr;
    return static_cast<const float*>(a.devicePointer);
}

}  // namespace

extern "C" int cpb_compute_masks_host_ex(const float* dP, const float* cellprob, const float* logits, int B, int H,
                                         int W, int C, const cpb_params* prm, void* masks, int32_t* counts,
                                         int32_t* cell_class, uint8_t* class_masks, const cpb_host_options* opt) {
    if (!dP || !cellprob || !prm || !masks || !counts || !opt || B <= 0 || H < 2 || W < 2) return CPB_E_ARG;
    if (logits && (!cell_class || C < 1)) return CPB_E_ARG;
    cudaError_t ce = cudaSetDevice(opt->device);
    if (ce != cudaSuccess) return (int)ce;
    const size_t N = (size_t)H * W;
    const int LC = cpb_label_capacity(H, W);
    int Bc = opt->tiles_per_chunk > 0 ? opt->tiles_per_chunk : 128;
    Bc = std::min(Bc, B);
    while ((long long)Bc * H * W >= (1LL << 31)) Bc /= 2;
    if (Bc < 1) return CPB_E_RANGE;
    const bool has_logits = logits != nullptr, has_cm = has_logits && class_masks != nullptr;
    const bool u16 = opt->masks_u16 != 0 && N % 4 == 0;
    if (opt->masks_u16 != 0 && !u16) return CPB_E_ARG;
    // logits through the mapped host pointer: only when the vote rides on the final pass (it then reads nothing but
    // the groups under a cell) and the buffer is device-visible
    const float* lg_alias = nullptr;
    if (has_logits && opt->logits_mode != CPB_HOST_LOGITS_UPLOAD && !has_cm && !prm->remove_border && N % 4 == 0 &&
        reinterpret_cast<uintptr_t>(logits) % 16 == 0)
        lg_alias = mapped_alias(logits);
    if (has_logits && opt->logits_mode == CPB_HOST_LOGITS_MAPPED && !lg_alias) return CPB_E_ARG;
    const bool upload_logits = has_logits && !lg_alias;
    // dP through the mapped pointer: the vectorised prep kernel reads only the groups with foreground
    const float* dp_alias = nullptr;
    if (opt->flows_mode != CPB_HOST_LOGITS_UPLOAD && W % 4 == 0 && reinterpret_cast<uintptr_t>(dP) % 16 == 0)
        dp_alias = mapped_alias(dP);
    if (opt->flows_mode == CPB_HOST_LOGITS_MAPPED && !dp_alias) return CPB_E_ARG;
    const size_t need = carve_chunk(nullptr, Bc, H, W, C, has_logits, upload_logits, has_cm, u16).total + kAlign;

    HostCtx& ctx = tl_ctx;
    if (ctx.device != opt->device) { ctx.release(); ctx.device = opt->device; }
    for (auto& s : ctx.slot) {
        if (!s.stream && (ce = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking)) != cudaSuccess) return (int)ce;
        if (s.cap < need) {
            if (s.blob) cudaFree(s.blob);
            s.blob = nullptr; s.cap = 0;
            if ((ce = cudaMalloc(&s.blob, need)) != cudaSuccess) return (int)ce;
            s.cap = need;
        }
    }
    const int ccw = std::min(LC, std::max(ctx.cc_width, 16));
    // chunk schedule: small chunks at both ends (the first upload and the last download are not overlapped by
    // anything), full chunks in between
    std::vector<int> sizes;
    {
        const int small = std::max(1, Bc / 8);
        int left = B, up = small;
        std::vector<int> tail;
        while (left > 0) {
            const int n = std::min(left, up);
            sizes.push_back(n); left -= n;
            if (left > 0 && up < Bc) {          // mirror the ramp at the end
                const int m = std::min(left, up);
                tail.push_back(m); left -= m;
            }
            up = std::min(Bc, up * 2);
        }
        sizes.insert(sizes.end(), tail.rbegin(), tail.rend());
    }
    int rc = 0;
    for (int b0 = 0, k = 0; b0 < B; b0 += sizes[k], k++) {
        const int nb = sizes[k];
        HostSlot& s = ctx.slot[k % kHostSlots];
        ChunkBufs cb = carve_chunk(s.blob, Bc, H, W, C, has_logits, upload_logits, has_cm, u16);
        cudaStream_t st = s.stream;
        if (!dp_alias)
            cudaMemcpyAsync(cb.dP, dP + (size_t)b0 * 2 * N, (size_t)nb * 2 * N * sizeof(float), cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(cb.cellprob, cellprob + (size_t)b0 * N, (size_t)nb * N * sizeof(float), cudaMemcpyHostToDevice, st);
        const float* lg = nullptr;
        if (upload_logits) {
            cudaMemcpyAsync(cb.logits, logits + (size_t)b0 * C * N, (size_t)nb * C * N * sizeof(float), cudaMemcpyHostToDevice, st);
            lg = cb.logits;
        } else if (has_logits) {
            lg = lg_alias + (size_t)b0 * C * N;
        }
        rc = compute_masks_impl(dp_alias ? dp_alias + (size_t)b0 * 2 * N : cb.dP, cb.cellprob, lg, nb, H, W, C, prm, cb.masks,
                                cb.counts, has_logits ? cb.cell_class : nullptr, has_cm ? cb.class_masks : nullptr, cb.ws,
                                cb.ws_bytes, st, nullptr, dp_alias ? cb.dP : nullptr);
        if (rc) break;
        if (u16) {
            const long long n4 = (long long)nb * (long long)(N / 4);
            CPB_LAUNCH_COUNTED(k_narrow_u16, dim3(blocks_for(n4, 256)), dim3(256), 0, st, reinterpret_cast<const int4*>(cb.masks),
                               n4, reinterpret_cast<uint2*>(cb.masks16));
            cudaMemcpyAsync(static_cast<uint16_t*>(masks) + (size_t)b0 * N, cb.masks16, (size_t)nb * N * sizeof(uint16_t),
                            cudaMemcpyDeviceToHost, st);
        } else {
            cudaMemcpyAsync(static_cast<int32_t*>(masks) + (size_t)b0 * N, cb.masks, (size_t)nb * N * sizeof(int32_t),
                            cudaMemcpyDeviceToHost, st);
        }
        cudaMemcpyAsync(counts + b0, cb.counts, (size_t)nb * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
        if (has_logits)
            cudaMemcpy2DAsync(cell_class + (size_t)b0 * LC, (size_t)LC * sizeof(int32_t), cb.cell_class,
                              (size_t)LC * sizeof(int32_t), (size_t)ccw * sizeof(int32_t), nb, cudaMemcpyDeviceToHost, st);
        if (has_cm)
            cudaMemcpyAsync(class_masks + (size_t)b0 * N, cb.class_masks, (size_t)nb * N, cudaMemcpyDeviceToHost, st);
        // a slot is reused every kHostSlots chunks: its previous D2H copies must have left before the next chunk
        // of this slot overwrites the buffers -- same stream, so stream order already guarantees that
    }
    for (auto& s : ctx.slot) {
        ce = cudaStreamSynchronize(s.stream);
        if (ce != cudaSuccess && rc == 0) rc = (int)ce;
    }
    if (rc) return rc;
    int top = 0;
    for (int b = 0; b < B; b++) {
        if (counts[b] < 0) return CPB_E_CAPACITY;
        top = std::max(top, counts[b]);
    }
    if (has_logits && top + 1 > ccw) {
        // rare: some tile holds more cells than the speculative row width.  The device rows of the LAST kHostSlots
        // chunks are still resident; earlier chunks are gone, so redo the call once with a width that fits.
        ctx.cc_width = std::min(LC, (top + 1) * 3 / 2);
        return cpb_compute_masks_host_ex(dP, cellprob, logits, B, H, W, C, prm, masks, counts, cell_class, class_masks, opt);
    }
    return 0;
}

extern "C" int cpb_compute_masks_host(const float* dP, const float* cellprob, const float* logits, int B, int H,
                                      int W, int C, const cpb_params* prm, int32_t* masks, int32_t* counts,
                                      int32_t* cell_class, uint8_t* class_masks, int tiles_per_chunk, int device) {
    cpb_host_options opt{};
    opt.tiles_per_chunk = tiles_per_chunk; opt.device = device; opt.logits_mode = CPB_HOST_LOGITS_AUTO;
    opt.flows_mode = CPB_HOST_LOGITS_AUTO; opt.masks_u16 = 0;
    return cpb_compute_masks_host_ex(dP, cellprob, logits, B, H, W, C, prm, masks, counts, cell_class, class_masks, &opt);
}
