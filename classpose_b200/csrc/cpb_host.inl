// Host-buffer entry point: H2D -> fused path -> D2H, chunked over two streams so the copies
// of one chunk overlap the kernels of the other.  Device buffers are cached per host thread
// (the reference calls in from >= 2 threads per process; nothing is shared between them).
#include <algorithm>
#include <vector>

namespace {

struct HostSlot {
    cudaStream_t stream = nullptr;
    char* blob = nullptr;
    size_t cap = 0;
};

struct HostCtx {
    int device = -1;
    HostSlot slot[2];
    ~HostCtx() {
        for (auto& s : slot) {
            if (s.blob) cudaFree(s.blob);
            if (s.stream) cudaStreamDestroy(s.stream);
        }
    }
};

thread_local HostCtx tl_ctx;

struct ChunkBufs {
    float* dP; float* cellprob; float* logits; int32_t* masks; int32_t* counts; int32_t* cell_class;
    uint8_t* class_masks; void* ws; size_t ws_bytes; size_t total;
};

ChunkBufs carve_chunk(char* base, int Bc, int H, int W, int C, bool has_logits, bool has_cm) {
    const size_t N = (size_t)H * W;
    const int LC = cpb_label_capacity(H, W);
    Carver c{base, 0};
    ChunkBufs b{};
    b.dP = c.take<float>(2 * N * Bc);
    b.cellprob = c.take<float>(N * Bc);
    b.logits = c.take<float>(has_logits ? (size_t)C * N * Bc : 0);
    b.masks = c.take<int32_t>(N * Bc);
    b.counts = c.take<int32_t>(Bc);
    b.cell_class = c.take<int32_t>(has_logits ? (size_t)LC * Bc : 0);
    b.class_masks = c.take<uint8_t>(has_cm ? N * Bc : 0);
    b.ws_bytes = cpb_workspace_bytes(Bc, H, W, has_logits ? C : 0, 0);
    b.ws = c.take<char>(b.ws_bytes);
    b.total = c.off;
    return b;
}

}  // namespace

extern "C" int cpb_compute_masks_host(const float* dP, const float* cellprob, const float* logits, int B, int H,
                                      int W, int C, const cpb_params* prm, int32_t* masks, int32_t* counts,
                                      int32_t* cell_class, uint8_t* class_masks, int tiles_per_chunk, int device) {
    if (!dP || !cellprob || !prm || !masks || !counts || B <= 0 || H < 2 || W < 2) return CPB_E_ARG;
    if (logits && (!cell_class || C < 1)) return CPB_E_ARG;
    cudaError_t ce = cudaSetDevice(device);
    if (ce != cudaSuccess) return (int)ce;
    const size_t N = (size_t)H * W;
    const int LC = cpb_label_capacity(H, W);
    int Bc = tiles_per_chunk > 0 ? tiles_per_chunk : 128;
    Bc = std::min(Bc, B);
    while ((long long)Bc * H * W >= (1LL << 31)) Bc /= 2;
    if (Bc < 1) return CPB_E_RANGE;
    const bool has_logits = logits != nullptr, has_cm = has_logits && class_masks != nullptr;
    const size_t need = carve_chunk(nullptr, Bc, H, W, C, has_logits, has_cm).total + kAlign;

    HostCtx& ctx = tl_ctx;
    if (ctx.device != device) {
        for (auto& s : ctx.slot) {
            if (s.blob) { cudaFree(s.blob); s.blob = nullptr; s.cap = 0; }
            if (s.stream) { cudaStreamDestroy(s.stream); s.stream = nullptr; }
        }
        ctx.device = device;
    }
    for (auto& s : ctx.slot) {
        if (!s.stream && (ce = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking)) != cudaSuccess) return (int)ce;
        if (s.cap < need) {
            if (s.blob) cudaFree(s.blob);
            s.blob = nullptr; s.cap = 0;
            if ((ce = cudaMalloc(&s.blob, need)) != cudaSuccess) return (int)ce;
            s.cap = need;
        }
    }
    int rc = 0;
    for (int b0 = 0, k = 0; b0 < B; b0 += Bc, k++) {
        const int nb = std::min(Bc, B - b0);
        HostSlot& s = ctx.slot[k & 1];
        ChunkBufs cb = carve_chunk(s.blob, Bc, H, W, C, has_logits, has_cm);
        cudaStream_t st = s.stream;
        cudaMemcpyAsync(cb.dP, dP + (size_t)b0 * 2 * N, (size_t)nb * 2 * N * sizeof(float), cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(cb.cellprob, cellprob + (size_t)b0 * N, (size_t)nb * N * sizeof(float), cudaMemcpyHostToDevice, st);
        if (has_logits)
            cudaMemcpyAsync(cb.logits, logits + (size_t)b0 * C * N, (size_t)nb * C * N * sizeof(float), cudaMemcpyHostToDevice, st);
        rc = cpb_compute_masks_device(cb.dP, cb.cellprob, has_logits ? cb.logits : nullptr, nb, H, W, C, prm, cb.masks,
                                      cb.counts, has_logits ? cb.cell_class : nullptr, has_cm ? cb.class_masks : nullptr,
                                      cb.ws, cb.ws_bytes, st);
        if (rc) break;
        cudaMemcpyAsync(masks + (size_t)b0 * N, cb.masks, (size_t)nb * N * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(counts + b0, cb.counts, (size_t)nb * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
        if (has_logits)
            cudaMemcpyAsync(cell_class + (size_t)b0 * LC, cb.cell_class, (size_t)nb * LC * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
        if (has_cm)
            cudaMemcpyAsync(class_masks + (size_t)b0 * N, cb.class_masks, (size_t)nb * N, cudaMemcpyDeviceToHost, st);
    }
    for (auto& s : ctx.slot) {
        ce = cudaStreamSynchronize(s.stream);
        if (ce != cudaSuccess && rc == 0) rc = (int)ce;
    }
    if (rc == 0)
        for (int b = 0; b < B; b++) if (counts[b] < 0) { rc = CPB_E_CAPACITY; break; }
    return rc;
}
