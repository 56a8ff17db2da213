// Single-tile plans: what the numpy hooks call once per tile (the reference's WSI loop hands over one tile at a
// time from two inference threads: predict_wsi.py:728-797).  A plan owns pinned host staging, device buffers, a
// workspace, a non-blocking stream and CUDA graphs of
//     upload dP, cellprob -> fused path -> download masks, count                      (cpb_tile_plan_run)
//     upload logits -> class vote on the labels still on the device -> download classes  (cpb_tile_plan_vote)
// so that a call is one memcpy into the staging buffer, ONE graph launch and one synchronise instead of ~30 kernel
// launches, several allocations and pageable copies.  Everything is captured with cudaStreamCaptureModeThreadLocal
// on the plan's own stream: other host threads (the second inference thread, the network's allocator) are not
// affected by the capture and cannot invalidate it.  A plan is used by one host thread at a time.
#include <new>

struct cpb_tile_plan {
    int H = 0, W = 0, device = 0, LC = 0;
    cpb_params prm{};
    cudaStream_t st = nullptr;
    float* h_dP = nullptr; float* h_cp = nullptr; int32_t* h_masks = nullptr; int32_t* h_count = nullptr;
    float* d_dP = nullptr; float* d_cp = nullptr; int32_t* d_masks = nullptr; int32_t* d_count = nullptr;
    void* ws = nullptr; size_t ws_bytes = 0;
    cudaGraphExec_t run_exec = nullptr;
    // vote
    int C = 0;
    float* h_lg = nullptr; float* d_lg = nullptr; int32_t* d_cc = nullptr; int32_t* h_cc = nullptr;
    uint8_t* d_cm = nullptr; uint8_t* h_cm = nullptr; void* vws = nullptr; size_t vws_bytes = 0;
    cudaGraphExec_t vote_exec = nullptr;
};

namespace {

void plan_free_vote(cpb_tile_plan* p) {
    if (p->vote_exec) { cudaGraphExecDestroy(p->vote_exec); p->vote_exec = nullptr; }
    if (p->h_lg) cudaFreeHost(p->h_lg);
    if (p->h_cc) cudaFreeHost(p->h_cc);
    if (p->h_cm) cudaFreeHost(p->h_cm);
    if (p->d_lg) cudaFree(p->d_lg);
    if (p->d_cc) cudaFree(p->d_cc);
    if (p->d_cm) cudaFree(p->d_cm);
    if (p->vws) cudaFree(p->vws);
    p->h_lg = nullptr; p->h_cc = nullptr; p->h_cm = nullptr; p->d_lg = nullptr; p->d_cc = nullptr; p->d_cm = nullptr;
    p->vws = nullptr; p->C = 0;
}

int plan_enqueue_run(cpb_tile_plan* p) {
    const size_t N = (size_t)p->H * p->W;
    cudaMemcpyAsync(p->d_dP, p->h_dP, 2 * N * sizeof(float), cudaMemcpyHostToDevice, p->st);
    cudaMemcpyAsync(p->d_cp, p->h_cp, N * sizeof(float), cudaMemcpyHostToDevice, p->st);
    const int rc = compute_masks_impl(p->d_dP, p->d_cp, nullptr, 1, p->H, p->W, 0, &p->prm, p->d_masks, p->d_count, nullptr,
                                      nullptr, p->ws, p->ws_bytes, p->st, nullptr);
    if (rc) return rc;
    cudaMemcpyAsync(p->h_masks, p->d_masks, N * sizeof(int32_t), cudaMemcpyDeviceToHost, p->st);
    cudaMemcpyAsync(p->h_count, p->d_count, sizeof(int32_t), cudaMemcpyDeviceToHost, p->st);
    return 0;
}

int plan_enqueue_vote(cpb_tile_plan* p) {
    const size_t N = (size_t)p->H * p->W;
    cudaMemcpyAsync(p->d_lg, p->h_lg, (size_t)p->C * N * sizeof(float), cudaMemcpyHostToDevice, p->st);
    const int rc = cpb_class_vote_counts_device(p->d_masks, p->d_lg, p->d_count, 1, p->H, p->W, p->C, p->LC, p->d_cc, p->d_cm,
                                                p->vws, p->vws_bytes, p->st);
    if (rc) return rc;
    cudaMemcpyAsync(p->h_cm, p->d_cm, N, cudaMemcpyDeviceToHost, p->st);
    cudaMemcpyAsync(p->h_cc, p->d_cc, (size_t)std::min(p->LC, 1024) * sizeof(int32_t), cudaMemcpyDeviceToHost, p->st);
    return 0;
}

// warm-up run outside capture (lazy module loading, function attributes), then the same sequence as a graph
int plan_capture(cpb_tile_plan* p, int (*enqueue)(cpb_tile_plan*), cudaGraphExec_t* exec) {
    int rc = enqueue(p);
    if (rc) return rc;
    cudaError_t ce = cudaStreamSynchronize(p->st);
    if (ce != cudaSuccess) return (int)ce;
    if ((ce = cudaStreamBeginCapture(p->st, cudaStreamCaptureModeThreadLocal)) != cudaSuccess) return (int)ce;
    rc = enqueue(p);
    cudaGraph_t g = nullptr;
    ce = cudaStreamEndCapture(p->st, &g);
    if (rc) { if (g) cudaGraphDestroy(g); return rc; }
    if (ce != cudaSuccess) return (int)ce;
    ce = cudaGraphInstantiate(exec, g, 0);
    cudaGraphDestroy(g);
    return ce == cudaSuccess ? 0 : (int)ce;
}

}  // namespace

extern "C" {

int cpb_tile_plan_create(int H, int W, const cpb_params* prm, int device, cpb_tile_plan** out) {
    if (!prm || !out || check_geom(1, H, W)) return CPB_E_ARG;
    cudaError_t ce = cudaSetDevice(device);
    if (ce != cudaSuccess) return (int)ce;
    cpb_tile_plan* p = new (std::nothrow) cpb_tile_plan();
    if (!p) return CPB_E_ARG;
    p->H = H; p->W = W; p->device = device; p->prm = *prm; p->LC = cpb_label_capacity(H, W);
    const size_t N = (size_t)H * W;
    p->ws_bytes = cpb_workspace_bytes(1, H, W, 0, 0);
    bool ok = cudaStreamCreateWithFlags(&p->st, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaHostAlloc(&p->h_dP, 2 * N * sizeof(float), cudaHostAllocDefault) == cudaSuccess;
    ok = ok && cudaHostAlloc(&p->h_cp, N * sizeof(float), cudaHostAllocDefault) == cudaSuccess;
    ok = ok && cudaHostAlloc(&p->h_masks, N * sizeof(int32_t), cudaHostAllocDefault) == cudaSuccess;
    ok = ok && cudaHostAlloc(&p->h_count, 64, cudaHostAllocDefault) == cudaSuccess;
    ok = ok && cudaMalloc(&p->d_dP, 2 * N * sizeof(float)) == cudaSuccess;
    ok = ok && cudaMalloc(&p->d_cp, N * sizeof(float)) == cudaSuccess;
    ok = ok && cudaMalloc(&p->d_masks, N * sizeof(int32_t)) == cudaSuccess;
    ok = ok && cudaMalloc(&p->d_count, 64) == cudaSuccess;
    ok = ok && cudaMalloc(&p->ws, p->ws_bytes) == cudaSuccess;
    if (!ok) { const int e = (int)cudaGetLastError(); cpb_tile_plan_destroy(p); return e ? e : CPB_E_ARG; }
    memset(p->h_dP, 0, 2 * N * sizeof(float));
    for (size_t i = 0; i < N; i++) p->h_cp[i] = -1.f;            // the warm-up run sees an empty tile
    const int rc = plan_capture(p, plan_enqueue_run, &p->run_exec);
    if (rc) { cpb_tile_plan_destroy(p); return rc; }
    *out = p;
    return 0;
}

void cpb_tile_plan_destroy(cpb_tile_plan* p) {
    if (!p) return;
    cudaSetDevice(p->device);
    if (p->st) cudaStreamSynchronize(p->st);
    plan_free_vote(p);
    if (p->run_exec) cudaGraphExecDestroy(p->run_exec);
    if (p->h_dP) cudaFreeHost(p->h_dP);
    if (p->h_cp) cudaFreeHost(p->h_cp);
    if (p->h_masks) cudaFreeHost(p->h_masks);
    if (p->h_count) cudaFreeHost(p->h_count);
    if (p->d_dP) cudaFree(p->d_dP);
    if (p->d_cp) cudaFree(p->d_cp);
    if (p->d_masks) cudaFree(p->d_masks);
    if (p->d_count) cudaFree(p->d_count);
    if (p->ws) cudaFree(p->ws);
    if (p->st) cudaStreamDestroy(p->st);
    delete p;
}

float* cpb_tile_plan_dp(cpb_tile_plan* p) { return p ? p->h_dP : nullptr; }
float* cpb_tile_plan_cellprob(cpb_tile_plan* p) { return p ? p->h_cp : nullptr; }
const int32_t* cpb_tile_plan_masks(cpb_tile_plan* p) { return p ? p->h_masks : nullptr; }

int cpb_tile_plan_run(cpb_tile_plan* p, int32_t* count) {
    if (!p || !p->run_exec) return CPB_E_ARG;
    cudaError_t ce = cudaGraphLaunch(p->run_exec, p->st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(p->st);
    if (ce != cudaSuccess) return (int)ce;
    if (count) *count = *p->h_count;
    return *p->h_count < 0 ? CPB_E_CAPACITY : 0;
}

float* cpb_tile_plan_logits(cpb_tile_plan* p, int C) {
    if (!p || C < 1 || C > 255) return nullptr;
    if (p->C == C) return p->h_lg;
    cudaSetDevice(p->device);
    cudaStreamSynchronize(p->st);
    plan_free_vote(p);
    const size_t N = (size_t)p->H * p->W;
    p->C = C;
    p->vws_bytes = cpb_workspace_bytes(1, p->H, p->W, C, 0);
    bool ok = cudaHostAlloc(&p->h_lg, (size_t)C * N * sizeof(float), cudaHostAllocDefault) == cudaSuccess;
    ok = ok && cudaHostAlloc(&p->h_cm, N, cudaHostAllocDefault) == cudaSuccess;
    ok = ok && cudaHostAlloc(&p->h_cc, (size_t)std::min(p->LC, 1024) * sizeof(int32_t), cudaHostAllocDefault) == cudaSuccess;
    ok = ok && cudaMalloc(&p->d_lg, (size_t)C * N * sizeof(float)) == cudaSuccess;
    ok = ok && cudaMalloc(&p->d_cc, (size_t)p->LC * sizeof(int32_t)) == cudaSuccess;
    ok = ok && cudaMalloc(&p->d_cm, N) == cudaSuccess;
    ok = ok && cudaMalloc(&p->vws, p->vws_bytes) == cudaSuccess;
    if (ok) { memset(p->h_lg, 0, (size_t)C * N * sizeof(float)); cudaMemsetAsync(p->d_cc, 0, (size_t)p->LC * sizeof(int32_t), p->st); }
    if (!ok || plan_capture(p, plan_enqueue_vote, &p->vote_exec)) { cudaGetLastError(); plan_free_vote(p); return nullptr; }
    return p->h_lg;
}

int cpb_tile_plan_vote(cpb_tile_plan* p, const uint8_t** class_masks, const int32_t** cell_class) {
    if (!p || !p->vote_exec) return CPB_E_ARG;
    cudaError_t ce = cudaGraphLaunch(p->vote_exec, p->st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(p->st);
    if (ce != cudaSuccess) return (int)ce;
    if (class_masks) *class_masks = p->h_cm;
    if (cell_class) *cell_class = p->h_cc;
    return 0;
}

}  // extern "C"
