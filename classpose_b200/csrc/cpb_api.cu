// C ABI of libclasspose_b200.so (see include/classpose_b200.h).  Host side only: argument
// checks, workspace carving, kernel launch sequences.  No host synchronisation, no global
// mutable state; every call runs on the caller's stream with the caller's workspace.
#include "classpose_b200.h"

#define CPB_QCTR_INTS 24
#include "cpb_platform.h"
#include "cpb_common.cuh"
#include "cpb_flow.cuh"
#include "cpb_masks.cuh"
#include "cpb_tables.cuh"
#include "cpb_qc.cuh"
#include "cpb_qc32.cuh"
#include "cpb_post.cuh"
#include "cpb_fused.cuh"
#include "cpb_contour.cuh"
#include "cpb_dedup.cuh"
#include "cpb_prep.cuh"

#include <atomic>
#include <cstdlib>
#ifndef CPB_SIM
#include <mutex>
#endif

static_assert(sizeof(cpb_params) == 40, "cpb_params layout is part of the ABI (python mirror: _abi.Params)");

namespace {

std::atomic<long long> g_launches{0};   // statistics only: kernels launched by this library
std::atomic<int> g_last_qc[CPB_QCTR_INTS];   // statistics only: flow-check counters of the last profiled call
#define CPB_LAUNCH_COUNTED(...) do { g_launches.fetch_add(1, std::memory_order_relaxed); CPB_LAUNCH(__VA_ARGS__); } while (0)

// Optional per-stage timing of the fused path (cpb_compute_masks_profiled_device).
enum Stage { S_PREP = 0, S_FOLLOW, S_SEEDS, S_LOOKUP, S_FINALIZE, S_MAP1, S_CENTRES, S_DIFFUSE, S_FLOWERR,
             S_DROP, S_SIZE1, S_MAP2, S_FILL, S_MAP3, S_SIZE2, S_MAP4, S_BORDER, S_VOTE, S_COUNT };
// (S_MAP1, S_DROP, S_MAP2 are only used by the stage-by-stage entry points; the fused path has no such passes)
const char* kStageNames[S_COUNT] = {"prep_flow", "follow_flows", "seeds", "lookup", "gm_finalize", "map_stats_1",
                                    "centres", "diffuse", "flow_err", "drop_bad_stats", "size_filter_1",
                                    "map_stats_2", "fill_holes", "recount_hole_tiles", "size_filter_2", "final_map",
                                    "border", "vote"};
struct Prof {
#ifndef CPB_SIM
    cudaEvent_t begin[S_COUNT], end[S_COUNT];
#endif
    bool used[S_COUNT];
    cudaStream_t st;
};
inline void prof_begin(Prof* p, int s) {
#ifndef CPB_SIM
    if (p) { cudaEventRecord(p->begin[s], p->st); p->used[s] = true; }
#endif
}
inline void prof_end(Prof* p, int s) {
#ifndef CPB_SIM
    if (p) cudaEventRecord(p->end[s], p->st);
#endif
}
struct ProfScope {
    Prof* p; int s;
    ProfScope(Prof* p_, int s_) : p(p_), s(s_) { prof_begin(p, s); }
    ~ProfScope() { prof_end(p, s); }
};

constexpr int kLabelBlocksPerTile = 24;        // block-per-label kernels of the stage calls: grid (kLabelBlocksPerTile, B)
constexpr int kWarpDiffuseBlocksPerTile = 16;  // warp-per-label kernels: grid (16, B) x 4 warps = 64 warp slots per tile
constexpr int kVoteSmemInts = 6 * 1024;   // 24 KB (instance, class) table per tile: 4 blocks of 512 threads per SM;
                                          // tiles with more than 6144/C labels use the global table
constexpr int kVoteSmemIntsMax = 24 * 1024; // 96 KB
constexpr size_t kAlign = 256;

inline size_t align_up(size_t v) { return (v + kAlign - 1) / kAlign * kAlign; }

struct Carver {
    char* base; size_t off;
    template <class T> T* take(size_t n) {
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off = align_up(off + n * sizeof(T));
        return p;
    }
};

struct Workspace {
    float2* flow;        // [B*N]
    int* pfinal;         // [B*N]
    int* hist;           // [B*N]
    int* M;              // [B*N]
    double* T;           // [B*N]
    double* T2;          // [B*N]   (aliases holekey)
    u64* holekey;        // [B*N]
    unsigned* list;      // [B*N]
    unsigned* list_n;    // [64]: [0] foreground count, [2..3] bump cursor of the hole-fill bitmap pool (u64)
    u64* skey;           // [B*LC]
    int* sidx;           // [B*LC]
    int* sinv;           // [B*LC]
    int* alive;          // [B*LC]
    int* vote;           // [B*LC*C]
    int* jobs;           // [B+1] diffusion job offsets, 2 queue counters, 2 counts of `todo` (front / back)
    int2* todo;          // [B*LC] (tile, label) work list of the block-per-label kernels
    Q32 q;               // float32 flow-check screen: label info, class lists, job queue, float64 list, counters
    int* cover;          // [B*LC] exact hole fill (CPB_FILL_EXACT): pixels a label loses to other labels' fills
    int* seq;            // [B]    tiles whose fill is replayed sequentially
    LabelTables t;
    size_t bytes;
    Prof* prof;          // optional stage timing
};

Workspace carve(void* base, int B, int H, int W, int C, int lcap) {
    Workspace w{};
    const size_t N = (size_t)H * W, BN = (size_t)B * N;
    const int LC = lcap > 0 ? lcap : cpb_label_capacity(H, W);
    const size_t BL = (size_t)B * LC;
    Carver c{reinterpret_cast<char*>(base), 0};
    w.flow = c.take<float2>((size_t)B * (H + 2) * (W + 2 * CPB_FLOW_PADX));
    w.pfinal = c.take<int>(BN);
    w.hist = c.take<int>(BN);
    w.M = c.take<int>(BN);
    w.T = c.take<double>(BN);
    w.T2 = c.take<double>(BN);
    w.holekey = reinterpret_cast<u64*>(w.T2);
    w.list = c.take<unsigned>(BN);
    w.list_n = c.take<unsigned>(64);
    w.skey = c.take<u64>(BL);
    w.sidx = c.take<int>(BL);
    w.sinv = c.take<int>(BL);
    w.alive = c.take<int>(BL);
    w.vote = c.take<int>(C > 0 ? BL * C : 0);
    w.jobs = c.take<int>((size_t)B + 1 + 4);
    w.todo = c.take<int2>(BL);
    w.q.info = c.take<int>(BL); w.q.ent = c.take<int>(BL); w.q.jobs = c.take<int4>(BL); w.q.sorted = c.take<int4>(BL); w.q.l64 = c.take<int2>(BL); w.q.lc = c.take<int2>(BL);
    w.q.cls_cnt = c.take<int>((size_t)B * CPB_Q32_NCLS); w.q.T32 = reinterpret_cast<float*>(w.M);
    w.q.ctr = c.take<int>(CPB_QCTR_INTS);
    LabelTables& t = w.t;
    t.LC = LC;
    t.cnt = c.take<int>(BL); t.first = c.take<int>(BL);
    t.ymin = c.take<int>(BL); t.ymax = c.take<int>(BL); t.xmin = c.take<int>(BL); t.xmax = c.take<int>(BL);
    t.sumy = c.take<u64>(BL); t.sumx = c.take<u64>(BL);
    t.remap = c.take<int>(BL); t.flag = c.take<int>(BL); t.alive = nullptr;
    t.cy = c.take<int>(BL); t.cx = c.take<int>(BL);
    t.err = c.take<double>(BL);
    t.done = c.take<int>(BL);
    t.lbound = c.take<int>(B); t.nlab = c.take<int>(B); t.niter = c.take<int>(B); t.misc = c.take<int>(B);
    t.fail = c.take<int>(B);
    w.cover = c.take<int>(BL); w.seq = c.take<int>(B);
    w.bytes = c.off;
    return w;
}

int check_geom(int B, int H, int W) {
    if (B <= 0 || H < 2 || W < 2 || H > 32767 || W > 32767) return CPB_E_ARG;
    if ((long long)B * H * W >= (1LL << 31)) return CPB_E_RANGE;
    return 0;
}

// bitmap pool of the block hole-fill kernel for crops beyond its shared memory: the float64 T plane, free once
// the flow check is over (2 words per pixel of the batch); cursor in the zeroed scratch words after list_n
inline FillPool fill_pool(const Workspace& w, int B, int H, int W) {
    return FillPool{reinterpret_cast<unsigned*>(w.T), reinterpret_cast<u64*>(w.list_n + 2), 2ull * (u64)B * H * W};
}

inline unsigned blocks_for(long long n, int per) { return (unsigned)((n + per - 1) / per); }

// blocks per tile of the pixel passes that only run on flagged tiles (hole plane reset, recount): 8 for a 256 x 256
// tile, more for larger tiles (a dense 512 x 512 batch has a hole in every tile)
inline unsigned tile_slices(int H, int W) { return (unsigned)std::min<long long>(std::max<long long>((long long)H * W / 8192, 8), 256); }

// per-tile table kernels (seed ranking, renumbering): one block per tile; big tiles can hold thousands of labels
inline unsigned table_threads(int H, int W) { return (long long)H * W > 256 * 256 ? 1024u : 256u; }

#ifndef CPB_SIM
// Function attributes and the SM count belong to a DEVICE (context), and one process may drive several devices
// (cpb_compute_masks_host takes a device ordinal; the python engine caches one Engine per device): both are
// tracked per device ordinal.
constexpr int kMaxDevices = 64;
std::atomic<int> g_attr_done[kMaxDevices];
std::atomic<int> g_sm_count[kMaxDevices];
int current_device() { int dev = 0; cudaGetDevice(&dev); return (dev >= 0 && dev < kMaxDevices) ? dev : 0; }
void ensure_attributes() {
    const int dev = current_device();
    if (g_attr_done[dev].load(std::memory_order_acquire)) return;
    cudaFuncSetAttribute(k_fill_holes, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * CPB_FILL_WORDS * 4);
    cudaFuncSetAttribute(k_vote, cudaFuncAttributeMaxDynamicSharedMemorySize, kVoteSmemIntsMax * 4);
    cudaFuncSetAttribute(k_follow_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, CPB_FS_WIN * CPB_FS_WIN * 8);
    g_attr_done[dev].store(1, std::memory_order_release);      // (setting them twice from two threads is harmless)
}
// Stream-ordered scratch (the blend's weight tables) comes from a pool of the library's own that keeps what is freed:
// the device's default pool hands its memory back at every synchronisation, and the next call then pays a real
// allocation (measured: the 9-way blend pair took 1.1 to 1.9 ms depending on the box).
cudaMemPool_t g_pool[kMaxDevices] = {};
std::mutex g_pool_mutex;
cudaError_t scratch_alloc(void** p, size_t bytes, cudaStream_t st) {
    const int dev = current_device();
    cudaMemPool_t pool;
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        if (!g_pool[dev]) {
            cudaMemPoolProps props = {};
            props.allocType = cudaMemAllocationTypePinned;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            cudaError_t e = cudaMemPoolCreate(&g_pool[dev], &props);
            if (e != cudaSuccess) { g_pool[dev] = nullptr; return e; }
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(g_pool[dev], cudaMemPoolAttrReleaseThreshold, &keep);
        }
        pool = g_pool[dev];
    }
    return cudaMallocFromPoolAsync(p, bytes, pool, st);
}
int sm_count() {
    const int dev = current_device();
    int n = g_sm_count[dev].load(std::memory_order_relaxed);
    if (n <= 0) {
        int v = 148;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        n = v > 0 ? v : 148;
        g_sm_count[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}
#else
void ensure_attributes() {}
int sm_count() { return 4; }
cudaError_t scratch_alloc(void** p, size_t bytes, cudaStream_t st) { return cudaMallocAsync(p, bytes, st); }
#endif

// follow_flows variant: 2 = trajectory pool with many merge points (default), 1 = two merge points per 256-pixel
// chunk, 0 = plain kernel, 3 = plain kernel with the flow window of every 32 x 32 patch staged by TMA bulk copies
// (the north-star experiment, needs cellprob: fused path and cpb_follow_flows_device); CPB_FOLLOW_MERGE in the
// environment or cpb_debug_set_follow_merge (A/B measurements and tests).  Results are bit-identical in every mode.
std::atomic<int> g_follow_merge{-1};     // -1: take CPB_FOLLOW_MERGE from the environment
int follow_merge_mode() {
    const int v = g_follow_merge.load(std::memory_order_relaxed);
    if (v >= 0) return v;
    static const int env = [] { const char* e = getenv("CPB_FOLLOW_MERGE"); return (e && e[0] >= '0' && e[0] <= '3') ? e[0] - '0' : 2; }();
    return env;
}

// merge points of k_follow_pool for `niter` Euler steps.  A merge costs about as much as three Euler steps of the
// whole chunk, so a few points where the live count falls fastest beat many (measured on B200, ms per 1024 conic tiles:
// 16 points 3.19, 11 points 2.89, 6 points 2.43, 4 points 2.31, 3 points 2.32); CPB_FOLLOW_SCHEDULE="a,b,c" (step numbers) overrides for experiments
FollowSchedule follow_schedule(int niter) {
    static const int kPer200[] = {36, 56, 88, 136};
    FollowSchedule s{};
    static const char* env = getenv("CPB_FOLLOW_SCHEDULE");
    int last = 0;
    if (env && env[0]) {
        const char* p = env;
        while (*p && s.n < CPB_FP_MAXMERGE) {
            char* q = nullptr;
            const long v = strtol(p, &q, 10);
            if (q == p) break;
            if (v > last && v < niter) { s.at[s.n++] = (int)v; last = (int)v; }
            p = (*q == ',') ? q + 1 : q;
        }
        return s;
    }
    for (int v : kPer200) {
        const int a = (int)((long long)v * niter / 200);
        if (a > last && a < niter && s.n < CPB_FP_MAXMERGE) { s.at[s.n++] = a; last = a; }
    }
    return s;
}

// A/B switches (environment at first use, or cpb_debug_set_switch): every setting gives the same results.
//   CPB_DIFFUSE_QUEUE=0  static (block, warp) -> label map instead of the job queue
//   CPB_QC_FUSED=0       every label's flow error from T in global memory (k_flow_err) instead of the diffusion tile
//   CPB_VOTE_FUSED=0     class vote as its own pass over the finished label image
//   CPB_QC_SCREEN=0      every label through the float64 diffusion (no float32 screen in front of it)
//   CPB_FILL_EXACT=1     (default 0) replay upstream's label-by-label hole fill on tiles in which a label lies partly inside
//                        another label's hole (the one configuration where taking every label's holes from the input
//                        image differs from it); simulator-validated only
//   CPB_SEED_CANDS=0     seed candidates (bins with more than 10 end points) found by streaming the histogram (k_seed_scan)
//                        instead of being listed by the kernel that counts the end points
//   CPB_FOLLOW_SMALL=0   1024-entry chunks in the trajectory pool for every batch size (1: 256-entry chunks for a handful of tiles)
//   CPB_BLEND_EFT=0      blend with float64 arithmetic per element (numpy's literal op sequence) instead of the float32
//                        error-free form (identical up to ~1e-6 of the elements by one ulp)
std::atomic<int> g_switch[9] = {{-1}, {-1}, {-1}, {-1}, {-1}, {-1}, {-1}, {-1}, {-1}};
bool switch_on(int which, const char* env_name) {
    int v = g_switch[which].load(std::memory_order_relaxed);
    if (v < 0) {
        const char* e = getenv(env_name);
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}
bool diffuse_queue_enabled() { return switch_on(CPB_SWITCH_DIFFUSE_QUEUE, "CPB_DIFFUSE_QUEUE"); }
bool qc_fused_enabled() { return switch_on(CPB_SWITCH_QC_FUSED, "CPB_QC_FUSED"); }
bool vote_fused_enabled() { return switch_on(CPB_SWITCH_VOTE_FUSED, "CPB_VOTE_FUSED"); }
bool qc_screen_enabled() { return switch_on(CPB_SWITCH_QC_SCREEN, "CPB_QC_SCREEN"); }
bool blend_eft_enabled() { return switch_on(CPB_SWITCH_BLEND_EFT, "CPB_BLEND_EFT"); }
// default OFF (the only switch that is): the exact replay for tangled labels has run on the simulator only
bool fill_exact_enabled() {
    int v = g_switch[CPB_SWITCH_FILL_EXACT].load(std::memory_order_relaxed);
    if (v < 0) { const char* e = getenv("CPB_FILL_EXACT"); v = (e && e[0] == '1') ? 1 : 0; }
    return v != 0;
}
bool seed_cands_enabled() { return switch_on(CPB_SWITCH_SEED_CANDS, "CPB_SEED_CANDS"); }
bool follow_small_enabled() { return switch_on(CPB_SWITCH_FOLLOW_SMALL, "CPB_FOLLOW_SMALL"); }
// value 2 (tests only): the screen also runs when the caller asks for the per-label errors, and reports
// (float32 error, bound) bit-packed into the float64 error of the labels it decided
bool qc_screen_debug() { return g_switch[CPB_SWITCH_QC_SCREEN].load(std::memory_order_relaxed) == 2; }

#define CPB_CHECK_LAUNCH() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return (int)e_; } while (0)

// The fused path over a large batch is cut into parts that run on parallel streams (cpb_compute_masks_device): kernels
// of different parts then share the SMs -- the latency-bound stages of one part (label look-up, label scan, seeds)
// fill issue slots the issue-bound Euler kernel of another leaves, and no kernel's tail wave leaves the GPU idle.
// Measured on the B200, 1024 conic tiles: 1 part 4.36 ms, 2 parts 4.20, 4 parts 4.12, 8 parts 4.16
// (profiles/r02/ab_two_streams.txt).  CPB_BATCH_PARTS overrides (1 = off).
constexpr int kMaxParts = 4;
int batch_parts(int B) {
    static const int env = [] { const char* e = getenv("CPB_BATCH_PARTS"); return e ? atoi(e) : 0; }();
    int np = env > 0 ? env : B / 128;        // 256 tiles: 2 parts (+6.7 %), 512 and 1024: 4 parts (+4.3 % / +6.0 %)
    return std::max(1, std::min(np, std::min(kMaxParts, B)));
}
inline int part_begin(int B, int np, int k) { return (int)((long long)B * k / np); }

// ---- stage launch sequences (all asynchronous on `st`) ---------------------------------------

int run_init_tables(const Workspace& w, int B, cudaStream_t st) {
    CPB_LAUNCH_COUNTED(k_init_tables, dim3(blocks_for(w.t.LC, 256), B), dim3(256), 0, st, w.t);
    CPB_CHECK_LAUNCH();
    return 0;
}

int run_set_lbound(const Workspace& w, int B, int v, cudaStream_t st) {
    CPB_LAUNCH_COUNTED(k_fill_i32, dim3(blocks_for(B, 256)), dim3(256), 0, st, w.t.lbound, B, v);
    CPB_CHECK_LAUNCH();
    return 0;
}

// map labels (remap / drop flags / hole keys) and optionally regather statistics
int run_map_stats(const Workspace& w, int32_t* lab, int B, int H, int W, int nch, const int* map,
                  const int* drop, const u64* holekey, bool stats, cudaStream_t st, int stage = -1) {
    ProfScope ps(stage >= 0 ? w.prof : nullptr, stage >= 0 ? stage : 0);
    if (stats) { int e = run_init_tables(w, B, st); if (e) return e; }
    CPB_LAUNCH_COUNTED(k_map_stats, dim3(blocks_for((long long)B * H * W, 256)), dim3(256), 0, st, lab, B, H, W, nch,
               map, drop, holekey, stats ? 1 : 0, w.t);
    CPB_CHECK_LAUNCH();
    return 0;
}

// zero_out != NULL (fused path): the prep kernel zeroes that label image instead of marking background in p_final
int run_follow(const Workspace& w, const float* dP, const float* cellprob, int B, int H, int W, int niter,
               float thr, int32_t* pfinal, float* pfloat, int* hist, cudaStream_t st, int32_t* zero_out = nullptr,
               float* dP_copy = nullptr, bool prep_done = false, bool seed_cands = false) {
    const long long BN = (long long)B * H * W;
    prof_begin(w.prof, S_PREP);
    if (!prep_done) {
        cudaMemsetAsync(w.list_n, 0, 64 * sizeof(unsigned), st);
        cudaMemsetAsync(w.t.fail, 0, B * sizeof(int), st);
    }
    if (hist) cudaMemsetAsync(hist, 0, BN * sizeof(int), st);
    // seed candidates (pixels whose bin passes 10 end points) are listed by the counting kernel: see cpb_hist_count
    SeedCands cands{nullptr, nullptr, 0};
    if (hist && seed_cands) {
        cands = SeedCands{w.skey, w.t.misc, w.t.LC};
        cudaMemsetAsync(w.t.misc, 0, B * sizeof(int), st);
    }
    const float sx = (float)(2.0 / (double)(W - 1)), sy = (float)(2.0 / (double)(H - 1));
    // tap indices are formed in float32 (exact below 2^24): one tile of more than ~4090 x 4090 pixels is out of range
    if ((long long)(H + 2) * (W + 2 * CPB_FLOW_PADX) >= (1LL << 24)) return CPB_E_RANGE;
    int32_t* bg_out = zero_out ? zero_out : pfinal;
    const int bg_value = zero_out ? 0 : -1;
    const bool vec4 = (W % 4 == 0) && (reinterpret_cast<uintptr_t>(dP) % 16 == 0) &&
                      (reinterpret_cast<uintptr_t>(cellprob) % 16 == 0) && (reinterpret_cast<uintptr_t>(bg_out) % 16 == 0);
    if (dP_copy && !(vec4 && reinterpret_cast<uintptr_t>(dP_copy) % 16 == 0)) return CPB_E_ARG;   // (only the host path asks)
    if (prep_done) {
        // (k_blend_prep already wrote the flow field, the foreground list and the zeroed label image)
    } else if (vec4) {
        const int patch = (W % 64 == 0) ? 1 : 0;
        const long long nblk = patch ? (long long)B * ((H + 2 + 15) / 16) * (W / 64)
                                     : (long long)blocks_for((long long)B * (H + 2) * (W / 4), 256);
        CPB_LAUNCH_COUNTED(k_prep_flow_v4, dim3((unsigned)nblk), dim3(256), 0, st,
                           reinterpret_cast<const float4*>(dP), reinterpret_cast<const float4*>(cellprob), B, H, W, thr, sx,
                           sy, reinterpret_cast<float4*>(w.flow), reinterpret_cast<int4*>(bg_out), w.list, w.list_n, patch, bg_value,
                           reinterpret_cast<float4*>(dP_copy));
    } else {
        CPB_LAUNCH_COUNTED(k_prep_flow, dim3(blocks_for((long long)B * (H + 2) * (W + 2 * CPB_FLOW_PADX), 256)), dim3(256),
                           0, st, dP, cellprob, B, H, W, thr, sx, sy, w.flow, bg_out, w.list, w.list_n, bg_value);
    }
    CPB_CHECK_LAUNCH();
    prof_end(w.prof, S_PREP);
    ProfScope ps(w.prof, S_FOLLOW);
    const unsigned grid = (unsigned)std::min<long long>(blocks_for(BN, 256), (long long)sm_count() * 16);
    const int mode = follow_merge_mode();
    // the pool kernel forms the tap index with an FADD (1.5 * 2^23 + index): the padded tile must stay below 2^22
    // pixels (about 2040 x 2040); larger tiles take the two-point merge kernel, whose index is a float -> int conversion
    const bool pool_ok = (long long)(H + 2) * (W + 2 * CPB_FLOW_PADX) < (1LL << 22);
#ifndef CPB_SIM
    if (mode == 3) {
        ensure_attributes();
        const int pb = ((W + CPB_FS_PATCH - 1) / CPB_FS_PATCH) * ((H + CPB_FS_PATCH - 1) / CPB_FS_PATCH);
        CPB_LAUNCH_COUNTED(k_follow_staged, dim3((unsigned)((long long)B * pb)), dim3(256), CPB_FS_WIN * CPB_FS_WIN * 8, st, w.flow,
                           cellprob, thr, B, H, W, niter, pfinal, pfloat, hist, cands);
        CPB_CHECK_LAUNCH();
        return 0;
    }
#endif
    if (mode == 2 && niter >= 32 && pool_ok) {
        // one block per chunk of the list (blocks past the end of the list exit at once).  A handful of tiles (the
        // numpy hooks hand over one) cannot fill the GPU with 1024-entry chunks: 256-entry chunks give four times
        // the blocks and one trajectory per thread, i.e. a quarter of the serial Euler steps per thread
        const bool small = BN <= (long long)CPB_FP_POOL * 4 * sm_count() && follow_small_enabled();
        const int pool = small ? CPB_FP_POOL_SMALL : CPB_FP_POOL;
#ifdef CPB_SIM
        const unsigned pgrid = (unsigned)std::min<long long>(blocks_for(BN, pool), 8);
#else
        const unsigned pgrid = blocks_for(BN, pool);
#endif
#define CPB_LAUNCH_POOL(WPv, POOLv) CPB_LAUNCH_COUNTED((k_follow_pool<WPv, POOLv>), dim3(pgrid), dim3(CPB_FP_THREADS), 0, st, w.flow, \
                               w.list, w.list_n, H, W, niter, follow_schedule(niter), pfinal, pfloat, hist, cands)
        if (W == 256) {       // the WSI tile width: row pitch as an immediate
            if (small) CPB_LAUNCH_POOL(256 + 2 * CPB_FLOW_PADX, CPB_FP_POOL_SMALL); else CPB_LAUNCH_POOL(256 + 2 * CPB_FLOW_PADX, CPB_FP_POOL);
        } else {
            if (small) CPB_LAUNCH_POOL(0, CPB_FP_POOL_SMALL); else CPB_LAUNCH_POOL(0, CPB_FP_POOL);
        }
#undef CPB_LAUNCH_POOL
    } else if (mode >= 1 && niter >= 32 && B < (1 << 28)) {
        // merge points at a quarter and a half of the integration (48 and 96 of 200 steps)
        CPB_LAUNCH_COUNTED(k_follow_merge, dim3(grid), dim3(CPB_FM_THREADS), 0, st, w.flow, w.list, w.list_n, H, W, niter,
                           (niter * 6) / 25, (niter * 12) / 25, pfinal, pfloat, hist, cands);
    } else {
        CPB_LAUNCH_COUNTED(k_follow, dim3(grid), dim3(256), 0, st, w.flow, w.list, w.list_n, H, W, niter, pfinal, pfloat,
                           hist, cands);
    }
    CPB_CHECK_LAUNCH();
    return 0;
}

// end-point histogram in w.hist -> seed labels painted into it, seed count per tile in t.lbound
int run_seeds(const Workspace& w, int B, int H, int W, cudaStream_t st, bool have_cands = false) {
    const long long BN = (long long)B * H * W;
    const int vec = ((long long)H * W % 4 == 0) && (reinterpret_cast<uintptr_t>(w.hist) % 16 == 0) ? 1 : 0;
    if (!have_cands) {      // (the fused path's counting kernel has listed the candidates already: cpb_hist_count)
        cudaMemsetAsync(w.t.misc, 0, B * sizeof(int), st);                 // candidate counters
        CPB_LAUNCH_COUNTED(k_seed_scan, dim3(blocks_for(vec ? BN / 4 : BN, 256)), dim3(256), 0, st, (const int*)w.hist, B, H, W,
                           w.t.LC, vec, w.skey, w.t.misc);
        CPB_CHECK_LAUNCH();
    }
    CPB_LAUNCH_COUNTED(k_seeds, dim3(B), dim3(table_threads(H, W)), 0, st, w.hist, H, W, w.t.LC, w.skey, w.sidx, w.t.lbound,
                       (const int*)w.t.misc);
    CPB_CHECK_LAUNCH();
    return 0;
}

// end points (+ histogram already in w.hist) -> contiguous labels in `masks`
int run_get_masks(const Workspace& w, const int32_t* pfinal, int B, int H, int W, double msf, int32_t* masks,
                  int32_t* counts, cudaStream_t st) {
    const long long BN = (long long)B * H * W;
    prof_begin(w.prof, S_SEEDS);
    { int e_ = run_seeds(w, B, H, W, st); if (e_) return e_; }
    prof_end(w.prof, S_SEEDS);
    prof_begin(w.prof, S_LOOKUP);
    int e = run_init_tables(w, B, st); if (e) return e;
    CPB_LAUNCH_COUNTED(k_lookup, dim3(blocks_for(BN, 256)), dim3(256), 0, st, pfinal, (const int*)w.hist, B, H, W, masks, w.t);
    CPB_CHECK_LAUNCH();
    prof_end(w.prof, S_LOOKUP);
    ProfScope ps(w.prof, S_FINALIZE);
    CPB_LAUNCH_COUNTED(k_gm_finalize, dim3(B), dim3(table_threads(H, W)), 0, st, w.t, H, W, msf, w.skey, w.sidx, counts, 0);
    CPB_CHECK_LAUNCH();
    return 0;   // caller applies w.t.remap
}

// n_iter per tile and the work list of the block-per-label kernels (labels beyond the warp kernels' bbox range)
int run_label_scan(const Workspace& w, int B, cudaStream_t st) {
    cudaMemsetAsync(w.t.niter, 0, B * sizeof(int), st);
    cudaMemsetAsync(w.jobs + B + 3, 0, 2 * sizeof(int), st);
    CPB_LAUNCH_COUNTED(k_qc_scan, dim3(B), dim3(256), 0, st, w.t, w.todo, w.jobs + B + 3);
    CPB_CHECK_LAUNCH();
    return 0;
}

inline LabelWork todo_work(const Workspace& w, int B, bool back) {
    return LabelWork{w.todo, w.jobs + B + 3, (int)std::min<size_t>((size_t)B * w.t.LC, 0x7fffffff), back ? 1 : 0};
}

// labels with statistics in the tables -> T (and mu / err / bad flags)
// exact_err: the caller wants the float64 error of every label (no float32 screen)
int run_flow_qc(const Workspace& w, const int32_t* masks, const float* dP, int B, int H, int W, double thr,
                double* mu_out, cudaStream_t st, bool exact_err = false) {
    const size_t smem = (size_t)CPB_DIFF_SMEM_CELLS * 17;
    // labels that touch no other live label get their flow error inside the diffusion warp (no T round trip); the
    // others are appended to the end of the work list for k_flow_err
    const float* qc_dP = (dP && !mu_out && qc_fused_enabled()) ? dP : nullptr;
    int* todo_n = w.jobs + B + 3;
    const LabelWork big = todo_work(w, B, false);
    if (qc_dP && diffuse_queue_enabled()) {
        // decision-exact path (cpb_qc32.cuh): k_qc_scan32 (centres, contact, classes) + k_qc_pack (jobs) -> float32
        // screen in registers (isolated labels decided in place, labels in contact via the float32 T plane and
        // k_flow_err32) -> float64 warp kernel for whatever the screen cannot hold or decide
        prof_begin(w.prof, S_CENTRES);
        const int screen = ((!exact_err || qc_screen_debug()) && qc_screen_enabled()) ? 1 : 0;
        const int pack_err = (exact_err && qc_screen_debug()) ? 1 : 0;
        cudaMemsetAsync(w.jobs + B + 3, 0, 2 * sizeof(int), st);
        cudaMemsetAsync(w.q.ctr, 0, CPB_QCTR_INTS * sizeof(int), st);
        cudaMemsetAsync(w.q.cls_cnt, 0, (size_t)B * CPB_Q32_NCLS * sizeof(int), st);
        cudaMemsetAsync(w.t.niter, 0, B * sizeof(int), st);
        CPB_LAUNCH_COUNTED(k_qc_scan32, dim3(tile_slices(H, W), B), dim3(256), 0, st, masks, H, W, w.t, w.q, w.todo, todo_n, screen);
        CPB_CHECK_LAUNCH();
        CPB_LAUNCH_COUNTED(k_qc_pack, dim3(B), dim3(32 * CPB_Q32_NCLS), 0, st, w.t, w.q);
        CPB_CHECK_LAUNCH();
        CPB_LAUNCH_COUNTED(k_centres, dim3(sm_count() * 8), dim3(CPB_QC_THREADS), 0, st, masks, H, W, w.t, 1, big);
        CPB_CHECK_LAUNCH();
        CPB_LAUNCH_COUNTED(k_q32_sort, dim3(sm_count()), dim3(256), 0, st, w.q);
        CPB_CHECK_LAUNCH();
        prof_end(w.prof, S_CENTRES);
        prof_begin(w.prof, S_DIFFUSE);
        CPB_LAUNCH_COUNTED(k_diffuse32, dim3(sm_count() * CPB_Q32_MINBLOCKS), dim3(128), 0, st, masks, qc_dP, H, W, w.t, w.q, thr, pack_err);
        CPB_CHECK_LAUNCH();
        CPB_LAUNCH_COUNTED(k_flow_err32, dim3(sm_count() * 8), dim3(128), 0, st, masks, qc_dP, H, W, w.t, w.q, thr, pack_err);
        CPB_CHECK_LAUNCH();
        CPB_LAUNCH_COUNTED(k_diffuse64_list, dim3(sm_count() * 6), dim3(CPB_DW_WARPS * 32), 0, st, masks, H, W, w.t, w.T, w.q,
                           qc_dP, thr, w.todo, todo_n, big.cap);
        CPB_CHECK_LAUNCH();
    } else {
        prof_begin(w.prof, S_CENTRES);
        { int e = run_label_scan(w, B, st); if (e) return e; }
        CPB_LAUNCH_COUNTED(k_centres, dim3(sm_count() * 8), dim3(CPB_QC_THREADS), 0, st, masks, H, W, w.t, 1, big);
        CPB_CHECK_LAUNCH();
        prof_end(w.prof, S_CENTRES);
        prof_begin(w.prof, S_DIFFUSE);
        if (diffuse_queue_enabled()) {
            // persistent warps pulling label pairs from one queue per size class (see k_diffuse_jobs)
            CPB_LAUNCH_COUNTED(k_diffuse_jobs, dim3(1), dim3(1024), 0, st, w.t.lbound, B, w.jobs, w.jobs + B + 1);
            CPB_CHECK_LAUNCH();
            int* ctr = w.jobs + B + 1;
            CPB_LAUNCH_COUNTED((k_diffuse_warp_q<CPB_DC_MIDH, 2>), dim3(sm_count() * CPB_DQ_MINBLOCKS), dim3(CPB_DW_WARPS * 32), 0, st, masks, B, H, W,
                               w.t, w.T, 0, w.jobs, ctr + 0, qc_dP, thr, w.todo, todo_n, big.cap);
            CPB_CHECK_LAUNCH();
            CPB_LAUNCH_COUNTED((k_diffuse_warp_q<CPB_DC_MAXH, 2>), dim3(sm_count() * 6), dim3(CPB_DW_WARPS * 32), 0, st, masks, B, H, W,
                               w.t, w.T, 0, w.jobs, ctr + 1, qc_dP, thr, w.todo, todo_n, big.cap);
            CPB_CHECK_LAUNCH();
        } else {
            CPB_LAUNCH_COUNTED(k_diffuse_warp<CPB_DC_MIDH>, dim3(kWarpDiffuseBlocksPerTile, B), dim3(CPB_DW_WARPS * 32), 0, st, masks,
                               H, W, w.t, w.T, 0, qc_dP, thr, w.todo, todo_n, big.cap);
            CPB_CHECK_LAUNCH();
            CPB_LAUNCH_COUNTED(k_diffuse_warp<CPB_DC_MAXH>, dim3(kWarpDiffuseBlocksPerTile, B), dim3(CPB_DW_WARPS * 32), 0, st, masks,
                               H, W, w.t, w.T, 0, qc_dP, thr, w.todo, todo_n, big.cap);
            CPB_CHECK_LAUNCH();
        }
    }
    CPB_LAUNCH_COUNTED(k_diffuse, dim3(sm_count() * 4), dim3(CPB_QC_THREADS), smem, st, masks, H, W, w.t, w.T, w.T2, 0, 1, big);
    CPB_CHECK_LAUNCH();
    prof_end(w.prof, S_DIFFUSE);
    ProfScope ps(w.prof, S_FLOWERR);
    CPB_LAUNCH_COUNTED(k_flow_err, dim3(sm_count() * 8), dim3(CPB_QC_THREADS), 0, st, masks, dP, H, W, w.t, w.T, thr, mu_out, 0,
                       todo_work(w, B, true));
    CPB_CHECK_LAUNCH();
    return 0;
}

// exact hole fill for tangled labels (see k_fill_sequential): detection + replay on the proposal plane in place
int run_fill_exact(const Workspace& w, const int32_t* masks, int B, int H, int W, bool order_by_remap, cudaStream_t st) {
    cudaMemsetAsync(w.cover, 0, (size_t)B * w.t.LC * sizeof(int), st);
    cudaMemsetAsync(w.seq, 0, B * sizeof(int), st);
    CPB_LAUNCH_COUNTED(k_fill_cover, dim3(tile_slices(H, W), B), dim3(256), 0, st, masks, (const u64*)w.holekey, H, W, w.t, w.cover);
    CPB_CHECK_LAUNCH();
    CPB_LAUNCH_COUNTED(k_fill_conflict, dim3(B), dim3(256), 0, st, w.t, (const int*)w.cover, w.seq);
    CPB_CHECK_LAUNCH();
    CPB_LAUNCH_COUNTED(k_fill_sequential, dim3(B), dim3(32), 0, st, masks, H, W, w.t, (const int*)w.seq, order_by_remap ? 1 : 0,
                       w.M, w.hist, w.pfinal, w.sinv, w.holekey);
    CPB_CHECK_LAUNCH();
    return 0;
}

// fill_holes_and_remove_small_masks on `masks` whose statistics are NOT yet in the tables
int run_fill_small(const Workspace& w, int32_t* masks, int B, int H, int W, int min_size, int32_t* counts,
                   bool have_stats, cudaStream_t st) {
    const long long BN = (long long)B * H * W;
    int e;
    cudaMemsetAsync(w.list_n, 0, 64 * sizeof(unsigned), st);          // bump cursor of the bitmap pool
    cudaMemsetAsync(w.t.fail, 0, B * sizeof(int), st);
    if (!have_stats) { e = run_map_stats(w, masks, B, H, W, 1, nullptr, nullptr, nullptr, true, st); if (e) return e; }
    const int mode = min_size > 0 ? 1 : 0;
    prof_begin(w.prof, S_SIZE1);
    CPB_LAUNCH_COUNTED(k_size_renumber, dim3(B), dim3(table_threads(H, W)), 0, st, w.t, H, W, min_size, mode, w.skey, w.sidx, (int*)nullptr);
    CPB_CHECK_LAUNCH();
    prof_end(w.prof, S_SIZE1);
    e = run_map_stats(w, masks, B, H, W, 1, w.t.remap, nullptr, nullptr, true, st, S_MAP2); if (e) return e;
    prof_begin(w.prof, S_FILL);
    cudaMemsetAsync(w.holekey, 0, BN * sizeof(u64), st);
    CPB_LAUNCH_COUNTED(k_fill_holes_warp, dim3(kWarpDiffuseBlocksPerTile, B), dim3(128), 0, st, masks, H, W, w.t, w.holekey,
                       CPB_FILL_BOTH);
    CPB_CHECK_LAUNCH();
    CPB_LAUNCH_COUNTED(k_fill_holes, dim3(kLabelBlocksPerTile, B), dim3(CPB_FILL_THREADS), 2 * CPB_FILL_WORDS * 4, st,
               masks, H, W, w.t, w.holekey, fill_pool(w, B, H, W), 1, (LabelWork{nullptr, nullptr, 0, 0}), CPB_FILL_BOTH);
    CPB_CHECK_LAUNCH();
    if (fill_exact_enabled()) { e = run_fill_exact(w, masks, B, H, W, false, st); if (e) return e; }
    prof_end(w.prof, S_FILL);
    e = run_map_stats(w, masks, B, H, W, 1, nullptr, nullptr, w.holekey, true, st, S_MAP3); if (e) return e;
    if (mode == 1) {
        prof_begin(w.prof, S_SIZE2);
        CPB_LAUNCH_COUNTED(k_size_renumber, dim3(B), dim3(table_threads(H, W)), 0, st, w.t, H, W, min_size, 1, w.skey, w.sidx, counts);
        CPB_CHECK_LAUNCH();
        prof_end(w.prof, S_SIZE2);
        e = run_map_stats(w, masks, B, H, W, 1, w.t.remap, nullptr, nullptr, false, st, S_MAP4); if (e) return e;
    } else if (counts) {
        cudaMemcpyAsync(counts, w.t.lbound, B * sizeof(int), cudaMemcpyDeviceToDevice, st);
    }
    if (counts) {
        CPB_LAUNCH_COUNTED(k_apply_fail, dim3(blocks_for(B, 256)), dim3(256), 0, st, (const int*)w.t.fail, B, counts);
        CPB_CHECK_LAUNCH();
    }
    return 0;
}

int run_vote(const Workspace& w, const int32_t* masks, const float* logits, int B, int H, int W, int C,
             int32_t* cell_class, uint8_t* class_masks, cudaStream_t st) {
    ProfScope ps(w.prof, S_VOTE);
    const long long BN = (long long)B * H * W;
    if ((H * W) % 4 == 0 && reinterpret_cast<uintptr_t>(masks) % 16 == 0 && reinterpret_cast<uintptr_t>(logits) % 16 == 0 &&
        (!class_masks || reinterpret_cast<uintptr_t>(class_masks) % 4 == 0)) {
        // streaming form: the whole grid reads labels and logits with 128-bit loads, histogram in the global table
        CPB_LAUNCH_COUNTED(k_vote_zero_lb, dim3(B), dim3(256), 0, st, (const int*)w.t.lbound, w.t.LC, C, w.vote);
        CPB_CHECK_LAUNCH();
#define CPB_VP_LAUNCH(CT) CPB_LAUNCH_COUNTED(k_vote_px_v4<CT>, dim3(blocks_for(BN / 4, 256)), dim3(256), 0, st,                \
                           reinterpret_cast<const int4*>(masks), reinterpret_cast<const float4*>(logits), B, H, W, C, w.t.LC, \
                           (const int*)w.t.lbound, w.vote)
        if (C == 7) { CPB_VP_LAUNCH(7); } else if (C == 10) { CPB_VP_LAUNCH(10); } else if (C == 5) { CPB_VP_LAUNCH(5); }
        else { CPB_VP_LAUNCH(0); }
#undef CPB_VP_LAUNCH
        CPB_CHECK_LAUNCH();
        CPB_LAUNCH_COUNTED(k_vote_finish_lb, dim3(B), dim3(256), 0, st, (const int*)w.t.lbound, w.t.LC, C, (const int*)w.vote, cell_class);
        CPB_CHECK_LAUNCH();
        if (class_masks) {
            CPB_LAUNCH_COUNTED(k_class_image_v4, dim3(blocks_for(BN / 4, 256)), dim3(256), 0, st, reinterpret_cast<const int4*>(masks),
                               B, H, W, w.t.LC, (const int*)w.t.lbound, (const int*)cell_class, reinterpret_cast<unsigned*>(class_masks));
            CPB_CHECK_LAUNCH();
        }
        return 0;
    }
    // (instance, class) table in shared memory: 24 KB for nuclei-scale tiles, up to 96 KB on big tiles
    const int smem_ints = (int)std::min<long long>(std::max<long long>((long long)H * W / 16, kVoteSmemInts), kVoteSmemIntsMax);
    CPB_LAUNCH_COUNTED(k_vote, dim3(B), dim3(512), smem_ints * 4, st, masks, logits, H, W, C, w.t.LC, w.t.lbound,
               smem_ints, w.vote, cell_class, class_masks);
    CPB_CHECK_LAUNCH();
    return 0;
}

int run_border(const Workspace& w, int32_t* masks, int B, int H, int W, int nch, cudaStream_t st) {
    ProfScope ps(w.prof, S_BORDER);
    int e = run_init_tables(w, B, st); if (e) return e;
    CPB_LAUNCH_COUNTED(k_border_flags, dim3(B), dim3(256), 0, st, masks, H, W, nch, w.t);
    CPB_CHECK_LAUNCH();
    return run_map_stats(w, masks, B, H, W, nch, nullptr, w.t.flag, nullptr, false, st);
}

}  // namespace

extern "C" {

int cpb_abi_version(void) { return CPB_ABI_VERSION; }

int cpb_label_capacity(int H, int W) {
    // every seed needs more than 10 end points, so a tile yields at most H*W/11 labels
    return (int)((long long)H * W / 11) + 2;
}

size_t cpb_workspace_bytes(int B, int H, int W, int C, int lcap) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    const size_t whole = carve(nullptr, B, H, W, C < 0 ? 0 : C, lcap).bytes + kAlign;
    // the fused path may cut the batch into parts that run on parallel streams, each with its own carve
    size_t parts = 0;
    const int np = batch_parts(B);
    for (int k = 0; k < np; k++) parts += carve(nullptr, part_begin(B, np, k + 1) - part_begin(B, np, k), H, W, C < 0 ? 0 : C, lcap).bytes + 2 * kAlign;
    return std::max(whole, parts);
}

#define CPB_PROLOGUE(Cval, lcapval)                                                            \
    { int e_ = check_geom(B, H, W); if (e_) return e_; }                                       \
    if (!workspace) return CPB_E_ARG;                                                          \
    ensure_attributes();                                                                       \
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);                                  \
    const uintptr_t wsa_ = (reinterpret_cast<uintptr_t>(workspace) + kAlign - 1) / kAlign * kAlign; \
    Workspace w = carve(reinterpret_cast<void*>(wsa_), B, H, W, (Cval), (lcapval));            \
    w.prof = nullptr;            \
    if (w.bytes + (wsa_ - reinterpret_cast<uintptr_t>(workspace)) > workspace_bytes) return CPB_E_WORKSPACE;

int cpb_follow_flows_device(const float* dP, const float* cellprob, int B, int H, int W, int niter,
                            float cellprob_threshold, int32_t* p_final, float* p_float, void* workspace,
                            size_t workspace_bytes, void* stream) {
    if (!dP || !cellprob || !p_final || niter < 0) return CPB_E_ARG;
    CPB_PROLOGUE(0, 0)
    return run_follow(w, dP, cellprob, B, H, W, niter, cellprob_threshold, p_final, p_float, nullptr, st);
}

CPB_KERNEL k_hist_from_pfinal(const int* CPB_RESTRICT pfinal, int B, int H, int W, int* CPB_RESTRICT hist) {
    const int N = H * W;
    const long long total = (long long)B * N;
    const long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= total) return;
    const int pf = pfinal[gi];
    if (pf < 0) return;
    const int b = (int)(gi / N);
    const int y = min(pf >> 16, H - 1), x = min(pf & 0xffff, W - 1);
    atomicAdd(&hist[(size_t)b * N + y * W + x], 1);
}

int cpb_get_masks_device(const int32_t* p_final, int B, int H, int W, double max_size_fraction, int32_t* masks,
                         int32_t* counts, void* workspace, size_t workspace_bytes, void* stream) {
    if (!p_final || !masks) return CPB_E_ARG;
    CPB_PROLOGUE(0, 0)
    const long long BN = (long long)B * H * W;
    cudaMemsetAsync(w.hist, 0, BN * sizeof(int), st);
    CPB_LAUNCH_COUNTED(k_hist_from_pfinal, dim3(blocks_for(BN, 256)), dim3(256), 0, st, p_final, B, H, W, w.hist);
    CPB_CHECK_LAUNCH();
    int e = run_get_masks(w, p_final, B, H, W, max_size_fraction, masks, counts, st); if (e) return e;
    return run_map_stats(w, masks, B, H, W, 1, w.t.remap, nullptr, nullptr, false, st);
}

int cpb_masks_to_flows_device(const int32_t* masks, int B, int H, int W, int lcap, double* mu, void* workspace,
                              size_t workspace_bytes, void* stream) {
    if (!masks || !mu || lcap < 2) return CPB_E_ARG;
    CPB_PROLOGUE(0, lcap)
    const long long BN = (long long)B * H * W;
    int e = run_set_lbound(w, B, lcap - 1, st); if (e) return e;
    e = run_map_stats(w, const_cast<int32_t*>(masks), B, H, W, 1, nullptr, nullptr, nullptr, true, st); if (e) return e;
    cudaMemsetAsync(mu, 0, 2 * BN * sizeof(double), st);
    return run_flow_qc(w, masks, nullptr, B, H, W, 0.0, mu, st);
}

int cpb_remove_bad_flow_masks_device(int32_t* masks, const float* dP, int B, int H, int W, int lcap,
                                     double threshold, double* flow_err, void* workspace, size_t workspace_bytes,
                                     void* stream) {
    if (!masks || !dP || lcap < 2) return CPB_E_ARG;
    CPB_PROLOGUE(0, lcap)
    int e = run_set_lbound(w, B, lcap - 1, st); if (e) return e;
    e = run_map_stats(w, masks, B, H, W, 1, nullptr, nullptr, nullptr, true, st); if (e) return e;
    if (flow_err) cudaMemsetAsync(w.t.err, 0, (size_t)B * w.t.LC * sizeof(double), st);
    e = run_flow_qc(w, masks, dP, B, H, W, threshold, nullptr, st, flow_err != nullptr); if (e) return e;
    if (flow_err) cudaMemcpyAsync(flow_err, w.t.err, (size_t)B * w.t.LC * sizeof(double), cudaMemcpyDeviceToDevice, st);
    return run_map_stats(w, masks, B, H, W, 1, nullptr, w.t.flag, nullptr, false, st);
}

int cpb_fill_holes_and_remove_small_masks_device(int32_t* masks, int B, int H, int W, int lcap, int min_size,
                                                 int32_t* counts, void* workspace, size_t workspace_bytes,
                                                 void* stream) {
    if (!masks || lcap < 2) return CPB_E_ARG;
    CPB_PROLOGUE(0, lcap)
    int e = run_set_lbound(w, B, lcap - 1, st); if (e) return e;
    return run_fill_small(w, masks, B, H, W, min_size, counts, false, st);
}

int cpb_class_vote_device(const int32_t* masks, const float* logits, int B, int H, int W, int C, int lcap,
                          int32_t* cell_class, uint8_t* class_masks, void* workspace, size_t workspace_bytes,
                          void* stream) {
    if (!masks || !logits || !cell_class || C < 1 || C > 255 || lcap < 2) return CPB_E_ARG;
    CPB_PROLOGUE(C, lcap)
    int e = run_set_lbound(w, B, lcap - 1, st); if (e) return e;
    return run_vote(w, masks, logits, B, H, W, C, cell_class, class_masks, st);
}

CPB_KERNEL k_lbound_from_counts(const int* CPB_RESTRICT counts, int B, int cap, int* CPB_RESTRICT lbound) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) lbound[b] = min(max(counts[b], 0), cap);
}

int cpb_class_vote_counts_device(const int32_t* masks, const float* logits, const int32_t* counts, int B, int H, int W,
                                 int C, int lcap, int32_t* cell_class, uint8_t* class_masks, void* workspace,
                                 size_t workspace_bytes, void* stream) {
    if (!masks || !logits || !counts || !cell_class || C < 1 || C > 255 || lcap < 2) return CPB_E_ARG;
    CPB_PROLOGUE(C, lcap)
    CPB_LAUNCH_COUNTED(k_lbound_from_counts, dim3(blocks_for(B, 256)), dim3(256), 0, st, counts, B, lcap - 1, w.t.lbound);
    CPB_CHECK_LAUNCH();
    return run_vote(w, masks, logits, B, H, W, C, cell_class, class_masks, st);
}

int cpb_remove_border_instances_device(int32_t* masks, int B, int H, int W, int nch, int lcap, void* workspace,
                                       size_t workspace_bytes, void* stream) {
    if (!masks || nch < 1 || lcap < 2) return CPB_E_ARG;
    { int e_ = check_geom(B, H, W); if (e_) return e_; }
    if ((long long)B * H * W * nch >= (1LL << 31)) return CPB_E_RANGE;
    CPB_PROLOGUE(0, lcap)
    int e = run_set_lbound(w, B, lcap - 1, st); if (e) return e;
    return run_border(w, masks, B, H, W, nch, st);
}

}  // extern "C"

// dP_copy (host path only): dP is a mapped HOST pointer; the prep kernel reads the groups that hold foreground
// through it and keeps them in dP_copy [B,2,H,W] on the device, which is what the flow check then reads
static int compute_masks_impl(const float* dP, const float* cellprob, const float* logits, int B, int H, int W,
                             int C, const cpb_params* prm, int32_t* masks, int32_t* counts, int32_t* cell_class,
                             uint8_t* class_masks, void* workspace, size_t workspace_bytes, void* stream, Prof* prof,
                             float* dP_copy = nullptr, bool prep_done = false) {
    if (!dP || !cellprob || !prm || !masks || !counts) return CPB_E_ARG;
    if (logits && (!cell_class || C < 1 || C > 255)) return CPB_E_ARG;
    if (prm->niter < 0) return CPB_E_ARG;
    CPB_PROLOGUE(logits ? C : 0, 0)
    w.prof = prof;
    if (prof) prof->st = st;
    int e;
    const long long BN = (long long)B * H * W;
    // (2) Euler integration + end-point histogram
    const bool seed_cands = seed_cands_enabled();
    e = run_follow(w, dP, cellprob, B, H, W, prm->niter, prm->cellprob_threshold, w.pfinal, nullptr, w.hist, st, masks, dP_copy,
                   prep_done, seed_cands);
    if (e) return e;
    if (dP_copy) dP = dP_copy;          // every later reader (flow check) sits on a foreground pixel
    // (3) seeds -> raw labels (seed order + 1) and their statistics; ids after get_masks live in t.remap
    prof_begin(w.prof, S_SEEDS);
    { int e_ = run_seeds(w, B, H, W, st, seed_cands); if (e_) return e_; }
    prof_end(w.prof, S_SEEDS);
    prof_begin(w.prof, S_LOOKUP);
    e = run_init_tables(w, B, st); if (e) return e;
    cudaMemsetAsync(w.t.misc, 0, B * sizeof(int), st);      // (the label image was zeroed by the prep kernel)
    {
        const unsigned grid = (unsigned)std::min<long long>(blocks_for(BN, 256), (long long)sm_count() * 8);
        CPB_LAUNCH_COUNTED(k_lookup_list, dim3(grid), dim3(256), 0, st, w.list, w.list_n, w.pfinal, (const int*)w.hist, H, W, masks, w.t);
        CPB_CHECK_LAUNCH();
    }
    prof_end(w.prof, S_LOOKUP);
    w.t.alive = w.alive;
    prof_begin(w.prof, S_FINALIZE);
    CPB_LAUNCH_COUNTED(k_gm_finalize, dim3(B), dim3(table_threads(H, W)), 0, st, w.t, H, W, prm->max_size_fraction, w.skey, w.sidx,
                       (int*)nullptr, 1);
    CPB_CHECK_LAUNCH();
    prof_end(w.prof, S_FINALIZE);
    // (4) flow-error check on the raw labels; bad labels are only flagged
    if (prm->flow_threshold > 0.0) {
        e = run_flow_qc(w, masks, dP, B, H, W, prm->flow_threshold, nullptr, st); if (e) return e;
    } else if (prm->fill_holes) {
        e = run_label_scan(w, B, st); if (e) return e;       // the hole fill below wants the list of large labels
    }
    // (5) size filter / hole fill / size filter as table operations, one final pixel pass
    if (prm->fill_holes) {
        prof_begin(w.prof, S_SIZE1);
        // min_size <= 0: upstream skips both size filters and with them every first-appearance renumbering
        CPB_LAUNCH_COUNTED(k_fuse_size, dim3(B), dim3(table_threads(H, W)), 0, st, w.t, H, W, prm->min_size,
                           prm->min_size > 0 ? 1 : 3, (const int*)nullptr, w.skey, w.sidx, w.sinv);
        CPB_CHECK_LAUNCH();
        prof_end(w.prof, S_SIZE1);
        prof_begin(w.prof, S_FILL);
        // detect -> zero the hole plane of the tiles that have a hole -> write the proposals of those tiles
        for (int pass = CPB_FILL_DETECT; pass <= CPB_FILL_WRITE; pass++) {
            CPB_LAUNCH_COUNTED(k_fill_holes_warp, dim3(kWarpDiffuseBlocksPerTile, B), dim3(128), 0, st, masks, H, W, w.t,
                               w.holekey, pass);
            CPB_CHECK_LAUNCH();
            CPB_LAUNCH_COUNTED(k_fill_holes, dim3(sm_count() * 3), dim3(CPB_FILL_THREADS), 2 * CPB_FILL_WORDS * 4, st,
                               masks, H, W, w.t, w.holekey, fill_pool(w, B, H, W), 1, todo_work(w, B, false), pass);
            CPB_CHECK_LAUNCH();
            if (pass == CPB_FILL_DETECT) {
                CPB_LAUNCH_COUNTED(k_zero_hole_tiles, dim3(tile_slices(H, W), B), dim3(256), 0, st, w.holekey, H, W, w.t);
                CPB_CHECK_LAUNCH();
            }
        }
        if (fill_exact_enabled()) { e = run_fill_exact(w, masks, B, H, W, true, st); if (e) return e; }
        prof_end(w.prof, S_FILL);
        prof_begin(w.prof, S_MAP3);
        CPB_LAUNCH_COUNTED(k_recount_reset, dim3(B), dim3(256), 0, st, w.t);
        CPB_CHECK_LAUNCH();
        CPB_LAUNCH_COUNTED(k_recount, dim3(tile_slices(H, W), B), dim3(256), 0, st, masks, w.holekey, H, W, w.t);
        CPB_CHECK_LAUNCH();
        prof_end(w.prof, S_MAP3);
        prof_begin(w.prof, S_SIZE2);
        // second filter on every tile: besides holes, labels that survived the positional first filter are
        // caught here (labels are contiguous again, so position == value)
        CPB_LAUNCH_COUNTED(k_fuse_size, dim3(B), dim3(table_threads(H, W)), 0, st, w.t, H, W, prm->min_size,
                           prm->min_size > 0 ? 1 : 4, (const int*)nullptr, w.skey, w.sidx, w.sinv);
        CPB_CHECK_LAUNCH();
        prof_end(w.prof, S_SIZE2);
    } else {
        prof_begin(w.prof, S_SIZE1);
        CPB_LAUNCH_COUNTED(k_fuse_size, dim3(B), dim3(table_threads(H, W)), 0, st, w.t, H, W, 0, 2, (const int*)nullptr, w.skey, w.sidx,
                           w.sinv);
        CPB_CHECK_LAUNCH();
        prof_end(w.prof, S_SIZE1);
    }
    // (6) the class vote rides on the final pass unless something between the two changes the labels (border
    // removal) or the per-pixel class image is wanted (it needs the finished per-instance classes)
    const bool vote_fused = logits && !class_masks && !prm->remove_border && vote_fused_enabled() && (H * W) % 4 == 0 &&
                            reinterpret_cast<uintptr_t>(masks) % 16 == 0 && reinterpret_cast<uintptr_t>(logits) % 16 == 0;
    if (vote_fused) {
        prof_begin(w.prof, S_MAP4);
        CPB_LAUNCH_COUNTED(k_vote_zero, dim3(B), dim3(256), 0, st, w.t, C, w.vote);
        CPB_CHECK_LAUNCH();
#define CPB_FV_LAUNCH(CT) CPB_LAUNCH_COUNTED(k_final_vote_v4<CT>, dim3(blocks_for(BN / 4, 256)), dim3(256), 0, st,        \
                           reinterpret_cast<int4*>(masks), prm->fill_holes ? (const u64*)w.holekey : (const u64*)nullptr, \
                           reinterpret_cast<const float4*>(logits), B, H, W, C, w.t, counts, w.vote)
        // class counts of the reference's model configurations (conic / consep / nucls 7, puma 10, monusac / glysac 5)
        if (C == 7) { CPB_FV_LAUNCH(7); } else if (C == 10) { CPB_FV_LAUNCH(10); } else if (C == 5) { CPB_FV_LAUNCH(5); }
        else { CPB_FV_LAUNCH(0); }
#undef CPB_FV_LAUNCH
        CPB_CHECK_LAUNCH();
        CPB_LAUNCH_COUNTED(k_finish_bounds, dim3(blocks_for(B, 256)), dim3(256), 0, st, w.t, B, counts);
        CPB_CHECK_LAUNCH();
        prof_end(w.prof, S_MAP4);
        w.t.alive = nullptr;
        ProfScope ps(w.prof, S_VOTE);
        CPB_LAUNCH_COUNTED(k_vote_finish, dim3(B), dim3(256), 0, st, w.t, C, (const int*)w.vote, cell_class);
        CPB_CHECK_LAUNCH();
        return 0;
    }
    prof_begin(w.prof, S_MAP4);
    if ((H * W) % 4 == 0 && reinterpret_cast<uintptr_t>(masks) % 16 == 0) {
        CPB_LAUNCH_COUNTED(k_final_v4, dim3(blocks_for(BN / 4, 256)), dim3(256), 0, st, reinterpret_cast<int4*>(masks),
                           prm->fill_holes ? (const u64*)w.holekey : (const u64*)nullptr, B, H, W, w.t, counts);
    } else {
        CPB_LAUNCH_COUNTED(k_final, dim3(blocks_for(BN, 256)), dim3(256), 0, st, masks,
                           prm->fill_holes ? (const u64*)w.holekey : (const u64*)nullptr, B, H, W, w.t, counts);
    }
    CPB_CHECK_LAUNCH();
    CPB_LAUNCH_COUNTED(k_finish_bounds, dim3(blocks_for(B, 256)), dim3(256), 0, st, w.t, B, counts);
    CPB_CHECK_LAUNCH();
    prof_end(w.prof, S_MAP4);
    w.t.alive = nullptr;
    // (7) optional border-instance removal (labels are not renumbered afterwards, as in the reference)
    if (prm->remove_border) { e = run_border(w, masks, B, H, W, 1, st); if (e) return e; }
    // (6) class vote
    if (logits) { e = run_vote(w, masks, logits, B, H, W, C, cell_class, class_masks, st); if (e) return e; }
    return 0;
}


#ifndef CPB_SIM
namespace {
struct PartStreams {            // helper streams and events of one host thread (created on first use, per device)
    int device = -1;
    cudaStream_t helper[kMaxParts - 1] = {};
    cudaEvent_t ready = nullptr, done[kMaxParts - 1] = {};
    bool ensure() {
        int dev = 0;
        cudaGetDevice(&dev);
        if (device == dev) return true;
        release();
        bool ok = cudaEventCreateWithFlags(&ready, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; i < kMaxParts - 1 && ok; i++)
            ok = cudaStreamCreateWithFlags(&helper[i], cudaStreamNonBlocking) == cudaSuccess &&
                 cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming) == cudaSuccess;
        if (!ok) { cudaGetLastError(); release(); return false; }
        device = dev;
        return true;
    }
    void release() {
        for (int i = 0; i < kMaxParts - 1; i++) {
            if (helper[i]) cudaStreamDestroy(helper[i]);
            if (done[i]) cudaEventDestroy(done[i]);
            helper[i] = nullptr; done[i] = nullptr;
        }
        if (ready) cudaEventDestroy(ready);
        ready = nullptr; device = -1;
    }
    ~PartStreams() { release(); }
};
thread_local PartStreams tl_parts;
}  // namespace
#endif

// fork / join: part(b0, nb, ws, ws_bytes, stream) runs sub-batch [b0, b0 + nb) of the call; part 0 stays on the caller's
// stream, the others run on this thread's helper streams behind an event, and the caller's stream waits for all of them.
// Still one asynchronous, stream-ordered call for the caller.  Not used while the caller's stream is being captured.
template <class F>
static int run_in_parts(int B, int H, int W, int C, void* workspace, size_t workspace_bytes, void* stream, F&& part) {
#ifndef CPB_SIM
    const int np = batch_parts(B);
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(reinterpret_cast<cudaStream_t>(stream), &cap) != cudaSuccess) cudaGetLastError();
    if (np > 1 && workspace && cap == cudaStreamCaptureStatusNone && tl_parts.ensure()) {
        cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
        char* ws = reinterpret_cast<char*>(workspace);
        size_t off = 0;
        cudaEventRecord(tl_parts.ready, st);
        int rc = 0;
        for (int k = 0; k < np && rc == 0; k++) {
            const int b0 = part_begin(B, np, k), nb = part_begin(B, np, k + 1) - b0;
            const size_t need = carve(nullptr, nb, H, W, C, 0).bytes + 2 * kAlign;
            if (off + need > workspace_bytes) { rc = CPB_E_WORKSPACE; break; }
            cudaStream_t sk = k == 0 ? st : tl_parts.helper[k - 1];
            if (k > 0) cudaStreamWaitEvent(sk, tl_parts.ready, 0);
            rc = part(b0, nb, static_cast<void*>(ws + off), need, static_cast<void*>(sk));
            if (k > 0) { cudaEventRecord(tl_parts.done[k - 1], sk); cudaStreamWaitEvent(st, tl_parts.done[k - 1], 0); }
            off += need;
        }
        return rc;
    }
#endif
    return part(0, B, workspace, workspace_bytes, stream);
}

extern "C" {

int cpb_compute_masks_device(const float* dP, const float* cellprob, const float* logits, int B, int H, int W,
                             int C, const cpb_params* prm, int32_t* masks, int32_t* counts, int32_t* cell_class,
                             uint8_t* class_masks, void* workspace, size_t workspace_bytes, void* stream) {
    if (B <= 0 || H < 2 || W < 2) return CPB_E_ARG;
    const size_t N = (size_t)H * W;
    const int LC = cpb_label_capacity(H, W);
    return run_in_parts(B, H, W, logits ? C : 0, workspace, workspace_bytes, stream,
                        [&](int b0, int nb, void* ws, size_t ws_bytes, void* st) {
        return compute_masks_impl(dP ? dP + (size_t)b0 * 2 * N : nullptr, cellprob ? cellprob + (size_t)b0 * N : nullptr,
                                  logits ? logits + (size_t)b0 * C * N : nullptr, nb, H, W, C, prm, masks ? masks + (size_t)b0 * N : nullptr,
                                  counts ? counts + b0 : nullptr, cell_class ? cell_class + (size_t)b0 * LC : nullptr,
                                  class_masks ? class_masks + (size_t)b0 * N : nullptr, ws, ws_bytes, st, nullptr);
    });
}

int cpb_num_stages(void) { return S_COUNT; }
const char* cpb_stage_name(int i) { return (i >= 0 && i < S_COUNT) ? kStageNames[i] : ""; }
void cpb_debug_set_follow_merge(int mode) { g_follow_merge.store(mode, std::memory_order_relaxed); }
void cpb_debug_set_switch(int which, int value) {
    if (which == CPB_SWITCH_FOLLOW_MERGE) g_follow_merge.store(value, std::memory_order_relaxed);
    else if (which > 0 && which < 9) g_switch[which].store(value, std::memory_order_relaxed);
}
long long cpb_debug_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
void cpb_debug_qc_stats(int32_t* out) {
    for (int i = 0; i < CPB_QCTR_INTS; i++) out[i] = g_last_qc[i].load(std::memory_order_relaxed);
}

int cpb_compute_masks_profiled_device(const float* dP, const float* cellprob, const float* logits, int B, int H,
                                      int W, int C, const cpb_params* prm, int32_t* masks, int32_t* counts,
                                      int32_t* cell_class, uint8_t* class_masks, void* workspace,
                                      size_t workspace_bytes, void* stream, float* stage_ms) {
#ifdef CPB_SIM
    (void)stage_ms;
    const int rc_ = compute_masks_impl(dP, cellprob, logits, B, H, W, C, prm, masks, counts, cell_class, class_masks, workspace,
                                       workspace_bytes, stream, nullptr);
    if (rc_ == 0) {
        const uintptr_t wsa = (reinterpret_cast<uintptr_t>(workspace) + kAlign - 1) / kAlign * kAlign;
        const Workspace w = carve(reinterpret_cast<void*>(wsa), B, H, W, logits ? C : 0, 0);
        for (int i = 0; i < CPB_QCTR_INTS; i++) g_last_qc[i].store(w.q.ctr[i], std::memory_order_relaxed);
    }
    return rc_;
#else
    if (!stage_ms) return CPB_E_ARG;
    Prof prof{};
    for (int i = 0; i < S_COUNT; i++) { cudaEventCreate(&prof.begin[i]); cudaEventCreate(&prof.end[i]); prof.used[i] = false; }
    int rc = compute_masks_impl(dP, cellprob, logits, B, H, W, C, prm, masks, counts, cell_class, class_masks,
                                workspace, workspace_bytes, stream, &prof);
    cudaError_t ce = cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream));
    if (rc == 0 && ce == cudaSuccess) {      // flow-check statistics of this call (cpb_debug_qc_stats)
        const uintptr_t wsa = (reinterpret_cast<uintptr_t>(workspace) + kAlign - 1) / kAlign * kAlign;
        const Workspace w = carve(reinterpret_cast<void*>(wsa), B, H, W, logits ? C : 0, 0);
        int host[CPB_QCTR_INTS] = {0};
        if (cudaMemcpy(host, w.q.ctr, sizeof(host), cudaMemcpyDeviceToHost) == cudaSuccess)
            for (int i = 0; i < CPB_QCTR_INTS; i++) g_last_qc[i].store(host[i], std::memory_order_relaxed);
    }
    for (int i = 0; i < S_COUNT; i++) {
        stage_ms[i] = 0.f;
        if (rc == 0 && ce == cudaSuccess && prof.used[i]) cudaEventElapsedTime(&stage_ms[i], prof.begin[i], prof.end[i]);
        cudaEventDestroy(prof.begin[i]); cudaEventDestroy(prof.end[i]);
    }
    return rc ? rc : (int)ce;
#endif
}

int cpb_prepare_tiles_device(const float* img, int B, int H, int W, int C, double lower, double upper, int pad_y,
                             int pad_x, int ntiles, int ly, int lx, const int32_t* y0, const int32_t* x0,
                             const int32_t* flip, float* tiles, float* lowhigh, int32_t* code, void* stream) {
    if (!img || !y0 || !x0 || !flip || !tiles || !lowhigh || !code || B <= 0 || H <= 0 || W <= 0 || C <= 0 || ntiles <= 0 ||
        ly <= 0 || lx <= 0 || !(lower >= 0.0 && upper <= 100.0 && lower <= upper))
        return CPB_E_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    CPB_LAUNCH_COUNTED(k_percentiles, dim3(B * C), dim3(1024), 0, st, img, H, W, C, lower, upper, lowhigh, code);
    CPB_CHECK_LAUNCH();
    const long long total = (long long)B * ntiles * C * ly * lx;
    CPB_LAUNCH_COUNTED(k_make_tiles, dim3(blocks_for(total, 256)), dim3(256), 0, st, img, B, H, W, C, pad_y, pad_x, ntiles, ly,
                       lx, y0, x0, flip, (const float*)lowhigh, (const int*)code, tiles);
    CPB_CHECK_LAUNCH();
    return 0;
}

static unsigned dedup_table_size(long long n) {
    unsigned m = 1024;
    while ((long long)m < 2 * n && m < (1u << 30)) m <<= 1;
    return m;
}

size_t cpb_dedup_workspace_bytes(int64_t n) {
    if (n <= 0) return kAlign;
    Carver c{nullptr, 0};
    c.take<int>(dedup_table_size(n)); c.take<int>(n); c.take<int>(n); c.take<u64>(n); c.take<int>(n);
    return c.off + kAlign;
}

int cpb_dedup_cells_device(const double* cx, const double* cy, const double* size, int64_t n, double max_dist,
                           int32_t* keep, int32_t* group, void* workspace, size_t workspace_bytes, void* stream) {
    if (n == 0) return 0;
    if (!cx || !cy || !size || !keep || !workspace || n < 0 || n >= (1LL << 31) || !(max_dist > 0.0)) return CPB_E_ARG;
    if (workspace_bytes < cpb_dedup_workspace_bytes(n)) return CPB_E_WORKSPACE;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const uintptr_t wsa = (reinterpret_cast<uintptr_t>(workspace) + kAlign - 1) / kAlign * kAlign;
    Carver c{reinterpret_cast<char*>(wsa), 0};
    const unsigned M = dedup_table_size(n);
    int* head = c.take<int>(M); int* next = c.take<int>(n); int* parent = c.take<int>(n);
    u64* best_size = c.take<u64>(n); int* best_idx = c.take<int>(n);
    cudaMemsetAsync(head, 0xff, (size_t)M * sizeof(int), st);
    const dim3 grid(blocks_for(n, 256));
    const double inv_cell = 1.0 / max_dist;
    CPB_LAUNCH_COUNTED(k_dedup_insert, grid, dim3(256), 0, st, cx, cy, (long long)n, inv_cell, M - 1, head, next, parent,
                       best_size, best_idx);
    CPB_CHECK_LAUNCH();
    CPB_LAUNCH_COUNTED(k_dedup_link, grid, dim3(256), 0, st, cx, cy, (long long)n, inv_cell, max_dist * max_dist, M - 1,
                       (const int*)head, (const int*)next, parent);
    CPB_CHECK_LAUNCH();
    for (int phase = 0; phase < 3; phase++) {
        CPB_LAUNCH_COUNTED(k_dedup_select, grid, dim3(256), 0, st, size, (long long)n, parent, best_size, best_idx, phase,
                           keep, group);
        CPB_CHECK_LAUNCH();
    }
    return 0;
}

int cpb_cell_contours_device(const int32_t* masks, int B, int H, int W, int lcap, int32_t* npoints, int64_t* offsets,
                             int64_t* total, int16_t* points, int64_t points_cap, int64_t* feat, double* perimeter,
                             int32_t* valid, void* workspace, size_t workspace_bytes, void* stream) {
    if (!masks || !npoints || !offsets || !total || !points || !feat || !perimeter || !valid || lcap < 2 || points_cap < 0)
        return CPB_E_ARG;
    CPB_PROLOGUE(0, lcap)
    const long long BN = (long long)B * H * W;
    int e = run_set_lbound(w, B, lcap - 1, st); if (e) return e;
    e = run_map_stats(w, const_cast<int32_t*>(masks), B, H, W, 1, nullptr, nullptr, nullptr, true, st); if (e) return e;
    cudaMemsetAsync(w.M, 0, BN * sizeof(int), st);                       // border marks
    const dim3 grid(4, B);
    long long* tile_base = reinterpret_cast<long long*>(w.skey);
    CPB_LAUNCH_COUNTED(k_contours, grid, dim3(128), 0, st, masks, H, W, w.t, w.M, npoints, reinterpret_cast<long long*>(feat),
                       perimeter, w.sidx, (const long long*)nullptr, (short*)nullptr, (long long)0, (int*)nullptr);
    CPB_CHECK_LAUNCH();
    CPB_LAUNCH_COUNTED(k_contour_tile_sums, dim3(B), dim3(256), 0, st, npoints, w.t, w.t.misc);
    CPB_CHECK_LAUNCH();
    CPB_LAUNCH_COUNTED(k_label_offsets, dim3(1), dim3(256), 0, st, w.t.misc, B, (long long)0, tile_base,
                       reinterpret_cast<long long*>(total));
    CPB_CHECK_LAUNCH();
    CPB_LAUNCH_COUNTED(k_contour_offsets, dim3(B), dim3(256), 0, st, npoints, w.t, tile_base,
                       reinterpret_cast<long long*>(offsets));
    CPB_CHECK_LAUNCH();
    CPB_LAUNCH_COUNTED(k_contours, grid, dim3(128), 0, st, masks, H, W, w.t, w.M, npoints, reinterpret_cast<long long*>(feat),
                       perimeter, w.sidx, reinterpret_cast<const long long*>(offsets), reinterpret_cast<short*>(points),
                       (long long)points_cap, valid);
    CPB_CHECK_LAUNCH();
    return 0;
}

int cpb_average_tiles_device(const float* y, int B, int ntiles, int nch, int ly, int lx, const int32_t* y0,
                             const int32_t* x0, const int32_t* flip, int negate_flow, const double* taper_y,
                             const double* taper_x, int Ly, int Lx, int cy0, int cy1, int cx0, int cx1, float* yf,
                             void* stream) {
    return cpb_average_tiles_ex_device(y, B, ntiles, nch, ly, lx, y0, x0, flip, negate_flow, taper_y, taper_x, Ly, Lx,
                                       cy0, cy1, cx0, cx1, yf, 0, 0, stream);
}

int cpb_average_tiles_ex_device(const float* y, int B, int ntiles, int nch, int ly, int lx, const int32_t* y0,
                                const int32_t* x0, const int32_t* flip, int negate_flow, const double* taper_y,
                                const double* taper_x, int Ly, int Lx, int cy0, int cy1, int cx0, int cx1, float* yf,
                                int x0_multiple_of_4, int max_cover, void* stream) {
    if (!y || !y0 || !x0 || !flip || !taper_y || !taper_x || !yf) return CPB_E_ARG;
    const int oH = Ly - cy0 - cy1, oW = Lx - cx0 - cx1;
    if (B <= 0 || ntiles <= 0 || nch <= 0 || ly <= 0 || lx <= 0 || oH <= 0 || oW <= 0 || cy0 < 0 || cx0 < 0)
        return CPB_E_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // vector path: the host vouches for the tile geometry (device arrays are not read back here)
    (void)max_cover;
    const bool vec4 = x0_multiple_of_4 && (blend_eft_enabled() || nch <= 16) && (lx % 4 == 0) && (cx0 % 4 == 0) && (oW % 4 == 0) &&
                      (reinterpret_cast<uintptr_t>(y) % 16 == 0) && (reinterpret_cast<uintptr_t>(yf) % 16 == 0);
    if (vec4 && blend_eft_enabled()) {
        // float32 error-free blend: weight table and per-pixel reciprocal normaliser from two tiny kernels, scratch from
        // the stream-ordered allocator (freed on the stream after the blend)
        float* tab = nullptr;
        const size_t nw = (size_t)ly * lx, nr = (size_t)oH * oW;
        if (scratch_alloc(reinterpret_cast<void**>(&tab), (2 * nw + 2 * nr) * sizeof(float), st) != cudaSuccess) return (int)cudaGetLastError();
        float* wh = tab; float* wl = tab + nw; float* rh = tab + 2 * nw; float* rl = rh + nr;
        CPB_LAUNCH_COUNTED(k_blend_weights, dim3(blocks_for((long long)nw, 256)), dim3(256), 0, st, taper_y, taper_x, ly, lx, wh, wl);
        CPB_LAUNCH_COUNTED(k_blend_rinv, dim3(blocks_for((long long)nr, 256)), dim3(256), 0, st, ntiles, ly, lx, y0, x0, taper_y, taper_x,
                           cy0, cx0, oH, oW, rh, rl);
        // channel groups of at most 5 (register budget), none of them padded: 3 -> 3, 5 -> 5, 7 -> 4 + 3, 10 -> 5 + 5
        const dim3 grid(blocks_for((long long)B * oH * (oW / 4), 256));
#define CPB_EFT_LAUNCH(N) CPB_LAUNCH_COUNTED(k_average_tiles_eft<N>, grid, dim3(256), 0, st, y, B, ntiles, nch, c0, ly, lx, y0, x0, flip, \
                           negate_flow, (const float*)wh, (const float*)wl, (const float*)rh, (const float*)rl, cy0, cx0, oH, oW, yf)
        for (int c0 = 0; c0 < nch;) {
            const int rem = nch - c0;
            const int n = rem <= 5 ? rem : (rem == 6 ? 3 : (rem == 7 ? 4 : 5));
            switch (n) {
                case 1: CPB_EFT_LAUNCH(1); break;
                case 2: CPB_EFT_LAUNCH(2); break;
                case 3: CPB_EFT_LAUNCH(3); break;
                case 4: CPB_EFT_LAUNCH(4); break;
                default: CPB_EFT_LAUNCH(5); break;
            }
            c0 += n;
        }
#undef CPB_EFT_LAUNCH
        cudaFreeAsync(tab, st);
    } else if (vec4) {
        const dim3 grid4(blocks_for((long long)B * oH * (oW / 4), 256));
#define CPB_BLEND_ARGS y, B, ntiles, nch, ly, lx, y0, x0, flip, negate_flow, taper_y, taper_x, cy0, cx0, oH, oW, yf
        if (nch <= 4)      { CPB_LAUNCH_COUNTED(k_average_tiles_v4<4>, grid4, dim3(256), 0, st, CPB_BLEND_ARGS); }
        else if (nch <= 8) { CPB_LAUNCH_COUNTED(k_average_tiles_v4<8>, grid4, dim3(256), 0, st, CPB_BLEND_ARGS); }
        else               { CPB_LAUNCH_COUNTED(k_average_tiles_v4<16>, grid4, dim3(256), 0, st, CPB_BLEND_ARGS); }
#undef CPB_BLEND_ARGS
    } else {
        const long long total = (long long)B * nch * oH * oW;
        CPB_LAUNCH_COUNTED(k_average_tiles, dim3(blocks_for(total, 256)), dim3(256), 0, st, y, B, ntiles, nch, ly, lx, y0,
                           x0, flip, negate_flow, taper_y, taper_x, cy0, cx0, oH, oW, yf);
    }
    CPB_CHECK_LAUNCH();
    return 0;
}

int cpb_eval_tail_device(const float* y_flows, const float* y_logits, int B, int ntiles, int C, int ly, int lx,
                         const int32_t* y0, const int32_t* x0, const int32_t* flip, int augment, const double* taper_y,
                         const double* taper_x, int Ly, int Lx, int cy0, int cy1, int cx0, int cx1, const cpb_params* prm,
                         float* dP, float* cellprob, float* logits, int32_t* masks, int32_t* counts, int32_t* cell_class,
                         uint8_t* class_masks, void* workspace, size_t workspace_bytes, void* stream) {
    if (!y_flows || !y0 || !x0 || !flip || !taper_y || !taper_x || !prm || !dP || !cellprob || !masks || !counts) return CPB_E_ARG;
    if (y_logits && (!logits || !cell_class || C < 1 || C > 255)) return CPB_E_ARG;
    const int H = Ly - cy0 - cy1, W = Lx - cx0 - cx1;
    if (B <= 0 || ntiles <= 0 || ly <= 0 || lx <= 0 || cy0 < 0 || cx0 < 0 || !workspace) return CPB_E_ARG;
    // the fused kernel's geometry promises (the WSI path: 256-px sub-tiles, 8 / 16-px pads): otherwise CPB_E_ARG and the
    // caller composes cpb_average_tiles_ex_device + cpb_compute_masks_device itself
    if (lx % 4 || cx0 % 4 || W % 64 || H < 2 || reinterpret_cast<uintptr_t>(y_flows) % 16 || reinterpret_cast<uintptr_t>(dP) % 16 ||
        reinterpret_cast<uintptr_t>(cellprob) % 16 || reinterpret_cast<uintptr_t>(masks) % 16)
        return CPB_E_ARG;
    { int e_ = check_geom(B, H, W); if (e_) return e_; }
    if ((long long)(H + 2) * (W + 2 * CPB_FLOW_PADX) >= (1LL << 24)) return CPB_E_RANGE;
    ensure_attributes();
    cudaStream_t st0 = reinterpret_cast<cudaStream_t>(stream);
    // weight table and reciprocal normaliser depend on the geometry only: built once, before the batch is cut into parts
    float* tab = nullptr;
    const size_t nw = (size_t)ly * lx, nr = (size_t)H * W, N = nr;
    if (scratch_alloc(reinterpret_cast<void**>(&tab), (2 * nw + 2 * nr) * sizeof(float), st0) != cudaSuccess) return (int)cudaGetLastError();
    float* wh = tab; float* wl = tab + nw; float* rh = tab + 2 * nw; float* rl = rh + nr;
    CPB_LAUNCH_COUNTED(k_blend_weights, dim3(blocks_for((long long)nw, 256)), dim3(256), 0, st0, taper_y, taper_x, ly, lx, wh, wl);
    CPB_LAUNCH_COUNTED(k_blend_rinv, dim3(blocks_for((long long)nr, 256)), dim3(256), 0, st0, ntiles, ly, lx, y0, x0, taper_y, taper_x,
                       cy0, cx0, H, W, rh, rl);
    const int LC = cpb_label_capacity(H, W);
    const size_t tplane = (size_t)ntiles * ly * lx;
    const float sx = (float)(2.0 / (double)(W - 1)), sy = (float)(2.0 / (double)(H - 1));
    // (one part: cutting this call into parallel sub-batches makes the HBM-bound blends of the parts contend --
    //  measured 3.15 ms vs 2.40 ms per 256 TTA tiles -- so only cpb_compute_masks_device forks)
    auto part = [&](int b0, int nb, void* ws, size_t ws_bytes, void* stp) -> int {
        cudaStream_t st = reinterpret_cast<cudaStream_t>(stp);
        const uintptr_t wsa = (reinterpret_cast<uintptr_t>(ws) + kAlign - 1) / kAlign * kAlign;
        Workspace w = carve(reinterpret_cast<void*>(wsa), nb, H, W, y_logits ? C : 0, 0);
        if (w.bytes + (wsa - reinterpret_cast<uintptr_t>(ws)) > ws_bytes) return CPB_E_WORKSPACE;
        float* dPp = dP + (size_t)b0 * 2 * N; float* cpp = cellprob + (size_t)b0 * N;
        float* lgp = y_logits ? logits + (size_t)b0 * C * N : nullptr;
        int32_t* mp = masks + (size_t)b0 * N;
        if (y_logits) {          // class logits: un-flip + blend (transforms/transforms.py:4-21, core.py:218-220)
            const dim3 grid(blocks_for((long long)nb * H * (W / 4), 256));
            const float* ysrc = y_logits + (size_t)b0 * C * tplane;
#define CPB_EFT_LAUNCH(Nc) CPB_LAUNCH_COUNTED(k_average_tiles_eft<Nc>, grid, dim3(256), 0, st, ysrc, nb, ntiles, C, c0, ly, lx, y0, x0, flip, \
                           0, (const float*)wh, (const float*)wl, (const float*)rh, (const float*)rl, cy0, cx0, H, W, lgp)
            for (int c0 = 0; c0 < C;) {
                const int rem = C - c0;
                const int n = rem <= 5 ? rem : (rem == 6 ? 3 : (rem == 7 ? 4 : 5));
                switch (n) {
                    case 1: CPB_EFT_LAUNCH(1); break;
                    case 2: CPB_EFT_LAUNCH(2); break;
                    case 3: CPB_EFT_LAUNCH(3); break;
                    case 4: CPB_EFT_LAUNCH(4); break;
                    default: CPB_EFT_LAUNCH(5); break;
                }
                c0 += n;
            }
#undef CPB_EFT_LAUNCH
        }
        cudaMemsetAsync(w.list_n, 0, 64 * sizeof(unsigned), st);
        cudaMemsetAsync(w.t.fail, 0, nb * sizeof(int), st);
        const long long nblk = (long long)nb * ((H + 15) / 16) * (W / 64);
        CPB_LAUNCH_COUNTED(k_blend_prep, dim3((unsigned)nblk), dim3(256), 0, st, y_flows + (size_t)b0 * 3 * tplane, nb, ntiles, ly, lx, y0, x0,
                           flip, augment ? 1 : 0, (const float*)wh, (const float*)wl, (const float*)rh, (const float*)rl, cy0, cx0, H, W,
                           prm->cellprob_threshold, sx, sy, dPp, cpp, reinterpret_cast<float4*>(w.flow), reinterpret_cast<int4*>(mp),
                           w.list, w.list_n);
        CPB_CHECK_LAUNCH();
        return compute_masks_impl(dPp, cpp, lgp, nb, H, W, C, prm, mp, counts + b0, cell_class ? cell_class + (size_t)b0 * LC : nullptr,
                                  class_masks ? class_masks + (size_t)b0 * N : nullptr, ws, ws_bytes, stp, nullptr, nullptr, true);
    };
    const int rc = part(0, B, workspace, workspace_bytes, stream);
    cudaFreeAsync(tab, st0);
    return rc;
}

int cpb_label_offsets_device(const int32_t* counts, int B, int64_t base, int64_t* offsets, int64_t* total,
                             void* stream) {
    if (!counts || !offsets || B <= 0) return CPB_E_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    CPB_LAUNCH_COUNTED(k_label_offsets, dim3(1), dim3(256), 0, st, counts, B, (long long)base,
               reinterpret_cast<long long*>(offsets), reinterpret_cast<long long*>(total));
    CPB_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"

#ifndef CPB_SIM
#include "cpb_host.inl"
#include "cpb_plan.inl"
#endif
