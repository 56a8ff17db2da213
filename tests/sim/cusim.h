// cusim -- a TEST-ONLY model of a CUDA grid on the CPU.
//
// Purpose: the build box has no GPU.  The kernels under classpose_b200/csrc are written
// against a small macro layer (cpb_platform.h); compiled with -DCPB_SIM they run here, one
// cooperative fiber per CUDA thread, so that their *logic* (indexing, barriers, warp
// collectives, atomics, table handling) can be checked against the oracle before any GPU
// time is spent.  It models semantics, not performance, and it is never loaded by the
// classpose_b200 package -- only by tests/ (see tests/sim/build_sim.py).
//
// Model: blocks are distributed over OS worker threads; inside a block every CUDA thread is
// a fiber on the worker's stack pool, switched cooperatively at __syncthreads() and at
// warp-synchronous intrinsics.  Global-memory atomics are real atomics (blocks run
// concurrently); everything inside a block is sequentially interleaved.
#pragma once

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>
#include <sys/mman.h>

#if !defined(__x86_64__)
#error "cusim's context switch is written for x86-64"
#endif

// ---------------------------------------------------------------- vector types / dims
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
static inline float2 make_float2(float x, float y) { return {x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
static inline int4 make_int4(int x, int y, int z, int w) { return {x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return {x, y}; }
static inline double2 make_double2(double x, double y) { return {x, y}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaMallocAsync(void** p, size_t n, cudaStream_t) { *p = malloc(n); return *p ? 0 : 2; }
static inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { free(p); return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "cusim"; }

extern "C" void cusim_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl cusim_switch
.type cusim_switch,@function
cusim_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size cusim_switch,.-cusim_switch
)");

namespace cusim {

constexpr size_t kStack = 96 * 1024;
constexpr int kMaxThreads = 1024;

struct WarpState {
    uint64_t slot[32];
    uint32_t want[32];          // mask of the collective each arrived lane is waiting in
    uint32_t arrived = 0, left = 0;
};

struct Worker {
    char* stacks = nullptr;
    void* sched_sp = nullptr;
    void* fiber_sp[kMaxThreads];
    bool done[kMaxThreads];
    dim3 tid[kMaxThreads];
    const char* wait_what[kMaxThreads];
    unsigned wait_mask[kMaxThreads];
    int cur = 0, nthreads = 0, live = 0;
    int bar_arrived = 0;
    unsigned bar_gen = 0;
    uint64_t progress = 0;
    WarpState warps[kMaxThreads / 32];
    std::vector<unsigned char> smem;
    const std::function<void()>* body = nullptr;
    dim3 block, grid, bid;
};

inline thread_local Worker* tl_worker = nullptr;

}  // namespace cusim

inline thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
static const int warpSize = 32;

namespace cusim {

inline void* dyn_smem() { return tl_worker->smem.data(); }

inline void yield() {
    Worker* w = tl_worker;
    int me = w->cur;
    cusim_switch(&w->fiber_sp[me], w->sched_sp);
    threadIdx = w->tid[me];
}

inline void release_barrier_if_complete(Worker* w) {
    if (w->bar_arrived > 0 && w->bar_arrived >= w->live) {
        w->bar_arrived = 0;
        w->bar_gen++;
        w->progress++;
    }
}

inline void fiber_main() {
    Worker* w = tl_worker;
    threadIdx = w->tid[w->cur];
    (*w->body)();
    w = tl_worker;
    w->done[w->cur] = true;
    w->live--;
    w->progress++;
    release_barrier_if_complete(w);
    void* dummy;
    cusim_switch(&dummy, w->sched_sp);
    abort();
}

inline const char*& tl_kernel_name();

inline void run_block(Worker* w) {
    const int n = w->nthreads;
    for (int i = 0; i < n; i++) {
        char* top = w->stacks + (size_t)(i + 1) * kStack;
        void** sp = reinterpret_cast<void**>(top);
        *--sp = nullptr;                              // fake return address of fiber_main
        *--sp = reinterpret_cast<void*>(&fiber_main); // popped by ret in cusim_switch
        for (int r = 0; r < 6; r++) *--sp = nullptr;  // rbp rbx r12..r15
        w->fiber_sp[i] = sp;
        w->done[i] = false;
        unsigned t = i;
        w->tid[i] = dim3(t % w->block.x, (t / w->block.x) % w->block.y, t / (w->block.x * w->block.y));
    }
    w->live = n;
    w->bar_arrived = 0;
    for (auto& ws : w->warps) ws.arrived = ws.left = 0;
    blockIdx = w->bid; blockDim = w->block; gridDim = w->grid;
    while (w->live > 0) {
        uint64_t before = w->progress;
        for (int i = 0; i < n; i++) {
            if (w->done[i]) continue;
            w->cur = i;
            cusim_switch(&w->sched_sp, w->fiber_sp[i]);
        }
        if (w->live > 0 && w->progress == before) {
            fprintf(stderr, "cusim: deadlock in %s block (%u,%u,%u): %d live threads, %d at barrier\n", tl_kernel_name(),
                    w->bid.x, w->bid.y, w->bid.z, w->live, w->bar_arrived);
            for (int i = 0; i < n; i++)
                if (!w->done[i]) fprintf(stderr, "  thread %d: %s mask=%08x  warp arrived=%08x left=%08x\n", i,
                                         w->wait_what[i] ? w->wait_what[i] : "?", w->wait_mask[i],
                                         w->warps[i >> 5].arrived, w->warps[i >> 5].left);
            abort();
        }
    }
}

inline const char*& tl_kernel_name() { static const char* n = "?"; return n; }

inline int num_workers() {
    static int n = [] {
        const char* e = getenv("CUSIM_THREADS");
        int v = e ? atoi(e) : (int)std::thread::hardware_concurrency();
        return std::max(1, v);
    }();
    return n;
}

template <class F>
inline void launch(const char* name, dim3 grid, dim3 block, size_t smem, F f) {
    tl_kernel_name() = name;
    std::function<void()> body = f;
    const size_t nblocks = (size_t)grid.x * grid.y * grid.z;
    const int nthreads = (int)(block.x * block.y * block.z);
    if (nthreads <= 0 || nthreads > kMaxThreads || nblocks == 0) { fprintf(stderr, "cusim: bad launch\n"); abort(); }
    std::atomic<size_t> next{0};
    auto work = [&]() {
        Worker* w = new Worker();
        w->stacks = (char*)mmap(nullptr, kStack * nthreads, PROT_READ | PROT_WRITE,
                                MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (w->stacks == MAP_FAILED) { perror("cusim mmap"); abort(); }
        w->smem.assign(smem + 16, 0);
        w->body = &body; w->block = block; w->grid = grid; w->nthreads = nthreads;
        tl_worker = w;
        for (;;) {
            size_t b = next.fetch_add(1);
            if (b >= nblocks) break;
            w->bid = dim3((unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((size_t)grid.x * grid.y)));
            run_block(w);
        }
        munmap(w->stacks, kStack * nthreads);
        tl_worker = nullptr;
        delete w;
    };
    int nw = (int)std::min<size_t>(num_workers(), nblocks);
    if (nw <= 1) { std::thread t(work); t.join(); return; }
    std::vector<std::thread> ts;
    for (int i = 0; i < nw; i++) ts.emplace_back(work);
    for (auto& t : ts) t.join();
}

// ---- warp collectives ---------------------------------------------------------------
template <class R, class Fn>
inline R warp_collective(unsigned mask, uint64_t v, Fn fn) {
    Worker* w = tl_worker;
    int t = w->cur;
    int lane = t & 31;
    WarpState& ws = w->warps[t >> 5];
    int nl = std::min(32, w->nthreads - (t & ~31));
    unsigned exist = nl == 32 ? 0xffffffffu : ((1u << nl) - 1);
    mask &= exist;
    unsigned bit = 1u << lane;
    if (!(mask & bit)) { fprintf(stderr, "cusim: lane %d not in mask %08x\n", lane, mask); abort(); }
    for (int l = 0; l < nl; l++)
        if ((mask >> l & 1) && w->done[(t & ~31) + l]) { fprintf(stderr, "cusim: mask %08x names exited lane %d in %s (thread %d)\n", mask, l, tl_kernel_name(), t); abort(); }
    w->wait_what[t] = "drain"; w->wait_mask[t] = mask;
    while (ws.arrived & bit) yield();  // previous collective of this lane not drained yet
    ws.slot[lane] = v;
    ws.want[lane] = mask;
    ws.arrived |= bit;
    w->progress++;
    w->wait_what[t] = "arrive";
    // complete when every lane of the mask has arrived *in this same collective* (lanes of a warp can be
    // inside different collectives at the same time, e.g. some in a subset reduce, others already waiting
    // in the next full-warp ballot)
    auto complete = [&]() {
        if ((ws.arrived & mask) != mask) return false;
        for (int l2 = 0; l2 < 32; l2++) if ((mask >> l2 & 1) && ws.want[l2] != mask) return false;
        return true;
    };
    while (!complete()) yield();
    R r = fn(ws.slot, mask, lane);
    ws.left |= bit;
    if ((ws.left & mask) == mask) { ws.arrived &= ~mask; ws.left &= ~mask; w->progress++; }
    return r;
}

template <class T> inline uint64_t to_bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, ""); memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

}  // namespace cusim

inline void __syncthreads() {
    cusim::Worker* w = cusim::tl_worker;
    unsigned gen = w->bar_gen;
    w->bar_arrived++;
    if (w->bar_arrived >= w->live) { w->bar_arrived = 0; w->bar_gen++; w->progress++; return; }
    w->wait_what[w->cur] = "barrier";
    while (w->bar_gen == gen) cusim::yield();
}
inline void __syncwarp(unsigned mask = 0xffffffffu) {
    cusim::warp_collective<int>(mask, 0, [](uint64_t*, unsigned, int) { return 0; });
}
template <class T> inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    return cusim::warp_collective<T>(mask, cusim::to_bits(v), [=](uint64_t* s, unsigned, int lane) {
        int base = lane & ~(width - 1);
        return cusim::from_bits<T>(s[base + (src & (width - 1))]); });
}
template <class T> inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32) {
    return cusim::warp_collective<T>(mask, cusim::to_bits(v), [=](uint64_t* s, unsigned, int lane) {
        int base = lane & ~(width - 1); int src = lane - (int)d;
        return cusim::from_bits<T>(s[src < base ? lane : src]); });
}
template <class T> inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32) {
    return cusim::warp_collective<T>(mask, cusim::to_bits(v), [=](uint64_t* s, unsigned, int lane) {
        int base = lane & ~(width - 1); int src = lane + (int)d;
        return cusim::from_bits<T>(s[src >= base + width ? lane : src]); });
}
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int x, int width = 32) {
    return cusim::warp_collective<T>(mask, cusim::to_bits(v), [=](uint64_t* s, unsigned, int lane) {
        return cusim::from_bits<T>(s[lane ^ x]); });
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
    return cusim::warp_collective<unsigned>(mask, pred ? 1 : 0, [](uint64_t* s, unsigned m, int) {
        unsigned r = 0; for (int l = 0; l < 32; l++) if ((m >> l & 1) && s[l]) r |= 1u << l; return r; });
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) {
    return cusim::warp_collective<int>(mask, pred ? 1 : 0, [](uint64_t* s, unsigned m, int) {
        for (int l = 0; l < 32; l++) if ((m >> l & 1) && !s[l]) return 0; return 1; });
}
template <class T> inline unsigned __match_any_sync(unsigned mask, T v) {
    return cusim::warp_collective<unsigned>(mask, cusim::to_bits(v), [](uint64_t* s, unsigned m, int lane) {
        unsigned r = 0; for (int l = 0; l < 32; l++) if ((m >> l & 1) && s[l] == s[lane]) r |= 1u << l; return r; });
}
inline int __reduce_add_sync(unsigned mask, int v) {
    return cusim::warp_collective<int>(mask, cusim::to_bits(v), [](uint64_t* s, unsigned m, int) {
        int r = 0; for (int l = 0; l < 32; l++) if (m >> l & 1) r += cusim::from_bits<int>(s[l]); return r; });
}
inline unsigned __reduce_add_sync(unsigned mask, unsigned v) { return (unsigned)__reduce_add_sync(mask, (int)v); }
inline int __reduce_min_sync(unsigned mask, int v) {
    return cusim::warp_collective<int>(mask, cusim::to_bits(v), [](uint64_t* s, unsigned m, int) {
        int r = INT32_MAX; for (int l = 0; l < 32; l++) if (m >> l & 1) r = std::min(r, cusim::from_bits<int>(s[l])); return r; });
}
inline int __reduce_max_sync(unsigned mask, int v) {
    return cusim::warp_collective<int>(mask, cusim::to_bits(v), [](uint64_t* s, unsigned m, int) {
        int r = INT32_MIN; for (int l = 0; l < 32; l++) if (m >> l & 1) r = std::max(r, cusim::from_bits<int>(s[l])); return r; });
}
inline unsigned __reduce_or_sync(unsigned mask, unsigned v) {
    return cusim::warp_collective<unsigned>(mask, v, [](uint64_t* s, unsigned m, int) {
        unsigned r = 0; for (int l = 0; l < 32; l++) if (m >> l & 1) r |= (unsigned)s[l]; return r; });
}

// ---- atomics (blocks run on different OS threads) ------------------------------------
template <class T> inline T cusim_atomic_add(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline int atomicAdd(int* p, int v) { return cusim_atomic_add(p, v); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return cusim_atomic_add(p, v); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return cusim_atomic_add(p, v); }
template <class T> inline T cusim_atomic_fadd(T* p, T v) {
    T old = *p, neu;
    do { neu = old + v; } while (!__atomic_compare_exchange(p, &old, &neu, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    return old;
}
inline float atomicAdd(float* p, float v) { return cusim_atomic_fadd(p, v); }
inline double atomicAdd(double* p, double v) { return cusim_atomic_fadd(p, v); }
template <class T, class Op> inline T cusim_atomic_rmw(T* p, T v, Op op) {
    T old = __atomic_load_n(p, __ATOMIC_RELAXED);
    for (;;) { T neu = op(old, v); if (__atomic_compare_exchange_n(p, &old, neu, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return old; }
}
inline int atomicMin(int* p, int v) { return cusim_atomic_rmw(p, v, [](int a, int b) { return std::min(a, b); }); }
inline int atomicMax(int* p, int v) { return cusim_atomic_rmw(p, v, [](int a, int b) { return std::max(a, b); }); }
inline unsigned atomicMin(unsigned* p, unsigned v) { return cusim_atomic_rmw(p, v, [](unsigned a, unsigned b) { return std::min(a, b); }); }
inline unsigned atomicMax(unsigned* p, unsigned v) { return cusim_atomic_rmw(p, v, [](unsigned a, unsigned b) { return std::max(a, b); }); }
inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) { return cusim_atomic_rmw(p, v, [](unsigned long long a, unsigned long long b) { return std::max(a, b); }); }
inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) { return cusim_atomic_rmw(p, v, [](unsigned long long a, unsigned long long b) { return std::min(a, b); }); }
inline int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
inline int atomicCAS(int* p, int cmp, int v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED); return cmp; }
inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long cmp, unsigned long long v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED); return cmp; }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
inline long long __double_as_longlong(double d) { long long v; memcpy(&v, &d, 8); return v; }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_block() {}

// ---- misc intrinsics ----------------------------------------------------------------
template <class T> inline T __ldg(const T* p) { return *p; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; i++) if (v >> i & 1) r |= 1u << (31 - i); return r; }
// round-to-nearest single ops that must not be contracted (build with -ffp-contract=off)
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __dsqrt_rn(double a) { return std::sqrt(a); }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline int __float2int_rz(float f) { return (int)f; }
inline float __int2float_rn(int i) { return (float)i; }
inline double __int2double_rn(int i) { return (double)i; }
inline double __ll2double_rn(long long i) { return (double)i; }
inline double __ull2double_rn(unsigned long long i) { return (double)i; }
inline float __ldcs(const float* p) { return *p; }
using std::min;
using std::max;
