// Study (GPU): bandwidth of kernel reads from PINNED HOST memory over PCIe as a function of the access pattern.
// The host-buffer path reads the class logits in place (cpb_host.inl); this measures what a sector-granular,
// run-structured pattern costs against full lines and against cudaMemcpyAsync.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/zc tests/studies/zerocopy_bw.cu && /tmp/zc
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

// every lane owns one float4 (16 B) of a contiguous stream; `mask` says which lanes of each group of 8 lanes
// (= one 128-byte line) load.  pattern: bit i of `mask8` = lane i of the line participates.
__global__ void k_read(const float4* __restrict__ src, long long n4, unsigned mask8, float* sink, int run_mode, int run_len,
                       int gap_len) {
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        bool on;
        if (run_mode) on = (int)(i % (run_len + gap_len)) < run_len;     // runs of run_len groups, gap_len groups apart
        else on = (mask8 >> (i & 7)) & 1;
        if (on) { const float4 v = src[i]; acc += v.x + v.y + v.z + v.w; }
    }
    if (acc == 123.456f) *sink = acc;
}

int main() {
    const size_t bytes = 1ull << 30;
    float4* h; cudaHostAlloc(&h, bytes, cudaHostAllocDefault);
    for (size_t i = 0; i < bytes / 16; i += 4096) h[i] = make_float4(1, 2, 3, 4);
    float4* d; cudaMalloc(&d, bytes);
    float* sink; cudaMalloc(&sink, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int r = 0; r < 2; r++) { cudaEventRecord(e0); cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice); cudaEventRecord(e1); cudaEventSynchronize(e1); }
    cudaEventElapsedTime(&ms, e0, e1);
    printf("cudaMemcpyAsync H2D pinned        : %6.1f GB/s\n", bytes / ms / 1e6);
    const long long n4 = bytes / 16;
    struct { const char* name; unsigned mask; int run, len, gap; double frac; } tests[] = {
        {"all 8 lanes of every line (128 B)", 0xff, 0, 0, 0, 1.0},
        {"4 lanes = sectors 0,1 (64 B)     ", 0x0f, 0, 0, 0, 0.5},
        {"2 lanes = sector 0 (32 B)        ", 0x03, 0, 0, 0, 0.25},
        {"1 lane  = half a sector (16 B)   ", 0x01, 0, 0, 0, 0.125},
        {"sectors 0 and 2 (2 x 32 B)       ", 0x33, 0, 0, 0, 0.5},
        {"runs of 4 groups every 16        ", 0, 1, 4, 12, 0.25},
        {"runs of 5 groups every 18        ", 0, 1, 5, 13, 5.0 / 18},
        {"runs of 8 groups every 32        ", 0, 1, 8, 24, 0.25},
        {"runs of 16 groups every 64       ", 0, 1, 16, 48, 0.25},
    };
    for (auto& t : tests) {
        for (int r = 0; r < 2; r++) {
            cudaEventRecord(e0);
            k_read<<<148 * 8, 256>>>(h, n4, t.mask, n4 ? sink : nullptr, t.run, t.len, t.gap);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        cudaEventElapsedTime(&ms, e0, e1);
        printf("%s: %7.2f ms, requested %6.1f GB/s, span-equivalent %6.1f GB/s\n", t.name, ms, bytes * t.frac / ms / 1e6, bytes / ms / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
