// Study: what does a packed FP32 instruction (FFMA2 / FADD2 / FMUL2, PTX *.f32x2) cost in issue slots on sm_100a?
// Every warp runs the same unrolled loop of independent accumulator chains; 16 warps per SM sub-partition hide
// the 4-cycle latency, so the time of a row relative to the scalar row is the issue / pipe cost of the instruction mix.
// Measured on B200 (profiles/r02/f32x2_issue.txt): 8 x FFMA 1.370 ms, 8 x FFMA2 2.708 ms (1.98 x), 8 x FADD2 / 8 x FADD
// 1.96 x, 4 x FFMA + 4 x FMNMX 1.122 ms (the alu pipe runs beside the fma pipe).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_issue f32x2_issue.cu && ./f32x2_issue
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float add1(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float mnmx(float a, float b) { float r; asm volatile("min.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

#define CHAINS 8
// MODE 0: 8 scalar FFMA per iteration; 1: 8 FFMA2; 2: 4 FFMA2 + 4 FFMA; 3: 8 scalar FADD; 4: 8 FADD2;
// 5: 4 FFMA + 4 FMNMX (fma pipe + alu pipe); 6: 4 FFMA2 + 4 FMNMX
template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, long long* cyc, int iters, float s) {
    float a[CHAINS]; u64 p[CHAINS];
    for (int i = 0; i < CHAINS; i++) { a[i] = s * (threadIdx.x + i); p[i] = pk(a[i], a[i] + 1.f); }
    const float m = s * 0.999f; const u64 m2 = pk(m, m);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            if (MODE == 0) a[i] = fma1(a[i], m, a[i]);
            if (MODE == 1) p[i] = fma2(p[i], m2, p[i]);
            if (MODE == 2) { if (i & 1) a[i] = fma1(a[i], m, a[i]); else p[i] = fma2(p[i], m2, p[i]); }
            if (MODE == 3) a[i] = add1(a[i], m);
            if (MODE == 4) p[i] = add2(p[i], m2);
            if (MODE == 5) { if (i & 1) a[i] = fma1(a[i], m, a[i]); else a[i] = mnmx(a[i], m); }
            if (MODE == 6) { if (i & 1) a[i] = mnmx(a[i], m); else p[i] = fma2(p[i], m2, p[i]); }
        }
    }
    const long long t1 = clock64();
    float r = 0; for (int i = 0; i < CHAINS; i++) { r += a[i]; r += (float)(p[i] & 0xffff); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char* name, int sms) {
    const int threads = 512, blocks = sms * 4, iters = 20000;      // 2048 threads / SM = 16 warps per sub-partition
    float* out; long long* cyc; cudaMalloc(&out, sizeof(float) * threads * blocks); cudaMalloc(&cyc, sizeof(long long) * blocks);
    k<MODE><<<blocks, threads>>>(out, cyc, 100, 1e-3f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(out, cyc, iters, 1e-3f); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long* h = new long long[blocks]; cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < blocks; i++) mean += (double)h[i]; mean /= blocks;
    // per sub-partition: 16 warps x iters x CHAINS instructions; the event time is the measurement (clock64 is printed
    // for reference only), compare the rows with the FFMA row
    const double instr = 16.0 * iters * CHAINS;
    printf("%-34s %8.3f ms  (%.3f clock64 ticks per warp instruction per sub-partition)\n", name, ms, mean / instr);
    cudaFree(out); cudaFree(cyc); delete[] h;
}
int main() {
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    run<0>("8 x FFMA", sms); run<1>("8 x FFMA2", sms); run<2>("4 x FFMA2 + 4 x FFMA", sms);
    run<3>("8 x FADD", sms); run<4>("8 x FADD2", sms); run<5>("4 x FFMA + 4 x FMNMX", sms); run<6>("4 x FFMA2 + 4 x FMNMX", sms);
    return 0;
}
