// CPU twin of div5_exhaustive.cu (x86 FMA hardware is IEEE like the GPU): gcc -O2 -mfma -ffp-contract=off -fopenmp div5_exhaustive_cpu.c -lm
// Result in this container: 0 mismatches over all 2^32 float32 bit patterns (NaNs excluded).
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <omp.h>
static inline float div5(float x) {
    float ax = fabsf(x);
    if (!(ax >= 1e-30f && ax <= 1e30f)) return x / 5.0f;
    const float r = 0.2f;
    float q = x * r;
    float rem = fmaf(-5.0f, q, x);
    return fmaf(rem, r, q);
}
int main() {
    unsigned long long bad = 0; 
    #pragma omp parallel for reduction(+:bad) schedule(static)
    for (long long i = 0; i < (1LL << 32); i++) {
        uint32_t u = (uint32_t)i; float x; memcpy(&x, &u, 4);
        if (x != x) continue;
        volatile float a = x / 5.0f; float b = div5(x);
        float aa = a; uint32_t ua, ub; memcpy(&ua, &aa, 4); memcpy(&ub, &b, 4);
        if (ua != ub) bad++;
    }
    printf("cpu exhaustive div5: %llu mismatches\n", bad);
    return bad != 0;
}
