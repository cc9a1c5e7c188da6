// microbench.cuh -- instruction-issue ceilings measured on the device the library runs on (BASELINE.md section 2 asks
// for the fp64 rate before K2 is called HBM-bound; K2's bin function now runs in packed fp32, so the fp32 / packed-fp32 /
// integer rates are what bound it beside HBM).  One CTA of 1024 threads per SM (8 warps per scheduler), every thread
// runs `iters` trips of 8 independent dependency chains x 4 operations; the kernel times itself with clock64().
#pragma once
#include "common.cuh"

namespace cdfgpu {

enum { kMbDfma = 0, kMbFfma = 1, kMbFfma2 = 2, kMbIadd = 3, kMbMixed = 4 };

template <int KIND>
__global__ void __launch_bounds__(1024, 1) microbench_kernel(int iters, float seed, unsigned long long *cycles, double *sink)
{
    const long long t0 = clock64();
    double out = 0.0;
    if (KIND == kMbDfma) {
        double a[8], m = 1.0 + 1e-9 * seed, c = 1e-9 * seed;
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = seed + i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, c);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) out += a[i];
    } else if (KIND == kMbFfma) {
        float a[8], m = 1.0f + 1e-6f * seed, c = 1e-6f * seed;
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = seed + i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = __fmaf_rn(a[i], m, c);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) out += a[i];
    } else if (KIND == kMbFfma2) {
        unsigned long long a[8], m, c;
        asm("mov.b64 %0, {%1, %1};" : "=l"(m) : "f"(1.0f + 1e-6f * seed));
        asm("mov.b64 %0, {%1, %1};" : "=l"(c) : "f"(1e-6f * seed));
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %1};" : "=l"(a[i]) : "f"(seed + i));
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(m), "l"(c));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float lo, hi;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[i]));
            out += lo + hi;
        }
    } else if (KIND == kMbIadd) {
        unsigned a[8], m = 0x9e3779b9u ^ (unsigned)seed;
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = (unsigned)seed + i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = (a[i] ^ m) + (a[i] >> 3);   // LOP3 + SHF/IADD: the integer pipe
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) out += a[i];
    } else {   // alternating fp32 FMA and integer operations (two pipes)
        float a[4], m = 1.0f + 1e-6f * seed, c = 1e-6f * seed;
        unsigned b[4], x = 0x9e3779b9u ^ (unsigned)seed;
#pragma unroll
        for (int i = 0; i < 4; ++i) { a[i] = seed + i; b[i] = (unsigned)seed + i; }
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 4; ++i) { a[i] = __fmaf_rn(a[i], m, c); b[i] = (b[i] ^ x) + 0x1234567u; }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) out += a[i] + b[i];
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (out == 12345.678) sink[0] = out;   // keeps the chains alive
}

}  // namespace cdfgpu
