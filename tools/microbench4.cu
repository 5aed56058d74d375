// Microbenchmark 4: candidate inner loops for the clash kernel, scheduled by ptxas (no volatile).
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0,{%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0,%1,%2,%3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float min3(float a, float b, float c) { float d; asm("min.f32 %0,%1,%2,%3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ void up(u64 v, float& lo, float& hi) { asm("mov.b64 {%0,%1},%2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 movv(u64 a) { u64 d; asm volatile("mov.b64 %0,%1;" : "=l"(d) : "l"(a)); return d; }

// VAR 1: v1  A dup'd in smem (2 LDS.128/atom), B packed pairs in regs, min3 across two A atoms
// VAR 2: v2  A natural pairs in smem (2 LDS.128 / 2 atoms), B dup'd in regs, min3(m, e.lo, e.hi)
// VAR 3: scalar FFMA, A {x,y,z,n} 1 LDS.128/atom, min3 across two A atoms
// VAR 4: v1 with A operands moved to fresh registers
// VAR 5: v2 without mins (sum instead, FMA pipe only) -- upper bound probe
template <int VAR, int TB, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k(float* out, int n_a, int reps, float seed) {
    extern __shared__ ulonglong2 sA[];
    for (int i = threadIdx.x; i < n_a * 2 + 8; i += blockDim.x) {
        float v = 1e-3f * (i + 1);
        sA[i] = make_ulonglong2(pk(v, v * 1.0001f), pk(-v, v * 0.5f));
    }
    __syncthreads();
    float res = 0.f;
    if (VAR == 1 || VAR == 4) {
        u64 bx[TB / 2], by[TB / 2], bz[TB / 2]; float m[TB];
        for (int q = 0; q < TB / 2; ++q) { bx[q] = pk(seed + q + threadIdx.x * 1e-3f, seed - q); by[q] = pk(seed * .5f + q, seed * .25f - q); bz[q] = pk(seed * .125f + q, seed * .0625f - q); m[2 * q] = m[2 * q + 1] = 3e38f; }
        for (int rep = 0; rep < reps; ++rep) {
#pragma unroll 2
            for (int i = 0; i < n_a; i += 2) {
                ulonglong2 u0 = sA[2 * i], u1 = sA[2 * i + 1], v0 = sA[2 * i + 2], v1 = sA[2 * i + 3];
                if (VAR == 4) { u0.x = movv(u0.x); u0.y = movv(u0.y); u1.x = movv(u1.x); u1.y = movv(u1.y); v0.x = movv(v0.x); v0.y = movv(v0.y); v1.x = movv(v1.x); v1.y = movv(v1.y); }
#pragma unroll
                for (int q = 0; q < TB / 2; ++q) {
                    u64 e0 = fma2(u0.x, bx[q], fma2(u0.y, by[q], fma2(u1.x, bz[q], u1.y)));
                    u64 e1 = fma2(v0.x, bx[q], fma2(v0.y, by[q], fma2(v1.x, bz[q], v1.y)));
                    float a, b, c, d; up(e0, a, b); up(e1, c, d);
                    m[2 * q] = min3(m[2 * q], a, c); m[2 * q + 1] = min3(m[2 * q + 1], b, d);
                }
            }
        }
        for (int j = 0; j < TB; ++j) res += m[j];
    } else if (VAR == 2 || VAR == 5) {
        u64 bx[TB], by[TB], bz[TB]; float m[TB];
        for (int q = 0; q < TB; ++q) { float x = seed + q + threadIdx.x * 1e-3f; bx[q] = pk(x, x); by[q] = pk(x * .5f, x * .5f); bz[q] = pk(x * .25f, x * .25f); m[q] = 3e38f; }
        for (int rep = 0; rep < reps; ++rep) {
#pragma unroll 2
            for (int i = 0; i < n_a; i += 2) {
                ulonglong2 u0 = sA[i], u1 = sA[i + 1];
#pragma unroll
                for (int q = 0; q < TB; ++q) {
                    u64 e = fma2(u0.x, bx[q], fma2(u0.y, by[q], fma2(u1.x, bz[q], u1.y)));
                    float a, b; up(e, a, b);
                    if (VAR == 2) m[q] = min3(m[q], a, b); else { m[q] += a; m[q] += b; }
                }
            }
        }
        for (int j = 0; j < TB; ++j) res += m[j];
    } else if (VAR == 3) {
        float bx[TB], by[TB], bz[TB], m[TB];
        for (int q = 0; q < TB; ++q) { float x = seed + q + threadIdx.x * 1e-3f; bx[q] = x; by[q] = x * .5f; bz[q] = x * .25f; m[q] = 3e38f; }
        const float4* sF = reinterpret_cast<const float4*>(sA);
        for (int rep = 0; rep < reps; ++rep) {
#pragma unroll 2
            for (int i = 0; i < n_a; i += 2) {
                float4 u = sF[i], v = sF[i + 1];
#pragma unroll
                for (int q = 0; q < TB; ++q) {
                    float e0 = fmaf(u.x, bx[q], fmaf(u.y, by[q], fmaf(u.z, bz[q], u.w)));
                    float e1 = fmaf(v.x, bx[q], fmaf(v.y, by[q], fmaf(v.z, bz[q], v.w)));
                    m[q] = min3(m[q], e0, e1);
                }
            }
        }
        for (int j = 0; j < TB; ++j) res += m[j];
    }
    if (res == 12345.678f) out[0] = res;
}

template <int VAR, int TB, int THREADS, int MINB>
void run(const char* name) {
    float* d; cudaMalloc(&d, 16);
    int n_a = 150, reps = 200, grid = 148 * MINB;
    size_t smem = (n_a * 2 + 8) * 16;
    auto kern = k<VAR, TB, THREADS, MINB>;
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<grid, THREADS, smem>>>(d, n_a, reps / 10, 1.f);
    cudaEventRecord(e0);
    kern<<<grid, THREADS, smem>>>(d, n_a, reps, 1.f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fmas = 3.0 * TB * n_a * (double)reps * THREADS * grid;
    printf("%-34s TB=%2d regs=%3d thr=%3d x%d (%2d warps/SM) %7.3f ms %6.2f TFLOP/s (%5.1f%%)\n", name, TB, fa.numRegs, THREADS, MINB, THREADS * MINB / 32, ms,
           2 * fmas / ms / 1e9, 2 * fmas / ms / 1e9 / 74.45 * 100);
    cudaFree(d);
}
int main() {
    run<1, 10, 480, 1>("v1 dupA f32x2");
    run<1, 10, 256, 2>("v1 dupA f32x2");
    run<1, 8, 256, 2>("v1 dupA f32x2");
    run<1, 16, 128, 3>("v1 dupA f32x2");
    run<4, 10, 256, 2>("v1 + mov A");
    run<2, 5, 256, 3>("v2 natA dupB f32x2");
    run<2, 8, 256, 2>("v2 natA dupB f32x2");
    run<2, 10, 256, 2>("v2 natA dupB f32x2");
    run<2, 15, 128, 3>("v2 natA dupB f32x2");
    run<5, 10, 256, 2>("v2 no-min (adds)");
    run<3, 8, 256, 3>("v3 scalar");
    run<3, 10, 256, 3>("v3 scalar");
    run<3, 16, 256, 2>("v3 scalar");
    run<3, 16, 128, 4>("v3 scalar");
    run<3, 24, 128, 3>("v3 scalar");
    return 0;
}
