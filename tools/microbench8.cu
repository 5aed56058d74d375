// Microbenchmark 8: FFMA2 operand-source matrix (what sustains 1 FFMA2 / 2 cycles?)
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0,{%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0,%1,%2,%3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ void up(u64 v, float& lo, float& hi) { asm("mov.b64 {%0,%1},%2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
#define NK 12
// S0: acc_k = fma2(acc_k, Aconst(UR), Bconst)                    (reference: 98%)
// S1: acc_k = fma2(Areg(vector pair, same for all k), b_k.F32, acc_k)
// S2: acc_k = fma2(A_UR (uniform), b_k.F32, acc_k)
// S3: acc_k = fma2(Areg, b.F32 (same b), acc_k)
// S4: acc_k = fma2(Areg_k (distinct pairs), b.F32 same, acc_k)
// S5: acc_k = fma2(A_UR, bpair_k (64-bit distinct), acc_k)
template <int S>
__global__ void __launch_bounds__(256) k(float* out, int iters, float seed, const float* __restrict__ g) {
    u64 acc[NK], ap[NK], bp[NK]; float b[NK];
    for (int i = 0; i < NK; ++i) { acc[i] = pk(seed + i, seed - i); b[i] = g[threadIdx.x + i]; ap[i] = pk(g[threadIdx.x + 32 + i], g[threadIdx.x + 64 + i]); bp[i] = pk(g[threadIdx.x + 96 + i], g[threadIdx.x + 128 + i]); }
    u64 aU = pk(1.0000001f + seed * 1e-9f, 0.9999999f);           // uniform (from param)
    u64 aR = pk(g[threadIdx.x + 200], g[threadIdx.x + 201]);      // per-thread vector pair
    float bs = g[threadIdx.x + 300];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NK; ++i) {
            if (S == 0) acc[i] = fma2(acc[i], aU, aR);
            if (S == 1) acc[i] = fma2(aR, pk(b[i], b[i]), acc[i]);
            if (S == 2) acc[i] = fma2(aU, pk(b[i], b[i]), acc[i]);
            if (S == 3) acc[i] = fma2(aR, pk(bs, bs), acc[i]);
            if (S == 4) acc[i] = fma2(ap[i], pk(bs, bs), acc[i]);
            if (S == 5) acc[i] = fma2(aU, bp[i], acc[i]);
        }
    }
    float s = 0.f;
    for (int i = 0; i < NK; ++i) { float lo, hi; up(acc[i], lo, hi); s += lo + hi; }
    if (s == 12345.678f) out[0] = s;
}
template <int S>
void run(const char* name, float* g) {
    float* d; cudaMalloc(&d, 16);
    int iters = 10000, grid = 148 * 8, threads = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<S><<<grid, threads>>>(d, iters / 10, 1.f, g);
    cudaEventRecord(e0);
    k<S><<<grid, threads>>>(d, iters, 1.f, g);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fmas = 2.0 * NK * iters * (double)threads * grid;
    printf("%-58s %7.3f ms %6.2f TFLOP/s (%5.1f%%)\n", name, ms, 2 * fmas / ms / 1e9, 2 * fmas / ms / 1e9 / 74.45 * 100);
    cudaFree(d);
}
int main() {
    float* g; cudaMalloc(&g, 4096 * 4); cudaMemset(g, 0, 4096 * 4);
    run<0>("S0 fma2(acc, A_UR, Breg-reuse)", g);
    run<1>("S1 fma2(Areg same, b_k.F32, acc_k)", g);
    run<2>("S2 fma2(A_UR, b_k.F32, acc_k)", g);
    run<3>("S3 fma2(Areg same, b.F32 same, acc_k)", g);
    run<4>("S4 fma2(Areg_k, b.F32 same, acc_k)", g);
    run<5>("S5 fma2(A_UR, bpair_k, acc_k)", g);
    return 0;
}
