// Microbenchmarks of the sm_100a FP32 pipes used to design the clash kernel (not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0,{%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0,%1,%2,%3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0,%1,%2,%3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float min3(float a, float b, float c) { float d; asm volatile("min.f32 %0,%1,%2,%3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ void up(u64 v, float& lo, float& hi) { asm("mov.b64 {%0,%1},%2;" : "=f"(lo), "=f"(hi) : "l"(v)); }

#define NK 8
// MODE 0: acc_k = fma2(acc_k, a, b)            one new pair per instr (a, b reusable)
// MODE 1: acc_k = fma2(x_k, y, acc_k)          two new pairs (x_k, acc_k), y reusable
// MODE 2: acc_k = fma2(x_k, y_k, acc_k)        three new pairs
// MODE 3: MODE 1 + one FMNMX3 per 3 FFMA2
// MODE 4: scalar acc_k = fma(x_k, y, acc_k)    (16 accumulators)
// MODE 5: scalar three distinct
// MODE 6: chain of 3 like the clash kernel: e = fma2(ax, bx_k, fma2(ay, by_k, fma2(az, bz_k, na))) + min3
// MODE 7: MODE 6 but B operand shared (A varies): e_k = fma2(a_kx, bx, ...) 
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float seed) {
    u64 acc[NK], x[NK], y[NK], z[NK];
    float m[NK], s1[2 * NK];
    for (int i = 0; i < NK; ++i) {
        acc[i] = pk(seed + i, seed - i);
        x[i] = pk(1.0f + 1e-7f * (i + threadIdx.x), 1.0f - 1e-7f * i);
        y[i] = pk(1e-9f * (i + 1), -1e-9f * (i + 2));
        z[i] = pk(1e-8f * (i + 1), -1e-8f * (i + 2));
        m[i] = 1e30f;
        s1[i] = seed + i; s1[i + NK] = seed - i;
    }
    u64 a = pk(1.0000001f, 0.9999999f), b = pk(seed * 1e-9f, -seed * 1e-9f), c = pk(seed * 1e-8f, seed * 2e-8f), d = pk(seed, seed);
    float fa = 1.0000001f + seed * 1e-9f, fb = seed * 1e-9f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NK; ++i) {
            if (MODE == 0) acc[i] = fma2(acc[i], a, b);
            if (MODE == 1) acc[i] = fma2(x[i], a, acc[i]);
            if (MODE == 2) acc[i] = fma2(x[i], y[i], acc[i]);
            if (MODE == 3) { acc[i] = fma2(x[i], a, acc[i]); }
            if (MODE == 4) { s1[i] = fma1(s1[i], fa, fb); s1[i + NK] = fma1(s1[i + NK], fa, fb); }
            if (MODE == 5) { float lo, hi, l2, h2; up(x[i], lo, hi); up(y[i], l2, h2); s1[i] = fma1(lo, l2, s1[i]); s1[i + NK] = fma1(hi, h2, s1[i + NK]); }
            if (MODE == 6) { u64 e = fma2(a, x[i], fma2(b, y[i], fma2(c, z[i], d))); float lo, hi; up(e, lo, hi); m[i] = min3(m[i], lo, hi); }
            if (MODE == 7) { u64 e = fma2(x[i], a, fma2(y[i], b, fma2(z[i], c, acc[i]))); float lo, hi; up(e, lo, hi); m[i] = min3(m[i], lo, hi); }
        }
        if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < NK; i += 3) { float lo, hi; up(acc[i], lo, hi); m[i] = min3(m[i], lo, hi); }
        }
        if (MODE == 6 || MODE == 7) { a = fma2(a, pk(1.f, 1.f), pk(1e-9f, 1e-9f)); }
    }
    float s = 0.f;
    for (int i = 0; i < NK; ++i) { float lo, hi; up(acc[i], lo, hi); s += lo + hi + m[i] + s1[i] + s1[i + NK]; }
    if (s == 12345.678f) out[0] = s;
}

template <int MODE>
void run(const char* name, double fma_per_thread_iter) {
    float* d; cudaMalloc(&d, 16);
    int iters = 20000, grid = 148 * 8, threads = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, threads>>>(d, iters / 10, 1.f);
    cudaEventRecord(e0);
    k<MODE><<<grid, threads>>>(d, iters, 1.f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fmas = fma_per_thread_iter * iters * (double)threads * grid;
    double clk = 1.965e9;
    printf("%-46s %8.3f ms  %7.2f TFLOP/s  (%.1f%% of 74.45)  fma-lanes/clk/SM=%.1f\n", name, ms, 2 * fmas / ms / 1e9, 2 * fmas / ms / 1e9 / 74.45 * 100, fmas / (ms * 1e-3) / clk / 148);
    cudaFree(d);
}
int main() {
    run<0>("ffma2 acc=fma2(acc,a,b) 1 new pair", 2.0 * NK);
    run<1>("ffma2 acc=fma2(x_k,a,acc) 2 new pairs", 2.0 * NK);
    run<2>("ffma2 acc=fma2(x_k,y_k,acc) 3 new pairs", 2.0 * NK);
    run<3>("ffma2 2 new pairs + fmnmx3 per 3", 2.0 * NK);
    run<4>("ffma scalar acc=fma(acc,a,b)", 2.0 * NK);
    run<5>("ffma scalar 3 distinct", 2.0 * NK);
    run<6>("clash-like chain (A shared, B_k regs) + min3", 6.0 * NK + 2);
    run<7>("clash-like chain (B shared, A_k regs) + min3", 6.0 * NK + 2);
    return 0;
}
