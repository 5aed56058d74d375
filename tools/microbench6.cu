// Microbenchmark 6: which companion instructions co-issue with a saturated FFMA2 stream?
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0,{%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0,%1,%2,%3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ void up(u64 v, float& lo, float& hi) { asm("mov.b64 {%0,%1},%2;" : "=f"(lo), "=f"(hi) : "l"(v)); }

// X: 0 none, 1 FMNMX3, 2 FMNMX, 3 min.s32, 4 xor, 5 add.s32, 6 FADD, 7 LDS.128 broadcast, 8 FMUL, 9 max.u32 x3 (vimnmx3?)
// PER: one companion op per PER FFMA2 (PER in 1,2,3,6)
template <int X, int PER>
__global__ void __launch_bounds__(256) k(float* out, int iters, float seed) {
    __shared__ float4 sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = make_float4(seed, seed, seed, seed);
    __syncthreads();
    u64 acc[12];
    float f[12]; int n[12];
    for (int i = 0; i < 12; ++i) { acc[i] = pk(seed + i, seed - i); f[i] = seed * i; n[i] = threadIdx.x + i; }
    u64 a = pk(1.0000001f, 0.9999999f), b = pk(seed * 1e-9f, -seed * 1e-9f);
    float g = seed * 0.5f, h = seed * 0.25f;
    int gi = threadIdx.x * 3, hi2 = threadIdx.x * 7;
    float4 ld = make_float4(0, 0, 0, 0);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            acc[i] = fma2(acc[i], a, b);
            if (i % PER == PER - 1) {
                if (X == 1) asm volatile("min.f32 %0,%0,%1,%2;" : "+f"(f[i]) : "f"(g), "f"(h));
                if (X == 2) asm volatile("min.f32 %0,%0,%1;" : "+f"(f[i]) : "f"(g));
                if (X == 3) asm volatile("min.s32 %0,%0,%1;" : "+r"(n[i]) : "r"(gi));
                if (X == 4) asm volatile("xor.b32 %0,%0,%1;" : "+r"(n[i]) : "r"(gi));
                if (X == 5) asm volatile("add.s32 %0,%0,%1;" : "+r"(n[i]) : "r"(gi));
                if (X == 6) asm volatile("add.f32 %0,%0,%1;" : "+f"(f[i]) : "f"(g));
                if (X == 7) { float4 t; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3},[%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"((unsigned)__cvta_generic_to_shared(&sm[(it + i) & 63]))); ld.x += t.x; }
                if (X == 8) asm volatile("mul.f32 %0,%0,%1;" : "+f"(f[i]) : "f"(g));
                if (X == 9) { int t; asm volatile("min.s32 %0,%1,%2;" : "=r"(t) : "r"(n[i]), "r"(gi)); asm volatile("min.s32 %0,%1,%2;" : "=r"(n[i]) : "r"(t), "r"(hi2)); }
            }
        }
    }
    float s = ld.x;
    for (int i = 0; i < 12; ++i) { float lo, hi; up(acc[i], lo, hi); s += lo + hi + f[i] + n[i]; }
    if (s == 12345.678f) out[0] = s;
}
template <int X, int PER>
void run(const char* name) {
    float* d; cudaMalloc(&d, 16);
    int iters = 10000, grid = 148 * 8, threads = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<X, PER><<<grid, threads>>>(d, iters / 10, 1.f);
    cudaEventRecord(e0);
    k<X, PER><<<grid, threads>>>(d, iters, 1.f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fmas = 2.0 * 12 * iters * (double)threads * grid;
    double cyc_per_ffma2 = ms * 1e-3 * 1.965e9 / (12.0 * iters * (threads / 32) * 8 / 4);
    printf("%-28s 1 per %d FFMA2: %7.3f ms  FFMA2-only %6.2f TFLOP/s (%5.1f%%)  cycles per FFMA2 (+companions) = %.2f\n", name, PER, ms, 2 * fmas / ms / 1e9, 2 * fmas / ms / 1e9 / 74.45 * 100, cyc_per_ffma2);
    cudaFree(d);
}
int main() {
    run<0, 3>("none");
    run<1, 3>("FMNMX3"); run<1, 1>("FMNMX3"); run<1, 6>("FMNMX3");
    run<2, 3>("FMNMX"); run<2, 1>("FMNMX");
    run<3, 3>("IMNMX(min.s32)"); run<3, 1>("IMNMX(min.s32)");
    run<4, 3>("LOP3 xor"); run<4, 1>("LOP3 xor");
    run<5, 3>("IADD"); run<5, 1>("IADD");
    run<6, 3>("FADD"); run<6, 1>("FADD");
    run<8, 3>("FMUL");
    run<7, 3>("LDS.128"); run<7, 6>("LDS.128");
    run<9, 3>("2x min.s32");
    return 0;
}
