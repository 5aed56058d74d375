// Microbenchmark 9: from the 98.7% S1 stream to the real clash loop, one factor at a time.
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0,{%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0,%1,%2,%3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ void up(u64 v, float& lo, float& hi) { asm("mov.b64 {%0,%1},%2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
// E1: acc_k = fma2(A, b_k, acc_k)                                   NK accumulators, same A
// E2: three passes with A[comp], b_k[comp] accumulating into acc_k
// E3: first pass starts from An (fresh temporaries each iteration): e_k = fma2(Az,bz_k,An); e_k = fma2(Ay,by_k,e_k); e_k=fma2(Ax,bx_k,e_k); sink: acc_k ^= e_k (LOP3 x2)
// E4: E3 with float mins as the sink
// E5: E4 with A (4 pairs) re-loaded from shared memory every iteration
// ORDER 0: component-major (all k for comp z, then y, then x); ORDER 1: chain-major (k outer, 3 dependent steps inner)
template <int E, int NK, int ORDER>
__global__ void __launch_bounds__(128) k(float* out, int iters, float seed, const float* __restrict__ g) {
    __shared__ ulonglong2 sA[64];
    if (threadIdx.x < 64) sA[threadIdx.x] = make_ulonglong2(pk(seed, seed * 1.1f), pk(seed * .5f, seed * .7f));
    __syncthreads();
    u64 acc[NK]; float bx[NK], by[NK], bz[NK], m[NK];
    for (int i = 0; i < NK; ++i) { acc[i] = pk(seed + i, seed - i); bx[i] = g[threadIdx.x + i]; by[i] = g[threadIdx.x + 64 + i]; bz[i] = g[threadIdx.x + 128 + i]; m[i] = 3e38f; }
    u64 Ax = pk(g[threadIdx.x + 200], g[threadIdx.x + 201]), Ay = pk(g[threadIdx.x + 202], g[threadIdx.x + 203]);
    u64 Az = pk(g[threadIdx.x + 204], g[threadIdx.x + 205]), An = pk(g[threadIdx.x + 206], g[threadIdx.x + 207]);
    for (int it = 0; it < iters; ++it) {
        if (E == 5) { ulonglong2 u0 = sA[(2 * it) & 63], u1 = sA[(2 * it + 1) & 63]; Ax = u0.x; Ay = u0.y; Az = u1.x; An = u1.y; }
        if (E == 1) {
#pragma unroll
            for (int i = 0; i < NK; ++i) { acc[i] = fma2(Ax, pk(bx[i], bx[i]), acc[i]); }
#pragma unroll
            for (int i = 0; i < NK; ++i) { acc[i] = fma2(Ax, pk(by[i], by[i]), acc[i]); }
#pragma unroll
            for (int i = 0; i < NK; ++i) { acc[i] = fma2(Ax, pk(bz[i], bz[i]), acc[i]); }
        } else if (E == 2) {
#pragma unroll
            for (int i = 0; i < NK; ++i) acc[i] = fma2(Az, pk(bz[i], bz[i]), acc[i]);
#pragma unroll
            for (int i = 0; i < NK; ++i) acc[i] = fma2(Ay, pk(by[i], by[i]), acc[i]);
#pragma unroll
            for (int i = 0; i < NK; ++i) acc[i] = fma2(Ax, pk(bx[i], bx[i]), acc[i]);
        } else {
            u64 e[NK];
            if (ORDER == 0) {
#pragma unroll
                for (int i = 0; i < NK; ++i) e[i] = fma2(Az, pk(bz[i], bz[i]), An);
#pragma unroll
                for (int i = 0; i < NK; ++i) e[i] = fma2(Ay, pk(by[i], by[i]), e[i]);
#pragma unroll
                for (int i = 0; i < NK; ++i) e[i] = fma2(Ax, pk(bx[i], bx[i]), e[i]);
            } else {
#pragma unroll
                for (int i = 0; i < NK; ++i) e[i] = fma2(Ax, pk(bx[i], bx[i]), fma2(Ay, pk(by[i], by[i]), fma2(Az, pk(bz[i], bz[i]), An)));
            }
#pragma unroll
            for (int i = 0; i < NK; ++i) {
                if (E == 3) acc[i] ^= e[i];
                else { float a, b; up(e[i], a, b); m[i] = fminf(fminf(m[i], a), b); }
            }
        }
    }
    float s = 0.f;
    for (int i = 0; i < NK; ++i) { float lo, hi; up(acc[i], lo, hi); s += lo + hi + m[i]; }
    if (s == 12345.678f) out[0] = s;
}
template <int E, int NK, int ORDER>
void run(const char* name, float* g) {
    float* d; cudaMalloc(&d, 16);
    int iters = 4000, grid = 148 * 16, threads = 128;
    auto kern = k<E, NK, ORDER>;
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<grid, threads>>>(d, iters / 10, 1.f, g);
    cudaEventRecord(e0);
    kern<<<grid, threads>>>(d, iters, 1.f, g);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fmas = 2.0 * 3 * NK * iters * (double)threads * grid;
    printf("%-46s NK=%2d order=%d regs=%3d warps/SM=%2d %7.3f ms %6.2f TFLOP/s (%5.1f%%)\n", name, NK, ORDER, fa.numRegs, occ * 4, ms, 2 * fmas / ms / 1e9, 2 * fmas / ms / 1e9 / 74.45 * 100);
    cudaFree(d);
}
int main() {
    float* g; cudaMalloc(&g, 4096 * 4); cudaMemset(g, 0, 4096 * 4);
    run<1, 10, 0>("E1 same A, 3 passes over b comps", g);  run<1, 30, 0>("E1", g);
    run<2, 10, 0>("E2 A[comp] rotates", g);                run<2, 30, 0>("E2", g);
    run<3, 10, 0>("E3 fresh temporaries + xor sink", g);   run<3, 30, 0>("E3", g);  run<3, 10, 1>("E3 chain-major", g); run<3, 30, 1>("E3 chain-major", g);
    run<4, 10, 0>("E4 fmin sink", g);                      run<4, 30, 0>("E4", g);  run<4, 10, 1>("E4 chain-major", g); run<4, 30, 1>("E4 chain-major", g);
    run<5, 10, 0>("E5 + A from smem", g);                  run<5, 30, 0>("E5", g);  run<5, 10, 1>("E5 chain-major", g); run<5, 30, 1>("E5 chain-major", g);
    return 0;
}
