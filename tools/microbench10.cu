// Microbenchmark 10: cost of the per-chain SINK of the clash inner loop (A pairs from shared memory, TB atoms
// of B in registers, 3 FFMA2 per chain).  ptxas orders the FFMA2s component-major over all TB chains whatever
// the PTX order is (volatile asm does not pin SASS order) and fuses min(min(m,lo),hi) into FMNMX3.
// SINK 0: fminf(fminf(m,lo),hi) -> FMNMX3      1: two accumulators, 2 x FMNMX      2: FMNMX + integer VIMNMX
//      3: sign-OR (one LOP3)                   4: predicate accumulate (FSETP)      5: xor on the pair (2 LOP3)
//      6: min of the low half only (1 FMNMX; timing reference, not a valid reduction)
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0,{%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0,%1,%2,%3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ void up(u64 v, float& lo, float& hi) { asm("mov.b64 {%0,%1},%2;" : "=f"(lo), "=f"(hi) : "l"(v)); }

template <int TB, int SINK, int UNR, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k(float* out, int n_pairs, int reps, const float* __restrict__ g, float thr) {
    extern __shared__ ulonglong2 sA[];
    for (int i = threadIdx.x; i < 2 * n_pairs; i += blockDim.x) { float v = g[i & 1023] + 1e-3f * (i + 1); sA[i] = make_ulonglong2(pk(v, v * 1.0001f), pk(-v, v * 0.5f)); }
    __syncthreads();
    float bx[TB], by[TB], bz[TB], m[TB], m2[TB];
    unsigned orr[TB];
    bool flag = false;
#pragma unroll
    for (int q = 0; q < TB; ++q) { bx[q] = g[threadIdx.x + q]; by[q] = g[threadIdx.x + 64 + q]; bz[q] = g[threadIdx.x + 128 + q]; m[q] = 3e38f; m2[q] = 3e38f; orr[q] = 0; }
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll UNR
        for (int i = 0; i < n_pairs; ++i) {
            ulonglong2 u0 = sA[2 * i], u1 = sA[2 * i + 1];
#pragma unroll
            for (int q = 0; q < TB; ++q) {
                u64 e = fma2(u0.x, pk(bx[q], bx[q]), fma2(u0.y, pk(by[q], by[q]), fma2(u1.x, pk(bz[q], bz[q]), u1.y)));
                float lo, hi; up(e, lo, hi);
                if (SINK == 0) m[q] = fminf(fminf(m[q], lo), hi);
                if (SINK == 1) { m[q] = fminf(m[q], lo); m2[q] = fminf(m2[q], hi); }
                if (SINK == 2) { m[q] = fminf(m[q], lo); m[q] = __int_as_float(min(__float_as_int(m[q]), __float_as_int(hi))); }
                if (SINK == 3) orr[q] |= __float_as_uint(lo) | __float_as_uint(hi);
                if (SINK == 4) flag = flag || (lo < thr) || (hi < thr);
                if (SINK == 5) { orr[q] ^= __float_as_uint(lo); m2[q] = __uint_as_float(__float_as_uint(m2[q]) ^ __float_as_uint(hi)); }
                if (SINK == 6) m[q] = fminf(m[q], lo);
            }
        }
    }
    float res = flag ? 1.f : 0.f;
#pragma unroll
    for (int j = 0; j < TB; ++j) res += m[j] + m2[j] + (float)orr[j];
    if (res == 12345.678f) out[0] = res;
}

template <int TB, int SINK, int UNR, int THREADS, int MINB>
void run(float* g) {
    float* d; cudaMalloc(&d, 16);
    int n_pairs = 75, reps = 300, grid = 148 * MINB;
    size_t smem = (size_t)2 * n_pairs * 16;
    auto kern = k<TB, SINK, UNR, THREADS, MINB>;
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<grid, THREADS, smem>>>(d, n_pairs, reps / 10, g, 1.f);
    cudaEventRecord(e0);
    kern<<<grid, THREADS, smem>>>(d, n_pairs, reps, g, 1.f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fmas = 2.0 * 3.0 * TB * n_pairs * (double)reps * THREADS * grid;
    printf("TB=%2d sink=%d unr=%d thr=%3d x%d regs=%3d occ=%d warps/SM=%2d %8.3f ms %6.2f TFLOP/s (%5.1f%%) %s\n", TB, SINK, UNR, THREADS, MINB, fa.numRegs, occ,
           occ * THREADS / 32, ms, 2 * fmas / ms / 1e9, 2 * fmas / ms / 1e9 / 74.45 * 100, cudaGetErrorString(cudaGetLastError()));
    fflush(stdout);
    cudaFree(d);
}
int main() {
    float* g; cudaMalloc(&g, 4096 * 4); cudaMemset(g, 0, 4096 * 4);
    run<30, 0, 1, 128, 2>(g);
    run<30, 1, 1, 128, 2>(g);
    run<30, 2, 1, 128, 2>(g);
    run<30, 3, 1, 128, 2>(g);
    run<30, 5, 1, 128, 2>(g);
    run<30, 6, 1, 128, 2>(g);
    run<30, 0, 2, 128, 2>(g);
    run<30, 1, 2, 128, 2>(g);
    run<30, 2, 2, 128, 2>(g);
    run<30, 3, 2, 128, 2>(g);
    run<25, 0, 1, 128, 2>(g);
    run<20, 0, 1, 128, 3>(g);
    run<20, 0, 1, 192, 2>(g);
    run<16, 0, 1, 128, 4>(g);
    run<10, 0, 1, 256, 3>(g);
    run<25, 1, 1, 128, 2>(g);
    run<20, 1, 1, 128, 3>(g);
    run<20, 1, 1, 192, 2>(g);
    run<16, 1, 1, 128, 4>(g);
    run<10, 1, 1, 256, 3>(g);
    run<25, 2, 1, 128, 2>(g);
    run<20, 2, 1, 128, 3>(g);
    run<20, 2, 1, 192, 2>(g);
    run<16, 2, 1, 128, 4>(g);
    run<10, 2, 1, 256, 3>(g);
    run<25, 3, 1, 128, 2>(g);
    run<20, 3, 1, 128, 3>(g);
    run<20, 3, 1, 192, 2>(g);
    run<16, 3, 1, 128, 4>(g);
    run<10, 3, 1, 256, 3>(g);
    return 0;
}
