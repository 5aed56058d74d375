// Microbenchmark 7: clash loop v3 candidates: B 32-bit broadcast operands, A natural pairs in smem.
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0,{%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0,%1,%2,%3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float min3(float a, float b, float c) { float d; asm("min.f32 %0,%1,%2,%3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ void up(u64 v, float& lo, float& hi) { asm("mov.b64 {%0,%1},%2;" : "=f"(lo), "=f"(hi) : "l"(v)); }

// MINV: 0 = FMNMX3, 1 = two FMNMX, 2 = integer min on the raw bits (two VIMNMX / one VIMNMX3), 3 = no min (xor)
// UNR: A pairs per unrolled step
template <int MINV, int TB, int THREADS, int MINB, int UNR>
__global__ void __launch_bounds__(THREADS, MINB) k(float* out, int n_a, int reps, float seed) {
    extern __shared__ ulonglong2 sA[];
    for (int i = threadIdx.x; i < n_a + 8; i += blockDim.x) { float v = 1e-3f * (i + 1); sA[i] = make_ulonglong2(pk(v, v * 1.0001f), pk(-v, v * 0.5f)); }
    __syncthreads();
    float bx[TB], by[TB], bz[TB], m[TB];
    for (int q = 0; q < TB; ++q) { float x = seed + q + threadIdx.x * 1e-3f; bx[q] = x; by[q] = x * .5f; bz[q] = x * .25f; m[q] = 3e38f; }
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll UNR
        for (int i = 0; i < n_a; i += 2) {
            ulonglong2 u0 = sA[i], u1 = sA[i + 1];
#pragma unroll
            for (int q = 0; q < TB; ++q) {
                u64 e = fma2(u0.x, pk(bx[q], bx[q]), fma2(u0.y, pk(by[q], by[q]), fma2(u1.x, pk(bz[q], bz[q]), u1.y)));
                float a, b; up(e, a, b);
                if (MINV == 0) m[q] = min3(m[q], a, b);
                if (MINV == 1) m[q] = fminf(fminf(m[q], a), b);
                if (MINV == 2) m[q] = __int_as_float(min(min(__float_as_int(m[q]), __float_as_int(a)), __float_as_int(b)));
                if (MINV == 3) m[q] = __int_as_float(__float_as_int(m[q]) ^ __float_as_int(a) ^ __float_as_int(b));
            }
        }
    }
    float res = 0.f;
    for (int j = 0; j < TB; ++j) res += m[j];
    if (res == 12345.678f) out[0] = res;
}
template <int MINV, int TB, int THREADS, int MINB, int UNR>
void run(const char* name) {
    float* d; cudaMalloc(&d, 16);
    int n_a = 150, reps = 200, grid = 148 * MINB;
    size_t smem = (n_a + 8) * 16;
    auto kern = k<MINV, TB, THREADS, MINB, UNR>;
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<grid, THREADS, smem>>>(d, n_a, reps / 10, 1.f);
    cudaEventRecord(e0);
    kern<<<grid, THREADS, smem>>>(d, n_a, reps, 1.f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fmas = 3.0 * TB * n_a * (double)reps * THREADS * grid;
    printf("%-12s TB=%2d unr=%d regs=%3d spill=%zu thr=%3d x%d (%2d warps/SM) %7.3f ms %6.2f TFLOP/s (%5.1f%%)\n", name, TB, UNR, fa.numRegs, fa.localSizeBytes, THREADS, MINB, THREADS * MINB / 32, ms,
           2 * fmas / ms / 1e9, 2 * fmas / ms / 1e9 / 74.45 * 100);
    cudaFree(d);
}
int main() {
    run<0, 10, 256, 2, 2>("min3");   run<1, 10, 256, 2, 2>("2xfmnmx"); run<2, 10, 256, 2, 2>("imin");  run<3, 10, 256, 2, 2>("xor");
    run<0, 16, 256, 2, 2>("min3");   run<1, 16, 256, 2, 2>("2xfmnmx"); run<2, 16, 256, 2, 2>("imin");  run<3, 16, 256, 2, 2>("xor");
    run<0, 20, 128, 3, 2>("min3");   run<1, 20, 128, 3, 2>("2xfmnmx"); run<2, 20, 128, 3, 2>("imin");
    run<0, 30, 128, 2, 2>("min3");   run<1, 30, 128, 2, 2>("2xfmnmx"); run<2, 30, 128, 2, 2>("imin");  run<3, 30, 128, 2, 2>("xor");
    run<1, 30, 128, 2, 1>("2xfmnmx"); run<1, 30, 160, 2, 2>("2xfmnmx"); run<1, 30, 192, 2, 2>("2xfmnmx");
    run<1, 38, 128, 2, 1>("2xfmnmx"); run<1, 50, 96, 2, 1>("2xfmnmx");
    return 0;
}
