// firecode_b200 -- small batched helpers that keep the reference's single-structure entry points on the GPU.
//
//  * fc_rmsd_and_max_batch: prism_pruner.rmsd.rmsd_and_max(ref, structure, center) for a batch of
//    structures against one reference -- the arithmetic of firecode/utils.py:494-504 `rmsd_similarity`
//    (center = 0, all atoms) and of embedder.py:1784-1786 (center = 1).  One warp per structure, FP64
//    covariance by warp shuffles, Jacobi-based Kabsch in registers, second pass for RMSD / max deviation.
//  * fc_self_clash_batch: the non-fragment branch of firecode/utils.py:523-542 `compenetration_check`
//    (ids = None) and firecode/algebra.py:52-54 `count_clashes`: per structure the number of ORDERED atom
//    pairs with 0 < d < 0.5 A and the number of ordered pairs i != j with d < thresh that are not bonded.
#include "fc_embed.cuh"

namespace fc {

__device__ __forceinline__ double misc_wsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double misc_wmax(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__global__ void __launch_bounds__(128) rmsd_and_max_kernel(const double* __restrict__ ref, const double* __restrict__ xs,
                                                           long long n, int n_atoms, int center,
                                                           double* __restrict__ rmsd_out, double* __restrict__ maxdev_out) {
    const int lane = threadIdx.x & 31;
    const long long s = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= n) return;
    const double* q = xs + (size_t)s * n_atoms * 3;
    double mp[3] = {0, 0, 0}, mq[3] = {0, 0, 0};
    if (center) {
        for (int k = lane; k < n_atoms; k += 32)
            for (int c = 0; c < 3; ++c) { mp[c] += ref[3 * k + c]; mq[c] += q[3 * k + c]; }
        for (int c = 0; c < 3; ++c) { mp[c] = misc_wsum(mp[c]) / n_atoms; mq[c] = misc_wsum(mq[c]) / n_atoms; }
    }
    double h[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = lane; k < n_atoms; k += 32) {
        double x[3] = {ref[3 * k] - mp[0], ref[3 * k + 1] - mp[1], ref[3 * k + 2] - mp[2]};
        double y[3] = {q[3 * k] - mq[0], q[3 * k + 1] - mq[1], q[3 * k + 2] - mq[2]};
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) h[3 * r + c] += x[r] * y[c];
    }
#pragma unroll
    for (int e = 0; e < 9; ++e) h[e] = misc_wsum(h[e]);
    M3 R = kabsch_from_cov(h, nullptr);
    double ss = 0.0, mx = 0.0;
    for (int k = lane; k < n_atoms; k += 32) {
        double x[3] = {ref[3 * k] - mp[0], ref[3 * k + 1] - mp[1], ref[3 * k + 2] - mp[2]};
        double y[3] = {q[3 * k] - mq[0], q[3 * k + 1] - mq[1], q[3 * k + 2] - mq[2]};
        double dx = (x[0] * R.m[0] + x[1] * R.m[3] + x[2] * R.m[6]) - y[0];
        double dy = (x[0] * R.m[1] + x[1] * R.m[4] + x[2] * R.m[7]) - y[1];
        double dz = (x[0] * R.m[2] + x[1] * R.m[5] + x[2] * R.m[8]) - y[2];
        double d2 = dx * dx + dy * dy + dz * dz;
        ss += d2;
        mx = fmax(mx, d2);
    }
    ss = misc_wsum(ss);
    mx = misc_wmax(mx);
    if (lane == 0) {
        rmsd_out[s] = sqrt(ss / n_atoms);
        maxdev_out[s] = sqrt(mx);
    }
}

// one CTA per structure; FP64 distances as scipy cdist computes them (sqrt of the summed squares)
__global__ void __launch_bounds__(256) self_clash_kernel(const double* __restrict__ coords, int n_atoms,
                                                         const unsigned char* __restrict__ bonded, double thresh,
                                                         long long* __restrict__ close_out, long long* __restrict__ nonbonded_out) {
    const double* x = coords + (size_t)blockIdx.x * n_atoms * 3;
    long long close = 0, nb = 0;
    const long long total = (long long)n_atoms * n_atoms;
    for (long long e = threadIdx.x; e < total; e += blockDim.x) {
        int i = (int)(e / n_atoms), j = (int)(e - (long long)i * n_atoms);
        if (i == j) continue;
        double dx = x[3 * i] - x[3 * j], dy = x[3 * i + 1] - x[3 * j + 1], dz = x[3 * i + 2] - x[3 * j + 2];
        double d = sqrt(dx * dx + dy * dy + dz * dz);
        close += (d < 0.5 && d > 0.0) ? 1 : 0;
        if (d < thresh && !(bonded && bonded[(size_t)i * n_atoms + j])) nb += 1;
    }
    __shared__ long long s_c[8], s_n[8];
    for (int o = 16; o > 0; o >>= 1) {
        close += __shfl_xor_sync(0xffffffffu, close, o);
        nb += __shfl_xor_sync(0xffffffffu, nb, o);
    }
    if ((threadIdx.x & 31) == 0) { s_c[threadIdx.x >> 5] = close; s_n[threadIdx.x >> 5] = nb; }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long c = 0, m = 0;
        for (int w = 0; w < 8; ++w) { c += s_c[w]; m += s_n[w]; }
        close_out[blockIdx.x] = c;
        nonbonded_out[blockIdx.x] = m;
    }
}

// compenetration_check(structure, ids, thresh, max_clashes) (utils.py:544-575) for a batch of complete
// structures (what RunEmbedding.compenetration_refining loops over, embedder.py:1954-1975): one CTA per
// structure, FP64 distances as scipy cdist computes them.  Two fragments: #{d < thresh} <= max_clashes.
// Three fragments: cumulative #{d <= thresh} over (m2,m1), (m3,m2), (m1,m3); the reference returns False as
// soon as the running count exceeds max_clashes, which is the same as testing the total.
__global__ void __launch_bounds__(256) structure_clash_kernel(const double* __restrict__ coords, int n_atoms, int n1, int n2,
                                                              int n3, double thresh, long long* __restrict__ count_out,
                                                              double* __restrict__ closest_out) {
    const double* x = coords + (size_t)blockIdx.x * n_atoms * 3;
    const int b1 = n1, b2 = n1 + n2;
    long long cnt = 0;
    double closest = 1e300;
    // pair blocks as (rows, cols): two fragments -> (m2, m1); three -> (m2, m1), (m3, m2), (m1, m3)
    const int n_blocks = n3 > 0 ? 3 : 1;
    for (int blk = 0; blk < n_blocks; ++blk) {
        int r0, r1, c0, c1;
        if (blk == 0) { r0 = b1; r1 = b2; c0 = 0; c1 = b1; }
        else if (blk == 1) { r0 = b2; r1 = n_atoms; c0 = b1; c1 = b2; }
        else { r0 = 0; r1 = b1; c0 = b2; c1 = n_atoms; }
        const int nr = r1 - r0, nc = c1 - c0;
        for (long long e = threadIdx.x; e < (long long)nr * nc; e += blockDim.x) {
            const int i = r0 + (int)(e / nc), j = c0 + (int)(e % nc);
            const double dx = x[3 * i] - x[3 * j], dy = x[3 * i + 1] - x[3 * j + 1], dz = x[3 * i + 2] - x[3 * j + 2];
            const double d = sqrt(dx * dx + dy * dy + dz * dz);
            cnt += (n3 > 0 ? d <= thresh : d < thresh) ? 1 : 0;
            closest = fmin(closest, fabs(d - thresh));
        }
    }
    __shared__ long long s_c[8];
    __shared__ double s_m[8];
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        closest = fmin(closest, __shfl_xor_sync(0xffffffffu, closest, o));
    }
    if ((threadIdx.x & 31) == 0) { s_c[threadIdx.x >> 5] = cnt; s_m[threadIdx.x >> 5] = closest; }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long c = 0;
        double m = 1e300;
        for (int w = 0; w < 8; ++w) { c += s_c[w]; m = fmin(m, s_m[w]); }
        count_out[blockIdx.x] = c;
        closest_out[blockIdx.x] = m;
    }
}

}  // namespace fc

using namespace fc;

extern "C" int fc_rmsd_and_max_batch(const double* ref, const double* structures, int64_t n, int32_t n_atoms,
                                     int32_t center, double* rmsd_out, double* maxdev_out) {
    FC_REQUIRE(n >= 0 && n_atoms > 0, "fc_rmsd_and_max_batch: bad sizes");
    if (n == 0) return FC_OK;
    FC_REQUIRE(ref && structures && rmsd_out && maxdev_out, "fc_rmsd_and_max_batch: null pointer");
    sm_count();
    cudaStream_t s;
    FC_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    cudaError_t e = cudaSuccess;
    {
        DevBuf<double> d_ref, d_x, d_r, d_m;
#define MS(call) do { if (e == cudaSuccess) e = (call); } while (0)
        MS(d_ref.alloc((size_t)n_atoms * 3, s));
        MS(d_x.alloc((size_t)n * n_atoms * 3, s));
        MS(d_r.alloc((size_t)n, s));
        MS(d_m.alloc((size_t)n, s));
        MS(cudaMemcpyAsync(d_ref.p, ref, (size_t)n_atoms * 24, cudaMemcpyHostToDevice, s));
        MS(cudaMemcpyAsync(d_x.p, structures, (size_t)n * n_atoms * 24, cudaMemcpyHostToDevice, s));
        if (e == cudaSuccess) {
            rmsd_and_max_kernel<<<(unsigned)((n + 3) / 4), 128, 0, s>>>(d_ref.p, d_x.p, n, n_atoms, center, d_r.p, d_m.p);
            e = cudaGetLastError();
        }
        MS(cudaMemcpyAsync(rmsd_out, d_r.p, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
        MS(cudaMemcpyAsync(maxdev_out, d_m.p, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
        MS(cudaStreamSynchronize(s));
    }
    cudaStreamSynchronize(s);
    cudaStreamDestroy(s);
    if (e != cudaSuccess) return cuda_fail(e, "fc_rmsd_and_max_batch", __FILE__, __LINE__);
    return FC_OK;
}

extern "C" int fc_self_clash_batch(const double* coords, int64_t n, int32_t n_atoms, const uint8_t* bonded, double thresh,
                                   int64_t* close_pairs_out, int64_t* nonbonded_out) {
    FC_REQUIRE(n >= 0 && n_atoms > 0 && n < ((int64_t)1 << 31), "fc_self_clash_batch: bad sizes");
    if (n == 0) return FC_OK;
    FC_REQUIRE(coords && close_pairs_out && nonbonded_out, "fc_self_clash_batch: null pointer");
    sm_count();
    cudaStream_t s;
    FC_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    cudaError_t e = cudaSuccess;
    {
        DevBuf<double> d_x;
        DevBuf<unsigned char> d_b;
        DevBuf<long long> d_c, d_n;
        MS(d_x.alloc((size_t)n * n_atoms * 3, s));
        MS(d_c.alloc((size_t)n, s));
        MS(d_n.alloc((size_t)n, s));
        MS(cudaMemcpyAsync(d_x.p, coords, (size_t)n * n_atoms * 24, cudaMemcpyHostToDevice, s));
        if (bonded) {
            MS(d_b.alloc((size_t)n_atoms * n_atoms, s));
            MS(cudaMemcpyAsync(d_b.p, bonded, (size_t)n_atoms * n_atoms, cudaMemcpyHostToDevice, s));
        }
        if (e == cudaSuccess) {
            self_clash_kernel<<<(unsigned)n, 256, 0, s>>>(d_x.p, n_atoms, bonded ? d_b.p : nullptr, thresh, d_c.p, d_n.p);
            e = cudaGetLastError();
        }
        MS(cudaMemcpyAsync(close_pairs_out, d_c.p, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
        MS(cudaMemcpyAsync(nonbonded_out, d_n.p, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
        MS(cudaStreamSynchronize(s));
#undef MS
    }
    cudaStreamSynchronize(s);
    cudaStreamDestroy(s);
    if (e != cudaSuccess) return cuda_fail(e, "fc_self_clash_batch", __FILE__, __LINE__);
    return FC_OK;
}

extern "C" int fc_structure_clash_batch(const double* coords, int64_t n, int32_t n_atoms, const int32_t* ids, int32_t n_ids,
                                        double thresh, int64_t* count_out, double* closest_out) {
    FC_REQUIRE(n >= 0 && n_atoms > 0 && n < ((int64_t)1 << 31), "fc_structure_clash_batch: bad sizes");
    FC_REQUIRE(ids && (n_ids == 2 || n_ids == 3), "fc_structure_clash_batch: ids must name two or three fragments");
    int tot = 0;
    for (int k = 0; k < n_ids; ++k) {
        FC_REQUIRE(ids[k] > 0, "fc_structure_clash_batch: empty fragment");
        tot += ids[k];
    }
    FC_REQUIRE(tot == n_atoms, "fc_structure_clash_batch: ids sum to %d, structures have %d atoms", tot, n_atoms);
    if (n == 0) return FC_OK;
    FC_REQUIRE(coords && count_out && closest_out, "fc_structure_clash_batch: null pointer");
    sm_count();
    cudaStream_t s;
    FC_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    cudaError_t e = cudaSuccess;
    {
        DevBuf<double> d_x, d_m;
        DevBuf<long long> d_c;
#define MS(call) do { if (e == cudaSuccess) e = (call); } while (0)
        MS(d_x.alloc((size_t)n * n_atoms * 3, s));
        MS(d_c.alloc((size_t)n, s));
        MS(d_m.alloc((size_t)n, s));
        MS(cudaMemcpyAsync(d_x.p, coords, (size_t)n * n_atoms * 24, cudaMemcpyHostToDevice, s));
        if (e == cudaSuccess) {
            structure_clash_kernel<<<(unsigned)n, 256, 0, s>>>(d_x.p, n_atoms, ids[0], ids[1], n_ids == 3 ? ids[2] : 0, thresh,
                                                              d_c.p, d_m.p);
            e = cudaGetLastError();
        }
        MS(cudaMemcpyAsync(count_out, d_c.p, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
        MS(cudaMemcpyAsync(closest_out, d_m.p, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
        MS(cudaStreamSynchronize(s));
#undef MS
    }
    cudaStreamSynchronize(s);
    cudaStreamDestroy(s);
    if (e != cudaSuccess) return cuda_fail(e, "fc_structure_clash_batch", __FILE__, __LINE__);
    return FC_OK;
}


// ---------------------------------------------------------------------------------------------
// fitness_check over a batch (optimization_methods.py:163-180, looped by RunEmbedding.fitness_refining,
// embedder.py:1997-2039): error[s] = sum over the structure's constraints with a target distance of
// (|x_a - x_b| - target), accumulated in constraint order in FP64 without contraction, as the reference's loop does.
// A constraint without target (None in the reference) is passed as NaN.
// ---------------------------------------------------------------------------------------------
namespace fc {
__global__ void fitness_kernel(const double* __restrict__ coords, long long n, int n_atoms, const int* __restrict__ pairs,
                               const double* __restrict__ targets, int n_constraints, double* __restrict__ error) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const double* x = coords + (size_t)s * n_atoms * 3;
    double err = 0.0;
    for (int c = 0; c < n_constraints; ++c) {
        const double t = targets[(size_t)s * n_constraints + c];
        if (t != t) continue;  // no target for this constraint
        const int a = pairs[((size_t)s * n_constraints + c) * 2], b = pairs[((size_t)s * n_constraints + c) * 2 + 1];
        const double dx = x[3 * a] - x[3 * b], dy = x[3 * a + 1] - x[3 * b + 1], dz = x[3 * a + 2] - x[3 * b + 2];
        const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        err = __dadd_rn(err, __dsub_rn(sqrt(d2), t));
    }
    error[s] = err;
}
}  // namespace fc

extern "C" int fc_fitness_batch(const double* coords, int64_t n, int32_t n_atoms, const int32_t* pairs, const double* targets,
                                int32_t n_constraints, double* error_out) {
    FC_REQUIRE(n >= 0 && n_atoms > 0 && n_constraints >= 0 && n < ((int64_t)1 << 31), "fc_fitness_batch: bad sizes");
    if (n == 0) return FC_OK;
    FC_REQUIRE(coords && error_out && (n_constraints == 0 || (pairs && targets)), "fc_fitness_batch: null pointer");
    for (int64_t i = 0; i < n * (int64_t)n_constraints * 2; ++i)
        FC_REQUIRE(pairs[i] >= 0 && pairs[i] < n_atoms, "fc_fitness_batch: atom index %d out of range", (int)pairs[i]);
    cudaStream_t s;
    FC_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    cudaError_t e = cudaSuccess;
    {
        DevBuf<double> d_x, d_t, d_e;
        DevBuf<int> d_p;
#define MS(call) do { if (e == cudaSuccess) e = (call); } while (0)
        const size_t nc = (size_t)n * (size_t)std::max(n_constraints, 1);
        MS(d_x.alloc((size_t)n * n_atoms * 3, s));
        MS(d_t.alloc(nc, s));
        MS(d_p.alloc(2 * nc, s));
        MS(d_e.alloc((size_t)n, s));
        MS(upload_rows_staged(d_x.p, coords, n, n_atoms, nullptr, n_atoms, s));
        if (n_constraints > 0) {
            MS(cudaMemcpyAsync(d_t.p, targets, (size_t)n * n_constraints * 8, cudaMemcpyHostToDevice, s));
            MS(cudaMemcpyAsync(d_p.p, pairs, (size_t)n * n_constraints * 8, cudaMemcpyHostToDevice, s));
        }
        if (e == cudaSuccess) {
            fitness_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_x.p, n, n_atoms, d_p.p, d_t.p, n_constraints, d_e.p);
            e = cudaGetLastError();
        }
        MS(cudaMemcpyAsync(error_out, d_e.p, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
        MS(cudaStreamSynchronize(s));
#undef MS
    }
    cudaStreamSynchronize(s);
    cudaStreamDestroy(s);
    if (e != cudaSuccess) return cuda_fail(e, "fc_fitness_batch", __FILE__, __LINE__);
    return FC_OK;
}

// Host-only test hook (no CUDA call): the Kabsch rotation the kernels use for a 3x3 cross-covariance h (row-major),
// r_out = U D V^T, sig_out = singular values with the sign of the third set by det(h).  tests/test_host_logic.py
// feeds it rank-deficient covariances (collinear atoms, a single atom).
extern "C" int fc_kabsch_host(const double* h, double* r_out, double* sig_out) {
    FC_REQUIRE(h && r_out, "fc_kabsch_host: null pointer");
    double sig[3];
    fc::M3 r = fc::kabsch_from_cov(h, sig);
    for (int i = 0; i < 9; ++i) r_out[i] = r.m[i];
    if (sig_out)
        for (int i = 0; i < 3; ++i) sig_out[i] = sig[i];
    return FC_OK;
}

