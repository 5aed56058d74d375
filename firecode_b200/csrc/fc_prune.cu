// firecode_b200 -- ensemble similarity pruning (RMSD and moment-of-inertia flavours).
//
// Replaces prism_pruner.pruner.prune_by_rmsd / prune_by_moment_of_inertia as FIRECODE calls them
// (/root/reference/firecode/embedder.py:1452,1472; ensemble.py:211,230; operators.py:613-624;
// atropisomer_module.py:504).  prism_pruner is absent from the reference tree: the algorithm is the
// one restated in oracle/prism_pruner/pruner.py (PARITY UNPINNED), whose in-tree
// structural analogue is firecode/torsion_module.py:957-1043.
//
// Driver (host, per pass k of the K schedule): the array is cut in k contiguous chunks; inside each chunk every
// pair of ACTIVE structures is screened on the GPU -- by default on the tensor cores (fc_gram_tc.cuh: TF32 Gram
// matrix of the centred coordinates, FP32 bounds on the sum of singular values in the epilogue), otherwise by
// prune_screen_f32_kernel (FP32 CUDA cores, 32x32 pair tiles) -- and the pairs a screen cannot rule out are
// decided in FP64 by prune_exact_kernel (one warp per pair: covariance, Jacobi, rotate-and-deviate as
// rmsd_and_max does).  FC_PRUNE_FP64=1 and the MOI flavour use prune_pairs_kernel alone.  Similar pairs are
// sorted by their deciding structure on the device and come back as one list, which the host resolves in a
// linear scan (prune_resolve); a pass costs one host wait.  Several GPUs: work items dealt to the ranks, lists
// all-gathered through the host (fc_prune_sharded) or GPU to GPU (fc_prune_sharded_dev).
//   similar(i, j)  <=>  rmsd < max_rmsd  and  max deviation < max_dev      (heavy atoms, centred)
//   MOI flavour    <=>  all three principal moments within max_deviation (relative to structure i)
// Keep rules: "greedy" updates the mask in place (NMS sweep), "snapshot" reads the mask of the pass
// start; "first" keeps the earlier structure of a similar pair, "last" the later one.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "fc_embed.cuh"
#include "fc_gram_tc.cuh"

namespace fc {

// ---------------------------------------------------------------------------------------------
// per-structure preparation
// ---------------------------------------------------------------------------------------------
// heavy-atom coordinates centred on their mean: out (n, nh, 3); g[i] = sum |x|^2
// outf: the same coordinates as float4 {x, y, z, 0} for the FP32 screen
// mom (optional): the six second moments xx, xy, xz, yy, yz, zz of the centred coordinates (prune_sigma_kernel turns
// them into the shape numbers the screen culls tiles with)
__global__ void prune_center_kernel(const double* __restrict__ coords, int n_atoms, const int* __restrict__ sel,
                                    int nh, long long n, double* __restrict__ out, double* __restrict__ g,
                                    float4* __restrict__ outf, double* __restrict__ mom) {
    long long s = blockIdx.x;
    if (s >= n) return;
    const double* src = coords + (size_t)s * n_atoms * 3;
    __shared__ double sm[3][32];
    double acc[3] = {0, 0, 0};
    for (int k = threadIdx.x; k < nh; k += blockDim.x) {
        const double* a = src + 3 * (sel ? sel[k] : k);
        acc[0] += a[0]; acc[1] += a[1]; acc[2] += a[2];
    }
    for (int c = 0; c < 3; ++c) {
        double v = acc[c];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) sm[c][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    double mean[3];
    for (int c = 0; c < 3; ++c) {
        double v = 0;
        for (int w = 0; w < (blockDim.x + 31) / 32; ++w) v += sm[c][w];
        mean[c] = v / nh;
    }
    __syncthreads();
    double gg = 0, m2[6] = {0, 0, 0, 0, 0, 0};  // xx, xy, xz, yy, yz, zz
    double* dst = out + (size_t)s * nh * 3;
    for (int k = threadIdx.x; k < nh; k += blockDim.x) {
        const double* a = src + 3 * (sel ? sel[k] : k);
        double x = a[0] - mean[0], y = a[1] - mean[1], z = a[2] - mean[2];
        dst[3 * k] = x; dst[3 * k + 1] = y; dst[3 * k + 2] = z;
        outf[(size_t)s * nh + k] = make_float4((float)x, (float)y, (float)z, 0.f);
        gg += x * x + y * y + z * z;
        m2[0] += x * x; m2[1] += x * y; m2[2] += x * z; m2[3] += y * y; m2[4] += y * z; m2[5] += z * z;
    }
    for (int o = 16; o > 0; o >>= 1) gg += __shfl_xor_sync(0xffffffffu, gg, o);
    if ((threadIdx.x & 31) == 0) sm[0][threadIdx.x >> 5] = gg;
    __shared__ double sm2[6][32];
    if (mom) {
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            double v = m2[c];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) == 0) sm2[c][threadIdx.x >> 5] = v;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0;
        for (int w = 0; w < (blockDim.x + 31) / 32; ++w) v += sm[0][w];
        g[s] = v;
        if (mom)
            for (int c = 0; c < 6; ++c) {
                double t = 0;
                for (int w = 0; w < (blockDim.x + 31) / 32; ++w) t += sm2[c][w];
                mom[6 * s + c] = t;
            }
    }
}

// sig[s] = the three singular values of the centred coordinate matrix of structure s, descending (square roots of the
// eigenvalues of its 3x3 Gram matrix): the rotation-invariant shape numbers of the tile culling.  One thread per structure.
__global__ void prune_sigma_kernel(const double* __restrict__ mom, long long n, float4* __restrict__ sig) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const double* t = mom + 6 * s;
    const double a9[9] = {t[0], t[1], t[2], t[1], t[3], t[4], t[2], t[4], t[5]};
    double w3[3], v9[9];
    jacobi_eig3(a9, w3, v9);
    if (w3[0] < w3[1]) { double q = w3[0]; w3[0] = w3[1]; w3[1] = q; }
    if (w3[1] < w3[2]) { double q = w3[1]; w3[1] = w3[2]; w3[2] = q; }
    if (w3[0] < w3[1]) { double q = w3[0]; w3[0] = w3[1]; w3[1] = q; }
    sig[s] = make_float4((float)sqrt(fmax(w3[0], 0.0)), (float)sqrt(fmax(w3[1], 0.0)), (float)sqrt(fmax(w3[2], 0.0)), 0.f);
}

// principal moments of inertia (ascending) about the centre of mass: moi (n, 3)
__global__ void prune_moi_kernel(const double* __restrict__ coords, const double* __restrict__ masses, int n_atoms,
                                 long long n, double* __restrict__ moi) {
    long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const double* x = coords + (size_t)s * n_atoms * 3;
    double com[3] = {0, 0, 0}, mt = 0;
    for (int k = 0; k < n_atoms; ++k) {
        com[0] += x[3 * k] * masses[k]; com[1] += x[3 * k + 1] * masses[k]; com[2] += x[3 * k + 2] * masses[k];
        mt += masses[k];
    }
    com[0] /= mt; com[1] /= mt; com[2] /= mt;
    double t[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < n_atoms; ++k) {
        double a = x[3 * k] - com[0], b = x[3 * k + 1] - com[1], c = x[3 * k + 2] - com[2], m = masses[k];
        t[0] += m * (b * b + c * c); t[4] += m * (a * a + c * c); t[8] += m * (a * a + b * b);
        t[1] -= m * a * b; t[2] -= m * a * c; t[5] -= m * b * c;
    }
    t[3] = t[1]; t[6] = t[2]; t[7] = t[5];
    double w[3], v[9];
    jacobi_eig3(t, w, v);
    if (w[0] > w[1]) { double q = w[0]; w[0] = w[1]; w[1] = q; }
    if (w[1] > w[2]) { double q = w[1]; w[1] = w[2]; w[2] = q; }
    if (w[0] > w[1]) { double q = w[0]; w[0] = w[1]; w[1] = q; }
    moi[3 * s] = w[0]; moi[3 * s + 1] = w[1]; moi[3 * s + 2] = w[2];
}

// ---------------------------------------------------------------------------------------------
// pair tiles
// ---------------------------------------------------------------------------------------------
struct PruneTile {
    int row0, col0;      // positions in the compacted active list
    int chunk_begin;     // first active position of the chunk
    int chunk_len;       // active members of the chunk
};

struct PruneArgs {
    const double* xc;        // (n, nh, 3) centred heavy-atom coordinates
    const double* g;         // (n)
    const double* moi;       // (n, 3) or null
    const double* energies;  // (n) or null
    const int* active;       // compacted active list -> structure index
    const PruneTile* tiles;
    int nh;
    int mode;                // 0 = rmsd, 1 = moi
    double max_rmsd, max_dev, max_dE, moi_dev, eps;
    const float4* xcf;           // (n, nh) centred heavy-atom coordinates, FP32
    int2* cand;                  // pairs the FP32 screen could not rule out (structure indices, x < y)
    unsigned long long* n_cand;
    long long cand_cap;
    int2* pairs;                 // similar pairs found in this pass (structure indices, x < y)
    unsigned long long* n_pairs; // ... their number (may exceed pair_cap: the pass is then repeated)
    long long pair_cap;
    int swap_xy;                 // 1: a pair is stored {later, earlier} (keep-first), 0: {earlier, later} (keep-last) -- the
                                 // structure that decides (the "head") is always .y, the droppable one .x, so that the
                                 // list sorts by head as 64-bit keys
    unsigned long long* n_eval;  // pairs fully evaluated
    TieRecord* ties;
    int* n_ties;
    int tie_cap;
};

#define PR_TS 32        // structures per tile side
#define PR_ATOMS 32     // atoms staged per step

// similar pairs go to a compact list (sparse: survivors of earlier passes are mutually dissimilar)
__device__ __forceinline__ void prune_push_pair(const PruneArgs& a, int s_lo, int s_hi) {
    unsigned long long slot = atomicAdd(a.n_pairs, 1ull);
    if ((long long)slot < a.pair_cap) a.pairs[slot] = a.swap_xy ? make_int2(s_hi, s_lo) : make_int2(s_lo, s_hi);
}

// 256 threads: thread (tx = column, ty) owns pairs (row ty + 8u, column tx), u = 0..3
__global__ void __launch_bounds__(256) prune_pairs_kernel(PruneArgs a) {
    const PruneTile t = a.tiles[blockIdx.x];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int chunk_end = t.chunk_begin + t.chunk_len;
    const int col = t.col0 + tx;
    const bool col_ok = col < chunk_end;
    const int s_col = col_ok ? a.active[col] : -1;
    int row[4], s_row[4];
    bool pair_ok[4];
    bool any = false;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        row[u] = t.row0 + ty + 8 * u;
        bool ok = row[u] < chunk_end;
        s_row[u] = ok ? a.active[row[u]] : -1;
        pair_ok[u] = ok && col_ok && row[u] < col;  // each unordered pair once (upper triangle)
        if (pair_ok[u] && a.energies) pair_ok[u] = fabs(a.energies[s_row[u]] - a.energies[s_col]) < a.max_dE;
        any |= pair_ok[u];
    }
    if (a.mode == 1) {  // moment-of-inertia flavour: O(1) per pair
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (!pair_ok[u]) continue;
            const double* mi = a.moi + 3 * (size_t)s_row[u];
            const double* mj = a.moi + 3 * (size_t)s_col;
            bool sim = true;
            for (int k = 0; k < 3; ++k) sim = sim && (fabs(mi[k] - mj[k]) / mi[k] < a.moi_dev);
            if (sim) prune_push_pair(a, s_row[u], s_col);
        }
        return;
    }
    __shared__ double sR[PR_ATOMS][3][PR_TS];  // rows of the tile
    __shared__ double sC[PR_ATOMS][3][PR_TS];  // columns of the tile
    double h[4][9];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int e = 0; e < 9; ++e) h[u][e] = 0.0;
    const int nh = a.nh;
    for (int a0 = 0; a0 < nh; a0 += PR_ATOMS) {
        const int na = min(PR_ATOMS, nh - a0);
        __syncthreads();
        // stage: element (structure s of the tile, atom k, coordinate c)
        for (int e = threadIdx.x; e < PR_TS * na * 3; e += 256) {
            int s = e / (na * 3), rem = e - s * (na * 3);
            int k = rem / 3, c = rem - 3 * k;
            int rpos = t.row0 + s, cpos = t.col0 + s;
            sR[k][c][s] = rpos < chunk_end ? a.xc[((size_t)a.active[rpos] * nh + a0 + k) * 3 + c] : 0.0;
            sC[k][c][s] = cpos < chunk_end ? a.xc[((size_t)a.active[cpos] * nh + a0 + k) * 3 + c] : 0.0;
        }
        __syncthreads();
        if (!__syncthreads_or(any)) continue;
        for (int k = 0; k < na; ++k) {
            const double qx = sC[k][0][tx], qy = sC[k][1][tx], qz = sC[k][2][tx];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = ty + 8 * u;
                const double px = sR[k][0][r], py = sR[k][1][r], pz = sR[k][2][r];
                h[u][0] += px * qx; h[u][1] += px * qy; h[u][2] += px * qz;
                h[u][3] += py * qx; h[u][4] += py * qy; h[u][5] += py * qz;
                h[u][6] += pz * qx; h[u][7] += pz * qy; h[u][8] += pz * qz;
            }
        }
    }
    unsigned long long evals = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        if (!pair_ok[u]) continue;
        const double e0 = a.g[s_row[u]] + a.g[s_col];
        // cheap bound: sum of singular values <= sqrt(3) |H|_F  =>  rmsd^2 >= (e0 - 2 sqrt(3)|H|_F) / nh
        double f2 = 0;
#pragma unroll
        for (int e = 0; e < 9; ++e) f2 += h[u][e] * h[u][e];
        const double thr2 = (a.max_rmsd + a.eps) * (a.max_rmsd + a.eps);
        if ((e0 - 2.0 * sqrt(3.0 * f2)) / nh >= thr2) continue;
        {   // closed-form singular values: clearly dissimilar pairs never reach the Jacobi solve
            const double lim = a.max_rmsd + 1e-4;
            if ((e0 - 2.0 * singular_sum3(h[u])) / nh > lim * lim) continue;
        }
        ++evals;
        double sig[3];
        M3 R = kabsch_from_cov(h[u], sig);
        double msd = (e0 - 2.0 * (sig[0] + sig[1] + sig[2])) / nh;
        if (msd < 0) msd = 0;
        if (msd >= thr2) continue;
        // the decision itself restates rmsd_and_max: rotate, difference, sum of squares, max norm
        const double* p = a.xc + (size_t)s_row[u] * nh * 3;
        const double* q = a.xc + (size_t)s_col * nh * 3;
        double ss = 0, mx = 0;
        for (int k = 0; k < nh; ++k) {
            double x = p[3 * k], y = p[3 * k + 1], z = p[3 * k + 2];
            double dx = (x * R.m[0] + y * R.m[3] + z * R.m[6]) - q[3 * k];
            double dy = (x * R.m[1] + y * R.m[4] + z * R.m[7]) - q[3 * k + 1];
            double dz = (x * R.m[2] + y * R.m[5] + z * R.m[8]) - q[3 * k + 2];
            double d2 = dx * dx + dy * dy + dz * dz;
            ss += d2;
            mx = fmax(mx, d2);
        }
        const double rmsd = sqrt(ss / nh), maxdev = sqrt(mx);
        const bool r_ok = rmsd < a.max_rmsd, m_ok = maxdev < a.max_dev;
        if (a.ties) {
            if (fabs(rmsd - a.max_rmsd) <= a.eps) {
                int slot = atomicAdd(a.n_ties, 1);
                if (slot < a.tie_cap) a.ties[slot] = TieRecord{s_col, s_row[u], rmsd, FC_TIE_RMSD, r_ok ? 1 : 0};
            }
            if (r_ok && fabs(maxdev - a.max_dev) <= a.eps) {
                int slot = atomicAdd(a.n_ties, 1);
                if (slot < a.tie_cap) a.ties[slot] = TieRecord{s_col, s_row[u], maxdev, FC_TIE_MAXDEV, m_ok ? 1 : 0};
            }
        }
        if (r_ok && m_ok) prune_push_pair(a, s_row[u], s_col);
    }
    if (a.n_eval) {
        for (int o = 16; o > 0; o >>= 1) evals += __shfl_xor_sync(0xffffffffu, evals, o);
        if (tx == 0 && evals) atomicAdd(a.n_eval, evals);
    }
}


// ---------------------------------------------------------------------------------------------
// FP32 screen + FP64 exact evaluation of the pairs it cannot rule out
// ---------------------------------------------------------------------------------------------
// closed-form UPPER BOUND on the singular-value sum in FP32 (see singular_sum3): the screen rejects a pair only when
// even this bound leaves the RMSD above the threshold by kScreenBand
__device__ __forceinline__ float singular_sum3f(const float* h) {
    float k0 = h[0] * h[0] + h[3] * h[3] + h[6] * h[6], k1 = h[0] * h[1] + h[3] * h[4] + h[6] * h[7];
    float k2 = h[0] * h[2] + h[3] * h[5] + h[6] * h[8], k3 = h[1] * h[1] + h[4] * h[4] + h[7] * h[7];
    float k4 = h[1] * h[2] + h[4] * h[5] + h[7] * h[8], k5 = h[2] * h[2] + h[5] * h[5] + h[8] * h[8];
    float e1, e2, e3;
    const float p1 = k1 * k1 + k2 * k2 + k4 * k4;
    const float q = (k0 + k3 + k5) * (1.0f / 3.0f);
    const float b0 = k0 - q, b3 = k3 - q, b5 = k5 - q;
    const float p2 = b0 * b0 + b3 * b3 + b5 * b5 + 2.0f * p1;
    if (!(p2 > 0.f)) {
        e1 = e2 = e3 = q;
    } else {
        const float p = sqrtf(p2 * (1.0f / 6.0f)), ip = 1.0f / p;
        const float c0 = b0 * ip, c3 = b3 * ip, c5 = b5 * ip, c1 = k1 * ip, c2 = k2 * ip, c4 = k4 * ip;
        float r = 0.5f * (c0 * (c3 * c5 - c4 * c4) - c1 * (c1 * c5 - c4 * c2) + c2 * (c1 * c4 - c3 * c2));
        r = fminf(1.0f, fmaxf(-1.0f, r));
        const float phi = acosf(r) * (1.0f / 3.0f);
        e1 = q + 2.0f * p * cosf(phi);
        e3 = q + 2.0f * p * cosf(phi + 2.0943951f);
        e2 = 3.0f * q - e1 - e3;
    }
    // An UPPER bound is what a screen may reject on: the smallest singular value is ADDED whatever the sign of det H
    // (for near-planar frames e3 is FP32 noise and the sign of an FP32 determinant is a coin toss; subtracting s3 there
    // would lower the sum by ~1e-3 sigma_1 and could screen out a truly similar pair), and the FP32 error of the
    // closed form (eigenvalues good to ~1e-6 sigma_1^2, i.e. ~1e-3 sigma_1 on the smallest root) is added explicitly.
    const float s1 = sqrtf(fmaxf(e1, 0.f));
    return s1 + sqrtf(fmaxf(e2, 0.f)) + sqrtf(fmaxf(e3, 0.f)) + 2e-3f * s1;
}

constexpr float kScreenBand = 0.05f;   // A: FP32 covariance + FP32 closed form are good to a few 1e-3 A
#define PS_ATOMS 64                    // atoms staged per step
#define PS_LD 33                       // row stride (float4) of the staged tiles: 32 structures + 1 pad

// 128 threads per 32 x 32 pair tile: thread (tc, tr) owns rows 4 tr .. 4 tr + 3 and columns 2 tc, 2 tc + 1
// (8 pairs, 72 FP32 accumulators); per atom 4 broadcast LDS.128 (rows) + 2 LDS.128 (columns) feed 72 FFMA.
__global__ void __launch_bounds__(128) prune_screen_f32_kernel(PruneArgs a) {
    const PruneTile t = a.tiles[blockIdx.x];
    const int tc = threadIdx.x & 15, tr = threadIdx.x >> 4;
    const int chunk_end = t.chunk_begin + t.chunk_len;
    extern __shared__ float4 s_stage[];  // [PS_ATOMS][PS_LD] rows, then [PS_ATOMS][PS_LD] columns
    float4* sR = s_stage;
    float4* sC = s_stage + PS_ATOMS * PS_LD;
    float h[8][9];
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int e = 0; e < 9; ++e) h[u][e] = 0.f;
    const int nh = a.nh;
    for (int a0 = 0; a0 < nh; a0 += PS_ATOMS) {
        const int na = min(PS_ATOMS, nh - a0);
        __syncthreads();
        // a warp stages one structure at a time: lanes read consecutive atoms (coalesced) and write with a row
        // stride of PS_LD = 33 float4, so the 8 lanes of a quarter-warp hit 8 different 4-bank groups
        for (int sidx = threadIdx.x >> 5; sidx < PR_TS; sidx += 4) {
            const int rpos = t.row0 + sidx, cpos = t.col0 + sidx;
            const float4* rsrc = rpos < chunk_end ? a.xcf + (size_t)a.active[rpos] * nh + a0 : nullptr;
            const float4* csrc = cpos < chunk_end ? a.xcf + (size_t)a.active[cpos] * nh + a0 : nullptr;
            for (int k = threadIdx.x & 31; k < na; k += 32) {
                sR[k * PS_LD + sidx] = rsrc ? rsrc[k] : make_float4(0.f, 0.f, 0.f, 0.f);
                sC[k * PS_LD + sidx] = csrc ? csrc[k] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        __syncthreads();
#pragma unroll 2
        for (int k = 0; k < na; ++k) {
            const float4 q0 = sC[k * PS_LD + 2 * tc], q1 = sC[k * PS_LD + 2 * tc + 1];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 pr = sR[k * PS_LD + 4 * tr + i];
                float* h0 = h[2 * i];
                float* h1 = h[2 * i + 1];
                h0[0] = fmaf(pr.x, q0.x, h0[0]); h0[1] = fmaf(pr.x, q0.y, h0[1]); h0[2] = fmaf(pr.x, q0.z, h0[2]);
                h0[3] = fmaf(pr.y, q0.x, h0[3]); h0[4] = fmaf(pr.y, q0.y, h0[4]); h0[5] = fmaf(pr.y, q0.z, h0[5]);
                h0[6] = fmaf(pr.z, q0.x, h0[6]); h0[7] = fmaf(pr.z, q0.y, h0[7]); h0[8] = fmaf(pr.z, q0.z, h0[8]);
                h1[0] = fmaf(pr.x, q1.x, h1[0]); h1[1] = fmaf(pr.x, q1.y, h1[1]); h1[2] = fmaf(pr.x, q1.z, h1[2]);
                h1[3] = fmaf(pr.y, q1.x, h1[3]); h1[4] = fmaf(pr.y, q1.y, h1[4]); h1[5] = fmaf(pr.y, q1.z, h1[5]);
                h1[6] = fmaf(pr.z, q1.x, h1[6]); h1[7] = fmaf(pr.z, q1.y, h1[7]); h1[8] = fmaf(pr.z, q1.z, h1[8]);
            }
        }
    }
    const float lim = (float)a.max_rmsd + kScreenBand;
    const float thr2 = lim * lim;
    const float inv_nh = 1.0f / (float)nh;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int row = t.row0 + 4 * tr + (u >> 1), col = t.col0 + 2 * tc + (u & 1);
        if (!(row < chunk_end && col < chunk_end && row < col)) continue;  // each unordered pair once
        const int s_row = a.active[row], s_col = a.active[col];
        if (a.energies && !(fabs(a.energies[s_row] - a.energies[s_col]) < a.max_dE)) continue;
        const float e0 = (float)(a.g[s_row] + a.g[s_col]);
        // cheap bound first: sum of singular values <= sqrt(3) |H|_F
        float f2 = 0.f;
#pragma unroll
        for (int e = 0; e < 9; ++e) f2 = fmaf(h[u][e], h[u][e], f2);
        if ((e0 - 2.0f * sqrtf(3.0f * f2)) * inv_nh >= thr2) continue;
        if ((e0 - 2.0f * singular_sum3f(h[u])) * inv_nh > thr2) continue;
        const unsigned long long slot = atomicAdd(a.n_cand, 1ull);
        if ((long long)slot < a.cand_cap) a.cand[slot] = make_int2(s_row, s_col);
    }
}

__device__ __forceinline__ double prune_wsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one warp per candidate pair: the FP64 evaluation that restates rmsd_and_max (centred heavy atoms)
__device__ __forceinline__ void prune_exact_one(const PruneArgs& a, long long c, int lane) {
    const int2 pr = a.cand[c];
    const int nh = a.nh;
    const double* p = a.xc + (size_t)pr.x * nh * 3;
    const double* q = a.xc + (size_t)pr.y * nh * 3;
    double h[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = lane; k < nh; k += 32) {
        const double px = p[3 * k], py = p[3 * k + 1], pz = p[3 * k + 2];
        const double qx = q[3 * k], qy = q[3 * k + 1], qz = q[3 * k + 2];
        h[0] += px * qx; h[1] += px * qy; h[2] += px * qz;
        h[3] += py * qx; h[4] += py * qy; h[5] += py * qz;
        h[6] += pz * qx; h[7] += pz * qy; h[8] += pz * qz;
    }
#pragma unroll
    for (int e = 0; e < 9; ++e) h[e] = prune_wsum(h[e]);
    const double e0 = a.g[pr.x] + a.g[pr.y];
    const double thr2 = (a.max_rmsd + a.eps) * (a.max_rmsd + a.eps);
    {
        const double lim = a.max_rmsd + 1e-4;
        if ((e0 - 2.0 * singular_sum3(h)) / nh > lim * lim) return;
    }
    if (lane == 0 && a.n_eval) atomicAdd(a.n_eval, 1ull);
    double sig[3];
    M3 R = kabsch_from_cov(h, sig);
    double msd = (e0 - 2.0 * (sig[0] + sig[1] + sig[2])) / nh;
    if (msd < 0) msd = 0;
    if (msd >= thr2) return;
    double ss = 0, mx = 0;
    for (int k = lane; k < nh; k += 32) {
        const double x = p[3 * k], y = p[3 * k + 1], z = p[3 * k + 2];
        const double dx = (x * R.m[0] + y * R.m[3] + z * R.m[6]) - q[3 * k];
        const double dy = (x * R.m[1] + y * R.m[4] + z * R.m[7]) - q[3 * k + 1];
        const double dz = (x * R.m[2] + y * R.m[5] + z * R.m[8]) - q[3 * k + 2];
        const double d2 = dx * dx + dy * dy + dz * dz;
        ss += d2;
        mx = fmax(mx, d2);
    }
    ss = prune_wsum(ss);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane != 0) return;
    const double rmsd = sqrt(ss / nh), maxdev = sqrt(mx);
    const bool r_ok = rmsd < a.max_rmsd, m_ok = maxdev < a.max_dev;
    if (a.ties) {
        if (fabs(rmsd - a.max_rmsd) <= a.eps) {
            int slot = atomicAdd(a.n_ties, 1);
            if (slot < a.tie_cap) a.ties[slot] = TieRecord{pr.y, pr.x, rmsd, FC_TIE_RMSD, r_ok ? 1 : 0};
        }
        if (r_ok && fabs(maxdev - a.max_dev) <= a.eps) {
            int slot = atomicAdd(a.n_ties, 1);
            if (slot < a.tie_cap) a.ties[slot] = TieRecord{pr.y, pr.x, maxdev, FC_TIE_MAXDEV, m_ok ? 1 : 0};
        }
    }
    if (r_ok && m_ok) prune_push_pair(a, pr.x, pr.y);
}

// The number of candidates is read on the device (the screen's counter), so the host does not wait between the
// screen and this launch.  A list that overflowed is left alone: the host repeats the pass with a larger one.
__global__ void __launch_bounds__(256) prune_exact_kernel(PruneArgs a) {
    const long long n_cand = (long long)*a.n_cand;
    if (n_cand > a.cand_cap) return;
    const int lane = threadIdx.x & 31;
    const long long step = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long c = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < n_cand; c += step)
        prune_exact_one(a, c, lane);
}

}  // namespace fc

using namespace fc;



static const int64_t kSchedule[] = {500000, 200000, 100000, 50000, 20000, 10000, 5000, 2000, 1000,
                                    500,    200,    100,    50,    20,    10,    5,    2,    1};

// Ordered resolution of one pass on the host: O(similar pairs).  Chunks are contiguous index ranges and
// pairs never cross chunks, so one sweep in index order resolves every chunk.
//   greedy   : NMS sweep -- a structure that is still alive drops every similar structure on its
//              droppable side (later ones for keep-first, earlier ones for keep-last);
//   snapshot : a structure is dropped if any structure on its keeper side that was active at the start of
//              the pass is similar to it.
struct ResolveScratch {
    std::vector<int> off, adj;  // CSR of the similar pairs by head, reused from pass to pass
};

// A pair is {.x = droppable structure, .y = head}: the head, while alive, drops the other one (PruneArgs::swap_xy).
// `sorted`: the list is ordered by head (ascending), as the single-rank driver gets it from the device sort -- one linear
// scan, forwards for keep-first (heads are the earlier structures) and backwards for keep-last.  Otherwise (lists
// gathered from several ranks) the pairs are bucketed by head first.
static void prune_resolve(std::vector<uint8_t>& mask, const std::vector<int2>& pairs, bool keep_first, bool snapshot,
                          bool sorted, ResolveScratch& sc) {
    if (pairs.empty()) return;
    if (snapshot) {
        for (const int2& p : pairs) mask[(size_t)p.x] = 0;
        return;
    }
    if (sorted) {
        if (keep_first) {
            for (const int2& p : pairs)
                if (mask[(size_t)p.y]) mask[(size_t)p.x] = 0;
        } else {
            for (size_t e = pairs.size(); e-- > 0;)
                if (mask[(size_t)pairs[e].y]) mask[(size_t)pairs[e].x] = 0;
        }
        return;
    }
    // heads lie in [lo, hi]: the counting and the sweep stay inside that range
    int lo = (int)mask.size(), hi = -1;
    for (const int2& p : pairs) {
        lo = std::min(lo, p.y);
        hi = std::max(hi, p.y);
    }
    const size_t span = (size_t)(hi - lo) + 2;
    if (sc.off.size() < span) sc.off.resize(span);
    if (sc.adj.size() < pairs.size()) sc.adj.resize(pairs.size());
    int* off = sc.off.data();
    int* adj = sc.adj.data();
    memset(off, 0, span * sizeof(int));
    for (const int2& p : pairs) off[p.y - lo + 1] += 1;
    for (size_t i = 1; i < span; ++i) off[i] += off[i - 1];
    // fill backwards: afterwards off[i + 1] holds the START of row i; its end is the start of row i + 1
    for (const int2& p : pairs) adj[off[p.y - lo + 1]-- - 1] = p.x;
    auto row_begin = [&](size_t i) { return off[i + 1]; };
    auto row_end = [&](size_t i) { return i + 2 < span ? off[i + 2] : (int)pairs.size(); };
    if (keep_first) {
        for (size_t i = 0; i + 1 < span; ++i)
            if (mask[(size_t)lo + i])
                for (int e = row_begin(i); e < row_end(i); ++e) mask[(size_t)adj[e]] = 0;
    } else {
        for (size_t i = span - 1; i-- > 0;)
            if (mask[(size_t)lo + i])
                for (int e = row_begin(i); e < row_end(i); ++e) mask[(size_t)adj[e]] = 0;
    }
}

// pairs (r, c) with r in [row0, min(row0 + 128, pend)), c in [max(col_lo, r + 1), col_hi): what one work item of the
// tensor-core screen evaluates
static int64_t gram_item_pairs(int64_t row0, int64_t col_lo, int64_t col_hi, int64_t pend) {
    const int64_t ra = row0, rb = std::min(row0 + 128, pend);
    if (col_hi <= col_lo) return 0;
    int64_t total = 0;
    const int64_t n1 = std::max<int64_t>(0, std::min(rb, col_lo) - ra);  // rows entirely left of the column range
    total += n1 * (col_hi - col_lo);
    const int64_t r1 = std::max(ra, col_lo), r2 = std::min(rb, col_hi - 1);  // rows inside it: col_hi - 1 - r columns each
    if (r2 > r1) total += (r2 - r1) * (col_hi - 1) - (r1 + r2 - 1) * (r2 - r1) / 2;
    return total;
}

// One pass of the tensor-core screen, planned on the host: the active structures (mask != 0) of each of the k chunks
// (n / k structures, the last takes the rest) get padded positions (every chunk starts at a multiple of 16) and, per
// 128-row block, the column tiles that still need evaluating.  Two survivors of a common chunk of the previous pass
// (chunk size prev_size, prev_k chunks; prev_size = 0 on the first pass) are known dissimilar, so the tiles whose
// structures all shared the previous chunk of the block's first row -- a prefix of the block's column range -- are
// left out.  Column ranges are cut into work items so that every SM gets several of similar size; item i goes to rank
// i % world.  Adds to pairs_tiled (pairs this rank evaluates) and pairs_skipped (pairs no rank evaluates).
// Chunk boundaries (k + 1 structure indices) of a pass.  "Full": k chunks of n / k consecutive STRUCTURES, the last takes
// the rest -- what the oracle shim and FIRECODE's in-tree sibling driver do (torsion_module.py:973-1041).  "Active": k
// chunks of n_active / k consecutive ACTIVE structures (SURVEY.md 8c's wording of the prism_pruner contract); as index
// ranges: chunk c starts at the (c * size)-th active structure.  Which one prism_pruner 0.0.7 uses is unpinned:
// conventions.PRUNE_CHUNK_OVER.
static std::vector<int64_t> chunk_bounds(const std::vector<uint8_t>& mask, int64_t n, int64_t k, bool over_active) {
    std::vector<int64_t> b((size_t)k + 1, n);
    if (!over_active) {
        const int64_t size = n / k;
        for (int64_t c = 0; c < k; ++c) b[(size_t)c] = c * size;
        return b;
    }
    int64_t n_active = 0;
    for (uint8_t m : mask) n_active += m ? 1 : 0;
    const int64_t size = std::max<int64_t>(1, n_active / k);
    int64_t seen = 0, c = 0;
    b[0] = 0;
    for (int64_t i = 0; i < n && c + 1 < k; ++i) {
        if (!mask[(size_t)i]) continue;
        if (seen == (c + 1) * size) b[(size_t)++c] = i;
        ++seen;
    }
    // chunks that found no active structure to start at stay empty ([n, n))
    return b;
}

static inline int64_t chunk_of(const std::vector<int64_t>& bounds, int64_t idx) {
    return (int64_t)(std::upper_bound(bounds.begin(), bounds.end() - 1, idx) - bounds.begin()) - 1;
}

// `seg` (optional): per padded position, the number of the segment it belongs to -- a run of positions whose structures
// shared a chunk of the previous pass (the whole chunk on the first pass); numbers increase along the positions and the
// padding behind a chunk carries the number of the chunk's last segment.  The order of the positions INSIDE a segment
// is free (pairs inside it are never evaluated, or -- first pass -- only by position order), which the driver uses to
// sort them by norm on the device so that whole tiles can be culled.
static void plan_gram_pass(const std::vector<uint8_t>& mask, int64_t n, const std::vector<int64_t>& bounds,
                           const std::vector<int64_t>& prev_bounds, int world,
                           int rank, int n_sms, std::vector<int>& spos, std::vector<GramWork>& work, int64_t& pairs_tiled,
                           int64_t& pairs_skipped, std::vector<int>* seg = nullptr) {
    struct RowBlock { int row0, c_min, tile_end, pend; };
    std::vector<RowBlock> blocks;
    const int64_t k = (int64_t)bounds.size() - 1;
    spos.clear();
    spos.reserve((size_t)n + 16 * (size_t)k + 256);
    work.clear();
    int seg_no = 0;
    size_t seg_pc = 0;  // previous-pass chunk of the structure the segment numbering has reached
    if (seg) {
        seg->clear();
        seg->reserve((size_t)n + 16 * (size_t)k + 256);
    }
    int64_t col_tiles_total = 0;
    for (int64_t c = 0; c < k; ++c) {
        const int64_t first = bounds[(size_t)c], last = bounds[(size_t)c + 1];
        while (spos.size() % 16) spos.push_back(-1);
        const int pbegin = (int)spos.size();
        {   // branch-free compaction of the chunk's active structures
            spos.resize((size_t)pbegin + (size_t)(last - first));
            int* out = spos.data() + pbegin;
            const uint8_t* m = mask.data();
            size_t w = 0;
            for (int64_t i = first; i < last; ++i) {
                out[w] = (int)i;
                w += m[i] != 0;
            }
            spos.resize((size_t)pbegin + w);
        }
        const int pend = (int)spos.size(), len = pend - pbegin;
        if (seg) {
            seg->resize((size_t)pbegin, seg_no);   // the padding in front of this chunk: last segment of the chunk before
            seg->resize((size_t)pend);
            int* sg = seg->data();
            if (prev_bounds.empty()) {
                if (len > 0) std::fill(sg + pbegin, sg + pend, ++seg_no);
            } else {
                // runs of positions whose structures shared a chunk of the previous pass (structures ascend over the whole
                // pass, so the pointer into prev_bounds only advances)
                for (int q = pbegin; q < pend;) {
                    const int64_t idx = spos[(size_t)q];
                    while (seg_pc + 2 < prev_bounds.size() && prev_bounds[seg_pc + 1] <= idx) ++seg_pc;
                    const int64_t idx_end = prev_bounds[seg_pc + 1];
                    const int q_end = (int)(std::lower_bound(spos.begin() + q, spos.begin() + pend, idx_end,
                                                             [](int v, int64_t lim) { return (int64_t)v < lim; }) - spos.begin());
                    std::fill(sg + q, sg + q_end, ++seg_no);
                    q = q_end;
                }
            }
        }
        if (len < 2) continue;
        const int tile_end = (pend + 15) / 16;
        int64_t evaluated = 0;
        for (int row0 = pbegin; row0 < pend - 1; row0 += 128) {
            int c_min = row0 / 16;
            if (!prev_bounds.empty()) {
                // first position of the chunk whose structure lies beyond the previous-pass chunk of the block's first
                // row: a tile is known dissimilar iff its last structure comes before that position
                const int64_t pc = chunk_of(prev_bounds, spos[(size_t)row0]);
                const int64_t idx_end = prev_bounds[(size_t)pc + 1];
                const int pos_e = (int)(std::lower_bound(spos.begin() + row0, spos.begin() + pend, idx_end,
                                                         [](int v, int64_t lim) { return (int64_t)v < lim; }) -
                                        spos.begin());
                c_min = pos_e >= pend ? tile_end : std::max(c_min, pos_e / 16);
            }
            if (c_min >= tile_end) continue;
            blocks.push_back(RowBlock{row0, c_min, tile_end, pend});
            col_tiles_total += tile_end - c_min;
            evaluated += gram_item_pairs(row0, 16 * c_min, pend, pend);
        }
        pairs_skipped += (int64_t)len * (len - 1) / 2 - evaluated;
    }
    // items per SM: FC_PRUNE_ITEMS_PER_SM overrides (experiments); more items = better balance, more A reloads
    static const int64_t items_per_sm = []() {
        const char* v = getenv("FC_PRUNE_ITEMS_PER_SM");
        return v && atoll(v) > 0 ? (int64_t)atoll(v) : (int64_t)32;
    }();
    const int64_t item_tiles = std::min<int64_t>(1024, std::max<int64_t>(16, col_tiles_total / ((int64_t)n_sms * items_per_sm)));
    int64_t item_no = 0;
    for (const RowBlock& b : blocks)
        for (int c0 = b.c_min; c0 < b.tile_end; c0 += (int)item_tiles) {
            const int nt = (int)std::min<int64_t>(item_tiles, b.tile_end - c0);
            if (item_no++ % world != rank) continue;
            work.push_back(GramWork{b.row0, c0, nt, b.pend});
            pairs_tiled += gram_item_pairs(b.row0, 16 * c0, std::min(16 * (c0 + nt), b.pend), b.pend);
        }
    while (spos.size() % 16) spos.push_back(-1);
    spos.resize(spos.size() + 128, -1);  // a row block may read 128 positions from its first row
    if (seg) seg->resize(spos.size(), seg_no + 1);  // trailing padding: a segment of its own, behind everything
}

// Host-only view of plan_gram_pass for the tests (no CUDA call): positions and work items of one pass.
extern "C" int fc_prune_plan(const uint8_t* mask, int64_t n, int64_t k, int64_t prev_k, int32_t world, int32_t rank, int32_t n_sms,
                             int32_t chunk_over_active, int32_t* spos_out, int64_t spos_cap, int64_t* n_spos, int32_t* work_out,
                             int64_t work_cap, int64_t* n_work, int64_t* counts_out) {
    FC_REQUIRE(mask && n > 0 && k >= 1 && k <= n && prev_k >= 0 && prev_k <= n && world >= 1 && rank >= 0 && rank < world && n_sms >= 1,
               "fc_prune_plan: bad arguments");
    FC_REQUIRE(n_spos && n_work && counts_out, "fc_prune_plan: null pointer");
    std::vector<uint8_t> m(mask, mask + n);
    std::vector<int> spos;
    std::vector<GramWork> work;
    int64_t tiled = 0, skipped = 0;
    // the previous pass is planned on the same mask here (the hook has no history): enough to test the coverage property
    const std::vector<int64_t> bounds = chunk_bounds(m, n, k, chunk_over_active != 0);
    const std::vector<int64_t> prev = prev_k ? chunk_bounds(m, n, prev_k, chunk_over_active != 0) : std::vector<int64_t>();
    plan_gram_pass(m, n, bounds, prev, world, rank, n_sms, spos, work, tiled, skipped);
    *n_spos = (int64_t)spos.size();
    *n_work = (int64_t)work.size();
    counts_out[0] = tiled;
    counts_out[1] = skipped;
    FC_REQUIRE((int64_t)spos.size() <= spos_cap && (int64_t)work.size() <= work_cap && spos_out && work_out,
               "fc_prune_plan: buffers too small (%lld positions, %lld items)", (long long)spos.size(), (long long)work.size());
    memcpy(spos_out, spos.data(), spos.size() * sizeof(int));
    for (size_t i = 0; i < work.size(); ++i) {
        work_out[4 * i] = work[i].row0; work_out[4 * i + 1] = work[i].col_tile0;
        work_out[4 * i + 2] = work[i].n_col_tiles; work_out[4 * i + 3] = work[i].pend;
    }
    return FC_OK;
}

static thread_local double g_prune_timing[6] = {0, 0, 0, 0, 0, 0};

// Host-only test hook: the segment number of every padded position of the same pass (see plan_gram_pass).
extern "C" int fc_prune_plan_segments(const uint8_t* mask, int64_t n, int64_t k, int64_t prev_k, int32_t chunk_over_active,
                                      int32_t* seg_out, int64_t seg_cap, int64_t* n_seg) {
    FC_REQUIRE(mask && n > 0 && k >= 1 && k <= n && prev_k >= 0 && prev_k <= n && n_seg, "fc_prune_plan_segments: bad arguments");
    std::vector<uint8_t> m(mask, mask + n);
    std::vector<int> spos, seg;
    std::vector<GramWork> work;
    int64_t tiled = 0, skipped = 0;
    const std::vector<int64_t> bounds = chunk_bounds(m, n, k, chunk_over_active != 0);
    const std::vector<int64_t> prev = prev_k ? chunk_bounds(m, n, prev_k, chunk_over_active != 0) : std::vector<int64_t>();
    plan_gram_pass(m, n, bounds, prev, 1, 0, 4, spos, work, tiled, skipped, &seg);
    *n_seg = (int64_t)seg.size();
    FC_REQUIRE((int64_t)seg.size() <= seg_cap && seg_out, "fc_prune_plan_segments: buffer too small (%lld positions)", (long long)seg.size());
    memcpy(seg_out, seg.data(), seg.size() * sizeof(int));
    return FC_OK;
}

static thread_local double g_prune_tiles[2] = {0, 0};

extern "C" int fc_prune_tiles(double* out2) {
    FC_REQUIRE(out2, "fc_prune_tiles: null pointer");
    out2[0] = g_prune_tiles[0];
    out2[1] = g_prune_tiles[1];
    return FC_OK;
}

extern "C" int fc_prune_timing(double* out6) {
    FC_REQUIRE(out6, "fc_prune_timing: null pointer");
    for (int i = 0; i < 6; ++i) out6[i] = g_prune_timing[i];
    return FC_OK;
}

// mode 0: RMSD, mode 1: MOI.  `sel` = indices of the atoms used for the RMSD (heavy atoms).
// rank / world / gather: pair tiles of every pass are dealt round-robin to the ranks; the similar pairs each
// rank finds are all-gathered through `gather` (NCCL or gloo behind the host language) and every rank resolves
// the pass on the union, so all ranks hold the same mask after every pass.
// gather      : host all-gather of byte buffers (structures replicated on every rank, lists staged through the host);
// gather_dev  : device all-gather of equal-sized pieces (NCCL over NVLink behind the host language): every rank uploads
//               only its 1 / world of the structures and the pieces are exchanged between the GPUs; per pass the ranks
//               exchange their counters (16 bytes each) and their similar-pair lists device to device, the union is
//               sorted by head on the device and every rank resolves the same list.
static int prune_impl(const double* structures, int64_t n, int32_t n_atoms, int32_t mode, const int32_t* sel,
                      int32_t n_sel, const double* masses, double max_rmsd, double max_dev, double moi_dev,
                      const double* energies, double max_dE, int32_t keep_first, int32_t snapshot,
                      int32_t min_per_chunk, uint8_t* mask_out, int64_t* stats_out, fc_tie* ties_out,
                      int64_t tie_cap, int64_t* n_ties_out, int32_t rank, int32_t world,
                      fc_allgather_fn gather, fc_allgather_dev_fn gather_dev, void* gather_ctx) {
    FC_REQUIRE(n >= 0 && n < ((int64_t)1 << 31) && n_atoms > 0, "fc_prune: bad sizes");
    FC_REQUIRE(world >= 1 && rank >= 0 && rank < world, "fc_prune: bad rank / world");
    FC_REQUIRE(world == 1 || gather || gather_dev, "fc_prune: a multi-rank call needs an all-gather callback");
    const bool dev_gather = world > 1 && gather_dev != nullptr;
    if (n_ties_out) *n_ties_out = 0;
    if (stats_out) stats_out[0] = stats_out[1] = stats_out[2] = stats_out[3] = 0;
    if (n == 0) return FC_OK;
    FC_REQUIRE(structures && mask_out, "fc_prune: null pointer");
    FC_REQUIRE(mode == 0 || mode == 1, "fc_prune: unknown mode %d", mode);
    if (mode == 0) {
        FC_REQUIRE(sel && n_sel > 0, "fc_prune: RMSD pruning needs at least one selected atom");
        for (int k = 0; k < n_sel; ++k) FC_REQUIRE(sel[k] >= 0 && sel[k] < n_atoms, "fc_prune: atom selection out of range");
    } else {
        FC_REQUIRE(masses, "fc_prune: MOI pruning needs masses");
    }
    sm_count();
    const bool trace = getenv("FC_PRUNE_TRACE") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    double t_tiles = 0, t_kernels = 0, t_resolve = 0, t_upload = 0, t_gram = 0;
    unsigned long long cand_total = 0;
    cudaEvent_t ev_g0 = nullptr, ev_g1 = nullptr;
    cudaEventCreate(&ev_g0);
    cudaEventCreate(&ev_g1);
    double screen_slots = 0;
    long long tiles_planned = 0;
    int64_t screen_launches = 0;
    // page-locked landing place of the per-pass counters (one per host thread, kept)
    struct PassCounters { unsigned long long n_cand, found, tiles_done; int gram_err; };
    static thread_local PassCounters* hb = nullptr;
    if (!hb) FC_CUDA(cudaHostAlloc((void**)&hb, sizeof(PassCounters), cudaHostAllocPortable));
    cudaStream_t s;
    FC_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    int rc = FC_OK;
    std::vector<uint8_t> mask((size_t)n, 1);
    ResolveScratch resolve_scratch;
    int64_t pairs_tiled = 0, passes = 0, pairs_skipped = 0, similar_total = 0;
    unsigned long long evals_total = 0;
    int64_t ties_total = 0;
    {
        DevBuf<double> d_coords, d_xc, d_g, d_moi, d_mass, d_energy;
        DevBuf<int> d_sel, d_active, d_nties;
        DevBuf<int2> d_pairs, d_cand, d_pairs_sorted, d_pairs_all;
        DevBuf<uint8_t> d_sort_tmp;
        DevBuf<unsigned long long> d_counts;       // device all-gather: {similar pairs, candidates} of every rank
        std::vector<unsigned long long> h_counts;
        DevBuf<float4> d_xcf;
        DevBuf<PruneTile> d_tiles;
        DevBuf<float> d_img, d_gp;       // tensor-core screen: operand image and |x|^2 per padded position
        DevBuf<int> d_spos, d_gram_err, d_spos_raw, d_seg;
        DevBuf<unsigned long long> d_keys;
        DevBuf<float4> d_tile_norm, d_sig;   // per tile: two float4 {lo1, hi1, lo2, hi2}, {lo3, hi3, -, -}; per structure: sigma
        std::vector<int> seg;
        DevBuf<GramWork> d_work;
        DevBuf<unsigned long long> d_eval;
        DevBuf<TieRecord> d_ties;
        const int cap = (int)std::min<int64_t>(std::max<int64_t>(tie_cap, 1), 1 << 22);
        cudaError_t e = cudaSuccess;
#define PR(call) do { if (e == cudaSuccess) e = (call); } while (0)
        // only the atoms the criterion reads travel: the selected (heavy) atoms for the RMSD, all of them for the moments
        const int n_up = mode == 0 ? n_sel : n_atoms;
        if (!dev_gather) {
            PR(d_coords.alloc((size_t)n * n_up * 3, s));
            PR(upload_rows_staged(d_coords.p, structures, n, n_atoms, mode == 0 ? sel : nullptr, n_up, s));
        } else {
            // this rank's 1 / world of the rows travels over PCIe; the pieces are exchanged between the GPUs
            const int64_t per = (n + world - 1) / world, lo = std::min<int64_t>(n, per * rank), hi = std::min<int64_t>(n, lo + per);
            const size_t piece = (size_t)per * n_up * 3;
            DevBuf<double> d_part;
            PR(d_coords.alloc(piece * (size_t)world, s));
            PR(d_part.alloc(piece, s));
            PR(cudaMemsetAsync(d_part.p, 0, piece * 8, s));
            if (hi > lo)
                PR(upload_rows_staged(d_part.p, structures + (size_t)lo * n_atoms * 3, hi - lo, n_atoms, mode == 0 ? sel : nullptr, n_up, s));
            if (e == cudaSuccess && gather_dev(d_part.p, d_coords.p, (int64_t)(piece * 8), (void*)s, gather_ctx) != 0) {
                set_error("fc_prune: device all-gather of the structures failed");
                rc = FC_ERR_INVALID;
            }
            PR(cudaStreamSynchronize(s));  // d_part is released below: the exchange must have read it
        }
        PR(d_nties.alloc(4, s));
        PR(cudaMemsetAsync(d_nties.p, 0, 16, s));
        PR(d_eval.alloc(4, s));  // [0] eigen-solves, [1] similar pairs, [2] screen candidates of the current pass
        PR(cudaMemsetAsync(d_eval.p, 0, 32, s));
        PR(d_ties.alloc(cap, s));
        if (energies) {
            PR(d_energy.alloc((size_t)n, s));
            PR(cudaMemcpyAsync(d_energy.p, energies, (size_t)n * 8, cudaMemcpyHostToDevice, s));
        }
        if (mode == 0) {
            PR(d_sel.alloc(n_sel, s));
            PR(cudaMemcpyAsync(d_sel.p, sel, (size_t)n_sel * 4, cudaMemcpyHostToDevice, s));
            PR(d_xc.alloc((size_t)n * n_sel * 3, s));
            PR(d_xcf.alloc((size_t)n * n_sel, s));
            PR(d_g.alloc((size_t)n, s));
            PR(d_sig.alloc((size_t)n, s));
            if (e == cudaSuccess) {
                DevBuf<double> d_mom;
                e = d_mom.alloc((size_t)n * 6, s);
                if (e == cudaSuccess) {
                    prune_center_kernel<<<(unsigned)n, 64, 0, s>>>(d_coords.p, n_sel, nullptr, n_sel, n, d_xc.p, d_g.p, d_xcf.p, d_mom.p);
                    prune_sigma_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_mom.p, n, d_sig.p);
                    e = cudaGetLastError();
                }
            }
        } else {
            PR(d_mass.alloc(n_atoms, s));
            PR(cudaMemcpyAsync(d_mass.p, masses, (size_t)n_atoms * 8, cudaMemcpyHostToDevice, s));
            PR(d_moi.alloc((size_t)n * 3, s));
            if (e == cudaSuccess) {
                prune_moi_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_coords.p, d_mass.p, n_atoms, n, d_moi.p);
                e = cudaGetLastError();
            }
        }
        PR(d_active.alloc((size_t)n, s));
        if (trace) { PR(cudaStreamSynchronize(s)); t_upload = now() - t_begin; }
        if (e != cudaSuccess) rc = cuda_fail(e, "fc_prune setup", __FILE__, __LINE__);

        std::vector<int> active;
        std::vector<PruneTile> tiles;
        std::vector<int2> pairs, all_pairs;
        std::vector<int64_t> prev_bounds;  // chunking of the last executed pass (empty: none yet)
        const bool over_active = (snapshot & 2) != 0;  // FC_PRUNE_CHUNK_ACTIVE
        snapshot &= 1;
        // a single rank's list, and the union the device all-gather builds, are sorted by head on the device
        const bool sort_pairs = (world == 1 || dev_gather) && !snapshot;
        if (dev_gather && rc == FC_OK) {
            e = d_counts.alloc((size_t)2 * world, s);
            h_counts.assign((size_t)2 * world, 0);
            if (e != cudaSuccess) rc = cuda_fail(e, "fc_prune setup", __FILE__, __LINE__);
        }
        // screen flavour: FC_PRUNE_FP64=1 -> FP64 pair kernel only; otherwise FP32 screen + FP64 exact stage, the screen
        // on the tensor cores (gram_tc_kernel) unless FC_PRUNE_TC=0 or the molecule has more than 176 selected atoms (kGramMaxKc k-cores of 8 FP16 values)
        const char* env64 = getenv("FC_PRUNE_FP64");
        const char* envtc = getenv("FC_PRUNE_TC");
        const bool two_stage = mode == 0 && !(env64 && atoi(env64));
        // operands of the tensor-core screen: FP16 (8 atoms per 16-byte k-core) unless FC_PRUNE_TF32=1 (4 atoms)
        const char* envtf = getenv("FC_PRUNE_TF32");
        const int gram_tf32 = envtf && atoi(envtf) ? 1 : 0;
        // FC_PRUNE_CULL=0: positions stay in index order and every tile is multiplied (the round-2 start behaviour)
        const char* envcull = getenv("FC_PRUNE_CULL");
        const bool gram_cull = !(envcull && !atoi(envcull));
        int kc = gram_tf32 ? (n_sel + 3) / 4 : (n_sel + 7) / 8;
        kc += kc & 1;
        const bool use_tc = two_stage && kc <= kGramMaxKc && !(envtc && !atoi(envtc));
        std::vector<int> spos;
        std::vector<GramWork> work;
        if (use_tc && rc == FC_OK) {
            e = d_gram_err.alloc(1, s);
            PR(cudaMemsetAsync(d_gram_err.p, 0, 4, s));
            if (e != cudaSuccess) rc = cuda_fail(e, "fc_prune setup", __FILE__, __LINE__);
        }
        for (int64_t k : kSchedule) {
            if (rc) break;
            int64_t n_active = 0;
            for (uint8_t b : mask) n_active += b;
            if (!(k == 1 || (int64_t)min_per_chunk * k < n_active)) continue;
            ++passes;
            double tp = now();
            const std::vector<int64_t> bounds = chunk_bounds(mask, n, k, over_active);
            if (use_tc) {
                plan_gram_pass(mask, n, bounds, prev_bounds, world, rank, sm_count(), spos, work, pairs_tiled, pairs_skipped,
                               gram_cull ? &seg : nullptr);
            } else {
                // compacted active list + chunks over the FULL array (chunk size n // k, last takes the rest)
                active.clear();
                tiles.clear();
                int64_t tile_no = 0;
                for (int64_t c = 0; c < k; ++c) {
                    int64_t first = bounds[(size_t)c], last = bounds[(size_t)c + 1];
                    int begin = (int)active.size();
                    for (int64_t i = first; i < last; ++i)
                        if (mask[(size_t)i]) active.push_back((int)i);
                    int len = (int)active.size() - begin;
                    if (len < 2) continue;  // nothing to compare
                    for (int r0 = 0; r0 < len; r0 += PR_TS)
                        for (int c0 = r0; c0 < len; c0 += PR_TS) {
                            const int rows = std::min(PR_TS, len - r0), cols = std::min(PR_TS, len - c0);
                            const int64_t tile_pairs = r0 == c0 ? (int64_t)rows * (rows - 1) / 2 : (int64_t)rows * cols;
                            // survivors that shared a chunk in an earlier pass are known to be dissimilar
                            if (!prev_bounds.empty()) {
                                const int64_t lo_idx = active[(size_t)(begin + r0)], hi_idx = active[(size_t)(begin + c0 + cols - 1)];
                                if (chunk_of(prev_bounds, lo_idx) == chunk_of(prev_bounds, hi_idx)) { pairs_skipped += tile_pairs; continue; }
                            }
                            if (tile_no++ % world != rank) continue;
                            pairs_tiled += tile_pairs;
                            tiles.push_back(PruneTile{begin + r0, begin + c0, begin, len});
                        }
                }
            }
            prev_bounds = bounds;
            pairs.clear();
            t_tiles += now() - tp;
            tp = now();
            if (use_tc ? !work.empty() : !tiles.empty()) {
                const size_t n_act = use_tc ? spos.size() : active.size();
                e = cudaSuccess;
                if (use_tc) {
                    const size_t n_pos = spos.size(), img_floats = (n_pos / 8) * (size_t)3 * kc * 32;
                    if (d_spos.n < n_pos) { PR(d_spos.alloc(n_pos + n_pos / 4, s)); PR(d_gp.alloc(n_pos + n_pos / 4, s)); }
                    if (d_img.n < img_floats) PR(d_img.alloc(img_floats + img_floats / 4, s));
                    if (d_work.n < work.size()) PR(d_work.alloc(work.size() + work.size() / 4, s));
                    PR(cudaMemcpyAsync(d_work.p, work.data(), work.size() * sizeof(GramWork), cudaMemcpyHostToDevice, s));
                    if (!gram_cull) {
                        PR(cudaMemcpyAsync(d_spos.p, spos.data(), n_pos * 4, cudaMemcpyHostToDevice, s));
                    } else {
                        // positions ordered by norm inside every segment (device radix sort on {segment, norm}): row blocks and
                        // column tiles then cover narrow ranges of the norm and the screen skips the tiles that are too far apart
                        if (d_spos_raw.n < n_pos) {
                            PR(d_spos_raw.alloc(n_pos + n_pos / 4, s));
                            PR(d_seg.alloc(n_pos + n_pos / 4, s));
                            PR(d_keys.alloc(2 * (n_pos + n_pos / 4), s));
                            PR(d_tile_norm.alloc(2 * ((n_pos + n_pos / 4) / 16 + 16), s));
                        }
                        PR(cudaMemcpyAsync(d_spos_raw.p, spos.data(), n_pos * 4, cudaMemcpyHostToDevice, s));
                        PR(cudaMemcpyAsync(d_seg.p, seg.data(), n_pos * 4, cudaMemcpyHostToDevice, s));
                        unsigned long long* k_in = d_keys.p;
                        unsigned long long* k_out = d_keys.p + d_keys.n / 2;
                        if (e == cudaSuccess) {
                            gram_sort_key_kernel<<<(unsigned)((n_pos + 255) / 256), 256, 0, s>>>(d_spos_raw.p, d_seg.p, d_sig.p, (int)n_pos, k_in);
                            e = cudaGetLastError();
                        }
                        int seg_bits = 1;
                        while (((long long)1 << seg_bits) <= (long long)seg.back()) ++seg_bits;
                        size_t need = 0;
                        PR(cub::DeviceRadixSort::SortPairs(nullptr, need, k_in, k_out, d_spos_raw.p, d_spos.p, (int)n_pos, 0, 16 + seg_bits, s));
                        if (d_sort_tmp.n < need) PR(d_sort_tmp.alloc(need + need / 4 + 256, s));
                        need = d_sort_tmp.n;
                        PR(cub::DeviceRadixSort::SortPairs(d_sort_tmp.p, need, k_in, k_out, d_spos_raw.p, d_spos.p, (int)n_pos, 0, 16 + seg_bits, s));
                    }
                    if (e == cudaSuccess) {
                        gram_pack_kernel<<<(unsigned)(n_pos / 8), 256, 0, s>>>(d_xcf.p, d_g.p, d_spos.p, n_sel, kc, (int)(n_pos / 8),
                                                                              gram_tf32, d_img.p, d_gp.p);
                        e = cudaGetLastError();
                    }
                    if (gram_cull && e == cudaSuccess) {
                        gram_tile_norm_kernel<<<(unsigned)((n_pos / 16 + 127) / 128), 128, 0, s>>>(d_sig.p, d_spos.p, (int)(n_pos / 16), d_tile_norm.p);
                        e = cudaGetLastError();
                    }
                } else {
                    e = d_tiles.alloc(tiles.size(), s);
                    PR(cudaMemcpyAsync(d_tiles.p, tiles.data(), tiles.size() * sizeof(PruneTile), cudaMemcpyHostToDevice, s));
                    PR(cudaMemcpyAsync(d_active.p, active.data(), active.size() * 4, cudaMemcpyHostToDevice, s));
                }
                if (e != cudaSuccess) { rc = cuda_fail(e, "fc_prune pass setup", __FILE__, __LINE__); break; }
                long long pair_cap = std::max<long long>(1 << 20, 8 * (long long)n_act);
                long long cand_cap = std::max<long long>(1 << 21, 16 * (long long)n_act);
                bool counted = false;  // ties / eigen-solve statistics are recorded by one exact run only
                for (int attempt = 0; attempt < 4 && !rc; ++attempt) {
                    if (d_pairs.n < (size_t)pair_cap) PR(d_pairs.alloc((size_t)pair_cap, s));
                    if (two_stage && d_cand.n < (size_t)cand_cap) PR(d_cand.alloc((size_t)cand_cap, s));
                    PR(cudaMemsetAsync(d_eval.p + 1, 0, 24, s));
                    if (e != cudaSuccess) { rc = cuda_fail(e, "fc_prune pair list", __FILE__, __LINE__); break; }
                    PruneArgs a{};
                    a.xc = d_xc.p; a.xcf = d_xcf.p; a.g = d_g.p; a.moi = d_moi.p; a.energies = energies ? d_energy.p : nullptr;
                    a.active = d_active.p; a.tiles = d_tiles.p; a.nh = n_sel; a.mode = mode;
                    a.max_rmsd = max_rmsd; a.max_dev = max_dev; a.max_dE = max_dE; a.moi_dev = moi_dev; a.eps = FC_NEAR_EPS;
                    a.pairs = d_pairs.p; a.n_pairs = d_eval.p + 1; a.pair_cap = pair_cap;
                    a.cand = d_cand.p; a.n_cand = d_eval.p + 2; a.cand_cap = cand_cap;
                    a.swap_xy = keep_first ? 1 : 0;
                    a.n_eval = !counted ? d_eval.p : nullptr;
                    a.ties = (ties_out && !counted) ? d_ties.p : nullptr; a.n_ties = d_nties.p; a.tie_cap = cap;
                    unsigned long long found = 0;
                    if (two_stage) {
                        // screen of every pair of the tiles -> candidates; FP64 exact evaluation of the candidates
                        if (use_tc) {
                            GramArgs ga{};
                            ga.img = d_img.p; ga.gp = d_gp.p; ga.spos = d_spos.p;
                            ga.energies = energies ? d_energy.p : nullptr; ga.max_dE = max_dE;
                            ga.work = d_work.p; ga.n_work = (int)work.size(); ga.kc = kc; ga.tf32 = gram_tf32;
                            const float lim = (float)max_rmsd + kScreenBand;
                            ga.thr_e = lim * lim * (float)n_sel;
                            ga.e0_scale = 1.0f - 1.7320508f * kGramTf32Eps;
                            ga.cand = d_cand.p; ga.n_cand = d_eval.p + 2; ga.cand_cap = cand_cap;
                            ga.dump = nullptr; ga.dump_ld = 0; ga.error = d_gram_err.p;
                            ga.tile_norm = gram_cull ? d_tile_norm.p : nullptr;
                            ga.cull_gap2 = ga.thr_e * 1.002f;
                            ga.tiles_done = d_eval.p + 3;
                            const size_t smem = gram_smem_bytes(kc);
                            PR(cudaFuncSetAttribute(gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                            const unsigned grid = (unsigned)std::min<size_t>((size_t)sm_count(), work.size());
                            cudaEventRecord(ev_g0, s);
                            gram_tc_kernel<<<grid, kGramThreads, smem, s>>>(ga);
                            cudaEventRecord(ev_g1, s);
                        } else {
                            const size_t smem = (size_t)2 * PS_ATOMS * PS_LD * sizeof(float4);
                            PR(cudaFuncSetAttribute(prune_screen_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                            prune_screen_f32_kernel<<<(unsigned)tiles.size(), 128, smem, s>>>(a);
                        }
                        PR(cudaGetLastError());
                        // the exact stage follows at once on the stream (it reads the candidate count on the device); one
                        // wait per pass brings back both counters and the barrier code of the screen
                        if (e == cudaSuccess) {
                            prune_exact_kernel<<<(unsigned)(sm_count() * 8), 256, 0, s>>>(a);
                            e = cudaGetLastError();
                        }
                        PR(cudaMemcpyAsync(&hb->n_cand, d_eval.p + 2, 8, cudaMemcpyDeviceToHost, s));
                        PR(cudaMemcpyAsync(&hb->found, d_eval.p + 1, 8, cudaMemcpyDeviceToHost, s));
                        PR(cudaMemcpyAsync(&hb->tiles_done, d_eval.p + 3, 8, cudaMemcpyDeviceToHost, s));
                        if (use_tc) PR(cudaMemcpyAsync(&hb->gram_err, d_gram_err.p, 4, cudaMemcpyDeviceToHost, s));
                        else hb->gram_err = 0;
                        PR(cudaStreamSynchronize(s));
                        const unsigned long long n_cand = hb->n_cand;
                        found = hb->found;
                        if (e == cudaSuccess && hb->gram_err) {  // a bounded mbarrier wait of the tensor-core screen expired
                            set_error("gram_tc_kernel: mbarrier wait %d timed out (pass k=%lld)", hb->gram_err, (long long)k);
                            rc = FC_ERR_CUDA;
                            break;
                        }
                        if (e != cudaSuccess) { rc = cuda_fail(e, "fc_prune screen", __FILE__, __LINE__); break; }
                        if (use_tc) {
                            float ms = 0;
                            cudaEventElapsedTime(&ms, ev_g0, ev_g1);
                            t_gram += ms;
                            long long tiles_in_pass = 0;
                            for (const GramWork& wk : work) tiles_in_pass += wk.n_col_tiles;
                            if ((long long)n_cand <= cand_cap) {
                                cand_total += n_cand;
                                screen_slots += 2048.0 * (double)hb->tiles_done;   // tiles actually multiplied
                                tiles_planned += tiles_in_pass;
                                ++screen_launches;
                            }
                            if (trace) fprintf(stderr, "  pass k=%lld active=%lld items=%zu col-tiles=%lld, %llu multiplied (%.3e pair slots) screen %.3f ms, %llu candidates\n",
                                    (long long)k, (long long)n_active, work.size(), tiles_in_pass, hb->tiles_done,
                                    2048.0 * (double)hb->tiles_done, ms, n_cand);
                        }
                        if (!dev_gather && (long long)n_cand > cand_cap) {  // list too small (the exact stage did nothing): repeat with the exact size
                            cand_cap = (long long)n_cand;
                            continue;
                        }
                        counted = (long long)n_cand <= cand_cap;  // (device all-gather: the ranks decide together below)
                    } else {
                        prune_pairs_kernel<<<(unsigned)tiles.size(), 256, 0, s>>>(a);
                        PR(cudaGetLastError());
                        counted = true;
                        PR(cudaMemcpyAsync(&hb->found, d_eval.p + 1, 8, cudaMemcpyDeviceToHost, s));
                        PR(cudaStreamSynchronize(s));
                        found = hb->found;
                        if (e != cudaSuccess) { rc = cuda_fail(e, "fc_prune pass", __FILE__, __LINE__); break; }
                    }
                    if (dev_gather) {
                        // every rank learns every rank's counters, so all of them take the same branch below
                        if (gather_dev(d_eval.p + 1, d_counts.p, 16, (void*)s, gather_ctx) != 0) {
                            set_error("fc_prune: device all-gather of the counters failed");
                            rc = FC_ERR_INVALID;
                            break;
                        }
                        PR(cudaMemcpyAsync(h_counts.data(), d_counts.p, (size_t)16 * world, cudaMemcpyDeviceToHost, s));
                        PR(cudaStreamSynchronize(s));
                        if (e != cudaSuccess) { rc = cuda_fail(e, "fc_prune counters", __FILE__, __LINE__); break; }
                        unsigned long long max_found = 0, max_cand = 0, total = 0;
                        for (int r = 0; r < world; ++r) {
                            max_found = std::max(max_found, h_counts[(size_t)2 * r]);
                            max_cand = std::max(max_cand, h_counts[(size_t)2 * r + 1]);
                            total += h_counts[(size_t)2 * r];
                        }
                        if (two_stage && (long long)max_cand > cand_cap) {  // some rank's candidate list overflowed: all repeat
                            cand_cap = (long long)max_cand;  // (`counted` stays: a rank whose own list fitted has recorded its ties)
                            continue;
                        }
                        if ((long long)max_found > pair_cap) {
                            pair_cap = (long long)max_found;
                            continue;
                        }
                        pairs.resize((size_t)total);
                        if (total) {
                            if (d_pairs_all.n < (size_t)world * max_found) PR(d_pairs_all.alloc((size_t)world * max_found, s));
                            if (e == cudaSuccess && gather_dev(d_pairs.p, d_pairs_all.p, (int64_t)(max_found * sizeof(int2)), (void*)s, gather_ctx) != 0) {
                                set_error("fc_prune: device all-gather of the similar pairs failed");
                                rc = FC_ERR_INVALID;
                                break;
                            }
                            // rank-order concatenation of the filled parts, then (greedy passes) ordered by head
                            if (d_pairs_sorted.n < (size_t)total) PR(d_pairs_sorted.alloc((size_t)total + (size_t)total / 4, s));
                            size_t at = 0;
                            for (int r = 0; r < world; ++r) {
                                const size_t cnt = (size_t)h_counts[(size_t)2 * r];
                                if (cnt) PR(cudaMemcpyAsync(d_pairs_sorted.p + at, d_pairs_all.p + (size_t)r * max_found, cnt * sizeof(int2),
                                                            cudaMemcpyDeviceToDevice, s));
                                at += cnt;
                            }
                            const int2* src = d_pairs_sorted.p;
                            if (sort_pairs && total > 1) {
                                int head_bits = 1;
                                while (((int64_t)1 << head_bits) < n) ++head_bits;
                                if (d_pairs.n < (size_t)total) PR(d_pairs.alloc((size_t)total + (size_t)total / 4, s));
                                size_t need = 0;
                                PR(cub::DeviceRadixSort::SortKeys(nullptr, need, (const unsigned long long*)d_pairs_sorted.p,
                                                                  (unsigned long long*)d_pairs.p, (int)total, 32, 32 + head_bits, s));
                                if (d_sort_tmp.n < need) PR(d_sort_tmp.alloc(need + need / 4 + 256, s));
                                need = d_sort_tmp.n;
                                PR(cub::DeviceRadixSort::SortKeys(d_sort_tmp.p, need, (const unsigned long long*)d_pairs_sorted.p,
                                                                  (unsigned long long*)d_pairs.p, (int)total, 32, 32 + head_bits, s));
                                src = d_pairs.p;
                            }
                            PR(cudaMemcpyAsync(pairs.data(), src, (size_t)total * sizeof(int2), cudaMemcpyDeviceToHost, s));
                            PR(cudaStreamSynchronize(s));
                            if (e != cudaSuccess) rc = cuda_fail(e, "fc_prune pair exchange", __FILE__, __LINE__);
                        }
                        break;
                    }
                    if ((long long)found > pair_cap) {  // list too small: repeat the pass with the exact size
                        pair_cap = (long long)found;
                        continue;
                    }
                    pairs.resize((size_t)found);
                    if (found) {
                        const int2* src = d_pairs.p;
                        if (sort_pairs && found > 1) {
                            // order the list by head on the device (radix sort on the bits of the head only): the host
                            // then resolves the pass in one linear scan instead of bucketing the pairs itself
                            int head_bits = 1;
                            while (((int64_t)1 << head_bits) < n) ++head_bits;
                            if (d_pairs_sorted.n < (size_t)pair_cap) PR(d_pairs_sorted.alloc((size_t)pair_cap, s));
                            size_t need = 0;
                            PR(cub::DeviceRadixSort::SortKeys(nullptr, need, (const unsigned long long*)d_pairs.p,
                                                              (unsigned long long*)d_pairs_sorted.p, (int)found, 32, 32 + head_bits, s));
                            if (d_sort_tmp.n < need) PR(d_sort_tmp.alloc(need + need / 4 + 256, s));
                            need = d_sort_tmp.n;
                            PR(cub::DeviceRadixSort::SortKeys(d_sort_tmp.p, need, (const unsigned long long*)d_pairs.p,
                                                              (unsigned long long*)d_pairs_sorted.p, (int)found, 32, 32 + head_bits, s));
                            src = d_pairs_sorted.p;
                        }
                        PR(cudaMemcpyAsync(pairs.data(), src, (size_t)found * sizeof(int2), cudaMemcpyDeviceToHost, s));
                        PR(cudaStreamSynchronize(s));
                        if (e != cudaSuccess) rc = cuda_fail(e, "fc_prune pair readback", __FILE__, __LINE__);
                    }
                    break;
                }
                if (rc) break;
            }
            t_kernels += now() - tp;
            tp = now();
            const std::vector<int2>* use = &pairs;
            if (world > 1 && !dev_gather) {
                const void* recv = nullptr;
                int64_t recv_bytes = 0;
                int grc = gather(pairs.data(), (int64_t)(pairs.size() * sizeof(int2)), &recv, &recv_bytes, gather_ctx);
                if (grc != 0 || recv_bytes < 0 || recv_bytes % (int64_t)sizeof(int2)) {
                    set_error("fc_prune: all-gather callback failed (%d)", grc);
                    rc = FC_ERR_INVALID;
                    break;
                }
                all_pairs.resize((size_t)(recv_bytes / (int64_t)sizeof(int2)));
                if (recv_bytes) memcpy(all_pairs.data(), recv, (size_t)recv_bytes);
                use = &all_pairs;
            }
            similar_total += (int64_t)use->size();
            prune_resolve(mask, *use, keep_first != 0, snapshot != 0, sort_pairs, resolve_scratch);
            t_resolve += now() - tp;
        }
        if (!rc) {
            int n_t = 0;
            e = cudaMemcpy(&n_t, d_nties.p, 4, cudaMemcpyDeviceToHost);
            PR(cudaMemcpy(&evals_total, d_eval.p, 8, cudaMemcpyDeviceToHost));
            ties_total = n_t;
            int n_copy = (int)std::min<int64_t>(std::min<int64_t>(n_t, cap), tie_cap);
            if (n_copy > 0 && ties_out && e == cudaSuccess) {
                std::vector<TieRecord> tmp(n_copy);
                e = cudaMemcpy(tmp.data(), d_ties.p, (size_t)n_copy * sizeof(TieRecord), cudaMemcpyDeviceToHost);
                for (int i = 0; i < n_copy; ++i) {
                    ties_out[i].a = tmp[i].a; ties_out[i].b = tmp[i].b; ties_out[i].value = tmp[i].value;
                    ties_out[i].kind = tmp[i].kind; ties_out[i].decision = tmp[i].decision;
                }
            }
            if (e != cudaSuccess) rc = cuda_fail(e, "fc_prune readback", __FILE__, __LINE__);
        }
#undef PR
    }
    cudaStreamSynchronize(s);
    cudaStreamDestroy(s);
    if (rc) {
        cudaEventDestroy(ev_g0);
        cudaEventDestroy(ev_g1);
        return rc;
    }
    if (trace)
        fprintf(stderr, "fc_prune: total %.1f ms: upload+centre %.1f, tile lists %.1f, kernels+readback %.1f (tensor-core screen %.1f, "
                "%.0f of %lld planned tiles multiplied, %llu candidates), resolve %.1f\n",
                now() - t_begin, t_upload, t_tiles, t_kernels, t_gram, screen_slots / 2048.0, tiles_planned, cand_total, t_resolve);
    cudaEventDestroy(ev_g0);
    cudaEventDestroy(ev_g1);
    g_prune_timing[0] = now() - t_begin;
    g_prune_timing[1] = t_gram;
    g_prune_timing[2] = (double)screen_launches;
    g_prune_timing[3] = screen_slots;
    g_prune_timing[4] = (double)cand_total;
    g_prune_timing[5] = (double)n_sel;
    g_prune_tiles[0] = (double)tiles_planned;
    g_prune_tiles[1] = screen_slots / 2048.0;
    memcpy(mask_out, mask.data(), (size_t)n);
    if (n_ties_out) *n_ties_out = ties_total;
    if (stats_out) {
        stats_out[0] = passes;
        stats_out[1] = pairs_tiled;               // pairs whose covariance was accumulated (on this rank)
        stats_out[2] = (int64_t)evals_total;      // pairs that needed the eigen-solve (on this rank)
        stats_out[3] = pairs_skipped;             // pairs known dissimilar from an earlier pass, not re-evaluated
    }
    return FC_OK;
}

extern "C" int fc_prune_sharded(const double* structures, int64_t n, int32_t n_atoms, int32_t mode, const int32_t* sel,
                                int32_t n_sel, const double* masses, double max_rmsd, double max_dev, double moi_dev,
                                const double* energies, double max_dE, int32_t keep_first, int32_t snapshot,
                                int32_t min_per_chunk, uint8_t* mask_out, int64_t* stats_out, fc_tie* ties_out,
                                int64_t tie_cap, int64_t* n_ties_out, int32_t rank, int32_t world,
                                fc_allgather_fn gather, void* gather_ctx) {
    return prune_impl(structures, n, n_atoms, mode, sel, n_sel, masses, max_rmsd, max_dev, moi_dev, energies, max_dE,
                      keep_first, snapshot, min_per_chunk, mask_out, stats_out, ties_out, tie_cap, n_ties_out, rank, world,
                      gather, nullptr, gather_ctx);
}

extern "C" int fc_prune_sharded_dev(const double* structures, int64_t n, int32_t n_atoms, int32_t mode, const int32_t* sel,
                                    int32_t n_sel, const double* masses, double max_rmsd, double max_dev, double moi_dev,
                                    const double* energies, double max_dE, int32_t keep_first, int32_t snapshot,
                                    int32_t min_per_chunk, uint8_t* mask_out, int64_t* stats_out, fc_tie* ties_out,
                                    int64_t tie_cap, int64_t* n_ties_out, int32_t rank, int32_t world,
                                    fc_allgather_dev_fn gather_dev, void* gather_ctx) {
    FC_REQUIRE(world == 1 || gather_dev, "fc_prune_sharded_dev: a multi-rank call needs the device all-gather callback");
    return prune_impl(structures, n, n_atoms, mode, sel, n_sel, masses, max_rmsd, max_dev, moi_dev, energies, max_dE,
                      keep_first, snapshot, min_per_chunk, mask_out, stats_out, ties_out, tie_cap, n_ties_out, rank, world,
                      nullptr, gather_dev, gather_ctx);
}

extern "C" int fc_prune(const double* structures, int64_t n, int32_t n_atoms, int32_t mode, const int32_t* sel,
                        int32_t n_sel, const double* masses, double max_rmsd, double max_dev, double moi_dev,
                        const double* energies, double max_dE, int32_t keep_first, int32_t snapshot,
                        int32_t min_per_chunk, uint8_t* mask_out, int64_t* stats_out, fc_tie* ties_out,
                        int64_t tie_cap, int64_t* n_ties_out) {
    return fc_prune_sharded(structures, n, n_atoms, mode, sel, n_sel, masses, max_rmsd, max_dev, moi_dev, energies,
                            max_dE, keep_first, snapshot, min_per_chunk, mask_out, stats_out, ties_out, tie_cap,
                            n_ties_out, 0, 1, nullptr, nullptr);
}
