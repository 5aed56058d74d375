// firecode_b200 -- candidate-pose generation and the in-loop similarity filters of the embeds.
//
// String embed  (reference: firecode/embeds.py:51-158):
//   tuple (c1, c2, ai1, ai2, angle) -> rigid transform of molecule 2 (string_xf_kernel, FP64)
//   -> clash screen (fc_clash.cu) -> order-preserving compaction of the survivors
//   -> torsion fingerprints of the survivors (tfd_fingerprint_kernel; torsion_module.py:1070-1076)
//   -> keep-first sweep against ALL earlier accepted poses (embeds.py:59-84: the "LRU" cache never
//      evicts) = the lexicographically first maximal independent set of the
//      similarity graph: all pairs of a block of candidates into a bit matrix (FP32 pre-screen with
//      a margin, FP64 decision), then rounds of final verdicts over the whole block in one
//      cooperative launch -- the accepted set and its order are the reference's
//   -> materialisation of the kept poses only (get_embed, embeds.py:808-817).
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

#include <string.h>

#include <algorithm>
#include <chrono>
#include <vector>

#include <cooperative_groups.h>

#include "fc_embed.cuh"

namespace fc {

// ---------------------------------------------------------------------------------------------
// compaction
// ---------------------------------------------------------------------------------------------
struct PassPred {
    const uint8_t* status;
    long long base;
    __host__ __device__ bool operator()(long long i) const { return status[i - base] & FC_STATUS_PASS; }
};

int compact_pass(const uint8_t* status, long long n, long long base, long long* out_idx, int* out_count,
                 cudaStream_t s) {
    thrust::counting_iterator<long long> it(base);
    PassPred pred{status, base};
    size_t tmp_bytes = 0;
    FC_CUDA(cub::DeviceSelect::If(nullptr, tmp_bytes, it, out_idx, out_count, (int)n, pred, s));
    void* tmp = nullptr;
    FC_CUDA(cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 16, s));
    cudaError_t e = cub::DeviceSelect::If(tmp, tmp_bytes, it, out_idx, out_count, (int)n, pred, s);
    cudaFreeAsync(tmp, s);
    FC_CUDA(e);
    return FC_OK;
}

// ---------------------------------------------------------------------------------------------
// string embed: tuple -> transform
// ---------------------------------------------------------------------------------------------
struct StringDev {
    const double *coords1, *coords2;    // (C1,N1,3), (C2,N2,3)
    const double *cen1, *vec1;          // (C1,K1,3)
    const double *cen2, *vec2;          // (C2,K2,3)
    const double* angles;               // (A)
    int c1, n1, c2, n2, k1, k2, na;
    int handed;
};

__device__ __forceinline__ void string_decode(const StringDev& p, long long pose, int& c1, int& c2, int& a1,
                                              int& a2, int& ai) {
    // cartesian_product order: first index fastest for conformers and centres
    long long per_conf = (long long)p.k1 * p.k2 * p.na;
    long long ci = pose / per_conf;
    int rest = (int)(pose - ci * per_conf);
    int ki = rest / p.na;
    ai = rest - ki * p.na;
    c1 = (int)(ci % p.c1);
    c2 = (int)(ci / p.c1);
    a1 = ki % p.k1;
    a2 = ki / p.k1;
}

// embeds.py:121-135
__device__ __forceinline__ void string_transform(const StringDev& p, int c1, int c2, int a1, int a2, int ai,
                                                 M3& rot, double* t) {
    const double* p1 = p.cen1 + ((size_t)c1 * p.k1 + a1) * 3;
    const double* p2 = p.cen2 + ((size_t)c2 * p.k2 + a2) * 3;
    const double* ref_vec = p.vec1 + ((size_t)c1 * p.k1 + a1) * 3;
    const double* mol_vec = p.vec2 + ((size_t)c2 * p.k2 + a2) * 3;
    double neg_ref[3] = {-ref_vec[0], -ref_vec[1], -ref_vec[2]};
    rot = rot_vec_to_vec(mol_vec, neg_ref, p.handed);
    double angle = p.angles[ai];
    if (angle != 0.0) {
        M3 delta = rot_from_pointer(ref_vec, angle, p.handed);
        rot = m3_mul(delta, rot);
    }
    double rp[3];
    m3_apply(rot, p2, rp);
    t[0] = p1[0] - rp[0];
    t[1] = p1[1] - rp[1];
    t[2] = p1[2] - rp[2];
}

__global__ void string_xf_kernel(StringDev p, long long lo, long long n, double* __restrict__ xf) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c1, c2, a1, a2, ai;
    string_decode(p, lo + i, c1, c2, a1, a2, ai);
    M3 rot;
    double t[3];
    string_transform(p, c1, c2, a1, a2, ai, rot, t);
    double* o = xf + i * 12;
#pragma unroll
    for (int k = 0; k < 9; ++k) o[k] = rot.m[k];
    o[9] = t[0];
    o[10] = t[1];
    o[11] = t[2];
}

// position of atom `atom` (cumulative numbering) of the pose with transform xf
__device__ __forceinline__ void string_atom(const StringDev& p, int c1, int c2, const double* xf, int atom,
                                            double* out) {
    if (atom < p.n1) {
        const double* a = p.coords1 + ((size_t)c1 * p.n1 + atom) * 3;
        out[0] = a[0]; out[1] = a[1]; out[2] = a[2];
    } else {
        const double* b = p.coords2 + ((size_t)c2 * p.n2 + (atom - p.n1)) * 3;
        // (R @ X.T).T + t  (embeds.py:815-817)
        out[0] = (xf[0] * b[0] + xf[1] * b[1] + xf[2] * b[2]) + xf[9];
        out[1] = (xf[3] * b[0] + xf[4] * b[1] + xf[5] * b[2]) + xf[10];
        out[2] = (xf[6] * b[0] + xf[7] * b[1] + xf[8] * b[2]) + xf[11];
    }
}

// torsion fingerprints of the survivors: one thread per (survivor, quadruplet)
__global__ void tfd_fingerprint_kernel(StringDev p, const long long* __restrict__ surv, int n_surv, long long lo,
                                       const double* __restrict__ xf, const long long* __restrict__ quads, int nq,
                                       double* __restrict__ fp) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n_surv * nq) return;
    int s = (int)(i / nq), q = (int)(i - (long long)s * nq);
    long long pose = surv[s];
    int c1, c2, a1, a2, ai;
    string_decode(p, pose, c1, c2, a1, a2, ai);
    const double* x = xf + (pose - lo) * 12;
    double pt[4][3];
#pragma unroll
    for (int k = 0; k < 4; ++k) string_atom(p, c1, c2, x, (int)quads[4 * q + k], pt[k]);
    fp[i] = dihedral_deg(pt[0], pt[1], pt[2], pt[3]);
}

// out[i] = table[idx[i]]
__global__ void gather_index_kernel(const long long* __restrict__ table, const int* __restrict__ idx, int n,
                                    long long* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = table[idx[i]];
}

__global__ void string_materialize_kernel(StringDev p, const long long* __restrict__ kept, int n_kept,
                                          double* __restrict__ out) {
    int k = blockIdx.x;
    if (k >= n_kept) return;
    long long pose = kept[k];
    int c1, c2, a1, a2, ai;
    string_decode(p, pose, c1, c2, a1, a2, ai);
    __shared__ double xf[12];
    if (threadIdx.x == 0) {
        M3 rot;
        double t[3];
        string_transform(p, c1, c2, a1, a2, ai, rot, t);
        for (int j = 0; j < 9; ++j) xf[j] = rot.m[j];
        xf[9] = t[0]; xf[10] = t[1]; xf[11] = t[2];
    }
    __syncthreads();
    int n_tot = p.n1 + p.n2;
    double* o = out + (size_t)k * n_tot * 3;
    for (int a = threadIdx.x; a < n_tot; a += blockDim.x) {
        double v[3];
        string_atom(p, c1, c2, xf, a, v);
        o[3 * a] = v[0];
        o[3 * a + 1] = v[1];
        o[3 * a + 2] = v[2];
    }
}

// ---------------------------------------------------------------------------------------------
// TFD keep-first sweep (torsion_module.py:1056-1067 + embeds.py:59-84)
// ---------------------------------------------------------------------------------------------
struct TfdArgs {
    const double* fp;        // (n, q)
    const long long* label;  // (n) pose index of each row (for the tie list) or null
    int q;
    double thr, eps;
    int* flag;   // (n) 1 = rejected
    int* acc;    // accepted rows, in order
    int* n_acc;
    TieRecord* ties;
    int* n_ties;
    int tie_cap;
};

// similar <=> sum_k wrap(|fa_k - fb_k|) < thr ; near-threshold sums are listed
__device__ __forceinline__ bool tfd_similar(const TfdArgs& a, int row_new, int row_ref) {
    const double* x = a.fp + (size_t)row_new * a.q;
    const double* y = a.fp + (size_t)row_ref * a.q;
    double sum = 0.0;
    const double stop = a.thr + a.eps;
    for (int k = 0; k < a.q; ++k) {
        double d = fabs(x[k] - y[k]);
        d = fabs(d - (d > 180.0 ? 360.0 : 0.0));
        sum += d;
        if (sum > stop) return false;  // partial sums only grow: cannot come back under thr + eps
    }
    bool sim = sum < a.thr;
    if (fabs(sum - a.thr) <= a.eps && a.ties) {
        int slot = atomicAdd(a.n_ties, 1);
        if (slot < a.tie_cap) {
            TieRecord r;
            r.a = a.label ? a.label[row_new] : row_new;
            r.b = a.label ? a.label[row_ref] : row_ref;
            r.value = sum;
            r.kind = FC_TIE_TFD;
            r.decision = sim ? 1 : 0;
            a.ties[slot] = r;
        }
    }
    return sim;
}

// candidates [b0, b1) against every row accepted before this block
__global__ void __launch_bounds__(256) tfd_vs_accepted_kernel(TfdArgs a, int b0, int b1) {
    const int n_acc = *a.n_acc;
    const long long total = (long long)(b1 - b0) * n_acc;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int cand = b0 + (int)(i / n_acc);
        int ref = a.acc[(int)(i % n_acc)];
        if (tfd_similar(a, cand, ref)) a.flag[cand] = 1;
    }
}

// ---- sweep inside a block of candidates ------------------------------------------------------------
// The keep-first rule is the lexicographically first maximal independent set of the similarity graph: row t is accepted
// iff no EARLIER accepted row is similar to it.  Instead of walking the rows one by one (a chain of n dependent steps),
// all pairs of a block are evaluated at once into a bit matrix (tfd_block_matrix_kernel: bit i of word i / 32 of row t
// says "t is similar to the earlier row i") and the rule is resolved by rounds over the whole block
// (tfd_block_resolve_kernel, one cooperative launch): an undecided row is REJECTED as soon as a similar earlier row is
// accepted and ACCEPTED once every similar earlier row is rejected; both verdicts are final, the lowest undecided row
// can always be decided, and the number of rounds is the longest chain of dependent decisions (a handful in practice),
// not the number of rows.  C1 (10 504 clash survivors): one block, 2 launches instead of 44.
constexpr int kTfdBlock = 16384;            // candidates per sweep step (bit matrix of a full block: 32 MB)
constexpr int kTfdWords = kTfdBlock / 32;
constexpr int kTfdRowsPerCta = 256;
constexpr int kTfdSegWords = 4;             // words of the bit matrix one CTA of the pair kernel walks
constexpr int kTfdHead = 8;                 // terms of phase A of the pair kernel (kept in registers)
constexpr int kTfdMaxOrderedQ = 1024;       // torsions per fingerprint up to which the pre-screen re-orders the terms

// similar <=> sum_k wrap(|x_k - y_k|) < thr, on explicit fingerprint rows (shared or global memory)
__device__ __forceinline__ bool tfd_similar_rows(const TfdArgs& a, const double* __restrict__ x, const double* __restrict__ y,
                                                 int row_new, int row_ref) {
    double sum = 0.0;
    const double stop = a.thr + a.eps;
    // terms are ADDED in the reference's order; they are LOADED eight at a time so that the loads of a chunk are in
    // flight together (one global round trip per chunk instead of one per term).  Leaving at a chunk boundary instead
    // of at the first term over the limit changes nothing: partial sums only grow.
    int k = 0;
    for (; k + 8 <= a.q; k += 8) {
        double xv[8], yv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { xv[u] = x[k + u]; yv[u] = y[k + u]; }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            double d = fabs(xv[u] - yv[u]);
            d = fabs(d - (d > 180.0 ? 360.0 : 0.0));
            sum += d;
        }
        if (sum > stop) return false;
    }
    for (; k < a.q; ++k) {
        double d = fabs(x[k] - y[k]);
        d = fabs(d - (d > 180.0 ? 360.0 : 0.0));
        sum += d;
    }
    if (sum > stop) return false;
    const bool sim = sum < a.thr;
    if (fabs(sum - a.thr) <= a.eps && a.ties) {
        int slot = atomicAdd(a.n_ties, 1);
        if (slot < a.tie_cap) {
            TieRecord r;
            r.a = a.label ? a.label[row_new] : row_new;
            r.b = a.label ? a.label[row_ref] : row_ref;
            r.value = sum;
            r.kind = FC_TIE_TFD;
            r.decision = sim ? 1 : 0;
            a.ties[slot] = r;
        }
    }
    return sim;
}

// Order in which the FP32 pre-screen adds the torsion terms: most scattered torsion first (circular variance over a
// sample of the rows), so that dissimilar pairs leave the loop after a few terms.  Speed only: the pre-screen has a
// margin and every pair it cannot rule out is decided by the FP64 sum in the reference's term order.
__global__ void __launch_bounds__(1024) tfd_order_kernel(const double* __restrict__ fp, int n, int q, int* __restrict__ perm) {
    __shared__ float s_score[kTfdMaxOrderedQ];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_sample = min(n, 1024), stride = max(1, n / n_sample);
    for (int k = warp; k < q; k += 32) {
        float c = 0.f, sn = 0.f;
        for (int r = lane; r < n_sample; r += 32) {
            float sv, cv;
            __sincosf((float)fp[(size_t)r * stride * q + k] * 0.017453292f, &sv, &cv);
            c += cv;
            sn += sv;
        }
        for (int o = 16; o > 0; o >>= 1) {
            c += __shfl_xor_sync(0xffffffffu, c, o);
            sn += __shfl_xor_sync(0xffffffffu, sn, o);
        }
        if (lane == 0) s_score[k] = c * c + sn * sn;  // small = scattered
    }
    __syncthreads();
    // rank of every term (ties by index): a stable sort by ascending score
    for (int k = threadIdx.x; k < q; k += blockDim.x) {
        int rank = 0;
        const float mine = s_score[k];
        for (int j = 0; j < q; ++j) rank += (s_score[j] < mine || (s_score[j] == mine && j < k)) ? 1 : 0;
        perm[rank] = k;
    }
}

// Group culling.  Rows come in the reference's enumeration order, so 32 consecutive rows mostly share their conformers
// and with them most torsions: per group of 32 rows (one word of the bit matrix) and torsion, the interval of the values
// (tfd_group_box_kernel); two groups whose intervals are apart by more than the threshold in the wrapped L1 norm
// cannot hold a similar pair, and the pair kernel skips the whole 32 x 32 word (tfd_group_sep_kernel: one bit per
// (row group, word)).  A lower bound only: groups that straddle +-180 degrees have wide intervals and are simply not culled.
__global__ void __launch_bounds__(128) tfd_group_box_kernel(const double* __restrict__ fp, int b0, int nb, int q, int W,
                                                            float* __restrict__ gmin, float* __restrict__ gmax) {
    const int g = blockIdx.x;
    const int r_end = min(32, nb - 32 * g);
    for (int k = threadIdx.x; k < q; k += blockDim.x) {
        float lo = 1e30f, hi = -1e30f;
        for (int r = 0; r < r_end; ++r) {
            const float v = (float)fp[(size_t)(b0 + 32 * g + r) * q + k];
            lo = fminf(lo, v);
            hi = fmaxf(hi, v);
        }
        gmin[(size_t)k * W + g] = lo;
        gmax[(size_t)k * W + g] = hi;
    }
}

// sep[rg * Wd + (w >> 5)] bit (w & 31) = 1: no row of group rg can be similar to a row of group w (w <= rg)
__global__ void __launch_bounds__(128) tfd_group_sep_kernel(const float* __restrict__ gmin, const float* __restrict__ gmax, int W,
                                                            int Wd, int q, float limit, unsigned* __restrict__ sep) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x, rg = blockIdx.y;
    bool apart = true;
    if (w < W && w <= rg) {
        float sum = 0.f;
        for (int k = 0; k < q && sum <= limit; ++k) {
            const float a0 = gmin[(size_t)k * W + rg], a1 = gmax[(size_t)k * W + rg];
            const float c0 = gmin[(size_t)k * W + w], c1 = gmax[(size_t)k * W + w];
            const float gap = fmaxf(0.f, fmaxf(c0 - a1, a0 - c1));       // linear gap of the intervals
            const float far = fmaxf(a1 - c0, c1 - a0);                   // largest linear difference of two members
            sum += fmaxf(0.f, fminf(gap, 360.0f - far));                 // wrapped distance of any two members is at least this
        }
        apart = sum > limit;
    }
    const unsigned m = __ballot_sync(0xffffffffu, apart);
    if ((threadIdx.x & 31) == 0 && w < 32 * Wd) sep[(size_t)rg * Wd + (w >> 5)] = m;
}

// CTA = 256 consecutive rows t of the block x a segment of kTfdSegWords words of the bit matrix (32 earlier rows each);
// thread = one row.  The rows' fingerprints are staged once per CTA as FP32 in the order of `perm`, the 32 earlier rows
// of a word per step; the lanes of a warp compare their own rows with the SAME earlier row at a time (broadcast).
// FP32 pre-screen: the wrapped L1 distance of FP32-rounded fingerprints differs from the FP64 one by less than
// `margin` (2e-4 per term, 0.02 floor: 50x the rounding analysis), so a pair whose FP32 sum exceeds thr + margin is
// dissimilar and not near the threshold; everything else is decided (and, near the threshold, listed) in FP64.
// Rows already rejected against rows accepted in earlier blocks take no part (their words are 0).
__global__ void __launch_bounds__(kTfdRowsPerCta) tfd_block_matrix_kernel(TfdArgs a, int b0, int nb, int W, const int* __restrict__ perm,
                                                                          const unsigned* __restrict__ sep, int Wd,
                                                                          unsigned* __restrict__ simT) {
    extern __shared__ float s_fp32[];
    const int n_seg = (W + kTfdSegWords - 1) / kTfdSegWords;
    const int seg = blockIdx.x % n_seg, rt = blockIdx.x / n_seg;
    const int rows_per_cta = (int)blockDim.x;  // 256, fewer when the fingerprints are long
    const int row0 = rt * rows_per_cta;
    const int t = row0 + threadIdx.x;
    const int t_hi = min(nb, row0 + rows_per_cta);  // rows of this CTA: [row0, t_hi)
    const int w_lo = seg * kTfdSegWords, w_hi = min(W, w_lo + kTfdSegWords);
    // words of the segment that hold an earlier row of some row of the CTA: 32 w < t_hi - 1
    const int w_live = min(w_hi, (t_hi - 1 + 31) / 32);
    if (w_live <= w_lo) {
        if (t < nb)
            for (int w = w_lo; w < w_hi; ++w) simT[(size_t)t * W + w] = 0u;
        return;
    }
    // group culling bits of this CTA's row groups (one per warp) for the words of the segment
    __shared__ unsigned s_sep[kTfdRowsPerCta / 32];
    if (threadIdx.x < (int)blockDim.x / 32) {
        const int rg = row0 / 32 + threadIdx.x;
        s_sep[threadIdx.x] = rg < W ? (sep[(size_t)rg * Wd + (w_lo >> 5)] >> (w_lo & 31)) : 0xffffffffu;
    }
    __syncthreads();
    {
        unsigned all_apart = 0xffffffffu;
        for (int k = 0; k < (int)blockDim.x / 32; ++k) all_apart &= s_sep[k];
        bool any_live = false;
        for (int w = w_lo; w < w_live; ++w) any_live |= !((all_apart >> (w - w_lo)) & 1u);
        if (!any_live) {  // CTA-uniform: nothing in this segment can be similar
            if (t < nb)
                for (int w = w_lo; w < w_hi; ++w) simT[(size_t)t * W + w] = 0u;
            return;
        }
    }
    const int q = a.q, qs = q | 1;  // odd stride: conflict-free rows
    float* s_rows = s_fp32;
    float* s_cols = s_fp32 + (size_t)rows_per_cta * qs;
    __shared__ int s_cflag[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < t_hi - row0; r += rows_per_cta / 32) {
        const double* src = a.fp + (size_t)(b0 + row0 + r) * q;
        for (int k = lane; k < q; k += 32) s_rows[r * qs + k] = (float)src[perm[k]];
    }
    const bool mine = t < nb && a.flag[b0 + t] == 0;
    const float margin = 0.02f + 2e-4f * (float)q;
    const float stop = (float)a.thr + margin;
    // Two phases per word, because the cost of a warp is the cost of its slowest lane: (A) every lane adds the first
    // kTfdHead terms of its row against each earlier row -- almost every pair is over the limit by then; (B) the few pairs
    // that are not go to a per-warp queue and are finished 32 at a time, one pair per lane (the remaining FP32 terms,
    // then the FP64 sum), so the long evaluations run on full warps.  Bits are set in a shared word per row.
    __shared__ unsigned s_word[kTfdRowsPerCta];
    __shared__ unsigned short s_queue[kTfdRowsPerCta / 32][64];
    unsigned short* queue = s_queue[warp];
    const int head = min(q, kTfdHead);
    float xh[kTfdHead];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kTfdHead; ++k) xh[k] = k < head ? s_rows[(size_t)threadIdx.x * qs + k] : 0.f;
    // finishes `cnt` queued pairs of this warp (entry = row within the warp << 5 | column within the word)
    auto flush = [&](int cnt, int w) {
        __syncwarp();
        if (lane < cnt) {
            const int ent = queue[lane], rl = ent >> 5, j = ent & 31;
            const float* xx = s_rows + (size_t)(warp * 32 + rl) * qs;
            const float* yy = s_cols + (size_t)j * qs;
            float sum = 0.f;
            for (int k = 0; k < q; ++k) {
                const float d = fabsf(xx[k] - yy[k]);
                sum += fminf(d, 360.0f - d);
            }
            if (sum <= stop) {
                const int tt = row0 + warp * 32 + rl, i = 32 * w + j;
                if (tfd_similar_rows(a, a.fp + (size_t)(b0 + tt) * q, a.fp + (size_t)(b0 + i) * q, b0 + tt, b0 + i))
                    atomicOr(&s_word[warp * 32 + rl], 1u << j);
            }
        }
        __syncwarp();
    };
    for (int w = w_lo; w < w_hi; ++w) {
        bool word_live = w < w_live;  // CTA-uniform
        if (word_live) {
            unsigned all_apart = 1u;
            for (int k = 0; k < (int)blockDim.x / 32; ++k) all_apart &= s_sep[k] >> (w - w_lo);
            word_live = !(all_apart & 1u);
        }
        if (!word_live) {
            if (t < nb) simT[(size_t)t * W + w] = 0u;
            continue;
        }
        __syncthreads();  // the previous word's columns are no longer read
        for (int r = warp; r < 32; r += rows_per_cta / 32) {
            const int i = 32 * w + r;
            if (lane == 0) s_cflag[r] = i < nb ? a.flag[b0 + i] : 1;
            if (i < nb) {
                const double* src = a.fp + (size_t)(b0 + i) * q;
                for (int k = lane; k < q; k += 32) s_cols[r * qs + k] = (float)src[perm[k]];
            }
        }
        s_word[threadIdx.x] = 0u;
        __syncthreads();
        int qn = 0;  // warp-uniform fill level of the queue
        // earlier rows of this word for the LAST row of the warp (the lanes' own limits are checked per pair)
        const int j_warp = ((s_sep[warp] >> (w - w_lo)) & 1u) ? 0 : min(32, row0 + warp * 32 + 31 - 32 * w);
        for (int j = 0; j < j_warp; ++j) {
            if (s_cflag[j]) continue;  // warp-uniform
            const float* y = s_cols + (size_t)j * qs;
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < kTfdHead; ++k) {
                const float d = fabsf(xh[k] - (k < head ? y[k] : 0.f));
                sum += fminf(d, 360.0f - d);
            }
            const bool go_on = mine && 32 * w + j < t && sum <= stop;
            const unsigned m = __ballot_sync(0xffffffffu, go_on);
            if (m == 0u) continue;
            const int add = __popc(m);
            if (qn + add > 64) {
                flush(min(qn, 32), w);  // qn <= 63 here: at most two rounds
                if (qn > 32) {
                    __syncwarp();
                    if (lane < qn - 32) queue[lane] = queue[32 + lane];
                    __syncwarp();
                    flush(qn - 32, w);
                }
                qn = 0;
            }
            if (go_on) queue[qn + __popc(m & ((1u << lane) - 1u))] = (unsigned short)((lane << 5) | j);
            qn += add;
        }
        if (qn > 0) {
            flush(min(qn, 32), w);
            if (qn > 32) {
                __syncwarp();
                if (lane < qn - 32) queue[lane] = queue[32 + lane];
                __syncwarp();
                flush(qn - 32, w);
            }
        }
        __syncwarp();
        if (t < nb) simT[(size_t)t * W + w] = s_word[threadIdx.x];
    }
}

// Rounds of the rule above over one block.  Cooperative launch, one CTA per SM.  Every CTA keeps its own copy of the
// accepted / undecided bit masks in shared memory; a round = every undecided row gets a verdict from the masks (one warp
// per row scans the row's words), grid-wide barrier, every CTA applies all verdicts to its copy.  The verdict array is
// double-buffered by round parity, so one barrier per round is enough.  At the end CTA 0 appends the accepted rows, in
// order, to the accepted list and writes the rejected flags.
__global__ void __launch_bounds__(1024, 1) tfd_block_resolve_kernel(TfdArgs a, int b0, int nb, int W,
                                                                    const unsigned* __restrict__ simT,
                                                                    unsigned char* __restrict__ verdict /* [2][nb] */,
                                                                    int* __restrict__ rounds_out) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    __shared__ unsigned s_acc[kTfdWords], s_und[kTfdWords];
    __shared__ int s_left;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g_warp = blockIdx.x * (blockDim.x >> 5) + warp, n_warps = gridDim.x * (blockDim.x >> 5);
    for (int w = threadIdx.x; w < W; w += blockDim.x) {
        unsigned u = 0u;
        for (int k = 0; k < 32; ++k) {
            const int r = 32 * w + k;
            if (r < nb && a.flag[b0 + r] == 0) u |= 1u << k;
        }
        s_und[w] = u;
        s_acc[w] = 0u;
    }
    __syncthreads();
    for (int round = 0;; ++round) {
        unsigned char* v = verdict + (size_t)(round & 1) * nb;
        for (int t = g_warp; t < nb; t += n_warps) {
            if (!((s_und[t >> 5] >> (t & 31)) & 1u)) continue;  // warp-uniform
            unsigned hit_acc = 0u, hit_und = 0u;
            const unsigned* row = simT + (size_t)t * W;
            const int last = t > 0 ? (t - 1) >> 5 : -1;
            for (int w = lane; w <= last; w += 32) {
                const unsigned m = row[w];
                hit_acc |= m & s_acc[w];
                hit_und |= m & s_und[w];
            }
            const bool any_acc = __any_sync(0xffffffffu, hit_acc != 0u), any_und = __any_sync(0xffffffffu, hit_und != 0u);
            if (lane == 0) v[t] = any_acc ? 2 : (any_und ? 0 : 1);  // 2 = rejected, 1 = accepted, 0 = still undecided
        }
        grid.sync();
        if (threadIdx.x == 0) s_left = 0;
        __syncthreads();
        int left = 0;
        for (int w = threadIdx.x; w < W; w += blockDim.x) {
            unsigned u = s_und[w], acc = s_acc[w];
            unsigned todo = u;
            while (todo) {
                const int k = __ffs(todo) - 1;
                todo &= todo - 1u;
                const unsigned char d = v[32 * w + k];
                if (d) {
                    u &= ~(1u << k);
                    if (d == 1) acc |= 1u << k;
                }
            }
            s_und[w] = u;
            s_acc[w] = acc;
            left += __popc(u);
        }
        if (left) atomicAdd(&s_left, left);
        __syncthreads();
        if (s_left == 0) {  // every CTA holds the same masks: the exit is grid-uniform
            if (blockIdx.x == 0 && threadIdx.x == 0 && rounds_out) *rounds_out += round + 1;
            break;
        }
        __syncthreads();
    }
    if (blockIdx.x != 0) return;
    // accepted rows in order -> list; flags of the rejected rows
    __shared__ int s_pref[kTfdWords + 1];
    if (threadIdx.x == 0) {
        int run = 0;
        for (int w = 0; w < W; ++w) { s_pref[w] = run; run += __popc(s_acc[w]); }
        s_pref[W] = run;
    }
    __syncthreads();
    const int n_acc0 = *a.n_acc;
    for (int w = threadIdx.x; w < W; w += blockDim.x) {
        unsigned acc = s_acc[w];
        int o = n_acc0 + s_pref[w];
        for (int k = 0; k < 32; ++k) {
            const int r = 32 * w + k;
            if (r >= nb) break;
            const bool kept = (acc >> k) & 1u;
            if (kept) a.acc[o++] = b0 + r;
            a.flag[b0 + r] = kept ? 0 : 1;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *a.n_acc = n_acc0 + s_pref[W];
}

int tfd_keepfirst_dev(const double* fp, const long long* label, int n, int q, double thr, double eps, int* flag,
                      int* acc, int* n_acc, TieRecord* ties, int* n_ties, int tie_cap, cudaStream_t s) {
    if (n == 0) return FC_OK;
    FC_CUDA(cudaMemsetAsync(flag, 0, (size_t)n * sizeof(int), s));
    FC_CUDA(cudaMemsetAsync(n_acc, 0, sizeof(int), s));
    const int nb_max = std::min(n, kTfdBlock), w_max = (nb_max + 31) / 32;
    unsigned* simT = nullptr;
    unsigned char* verdict = nullptr;
    const bool trace = getenv("FC_STRING_TRACE") != nullptr;
    const size_t verdict_bytes = ((size_t)2 * nb_max + 15) / 16 * 16;
    FC_CUDA(cudaMallocAsync((void**)&simT, (size_t)nb_max * w_max * sizeof(unsigned), s));
    FC_CUDA(cudaMallocAsync((void**)&verdict, verdict_bytes + 16, s));
    int* rounds = reinterpret_cast<int*>(verdict + verdict_bytes);
    FC_CUDA(cudaMemsetAsync(rounds, 0, 16, s));
    TfdArgs a{fp, label, q, thr, eps, flag, acc, n_acc, ties, n_ties, tie_cap};
    // FP32 copies of the rows of a CTA (256, fewer for long fingerprints) and of one word's earlier rows (32), staged in
    // the pre-screen's term order
    int rows_per_cta = kTfdRowsPerCta;
    while (rows_per_cta > 32 && (size_t)(rows_per_cta + 32) * (size_t)(q | 1) * sizeof(float) > 200 * 1024) rows_per_cta /= 2;
    const size_t fp_smem = (size_t)(rows_per_cta + 32) * (size_t)(q | 1) * sizeof(float);
    FC_REQUIRE(fp_smem <= 200 * 1024 && q <= kTfdMaxOrderedQ, "fc_tfd: %d torsions per fingerprint are more than the sweep stages (800)", q);
    FC_CUDA(cudaFuncSetAttribute(tfd_block_matrix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fp_smem));
    int* perm = nullptr;
    float* gbox = nullptr;    // [2][q][w_max] interval of every torsion over every group of 32 rows
    unsigned* sep = nullptr;  // [w_max][wd_max] culling bits
    const int wd_max = (w_max + 31) / 32;
    FC_CUDA(cudaMallocAsync((void**)&perm, (size_t)std::max(q, 1) * sizeof(int), s));
    FC_CUDA(cudaMallocAsync((void**)&gbox, (size_t)2 * std::max(q, 1) * w_max * sizeof(float), s));
    FC_CUDA(cudaMallocAsync((void**)&sep, (size_t)w_max * wd_max * sizeof(unsigned), s));
    // the same margins as the FP32 pre-screen of the pair kernel, once more for the FP32 interval arithmetic
    const float sep_limit = (float)thr + 2.0f * (0.02f + 2e-4f * (float)q);
    tfd_order_kernel<<<1, 1024, 0, s>>>(fp, n, q, perm);
    const int grid = sm_count() * 8;
    cudaError_t e = cudaSuccess;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    if (trace) for (auto& x : ev) cudaEventCreate(&x);
    for (int b0 = 0; b0 < n && e == cudaSuccess; b0 += kTfdBlock) {
        int nb = std::min(n - b0, kTfdBlock), W = (nb + 31) / 32;
        if (b0 > 0) tfd_vs_accepted_kernel<<<grid, 256, 0, s>>>(a, b0, b0 + nb);
        if (trace) cudaEventRecord(ev[0], s);
        const unsigned ctas = (unsigned)(((nb + rows_per_cta - 1) / rows_per_cta) * ((W + kTfdSegWords - 1) / kTfdSegWords));
        const int Wd = (W + 31) / 32;
        float* gmin = gbox;
        float* gmax = gbox + (size_t)std::max(q, 1) * w_max;
        tfd_group_box_kernel<<<(unsigned)W, 128, 0, s>>>(fp, b0, nb, q, W, gmin, gmax);
        tfd_group_sep_kernel<<<dim3((unsigned)(32 * Wd + 127) / 128, (unsigned)W), 128, 0, s>>>(gmin, gmax, W, Wd, q, sep_limit, sep);
        tfd_block_matrix_kernel<<<ctas, rows_per_cta, fp_smem, s>>>(a, b0, nb, W, perm, sep, Wd, simT);
        e = cudaGetLastError();
        if (e != cudaSuccess) break;
        if (trace) cudaEventRecord(ev[1], s);
        const unsigned* simT_c = simT;
        void* args[] = {(void*)&a, (void*)&b0, (void*)&nb, (void*)&W, (void*)&simT_c, (void*)&verdict, (void*)&rounds};
        e = cudaLaunchCooperativeKernel((const void*)tfd_block_resolve_kernel, dim3((unsigned)sm_count()), dim3(1024), args, 0, s);
        if (trace && e == cudaSuccess) {
            cudaEventRecord(ev[2], s);
            cudaEventSynchronize(ev[2]);
            float m1 = 0, m2 = 0;
            int h_rounds = 0;
            cudaEventElapsedTime(&m1, ev[0], ev[1]);
            cudaEventElapsedTime(&m2, ev[1], ev[2]);
            cudaMemcpy(&h_rounds, rounds, 4, cudaMemcpyDeviceToHost);
            fprintf(stderr, "  tfd sweep block %d (+%d rows, q = %d): pair matrix %.1f us, resolve %.1f us (%d rounds so far)\n", b0, nb, q,
                    1e3 * m1, 1e3 * m2, h_rounds);
        }
    }
    if (trace) for (auto& x : ev) cudaEventDestroy(x);
    cudaFreeAsync(simT, s);
    cudaFreeAsync(verdict, s);
    cudaFreeAsync(perm, s);
    cudaFreeAsync(gbox, s);
    cudaFreeAsync(sep, s);
    FC_CUDA(e);
    return FC_OK;
}

}  // namespace fc

using namespace fc;

// ---------------------------------------------------------------------------------------------
// result object
// ---------------------------------------------------------------------------------------------
extern "C" void fc_result_free(fc_result* r) { delete r; }
extern "C" int fc_result_counts(const fc_result* r, int64_t* out10) {
    FC_REQUIRE(r && out10, "fc_result_counts: null pointer");
    out10[0] = r->n_poses;
    out10[1] = r->n_clash_pass;
    out10[2] = r->n_rechecked;
    out10[3] = r->n_kept;
    out10[4] = r->ties_total;
    out10[5] = r->n_atoms;
    out10[6] = r->n_surv;
    out10[7] = r->n_quads;
    out10[8] = r->n_pairs;
    out10[9] = r->n_groups;
    return FC_OK;
}

namespace fc {
fc_result* result_new() { return new fc_result(); }
void result_set_cyclical(fc_result* r, int64_t n_poses, int64_t n_atoms, int64_t n_pass, std::vector<uint8_t>&& status,
                         std::vector<int64_t>&& kept, std::vector<double>&& coords, std::vector<int32_t>&& constrained,
                         int n_pairs, std::vector<fc_tie>&& ties, int64_t ties_total) {
    r->n_poses = n_poses;
    r->n_atoms = n_atoms;
    r->n_clash_pass = n_pass;
    r->n_surv = n_pass;
    r->status = std::move(status);
    for (uint8_t st : r->status) r->n_rechecked += (st & FC_STATUS_RECHECKED) ? 1 : 0;
    r->kept = std::move(kept);
    r->n_kept = (int64_t)r->kept.size();
    r->coords = std::move(coords);
    r->constrained = std::move(constrained);
    r->n_pairs = n_pairs;
    r->ties = std::move(ties);
    r->ties_total = ties_total;
}
}  // namespace fc
#define FC_COPY_OUT(vec, out)                                  \
    do {                                                       \
        FC_REQUIRE(r, "null result");                          \
        if (!(vec).empty()) {                                  \
            FC_REQUIRE(out, "null output pointer");            \
            memcpy(out, (vec).data(), (vec).size() * sizeof((vec)[0])); \
        }                                                      \
        return FC_OK;                                          \
    } while (0)
extern "C" int fc_result_status(const fc_result* r, uint8_t* out) { FC_COPY_OUT(r->status, out); }
extern "C" int fc_result_survivors(const fc_result* r, int64_t* out) { FC_COPY_OUT(r->survivors, out); }
extern "C" int fc_result_fingerprints(const fc_result* r, double* out) { FC_COPY_OUT(r->fingerprints, out); }
extern "C" int fc_result_kept_indices(const fc_result* r, int64_t* out) { FC_COPY_OUT(r->kept, out); }
namespace fc {
// atoms of the molecules placed by per-pose transforms: out (n, sum of atoms, 3)
__global__ void __launch_bounds__(128) place_kept_kernel(const double* __restrict__ xf, const int32_t* __restrict__ conf,
                                                         const double* c0, const double* c1, const double* c2, int n0,
                                                         int n1, int n2, int n_mols, long long n, double* __restrict__ out) {
    const long long k = blockIdx.x;
    if (k >= n) return;
    __shared__ double s_xf[3][12];
    if (threadIdx.x < n_mols * 12) s_xf[threadIdx.x / 12][threadIdx.x % 12] = xf[k * n_mols * 12 + threadIdx.x];
    __syncthreads();
    const int n_tot = n0 + n1 + n2;
    double* o = out + (size_t)k * n_tot * 3;
    for (int at = threadIdx.x; at < n_tot; at += blockDim.x) {
        const int m = at < n0 ? 0 : (at < n0 + n1 ? 1 : 2);
        const int local = at - (m == 0 ? 0 : (m == 1 ? n0 : n0 + n1));
        const int na = m == 0 ? n0 : (m == 1 ? n1 : n2);
        const double* base = m == 0 ? c0 : (m == 1 ? c1 : c2);
        const double* b = base + ((size_t)conf[k * n_mols + m] * na + local) * 3;
        const double* r = s_xf[m];
        // (R @ b) + t, the get_embed expression embeds.py:815-817
        o[3 * at] = (r[0] * b[0] + r[1] * b[1] + r[2] * b[2]) + r[9];
        o[3 * at + 1] = (r[3] * b[0] + r[4] * b[1] + r[5] * b[2]) + r[10];
        o[3 * at + 2] = (r[6] * b[0] + r[7] * b[1] + r[8] * b[2]) + r[11];
    }
}
}  // namespace fc

fc_result::~fc_result() {
    if (lazy.empty() && !lazy_coords[0] && !d_kept_coords) return;
    int cur = -1;
    cudaGetDevice(&cur);
    if (lazy_device >= 0 && lazy_device != cur) cudaSetDevice(lazy_device);
    if (d_kept_coords) cudaFree(d_kept_coords);
    for (LazySegment& sgm : lazy) {
        if (sgm.d_xf) cudaFree(sgm.d_xf);
        if (sgm.d_conf) cudaFree(sgm.d_conf);
    }
    for (int m = 0; m < 3; ++m)
        if (lazy_coords[m]) cudaFree(lazy_coords[m]);
    if (lazy_device >= 0 && lazy_device != cur && cur >= 0) cudaSetDevice(cur);
}

extern "C" int fc_result_kept_coords(const fc_result* r, double* out) {
    FC_REQUIRE(r, "null result");
    if (r->lazy.empty() && !r->d_kept_coords) FC_COPY_OUT(r->coords, out);
    FC_REQUIRE(out, "null output pointer");
    int cur = -1;
    FC_CUDA(cudaGetDevice(&cur));
    if (r->lazy_device >= 0 && r->lazy_device != cur) FC_CUDA(cudaSetDevice(r->lazy_device));
    if (r->d_kept_coords) {  // string embed: already placed, one copy into the caller's array
        cudaError_t ce = cudaMemcpy(out, r->d_kept_coords, r->d_kept_bytes, cudaMemcpyDeviceToHost);
        if (r->lazy_device >= 0 && r->lazy_device != cur) cudaSetDevice(cur);
        if (ce != cudaSuccess) return fc::cuda_fail(ce, "fc_result_kept_coords", __FILE__, __LINE__);
        return FC_OK;
    }
    const int n_tot = r->lazy_n_atoms[0] + r->lazy_n_atoms[1] + r->lazy_n_atoms[2];
    const size_t pose_bytes = (size_t)n_tot * 24;
    const int64_t slice = std::max<int64_t>(1, ((int64_t)256 << 20) / (int64_t)pose_bytes);  // 256 MB of coordinates per slice
    cudaStream_t s = nullptr;
    cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    double* d_out[2] = {nullptr, nullptr};
    for (int b = 0; b < 2 && e == cudaSuccess; ++b) e = cudaMalloc((void**)&d_out[b], (size_t)slice * pose_bytes);
    size_t done = 0;
    int b = 0;
    for (const fc_result::LazySegment& sgm : r->lazy) {
        for (int64_t lo = 0; lo < sgm.count && e == cudaSuccess; lo += slice, b ^= 1) {
            const int64_t n = std::min<int64_t>(slice, sgm.count - lo);
            fc::place_kept_kernel<<<(unsigned)n, 128, 0, s>>>(sgm.d_xf + lo * r->lazy_n_mols * 12, sgm.d_conf + lo * r->lazy_n_mols,
                                                              r->lazy_coords[0], r->lazy_coords[1], r->lazy_coords[2],
                                                              r->lazy_n_atoms[0], r->lazy_n_atoms[1], r->lazy_n_atoms[2],
                                                              r->lazy_n_mols, n, d_out[b]);
            e = cudaGetLastError();
            // download_staged returns once the slice sits in the caller's buffer; the kernel of the next slice is
            // short, so the copy engine and the host threads are what the loop keeps busy
            if (e == cudaSuccess) e = fc::download_staged(reinterpret_cast<char*>(out) + done, d_out[b], (size_t)n * pose_bytes, s);
            done += (size_t)n * pose_bytes;
        }
    }
    if (s) cudaStreamSynchronize(s);
    for (int k = 0; k < 2; ++k)
        if (d_out[k]) cudaFree(d_out[k]);
    if (s) cudaStreamDestroy(s);
    if (r->lazy_device >= 0 && r->lazy_device != cur) cudaSetDevice(cur);
    if (e != cudaSuccess) return fc::cuda_fail(e, "fc_result_kept_coords", __FILE__, __LINE__);
    return FC_OK;
}
extern "C" int fc_result_constrained(const fc_result* r, int32_t* out) { FC_COPY_OUT(r->constrained, out); }
extern "C" int fc_result_groups(const fc_result* r, int32_t* choice, double* gap) {
    FC_REQUIRE(r, "null result");
    if (choice && !r->group_choice.empty()) memcpy(choice, r->group_choice.data(), r->group_choice.size() * 4);
    if (gap && !r->group_gap.empty()) memcpy(gap, r->group_gap.data(), r->group_gap.size() * 8);
    return FC_OK;
}
extern "C" int64_t fc_result_ties(const fc_result* r, fc_tie* out, int64_t cap) {
    if (!r) return -1;
    int64_t n = std::min<int64_t>(cap, (int64_t)r->ties.size());
    if (n > 0 && out) memcpy(out, r->ties.data(), (size_t)n * sizeof(fc_tie));
    return (int64_t)r->ties.size();
}

// ---------------------------------------------------------------------------------------------
// string embed driver
// ---------------------------------------------------------------------------------------------
namespace {

struct StringCtx {
    cudaStream_t s = nullptr;
    DevBuf<double> coords1, coords2, cen1, vec1, cen2, vec2, angles;
    DevBuf<long long> quads;
    StringDev dev{};
    ~StringCtx() {
        // buffers release on the stream before it is destroyed
        coords1.release(); coords2.release(); cen1.release(); vec1.release(); cen2.release();
        vec2.release(); angles.release(); quads.release();
        if (s) {
            cudaStreamSynchronize(s);
            cudaStreamDestroy(s);
        }
    }
};

int upload(DevBuf<double>& b, const double* src, size_t n, cudaStream_t s) {
    FC_CUDA(b.alloc(n, s));
    if (n) FC_CUDA(cudaMemcpyAsync(b.p, src, n * 8, cudaMemcpyHostToDevice, s));
    return FC_OK;
}

int string_ctx_init(StringCtx& c, const fc_string_problem* p) {
    FC_REQUIRE(p, "null problem");
    FC_REQUIRE(p->n_conf1 > 0 && p->n_conf2 > 0 && p->n_atoms1 > 0 && p->n_atoms2 > 0, "empty ensemble");
    FC_REQUIRE(p->k1 > 0 && p->k2 > 0 && p->n_angles > 0, "no orbital centres / angles");
    FC_REQUIRE(p->coords1 && p->coords2 && p->centers1 && p->centers2 && p->vecs1 && p->vecs2 && p->angles,
               "null pointer in fc_string_problem");
    FC_REQUIRE(p->n_quads == 0 || p->quadruplets, "null quadruplets");
    sm_count();
    FC_CUDA(cudaStreamCreateWithFlags(&c.s, cudaStreamNonBlocking));
    int rc;
    if ((rc = upload(c.coords1, p->coords1, (size_t)p->n_conf1 * p->n_atoms1 * 3, c.s))) return rc;
    if ((rc = upload(c.coords2, p->coords2, (size_t)p->n_conf2 * p->n_atoms2 * 3, c.s))) return rc;
    if ((rc = upload(c.cen1, p->centers1, (size_t)p->n_conf1 * p->k1 * 3, c.s))) return rc;
    if ((rc = upload(c.vec1, p->vecs1, (size_t)p->n_conf1 * p->k1 * 3, c.s))) return rc;
    if ((rc = upload(c.cen2, p->centers2, (size_t)p->n_conf2 * p->k2 * 3, c.s))) return rc;
    if ((rc = upload(c.vec2, p->vecs2, (size_t)p->n_conf2 * p->k2 * 3, c.s))) return rc;
    if ((rc = upload(c.angles, p->angles, (size_t)p->n_angles, c.s))) return rc;
    FC_CUDA(c.quads.alloc((size_t)p->n_quads * 4, c.s));
    if (p->n_quads)
        FC_CUDA(cudaMemcpyAsync(c.quads.p, p->quadruplets, (size_t)p->n_quads * 32, cudaMemcpyHostToDevice, c.s));
    int n_tot = p->n_atoms1 + p->n_atoms2;
    for (int i = 0; i < p->n_quads * 4; ++i)
        FC_REQUIRE(p->quadruplets[i] >= 0 && p->quadruplets[i] < n_tot, "quadruplet index out of range");
    StringDev& d = c.dev;
    d.coords1 = c.coords1.p; d.coords2 = c.coords2.p;
    d.cen1 = c.cen1.p; d.vec1 = c.vec1.p; d.cen2 = c.cen2.p; d.vec2 = c.vec2.p;
    d.angles = c.angles.p;
    d.c1 = p->n_conf1; d.n1 = p->n_atoms1; d.c2 = p->n_conf2; d.n2 = p->n_atoms2;
    d.k1 = p->k1; d.k2 = p->k2; d.na = p->n_angles;
    d.handed = p->rot_handedness >= 0 ? 1 : -1;
    return FC_OK;
}

int64_t string_total(const fc_string_problem* p) {
    return (int64_t)p->n_conf1 * p->n_conf2 * p->k1 * p->k2 * p->n_angles;
}

// tiles of the clash screen for poses [lo, hi): runs of equal conformer pair cut to tile_poses
void string_tiles(const fc_string_problem* p, int64_t lo, int64_t hi, int tile_poses, std::vector<int32_t>& out) {
    const int64_t per_conf = (int64_t)p->k1 * p->k2 * p->n_angles;
    int64_t pose = lo;
    while (pose < hi) {
        int64_t ci = pose / per_conf;
        int64_t run_end = std::min<int64_t>(hi, (ci + 1) * per_conf);
        int c1 = (int)(ci % p->n_conf1), c2 = (int)(ci / p->n_conf1);
        while (pose < run_end) {
            int cnt = (int)std::min<int64_t>(tile_poses, run_end - pose);
            out.push_back(c1);
            out.push_back(c2);
            out.push_back((int32_t)(pose - lo));
            out.push_back(cnt);
            pose += cnt;
        }
    }
}

// stage 1 on the device: transforms, clash screen, survivors, fingerprints. Leaves results on host.
int string_stage1(StringCtx& c, const fc_string_problem* p, int64_t lo, int64_t hi, fc_result* r,
                  DevBuf<double>* keep_xf, DevBuf<long long>* keep_surv, DevBuf<double>* keep_fp,
                  DevBuf<TieRecord>& d_ties, DevBuf<int>& d_nties, int tie_cap) {
    const int64_t n = hi - lo;
    r->n_poses = n;
    r->n_atoms = p->n_atoms1 + p->n_atoms2;
    r->n_quads = p->n_quads;
    if (n == 0) return FC_OK;
    FC_REQUIRE(n < ((int64_t)1 << 31), "pose range too large for one call (%lld)", (long long)n);
    cudaStream_t s = c.s;
    DevBuf<double>& xf = *keep_xf;
    FC_CUDA(xf.alloc((size_t)n * 12, s));
    string_xf_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(c.dev, lo, n, xf.p);
    FC_CUDA(cudaGetLastError());

    std::vector<int32_t> tiles;
    string_tiles(p, lo, hi, fc_clash_tile_poses(p->n_atoms2), tiles);
    DevBuf<int32_t> d_tiles;
    FC_CUDA(d_tiles.alloc(tiles.size(), s));
    FC_CUDA(cudaMemcpyAsync(d_tiles.p, tiles.data(), tiles.size() * 4, cudaMemcpyHostToDevice, s));
    DevBuf<uint8_t> status;
    FC_CUDA(status.alloc((size_t)n, s));
    DevBuf<int32_t> near_count;
    DevBuf<int64_t> near_idx;
    DevBuf<double> near_dist;
    const int near_cap = 1 << 16;
    FC_CUDA(near_count.alloc(4, s));
    FC_CUDA(near_idx.alloc(near_cap, s));
    FC_CUDA(near_dist.alloc(near_cap, s));
    FC_CUDA(cudaMemsetAsync(near_count.p, 0, 16, s));
    int rc = fc_clash_screen_dev(c.coords1.p, p->n_conf1, p->n_atoms1, c.coords2.p, p->n_conf2, p->n_atoms2, xf.p,
                                 n, d_tiles.p, (int64_t)tiles.size() / 4, p->thresh, p->max_clashes, 1, status.p,
                                 nullptr, near_count.p, near_idx.p, near_dist.p, near_cap, lo, (void*)s);
    if (rc) return rc;

    DevBuf<long long>& surv = *keep_surv;
    DevBuf<int> d_count;
    FC_CUDA(surv.alloc((size_t)n, s));
    FC_CUDA(d_count.alloc(4, s));
    if ((rc = compact_pass(status.p, n, lo, surv.p, d_count.p, s))) return rc;
    int n_surv = 0, n_near = 0;
    FC_CUDA(cudaMemcpyAsync(&n_surv, d_count.p, 4, cudaMemcpyDeviceToHost, s));
    FC_CUDA(cudaMemcpyAsync(&n_near, near_count.p, 4, cudaMemcpyDeviceToHost, s));
    r->status.resize((size_t)n);
    FC_CUDA(cudaMemcpyAsync(r->status.data(), status.p, (size_t)n, cudaMemcpyDeviceToHost, s));
    FC_CUDA(cudaStreamSynchronize(s));
    r->n_surv = n_surv;
    r->n_clash_pass = n_surv;
    for (uint8_t st : r->status) r->n_rechecked += (st & FC_STATUS_RECHECKED) ? 1 : 0;
    // near-threshold clash decisions -> tie list
    n_near = std::min(n_near, near_cap);
    if (n_near > 0) {
        std::vector<int64_t> idx(n_near);
        std::vector<double> dist(n_near);
        FC_CUDA(cudaMemcpy(idx.data(), near_idx.p, (size_t)n_near * 8, cudaMemcpyDeviceToHost));
        FC_CUDA(cudaMemcpy(dist.data(), near_dist.p, (size_t)n_near * 8, cudaMemcpyDeviceToHost));
        for (int i = 0; i < n_near; ++i) {
            fc_tie t;
            t.a = idx[i];
            t.b = -1;
            t.value = dist[i];
            t.kind = FC_TIE_CLASH;
            t.decision = (r->status[(size_t)(idx[i] - lo)] & FC_STATUS_PASS) ? 0 : 1;  // 1 = "below threshold"
            r->ties.push_back(t);
        }
        r->ties_total += n_near;
    }
    r->survivors.resize((size_t)n_surv);
    if (n_surv)
        FC_CUDA(cudaMemcpyAsync(r->survivors.data(), surv.p, (size_t)n_surv * 8, cudaMemcpyDeviceToHost, s));
    // fingerprints
    DevBuf<double>& fp = *keep_fp;
    FC_CUDA(fp.alloc((size_t)n_surv * std::max(1, p->n_quads), s));
    if (n_surv && p->n_quads) {
        long long total = (long long)n_surv * p->n_quads;
        tfd_fingerprint_kernel<<<(unsigned)((total + 127) / 128), 128, 0, s>>>(c.dev, surv.p, n_surv, lo, xf.p,
                                                                               c.quads.p, p->n_quads, fp.p);
        FC_CUDA(cudaGetLastError());
    }
    (void)d_ties; (void)d_nties; (void)tie_cap;
    return FC_OK;
}

int fetch_ties(fc_result* r, DevBuf<TieRecord>& d_ties, const int* d_count, int tie_cap, cudaStream_t s) {
    int n_ties = 0;
    FC_CUDA(cudaMemcpyAsync(&n_ties, d_count, 4, cudaMemcpyDeviceToHost, s));
    FC_CUDA(cudaStreamSynchronize(s));
    r->ties_total += n_ties;
    int n_copy = std::min(n_ties, tie_cap);
    if (n_copy > 0) {
        std::vector<TieRecord> tmp(n_copy);
        FC_CUDA(cudaMemcpy(tmp.data(), d_ties.p, (size_t)n_copy * sizeof(TieRecord), cudaMemcpyDeviceToHost));
        for (const TieRecord& t : tmp) {
            fc_tie o;
            o.a = t.a; o.b = t.b; o.value = t.value; o.kind = t.kind; o.decision = t.decision;
            r->ties.push_back(o);
        }
    }
    return FC_OK;
}

}  // namespace

extern "C" int64_t fc_string_n_poses(const fc_string_problem* p) { return p ? string_total(p) : -1; }

extern "C" int fc_string_stage1(const fc_string_problem* p, int64_t pose_lo, int64_t pose_hi, fc_result** out) {
    FC_REQUIRE(out, "null output");
    *out = nullptr;
    StringCtx c;
    int rc = string_ctx_init(c, p);
    if (rc) return rc;
    int64_t total = string_total(p);
    if (pose_hi < 0 || pose_hi > total) pose_hi = total;
    FC_REQUIRE(pose_lo >= 0 && pose_lo <= pose_hi, "bad pose range");
    fc_result* r = new fc_result();
    DevBuf<double> xf, fp;
    DevBuf<long long> surv;
    DevBuf<TieRecord> d_ties;
    DevBuf<int> d_nties;
    rc = string_stage1(c, p, pose_lo, pose_hi, r, &xf, &surv, &fp, d_ties, d_nties, 0);
    if (!rc && r->n_surv && p->n_quads) {
        r->fingerprints.resize((size_t)r->n_surv * p->n_quads);
        cudaError_t e = cudaMemcpyAsync(r->fingerprints.data(), fp.p, r->fingerprints.size() * 8,
                                        cudaMemcpyDeviceToHost, c.s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c.s);
        if (e != cudaSuccess) rc = cuda_fail(e, "fingerprint copy", __FILE__, __LINE__);
    }
    if (rc) {
        delete r;
        return rc;
    }
    cudaStreamSynchronize(c.s);
    *out = r;
    return FC_OK;
}

// Ordered keep-first sweep on host fingerprints (the merge step every rank runs after the
// all-gather of the survivors' fingerprints).  keep_out[i] = 1 if row i is accepted.
extern "C" int fc_tfd_keepfirst(const double* fingerprints, const int64_t* labels, int64_t n, int32_t n_quads,
                                double thresh, uint8_t* keep_out, fc_tie* ties_out, int64_t tie_cap,
                                int64_t* n_ties_out) {
    FC_REQUIRE(n >= 0 && n < ((int64_t)1 << 31) && n_quads >= 0, "fc_tfd_keepfirst: bad size");
    if (n_ties_out) *n_ties_out = 0;
    if (n == 0) return FC_OK;
    FC_REQUIRE(keep_out && (fingerprints || n_quads == 0), "fc_tfd_keepfirst: null pointer");
    sm_count();
    cudaStream_t s;
    FC_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    int rc = FC_OK;
    {
        DevBuf<double> fp;
        DevBuf<long long> lab;
        DevBuf<int> flag, acc, cnt;
        DevBuf<TieRecord> d_ties;
        const int cap = (int)std::min<int64_t>(std::max<int64_t>(tie_cap, 1), 1 << 22);
        cudaError_t e = fp.alloc((size_t)n * std::max(1, n_quads), s);
        if (e == cudaSuccess) e = lab.alloc((size_t)n, s);
        if (e == cudaSuccess) e = flag.alloc((size_t)n, s);
        if (e == cudaSuccess) e = acc.alloc((size_t)n, s);
        if (e == cudaSuccess) e = cnt.alloc(8, s);
        if (e == cudaSuccess) e = d_ties.alloc((size_t)cap, s);
        if (e == cudaSuccess && n_quads) e = cudaMemcpyAsync(fp.p, fingerprints, (size_t)n * n_quads * 8, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess && labels) e = cudaMemcpyAsync(lab.p, labels, (size_t)n * 8, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemsetAsync(cnt.p, 0, 32, s);
        if (e != cudaSuccess) rc = cuda_fail(e, "fc_tfd_keepfirst setup", __FILE__, __LINE__);
        if (!rc)
            rc = tfd_keepfirst_dev(fp.p, labels ? lab.p : nullptr, (int)n, n_quads, thresh, FC_NEAR_EPS, flag.p, acc.p,
                                   cnt.p, ties_out ? d_ties.p : nullptr, cnt.p + 1, cap, s);
        if (!rc) {
            std::vector<int> h_flag((size_t)n);
            int h_cnt[2] = {0, 0};
            e = cudaMemcpyAsync(h_flag.data(), flag.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaMemcpyAsync(h_cnt, cnt.p, 8, cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) rc = cuda_fail(e, "fc_tfd_keepfirst readback", __FILE__, __LINE__);
            if (!rc) {
                for (int64_t i = 0; i < n; ++i) keep_out[i] = h_flag[(size_t)i] ? 0 : 1;
                if (n_ties_out) *n_ties_out = h_cnt[1];
                int n_copy = (int)std::min<int64_t>(std::min<int64_t>(h_cnt[1], cap), tie_cap);
                if (n_copy > 0 && ties_out) {
                    std::vector<TieRecord> tmp(n_copy);
                    e = cudaMemcpy(tmp.data(), d_ties.p, (size_t)n_copy * sizeof(TieRecord), cudaMemcpyDeviceToHost);
                    if (e != cudaSuccess) rc = cuda_fail(e, "tie copy", __FILE__, __LINE__);
                    for (int i = 0; i < n_copy && !rc; ++i) {
                        ties_out[i].a = tmp[i].a; ties_out[i].b = tmp[i].b; ties_out[i].value = tmp[i].value;
                        ties_out[i].kind = tmp[i].kind; ties_out[i].decision = tmp[i].decision;
                    }
                }
            }
        }
    }
    cudaStreamSynchronize(s);
    cudaStreamDestroy(s);
    return rc;
}

extern "C" int fc_string_materialize(const fc_string_problem* p, const int64_t* kept, int64_t n_kept, double* out) {
    FC_REQUIRE(n_kept >= 0, "negative size");
    if (n_kept == 0) return FC_OK;
    FC_REQUIRE(kept && out, "null pointer");
    StringCtx c;
    int rc = string_ctx_init(c, p);
    if (rc) return rc;
    const int64_t total = string_total(p);
    for (int64_t i = 0; i < n_kept; ++i) FC_REQUIRE(kept[i] >= 0 && kept[i] < total, "kept index out of range");
    const size_t n_tot = (size_t)p->n_atoms1 + p->n_atoms2;
    DevBuf<long long> d_kept;
    DevBuf<double> d_out;
    FC_CUDA(d_kept.alloc((size_t)n_kept, c.s));
    FC_CUDA(d_out.alloc((size_t)n_kept * n_tot * 3, c.s));
    FC_CUDA(cudaMemcpyAsync(d_kept.p, kept, (size_t)n_kept * 8, cudaMemcpyHostToDevice, c.s));
    string_materialize_kernel<<<(unsigned)n_kept, 128, 0, c.s>>>(c.dev, d_kept.p, (int)n_kept, d_out.p);
    FC_CUDA(cudaGetLastError());
    FC_CUDA(cudaMemcpyAsync(out, d_out.p, (size_t)n_kept * n_tot * 24, cudaMemcpyDeviceToHost, c.s));
    FC_CUDA(cudaStreamSynchronize(c.s));
    return FC_OK;
}

// Whole string screen on one GPU: stage 1 -> keep-first sweep -> materialisation, device resident.
extern "C" int fc_string_screen(const fc_string_problem* p, fc_result** out) {
    FC_REQUIRE(out, "null output");
    *out = nullptr;
    // FC_STRING_TRACE=1: wall-clock stage times on stderr (adds a stream synchronisation per stage)
    const bool trace = getenv("FC_STRING_TRACE") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    double t_prev = t_begin;
    auto mark = [&](const char* what, cudaStream_t st) {
        if (!trace) return;
        if (st) cudaStreamSynchronize(st);
        const double t = now();
        fprintf(stderr, "  fc_string_screen %-28s %8.1f us\n", what, t - t_prev);
        t_prev = t;
    };
    StringCtx c;
    int rc = string_ctx_init(c, p);
    if (rc) return rc;
    mark("context + uploads", c.s);
    const int64_t total = string_total(p);
    fc_result* r = new fc_result();
    DevBuf<double> xf, fp, d_coords;
    DevBuf<long long> surv, d_kept;
    DevBuf<TieRecord> d_ties;
    DevBuf<int> d_cnt, flag, acc;
    const int tie_cap = 1 << 20;
    cudaStream_t s = c.s;
    rc = string_stage1(c, p, 0, total, r, &xf, &surv, &fp, d_ties, d_cnt, tie_cap);
    mark("stage 1 (xf, clash, fp)", s);
    const int n_surv = (int)r->n_surv;
    if (!rc && n_surv > 0) {
        cudaError_t e = d_ties.alloc(tie_cap, s);
        if (e == cudaSuccess) e = d_cnt.alloc(8, s);
        if (e == cudaSuccess) e = flag.alloc(n_surv, s);
        if (e == cudaSuccess) e = acc.alloc(n_surv, s);
        if (e == cudaSuccess) e = cudaMemsetAsync(d_cnt.p, 0, 32, s);
        if (e != cudaSuccess) rc = cuda_fail(e, "fc_string_screen alloc", __FILE__, __LINE__);
        if (!rc)
            rc = tfd_keepfirst_dev(fp.p, surv.p, n_surv, p->n_quads, p->tfd_thresh, FC_NEAR_EPS, flag.p, acc.p, d_cnt.p,
                                   d_ties.p, d_cnt.p + 1, tie_cap, s);
        int n_kept = 0;
        if (!rc) {
            e = cudaMemcpyAsync(&n_kept, d_cnt.p, 4, cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) rc = cuda_fail(e, "n_kept readback", __FILE__, __LINE__);
        }
        mark("keep-first sweep", s);
        if (!rc && n_kept > 0) {
            // accepted rows (ascending) -> absolute pose indices, on the device; the coordinates of the kept poses are
            // placed there too and stay with the result until the caller asks for them
            const size_t n_tot = (size_t)r->n_atoms;
            e = d_kept.alloc(n_kept, s);
            if (e == cudaSuccess) {
                r->d_kept_bytes = (size_t)n_kept * n_tot * 24;
                e = cudaMallocAsync((void**)&r->d_kept_coords, r->d_kept_bytes, s);
                cudaGetDevice(&r->lazy_device);
            }
            if (e == cudaSuccess) {
                gather_index_kernel<<<(unsigned)((n_kept + 255) / 256), 256, 0, s>>>(surv.p, acc.p, n_kept, d_kept.p);
                string_materialize_kernel<<<n_kept, 128, 0, s>>>(c.dev, d_kept.p, n_kept, r->d_kept_coords);
                e = cudaGetLastError();
            }
            r->kept.resize(n_kept);
            if (e == cudaSuccess) e = cudaMemcpyAsync(r->kept.data(), d_kept.p, (size_t)n_kept * 8, cudaMemcpyDeviceToHost, s);
            if (e != cudaSuccess) rc = cuda_fail(e, "materialize", __FILE__, __LINE__);
            r->n_kept = n_kept;
        }
        if (!rc) rc = fetch_ties(r, d_ties, d_cnt.p + 1, tie_cap, s);
        mark("materialise + ties", s);
    }
    if (rc) {
        delete r;
        return rc;
    }
    *out = r;
    if (trace) fprintf(stderr, "  fc_string_screen total %.1f us\n", now() - t_begin);
    return FC_OK;
}
