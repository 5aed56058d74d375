// firecode_b200 -- torsion-fingerprint (TFD) ensemble pruning: the O(n^2 Q) part of
// firecode/torsion_module.py:957-1043 `prune_conformers_tfd` (called from the `tfd` branch of
// similarity_refining, embedder.py:1430-1437, and from the conformational search, torsion_module.py:875).
//
//   fc_tfd_fingerprints : `_get_tf_mat` (torsion_module.py:1046-1053): dihedral of every quadruplet of every
//                         structure, degrees, atan2 form (prism_pruner.algebra.dihedral).
//   fc_tfd_first_match  : for every structure i of every chunk of a pass, the FIRST later structure j of the
//                         same chunk with  sum_q wrap180(|tf_i[q] - tf_j[q]|) < thresh  (tfd_similarity,
//                         torsion_module.py:1056-1067) -- exactly what the reference's pair loops record
//                         (they `break` at the first match; the cache of dissimilar pairs only saves work).
// The order-dependent part (match graph, connected components, first node of each cluster survives) stays on
// the host in Python with the same networkx calls as the reference: its outcome depends on CPython set and
// networkx iteration order, which only the same code on the same inputs reproduces.
#include <algorithm>
#include <vector>

#include "fc_embed.cuh"

namespace fc {

__global__ void tfd_matrix_kernel(const double* __restrict__ x, long long n, int n_atoms, const long long* __restrict__ quads,
                                  int nq, double* __restrict__ tf) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * nq) return;
    const long long s = i / nq;
    const int q = (int)(i - s * nq);
    const double* c = x + (size_t)s * n_atoms * 3;
    const double* p0 = c + 3 * quads[4 * q];
    const double* p1 = c + 3 * quads[4 * q + 1];
    const double* p2 = c + 3 * quads[4 * q + 2];
    const double* p3 = c + 3 * quads[4 * q + 3];
    tf[i] = dihedral_deg(p0, p1, p2, p3);
}

struct TfdArgs {
    const double* tf;        // (n, nq)
    int nq;
    const long long* start;  // per chunk
    const long long* len;    // per chunk
    const long long* row0;   // per chunk: prefix sum of len (n_chunks + 1)
    long long n_chunks, n_rows;
    double thresh, eps;
    long long* first;        // (n) absolute index of the first match, -1 if none
    TieRecord* ties;
    int* n_ties;
    int tie_cap;
};

// one warp per structure of a chunk; lanes scan 32 later structures at a time
__global__ void __launch_bounds__(128) tfd_first_match_kernel(TfdArgs a) {
    extern __shared__ double s_tf[];  // one fingerprint per warp
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + wib;
    if (r >= a.n_rows) return;
    long long lo = 0, hi = a.n_chunks - 1;  // chunk of this row
    while (lo < hi) {
        long long mid = (lo + hi + 1) >> 1;
        if (a.row0[mid] <= r) lo = mid;
        else hi = mid - 1;
    }
    const long long i = a.start[lo] + (r - a.row0[lo]);
    const long long end = a.start[lo] + a.len[lo];
    double* mine = s_tf + (size_t)wib * a.nq;
    for (int q = lane; q < a.nq; q += 32) mine[q] = a.tf[(size_t)i * a.nq + q];
    __syncwarp();
    long long found = -1;
    for (long long j0 = i + 1; j0 < end && found < 0; j0 += 32) {
        const long long j = j0 + lane;
        bool similar = false;
        if (j < end) {
            const double* other = a.tf + (size_t)j * a.nq;
            double sum = 0.0;
            for (int q = 0; q < a.nq; ++q) {
                double d = fabs(mine[q] - other[q]);
                d = fabs(d - (d > 180.0 ? 360.0 : 0.0));  // torsion_module.py:1062
                sum += d;
            }
            similar = sum < a.thresh;
            if (a.ties && fabs(sum - a.thresh) <= a.eps) {
                int slot = atomicAdd(a.n_ties, 1);
                if (slot < a.tie_cap) a.ties[slot] = TieRecord{j, i, sum, FC_TIE_TFD, similar ? 1 : 0};
            }
        }
        const unsigned hit = __ballot_sync(0xffffffffu, similar);
        if (hit) found = j0 + __ffs(hit) - 1;
    }
    if (lane == 0) a.first[i] = found;
}

}  // namespace fc

using namespace fc;

extern "C" int fc_tfd_fingerprints(const double* structures, int64_t n, int32_t n_atoms, const int64_t* quadruplets,
                                   int32_t n_quads, double* tf_out) {
    FC_REQUIRE(n >= 0 && n_atoms > 0 && n_quads >= 0, "fc_tfd_fingerprints: bad sizes");
    if (n == 0 || n_quads == 0) return FC_OK;
    FC_REQUIRE(structures && quadruplets && tf_out, "fc_tfd_fingerprints: null pointer");
    for (int64_t k = 0; k < (int64_t)n_quads * 4; ++k)
        FC_REQUIRE(quadruplets[k] >= 0 && quadruplets[k] < n_atoms, "fc_tfd_fingerprints: quadruplet index out of range");
    sm_count();
    cudaStream_t s;
    FC_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    cudaError_t e = cudaSuccess;
    {
        DevBuf<double> d_x, d_tf;
        DevBuf<long long> d_q;
#define TF(call) do { if (e == cudaSuccess) e = (call); } while (0)
        TF(d_x.alloc((size_t)n * n_atoms * 3, s));
        TF(d_tf.alloc((size_t)n * n_quads, s));
        TF(d_q.alloc((size_t)n_quads * 4, s));
        TF(cudaMemcpyAsync(d_x.p, structures, (size_t)n * n_atoms * 24, cudaMemcpyHostToDevice, s));
        TF(cudaMemcpyAsync(d_q.p, quadruplets, (size_t)n_quads * 32, cudaMemcpyHostToDevice, s));
        if (e == cudaSuccess) {
            const long long total = (long long)n * n_quads;
            tfd_matrix_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(d_x.p, n, n_atoms, d_q.p, n_quads, d_tf.p);
            e = cudaGetLastError();
        }
        TF(cudaMemcpyAsync(tf_out, d_tf.p, (size_t)n * n_quads * 8, cudaMemcpyDeviceToHost, s));
        TF(cudaStreamSynchronize(s));
    }
    cudaStreamSynchronize(s);
    cudaStreamDestroy(s);
    if (e != cudaSuccess) return cuda_fail(e, "fc_tfd_fingerprints", __FILE__, __LINE__);
    return FC_OK;
}

extern "C" int fc_tfd_first_match(const double* tf, int64_t n, int32_t n_quads, const int64_t* chunk_start,
                                  const int64_t* chunk_len, int64_t n_chunks, double thresh, int64_t* first_out,
                                  fc_tie* ties_out, int64_t tie_cap, int64_t* n_ties_out) {
    FC_REQUIRE(n >= 0 && n_quads >= 0 && n_chunks >= 0, "fc_tfd_first_match: bad sizes");
    if (n_ties_out) *n_ties_out = 0;
    if (n == 0) return FC_OK;
    FC_REQUIRE(first_out, "fc_tfd_first_match: null output");
    for (int64_t i = 0; i < n; ++i) first_out[i] = -1;
    if (n_chunks == 0) return FC_OK;
    FC_REQUIRE(chunk_start && chunk_len && (tf || n_quads == 0), "fc_tfd_first_match: null pointer");
    std::vector<long long> start, len, row0;
    row0.push_back(0);
    for (int64_t c = 0; c < n_chunks; ++c) {
        FC_REQUIRE(chunk_start[c] >= 0 && chunk_len[c] >= 0 && chunk_start[c] + chunk_len[c] <= n,
                   "fc_tfd_first_match: chunk %lld out of range", (long long)c);
        if (chunk_len[c] < 2) continue;  // nothing to compare
        start.push_back(chunk_start[c]);
        len.push_back(chunk_len[c]);
        row0.push_back(row0.back() + chunk_len[c]);
    }
    if (start.empty()) return FC_OK;
    sm_count();
    cudaStream_t s;
    FC_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    cudaError_t e = cudaSuccess;
    int n_t = 0;
    {
        DevBuf<double> d_tf;
        DevBuf<long long> d_start, d_len, d_row0, d_first;
        DevBuf<TieRecord> d_ties;
        DevBuf<int> d_nt;
        const int cap = (int)std::min<int64_t>(std::max<int64_t>(tie_cap, 1), 1 << 20);
        TF(d_tf.alloc((size_t)n * std::max(n_quads, 1), s));
        TF(d_start.alloc(start.size(), s));
        TF(d_len.alloc(len.size(), s));
        TF(d_row0.alloc(row0.size(), s));
        TF(d_first.alloc((size_t)n, s));
        TF(d_ties.alloc(cap, s));
        TF(d_nt.alloc(4, s));
        if (n_quads) TF(cudaMemcpyAsync(d_tf.p, tf, (size_t)n * n_quads * 8, cudaMemcpyHostToDevice, s));
        TF(cudaMemcpyAsync(d_start.p, start.data(), start.size() * 8, cudaMemcpyHostToDevice, s));
        TF(cudaMemcpyAsync(d_len.p, len.data(), len.size() * 8, cudaMemcpyHostToDevice, s));
        TF(cudaMemcpyAsync(d_row0.p, row0.data(), row0.size() * 8, cudaMemcpyHostToDevice, s));
        TF(cudaMemsetAsync(d_first.p, 0xff, (size_t)n * 8, s));
        TF(cudaMemsetAsync(d_nt.p, 0, 16, s));
        if (e == cudaSuccess) {
            TfdArgs a{};
            a.tf = d_tf.p; a.nq = n_quads;
            a.start = d_start.p; a.len = d_len.p; a.row0 = d_row0.p;
            a.n_chunks = (long long)start.size(); a.n_rows = row0.back();
            a.thresh = thresh; a.eps = FC_NEAR_EPS;
            a.first = d_first.p;
            a.ties = ties_out ? d_ties.p : nullptr; a.n_ties = d_nt.p; a.tie_cap = cap;
            const size_t smem = (size_t)4 * std::max(n_quads, 1) * sizeof(double);
            if (smem > 48 * 1024) TF(cudaFuncSetAttribute(tfd_first_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (e == cudaSuccess) {
                tfd_first_match_kernel<<<(unsigned)((a.n_rows + 3) / 4), 128, smem, s>>>(a);
                e = cudaGetLastError();
            }
        }
        TF(cudaMemcpyAsync(first_out, d_first.p, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
        TF(cudaMemcpyAsync(&n_t, d_nt.p, 4, cudaMemcpyDeviceToHost, s));
        TF(cudaStreamSynchronize(s));
        if (e == cudaSuccess && ties_out && n_t > 0) {
            const int n_copy = (int)std::min<int64_t>(std::min<int64_t>(n_t, cap), tie_cap);
            std::vector<TieRecord> tmp(n_copy);
            e = cudaMemcpy(tmp.data(), d_ties.p, (size_t)n_copy * sizeof(TieRecord), cudaMemcpyDeviceToHost);
            for (int i = 0; i < n_copy; ++i) {
                ties_out[i].a = tmp[i].a; ties_out[i].b = tmp[i].b; ties_out[i].value = tmp[i].value;
                ties_out[i].kind = tmp[i].kind; ties_out[i].decision = tmp[i].decision;
            }
        }
#undef TF
    }
    cudaStreamSynchronize(s);
    cudaStreamDestroy(s);
    if (e != cudaSuccess) return cuda_fail(e, "fc_tfd_first_match", __FILE__, __LINE__);
    if (n_ties_out) *n_ties_out = n_t;
    return FC_OK;
}
