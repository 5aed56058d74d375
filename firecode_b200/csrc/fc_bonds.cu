// firecode_b200 -- batched bond-graph checks: the post-filters `scramble_check` and `molecule_check`
// (/root/reference/firecode/utils.py:341-400; called per optimised structure from embedder.py:2181-2195, 2430-2444,
// optimization_methods.py:139-147, operators.py:502, interfaces/goat.py:309).
//
// Both compare the bond set of a structure -- `graphize(atoms, coords)` of prism_pruner.graph_manipulations: atoms i < j
// are bonded iff  |x_i - x_j| < d_min_bond(e_i, e_j) = factor * (r_cov[e_i] + r_cov[e_j])  -- with an expected bond set
// (the union of the fragments' graphs for scramble_check, the graph of the un-optimised structure for molecule_check)
// and count the bonds that appear in exactly one of the two, leaving out bonds that touch an excluded (constrained)
// atom.  The reference builds two networkx graphs per structure (O(n^2) Python each); here one CTA per structure walks
// the atom pairs out of shared memory and compares against a bit matrix.  prism_pruner is absent from the reference
// tree: the criterion is the one restated in oracle/prism_pruner/graph_manipulations.py (PARITY UNPINNED at that
// boundary, like the pruner); the set arithmetic around it is the reference's own and is pinned to it.
//
// Distances are FP64 with the reference's expression (sqrt of the summed squares, strict <).  A pair whose distance
// lies within FC_NEAR_EPS of its limit is counted in near_out, so a caller can tell a delicate verdict.
#include "fc_embed.cuh"

namespace fc {

struct BondArgs {
    const double* coords;      // (n_struct, n, 3)
    const double* radii;       // (n) covalent radius of every atom
    double factor;
    int n, words;              // words = ceil(n / 32)
    const unsigned* expected;  // (n, words) or (n_struct, n, words): bit j of row i = "i - j bonded" (symmetric)
    int expected_per_structure;
    const unsigned char* excluded;  // (n) or null
    unsigned* adj_out;         // (n_struct, n, words) or null: the structure's own bond bits
    int* delta_out;            // (n_struct) or null
    int* near_out;             // (n_struct) or null
};

__global__ void __launch_bounds__(256) bond_graph_kernel(BondArgs a) {
    extern __shared__ double s_x[];  // n x 3 coordinates, n radii
    const long long p = blockIdx.x;
    const int n = a.n;
    const double* x = a.coords + (size_t)p * n * 3;
    for (int e = threadIdx.x; e < 3 * n; e += blockDim.x) s_x[e] = x[e];
    double* s_r = s_x + 3 * n;
    for (int e = threadIdx.x; e < n; e += blockDim.x) s_r[e] = a.radii[e];
    __syncthreads();
    const unsigned* expected = a.expected ? a.expected + (a.expected_per_structure ? (size_t)p * n * a.words : 0) : nullptr;
    unsigned* adj = a.adj_out ? a.adj_out + (size_t)p * n * a.words : nullptr;
    int delta = 0, near = 0;
    // thread = (row i, word w): the 32 partners j = 32 w .. 32 w + 31 of atom i
    for (int item = threadIdx.x; item < n * a.words; item += blockDim.x) {
        const int i = item / a.words, w = item - i * a.words;
        const double xi = s_x[3 * i], yi = s_x[3 * i + 1], zi = s_x[3 * i + 2], ri = s_r[i];
        unsigned bits = 0u;
        const int j_end = min(32, n - 32 * w);
        for (int b = 0; b < j_end; ++b) {
            const int j = 32 * w + b;
            if (j == i) continue;
            const double dx = xi - s_x[3 * j], dy = yi - s_x[3 * j + 1], dz = zi - s_x[3 * j + 2];
            // uncontracted, in numpy's order: ((dx^2 + dy^2) + dz^2)
            const double d = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
            const double lim = a.factor * (ri + s_r[j]);
            if (d < lim) bits |= 1u << b;
            if (j > i && fabs(d - lim) <= FC_NEAR_EPS) ++near;
        }
        if (adj) adj[item] = bits;
        if (expected && a.delta_out) {
            unsigned diff = bits ^ expected[item];
            // every differing bond is seen from both of its atoms: count it at the smaller index only
            for (; diff; diff &= diff - 1u) {
                const int j = 32 * w + __ffs(diff) - 1;
                if (j <= i) continue;
                if (a.excluded && (a.excluded[i] || a.excluded[j])) continue;
                ++delta;
            }
        }
    }
    __shared__ int s_red[2];
    if (threadIdx.x < 2) s_red[threadIdx.x] = 0;
    __syncthreads();
    for (int o = 16; o > 0; o >>= 1) {
        delta += __shfl_xor_sync(0xffffffffu, delta, o);
        near += __shfl_xor_sync(0xffffffffu, near, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (delta) atomicAdd(&s_red[0], delta);
        if (near) atomicAdd(&s_red[1], near);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (a.delta_out) a.delta_out[p] = s_red[0];
        if (a.near_out) a.near_out[p] = s_red[1];
    }
}

}  // namespace fc

using namespace fc;

static int bond_run(const double* coords, int64_t n_struct, int32_t n_atoms, const double* radii, double factor,
                    const uint32_t* expected, int32_t expected_per_structure, const uint8_t* excluded, uint32_t* adj_out,
                    int32_t* delta_out, int32_t* near_out) {
    FC_REQUIRE(n_struct >= 0 && n_atoms > 0 && n_struct < ((int64_t)1 << 31), "fc_bond: bad sizes");
    if (n_struct == 0) return FC_OK;
    FC_REQUIRE(coords && radii, "fc_bond: null pointer");
    FC_REQUIRE(factor > 0.0, "fc_bond: the bond factor must be positive");
    const size_t smem = (size_t)n_atoms * 4 * sizeof(double);
    FC_REQUIRE(smem <= 200 * 1024, "fc_bond: %d atoms per structure exceed the shared-memory staging (6400)", n_atoms);
    const int words = (n_atoms + 31) / 32;
    cudaStream_t s;
    FC_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    int rc = FC_OK;
    {
        DevBuf<double> d_x, d_r;
        DevBuf<unsigned> d_exp, d_adj;
        DevBuf<unsigned char> d_excl;
        DevBuf<int> d_delta, d_near;
        cudaError_t e = d_x.alloc((size_t)n_struct * n_atoms * 3, s);
        const size_t exp_words = (size_t)(expected_per_structure ? n_struct : 1) * n_atoms * words;
#define BR(call) do { if (e == cudaSuccess) e = (call); } while (0)
        BR(d_r.alloc((size_t)n_atoms, s));
        BR(cudaMemcpyAsync(d_x.p, coords, (size_t)n_struct * n_atoms * 24, cudaMemcpyHostToDevice, s));
        BR(cudaMemcpyAsync(d_r.p, radii, (size_t)n_atoms * 8, cudaMemcpyHostToDevice, s));
        if (expected) {
            BR(d_exp.alloc(exp_words, s));
            BR(cudaMemcpyAsync(d_exp.p, expected, exp_words * 4, cudaMemcpyHostToDevice, s));
        }
        if (excluded) {
            BR(d_excl.alloc((size_t)n_atoms, s));
            BR(cudaMemcpyAsync(d_excl.p, excluded, (size_t)n_atoms, cudaMemcpyHostToDevice, s));
        }
        if (adj_out) BR(d_adj.alloc((size_t)n_struct * n_atoms * words, s));
        if (delta_out) BR(d_delta.alloc((size_t)n_struct, s));
        if (near_out) BR(d_near.alloc((size_t)n_struct, s));
        BR(cudaFuncSetAttribute(bond_graph_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (e == cudaSuccess) {
            BondArgs a{d_x.p, d_r.p, factor, n_atoms, words, expected ? d_exp.p : nullptr, expected_per_structure,
                       excluded ? d_excl.p : nullptr, adj_out ? d_adj.p : nullptr, delta_out ? d_delta.p : nullptr,
                       near_out ? d_near.p : nullptr};
            bond_graph_kernel<<<(unsigned)n_struct, 256, smem, s>>>(a);
            e = cudaGetLastError();
        }
        if (adj_out) BR(cudaMemcpyAsync(adj_out, d_adj.p, (size_t)n_struct * n_atoms * words * 4, cudaMemcpyDeviceToHost, s));
        if (delta_out) BR(cudaMemcpyAsync(delta_out, d_delta.p, (size_t)n_struct * 4, cudaMemcpyDeviceToHost, s));
        if (near_out) BR(cudaMemcpyAsync(near_out, d_near.p, (size_t)n_struct * 4, cudaMemcpyDeviceToHost, s));
        BR(cudaStreamSynchronize(s));
#undef BR
        if (e != cudaSuccess) rc = cuda_fail(e, "fc_bond", __FILE__, __LINE__);
    }
    cudaStreamSynchronize(s);
    cudaStreamDestroy(s);
    return rc;
}

extern "C" int fc_bond_graph_batch(const double* coords, int64_t n_struct, int32_t n_atoms, const double* radii,
                                   double factor, uint32_t* adj_out, int32_t* near_out) {
    FC_REQUIRE(adj_out || n_struct == 0, "fc_bond_graph_batch: null output");
    return bond_run(coords, n_struct, n_atoms, radii, factor, nullptr, 0, nullptr, adj_out, nullptr, near_out);
}

extern "C" int fc_bond_delta_batch(const double* coords, int64_t n_struct, int32_t n_atoms, const double* radii,
                                   double factor, const uint32_t* expected, int32_t expected_per_structure,
                                   const uint8_t* excluded, int32_t* delta_out, int32_t* near_out) {
    FC_REQUIRE((expected && delta_out) || n_struct == 0, "fc_bond_delta_batch: null pointer");
    return bond_run(coords, n_struct, n_atoms, radii, factor, expected, expected_per_structure, excluded, nullptr, delta_out,
                    near_out);
}
