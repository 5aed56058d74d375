// firecode_b200 -- tensor-core screen of the similarity pruning (tcgen05 / TMEM, sm_100a).
//
// The pair covariance of the Kabsch RMSD,  H(i, j)[a][b] = sum_k x_i[k][a] * x_j[k][b]  (centred heavy
// atoms; what prism_pruner.rmsd.rmsd_and_max contracts per pair, call sites
// /root/reference/firecode/embedder.py:1472, ensemble.py:230), over all pairs of a chunk IS a Gram
// matrix: rows (structure, component), K = atoms.  gram_tc_kernel computes it on the 5th-generation
// tensor cores in TF32 with FP32 accumulation in tensor memory and turns each 3x3 block into a
// screening decision in the epilogue; pairs it cannot rule out go to the FP64 exact kernel, so the
// kept set does not depend on TF32 rounding (the screen widens its band by a rigorous bound on it).
//
// Operand image (written per pass by gram_pack_kernel, read by 1-D bulk async copies): the active
// structures of the pass sit at "padded positions" (every chunk starts at a multiple of 16); 8
// consecutive positions form a group stored as [component a][k-core c][row r][4 floats] = the
// no-swizzle K-major core-matrix layout of a UMMA shared-memory descriptor (core matrix = 8 rows x
// 16 bytes).  The same bytes serve as operand A (M = 128 positions of one component: 16 groups,
// stride 3 * kc * 128 B) and as operand B (N = 48 = 16 positions x 3 components: 6 row groups,
// stride kc * 128 B), so a tile is one contiguous copy.
//
// CTA = 736 threads, one per SM, persistent over work items (128-row block x range of 16-column
// tiles): warp 0 = copy producer (A once per item, B tiles through a 4-stage ring), warps 1-6 = MMA
// issuers (tile parity x component: kc / 2 k-steps of tcgen05.mma M128 N48 K16 (FP16) per tile, three
// accumulator stages in TMEM), warps 7-22 = epilogue (tcgen05.ld of 4 pairs' 3x3 blocks per thread,
// Frobenius bound, cofactor bound, and a shared-memory queue that finishes the few remaining pairs with
// a Newton iteration on the QCP quartic, 32 at a time).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fc {

struct GramWork {
    int row0;         // first padded position of the 128-row block (multiple of 8)
    int col_tile0;    // first column tile; tile t covers positions [16 t, 16 t + 16)
    int n_col_tiles;
    int pend;         // end of the chunk (padded position): a pair needs prow < pcol < pend
};

struct GramArgs {
    const float* img;         // operand image, (n_pos / 8) groups x 3 x kc x 8 x 4 floats
    const float* gp;          // (n_pos) sum |x|^2 of the structure at the position, 0 for padding
    const int* spos;          // (n_pos) structure index at the position, -1 for padding
    const double* energies;   // (n structures) or null
    double max_dE;
    const GramWork* work;
    int n_work;
    int kc;                   // 16-byte k-cores per row (even, <= kGramMaxKc): atoms padded to 8 * kc (FP16) / 4 * kc (TF32)
    int tf32;                 // 0: operands are FP16 (default: same 11-bit significand as TF32, K = 16 per instruction,
                              // half the operand bytes and half the tcgen05.mma count), 1: TF32 (FC_PRUNE_TF32=1)
    float thr_e;              // (max_rmsd + band)^2 * nh
    float e0_scale;           // 1 - sqrt(3) * (bound on the relative TF32 product error)
    int2* cand;               // pairs the screen could not rule out (structure indices, x < y)
    unsigned long long* n_cand;
    long long cand_cap;
    float* dump;              // debug: H of every (prow, pcol) visited, [prow][dump_ld][9]; null in production
    int dump_ld;
    const float4* tile_norm;  // per 16-position tile the ranges {lo1, hi1, lo2, hi2}, {lo3, hi3, -, -} of the three singular
                              // values of its structures' centred coordinates ({+inf, -inf}: no structure), or null; with
                              // it, a (row block, column tile) whose ranges are apart by more than cull_gap2 in squared
                              // distance is skipped by all three roles: E >= sum_k (sigma_k(P) - sigma_k(Q))^2 (von Neumann)
    float cull_gap2;          // the (banded) threshold on the summed squared deviation, with a safety factor
    unsigned long long* tiles_done;  // += column tiles actually multiplied (statistics; may be null)
    int* error;               // set when a barrier wait times out (the kernel then runs out, the host reports it)
    long long* prof;          // debug: cycle counters of CTA 0 (see tools/gram_tc_test.cu); null in production
    int no_math;              // debug bits (tools/gram_tc_test.cu; 0 in production): 1 = the epilogue only drains tensor
                              // memory, 2 = one k-step per tile, 4 = no B copies, 64 = every valid pair takes pass 2
};

constexpr int kGramEpiWarps = 16;
constexpr int kGramIssuers = 6;     // MMA-issuing warps: (tile parity, component)
constexpr int kGramThreads = 32 * (1 + kGramIssuers + kGramEpiWarps);  // producer, MMA issuers, epilogue warps
constexpr int kGramBStages = 4;
constexpr int kGramDStages = 3;     // accumulator stages in tensor memory: 3 x (3 components x 48 columns) = 432 of 512 columns
constexpr unsigned kGramDStageCols = 160, kGramDCompCols = 48;
constexpr int kGramMaxKc = 22;   // <= 176 selected atoms as FP16 (88 as TF32): operands + pair queues must fit the 227 KB of shared memory
constexpr unsigned kGramWaitHintNs = 20000;  // suspend-time hint of the mbarrier waits (the phase completing ends the sleep)
constexpr int kGramQueue = 32;   // entries of an epilogue warp's queue of pairs that need the Newton iteration
// Relative bound on |H_tf32 - H|_F / (|p| |q|): operands are rounded to nearest TF32 (2^-11 each, so 2^-10 on a
// product, Cauchy-Schwarz over the atoms); doubled to cover the accumulation inside the tensor core.  Measured on
// random ensembles: 1.1e-4 (tools/gram_tc_test.cu).
constexpr float kGramTf32Eps = 1.0f / 512.0f;

__host__ __device__ inline size_t gram_smem_bytes(int kc) {
    return (size_t)6144 * kc + (size_t)kGramBStages * 768 * kc + 256 + (size_t)kGramEpiWarps * 6 * kGramQueue * sizeof(float);
}

// ---- PTX plumbing -----------------------------------------------------------------------------------
namespace gram {
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_expect(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_load(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// Bounded mbarrier wait.  A protocol error must neither hang the GPU nor poison the CUDA context: after a short burst of
// polls the wait looks at the clock (%globaltimer, two seconds: a TIME bound, so time-slicing, MPS or a debugger cannot
// expire it spuriously) and at the abort flag; on expiry it records `code` in *error and gives up.  Every other wait of
// the kernel then falls through as soon as it sees the flag, the kernel runs to its end with garbage that nobody reads,
// and the host turns the flag into an error code (fc_prune: "gram_tc_kernel: mbarrier wait N timed out").
// The wait itself is mbarrier.try_wait WITH a suspend-time hint: the thread sleeps in hardware until the phase completes
// (or the hint expires) instead of re-issuing polls.  ncu of the poll-loop version showed the seven single-thread
// producer / issuer warps executing a third of all warp instructions of the kernel in their wait loops, competing with
// the sixteen epilogue warps for issue slots ("not selected" was the top stall of the epilogue).
__device__ __forceinline__ void bar_wait(unsigned bar, unsigned parity, int* error, int code) {
    unsigned long long t0 = 0;
    for (unsigned spin = 0;; ++spin) {
        unsigned ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"(kGramWaitHintNs)
            : "memory");
        if (ok) return;
        if ((spin & 0x3Fu) == 0x3Fu) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            const bool aborted = error && *reinterpret_cast<volatile int*>(error) != 0;
            if (aborted || now - t0 > 2000000000ull) {
                if (error && !aborted) atomicCAS(error, 0, code);
                return;
            }
        }
    }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// no-swizzle K-major shared-memory matrix descriptor: lbo = byte distance of the two k-cores of one MMA,
// sbo = byte distance of consecutive 8-row groups
__device__ __forceinline__ uint64_t smem_desc(unsigned addr, unsigned lbo, unsigned sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// D[tmem] (+)= A[smem] * B[smem]^T, TF32 inputs, FP32 accumulate, M = 128, N from the instruction descriptor
__device__ __forceinline__ void mma_tf32(unsigned d_tmem, uint64_t a_desc, uint64_t b_desc, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the same with FP16 inputs (K = 16 per instruction)
__device__ __forceinline__ void mma_f16(unsigned d_tmem, uint64_t a_desc, uint64_t b_desc, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(unsigned taddr, float* v) {
    unsigned r0, r1, r2, r3, r4, r5, r6, r7;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
                 : "r"(taddr));
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
    v[4] = __uint_as_float(r4); v[5] = __uint_as_float(r5); v[6] = __uint_as_float(r6); v[7] = __uint_as_float(r7);
}
__device__ __forceinline__ void tmem_ld4(unsigned taddr, float* v) {
    unsigned r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(taddr));
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace gram

// ---- operand image --------------------------------------------------------------------------------
// one block per group of 8 positions; xcf = centred heavy-atom coordinates (n, nh) as float4
__global__ void __launch_bounds__(256) gram_pack_kernel(const float4* __restrict__ xcf, const double* __restrict__ g,
                                                        const int* __restrict__ spos, int nh, int kc, int n_groups, int tf32,
                                                        float* __restrict__ img, float* __restrict__ gp) {
    const int grp = blockIdx.x;
    if (grp >= n_groups) return;
    __shared__ int s_idx[8];
    if (threadIdx.x < 8) {
        const int s = spos[grp * 8 + threadIdx.x];
        s_idx[threadIdx.x] = s;
        gp[grp * 8 + threadIdx.x] = s >= 0 ? (float)g[s] : 0.f;
    }
    __syncthreads();
    const int per_group = 3 * kc * 32;  // 32-bit slots: one TF32 value or two FP16 values each
    float* dst = img + (size_t)grp * per_group;
    for (int e = threadIdx.x; e < per_group; e += blockDim.x) {
        const int t = e & 3, r = (e >> 2) & 7, c = (e >> 5) % kc, comp = (e >> 5) / kc;
        const int s = s_idx[r];
        if (tf32) {
            const int k = 4 * c + t;
            float v = 0.f;
            if (s >= 0 && k < nh) {
                const float4 x = xcf[(size_t)s * nh + k];
                v = comp == 0 ? x.x : (comp == 1 ? x.y : x.z);
            }
            unsigned bits;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(bits) : "f"(v));
            dst[e] = __uint_as_float(bits);
        } else {
            const int k = 8 * c + 2 * t;  // atoms k, k + 1 of the 8 this k-core row holds
            float v0 = 0.f, v1 = 0.f;
            if (s >= 0 && k < nh) {
                const float4 x = xcf[(size_t)s * nh + k];
                v0 = comp == 0 ? x.x : (comp == 1 ? x.y : x.z);
            }
            if (s >= 0 && k + 1 < nh) {
                const float4 x = xcf[(size_t)s * nh + k + 1];
                v1 = comp == 0 ? x.x : (comp == 1 ? x.y : x.z);
            }
            const __half2 h = __floats2half2_rn(v0, v1);  // low half = atom k (lower address)
            dst[e] = __uint_as_float(*reinterpret_cast<const unsigned*>(&h));
        }
    }
}

// ---- tile culling -----------------------------------------------------------------------------------
// For centred coordinate matrices P, Q (atoms x 3) the best superposition gives
//   E = |P|^2 + |Q|^2 - 2 max_R tr(R P^T Q)  >=  sum_k (sigma_k(P) - sigma_k(Q))^2
// (von Neumann's trace inequality; sigma_k = singular values, descending), so two structures whose shape numbers differ
// by more than the threshold cannot be similar whatever their orientation.  The positions of a segment are sorted by
// sigma_1; per 16-position tile the ranges of the three sigma_k; a (row block, column tile) pair whose ranges are too
// far apart is skipped before anything is copied or multiplied.
struct SigRange { float lo[3], hi[3]; };

__global__ void gram_tile_norm_kernel(const float4* __restrict__ sig, const int* __restrict__ spos, int n_tiles,
                                      float4* __restrict__ tile_norm) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const float inf = __int_as_float(0x7f800000);
    float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
    for (int k = 0; k < 16; ++k) {
        const int sidx = spos[16 * t + k];
        if (sidx >= 0) {
            const float4 v = sig[sidx];
            lo[0] = fminf(lo[0], v.x); hi[0] = fmaxf(hi[0], v.x);
            lo[1] = fminf(lo[1], v.y); hi[1] = fmaxf(hi[1], v.y);
            lo[2] = fminf(lo[2], v.z); hi[2] = fmaxf(hi[2], v.z);
        }
    }
    tile_norm[2 * t] = make_float4(lo[0], hi[0], lo[1], hi[1]);
    tile_norm[2 * t + 1] = make_float4(lo[2], hi[2], 0.f, 0.f);
}

// sort keys of the positions: (segment << 16) | the top 16 bits of sigma_1 (non-negative floats order like their bits; the
// order inside a segment only has to be roughly by sigma_1: any order is valid, a rough one widens a tile's range by
// less than 1 %); padding sorts behind the structures of its segment
__global__ void gram_sort_key_kernel(const int* __restrict__ spos, const int* __restrict__ seg, const float4* __restrict__ sig,
                                     int n_pos, unsigned long long* __restrict__ keys) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_pos) return;
    const int sidx = spos[q];
    const unsigned low = sidx >= 0 ? (__float_as_uint(sig[sidx].x) >> 16) : 0xffffu;
    keys[q] = ((unsigned long long)(unsigned)seg[q] << 16) | low;
}

namespace gram {
// ranges of the 128-row block starting at position row0 (8 tiles)
__device__ __forceinline__ SigRange block_norm_range(const float4* __restrict__ tile_norm, int row0) {
    const float inf = __int_as_float(0x7f800000);
    SigRange r{{inf, inf, inf}, {-inf, -inf, -inf}};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float4 a = tile_norm[2 * ((row0 >> 4) + k)], b = tile_norm[2 * ((row0 >> 4) + k) + 1];
        r.lo[0] = fminf(r.lo[0], a.x); r.hi[0] = fmaxf(r.hi[0], a.y);
        r.lo[1] = fminf(r.lo[1], a.z); r.hi[1] = fmaxf(r.hi[1], a.w);
        r.lo[2] = fminf(r.lo[2], b.x); r.hi[2] = fmaxf(r.hi[2], b.y);
    }
    return r;
}
// true: no pair of (row block with ranges `rows`, column tile `ct`) can be similar.  Evaluated identically by the copy
// producer, the MMA issuers and the epilogue warps, which therefore agree on the sequence of tiles.  An empty range
// ({+inf, -inf}) gives an infinite gap: culled; inf - inf does not occur.
__device__ __forceinline__ bool tile_culled(const float4* __restrict__ tile_norm, const SigRange& rows, int ct, float cull_gap2) {
    const float4 a = tile_norm[2 * ct], b = tile_norm[2 * ct + 1];
    const float g0 = fmaxf(0.f, fmaxf(a.x - rows.hi[0], rows.lo[0] - a.y));
    const float g1 = fmaxf(0.f, fmaxf(a.z - rows.hi[1], rows.lo[1] - a.w));
    const float g2 = fmaxf(0.f, fmaxf(b.x - rows.hi[2], rows.lo[2] - b.y));
    return fmaf(g0, g0, fmaf(g1, g1, g2 * g2)) > cull_gap2;
}
}  // namespace gram

// ---- the Gram screen --------------------------------------------------------------------------------
// Finishes the queued pairs of one warp, one lane per pair.  Entry = {u, f2, cc, det, row structure, column structure}.
// Newton from above on  P(x) = (x^2 - f2)^2 - 8 det x - 4 cc  (the QCP characteristic polynomial, largest root = sum of the
// singular values with the sign of det on the smallest): every iterate stays above the root, so each is a valid upper
// bound; a pair whose bound never rules it out becomes a candidate for the FP64 kernel.
__device__ __forceinline__ void gram_newton_flush(const GramArgs& a, const float* queue, int qn, int lane) {
    __syncwarp();
    if (lane < qn) {
        const float* e = queue + 6 * lane;
        const float u = e[0], f2 = e[1], cc = e[2], det = e[3];
        float lam = sqrtf(f2 + 2.0f * sqrtf(3.0f * cc)) * 1.000001f;  // the pass-2 bound: the iteration starts above the root
        bool reject = u - 2.0f * lam > 0.f;
        for (int it = 0; it < 8 && !reject; ++it) {
            const float tt = lam * lam - f2;
            const float pv = tt * tt - 8.0f * det * lam - 4.0f * cc, dp = 4.0f * lam * tt - 8.0f * det;
            if (!(dp > 0.f) || !(pv > 0.f)) break;
            const float step = __fdividef(pv, dp);
            lam -= step;
            reject = u - 2.0f * lam > 0.f;
            if (step <= 1e-5f * lam) break;
        }
        if (!reject) {
            const unsigned long long slot = atomicAdd(a.n_cand, 1ull);
            // (structure indices ascending: positions are ordered by norm inside a segment, not by index)
            const int s1 = __float_as_int(e[4]), s2 = __float_as_int(e[5]);
            if ((long long)slot < a.cand_cap) a.cand[slot] = make_int2(min(s1, s2), max(s1, s2));
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(kGramThreads, 1) gram_tc_kernel(GramArgs a) {
    using namespace gram;
    extern __shared__ __align__(128) uint8_t gram_smem[];
    const int kc = a.kc;
    const unsigned a_bytes = 6144u * kc, b_bytes = 768u * kc;
    uint8_t* sA = gram_smem;
    uint8_t* sB = gram_smem + a_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + kGramBStages * b_bytes);
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 16);
    const unsigned bar0 = smem_addr(bars);
    // barrier ids
    const unsigned A_FULL = bar0, A_EMPTY = bar0 + 8;
    auto B_FULL = [&](unsigned s) { return bar0 + 16 + 8 * s; };
    auto B_EMPTY = [&](unsigned s) { return bar0 + 48 + 8 * s; };
    auto D_FULL = [&](unsigned s) { return bar0 + 80 + 8 * s; };    // 80, 88, 96
    auto D_EMPTY = [&](unsigned s) { return bar0 + 104 + 8 * s; };  // 104, 112, 120 (16 barrier slots in all)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        bar_init(A_FULL, 1);
        bar_init(A_EMPTY, kGramIssuers);
        for (unsigned s = 0; s < kGramBStages; ++s) { bar_init(B_FULL(s), 1); bar_init(B_EMPTY(s), 3); }
        for (unsigned s = 0; s < kGramDStages; ++s) { bar_init(D_FULL(s), 3); bar_init(D_EMPTY(s), kGramEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // 512 columns of tensor memory: three accumulator stages x 3 components x 48 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmem_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem_base = *tmem_slot;
    const size_t group_floats = (size_t)3 * kc * 32;

    if (warp == 0) {
        if (lane == 0) {  // ---- copy producer ----
            unsigned bstage = 0, bphase = 0, aphase = 0;
            unsigned long long n_done = 0;
            for (int w = blockIdx.x; w < a.n_work; w += gridDim.x) {
                const GramWork wk = a.work[w];
                bar_wait(A_EMPTY, aphase ^ 1, a.error, 1);
                bar_expect(A_FULL, a_bytes);
                const float* src = a.img + (size_t)(wk.row0 >> 3) * group_floats;
                for (unsigned g = 0; g < 16; ++g)
                    bulk_load(smem_addr(sA) + g * (a_bytes / 16), src + g * group_floats, a_bytes / 16, A_FULL);
                aphase ^= 1;
                const SigRange rows = a.tile_norm ? block_norm_range(a.tile_norm, wk.row0) : SigRange{};
                for (int t = 0; t < wk.n_col_tiles; ++t) {
                    if (a.tile_norm && tile_culled(a.tile_norm, rows, wk.col_tile0 + t, a.cull_gap2)) continue;
                    ++n_done;
                    bar_wait(B_EMPTY(bstage), bphase ^ 1, a.error, 2);
                    if (a.no_math & 4) {
                        bar_arrive(B_FULL(bstage));
                    } else {
                    bar_expect(B_FULL(bstage), b_bytes);
                    bulk_load(smem_addr(sB) + bstage * b_bytes, a.img + (size_t)(2 * (wk.col_tile0 + t)) * group_floats, b_bytes,
                              B_FULL(bstage));
                    }
                    if (++bstage == kGramBStages) { bstage = 0; bphase ^= 1; }
                }
            }
            if (a.tiles_done && n_done) atomicAdd(a.tiles_done, n_done);
        }
        __syncwarp();
    } else if (warp <= kGramIssuers) {
        if (lane == 0) {  // ---- MMA issuers ----
            // Six issuing threads: a tcgen05.mma costs its issuing thread 65-130 cycles whatever N is (the operands go
            // through uniform registers; measured, tools/gram_tc_test.cu), while the tensor pipe needs ~24 cycles for this
            // shape, so one issuer caps the kernel at a fifth of what the operand fetch allows.
            // instruction descriptor: D = F32 (bit 4), A = B = TF32 (bits 7, 10), both K-major, N = 48 (>> 3 at bit 17),
            // M = 128 (>> 4 at bit 24)
            // Tile g (counted over the whole CTA) uses operand slot g % 4 and accumulator stage g % 3 (its k-th use has
            // barrier parity k & 1, k = g / 3); issuer (p, comp) issues component comp of the tiles with g % 2 == p.  Three
            // stages let the MMAs run two tiles ahead of the slowest epilogue warp (with two, the sixteen epilogue warps
            // spent a quarter of their time waiting for a stage that the slowest of them had not released yet).
            const unsigned stage = (unsigned)(warp - 1) / 3u, comp = (unsigned)(warp - 1) % 3u;  // stage = the parity p
            // (FP16 operands: format code 0 in both fields, K = 16 per instruction -- the same 256 bytes per k-step)
            const unsigned fmt = a.tf32 ? 2u : 0u;
            const unsigned idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (6u << 17) | (8u << 24);
            unsigned g = 0, bphases = 0, aphase = 0;  // one phase bit per operand slot
            const unsigned sA_addr = smem_addr(sA) + comp * kc * 128, sB_addr = smem_addr(sB);
            for (int w = blockIdx.x; w < a.n_work; w += gridDim.x) {
                const GramWork wk = a.work[w];
                bar_wait(A_FULL, aphase, a.error, 3);
                aphase ^= 1;
                const SigRange rows = a.tile_norm ? block_norm_range(a.tile_norm, wk.row0) : SigRange{};
                for (int t = 0; t < wk.n_col_tiles; ++t) {
                    if (a.tile_norm && tile_culled(a.tile_norm, rows, wk.col_tile0 + t, a.cull_gap2)) continue;
                    const unsigned g_now = g++;
                    if ((g_now & 1u) != stage) continue;
                    const unsigned bs = g_now % kGramBStages;
                    const long long c0 = a.prof ? clock64() : 0;
                    bar_wait(B_FULL(bs), (bphases >> bs) & 1u, a.error, 4);
                    const long long c1 = a.prof ? clock64() : 0;
                    const unsigned dst = g_now % kGramDStages, dpar = (g_now / kGramDStages) & 1u;
                    const unsigned d_tmem = tmem_base + dst * kGramDStageCols + comp * kGramDCompCols;
                    bar_wait(D_EMPTY(dst), dpar ^ 1u, a.error, 5);
                    const long long c2 = a.prof ? clock64() : 0;
                    tc_fence_after();
                    const int ksteps = (a.no_math & 2) ? 1 : kc / 2;
                    uint64_t da = smem_desc(sA_addr, 128, 3 * kc * 128), db = smem_desc(sB_addr + bs * b_bytes, 128, kc * 128);
                    for (int k = 0; k < ksteps; ++k) {
                        if (a.tf32) mma_tf32(d_tmem, da, db, idesc, k > 0 ? 1u : 0u);
                        else mma_f16(d_tmem, da, db, idesc, k > 0 ? 1u : 0u);
                        da += 16;  // next k-step: 256 bytes further, in the 16-byte units of the start-address field
                        db += 16;
                    }
                    tc_commit(B_EMPTY(bs));
                    tc_commit(D_FULL(dst));
                    bphases ^= 1u << bs;
                    if (a.prof && blockIdx.x == 0 && warp == 1) {
                        const long long c3 = clock64();
                        a.prof[0] += c1 - c0; a.prof[1] += c2 - c1; a.prof[2] += c3 - c2; a.prof[3] += 1;
                    }
                }
                tc_commit(A_EMPTY);
            }
        }
        __syncwarp();
    } else {
        // ---- epilogue: thread = one row of the block (TMEM lane), 4 of the 16 columns of the tile ----
        // 16 warps: four per scheduler, so that the dependent FP32 chains of one warp are covered by the others
        const int quarter = warp & 3;          // TMEM lanes 32 * (warp % 4) ... are the ones this warp may read
        const int part = (warp - 1 - kGramIssuers) >> 2;      // which 4 of the 16 column positions
        const int row_in_block = quarter * 32 + lane;
        float* queue = reinterpret_cast<float*>(tmem_slot + 4) + (warp - 1 - kGramIssuers) * (6 * kGramQueue);  // this warp's pair queue
        int qn = 0;                                                                              // warp-uniform fill level
        unsigned ge = 0;  // tiles seen so far: stage ge % 3, parity (ge / 3) & 1
        // accumulator column of (column position 8 g + r, component cb) = 24 g + 8 cb + r
        const unsigned tcol = (unsigned)(24 * (part >> 1) + 4 * (part & 1));
        for (int w = blockIdx.x; w < a.n_work; w += gridDim.x) {
            const GramWork wk = a.work[w];
            const int prow = wk.row0 + row_in_block;
            const int s_row = prow < wk.pend ? a.spos[prow] : -1;
            const float g_row = prow < wk.pend ? a.gp[prow] : 0.f;
            const double e_row = (a.energies && s_row >= 0) ? a.energies[s_row] : 0.0;
            const SigRange rows = a.tile_norm ? block_norm_range(a.tile_norm, wk.row0) : SigRange{};
            for (int t = 0; t < wk.n_col_tiles; ++t) {
                if (a.tile_norm && tile_culled(a.tile_norm, rows, wk.col_tile0 + t, a.cull_gap2)) continue;  // warp-uniform
                __syncwarp();  // the tensor-memory loads below are warp-collective
                // column metadata first: these loads fly while the accumulator is still being produced
                const int pcol0 = 16 * (wk.col_tile0 + t) + 4 * part;
                const int4 sc = *reinterpret_cast<const int4*>(a.spos + pcol0);
                const float4 gc = *reinterpret_cast<const float4*>(a.gp + pcol0);
                const long long e0 = a.prof ? clock64() : 0;
                const unsigned dstage = ge % kGramDStages, dphase = (ge / kGramDStages) & 1u;
                ++ge;
                bar_wait(D_FULL(dstage), dphase, a.error, 6);
                const long long e1 = a.prof ? clock64() : 0;
                tc_fence_after();
                float h[3][3][4];  // [component of the row structure][component of the column structure][column]
                const unsigned taddr = tmem_base + ((unsigned)(quarter * 32) << 16) + dstage * kGramDStageCols + tcol;
#pragma unroll
                for (int ca = 0; ca < 3; ++ca)
#pragma unroll
                    for (int cb = 0; cb < 3; ++cb) tmem_ld4(taddr + ca * kGramDCompCols + cb * 8, h[ca][cb]);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) bar_arrive(D_EMPTY(dstage));
                if (a.prof && blockIdx.x == 0 && threadIdx.x == 32 * (1 + kGramIssuers)) {
                    const long long e2 = clock64();
                    a.prof[4] += e1 - e0; a.prof[5] += e2 - e1; a.prof[6] += 1;
                }

                if (a.dump) {
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int ca = 0; ca < 3; ++ca)
#pragma unroll
                            for (int cb = 0; cb < 3; ++cb)
                                a.dump[((size_t)prow * a.dump_ld + pcol0 + r) * 9 + 3 * ca + cb] = h[ca][cb][r];
                }
                const bool live = !((a.no_math & 1) || s_row < 0 || pcol0 + 3 <= prow || pcol0 >= wk.pend);
                const int s_col[4] = {sc.x, sc.y, sc.z, sc.w};
                const float g_col[4] = {gc.x, gc.y, gc.z, gc.w};
                // Lower bound on the true sum of squared deviations: E >= e0 - 2 S, S = sum of the singular values of H (largest
                // root of the quartic in gram_newton_flush); the TF32 rounding of H moves S by at most
                // sqrt(3) eps |p| |q| <= sqrt(3) eps e0 / 2, which e0_scale takes off e0.  A pair is ruled out as soon as an UPPER
                // bound on S gives  u - 2 S > 0.
                // Pass 1, branch-free over the columns: S <= sqrt(3) |H|_F, i.e. u >= 0 and u^2 >= 12 |H|_F^2.
                unsigned und = 0;
                float uu[4], ff[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int pcol = pcol0 + r;
                    const bool valid = live && prow < pcol && pcol < wk.pend && s_col[r] >= 0;
                    uu[r] = (g_row + g_col[r]) * a.e0_scale - a.thr_e;
                    const float fa = h[0][0][r] * h[0][0][r] + h[0][1][r] * h[0][1][r] + h[0][2][r] * h[0][2][r];
                    const float fb = h[1][0][r] * h[1][0][r] + h[1][1][r] * h[1][1][r] + h[1][2][r] * h[1][2][r];
                    const float fc2 = h[2][0][r] * h[2][0][r] + h[2][1][r] * h[2][1][r] + h[2][2][r] * h[2][2][r];
                    ff[r] = fa + fb + fc2;
                    if (valid && (!(uu[r] >= 0.f && uu[r] * uu[r] >= 12.0f * ff[r]) || (a.no_math & 64))) und |= 1u << r;
                }
                if (!__any_sync(0xffffffffu, und != 0)) continue;
                // Pass 2 (about a third of the pairs of a conformer ensemble get here), again branch-free over the columns:
                // S^2 = f2 + 2 e2 with e2 = sum of the pairwise products of the singular values <= sqrt(3 cc), cc = |cof H|_F^2.
                // Without square roots:  u - 2 sqrt(f2 + 2 sqrt(3 cc)) > 0  <=>  u > 0,  v = u^2 - 4 f2 > 0,  v^2 > 192 cc.
                float ccs[4], dets[4];
                unsigned surv = 0;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float h0 = h[0][0][r], h1 = h[0][1][r], h2 = h[0][2][r], h3 = h[1][0][r], h4 = h[1][1][r],
                                h5 = h[1][2][r], h6 = h[2][0][r], h7 = h[2][1][r], h8 = h[2][2][r];
                    const float c0 = h4 * h8 - h5 * h7, c1 = h5 * h6 - h3 * h8, c2 = h3 * h7 - h4 * h6;
                    const float c3 = h2 * h7 - h1 * h8, c4 = h0 * h8 - h2 * h6, c5 = h1 * h6 - h0 * h7;
                    const float c6 = h1 * h5 - h2 * h4, c7 = h2 * h3 - h0 * h5, c8 = h0 * h4 - h1 * h3;
                    ccs[r] = (c0 * c0 + c1 * c1 + c2 * c2) + (c3 * c3 + c4 * c4 + c5 * c5) + (c6 * c6 + c7 * c7 + c8 * c8);
                    dets[r] = h0 * c0 + h1 * c1 + h2 * c2;
                    const float v = uu[r] * uu[r] - 4.0f * ff[r];
                    const bool reject = uu[r] > 0.f && v > 0.f && v * v > 192.0f * ccs[r];
                    if ((und & (1u << r)) && !reject) surv |= 1u << r;
                }
                if (!__any_sync(0xffffffffu, surv != 0)) continue;
                // The few pairs that survive this too are queued (per warp, shared memory) and finished 32 at a time by
                // gram_newton_flush, so the rare expensive path never runs with one active lane.
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    bool push = (surv >> r) & 1u;
                    if (push && a.energies && !(fabs(e_row - a.energies[s_col[r]]) < a.max_dE)) push = false;
                    const unsigned m = __ballot_sync(0xffffffffu, push);
                    if (m == 0) continue;
                    const int add = __popc(m);
                    if (qn + add > kGramQueue) {
                        gram_newton_flush(a, queue, qn, lane);
                        qn = 0;
                    }
                    if (push) {
                        float* e = queue + 6 * (qn + __popc(m & ((1u << lane) - 1u)));
                        e[0] = uu[r]; e[1] = ff[r]; e[2] = ccs[r]; e[3] = dets[r];
                        e[4] = __int_as_float(s_row); e[5] = __int_as_float(s_col[r]);
                    }
                    qn += add;
                    __syncwarp();
                }
            }
            if (qn) {
                gram_newton_flush(a, queue, qn, lane);
                qn = 0;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace fc
