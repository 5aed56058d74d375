// firecode_b200 -- compenetration (clash) screen for sm_100a.
//
// Replaces the per-pose call `compenetration_check(pose, ids=..., thresh=...)`
// (/root/reference/firecode/utils.py:507-575, called at embeds.py:139, 559, 718) by a batched
// screen over (conformer_a, conformer_b, rigid transform) poses.
//
// FP32 pass (clash_f32_kernel): the squared distance is evaluated in Gram form
//      |a - b'|^2 = (|a|^2 - 2 a.b') + |b'|^2
// so one atom pair costs 3 FMA lanes instead of the 6 FMA-pipe slots of the difference form.
//   * thread (pose_local, chunk) owns TB atoms of the transformed fragment B as 3*TB scalar
//     registers (ptxas feeds them to FFMA2 as 32-bit broadcast operands, `Rn.F32`) and TB running
//     minima;
//   * fragment A sits in shared memory as atom PAIRS {-2a0x,-2a1x,-2a0y,-2a1y | -2a0z,-2a1z,
//     |a0|^2,|a1|^2}: two warp-wide broadcast LDS.128 deliver the four FFMA2 operands of two atoms;
//   * per A pair and B atom: 3 FFMA2 + 2 FMNMX.  tools/microbench*.cu record why this shape was
//     chosen (FFMA2 register-operand limits, LDS.128 and FMNMX3 issue costs on B200).
// The FP32 result is trusted only outside a rigorous rounding band around thresh^2; poses inside
// the band (and every pose when max_clashes > 0 and a clash is possible) go to the FP64 recheck
// kernel, which restates the reference arithmetic (f64 transform, sqrt of the summed squares,
// `<` / `<=` compare, clash count).  Poses whose FP64 minimum distance lies within FC_NEAR_EPS of
// the threshold are flagged and listed.
#include <stdlib.h>

#include "fc_common.cuh"

namespace fc {

struct ClashGeom {
    int tb;       // B atoms per thread (even)
    int chunks;   // threads per pose
    int poses;    // poses per tile
    int threads;  // block size (multiple of 32)
};

static const int kTBs[] = {30, 25, 20, 16, 15, 12, 10, 8, 6, 4, 2};
// measured FMA-pipe efficiency of the main loop per TB (tools/microbench9.cu, B200)
static double tb_efficiency(int tb) {
    if (tb >= 30) return 0.76;
    if (tb >= 25) return 0.74;
    if (tb >= 20) return 0.71;
    if (tb >= 15) return 0.68;
    if (tb >= 10) return 0.62;
    if (tb >= 6) return 0.52;
    return 0.40;
}
static int tb_max_threads(int tb) {  // register budget: 65536 / (registers per thread)
    if (tb >= 25) return 256;
    if (tb >= 20) return 384;
    return 512;
}

static ClashGeom choose_geom(int n_b) {
    ClashGeom best{0, 0, 0, 0};
    double best_cost = 1e300;
    const char* force = getenv("FC_CLASH_TB");
    for (int tb : kTBs) {
        if (force && atoi(force) != tb) continue;
        int chunks = (n_b + tb - 1) / tb;
        if (chunks > tb_max_threads(tb)) continue;
        double cost = (double)chunks * tb / tb_efficiency(tb);
        if (cost < best_cost) {
            best_cost = cost;
            best.tb = tb;
            best.chunks = chunks;
        }
    }
    if (best.tb == 0) return best;
    // poses per tile: CTAs of ~4-5 warps, several per SM, as few idle lanes as possible
    int c = best.chunks;
    int t_max = tb_max_threads(best.tb);
    double best_score = -1.0;
    for (int p = 1; p * c <= t_max; ++p) {
        int used = p * c;
        int threads = (used + 31) / 32 * 32;
        if (threads > t_max) break;
        int ctas = t_max / threads;
        if (ctas > 8) ctas = 8;
        double util = (double)used / threads;
        double warps = (double)ctas * threads / 32.0;
        double fill = warps >= 8.0 ? 1.0 : warps / 8.0;
        // small penalty for very large CTAs (barrier + prologue are not overlapped inside one CTA)
        double size_pen = threads > 160 ? 0.97 : 1.0;
        // warps spread evenly over the four SM sub-partitions only in multiples of four
        double balance = ((int)warps % 4 == 0) ? 1.0 : 0.85;
        double score = util * fill * size_pen * balance;
        if (score > best_score + 1e-9) {
            best_score = score;
            best.poses = p;
            best.threads = threads;
        }
    }
    return best;
}

// ---------------------------------------------------------------------------------------------
// table preparation: f64 ensembles -> FP32 layouts used by the screen
// ---------------------------------------------------------------------------------------------
// a_tab: [conf][n_a_pad/2][2] float4 per atom pair {-2a0x,-2a1x,-2a0y,-2a1y},{-2a0z,-2a1z,|a0|^2,
//        |a1|^2}  (n_a_pad even; padding atoms: a' = 0, |a|^2 = 1e30)
// b_tab: [conf][n_b_pad]    float4  {x, y, z, 0}  (padding atoms carry w = 1e30)
// rad  : [conf] max |x| over the conformer (f32, rounded up)
__global__ void clash_prep_kernel(const double* __restrict__ coords, int n_conf, int n_atoms,
                                  int n_pad, int as_a, float4* __restrict__ tab,
                                  float* __restrict__ rad) {
    int conf = blockIdx.x;
    const double* src = coords + (size_t)conf * n_atoms * 3;
    float r2max = 0.f;
    for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
        if (i < n_atoms) {
            float x = (float)src[3 * i], y = (float)src[3 * i + 1], z = (float)src[3 * i + 2];
            float n2 = fmaf(x, x, fmaf(y, y, z * z));
            r2max = fmaxf(r2max, n2);
            if (as_a) {
                float* dst = reinterpret_cast<float*>(tab + ((size_t)conf * n_pad + (i & ~1))) + (i & 1);
                dst[0] = -2.f * x;
                dst[2] = -2.f * y;
                dst[4] = -2.f * z;
                dst[6] = n2;
            } else {
                tab[(size_t)conf * n_pad + i] = make_float4(x, y, z, 0.f);
            }
        } else {
            if (as_a) {
                float* dst = reinterpret_cast<float*>(tab + ((size_t)conf * n_pad + (i & ~1))) + (i & 1);
                dst[0] = 0.f;
                dst[2] = 0.f;
                dst[4] = 0.f;
                dst[6] = 1e30f;
            } else {
                tab[(size_t)conf * n_pad + i] = make_float4(0.f, 0.f, 0.f, 1e30f);
            }
        }
    }
    __shared__ float s_r[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r2max = fmaxf(r2max, __shfl_xor_sync(0xffffffffu, r2max, o));
    if ((threadIdx.x & 31) == 0) s_r[threadIdx.x >> 5] = r2max;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = 0.f;
        for (int w = 0; w < (blockDim.x + 31) / 32; ++w) m = fmaxf(m, s_r[w]);
        rad[conf] = sqrtf(m) * 1.000001f + 1e-6f;
    }
}

struct UncEntry {  // a pose the FP32 pass could not decide
    long long pose;
    int conf_a, conf_b;
};

struct ClashArgs {
    const float4* a_tab;
    const float* a_rad;
    const float4* b_tab;
    const float* b_rad;
    const double* xf;
    const int4* tiles;  // may be null: implicit single-group tiling
    long long n_tiles;
    long long n_poses;
    int n_a_pad, n_b_pad, chunks, poses;
    float thr2;
    int count_mode;  // max_clashes > 0: FP32 pass may only prove "no pair can clash"
    int use_tma;     // transforms of a tile arrive by cp.async.bulk (xf 16-byte aligned)
    uint8_t* status;
    float* min_dist;
    int* unc_count;
    UncEntry* unc_list;
};

template <int TB>
struct ClashLaunch {
    static constexpr int kMaxThreads = TB >= 25 ? 256 : (TB >= 20 ? 384 : 512);
};

// ---- TMA (bulk async copy) + mbarrier plumbing for the per-tile transform block ----------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tma_load_1d(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}

// tile -> (conformers, first pose, pose count)
__device__ __forceinline__ void clash_tile(const ClashArgs& p, long long tile, int& conf_a, int& conf_b, long long& first,
                                           int& count) {
    conf_a = 0;
    conf_b = 0;
    if (p.tiles) {
        int4 t = p.tiles[tile];
        conf_a = t.x;
        conf_b = t.y;
        first = t.z;
        count = t.w;
    } else {
        first = tile * p.poses;
        long long left = p.n_poses - first;
        count = left < p.poses ? (int)left : p.poses;
    }
}

// UNR: unroll factor of the loop over A atom pairs (ptxas orders the FFMA2s of one A pair component-major
// over all TB chains and fuses the two minima into FMNMX3 whatever the source order is)
template <int TB, int UNR, int PF>
__global__ void __launch_bounds__(ClashLaunch<TB>::kMaxThreads, 1) clash_f32_kernel(ClashArgs p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ulonglong2* sA = reinterpret_cast<ulonglong2*>(smem_raw);                  // n_a_pad entries
    float4* sB = reinterpret_cast<float4*>(smem_raw + (size_t)(p.n_a_pad + 2) * 16);  // n_b_pad entries
    size_t off = (size_t)(p.n_a_pad + 2) * 16 + (size_t)p.n_b_pad * 16;
    double* sXf = reinterpret_cast<double*>(smem_raw + off);                   // 2 x poses x 12 transforms
    off += (size_t)2 * p.poses * 96;
    unsigned long long* sBar = reinterpret_cast<unsigned long long*>(smem_raw + off);  // 2 mbarriers
    off += 16;
    int* sMin = reinterpret_cast<int*>(smem_raw + off);

    const int tid = threadIdx.x;
    const int pose_local = tid / p.chunks;
    const int chunk = tid - pose_local * p.chunks;
    const bool lane_used = pose_local < p.poses;
    const unsigned bar0 = smem_u32(sBar), xf0 = smem_u32(sXf);

    int cur_a = -1, cur_b = -1;
    for (int i = tid; i < p.poses; i += blockDim.x) sMin[i] = 0x7fffffff;
    if (p.use_tma && tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if ((long long)blockIdx.x < p.n_tiles) {  // transforms of the first tile
            int ca, cb, cnt;
            long long fst;
            clash_tile(p, blockIdx.x, ca, cb, fst, cnt);
            tma_load_1d(xf0, p.xf + fst * 12, (unsigned)cnt * 96u, bar0);
        }
    }
    __syncthreads();

    int it = 0;
    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
        int conf_a, conf_b, count;
        long long first;
        clash_tile(p, tile, conf_a, conf_b, first, count);
        if (conf_a != cur_a || conf_b != cur_b) {
            __syncthreads();  // everybody is done with the previous tables
            if (conf_a != cur_a) {
                const ulonglong2* src =
                    reinterpret_cast<const ulonglong2*>(p.a_tab) + (size_t)conf_a * p.n_a_pad;
                for (int i = tid; i < p.n_a_pad; i += blockDim.x) sA[i] = src[i];
                if (tid < 2) sA[p.n_a_pad + tid] = make_ulonglong2(0ull, 0ull);
            }
            if (conf_b != cur_b) {
                const float4* src = p.b_tab + (size_t)conf_b * p.n_b_pad;
                for (int i = tid; i < p.n_b_pad; i += blockDim.x) sB[i] = src[i];
            }
            cur_a = conf_a;
            cur_b = conf_b;
        }
        __syncthreads();  // tables + sMin reset visible; previous tile's transform buffer is free
        const int buf = it & 1;
        if (p.use_tma) {
            if (tid == 0 && tile + gridDim.x < p.n_tiles) {  // prefetch the next tile's transforms
                int ca, cb, cnt;
                long long fst;
                clash_tile(p, tile + gridDim.x, ca, cb, fst, cnt);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                tma_load_1d(xf0 + (unsigned)(buf ^ 1) * (unsigned)p.poses * 96u, p.xf + fst * 12, (unsigned)cnt * 96u,
                            bar0 + 8u * (unsigned)(buf ^ 1));
            }
            mbar_wait(bar0 + 8u * (unsigned)buf, (unsigned)(it >> 1) & 1u);
        }

        const bool active = lane_used && pose_local < count;
        const long long pose = first + pose_local;
        float tnorm = 0.f;
        if (active) {
            // ---- transform this thread's TB atoms of fragment B --------------------------------
            float r[12];
            if (p.use_tma) {
                const double* x = sXf + (size_t)buf * p.poses * 12 + (size_t)pose_local * 12;
#pragma unroll
                for (int k = 0; k < 12; k += 2) {
                    double2 v = *reinterpret_cast<const double2*>(x + k);
                    r[k] = (float)v.x;
                    r[k + 1] = (float)v.y;
                }
            } else {
                const double* x = p.xf + pose * 12;
#pragma unroll
                for (int k = 0; k < 12; ++k) r[k] = (float)__ldg(x + k);
            }
            tnorm = sqrtf(fmaf(r[9], r[9], fmaf(r[10], r[10], r[11] * r[11])));

            float bx[TB], by[TB], bz[TB], m[TB];
#pragma unroll
            for (int j = 0; j < TB; ++j) {
                float4 b = sB[chunk * TB + j];
                bx[j] = fmaf(r[0], b.x, fmaf(r[1], b.y, fmaf(r[2], b.z, r[9])));
                by[j] = fmaf(r[3], b.x, fmaf(r[4], b.y, fmaf(r[5], b.z, r[10])));
                bz[j] = fmaf(r[6], b.x, fmaf(r[7], b.y, fmaf(r[8], b.z, r[11])));
                m[j] = 3.0e38f;
            }

            // ---- all atom pairs: two A atoms per step, broadcast from shared memory ------------
            const int n_pairs = p.n_a_pad >> 1;
            if (PF) {
                // A pair of step i + 1 is fetched into its own registers while step i computes, so the
                // first FFMA2 of a step never waits for shared memory (sA holds one padding pair)
                ulonglong2 u0 = sA[0], u1 = sA[1];
#pragma unroll UNR
                for (int i = 0; i < n_pairs; ++i) {
                    ulonglong2 v0 = sA[2 * i + 2], v1 = sA[2 * i + 3];
#pragma unroll
                    for (int j = 0; j < TB; ++j) {
                        f32x2 e = fma2(u0.x, pack2(bx[j], bx[j]),
                                       fma2(u0.y, pack2(by[j], by[j]), fma2(u1.x, pack2(bz[j], bz[j]), u1.y)));
                        float lo, hi;
                        unpack2(e, lo, hi);
                        m[j] = fminf(fminf(m[j], lo), hi);
                    }
                    u0 = v0;
                    u1 = v1;
                }
            } else {
#pragma unroll UNR
                for (int i = 0; i < n_pairs; ++i) {
                    ulonglong2 u0 = sA[2 * i], u1 = sA[2 * i + 1];
#pragma unroll
                    for (int j = 0; j < TB; ++j) {
                        f32x2 e = fma2(u0.x, pack2(bx[j], bx[j]),
                                       fma2(u0.y, pack2(by[j], by[j]), fma2(u1.x, pack2(bz[j], bz[j]), u1.y)));
                        float lo, hi;
                        unpack2(e, lo, hi);
                        m[j] = fminf(fminf(m[j], lo), hi);
                    }
                }
            }
            float lane_min = 3.0e38f;
#pragma unroll
            for (int j = 0; j < TB; ++j) {
                float nb = fmaf(bx[j], bx[j], fmaf(by[j], by[j], bz[j] * bz[j])) + sB[chunk * TB + j].w;
                lane_min = fminf(lane_min, m[j] + nb);
            }
            atomicMin(&sMin[pose_local], float_key(lane_min));
        }
        __syncthreads();
        if (active && chunk == 0) {
            float d2 = key_float(sMin[pose_local]);
            sMin[pose_local] = 0x7fffffff;  // reset for the next tile (ordered by the barrier above)
            // rounding band of the Gram-form FP32 evaluation (DESIGN.md "clash band")
            float ext = p.a_rad[conf_a] + p.b_rad[conf_b] + tnorm;
            float band = fmaf(1.5e-6f * ext, ext, 1e-6f);
            uint8_t st;
            bool uncertain;
            if (p.count_mode) {
                uncertain = !(d2 > p.thr2 + band);
                st = FC_STATUS_PASS;
            } else {
                uncertain = fabsf(d2 - p.thr2) <= band;
                st = d2 > p.thr2 ? FC_STATUS_PASS : 0;
            }
            if (uncertain) {
                int slot = atomicAdd(p.unc_count, 1);
                UncEntry e;
                e.pose = pose;
                e.conf_a = conf_a;
                e.conf_b = conf_b;
                p.unc_list[slot] = e;
            }
            p.status[pose] = st;
            if (p.min_dist) p.min_dist[pose] = sqrtf(fmaxf(d2, 0.f));
        }
        // next iteration's first barrier orders the sMin reset before any atomicMin
    }
}

// ---------------------------------------------------------------------------------------------
// Cell-list screen: the same decision without touching far atom pairs.
// ---------------------------------------------------------------------------------------------
// Fragment A (per conformer) is binned once per call into a G^3 grid over the bounding box of the whole
// ensemble, padded by the clash radius.  A cell record (16 bytes, one LDG.128) holds the number of
// candidate atoms and up to 15 of their indices: every atom whose distance to the cell CENTRE is at
// most  rc = thresh + kCellPad + h*sqrt(3)/2  -- a superset of the atoms within thresh + kCellPad of ANY
// point of the cell, hence of every atom that can decide the pose (the FP32 band is far below kCellPad;
// poses whose band is not are sent to the FP64 recheck).  A query (pose, atom of B) transforms the atom,
// finds its cell and evaluates only the listed atoms, in the difference form; cells with more than 15
// candidates (count byte 255) fall back to scanning all atoms of A for that query.
// One thread per pose, one warp per 32 consecutive poses; fragments and grid are read through L1/L2
// (150 atoms = 2.4 KB, grid = G^3 * 16 B per conformer).  The FP32 band / FP64 recheck protocol of the
// all-pairs kernel is unchanged, so both paths produce identical status bytes.
constexpr float kCellPad = 0.05f;

struct CellMeta {        // written by clash_bbox_kernel, read by the grid / query kernels
    float ox, oy, oz;    // grid origin
    float h, inv_h;      // cell edge
    float rc2;           // squared candidate radius around a cell centre
    int g;               // cells per axis
    int gmask;           // 2^ceil(log2 g) - 1
    unsigned cell_bias;  // 0x4B400000 * (g*g + g + 1) mod 2^32 (biased -> linear cell index)
};

__global__ void __launch_bounds__(256) clash_bbox_kernel(const double* __restrict__ a_coords, long long n_atoms_total,
                                                         float thresh, int g, CellMeta* __restrict__ meta) {
    __shared__ float s_lo[3][8], s_hi[3][8];
    float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
    for (long long i = threadIdx.x; i < n_atoms_total; i += blockDim.x)
        for (int c = 0; c < 3; ++c) {
            float v = (float)a_coords[3 * i + c];
            lo[c] = fminf(lo[c], v);
            hi[c] = fmaxf(hi[c], v);
        }
    for (int c = 0; c < 3; ++c) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
        if ((threadIdx.x & 31) == 0) { s_lo[c][threadIdx.x >> 5] = lo[c]; s_hi[c][threadIdx.x >> 5] = hi[c]; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float ext = 0.f, l[3];
        const float pad = thresh + kCellPad + 0.01f;
        for (int c = 0; c < 3; ++c) {
            float a = 3e38f, b = -3e38f;
            for (int w = 0; w < 8; ++w) { a = fminf(a, s_lo[c][w]); b = fmaxf(b, s_hi[c][w]); }
            l[c] = a - pad;
            ext = fmaxf(ext, (b + pad) - l[c]);
        }
        CellMeta m;
        m.ox = l[0]; m.oy = l[1]; m.oz = l[2];
        m.h = ext / (float)g * 1.0001f;
        m.inv_h = 1.0f / m.h;
        float rc = thresh + kCellPad + m.h * 0.8661f + 1e-3f;
        m.rc2 = rc * rc;
        m.g = g;
        int pw = 1;
        while (pw < g) pw <<= 1;
        m.gmask = pw - 1;
        m.cell_bias = 0x4B400000u * (unsigned)(g * g + g + 1);
        *meta = m;
    }
}

// a_xyz: [conf][n_a] float4 {x, y, z, 0};  grid: [conf][g^3] uint4 = {count, idx0..14} as bytes
// occ:   [conf][g^3 / 32] one bit per cell: the cell has candidates (32 KB per conformer for g = 64: L1-resident)
__global__ void __launch_bounds__(128) clash_grid_kernel(const double* __restrict__ a_coords, int n_a,
                                                         const CellMeta* __restrict__ meta, float4* __restrict__ a_xyz,
                                                         uint4* __restrict__ grid, unsigned* __restrict__ occ) {
    extern __shared__ double s_a[];  // n_a * 3
    const int conf = blockIdx.y;
    const double* src = a_coords + (size_t)conf * n_a * 3;
    for (int i = threadIdx.x; i < n_a * 3; i += blockDim.x) s_a[i] = src[i];
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < n_a; i += blockDim.x)
            a_xyz[(size_t)conf * n_a + i] = make_float4((float)src[3 * i], (float)src[3 * i + 1], (float)src[3 * i + 2], 0.f);
    __syncthreads();
    const CellMeta m = *meta;
    const int g = m.g;
    const long long n_cells = (long long)g * g * g;
    const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // g^3 is a multiple of 128
    if (cell >= n_cells) return;
    const int cx = (int)(cell % g), cy = (int)((cell / g) % g), cz = (int)(cell / ((long long)g * g));
    const double px = (double)m.ox + ((double)cx + 0.5) * (double)m.h;
    const double py = (double)m.oy + ((double)cy + 0.5) * (double)m.h;
    const double pz = (double)m.oz + ((double)cz + 0.5) * (double)m.h;
    unsigned bytes[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) bytes[k] = 0;
    int count = 0;
    for (int i = 0; i < n_a; ++i) {
        double dx = s_a[3 * i] - px, dy = s_a[3 * i + 1] - py, dz = s_a[3 * i + 2] - pz;
        if (dx * dx + dy * dy + dz * dz <= (double)m.rc2) {
            if (count < 15) {
#pragma unroll
                for (int k = 1; k < 16; ++k)
                    if (k == count + 1) bytes[k] = (unsigned)i;
            }
            ++count;
        }
    }
    bytes[0] = count > 15 ? 255u : (unsigned)count;
    uint4 rec;
    rec.x = bytes[0] | (bytes[1] << 8) | (bytes[2] << 16) | (bytes[3] << 24);
    rec.y = bytes[4] | (bytes[5] << 8) | (bytes[6] << 16) | (bytes[7] << 24);
    rec.z = bytes[8] | (bytes[9] << 8) | (bytes[10] << 16) | (bytes[11] << 24);
    rec.w = bytes[12] | (bytes[13] << 8) | (bytes[14] << 16) | (bytes[15] << 24);
    grid[(size_t)conf * n_cells + cell] = rec;
    const unsigned any = __ballot_sync(0xffffffffu, count > 0);
    if ((threadIdx.x & 31) == 0) occ[(size_t)conf * (n_cells / 32 + 1) + (cell >> 5)] = any;
    if (cell == 0) occ[(size_t)conf * (n_cells / 32 + 1) + n_cells / 32] = 0u;  // spare word: the dummy cell
}

struct CellArgs {
    const float4* a_xyz;
    const float* a_rad;
    const float4* b_tab;
    const float* b_rad;
    const uint4* grid;
    const unsigned* occ;
    const CellMeta* meta;
    const double* xf;
    const int4* tiles;
    long long n_tiles, n_poses;
    int n_a, n_b, n_b_pad, tile_poses;
    float thr2;
    int count_mode;
    uint8_t* status;
    int* unc_count;
    UncEntry* unc_list;
};

__global__ void __launch_bounds__(128) clash_cell_kernel(CellArgs p) {
    const long long pose = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pose >= p.n_poses) return;
    int conf_a = 0, conf_b = 0;
    if (p.tiles) {  // tiles are sorted by first pose: binary search
        long long lo = 0, hi = p.n_tiles - 1;
        while (lo < hi) {
            long long mid = (lo + hi + 1) >> 1;
            if ((long long)p.tiles[mid].z <= pose) lo = mid;
            else hi = mid - 1;
        }
        int4 t = p.tiles[lo];
        if (pose >= (long long)t.z + t.w) return;  // pose not covered by any tile
        conf_a = t.x;
        conf_b = t.y;
    }
    const CellMeta m = *p.meta;
    const int g = m.g;
    const double* x = p.xf + pose * 12;
    float r[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) r[k] = (float)__ldg(x + k);
    const float tnorm = sqrtf(fmaf(r[9], r[9], fmaf(r[10], r[10], r[11] * r[11])));
    const float4* bt = p.b_tab + (size_t)conf_b * p.n_b_pad;
    const float4* at = p.a_xyz + (size_t)conf_a * p.n_a;
    const uint4* gr = p.grid + (size_t)conf_a * g * g * g;
    const unsigned* oc = p.occ + ((size_t)conf_a * ((size_t)g * g * g / 32 + 1));
    // same band as the all-pairs kernel (the difference form is at least as accurate as the Gram form)
    const float ext = p.a_rad[conf_a] + p.b_rad[conf_b] + tnorm;
    const float band = fmaf(1.5e-6f * ext, ext, 1e-6f);
    // once the minimum is below this value the pose is decided (certain clash, or -- with max_clashes > 0 --
    // certain to need the FP64 count): the remaining atoms are skipped
    const float settle = p.count_mode ? p.thr2 + band : p.thr2 - band;
    float dmin2 = 3.0e38f;
    // grid coordinates straight from the atom: (R b + t - o) / h, with the float->int conversion done by the
    // FP32 adder (adding 1.5 * 2^23 leaves round-to-nearest(x - 0.5) in the low mantissa bits; which of two
    // cells a point ON a cell face goes to is irrelevant, the candidate lists carry that slack, and both
    // phases use the same arithmetic).  F2I would run on the quarter-rate conversion unit.
    float q[12];
#pragma unroll
    for (int k = 0; k < 9; ++k) q[k] = r[k] * m.inv_h;
    q[9] = (r[9] - m.ox) * m.inv_h - 0.5f;
    q[10] = (r[10] - m.oy) * m.inv_h - 0.5f;
    q[11] = (r[11] - m.oz) * m.inv_h - 0.5f;
    const unsigned kMagicBits = 0x4B400000u;
    const unsigned range_mask = ~(unsigned)(m.gmask);  // bits that must equal kMagicBits for 0 <= n < 2^k
    const unsigned dummy_cell = (unsigned)g * g * g;    // one spare, always-empty word behind the bit-grid
    for (int j0 = 0; j0 < p.n_b && !(dmin2 < settle); j0 += 32) {
        const int jn = min(32, p.n_b - j0);
        // ---- phase 1: flag the atoms of B that land in a cell with candidates (no distances yet), so that
        //      the lanes of a warp do not wait for each other's candidate loops on every atom.  Integer work
        //      is kept minimal: the ALU pipe (16 lanes) is what bounds this loop.
        unsigned bits = 0u;
#pragma unroll 4
        for (int k = 0; k < jn; ++k) {
            const float4 b = __ldg(bt + j0 + k);
            const unsigned ix = __float_as_uint(fmaf(q[0], b.x, fmaf(q[1], b.y, fmaf(q[2], b.z, q[9]))) + 12582912.0f);
            const unsigned iy = __float_as_uint(fmaf(q[3], b.x, fmaf(q[4], b.y, fmaf(q[5], b.z, q[10]))) + 12582912.0f);
            const unsigned iz = __float_as_uint(fmaf(q[6], b.x, fmaf(q[7], b.y, fmaf(q[8], b.z, q[11]))) + 12582912.0f);
            // all three of the form kMagicBits + n with n < 2^k  <=>  their OR is (a point that is really inside
            // always passes; a far-away point that slips through only costs a phase-2 visit, which re-checks)
            const bool inside = ((ix | iy | iz) & range_mask) == kMagicBits;
            const unsigned cell = inside ? (iz * (unsigned)(g * g) + iy * (unsigned)g + ix - m.cell_bias) : dummy_cell;
            const unsigned word = __ldg(oc + (cell >> 5));
            bits = __funnelshift_r(bits, __funnelshift_r(word, 0u, cell), 1);  // bit (cell & 31) of word -> top of bits
        }
        bits >>= (32 - jn);
        // ---- phase 2: distances to the candidate atoms of the flagged atoms only
        while (bits && !(dmin2 < settle)) {
            const int k = __ffs(bits) - 1;
            bits &= bits - 1u;
            const float4 b = __ldg(bt + j0 + k);
            const int cx = (int)(__float_as_uint(fmaf(q[0], b.x, fmaf(q[1], b.y, fmaf(q[2], b.z, q[9]))) + 12582912.0f) - kMagicBits);
            const int cy = (int)(__float_as_uint(fmaf(q[3], b.x, fmaf(q[4], b.y, fmaf(q[5], b.z, q[10]))) + 12582912.0f) - kMagicBits);
            const int cz = (int)(__float_as_uint(fmaf(q[6], b.x, fmaf(q[7], b.y, fmaf(q[8], b.z, q[11]))) + 12582912.0f) - kMagicBits);
            if ((unsigned)cx >= (unsigned)g || (unsigned)cy >= (unsigned)g || (unsigned)cz >= (unsigned)g) continue;
            const float bx = fmaf(r[0], b.x, fmaf(r[1], b.y, fmaf(r[2], b.z, r[9])));
            const float by = fmaf(r[3], b.x, fmaf(r[4], b.y, fmaf(r[5], b.z, r[10])));
            const float bz = fmaf(r[6], b.x, fmaf(r[7], b.y, fmaf(r[8], b.z, r[11])));
            const uint4 rec = __ldg(gr + ((size_t)cz * g + cy) * g + cx);
            const unsigned count = rec.x & 0xffu;
            if (count == 255u) {  // crowded cell: all atoms of A
                for (int i = 0; i < p.n_a; ++i) {
                    const float4 a = __ldg(at + i);
                    const float dx = a.x - bx, dy = a.y - by, dz = a.z - bz;
                    dmin2 = fminf(dmin2, fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
                }
                continue;
            }
            unsigned long long q0 = (((unsigned long long)rec.y << 32) | rec.x) >> 8;  // indices 0..6
            unsigned long long q1 = ((unsigned long long)rec.w << 32) | rec.z;         // indices 7..14
            q0 |= q1 << 56;
            q1 >>= 8;
            for (unsigned c = 0; c < count; ++c) {
                const unsigned ai = (unsigned)(q0 & 0xffull);
                q0 = (q0 >> 8) | (q1 << 56);
                q1 >>= 8;
                const float4 a = __ldg(at + ai);
                const float dx = a.x - bx, dy = a.y - by, dz = a.z - bz;
                dmin2 = fminf(dmin2, fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
            }
        }
    }
    uint8_t st;
    bool uncertain;
    if (p.count_mode) {
        uncertain = !(dmin2 > p.thr2 + band);
        st = FC_STATUS_PASS;
    } else {
        uncertain = fabsf(dmin2 - p.thr2) <= band;
        st = dmin2 > p.thr2 ? FC_STATUS_PASS : 0;
    }
    // the candidate lists only cover thresh + kCellPad: a band that large cannot be trusted to them
    if (band > 2.0f * kCellPad * sqrtf(p.thr2) * 0.5f) uncertain = true;
    if (uncertain) {
        int slot = atomicAdd(p.unc_count, 1);
        UncEntry e;
        e.pose = pose;
        e.conf_a = conf_a;
        e.conf_b = conf_b;
        p.unc_list[slot] = e;
    }
    p.status[pose] = st;
}

// ---------------------------------------------------------------------------------------------
// FP64 recheck: one warp per undecided pose, reference arithmetic (utils.py:544-575)
// ---------------------------------------------------------------------------------------------
struct RecheckArgs {
    const double* a_coords;
    const double* b_coords;
    const double* xf;
    int n_a, n_b;
    double thresh;
    int max_clashes;
    int strict;
    const int* unc_count;
    const UncEntry* unc_list;
    uint8_t* status;
    float* min_dist;
    int* near_count;
    long long* near_idx;
    double* near_dist;
    long long near_cap;
    long long pose_base;
};

__global__ void __launch_bounds__(256) clash_recheck_f64_kernel(RecheckArgs p) {
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int n_unc = *p.unc_count;
    for (int u = blockIdx.x * warps_per_block + (threadIdx.x >> 5); u < n_unc;
         u += gridDim.x * warps_per_block) {
        UncEntry e = p.unc_list[u];
        const double* a = p.a_coords + (size_t)e.conf_a * p.n_a * 3;
        const double* b = p.b_coords + (size_t)e.conf_b * p.n_b * 3;
        const double* x = p.xf + e.pose * 12;
        double r[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) r[k] = x[k];
        int clashes = 0;
        double dmin = 1e300, closest = 1e300;
        for (int j = lane; j < p.n_b; j += 32) {
            double bx = b[3 * j], by = b[3 * j + 1], bz = b[3 * j + 2];
            // (R @ b) + t, the get_embed expression embeds.py:815-817
            double px = (r[0] * bx + r[1] * by + r[2] * bz) + r[9];
            double py = (r[3] * bx + r[4] * by + r[5] * bz) + r[10];
            double pz = (r[6] * bx + r[7] * by + r[8] * bz) + r[11];
            for (int i = 0; i < p.n_a; ++i) {
                double dx = px - a[3 * i], dy = py - a[3 * i + 1], dz = pz - a[3 * i + 2];
                double d = sqrt(dx * dx + dy * dy + dz * dz);
                bool hit = p.strict ? (d < p.thresh) : (d <= p.thresh);
                clashes += hit ? 1 : 0;
                dmin = fmin(dmin, d);
                closest = fmin(closest, fabs(d - p.thresh));
            }
        }
        clashes = warp_sum(clashes);
        dmin = warp_min(dmin);
        closest = warp_min(closest);
        if (lane == 0) {
            uint8_t st = FC_STATUS_RECHECKED;
            if (clashes <= p.max_clashes) st |= FC_STATUS_PASS;
            // with max_clashes == 0 only the minimum distance decides; otherwise any pair may
            const double margin = p.max_clashes == 0 ? fabs(dmin - p.thresh) : closest;
            if (margin <= FC_NEAR_EPS) {
                st |= FC_STATUS_NEAR;
                if (p.near_count) {
                    int slot = atomicAdd(p.near_count, 1);
                    if (slot < p.near_cap) {
                        if (p.near_idx) p.near_idx[slot] = e.pose + p.pose_base;
                        if (p.near_dist) p.near_dist[slot] = dmin;
                    }
                }
            }
            p.status[e.pose] = st;
            if (p.min_dist) p.min_dist[e.pose] = (float)dmin;
        }
    }
}

// optional CUDA-event timing of the dominant kernel (bench.py roofline leg)
static thread_local bool g_time_on = false;
static thread_local cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;
static thread_local double g_ms_sum = 0.0;
static thread_local long g_ms_n = 0;
static thread_local bool g_ev_pending = false;

static void timing_flush() {
    if (g_ev_pending) {
        float ms = 0.f;
        if (cudaEventSynchronize(g_ev1) == cudaSuccess && cudaEventElapsedTime(&ms, g_ev0, g_ev1) == cudaSuccess) {
            g_ms_sum += ms;
            g_ms_n += 1;
        }
        g_ev_pending = false;
    }
}

static void timing_begin(cudaStream_t s) {
    if (!g_time_on) return;
    timing_flush();
    if (!g_ev0) {
        cudaEventCreate(&g_ev0);
        cudaEventCreate(&g_ev1);
    }
    cudaEventRecord(g_ev0, s);
}
static void timing_end(cudaStream_t s) {
    if (!g_time_on) return;
    cudaEventRecord(g_ev1, s);
    g_ev_pending = true;
}

template <int TB, int UNR, int PF>
static cudaError_t launch_f32v(const ClashArgs& args, int threads, size_t smem, int grid,
                               cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(clash_f32_kernel<TB, UNR, PF>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    timing_begin(s);
    clash_f32_kernel<TB, UNR, PF><<<grid, threads, smem, s>>>(args);
    e = cudaGetLastError();
    timing_end(s);
    return e;
}

template <int TB>
static cudaError_t launch_f32(const ClashArgs& args, int threads, size_t smem, int grid,
                              cudaStream_t s) {
    static int unr = -1;
    if (unr < 0) {
        const char* v = getenv("FC_CLASH_UNROLL");
        unr = v ? atoi(v) : 11;  // default: A pairs register-prefetched, no unrolling (measured fastest on B200)
    }
    switch (unr) {
        case 1: return launch_f32v<TB, 1, 0>(args, threads, smem, grid, s);
        case 2: return launch_f32v<TB, 2, 0>(args, threads, smem, grid, s);
        case 13: return launch_f32v<TB, 3, 1>(args, threads, smem, grid, s);
        case 15: return launch_f32v<TB, 5, 1>(args, threads, smem, grid, s);
        default: return launch_f32v<TB, 1, 1>(args, threads, smem, grid, s);
    }
}

}  // namespace fc

using namespace fc;

extern "C" int fc_clash_timing(int enable, double* ms_sum, int64_t* launches) {
    timing_flush();
    if (ms_sum) *ms_sum = g_ms_sum;
    if (launches) *launches = g_ms_n;
    g_time_on = enable != 0;
    g_ms_sum = 0.0;
    g_ms_n = 0;
    return FC_OK;
}

extern "C" int fc_clash_geometry(int n_b, int32_t* out4) {
    ClashGeom g = choose_geom(n_b);
    if (out4) {
        out4[0] = g.tb;
        out4[1] = g.chunks;
        out4[2] = g.poses;
        out4[3] = g.threads;
    }
    return g.tb > 0 ? FC_OK : FC_ERR_INVALID;
}

extern "C" int fc_clash_tile_poses(int n_b) {
    ClashGeom g = choose_geom(n_b);
    return g.poses;
}

// A-side tables of a screen, reusable across calls that share fragment A and the threshold
struct fc_clash_prep {
    int n_conf_a = 0, n_a = 0, n_a_pad = 0, cell_g = 0;
    double thresh = 0.0;
    unsigned char* buf = nullptr;  // stream-ordered allocation holding everything below
    const float4* a_tab = nullptr;  // Gram-form atom pairs (all-pairs kernel)
    const float* a_rad = nullptr;
    const float4* a_xyz = nullptr;  // cell-list tables (cell_g > 0)
    const CellMeta* meta = nullptr;
    const uint4* grid = nullptr;
    const unsigned* occ = nullptr;
};

extern "C" int fc_clash_prepare_dev(const double* a_coords, int n_conf_a, int n_a, double thresh, int want_cells,
                                    fc_clash_prep** out, void* stream) {
    FC_REQUIRE(out, "fc_clash_prepare_dev: null output");
    *out = nullptr;
    FC_REQUIRE(a_coords && n_conf_a > 0 && n_a > 0, "fc_clash_prepare_dev: empty fragment");
    cudaStream_t s = (cudaStream_t)stream;
    sm_count();
    fc_clash_prep* p = new fc_clash_prep();
    p->n_conf_a = n_conf_a;
    p->n_a = n_a;
    p->n_a_pad = (n_a + 1) / 2 * 2;
    p->thresh = thresh;
    if (want_cells && n_a <= 254 && thresh > 0.0 && thresh < 1e3) {
        const size_t budget = (size_t)512 << 20;
        for (int g_try : {64, 32}) {  // powers of two: the flag phase range-checks with one mask
            if ((size_t)n_conf_a * g_try * g_try * g_try * 16 <= budget) {
                p->cell_g = g_try;
                break;
            }
        }
    }
    const int cell_g = p->cell_g;
    const size_t n_cells = (size_t)cell_g * cell_g * cell_g;
    size_t off_a = 0;
    size_t off_ra = off_a + (size_t)n_conf_a * p->n_a_pad * 16;
    size_t off_axyz = (off_ra + (size_t)n_conf_a * 4 + 15) / 16 * 16;
    size_t off_meta = off_axyz + (cell_g ? (size_t)n_conf_a * n_a * 16 : 0);
    size_t off_grid = off_meta + (cell_g ? 64 : 0);
    size_t off_occ = off_grid + (cell_g ? (size_t)n_conf_a * n_cells * 16 : 0);
    size_t total = off_occ + (cell_g ? (size_t)n_conf_a * (n_cells / 32 + 1) * 4 : 0);
    cudaError_t e = cudaMallocAsync((void**)&p->buf, total, s);
    if (e != cudaSuccess) {
        delete p;
        return cuda_fail(e, "fc_clash_prepare_dev: cudaMallocAsync", __FILE__, __LINE__);
    }
    clash_prep_kernel<<<n_conf_a, 128, 0, s>>>(a_coords, n_conf_a, n_a, p->n_a_pad, 1, (float4*)(p->buf + off_a),
                                               (float*)(p->buf + off_ra));
    p->a_tab = (const float4*)(p->buf + off_a);
    p->a_rad = (const float*)(p->buf + off_ra);
    if (cell_g) {
        CellMeta* meta = (CellMeta*)(p->buf + off_meta);
        float4* a_xyz = (float4*)(p->buf + off_axyz);
        uint4* grid_tab = (uint4*)(p->buf + off_grid);
        unsigned* occ = (unsigned*)(p->buf + off_occ);
        clash_bbox_kernel<<<1, 256, 0, s>>>(a_coords, (long long)n_conf_a * n_a, (float)thresh, cell_g, meta);
        dim3 gg((unsigned)((n_cells + 127) / 128), (unsigned)n_conf_a);
        clash_grid_kernel<<<gg, 128, (size_t)n_a * 24, s>>>(a_coords, n_a, meta, a_xyz, grid_tab, occ);
        p->a_xyz = a_xyz;
        p->meta = meta;
        p->grid = grid_tab;
        p->occ = occ;
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) {
        cudaFreeAsync(p->buf, s);
        delete p;
        return cuda_fail(e, "fc_clash_prepare_dev kernels", __FILE__, __LINE__);
    }
    *out = p;
    return FC_OK;
}

extern "C" void fc_clash_prep_free(fc_clash_prep* p, void* stream) {
    if (!p) return;
    if (p->buf) cudaFreeAsync(p->buf, (cudaStream_t)stream);
    delete p;
}

extern "C" int fc_clash_screen_prepared_dev(const fc_clash_prep* prep, const double* a_coords, const double* b_coords,
                                            int n_conf_b, int n_b, const double* xf, int64_t n_poses,
                                            const int32_t* tiles, int64_t n_tiles, int max_clashes, int strict,
                                            uint8_t* status, float* min_dist, int32_t* near_count, int64_t* near_idx,
                                            double* near_dist, int64_t near_cap, int64_t pose_index_base, void* stream) {
    FC_REQUIRE(prep, "fc_clash_screen_prepared_dev: null preparation");
    const int n_conf_a = prep->n_conf_a, n_a = prep->n_a;
    const double thresh = prep->thresh;
    FC_REQUIRE(n_b > 0 && n_conf_b > 0, "fc_clash_screen_dev: empty fragment");
    FC_REQUIRE(n_poses >= 0 && max_clashes >= 0, "fc_clash_screen_dev: negative size");
    if (n_poses == 0) return FC_OK;
    FC_REQUIRE(a_coords && b_coords && xf && status, "fc_clash_screen_dev: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    const int sms = sm_count();  // also configures the stream-ordered memory pool on first use
    ClashGeom g = choose_geom(n_b);
    FC_REQUIRE(g.tb > 0, "fc_clash_screen_dev: fragment B too large (%d atoms)", n_b);
    const int n_a_pad = prep->n_a_pad;
    const int n_b_pad = g.chunks * g.tb;
    if (!tiles) n_tiles = (n_poses + g.poses - 1) / g.poses;
    FC_REQUIRE(n_tiles > 0, "fc_clash_screen_dev: empty tile list");

    size_t smem = (size_t)(n_a_pad + 2) * 16 + (size_t)n_b_pad * 16 + (size_t)2 * g.poses * 96 + 16 + (size_t)g.poses * 4;
    FC_REQUIRE(smem <= 227 * 1024, "fc_clash_screen_dev: fragments need %zu B of shared memory", smem);

    // ---- path: all atom pairs (Gram-form FP32 kernel) or cell lists over fragment A -----------------
    int cell_g = 0;
    {
        // FC_CLASH_MODE: 0 = all pairs, 1 = cell lists whenever possible, unset = automatic
        const char* v = getenv("FC_CLASH_MODE");
        const int mode = (v && *v) ? (atoi(v) ? 1 : 0) : 2;
        const bool possible = !min_dist && prep->cell_g > 0;
        const bool wanted = mode == 1 || (mode == 2 && n_poses >= 16384);
        if (possible && wanted) cell_g = prep->cell_g;
    }

    // stream-ordered scratch: B tables, undecided-pose list
    size_t b_bytes = (size_t)n_conf_b * n_b_pad * 16;
    size_t off_b = 0, off_rb = off_b + b_bytes, off_cnt = (off_rb + (size_t)n_conf_b * 4 + 15) / 16 * 16;
    size_t off_list = off_cnt + 16;
    size_t total = off_list + (size_t)n_poses * sizeof(UncEntry);
    unsigned char* scratch = nullptr;
    FC_CUDA(cudaMallocAsync((void**)&scratch, total, s));

    clash_prep_kernel<<<n_conf_b, 128, 0, s>>>(b_coords, n_conf_b, n_b, n_b_pad, 0,
                                               (float4*)(scratch + off_b), (float*)(scratch + off_rb));
    FC_CUDA(cudaMemsetAsync(scratch + off_cnt, 0, 16, s));

    ClashArgs a;
    a.a_tab = prep->a_tab;
    a.a_rad = prep->a_rad;
    a.b_tab = (const float4*)(scratch + off_b);
    a.b_rad = (const float*)(scratch + off_rb);
    a.xf = xf;
    a.tiles = (const int4*)tiles;
    a.n_tiles = n_tiles;
    a.n_poses = n_poses;
    a.n_a_pad = n_a_pad;
    a.n_b_pad = n_b_pad;
    a.chunks = g.chunks;
    a.poses = g.poses;
    a.thr2 = (float)(thresh * thresh);
    a.count_mode = max_clashes > 0;
    {
        const char* v = getenv("FC_CLASH_TMA");
        a.use_tma = ((reinterpret_cast<uintptr_t>(xf) & 15) == 0) && !(v && atoi(v) == 0);
    }
    a.status = status;
    a.min_dist = min_dist;
    a.unc_count = (int*)(scratch + off_cnt);
    a.unc_list = (UncEntry*)(scratch + off_list);

    if (cell_g) {
        CellArgs c;
        c.a_xyz = prep->a_xyz;
        c.a_rad = a.a_rad;
        c.b_tab = a.b_tab;
        c.b_rad = a.b_rad;
        c.grid = prep->grid;
        c.occ = prep->occ;
        c.meta = prep->meta;
        c.xf = xf;
        c.tiles = (const int4*)tiles;
        c.n_tiles = n_tiles;
        c.n_poses = n_poses;
        c.n_a = n_a;
        c.n_b = n_b;
        c.n_b_pad = n_b_pad;
        c.tile_poses = g.poses;
        c.thr2 = a.thr2;
        c.count_mode = a.count_mode;
        c.status = status;
        c.unc_count = a.unc_count;
        c.unc_list = a.unc_list;
        timing_begin(s);
        clash_cell_kernel<<<(unsigned)((n_poses + 127) / 128), 128, 0, s>>>(c);
        timing_end(s);
    }

    // persistent grid: resident CTAs per SM follow from the register/thread budget
    int ctas_per_sm = tb_max_threads(g.tb) / g.threads;
    if (ctas_per_sm > 8) ctas_per_sm = 8;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    long long grid_ll = (long long)sms * ctas_per_sm;
    if (grid_ll > n_tiles) grid_ll = n_tiles;
    int grid = (int)grid_ll;
    cudaError_t e;
    if (cell_g) {
        e = cudaGetLastError();
    } else
    switch (g.tb) {
        case 2: e = launch_f32<2>(a, g.threads, smem, grid, s); break;
        case 4: e = launch_f32<4>(a, g.threads, smem, grid, s); break;
        case 6: e = launch_f32<6>(a, g.threads, smem, grid, s); break;
        case 8: e = launch_f32<8>(a, g.threads, smem, grid, s); break;
        case 10: e = launch_f32<10>(a, g.threads, smem, grid, s); break;
        case 12: e = launch_f32<12>(a, g.threads, smem, grid, s); break;
        case 15: e = launch_f32<15>(a, g.threads, smem, grid, s); break;
        case 16: e = launch_f32<16>(a, g.threads, smem, grid, s); break;
        case 20: e = launch_f32<20>(a, g.threads, smem, grid, s); break;
        case 25: e = launch_f32<25>(a, g.threads, smem, grid, s); break;
        default: e = launch_f32<30>(a, g.threads, smem, grid, s); break;
    }
    if (e != cudaSuccess) {
        cudaFreeAsync(scratch, s);
        return cuda_fail(e, "clash_f32_kernel launch", __FILE__, __LINE__);
    }

    RecheckArgs r;
    r.a_coords = a_coords;
    r.b_coords = b_coords;
    r.xf = xf;
    r.n_a = n_a;
    r.n_b = n_b;
    r.thresh = thresh;
    r.max_clashes = max_clashes;
    r.strict = strict;
    r.unc_count = a.unc_count;
    r.unc_list = a.unc_list;
    r.status = status;
    r.min_dist = min_dist;
    r.near_count = near_count;
    r.near_idx = (long long*)near_idx;
    r.near_dist = near_dist;
    r.near_cap = near_cap;
    r.pose_base = pose_index_base;
    clash_recheck_f64_kernel<<<sms * 2, 256, 0, s>>>(r);
    e = cudaGetLastError();
    cudaError_t e2 = cudaFreeAsync(scratch, s);
    if (e != cudaSuccess) return cuda_fail(e, "clash_recheck_f64_kernel launch", __FILE__, __LINE__);
    if (e2 != cudaSuccess) return cuda_fail(e2, "cudaFreeAsync", __FILE__, __LINE__);
    return FC_OK;
}

extern "C" int fc_clash_screen_dev(const double* a_coords, int n_conf_a, int n_a,
                                   const double* b_coords, int n_conf_b, int n_b, const double* xf,
                                   int64_t n_poses, const int32_t* tiles, int64_t n_tiles,
                                   double thresh, int max_clashes, int strict, uint8_t* status,
                                   float* min_dist, int32_t* near_count, int64_t* near_idx,
                                   double* near_dist, int64_t near_cap, int64_t pose_index_base,
                                   void* stream) {
    FC_REQUIRE(n_a > 0 && n_b > 0 && n_conf_a > 0 && n_conf_b > 0, "fc_clash_screen_dev: empty fragment");
    FC_REQUIRE(n_poses >= 0 && max_clashes >= 0, "fc_clash_screen_dev: negative size");
    if (n_poses == 0) return FC_OK;
    FC_REQUIRE(a_coords && b_coords && xf && status, "fc_clash_screen_dev: null pointer");
    fc_clash_prep* prep = nullptr;
    const char* v = getenv("FC_CLASH_MODE");
    const int mode = (v && *v) ? (atoi(v) ? 1 : 0) : 2;
    const int want_cells = !min_dist && (mode == 1 || (mode == 2 && n_poses >= 16384));
    int rc = fc_clash_prepare_dev(a_coords, n_conf_a, n_a, thresh, want_cells, &prep, stream);
    if (rc) return rc;
    rc = fc_clash_screen_prepared_dev(prep, a_coords, b_coords, n_conf_b, n_b, xf, n_poses, tiles, n_tiles, max_clashes,
                                      strict, status, min_dist, near_count, near_idx, near_dist, near_cap,
                                      pose_index_base, stream);
    fc_clash_prep_free(prep, stream);
    return rc;
}
