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
#include <string.h>

#include <algorithm>
#include <chrono>

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
// Cell-list screen (v2): the same decision without touching far atom pairs.
// ---------------------------------------------------------------------------------------------
// Fragment A (per conformer) is binned once per call into a G^3 grid of cubic cells CENTRED on the lattice
// points o + n*h (n = 0 .. G-1 per axis).  A cell record (16 bytes, one LDG.128) holds the number of candidate
// atoms and up to 15 of their indices: every atom whose distance to the cell centre is at most
//      rc = thresh + kCellPad + h*sqrt(3)/2 + 1e-3
// -- a superset of the atoms within thresh + kCellPad of ANY point of the cell, hence of every atom that can
// decide the pose (the FP32 band and the cell-assignment rounding are far below kCellPad; poses whose band is
// not are sent to the FP64 recheck).  Next to it a 1-bit-per-cell occupancy grid (32 KB per conformer at G = 64:
// L1-resident).  The box is padded so that every BORDER cell is empty; a query point is clamped into the box for
// free by the .SAT modifier of its last FFMA, so a point outside the box reads an (empty) border cell -- which
// is the right answer, it is further than rc from every atom -- and neither phase needs a range check.
//
// One thread per pose; atoms of B in blocks of 32:
//   phase 1 flags the atoms that land in an occupied cell: normalised box coordinates u = sat((R b + t - o) / ((G-1) h))
//           (9 FFMA), one FFMA per axis scales u to the cell coordinate and adds 1.5 * 2^23 so that the cell number
//           appears in the low mantissa bits -- scaled per axis so that x, y, z land in disjoint bit fields and the
//           linear cell index is two LOP3s -- then one occupancy word and two funnel shifts;
//   phase 2 visits only the flagged atoms: one LDG.128 cell record, difference-form distances to its candidates,
//           early exit once the pose is decided (certain clash).
// Early exit only pays if the lanes of a warp exit together, so the screen runs in LEVELS over the atoms of B
// (atoms [0,32), [32,64), [64,n_b)), B being re-ordered once per call so that its first atoms are spread over its
// surface (farthest-point order): a level finalises the poses it can decide and appends the others -- (pose, running
// minimum) -- to a compact list that the next level processes with dense warps.
// The FP32 band / FP64 recheck protocol of the all-pairs kernel is unchanged, so both paths produce identical
// status bytes.  Poses arrive either as 12 doubles (R row-major, t) or in the compact form of fc_clash_screen_pose7_dev.
constexpr float kCellPad = 0.05f;
constexpr float kCellMagic = 12582912.0f;  // 1.5 * 2^23: the adder leaves round-to-nearest(x) in the low mantissa bits

struct CellMeta {        // written by clash_bbox_kernel, read by the grid / query kernels
    float ox, oy, oz;    // centre of cell (0, 0, 0)
    float h;             // cell edge
    float inv_span;      // 1 / ((g - 1) h): box coordinates u in [0, 1]
    float rc2;           // squared candidate radius around a cell centre
    int g;               // cells per axis (power of two)
    int shift;           // log2 g
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
        float span = 0.f, mid[3];
        for (int c = 0; c < 3; ++c) {
            float a = 3e38f, b = -3e38f;
            for (int w = 0; w < 8; ++w) { a = fminf(a, s_lo[c][w]); b = fmaxf(b, s_hi[c][w]); }
            mid[c] = 0.5f * (a + b);
            span = fmaxf(span, b - a);
        }
        // lattice of g points per axis over the atoms' extent plus, on both sides, more than the candidate radius:
        // (g-1) h = span + 2 (thresh + kCellPad) + 0.05 + 1.8 h  =>  border lattice points are further than rc from
        // every atom along their axis (margin 0.024 + 0.03 h), so border cells have no candidates
        const float tp = thresh + kCellPad;
        CellMeta m;
        m.h = (span + 2.f * tp + 0.05f) / ((float)g - 2.8f);
        const float ext = (float)(g - 1) * m.h;
        m.ox = mid[0] - 0.5f * ext;
        m.oy = mid[1] - 0.5f * ext;
        m.oz = mid[2] - 0.5f * ext;
        m.inv_span = 1.0f / ext;
        float rc = tp + m.h * 0.8661f + 1e-3f;
        m.rc2 = rc * rc;
        m.g = g;
        int sh = 0;
        while ((1 << sh) < g) ++sh;
        m.shift = sh;
        *meta = m;
    }
}

// a_u:  [conf][n_a] float4 {x, y, z, 0} in BOX coordinates u = (a - o) / ((g-1) h)  (the frame phase 2 measures in)
// grid: [conf][g^3] uint4 = {count, idx0..14} as bytes
// occ:  [conf][g^3 / 32] one bit per cell: the cell has candidates (32 KB per conformer for g = 64; the screen keeps
//       the current conformer's copy in shared memory)
__global__ void __launch_bounds__(128) clash_grid_kernel(const double* __restrict__ a_coords, int n_a,
                                                         const CellMeta* __restrict__ meta, float4* __restrict__ a_u,
                                                         uint4* __restrict__ grid, unsigned* __restrict__ occ) {
    extern __shared__ double s_a[];  // n_a * 3
    const int conf = blockIdx.y;
    const double* src = a_coords + (size_t)conf * n_a * 3;
    for (int i = threadIdx.x; i < n_a * 3; i += blockDim.x) s_a[i] = src[i];
    const CellMeta m = *meta;
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < n_a; i += blockDim.x) {
            const double inv = 1.0 / ((double)(m.g - 1) * (double)m.h);
            a_u[(size_t)conf * n_a + i] = make_float4((float)((src[3 * i] - (double)m.ox) * inv), (float)((src[3 * i + 1] - (double)m.oy) * inv),
                                                      (float)((src[3 * i + 2] - (double)m.oz) * inv), 0.f);
        }
    __syncthreads();
    const int g = m.g;
    const long long n_cells = (long long)g * g * g;
    const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // g^3 is a multiple of 128
    if (cell >= n_cells) return;
    const int cx = (int)(cell % g), cy = (int)((cell / g) % g), cz = (int)(cell / ((long long)g * g));
    const double px = (double)m.ox + (double)cx * (double)m.h;
    const double py = (double)m.oy + (double)cy * (double)m.h;
    const double pz = (double)m.oz + (double)cz * (double)m.h;
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
    if ((threadIdx.x & 31) == 0) occ[(size_t)conf * (n_cells / 32) + (cell >> 5)] = any;
}

// Fragment B for the cell-list screen: [conf][n_b] float4 {x, y, z, 0} in farthest-point order for the first
// `n_spread` positions (seed: the atom farthest from the centroid; then repeatedly the atom farthest from everything
// chosen so far), original order behind them.  The screen's result does not depend on the order (a minimum over all
// atoms); the order only decides how early a clashing pose is recognised.  One warp per conformer.
__global__ void __launch_bounds__(32) clash_order_b_kernel(const double* __restrict__ coords, int n_b, int n_spread,
                                                           float4* __restrict__ b_ord) {
    extern __shared__ float s_ob[];  // x[n_b], y[n_b], z[n_b], d[n_b], then int taken-order [n_b]
    float* sx = s_ob;
    float* sy = sx + n_b;
    float* sz = sy + n_b;
    float* sd = sz + n_b;
    int* sel = reinterpret_cast<int*>(sd + n_b);
    const int conf = blockIdx.x, lane = threadIdx.x;
    const double* src = coords + (size_t)conf * n_b * 3;
    float cx = 0.f, cy = 0.f, cz = 0.f;
    for (int i = lane; i < n_b; i += 32) {
        sx[i] = (float)src[3 * i];
        sy[i] = (float)src[3 * i + 1];
        sz[i] = (float)src[3 * i + 2];
        cx += sx[i];
        cy += sy[i];
        cz += sz[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cx += __shfl_xor_sync(0xffffffffu, cx, o);
        cy += __shfl_xor_sync(0xffffffffu, cy, o);
        cz += __shfl_xor_sync(0xffffffffu, cz, o);
    }
    // distance to the chosen set; first round: distance to the centroid
    float px = cx / (float)n_b, py = cy / (float)n_b, pz = cz / (float)n_b;
    for (int i = lane; i < n_b; i += 32) sd[i] = 3e38f;
    __syncwarp();
    // arg-max in ONE warp reduction: key = distance bits (a non-negative float orders like an unsigned) with the low 13
    // bits replaced by 8191 - index (ties and the lost mantissa bits do not matter for a visiting order)
    const int rounds = n_spread < n_b ? n_spread : n_b;
    for (int r = 0; r < rounds; ++r) {
        unsigned best = 0u;
        for (int i = lane; i < n_b; i += 32) {
            float d = sd[i];
            if (d >= 0.f) {  // not taken yet
                const float dx = sx[i] - px, dy = sy[i] - py, dz = sz[i] - pz;
                d = fminf(d, fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
                sd[i] = d;
                best = max(best, ((__float_as_uint(d) & ~8191u) | (unsigned)(8191 - i)) + 8192u);  // +8192: never 0
            }
        }
        best = __reduce_max_sync(0xffffffffu, best);
        const int best_i = 8191 - (int)(best & 8191u);
        if (lane == 0) {
            sel[r] = best_i;
            sd[best_i] = -1.f;  // taken
        }
        px = sx[best_i];
        py = sy[best_i];
        pz = sz[best_i];
        __syncwarp();
        if (r == 0) {  // from now on: distance to the chosen atoms only
            for (int i = lane; i < n_b; i += 32)
                if (sd[i] >= 0.f) sd[i] = 3e38f;
            __syncwarp();
        }
    }
    if (lane == 0) {  // the rest in original order
        int w = rounds;
        for (int i = 0; i < n_b; ++i)
            if (sd[i] >= 0.f) sel[w++] = i;
    }
    __syncwarp();
    for (int i = lane; i < n_b; i += 32) {
        const int k = sel[i];
        b_ord[(size_t)conf * n_b + i] = make_float4(sx[k], sy[k], sz[k], 0.f);
    }
}

// pose formats of the cell-list screen and the FP64 recheck
enum { kPoseXf64 = 0, kPoseQ7 = 1 };

// compact pose (fc_clash_screen_pose7_dev): q = (x, y, z, w) any non-zero quaternion, t; the rotation is
//   R = I + (2 / |q|^2) * [[-(yy+zz), xy-zw, xz+yw], [xy+zw, -(xx+zz), yz-xw], [xz-yw, yz+xw, -(xx+yy)]]
// FP64 expansion with every operation rounded on its own (no contraction), the order the header documents, so a host
// restatement in plain numpy produces the same doubles.
__device__ __forceinline__ void pose7_expand_f64(const float* __restrict__ p, double* r) {
    const double x = (double)p[0], y = (double)p[1], z = (double)p[2], w = (double)p[3];
    const double xx = __dmul_rn(x, x), yy = __dmul_rn(y, y), zz = __dmul_rn(z, z), ww = __dmul_rn(w, w);
    const double n = __dadd_rn(__dadd_rn(__dadd_rn(xx, yy), zz), ww);
    const double s = __ddiv_rn(2.0, n);
    const double xy = __dmul_rn(x, y), xz = __dmul_rn(x, z), yz = __dmul_rn(y, z);
    const double xw = __dmul_rn(x, w), yw = __dmul_rn(y, w), zw = __dmul_rn(z, w);
    r[0] = __dsub_rn(1.0, __dmul_rn(s, __dadd_rn(yy, zz)));
    r[1] = __dmul_rn(s, __dsub_rn(xy, zw));
    r[2] = __dmul_rn(s, __dadd_rn(xz, yw));
    r[3] = __dmul_rn(s, __dadd_rn(xy, zw));
    r[4] = __dsub_rn(1.0, __dmul_rn(s, __dadd_rn(xx, zz)));
    r[5] = __dmul_rn(s, __dsub_rn(yz, xw));
    r[6] = __dmul_rn(s, __dsub_rn(xz, yw));
    r[7] = __dmul_rn(s, __dadd_rn(yz, xw));
    r[8] = __dsub_rn(1.0, __dmul_rn(s, __dadd_rn(xx, yy)));
    r[9] = (double)p[4];
    r[10] = (double)p[5];
    r[11] = (double)p[6];
}

template <int FMT>
struct RawPose {  // a pose as it sits in memory, converted to float
    float v[FMT == kPoseXf64 ? 12 : 7];
};

template <int FMT>
__device__ __forceinline__ void pose_load_raw(const void* __restrict__ poses, long long pose, RawPose<FMT>& raw) {
    if (FMT == kPoseXf64) {
        const double2* x = reinterpret_cast<const double2*>(reinterpret_cast<const double*>(poses) + pose * 12);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const double2 v = __ldg(x + k);
            raw.v[2 * k] = (float)v.x;
            raw.v[2 * k + 1] = (float)v.y;
        }
    } else {
        const float* q = reinterpret_cast<const float*>(poses) + pose * 7;
#pragma unroll
        for (int k = 0; k < 7; ++k) raw.v[k] = __ldg(q + k);
    }
}

template <int FMT>
__device__ __forceinline__ void pose_raw_to_f32(const RawPose<FMT>& raw, float* r) {
    if (FMT == kPoseXf64) {
#pragma unroll
        for (int k = 0; k < 12; ++k) r[k] = raw.v[k];
    } else {
        const float x = raw.v[0], y = raw.v[1], z = raw.v[2], w = raw.v[3];
        const float xx = x * x, yy = y * y, zz = z * z;
        const float s = 2.0f / (xx + yy + zz + w * w);
        const float xy = x * y, xz = x * z, yz = y * z, xw = x * w, yw = y * w, zw = z * w;
        r[0] = 1.f - s * (yy + zz);
        r[1] = s * (xy - zw);
        r[2] = s * (xz + yw);
        r[3] = s * (xy + zw);
        r[4] = 1.f - s * (xx + zz);
        r[5] = s * (yz - xw);
        r[6] = s * (xz - yw);
        r[7] = s * (yz + xw);
        r[8] = 1.f - s * (xx + yy);
        r[9] = raw.v[4];
        r[10] = raw.v[5];
        r[11] = raw.v[6];
    }
}

// a pose a level could not decide, handed to the next level with everything that level needs (64 bytes, four
// LDG.128): the box-frame transform, the running minimum (box units) and the FP32 band (A^2)
struct __align__(16) CellItem {
    unsigned pose;
    float dmin2;
    unsigned spare;
    float band;
    float q[12];
};

struct CellArgs {
    const float4* a_u;
    const float* a_rad;
    const float4* b_ord;   // [conf][n_b] farthest-point order
    const float* b_rad;
    const uint4* grid;
    const unsigned* occ;
    const CellMeta* meta;
    const void* poses;     // (n_poses, 12) f64 or (n_poses, 7) f32
    const int4* tiles;
    long long n_tiles, n_poses;
    int n_a, n_b;
    int j_lo, j_hi;        // atoms of B this level visits
    int last;              // no level behind this one
    float thr2;
    int count_mode;
    const CellItem* in_list;   // null: level 0, item i is pose i
    const unsigned* in_count;
    CellItem* out_list;
    unsigned* out_count;
    uint8_t* status;       // may be null
    unsigned* bits;        // may be null: survivor bitmask, bit (p & 31) of word p >> 5
    int* unc_count;
    UncEntry* unc_list;
};

struct CellConst {  // per-axis scale / magic constants and field masks of the cell number
    float sx, sy, sz, mx, my, mz;
    unsigned fx, fy, fz;
};

// box coordinates of atom b under the box-frame transform q, clamped into the box by the .SAT of the last FFMA
#define FC_CELL_U(b)                                                                              \
    const float ux = __saturatef(fmaf(q[0], (b).x, fmaf(q[1], (b).y, fmaf(q[2], (b).z, q[9]))));  \
    const float uy = __saturatef(fmaf(q[3], (b).x, fmaf(q[4], (b).y, fmaf(q[5], (b).z, q[10])))); \
    const float uz = __saturatef(fmaf(q[6], (b).x, fmaf(q[7], (b).y, fmaf(q[8], (b).z, q[11]))))

// phase 1 for one block of <= 32 atoms: bit k of the result = atom j0 + k landed in a cell with candidates.
// SMEM: the occupancy words of the pose's conformer are the CTA's shared-memory copy (byte offset = (cell >> 3) & mask);
// otherwise they are read from global memory through a pointer that already discounts the magic constant's high bits.
template <bool SMEM>
__device__ __forceinline__ unsigned cell_flag_block(const float (&q)[12], const float4* __restrict__ bt, int jn,
                                                    const CellConst& k, const unsigned* __restrict__ oc_global,
                                                    const unsigned* s_occ_words, unsigned off_mask) {
    unsigned flagged = 0u;
#pragma unroll 4
    for (int a = 0; a < jn; ++a) {
        const float4 b = __ldg(bt + a);
        FC_CELL_U(b);
        // x keeps the bits of 1.5 * 2^23 above its field; y and z are merged in by bitwise select
        unsigned cell = __float_as_uint(fmaf(ux, k.sx, k.mx));
        cell = (cell & ~k.fy) | (__float_as_uint(fmaf(uy, k.sy, k.my)) & k.fy);
        cell = (cell & ~k.fz) | (__float_as_uint(fmaf(uz, k.sz, k.mz)) & k.fz);
        unsigned word;
        if (SMEM) {
            word = *reinterpret_cast<const unsigned*>(reinterpret_cast<const char*>(s_occ_words) + ((cell >> 3) & off_mask));
        } else {
            word = __ldg(oc_global + (cell >> 5));
        }
        flagged = __funnelshift_r(flagged, __funnelshift_r(word, 0u, cell), 1);  // bit (cell & 31) of word -> top of flagged
    }
    return flagged >> (32 - jn);
}

constexpr int kCellThreads = 256;

template <int FMT>
__global__ void __launch_bounds__(kCellThreads, 3) clash_cell_kernel(CellArgs p) {
    extern __shared__ unsigned s_occ[];  // occupancy bits of the cached conformer (g^3 / 32 words)
    __shared__ int s_want;
    const CellMeta m = *p.meta;
    const int g = m.g, sh = m.shift;
    const unsigned lane = threadIdx.x & 31u;
    const bool first = p.in_list == nullptr;
    const long long n_items = first ? p.n_poses : (long long)*p.in_count;
    // cell numbers in disjoint mantissa fields (x: bits [0,sh), y: [sh,2sh), z: [2sh,3sh)); x rounds to the nearest
    // lattice point; the y and z fields truncate, so half a cell is added to them
    CellConst kc;
    kc.sx = (float)(g - 1);
    kc.sy = (float)((g - 1) << sh);
    kc.sz = (float)((g - 1) << (2 * sh));
    kc.mx = kCellMagic;
    kc.my = kCellMagic + (float)(1 << (sh - 1));
    kc.mz = kCellMagic + (float)(1 << (2 * sh - 1));
    kc.fx = (unsigned)(g - 1);
    kc.fy = kc.fx << sh;
    kc.fz = kc.fx << (2 * sh);
    const size_t n_cells = (size_t)g * g * g;
    const unsigned n_words = (unsigned)(n_cells / 32);
    const unsigned off_mask = (n_words - 1u) << 2;  // byte offset of a word inside the shared copy
    const float u2 = m.inv_span * m.inv_span;  // A^2 -> box units
    const float thr2u = p.thr2 * u2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long cta0 = (long long)blockIdx.x * blockDim.x;
    int cached = -1;  // conformer of A whose occupancy bits sit in shared memory (CTA-uniform)
    if (!p.tiles) {   // one conformer pair for every pose
        for (unsigned w = threadIdx.x; w < n_words; w += blockDim.x) s_occ[w] = __ldg(p.occ + w);
        cached = 0;
        __syncthreads();
    }
    RawPose<FMT> nxt;
#pragma unroll
    for (int k = 0; k < (int)(sizeof(nxt.v) / 4); ++k) nxt.v[k] = 0.f;
    if (first && cta0 + threadIdx.x < n_items) pose_load_raw<FMT>(p.poses, cta0 + threadIdx.x, nxt);
    for (long long cta_base = cta0; cta_base < n_items; cta_base += stride) {
        const long long base = cta_base + (threadIdx.x & ~31u);
        const long long item = base + lane;
        const bool valid = item < n_items;
        long long pose = item;
        float dmin2 = 3.0e38f, band = 0.f;
        float q[12];
        RawPose<FMT> cur = nxt;
        if (first) {  // the next item's pose travels while this one is screened
            if (item + stride < n_items) pose_load_raw<FMT>(p.poses, item + stride, nxt);
        } else if (valid) {
            const float4* it = reinterpret_cast<const float4*>(p.in_list + item);
            const float4 h0 = it[0], h1 = it[1], h2 = it[2], h3 = it[3];
            pose = __float_as_uint(h0.x);
            dmin2 = h0.y;
            band = h0.w;
            q[0] = h1.x; q[1] = h1.y; q[2] = h1.z; q[3] = h1.w;
            q[4] = h2.x; q[5] = h2.y; q[6] = h2.z; q[7] = h2.w;
            q[8] = h3.x; q[9] = h3.y; q[10] = h3.z; q[11] = h3.w;
        }
        int conf_a = 0, conf_b = 0;
        bool covered = valid;
        if (p.tiles) {
            if (valid) {  // tiles are sorted by first pose: binary search
                long long lo = 0, hi = p.n_tiles - 1;
                while (lo < hi) {
                    long long mid = (lo + hi + 1) >> 1;
                    if ((long long)p.tiles[mid].z <= pose) lo = mid;
                    else hi = mid - 1;
                }
                const int4 t = p.tiles[lo];
                covered = pose >= (long long)t.z && pose < (long long)t.z + t.w;  // else: pose not covered by any tile
                conf_a = t.x;
                conf_b = t.y;
            }
            // the CTA keeps the occupancy bits of ONE conformer of A in shared memory: the one its first item uses
            // (poses are ordered by tile, so a batch rarely mixes conformers; the others read global memory)
            if (threadIdx.x == 0) s_want = covered ? conf_a : cached;
            __syncthreads();  // also: every warp is done with the previous batch's shared-memory reads
            const int want = s_want;
            if (want != cached) {
                const unsigned* src = p.occ + (size_t)want * n_words;
                for (unsigned w = threadIdx.x; w < n_words; w += blockDim.x) s_occ[w] = __ldg(src + w);
                cached = want;
                __syncthreads();
            }
        }
        bool push = false, pass = false;
        if (covered) {
            if (first) {
                float r[12];
                pose_raw_to_f32<FMT>(cur, r);
                const float tnorm = sqrtf(fmaf(r[9], r[9], fmaf(r[10], r[10], r[11] * r[11])));
                // same band as the all-pairs kernel (the difference form is at least as accurate as the Gram form)
                const float ext = p.a_rad[conf_a] + p.b_rad[conf_b] + tnorm;
                band = fmaf(1.5e-6f * ext, ext, 1e-6f);
#pragma unroll
                for (int k = 0; k < 9; ++k) q[k] = r[k] * m.inv_span;
                q[9] = (r[9] - m.ox) * m.inv_span;
                q[10] = (r[10] - m.oy) * m.inv_span;
                q[11] = (r[11] - m.oz) * m.inv_span;
            }
            const float4* bt = p.b_ord + (size_t)conf_b * p.n_b;
            const float4* at = p.a_u + (size_t)conf_a * p.n_a;
            const uint4* gr = p.grid + (size_t)conf_a * n_cells;
            // phase 1 leaves the magic constant's high bits in its cell number (two LOP3s instead of three): the
            // constant is a multiple of 32 cells, i.e. a fixed number of words, taken off the table pointer here
            const unsigned* oc = p.occ + (size_t)conf_a * n_words - (size_t)(0x4B400000u >> 5);
            asm volatile("" : "+l"(oc));  // keep it a materialised pointer: word address = one IMAD.WIDE.U32
            const bool in_smem = conf_a == cached;
            // once the minimum is below this value the pose is decided (certain clash, or -- with max_clashes > 0 --
            // certain to need the FP64 count): the remaining atoms are skipped
            const float bandu = band * u2;
            const float settle = p.count_mode ? thr2u + bandu : thr2u - bandu;
            for (int j0 = p.j_lo; j0 < p.j_hi && !(dmin2 < settle); j0 += 32) {
                const int jn = min(32, p.j_hi - j0);
                // ---- phase 1: flag the atoms of B that land in a cell with candidates (no distances yet), so that
                //      the lanes of a warp do not wait for each other's candidate loops on every atom
                unsigned flagged = in_smem ? cell_flag_block<true>(q, bt + j0, jn, kc, oc, s_occ, off_mask)
                                           : cell_flag_block<false>(q, bt + j0, jn, kc, oc, s_occ, off_mask);
                // ---- phase 2: distances (box units) from the flagged atoms to the candidates of their cells
                while (flagged && !(dmin2 < settle)) {
                    const int k = __ffs(flagged) - 1;
                    flagged &= flagged - 1u;
                    const float4 b = __ldg(bt + j0 + k);
                    FC_CELL_U(b);
                    const unsigned cell = (__float_as_uint(fmaf(ux, kc.sx, kc.mx)) & kc.fx) |
                                          (__float_as_uint(fmaf(uy, kc.sy, kc.my)) & kc.fy) |
                                          (__float_as_uint(fmaf(uz, kc.sz, kc.mz)) & kc.fz);
                    const uint4 rec = __ldg(gr + cell);
                    const unsigned count = rec.x & 0xffu;
                    if (count == 255u) {  // crowded cell: all atoms of A
                        for (int i = 0; i < p.n_a; ++i) {
                            const float4 a = __ldg(at + i);
                            const float dx = a.x - ux, dy = a.y - uy, dz = a.z - uz;
                            dmin2 = fminf(dmin2, fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
                        }
                        continue;
                    }
                    // candidates four at a time (independent loads); unused index bytes are 0 = atom 0, a real atom of A,
                    // so measuring it as well cannot make the minimum wrong
                    unsigned long long q0 = (((unsigned long long)rec.y << 32) | rec.x) >> 8;  // indices 0..6
                    unsigned long long q1 = ((unsigned long long)rec.w << 32) | rec.z;         // indices 7..14
                    q0 |= q1 << 56;  // indices 0..7
                    q1 >>= 8;        // indices 8..14
                    for (unsigned c = 0; c < count; c += 4) {
                        const unsigned i4 = (unsigned)q0;
                        q0 = (q0 >> 32) | (q1 << 32);
                        q1 >>= 32;
                        const float4 a0 = __ldg(at + (i4 & 0xffu)), a1 = __ldg(at + ((i4 >> 8) & 0xffu));
                        const float4 a2 = __ldg(at + ((i4 >> 16) & 0xffu)), a3 = __ldg(at + (i4 >> 24));
                        float dx = a0.x - ux, dy = a0.y - uy, dz = a0.z - uz;
                        const float d0 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                        dx = a1.x - ux, dy = a1.y - uy, dz = a1.z - uz;
                        const float d1 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                        dx = a2.x - ux, dy = a2.y - uy, dz = a2.z - uz;
                        const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                        dx = a3.x - ux, dy = a3.y - uy, dz = a3.z - uz;
                        const float d3 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                        dmin2 = fminf(fminf(dmin2, fminf(d0, d1)), fminf(d2, d3));
                    }
                }
            }
            if (!p.last && !(dmin2 < settle)) {
                push = true;  // undecided so far: the next level continues with the remaining atoms
            } else {
                uint8_t st;
                bool uncertain;
                if (p.count_mode) {
                    uncertain = !(dmin2 > thr2u + bandu);
                    st = FC_STATUS_PASS;
                } else {
                    uncertain = fabsf(dmin2 - thr2u) <= bandu;
                    st = dmin2 > thr2u ? FC_STATUS_PASS : 0;
                }
                // the candidate lists only cover thresh + kCellPad: a band that large cannot be trusted to them
                if (band > kCellPad * sqrtf(p.thr2)) uncertain = true;
                if (uncertain) {
                    int slot = atomicAdd(p.unc_count, 1);
                    UncEntry e;
                    e.pose = pose;
                    e.conf_a = conf_a;
                    e.conf_b = conf_b;
                    p.unc_list[slot] = e;
                }
                if (p.status) p.status[pose] = st;
                pass = (st & FC_STATUS_PASS) && !uncertain;  // the FP64 recheck sets the bit of an undecided pose
            }
        }
        // survivors of this level, compacted (one atomic per warp)
        const unsigned pmask = __ballot_sync(0xffffffffu, push);
        if (pmask) {
            unsigned slot0 = 0;
            if (lane == 0) slot0 = atomicAdd(p.out_count, (unsigned)__popc(pmask));
            slot0 = __shfl_sync(0xffffffffu, slot0, 0);
            if (push) {
                float4* it = reinterpret_cast<float4*>(p.out_list + slot0 + __popc(pmask & ((1u << lane) - 1u)));
                it[0] = make_float4(__uint_as_float((unsigned)pose), dmin2, 0.f, band);
                it[1] = make_float4(q[0], q[1], q[2], q[3]);
                it[2] = make_float4(q[4], q[5], q[6], q[7]);
                it[3] = make_float4(q[8], q[9], q[10], q[11]);
            }
        }
        if (p.bits) {
            if (first) {  // level 0 owns the words: 32 consecutive poses per warp
                const unsigned word = __ballot_sync(0xffffffffu, pass);
                if (lane == 0 && base < n_items) p.bits[base >> 5] = word;
            } else if (pass) {
                atomicOr(p.bits + (pose >> 5), 1u << (pose & 31));
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// FP64 recheck: one warp per undecided pose, reference arithmetic (utils.py:544-575)
// ---------------------------------------------------------------------------------------------
struct RecheckArgs {
    const double* a_coords;
    const double* b_coords;
    const void* poses;  // (n_poses, 12) f64 or (n_poses, 7) f32
    int n_a, n_b;
    double thresh;
    int max_clashes;
    int strict;
    const int* unc_count;
    const UncEntry* unc_list;
    uint8_t* status;    // may be null
    unsigned* bits;     // may be null: survivor bitmask (the FP32 pass left the bit of an undecided pose at 0)
    float* min_dist;
    int* near_count;
    long long* near_idx;
    double* near_dist;
    long long near_cap;
    long long pose_base;
    int* recheck_total;  // may be null: running number of rechecked poses (host-buffer entry points)
};

// One CTA per undecided pose: the placed atoms of B go to shared memory once, warp w takes atoms w, w + 8, ... of B and
// its lanes the atoms of A, so a pose costs (n_b / 8) x (n_a / 32) square roots per thread instead of n_a x n_b / 32.
template <int FMT>
__global__ void __launch_bounds__(256) clash_recheck_f64_kernel(RecheckArgs p) {
    extern __shared__ double s_placed[];  // n_b * 3
    __shared__ int s_cl[8];
    __shared__ double s_dm[8], s_cz[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_unc = *p.unc_count;
    if (p.recheck_total && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(p.recheck_total, n_unc);
    for (int u = blockIdx.x; u < n_unc; u += gridDim.x) {
        UncEntry e = p.unc_list[u];
        const double* a = p.a_coords + (size_t)e.conf_a * p.n_a * 3;
        const double* b = p.b_coords + (size_t)e.conf_b * p.n_b * 3;
        double r[12];
        if (FMT == kPoseXf64) {
            const double* x = reinterpret_cast<const double*>(p.poses) + e.pose * 12;
#pragma unroll
            for (int k = 0; k < 12; ++k) r[k] = x[k];
        } else {
            pose7_expand_f64(reinterpret_cast<const float*>(p.poses) + e.pose * 7, r);
        }
        __syncthreads();  // previous pose's placed atoms are no longer read
        for (int j = threadIdx.x; j < p.n_b; j += blockDim.x) {
            double bx = b[3 * j], by = b[3 * j + 1], bz = b[3 * j + 2];
            // (R @ b) + t, the get_embed expression embeds.py:815-817
            s_placed[3 * j] = (r[0] * bx + r[1] * by + r[2] * bz) + r[9];
            s_placed[3 * j + 1] = (r[3] * bx + r[4] * by + r[5] * bz) + r[10];
            s_placed[3 * j + 2] = (r[6] * bx + r[7] * by + r[8] * bz) + r[11];
        }
        __syncthreads();
        int clashes = 0;
        double dmin = 1e300, closest = 1e300;
        for (int i = lane; i < p.n_a; i += 32) {
            const double ax = a[3 * i], ay = a[3 * i + 1], az = a[3 * i + 2];
            for (int j = warp; j < p.n_b; j += 8) {
                double dx = s_placed[3 * j] - ax, dy = s_placed[3 * j + 1] - ay, dz = s_placed[3 * j + 2] - az;
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
            s_cl[warp] = clashes;
            s_dm[warp] = dmin;
            s_cz[warp] = closest;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < 8; ++w) {
                clashes += s_cl[w];
                dmin = fmin(dmin, s_dm[w]);
                closest = fmin(closest, s_cz[w]);
            }
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
            if (p.status) p.status[e.pose] = st;
            if (p.bits && (st & FC_STATUS_PASS)) atomicOr(p.bits + (e.pose >> 5), 1u << (e.pose & 31));
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
    const float4* a_u = nullptr;    // cell-list tables (cell_g > 0): atoms in box coordinates
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
        for (int g_try : {64, 32}) {  // powers of two: the cell number of an axis is a mantissa bit field
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
    size_t total = off_occ + (cell_g ? (size_t)n_conf_a * (n_cells / 32) * 4 : 0);
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
        float4* a_u = (float4*)(p->buf + off_axyz);
        uint4* grid_tab = (uint4*)(p->buf + off_grid);
        unsigned* occ = (unsigned*)(p->buf + off_occ);
        clash_bbox_kernel<<<1, 256, 0, s>>>(a_coords, (long long)n_conf_a * n_a, (float)thresh, cell_g, meta);
        dim3 gg((unsigned)((n_cells + 127) / 128), (unsigned)n_conf_a);
        clash_grid_kernel<<<gg, 128, (size_t)n_a * 24, s>>>(a_coords, n_a, meta, a_u, grid_tab, occ);
        p->a_u = a_u;
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

// test hook (host pointers): geometry of the cell grid the screen builds for this ensemble and threshold
extern "C" int fc_clash_cell_meta(const double* a_coords, int n_conf_a, int n_a, double thresh, float* out8) {
    FC_REQUIRE(a_coords && out8 && n_conf_a > 0 && n_a > 0, "fc_clash_cell_meta: null pointer");
    for (int k = 0; k < 8; ++k) out8[k] = 0.f;
    double* d_a = nullptr;
    const size_t bytes = (size_t)n_conf_a * n_a * 24;
    FC_CUDA(cudaMalloc((void**)&d_a, bytes));
    cudaError_t e = cudaMemcpy(d_a, a_coords, bytes, cudaMemcpyHostToDevice);
    fc_clash_prep* prep = nullptr;
    int rc = e == cudaSuccess ? fc_clash_prepare_dev(d_a, n_conf_a, n_a, thresh, 1, &prep, nullptr)
                              : cuda_fail(e, "cudaMemcpy", __FILE__, __LINE__);
    if (rc == FC_OK && prep->cell_g) {
        CellMeta m;
        e = cudaMemcpy(&m, prep->meta, sizeof(m), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpy", __FILE__, __LINE__);
        out8[0] = m.ox;
        out8[1] = m.oy;
        out8[2] = m.oz;
        out8[3] = m.h;
        out8[4] = (float)m.g;
        out8[5] = sqrtf(m.rc2);
        out8[6] = m.inv_span;
        out8[7] = (float)m.shift;
    }
    if (prep) fc_clash_prep_free(prep, nullptr);
    cudaDeviceSynchronize();
    cudaFree(d_a);
    return rc;
}

extern "C" void fc_clash_prep_free(fc_clash_prep* p, void* stream) {
    if (!p) return;
    if (p->buf) cudaFreeAsync(p->buf, (cudaStream_t)stream);
    delete p;
}

namespace fc {

// B-side tables of a screen: all-pairs layout (padded float4), bounding radii and the farthest-point ordered copy
// the cell-list levels read
struct ClashBTabs {
    unsigned char* buf = nullptr;
    const float4* b_tab = nullptr;
    const float* b_rad = nullptr;
    const float4* b_ord = nullptr;
    int n_b_pad = 0;
};

static int clash_prepare_b(const double* b_coords, int n_conf_b, int n_b, ClashBTabs* out, cudaStream_t s) {
    ClashGeom g = choose_geom(n_b);
    FC_REQUIRE(g.tb > 0, "fc_clash_screen_dev: fragment B too large (%d atoms)", n_b);
    FC_REQUIRE(n_b <= 8191, "fc_clash_screen_dev: fragment B too large (%d atoms)", n_b);
    out->n_b_pad = g.chunks * g.tb;
    const size_t off_rb = (size_t)n_conf_b * out->n_b_pad * 16;
    const size_t off_ord = (off_rb + (size_t)n_conf_b * 4 + 15) / 16 * 16;
    const size_t total = off_ord + (size_t)n_conf_b * n_b * 16;
    FC_CUDA(cudaMallocAsync((void**)&out->buf, total, s));
    clash_prep_kernel<<<n_conf_b, 128, 0, s>>>(b_coords, n_conf_b, n_b, out->n_b_pad, 0, (float4*)out->buf,
                                               (float*)(out->buf + off_rb));
    if ((size_t)n_b * 20 > 48 * 1024)
        cudaFuncSetAttribute(clash_order_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, n_b * 20);
    clash_order_b_kernel<<<n_conf_b, 32, (size_t)n_b * 20, s>>>(b_coords, n_b, 32, (float4*)(out->buf + off_ord));
    out->b_tab = (const float4*)out->buf;
    out->b_rad = (const float*)(out->buf + off_rb);
    out->b_ord = (const float4*)(out->buf + off_ord);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        cudaFreeAsync(out->buf, s);
        out->buf = nullptr;
        return cuda_fail(e, "clash_prepare_b kernels", __FILE__, __LINE__);
    }
    return FC_OK;
}

static void clash_free_b(ClashBTabs* t, cudaStream_t s) {
    if (t->buf) cudaFreeAsync(t->buf, s);
    t->buf = nullptr;
}

__global__ void pose7_expand_kernel(const float* __restrict__ pose7, long long n, double* __restrict__ xf) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double r[12];
    pose7_expand_f64(pose7 + i * 7, r);
#pragma unroll
    for (int k = 0; k < 12; ++k) xf[i * 12 + k] = r[k];
}

struct ScreenIO {
    const void* poses = nullptr;
    int fmt = kPoseXf64;
    int64_t n_poses = 0;
    const int32_t* tiles = nullptr;
    int64_t n_tiles = 0;
    int max_clashes = 0, strict = 1;
    uint8_t* status = nullptr;
    uint32_t* bits = nullptr;
    float* min_dist = nullptr;
    int32_t* near_count = nullptr;
    int64_t* near_idx = nullptr;
    double* near_dist = nullptr;
    int64_t near_cap = 0;
    int64_t pose_index_base = 0;
    int32_t* recheck_total = nullptr;
};

template <int FMT>
static int cell_grid_blocks(int sms, size_t smem) {
    static int per_sm = 0;
    static size_t for_smem = 0;
    if (per_sm == 0 || for_smem != smem) {
        cudaFuncSetAttribute(clash_cell_kernel<FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, clash_cell_kernel<FMT>, kCellThreads, smem) != cudaSuccess || n <= 0) n = 2;
        per_sm = n;
        for_smem = smem;
    }
    return sms * per_sm;
}

// level boundaries over the (re-ordered) atoms of B: FC_CLASH_LEVELS="32,64" overrides
static int cell_levels(int n_b, int* bounds /* [<= 8] */) {
    int n = 0;
    bounds[n++] = 0;
    const char* v = getenv("FC_CLASH_LEVELS");
    if (v && *v) {
        while (*v && n < 6) {
            int b = atoi(v);
            if (b > bounds[n - 1] && b < n_b) bounds[n++] = b;
            while (*v && *v != ',') ++v;
            if (*v == ',') ++v;
        }
    } else {
        if (n_b > 48) bounds[n++] = 32;  // measured on B200 (C3): two levels beat one (2.39 -> 1.81 ms) and three (1.93 ms)
    }
    bounds[n] = n_b;
    return n;  // number of levels
}

static int clash_screen_core(const fc_clash_prep* prep, const double* a_coords, const double* b_coords, int n_conf_b,
                             int n_b, const ClashBTabs& bt, const ScreenIO& io, cudaStream_t s) {
    const int n_a = prep->n_a;
    const double thresh = prep->thresh;
    const int64_t n_poses = io.n_poses;
    if (n_poses == 0) return FC_OK;
    const int sms = sm_count();  // also configures the stream-ordered memory pool on first use
    ClashGeom g = choose_geom(n_b);
    const int n_a_pad = prep->n_a_pad;
    const int n_b_pad = bt.n_b_pad;
    int64_t n_tiles = io.n_tiles;
    if (!io.tiles) n_tiles = (n_poses + g.poses - 1) / g.poses;
    FC_REQUIRE(n_tiles > 0, "fc_clash_screen_dev: empty tile list");

    // ---- path: all atom pairs (Gram-form FP32 kernel) or cell lists over fragment A -----------------
    int cell_g = 0;
    {
        // FC_CLASH_MODE: 0 = all pairs, 1 = cell lists whenever possible, unset = automatic
        const char* v = getenv("FC_CLASH_MODE");
        const int mode = (v && *v) ? (atoi(v) ? 1 : 0) : 2;
        const bool possible = !io.min_dist && prep->cell_g > 0 && n_poses < ((int64_t)1 << 31);
        const bool wanted = mode == 1 || (mode == 2 && n_poses >= 16384);
        if (possible && wanted) cell_g = prep->cell_g;
    }

    // stream-ordered scratch: undecided-pose list, counters, (cell path) the two item lists of the levels,
    // (all-pairs path) expanded transforms of compact poses and status bytes when only the bitmask is wanted
    const size_t off_cnt = 0;
    const size_t off_unc = 64;
    size_t off_l0 = off_unc + (size_t)n_poses * sizeof(UncEntry);
    size_t off_l1 = off_l0, off_xf = off_l0, off_st = off_l0, total = off_l0;
    const bool need_xf = !cell_g && io.fmt == kPoseQ7;
    const bool need_st = !cell_g && !io.status;
    if (cell_g) {
        off_l1 = off_l0 + (size_t)n_poses * sizeof(CellItem);
        total = off_l1 + (size_t)n_poses * sizeof(CellItem);
    } else {
        off_st = off_xf + (need_xf ? (size_t)n_poses * 96 : 0);
        total = off_st + (need_st ? ((size_t)n_poses + 15) / 16 * 16 : 0);
    }
    unsigned char* scratch = nullptr;
    FC_CUDA(cudaMallocAsync((void**)&scratch, total, s));
    FC_CUDA(cudaMemsetAsync(scratch + off_cnt, 0, 64, s));
    int* unc_count = (int*)(scratch + off_cnt);
    UncEntry* unc_list = (UncEntry*)(scratch + off_unc);
    uint8_t* status = io.status ? io.status : (need_st ? scratch + off_st : nullptr);
    const void* poses_f64 = io.poses;  // what the FP64 recheck reads (format io.fmt unless expanded below)
    int recheck_fmt = io.fmt;
    cudaError_t e = cudaSuccess;

    if (cell_g) {
        CellArgs c;
        c.a_u = prep->a_u;
        c.a_rad = prep->a_rad;
        c.b_ord = bt.b_ord;
        c.b_rad = bt.b_rad;
        c.grid = prep->grid;
        c.occ = prep->occ;
        c.meta = prep->meta;
        c.poses = io.poses;
        c.tiles = (const int4*)io.tiles;
        c.n_tiles = n_tiles;
        c.n_poses = n_poses;
        c.n_a = n_a;
        c.n_b = n_b;
        c.thr2 = (float)(thresh * thresh);
        c.count_mode = io.max_clashes > 0;
        c.status = status;
        c.bits = io.bits;
        c.unc_count = unc_count;
        c.unc_list = unc_list;
        int bounds[8];
        const int n_levels = cell_levels(n_b, bounds);
        CellItem* lists[2] = {(CellItem*)(scratch + off_l0), (CellItem*)(scratch + off_l1)};
        unsigned* counters = (unsigned*)(scratch + off_cnt) + 4;  // one per level, zeroed above
        const size_t cell_smem = (size_t)cell_g * cell_g * cell_g / 8;  // the occupancy bits of one conformer
        const int full_grid = io.fmt == kPoseQ7 ? cell_grid_blocks<kPoseQ7>(sms, cell_smem) : cell_grid_blocks<kPoseXf64>(sms, cell_smem);
        timing_begin(s);
        for (int lv = 0; lv < n_levels; ++lv) {
            c.j_lo = bounds[lv];
            c.j_hi = bounds[lv + 1];
            c.last = lv == n_levels - 1;
            c.in_list = lv ? lists[(lv - 1) & 1] : nullptr;
            c.in_count = lv ? counters + (lv - 1) : nullptr;
            c.out_list = lists[lv & 1];
            c.out_count = counters + lv;
            int grid = full_grid;
            if (lv == 0) grid = (int)std::min<long long>(grid, (n_poses + kCellThreads - 1) / kCellThreads);
            if (io.fmt == kPoseQ7) clash_cell_kernel<kPoseQ7><<<grid, kCellThreads, cell_smem, s>>>(c);
            else clash_cell_kernel<kPoseXf64><<<grid, kCellThreads, cell_smem, s>>>(c);
        }
        timing_end(s);
        e = cudaGetLastError();
    } else {
        if (need_xf) {
            pose7_expand_kernel<<<(unsigned)((n_poses + 255) / 256), 256, 0, s>>>((const float*)io.poses, n_poses,
                                                                                 (double*)(scratch + off_xf));
            poses_f64 = scratch + off_xf;
            recheck_fmt = kPoseXf64;
        }
        size_t smem = (size_t)(n_a_pad + 2) * 16 + (size_t)n_b_pad * 16 + (size_t)2 * g.poses * 96 + 16 + (size_t)g.poses * 4;
        if (smem > 227 * 1024) {
            cudaFreeAsync(scratch, s);
            FC_REQUIRE(false, "fc_clash_screen_dev: fragments need %zu B of shared memory", smem);
        }
        ClashArgs a;
        a.a_tab = prep->a_tab;
        a.a_rad = prep->a_rad;
        a.b_tab = bt.b_tab;
        a.b_rad = bt.b_rad;
        a.xf = (const double*)poses_f64;
        a.tiles = (const int4*)io.tiles;
        a.n_tiles = n_tiles;
        a.n_poses = n_poses;
        a.n_a_pad = n_a_pad;
        a.n_b_pad = n_b_pad;
        a.chunks = g.chunks;
        a.poses = g.poses;
        a.thr2 = (float)(thresh * thresh);
        a.count_mode = io.max_clashes > 0;
        {
            const char* v = getenv("FC_CLASH_TMA");
            a.use_tma = ((reinterpret_cast<uintptr_t>(a.xf) & 15) == 0) && !(v && atoi(v) == 0);
        }
        a.status = status;
        a.min_dist = io.min_dist;
        a.unc_count = unc_count;
        a.unc_list = unc_list;
        // persistent grid: resident CTAs per SM follow from the register/thread budget
        int ctas_per_sm = tb_max_threads(g.tb) / g.threads;
        if (ctas_per_sm > 8) ctas_per_sm = 8;
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        long long grid_ll = (long long)sms * ctas_per_sm;
        if (grid_ll > n_tiles) grid_ll = n_tiles;
        int grid = (int)grid_ll;
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
    }
    if (e != cudaSuccess) {
        cudaFreeAsync(scratch, s);
        return cuda_fail(e, "clash screen kernel launch", __FILE__, __LINE__);
    }

    RecheckArgs r;
    r.a_coords = a_coords;
    r.b_coords = b_coords;
    r.poses = poses_f64;
    r.n_a = n_a;
    r.n_b = n_b;
    r.thresh = thresh;
    r.max_clashes = io.max_clashes;
    r.strict = io.strict;
    r.unc_count = unc_count;
    r.unc_list = unc_list;
    r.status = status;
    r.bits = cell_g ? io.bits : nullptr;  // the all-pairs path packs the final status bytes below
    r.min_dist = io.min_dist;
    r.near_count = io.near_count;
    r.near_idx = (long long*)io.near_idx;
    r.near_dist = io.near_dist;
    r.near_cap = io.near_cap;
    r.pose_base = io.pose_index_base;
    r.recheck_total = io.recheck_total;
    {
        const size_t rsmem = (size_t)n_b * 24;
        if (rsmem > 48 * 1024) {
            cudaFuncSetAttribute(clash_recheck_f64_kernel<kPoseQ7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem);
            cudaFuncSetAttribute(clash_recheck_f64_kernel<kPoseXf64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem);
        }
        if (recheck_fmt == kPoseQ7) clash_recheck_f64_kernel<kPoseQ7><<<sms * 4, 256, rsmem, s>>>(r);
        else clash_recheck_f64_kernel<kPoseXf64><<<sms * 4, 256, rsmem, s>>>(r);
    }
    e = cudaGetLastError();
    int rc = FC_OK;
    if (e == cudaSuccess && !cell_g && io.bits) rc = fc_pack_mask_dev(status, n_poses, io.bits, (void*)s);
    cudaError_t e2 = cudaFreeAsync(scratch, s);
    if (e != cudaSuccess) return cuda_fail(e, "clash_recheck_f64_kernel launch", __FILE__, __LINE__);
    if (rc != FC_OK) return rc;
    if (e2 != cudaSuccess) return cuda_fail(e2, "cudaFreeAsync", __FILE__, __LINE__);
    return FC_OK;
}

}  // namespace fc

extern "C" int fc_clash_screen_ex_dev(const fc_clash_prep* prep, const double* a_coords, const double* b_coords,
                                      int n_conf_b, int n_b, const void* poses, int pose_format, int64_t n_poses,
                                      const int32_t* tiles, int64_t n_tiles, int max_clashes, int strict,
                                      uint8_t* status, uint32_t* bits, float* min_dist, int32_t* near_count,
                                      int64_t* near_idx, double* near_dist, int64_t near_cap, int64_t pose_index_base,
                                      int32_t* recheck_count, void* stream) {
    FC_REQUIRE(prep, "fc_clash_screen_ex_dev: null preparation");
    FC_REQUIRE(n_b > 0 && n_conf_b > 0, "fc_clash_screen_dev: empty fragment");
    FC_REQUIRE(n_poses >= 0 && max_clashes >= 0, "fc_clash_screen_dev: negative size");
    FC_REQUIRE(pose_format == FC_POSE_XF64 || pose_format == FC_POSE_Q7, "fc_clash_screen_ex_dev: unknown pose format %d",
               pose_format);
    if (n_poses == 0) return FC_OK;
    FC_REQUIRE(a_coords && b_coords && poses && (status || bits), "fc_clash_screen_dev: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    ClashBTabs bt;
    int rc = clash_prepare_b(b_coords, n_conf_b, n_b, &bt, s);
    if (rc) return rc;
    ScreenIO io;
    io.poses = poses;
    io.fmt = pose_format == FC_POSE_Q7 ? kPoseQ7 : kPoseXf64;
    io.n_poses = n_poses;
    io.tiles = tiles;
    io.n_tiles = n_tiles;
    io.max_clashes = max_clashes;
    io.strict = strict;
    io.status = status;
    io.bits = bits;
    io.min_dist = min_dist;
    io.near_count = near_count;
    io.near_idx = near_idx;
    io.near_dist = near_dist;
    io.near_cap = near_cap;
    io.pose_index_base = pose_index_base;
    io.recheck_total = recheck_count;
    rc = clash_screen_core(prep, a_coords, b_coords, n_conf_b, n_b, bt, io, s);
    clash_free_b(&bt, s);
    return rc;
}

extern "C" int fc_clash_screen_prepared_dev(const fc_clash_prep* prep, const double* a_coords, const double* b_coords,
                                            int n_conf_b, int n_b, const double* xf, int64_t n_poses,
                                            const int32_t* tiles, int64_t n_tiles, int max_clashes, int strict,
                                            uint8_t* status, float* min_dist, int32_t* near_count, int64_t* near_idx,
                                            double* near_dist, int64_t near_cap, int64_t pose_index_base, void* stream) {
    FC_REQUIRE(n_poses <= 0 || status, "fc_clash_screen_dev: null pointer");
    return fc_clash_screen_ex_dev(prep, a_coords, b_coords, n_conf_b, n_b, xf, FC_POSE_XF64, n_poses, tiles, n_tiles,
                                  max_clashes, strict, status, nullptr, min_dist, near_count, near_idx, near_dist, near_cap,
                                  pose_index_base, nullptr, stream);
}

static int clash_want_cells(int64_t n_poses, const float* min_dist) {
    const char* v = getenv("FC_CLASH_MODE");
    const int mode = (v && *v) ? (atoi(v) ? 1 : 0) : 2;
    return !min_dist && (mode == 1 || (mode == 2 && n_poses >= 16384));
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
    int rc = fc_clash_prepare_dev(a_coords, n_conf_a, n_a, thresh, clash_want_cells(n_poses, min_dist), &prep, stream);
    if (rc) return rc;
    rc = fc_clash_screen_prepared_dev(prep, a_coords, b_coords, n_conf_b, n_b, xf, n_poses, tiles, n_tiles, max_clashes,
                                      strict, status, min_dist, near_count, near_idx, near_dist, near_cap,
                                      pose_index_base, stream);
    fc_clash_prep_free(prep, stream);
    return rc;
}

// ------------------------------------------------------------------------------------------------
// host-buffer screen of compact poses: chunked, double-buffered H2D(pose7) -> screen -> D2H(bitmask)
// ------------------------------------------------------------------------------------------------
extern "C" int fc_clash_batch_pose7(const double* a_coords, int n_conf_a, int n_a, const double* b_coords, int n_conf_b,
                                    int n_b, const float* pose7, int64_t n_poses, const int32_t* tiles, int64_t n_tiles,
                                    double thresh, int max_clashes, int strict, uint32_t* bits_out, uint8_t* status_out,
                                    int64_t* counts, int64_t* near_idx, double* near_dist, int64_t near_cap) {
    FC_REQUIRE(n_a > 0 && n_b > 0 && n_conf_a > 0 && n_conf_b > 0, "fc_clash_batch_pose7: empty fragment");
    FC_REQUIRE(n_poses >= 0 && max_clashes >= 0, "fc_clash_batch_pose7: negative size");
    if (counts) counts[0] = counts[1] = counts[2] = 0;
    if (n_poses == 0) return FC_OK;
    FC_REQUIRE(a_coords && b_coords && pose7 && bits_out, "fc_clash_batch_pose7: null pointer");
    const int tile_poses = fc_clash_tile_poses(n_b);
    FC_REQUIRE(tile_poses > 0, "fc_clash_batch_pose7: fragment B too large (%d atoms)", n_b);
    if (tiles) {
        for (int64_t t = 0; t < n_tiles; ++t) {
            const int32_t* q = tiles + 4 * t;
            FC_REQUIRE(q[0] >= 0 && q[0] < n_conf_a && q[1] >= 0 && q[1] < n_conf_b && q[3] >= 0 && q[3] <= tile_poses &&
                           q[2] >= 0 && (int64_t)q[2] + q[3] <= n_poses,
                       "fc_clash_batch_pose7: tile %lld out of range", (long long)t);
        }
    }
    const int64_t n_words = (n_poses + 31) / 32;
    // grow-only pinned staging for what comes back (per host thread): bitmask words, then status bytes
    static thread_local unsigned char* h_stage = nullptr;
    static thread_local size_t h_stage_cap = 0;
    const size_t stage_need = 64 + (size_t)n_words * 4 + (status_out ? (size_t)n_poses : 0);
    int64_t n_pass = 0;
    if (h_stage_cap < stage_need) {
        if (h_stage) cudaFreeHost(h_stage);
        h_stage = nullptr;
        h_stage_cap = 0;
        size_t want = std::max<size_t>(stage_need, (size_t)1 << 20);
        cudaError_t he = cudaHostAlloc((void**)&h_stage, want, cudaHostAllocDefault);
        if (he != cudaSuccess) return cuda_fail(he, "cudaHostAlloc(result staging)", __FILE__, __LINE__);
        h_stage_cap = want;
    }
    int32_t* h_cnt = (int32_t*)h_stage;
    uint32_t* h_bits = (uint32_t*)(h_stage + 64);
    uint8_t* h_status = status_out ? h_stage + 64 + (size_t)n_words * 4 : nullptr;

    const int kBuf = 2;
    // streams and events of this host thread live as long as the thread does (creating them costs ~0.1 ms per call)
    static thread_local cudaStream_t tl_st[kBuf] = {nullptr, nullptr};
    static thread_local cudaEvent_t tl_ev[kBuf + 1] = {nullptr, nullptr, nullptr};
    static thread_local int tl_dev = -1;
    cudaStream_t st[kBuf] = {nullptr, nullptr};
    cudaEvent_t ev_setup = nullptr, ev_done[kBuf] = {nullptr, nullptr};
    float* d_pose[kBuf] = {nullptr, nullptr};
    uint32_t* d_bits = nullptr;
    uint8_t* d_status = nullptr;
    double *d_a = nullptr, *d_b = nullptr, *d_near_dist = nullptr;
    int32_t *d_tiles = nullptr, *d_cnt = nullptr;
    int64_t* d_near_idx = nullptr;
    fc_clash_prep* prep = nullptr;
    ClashBTabs bt;
    int rc = FC_OK;
    // with an explicit tile list the whole batch is one chunk (tiles index absolute poses); chunks are multiples of
    // 32 poses so that every chunk owns whole bitmask words (FC_CLASH_CHUNK overrides the default of 512 k poses =
    // 14 MB per copy; measured on B200 at 10 M poses: 256 k 6.5 ms, 512 k 5.7 ms, 1 M 5.8 ms, 2 M 6.8 ms per call)
    int64_t chunk_env = 0;
    if (const char* v = getenv("FC_CLASH_CHUNK")) chunk_env = atoll(v);
    int64_t chunk = tiles ? n_poses : std::min<int64_t>(n_poses, chunk_env > 0 ? chunk_env : (int64_t)1 << 19);
    chunk = (chunk + 31) / 32 * 32;

#define FC_TRY(call)                                                 \
    do {                                                             \
        cudaError_t _e = (call);                                     \
        if (_e != cudaSuccess) {                                     \
            rc = cuda_fail(_e, #call, __FILE__, __LINE__);           \
            goto done;                                               \
        }                                                            \
    } while (0)

    const bool trace = getenv("FC_CLASH_TRACE") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    double t_setup = 0, t_issued = 0, t_synced = 0;
    sm_count();  // configures the memory pool on first use
    {
        int dev = 0;
        FC_TRY(cudaGetDevice(&dev));
        if (tl_dev != dev) {  // first call of this thread, or the thread moved to another device
            for (int i = 0; i < kBuf; ++i) {
                if (tl_st[i]) cudaStreamDestroy(tl_st[i]);
                tl_st[i] = nullptr;
            }
            for (int i = 0; i < kBuf + 1; ++i) {
                if (tl_ev[i]) cudaEventDestroy(tl_ev[i]);
                tl_ev[i] = nullptr;
            }
            tl_dev = -1;
            for (int i = 0; i < kBuf; ++i) FC_TRY(cudaStreamCreateWithFlags(&tl_st[i], cudaStreamNonBlocking));
            for (int i = 0; i < kBuf + 1; ++i) FC_TRY(cudaEventCreateWithFlags(&tl_ev[i], cudaEventDisableTiming));
            tl_dev = dev;
        }
        for (int i = 0; i < kBuf; ++i) {
            st[i] = tl_st[i];
            ev_done[i] = tl_ev[i];
        }
        ev_setup = tl_ev[kBuf];
    }
    FC_TRY(cudaMallocAsync((void**)&d_a, (size_t)n_conf_a * n_a * 24, st[0]));
    FC_TRY(cudaMallocAsync((void**)&d_b, (size_t)n_conf_b * n_b * 24, st[0]));
    FC_TRY(cudaMallocAsync((void**)&d_bits, (size_t)n_words * 4, st[0]));
    if (status_out) FC_TRY(cudaMallocAsync((void**)&d_status, (size_t)n_poses, st[0]));
    FC_TRY(cudaMallocAsync((void**)&d_cnt, 16, st[0]));
    if (near_cap > 0) {
        FC_TRY(cudaMallocAsync((void**)&d_near_idx, (size_t)near_cap * 8, st[0]));
        FC_TRY(cudaMallocAsync((void**)&d_near_dist, (size_t)near_cap * 8, st[0]));
    }
    for (int i = 0; i < kBuf; ++i) FC_TRY(cudaMallocAsync((void**)&d_pose[i], (size_t)chunk * 28, st[0]));
    FC_TRY(cudaMemcpyAsync(d_a, a_coords, (size_t)n_conf_a * n_a * 24, cudaMemcpyHostToDevice, st[0]));
    FC_TRY(cudaMemcpyAsync(d_b, b_coords, (size_t)n_conf_b * n_b * 24, cudaMemcpyHostToDevice, st[0]));
    FC_TRY(cudaMemsetAsync(d_cnt, 0, 16, st[0]));
    if (tiles) {
        FC_TRY(cudaMallocAsync((void**)&d_tiles, (size_t)n_tiles * 16, st[0]));
        FC_TRY(cudaMemcpyAsync(d_tiles, tiles, (size_t)n_tiles * 16, cudaMemcpyHostToDevice, st[0]));
    }
    // both fragments' tables once for the whole batch
    rc = fc_clash_prepare_dev(d_a, n_conf_a, n_a, thresh, clash_want_cells(n_poses, nullptr), &prep, (void*)st[0]);
    if (rc) goto done;
    rc = clash_prepare_b(d_b, n_conf_b, n_b, &bt, st[0]);
    if (rc) goto done;
    FC_TRY(cudaEventRecord(ev_setup, st[0]));
    for (int i = 1; i < kBuf; ++i) FC_TRY(cudaStreamWaitEvent(st[i], ev_setup, 0));
    t_setup = now();
    {
        // chunk c rides stream c % 2.  While the copy engine moves the next chunks' poses the host thread drains a finished
        // chunk: bitmask words (and status bytes) pinned staging -> caller's buffers, and the pass count
        auto drain = [&](int64_t first, int64_t n) {
            const int64_t w0 = first / 32, nw = (n + 31) / 32;
            for (int64_t w = w0; w < w0 + nw; ++w) n_pass += __builtin_popcount(h_bits[w]);
            memcpy(bits_out + w0, h_bits + w0, (size_t)nw * 4);
            if (status_out) memcpy(status_out + first, h_status + first, (size_t)n);
        };
        int64_t pend_first[kBuf] = {-1, -1}, pend_n[kBuf] = {0, 0};
        int b = 0;
        for (int64_t first = 0; first < n_poses; first += chunk, b = (b + 1) % kBuf) {
            const int64_t n = std::min<int64_t>(chunk, n_poses - first);
            const int64_t w0 = first / 32, nw = (n + 31) / 32;
            cudaStream_t s = st[b];
            if (pend_first[b] >= 0) {  // the chunk that used this stream two steps ago
                FC_TRY(cudaEventSynchronize(ev_done[b]));
                drain(pend_first[b], pend_n[b]);
                pend_first[b] = -1;
            }
            FC_TRY(cudaMemcpyAsync(d_pose[b], pose7 + first * 7, (size_t)n * 28, cudaMemcpyHostToDevice, s));
            ScreenIO io;
            io.poses = d_pose[b];
            io.fmt = kPoseQ7;
            io.n_poses = n;
            io.tiles = d_tiles;
            io.n_tiles = n_tiles;
            io.max_clashes = max_clashes;
            io.strict = strict;
            io.status = d_status ? d_status + first : nullptr;
            io.bits = d_bits + w0;
            io.near_count = d_cnt;
            io.near_idx = d_near_idx;
            io.near_dist = d_near_dist;
            io.near_cap = near_cap;
            io.pose_index_base = first;
            io.recheck_total = d_cnt + 1;
            rc = clash_screen_core(prep, d_a, d_b, n_conf_b, n_b, bt, io, s);
            if (rc != FC_OK) goto done;
            FC_TRY(cudaMemcpyAsync(h_bits + w0, d_bits + w0, (size_t)nw * 4, cudaMemcpyDeviceToHost, s));
            if (d_status) FC_TRY(cudaMemcpyAsync(h_status + first, d_status + first, (size_t)n, cudaMemcpyDeviceToHost, s));
            FC_TRY(cudaEventRecord(ev_done[b], s));
            pend_first[b] = first;
            pend_n[b] = n;
        }
        t_issued = now();
        for (int k = 0; k < kBuf; ++k, b = (b + 1) % kBuf) {  // remaining chunks in issue order
            if (pend_first[b] < 0) continue;
            FC_TRY(cudaEventSynchronize(ev_done[b]));
            drain(pend_first[b], pend_n[b]);
        }
    }
    t_synced = now();
    {
        int32_t* cnt = h_cnt;
        FC_TRY(cudaMemcpyAsync(cnt, d_cnt, 16, cudaMemcpyDeviceToHost, st[0]));
        FC_TRY(cudaStreamSynchronize(st[0]));
        const int64_t n_copy = std::min<int64_t>(cnt[0], near_cap);
        if (n_copy > 0 && near_idx) FC_TRY(cudaMemcpy(near_idx, d_near_idx, (size_t)n_copy * 8, cudaMemcpyDeviceToHost));
        if (n_copy > 0 && near_dist) FC_TRY(cudaMemcpy(near_dist, d_near_dist, (size_t)n_copy * 8, cudaMemcpyDeviceToHost));
        if (counts) {
            counts[0] = n_pass;
            counts[1] = cnt[1];
            counts[2] = cnt[0];
        }
    }
    if (trace)
        fprintf(stderr, "fc_clash_batch_pose7: setup %.2f ms, issue %.2f ms, wait %.2f ms, readback %.2f ms\n", t_setup - t_begin,
                t_issued - t_setup, t_synced - t_issued, now() - t_synced);
done:
    if (st[0]) {
        for (int i = 0; i < kBuf; ++i)
            if (st[i]) cudaStreamSynchronize(st[i]);
        clash_free_b(&bt, st[0]);
        fc_clash_prep_free(prep, (void*)st[0]);
        void* bufs[] = {d_pose[0], d_pose[1], d_a, d_b, d_bits, d_status, d_tiles, d_cnt, d_near_idx, d_near_dist};
        for (void* q : bufs)
            if (q) cudaFreeAsync(q, st[0]);
    }
    return rc;
#undef FC_TRY
}
