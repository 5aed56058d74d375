// firecode_b200 -- cyclical embed, bimolecular path.
//
// Reference: firecode/embeds.py:588-750 `_fast_bimol_rigid_cyclical_embed` (reached through
// `cyclical_embed`, embeds.py:180-185, for two molecules and for "chelotropic" embeds).
//   group  = (conformer pair, pivot pair, polygon orientation v) that passed the norm / pairing
//            filters (enumerated on the host in the reference's loop order, O(conformers*pivots));
//   pose   = group x angle tuple of embedder.systematic_angles.
// Per group and molecule the angle-independent part of the transform (alignment rotation A,
// rotation centre c, offset pos, step axis) is computed once (cyc_group_setup_kernel, FP64 Kabsch
// on the 2-vector covariance = algebra.py:28-49).  Per pose: S = rot(axis, angle), R = S A,
// t = c - S c + pos (embeds.py:694-709); the clash screen receives the transform of molecule 1
// relative to molecule 0.  In-loop similarity (embeds.py:723-729 -> utils.py:494-504): one warp
// per group walks its clash survivors in angle order and keeps a pose unless an earlier kept pose
// of the same group has all-atom, UNCENTRED Kabsch RMSD < 1 and max deviation < 2 (quirk N8).
#include <string.h>

#include <algorithm>
#include <vector>

#include "fc_embed.cuh"

namespace fc {

struct CycGroupXf {  // angle-independent transform data of one molecule in one group
    double a[9];     // alignment rotation
    double c[3];     // centre of the step rotation (A @ mean of reactive atoms)
    double pos[3];   // mean(vec_pair) - A @ pivot.meanpoint
    double axis[3];  // axis of the step rotation
};

struct CycDev {
    const double* coords[2];
    int n_conf[2], n_atoms[2], n_react[2];
    const long long* reactive[2];
    long long n_groups;
    const int* g_conf;       // (G, 2)
    const double* g_pivot;   // (G, 2, 3)
    const double* g_mean;    // (G, 2, 3)
    const double* g_vecs;    // (G, 2, 2, 3)
    const double* g_dirs;    // (G, 2, 3)
    const double* angles;    // (A, 2)
    int n_angles;
    int handed;
    CycGroupXf* gx;          // (G, 2)
};

// embeds.py:657-709, everything that does not depend on the angle
__global__ void cyc_group_setup_kernel(CycDev p) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n_groups * 2) return;
    long long g = i >> 1;
    int m = (int)(i & 1);
    int conf = p.g_conf[2 * g + m];
    const double* x = p.coords[m] + (size_t)conf * p.n_atoms[m] * 3;
    const double* pivot = p.g_pivot + (2 * g + m) * 3;
    const double* mean = p.g_mean + (2 * g + m) * 3;
    const double* start = p.g_vecs + ((2 * g + m) * 2) * 3;
    const double* end = start + 3;
    const double* dir = p.g_dirs + (2 * g + m) * 3;
    // mean position of the reactive atoms
    double r0[3], r1[3] = {0, 0, 0}, apm[3];
    const double* a0 = x + 3 * p.reactive[m][0];
    r0[0] = a0[0]; r0[1] = a0[1]; r0[2] = a0[2];
    if (p.n_react[m] == 2) {
        const double* a1 = x + 3 * p.reactive[m][1];
        r1[0] = a1[0]; r1[1] = a1[1]; r1[2] = a1[2];
        // np.mean over two rows: (r0 + r1) / 2
        apm[0] = (r0[0] + r1[0]) / 2.0; apm[1] = (r0[1] + r1[1]) / 2.0; apm[2] = (r0[2] + r1[2]) / 2.0;
    } else {
        apm[0] = r0[0]; apm[1] = r0[1]; apm[2] = r0[2];
    }
    double md[3] = {mean[0] - apm[0], mean[1] - apm[1], mean[2] - apm[2]};
    if (md[0] == 0.0 && md[1] == 0.0 && md[2] == 0.0) { md[0] = mean[0]; md[1] = mean[1]; md[2] = mean[2]; }
    // align_vec_pair(ref = [end - start, direction], tgt = [pivot, mol_direction])
    double ref0[3] = {end[0] - start[0], end[1] - start[1], end[2] - start[2]};
    double h[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) h[3 * r + c] = ref0[r] * pivot[c] + dir[r] * md[c];
    M3 A = kabsch_from_cov(h, nullptr);
    CycGroupXf o;
    for (int k = 0; k < 9; ++k) o.a[k] = A.m[k];
    double ax_src[3];
    if (p.n_react[m] == 2) { ax_src[0] = r0[0] - r1[0]; ax_src[1] = r0[1] - r1[1]; ax_src[2] = r0[2] - r1[2]; }
    else { ax_src[0] = pivot[0]; ax_src[1] = pivot[1]; ax_src[2] = pivot[2]; }
    m3_apply(A, ax_src, o.axis);
    m3_apply(A, apm, o.c);
    double am[3];
    m3_apply(A, mean, am);
    // np.mean(vec_pair, axis=0) = (start + end) / 2
    o.pos[0] = (start[0] + end[0]) / 2.0 - am[0];
    o.pos[1] = (start[1] + end[1]) / 2.0 - am[1];
    o.pos[2] = (start[2] + end[2]) / 2.0 - am[2];
    p.gx[2 * g + m] = o;
}

// absolute transform of molecule m for (group, angle): R = S A, t = c - S c + pos
__device__ __forceinline__ void cyc_mol_xf(const CycDev& p, long long g, int m, int ai, M3& rot, double* t) {
    const CycGroupXf& q = p.gx[2 * g + m];
    double angle = p.angles[2 * ai + m];
    M3 S = rot_from_pointer(q.axis, angle, p.handed);
    M3 A;
    for (int k = 0; k < 9; ++k) A.m[k] = q.a[k];
    rot = m3_mul(S, A);
    double sc[3];
    m3_apply(S, q.c, sc);
    t[0] = q.c[0] - sc[0] + q.pos[0];
    t[1] = q.c[1] - sc[1] + q.pos[1];
    t[2] = q.c[2] - sc[2] + q.pos[2];
}

// transform of molecule 1 in the frame of molecule 0, for the clash screen
__global__ void cyc_pose_xf_kernel(CycDev p, long long n_poses, double* __restrict__ xf_rel) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_poses) return;
    long long g = i / p.n_angles;
    int ai = (int)(i - g * p.n_angles);
    M3 r0, r1;
    double t0[3], t1[3];
    cyc_mol_xf(p, g, 0, ai, r0, t0);
    cyc_mol_xf(p, g, 1, ai, r1, t1);
    M3 r0t = m3_transpose(r0);
    M3 rel = m3_mul(r0t, r1);
    double dt[3] = {t1[0] - t0[0], t1[1] - t0[1], t1[2] - t0[2]}, trel[3];
    m3_apply(r0t, dt, trel);
    double* o = xf_rel + i * 12;
    for (int k = 0; k < 9; ++k) o[k] = rel.m[k];
    o[9] = trel[0]; o[10] = trel[1]; o[11] = trel[2];
}

__device__ __forceinline__ void cyc_atom(const CycDev& p, const int* conf, const M3* rot, const double (*t)[3],
                                         int atom, double* out) {
    int m = atom < p.n_atoms[0] ? 0 : 1;
    int local = atom - (m ? p.n_atoms[0] : 0);
    const double* b = p.coords[m] + ((size_t)conf[m] * p.n_atoms[m] + local) * 3;
    const double* r = rot[m].m;
    out[0] = (r[0] * b[0] + r[1] * b[1] + r[2] * b[2]) + t[m][0];
    out[1] = (r[3] * b[0] + r[4] * b[1] + r[5] * b[2]) + t[m][1];
    out[2] = (r[6] * b[0] + r[7] * b[1] + r[8] * b[2]) + t[m][2];
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

struct CycSimArgs {
    CycDev p;
    const uint8_t* status;   // per pose
    uint8_t* keep;           // per pose, out
    double rmsd_thr, eps;
    TieRecord* ties;
    int* n_ties;
    int tie_cap;
};

__device__ __forceinline__ void push_tie(const CycSimArgs& a, long long pose, long long ref, double value, int kind,
                                         bool decision) {
    int slot = atomicAdd(a.n_ties, 1);
    if (slot < a.tie_cap) {
        TieRecord r;
        r.a = pose; r.b = ref; r.value = value; r.kind = kind; r.decision = decision ? 1 : 0;
        a.ties[slot] = r;
    }
}

// one warp per group: keep-first over the clash survivors in angle order
__global__ void __launch_bounds__(128) cyc_group_similarity_kernel(CycSimArgs a) {
    const CycDev& p = a.p;
    const int lane = threadIdx.x & 31;
    const long long g = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= p.n_groups) return;
    const int n_ang = p.n_angles;
    const int n_tot = p.n_atoms[0] + p.n_atoms[1];
    int conf[2] = {p.g_conf[2 * g], p.g_conf[2 * g + 1]};
    // accepted angle indices of this group live in a per-warp bitmask list (n_ang <= 1024)
    extern __shared__ int s_acc_all[];
    int* s_acc = s_acc_all + (threadIdx.x >> 5) * n_ang;
    int n_acc = 0;
    for (int ai = 0; ai < n_ang; ++ai) {
        const long long pose = g * n_ang + ai;
        if (!(a.status[pose] & FC_STATUS_PASS)) {
            if (lane == 0) a.keep[pose] = 0;
            continue;
        }
        M3 rp[2];
        double tp[2][3];
        cyc_mol_xf(p, g, 0, ai, rp[0], tp[0]);
        cyc_mol_xf(p, g, 1, ai, rp[1], tp[1]);
        bool similar = false;
        for (int k = 0; k < n_acc && !similar; ++k) {
            const int aj = s_acc[k];
            M3 rq[2];
            double tq[2][3];
            cyc_mol_xf(p, g, 0, aj, rq[0], tq[0]);
            cyc_mol_xf(p, g, 1, aj, rq[1], tq[1]);
            // H = p^T q over all atoms, uncentred (rmsd_and_max(center=False))
            double h[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            double gsum = 0.0;  // |p|^2 + |q|^2
            for (int at = lane; at < n_tot; at += 32) {
                double x[3], y[3];
                cyc_atom(p, conf, rp, tp, at, x);
                cyc_atom(p, conf, rq, tq, at, y);
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    gsum += x[r] * x[r] + y[r] * y[r];
#pragma unroll
                    for (int c = 0; c < 3; ++c) h[3 * r + c] += x[r] * y[c];
                }
            }
#pragma unroll
            for (int e = 0; e < 9; ++e) h[e] = warp_sum_d(h[e]);
            gsum = warp_sum_d(gsum);
            {   // closed-form singular values: clearly dissimilar pairs skip the Jacobi solve and the second pass
                const double lim = a.rmsd_thr + 1e-4;
                if ((gsum - 2.0 * singular_sum3(h)) / n_tot > lim * lim) continue;
            }
            M3 R = kabsch_from_cov(h, nullptr);
            double ss = 0.0, mx = 0.0;
            for (int at = lane; at < n_tot; at += 32) {
                double x[3], y[3];
                cyc_atom(p, conf, rp, tp, at, x);
                cyc_atom(p, conf, rq, tq, at, y);
                // diff = p @ R - q
                double dx = (x[0] * R.m[0] + x[1] * R.m[3] + x[2] * R.m[6]) - y[0];
                double dy = (x[0] * R.m[1] + x[1] * R.m[4] + x[2] * R.m[7]) - y[1];
                double dz = (x[0] * R.m[2] + x[1] * R.m[5] + x[2] * R.m[8]) - y[2];
                double d2 = dx * dx + dy * dy + dz * dz;
                ss += d2;
                mx = fmax(mx, d2);
            }
            ss = warp_sum_d(ss);
            mx = warp_max_d(mx);
            const double rmsd = sqrt(ss / n_tot), maxdev = sqrt(mx);
            const bool rm_ok = rmsd < a.rmsd_thr, md_ok = maxdev < 2.0 * a.rmsd_thr;
            if (lane == 0) {
                const long long ref = g * n_ang + aj;
                if (fabs(rmsd - a.rmsd_thr) <= a.eps) push_tie(a, pose, ref, rmsd, FC_TIE_RMSD, rm_ok);
                if (fabs(maxdev - 2.0 * a.rmsd_thr) <= a.eps) push_tie(a, pose, ref, maxdev, FC_TIE_MAXDEV, md_ok);
            }
            similar = rm_ok && md_ok;
        }
        if (!similar) {
            if (lane == 0) s_acc[n_acc] = ai;
            ++n_acc;
            __syncwarp();
        }
        if (lane == 0) a.keep[pose] = similar ? 0 : FC_STATUS_PASS;
    }
}

__global__ void cyc_materialize_kernel(CycDev p, const long long* __restrict__ kept, int n_kept,
                                       double* __restrict__ out) {
    int k = blockIdx.x;
    if (k >= n_kept) return;
    long long pose = kept[k];
    long long g = pose / p.n_angles;
    int ai = (int)(pose - g * p.n_angles);
    __shared__ M3 rot[2];
    __shared__ double t[2][3];
    if (threadIdx.x < 2) cyc_mol_xf(p, g, threadIdx.x, ai, rot[threadIdx.x], t[threadIdx.x]);
    __syncthreads();
    int conf[2] = {p.g_conf[2 * g], p.g_conf[2 * g + 1]};
    int n_tot = p.n_atoms[0] + p.n_atoms[1];
    double* o = out + (size_t)k * n_tot * 3;
    for (int at = threadIdx.x; at < n_tot; at += blockDim.x) {
        double v[3];
        cyc_atom(p, conf, rot, t, at, v);
        o[3 * at] = v[0]; o[3 * at + 1] = v[1]; o[3 * at + 2] = v[2];
    }
}

}  // namespace fc

using namespace fc;

// fc_result is defined in fc_embed.cu; the accessors used here are the public ones
struct fc_result;
namespace fc {
fc_result* result_new();
void result_set_cyclical(fc_result* r, int64_t n_poses, int64_t n_atoms, int64_t n_pass, std::vector<uint8_t>&& status,
                         std::vector<int64_t>&& kept, std::vector<double>&& coords, std::vector<int32_t>&& constrained,
                         int n_pairs, std::vector<fc_tie>&& ties, int64_t ties_total);
}  // namespace fc

extern "C" int fc_cyclical_screen(const fc_cyclical_problem* p, fc_result** out) {
    FC_REQUIRE(out, "null output");
    *out = nullptr;
    FC_REQUIRE(p, "null problem");
    FC_REQUIRE(p->n_mols == 2, "fc_cyclical_screen: only bimolecular embeds are built (n_mols = %d)", p->n_mols);
    FC_REQUIRE(p->n_angles > 0 && p->n_angles <= 4096, "bad angle count");
    for (int m = 0; m < 2; ++m) {
        FC_REQUIRE(p->coords[m] && p->n_conf[m] > 0 && p->n_atoms[m] > 0, "empty ensemble %d", m);
        FC_REQUIRE(p->reactive[m] && (p->n_reactive[m] == 1 || p->n_reactive[m] == 2), "molecule %d needs 1 or 2 reactive atoms", m);
        for (int k = 0; k < p->n_reactive[m]; ++k)
            FC_REQUIRE(p->reactive[m][k] >= 0 && p->reactive[m][k] < p->n_atoms[m], "reactive index out of range");
    }
    const int64_t G = p->n_groups, A = p->n_angles;
    const int64_t n_poses = G * A;
    FC_REQUIRE(G >= 0 && n_poses < ((int64_t)1 << 31), "too many poses for one call (%lld)", (long long)n_poses);
    const int64_t n_tot = (int64_t)p->n_atoms[0] + p->n_atoms[1];
    fc_result* r = result_new();
    if (G == 0) {
        result_set_cyclical(r, 0, n_tot, 0, {}, {}, {}, {}, p->n_pairs, {}, 0);
        *out = r;
        return FC_OK;
    }
    FC_REQUIRE(p->group_conf && p->group_pivot && p->group_mean && p->group_vecs && p->group_dirs && p->angles,
               "null group table");
    for (int64_t g = 0; g < G; ++g)
        for (int m = 0; m < 2; ++m)
            FC_REQUIRE(p->group_conf[2 * g + m] >= 0 && p->group_conf[2 * g + m] < p->n_conf[m], "group conformer out of range");
    sm_count();
    cudaStream_t s;
    {
        cudaError_t se = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
        if (se != cudaSuccess) {
            fc_result_free(r);
            return cuda_fail(se, "cudaStreamCreateWithFlags", __FILE__, __LINE__);
        }
    }
    int rc = FC_OK;
    std::vector<uint8_t> h_status, h_keep;
    std::vector<int64_t> kept;
    std::vector<double> coords;
    std::vector<int32_t> constrained;
    std::vector<fc_tie> ties;
    int64_t ties_total = 0, n_pass = 0;
    {
        DevBuf<double> d_coords[2], d_pivot, d_mean, d_vecs, d_dirs, d_angles, d_xf, d_out, d_near_dist;
        DevBuf<long long> d_react[2], d_kept;
        DevBuf<int> d_conf, d_cnt;
        DevBuf<CycGroupXf> d_gx;
        DevBuf<uint8_t> d_status, d_keep;
        DevBuf<int32_t> d_tiles, d_near_count;
        DevBuf<int64_t> d_near_idx;
        DevBuf<TieRecord> d_ties;
        const int tie_cap = 1 << 20, near_cap = 1 << 16;
        cudaError_t e = cudaSuccess;
#define CY(call) do { if (e == cudaSuccess) e = (call); } while (0)
        for (int m = 0; m < 2; ++m) {
            size_t n = (size_t)p->n_conf[m] * p->n_atoms[m] * 3;
            CY(d_coords[m].alloc(n, s));
            CY(cudaMemcpyAsync(d_coords[m].p, p->coords[m], n * 8, cudaMemcpyHostToDevice, s));
            CY(d_react[m].alloc(2, s));
            CY(cudaMemcpyAsync(d_react[m].p, p->reactive[m], (size_t)p->n_reactive[m] * 8, cudaMemcpyHostToDevice, s));
        }
        CY(d_conf.alloc((size_t)G * 2, s));
        CY(cudaMemcpyAsync(d_conf.p, p->group_conf, (size_t)G * 8, cudaMemcpyHostToDevice, s));
        CY(d_pivot.alloc((size_t)G * 6, s));
        CY(cudaMemcpyAsync(d_pivot.p, p->group_pivot, (size_t)G * 48, cudaMemcpyHostToDevice, s));
        CY(d_mean.alloc((size_t)G * 6, s));
        CY(cudaMemcpyAsync(d_mean.p, p->group_mean, (size_t)G * 48, cudaMemcpyHostToDevice, s));
        CY(d_vecs.alloc((size_t)G * 12, s));
        CY(cudaMemcpyAsync(d_vecs.p, p->group_vecs, (size_t)G * 96, cudaMemcpyHostToDevice, s));
        CY(d_dirs.alloc((size_t)G * 6, s));
        CY(cudaMemcpyAsync(d_dirs.p, p->group_dirs, (size_t)G * 48, cudaMemcpyHostToDevice, s));
        CY(d_angles.alloc((size_t)A * 2, s));
        CY(cudaMemcpyAsync(d_angles.p, p->angles, (size_t)A * 16, cudaMemcpyHostToDevice, s));
        CY(d_gx.alloc((size_t)G * 2, s));
        CY(d_xf.alloc((size_t)n_poses * 12, s));
        CY(d_status.alloc((size_t)n_poses, s));
        CY(d_keep.alloc((size_t)n_poses, s));
        CY(d_ties.alloc(tie_cap, s));
        CY(d_cnt.alloc(8, s));
        CY(cudaMemsetAsync(d_cnt.p, 0, 32, s));
        CY(d_near_count.alloc(4, s));
        CY(cudaMemsetAsync(d_near_count.p, 0, 16, s));
        CY(d_near_idx.alloc(near_cap, s));
        CY(d_near_dist.alloc(near_cap, s));
        if (e != cudaSuccess) rc = cuda_fail(e, "fc_cyclical_screen setup", __FILE__, __LINE__);

        CycDev d{};
        if (!rc) {
            for (int m = 0; m < 2; ++m) {
                d.coords[m] = d_coords[m].p;
                d.n_conf[m] = p->n_conf[m];
                d.n_atoms[m] = p->n_atoms[m];
                d.n_react[m] = p->n_reactive[m];
                d.reactive[m] = d_react[m].p;
            }
            d.n_groups = G;
            d.g_conf = d_conf.p; d.g_pivot = d_pivot.p; d.g_mean = d_mean.p; d.g_vecs = d_vecs.p; d.g_dirs = d_dirs.p;
            d.angles = d_angles.p;
            d.n_angles = (int)A;
            d.handed = p->rot_handedness >= 0 ? 1 : -1;
            d.gx = d_gx.p;
            cyc_group_setup_kernel<<<(unsigned)((G * 2 + 127) / 128), 128, 0, s>>>(d);
            cyc_pose_xf_kernel<<<(unsigned)((n_poses + 127) / 128), 128, 0, s>>>(d, n_poses, d_xf.p);
            e = cudaGetLastError();
            if (e != cudaSuccess) rc = cuda_fail(e, "cyclical transform kernels", __FILE__, __LINE__);
        }
        if (!rc) {
            // tiles: poses of one group share the conformer pair; consecutive groups often do too
            const int tp = fc_clash_tile_poses(p->n_atoms[1]);
            std::vector<int32_t> tiles;
            int64_t g = 0;
            while (g < G) {
                int64_t g_end = g + 1;
                while (g_end < G && p->group_conf[2 * g_end] == p->group_conf[2 * g] &&
                       p->group_conf[2 * g_end + 1] == p->group_conf[2 * g + 1])
                    ++g_end;
                for (int64_t pose = g * A; pose < g_end * A;) {
                    int cnt = (int)std::min<int64_t>(tp, g_end * A - pose);
                    tiles.push_back(p->group_conf[2 * g]);
                    tiles.push_back(p->group_conf[2 * g + 1]);
                    tiles.push_back((int32_t)pose);
                    tiles.push_back(cnt);
                    pose += cnt;
                }
                g = g_end;
            }
            e = d_tiles.alloc(tiles.size(), s);
            CY(cudaMemcpyAsync(d_tiles.p, tiles.data(), tiles.size() * 4, cudaMemcpyHostToDevice, s));
            if (e != cudaSuccess) rc = cuda_fail(e, "tile upload", __FILE__, __LINE__);
            if (!rc)
                rc = fc_clash_screen_dev(d_coords[0].p, p->n_conf[0], p->n_atoms[0], d_coords[1].p, p->n_conf[1],
                                         p->n_atoms[1], d_xf.p, n_poses, d_tiles.p, (int64_t)tiles.size() / 4, p->thresh,
                                         p->max_clashes, 1, d_status.p, nullptr, d_near_count.p, d_near_idx.p,
                                         d_near_dist.p, near_cap, 0, (void*)s);
            if (!rc) {
                e = cudaStreamSynchronize(s);  // the host tile vector must outlive the copy
                if (e != cudaSuccess) rc = cuda_fail(e, "clash screen", __FILE__, __LINE__);
            }
        }
        if (!rc) {
            CycSimArgs a{d, d_status.p, d_keep.p, p->rmsd_thresh, FC_NEAR_EPS, d_ties.p, d_cnt.p + 1, tie_cap};
            const int warps = 4;
            size_t smem = (size_t)warps * A * sizeof(int);
            cyc_group_similarity_kernel<<<(unsigned)((G + warps - 1) / warps), warps * 32, smem, s>>>(a);
            e = cudaGetLastError();
            h_status.resize((size_t)n_poses);
            h_keep.resize((size_t)n_poses);
            CY(cudaMemcpyAsync(h_status.data(), d_status.p, (size_t)n_poses, cudaMemcpyDeviceToHost, s));
            CY(cudaMemcpyAsync(h_keep.data(), d_keep.p, (size_t)n_poses, cudaMemcpyDeviceToHost, s));
            int h_cnt[2] = {0, 0}, n_near = 0;
            CY(cudaMemcpyAsync(h_cnt, d_cnt.p, 8, cudaMemcpyDeviceToHost, s));
            CY(cudaMemcpyAsync(&n_near, d_near_count.p, 4, cudaMemcpyDeviceToHost, s));
            CY(cudaStreamSynchronize(s));
            if (e != cudaSuccess) rc = cuda_fail(e, "similarity stage", __FILE__, __LINE__);
            if (!rc) {
                for (int64_t i = 0; i < n_poses; ++i) {
                    n_pass += h_status[(size_t)i] & FC_STATUS_PASS;
                    if (h_keep[(size_t)i]) kept.push_back(i);
                }
                n_near = std::min(n_near, near_cap);
                if (n_near > 0) {
                    std::vector<int64_t> idx(n_near);
                    std::vector<double> dist(n_near);
                    e = cudaMemcpy(idx.data(), d_near_idx.p, (size_t)n_near * 8, cudaMemcpyDeviceToHost);
                    CY(cudaMemcpy(dist.data(), d_near_dist.p, (size_t)n_near * 8, cudaMemcpyDeviceToHost));
                    for (int i = 0; i < n_near && e == cudaSuccess; ++i) {
                        fc_tie t;
                        t.a = idx[i]; t.b = -1; t.value = dist[i]; t.kind = FC_TIE_CLASH;
                        t.decision = (h_status[(size_t)idx[i]] & FC_STATUS_PASS) ? 0 : 1;
                        ties.push_back(t);
                    }
                    ties_total += n_near;
                }
                int n_t = std::min(h_cnt[1], tie_cap);
                ties_total += h_cnt[1];
                if (n_t > 0 && e == cudaSuccess) {
                    std::vector<TieRecord> tmp(n_t);
                    e = cudaMemcpy(tmp.data(), d_ties.p, (size_t)n_t * sizeof(TieRecord), cudaMemcpyDeviceToHost);
                    for (const TieRecord& t : tmp) {
                        fc_tie o;
                        o.a = t.a; o.b = t.b; o.value = t.value; o.kind = t.kind; o.decision = t.decision;
                        ties.push_back(o);
                    }
                }
                if (e != cudaSuccess) rc = cuda_fail(e, "tie readback", __FILE__, __LINE__);
            }
        }
        if (!rc && !kept.empty()) {
            const int n_kept = (int)kept.size();
            e = d_kept.alloc(n_kept, s);
            CY(d_out.alloc((size_t)n_kept * n_tot * 3, s));
            CY(cudaMemcpyAsync(d_kept.p, kept.data(), (size_t)n_kept * 8, cudaMemcpyHostToDevice, s));
            if (e == cudaSuccess) {
                cyc_materialize_kernel<<<n_kept, 128, 0, s>>>(d, d_kept.p, n_kept, d_out.p);
                e = cudaGetLastError();
            }
            coords.resize((size_t)n_kept * n_tot * 3);
            CY(cudaMemcpyAsync(coords.data(), d_out.p, coords.size() * 8, cudaMemcpyDeviceToHost, s));
            CY(cudaStreamSynchronize(s));
            if (e != cudaSuccess) rc = cuda_fail(e, "materialize", __FILE__, __LINE__);
            if (!rc && p->group_ids && p->n_pairs > 0) {
                constrained.resize((size_t)n_kept * p->n_pairs * 2);
                for (int k = 0; k < n_kept; ++k) {
                    int64_t g = kept[k] / A;
                    memcpy(&constrained[(size_t)k * p->n_pairs * 2], p->group_ids + g * p->n_pairs * 2,
                           (size_t)p->n_pairs * 8);
                }
            }
        }
#undef CY
    }
    cudaStreamSynchronize(s);
    cudaStreamDestroy(s);
    if (rc) {
        fc_result_free(r);
        return rc;
    }
    result_set_cyclical(r, n_poses, n_tot, n_pass, std::move(h_status), std::move(kept), std::move(coords),
                        std::move(constrained), p->n_pairs, std::move(ties), ties_total);
    *out = r;
    return FC_OK;
}

// ------------------------------------------------------------------------------------------------
// Group table of the bimolecular cyclical embed, host only (no CUDA call): the reference's loops
// embeds.py:596-641 -- conformer pairs (cartesian_product order: first index fastest), pivot pairs (first index
// fastest), two orientations -- minus pivot pairs whose norms differ by more than max_norm_delta (embeds.py:624) and
// arrangements that miss a user pairing (embeds.py:638-641; a pairing listed in `internal` counts as satisfied).
// Pivot tables are CSR over conformers (rows off[c] .. off[c+1]).  Writes at most `cap` groups, returns the number
// there are in *n_groups (call again with a larger cap if it exceeds it).
// ------------------------------------------------------------------------------------------------
extern "C" int fc_cyclical_groups(const int32_t* n_conf, const int64_t* off0, const double* vec0, const double* mean0,
                                  const int64_t* ids0, const int64_t* off1, const double* vec1, const double* mean1,
                                  const int64_t* ids1, double max_norm_delta, const int64_t* pairings, int32_t n_pairings,
                                  const int64_t* internal, int32_t n_internal, int64_t cap, int64_t* n_groups,
                                  int32_t* conf_out, double* pivot_out, double* mean_out, double* vecs_out, double* dirs_out,
                                  int32_t* ids_out) {
    FC_REQUIRE(n_conf && off0 && off1 && n_groups, "fc_cyclical_groups: null pointer");
    FC_REQUIRE(cap == 0 || (conf_out && pivot_out && mean_out && vecs_out && dirs_out && ids_out), "fc_cyclical_groups: null output");
    auto norm3 = [](const double* v) { return sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]); };  // np.linalg.norm(axis=1)
    int64_t g = 0;
    const int64_t n_cp = (int64_t)n_conf[0] * n_conf[1];
    for (int64_t cp = 0; cp < n_cp; ++cp) {
        const int64_t c0 = cp % n_conf[0], c1 = cp / n_conf[0];
        const int64_t k0 = off0[c0 + 1] - off0[c0], k1 = off1[c1 + 1] - off1[c1];
        for (int64_t p1 = 0; p1 < k1; ++p1)
            for (int64_t p0 = 0; p0 < k0; ++p0) {
                const int64_t r0 = off0[c0] + p0, r1 = off1[c1] + p1;
                const double l0 = norm3(vec0 + 3 * r0), l1 = norm3(vec1 + 3 * r1);
                if (fabs(l0 - l1) > max_norm_delta) continue;
                for (int v = 0; v < 2; ++v) {
                    // couples facing each other: swaps = [(0, 0), (0, 1)] (embeds.py:767)
                    const int64_t a0 = ids0[2 * r0], a1 = ids0[2 * r0 + 1];
                    const int64_t b0 = v ? ids1[2 * r1 + 1] : ids1[2 * r1], b1 = v ? ids1[2 * r1] : ids1[2 * r1 + 1];
                    bool ok = true;
                    for (int32_t q = 0; q < n_pairings && ok; ++q) {
                        const int64_t pa = pairings[2 * q], pb = pairings[2 * q + 1];
                        bool is_internal = false;
                        for (int32_t t = 0; t < n_internal; ++t) is_internal = is_internal || (internal[2 * t] == pa && internal[2 * t + 1] == pb);
                        if (is_internal) continue;
                        ok = (a0 == pa && b0 == pb) || (a1 == pa && b1 == pb);
                    }
                    if (!ok) continue;
                    if (g < cap) {
                        conf_out[2 * g] = (int32_t)c0;
                        conf_out[2 * g + 1] = (int32_t)c1;
                        for (int k = 0; k < 3; ++k) {
                            pivot_out[6 * g + k] = vec0[3 * r0 + k];
                            pivot_out[6 * g + 3 + k] = vec1[3 * r1 + k];
                            mean_out[6 * g + k] = mean0[3 * r0 + k];
                            mean_out[6 * g + 3 + k] = mean1[3 * r1 + k];
                        }
                        // polygonize for two lengths (utils.py:262-271): centred collinear segments on x; the second
                        // orientation multiplies segment 2 by -1 (its zeros become -0.0, as numpy's in-place product does)
                        double* vv = vecs_out + 12 * g;
                        for (int k = 0; k < 12; ++k) vv[k] = 0.0;
                        vv[0] = -l0 / 2;
                        vv[3] = l0 / 2;
                        vv[6] = -l1 / 2;
                        vv[9] = l1 / 2;
                        if (v)
                            for (int k = 6; k < 12; ++k) vv[k] *= -1.0;
                        const double d[6] = {0.0, 1.0, 0.0, 0.0, -1.0, 0.0};
                        for (int k = 0; k < 6; ++k) dirs_out[6 * g + k] = d[k];
                        ids_out[4 * g] = (int32_t)a0;
                        ids_out[4 * g + 1] = (int32_t)b0;
                        ids_out[4 * g + 2] = (int32_t)a1;
                        ids_out[4 * g + 3] = (int32_t)b1;
                    }
                    ++g;
                }
            }
    }
    *n_groups = g;
    return FC_OK;
}

