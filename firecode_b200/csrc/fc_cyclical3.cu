// firecode_b200 -- cyclical embed, trimolecular body.
//
// Reference: firecode/embeds.py:409-585 (`cyclical_embed` for three molecules) with its closures
// `_get_directions` (embeds.py:188-254) and `_adjust_directions` (embeds.py:256-407).
//   super-group = (conformer triple, pivot triple) whose pivot norms form a triangle
//                 (embeds.py:447-462); enumerated on the host (C++) in the reference's loop order;
//   group       = (super-group, polygon orientation v in 0..7) that passes the pairing filter
//                 (embeds.py:473-476);
//   pose        = group x angle triple of embedder.systematic_angles.
// Per super-group one warp restates the STATEFUL direction search (quirk N6): the circumcentre
// directions of `_get_directions`, then for every active orientation in order the 343-point grid
// search of `_adjust_directions` (7 x 7 x 7 rotations of the reactive-atom positions of CONFORMER 0
// about the triangle sides, cost = sum of three orbital-alignment angles, first minimum wins), whose
// result seeds the next orientation.  The same warp then stores the angle-independent transform of
// each molecule (alignment rotation, step-rotation axis / centre, offset: embeds.py:494-554).
// Clash test: utils.py:553-575, three blocks (m2,m1), (m3,m2), (m1,m3) with `<=` and max_clashes = 0
// => a pose passes iff every block is clash-free; block (i,j) depends only on (angle_i, angle_j), so
// it is screened once per distinct angle pair (fc_clash_screen_dev on relative transforms).
// In-loop similarity (embeds.py:564-569 -> utils.py:494-504): one warp per group, keep-first over
// the clash survivors in angle order, all-atom UNCENTRED Kabsch RMSD < 1 and max deviation < 2.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <functional>
#include <map>
#include <thread>
#include <utility>
#include <vector>

#include "fc_embed.cuh"

namespace fc {

struct Cyc3Xf {      // angle-independent transform data of one molecule in one group
    double a[9];     // alignment rotation
    double c[3];     // centre of the step rotation (A @ mean of reactive atoms)
    double pos[3];   // mean(vec_pair) - A @ pivot.meanpoint
    double axis[3];  // axis of the step rotation
};

struct Cyc3Super {   // (conformer triple, pivot triple)
    int conf[3];
    int piv[3];      // row in the per-molecule pivot tables
    int first_group;
    int active;      // bit v: orientation v passes the pairing filter
};

struct Cyc3Group {
    int super, v;
    int r[6];        // r01, r02, r10, r12, r20, r21: reactive atom of molecule m facing the partner
};

struct Cyc3Dev {
    const double* coords[3];
    int n_conf[3], n_atoms[3], n_react[3];
    const long long* reactive[3];
    const double* pvec[3];
    const double* pmean[3];
    const double* pnorm[3];
    const Cyc3Super* supers;
    int n_supers;
    const Cyc3Group* groups;
    int n_groups;
    const double* angles;  // (A, 3)
    int n_angles;
    int handed;
    Cyc3Xf* gx;            // (G, 3)
    int* choice;           // (G)
    double* gap;           // (G)
};

// ---- planar triangle helpers (utils.py:252-312, embeds.py:198-209) -------------------------------
__device__ __forceinline__ void tri_vertex2(const double* n, double& x, double& y) {
    double a = n[0] * n[0], b = n[1] * n[1], c = n[2] * n[2];
    x = (a - b + c) / (2.0 * sqrt(a));
    y = sqrt(c - x * x);
}

// start / end of side m for orientation v (polygonize, swap table utils.py:293-310)
__device__ __forceinline__ void poly3(const double* n, int v, int m, double* start, double* end) {
    double x, y;
    tri_vertex2(n, x, y);
    double pts[3][3] = {{0.0, 0.0, 0.0}, {n[0], 0.0, 0.0}, {x, y, 0.0}};
    const int swapmask[8] = {0, 4, 2, 6, 1, 3, 5, 7};
    int s = m, e = (m + 1) % 3;
    if ((swapmask[v] >> m) & 1) { int t = s; s = e; e = t; }
    for (int k = 0; k < 3; ++k) { start[k] = pts[s][k]; end[k] = pts[e][k]; }
}

__device__ __forceinline__ double vec_angle_deg(const double* a, const double* b, int dim) {
    const double kPi = 3.14159265358979323846;
    double na = 0.0, nb = 0.0;
    for (int k = 0; k < dim; ++k) { na += a[k] * a[k]; nb += b[k] * b[k]; }
    na = sqrt(na); nb = sqrt(nb);
    double d = 0.0;
    for (int k = 0; k < dim; ++k) d += (a[k] / na) * (b[k] / nb);
    d = fmin(1.0, fmax(-1.0, d));
    return acos(d) * 180.0 / kPi;
}

// 2D circumcentre directions before sign / normalisation; returns true if one of them is exactly 0
__device__ __forceinline__ bool raw_dirs(const double* n, double d[3][2], double vtx[3][2]) {
    double x, y;
    tri_vertex2(n, x, y);
    vtx[0][0] = 0.0; vtx[0][1] = 0.0; vtx[1][0] = n[0]; vtx[1][1] = 0.0; vtx[2][0] = x; vtx[2][1] = y;
    double a = vtx[1][0], b = vtx[2][0], c = vtx[2][1];
    double ccx = a / 2.0, ccy = (b * b + c * c - a * b) / (2.0 * c);
    double mp[3][2] = {{(vtx[0][0] + vtx[1][0]) / 2.0, (vtx[0][1] + vtx[1][1]) / 2.0},
                       {(vtx[1][0] + vtx[2][0]) / 2.0, (vtx[1][1] + vtx[2][1]) / 2.0},
                       {(vtx[2][0] + vtx[0][0]) / 2.0, (vtx[2][1] + vtx[0][1]) / 2.0}};
    bool zero = false;
    for (int i = 0; i < 3; ++i) {
        d[i][0] = ccx - mp[i][0];
        d[i][1] = ccy - mp[i][1];
        zero = zero || (d[i][0] == 0.0 && d[i][1] == 0.0);
    }
    return zero;
}

__device__ __forceinline__ void finish_dirs(double d[3][2], const double vtx[3][2], double out[3][3]) {
    double e01[2] = {vtx[1][0] - vtx[0][0], vtx[1][1] - vtx[0][1]}, e02[2] = {vtx[2][0] - vtx[0][0], vtx[2][1] - vtx[0][1]};
    double e10[2] = {-e01[0], -e01[1]}, e12[2] = {vtx[2][0] - vtx[1][0], vtx[2][1] - vtx[1][1]};
    double e20[2] = {-e02[0], -e02[1]}, e21[2] = {-e12[0], -e12[1]};
    bool ob0 = vec_angle_deg(e01, e02, 2) > 90.0;
    bool ob1 = vec_angle_deg(e10, e12, 2) > 90.0;
    bool ob2 = vec_angle_deg(e20, e21, 2) > 90.0;
    const bool flip[3] = {ob2, ob0, ob1};  // embeds.py:243-245
    for (int i = 0; i < 3; ++i) {
        double x = flip[i] ? -d[i][0] : d[i][0], y = flip[i] ? -d[i][1] : d[i][1];
        double n = sqrt(x * x + y * y + 0.0);
        out[i][0] = x / n; out[i][1] = y / n; out[i][2] = 0.0 / n;
    }
}

// embeds.py:188-254; norms[0] is perturbed in place in the right-triangle case (quirk N7)
__device__ void get_directions3(double* norms, double out[3][3]) {
    double d[3][2], vtx[3][2];
    if (raw_dirs(norms, d, vtx)) {
        norms[0] += 1e-5;
        double d2[3][2], vtx2[3][2], rec[3][3];
        raw_dirs(norms, d2, vtx2);
        finish_dirs(d2, vtx2, rec);
        for (int i = 0; i < 3; ++i) { d[i][0] = rec[i][0]; d[i][1] = rec[i][1]; }
    }
    finish_dirs(d, vtx, out);
}

// alignment of molecule m of a group onto side (start, end) with facing direction dir:
// algebra.py:28-49 align_vec_pair([end - start, dir], [pivot, mol_direction])
__device__ __forceinline__ M3 align_mol(const double* start, const double* end, const double* dir, const double* pivot,
                                        const double* md) {
    double ref0[3] = {end[0] - start[0], end[1] - start[1], end[2] - start[2]};
    double h[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) h[3 * r + c] = ref0[r] * pivot[c] + dir[r] * md[c];
    return kabsch_from_cov(h, nullptr);
}

__global__ void __launch_bounds__(128) cyc3_directions_kernel(Cyc3Dev p) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int sgi = blockIdx.x * 4 + wib;
    __shared__ double s_pts[4][6][7][3];  // rotated reactive-atom positions [2 * molecule + slot][angle step]
    __shared__ double s_pm[4][3][3];      // mean point of each triangle side
    if (sgi >= p.n_supers) return;
    const Cyc3Super sg = p.supers[sgi];
    double norms[3], norms_poly[3];
    for (int m = 0; m < 3; ++m) norms[m] = norms_poly[m] = p.pnorm[m][sg.piv[m]];
    double dirs[3][3];
    get_directions3(norms, dirs);
    double V[3][3] = {{0.0, 0.0, 0.0}, {norms[0], 0.0, 0.0}, {0.0, 0.0, 0.0}};
    tri_vertex2(norms, V[2][0], V[2][1]);

    const int m = lane < 3 ? lane : 0;
    // data of "my" molecule (lanes 0..2)
    const double* x = p.coords[m] + (size_t)sg.conf[m] * p.n_atoms[m] * 3;
    const double* x0 = p.coords[m];  // conformer 0 (embeds.py:359-366)
    const double* pivot = p.pvec[m] + 3 * (size_t)sg.piv[m];
    const double* mean = p.pmean[m] + 3 * (size_t)sg.piv[m];
    double r0[3], r1[3] = {0, 0, 0}, apm[3];
    {
        const double* a0 = x + 3 * p.reactive[m][0];
        r0[0] = a0[0]; r0[1] = a0[1]; r0[2] = a0[2];
        if (p.n_react[m] == 2) {
            const double* a1 = x + 3 * p.reactive[m][1];
            r1[0] = a1[0]; r1[1] = a1[1]; r1[2] = a1[2];
            apm[0] = (r0[0] + r1[0]) / 2.0; apm[1] = (r0[1] + r1[1]) / 2.0; apm[2] = (r0[2] + r1[2]) / 2.0;
        } else {
            apm[0] = r0[0]; apm[1] = r0[1]; apm[2] = r0[2];
        }
    }
    double md[3] = {mean[0] - apm[0], mean[1] - apm[1], mean[2] - apm[2]};
    if (md[0] == 0.0 && md[1] == 0.0 && md[2] == 0.0) { md[0] = mean[0]; md[1] = mean[1]; md[2] = mean[2]; }

    int g = sg.first_group;
    for (int v = 0; v < 8; ++v) {
        if (!((sg.active >> v) & 1)) continue;
        const Cyc3Group gr = p.groups[g];
        double start[3], end[3];
        poly3(norms_poly, v, m, start, end);
        if (lane < 3) {
            M3 A = align_mol(start, end, dirs[m], pivot, md);
            double am[3];
            m3_apply(A, mean, am);
            double pos[3] = {(start[0] + end[0]) / 2.0 - am[0], (start[1] + end[1]) / 2.0 - am[1],
                             (start[2] + end[2]) / 2.0 - am[2]};
            double side[3] = {end[0] - start[0], end[1] - start[1], end[2] - start[2]};
            for (int k = 0; k < 3; ++k) s_pm[wib][m][k] = (end[k] + start[k]) / 2.0;
            for (int slot = 0; slot < 2; ++slot) {
                const double* atom = x0 + 3 * (size_t)gr.r[2 * m + slot];
                double a[3];
                m3_apply(A, atom, a);
                a[0] += pos[0]; a[1] += pos[1]; a[2] += pos[2];
                for (int k = 0; k < 7; ++k) {
                    M3 R = rot_from_pointer(side, -30.0 + 10.0 * k, p.handed);
                    m3_apply(R, a, s_pts[wib][2 * m + slot][k]);
                }
            }
        }
        __syncwarp();
        // ---- 343 candidates, cartesian_product order: third index fastest, first middle, second outermost
        double best = 1e300, second = 1e300;
        int bidx = 0x7fffffff;
        for (int c = lane; c < 343; c += 32) {
            const int k2 = c % 7, k0 = (c / 7) % 7, k1 = c / 49;
            const double* a01 = s_pts[wib][0][k0];
            const double* a02 = s_pts[wib][1][k0];
            const double* a10 = s_pts[wib][2][k1];
            const double* a12 = s_pts[wib][3][k1];
            const double* a20 = s_pts[wib][4][k2];
            const double* a21 = s_pts[wib][5][k2];
            double u[3], w[3], cost = 0.0;
            for (int k = 0; k < 3; ++k) { u[k] = V[0][k] - a02[k]; w[k] = a20[k] - V[0][k]; }
            cost += vec_angle_deg(u, w, 3);
            for (int k = 0; k < 3; ++k) { u[k] = V[1][k] - a01[k]; w[k] = a10[k] - V[1][k]; }
            cost += vec_angle_deg(u, w, 3);
            for (int k = 0; k < 3; ++k) { u[k] = V[2][k] - a21[k]; w[k] = a12[k] - V[2][k]; }
            cost += vec_angle_deg(u, w, 3);
            if (cost < best) { second = best; best = cost; bidx = c; }
            else if (cost < second) second = cost;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ob = __shfl_xor_sync(0xffffffffu, best, o), os = __shfl_xor_sync(0xffffffffu, second, o);
            int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
            if (ob < best || (ob == best && oi < bidx)) { second = fmin(best, os); best = ob; bidx = oi; }
            else second = fmin(second, ob);
        }
        if (lane < 3) {
            const int km = m == 0 ? (bidx / 7) % 7 : (m == 1 ? bidx / 49 : bidx % 7);
            const double* q0 = s_pts[wib][2 * m][km];
            const double* q1 = s_pts[wib][2 * m + 1][km];
            for (int k = 0; k < 3; ++k) dirs[m][k] = s_pm[wib][m][k] - (q0[k] + q1[k]) / 2.0;
            // ---- angle-independent transform of molecule m for this group (embeds.py:494-554)
            M3 A = align_mol(start, end, dirs[m], pivot, md);
            Cyc3Xf o;
            for (int k = 0; k < 9; ++k) o.a[k] = A.m[k];
            double ax_src[3];
            if (p.n_react[m] == 2) { ax_src[0] = r0[0] - r1[0]; ax_src[1] = r0[1] - r1[1]; ax_src[2] = r0[2] - r1[2]; }
            else { ax_src[0] = pivot[0]; ax_src[1] = pivot[1]; ax_src[2] = pivot[2]; }
            m3_apply(A, ax_src, o.axis);
            m3_apply(A, apm, o.c);
            double am[3];
            m3_apply(A, mean, am);
            o.pos[0] = (start[0] + end[0]) / 2.0 - am[0];
            o.pos[1] = (start[1] + end[1]) / 2.0 - am[1];
            o.pos[2] = (start[2] + end[2]) / 2.0 - am[2];
            p.gx[(size_t)g * 3 + m] = o;
        }
        if (lane == 0) { p.choice[g] = bidx; p.gap[g] = second - best; }
        __syncwarp();
        ++g;
    }
}

// absolute transform of molecule m for (group, angle): R = S A, t = c - S c + pos
__device__ __forceinline__ void cyc3_mol_xf(const Cyc3Dev& p, long long g, int m, double angle, M3& rot, double* t) {
    const Cyc3Xf& q = p.gx[g * 3 + m];
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

// transform of molecule mj in the frame of molecule mi for every (group, distinct angle pair)
__global__ void cyc3_pair_xf_kernel(Cyc3Dev p, int mi, int mj, const double* __restrict__ ua, int n_u,
                                    double* __restrict__ xf_rel) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)p.n_groups * n_u) return;
    long long g = i / n_u;
    int u = (int)(i - g * n_u);
    M3 ri, rj;
    double ti[3], tj[3];
    cyc3_mol_xf(p, g, mi, ua[2 * u], ri, ti);
    cyc3_mol_xf(p, g, mj, ua[2 * u + 1], rj, tj);
    M3 rit = m3_transpose(ri);
    M3 rel = m3_mul(rit, rj);
    double dt[3] = {tj[0] - ti[0], tj[1] - ti[1], tj[2] - ti[2]}, trel[3];
    m3_apply(rit, dt, trel);
    double* o = xf_rel + i * 12;
    for (int k = 0; k < 9; ++k) o[k] = rel.m[k];
    o[9] = trel[0]; o[10] = trel[1]; o[11] = trel[2];
}

// absolute transform (R row-major, t) of molecule m of group g for each DISTINCT angle of that molecule's column
__global__ void cyc3_abs_xf_kernel(Cyc3Dev p, const double* __restrict__ uang, int u_max, int n_u0, int n_u1, int n_u2,
                                   double* __restrict__ xf_abs) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)p.n_groups * 3 * u_max) return;
    const int u = (int)(i % u_max);
    const int m = (int)((i / u_max) % 3);
    const long long g = i / ((long long)u_max * 3);
    const int n_u = m == 0 ? n_u0 : (m == 1 ? n_u1 : n_u2);
    if (u >= n_u) return;
    M3 r;
    double t[3];
    cyc3_mol_xf(p, g, m, uang[m * u_max + u], r, t);
    double* o = xf_abs + i * 12;
    for (int k = 0; k < 9; ++k) o[k] = r.m[k];
    o[9] = t[0]; o[10] = t[1]; o[11] = t[2];
}

__device__ __forceinline__ void cyc3_atom(const Cyc3Dev& p, const int* conf, const M3* rot, const double (*t)[3],
                                          int atom, double* out) {
    int m = atom < p.n_atoms[0] ? 0 : (atom < p.n_atoms[0] + p.n_atoms[1] ? 1 : 2);
    int local = atom - (m == 0 ? 0 : (m == 1 ? p.n_atoms[0] : p.n_atoms[0] + p.n_atoms[1]));
    const double* b = p.coords[m] + ((size_t)conf[m] * p.n_atoms[m] + local) * 3;
    const double* r = rot[m].m;
    out[0] = (r[0] * b[0] + r[1] * b[1] + r[2] * b[2]) + t[m][0];
    out[1] = (r[3] * b[0] + r[4] * b[1] + r[5] * b[2]) + t[m][1];
    out[2] = (r[6] * b[0] + r[7] * b[1] + r[8] * b[2]) + t[m][2];
}

__device__ __forceinline__ double wsum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double wmax_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

struct Cyc3SimArgs {
    Cyc3Dev p;
    const uint8_t* st[3];   // per pair: status over (group, distinct angle pair)
    const int* umap[3];     // per pair: angle index -> distinct angle pair
    int n_u[3];
    const double* xf_abs;   // (G, 3, u_max, 12): absolute transform of every molecule for its distinct angles
    const int* amap[3];     // per molecule: angle index -> distinct angle of that molecule
    int u_max;
    const double* mom[3];   // per molecule (n_conf, 10): sum b b^T (00 01 02 11 12 22), sum b (3), n_atoms
    uint8_t* status;        // per pose, out: combined FC_STATUS_* bits
    uint8_t* keep;          // per pose, out
    double rmsd_thr, eps;
    TieRecord* ties;
    int* n_ties;
    int tie_cap;
    long long pose_base;
};

__device__ __forceinline__ void push_tie3(const Cyc3SimArgs& a, long long pose, long long ref, double value, int kind,
                                          bool decision) {
    int slot = atomicAdd(a.n_ties, 1);
    if (slot < a.tie_cap) {
        TieRecord r;
        r.a = pose + a.pose_base; r.b = ref + a.pose_base; r.value = value; r.kind = kind; r.decision = decision ? 1 : 0;
        a.ties[slot] = r;
    }
}

// second / first moments of every conformer: with them the cross-covariance of two poses of the same
// conformers needs no loop over atoms,
//   sum (Rp b + tp)(Rq b + tq)^T = Rp M Rq^T + (Rp s) tq^T + tp (Rq s)^T + n tp tq^T,   M = sum b b^T, s = sum b
__global__ void cyc3_moments_kernel(const double* __restrict__ coords, int n_conf, int n_atoms, double* __restrict__ mom) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= n_conf) return;
    const double* x = coords + (size_t)c * n_atoms * 3;
    double v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = lane; k < n_atoms; k += 32) {
        const double a = x[3 * k], b = x[3 * k + 1], d = x[3 * k + 2];
        v[0] += a * a; v[1] += a * b; v[2] += a * d; v[3] += b * b; v[4] += b * d; v[5] += d * d;
        v[6] += a; v[7] += b; v[8] += d;
    }
#pragma unroll
    for (int e = 0; e < 9; ++e) v[e] = wsum_d(v[e]);
    if (lane == 0) {
        double* o = mom + (size_t)c * 10;
        for (int e = 0; e < 9; ++e) o[e] = v[e];
        o[9] = (double)n_atoms;
    }
}

// cross-covariance H (+=) and |p|^2 + |q|^2 (+=) of one molecule placed by (rp, tp) and (rq, tq), from its moments
__device__ __forceinline__ void cov_from_moments(const double* mo, const double* rp, const double* rq, double* h, double& gsum) {
    const double M[9] = {mo[0], mo[1], mo[2], mo[1], mo[3], mo[4], mo[2], mo[4], mo[5]};
    const double n = mo[9];
    double A[9];  // Rp M
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) A[3 * i + j] = rp[3 * i] * M[j] + rp[3 * i + 1] * M[3 + j] + rp[3 * i + 2] * M[6 + j];
    double u[3], w[3];  // Rp s, Rq s
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        u[i] = rp[3 * i] * mo[6] + rp[3 * i + 1] * mo[7] + rp[3 * i + 2] * mo[8];
        w[i] = rq[3 * i] * mo[6] + rq[3 * i + 1] * mo[7] + rq[3 * i + 2] * mo[8];
    }
    const double* tp = rp + 9;
    const double* tq = rq + 9;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            h[3 * i + j] += (A[3 * i] * rq[3 * j] + A[3 * i + 1] * rq[3 * j + 1] + A[3 * i + 2] * rq[3 * j + 2]) + u[i] * tq[j] +
                            tp[i] * w[j] + n * tp[i] * tq[j];
    // sum |R b + t|^2 = tr(M) + 2 t.(R s) + n |t|^2 (R orthogonal)
    const double tr = mo[0] + mo[3] + mo[5];
    gsum += 2.0 * tr + 2.0 * (tp[0] * u[0] + tp[1] * u[1] + tp[2] * u[2]) + 2.0 * (tq[0] * w[0] + tq[1] * w[1] + tq[2] * w[2]) +
            n * (tp[0] * tp[0] + tp[1] * tp[1] + tp[2] * tp[2] + tq[0] * tq[0] + tq[1] * tq[1] + tq[2] * tq[2]);
}

// combined status of every pose: PASS iff all three block screens pass (utils.py:553-575 with max_clashes = 0),
// RECHECKED / NEAR if any block was; also counts passes and rechecks
__global__ void __launch_bounds__(256) cyc3_combine_kernel(Cyc3SimArgs a, long long n_poses, unsigned long long* __restrict__ counts) {
    const long long pose = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    uint8_t comb = 0;
    if (pose < n_poses) {
        const int n_ang = a.p.n_angles;
        const long long g = pose / n_ang;
        const int ai = (int)(pose - g * n_ang);
        uint8_t s0 = a.st[0][g * a.n_u[0] + a.umap[0][ai]];
        uint8_t s1 = a.st[1][g * a.n_u[1] + a.umap[1][ai]];
        uint8_t s2 = a.st[2][g * a.n_u[2] + a.umap[2][ai]];
        comb = (uint8_t)((s0 & s1 & s2 & FC_STATUS_PASS) | ((s0 | s1 | s2) & (FC_STATUS_RECHECKED | FC_STATUS_NEAR)));
        a.status[pose] = comb;
        a.keep[pose] = 0;
    }
    const unsigned pass = __ballot_sync(0xffffffffu, comb & FC_STATUS_PASS);
    const unsigned rech = __ballot_sync(0xffffffffu, comb & FC_STATUS_RECHECKED);
    if ((threadIdx.x & 31) == 0 && (pass | rech)) {
        if (pass) atomicAdd(counts, (unsigned long long)__popc(pass));
        if (rech) atomicAdd(counts + 1, (unsigned long long)__popc(rech));
    }
}

// one warp per group: keep-first over the clash survivors in angle order
__global__ void __launch_bounds__(128, 3) cyc3_group_similarity_kernel(Cyc3SimArgs a) {
    const Cyc3Dev& p = a.p;
    const int lane = threadIdx.x & 31;
    const long long g = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= p.n_groups) return;
    const int n_ang = p.n_angles;
    const int n_tot = p.n_atoms[0] + p.n_atoms[1] + p.n_atoms[2];
    const Cyc3Super sg = p.supers[p.groups[g].super];
    const int conf[3] = {sg.conf[0], sg.conf[1], sg.conf[2]};
    extern __shared__ int s_acc_all[];
    int* s_acc = s_acc_all + (threadIdx.x >> 5) * n_ang;
    int n_acc = 0;
    for (int base = 0; base < n_ang; base += 32) {
        const int mine = base + lane;
        const bool ok = mine < n_ang && (a.status[g * n_ang + mine] & FC_STATUS_PASS);
        unsigned survivors = __ballot_sync(0xffffffffu, ok);
        while (survivors) {
            const int ai = base + __ffs(survivors) - 1;
            survivors &= survivors - 1u;
            const long long pose = g * n_ang + ai;
            const double* xp[3];
#pragma unroll
            for (int m = 0; m < 3; ++m) xp[m] = a.xf_abs + (((size_t)g * 3 + m) * a.u_max + a.amap[m][ai]) * 12;
            bool similar = false;
            for (int k0 = 0; k0 < n_acc && !similar; k0 += 32) {
                // ---- screen, one accepted pose per lane: covariance from the conformers' moments (no atom loop),
                //      closed-form singular values; kept poses are mutually dissimilar, so this is where almost
                //      every comparison ends
                const int k = k0 + lane;
                bool need = false;
                if (k < n_acc) {
                    const int ak = s_acc[k];
                    double hs[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
                    double gs = 0.0;
#pragma unroll
                    for (int m = 0; m < 3; ++m) {
                        const double* xq = a.xf_abs + (((size_t)g * 3 + m) * a.u_max + a.amap[m][ak]) * 12;
                        double rp[12], rq[12];
#pragma unroll
                        for (int e = 0; e < 12; ++e) { rp[e] = xp[m][e]; rq[e] = xq[e]; }
                        cov_from_moments(a.mom[m] + (size_t)conf[m] * 10, rp, rq, hs, gs);
                    }
                    const double lim = a.rmsd_thr + 1e-4;
                    need = !((gs - 2.0 * singular_sum3(hs)) / n_tot > lim * lim);
                }
                unsigned todo = __ballot_sync(0xffffffffu, need);
                // ---- exact evaluation (whole warp per pair, atoms over the lanes) of the few pairs left, in order
                while (todo && !similar) {
                    const int aj = s_acc[k0 + __ffs(todo) - 1];
                    todo &= todo - 1u;
                    double h[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
                    double gsum = 0.0;  // |p|^2 + |q|^2
    #pragma unroll
                    for (int m = 0; m < 3; ++m) {
                        const double* xq = a.xf_abs + (((size_t)g * 3 + m) * a.u_max + a.amap[m][aj]) * 12;
                        double rp[12], rq[12];
    #pragma unroll
                        for (int e = 0; e < 12; ++e) { rp[e] = xp[m][e]; rq[e] = xq[e]; }
                        const double* base = p.coords[m] + (size_t)conf[m] * p.n_atoms[m] * 3;
                        for (int at = lane; at < p.n_atoms[m]; at += 32) {
                            const double bx = base[3 * at], by = base[3 * at + 1], bz = base[3 * at + 2];
                            const double x[3] = {(rp[0] * bx + rp[1] * by + rp[2] * bz) + rp[9], (rp[3] * bx + rp[4] * by + rp[5] * bz) + rp[10],
                                                 (rp[6] * bx + rp[7] * by + rp[8] * bz) + rp[11]};
                            const double y[3] = {(rq[0] * bx + rq[1] * by + rq[2] * bz) + rq[9], (rq[3] * bx + rq[4] * by + rq[5] * bz) + rq[10],
                                                 (rq[6] * bx + rq[7] * by + rq[8] * bz) + rq[11]};
    #pragma unroll
                            for (int r = 0; r < 3; ++r) {
                                gsum += x[r] * x[r] + y[r] * y[r];
    #pragma unroll
                                for (int c = 0; c < 3; ++c) h[3 * r + c] += x[r] * y[c];
                            }
                        }
                    }
    #pragma unroll
                    for (int e = 0; e < 9; ++e) h[e] = wsum_d(h[e]);
                    gsum = wsum_d(gsum);
                    // screen: RMSD from the closed-form singular values; clearly dissimilar pairs (the common case
                    // among kept poses) skip the Jacobi solve and the second pass over the atoms
                    {
                        const double msd = (gsum - 2.0 * singular_sum3(h)) / n_tot;
                        const double lim = a.rmsd_thr + 1e-4;
                        if (msd > lim * lim) continue;
                    }
                    M3 R = kabsch_from_cov(h, nullptr);
                    double ss = 0.0, mx = 0.0;
    #pragma unroll
                    for (int m = 0; m < 3; ++m) {
                        const double* xq = a.xf_abs + (((size_t)g * 3 + m) * a.u_max + a.amap[m][aj]) * 12;
                        double rp[12], rq[12];
    #pragma unroll
                        for (int e = 0; e < 12; ++e) { rp[e] = xp[m][e]; rq[e] = xq[e]; }
                        const double* base = p.coords[m] + (size_t)conf[m] * p.n_atoms[m] * 3;
                        for (int at = lane; at < p.n_atoms[m]; at += 32) {
                            const double bx = base[3 * at], by = base[3 * at + 1], bz = base[3 * at + 2];
                            const double x[3] = {(rp[0] * bx + rp[1] * by + rp[2] * bz) + rp[9], (rp[3] * bx + rp[4] * by + rp[5] * bz) + rp[10],
                                                 (rp[6] * bx + rp[7] * by + rp[8] * bz) + rp[11]};
                            const double y[3] = {(rq[0] * bx + rq[1] * by + rq[2] * bz) + rq[9], (rq[3] * bx + rq[4] * by + rq[5] * bz) + rq[10],
                                                 (rq[6] * bx + rq[7] * by + rq[8] * bz) + rq[11]};
                            // diff = p @ R - q
                            const double dx = (x[0] * R.m[0] + x[1] * R.m[3] + x[2] * R.m[6]) - y[0];
                            const double dy = (x[0] * R.m[1] + x[1] * R.m[4] + x[2] * R.m[7]) - y[1];
                            const double dz = (x[0] * R.m[2] + x[1] * R.m[5] + x[2] * R.m[8]) - y[2];
                            const double d2 = dx * dx + dy * dy + dz * dz;
                            ss += d2;
                            mx = fmax(mx, d2);
                        }
                    }
                    ss = wsum_d(ss);
                    mx = wmax_d(mx);
                    const double rmsd = sqrt(ss / n_tot), maxdev = sqrt(mx);
                    const bool rm_ok = rmsd < a.rmsd_thr, md_ok = maxdev < 2.0 * a.rmsd_thr;
                    if (lane == 0) {
                        const long long ref = g * n_ang + aj;
                        if (fabs(rmsd - a.rmsd_thr) <= a.eps) push_tie3(a, pose, ref, rmsd, FC_TIE_RMSD, rm_ok);
                        if (fabs(maxdev - 2.0 * a.rmsd_thr) <= a.eps) push_tie3(a, pose, ref, maxdev, FC_TIE_MAXDEV, md_ok);
                    }
                    similar = rm_ok && md_ok;
                }
            }
            if (!similar) {
                if (lane == 0) {
                    s_acc[n_acc] = ai;
                    a.keep[pose] = FC_STATUS_PASS;
                }
                ++n_acc;
                __syncwarp();
            }
        }
    }
}

__global__ void cyc3_materialize_kernel(Cyc3Dev p, const long long* __restrict__ kept, int n_kept,
                                        double* __restrict__ out) {
    int k = blockIdx.x;
    if (k >= n_kept) return;
    long long pose = kept[k];
    long long g = pose / p.n_angles;
    int ai = (int)(pose - g * p.n_angles);
    __shared__ M3 rot[3];
    __shared__ double t[3][3];
    if (threadIdx.x < 3) cyc3_mol_xf(p, g, threadIdx.x, p.angles[3 * ai + threadIdx.x], rot[threadIdx.x], t[threadIdx.x]);
    __syncthreads();
    const Cyc3Super sg = p.supers[p.groups[g].super];
    const int conf[3] = {sg.conf[0], sg.conf[1], sg.conf[2]};
    int n_tot = p.n_atoms[0] + p.n_atoms[1] + p.n_atoms[2];
    double* o = out + (size_t)k * n_tot * 3;
    for (int at = threadIdx.x; at < n_tot; at += blockDim.x) {
        double v[3];
        cyc3_atom(p, conf, rot, t, at, v);
        o[3 * at] = v[0]; o[3 * at + 1] = v[1]; o[3 * at + 2] = v[2];
    }
}

// per kept pose: the rigid transform (R row-major, t) and the conformer of every molecule -- what is needed to place
// its atoms later (fc_result_kept_coords streams the coordinates into the caller's buffer)
__global__ void cyc3_kept_xf_kernel(Cyc3Dev p, const long long* __restrict__ kept, int n_kept, double* __restrict__ xf,
                                    int32_t* __restrict__ conf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_kept * 3) return;
    const int k = i / 3, m = i - 3 * k;
    const long long pose = kept[k];
    const long long g = pose / p.n_angles;
    const int ai = (int)(pose - g * p.n_angles);
    M3 rot;
    double t[3];
    cyc3_mol_xf(p, g, m, p.angles[3 * ai + m], rot, t);
    double* o = xf + (size_t)i * 12;
    for (int e = 0; e < 9; ++e) o[e] = rot.m[e];
    o[9] = t[0]; o[10] = t[1]; o[11] = t[2];
    conf[i] = p.supers[p.groups[g].super].conf[m];
}

// ---- host-side enumeration (reference loop order) ------------------------------------------------
struct HostGroup {
    int ids[6];  // three sorted atom couples (embeds.py:774-784)
};

static bool pair_in(const int64_t* list, int n, int64_t a, int64_t b) {
    for (int i = 0; i < n; ++i)
        if (list[2 * i] == a && list[2 * i + 1] == b) return true;
    return false;
}

}  // namespace fc

using namespace fc;

extern "C" int fc_cyclical3_screen(const fc_cyclical3_problem* p, fc_result** out) {
    FC_REQUIRE(out, "null output");
    *out = nullptr;
    FC_REQUIRE(p, "null problem");
    FC_REQUIRE(p->n_angles > 0 && p->n_angles <= 4096 && p->angles, "bad angle table");
    for (int m = 0; m < 3; ++m) {
        FC_REQUIRE(p->coords[m] && p->n_conf[m] > 0 && p->n_atoms[m] > 0, "empty ensemble %d", m);
        FC_REQUIRE(p->reactive[m] && (p->n_reactive[m] == 1 || p->n_reactive[m] == 2), "molecule %d needs 1 or 2 reactive atoms", m);
        for (int k = 0; k < p->n_reactive[m]; ++k)
            FC_REQUIRE(p->reactive[m][k] >= 0 && p->reactive[m][k] < p->n_atoms[m], "reactive index out of range");
        FC_REQUIRE(p->pivot_offsets[m], "null pivot table %d", m);
        int64_t rows = p->pivot_offsets[m][p->n_conf[m]];
        FC_REQUIRE(rows >= 0 && rows < ((int64_t)1 << 30), "bad pivot table %d", m);
        FC_REQUIRE(rows == 0 || (p->pivot_vec[m] && p->pivot_mean[m] && p->pivot_ids[m]), "null pivot table %d", m);
        FC_REQUIRE(p->n_ratoms0[m] == 0 || p->ratoms0[m], "null reactive atom table %d", m);
        for (int k = 0; k < p->n_ratoms0[m]; ++k)
            FC_REQUIRE(p->ratoms0[m][2 * k] >= 0 && p->ratoms0[m][2 * k] < p->n_atoms[m], "reactive atom index out of range");
    }
    const bool trace = getenv("FC_CLASH_TRACE") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    double t_enum = 0, t_dir = 0, t_clash = 0, t_sim = 0, t_host = 0, t_mat = 0;
    const int A = p->n_angles;
    const int64_t n_tot = (int64_t)p->n_atoms[0] + p->n_atoms[1] + p->n_atoms[2];
    const int n0 = p->n_conf[0], n1 = p->n_conf[1], n2 = p->n_conf[2];
    const int64_t n_tuples = (int64_t)n0 * n1 * n2;
    int64_t t_lo = 0, t_hi = n_tuples;
    if (p->conf_tuple_hi > 0) {
        t_lo = std::max<int64_t>(0, p->conf_tuple_lo);
        t_hi = std::min<int64_t>(n_tuples, p->conf_tuple_hi);
    }

    // pivot norms: np.linalg.norm(axis=1) = sqrt((x*x + y*y) + z*z)
    std::vector<double> pnorm[3];
    for (int m = 0; m < 3; ++m) {
        int64_t rows = p->pivot_offsets[m][p->n_conf[m]];
        pnorm[m].resize((size_t)rows);
        for (int64_t r = 0; r < rows; ++r) {
            const double* v = p->pivot_vec[m] + 3 * r;
            volatile double s = v[0] * v[0];
            s = s + v[1] * v[1];
            s = s + v[2] * v[2];
            pnorm[m][(size_t)r] = sqrt(s);
        }
    }

    // ---- enumerate super-groups and groups (host threads over contiguous conformer-triple ranges; the
    //      concatenation in range order is the reference's loop order) ------------------------------------
    struct EnumPart {
        std::vector<Cyc3Super> supers;
        std::vector<Cyc3Group> groups;
        std::vector<HostGroup> hgroups;
        bool bad = false;
    };
    static const int swaps[8][3] = {{0, 0, 0}, {0, 0, 1}, {0, 1, 0}, {0, 1, 1}, {1, 0, 0}, {1, 1, 0}, {1, 0, 1}, {1, 1, 1}};
    auto enumerate = [&](int64_t ta, int64_t tb, EnumPart& L) {
        for (int64_t t = ta; t < tb; ++t) {
            // cartesian_product of three ranges: third fastest, first middle, second outermost
            const int c2 = (int)(t % n2), c0 = (int)((t / n2) % n0), c1 = (int)(t / ((int64_t)n2 * n0));
            const int conf[3] = {c0, c1, c2};
            int64_t lo[3], cnt[3];
            for (int m = 0; m < 3; ++m) {
                lo[m] = p->pivot_offsets[m][conf[m]];
                cnt[m] = p->pivot_offsets[m][conf[m] + 1] - lo[m];
            }
            if (cnt[0] <= 0 || cnt[1] <= 0 || cnt[2] <= 0) continue;
            const int64_t n_pt = cnt[0] * cnt[1] * cnt[2];
            for (int64_t q = 0; q < n_pt; ++q) {
                const int64_t q2 = q % cnt[2], q0 = (q / cnt[2]) % cnt[0], q1 = q / (cnt[2] * cnt[0]);
                const int64_t row[3] = {lo[0] + q0, lo[1] + q1, lo[2] + q2};
                const double nn[3] = {pnorm[0][(size_t)row[0]], pnorm[1][(size_t)row[1]], pnorm[2][(size_t)row[2]]};
                // embeds.py:447: all(norms[i] < norms[i-1] + norms[i-2] for i in (0, 1, 2))
                if (!(nn[0] < nn[2] + nn[1] && nn[1] < nn[0] + nn[2] && nn[2] < nn[1] + nn[0])) continue;
                Cyc3Super sg;
                for (int m = 0; m < 3; ++m) { sg.conf[m] = conf[m]; sg.piv[m] = (int)row[m]; }
                sg.first_group = (int)L.groups.size();
                sg.active = 0;
                for (int v = 0; v < 8; ++v) {
                    int64_t o[3][2];
                    for (int m = 0; m < 3; ++m) {
                        const int64_t* id = p->pivot_ids[m] + 2 * row[m];
                        o[m][0] = swaps[v][m] ? id[1] : id[0];
                        o[m][1] = swaps[v][m] ? id[0] : id[1];
                    }
                    int64_t cp[3][2] = {{o[0][1], o[1][0]}, {o[1][1], o[2][0]}, {o[2][1], o[0][0]}};
                    for (int k = 0; k < 3; ++k)
                        if (cp[k][0] > cp[k][1]) std::swap(cp[k][0], cp[k][1]);
                    bool ok = true;
                    for (int i = 0; i < p->n_pairings && ok; ++i) {
                        const int64_t a = p->pairings[2 * i], b = p->pairings[2 * i + 1];
                        bool in_ids = false;
                        for (int k = 0; k < 3; ++k) in_ids = in_ids || (cp[k][0] == a && cp[k][1] == b);
                        ok = in_ids || pair_in(p->internal, p->n_internal, a, b);
                    }
                    if (!ok) continue;
                    sg.active |= 1 << v;
                    Cyc3Group gr;
                    gr.super = (int)L.supers.size();
                    gr.v = v;
                    // facing table r[m][partner] (embeds.py:328-353)
                    int pr[3][2][2];  // [couple][end] -> (mol, index)
                    for (int k = 0; k < 3; ++k)
                        for (int e = 0; e < 2; ++e) {
                            pr[k][e][0] = -1; pr[k][e][1] = -1;
                            for (int m = 0; m < 3; ++m)
                                for (int j = 0; j < p->n_ratoms0[m]; ++j)
                                    if (p->ratoms0[m][2 * j + 1] == cp[k][e]) { pr[k][e][0] = m; pr[k][e][1] = (int)p->ratoms0[m][2 * j]; }
                        }
                    int rt[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
                    for (int k = 0; k < 3; ++k) {
                        // python negative indices wrap: -1 -> 2
                        int m0 = pr[k][0][0] < 0 ? 2 : pr[k][0][0], m1 = pr[k][1][0] < 0 ? 2 : pr[k][1][0];
                        rt[m0][m1] = pr[k][0][1];
                        rt[m1][m0] = pr[k][1][1];
                    }
                    const int rr[6] = {rt[0][1], rt[0][2], rt[1][0], rt[1][2], rt[2][0], rt[2][1]};
                    for (int k = 0; k < 6; ++k) {
                        int mol = k / 2;
                        // an unmatched couple leaves -1 (python: coords[0][-1] = last atom)
                        gr.r[k] = rr[k] < 0 ? p->n_atoms[mol] + rr[k] : rr[k];
                        if (!(gr.r[k] >= 0 && gr.r[k] < p->n_atoms[mol])) { L.bad = true; return; }
                    }
                    L.groups.push_back(gr);
                    HostGroup hg;
                    for (int k = 0; k < 3; ++k) { hg.ids[2 * k] = (int)cp[k][0]; hg.ids[2 * k + 1] = (int)cp[k][1]; }
                    L.hgroups.push_back(hg);
                }
                if (sg.active) L.supers.push_back(sg);
                if (L.groups.size() >= ((size_t)1 << 30)) { L.bad = true; return; }
            }
        }
    };
    const int64_t span = t_hi - t_lo;
    int n_threads = (int)std::min<int64_t>(std::max(1u, std::min(16u, std::thread::hardware_concurrency())), std::max<int64_t>(1, span / 512));
    std::vector<EnumPart> parts((size_t)n_threads);
    {
        std::vector<std::thread> workers;
        for (int w = 0; w < n_threads; ++w) {
            const int64_t ta = t_lo + span * w / n_threads, tb = t_lo + span * (w + 1) / n_threads;
            if (w + 1 == n_threads) enumerate(ta, tb, parts[(size_t)w]);
            else workers.emplace_back(enumerate, ta, tb, std::ref(parts[(size_t)w]));
        }
        for (auto& th : workers) th.join();
    }
    std::vector<Cyc3Super> supers;
    std::vector<Cyc3Group> groups;
    std::vector<HostGroup> hgroups;
    {
        // concatenate the parts in range order (= the reference's loop order); every part is rebased and copied by
        // its own thread
        std::vector<size_t> s_off((size_t)n_threads + 1, 0), g_off((size_t)n_threads + 1, 0);
        for (int w = 0; w < n_threads; ++w) {
            FC_REQUIRE(!parts[(size_t)w].bad, "fc_cyclical3_screen: facing atom out of range or too many groups");
            s_off[(size_t)w + 1] = s_off[(size_t)w] + parts[(size_t)w].supers.size();
            g_off[(size_t)w + 1] = g_off[(size_t)w] + parts[(size_t)w].groups.size();
        }
        FC_REQUIRE(g_off.back() < ((size_t)1 << 30), "too many groups for one call");
        supers.resize(s_off.back());
        groups.resize(g_off.back());
        hgroups.resize(g_off.back());
        auto place = [&](int w) {
            EnumPart& L = parts[(size_t)w];
            const int so = (int)s_off[(size_t)w], go = (int)g_off[(size_t)w];
            for (size_t i = 0; i < L.supers.size(); ++i) {
                Cyc3Super x = L.supers[i];
                x.first_group += go;
                supers[(size_t)so + i] = x;
            }
            for (size_t i = 0; i < L.groups.size(); ++i) {
                Cyc3Group x = L.groups[i];
                x.super += so;
                groups[(size_t)go + i] = x;
            }
            if (!L.hgroups.empty()) memcpy(&hgroups[(size_t)go], L.hgroups.data(), L.hgroups.size() * sizeof(HostGroup));
        };
        std::vector<std::thread> workers;
        for (int w = 0; w + 1 < n_threads; ++w) workers.emplace_back(place, w);
        place(n_threads - 1);
        for (auto& th : workers) th.join();
    }
    const int64_t G = (int64_t)groups.size();
    t_enum = now() - t0;
    fc_result* r = result_new();
    r->n_atoms = n_tot;
    r->n_pairs = 3;
    r->n_groups = G;
    r->n_poses = G * A;
    if (G == 0) {
        *out = r;
        return FC_OK;
    }

    // ---- distinct angle pairs per block --------------------------------------------------------
    static const int pair_m[3][2] = {{0, 1}, {1, 2}, {2, 0}};
    std::vector<double> ua[3];
    std::vector<int> umap[3];
    for (int k = 0; k < 3; ++k) {
        std::map<std::pair<double, double>, int> seen;
        umap[k].resize(A);
        for (int ai = 0; ai < A; ++ai) {
            std::pair<double, double> key(p->angles[3 * ai + pair_m[k][0]], p->angles[3 * ai + pair_m[k][1]]);
            auto it = seen.find(key);
            if (it == seen.end()) {
                it = seen.emplace(key, (int)seen.size()).first;
                ua[k].push_back(key.first);
                ua[k].push_back(key.second);
            }
            umap[k][ai] = it->second;
        }
    }

    // distinct angles per molecule column (6 for the default grid): absolute transforms are tabulated per group
    std::vector<double> uang_m[3];
    std::vector<int> amap[3];
    int u_max = 1;
    for (int m = 0; m < 3; ++m) {
        std::map<double, int> seen;
        amap[m].resize(A);
        for (int ai = 0; ai < A; ++ai) {
            const double v = p->angles[3 * ai + m];
            auto it = seen.find(v);
            if (it == seen.end()) {
                it = seen.emplace(v, (int)seen.size()).first;
                uang_m[m].push_back(v);
            }
            amap[m][ai] = it->second;
        }
        u_max = std::max(u_max, (int)uang_m[m].size());
    }
    std::vector<double> uang((size_t)3 * u_max, 0.0);
    for (int m = 0; m < 3; ++m) std::copy(uang_m[m].begin(), uang_m[m].end(), uang.begin() + (size_t)m * u_max);

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
    const bool want_status = !(p->flags & FC_CYC3_NO_STATUS), want_coords = !(p->flags & FC_CYC3_NO_COORDS);
    if (want_status) r->status.resize((size_t)(G * A));
    r->group_choice.resize((size_t)G);
    r->group_gap.resize((size_t)G);
    {
        DevBuf<double> d_coords[3], d_pvec[3], d_pmean[3], d_pnorm[3], d_angles, d_ua[3], d_xf, d_out, d_near_dist, d_gap;
        DevBuf<long long> d_react[3], d_kept;
        DevBuf<int> d_umap[3], d_amap[3], d_choice, d_cnt;
        DevBuf<double> d_uang, d_xf_abs, d_mom[3];
        DevBuf<Cyc3Super> d_supers;
        DevBuf<Cyc3Group> d_groups;
        DevBuf<Cyc3Xf> d_gx;
        DevBuf<uint8_t> d_st[3], d_status, d_keep;
        DevBuf<int32_t> d_tiles, d_near_count;
        DevBuf<int64_t> d_near_idx;
        DevBuf<TieRecord> d_ties;
        const int tie_cap = 1 << 20, near_cap = 1 << 16;
        cudaError_t e = cudaSuccess;
#define CY(call) do { if (e == cudaSuccess) e = (call); } while (0)
        for (int m = 0; m < 3; ++m) {
            size_t n = (size_t)p->n_conf[m] * p->n_atoms[m] * 3;
            CY(d_coords[m].alloc(n, s));
            CY(cudaMemcpyAsync(d_coords[m].p, p->coords[m], n * 8, cudaMemcpyHostToDevice, s));
            CY(d_react[m].alloc(2, s));
            CY(cudaMemcpyAsync(d_react[m].p, p->reactive[m], (size_t)p->n_reactive[m] * 8, cudaMemcpyHostToDevice, s));
            size_t rows = pnorm[m].size();
            CY(d_pvec[m].alloc(rows * 3, s));
            CY(d_pmean[m].alloc(rows * 3, s));
            CY(d_pnorm[m].alloc(rows, s));
            CY(cudaMemcpyAsync(d_pvec[m].p, p->pivot_vec[m], rows * 24, cudaMemcpyHostToDevice, s));
            CY(cudaMemcpyAsync(d_pmean[m].p, p->pivot_mean[m], rows * 24, cudaMemcpyHostToDevice, s));
            CY(cudaMemcpyAsync(d_pnorm[m].p, pnorm[m].data(), rows * 8, cudaMemcpyHostToDevice, s));
        }
        for (int m = 0; m < 3; ++m) {
            CY(d_mom[m].alloc((size_t)p->n_conf[m] * 10, s));
            if (e == cudaSuccess)
                cyc3_moments_kernel<<<(unsigned)((p->n_conf[m] + 3) / 4), 128, 0, s>>>(d_coords[m].p, p->n_conf[m], p->n_atoms[m], d_mom[m].p);
        }
        CY(d_angles.alloc((size_t)A * 3, s));
        CY(cudaMemcpyAsync(d_angles.p, p->angles, (size_t)A * 24, cudaMemcpyHostToDevice, s));
        int n_u[3];
        for (int k = 0; k < 3; ++k) {
            n_u[k] = (int)ua[k].size() / 2;
            CY(d_ua[k].alloc(ua[k].size(), s));
            CY(cudaMemcpyAsync(d_ua[k].p, ua[k].data(), ua[k].size() * 8, cudaMemcpyHostToDevice, s));
            CY(d_umap[k].alloc(A, s));
            CY(cudaMemcpyAsync(d_umap[k].p, umap[k].data(), (size_t)A * 4, cudaMemcpyHostToDevice, s));
        }
        CY(d_uang.alloc(uang.size(), s));
        CY(cudaMemcpyAsync(d_uang.p, uang.data(), uang.size() * 8, cudaMemcpyHostToDevice, s));
        for (int m = 0; m < 3; ++m) {
            CY(d_amap[m].alloc(A, s));
            CY(cudaMemcpyAsync(d_amap[m].p, amap[m].data(), (size_t)A * 4, cudaMemcpyHostToDevice, s));
        }
        CY(d_ties.alloc(tie_cap, s));
        CY(d_cnt.alloc(8, s));
        CY(d_near_count.alloc(4, s));
        CY(d_near_idx.alloc(near_cap, s));
        CY(d_near_dist.alloc(near_cap, s));
        if (e != cudaSuccess) rc = cuda_fail(e, "fc_cyclical3_screen setup", __FILE__, __LINE__);

        // fragment-A tables of the three block screens (A = molecule 0, 1, 2 of the blocks (0,1), (1,2), (2,0)):
        // built once, reused by every chunk
        fc_clash_prep* prep[3] = {nullptr, nullptr, nullptr};
        for (int k = 0; k < 3 && !rc; ++k)
            rc = fc_clash_prepare_dev(d_coords[k].p, p->n_conf[k], p->n_atoms[k], p->thresh, 1, &prep[k], (void*)s);

        // ---- chunks of whole super-groups ------------------------------------------------------
        // poses per chunk (FC_CYC3_CHUNK_POSES overrides): device buffers are ~20 B per pose + 96 B per pair-pose
        int64_t chunk_poses = (int64_t)12 << 20;
        if (const char* v = getenv("FC_CYC3_CHUNK_POSES")) chunk_poses = std::max<int64_t>(1 << 16, atoll(v));
        const int64_t max_chunk_groups = std::max<int64_t>(64, chunk_poses / A);
        size_t sg_lo = 0;
        std::vector<uint8_t> h_status, h_keep, h_st;
        while (!rc && sg_lo < supers.size()) {
            size_t sg_hi = sg_lo;
            const int64_t g_lo = supers[sg_lo].first_group;
            int64_t g_hi = g_lo;
            while (sg_hi < supers.size()) {
                int64_t next = sg_hi + 1 < supers.size() ? supers[sg_hi + 1].first_group : G;
                if (sg_hi > sg_lo && next - g_lo > max_chunk_groups) break;
                g_hi = next;
                ++sg_hi;
            }
            const int64_t cg = g_hi - g_lo, cposes = cg * A;
            // chunk-local copies (group.super and super.first_group rebased)
            std::vector<Cyc3Super> cs(supers.begin() + sg_lo, supers.begin() + sg_hi);
            std::vector<Cyc3Group> cgv(groups.begin() + g_lo, groups.begin() + g_hi);
            for (auto& x : cs) x.first_group -= (int)g_lo;
            for (auto& x : cgv) x.super -= (int)sg_lo;
            e = cudaSuccess;
            CY(d_supers.alloc(cs.size(), s));
            CY(d_groups.alloc(cgv.size(), s));
            CY(cudaMemcpyAsync(d_supers.p, cs.data(), cs.size() * sizeof(Cyc3Super), cudaMemcpyHostToDevice, s));
            CY(cudaMemcpyAsync(d_groups.p, cgv.data(), cgv.size() * sizeof(Cyc3Group), cudaMemcpyHostToDevice, s));
            CY(d_gx.alloc((size_t)cg * 3, s));
            CY(d_choice.alloc((size_t)cg, s));
            CY(d_gap.alloc((size_t)cg, s));
            CY(d_status.alloc((size_t)cposes, s));
            CY(d_keep.alloc((size_t)cposes, s));
            CY(cudaMemsetAsync(d_cnt.p, 0, 32, s));
            if (e != cudaSuccess) { rc = cuda_fail(e, "chunk setup", __FILE__, __LINE__); break; }

            Cyc3Dev d{};
            for (int m = 0; m < 3; ++m) {
                d.coords[m] = d_coords[m].p;
                d.n_conf[m] = p->n_conf[m];
                d.n_atoms[m] = p->n_atoms[m];
                d.n_react[m] = p->n_reactive[m];
                d.reactive[m] = d_react[m].p;
                d.pvec[m] = d_pvec[m].p;
                d.pmean[m] = d_pmean[m].p;
                d.pnorm[m] = d_pnorm[m].p;
            }
            d.supers = d_supers.p; d.n_supers = (int)cs.size();
            d.groups = d_groups.p; d.n_groups = (int)cg;
            d.angles = d_angles.p; d.n_angles = A;
            d.handed = p->rot_handedness >= 0 ? 1 : -1;
            d.gx = d_gx.p; d.choice = d_choice.p; d.gap = d_gap.p;
            double tc = now();
            cyc3_directions_kernel<<<(unsigned)((cs.size() + 3) / 4), 128, 0, s>>>(d);
            e = cudaGetLastError();
            if (trace) { cudaStreamSynchronize(s); t_dir += now() - tc; tc = now(); }
            if (e != cudaSuccess) { rc = cuda_fail(e, "cyc3_directions_kernel", __FILE__, __LINE__); break; }

            // ---- three block screens over (group, distinct angle pair) -------------------------
            std::vector<std::vector<int32_t>> tiles(3);
            for (int k = 0; k < 3 && !rc; ++k) {
                const int mi = pair_m[k][0], mj = pair_m[k][1];
                const int64_t n_pp = cg * n_u[k];
                CY(d_xf.alloc((size_t)n_pp * 12, s));
                CY(d_st[k].alloc((size_t)n_pp, s));
                CY(cudaMemsetAsync(d_near_count.p, 0, 16, s));
                if (e != cudaSuccess) { rc = cuda_fail(e, "pair buffers", __FILE__, __LINE__); break; }
                cyc3_pair_xf_kernel<<<(unsigned)((n_pp + 127) / 128), 128, 0, s>>>(d, mi, mj, d_ua[k].p, n_u[k], d_xf.p);
                const int tp = fc_clash_tile_poses(p->n_atoms[mj]);
                std::vector<int32_t>& tl = tiles[k];
                int64_t g = 0;
                while (g < cg) {
                    const Cyc3Super& s0 = cs[cgv[(size_t)g].super];
                    int64_t g_end = g + 1;
                    while (g_end < cg) {
                        const Cyc3Super& s1 = cs[cgv[(size_t)g_end].super];
                        if (s1.conf[mi] != s0.conf[mi] || s1.conf[mj] != s0.conf[mj]) break;
                        ++g_end;
                    }
                    for (int64_t pose = g * n_u[k]; pose < g_end * n_u[k];) {
                        int cnt = (int)std::min<int64_t>(tp, g_end * n_u[k] - pose);
                        tl.push_back(s0.conf[mi]);
                        tl.push_back(s0.conf[mj]);
                        tl.push_back((int32_t)pose);
                        tl.push_back(cnt);
                        pose += cnt;
                    }
                    g = g_end;
                }
                CY(d_tiles.alloc(tl.size(), s));
                CY(cudaMemcpyAsync(d_tiles.p, tl.data(), tl.size() * 4, cudaMemcpyHostToDevice, s));
                if (e != cudaSuccess) { rc = cuda_fail(e, "tile upload", __FILE__, __LINE__); break; }
                rc = fc_clash_screen_prepared_dev(prep[mi], d_coords[mi].p, d_coords[mj].p, p->n_conf[mj], p->n_atoms[mj], d_xf.p,
                                                  n_pp, d_tiles.p, (int64_t)tl.size() / 4, 0, /*strict=*/0, d_st[k].p, nullptr,
                                                  d_near_count.p, d_near_idx.p, d_near_dist.p, near_cap, 0, (void*)s);
                if (rc) break;
                // near-threshold block decisions -> one tie per affected pose (b = -1 - block)
                int n_near = 0;
                CY(cudaMemcpyAsync(&n_near, d_near_count.p, 4, cudaMemcpyDeviceToHost, s));
                CY(cudaStreamSynchronize(s));  // also: the host tile vector must outlive its copy
                if (e != cudaSuccess) { rc = cuda_fail(e, "block screen", __FILE__, __LINE__); break; }
                n_near = std::min(n_near, near_cap);
                if (n_near > 0) {
                    std::vector<int64_t> idx(n_near);
                    std::vector<double> dist(n_near);
                    h_st.resize((size_t)n_pp);
                    CY(cudaMemcpy(idx.data(), d_near_idx.p, (size_t)n_near * 8, cudaMemcpyDeviceToHost));
                    CY(cudaMemcpy(dist.data(), d_near_dist.p, (size_t)n_near * 8, cudaMemcpyDeviceToHost));
                    CY(cudaMemcpy(h_st.data(), d_st[k].p, (size_t)n_pp, cudaMemcpyDeviceToHost));
                    if (e != cudaSuccess) { rc = cuda_fail(e, "near list", __FILE__, __LINE__); break; }
                    for (int i = 0; i < n_near; ++i) {
                        const int64_t gg = idx[i] / n_u[k];
                        const int u = (int)(idx[i] - gg * n_u[k]);
                        for (int ai = 0; ai < A; ++ai) {
                            if (umap[k][ai] != u) continue;
                            fc_tie t;
                            t.a = (g_lo + gg) * A + ai;
                            t.b = -1 - k;
                            t.value = dist[i];
                            t.kind = FC_TIE_CLASH;
                            t.decision = (h_st[(size_t)idx[i]] & FC_STATUS_PASS) ? 0 : 1;
                            r->ties.push_back(t);
                            r->ties_total += 1;
                        }
                    }
                }
            }
            if (rc) break;
            if (trace) { cudaStreamSynchronize(s); t_clash += now() - tc; tc = now(); }

            // ---- combine + in-group similarity -------------------------------------------------
            Cyc3SimArgs a{};
            a.p = d;
            for (int k = 0; k < 3; ++k) { a.st[k] = d_st[k].p; a.umap[k] = d_umap[k].p; a.n_u[k] = n_u[k]; }
            {
                const long long n_abs = (long long)cg * 3 * u_max;
                CY(d_xf_abs.alloc((size_t)n_abs * 12, s));
                if (e != cudaSuccess) { rc = cuda_fail(e, "transform table", __FILE__, __LINE__); break; }
                cyc3_abs_xf_kernel<<<(unsigned)((n_abs + 127) / 128), 128, 0, s>>>(d, d_uang.p, u_max, (int)uang_m[0].size(),
                                                                                  (int)uang_m[1].size(), (int)uang_m[2].size(), d_xf_abs.p);
            }
            a.xf_abs = d_xf_abs.p;
            for (int m = 0; m < 3; ++m) { a.amap[m] = d_amap[m].p; a.mom[m] = d_mom[m].p; }
            a.u_max = u_max;
            a.status = d_status.p; a.keep = d_keep.p;
            a.rmsd_thr = p->rmsd_thresh; a.eps = FC_NEAR_EPS;
            a.ties = d_ties.p; a.n_ties = d_cnt.p + 1; a.tie_cap = tie_cap;
            a.pose_base = g_lo * A;
            CY(cudaMemsetAsync(d_cnt.p + 2, 0, 16, s));  // [2..5]: pass / recheck counters (2 x u64)
            cyc3_combine_kernel<<<(unsigned)((cposes + 255) / 256), 256, 0, s>>>(a, cposes, (unsigned long long*)(d_cnt.p + 2));
            const int warps = 4;
            cyc3_group_similarity_kernel<<<(unsigned)((cg + warps - 1) / warps), warps * 32, (size_t)warps * A * sizeof(int), s>>>(a);
            e = cudaGetLastError();
            // kept poses: order-preserving compaction of the keep flags on the device
            CY(d_kept.alloc((size_t)cposes, s));
            if (e == cudaSuccess && compact_pass(d_keep.p, cposes, 0, d_kept.p, d_cnt.p, s) != FC_OK) { rc = FC_ERR_CUDA; break; }
            if (want_status) {
                h_status.resize((size_t)cposes);
                CY(cudaMemcpyAsync(h_status.data(), d_status.p, (size_t)cposes, cudaMemcpyDeviceToHost, s));
            }
            CY(cudaMemcpyAsync(r->group_choice.data() + g_lo, d_choice.p, (size_t)cg * 4, cudaMemcpyDeviceToHost, s));
            CY(cudaMemcpyAsync(r->group_gap.data() + g_lo, d_gap.p, (size_t)cg * 8, cudaMemcpyDeviceToHost, s));
            int h_cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            CY(cudaMemcpyAsync(h_cnt, d_cnt.p, 32, cudaMemcpyDeviceToHost, s));
            CY(cudaStreamSynchronize(s));
            if (e != cudaSuccess) { rc = cuda_fail(e, "similarity stage", __FILE__, __LINE__); break; }
            if (trace) { t_sim += now() - tc; tc = now(); }
            unsigned long long pc[2];
            memcpy(pc, h_cnt + 2, 16);
            r->n_clash_pass += (int64_t)pc[0];
            r->n_rechecked += (int64_t)pc[1];
            std::vector<int64_t> ckept((size_t)h_cnt[0]);
            if (h_cnt[0] > 0) {
                static_assert(sizeof(long long) == sizeof(int64_t), "index width");
                e = cudaMemcpy(ckept.data(), d_kept.p, (size_t)h_cnt[0] * 8, cudaMemcpyDeviceToHost);
                if (e != cudaSuccess) { rc = cuda_fail(e, "kept readback", __FILE__, __LINE__); break; }
            }
            if (want_status) memcpy(r->status.data() + g_lo * A, h_status.data(), (size_t)cposes);
            {
                int n_t = std::min(h_cnt[1], tie_cap);
                r->ties_total += h_cnt[1];
                if (n_t > 0) {
                    std::vector<TieRecord> tmp(n_t);
                    e = cudaMemcpy(tmp.data(), d_ties.p, (size_t)n_t * sizeof(TieRecord), cudaMemcpyDeviceToHost);
                    if (e != cudaSuccess) { rc = cuda_fail(e, "tie readback", __FILE__, __LINE__); break; }
                    for (const TieRecord& t : tmp) {
                        fc_tie o;
                        o.a = t.a; o.b = t.b; o.value = t.value; o.kind = t.kind; o.decision = t.decision;
                        r->ties.push_back(o);
                    }
                }
            }
            if (trace) { t_host += now() - tc; tc = now(); }
            // ---- kept poses of this chunk ----------------------------------------------------------
            if (!ckept.empty()) {
                const int n_kept = (int)ckept.size();
                const size_t base = r->kept.size();
                if (want_coords) {
                    // the coordinates are NOT produced here (5.8 GB at BASELINE config C2): the kept poses' transforms
                    // stay on the device with the result, fc_result_kept_coords streams the atoms to the caller
                    fc_result::LazySegment sgm;
                    sgm.count = n_kept;
                    e = cudaMalloc((void**)&sgm.d_xf, (size_t)n_kept * 36 * sizeof(double));
                    if (e == cudaSuccess) e = cudaMalloc((void**)&sgm.d_conf, (size_t)n_kept * 3 * sizeof(int32_t));
                    if (e == cudaSuccess) {
                        cyc3_kept_xf_kernel<<<(unsigned)((n_kept * 3 + 127) / 128), 128, 0, s>>>(d, d_kept.p, n_kept, sgm.d_xf, sgm.d_conf);
                        e = cudaGetLastError();
                    }
                    r->lazy.push_back(sgm);  // owned by the result from here on (freed with it, also on failure)
                    CY(cudaStreamSynchronize(s));
                    if (e != cudaSuccess) { rc = cuda_fail(e, "kept transforms", __FILE__, __LINE__); break; }
                }
                r->constrained.resize((base + n_kept) * 6);
                for (int k = 0; k < n_kept; ++k) {
                    const int64_t gg = g_lo + ckept[k] / A;
                    r->kept.push_back(g_lo * A + ckept[k]);
                    memcpy(&r->constrained[(base + k) * 6], hgroups[(size_t)gg].ids, 24);
                }
            }
            if (trace) { t_mat += now() - tc; }
            sg_lo = sg_hi;
        }
        if (rc == FC_OK && !r->lazy.empty()) {  // device copies of the ensembles for the deferred placement
            cudaGetDevice(&r->lazy_device);
            r->lazy_n_mols = 3;
            for (int m = 0; m < 3 && e == cudaSuccess; ++m) {
                const size_t nb = (size_t)p->n_conf[m] * p->n_atoms[m] * 24;
                r->lazy_n_atoms[m] = p->n_atoms[m];
                e = cudaMalloc((void**)&r->lazy_coords[m], nb);
                CY(cudaMemcpyAsync(r->lazy_coords[m], d_coords[m].p, nb, cudaMemcpyDeviceToDevice, s));
            }
            CY(cudaStreamSynchronize(s));
            if (e != cudaSuccess) rc = cuda_fail(e, "ensemble copies", __FILE__, __LINE__);
        }
        for (int k = 0; k < 3; ++k) fc_clash_prep_free(prep[k], (void*)s);
#undef CY
    }
    cudaStreamSynchronize(s);
    cudaStreamDestroy(s);
    if (rc) {
        fc_result_free(r);
        return rc;
    }
    if (trace)
        fprintf(stderr, "fc_cyclical3_screen: total %.1f ms: enumerate %.1f, directions %.1f, block screens %.1f, similarity+D2H %.1f, "
                "host scan %.1f, kept/materialise %.1f\n", now() - t0, t_enum, t_dir, t_clash, t_sim, t_host, t_mat);
    r->n_kept = (int64_t)r->kept.size();
    r->n_surv = r->n_clash_pass;
    *out = r;
    return FC_OK;
}
