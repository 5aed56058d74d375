// firecode_b200 -- symmetry-corrected RMSD between structure pairs: the per-pair arithmetic of
// prism_pruner.pruner.prune_by_rmsd_rot_corr as FIRECODE calls it (embedder.py:1485-1496, ensemble.py:253,
// operators.py:626).  prism_pruner is absent from the reference tree; the algorithm restated here is
// [UNVERIFIED-RECALL] of its `rmsd_and_max_rot_corr` (the oracle, oracle/prism_pruner/pruner.py, states the same
// steps in numpy and the parity tests compare the two):
//
//   coord = copy of the later structure; for every symmetric torsion (i1, i2, i3, i4) IN ORDER, with its symmetry
//   angles (0, 120, 240 for a three-fold rotor, ...):
//     for each angle: rotate ONLY atom i4 about the i2 -> i3 axis (rotate_dihedral, pivot i3) and take the RMSD of the
//       four torsion atoms against the reference structure's (centred Kabsch, rmsd_and_max(center=True));
//     the first angle with the smallest RMSD wins; if it is not 0 the whole rotating group (the torsion's mask,
//     torsion_module.py:354-382) is rotated by it before the next torsion is looked at;
//   result = rmsd_and_max(ref[heavy], coord[heavy], center=True).
//
// One warp per pair: `coord` lives in shared memory, lanes try the angles of a torsion in parallel (one angle per
// lane, 4-atom Kabsch in registers), the warp picks the first minimum, rotates the group, and ends with the
// warp-wide covariance / Jacobi Kabsch / deviation pass of rmsd_and_max_kernel.  Per torsion the chosen angle
// index and the gap to the runner-up are returned so that a test can tell a genuine tie from a disagreement.
#include "fc_embed.cuh"

namespace fc {

struct RotCorrArgs {
    const double* structures;   // (n, N, 3)
    int n_atoms;
    const int* sel;             // (n_sel) atoms entering the final RMSD
    int n_sel;
    const int* torsions;        // (T, 4)
    int n_tors;
    const unsigned char* masks; // (T, N)
    const double* angles;       // flat
    const int* angle_off;       // (T + 1)
    const int* pairs;           // (P, 2) {reference, coord}
    long long n_pairs;
    int handed, axis_sign;
    double* rmsd_out;
    double* maxdev_out;
    int* choice_out;            // (P, T) or null
    double* gap_out;            // (P, T) or null
};

__device__ __forceinline__ double rc_wsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// RMSD of four points p against q after centring and optimal rotation (rmsd_and_max(center=True)[0])
__device__ inline double rmsd4(const double (*p)[3], const double (*q)[3]) {
    double mp[3] = {0, 0, 0}, mq[3] = {0, 0, 0};
    for (int k = 0; k < 4; ++k)
        for (int c = 0; c < 3; ++c) { mp[c] += p[k][c]; mq[c] += q[k][c]; }
    for (int c = 0; c < 3; ++c) { mp[c] /= 4.0; mq[c] /= 4.0; }
    double h[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < 4; ++k)
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) h[3 * r + c] += (p[k][r] - mp[r]) * (q[k][c] - mq[c]);
    M3 R = kabsch_from_cov(h, nullptr);
    double ss = 0.0;
    for (int k = 0; k < 4; ++k) {
        const double x[3] = {p[k][0] - mp[0], p[k][1] - mp[1], p[k][2] - mp[2]};
        const double y[3] = {q[k][0] - mq[0], q[k][1] - mq[1], q[k][2] - mq[2]};
        const double dx = (x[0] * R.m[0] + x[1] * R.m[3] + x[2] * R.m[6]) - y[0];
        const double dy = (x[0] * R.m[1] + x[1] * R.m[4] + x[2] * R.m[7]) - y[1];
        const double dz = (x[0] * R.m[2] + x[1] * R.m[5] + x[2] * R.m[8]) - y[2];
        ss += dx * dx + dy * dy + dz * dz;
    }
    return sqrt(ss / 4.0);
}

__global__ void __launch_bounds__(128) rot_corr_pairs_kernel(RotCorrArgs a) {
    extern __shared__ double s_rc[];  // per warp: coord (N, 3)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long pair = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (pair >= a.n_pairs) return;
    const int n = a.n_atoms;
    double* x = s_rc + (size_t)warp * n * 3;
    const double* ref = a.structures + (size_t)a.pairs[2 * pair] * n * 3;
    const double* src = a.structures + (size_t)a.pairs[2 * pair + 1] * n * 3;
    for (int k = lane; k < n * 3; k += 32) x[k] = src[k];
    __syncwarp();
    for (int t = 0; t < a.n_tors; ++t) {
        const int* tor = a.torsions + 4 * t;
        const int i2 = tor[1], i3 = tor[2], i4 = tor[3];
        const int a0 = a.angle_off[t], n_ang = a.angle_off[t + 1] - a0;
        const double axis[3] = {a.axis_sign * (x[3 * i2] - x[3 * i3]), a.axis_sign * (x[3 * i2 + 1] - x[3 * i3 + 1]),
                                a.axis_sign * (x[3 * i2 + 2] - x[3 * i3 + 2])};
        const double org[3] = {x[3 * i3], x[3 * i3 + 1], x[3 * i3 + 2]};
        // every lane tries its angles (round-robin); first smallest RMSD wins, as the sequential `<` scan does
        double best = 1e10, second = 1e300;
        int best_k = 0;
        for (int k0 = 0; k0 < n_ang; k0 += 32) {
            const int k = k0 + lane;
            double r = 1e300;
            if (k < n_ang) {
                double p[4][3], q[4][3];
                for (int m = 0; m < 4; ++m)
                    for (int c = 0; c < 3; ++c) { p[m][c] = ref[3 * tor[m] + c]; q[m][c] = x[3 * tor[m] + c]; }
                const M3 rot = rot_from_pointer(axis, a.angles[a0 + k], a.handed);
                const double d[3] = {q[3][0] - org[0], q[3][1] - org[1], q[3][2] - org[2]};
                double moved[3];
                m3_apply(rot, d, moved);
                for (int c = 0; c < 3; ++c) q[3][c] = moved[c] + org[c];
                r = rmsd4(p, q);
            }
            // warp arg-min with the lowest angle index on ties; the runner-up value gives the gap
            for (int l = 0; l < 32 && k0 + l < n_ang; ++l) {
                const double rl = __shfl_sync(0xffffffffu, r, l);
                if (rl < best) { second = best; best = rl; best_k = k0 + l; }
                else if (rl < second) second = rl;
            }
        }
        if (a.choice_out && lane == 0) {
            a.choice_out[pair * a.n_tors + t] = best_k;
            a.gap_out[pair * a.n_tors + t] = n_ang > 1 ? second - best : 1e300;
        }
        const double ang = a.angles[a0 + best_k];
        if (ang != 0.0) {  // rotate the whole group by the winning angle
            const M3 rot = rot_from_pointer(axis, ang, a.handed);
            const unsigned char* mask = a.masks + (size_t)t * n;
            __syncwarp();
            for (int k = lane; k < n; k += 32) {
                if (!mask[k]) continue;
                const double d[3] = {x[3 * k] - org[0], x[3 * k + 1] - org[1], x[3 * k + 2] - org[2]};
                double moved[3];
                m3_apply(rot, d, moved);
                x[3 * k] = moved[0] + org[0];
                x[3 * k + 1] = moved[1] + org[1];
                x[3 * k + 2] = moved[2] + org[2];
            }
        }
        __syncwarp();
    }
    // rmsd_and_max(ref[sel], coord[sel], center=True)
    const int ns = a.n_sel;
    double mp[3] = {0, 0, 0}, mq[3] = {0, 0, 0};
    for (int k = lane; k < ns; k += 32) {
        const int s = a.sel[k];
        for (int c = 0; c < 3; ++c) { mp[c] += ref[3 * s + c]; mq[c] += x[3 * s + c]; }
    }
    for (int c = 0; c < 3; ++c) { mp[c] = rc_wsum(mp[c]) / ns; mq[c] = rc_wsum(mq[c]) / ns; }
    double h[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = lane; k < ns; k += 32) {
        const int s = a.sel[k];
        const double p[3] = {ref[3 * s] - mp[0], ref[3 * s + 1] - mp[1], ref[3 * s + 2] - mp[2]};
        const double q[3] = {x[3 * s] - mq[0], x[3 * s + 1] - mq[1], x[3 * s + 2] - mq[2]};
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) h[3 * r + c] += p[r] * q[c];
    }
#pragma unroll
    for (int e = 0; e < 9; ++e) h[e] = rc_wsum(h[e]);
    const M3 R = kabsch_from_cov(h, nullptr);
    double ss = 0.0, mx = 0.0;
    for (int k = lane; k < ns; k += 32) {
        const int s = a.sel[k];
        const double p[3] = {ref[3 * s] - mp[0], ref[3 * s + 1] - mp[1], ref[3 * s + 2] - mp[2]};
        const double q[3] = {x[3 * s] - mq[0], x[3 * s + 1] - mq[1], x[3 * s + 2] - mq[2]};
        const double dx = (p[0] * R.m[0] + p[1] * R.m[3] + p[2] * R.m[6]) - q[0];
        const double dy = (p[0] * R.m[1] + p[1] * R.m[4] + p[2] * R.m[7]) - q[1];
        const double dz = (p[0] * R.m[2] + p[1] * R.m[5] + p[2] * R.m[8]) - q[2];
        const double d2 = dx * dx + dy * dy + dz * dz;
        ss += d2;
        mx = fmax(mx, d2);
    }
    ss = rc_wsum(ss);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) {
        a.rmsd_out[pair] = sqrt(ss / ns);
        a.maxdev_out[pair] = sqrt(mx);
    }
}

}  // namespace fc

using namespace fc;

extern "C" int fc_rmsd_rot_corr_pairs(const double* structures, int64_t n, int32_t n_atoms, const int32_t* sel, int32_t n_sel,
                                      const int32_t* torsions, int32_t n_tors, const uint8_t* masks, const double* angles,
                                      const int32_t* angle_offsets, const int32_t* pairs, int64_t n_pairs,
                                      int32_t rot_handedness, int32_t axis_sign, double* rmsd_out, double* maxdev_out,
                                      int32_t* choice_out, double* gap_out) {
    FC_REQUIRE(n >= 0 && n_pairs >= 0 && n_atoms > 0 && n_sel > 0 && n_tors >= 0, "fc_rmsd_rot_corr_pairs: bad sizes");
    if (n_pairs == 0) return FC_OK;
    FC_REQUIRE(structures && sel && pairs && rmsd_out && maxdev_out, "fc_rmsd_rot_corr_pairs: null pointer");
    FC_REQUIRE(n_tors == 0 || (torsions && masks && angles && angle_offsets), "fc_rmsd_rot_corr_pairs: null torsion tables");
    FC_REQUIRE((choice_out == nullptr) == (gap_out == nullptr), "fc_rmsd_rot_corr_pairs: choice and gap come together");
    for (int32_t k = 0; k < n_sel; ++k) FC_REQUIRE(sel[k] >= 0 && sel[k] < n_atoms, "fc_rmsd_rot_corr_pairs: selection out of range");
    for (int32_t t = 0; t < n_tors; ++t) {
        for (int m = 0; m < 4; ++m)
            FC_REQUIRE(torsions[4 * t + m] >= 0 && torsions[4 * t + m] < n_atoms, "fc_rmsd_rot_corr_pairs: torsion %d out of range", t);
        FC_REQUIRE(angle_offsets[t + 1] > angle_offsets[t], "fc_rmsd_rot_corr_pairs: torsion %d has no angles", t);
    }
    for (int64_t p = 0; p < 2 * n_pairs; ++p) FC_REQUIRE(pairs[p] >= 0 && pairs[p] < n, "fc_rmsd_rot_corr_pairs: pair out of range");
    const size_t smem = (size_t)4 * n_atoms * 24;
    FC_REQUIRE(smem <= 200 * 1024, "fc_rmsd_rot_corr_pairs: %d atoms do not fit in shared memory", n_atoms);
    sm_count();
    cudaStream_t s = nullptr;
    FC_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    int rc = FC_OK;
    {
        const int n_ang = n_tors ? angle_offsets[n_tors] : 0;
        DevBuf<double> d_x, d_ang, d_rmsd, d_dev, d_gap;
        DevBuf<int> d_sel, d_tor, d_off, d_pairs, d_choice;
        DevBuf<unsigned char> d_mask;
        cudaError_t e = cudaSuccess;
        auto ok = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
        ok(d_x.alloc((size_t)n * n_atoms * 3, s));
        ok(d_sel.alloc(n_sel, s));
        ok(d_tor.alloc((size_t)std::max(1, n_tors) * 4, s));
        ok(d_mask.alloc((size_t)std::max(1, n_tors) * n_atoms, s));
        ok(d_ang.alloc(std::max(1, n_ang), s));
        ok(d_off.alloc(n_tors + 1, s));
        ok(d_pairs.alloc((size_t)n_pairs * 2, s));
        ok(d_rmsd.alloc(n_pairs, s));
        ok(d_dev.alloc(n_pairs, s));
        if (choice_out) {
            ok(d_choice.alloc((size_t)n_pairs * std::max(1, n_tors), s));
            ok(d_gap.alloc((size_t)n_pairs * std::max(1, n_tors), s));
        }
        if (e == cudaSuccess) {
            ok(cudaMemcpyAsync(d_x.p, structures, (size_t)n * n_atoms * 24, cudaMemcpyHostToDevice, s));
            ok(cudaMemcpyAsync(d_sel.p, sel, (size_t)n_sel * 4, cudaMemcpyHostToDevice, s));
            if (n_tors) {
                ok(cudaMemcpyAsync(d_tor.p, torsions, (size_t)n_tors * 16, cudaMemcpyHostToDevice, s));
                ok(cudaMemcpyAsync(d_mask.p, masks, (size_t)n_tors * n_atoms, cudaMemcpyHostToDevice, s));
                ok(cudaMemcpyAsync(d_ang.p, angles, (size_t)n_ang * 8, cudaMemcpyHostToDevice, s));
                ok(cudaMemcpyAsync(d_off.p, angle_offsets, (size_t)(n_tors + 1) * 4, cudaMemcpyHostToDevice, s));
            } else {
                ok(cudaMemsetAsync(d_off.p, 0, 4, s));
            }
            ok(cudaMemcpyAsync(d_pairs.p, pairs, (size_t)n_pairs * 8, cudaMemcpyHostToDevice, s));
        }
        if (e == cudaSuccess) {
            RotCorrArgs a;
            a.structures = d_x.p;
            a.n_atoms = n_atoms;
            a.sel = d_sel.p;
            a.n_sel = n_sel;
            a.torsions = d_tor.p;
            a.n_tors = n_tors;
            a.masks = d_mask.p;
            a.angles = d_ang.p;
            a.angle_off = d_off.p;
            a.pairs = d_pairs.p;
            a.n_pairs = n_pairs;
            a.handed = rot_handedness;
            a.axis_sign = axis_sign;
            a.rmsd_out = d_rmsd.p;
            a.maxdev_out = d_dev.p;
            a.choice_out = choice_out ? d_choice.p : nullptr;
            a.gap_out = choice_out ? d_gap.p : nullptr;
            if (smem > 48 * 1024) ok(cudaFuncSetAttribute(rot_corr_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            rot_corr_pairs_kernel<<<(unsigned)((n_pairs + 3) / 4), 128, smem, s>>>(a);
            ok(cudaGetLastError());
            ok(cudaMemcpyAsync(rmsd_out, d_rmsd.p, (size_t)n_pairs * 8, cudaMemcpyDeviceToHost, s));
            ok(cudaMemcpyAsync(maxdev_out, d_dev.p, (size_t)n_pairs * 8, cudaMemcpyDeviceToHost, s));
            if (choice_out && n_tors) {
                ok(cudaMemcpyAsync(choice_out, d_choice.p, (size_t)n_pairs * n_tors * 4, cudaMemcpyDeviceToHost, s));
                ok(cudaMemcpyAsync(gap_out, d_gap.p, (size_t)n_pairs * n_tors * 8, cudaMemcpyDeviceToHost, s));
            }
            ok(cudaStreamSynchronize(s));
        }
        if (e != cudaSuccess) rc = cuda_fail(e, "fc_rmsd_rot_corr_pairs", __FILE__, __LINE__);
    }
    cudaStreamSynchronize(s);
    cudaStreamDestroy(s);
    return rc;
}
