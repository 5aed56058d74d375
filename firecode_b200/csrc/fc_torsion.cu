// firecode_b200 -- batched rigid-subtree torsion rotation with clash filtering.
//
// Replaces the primitive pair used by the torsion drivers of the reference
// (/root/reference/firecode/torsion_module.py:523-552, 813-856):
//     new = rotate_dihedral(coords, torsion, angle, mask=mask)          [prism_pruner.utils]
//     ok  = torsion_comp_check(new, torsion, mask, thresh)              [torsion_module.py:894-918]
// applied to every (conformer, torsion, angle) item of a regular scan (the 36 x 10 degree schedule
// of atropisomer_module.py:118 is one instance).  One CTA per item, all FP64:
//   axis = sign * (x[i2] - x[i3]);  M = rot_mat_from_pointer(axis, angle);
//   x[mask] = M (x[mask] - x[i3]) + x[i3];
//   pass <=> #{(s, m): |x_s - x_m| < thresh, m moved, s static and not i2 / i3} <= max_clashes.
#include <algorithm>
#include <vector>

#include "fc_embed.cuh"

namespace fc {

struct TorsionArgs {
    const double* coords;        // (C, N, 3)
    const int* torsions;         // (T, 4)
    const unsigned char* masks;  // (T, N) 1 = atom rotates
    const double* angles;        // (A)
    int n_conf, n_atoms, n_tors, n_angles;
    double thresh;
    int max_clashes, handed, axis_sign;
    double* out_coords;          // (C, T, A, N, 3) or null
    unsigned char* status;       // (C, T, A) FC_STATUS_*
    double* min_dist;            // (C, T, A) or null
    const short* lists;          // (T, 2, N): moved atoms, then static atoms without i2 / i3 (compacted)
    const int* list_len;         // (T, 2)
};

__global__ void __launch_bounds__(128) torsion_scan_kernel(TorsionArgs p) {
    extern __shared__ double sx[];  // rotated coordinates (N, 3)
    const long long item = blockIdx.x;
    const int ai = (int)(item % p.n_angles);
    const int t = (int)((item / p.n_angles) % p.n_tors);
    const int c = (int)(item / ((long long)p.n_angles * p.n_tors));
    const double* x = p.coords + (size_t)c * p.n_atoms * 3;
    const int* tor = p.torsions + 4 * t;
    const unsigned char* mask = p.masks + (size_t)t * p.n_atoms;
    const int i2 = tor[1], i3 = tor[2];
    __shared__ M3 rot;
    __shared__ double origin[3];
    if (threadIdx.x == 0) {
        double axis[3] = {p.axis_sign * (x[3 * i2] - x[3 * i3]), p.axis_sign * (x[3 * i2 + 1] - x[3 * i3 + 1]),
                          p.axis_sign * (x[3 * i2 + 2] - x[3 * i3 + 2])};
        rot = rot_from_pointer(axis, p.angles[ai], p.handed);
        origin[0] = x[3 * i3]; origin[1] = x[3 * i3 + 1]; origin[2] = x[3 * i3 + 2];
    }
    __syncthreads();
    for (int a = threadIdx.x; a < p.n_atoms; a += blockDim.x) {
        double v[3] = {x[3 * a], x[3 * a + 1], x[3 * a + 2]};
        if (mask[a]) {
            double d[3] = {v[0] - origin[0], v[1] - origin[1], v[2] - origin[2]}, r[3];
            m3_apply(rot, d, r);
            v[0] = r[0] + origin[0]; v[1] = r[1] + origin[1]; v[2] = r[2] + origin[2];
        }
        sx[3 * a] = v[0]; sx[3 * a + 1] = v[1]; sx[3 * a + 2] = v[2];
    }
    __syncthreads();
    if (p.out_coords) {
        double* o = p.out_coords + (size_t)item * p.n_atoms * 3;
        for (int e = threadIdx.x; e < p.n_atoms * 3; e += blockDim.x) o[e] = sx[e];
    }
    // clash count between moved atoms and static atoms (bond atoms i2, i3 excluded): only those pairs are
    // visited, through the per-torsion compacted index lists
    int clashes = 0;
    double dmin = 1e300;
    const int n = p.n_atoms;
    const short* moved = p.lists + (size_t)t * 2 * n;
    const short* stat = moved + n;
    const int n_m = p.list_len[2 * t], n_s = p.list_len[2 * t + 1];
    for (int e = threadIdx.x; e < n_m * n_s; e += blockDim.x) {
        const int s = stat[e / n_m], m = moved[e % n_m];
        double dx = sx[3 * s] - sx[3 * m], dy = sx[3 * s + 1] - sx[3 * m + 1], dz = sx[3 * s + 2] - sx[3 * m + 2];
        double d = sqrt(dx * dx + dy * dy + dz * dz);
        clashes += d < p.thresh ? 1 : 0;
        dmin = fmin(dmin, d);
    }
    __shared__ int s_cnt[4];
    __shared__ double s_min[4];
    clashes = warp_sum(clashes);
    dmin = warp_min(dmin);
    if ((threadIdx.x & 31) == 0) { s_cnt[threadIdx.x >> 5] = clashes; s_min[threadIdx.x >> 5] = dmin; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        double mn = 1e300;
        for (int w = 0; w < (blockDim.x >> 5); ++w) { tot += s_cnt[w]; mn = fmin(mn, s_min[w]); }
        unsigned char st = tot <= p.max_clashes ? FC_STATUS_PASS : 0;
        if (fabs(mn - p.thresh) <= FC_NEAR_EPS) st |= FC_STATUS_NEAR;
        p.status[item] = st;
        if (p.min_dist) p.min_dist[item] = mn;
    }
}


// ---------------------------------------------------------------------------------------------
// Conformational-search inner loop: firecode/torsion_module.py:512-552 (random_csearch) = 813-856
// (clustered_csearch).  One CTA per (starting structure, angle set); the torsions of a set are applied IN
// ORDER to the running coordinates:
//     temp = rotate_dihedral(cur, torsion, angle)                       (skipped when angle == 0)
//     if not torsion_comp_check(temp):  back off 5 degrees at a time, at most angle // 5 times, until the check
//                                       passes (then the bond counts as rotated)
//     else: the bond counts as rotated
//     cur = temp                                                        (also when the back-off never passed)
// Output: final coordinates and the number of rotated bonds of every item (the drivers keep items with > 0).
// ---------------------------------------------------------------------------------------------
struct CsearchArgs {
    const double* starts;        // (S, N, 3)
    const int* torsions;         // (T, 4)
    const unsigned char* masks;  // (T, N)
    const int* angle_sets;       // (A, T) degrees (integers, Torsion.get_angles)
    int n_starts, n_atoms, n_tors;
    long long n_sets;
    double thresh;
    int handed, axis_sign;
    const short* lists;          // (T, 2, N): moved atoms, then static atoms without i2 / i3 (compacted)
    const int* list_len;         // (T, 2)
    double* out;                 // (S, A, N, 3)
    int* rotated;                // (S, A)
    unsigned char* near;         // (S, A) some clash distance met on the way was within FC_NEAR_EPS of thresh
};

// rotate the masked atoms of x (in shared memory) in place about the i2-i3 axis through i3
__device__ __forceinline__ void csearch_rotate(double* x, const unsigned char* mask, int n, int i2, int i3, double angle,
                                               int handed, int axis_sign, M3* s_rot, double* s_origin) {
    if (threadIdx.x == 0) {
        double axis[3] = {axis_sign * (x[3 * i2] - x[3 * i3]), axis_sign * (x[3 * i2 + 1] - x[3 * i3 + 1]),
                          axis_sign * (x[3 * i2 + 2] - x[3 * i3 + 2])};
        *s_rot = rot_from_pointer(axis, angle, handed);
        s_origin[0] = x[3 * i3]; s_origin[1] = x[3 * i3 + 1]; s_origin[2] = x[3 * i3 + 2];
    }
    __syncthreads();
    for (int a = threadIdx.x; a < n; a += blockDim.x) {
        if (!mask[a]) continue;
        double d[3] = {x[3 * a] - s_origin[0], x[3 * a + 1] - s_origin[1], x[3 * a + 2] - s_origin[2]}, r[3];
        m3_apply(*s_rot, d, r);
        x[3 * a] = r[0] + s_origin[0]; x[3 * a + 1] = r[1] + s_origin[1]; x[3 * a + 2] = r[2] + s_origin[2];
    }
    __syncthreads();
}

// torsion_comp_check with max_clashes = 0: no (static, moved) pair closer than thresh; bond atoms excluded
__device__ __forceinline__ bool csearch_check(const double* x, const short* moved, int n_m, const short* stat, int n_s,
                                              double thresh, int* near_flag) {
    bool clash = false, near = false;
    for (int e = threadIdx.x; e < n_m * n_s; e += blockDim.x) {
        const int s = stat[e / n_m], m = moved[e % n_m];
        double dx = x[3 * s] - x[3 * m], dy = x[3 * s + 1] - x[3 * m + 1], dz = x[3 * s + 2] - x[3 * m + 2];
        double d = sqrt(dx * dx + dy * dy + dz * dz);
        clash = clash || d < thresh;
        near = near || fabs(d - thresh) <= FC_NEAR_EPS;
    }
    if (near) *near_flag = 1;
    return !__syncthreads_or(clash ? 1 : 0);
}

__global__ void __launch_bounds__(128) csearch_apply_kernel(CsearchArgs p) {
    extern __shared__ double s_x[];  // running coordinates (N, 3)
    __shared__ M3 s_rot;
    __shared__ double s_origin[3];
    __shared__ int s_near;
    const long long item = blockIdx.x;
    const long long set = item % p.n_sets;
    const int start = (int)(item / p.n_sets);
    const int n = p.n_atoms;
    const double* src = p.starts + (size_t)start * n * 3;
    for (int e = threadIdx.x; e < n * 3; e += blockDim.x) s_x[e] = src[e];
    if (threadIdx.x == 0) s_near = 0;
    __syncthreads();
    int rotated = 0;
    for (int t = 0; t < p.n_tors; ++t) {
        const int angle = p.angle_sets[set * p.n_tors + t];
        if (angle == 0) continue;  // uniform over the block
        const int* tor = p.torsions + 4 * t;
        const unsigned char* mask = p.masks + (size_t)t * n;
        const int i2 = tor[1], i3 = tor[2];
        const short* moved = p.lists + (size_t)t * 2 * n;
        const short* stat = moved + n;
        const int n_m = p.list_len[2 * t], n_s = p.list_len[2 * t + 1];
        csearch_rotate(s_x, mask, n, i2, i3, (double)angle, p.handed, p.axis_sign, &s_rot, s_origin);
        if (csearch_check(s_x, moved, n_m, stat, n_s, p.thresh, &s_near)) {
            ++rotated;
        } else {
            // python: range(angle // 5) -- floor division, empty for negative angles
            const int steps = angle >= 0 ? angle / 5 : 0;
            for (int b = 0; b < steps; ++b) {
                csearch_rotate(s_x, mask, n, i2, i3, -5.0, p.handed, p.axis_sign, &s_rot, s_origin);
                if (csearch_check(s_x, moved, n_m, stat, n_s, p.thresh, &s_near)) {
                    ++rotated;
                    break;
                }
            }
        }
    }
    double* o = p.out + (size_t)item * n * 3;
    for (int e = threadIdx.x; e < n * 3; e += blockDim.x) o[e] = s_x[e];
    if (threadIdx.x == 0) {
        p.rotated[item] = rotated;
        p.near[item] = (unsigned char)(s_near ? 1 : 0);
    }
}

}  // namespace fc

using namespace fc;

extern "C" int fc_torsion_scan(const double* coords, int32_t n_conf, int32_t n_atoms, const int32_t* torsions,
                               int32_t n_tors, const uint8_t* masks, const double* angles, int32_t n_angles,
                               double thresh, int32_t max_clashes, int32_t rot_handedness, int32_t axis_sign,
                               double* out_coords, uint8_t* status_out, double* min_dist_out) {
    FC_REQUIRE(n_conf >= 0 && n_atoms > 0 && n_tors >= 0 && n_angles >= 0, "fc_torsion_scan: bad sizes");
    const int64_t items = (int64_t)n_conf * n_tors * n_angles;
    if (items == 0) return FC_OK;
    FC_REQUIRE(items < ((int64_t)1 << 31), "fc_torsion_scan: too many items (%lld)", (long long)items);
    FC_REQUIRE(n_atoms < 32768, "fc_torsion_scan: more than 32767 atoms");
    FC_REQUIRE(coords && torsions && masks && angles && status_out, "fc_torsion_scan: null pointer");
    for (int t = 0; t < n_tors * 4; ++t)
        FC_REQUIRE(torsions[t] >= 0 && torsions[t] < n_atoms, "fc_torsion_scan: torsion index out of range");
    sm_count();
    cudaStream_t s;
    FC_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    int rc = FC_OK;
    {
        DevBuf<double> d_coords, d_angles, d_out, d_min;
        DevBuf<int> d_tors;
        DevBuf<unsigned char> d_masks, d_status;
        cudaError_t e = cudaSuccess;
#define TS(call) do { if (e == cudaSuccess) e = (call); } while (0)
        TS(d_coords.alloc((size_t)n_conf * n_atoms * 3, s));
        TS(cudaMemcpyAsync(d_coords.p, coords, (size_t)n_conf * n_atoms * 24, cudaMemcpyHostToDevice, s));
        TS(d_tors.alloc((size_t)n_tors * 4, s));
        TS(cudaMemcpyAsync(d_tors.p, torsions, (size_t)n_tors * 16, cudaMemcpyHostToDevice, s));
        TS(d_masks.alloc((size_t)n_tors * n_atoms, s));
        TS(cudaMemcpyAsync(d_masks.p, masks, (size_t)n_tors * n_atoms, cudaMemcpyHostToDevice, s));
        TS(d_angles.alloc(n_angles, s));
        TS(cudaMemcpyAsync(d_angles.p, angles, (size_t)n_angles * 8, cudaMemcpyHostToDevice, s));
        std::vector<short> h_lists((size_t)std::max(n_tors, 1) * 2 * n_atoms, 0);
        std::vector<int> h_len((size_t)std::max(n_tors, 1) * 2, 0);
        for (int t = 0; t < n_tors; ++t) {
            const int i2 = torsions[4 * t + 1], i3 = torsions[4 * t + 2];
            short* moved = h_lists.data() + (size_t)t * 2 * n_atoms;
            short* stat = moved + n_atoms;
            int nm = 0, ns = 0;
            for (int a = 0; a < n_atoms; ++a) {
                if (masks[(size_t)t * n_atoms + a]) moved[nm++] = (short)a;
                else if (a != i2 && a != i3) stat[ns++] = (short)a;
            }
            h_len[2 * t] = nm;
            h_len[2 * t + 1] = ns;
        }
        DevBuf<short> d_lists;
        DevBuf<int> d_len;
        TS(d_lists.alloc(h_lists.size(), s));
        TS(d_len.alloc(h_len.size(), s));
        TS(cudaMemcpyAsync(d_lists.p, h_lists.data(), h_lists.size() * sizeof(short), cudaMemcpyHostToDevice, s));
        TS(cudaMemcpyAsync(d_len.p, h_len.data(), h_len.size() * sizeof(int), cudaMemcpyHostToDevice, s));
        TS(d_status.alloc((size_t)items, s));
        if (out_coords) TS(d_out.alloc((size_t)items * n_atoms * 3, s));
        if (min_dist_out) TS(d_min.alloc((size_t)items, s));
        if (e == cudaSuccess) {
            TorsionArgs a{d_coords.p, d_tors.p, d_masks.p, d_angles.p, n_conf, n_atoms, n_tors, n_angles, thresh,
                          max_clashes, rot_handedness >= 0 ? 1 : -1, axis_sign >= 0 ? 1 : -1,
                          out_coords ? d_out.p : nullptr, d_status.p, min_dist_out ? d_min.p : nullptr, d_lists.p, d_len.p};
            size_t smem = (size_t)n_atoms * 24;
            if (smem > 48 * 1024)
                TS(cudaFuncSetAttribute(torsion_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (e == cudaSuccess) {
                torsion_scan_kernel<<<(unsigned)items, 128, smem, s>>>(a);
                e = cudaGetLastError();
            }
        }
        TS(cudaMemcpyAsync(status_out, d_status.p, (size_t)items, cudaMemcpyDeviceToHost, s));
        if (out_coords) TS(cudaMemcpyAsync(out_coords, d_out.p, (size_t)items * n_atoms * 24, cudaMemcpyDeviceToHost, s));
        if (min_dist_out) TS(cudaMemcpyAsync(min_dist_out, d_min.p, (size_t)items * 8, cudaMemcpyDeviceToHost, s));
        TS(cudaStreamSynchronize(s));
#undef TS
        if (e != cudaSuccess) rc = cuda_fail(e, "fc_torsion_scan", __FILE__, __LINE__);
    }
    cudaStreamSynchronize(s);
    cudaStreamDestroy(s);
    return rc;
}

extern "C" int fc_csearch_apply(const double* starts, int32_t n_starts, int32_t n_atoms, const int32_t* torsions,
                                int32_t n_tors, const uint8_t* masks, const int32_t* angle_sets, int64_t n_sets,
                                double thresh, int32_t rot_handedness, int32_t axis_sign, double* out_coords,
                                int32_t* rotated_out, uint8_t* near_out) {
    FC_REQUIRE(n_starts >= 0 && n_atoms > 0 && n_tors >= 0 && n_sets >= 0, "fc_csearch_apply: bad sizes");
    const int64_t items = (int64_t)n_starts * n_sets;
    if (items == 0) return FC_OK;
    FC_REQUIRE(items < ((int64_t)1 << 31), "fc_csearch_apply: too many items (%lld)", (long long)items);
    FC_REQUIRE(n_atoms < 32768, "fc_csearch_apply: more than 32767 atoms");
    FC_REQUIRE(starts && out_coords && rotated_out && near_out && (n_tors == 0 || (torsions && masks && angle_sets)),
               "fc_csearch_apply: null pointer");
    for (int t = 0; t < n_tors; ++t) {
        for (int k = 0; k < 4; ++k)
            FC_REQUIRE(torsions[4 * t + k] >= 0 && torsions[4 * t + k] < n_atoms, "fc_csearch_apply: torsion index out of range");
        FC_REQUIRE(!masks[(size_t)t * n_atoms + torsions[4 * t + 1]] && !masks[(size_t)t * n_atoms + torsions[4 * t + 2]],
                   "fc_csearch_apply: the bond atoms of torsion %d must not be in its rotation mask", t);
    }
    sm_count();
    cudaStream_t s;
    FC_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    int rc = FC_OK;
    {
        DevBuf<double> d_starts, d_out;
        DevBuf<int> d_tors, d_sets, d_rot;
        DevBuf<unsigned char> d_masks, d_near;
        cudaError_t e = cudaSuccess;
#define CS(call) do { if (e == cudaSuccess) e = (call); } while (0)
        CS(d_starts.alloc((size_t)n_starts * n_atoms * 3, s));
        CS(cudaMemcpyAsync(d_starts.p, starts, (size_t)n_starts * n_atoms * 24, cudaMemcpyHostToDevice, s));
        CS(d_tors.alloc((size_t)std::max(n_tors, 1) * 4, s));
        CS(d_masks.alloc((size_t)std::max(n_tors, 1) * n_atoms, s));
        CS(d_sets.alloc((size_t)n_sets * std::max(n_tors, 1), s));
        if (n_tors) {
            CS(cudaMemcpyAsync(d_tors.p, torsions, (size_t)n_tors * 16, cudaMemcpyHostToDevice, s));
            CS(cudaMemcpyAsync(d_masks.p, masks, (size_t)n_tors * n_atoms, cudaMemcpyHostToDevice, s));
            CS(cudaMemcpyAsync(d_sets.p, angle_sets, (size_t)n_sets * n_tors * 4, cudaMemcpyHostToDevice, s));
        }
        std::vector<short> h_lists((size_t)std::max(n_tors, 1) * 2 * n_atoms, 0);
        std::vector<int> h_len((size_t)std::max(n_tors, 1) * 2, 0);
        for (int t = 0; t < n_tors; ++t) {
            const int i2 = torsions[4 * t + 1], i3 = torsions[4 * t + 2];
            short* moved = h_lists.data() + (size_t)t * 2 * n_atoms;
            short* stat = moved + n_atoms;
            int nm = 0, ns = 0;
            for (int a = 0; a < n_atoms; ++a) {
                if (masks[(size_t)t * n_atoms + a]) moved[nm++] = (short)a;
                else if (a != i2 && a != i3) stat[ns++] = (short)a;
            }
            h_len[2 * t] = nm;
            h_len[2 * t + 1] = ns;
        }
        DevBuf<short> d_lists;
        DevBuf<int> d_len;
        CS(d_lists.alloc(h_lists.size(), s));
        CS(d_len.alloc(h_len.size(), s));
        CS(cudaMemcpyAsync(d_lists.p, h_lists.data(), h_lists.size() * sizeof(short), cudaMemcpyHostToDevice, s));
        CS(cudaMemcpyAsync(d_len.p, h_len.data(), h_len.size() * sizeof(int), cudaMemcpyHostToDevice, s));
        CS(d_out.alloc((size_t)items * n_atoms * 3, s));
        CS(d_rot.alloc((size_t)items, s));
        CS(d_near.alloc((size_t)items, s));
        if (e == cudaSuccess) {
            CsearchArgs a{d_starts.p, d_tors.p, d_masks.p, d_sets.p, n_starts, n_atoms, n_tors, (long long)n_sets, thresh,
                          rot_handedness >= 0 ? 1 : -1, axis_sign >= 0 ? 1 : -1, d_lists.p, d_len.p, d_out.p, d_rot.p, d_near.p};
            size_t smem = (size_t)n_atoms * 24;
            if (smem > 48 * 1024)
                CS(cudaFuncSetAttribute(csearch_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (e == cudaSuccess) {
                csearch_apply_kernel<<<(unsigned)items, 128, smem, s>>>(a);
                e = cudaGetLastError();
            }
        }
        CS(cudaMemcpyAsync(out_coords, d_out.p, (size_t)items * n_atoms * 24, cudaMemcpyDeviceToHost, s));
        CS(cudaMemcpyAsync(rotated_out, d_rot.p, (size_t)items * 4, cudaMemcpyDeviceToHost, s));
        CS(cudaMemcpyAsync(near_out, d_near.p, (size_t)items, cudaMemcpyDeviceToHost, s));
        CS(cudaStreamSynchronize(s));
#undef CS
        if (e != cudaSuccess) rc = cuda_fail(e, "fc_csearch_apply", __FILE__, __LINE__);
    }
    cudaStreamSynchronize(s);
    cudaStreamDestroy(s);
    return rc;
}

