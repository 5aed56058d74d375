// firecode_b200 -- FP64 device geometry shared by the embed / pruning / torsion kernels.
// Every routine restates a reference formula (cited) in double precision.
#pragma once

#include "fc_common.cuh"

namespace fc {

struct M3 {
    double m[9];  // row-major
};

__host__ __device__ __forceinline__ M3 m3_identity() {
    M3 r;
    r.m[0] = 1; r.m[1] = 0; r.m[2] = 0;
    r.m[3] = 0; r.m[4] = 1; r.m[5] = 0;
    r.m[6] = 0; r.m[7] = 0; r.m[8] = 1;
    return r;
}

__host__ __device__ __forceinline__ M3 m3_mul(const M3& a, const M3& b) {
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            r.m[3 * i + j] = a.m[3 * i] * b.m[j] + a.m[3 * i + 1] * b.m[3 + j] + a.m[3 * i + 2] * b.m[6 + j];
    return r;
}

__host__ __device__ __forceinline__ M3 m3_transpose(const M3& a) {
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) r.m[3 * i + j] = a.m[3 * j + i];
    return r;
}

__host__ __device__ __forceinline__ void m3_apply(const M3& a, const double* v, double* out) {
    double x = v[0], y = v[1], z = v[2];
    out[0] = a.m[0] * x + a.m[1] * y + a.m[2] * z;
    out[1] = a.m[3] * x + a.m[4] * y + a.m[5] * z;
    out[2] = a.m[6] * x + a.m[7] * y + a.m[8] * z;
}

__host__ __device__ __forceinline__ double det3(const M3& a) {
    return a.m[0] * (a.m[4] * a.m[8] - a.m[5] * a.m[7]) - a.m[1] * (a.m[3] * a.m[8] - a.m[5] * a.m[6]) +
           a.m[2] * (a.m[3] * a.m[7] - a.m[4] * a.m[6]);
}

// prism_pruner.algebra.rot_mat_from_pointer (contract: SURVEY.md 8c): rotation by angle_deg about
// `axis` through the unit quaternion (x, y, z, w); handed = +1 right-handed, -1 left-handed.
__host__ __device__ __forceinline__ M3 rot_from_pointer(const double* axis, double angle_deg, int handed) {
    const double kPi = 3.14159265358979323846;
    double n = sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
    double half = (double)handed * angle_deg * kPi / 180.0 / 2.0;
    double s = sin(half), w = cos(half);
    double x = axis[0] / n * s, y = axis[1] / n * s, z = axis[2] / n * s;
    double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
    double xy = x * y, zw = z * w, xz = x * z, yw = y * w, yz = y * z, xw = x * w;
    M3 r;
    r.m[0] = x2 - y2 - z2 + w2;
    r.m[3] = 2 * (xy + zw);
    r.m[6] = 2 * (xz - yw);
    r.m[1] = 2 * (xy - zw);
    r.m[4] = -x2 + y2 - z2 + w2;
    r.m[7] = 2 * (yz + xw);
    r.m[2] = 2 * (xz + yw);
    r.m[5] = 2 * (yz - xw);
    r.m[8] = -x2 - y2 + z2 + w2;
    return r;
}

// firecode/utils.py:224-249 rotation_matrix_from_vectors(vec1, vec2)
__host__ __device__ __forceinline__ M3 rot_vec_to_vec(const double* v1, const double* v2, int handed) {
    double n1 = sqrt(v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2]);
    double n2 = sqrt(v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2]);
    double a[3] = {v1[0] / n1, v1[1] / n1, v1[2] / n1};
    double b[3] = {v2[0] / n2, v2[1] / n2, v2[2] / n2};
    double v[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    double s = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (s != 0.0) {
        double c = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
        M3 k;
        k.m[0] = 0; k.m[1] = -v[2]; k.m[2] = v[1];
        k.m[3] = v[2]; k.m[4] = 0; k.m[5] = -v[0];
        k.m[6] = -v[1]; k.m[7] = v[0]; k.m[8] = 0;
        M3 k2 = m3_mul(k, k);
        double f = (1.0 - c) / (s * s);
        M3 r = m3_identity();
#pragma unroll
        for (int i = 0; i < 9; ++i) r.m[i] = r.m[i] + k.m[i] + k2.m[i] * f;
        return r;
    }
    double sx = a[0] + b[0], sy = a[1] + b[1], sz = a[2] + b[2];
    if (sqrt(sx * sx + sy * sy + sz * sz) == 0.0) {
        double zaxis[3] = {0.0, 0.0, 1.0};
        return rot_from_pointer(zaxis, 180.0, handed);
    }
    return m3_identity();
}

// prism_pruner.algebra.dihedral (degrees, atan2 form)
__host__ __device__ __forceinline__ double dihedral_deg(const double* p0, const double* p1, const double* p2,
                                                        const double* p3) {
    const double kPi = 3.14159265358979323846;
    double b0[3] = {-(p1[0] - p0[0]), -(p1[1] - p0[1]), -(p1[2] - p0[2])};
    double b1[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
    double b2[3] = {p3[0] - p2[0], p3[1] - p2[1], p3[2] - p2[2]};
    double n = sqrt(b1[0] * b1[0] + b1[1] * b1[1] + b1[2] * b1[2]);
    b1[0] /= n; b1[1] /= n; b1[2] /= n;
    double d0 = b0[0] * b1[0] + b0[1] * b1[1] + b0[2] * b1[2];
    double d2 = b2[0] * b1[0] + b2[1] * b1[1] + b2[2] * b1[2];
    double v[3] = {b0[0] - d0 * b1[0], b0[1] - d0 * b1[1], b0[2] - d0 * b1[2]};
    double w[3] = {b2[0] - d2 * b1[0], b2[1] - d2 * b1[1], b2[2] - d2 * b1[2]};
    double x = v[0] * w[0] + v[1] * w[1] + v[2] * w[2];
    double c[3] = {b1[1] * v[2] - b1[2] * v[1], b1[2] * v[0] - b1[0] * v[2], b1[0] * v[1] - b1[1] * v[0]};
    double y = c[0] * w[0] + c[1] * w[1] + c[2] * w[2];
    return atan2(y, x) * 180.0 / kPi;
}

// Symmetric 3x3 eigen-decomposition (cyclic Jacobi, FP64).  a: symmetric matrix (row-major), on
// exit eigenvalues in w (unsorted) and eigenvectors in the COLUMNS of v.
__host__ __device__ inline void jacobi_eig3(const double* a_in, double* w, double* v) {
    double a[9];
    for (int i = 0; i < 9; ++i) a[i] = a_in[i];
    v[0] = 1; v[1] = 0; v[2] = 0; v[3] = 0; v[4] = 1; v[5] = 0; v[6] = 0; v[7] = 0; v[8] = 1;
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = a[1] * a[1] + a[2] * a[2] + a[5] * a[5];
        double diag = a[0] * a[0] + a[4] * a[4] + a[8] * a[8];
        if (off <= 1e-32 * diag || off == 0.0) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double apq = a[3 * p + q];
                if (apq == 0.0) continue;
                double theta = (a[3 * q + q] - a[3 * p + p]) / (2.0 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {  // A <- A J
                    double akp = a[3 * k + p], akq = a[3 * k + q];
                    a[3 * k + p] = c * akp - s * akq;
                    a[3 * k + q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {  // A <- J^T A
                    double apk = a[3 * p + k], aqk = a[3 * q + k];
                    a[3 * p + k] = c * apk - s * aqk;
                    a[3 * q + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {  // V <- V J
                    double vkp = v[3 * k + p], vkq = v[3 * k + q];
                    v[3 * k + p] = c * vkp - s * vkq;
                    v[3 * k + q] = s * vkp + c * vkq;
                }
            }
    }
    w[0] = a[0]; w[1] = a[4]; w[2] = a[8];
}

// Sum of the singular values of a 3x3 matrix H with the sign of the smallest set by det(H) -- the quantity the
// Kabsch RMSD needs, msd = (|p|^2 + |q|^2 - 2 * sum) / n -- from the closed-form (trigonometric) eigenvalues of
// H^T H.  About ten times cheaper than the Jacobi solve; used as a SCREEN only: callers fall back to
// kabsch_from_cov when the value lands within a guard band of their threshold (the smallest singular value
// carries an absolute error of about 3e-8 * sigma_1).
__host__ __device__ inline double singular_sum3(const double* h) {
    double k[6];  // upper triangle of H^T H: 00 01 02 11 12 22
    k[0] = h[0] * h[0] + h[3] * h[3] + h[6] * h[6];
    k[1] = h[0] * h[1] + h[3] * h[4] + h[6] * h[7];
    k[2] = h[0] * h[2] + h[3] * h[5] + h[6] * h[8];
    k[3] = h[1] * h[1] + h[4] * h[4] + h[7] * h[7];
    k[4] = h[1] * h[2] + h[4] * h[5] + h[7] * h[8];
    k[5] = h[2] * h[2] + h[5] * h[5] + h[8] * h[8];
    double e1, e2, e3;
    const double p1 = k[1] * k[1] + k[2] * k[2] + k[4] * k[4];
    const double q = (k[0] + k[3] + k[5]) / 3.0;
    const double b0 = k[0] - q, b3 = k[3] - q, b5 = k[5] - q;
    const double p2 = b0 * b0 + b3 * b3 + b5 * b5 + 2.0 * p1;
    if (p2 <= 0.0) {
        e1 = e2 = e3 = q;
    } else {
        const double p = sqrt(p2 / 6.0), ip = 1.0 / p;
        const double c0 = b0 * ip, c3 = b3 * ip, c5 = b5 * ip, c1 = k[1] * ip, c2 = k[2] * ip, c4 = k[4] * ip;
        double r = 0.5 * (c0 * (c3 * c5 - c4 * c4) - c1 * (c1 * c5 - c4 * c2) + c2 * (c1 * c4 - c3 * c2));
        r = fmin(1.0, fmax(-1.0, r));
        const double phi = acos(r) / 3.0;
        e1 = q + 2.0 * p * cos(phi);
        e3 = q + 2.0 * p * cos(phi + 2.0943951023931954923);
        e2 = 3.0 * q - e1 - e3;
    }
    const double det = h[0] * (h[4] * h[8] - h[5] * h[7]) - h[1] * (h[3] * h[8] - h[5] * h[6]) + h[2] * (h[3] * h[7] - h[4] * h[6]);
    const double s3 = sqrt(fmax(e3, 0.0));
    return sqrt(fmax(e1, 0.0)) + sqrt(fmax(e2, 0.0)) + (det < 0.0 ? -s3 : s3);
}

// Kabsch rotation for a 3x3 cross-covariance H (row-major): returns the proper rotation R = U D V^T
// of the SVD H = U S V^T with D = diag(1, 1, det(U V^T)) -- what numpy's svd + sign fix yields in
// firecode/algebra.py:42-49 (align_vec_pair, H = ref^T tgt) and prism_pruner.rmsd.get_alignment_
// matrix (H = p^T q).  Built from the two leading singular triplets:
//      R = u1 v1^T + u2 v2^T + (u1 x u2)(v1 x v2)^T
// which is exact also for rank-2 H (align_vec_pair always is) and for reflections.
// Also returns the singular values (descending) with the sign of the third set by det(H).
__host__ __device__ inline M3 kabsch_from_cov(const double* h, double* sig /*3, may be null*/) {
    double hth[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            hth[3 * i + j] = h[i] * h[j] + h[3 + i] * h[3 + j] + h[6 + i] * h[6 + j];
    double w[3], v[9];
    jacobi_eig3(hth, w, v);
    int o[3] = {0, 1, 2};
    if (w[o[0]] < w[o[1]]) { int t = o[0]; o[0] = o[1]; o[1] = t; }
    if (w[o[1]] < w[o[2]]) { int t = o[1]; o[1] = o[2]; o[2] = t; }
    if (w[o[0]] < w[o[1]]) { int t = o[0]; o[0] = o[1]; o[1] = t; }
    double v1[3] = {v[o[0]], v[3 + o[0]], v[6 + o[0]]};
    double v2[3] = {v[o[1]], v[3 + o[1]], v[6 + o[1]]};
    double u1[3], u2[3];
    for (int i = 0; i < 3; ++i) {
        u1[i] = h[3 * i] * v1[0] + h[3 * i + 1] * v1[1] + h[3 * i + 2] * v1[2];
        u2[i] = h[3 * i] * v2[0] + h[3 * i + 1] * v2[1] + h[3 * i + 2] * v2[2];
    }
    double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
    if (!(n1 > 1e-150)) {  // H = 0 (e.g. a single selected atom after centring): every rotation is optimal
        M3 id;
        for (int i = 0; i < 9; ++i) id.m[i] = (i % 4 == 0) ? 1.0 : 0.0;
        if (sig) sig[0] = sig[1] = sig[2] = 0.0;
        return id;
    }
    for (int i = 0; i < 3; ++i) u1[i] /= n1;
    // Gram-Schmidt keeps u2 orthogonal to u1 when the second singular value is tiny
    double d = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];
    for (int i = 0; i < 3; ++i) u2[i] -= d * u1[i];
    double n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
    if (n2 > 1e-13 * n1) {
        for (int i = 0; i < 3; ++i) u2[i] /= n2;
    } else {
        // rank <= 1 (collinear atoms): any unit vector orthogonal to u1 completes an optimal rotation (numpy's SVD
        // returns one such completion too); take the coordinate axis least aligned with u1
        int k = fabs(u1[0]) <= fabs(u1[1]) ? (fabs(u1[0]) <= fabs(u1[2]) ? 0 : 2) : (fabs(u1[1]) <= fabs(u1[2]) ? 1 : 2);
        double e[3] = {0.0, 0.0, 0.0};
        e[k] = 1.0;
        const double de = u1[k];
        for (int i = 0; i < 3; ++i) u2[i] = e[i] - de * u1[i];
        n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
        for (int i = 0; i < 3; ++i) u2[i] /= n2;
    }
    double u3[3] = {u1[1] * u2[2] - u1[2] * u2[1], u1[2] * u2[0] - u1[0] * u2[2], u1[0] * u2[1] - u1[1] * u2[0]};
    double v3[3] = {v1[1] * v2[2] - v1[2] * v2[1], v1[2] * v2[0] - v1[0] * v2[2], v1[0] * v2[1] - v1[1] * v2[0]};
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.m[3 * i + j] = u1[i] * v1[j] + u2[i] * v2[j] + u3[i] * v3[j];
    if (sig) {
        sig[0] = sqrt(fmax(w[o[0]], 0.0));
        sig[1] = sqrt(fmax(w[o[1]], 0.0));
        M3 hm;
        for (int i = 0; i < 9; ++i) hm.m[i] = h[i];
        double s3 = sqrt(fmax(w[o[2]], 0.0));
        sig[2] = det3(hm) < 0 ? -s3 : s3;
    }
    return r;
}

}  // namespace fc
