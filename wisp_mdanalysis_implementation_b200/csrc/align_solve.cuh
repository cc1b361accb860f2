// align_solve.cuh -- optimal rotation of one frame from its 3x3 landmark correlation matrix (K7).
//
// Replaces MDAnalysis.analysis.align.rotation_matrix as called at
// trajectory_analysis_functions.py:102 (QCP: largest eigenvalue of Horn's 4x4 quaternion matrix,
// eigenvector -> rotation).  Same minimiser, different arithmetic:
//   1. lambda_max by Newton's method on the characteristic quartic from an upper bound (monotone from
//      above, the QCP idea) -- a handful of dependent FP64 operations instead of ~6 Jacobi sweeps;
//   2. eigenvector = the best-conditioned column of adj(N - lambda I);
//   3. one Rayleigh-quotient polish (mu = q^T N q, column of adj(N - mu I) again): the polynomial
//      root is only good to ~1e-11 relative, the Rayleigh quotient of the resulting vector is exact
//      to rounding, and so is the vector recomputed from it;
//   4. degenerate spectra (lambda_1 ~ lambda_2: collinear landmarks, 180-degree ambiguities) make adj
//      vanish -> cyclic Jacobi as the fallback (the round-1 solver, kept verbatim).
// Everything is __host__ __device__ so that tools/align_solve_check.cpp can hold it against a
// reference eigen-solver on the CPU; the kernels call horn_rotation() from one lane per frame.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define WISP_HD __host__ __device__ __forceinline__
#define WISP_HD_NOINLINE __host__ __device__ __noinline__
#else
#define WISP_HD inline
#define WISP_HD_NOINLINE inline
#endif

WISP_HD_NOINLINE void jacobi4(double A[4][4], double V[4][4]) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 32; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < 4; ++p)
            for (int q = p + 1; q < 4; ++q) off += A[p][q] * A[p][q];
        double diag = 0.0;
        for (int p = 0; p < 4; ++p) diag += A[p][p] * A[p][p];
        if (off <= 1e-32 * diag || off == 0.0) break;
        for (int p = 0; p < 4; ++p)
            for (int q = p + 1; q < 4; ++q) {
                if (A[p][q] == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 4; ++k) {
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq;
                    A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 4; ++k) {
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk;
                    A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 4; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
    }
}

// symmetric 4x4 held as its 10 upper entries
struct Sym4 {
    double a00, a01, a02, a03, a11, a12, a13, a22, a23, a33;
};

WISP_HD double det3(double a, double b, double c, double d, double e, double f, double g, double h, double i) {
    return a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
}

// column `col` of adj(M) for symmetric M (adj is symmetric too): out[r] = cofactor C(r, col)
WISP_HD void adj_column(const Sym4& m, int col, double out[4]) {
    const double M[4][4] = {{m.a00, m.a01, m.a02, m.a03}, {m.a01, m.a11, m.a12, m.a13},
                            {m.a02, m.a12, m.a22, m.a23}, {m.a03, m.a13, m.a23, m.a33}};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 4; ++r) {
        // minor: delete row r and column col
        double s[9];
        int k = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < 4; ++i) {
            if (i == r) continue;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int j = 0; j < 4; ++j) {
                if (j == col) continue;
                s[k++] = M[i][j];
            }
        }
        const double d = det3(s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], s[8]);
        out[r] = ((r + col) & 1) ? -d : d;
    }
}

WISP_HD double adj_diag(const Sym4& m, int j) {
    switch (j) {
        case 0: return det3(m.a11, m.a12, m.a13, m.a12, m.a22, m.a23, m.a13, m.a23, m.a33);
        case 1: return det3(m.a00, m.a02, m.a03, m.a02, m.a22, m.a23, m.a03, m.a23, m.a33);
        case 2: return det3(m.a00, m.a01, m.a03, m.a01, m.a11, m.a13, m.a03, m.a13, m.a33);
        default: return det3(m.a00, m.a01, m.a02, m.a01, m.a11, m.a12, m.a02, m.a12, m.a22);
    }
}

WISP_HD void adj_column_sw(const Sym4& m, int col, double out[4]) {
    switch (col) {
        case 0: adj_column(m, 0, out); break;
        case 1: adj_column(m, 1, out); break;
        case 2: adj_column(m, 2, out); break;
        default: adj_column(m, 3, out); break;
    }
}

WISP_HD void quaternion_to_rotation(double q0, double qx, double qy, double qz, double R[9]) {
    const double nq = sqrt(q0 * q0 + qx * qx + qy * qy + qz * qz);
    q0 /= nq; qx /= nq; qy /= nq; qz /= nq;
    R[0] = q0 * q0 + qx * qx - qy * qy - qz * qz; R[1] = 2 * (qx * qy - q0 * qz); R[2] = 2 * (qx * qz + q0 * qy);
    R[3] = 2 * (qy * qx + q0 * qz); R[4] = q0 * q0 - qx * qx + qy * qy - qz * qz; R[5] = 2 * (qy * qz - q0 * qx);
    R[6] = 2 * (qz * qx - q0 * qy); R[7] = 2 * (qz * qy + q0 * qx); R[8] = q0 * q0 - qx * qx - qy * qy + qz * qz;
}

// s[9] = S = sum_i a_i b_i^T (a = the frame's landmarks, b = the reference), row-major.
// R (row-major) is the proper rotation minimising sum_i |R a_i - b_i|^2.  Returns 0 when the fast
// path was taken, 1 for the Jacobi fallback.
WISP_HD int horn_rotation(const double s[9], double R[9]) {
    const double Sxx = s[0], Sxy = s[1], Sxz = s[2], Syx = s[3], Syy = s[4], Syz = s[5],
                 Szx = s[6], Szy = s[7], Szz = s[8];
    Sym4 n;
    n.a00 = Sxx + Syy + Szz; n.a01 = Syz - Szy; n.a02 = Szx - Sxz; n.a03 = Sxy - Syx;
    n.a11 = Sxx - Syy - Szz; n.a12 = Sxy + Syx; n.a13 = Szx + Sxz;
    n.a22 = -Sxx + Syy - Szz; n.a23 = Syz + Szy;
    n.a33 = -Sxx - Syy + Szz;
    double ss = 0.0;
    for (int k = 0; k < 9; ++k) ss += s[k] * s[k];
    bool fast = ss > 0.0 && ss < 1e280;
    double q[4] = {1.0, 0.0, 0.0, 0.0};
    if (fast) {
        // characteristic polynomial l^4 + c2 l^2 + c1 l + c0 (trace N = 0)
        const double c2 = -2.0 * ss;
        const double c1 = -8.0 * det3(Sxx, Sxy, Sxz, Syx, Syy, Syz, Szx, Szy, Szz);
        double col0[4];
        adj_column(n, 0, col0);
        const double c0 = n.a00 * col0[0] + n.a01 * col0[1] + n.a02 * col0[2] + n.a03 * col0[3];
        // sum lambda_i = 0, sum lambda_i^2 = 4 ss  ->  lambda_max <= sqrt(3 ss)
        double lam = sqrt(3.0 * ss);
        for (int it = 0; it < 64; ++it) {
            const double l2 = lam * lam;
            const double p = (l2 + c2) * l2 + c1 * lam + c0;
            const double dp = (4.0 * l2 + 2.0 * c2) * lam + c1;
            if (!(dp > 0.0)) break;
            const double nl = lam - p / dp;
            if (!(nl < lam)) break;                      // no further decrease: converged to rounding
            const bool done = lam - nl <= 2e-16 * lam;
            lam = nl;
            if (done) break;
        }
        const double scale3 = ss * sqrt(ss);             // ~ |lambda|^3
        int best = -1;
        for (int pass = 0; pass < 2 && fast; ++pass) {
            Sym4 m = n;
            m.a00 -= lam; m.a11 -= lam; m.a22 -= lam; m.a33 -= lam;
            if (pass == 0) {
                double dbest = 0.0;
                for (int j = 0; j < 4; ++j) {
                    const double d = fabs(adj_diag(m, j));
                    if (d > dbest) { dbest = d; best = j; }
                }
                // adj = prod_{i>1}(lambda_i - lambda_1) q q^T: tiny when lambda_1 is (nearly) repeated
                if (!(dbest > 1e-4 * scale3)) { fast = false; break; }
            }
            adj_column_sw(m, best, q);
            const double nq2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
            if (!(nq2 > 0.0)) { fast = false; break; }
            if (pass == 0) {
                // Rayleigh quotient of q
                const double y0 = n.a00 * q[0] + n.a01 * q[1] + n.a02 * q[2] + n.a03 * q[3];
                const double y1 = n.a01 * q[0] + n.a11 * q[1] + n.a12 * q[2] + n.a13 * q[3];
                const double y2 = n.a02 * q[0] + n.a12 * q[1] + n.a22 * q[2] + n.a23 * q[3];
                const double y3 = n.a03 * q[0] + n.a13 * q[1] + n.a23 * q[2] + n.a33 * q[3];
                lam = (q[0] * y0 + q[1] * y1 + q[2] * y2 + q[3] * y3) / nq2;
            }
        }
    }
    if (!fast) {
        double A[4][4] = {{n.a00, n.a01, n.a02, n.a03}, {n.a01, n.a11, n.a12, n.a13},
                          {n.a02, n.a12, n.a22, n.a23}, {n.a03, n.a13, n.a23, n.a33}};
        double V[4][4];
        jacobi4(A, V);
        int best = 0;
        for (int k = 1; k < 4; ++k)
            if (A[k][k] > A[best][best]) best = k;
        q[0] = V[0][best]; q[1] = V[1][best]; q[2] = V[2][best]; q[3] = V[3][best];
    }
    quaternion_to_rotation(q[0], q[1], q[2], q[3], R);
    return fast ? 0 : 1;
}
