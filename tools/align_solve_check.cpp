// CPU check of csrc/align_solve.cuh (the K7 rotation solver) against the cyclic-Jacobi solver:
//   g++ -O2 -o /tmp/align_solve_check tools/align_solve_check.cpp && /tmp/align_solve_check
// Random landmark sets under random rotations + noise of every size, nearly aligned frames (what every
// iteration after the first sees), planar / collinear / single-point sets, exact 180-degree turns.
#include "../wisp_mdanalysis_implementation_b200/csrc/align_solve.cuh"
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include <algorithm>

static void jacobi_rotation(const double s[9], double R[9]) {
    const double Sxx = s[0], Sxy = s[1], Sxz = s[2], Syx = s[3], Syy = s[4], Syz = s[5], Szx = s[6], Szy = s[7], Szz = s[8];
    double A[4][4] = {
        {Sxx + Syy + Szz, Syz - Szy, Szx - Sxz, Sxy - Syx},
        {Syz - Szy, Sxx - Syy - Szz, Sxy + Syx, Szx + Sxz},
        {Szx - Sxz, Sxy + Syx, -Sxx + Syy - Szz, Syz + Szy},
        {Sxy - Syx, Szx + Sxz, Syz + Szy, -Sxx - Syy + Szz}};
    double V[4][4];
    jacobi4(A, V);
    int best = 0;
    for (int k = 1; k < 4; ++k) if (A[k][k] > A[best][best]) best = k;
    quaternion_to_rotation(V[0][best], V[1][best], V[2][best], V[3][best], R);
}

static double objective(const double s[9], const double R[9]) {   // trace(R S^T) ~ sum b.(R a), larger = better
    double t = 0;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) t += R[3 * i + j] * s[3 * j + i];
    return t;
}

int main() {
    std::mt19937_64 rng(7);
    std::normal_distribution<double> g(0.0, 1.0);
    std::uniform_real_distribution<double> u(0.0, 1.0);
    double worst = 0.0, worst_obj = 0.0;
    long fallbacks = 0, cases = 0;
    for (int rep = 0; rep < 200000; ++rep) {
        const int kind = rep % 8;
        int n = 3 + (int)(u(rng) * 60);
        if (kind == 5) n = 1 + rep % 3;
        std::vector<double> a(3 * n), b(3 * n);
        const double ext[3] = {10 + 40 * u(rng), 5 + 20 * u(rng), kind == 3 ? 0.0 : 2 + 10 * u(rng)};   // kind 3: planar
        for (int i = 0; i < n; ++i)
            for (int c = 0; c < 3; ++c) a[3 * i + c] = ext[c] * g(rng) * ((kind == 4 && c > 0) ? 0.0 : 1.0);   // kind 4: collinear
        // random rotation (quaternion)
        double q[4] = {g(rng), g(rng), g(rng), g(rng)};
        if (kind == 1) { q[0] = 1.0; q[1] *= 1e-3; q[2] *= 1e-3; q[3] *= 1e-3; }      // nearly aligned
        if (kind == 2) { q[0] = 0.0; }                                               // exact 180-degree turn
        if (kind == 6) { q[0] = 1.0; q[1] = q[2] = q[3] = 0.0; }                      // identity
        double Rt[9];
        quaternion_to_rotation(q[0], q[1], q[2], q[3], Rt);
        const double noise = (kind == 7) ? 5.0 : (kind == 6 ? 0.0 : 0.5 * u(rng));
        for (int i = 0; i < n; ++i)
            for (int c = 0; c < 3; ++c)
                b[3 * i + c] = Rt[3 * c] * a[3 * i] + Rt[3 * c + 1] * a[3 * i + 1] + Rt[3 * c + 2] * a[3 * i + 2] + noise * g(rng);
        double s[9] = {0};
        for (int i = 0; i < n; ++i)
            for (int p = 0; p < 3; ++p) for (int r = 0; r < 3; ++r) s[3 * p + r] += a[3 * i + p] * b[3 * i + r];
        double R1[9], R2[9];
        const int fb = horn_rotation(s, R1);
        jacobi_rotation(s, R2);
        fallbacks += fb;
        ++cases;
        // proper rotation?
        const double det = det3(R1[0], R1[1], R1[2], R1[3], R1[4], R1[5], R1[6], R1[7], R1[8]);
        double orth = 0;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
            double d = 0; for (int k = 0; k < 3; ++k) d += R1[3 * i + k] * R1[3 * j + k];
            orth = std::max(orth, fabs(d - (i == j)));
        }
        if (fabs(det - 1) > 1e-12 || orth > 1e-12) { printf("not a rotation: rep %d kind %d det %.3e orth %.3e\n", rep, kind, det - 1, orth); return 1; }
        // the objective must match the Jacobi optimum; where the optimum is unique (non-degenerate kinds) the matrices must too
        const double o1 = objective(s, R1), o2 = objective(s, R2);
        double sc = 0; for (int k = 0; k < 9; ++k) sc += s[k] * s[k]; sc = sqrt(sc) + 1e-300;
        worst_obj = std::max(worst_obj, (o2 - o1) / sc);
        if (kind == 0 || kind == 1 || kind == 6 || kind == 7) {
            if (n >= 4) {
                double d = 0; for (int k = 0; k < 9; ++k) d = std::max(d, fabs(R1[k] - R2[k]));
                if (d > worst) { worst = d; }
                if (d > 1e-11) { printf("rep %d kind %d n %d: |R - R_jacobi| = %.3e (fallback %d)\n", rep, kind, n, d, fb); }
            }
        }
    }
    printf("%ld cases, %ld Jacobi fallbacks, worst |R - R_jacobi| (generic kinds) %.3e, worst objective deficit %.3e\n",
           cases, fallbacks, worst, worst_obj);
    return (worst < 1e-11 && worst_obj < 1e-12) ? 0 : 1;
}
