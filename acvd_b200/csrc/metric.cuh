// Per-metric cluster algebra on the device: energy of a cluster from its accumulated sums,
// representative point (truncated pseudo-inverse of the 3x3 quadric block).
//
// Follows the behaviour of the reference Metric classes:
//   iso    E = -(S.S)/W                                   vtkIsotropicMetricForClustering.h:114-121
//   qem    C = S/W [+ A^+(b - A C)], E = |C|^2 W - 2 C.S   vtkQEMetricForClustering.h:194-208, 268-285
//   aniso  C = S/W, E = C^T T C - 2 C.X                    vtkAnisotropicMetricForClustering.h:144-169
//   anisoq C = S/W + A^+(b - A S/W), E as aniso            vtkQuadricAnisotropicMetricForClustering.h:168-184, 282-288
// and of vtkQuadricTools::ComputeDisplacement (Common/vtkQuadricTools.cxx:83-163).
#pragma once
#include "common.cuh"

namespace acvd {

enum { M_ISO = 0, M_QEM = 1, M_ANISO = 2, M_ANISOQ = 3 };

// payload layout (doubles): [0..2] S, [3] W, then
//   QEM    [4..12] Q9                 (+1 pad  -> 14)
//   ANISO  [4..9] T6, [10..12] X3     (+1 pad  -> 14)
//   ANISOQ [4..9] T6, [10..12] X3, [13..21] Q9   (22)
template <int M> struct MetricTraits;
template <> struct MetricTraits<M_ISO>    { static constexpr int NP = 4,  NPAD = 4,  QOFF = -1; };
template <> struct MetricTraits<M_QEM>    { static constexpr int NP = 13, NPAD = 14, QOFF = 4;  };
template <> struct MetricTraits<M_ANISO>  { static constexpr int NP = 13, NPAD = 14, QOFF = -1; };
template <> struct MetricTraits<M_ANISOQ> { static constexpr int NP = 22, NPAD = 22, QOFF = 13; };

__host__ __device__ inline int payload_np(int m) { return m == M_ISO ? 4 : (m == M_ANISOQ ? 22 : 13); }
__host__ __device__ inline int payload_npad(int m) { return m == M_ISO ? 4 : (m == M_ANISOQ ? 22 : 14); }

struct EvalCfg {
    int constrained;   // QEM ActiveConstraints
    int qlevel;        // QuadricsOptimizationLevel
    double thr;        // singular value threshold
};

// One Jacobi rotation annihilating a[P][Q] of the symmetric 3x3 matrix a, accumulating V.
template <int P, int Q>
__device__ __forceinline__ void jacobi_rotate(double (&a)[3][3], double (&V)[3][3]) {
    constexpr int R = 3 - P - Q;
    double apq = a[P][Q];
    if (apq == 0.0) return;
    double theta = (a[Q][Q] - a[P][P]) / (2.0 * apq);
    double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
    double c = rsqrt(t * t + 1.0), s = t * c;
    double arp = a[R][P], arq = a[R][Q];
    a[P][P] -= t * apq;
    a[Q][Q] += t * apq;
    a[P][Q] = a[Q][P] = 0.0;
    a[R][P] = a[P][R] = c * arp - s * arq;
    a[R][Q] = a[Q][R] = s * arp + c * arq;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        double vkp = V[k][P], vkq = V[k][Q];
        V[k][P] = c * vkp - s * vkq;
        V[k][Q] = s * vkp + c * vkq;
    }
}

// P += A^+ (b - A P) with the reference's singular-value selection: the i-th largest |w| is kept iff
// |w|/|w|max > thr and i < max_sv.  Returns the rank deficiency.
__device__ __forceinline__ int representative_point(const double* Q, double (&P)[3], int max_sv, double thr) {
    double a[3][3] = {{Q[0], Q[1], Q[2]}, {Q[1], Q[4], Q[5]}, {Q[2], Q[5], Q[7]}};
    double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    double r[3];
#pragma unroll
    for (int i = 0; i < 3; i++) r[i] = -Q[i == 0 ? 3 : (i == 1 ? 6 : 8)] - (a[i][0] * P[0] + a[i][1] * P[1] + a[i][2] * P[2]);
    for (int sweep = 0; sweep < 32; sweep++) {
        double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        double diag = fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]);
        if (off == 0.0 || off <= 1e-300 || off <= diag * 1e-17) break;
        jacobi_rotate<0, 1>(a, V);
        jacobi_rotate<0, 2>(a, V);
        jacobi_rotate<1, 2>(a, V);
    }
    double w[3] = {a[0][0], a[1][1], a[2][2]};
    double aw[3] = {fabs(w[0]), fabs(w[1]), fabs(w[2])};
    double inv_max = 1.0 / fmax(aw[0], fmax(aw[1], aw[2]));
    int rank_def = 0;
    double d[3] = {0, 0, 0};
#pragma unroll
    for (int j = 0; j < 3; j++) {
        int rank = 0;  // position of j in decreasing-|w| order, earlier index first on ties
#pragma unroll
        for (int k = 0; k < 3; k++) rank += (aw[k] > aw[j]) || (k < j && aw[k] == aw[j]);
        bool keep = (aw[j] * inv_max > thr) && (rank < max_sv);
        if (keep) {
            double proj = (V[0][j] * r[0] + V[1][j] * r[1] + V[2][j] * r[2]) / w[j];
            d[0] += V[0][j] * proj; d[1] += V[1][j] * proj; d[2] += V[2][j] * proj;
        } else rank_def++;
    }
    P[0] += d[0]; P[1] += d[1]; P[2] += d[2];
    return rank_def;
}

// Energy (and optionally the representative point) of a cluster with sums s.
// anchor_pt: non-null => the cluster is anchored to that point (QEM fixed clusters).
template <int M>
__device__ __forceinline__ double cluster_energy(const double* s, const EvalCfg& cfg, double* centroid_out,
                                                 const double* anchor_pt = nullptr, int* rank_def_out = nullptr) {
    double W = s[3];
    double C[3] = {s[0] / W, s[1] / W, s[2] / W};
    double E;
    if (M == M_ISO) {
        E = (-s[0] * s[0] - s[1] * s[1] - s[2] * s[2]) / W;
    } else if (M == M_QEM) {
        if (anchor_pt) { C[0] = anchor_pt[0]; C[1] = anchor_pt[1]; C[2] = anchor_pt[2]; }
        else if (cfg.constrained && cfg.qlevel) {
            int rd = representative_point(s + 4, C, cfg.qlevel, cfg.thr);
            if (rank_def_out) *rank_def_out = rd;
        }
        E = (C[0] * C[0] + C[1] * C[1] + C[2] * C[2]) * W - 2.0 * (C[0] * s[0] + C[1] * s[1] + C[2] * s[2]);
    } else {
        if (M == M_ANISOQ) representative_point(s + 13, C, cfg.qlevel, cfg.thr);
        const double* T = s + 4;
        const double* X = s + 10;
        double x = C[0], y = C[1], z = C[2];
        E = T[0] * x * x + T[3] * y * y + T[5] * z * z + 2.0 * T[1] * x * y + 2.0 * T[2] * x * z + 2.0 * T[4] * y * z;
        E -= 2.0 * (x * X[0] + y * X[1] + z * X[2]);
    }
    if (centroid_out) { centroid_out[0] = C[0]; centroid_out[1] = C[1]; centroid_out[2] = C[2]; }
    return E;
}

// 16-byte vectorised payload row load/store (row starts are 16-byte aligned, N even)
template <int N>
__device__ __forceinline__ void load_row(const double* row, double* out) {
    const double2* p = reinterpret_cast<const double2*>(row);
#pragma unroll
    for (int i = 0; i < N / 2; i++) { double2 t = p[i]; out[2 * i] = t.x; out[2 * i + 1] = t.y; }
}
template <int N>
__device__ __forceinline__ void load_row_ro(const double* __restrict__ row, double* out) {
    const double2* p = reinterpret_cast<const double2*>(row);
#pragma unroll
    for (int i = 0; i < N / 2; i++) { double2 t = __ldg(p + i); out[2 * i] = t.x; out[2 * i + 1] = t.y; }
}
template <int N>
__device__ __forceinline__ void store_row(double* row, const double* in) {
    double2* p = reinterpret_cast<double2*>(row);
#pragma unroll
    for (int i = 0; i < N / 2; i++) p[i] = make_double2(in[2 * i], in[2 * i + 1]);
}

}  // namespace acvd
