// Per-vertex curvature by polynomial fitting: the step in front of the hot path for every gradation > 0 run.
//
// Replaces vtkCurvatureMeasure with ComputationMethod 1, ElementsType 1 (vertices), n-ring neighbourhood
// (reference Common/vtkCurvatureMeasure.cxx:188-508 ComputeFitting, :510-621 SolveLeastSquares, :625-718 driver,
// defaults :1175-1196; neighbourhood Common/vtkNeighbourhoodComputation.cxx:38-110).  Per vertex:
//   faces  = all faces with a vertex within graph distance ring_size - 1 (what ComputeNRingCells collects);
//   frame  = area-weighted mean normal n, t1 = normalize(perm(n) x n) with perm(n) = (n_z, n_x, n_y), t2 = n x t1;
//   fit    z = q0 + q1 x + q2 y + q3 x^2 + q4 xy + q5 y^2 over the face barycentres (normal equations, columns
//          conditioned by h = mean sqrt(x^2 + y^2), 6x6 inverse by LU with implicit-scaling pivoting);
//   shape operator from the fundamental forms, eigen-decomposition by vtkMath::JacobiN on its UPPER triangle;
//   out    indicator = sqrt(k1^2 + k2^2), info = (sqrt|k_a| d_a, sqrt|k_b| d_b), larger |k| first, float32.
// vtkMath::InvertMatrix, vtkMath::JacobiN and vtkTriangle::ComputeNormal are restated [VTK, from memory].
// The reference sums the faces in BFS order; here each face is taken at the first neighbourhood vertex that
// touches it, so sums differ by rounding only.  One thread per vertex; the neighbourhood list lives in local
// memory (kCurvLocalCap vertices), larger neighbourhoods are redone with a global scratch list.
#pragma once
#include "mesh.cuh"

namespace acvd {

constexpr int kCurvLocalCap = 96;
constexpr int kCurvGlobalCap = 8192;

// [VTK, from memory] vtkMath::InvertMatrix: LUFactorLinearSystem (Crout, implicit scaling, |pivot| <= 1e-12 ->
// singular) and one LUSolveLinearSystem per column of the identity.  A is destroyed.
template <int N>
__device__ int vtk_invert_matrix(double (&A)[N][N], double (&AI)[N][N]) {
    int index[N];
    double scale[N], colv[N];
    for (int i = 0; i < N; i++) {
        double largest = 0;
        for (int j = 0; j < N; j++) largest = fmax(largest, fabs(A[i][j]));
        if (largest == 0.0) return 0;
        scale[i] = 1.0 / largest;
    }
    for (int j = 0; j < N; j++) {
        for (int i = 0; i < j; i++) {
            double sum = A[i][j];
            for (int k = 0; k < i; k++) sum -= A[i][k] * A[k][j];
            A[i][j] = sum;
        }
        double largest = 0;
        int maxI = j;
        for (int i = j; i < N; i++) {
            double sum = A[i][j];
            for (int k = 0; k < j; k++) sum -= A[i][k] * A[k][j];
            A[i][j] = sum;
            const double t = scale[i] * fabs(sum);
            if (t >= largest) { largest = t; maxI = i; }
        }
        if (j != maxI) {
            for (int k = 0; k < N; k++) { const double t = A[maxI][k]; A[maxI][k] = A[j][k]; A[j][k] = t; }
            scale[maxI] = scale[j];
        }
        index[j] = maxI;
        if (fabs(A[j][j]) <= 1.0e-12) return 0;
        if (j != N - 1) {
            const double t = 1.0 / A[j][j];
            for (int i = j + 1; i < N; i++) A[i][j] *= t;
        }
    }
    for (int c = 0; c < N; c++) {
        for (int i = 0; i < N; i++) colv[i] = (i == c) ? 1.0 : 0.0;
        int ii = -1;
        for (int i = 0; i < N; i++) {
            const int idx = index[i];
            double sum = colv[idx];
            colv[idx] = colv[i];
            if (ii >= 0) { for (int j = ii; j <= i - 1; j++) sum -= A[i][j] * colv[j]; }
            else if (sum != 0.0) ii = i;
            colv[i] = sum;
        }
        for (int i = N - 1; i >= 0; i--) {
            double sum = colv[i];
            for (int j = i + 1; j < N; j++) sum -= A[i][j] * colv[j];
            colv[i] = sum / A[i][i];
        }
        for (int i = 0; i < N; i++) AI[i][c] = colv[i];
    }
    return 1;
}

// [VTK, from memory] vtkMath::JacobiN for n = 2: Jacobi rotations on the upper triangle (at most 20 sweeps),
// eigenvalues in decreasing order, an eigenvector (column of v) is negated when both components are negative.
__device__ inline int vtk_jacobi2(double (&a)[2][2], double (&w)[2], double (&v)[2][2]) {
    double b[2], z[2];
    v[0][0] = 1.0; v[0][1] = 0.0; v[1][0] = 0.0; v[1][1] = 1.0;
    b[0] = w[0] = a[0][0]; b[1] = w[1] = a[1][1];
    z[0] = z[1] = 0.0;
    int i;
    for (i = 0; i < 20; i++) {
        const double sm = fabs(a[0][1]);
        if (sm == 0.0) break;
        const double tresh = (i < 3) ? 0.2 * sm / 4.0 : 0.0;
        const double g = 100.0 * fabs(a[0][1]);
        if (i > 3 && (fabs(w[0]) + g) == fabs(w[0]) && (fabs(w[1]) + g) == fabs(w[1])) a[0][1] = 0.0;
        else if (fabs(a[0][1]) > tresh) {
            double h = w[1] - w[0], t;
            if ((fabs(h) + g) == fabs(h)) t = a[0][1] / h;
            else {
                const double theta = 0.5 * h / a[0][1];
                t = 1.0 / (fabs(theta) + sqrt(1.0 + theta * theta));
                if (theta < 0.0) t = -t;
            }
            const double c = 1.0 / sqrt(1 + t * t), sn = t * c, tau = sn / (1.0 + c);
            h = t * a[0][1];
            z[0] -= h; z[1] += h; w[0] -= h; w[1] += h;
            a[0][1] = 0.0;
            for (int j = 0; j < 2; j++) {
                const double gg = v[j][0], hh = v[j][1];
                v[j][0] = gg - sn * (hh + gg * tau);
                v[j][1] = hh + sn * (gg - hh * tau);
            }
        }
        b[0] += z[0]; w[0] = b[0]; z[0] = 0.0;
        b[1] += z[1]; w[1] = b[1]; z[1] = 0.0;
    }
    if (i >= 20) return 0;
    if (w[1] >= w[0]) {
        double t = w[0]; w[0] = w[1]; w[1] = t;
        t = v[0][0]; v[0][0] = v[0][1]; v[0][1] = t;
        t = v[1][0]; v[1][0] = v[1][1]; v[1][1] = t;
    }
    for (int j = 0; j < 2; j++)
        if (v[0][j] < 0.0 && v[1][j] < 0.0) { v[0][j] = -v[0][j]; v[1][j] = -v[1][j]; }
    return 1;
}

struct CurvMesh {
    int V;
    const int* __restrict__ row_ptr;
    const int* __restrict__ col;
    const int* __restrict__ vf_ptr;
    const unsigned long long* __restrict__ vf_keys;
    const float* __restrict__ xyz;
    const int* __restrict__ tri;
};

// vertices within graph distance ring_size - 1 of v0, breadth first; returns their number or -1 when `cap` is too small
__device__ inline int curv_collect_vertices(const CurvMesh& M, int v0, int ring_size, int* verts, int cap) {
    int n = 1, level_begin = 0;
    verts[0] = v0;
    for (int level = 0; level + 1 < ring_size; level++) {
        const int level_end = n;
        for (int i = level_begin; i < level_end; i++) {
            const int u = verts[i];
            for (int e = M.row_ptr[u]; e < M.row_ptr[u + 1]; e++) {
                const int w = M.col[e];
                bool seen = false;
                for (int k = 0; k < n; k++) seen |= (verts[k] == w);
                if (seen) continue;
                if (n == cap) return -1;
                verts[n++] = w;
            }
        }
        level_begin = level_end;
    }
    return n;
}

// calls fn(p1, p2, p3) once for every face that has a vertex in verts[0..n): a face is taken at the first list
// vertex that touches it
template <typename Fn>
__device__ inline void curv_for_each_face(const CurvMesh& M, const int* verts, int n, Fn&& fn) {
    for (int iu = 0; iu < n; iu++) {
        const int u = verts[iu];
        for (int i = M.vf_ptr[u]; i < M.vf_ptr[u + 1]; i++) {
            const int f = (int)(M.vf_keys[i] & 0xffffffffull);
            const int a = M.tri[3 * (int64_t)f], b = M.tri[3 * (int64_t)f + 1], c = M.tri[3 * (int64_t)f + 2];
            bool earlier = false;
            for (int k = 0; k < iu; k++) earlier |= (verts[k] == a) | (verts[k] == b) | (verts[k] == c);
            if (earlier) continue;
            double p1[3], p2[3], p3[3];
#pragma unroll
            for (int k = 0; k < 3; k++) { p1[k] = M.xyz[3 * (int64_t)a + k]; p2[k] = M.xyz[3 * (int64_t)b + k]; p3[k] = M.xyz[3 * (int64_t)c + k]; }
            fn(p1, p2, p3);
        }
    }
}

__device__ inline void curv_normalize(double* x) {
    const double l = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    if (l != 0.0) { x[0] /= l; x[1] /= l; x[2] /= l; }
}
__device__ inline void curv_cross(const double* a, const double* b, double* c) {
    const double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    c[0] = x; c[1] = y; c[2] = z;
}

// vtkSinglePolynomialMeasure::ComputeFitting over the neighbourhood verts[0..n); info = 6 doubles
__device__ inline double curv_fit(const CurvMesh& M, const int* verts, int n, double* info) {
    for (int i = 0; i < 6; i++) info[i] = 0.0;
    double SArea = 0, Origin[3] = {0, 0, 0}, Frame[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    int n_faces = 0;
    curv_for_each_face(M, verts, n, [&](const double* p1, const double* p2, const double* p3) {
        // [VTK, from memory] vtkTriangle::ComputeNormal: (p3 - p2) x (p1 - p2), normalised when non-zero
        const double ax = p3[0] - p2[0], ay = p3[1] - p2[1], az = p3[2] - p2[2];
        const double bx = p1[0] - p2[0], by = p1[1] - p2[1], bz = p1[2] - p2[2];
        double N[3] = {ay * bz - az * by, az * bx - ax * bz, ax * by - ay * bx};
        const double len = sqrt(N[0] * N[0] + N[1] * N[1] + N[2] * N[2]);
        const double Area = 0.5 * len;                       // vtkTriangle::TriangleArea of the same triangle
        if (len != 0.0) { N[0] /= len; N[1] /= len; N[2] /= len; }
#pragma unroll
        for (int k = 0; k < 3; k++) {
            double B = Area * (p1[k] + p2[k] + p3[k]) / 3.0;  // vtkSurface::GetCellMassProperties (Common/vtkSurface.cxx:1380-1420)
            if (Area > 0) B /= Area;
            Origin[k] += Area * B;
            Frame[0][k] += Area * N[k];
        }
        SArea += Area;
        n_faces++;
    });
    for (int k = 0; k < 3; k++) Origin[k] /= SArea;
    curv_normalize(Frame[0]);
    for (int k = 0; k < 3; k++) Frame[2][k] = Frame[0][k];
    Frame[1][1] = Frame[2][0]; Frame[1][2] = Frame[2][1]; Frame[1][0] = Frame[2][2];
    curv_cross(Frame[1], Frame[2], Frame[0]);
    curv_normalize(Frame[0]);
    curv_cross(Frame[2], Frame[0], Frame[1]);
    if (n_faces <= 6) return 0.0;
    // normal equations from the unconditioned monomials r = (1, x, y, x^2, xy, y^2); the conditioning by h
    // (h, h, h^2, h^2, h^2 on columns 1..5) is applied to the sums, which is the same algebra
    double XX[6][6], XY[6], h = 0;
    for (int i = 0; i < 6; i++) { XY[i] = 0; for (int j = 0; j < 6; j++) XX[i][j] = 0; }
    curv_for_each_face(M, verts, n, [&](const double* p1, const double* p2, const double* p3) {
        const double ax = p3[0] - p2[0], ay = p3[1] - p2[1], az = p3[2] - p2[2];
        const double bx = p1[0] - p2[0], by = p1[1] - p2[1], bz = p1[2] - p2[2];
        const double nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
        const double Area = 0.5 * sqrt(nx * nx + ny * ny + nz * nz);
        double d[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            double B = Area * (p1[k] + p2[k] + p3[k]) / 3.0;
            if (Area > 0) B /= Area;
            d[k] = B - Origin[k];
        }
        const double x = d[0] * Frame[0][0] + d[1] * Frame[0][1] + d[2] * Frame[0][2];
        const double y = d[0] * Frame[1][0] + d[1] * Frame[1][1] + d[2] * Frame[1][2];
        const double z = d[0] * Frame[2][0] + d[1] * Frame[2][1] + d[2] * Frame[2][2];
        const double r[6] = {1.0, x, y, x * x, x * y, y * y};
#pragma unroll
        for (int i = 0; i < 6; i++) {
#pragma unroll
            for (int j = i; j < 6; j++) XX[i][j] += r[i] * r[j];
            XY[i] += r[i] * z;
        }
        h += sqrt(x * x + y * y);
    });
    h /= (double)n_faces;
    const double sc[6] = {1.0, h, h, h * h, h * h, h * h};
    for (int i = 0; i < 6; i++) {
        for (int j = i; j < 6; j++) { XX[i][j] /= sc[i] * sc[j]; XX[j][i] = XX[i][j]; }
        XY[i] /= sc[i];
    }
    double XXI[6][6], Q[6];
    if (!vtk_invert_matrix<6>(XX, XXI)) return 0.0;
    for (int i = 0; i < 6; i++) { Q[i] = 0; for (int k = 0; k < 6; k++) Q[i] += XXI[i][k] * XY[k]; }
    Q[1] /= h; Q[2] /= h; Q[3] /= h * h; Q[4] /= h * h; Q[5] /= h * h;
    const double E = 1.0 + Q[1] * Q[1], Fm = Q[1] * Q[2], G = 1.0 + Q[2] * Q[2];
    const double den = sqrt(Q[1] * Q[1] + 1.0 + Q[2] * Q[2]);
    const double e = 2.0 * Q[3] / den, f = 2.0 * Q[4] / den, g = 2.0 * Q[5] / den;
    double A[2][2] = {{E, Fm}, {Fm, G}}, B[2][2];
    if (!vtk_invert_matrix<2>(A, B)) return 0.0;
    double S[2][2];
    S[0][0] = -(e * B[0][0] + f * B[1][0]);
    S[1][0] = -(e * B[0][1] + f * B[1][1]);
    S[0][1] = -(f * B[0][0] + g * B[1][0]);
    S[1][1] = -(f * B[0][1] + g * B[1][1]);
    double w[2], ev[2][2];
    if (!vtk_jacobi2(S, w, ev)) return 0.0;
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++)
            for (int k = 0; k < 3; k++) info[k + 3 * j] += sqrt(fabs(w[j])) * ev[i][j] * Frame[i][k];
    if (fabs(w[0]) < fabs(w[1]))
        for (int i = 0; i < 3; i++) { const double t = info[i]; info[i] = info[i + 3]; info[i + 3] = t; }
    return sqrt(w[0] * w[0] + w[1] * w[1]);
}

// main pass: neighbourhood list in local memory; vertices whose neighbourhood does not fit are appended to `big`
__global__ void __launch_bounds__(128) k_curvature(CurvMesh M, int ring_size, double* indicator, float* info6, int* big,
                                                   unsigned long long* n_big) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < M.V; v += gridDim.x * blockDim.x) {
        int verts[kCurvLocalCap];
        const int n = curv_collect_vertices(M, v, ring_size, verts, kCurvLocalCap);
        if (n < 0) { big[(int)atomicAdd(n_big, 1ull)] = v; continue; }
        double info[6];
        indicator[v] = curv_fit(M, verts, n, info);
        if (info6)
            for (int i = 0; i < 6; i++) info6[6 * (int64_t)v + i] = (float)info[i];
    }
}

// second pass for the (rare) large neighbourhoods: list in a global scratch row per vertex
__global__ void __launch_bounds__(128) k_curvature_big(CurvMesh M, int ring_size, double* indicator, float* info6, const int* __restrict__ big,
                                                       int n_big, int* scratch, int* failed) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_big; i += gridDim.x * blockDim.x) {
        const int v = big[i];
        int* verts = scratch + (int64_t)i * kCurvGlobalCap;
        const int n = curv_collect_vertices(M, v, ring_size, verts, kCurvGlobalCap);
        if (n < 0) { atomicExch(failed, 1); continue; }
        double info[6];
        indicator[v] = curv_fit(M, verts, n, info);
        if (info6)
            for (int k = 0; k < 6; k++) info6[6 * (int64_t)v + k] = (float)info[k];
    }
}

}  // namespace acvd
