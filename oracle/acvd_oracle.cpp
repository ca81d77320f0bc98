// ============================================================================
// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.
//
// CPU restatement of the sequential ACVD clustering path of valette/ACVD, written
// from the reference's *behaviour* (control flow, operation order, quirks), with
// every function citing the reference file:line it follows.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.  The product path (acvd_b200/csrc) never links or calls it.
//
// PARITY STATUS: *unpinned by the reference's own tests* — the reference ships no
// unit tests, golden vectors or fixtures for this path and
// cannot be compiled here (hard VTK dependency, VTK absent).  Arithmetic that
// lives in VTK (un-vendored, version unpinned, any >= 9.0) is restated from its
// published algorithm and marked [VTK, from memory]:
//   vtkTriangle::ComputeQuadric, vtkTriangle::TriangleArea,
//   vtkMath::SingularValueDecomposition3x3 (replaced by a cyclic Jacobi
//   eigen-decomposition: for the symmetric matrices on this path the truncated
//   pseudo-inverse V diag(1/w) U^T is identical up to rounding).
// The pins we create ourselves are the known-answer tests in tests/ (SURVEY §8c).
// ============================================================================
#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <mutex>
#include <queue>
#include <random>
#include <thread>
#include <vector>

namespace {

enum MetricKind { ISO = 0, QEM = 1, ANISO = 2, ANISOQ = 3 };

// number of accumulated doubles per cluster: S(3) W(1) [Q(9)] | [T(6) X(3) [Q(9)]]
inline int payload_size(int m) { return m == ISO ? 4 : (m == QEM ? 13 : (m == ANISO ? 13 : 22)); }

struct Cluster {
    double s[22];
    double centroid[3];
    double energy;
    int64_t anchor;  // QEM only: -1 none, >=0 anchored item, -2 broken (vtkQEMetricForClustering.h:127-129,259)
    char rank_def;
};

// ---------------------------------------------------------------------------
// 3x3 symmetric eigen-decomposition, cyclic Jacobi.  Stands in for
// vtkMath::SingularValueDecomposition3x3 at reference Common/vtkQuadricTools.cxx:105.
// ---------------------------------------------------------------------------
void sym_eig3(const double Ain[3][3], double w[3], double V[3][3]) {
    double a[3][3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) { a[i][j] = Ain[i][j]; V[i][j] = (i == j) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 32; sweep++) {
        double off = std::fabs(a[0][1]) + std::fabs(a[0][2]) + std::fabs(a[1][2]);
        if (off == 0.0) break;
        double diag = std::fabs(a[0][0]) + std::fabs(a[1][1]) + std::fabs(a[2][2]);
        if (off <= 1e-300 || off <= diag * 1e-17) break;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                double apq = a[p][q];
                if (apq == 0.0) continue;
                double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                int r = 3 - p - q;
                double app = a[p][p], aqq = a[q][q], arp = a[r][p], arq = a[r][q];
                a[p][p] = app - t * apq;
                a[q][q] = aqq + t * apq;
                a[p][q] = a[q][p] = 0.0;
                a[r][p] = a[p][r] = c * arp - s * arq;
                a[r][q] = a[q][r] = s * arp + c * arq;
                for (int k = 0; k < 3; k++) {
                    double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
    }
    w[0] = a[0][0]; w[1] = a[1][1]; w[2] = a[2][2];
}

// reference Common/vtkQuadricTools.cxx:83-163 (ComputeDisplacement) + :168-177 (ComputeRepresentativePoint)
int representative_point(const double* Q, double* P, int max_sv, double thr) {
    double A[3][3] = {{Q[0], Q[1], Q[2]}, {Q[1], Q[4], Q[5]}, {Q[2], Q[5], Q[7]}};
    double b[3] = {-Q[3], -Q[6], -Q[8]};
    double w[3], V[3][3];
    sym_eig3(A, w, V);
    double absw[3], maxw = -1.0;
    for (int j = 0; j < 3; j++) { absw[j] = std::fabs(w[j]); if (absw[j] > maxw) maxw = absw[j]; }
    double inv_max = 1.0 / maxw;
    double inv[3] = {0, 0, 0};
    int rank_def = 0;
    for (int i = 0; i < 3; i++) {
        double lm = -1; int im = -1;
        for (int j = 0; j < 3; j++) if (lm < absw[j]) { lm = absw[j]; im = j; }
        if ((absw[im] * inv_max > thr) && (max_sv > 0)) inv[im] = 1.0 / w[im];
        else { inv[im] = 0.0; rank_def++; }
        absw[im] = -2; max_sv--;
    }
    double r[3];
    for (int i = 0; i < 3; i++) r[i] = b[i] - (A[i][0] * P[0] + A[i][1] * P[1] + A[i][2] * P[2]);
    double d[3] = {0, 0, 0};
    for (int k = 0; k < 3; k++) {
        double proj = (V[0][k] * r[0] + V[1][k] * r[1] + V[2][k] * r[2]) * inv[k];
        d[0] += V[0][k] * proj; d[1] += V[1][k] * proj; d[2] += V[2][k] * proj;
    }
    P[0] += d[0]; P[1] += d[1]; P[2] += d[2];
    return rank_def;
}

// [VTK, from memory] vtkTriangle::ComputeQuadric: n = x1 x x2 + x2 x x3 + x3 x x1, d = -det[x1;x2;x3],
// quadric = (n,d)(n,d)^T.  Coefficient order of reference Common/vtkQuadricTools.cxx:68-78.
void triangle_quadric(const double* x1, const double* x2, const double* x3, double* Q10) {
    auto cross = [](const double* a, const double* b, double* c) {
        c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
    };
    double c12[3], c23[3], c31[3];
    cross(x1, x2, c12); cross(x2, x3, c23); cross(x3, x1, c31);
    double det = x1[0] * x2[1] * x3[2] + x2[0] * x3[1] * x1[2] + x3[0] * x1[1] * x2[2]
               - x1[0] * x3[1] * x2[2] - x2[0] * x1[1] * x3[2] - x3[0] * x2[1] * x1[2];
    double n[4] = {c12[0] + c23[0] + c31[0], c12[1] + c23[1] + c31[1], c12[2] + c23[2] + c31[2], -det};
    Q10[0] = n[0] * n[0]; Q10[1] = n[0] * n[1]; Q10[2] = n[0] * n[2]; Q10[3] = n[0] * n[3];
    Q10[4] = n[1] * n[1]; Q10[5] = n[1] * n[2]; Q10[6] = n[1] * n[3];
    Q10[7] = n[2] * n[2]; Q10[8] = n[2] * n[3]; Q10[9] = n[3] * n[3];
}

// [VTK, from memory] vtkTriangle::TriangleArea (VTK 9): 0.5 * |(p3-p2) x (p1-p2)|
double triangle_area(const double* p1, const double* p2, const double* p3) {
    double ax = p3[0] - p2[0], ay = p3[1] - p2[1], az = p3[2] - p2[2];
    double bx = p1[0] - p2[0], by = p1[1] - p2[1], bz = p1[2] - p2[2];
    double nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
    return 0.5 * std::sqrt(nx * nx + ny * ny + nz * nz);
}

// ---- [VTK, from memory] helpers of vtkCurvatureMeasure's polynomial fitting ------------------------------------
// vtkMath::InvertMatrix = LUFactorLinearSystem (Crout, implicit scaling, |pivot| <= 1e-12 -> singular) + one
// LUSolveLinearSystem per column of the identity.  A is n x n row-major and is destroyed.
int vtk_invert_matrix(double* A, double* AI, int n) {
    std::vector<int> index(n);
    std::vector<double> scale(n), col(n);
    for (int i = 0; i < n; i++) {
        double largest = 0;
        for (int j = 0; j < n; j++) largest = std::max(largest, std::fabs(A[i * n + j]));
        if (largest == 0.0) return 0;
        scale[i] = 1.0 / largest;
    }
    for (int j = 0; j < n; j++) {
        for (int i = 0; i < j; i++) {
            double sum = A[i * n + j];
            for (int k = 0; k < i; k++) sum -= A[i * n + k] * A[k * n + j];
            A[i * n + j] = sum;
        }
        double largest = 0;
        int maxI = j;
        for (int i = j; i < n; i++) {
            double sum = A[i * n + j];
            for (int k = 0; k < j; k++) sum -= A[i * n + k] * A[k * n + j];
            A[i * n + j] = sum;
            double t = scale[i] * std::fabs(sum);
            if (t >= largest) { largest = t; maxI = i; }
        }
        if (j != maxI) {
            for (int k = 0; k < n; k++) std::swap(A[maxI * n + k], A[j * n + k]);
            scale[maxI] = scale[j];
        }
        index[j] = maxI;
        if (std::fabs(A[j * n + j]) <= 1.0e-12) return 0;
        if (j != n - 1) {
            double t = 1.0 / A[j * n + j];
            for (int i = j + 1; i < n; i++) A[i * n + j] *= t;
        }
    }
    for (int c = 0; c < n; c++) {
        for (int i = 0; i < n; i++) col[i] = (i == c) ? 1.0 : 0.0;
        int ii = -1;
        for (int i = 0; i < n; i++) {            // forward substitution with the row permutation
            int idx = index[i];
            double sum = col[idx];
            col[idx] = col[i];
            if (ii >= 0) for (int j = ii; j <= i - 1; j++) sum -= A[i * n + j] * col[j];
            else if (sum != 0.0) ii = i;
            col[i] = sum;
        }
        for (int i = n - 1; i >= 0; i--) {       // back substitution
            double sum = col[i];
            for (int j = i + 1; j < n; j++) sum -= A[i * n + j] * col[j];
            col[i] = sum / A[i * n + i];
        }
        for (int i = 0; i < n; i++) AI[i * n + c] = col[i];
    }
    return 1;
}

// vtkMath::JacobiN for n = 2 (Numerical Recipes "jacobi" on the UPPER triangle, at most 20 sweeps), eigenvalues
// sorted in decreasing order (">=" comparison), each eigenvector (a column of v) negated when it has fewer than
// ceil(n/2) non-negative components.
int vtk_jacobi2(double a[2][2], double w[2], double v[2][2]) {
    const int n = 2;
    double b[2], z[2];
    for (int ip = 0; ip < n; ip++) {
        for (int iq = 0; iq < n; iq++) v[ip][iq] = 0.0;
        v[ip][ip] = 1.0;
        b[ip] = w[ip] = a[ip][ip];
        z[ip] = 0.0;
    }
    int i;
    for (i = 0; i < 20; i++) {
        double sm = std::fabs(a[0][1]);
        if (sm == 0.0) break;
        double tresh = (i < 3) ? 0.2 * sm / (n * n) : 0.0;
        {
            const int ip = 0, iq = 1;
            double g = 100.0 * std::fabs(a[ip][iq]);
            if (i > 3 && (std::fabs(w[ip]) + g) == std::fabs(w[ip]) && (std::fabs(w[iq]) + g) == std::fabs(w[iq])) a[ip][iq] = 0.0;
            else if (std::fabs(a[ip][iq]) > tresh) {
                double h = w[iq] - w[ip], t;
                if ((std::fabs(h) + g) == std::fabs(h)) t = a[ip][iq] / h;
                else {
                    double theta = 0.5 * h / a[ip][iq];
                    t = 1.0 / (std::fabs(theta) + std::sqrt(1.0 + theta * theta));
                    if (theta < 0.0) t = -t;
                }
                double c = 1.0 / std::sqrt(1 + t * t), sn = t * c, tau = sn / (1.0 + c);
                h = t * a[ip][iq];
                z[ip] -= h; z[iq] += h; w[ip] -= h; w[iq] += h;
                a[ip][iq] = 0.0;
                for (int j = 0; j < n; j++) {       // rotate the eigenvector columns ip, iq
                    double gg = v[j][ip], hh = v[j][iq];
                    v[j][ip] = gg - sn * (hh + gg * tau);
                    v[j][iq] = hh + sn * (gg - hh * tau);
                }
            }
        }
        for (int ip = 0; ip < n; ip++) { b[ip] += z[ip]; w[ip] = b[ip]; z[ip] = 0.0; }
    }
    if (i >= 20) return 0;
    if (w[1] >= w[0]) { std::swap(w[0], w[1]); std::swap(v[0][0], v[0][1]); std::swap(v[1][0], v[1][1]); }
    for (int j = 0; j < n; j++) {
        int num_pos = 0;
        for (int k = 0; k < n; k++) if (v[k][j] >= 0.0) num_pos++;
        if (num_pos < 1) for (int k = 0; k < n; k++) v[k][j] *= -1.0;
    }
    return 1;
}

struct Ctx {
    // ---- mesh in the reference's edge order (Common/vtkSurfaceBase.cxx:1166-1221, 1407-1468) ----
    int V = 0, F = 0, E = 0;
    std::vector<float> xyz;
    std::vector<int> tri;
    std::vector<int> ev1, ev2, ep1, ep2, enm; // Edges[e].Vertex1/2, Poly1/2, size of NonManifoldFaces
    std::vector<int64_t> ring_ptr;           // V+1, capacity slots
    std::vector<int> ring_len, ring;         // edge ids in insertion order (vtkSurfaceBase.cxx:1057-1068)
    // ---- metric ----
    int metric = ISO, np = 4;
    std::vector<double> item;                // V x np (float-valued entries for the anisotropic metrics)
    std::vector<double> weight;              // GetItemWeight
    int qlevel = 3;                          // QuadricsOptimizationLevel
    int constrained = 1;                     // QEM ActiveConstraints
    // ---- engine state (Common/vtkUniformClustering.h:183-333) ----
    int K = 0;
    std::vector<Cluster> clusters;
    std::vector<int> clustering, sizes, last_mod;
    std::vector<unsigned char> edges_last_loop, frozen;
    std::vector<int64_t> fixed;              // FixedClusters
    std::deque<int64_t> queue;
    unsigned char rel_loops = 1;
    int n_loops = 0;
    int connexity = 0;
    int unconstrained_init = 0;
    int max_loops = 5000000, max_conv = 1000000000;
    // ---- counters ----
    int64_t n_tests = 0, n_mods = 0;
    int n_conv = 0;
    double seconds = 0;
    std::vector<double> energy_log;          // one entry per loop if enabled
    int log_energy = 0;

    void point(int v, double* p) const { p[0] = xyz[3 * v]; p[1] = xyz[3 * v + 1]; p[2] = xyz[3 * v + 2]; }
    int other(int e, int v) const { return ev1[e] == v ? ev2[e] : ev1[e]; }

    // ---------------- mesh ----------------
    void build_edges() {
        std::vector<int> cap(V, 0);
        for (int f = 0; f < F; f++) for (int k = 0; k < 3; k++) cap[tri[3 * f + k]] += 2;
        ring_ptr.assign(V + 1, 0);
        for (int v = 0; v < V; v++) ring_ptr[v + 1] = ring_ptr[v] + cap[v];
        ring.assign(ring_ptr[V], -1);
        ring_len.assign(V, 0);
        ev1.clear(); ev2.clear(); ep1.clear(); ep2.clear(); enm.clear();
        ev1.reserve(3 * (size_t)F / 2 + 16); ev2.reserve(3 * (size_t)F / 2 + 16);
        ep1.reserve(3 * (size_t)F / 2 + 16); ep2.reserve(3 * (size_t)F / 2 + 16);
        for (int f = 0; f < F; f++) {
            const int* t = &tri[3 * f];
            if (t[0] == t[1]) continue;  // inactive face, vtkSurfaceBase.cxx:1443
            for (int k = 0; k < 3; k++) {
                int a = t[k], b = t[(k + 1) % 3];
                if (a == b) continue;    // self loop rejected, :1168-1172
                int found = -1;          // IsEdge scans ring(a) from the back, vtkSurfaceBase.h:543-557
                for (int i = ring_len[a] - 1; i >= 0; i--) {
                    int e = ring[ring_ptr[a] + i];
                    if ((ev1[e] == b) || (ev2[e] == b)) { found = e; break; }
                }
                if (found >= 0) {
                    if (ep1[found] < 0) ep1[found] = f;
                    else if (ep2[found] < 0) ep2[found] = f;
                    else enm[found]++;   // NonManifoldFaces list (:1187-1196): only IsEdgeManifold looks at it
                    continue;
                }
                int e = (int)ev1.size();
                ev1.push_back(a); ev2.push_back(b); ep1.push_back(f); ep2.push_back(-1); enm.push_back(0);
                ring[ring_ptr[a] + ring_len[a]++] = e;
                ring[ring_ptr[b] + ring_len[b]++] = e;
            }
        }
        E = (int)ev1.size();
    }

    // GetVertexNeighbourFaces order: ring edges, Poly1 then Poly2, unique (vtkSurfaceBase.cxx:964-984)
    int vertex_faces(int v, int* out) const {
        int n = 0;
        for (int i = 0; i < ring_len[v]; i++) {
            int e = ring[ring_ptr[v] + i];
            int fs[2] = {ep1[e], ep2[e]};
            for (int k = 0; k < 2; k++) {
                if (fs[k] < 0) continue;
                bool dup = false;
                for (int j = 0; j < n; j++) if (out[j] == fs[k]) { dup = true; break; }
                if (!dup && n < 256) out[n++] = fs[k];
            }
        }
        return n;
    }
    double face_area(int f) const {
        double a[3], b[3], c[3];
        point(tri[3 * f], a); point(tri[3 * f + 1], b); point(tri[3 * f + 2], c);
        return triangle_area(a, b, c);
    }
    // ---------------- vtkCurvatureMeasure, polynomial fitting on vertices ----------------
    // vtkNeighbourhoodComputation::ComputeNRingCells, CellType 1 (Common/vtkNeighbourhoodComputation.cxx:38-110):
    // faces around the vertex, then ring_size expansions; every expansion visits the edge ring of each not yet
    // visited vertex of the current front, appends the far end to the next front and the edge's faces to the list.
    void nring_faces(int v0, int ring_size, std::vector<int>& flist, std::vector<unsigned char>& face_tag,
                     std::vector<unsigned char>& vert_tag, std::vector<int>& touched_v) const {
        flist.clear();
        touched_v.clear();
        int fl[256];
        int n = vertex_faces(v0, fl);
        for (int i = 0; i < n; i++) { flist.push_back(fl[i]); face_tag[fl[i]] = 1; }
        std::vector<int> vlist{v0}, vlist2;
        for (int j = 0; j < ring_size; j++) {
            vlist2.clear();
            for (int v1 : vlist) {
                if (vert_tag[v1]) continue;
                vert_tag[v1] = 1;
                touched_v.push_back(v1);
                for (int l = 0; l < ring_len[v1]; l++) {
                    int e = ring[ring_ptr[v1] + l];
                    vlist2.push_back(other(e, v1));
                    int f1 = ep1[e], f2 = ep2[e];
                    if (f1 >= 0) {
                        if (!face_tag[f1]) { flist.push_back(f1); face_tag[f1] = 1; }
                        if (f2 >= 0 && !face_tag[f2]) { flist.push_back(f2); face_tag[f2] = 1; }
                    }
                }
            }
            vlist.swap(vlist2);
        }
        for (int f : flist) face_tag[f] = 0;
        for (int u : touched_v) vert_tag[u] = 0;
    }

    // vtkSinglePolynomialMeasure::ComputeFitting (Common/vtkCurvatureMeasure.cxx:188-508): z = q0 + q1 x + q2 y +
    // q3 x^2 + q4 xy + q5 y^2 over the face barycentres in the frame of the area-weighted mean normal; returns
    // sqrt(k1^2 + k2^2), info6 = (sqrt|k_a| d_a, sqrt|k_b| d_b) with the larger |k| first.
    double curvature_fit(const std::vector<int>& flist, double* info) const {
        for (int i = 0; i < 6; i++) info[i] = 0;
        const int n = (int)flist.size();
        double SArea = 0, Origin[3] = {0, 0, 0}, Frame[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        std::vector<double> bary(3 * (size_t)n);
        for (int j = 0; j < n; j++) {
            int f = flist[j];
            double p1[3], p2[3], p3[3];
            point(tri[3 * f], p1); point(tri[3 * f + 1], p2); point(tri[3 * f + 2], p3);
            // [VTK, from memory] vtkTriangle::ComputeNormal: (p3 - p2) x (p1 - p2), normalised when non-zero
            double ax = p3[0] - p2[0], ay = p3[1] - p2[1], az = p3[2] - p2[2];
            double bx = p1[0] - p2[0], by = p1[1] - p2[1], bz = p1[2] - p2[2];
            double N[3] = {ay * bz - az * by, az * bx - ax * bz, ax * by - ay * bx};
            double len = std::sqrt(N[0] * N[0] + N[1] * N[1] + N[2] * N[2]);
            if (len != 0.0) { N[0] /= len; N[1] /= len; N[2] /= len; }
            // vtkSurface::GetCellMassProperties (Common/vtkSurface.cxx:1380-1420)
            double Area = triangle_area(p1, p2, p3), B[3];
            for (int k = 0; k < 3; k++) {
                B[k] = Area * (p1[k] + p2[k] + p3[k]) / 3.0;
                if (Area > 0) B[k] /= Area;
                bary[3 * j + k] = B[k];
                Origin[k] += Area * B[k];
                Frame[0][k] += Area * N[k];
            }
            SArea += Area;
        }
        for (int k = 0; k < 3; k++) Origin[k] /= SArea;
        auto normalize = [](double* x) { double l = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]); if (l != 0.0) { x[0] /= l; x[1] /= l; x[2] /= l; } };
        auto cross = [](const double* a, const double* b, double* c) {
            double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
            c[0] = x; c[1] = y; c[2] = z;
        };
        normalize(Frame[0]);
        for (int k = 0; k < 3; k++) Frame[2][k] = Frame[0][k];
        Frame[1][1] = Frame[2][0]; Frame[1][2] = Frame[2][1]; Frame[1][0] = Frame[2][2];
        cross(Frame[1], Frame[2], Frame[0]);
        normalize(Frame[0]);
        cross(Frame[2], Frame[0], Frame[1]);
        if (n <= 6) return 0.0;                              // NumberOfCellsWithSmallNeighbourhood
        std::vector<double> X(6 * (size_t)n), Z(n);
        double h = 0;
        for (int j = 0; j < n; j++) {
            double d[3] = {bary[3 * j] - Origin[0], bary[3 * j + 1] - Origin[1], bary[3 * j + 2] - Origin[2]};
            double x = d[0] * Frame[0][0] + d[1] * Frame[0][1] + d[2] * Frame[0][2];
            double y = d[0] * Frame[1][0] + d[1] * Frame[1][1] + d[2] * Frame[1][2];
            double z = d[0] * Frame[2][0] + d[1] * Frame[2][1] + d[2] * Frame[2][2];
            double* r = &X[6 * (size_t)j];
            r[0] = 1.0; r[1] = x; r[2] = y; r[3] = x * x; r[4] = x * y; r[5] = y * y;
            Z[j] = z;
            h += std::sqrt(x * x + y * y);
        }
        h /= (double)n;
        for (int j = 0; j < n; j++) {
            double* r = &X[6 * (size_t)j];
            r[1] /= h; r[2] /= h; r[3] /= h * h; r[4] /= h * h; r[5] /= h * h;
        }
        // SolveLeastSquares (:510-621): normal equations, upper half accumulated then mirrored, vtkMath::InvertMatrix
        double XXt[36], XXtI[36], XYt[6], Q[6];
        for (int i = 0; i < 36; i++) XXt[i] = 0;
        for (int i = 0; i < 6; i++) XYt[i] = 0;
        for (int k = 0; k < n; k++) {
            const double* r = &X[6 * (size_t)k];
            for (int i = 0; i < 6; i++) {
                for (int j = i; j < 6; j++) XXt[i * 6 + j] += r[i] * r[j];
                XYt[i] += r[i] * Z[k];
            }
        }
        for (int i = 0; i < 6; i++) for (int j = 0; j < i; j++) XXt[i * 6 + j] = XXt[j * 6 + i];
        if (!vtk_invert_matrix(XXt, XXtI, 6)) return 0.0;   // NumberOfBadMatrices
        for (int i = 0; i < 6; i++) { Q[i] = 0; for (int k = 0; k < 6; k++) Q[i] += XXtI[i * 6 + k] * XYt[k]; }
        Q[1] /= h; Q[2] /= h; Q[3] /= h * h; Q[4] /= h * h; Q[5] /= h * h;
        double E = 1.0 + Q[1] * Q[1], Fm = Q[1] * Q[2], G = 1.0 + Q[2] * Q[2];
        double den = std::sqrt(Q[1] * Q[1] + 1.0 + Q[2] * Q[2]);
        double e = 2.0 * Q[3] / den, f = 2.0 * Q[4] / den, g = 2.0 * Q[5] / den;
        double A[4] = {E, Fm, Fm, G}, B[4];
        if (!vtk_invert_matrix(A, B, 2)) return 0.0;
        double S[2][2];
        S[0][0] = -(e * B[0] + f * B[2]);
        S[1][0] = -(e * B[1] + f * B[3]);
        S[0][1] = -(f * B[0] + g * B[2]);
        S[1][1] = -(f * B[1] + g * B[3]);
        double w[2], ev[2][2];
        if (!vtk_jacobi2(S, w, ev)) return 0.0;
        for (int i = 0; i < 2; i++)
            for (int j = 0; j < 2; j++)
                for (int k = 0; k < 3; k++) info[k + 3 * j] += std::sqrt(std::fabs(w[j])) * ev[i][j] * Frame[i][k];
        if (std::fabs(w[0]) < std::fabs(w[1]))
            for (int i = 0; i < 3; i++) std::swap(info[i], info[i + 3]);
        return std::sqrt(w[0] * w[0] + w[1] * w[1]);
    }

    // Common/vtkSurface.cxx:1342-1359
    double vertex_area(int v) const {
        int fl[256]; int n = vertex_faces(v, fl);
        double A = 0;
        for (int i = 0; i < n; i++) A += face_area(fl[i]) / 3.0;
        return A;
    }
    void vertex_quadric(int v, double* Q9) const {  // vtkQEMetricForClustering.h:151-167
        int fl[256]; int n = vertex_faces(v, fl);
        for (int i = 0; i < 9; i++) Q9[i] = 0;
        for (int i = 0; i < n; i++) {
            double a[3], b[3], c[3], q[10];
            int f = fl[i];
            point(tri[3 * f], a); point(tri[3 * f + 1], b); point(tri[3 * f + 2], c);
            triangle_quadric(a, b, c, q);
            for (int k = 0; k < 9; k++) Q9[k] += q[k];
        }
    }

    // ---------------- metric build ----------------
    // iso: vtkIsotropicMetricForClustering.h:249-291; qem: vtkQEMetricForClustering.h:323-362;
    // aniso(q): vtkQuadricAnisotropicMetricForClustering.h:366-487 / vtkAnisotropicMetricForClustering.h:300-430
    void build_metric(int m, double gradation, const double* custom, const float* pd) {
        metric = m; np = payload_size(m);
        item.assign((size_t)V * np, 0.0);
        weight.assign(V, 0.0);
        std::vector<double> area(V);
        for (int v = 0; v < V; v++) area[v] = vertex_area(v);
        bool is_float = (m == ANISO || m == ANISOQ);
        for (int v = 0; v < V; v++) {
            double w = area[v];
            bool use_custom = (m == ISO) ? (custom != nullptr) : (m == QEM ? gradation > 0 : gradation != 0);
            if (use_custom && custom) w = area[v] * std::pow(custom[v], gradation);
            weight[v] = is_float ? (double)(float)w : w;
        }
        double ratio = (m == QEM) ? 10000.0 : 100000.0;
        double avg = 0;
        for (int v = 0; v < V; v++) avg += weight[v];
        avg /= (double)V;
        double mn = avg / ratio, mx = avg * ratio;
        for (int v = 0; v < V; v++) {
            if (weight[v] > mx) weight[v] = is_float ? (double)(float)mx : mx;
            if (weight[v] < mn) weight[v] = is_float ? (double)(float)mn : mn;
        }
        for (int v = 0; v < V; v++) {
            double* it = &item[(size_t)v * np];
            double p[3]; point(v, p);
            if (!is_float) {
                for (int k = 0; k < 3; k++) it[k] = p[k] * weight[v];
                it[3] = weight[v];
                if (m == QEM) vertex_quadric(v, it + 4);
            } else {
                float val[3] = {(float)p[0], (float)p[1], (float)p[2]};
                float wf = (float)weight[v];
                double A = area[v];
                double d[6];
                for (int j = 0; j < 6; j++) d[j] = pd ? (double)pd[(size_t)v * 6 + j] : 0.0;
                float T[6];
                T[0] = (float)(A * d[0] * d[0] + A * d[3] * d[3]);
                T[1] = (float)(A * d[0] * d[1] + A * d[3] * d[4]);
                T[2] = (float)(A * d[0] * d[2] + A * d[3] * d[5]);
                T[3] = (float)(A * d[1] * d[1] + A * d[5] * d[5]);  // reference quirk: d[5], not d[4]
                T[4] = (float)(A * d[1] * d[2] + A * d[4] * d[5]);
                T[5] = (float)(A * d[2] * d[2] + A * d[5] * d[5]);
                float X[3];
                X[0] = T[0] * val[0] + T[1] * val[1] + T[2] * val[2];
                X[1] = T[1] * val[0] + T[3] * val[1] + T[4] * val[2];
                X[2] = T[2] * val[0] + T[4] * val[1] + T[5] * val[2];
                for (int k = 0; k < 3; k++) { val[k] *= wf; it[k] = (double)val[k]; }
                it[3] = (double)wf;
                for (int k = 0; k < 6; k++) it[4 + k] = (double)T[k];
                for (int k = 0; k < 3; k++) it[10 + k] = (double)X[k];
                if (m == ANISOQ) vertex_quadric(v, it + 13);
            }
        }
    }

    // ---------------- per-metric cluster ops ----------------
    void reset_cluster(Cluster& c) const {
        for (int i = 0; i < 22; i++) c.s[i] = 0;
        c.centroid[0] = c.centroid[1] = c.centroid[2] = 0; c.energy = 0;
    }
    void add_item(int i, Cluster& c) const {
        const double* it = &item[(size_t)i * np];
        for (int k = 0; k < np; k++) c.s[k] += it[k];
    }
    void sub_item(int i, Cluster& c) const {
        const double* it = &item[(size_t)i * np];
        for (int k = 0; k < np; k++) c.s[k] -= it[k];
        if (metric == QEM && c.anchor == i) c.anchor = -2;
    }
    void compute_centroid(Cluster& c) const {
        switch (metric) {
        case ISO: case ANISO:
            for (int k = 0; k < 3; k++) c.centroid[k] = c.s[k] / c.s[3];
            break;
        case QEM:  // vtkQEMetricForClustering.h:268-285
            if (c.anchor >= 0) { double p[3]; point((int)c.anchor, p); for (int k = 0; k < 3; k++) c.centroid[k] = p[k]; return; }
            for (int k = 0; k < 3; k++) c.centroid[k] = c.s[k] / c.s[3];
            if (!constrained || !qlevel) return;
            c.rank_def = (char)representative_point(c.s + 4, c.centroid, qlevel, 1e-3);
            break;
        case ANISOQ:  // vtkQuadricAnisotropicMetricForClustering.h:282-288
            for (int k = 0; k < 3; k++) c.centroid[k] = c.s[k] / c.s[3];
            representative_point(c.s + 13, c.centroid, qlevel, 1e-3);
            break;
        }
    }
    void compute_energy(Cluster& c) const {
        switch (metric) {
        case ISO:  // vtkIsotropicMetricForClustering.h:114-121
            c.energy = (-c.s[0] * c.s[0] - c.s[1] * c.s[1] - c.s[2] * c.s[2]) / c.s[3];
            break;
        case QEM:  // vtkQEMetricForClustering.h:194-208
            if (c.anchor == -2) { c.energy = 1e100; return; }
            c.energy = (c.centroid[0] * c.centroid[0] + c.centroid[1] * c.centroid[1] + c.centroid[2] * c.centroid[2]) * c.s[3]
                     - 2.0 * (c.centroid[0] * c.s[0] + c.centroid[1] * c.s[1] + c.centroid[2] * c.s[2]);
            break;
        case ANISO: case ANISOQ: {  // vtkAnisotropicMetricForClustering.h:144-169 / QuadricAniso :168-184
            double x = c.centroid[0], y = c.centroid[1], z = c.centroid[2];
            const double* T = c.s + 4; const double* X = c.s + 10;
            double e = T[0] * x * x + T[3] * y * y + T[5] * z * z + 2.0 * T[1] * x * y + 2.0 * T[2] * x * z + 2.0 * T[4] * y * z;
            e -= 2.0 * (x * X[0] + y * X[1] + z * X[2]);
            c.energy = e;
            break; }
        }
    }

    // ---------------- engine ----------------
    void set_num_clusters(int k) {  // vtkUniformClustering.h:62-69, 1349-1381
        K = k; clusters.assign(K, Cluster());
        for (auto& c : clusters) { reset_cluster(c); c.anchor = -1; c.rank_def = 0; }
        sizes.assign(K, 0); last_mod.assign(K, 0); frozen.assign(K, 0);
        clustering.assign(V, K);
        edges_last_loop.assign(E, 0);
        rel_loops = 1; n_loops = 0; queue.clear();
        n_tests = n_mods = 0; n_conv = 0;
    }

    // vtkVerticesProcessing.h:168-237
    int connexity_problem(int it, int cluster) {
        if (!connexity) return 0;
        int L[64]; int n = 0;
        for (int i = 0; i < ring_len[it]; i++) {
            int e = ring[ring_ptr[it] + i];
            int a = ev1[e], b = ev2[e];
            if (a != it && clustering[a] == cluster && n < 64) L[n++] = a;
            if (b != it && clustering[b] == cluster && n < 64) L[n++] = b;
        }
        if (n == 0) return 0;
        int total = n, visited = 1;
        int q[64]; int qh = 0, qt = 0;
        q[qt++] = L[0];
        { int v = L[0]; int m = 0; for (int j = 0; j < n; j++) if (L[j] != v) L[m++] = L[j]; n = m; }
        while (qh < qt) {
            int v = q[qh++];
            for (int i = 0; i < ring_len[v]; i++) {
                int e = ring[ring_ptr[v] + i];
                int a = ev1[e], b = ev2[e];
                for (int j = 0; j < n; j++) {
                    int c = L[j];
                    int hit = (c == a) ? a : ((c == b) ? b : -1);
                    if (hit >= 0) {
                        q[qt++] = hit; visited++;
                        int m = 0; for (int jj = 0; jj < n; jj++) if (L[jj] != hit) L[m++] = L[jj]; n = m;
                        break;
                    }
                }
            }
        }
        return (total == visited) ? 0 : 1;
    }

    void recompute_sizes() {  // :353-373
        std::fill(sizes.begin(), sizes.end(), 0);
        for (int i = 0; i < V; i++) { int c = clustering[i]; if (c >= 0 && c < K) sizes[c]++; }
    }
    void recompute_statistics() {  // :376-403
        recompute_sizes();
        for (auto& c : clusters) reset_cluster(c);
        for (int i = 0; i < V; i++) { int c = clustering[i]; if (c >= 0 && c < K) add_item(i, clusters[c]); }
        for (auto& c : clusters) { compute_centroid(c); compute_energy(c); }
    }
    void set_all_modified() { for (int i = 0; i < K; i++) last_mod[i] = n_loops; }  // :717-722
    void fill_queue() {  // :636-652
        queue.clear();
        for (int e = 0; e < E; e++) if (clustering[ev1[e]] != clustering[ev2[e]]) queue.push_back(e);
        queue.push_back(-1);
    }
    void push_ring(int v) { for (int i = 0; i < ring_len[v]; i++) queue.push_back(ring[ring_ptr[v] + i]); }

    long double global_energy() const {  // :1319-1346 (binning quirk does not change the sum order materially)
        int nb = (int)std::sqrt((double)K); if (nb < 1) nb = 1;
        std::vector<long double> bins(nb, 0.0L);
        int bin = 0, cnt = 0;
        for (int i = 0; i < K; i++) {
            bins[bin] += (long double)clusters[i].energy;
            cnt++;
            if (cnt == nb && bin < nb - 1) bin++;
        }
        long double e = 0; for (auto b : bins) e += b;
        return e;
    }

    int clean_clustering() {  // :406-549
        std::vector<char> visited(V, 0);
        std::vector<std::vector<int64_t>> lists(K);
        std::vector<char> has_list(K, 0);
        std::vector<int> csz(K, 0);
        std::vector<int64_t> vis_cluster(K, 0);
        int n_fixed = fixed.empty() ? -1 : (int)fixed.size();
        std::queue<int> q;
        int number = 0;
        for (int i = 0; i < V; i++) {
            while (!q.empty()) q.pop();
            if (visited[i] || clustering[i] == K) continue;
            int size = 0; q.push(i);
            int type = clustering[i];
            int64_t fixed_item = type < n_fixed ? fixed[type] : -1;
            while (!q.empty()) {
                int a = q.front(); q.pop();
                if (visited[a]) continue;
                size += (a == fixed_item) ? (int)1e9 : 1;
                visited[a] = 1;
                for (int k = 0; k < ring_len[a]; k++) {
                    int b = other(ring[ring_ptr[a] + k], a);
                    if (!visited[b] && clustering[b] == type) q.push(b);
                }
            }
            if (vis_cluster[type] == 0) { vis_cluster[type] = i; csz[type] = size; }  // quirk: item 0 looks "unvisited"
            else {
                if (!has_list[type]) { has_list[type] = 1; lists[type].push_back(vis_cluster[type]); lists[type].push_back(csz[type]); }
                lists[type].push_back(i); lists[type].push_back(size);
            }
        }
        std::fill(visited.begin(), visited.end(), 0);
        for (int c = 0; c < K; c++) {
            if (!has_list[c]) continue;
            number++;
            int64_t smax = 0; int imax = 0;
            int nc = (int)lists[c].size() / 2;
            for (int j = 0; j < nc; j++) { if (smax >= lists[c][2 * j + 1]) continue; smax = lists[c][2 * j + 1]; imax = j; }
            for (int j = 0; j < nc; j++) {
                if (j == imax) continue;
                while (!q.empty()) q.pop();
                q.push((int)lists[c][2 * j]);
                int type = clustering[lists[c][2 * j]];
                while (!q.empty()) {
                    int a = q.front(); q.pop();
                    if (visited[a]) continue;
                    visited[a] = 1; clustering[a] = K;
                    for (int k = 0; k < ring_len[a]; k++) {
                        int b = other(ring[ring_ptr[a] + k], a);
                        if (clustering[b] == type) q.push(b);
                    }
                }
            }
        }
        return number;
    }

    void fill_holes() {  // :552-633
        std::deque<int> q;
        auto bad = [&](int c) { return c < 0 || c >= K; };
        for (int e = 0; e < E; e++) {
            int c1 = clustering[ev1[e]], c2 = clustering[ev2[e]];
            if (bad(c1) != bad(c2)) q.push_back(e);
        }
        while (!q.empty()) {
            int e = q.front(); q.pop_front();
            int i1 = ev1[e], i2 = ev2[e];
            int c1 = clustering[i1], c2 = clustering[i2];
            if (c1 == K) { std::swap(i1, i2); c1 = c2; c2 = K; }
            if (c1 != K && c2 == K && connexity_problem(i2, c2) == 0) {
                clustering[i2] = c1;
                for (int k = 0; k < ring_len[i2]; k++) q.push_back(ring[ring_ptr[i2] + k]);
            }
        }
    }

    // ComputeInitialRandomSampling, :1178-1316 (shuffle is NOT Fisher-Yates; copied verbatim in behaviour)
    void initial_sampling() {
        std::fill(clustering.begin(), clustering.end(), K);
        int offset = 0;
        for (; offset < (int)fixed.size(); offset++) clustering[fixed[offset]] = offset;
        std::vector<int> items(V);
        int remaining_items = V, remaining_regions = K - offset;
        for (int i = 0; i < V; i++) items[i] = i;
        std::mt19937 rng; rng.seed(0);
        int n = V;
        for (int i = n - 1; i > 0; --i) std::swap(items[i], items[rng() % n]);
        double sw = 0;
        for (int i = 0; i < V; i++) sw += weight[i];
        double target = sw / (double)K;
        int first = 0;
        std::queue<int> q;
        while (remaining_items > 0 && remaining_regions > 0) {
            bool found = false; int it = 0;
            while (!found && first < V) { it = items[first]; if (clustering[it] == K) found = true; else first++; }
            if (!found) break;  // reference would spin forever here; unreachable without FixedClusters
            while (!q.empty()) q.pop();
            q.push(it); sw = 0; remaining_regions--;
            while (!q.empty()) {
                it = q.front(); q.pop();
                if (clustering[it] != K) continue;
                clustering[it] = remaining_regions + offset;
                sw += weight[it];
                remaining_items--;
                for (int k = 0; k < ring_len[it]; k++) q.push(other(ring[ring_ptr[it] + k], it));
                if (sw > target) break;
            }
        }
        if (remaining_regions == 0) return;
        std::fill(sizes.begin(), sizes.end(), 0);
        for (int i = 0; i < V; i++) { items[i] = i; int c = clustering[i]; if (c != K) sizes[c]++; }
        for (int i = n - 1; i > 0; --i) std::swap(items[i], items[rng() % n]);
        first = 0;
        while (remaining_regions) {
            int it;
            while (true) {
                if (first >= V) return;  // reference reads past the array here (UB); unreachable when K <= V
                it = items[first++];
                int c = clustering[it];
                if (c == K) break;
                if (sizes[c] == 1) continue;
                clustering[it] = remaining_regions + offset;
                sizes[c]--;
                if (remaining_regions + offset < K) sizes[remaining_regions + offset]++;  // NB reference bumps the id one past the one it assigns
                break;
            }
            remaining_regions--;
            clustering[it] = remaining_regions + offset;
        }
    }

    // ProcessOneLoop, :833-995
    int process_one_loop() {
        int mods = 0;
        Cluster c21, c22, c31, c32;
        while (true) {
            int64_t e = queue.front(); queue.pop_front();
            if (e == -1) { queue.push_back(-1); return mods; }
            if (edges_last_loop[e] == rel_loops) continue;
            edges_last_loop[e] = rel_loops;
            int i1 = ev1[e], i2 = ev2[e];
            int v1 = clustering[i1], v2 = clustering[i2];
            if (v1 == v2) continue;
            if (v1 == K) {
                Cluster& c = clusters[v2];
                add_item(i1, c); compute_centroid(c); compute_energy(c);
                sizes[v2]++; push_ring(i1); mods++; clustering[i1] = v2; last_mod[v2] = n_loops;
                continue;
            } else if (v2 == K) {
                Cluster& c = clusters[v1];
                add_item(i2, c); compute_centroid(c); compute_energy(c);
                sizes[v1]++; push_ring(i2); mods++; clustering[i2] = v1; last_mod[v1] = n_loops;
                continue;
            }
            if (((last_mod[v1] < n_loops - 1) && (last_mod[v2] < n_loops - 1)) || frozen[v1] || frozen[v2]) {
                queue.push_back(e); continue;
            }
            Cluster& k1 = clusters[v1]; Cluster& k2 = clusters[v2];
            volatile double try1, try2, try3;
            try1 = k1.energy + k2.energy;
            n_tests += 2;
            if (sizes[v1] == 1 || connexity_problem(i1, v1) == 1) try2 = 100000000.0;
            else {
                c21 = k1; sub_item(i1, c21); c22 = k2; add_item(i1, c22);
                compute_centroid(c21); compute_centroid(c22); compute_energy(c21); compute_energy(c22);
                try2 = c21.energy + c22.energy;
            }
            if (sizes[v2] == 1 || connexity_problem(i2, v2) == 1) try3 = 1000000000.0;
            else {
                c32 = k2; sub_item(i2, c32); c31 = k1; add_item(i2, c31);
                compute_centroid(c31); compute_centroid(c32); compute_energy(c31); compute_energy(c32);
                try3 = c31.energy + c32.energy;
            }
            if (try1 <= try2 && try1 <= try3) queue.push_back(e);
            else if (try2 < try1 && try2 < try3) {
                clustering[i1] = v2; sizes[v2]++; sizes[v1]--; k1 = c21; k2 = c22;
                push_ring(i1); mods++; last_mod[v1] = n_loops; last_mod[v2] = n_loops;
            } else {
                clustering[i2] = v1; sizes[v1]++; sizes[v2]--; k1 = c31; k2 = c32;
                push_ring(i2); mods++; last_mod[v1] = n_loops; last_mod[v2] = n_loops;
            }
        }
    }

    // MinimizeEnergy, :725-830.  loop_budget > 0 stops after that many loops (bounded CPU-baseline samples).
    void minimize_energy(int loop_budget) {
        auto t0 = std::chrono::steady_clock::now();
        fill_holes(); fill_queue(); recompute_statistics(); set_all_modified();
        int nconv = 0; int early = 0;
        for (int i = 0; i < K; i++) if (!frozen[i]) early += sizes[i];
        int loops_here = 0;
        while (true) {
            int mods = process_one_loop();
            n_mods += mods;
            if (log_energy) energy_log.push_back((double)global_energy());
            n_loops++; loops_here++;
            if (rel_loops == 255) { std::fill(edges_last_loop.begin(), edges_last_loop.end(), 0); rel_loops = 1; }
            else rel_loops++;
            if (loop_budget > 0 && loops_here >= loop_budget) break;
            if (mods == 0 || n_loops > max_loops || (mods <= early / 1000 && nconv <= 1)) {
                if (unconstrained_init && nconv == 0) constrained = 1;
                if (nconv >= 1) connexity = 1;
                nconv++; n_conv++;
                int disc = clean_clustering();
                fill_holes();
                recompute_sizes();
                if (disc == 0 && mods == 0) break;
                if (n_loops >= max_loops) break;
                if (nconv >= max_conv) break;
                recompute_statistics(); fill_queue(); set_all_modified();
            }
        }
        seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }

    // -------- threaded restatement (DiscreteRemeshing/vtkThreadedClustering.h) --------
    // Regions: PoolSize = 5*T+1; vertices clustered into PoolSize-1 spatial regions with
    // initial sampling + fill holes (:813-822); an edge belongs to its endpoints' common
    // region, else to the seam region PoolSize-1 (:860-878).  Per loop T threads drain whole
    // regions (one region at a time, slowest-first by previous duration :261-278), locking the
    // two cluster mutexes in id order (:323-348); the seam region is then run by the main thread.
    std::vector<int> edge_region;
    int pool = 0;
    void threaded_layout(int T) {
        pool = 5 * T + 1;
        int nreg = pool - 1;
        // spatial regions via the same sampling + fill on a scratch clustering
        std::vector<int> save = clustering; int saveK = K; std::vector<int> save_sizes = sizes;
        int save_conn = connexity; connexity = 0;
        K = nreg; sizes.assign(K, 0); clustering.assign(V, K);
        std::vector<int64_t> fx; fx.swap(fixed);
        initial_sampling(); fill_holes();
        fx.swap(fixed);
        std::vector<int> region = clustering;
        K = saveK; clustering = save; sizes = save_sizes; connexity = save_conn;
        edge_region.resize(E);
        for (int e = 0; e < E; e++) {
            int r1 = region[ev1[e]], r2 = region[ev2[e]];
            edge_region[e] = (r1 == r2 && r1 < nreg) ? r1 : pool - 1;
        }
    }
    // connexity test on thread-local scratch (same algorithm; reads clustering unlocked as the reference does)
    int connexity_problem_ts(int it, int cluster) { return connexity_problem(it, cluster); }

    void minimize_energy_threaded(int T, int loop_budget) {
        auto t0 = std::chrono::steady_clock::now();
        if ((int)edge_region.size() != E || pool != 5 * T + 1) threaded_layout(T);
        fill_holes(); recompute_statistics(); set_all_modified();
        std::vector<std::mutex> locks(K);
        // per-region double-buffered queues
        std::vector<std::vector<int>> cur(pool), nxt(pool);
        std::vector<std::mutex> qlock(pool);
        auto fill = [&]() {
            for (auto& q : cur) q.clear();
            for (int e = 0; e < E; e++) if (clustering[ev1[e]] != clustering[ev2[e]]) cur[edge_region[e]].push_back(e);
        };
        fill();
        std::vector<double> prio(pool, 0.0);
        int nconv = 0, early = 0;
        for (int i = 0; i < K; i++) if (!frozen[i]) early += sizes[i];
        int loops_here = 0;
        std::atomic<int64_t> tests{0};
        auto run_region = [&](int r) -> int {
            int mods = 0;
            Cluster c21, c22, c31, c32;
            int64_t local_tests = 0;
            auto push = [&](int e) { int rr = edge_region[e]; std::lock_guard<std::mutex> g(qlock[rr]); nxt[rr].push_back(e); };
            auto push_ring_t = [&](int v) { for (int k = 0; k < ring_len[v]; k++) push(ring[ring_ptr[v] + k]); };
            for (size_t qi = 0; qi < cur[r].size(); qi++) {
                int e = cur[r][qi];
                if (edges_last_loop[e] == rel_loops) continue;
                edges_last_loop[e] = rel_loops;
                int i1 = ev1[e], i2 = ev2[e];
                int v1 = clustering[i1], v2 = clustering[i2];
                if (v1 == v2) continue;
                int lo = std::min(v1, v2), hi = std::max(v1, v2);
                if (lo >= K) continue;
                std::unique_lock<std::mutex> g1(locks[lo], std::defer_lock), g2;
                g1.lock();
                if (hi < K) { g2 = std::unique_lock<std::mutex>(locks[hi]); }
                // re-read after locking (ids may have moved while waiting)
                v1 = clustering[i1]; v2 = clustering[i2];
                if (v1 == v2 || std::min(v1, v2) != lo || std::max(v1, v2) != hi) { push(e); continue; }
                if (v1 == K) {  // threaded adoption does not stamp last_mod
                    Cluster& c = clusters[v2]; add_item(i1, c); compute_centroid(c); compute_energy(c);
                    sizes[v2]++; clustering[i1] = v2; push_ring_t(i1); mods++; continue;
                } else if (v2 == K) {
                    Cluster& c = clusters[v1]; add_item(i2, c); compute_centroid(c); compute_energy(c);
                    sizes[v1]++; clustering[i2] = v1; push_ring_t(i2); mods++; continue;
                }
                if (((last_mod[v1] < n_loops - 1) && (last_mod[v2] < n_loops - 1)) || frozen[v1] || frozen[v2]) { push(e); continue; }
                Cluster& k1 = clusters[v1]; Cluster& k2 = clusters[v2];
                double try1 = k1.energy + k2.energy, try2, try3;
                local_tests += 2;
                if (sizes[v1] == 1 || connexity_problem_ts(i1, v1) == 1) try2 = 100000000.0;
                else {
                    c21 = k1; sub_item(i1, c21); c22 = k2; add_item(i1, c22);
                    compute_centroid(c21); compute_centroid(c22); compute_energy(c21); compute_energy(c22);
                    try2 = c21.energy + c22.energy;
                }
                if (sizes[v2] == 1 || connexity_problem_ts(i2, v2) == 1) try3 = 1000000000.0;
                else {
                    c32 = k2; sub_item(i2, c32); c31 = k1; add_item(i2, c31);
                    compute_centroid(c31); compute_centroid(c32); compute_energy(c31); compute_energy(c32);
                    try3 = c31.energy + c32.energy;
                }
                if (try1 <= try2 && try1 <= try3) push(e);
                else if (try2 < try1 && try2 < try3) {
                    clustering[i1] = v2; sizes[v2]++; sizes[v1]--; k1 = c21; k2 = c22;
                    push_ring_t(i1); mods++; last_mod[v1] = n_loops; last_mod[v2] = n_loops;
                } else {
                    clustering[i2] = v1; sizes[v1]++; sizes[v2]--; k1 = c31; k2 = c32;
                    push_ring_t(i2); mods++; last_mod[v1] = n_loops; last_mod[v2] = n_loops;
                }
            }
            tests += local_tests;
            return mods;
        };
        while (true) {
            // schedule regions slowest-first (:261-278), threads spawned and joined per loop (:671-676)
            std::vector<int> order(pool - 1);
            for (int i = 0; i < pool - 1; i++) order[i] = i;
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return prio[a] > prio[b]; });
            std::atomic<int> next{0}; std::atomic<int> mods_total{0};
            std::vector<std::thread> th;
            for (int t = 0; t < T; t++) th.emplace_back([&]() {
                while (true) {
                    int k = next.fetch_add(1);
                    if (k >= pool - 1) break;
                    int r = order[k];
                    auto s0 = std::chrono::steady_clock::now();
                    mods_total += run_region(r);
                    prio[r] = std::chrono::duration<double>(std::chrono::steady_clock::now() - s0).count();
                }
            });
            for (auto& t : th) t.join();
            int mods = mods_total.load() + run_region(pool - 1);
            for (int r = 0; r < pool; r++) { cur[r].swap(nxt[r]); nxt[r].clear(); }
            n_mods += mods;
            n_loops++; loops_here++;
            if (rel_loops == 255) { std::fill(edges_last_loop.begin(), edges_last_loop.end(), 0); rel_loops = 1; } else rel_loops++;
            if (loop_budget > 0 && loops_here >= loop_budget) break;
            if (mods == 0 || n_loops > max_loops || (mods <= early / 1000 && nconv <= 1)) {
                if (unconstrained_init && nconv == 0) constrained = 1;
                if (nconv >= 1) connexity = 1;
                nconv++; n_conv++;
                int disc = clean_clustering(); fill_holes(); recompute_sizes();
                if (disc == 0 && mods == 0) break;
                if (n_loops >= max_loops || nconv >= max_conv) break;
                recompute_statistics(); fill(); set_all_modified();
            }
        }
        n_tests += tests.load();
        seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
};

}  // namespace

int orc_dual_triangles_impl(const Ctx* c, int* out, int cap);

// ---------------------------------------------------------------------------------------------------------
// The slice of vtkSurfaceBase the -m 1 loop walks on: edge table with Poly1 / Poly2 / NonManifoldFaces and vertex
// rings in creation order (AddEdge Common/vtkSurfaceBase.cxx:1166-1221, AddPolygon :1235-1300, IsEdge
// vtkSurfaceBase.h:543-557, GetThirdPoint :384-400, Conquer :402-425, IsEdgeManifold :521-528,
// IsVertexManifold vtkSurfaceBase.cxx:259-317), restated literally.
struct MiniSurf {
    std::vector<int> v1, v2, p1, p2, nm;
    std::vector<std::vector<int>> ring;
    std::vector<std::array<int, 3>> faces;
    explicit MiniSurf(int nv) : ring(nv) {}
    int is_edge(int a, int b) const {
        const auto& r = ring[a];
        for (int i = (int)r.size() - 1; i >= 0; i--) { int e = r[i]; if (v1[e] == b || v2[e] == b) return e; }
        return -1;
    }
    int add_edge(int a, int b, int f) {
        if (a == b) return -1;
        int e = is_edge(a, b);
        if (e >= 0) {
            if (p1[e] < 0) { p1[e] = f; return e; }
            if (p2[e] >= 0) { nm[e]++; return e; }
            p2[e] = f; return e;
        }
        e = (int)v1.size();
        v1.push_back(a); v2.push_back(b); p1.push_back(f); p2.push_back(-1); nm.push_back(0);
        ring[a].push_back(e); ring[b].push_back(e);
        return e;
    }
    int add_face(int a, int b, int c) {
        int f = (int)faces.size();
        faces.push_back({a, b, c});
        add_edge(a, b, f); add_edge(b, c, f); add_edge(c, a, f);
        return f;
    }
    int third_point(int f, int a, int b) const {
        const auto& t = faces[f];
        if (a != t[0] && b != t[0]) return t[0];
        if (a != t[1] && b != t[1]) return t[1];
        return t[2];
    }
    void conquer(int f1, int a, int b, int& f2, int& v3) const {
        int e = is_edge(a, b);
        if (e < 0) { f2 = -1; v3 = -1; return; }
        f2 = p2[e];
        if (f2 == -1) { v3 = -1; return; }
        if (f2 == f1) f2 = p1[e];
        v3 = third_point(f2, a, b);
    }
    bool edge_manifold(int e) const { return !(p2[e] < 0 || nm[e] != 0); }
    bool vertex_manifold(int iv) const {
        const auto& r = ring[iv];
        int remaining = (int)r.size();
        if (remaining < 2) return false;
        for (int e : r) if (!edge_manifold(e)) return false;
        int first_edge = r[0];
        int first_vertex = v1[first_edge], a = v2[first_edge];
        if (first_vertex == iv) first_vertex = a;
        a = first_vertex;
        remaining--;
        int f1 = p1[first_edge], f2 = p2[first_edge], f3;
        int b = third_point(f1, iv, first_vertex);
        do {
            if (--remaining == 0) return true;
            conquer(f1, iv, b, f3, a);
            f1 = f3; b = a;
        } while (f1 >= 0 && b != first_vertex);
        if (f2 < 0 || b == first_vertex) return false;
        a = first_vertex;
        b = third_point(f2, iv, first_vertex);
        do {
            if (--remaining == 0) return true;
            conquer(f2, iv, b, f3, a);
            f2 = f3; b = a;
        } while (f2 >= 0 && b != first_vertex);
        return false;
    }
    void neighbours(int v, std::vector<int>& out) const {
        out.clear();
        for (int e : ring[v]) out.push_back(v1[e] == v ? v2[e] : v1[e]);
    }
};

// the input mesh as a MiniSurf (same face order as Ctx::build_edges, degenerate faces skipped)
static MiniSurf input_surface(const Ctx* c) {
    MiniSurf s(c->V);
    for (int f = 0; f < c->F; f++) {
        const int* t = &c->tri[3 * f];
        if (t[0] == t[1]) { s.faces.push_back({t[0], t[1], t[2]}); continue; }
        s.add_face(t[0], t[1], t[2]);
    }
    return s;
}

// BuildDelaunayTriangulation, vertex mode (DiscreteRemeshing/vtkDiscreteRemeshing.h:1003-1133): one output face per
// input face whose three clusters are distinct (first occurrence), and under ForceManifold the dual edges between
// adjacent clusters that share no output face (:1114-1133)
static MiniSurf output_surface(const Ctx* c, int force_manifold) {
    MiniSurf s(c->K);
    std::vector<int> tri(3 * (size_t)(4 * c->K + 64));
    int n = orc_dual_triangles_impl(c, tri.data(), (int)(tri.size() / 3));
    if (n > (int)(tri.size() / 3)) { tri.resize(3 * (size_t)n); n = orc_dual_triangles_impl(c, tri.data(), n); }
    for (int i = 0; i < n; i++) s.add_face(tri[3 * i], tri[3 * i + 1], tri[3 * i + 2]);
    if (force_manifold)
        for (int e = 0; e < c->E; e++) {
            int a = c->clustering[c->ev1[e]], b = c->clustering[c->ev2[e]];
            if (a != b && a >= 0 && a < c->K && b >= 0 && b < c->K && s.is_edge(a, b) < 0) s.add_edge(a, b, -1);
        }
    return s;
}

// DetectNonManifoldOutputVertices (DiscreteRemeshing/vtkDiscreteRemeshing.h:166-383): freezes every cluster, unfreezes
// the non-manifold output vertices (whose items are all manifold input vertices) and their output neighbours, and
// appends one new cluster per issue, seeded with the first item of the offending cluster or, for a single-item
// cluster, with the first ring neighbour whose cluster has more than one item.  Grows K; returns the issue count.
static int detect_non_manifold(Ctx* c, int force_manifold, std::vector<int>* flagged) {
    MiniSurf out = output_surface(c, force_manifold);
    MiniSurf in = input_surface(c);
    int K = c->K;
    std::vector<std::vector<int>> items(K);
    c->frozen.assign(K, 1);
    for (int i = 0; i < c->V; i++) { int cl = c->clustering[i]; if (cl >= 0 && cl < K) items[cl].push_back(i); }
    std::vector<int> issues, nb;
    for (int cl = 0; cl < K; cl++) {
        if (out.vertex_manifold(cl)) continue;
        if (items[cl].empty()) continue;
        bool problem = true;
        for (int it : items[cl]) if (!in.vertex_manifold(it)) { problem = false; break; }
        if (!problem) continue;
        issues.push_back(cl);
        c->frozen[cl] = 0;
        out.neighbours(cl, nb);
        for (int x : nb) c->frozen[x] = 0;
    }
    if (flagged) *flagged = issues;
    // unassigned items carry the NULL id = the cluster count, which is about to grow: parked at -1 meanwhile (the
    // reference leaves them at the old count, where they would silently join the first appended cluster; fixed on
    // both sides, SURVEY A.4 "fix in both and say so")
    const int null_old = K;
    if (!issues.empty()) for (int i = 0; i < c->V; i++) if (c->clustering[i] == null_old) c->clustering[i] = -1;
    for (int cl : issues) {
        const int fresh = K;
        items.emplace_back();
        bool found = false;
        if (items[cl].size() > 1) {
            int it = items[cl][0];
            c->clustering[it] = fresh;
            items[fresh].push_back(it);
            items[cl].erase(items[cl].begin());
            found = true;
        } else {
            int it = items[cl][0];
            for (int k = 0; k < c->ring_len[it] && !found; k++) {
                int u = c->other(c->ring[c->ring_ptr[it] + k], it);
                int cu = c->clustering[u];
                if (cu < 0 || cu >= null_old) continue;     // NULL, or a cluster created by this very pass (one item)
                if (items[cu].size() > 1) {
                    c->clustering[u] = fresh;
                    items[cu].erase(std::find(items[cu].begin(), items[cu].end(), u));
                    items[fresh].push_back(u);
                    found = true;
                }
            }
        }
        if (found) { K++; c->frozen.push_back(0); }
        else items.pop_back();
    }
    if (!issues.empty()) for (int i = 0; i < c->V; i++) if (c->clustering[i] < 0) c->clustering[i] = K;
    if (K != c->K) {
        int oldK = c->K;
        c->clusters.resize(K); c->sizes.resize(K, 0); c->last_mod.resize(K, c->n_loops);
        for (int cl = oldK; cl < K; cl++) { c->reset_cluster(c->clusters[cl]); c->clusters[cl].anchor = -1; c->clusters[cl].rank_def = 0; }
        c->K = K;
    }
    return (int)issues.size();
}

extern "C" {

void* orc_create(int V, int F, const float* xyz, const int* tri) {
    Ctx* c = new Ctx();
    c->V = V; c->F = F;
    c->xyz.assign(xyz, xyz + 3 * (size_t)V);
    c->tri.assign(tri, tri + 3 * (size_t)F);
    c->build_edges();
    return c;
}
void orc_destroy(void* h) { delete (Ctx*)h; }
int orc_num_edges(void* h) { return ((Ctx*)h)->E; }
void orc_get_edges(void* h, int* v1, int* v2) {
    Ctx* c = (Ctx*)h;
    std::copy(c->ev1.begin(), c->ev1.end(), v1); std::copy(c->ev2.begin(), c->ev2.end(), v2);
}
// adjacency in the reference's ring order (GetVertexNeighbours, vtkSurfaceBase.cxx:988-1000)
void orc_get_csr(void* h, int* row_ptr, int* col) {
    Ctx* c = (Ctx*)h; int64_t p = 0;
    for (int v = 0; v < c->V; v++) {
        row_ptr[v] = (int)p;
        for (int k = 0; k < c->ring_len[v]; k++) col[p++] = c->other(c->ring[c->ring_ptr[v] + k], v);
    }
    row_ptr[c->V] = (int)p;
}
void orc_vertex_areas(void* h, double* out) { Ctx* c = (Ctx*)h; for (int v = 0; v < c->V; v++) out[v] = c->vertex_area(v); }
void orc_build_metric(void* h, int metric, double gradation, const double* custom, const float* pd) {
    ((Ctx*)h)->build_metric(metric, gradation, custom, pd);
}
int orc_payload_size(void* h) { return ((Ctx*)h)->np; }
void orc_get_items(void* h, double* out) { Ctx* c = (Ctx*)h; std::copy(c->item.begin(), c->item.end(), out); }
void orc_set_num_clusters(void* h, int K) { ((Ctx*)h)->set_num_clusters(K); }
void orc_set_params(void* h, int unconstrained_init, int qlevel, int max_loops, int max_conv, int connexity, int log_energy) {
    Ctx* c = (Ctx*)h;
    c->unconstrained_init = unconstrained_init; c->qlevel = qlevel;
    if (max_loops > 0) c->max_loops = max_loops;
    if (max_conv > 0) c->max_conv = max_conv;
    c->connexity = connexity; c->log_energy = log_energy;
}
void orc_set_constrained(void* h, int on) { ((Ctx*)h)->constrained = on; }
void orc_set_fixed(void* h, const int64_t* items, int n) {
    Ctx* c = (Ctx*)h; c->fixed.assign(items, items + n);
    for (int i = 0; i < n && i < c->K; i++) c->clusters[i].anchor = items[i];  // ACVDQ.cxx:327-336
}
void orc_set_frozen(void* h, const unsigned char* f) { Ctx* c = (Ctx*)h; c->frozen.assign(f, f + c->K); }
void orc_initial_sampling(void* h) { ((Ctx*)h)->initial_sampling(); }
void orc_set_clustering(void* h, const int* cl) { Ctx* c = (Ctx*)h; c->clustering.assign(cl, cl + c->V); }
void orc_get_clustering(void* h, int* cl) { Ctx* c = (Ctx*)h; std::copy(c->clustering.begin(), c->clustering.end(), cl); }
// ProcessClustering minus Init/InitSamples (vtkUniformClustering.h:683-693)
void orc_minimize(void* h, int loop_budget) {
    Ctx* c = (Ctx*)h;
    if (c->unconstrained_init && c->n_loops == 0) c->constrained = 0;
    c->minimize_energy(loop_budget);
}
void orc_minimize_threaded(void* h, int threads, int loop_budget) {
    Ctx* c = (Ctx*)h;
    if (c->unconstrained_init && c->n_loops == 0) c->constrained = 0;
    c->minimize_energy_threaded(threads, loop_budget);
}
int orc_process_one_loop(void* h) { Ctx* c = (Ctx*)h; int m = c->process_one_loop(); c->n_loops++; if (c->rel_loops == 255) { std::fill(c->edges_last_loop.begin(), c->edges_last_loop.end(), 0); c->rel_loops = 1; } else c->rel_loops++; return m; }
void orc_prime(void* h) { Ctx* c = (Ctx*)h; c->fill_holes(); c->fill_queue(); c->recompute_statistics(); c->set_all_modified(); }
void orc_recompute_statistics(void* h) { ((Ctx*)h)->recompute_statistics(); }
int orc_clean_clustering(void* h) { return ((Ctx*)h)->clean_clustering(); }
void orc_fill_holes(void* h) { ((Ctx*)h)->fill_holes(); }
void orc_set_connexity(void* h, int on) { ((Ctx*)h)->connexity = on; }
int orc_connexity_problem(void* h, int item, int cluster) { return ((Ctx*)h)->connexity_problem(item, cluster); }
double orc_global_energy(void* h) { return (double)((Ctx*)h)->global_energy(); }
// sums: K x payload; centroid K x 3; energy K; sizes K (any pointer may be null)
void orc_get_cluster_stats(void* h, double* sums, double* centroid, double* energy, int* sizes) {
    Ctx* c = (Ctx*)h;
    for (int i = 0; i < c->K; i++) {
        if (sums) for (int k = 0; k < c->np; k++) sums[(size_t)i * c->np + k] = c->clusters[i].s[k];
        if (centroid) for (int k = 0; k < 3; k++) centroid[3 * i + k] = c->clusters[i].centroid[k];
        if (energy) energy[i] = c->clusters[i].energy;
        if (sizes) sizes[i] = c->sizes[i];
    }
}
void orc_get_report(void* h, double* out) {
    Ctx* c = (Ctx*)h;
    out[0] = c->n_loops; out[1] = c->n_conv; out[2] = (double)c->n_tests; out[3] = (double)c->n_mods; out[4] = c->seconds;
}
int orc_energy_log(void* h, double* out, int cap) {
    Ctx* c = (Ctx*)h; int n = std::min((int)c->energy_log.size(), cap);
    if (out) std::copy(c->energy_log.begin(), c->energy_log.begin() + n, out);
    return (int)c->energy_log.size();
}
int orc_representative_point(const double* Q9, double* P3, int level, double thr) { return representative_point(Q9, P3, level, thr); }
void orc_triangle_quadric(const double* x1, const double* x2, const double* x3, double* Q10) { triangle_quadric(x1, x2, x3, Q10); }
double orc_triangle_area(const double* a, const double* b, const double* c) { return triangle_area(a, b, c); }
void orc_mt19937_first(unsigned* out, int n) { std::mt19937 r; r.seed(0); for (int i = 0; i < n; i++) out[i] = (unsigned)r(); }

// Dual-mesh triangle extraction, vertex mode (vtkDiscreteRemeshing.h:1003-1100, 956-1000):
// per input face in order, the clusters of its 3 vertices (unique, < K); exactly 3 -> AddFace
// unless an output face with the same vertex set exists.  Returns number of output triangles.
int orc_dual_triangles(void* h, int* out, int cap) { return orc_dual_triangles_impl((const Ctx*)h, out, cap); }
}  // extern "C"
int orc_dual_triangles_impl(const Ctx* c, int* out, int cap) {
    std::vector<std::array<int, 3>> keys;
    keys.reserve(4 * (size_t)c->K);
    std::vector<uint64_t> seen;  // sorted-triple hash set via std::sort at the end is order-destroying; use open addressing
    size_t hcap = 1; while (hcap < 8 * (size_t)c->K + 64) hcap <<= 1;
    std::vector<int64_t> table(hcap, -1);
    int n = 0;
    for (int f = 0; f < c->F; f++) {
        int a = c->clustering[c->tri[3 * f]], b = c->clustering[c->tri[3 * f + 1]], d = c->clustering[c->tri[3 * f + 2]];
        if (a >= c->K || b >= c->K || d >= c->K || a < 0 || b < 0 || d < 0) continue;
        if (a == b || a == d || b == d) continue;
        int s[3] = {a, b, d}; std::sort(s, s + 3);
        uint64_t key = ((uint64_t)s[0] * 2654435761ULL) ^ ((uint64_t)s[1] * 40503ULL << 20) ^ ((uint64_t)s[2] * 0x9E3779B97F4A7C15ULL);
        size_t pos = key & (hcap - 1); bool dup = false;
        while (table[pos] >= 0) {
            auto& t = keys[table[pos]];
            if (t[0] == s[0] && t[1] == s[1] && t[2] == s[2]) { dup = true; break; }
            pos = (pos + 1) & (hcap - 1);
        }
        if (dup) continue;
        table[pos] = (int64_t)keys.size(); keys.push_back({s[0], s[1], s[2]});
        if (n < cap) { out[3 * n] = a; out[3 * n + 1] = b; out[3 * n + 2] = d; }
        n++;
    }
    return n;
}
extern "C" {
// per output vertex (cluster): vtkSurfaceBase::IsVertexManifold on the dual mesh; force_manifold adds the -m edges
void orc_output_vertex_manifold(void* h, int force_manifold, unsigned char* out) {
    Ctx* c = (Ctx*)h;
    MiniSurf s = output_surface(c, force_manifold);
    for (int k = 0; k < c->K; k++) out[k] = s.vertex_manifold(k) ? 1 : 0;
}
void orc_input_vertex_manifold(void* h, unsigned char* out) {
    Ctx* c = (Ctx*)h;
    MiniSurf s = input_surface(c);
    for (int v = 0; v < c->V; v++) out[v] = s.vertex_manifold(v) ? 1 : 0;
}
// one DetectNonManifoldOutputVertices step; flagged (may be null) receives up to cap offending cluster ids
int orc_detect_non_manifold(void* h, int force_manifold, int* flagged, int cap) {
    Ctx* c = (Ctx*)h;
    std::vector<int> fl;
    int n = detect_non_manifold(c, force_manifold, &fl);
    for (int i = 0; i < n && i < cap && flagged; i++) flagged[i] = fl[i];
    return n;
}
int orc_num_clusters(void* h) { return ((Ctx*)h)->K; }
void orc_get_frozen(void* h, unsigned char* out) { Ctx* c = (Ctx*)h; for (int i = 0; i < c->K; i++) out[i] = c->frozen[i]; }
// boundary flag per vertex: has a ring neighbour in another cluster (bit-exact integer stage)
void orc_boundary_flags(void* h, unsigned char* out) {
    Ctx* c = (Ctx*)h;
    for (int v = 0; v < c->V; v++) {
        unsigned char b = 0;
        for (int k = 0; k < c->ring_len[v]; k++) if (c->clustering[c->other(c->ring[c->ring_ptr[v] + k], v)] != c->clustering[v]) { b = 1; break; }
        out[v] = b;
    }
}
// cluster adjacency: sorted unique (lo,hi) pairs of clusters joined by a mesh edge. Returns count.
// vtkCurvatureMeasure (ComputationMethod 1 = polynomial fitting, ElementsType 1 = vertices, n-ring neighbourhood;
// Common/vtkCurvatureMeasure.cxx:625-718, defaults :1175-1196): indicator[V] (double), info[6 V] (float, as the
// reference stores CellsCurvatureInfo in a vtkFloatArray, :742)
void orc_curvature(void* h, int ring_size, double* indicator, float* info6) {
    Ctx* c = (Ctx*)h;
    std::vector<int> flist, touched;
    std::vector<unsigned char> face_tag(c->F, 0), vert_tag(c->V, 0);
    for (int v = 0; v < c->V; v++) {
        c->nring_faces(v, ring_size, flist, face_tag, vert_tag, touched);
        double info[6];
        indicator[v] = c->curvature_fit(flist, info);
        if (info6) for (int i = 0; i < 6; i++) info6[6 * (size_t)v + i] = (float)info[i];
    }
}

int64_t orc_cluster_adjacency(void* h, int64_t* out, int64_t cap) {
    Ctx* c = (Ctx*)h; std::vector<int64_t> p;
    for (int e = 0; e < c->E; e++) {
        int a = c->clustering[c->ev1[e]], b = c->clustering[c->ev2[e]];
        if (a == b || a >= c->K || b >= c->K) continue;
        p.push_back(((int64_t)std::min(a, b) << 32) | (int64_t)std::max(a, b));
    }
    std::sort(p.begin(), p.end()); p.erase(std::unique(p.begin(), p.end()), p.end());
    for (int64_t i = 0; i < (int64_t)p.size() && i < cap; i++) out[i] = p[i];
    return (int64_t)p.size();
}

// vtkSurface::SplitLongEdges (Common/vtkSurface.cxx:444-604) with Split2 / Split3 (:429-443): threshold = ratio x mean
// edge length of the mesh as given, then passes of "midpoint on every edge above the threshold (edge-id order), faces
// replaced by pattern" until nothing is cut.  Edge ids inside a pass are the first-seen ids of the current face list
// (upstream recycles deleted face / edge slots, SURVEY A.6: its numbering is unobservable here), faces keep their order
// with the children in the pattern's order.  Two calls: sizes first (xyz_out == null), then the arrays.
static std::vector<float> g_split_xyz;
static std::vector<int> g_split_tri, g_split_p1, g_split_p2;
int orc_split_long_edges(int V, int F, const float* xyz_in, const int* tri_in, double ratio, int* nv, int* nf,
                         float* xyz_out, int* tri_out, int* parent1, int* parent2) {
    if (xyz_out) {
        std::copy(g_split_xyz.begin(), g_split_xyz.end(), xyz_out);
        std::copy(g_split_tri.begin(), g_split_tri.end(), tri_out);
        if (parent1) std::copy(g_split_p1.begin(), g_split_p1.end(), parent1);
        if (parent2) std::copy(g_split_p2.begin(), g_split_p2.end(), parent2);
        return 0;
    }
    std::vector<float> xyz(xyz_in, xyz_in + 3 * (size_t)V);
    std::vector<int> tri(tri_in, tri_in + 3 * (size_t)F), p1(V), p2(V);
    for (int v = 0; v < V; v++) p1[v] = p2[v] = v;
    double threshold = 0;
    int passes = 0;
    for (;; passes++) {
        const int nf_cur = (int)(tri.size() / 3), nv_cur = (int)(xyz.size() / 3);
        // edge table, first-seen order (AddEdge, Common/vtkSurfaceBase.cxx:1166-1221)
        std::vector<std::vector<std::pair<int, int>>> ring(nv_cur);     // (other end, edge id)
        std::vector<int> ea, eb;
        std::vector<int> eos(3 * (size_t)nf_cur, -1);
        auto find = [&](int a, int b) { for (auto& q : ring[a]) if (q.first == b) return q.second; return -1; };
        for (int f = 0; f < nf_cur; f++) {
            const int* t = &tri[3 * (size_t)f];
            if (t[0] == t[1]) continue;
            for (int k = 0; k < 3; k++) {
                int a = t[k], b = t[(k + 1) % 3];
                if (a == b) continue;
                int e = find(a, b);
                if (e < 0) { e = (int)ea.size(); ea.push_back(a); eb.push_back(b); ring[a].push_back({b, e}); ring[b].push_back({a, e}); }
                eos[3 * (size_t)f + k] = e;
            }
        }
        const int E = (int)ea.size();
        if (E == 0) break;
        std::vector<double> len(E);
        for (int e = 0; e < E; e++) {
            double d2 = 0;
            for (int d = 0; d < 3; d++) { double t = (double)xyz[3 * (size_t)ea[e] + d] - (double)xyz[3 * (size_t)eb[e] + d]; d2 += t * t; }
            len[e] = std::sqrt(d2);
        }
        if (passes == 0) { double total = 0; for (int e = 0; e < E; e++) total += len[e]; threshold = ratio * total / (double)E; }
        std::vector<int> mid(E, -1);
        int n_cut = 0;
        for (int e = 0; e < E; e++) if (len[e] > threshold) {
            mid[e] = nv_cur + n_cut++;
            for (int d = 0; d < 3; d++) xyz.push_back((float)(0.5 * ((double)xyz[3 * (size_t)ea[e] + d] + (double)xyz[3 * (size_t)eb[e] + d])));
            p1.push_back(ea[e]); p2.push_back(eb[e]);
        }
        if (n_cut == 0) break;
        std::vector<int> out;
        out.reserve(tri.size() * 2);
        auto face = [&](int a, int b, int c) { out.push_back(a); out.push_back(b); out.push_back(c); };
        auto split2 = [&](int a, int b, int c, int ab) { face(a, ab, c); face(ab, b, c); };
        auto split3 = [&](int a, int b, int c, int ab, int ac) { face(a, ab, ac); face(ab, b, c); face(c, ac, ab); };
        for (int f = 0; f < nf_cur; f++) {
            const int v1 = tri[3 * (size_t)f], v2 = tri[3 * (size_t)f + 1], v3 = tri[3 * (size_t)f + 2];
            auto m = [&](int k) { int e = eos[3 * (size_t)f + k]; return e >= 0 ? mid[e] : -1; };
            const int v12 = m(0), v23 = m(1), v13 = m(2);
            if (v12 < 0) {
                if (v13 < 0) { if (v23 < 0) face(v1, v2, v3); else split2(v2, v3, v1, v23); }
                else { if (v23 < 0) split2(v3, v1, v2, v13); else split3(v3, v1, v2, v13, v23); }
            } else {
                if (v13 < 0) { if (v23 < 0) split2(v1, v2, v3, v12); else split3(v2, v3, v1, v23, v12); }
                else {
                    if (v23 < 0) split3(v1, v2, v3, v12, v13);
                    else { face(v1, v12, v13); face(v12, v2, v23); face(v23, v3, v13); face(v12, v23, v13); }
                }
            }
        }
        tri.swap(out);
        if (passes > 64) break;
    }
    g_split_xyz.swap(xyz); g_split_tri.swap(tri); g_split_p1.swap(p1); g_split_p2.swap(p2);
    *nv = (int)(g_split_xyz.size() / 3); *nf = (int)(g_split_tri.size() / 3);
    return passes;
}
}
