// Device-side mesh preparation: CSR vertex adjacency from the triangle list, vertex->face incidence,
// and the per-vertex item payloads of the four metrics.
//
// Replaces, for this path, the edge table / vertex rings of vtkSurfaceBase
// (reference Common/vtkSurfaceBase.cxx:1166-1221, 1407-1468; accessors used by the engine:
// DiscreteRemeshing/vtkVerticesProcessing.h:137-157) with int32 CSR, and Metric::BuildMetric
// (vtkIsotropicMetricForClustering.h:214-291, vtkQEMetricForClustering.h:151-167, 297-362,
// vtkQuadricAnisotropicMetricForClustering.h:366-487, vtkAnisotropicMetricForClustering.h:300-430).
#pragma once
#include "metric.cuh"

namespace acvd {

// ---------------------------------------------------------------------------------------------------
// CSR adjacency and vertex -> face incidence by counting instead of a global sort of the half-edges:
// corners per vertex (atomics) -> exclusive scan -> every corner drops its face id and its two ring neighbours into
// the vertex' segment (atomic cursor) -> one thread per vertex sorts its few entries and removes duplicate
// neighbours.  The per-row sort makes the result independent of the order the atomics were served in
// (rows ascending, as the sorted-key build produced them).
constexpr int kNoNeighbour = 0x7fffffff;

__global__ void k_count_corners(int F, const int* __restrict__ tri, int* cnt) {
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < F; f += gridDim.x * blockDim.x) {
        const int v[3] = {tri[3 * (int64_t)f], tri[3 * (int64_t)f + 1], tri[3 * (int64_t)f + 2]};
        if (v[0] == v[1]) continue;                      // inactive face (vtkSurfaceBase.cxx:1443)
#pragma unroll
        for (int k = 0; k < 3; k++) atomicAdd(cnt + v[k], 1);
    }
}

// every corner drops ONE 16-byte record (face, next vertex, previous vertex) into its vertex' segment: one store per corner
__global__ void k_scatter_corners(int F, const int* __restrict__ tri, const int* __restrict__ off, int* cursor, int4* rec) {
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < F; f += gridDim.x * blockDim.x) {
        const int v[3] = {tri[3 * (int64_t)f], tri[3 * (int64_t)f + 1], tri[3 * (int64_t)f + 2]};
        if (v[0] == v[1]) continue;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int u = v[k], nxt = v[(k + 1) % 3], prv = v[(k + 2) % 3];
            const int64_t pos = (int64_t)off[u] + atomicAdd(cursor + u, 1);
            rec[pos] = make_int4(f, nxt != u ? nxt : kNoNeighbour, prv != u ? prv : kNoNeighbour, 0);   // self loops are rejected (:1168-1172)
        }
    }
}

// Batcher's odd-even merge sort as a compile-time network (N a power of two): every index is static, so the keys stay in registers
template <typename T, int N>
__device__ __forceinline__ void sort_network(T (&a)[N]) {
#pragma unroll
    for (int p = 1; p < N; p *= 2)
#pragma unroll
        for (int k = p; k >= 1; k /= 2)
#pragma unroll
            for (int j = k % p; j <= N - 1 - k; j += 2 * k)
#pragma unroll
                for (int i = 0; i < k; i++)
                    if (i + j + k < N && (i + j) / (2 * p) == (i + j + k) / (2 * p)) {
                        const T lo = a[i + j] < a[i + j + k] ? a[i + j] : a[i + j + k];
                        const T hi = a[i + j] < a[i + j + k] ? a[i + j + k] : a[i + j];
                        a[i + j] = lo; a[i + j + k] = hi;
                    }
}

// one thread per vertex: its records -> faces ascending (vf_keys, kept) and distinct neighbours ascending (he, for k_compact_rows)
__global__ void k_sort_rows(int V, const int* __restrict__ off, const int4* __restrict__ rec, unsigned long long* vf_keys, int* he, int* deg) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
        const int64_t b = off[v];
        const int m = off[v + 1] - off[v];
        unsigned long long* fk = vf_keys + b;
        int* h = he + 2 * b;
        const unsigned long long vhi = (unsigned long long)(unsigned)v << 32;
        if (m <= 8) {
            // the common row (up to 8 incident faces, 16 half-edge ends): sorted in registers by a network, one pass over memory
            unsigned long long f[8];
            int hh[16];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                int4 r = make_int4(0, kNoNeighbour, kNoNeighbour, 0);
                if (i < m) r = __ldg(rec + b + i);
                f[i] = i < m ? (vhi | (unsigned)r.x) : ~0ull;
                hh[2 * i] = r.y; hh[2 * i + 1] = r.z;
            }
            sort_network<unsigned long long, 8>(f);
            sort_network<int, 16>(hh);
#pragma unroll
            for (int i = 0; i < 8; i++) if (i < m) fk[i] = f[i];
            int d = 0, last = -1;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int x = hh[i];
                if (x != kNoNeighbour && x != last) { h[d++] = x; last = x; }      // ascending: duplicates are adjacent
            }
            deg[v] = d;
            continue;
        }
        for (int i = 0; i < m; i++) { const int4 r = rec[b + i]; fk[i] = vhi | (unsigned)r.x; h[2 * i] = r.y; h[2 * i + 1] = r.z; }
        for (int i = 1; i < m; i++) {                    // faces ascending
            const unsigned long long x = fk[i];
            int j = i - 1;
            while (j >= 0 && fk[j] > x) { fk[j + 1] = fk[j]; j--; }
            fk[j + 1] = x;
        }
        for (int i = 1; i < 2 * m; i++) {                // neighbours ascending, kNoNeighbour last
            const int x = h[i];
            int j = i - 1;
            while (j >= 0 && h[j] > x) { h[j + 1] = h[j]; j--; }
            h[j + 1] = x;
        }
        int d = 0;
        for (int i = 0; i < 2 * m; i++) {
            const int x = h[i];
            if (x == kNoNeighbour) break;
            if (d == 0 || h[d - 1] != x) h[d++] = x;
        }
        deg[v] = d;
    }
}

__global__ void k_compact_rows(int V, const int* __restrict__ off, const int* __restrict__ row_ptr, const int* __restrict__ he, int* col) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
        const int* h = he + 2 * (int64_t)off[v];
        const int b = row_ptr[v], d = row_ptr[v + 1] - b;
        for (int i = 0; i < d; i++) col[b + i] = h[i];
    }
}

struct Tri3 { double a[3], b[3], c[3]; };

__device__ __forceinline__ void load_face(const float* __restrict__ xyz, const int* __restrict__ tri, int f, Tri3& t) {
    int i0 = tri[3 * f], i1 = tri[3 * f + 1], i2 = tri[3 * f + 2];
#pragma unroll
    for (int k = 0; k < 3; k++) { t.a[k] = xyz[3 * (int64_t)i0 + k]; t.b[k] = xyz[3 * (int64_t)i1 + k]; t.c[k] = xyz[3 * (int64_t)i2 + k]; }
}

// [VTK, from memory] vtkTriangle::TriangleArea: 0.5 |(p3 - p2) x (p1 - p2)|
__device__ __forceinline__ double tri_area(const Tri3& t) {
    double ax = t.c[0] - t.b[0], ay = t.c[1] - t.b[1], az = t.c[2] - t.b[2];
    double bx = t.a[0] - t.b[0], by = t.a[1] - t.b[1], bz = t.a[2] - t.b[2];
    double nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
    return 0.5 * sqrt(nx * nx + ny * ny + nz * nz);
}

// [VTK, from memory] vtkTriangle::ComputeQuadric: n = x1 x x2 + x2 x x3 + x3 x x1, d = -det[x1;x2;x3];
// coefficient order of vtkQuadricTools::AddTriangleQuadric (Common/vtkQuadricTools.cxx:68-78).
__device__ __forceinline__ void tri_quadric_add(const Tri3& t, double* Q9) {
    const double *x1 = t.a, *x2 = t.b, *x3 = t.c;
    double n0 = (x1[1] * x2[2] - x1[2] * x2[1]) + (x2[1] * x3[2] - x2[2] * x3[1]) + (x3[1] * x1[2] - x3[2] * x1[1]);
    double n1 = (x1[2] * x2[0] - x1[0] * x2[2]) + (x2[2] * x3[0] - x2[0] * x3[2]) + (x3[2] * x1[0] - x3[0] * x1[2]);
    double n2 = (x1[0] * x2[1] - x1[1] * x2[0]) + (x2[0] * x3[1] - x2[1] * x3[0]) + (x3[0] * x1[1] - x3[1] * x1[0]);
    double det = x1[0] * x2[1] * x3[2] + x2[0] * x3[1] * x1[2] + x3[0] * x1[1] * x2[2]
               - x1[0] * x3[1] * x2[2] - x2[0] * x1[1] * x3[2] - x3[0] * x2[1] * x1[2];
    double d = -det;
    Q9[0] += n0 * n0; Q9[1] += n0 * n1; Q9[2] += n0 * n2; Q9[3] += n0 * d;
    Q9[4] += n1 * n1; Q9[5] += n1 * n2; Q9[6] += n1 * d;
    Q9[7] += n2 * n2; Q9[8] += n2 * d;
}

// vertex area = sum incident face areas / 3 (Common/vtkSurface.cxx:1342-1359), faces in ascending id
__global__ void k_vertex_area(int V, const int* __restrict__ vf_ptr, const unsigned long long* __restrict__ vf_keys,
                              const float* __restrict__ xyz, const int* __restrict__ tri, double* area) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
        double A = 0;
        for (int i = vf_ptr[v]; i < vf_ptr[v + 1]; i++) {
            Tri3 t; load_face(xyz, tri, (int)(vf_keys[i] & 0xffffffffull), t);
            A += tri_area(t) / 3.0;
        }
        area[v] = A;
    }
}

// weight = area [* indicator^gradation]; float-rounded for the anisotropic metrics (their Item::Weight is float)
__global__ void k_raw_weight(int V, const double* __restrict__ area, const double* __restrict__ custom, double gradation,
                             int use_custom, int as_float, double* weight) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
        double w = area[v];
        if (use_custom) w = w * pow(custom[v], gradation);
        weight[v] = as_float ? (double)(float)w : w;
    }
}

// clamp to [avg/ratio, avg*ratio] (ClampWeights) and compose the payload row
template <int M>
__global__ void k_compose_items(int V, const double* __restrict__ sum_w, double ratio, double* weight,
                                const double* __restrict__ area, const float* __restrict__ xyz, const int* __restrict__ tri,
                                const int* __restrict__ vf_ptr, const unsigned long long* __restrict__ vf_keys,
                                const float* __restrict__ pd, double* items) {
    constexpr int NPAD = MetricTraits<M>::NPAD;
    const double avg = *sum_w / (double)V;
    const double mn = avg / ratio, mx = avg * ratio;
    constexpr bool is_float = (M == M_ANISO || M == M_ANISOQ);
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
        double w = weight[v];
        if (w > mx) w = is_float ? (double)(float)mx : mx;
        if (w < mn) w = is_float ? (double)(float)mn : mn;
        weight[v] = w;
        double row[NPAD];
#pragma unroll
        for (int k = 0; k < NPAD; k++) row[k] = 0.0;
        double p[3] = {(double)xyz[3 * (int64_t)v], (double)xyz[3 * (int64_t)v + 1], (double)xyz[3 * (int64_t)v + 2]};
        if (!is_float) {
            row[0] = p[0] * w; row[1] = p[1] * w; row[2] = p[2] * w; row[3] = w;
        } else {
            float val[3] = {(float)p[0], (float)p[1], (float)p[2]};
            float wf = (float)w;
            double A = area[v];
            double d[6];
#pragma unroll
            for (int j = 0; j < 6; j++) d[j] = pd ? (double)pd[6 * (int64_t)v + j] : 0.0;
            float T[6];
            T[0] = (float)(A * d[0] * d[0] + A * d[3] * d[3]);
            T[1] = (float)(A * d[0] * d[1] + A * d[3] * d[4]);
            T[2] = (float)(A * d[0] * d[2] + A * d[3] * d[5]);
            T[3] = (float)(A * d[1] * d[1] + A * d[5] * d[5]);   // reference quirk: d[5] where d[4] is expected
            T[4] = (float)(A * d[1] * d[2] + A * d[4] * d[5]);
            T[5] = (float)(A * d[2] * d[2] + A * d[5] * d[5]);
            float X0 = T[0] * val[0] + T[1] * val[1] + T[2] * val[2];
            float X1 = T[1] * val[0] + T[3] * val[1] + T[4] * val[2];
            float X2 = T[2] * val[0] + T[4] * val[1] + T[5] * val[2];
            row[0] = (double)(val[0] * wf); row[1] = (double)(val[1] * wf); row[2] = (double)(val[2] * wf); row[3] = (double)wf;
#pragma unroll
            for (int k = 0; k < 6; k++) row[4 + k] = (double)T[k];
            row[10] = (double)X0; row[11] = (double)X1; row[12] = (double)X2;
        }
        if (M == M_QEM || M == M_ANISOQ) {
            constexpr int QO = MetricTraits<M>::QOFF;
            double Q[9];
#pragma unroll
            for (int k = 0; k < 9; k++) Q[k] = 0.0;
            for (int i = vf_ptr[v]; i < vf_ptr[v + 1]; i++) {
                Tri3 t; load_face(xyz, tri, (int)(vf_keys[i] & 0xffffffffull), t);
                tri_quadric_add(t, Q);
            }
#pragma unroll
            for (int k = 0; k < 9; k++) row[QO + k] = Q[k];
        }
        store_row<NPAD>(items + (int64_t)v * NPAD, row);
    }
}

// host payload (V x NP, packed) <-> device rows (V x NPAD)
__global__ void k_pack_rows(int64_t n, int np, int npad, const double* __restrict__ in, double* out, int to_padded) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n * npad; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / npad; int k = (int)(i % npad);
        if (to_padded) out[i] = (k < np) ? in[r * np + k] : 0.0;
        else if (k < np) out[r * np + k] = in[i];
    }
}

// ELL copy of the adjacency for the frontier scan: column-major, ell[k * vpad + v] = k-th neighbour of v, v itself when
// the row is shorter, and -2 in the last column when the row is longer than W (the scan then finishes the row
// from the CSR).  Thread v reads ell[k * vpad + v]: every load of a warp is one coalesced 128-byte line.
__global__ void k_build_ell(int V, int64_t vpad, int W, const int* __restrict__ row_ptr, const int* __restrict__ col, int* ell) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < vpad; v += gridDim.x * blockDim.x) {
        const int beg = v < V ? row_ptr[v] : 0, deg = v < V ? row_ptr[v + 1] - beg : 0;
        for (int k = 0; k < W; k++) {
            int val = k < deg ? col[beg + k] : (v < V ? v : 0);   // short rows are padded with the vertex itself
            if (k == W - 1 && deg > W) val = -2;
            ell[(int64_t)k * vpad + v] = val;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// 1 -> 4 subdivision (vtkSurface::Subdivide, reference Common/vtkSurface.cxx:605-677): old points first, then one
// midpoint per edge in EDGE-ID order -- the reference numbers edges in the order AddEdge first sees them over the
// faces (Common/vtkSurfaceBase.cxx:1166-1221) --, then per face (V1,V4,V6) (V4,V2,V5) (V5,V3,V6) (V4,V5,V6).
// Edge ids on the device: sort the 3F undirected half-edges by (min,max) with their slot 3f+k as payload (stable, so
// the first entry of a run is the first occurrence), rank the runs by that first slot.

// undirected edge key of slot 3f+k, ~0 for inactive faces (first two vertices equal, vtkSurfaceBase.cxx:1443) / self loops
__global__ void k_sub_edge_keys(int F, const int* __restrict__ tri, unsigned long long* keys, int* slots) {
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < F; f += gridDim.x * blockDim.x) {
        const int v[3] = {tri[3 * f], tri[3 * f + 1], tri[3 * f + 2]};
        const bool active = v[0] != v[1];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const unsigned a = (unsigned)v[k], b = (unsigned)v[(k + 1) % 3];
            const bool ok = active && a != b;
            keys[3 * (int64_t)f + k] = ok ? (((unsigned long long)min(a, b) << 32) | max(a, b)) : ~0ull;
            slots[3 * (int64_t)f + k] = 3 * f + k;
        }
    }
}
// head[i] = 1 where a run of equal valid keys starts
__global__ void k_sub_heads(int64_t n, const unsigned long long* __restrict__ keys, int* head) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        head[i] = (keys[i] != ~0ull && (i == 0 || keys[i - 1] != keys[i])) ? 1 : 0;
}
// run r (= inclusive scan of head - 1) -> slot of its first occurrence
__global__ void k_sub_first_slots(int64_t n, const int* __restrict__ head, const int* __restrict__ run_incl, const int* __restrict__ slots,
                                  int* first_slot, int* run_id) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (head[i]) { first_slot[run_incl[i] - 1] = slots[i]; run_id[run_incl[i] - 1] = run_incl[i] - 1; }
}
// after sorting the runs by first slot: edge id of run run_sorted[e] is e; the edge's endpoints in first-seen order
__global__ void k_edge_of_run(int E, const int* __restrict__ run_sorted, int* edge_of_run) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) edge_of_run[run_sorted[e]] = e;
}
__global__ void k_sub_edges(int E, int V, const int* __restrict__ first_sorted, const int* __restrict__ tri,
                            const float* __restrict__ xyz, float* xyz_out, int* parent1, int* parent2) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
        const int s = first_sorted[e], f = s / 3, k = s % 3;
        const int a = tri[3 * (int64_t)f + k], b = tri[3 * (int64_t)f + (k + 1) % 3];
        parent1[V + e] = a; parent2[V + e] = b;
#pragma unroll
        for (int d = 0; d < 3; d++)
            xyz_out[3 * ((int64_t)V + e) + d] = (float)(0.5 * ((double)xyz[3 * (int64_t)a + d] + (double)xyz[3 * (int64_t)b + d]));
    }
}
__global__ void k_sub_old_points(int V, const float* __restrict__ xyz, float* xyz_out, int* parent1, int* parent2) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
        xyz_out[3 * (int64_t)v] = xyz[3 * (int64_t)v]; xyz_out[3 * (int64_t)v + 1] = xyz[3 * (int64_t)v + 1]; xyz_out[3 * (int64_t)v + 2] = xyz[3 * (int64_t)v + 2];
        parent1[v] = v; parent2[v] = v;
    }
}
// sorted position -> slot: the edge id of every half-edge slot
__global__ void k_sub_slot_edges(int64_t n, const unsigned long long* __restrict__ keys, const int* __restrict__ run_incl, const int* __restrict__ slots,
                                 const int* __restrict__ edge_of_run, int* edge_of_slot) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (keys[i] != ~0ull) edge_of_slot[slots[i]] = edge_of_run[run_incl[i] - 1];
}
// faces: 4 per face with three distinct vertices, written at 4 x (rank of the face among those)
__global__ void k_sub_face_flags(int F, const int* __restrict__ tri, int* flag) {
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < F; f += gridDim.x * blockDim.x) {
        const int a = tri[3 * f], b = tri[3 * f + 1], c = tri[3 * f + 2];
        flag[f] = (a != b && b != c && a != c) ? 1 : 0;
    }
}
__global__ void k_sub_faces(int F, int V, const int* __restrict__ tri, const int* __restrict__ flag, const int* __restrict__ rank_incl,
                            const int* __restrict__ edge_of_slot, int* tri_out) {
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < F; f += gridDim.x * blockDim.x) {
        if (!flag[f]) continue;
        const int v1 = tri[3 * f], v2 = tri[3 * f + 1], v3 = tri[3 * f + 2];
        const int v4 = V + edge_of_slot[3 * (int64_t)f], v5 = V + edge_of_slot[3 * (int64_t)f + 1], v6 = V + edge_of_slot[3 * (int64_t)f + 2];
        int* o = tri_out + 12 * (int64_t)(rank_incl[f] - 1);
        o[0] = v1; o[1] = v4; o[2] = v6;
        o[3] = v4; o[4] = v2; o[5] = v5;
        o[6] = v5; o[7] = v3; o[8] = v6;
        o[9] = v4; o[10] = v5; o[11] = v6;
    }
}

// ---------------------------------------------------------------------------------------------------
// vtkSurface::SplitLongEdges (reference Common/vtkSurface.cxx:444-604, option -l): threshold = ratio x mean edge length
// of the mesh as given; then passes until nothing is split: every edge longer than the threshold gets a midpoint vertex
// (added in edge-id order), every triangle is replaced according to which of its edges were cut -- Split2 / Split3
// (:429-443) or the 1 -> 4 pattern (:569-576).  Each pass works on a snapshot of the mesh, exactly like the reference's
// loop over the edges and the first NumCells faces, so it is one set of data-parallel kernels per pass.  Edge ids inside a
// pass are first-seen ids of the current mesh (upstream recycles deleted slots, its numbering is unobservable).
__global__ void k_split_lengths(int E, const int* __restrict__ first_sorted, const int* __restrict__ tri, const float* __restrict__ xyz, double* len) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
        const int s = first_sorted[e], f = s / 3, k = s % 3;
        const int a = tri[3 * (int64_t)f + k], b = tri[3 * (int64_t)f + (k + 1) % 3];
        double d2 = 0;
#pragma unroll
        for (int d = 0; d < 3; d++) { const double t = (double)xyz[3 * (int64_t)a + d] - (double)xyz[3 * (int64_t)b + d]; d2 += t * t; }
        len[e] = sqrt(d2);
    }
}
__global__ void k_split_mark(int E, const double* __restrict__ len, double threshold, int* mark) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) mark[e] = len[e] > threshold ? 1 : 0;
}
// midpoints of the marked edges: vertex V + (rank of the edge among the marked ones, in edge-id order)
__global__ void k_split_points(int E, int V, const int* __restrict__ mark, const int* __restrict__ rank_excl, const int* __restrict__ first_sorted,
                               const int* __restrict__ tri, float* xyz, int* parent1, int* parent2) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
        if (!mark[e]) continue;
        const int s = first_sorted[e], f = s / 3, k = s % 3;
        const int a = tri[3 * (int64_t)f + k], b = tri[3 * (int64_t)f + (k + 1) % 3];
        const int64_t m = (int64_t)V + rank_excl[e];
        parent1[m] = a; parent2[m] = b;
#pragma unroll
        for (int d = 0; d < 3; d++) xyz[3 * m + d] = (float)(0.5 * ((double)xyz[3 * (int64_t)a + d] + (double)xyz[3 * (int64_t)b + d]));
    }
}
// mid-vertex (or -1) of the three sides 12, 23, 31 of face f
__device__ __forceinline__ void split_mids(int f, int V, const int* __restrict__ tri, const int* __restrict__ edge_of_slot, const int* __restrict__ mark,
                                           const int* __restrict__ rank_excl, int& m12, int& m23, int& m13) {
    const bool active = tri[3 * (int64_t)f] != tri[3 * (int64_t)f + 1];
    int m[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const bool loop = tri[3 * (int64_t)f + k] == tri[3 * (int64_t)f + (k + 1) % 3];
        const int e = (active && !loop) ? edge_of_slot[3 * (int64_t)f + k] : -1;
        m[k] = (e >= 0 && mark[e]) ? V + rank_excl[e] : -1;
    }
    m12 = m[0]; m23 = m[1]; m13 = m[2];
}
__global__ void k_split_count(int F, int V, const int* __restrict__ tri, const int* __restrict__ edge_of_slot, const int* __restrict__ mark,
                              const int* __restrict__ rank_excl, int* n_child) {
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < F; f += gridDim.x * blockDim.x) {
        int m12, m23, m13;
        split_mids(f, V, tri, edge_of_slot, mark, rank_excl, m12, m23, m13);
        n_child[f] = 1 + (m12 >= 0) + (m23 >= 0) + (m13 >= 0);
    }
}
__global__ void k_split_emit(int F, int V, const int* __restrict__ tri, const int* __restrict__ edge_of_slot, const int* __restrict__ mark,
                             const int* __restrict__ rank_excl, const int* __restrict__ child_excl, int* tri_out) {
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < F; f += gridDim.x * blockDim.x) {
        int v12, v23, v13;
        split_mids(f, V, tri, edge_of_slot, mark, rank_excl, v12, v23, v13);
        const int v1 = tri[3 * (int64_t)f], v2 = tri[3 * (int64_t)f + 1], v3 = tri[3 * (int64_t)f + 2];
        int* o = tri_out + 3 * (int64_t)child_excl[f];
        auto face = [&](int i, int a, int b, int c) { o[3 * i] = a; o[3 * i + 1] = b; o[3 * i + 2] = c; };
        // Split2(f, a, b, c, ab): (a, ab, c) (ab, b, c);  Split3(f, a, b, c, ab, ac): (a, ab, ac) (ab, b, c) (c, ac, ab)
        if (v12 < 0) {
            if (v13 < 0) {
                if (v23 < 0) face(0, v1, v2, v3);
                else { face(0, v2, v23, v1); face(1, v23, v3, v1); }                                   // Split2(v2, v3, v1, v23)
            } else {
                if (v23 < 0) { face(0, v3, v13, v2); face(1, v13, v1, v2); }                            // Split2(v3, v1, v2, v13)
                else { face(0, v3, v13, v23); face(1, v13, v1, v2); face(2, v2, v23, v13); }            // Split3(v3, v1, v2, v13, v23)
            }
        } else {
            if (v13 < 0) {
                if (v23 < 0) { face(0, v1, v12, v3); face(1, v12, v2, v3); }                            // Split2(v1, v2, v3, v12)
                else { face(0, v2, v23, v12); face(1, v23, v3, v1); face(2, v1, v12, v23); }            // Split3(v2, v3, v1, v23, v12)
            } else {
                if (v23 < 0) { face(0, v1, v12, v13); face(1, v12, v2, v3); face(2, v3, v13, v12); }    // Split3(v1, v2, v3, v12, v13)
                else { face(0, v1, v12, v13); face(1, v12, v2, v23); face(2, v23, v3, v13); face(3, v12, v23, v13); }
            }
        }
    }
}

__global__ void k_max_degree(int V, const int* __restrict__ row_ptr, int* out) {
    int m = 0;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) m = max(m, row_ptr[v + 1] - row_ptr[v]);
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

}  // namespace acvd
