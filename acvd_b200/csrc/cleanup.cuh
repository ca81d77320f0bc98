// Small helpers and the integer stages that follow the hot path.
// (Statistic accumulation and CleanClustering live in sparse.cuh: k_cluster_pass; FillHoles in fill.cuh.)
//
//   boundary detection / cluster adjacency / dual triangles  DiscreteRemeshing/vtkDiscreteRemeshing.h:1003-1133
#pragma once
#include "metric.cuh"

namespace acvd {

__global__ void k_fill(int n, int value, int* out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = value;
}
__global__ void k_iota(int n, int* out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = i;
}

// ---------------- hole filling (fill.cuh): list of the NULL vertices ----------------
__global__ void k_collect_null(int V, int K, const int* __restrict__ cid, int* list, unsigned long long* n) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x)
        if (cid[v] >= K || cid[v] < 0) { int s = (int)atomicAdd(n, 1ull); list[s] = v; }   // ids are normalised to [0, K] on upload
}
// ---------------- integer stages ----------------
__global__ void k_boundary_flags(int V, const int* __restrict__ row_ptr, const int* __restrict__ col,
                                 const int* __restrict__ cid, unsigned char* flags) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
        int a = cid[v];
        unsigned char b = 0;
        for (int e = row_ptr[v]; e < row_ptr[v + 1]; e++) if (cid[col[e]] != a) { b = 1; break; }
        flags[v] = b;
    }
}
// one key per directed CSR entry u<v with different, valid clusters: (lo << 32 | hi), else ~0
__global__ void k_adjacency_keys(int V, int K, const int* __restrict__ row_ptr, const int* __restrict__ col,
                                 const int* __restrict__ cid, unsigned long long* keys) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
        int a = cid[v];
        for (int e = row_ptr[v]; e < row_ptr[v + 1]; e++) {
            int u = col[e];
            int b = cid[u];
            unsigned long long k = ~0ull;
            if (u > v && a != b && a >= 0 && a < K && b >= 0 && b < K) {
                unsigned lo = (unsigned)min(a, b), hi = (unsigned)max(a, b);
                k = ((unsigned long long)lo << 32) | hi;
            }
            keys[e] = k;
        }
    }
}
// per face: the sorted cluster triple when the three clusters are distinct and valid -- (mid << 32 | hi) as the first
// sort key, lo as the second -- else all-ones sentinels (they sort to the end)
__global__ void k_dual_keys(int F, int K, const int* __restrict__ tri, const int* __restrict__ cid,
                            unsigned long long* key_mid_hi, unsigned* key_lo, int* face_id) {
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < F; f += gridDim.x * blockDim.x) {
        const int a = cid[tri[3 * (int64_t)f]], b = cid[tri[3 * (int64_t)f + 1]], c = cid[tri[3 * (int64_t)f + 2]];
        unsigned long long k = ~0ull;
        unsigned l = 0xffffffffu;
        if (a >= 0 && b >= 0 && c >= 0 && a < K && b < K && c < K && a != b && a != c && b != c) {
            const int lo = min(a, min(b, c)), hi = max(a, max(b, c)), mid = a + b + c - lo - hi;
            k = ((unsigned long long)(unsigned)mid << 32) | (unsigned)hi;
            l = (unsigned)lo;
        }
        key_mid_hi[f] = k;
        key_lo[f] = l;
        face_id[f] = f;
    }
}
__global__ void k_gather_u32(int n, const int* __restrict__ idx, const unsigned* __restrict__ in, unsigned* out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = in[idx[i]];
}
// faces ordered by (lo, mid, hi), equal triples in ascending face order: the first face of every run of a valid triple
__global__ void k_dual_first(int F, int K, const int* __restrict__ face_sorted, const int* __restrict__ tri, const int* __restrict__ cid,
                             int* first_face, unsigned long long* n_first) {
    unsigned cnt = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < F; i += gridDim.x * blockDim.x) {
        auto triple = [&](int f, int* t) {
            const int a = cid[tri[3 * (int64_t)f]], b = cid[tri[3 * (int64_t)f + 1]], c = cid[tri[3 * (int64_t)f + 2]];
            if (!(a >= 0 && b >= 0 && c >= 0 && a < K && b < K && c < K && a != b && a != c && b != c)) return false;
            t[0] = min(a, min(b, c)); t[2] = max(a, max(b, c)); t[1] = a + b + c - t[0] - t[2];
            return true;
        };
        int t[3], u[3];
        const int f = face_sorted[i];
        bool first = triple(f, t);
        if (first && i > 0 && triple(face_sorted[i - 1], u)) first = !(t[0] == u[0] && t[1] == u[1] && t[2] == u[2]);
        first_face[i] = first ? f : 0x7fffffff;
        cnt += first ? 1u : 0u;
    }
    warp_count_add(n_first, cnt);
}
__global__ void k_dual_emit(int n, const int* __restrict__ faces, const int* __restrict__ tri, const int* __restrict__ cid, int* out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int f = faces[i];
        out[3 * i] = cid[tri[3 * f]]; out[3 * i + 1] = cid[tri[3 * f + 1]]; out[3 * i + 2] = cid[tri[3 * f + 2]];
    }
}

// batched vtkQuadricTools::ComputeRepresentativePoint
__global__ void k_representative_points(int n, const double* __restrict__ Q9, double* P3, int max_sv, double thr, int* rank_def) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double q[9], p[3];
#pragma unroll
        for (int k = 0; k < 9; k++) q[k] = Q9[9 * (int64_t)i + k];
        p[0] = P3[3 * (int64_t)i]; p[1] = P3[3 * (int64_t)i + 1]; p[2] = P3[3 * (int64_t)i + 2];
        int rd = representative_point(q, p, max_sv, thr);
        P3[3 * (int64_t)i] = p[0]; P3[3 * (int64_t)i + 1] = p[1]; P3[3 * (int64_t)i + 2] = p[2];
        if (rank_def) rank_def[i] = rd;
    }
}

}  // namespace acvd
