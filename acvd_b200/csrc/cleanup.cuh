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
// per face: sorted cluster triple packed 3 x 21 bits when the three clusters are distinct and valid, else ~0
__global__ void k_dual_keys(int F, int K, const int* __restrict__ tri, const int* __restrict__ cid,
                            unsigned long long* keys, int* face_id) {
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < F; f += gridDim.x * blockDim.x) {
        int a = cid[tri[3 * f]], b = cid[tri[3 * f + 1]], c = cid[tri[3 * f + 2]];
        unsigned long long k = ~0ull;
        if (a >= 0 && b >= 0 && c >= 0 && a < K && b < K && c < K && a != b && a != c && b != c) {
            int lo = min(a, min(b, c)), hi = max(a, max(b, c)), mid = a + b + c - lo - hi;
            k = ((unsigned long long)lo << 42) | ((unsigned long long)mid << 21) | (unsigned long long)hi;
        }
        keys[f] = k;
        face_id[f] = f;
    }
}
// after a stable sort by key: flag[i] = 1 for the first face of every distinct valid key
__global__ void k_dual_first(int F, const unsigned long long* __restrict__ keys, const int* __restrict__ face_id, int* first_face) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < F; i += gridDim.x * blockDim.x) {
        bool first = keys[i] != ~0ull && (i == 0 || keys[i - 1] != keys[i]);
        first_face[i] = first ? face_id[i] : 0x7fffffff;
    }
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
