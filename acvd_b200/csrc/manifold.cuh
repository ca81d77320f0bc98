// Manifoldness of mesh vertices on the device: the predicate of vtkSurfaceBase::IsVertexManifold
// (reference Common/vtkSurfaceBase.cxx:259-317) for the input mesh and for the dual (output) mesh that
// vtkDiscreteRemeshing::BuildDelaunayTriangulation produces (DiscreteRemeshing/vtkDiscreteRemeshing.h:1003-1133),
// as DetectNonManifoldOutputVertices needs it (:166-383, the -m 1 loop).
//
// The reference walks the fan of faces around the vertex from its first edge.  It first rejects the vertex when it has
// fewer than two edges or when one of its edges is not "manifold", which for the reference means: does not carry
// exactly two faces (IsEdgeManifold, vtkSurfaceBase.h:521-528: Poly2 < 0 or a non-empty NonManifoldFaces list).
// With every edge on exactly two faces the faces around the vertex form closed fans, and the walk returns true iff the
// fan it starts on reaches every edge -- i.e. iff there is exactly one fan.  So:
//     manifold(v)  <=>  #edges >= 2  and  every edge of v lies on exactly 2 faces of v  and  the link graph of v
//                       (edges of v as nodes, joined when they share a face) is connected,
// which does not depend on where the walk starts.  One thread per vertex, neighbours in a local list, union-find on
// the list slots.
#pragma once
#include "common.cuh"

namespace acvd {

struct FanMesh {
    int n;                                   // vertices
    const int* __restrict__ nb_ptr;          // neighbours (edges) of every vertex
    const int* __restrict__ nb;
    const int* __restrict__ f_ptr;           // incident faces of every vertex: ids in f_ids, or in the low words of f_keys
    const int* __restrict__ f_ids;
    const unsigned long long* __restrict__ f_keys;
    const int* __restrict__ tri;             // 3 vertex ids per face
};

// flags[v] = 1 manifold, 0 not manifold, 2 more than kMaxRing edges (left to the caller)
__global__ void __launch_bounds__(128) k_vertex_manifold(FanMesh M, unsigned char* flags) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < M.n; v += gridDim.x * blockDim.x) {
        const int b0 = M.nb_ptr[v], ne = M.nb_ptr[v + 1] - b0;
        if (ne < 2) { flags[v] = 0; continue; }
        if (ne > kMaxRing) { flags[v] = 2; continue; }
        unsigned char cnt[kMaxRing], par[kMaxRing];
        for (int j = 0; j < ne; j++) { cnt[j] = 0; par[j] = (unsigned char)j; }
        auto slot = [&](int u) { for (int j = 0; j < ne; j++) if (M.nb[b0 + j] == u) return j; return -1; };
        auto find = [&](int x) { while (par[x] != x) { par[x] = par[par[x]]; x = par[x]; } return x; };
        bool ok = true;
        for (int i = M.f_ptr[v]; i < M.f_ptr[v + 1]; i++) {
            const int f = M.f_ids ? M.f_ids[i] : (int)(M.f_keys[i] & 0xffffffffull);
            const int t[3] = {M.tri[3 * (int64_t)f], M.tri[3 * (int64_t)f + 1], M.tri[3 * (int64_t)f + 2]};
            int o[2], m = 0;
            for (int k = 0; k < 3; k++) if (t[k] != v) { if (m < 2) o[m] = t[k]; m++; }
            if (m != 2 || o[0] == o[1]) continue;                   // degenerate face
            const int ja = slot(o[0]), jb = slot(o[1]);
            if (ja < 0 || jb < 0) { ok = false; break; }
            if (cnt[ja] < 255) cnt[ja]++;
            if (cnt[jb] < 255) cnt[jb]++;
            const int ra = find(ja), rb = find(jb);
            if (ra != rb) par[ra > rb ? ra : rb] = (unsigned char)(ra > rb ? rb : ra);
        }
        if (ok) {
            const int r0 = find(0);
            for (int j = 0; j < ne && ok; j++) ok = cnt[j] == 2 && find(j) == r0;
        }
        flags[v] = ok ? 1 : 0;
    }
}

// ---- incidence of the dual mesh, by counting
__global__ void k_count_tri_corners(int n_tri, const int* __restrict__ tri, int* cnt) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 3 * n_tri; i += gridDim.x * blockDim.x) atomicAdd(cnt + tri[i], 1);
}
__global__ void k_scatter_tri_corners(int n_tri, const int* __restrict__ tri, const int* __restrict__ ptr, int* cursor, int* out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 3 * n_tri; i += gridDim.x * blockDim.x) {
        const int c = tri[i];
        out[ptr[c] + atomicAdd(cursor + c, 1)] = i / 3;
    }
}
// pairs: (lo << 32 | hi), unique
__global__ void k_count_pair_ends(int64_t n, const unsigned long long* __restrict__ pairs, int* cnt) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        atomicAdd(cnt + (int)(pairs[i] >> 32), 1);
        atomicAdd(cnt + (int)(pairs[i] & 0xffffffffull), 1);
    }
}
__global__ void k_scatter_pair_ends(int64_t n, const unsigned long long* __restrict__ pairs, const int* __restrict__ ptr, int* cursor, int* out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int a = (int)(pairs[i] >> 32), b = (int)(pairs[i] & 0xffffffffull);
        out[ptr[a] + atomicAdd(cursor + a, 1)] = b;
        out[ptr[b] + atomicAdd(cursor + b, 1)] = a;
    }
}
// the three edges of every dual triangle as (lo << 32 | hi) keys
__global__ void k_tri_edge_keys(int n_tri, const int* __restrict__ tri, unsigned long long* keys) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 3 * n_tri; i += gridDim.x * blockDim.x) {
        const int f = i / 3, k = i % 3;
        const unsigned a = (unsigned)tri[3 * f + k], b = (unsigned)tri[3 * f + (k + 1) % 3];
        keys[i] = ((unsigned long long)min(a, b) << 32) | max(a, b);
    }
}

}  // namespace acvd
