// Per-cluster statistic accumulation, cluster cleaning (connected components), hole filling, and the
// integer stages that follow the hot path.
//
//   ReComputeStatistics / ReComputeClustersSize  reference Common/vtkUniformClustering.h:353-403
//   CleanClustering                              :406-549
//   FillHolesInClustering                        :552-633
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

// seg[c] = first index i with sorted_keys[i] >= c, for c in [0, K+1]; seg[K+1] = V
__global__ void k_segments(int V, int K, const int* __restrict__ sorted_keys, int* seg) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c <= K + 1; c += gridDim.x * blockDim.x) {
        int lo = 0, hi = V;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (sorted_keys[mid] < c) lo = mid + 1; else hi = mid; }
        seg[c] = (c == K + 1) ? V : lo;
    }
}

// One warp per cluster: lanes stride over the cluster's items (sorted by vertex id, so the summation
// order is fixed), each payload component is reduced with warp shuffles; lane 0 stores sums, size,
// representative point and energy.  Deterministic.  M: metric of the stored rows; EM: metric whose
// energy formula is evaluated (QEM's unconstrained phase uses the isotropic one).
template <int M, int EM>
__global__ void __launch_bounds__(kThreads) k_cluster_stats(int K, const int* __restrict__ seg, const int* __restrict__ sorted_v,
                                                            const double* __restrict__ items, double* csum, double* cenergy,
                                                            double* ccentroid, int* csize, const int* __restrict__ anchor,
                                                            const float* __restrict__ xyz, EvalCfg cfg) {
    constexpr int NPAD = MetricTraits<M>::NPAD;
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    for (int c = blockIdx.x * warps_per_block + (threadIdx.x >> 5); c < K; c += gridDim.x * warps_per_block) {
        const int b = seg[c], e = seg[c + 1];
        double acc[NPAD];
#pragma unroll
        for (int k = 0; k < NPAD; k++) acc[k] = 0.0;
        for (int i = b + lane; i < e; i += 32) {
            double it[NPAD];
            load_row_ro<NPAD>(items + (int64_t)sorted_v[i] * NPAD, it);
#pragma unroll
            for (int k = 0; k < NPAD; k++) acc[k] += it[k];
        }
#pragma unroll
        for (int k = 0; k < NPAD; k++) acc[k] = warp_sum(acc[k]);
        if (lane == 0) {
            store_row<NPAD>(csum + (int64_t)c * NPAD, acc);
            csize[c] = e - b;
            double cen[3], apt[3];
            const double* ap = nullptr;
            if (EM == M_QEM && anchor && anchor[c] >= 0) {
                int av = anchor[c];
                apt[0] = xyz[3 * av]; apt[1] = xyz[3 * av + 1]; apt[2] = xyz[3 * av + 2];
                ap = apt;
            }
            cenergy[c] = cluster_energy<EM>(acc, cfg, cen, ap);
            ccentroid[3 * c] = cen[0]; ccentroid[3 * c + 1] = cen[1]; ccentroid[3 * c + 2] = cen[2];
        }
    }
}

// ---------------- connected components of every cluster (label = min vertex id of the component) ----------------
// Connected components of every cluster by hooking (union-find on the GPU): label[] is a parent array
// initialised to the identity; every same-cluster edge (u < v) joins the two trees by atomically hooking
// the larger root under the smaller one, so a component's root is its minimum vertex id -- which is also the
// vertex at which the reference's index-ordered BFS discovers the component (:428-437).  One pass over the
// edges (k_cc_hook) between two flattening passes (k_cc_flatten), and only for the clusters whose atomic-free
// initial forest (k_cc_init) has more than one root.
__device__ __forceinline__ int cc_find(const int* label, int x) {
    int p = __ldcg(label + x);
    while (p != x) { x = p; p = __ldcg(label + x); }
    return x;
}
// find with path halving; the shortcut is written with atomicMin so parents only ever decrease
__device__ __forceinline__ int cc_find_compress(int* label, int x) {
    while (true) {
        const int p = __ldcg(label + x);
        if (p == x) return x;
        const int gp = __ldcg(label + p);
        if (gp == p) return p;
        atomicMin(label + x, gp);
        x = gp;
    }
}

// joins the trees rooted at (or above) ru and rv; returns the smaller representative reached.  A failed CAS hands
// back the parent somebody else installed, which is closer to the root: the walk continues from there (ECL-CC).
__device__ __forceinline__ int cc_hook_roots(int* label, int ru, int rv) {
    while (ru != rv) {
        if (rv < ru) {
            const int ret = atomicCAS(label + ru, ru, rv);
            if (ret == ru) return rv;
            ru = ret;
        } else {
            const int ret = atomicCAS(label + rv, rv, ru);
            if (ret == rv) return ru;
            rv = ret;
        }
    }
    return rv;
}

// Initial forest without atomics (the first step of ECL-CC): every vertex points at its smallest same-cluster
// neighbour with a smaller id, or at itself.  parent <= self everywhere and every link is a real same-cluster edge,
// so the hooking pass that follows only has to join the few trees this leaves per cluster; n_roots[c] counts them.
template <int W>
__global__ void __launch_bounds__(kThreads) k_cc_init(int V, int K, int64_t vpad, const int* __restrict__ ell,
                                                      const int* __restrict__ row_ptr, const int* __restrict__ col,
                                                      const int* __restrict__ cid, int* label, int* n_roots) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
        const int c = cid[v];
        int best = v;
        if (c < K) {
            int nb[W];
#pragma unroll
            for (int k = 0; k < W; k++) nb[k] = __ldg(ell + (int64_t)k * vpad + v);
            const bool overflow = nb[W - 1] == -2;
#pragma unroll
            for (int k = 0; k < W; k++) {
                const int u = nb[k];
                if (u >= 0 && u < best && cid[u] == c) best = u;
            }
            if (overflow)
                for (int e = row_ptr[v] + W - 1; e < row_ptr[v + 1]; e++) {
                    const int u = col[e];
                    if (u < best && cid[u] == c) best = u;
                }
        }
        label[v] = best;
        // a cluster whose initial forest has a single root is connected (one tree spans it): only the clusters with
        // several roots go through the hooking and the component bookkeeping
        if (c < K && best == v) atomicAdd(n_roots + c, 1);
    }
}

template <int W>
__global__ void __launch_bounds__(kThreads) k_cc_hook(int V, int K, int64_t vpad, const int* __restrict__ ell,
                                                      const int* __restrict__ row_ptr, const int* __restrict__ col,
                                                      const int* __restrict__ cid, int* label, const int* __restrict__ n_roots) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
        const int c = cid[v];
        if (c >= K || n_roots[c] <= 1) continue;
        int nb[W];
#pragma unroll
        for (int k = 0; k < W; k++) nb[k] = __ldg(ell + (int64_t)k * vpad + v);
        const bool overflow = nb[W - 1] == -2;
        int rv = cc_find_compress(label, v);          // own representative: found once, updated by every hook
#pragma unroll
        for (int k = 0; k < W; k++) {
            const int u = nb[k];
            if (u >= 0 && u < v && cid[u] == c) rv = cc_hook_roots(label, cc_find_compress(label, u), rv);
        }
        if (overflow)
            for (int e = row_ptr[v] + W - 1; e < row_ptr[v + 1]; e++) {
                const int u = col[e];
                if (u < v && cid[u] == c) rv = cc_hook_roots(label, cc_find_compress(label, u), rv);
            }
    }
}

__global__ void __launch_bounds__(kThreads) k_cc_flatten(int V, int K, const int* __restrict__ cid, const int* __restrict__ n_roots, int* label) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
        const int c = cid[v];
        if (c >= K || n_roots[c] <= 1) continue;
        label[v] = cc_find(label, v);
    }
}

__global__ void k_cc_sizes(int V, int K, const int* __restrict__ cid, const int* __restrict__ label, int* comp_size,
                           const int* __restrict__ anchor, const int* __restrict__ n_roots) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
        int c = cid[v];
        if (c >= K || n_roots[c] <= 1) continue;
        // an anchored item weighs 1e9 so that its component always wins (:440-447)
        int w = (anchor && anchor[c] == v) ? 1000000000 : 1;
        atomicAdd(&comp_size[label[v]], w);
    }
}

// Per cluster: number of recorded components and the winner (largest; first-discovered wins ties, :503).
// Reference quirk kept: the component discovered at item 0 is never recorded because 0 doubles as the
// "unvisited" sentinel of VisitedCluster (:463-467), so it is neither counted nor ever reset.
__global__ void k_cc_winner(int V, int K, const int* __restrict__ cid, const int* __restrict__ label,
                            const int* __restrict__ comp_size, int* n_comp, unsigned long long* winner, const int* __restrict__ n_roots) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
        int c = cid[v];
        if (c >= K || n_roots[c] <= 1 || label[v] != v || v == 0) continue;
        atomicAdd(&n_comp[c], 1);
        unsigned long long key = ((unsigned long long)(unsigned)comp_size[v] << 32) | (unsigned)(0xffffffffu - (unsigned)v);
        atomicMax(&winner[c], key);
    }
}

__global__ void k_cc_apply(int V, int K, int* cid, const int* __restrict__ label, const int* __restrict__ n_comp,
                           const unsigned long long* __restrict__ winner, unsigned long long* n_reset) {
    unsigned cnt = 0;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
        int c = cid[v];
        if (c >= K) continue;
        int root = label[v];
        if (root == 0 || n_comp[c] < 2) continue;
        unsigned win_root = 0xffffffffu - (unsigned)(winner[c] & 0xffffffffull);
        if ((unsigned)root != win_root) { cid[v] = K; cnt++; }
    }
    warp_count_add(n_reset, cnt);
}

__global__ void k_count_ge2(int K, const int* __restrict__ n_comp, unsigned long long* out) {
    unsigned cnt = 0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < K; c += gridDim.x * blockDim.x) cnt += n_comp[c] >= 2;
    warp_count_add(out, cnt);
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
