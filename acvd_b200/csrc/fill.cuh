// FillHolesInClustering (reference Common/vtkUniformClustering.h:552-633), order-exact.
//
// The reference drains a FIFO of edges: first every edge with exactly one unassigned (NULL, id K) end, in edge-id
// order; a popped edge whose one end is assigned and whose other end is still NULL hands the cluster over -- unless
// ConnexityConstraintProblem(item, edge, K, cluster) objects, i.e. unless taking the item out of the NULL "cluster"
// would split the item's NULL ring neighbours (:606-607, active only while ConnexityConstraint is on) -- and pushes
// the edge ring of the adopted item.  Which cluster a NULL vertex ends up in therefore depends on the FIFO order.
//
// Edge ids are "first seen over the faces" (Common/vtkSurfaceBase.cxx:1166-1221, 1446-1451) and a vertex ring lists its
// edges in creation order (:1057-1068), so both orders are the order of the edges' first half-edge slot
//     slot(a, b) = min over the faces f holding a and b of 3 f + side,      side k of f = (t[k], t[k+1 mod 3]),
// which is computed locally from the vertex -> face incidence: no global edge table is needed on the device.
//
//  * connexity off (the priming call and the first convergence event, :727, :790): the FIFO is a multi-source BFS.
//    Level 1 = NULL vertices with an assigned neighbour; each adopts over its smallest-slot edge.  A level-(L+1)
//    vertex is reached first by the edge pushed earliest: smallest (adoption sequence number of the level-L parent,
//    slot of the edge in the parent's ring).  One pick + apply pass per level, a sort of the level's keys in between
//    gives the adoption sequence numbers.  Bit-identical to the sequential FIFO (tests/test_gpu_parity.py).
//  * connexity on: the guard makes every adoption depend on all earlier ones.  The holes CleanClustering leaves at that
//    stage are a handful of vertices, so the FIFO itself is replayed by one thread (k_fill_sequential) on the
//    slot-sorted initial edge list -- still bit-identical.  Only above kFillSequentialCap NULL vertices does the
//    driver fall back to the level-synchronous passes with the guard evaluated per level (documented deviation).
#pragma once
#include "reassign.cuh"

namespace acvd {

constexpr int kFillSequentialCap = 8192;      // NULL vertices the single-thread FIFO replay is used for

struct FillMesh {
    int V, K;
    const int* __restrict__ row_ptr;
    const int* __restrict__ col;
    const int* __restrict__ vf_ptr;
    const unsigned long long* __restrict__ vf_keys;
    const int* __restrict__ tri;
};

// first half-edge slot of the mesh edge (a, b): the order of the reference's edge ids and ring entries
__device__ __forceinline__ unsigned edge_slot(const FillMesh& M, int a, int b) {
    unsigned best = 0xffffffffu;
    for (int i = M.vf_ptr[a]; i < M.vf_ptr[a + 1]; i++) {
        const int f = (int)(M.vf_keys[i] & 0xffffffffull);
        const int t0 = M.tri[3 * (int64_t)f], t1 = M.tri[3 * (int64_t)f + 1], t2 = M.tri[3 * (int64_t)f + 2];
        const int ia = (t0 == a) ? 0 : ((t1 == a) ? 1 : 2);
        const int ib = (t0 == b) ? 0 : ((t1 == b) ? 1 : ((t2 == b) ? 2 : -1));
        if (ib < 0 || ib == ia) continue;
        const int lo = min(ia, ib), hi = max(ia, ib);
        const int side = (lo == 0 && hi == 1) ? 0 : ((lo == 1) ? 1 : 2);
        best = min(best, 3u * (unsigned)f + (unsigned)side);
    }
    return best;
}

// The neighbours of every vertex in the reference's ring order (edge creation order = ascending first half-edge slot),
// written over the CSR layout: what ComputeInitialRandomSampling's region growing walks (host_sampling.hpp).
__global__ void __launch_bounds__(kThreads) k_ring_order(FillMesh M, int* ring_col) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < M.V; v += gridDim.x * blockDim.x) {
        const int beg = M.row_ptr[v], deg = M.row_ptr[v + 1] - beg;
        if (deg <= 16) {
            unsigned s[16];
            int u[16];
            for (int k = 0; k < deg; k++) {
                const int w = M.col[beg + k];
                const unsigned sl = edge_slot(M, v, w);
                int j = k;
                while (j > 0 && s[j - 1] > sl) { s[j] = s[j - 1]; u[j] = u[j - 1]; j--; }
                s[j] = sl; u[j] = w;
            }
            for (int k = 0; k < deg; k++) ring_col[beg + k] = u[k];
        } else {        // long rows: selection sort in place over the output row
            for (int k = 0; k < deg; k++) ring_col[beg + k] = M.col[beg + k];
            for (int k = 0; k < deg; k++) {
                int best = k;
                unsigned sb = edge_slot(M, v, ring_col[beg + k]);
                for (int j = k + 1; j < deg; j++) { const unsigned sj = edge_slot(M, v, ring_col[beg + j]); if (sj < sb) { sb = sj; best = j; } }
                const int t = ring_col[beg + k]; ring_col[beg + k] = ring_col[beg + best]; ring_col[beg + best] = t;
            }
        }
    }
}

// NULL vertices in ascending order are collected by k_collect_null + a sort (cleanup.cuh); ids below 0 are
// normalised to K first so that "NULL" is one value everywhere.
__global__ void k_normalise_null(int V, int K, int* cid) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x)
        if (cid[v] < 0 || cid[v] > K) cid[v] = K;
}

// one BFS level: key / pick for every still-NULL listed vertex that touches the previous level (level 1: an assigned vertex)
__global__ void __launch_bounds__(kThreads) k_fill_level_pick(FillMesh M, int n, const int* __restrict__ list, const int* __restrict__ cid,
                                                              const int* __restrict__ lvl, const int* __restrict__ seq, int level, int guard,
                                                              unsigned long long* key_out, int* pick_out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int v = list[i];
        unsigned long long key = ~0ull;
        int pick = -1;
        if (cid[v] == M.K) {
            for (int e = M.row_ptr[v]; e < M.row_ptr[v + 1]; e++) {
                const int u = M.col[e];
                const int cu = cid[u];
                if (cu == M.K) continue;
                if (level > 1 && lvl[u] != level - 1) continue;
                const unsigned long long k = (level == 1 ? 0ull : ((unsigned long long)(unsigned)seq[u] << 32)) | edge_slot(M, v, u);
                if (k < key) { key = k; pick = cu; }
            }
            // level-synchronous form of the connexity guard (only used above kFillSequentialCap NULL vertices)
            if (pick >= 0 && guard && connexity_problem(v, M.K, M.row_ptr, M.col, cid)) pick = -1;
        }
        key_out[i] = key;
        pick_out[i] = pick;
    }
}

__global__ void __launch_bounds__(kThreads) k_fill_level_apply(int n, const int* __restrict__ list, const unsigned long long* __restrict__ key,
                                                               const int* __restrict__ pick, int level, int* cid, int* lvl,
                                                               unsigned long long* front_key, int* front_v, unsigned long long* n_front) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (pick[i] < 0) continue;
        const int v = list[i];
        cid[v] = pick[i];
        lvl[v] = level;
        const int s = (int)atomicAdd(n_front, 1ull);
        front_key[s] = key[i];
        front_v[s] = v;
    }
}

// after the level's (key, vertex) pairs were sorted by key: adoption sequence numbers
__global__ void k_fill_level_seq(int n, const int* __restrict__ sorted_v, int base, int* seq) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) seq[sorted_v[i]] = base + i;
}

// ---- sequential replay (connexity guard on, few NULL vertices) ----
// initial FIFO content: every (NULL w, assigned u) edge with its slot; also the capacity the FIFO can need
__global__ void __launch_bounds__(kThreads) k_fill_initial_edges(FillMesh M, int n, const int* __restrict__ list, const int* __restrict__ cid,
                                                                 unsigned* slot_out, int2* edge_out, unsigned long long* counters) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int v = list[i];
        if (cid[v] != M.K) continue;
        const int beg = M.row_ptr[v], end = M.row_ptr[v + 1];
        atomicAdd(counters + 1, (unsigned long long)(end - beg));
        for (int e = beg; e < end; e++) {
            const int u = M.col[e];
            if (cid[u] == M.K) continue;
            const int s = (int)atomicAdd(counters, 1ull);
            slot_out[s] = edge_slot(M, v, u);
            edge_out[s] = make_int2(v, u);
        }
    }
}

__global__ void k_gather_int2(int n, const int* __restrict__ idx, const int2* __restrict__ in, int2* out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = in[idx[i]];
}

// One thread replays the reference's FIFO.  q holds (a, b) vertex pairs: [0, n_init) the slot-sorted initial edges.
// Edges to already assigned neighbours are not pushed (the reference pops and drops them: same outcome).
__global__ void k_fill_sequential(FillMesh M, int n_init, long long cap, int2* q, int* cid, int connexity, unsigned long long* out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    long long head = 0, tail = n_init;
    unsigned long long filled = 0, overflow = 0;
    while (head < tail) {
        const int2 e = q[head++];
        int i1 = e.x, i2 = e.y;
        int c1 = cid[i1], c2 = cid[i2];
        if (c1 == M.K) { const int t = i1; i1 = i2; i2 = t; c1 = c2; c2 = M.K; }
        if (c1 == M.K || c2 != M.K) continue;
        if (connexity && connexity_problem(i2, M.K, M.row_ptr, M.col, cid)) continue;
        cid[i2] = c1;
        filled++;
        unsigned slots[kMaxRing];
        int nbr[kMaxRing];
        int m = 0;
        for (int k = M.row_ptr[i2]; k < M.row_ptr[i2 + 1]; k++) {
            const int u = M.col[k];
            if (cid[u] != M.K) continue;
            if (m == kMaxRing) { overflow = 1; break; }
            const unsigned s = edge_slot(M, i2, u);
            int j = m++;
            while (j > 0 && slots[j - 1] > s) { slots[j] = slots[j - 1]; nbr[j] = nbr[j - 1]; j--; }
            slots[j] = s; nbr[j] = u;
        }
        if (overflow || tail + m > cap) { overflow = 1; break; }
        for (int j = 0; j < m; j++) q[tail++] = make_int2(i2, nbr[j]);
    }
    out[0] = filled;
    out[1] = overflow;
}

// ---- device connexity predicate on caller-given (item, cluster) pairs (parity hook: acvd_connexity_problem) ----
__global__ void __launch_bounds__(kThreads) k_connexity_query(int n, const int* __restrict__ items, const int* __restrict__ clusters,
                                                              const int* __restrict__ row_ptr, const int* __restrict__ col,
                                                              const int* __restrict__ cid, const unsigned long long* __restrict__ ringadj,
                                                              int force_generic, unsigned char* out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int v = items[i], a = clusters[i];
        const int beg = row_ptr[v], deg = row_ptr[v + 1] - beg;
        bool p;
        if (deg <= kRingW && !force_generic) {
            unsigned L = 0;
            for (int k = 0; k < deg; k++) L |= (cid[col[beg + k]] == a ? 1u : 0u) << k;
            p = connexity_problem_ring(L, ringadj[v]);
        } else p = connexity_problem(v, a, row_ptr, col, cid);
        out[i] = p ? 1 : 0;
    }
}

}  // namespace acvd
