// The boundary-item reassignment loop as conflict-free parallel rounds.
//
// Replaces vtkUniformClustering::ProcessOneLoop (reference Common/vtkUniformClustering.h:833-995).
// The reference pops boundary edges from a FIFO and, per edge (I1,I2) with clusters (c1,c2), compares
// E(c1)+E(c2) with the two single-item moves and commits the best immediately (Gauss-Seidel).
// Its per-edge candidate set over all boundary edges is exactly {(v, b): v boundary vertex, b a cluster
// adjacent to v}, so one round here is:
//
//   propose : every boundary vertex whose own or adjacent clusters changed since its last evaluation
//             ("recently modified" rule, :909-920) evaluates all adjacent clusters b:
//                 try = E(a - v) + E(b + v)   vs   cur = E(a) + E(b)          (:922-958)
//             blocked when size(a)==1 or the connexity test fails (:929,:945), or a cluster is frozen
//             (:914-915).  The best strictly improving candidate becomes the vertex's proposal and is
//             submitted with a 64-bit priority key (delta-E, vertex) to both clusters via atomicMin.
//   commit  : a proposal that holds the minimum key on BOTH of its clusters wins; winners touch pairwise
//             disjoint cluster pairs, so their energy deltas are exact and the sums are updated without
//             atomics (:960-991).  Total energy strictly decreases every round.
//
// Vertices in the NULL cluster (id K) adopt an adjacent cluster unconditionally (:881-907).
#pragma once
#include "metric.cuh"

namespace acvd {

struct RoundCounters {
    unsigned long long proposals;   // live proposals submitted this round
    unsigned long long mods;        // committed moves
    unsigned long long tests;       // vertex tests (evaluated candidates incl. blocked)
    unsigned long long evaluated;   // vertices fully evaluated this round (dirty boundary vertices)
    unsigned long long boundary;    // boundary vertices seen
    unsigned long long pad[3];
};

struct ReassignArgs {
    int V, K;
    const int* __restrict__ row_ptr;
    const int* __restrict__ col;
    int* cid;
    const double* __restrict__ items;   // V x stride
    double* csum;                       // K x stride
    double* cenergy;                    // K
    int* csize;                         // K
    int* mod_round;                     // K: last round a cluster was modified
    const unsigned char* __restrict__ frozen;   // K or null
    const int* __restrict__ anchor;     // K or null (QEM fixed clusters)
    const float* __restrict__ xyz;      // V x 3 (anchor coordinates)
    unsigned long long* best;           // K: min priority key per cluster this round
    int* prop_dst;                      // V: proposed destination or -1
    unsigned long long* prop_key;       // V
    double2* prop_e;                    // V: (E(a - v), E(b + v)) of the proposal
    int* plist;                         // compact list of proposing vertices this round
    RoundCounters* ctr;
    int round;
    int force_all;                      // SetAllClustersToModified (:717-722)
    int connexity;
    EvalCfg cfg;
};

// Coordinates of the anchor item of cluster c (QEM fixed clusters, vtkQEMetricForClustering.h:270-275), or null.
template <int EM>
__device__ __forceinline__ const double* anchor_point(const ReassignArgs& A, int c, double* buf) {
    if (EM != M_QEM || !A.anchor) return nullptr;
    int av = A.anchor[c];
    if (av < 0) return nullptr;
    buf[0] = A.xyz[3 * av]; buf[1] = A.xyz[3 * av + 1]; buf[2] = A.xyz[3 * av + 2];
    return buf;
}

// Does removing v from cluster a disconnect v's same-cluster ring neighbours?
// Same predicate as vtkVerticesProcessing::ConnexityConstraintProblemLocal
// (reference DiscreteRemeshing/vtkVerticesProcessing.h:168-237): "L = ring(v) ∩ a is connected in
// the sub-graph induced by L".  More than kMaxRing members => conservatively a problem.
static __device__ __noinline__ bool connexity_problem(int v, int a, const int* __restrict__ row_ptr,
                                               const int* __restrict__ col, const int* cid) {
    int L[kMaxRing];
    int n = 0;
    int beg = row_ptr[v], end = row_ptr[v + 1];
    for (int e = beg; e < end; e++) {
        int u = col[e];
        if (cid[u] == a) {
            if (n == kMaxRing) return true;
            L[n++] = u;
        }
    }
    if (n <= 1) return false;
    unsigned long long reach = 1ull, frontier = 1ull;
    const unsigned long long full = (n == 64) ? ~0ull : ((1ull << n) - 1ull);
    while (frontier) {
        int f = __ffsll((long long)frontier) - 1;
        frontier &= frontier - 1;
        int u = L[f];
        int b2 = row_ptr[u], e2 = row_ptr[u + 1];
        unsigned long long adj = 0;
        for (int e = b2; e < e2; e++) {
            int w = col[e];
            for (int j = 0; j < n; j++)
                if (L[j] == w) adj |= 1ull << j;
        }
        unsigned long long fresh = adj & ~reach;
        reach |= fresh;
        frontier |= fresh;
        if (reach == full) return false;
    }
    return reach != full;
}

// EM: metric whose energy is evaluated; STRIDE: doubles per payload row in memory.
// (QEM's unconstrained phase evaluates the isotropic energy on the first 4 doubles of its rows.)
template <int EM, int STRIDE>
__global__ void __launch_bounds__(kThreads) k_propose(ReassignArgs A) {
    constexpr int NL = MetricTraits<EM>::NPAD;   // doubles loaded per row
    const int K = A.K;
    unsigned n_tests = 0, n_eval = 0, n_bnd = 0;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < A.V; v += gridDim.x * blockDim.x) {
        const int a = A.cid[v];
        const int beg = A.row_ptr[v], end = A.row_ptr[v + 1];
        bool boundary = false;
        const int rm1 = A.round - 1;
        bool a_dirty = A.force_all || (a < K && A.mod_round[a] >= rm1);
        bool dirty = a_dirty;
        for (int e = beg; e < end; e++) {
            int b = A.cid[A.col[e]];
            if (b != a) {
                boundary = true;
                if (b < K && A.mod_round[b] >= rm1) dirty = true;
            }
        }
        if (!boundary) {
            if (a_dirty) A.prop_dst[v] = -1;   // became interior: drop any stale proposal
            continue;
        }
        n_bnd++;
        if (!dirty) {   // clusters unchanged since last evaluation: the stored proposal is still exact
            int d = A.prop_dst[v];
            if (d >= 0) {
                unsigned long long key = A.prop_key[v];
                if (a < K) atomicMin(&A.best[a], key);
                atomicMin(&A.best[d], key);
                int slot = (int)atomicAdd(&A.ctr->proposals, 1ull);
                A.plist[slot] = v;
            }
            continue;
        }
        n_eval++;
        int best_b = -1;
        double best_delta = 0.0, best_ea = 0.0, best_eb = 0.0;
        unsigned long long key = 0;
        if (a >= K) {
            // NULL cluster: adopt the first assigned, non-frozen neighbour cluster; top priority
            for (int e = beg; e < end && best_b < 0; e++) {
                int b = A.cid[A.col[e]];
                if (b < K && !(A.frozen && A.frozen[b])) best_b = b;
            }
            key = (unsigned long long)(unsigned)v;
        } else if (!(A.frozen && A.frozen[a])) {
            bool blocked = (A.csize[a] == 1) || (A.anchor && A.anchor[a] == v);
            if (!blocked && A.connexity) blocked = connexity_problem(v, a, A.row_ptr, A.col, A.cid);
            double it[NL], s[NL];
            double ea_new = 0.0;
            const double cur_a = A.cenergy[a];
            double anchor_pt[3];
            if (!blocked) {
                load_row_ro<NL>(A.items + (int64_t)v * STRIDE, it);
                load_row<NL>(A.csum + (int64_t)a * STRIDE, s);
#pragma unroll
                for (int i = 0; i < NL; i++) s[i] -= it[i];
                ea_new = cluster_energy<EM>(s, A.cfg, nullptr, anchor_point<EM>(A, a, anchor_pt));
            }
            for (int e = beg; e < end; e++) {
                int b = A.cid[A.col[e]];
                if (b == a || b >= K) continue;
                bool seen = false;
                for (int e2 = beg; e2 < e; e2++) seen |= (A.cid[A.col[e2]] == b);
                if (seen) continue;
                if (A.frozen && A.frozen[b]) continue;
                n_tests++;
                if (blocked) continue;
                const double2* ps = reinterpret_cast<const double2*>(A.csum + (int64_t)b * STRIDE);
#pragma unroll
                for (int i = 0; i < NL / 2; i++) { double2 u = ps[i]; s[2 * i] = u.x + it[2 * i]; s[2 * i + 1] = u.y + it[2 * i + 1]; }
                double eb_new = cluster_energy<EM>(s, A.cfg, nullptr, anchor_point<EM>(A, b, anchor_pt));
                double tr = ea_new + eb_new;
                double cur = cur_a + A.cenergy[b];
                if (tr < cur) {
                    double delta = tr - cur;
                    if (best_b < 0 || delta < best_delta) { best_b = b; best_delta = delta; best_ea = ea_new; best_eb = eb_new; }
                }
            }
            if (best_b >= 0)
                key = ((unsigned long long)ordered_float_bits(__double2float_rn(best_delta)) << 32) | (unsigned)v;
        }
        A.prop_dst[v] = best_b;
        if (best_b >= 0) {
            A.prop_key[v] = key;
            A.prop_e[v] = make_double2(best_ea, best_eb);
            if (a < K) atomicMin(&A.best[a], key);
            atomicMin(&A.best[best_b], key);
            int slot = (int)atomicAdd(&A.ctr->proposals, 1ull);
            A.plist[slot] = v;
        }
    }
    warp_count_add(&A.ctr->tests, n_tests);
    warp_count_add(&A.ctr->evaluated, n_eval);
    warp_count_add(&A.ctr->boundary, n_bnd);
}

// UM: metric of the stored sums (all UM::NPAD doubles of a row are updated);
// EM: metric used to re-evaluate an adopting cluster's energy.
template <int EM, int UM>
__global__ void __launch_bounds__(kThreads) k_commit(ReassignArgs A) {
    constexpr int NU = MetricTraits<UM>::NPAD;
    const int K = A.K;
    const int n_props = (int)A.ctr->proposals;   // written by k_propose of this round
    unsigned n_mods = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_props; i += gridDim.x * blockDim.x) {
        const int v = A.plist[i];
        const int d = A.prop_dst[v];
        const unsigned long long key = A.prop_key[v];
        const int a = A.cid[v];
        bool win = (A.best[d] == key) && (a >= K || A.best[a] == key);
        if (!win) continue;
        double it[NU], s[NU];
        load_row_ro<NU>(A.items + (int64_t)v * NU, it);
        // destination += item
        load_row<NU>(A.csum + (int64_t)d * NU, s);
#pragma unroll
        for (int k = 0; k < NU; k++) s[k] += it[k];
        store_row<NU>(A.csum + (int64_t)d * NU, s);
        double2 pe = A.prop_e[v];
        if (a >= K) {
            // adoption: energy of the grown cluster evaluated here (:884-886)
            double anchor_pt[3];
            A.cenergy[d] = cluster_energy<EM>(s, A.cfg, nullptr, anchor_point<EM>(A, d, anchor_pt));
        } else {
            A.cenergy[d] = pe.y;
            load_row<NU>(A.csum + (int64_t)a * NU, s);
#pragma unroll
            for (int k = 0; k < NU; k++) s[k] -= it[k];
            store_row<NU>(A.csum + (int64_t)a * NU, s);
            A.cenergy[a] = pe.x;
            A.csize[a] -= 1;
            A.mod_round[a] = A.round;
        }
        A.csize[d] += 1;
        A.mod_round[d] = A.round;
        A.cid[v] = d;
        A.prop_dst[v] = -1;
        n_mods++;
    }
    warp_count_add(&A.ctr->mods, n_mods);
}

}  // namespace acvd
