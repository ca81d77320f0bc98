// Cluster-centric side of the reassignment loop (sm_100a):
//
//  * per-cluster member arrays -- built by counting (no sort), kept up to date by the commits of the exact rounds;
//  * k_cluster_pass: ONE pass over the clusters (a warp per cluster, members staged in shared memory) that sorts the
//    members (rank sort: the summation order is the reference's item order and does not depend on atomics), checks the
//    cluster's connectivity (CleanClustering, Common/vtkUniformClustering.h:406-549: union-find on the members' local
//    indices in shared memory), and accumulates the statistics (ReComputeStatistics, :376-403: warp-shuffle reduction of
//    the payload rows) -- what used to be a global radix sort + four union-find sweeps over the whole mesh + a
//    statistics kernel per convergence event;
//  * sparse rounds: once a phase has left its opening round, the vertices that must be (re)evaluated are exactly the
//    boundary members of the clusters modified in the previous round and their foreign neighbours ("recently modified"
//    rule, :909-920).  They are enumerated from the member arrays of the modified clusters -- the work of a round is
//    proportional to what changed, not to the mesh -- and the whole tail of a phase runs inside ONE persistent
//    cooperative kernel (k_sparse_rounds: enumerate -> evaluate -> select/commit passes, grid-wide barriers in
//    between, convergence decided on the device), instead of ~10 launches and a host poll per round.
//
// The decisions are those of the tile-filter path (k_tile_filter / k_scan / k_carry in reassign.cuh, still used for the
// opening round of every phase and as the A/B partner, ACVD_NO_SPARSE=1): same dirty set, same keys, same winners.
#pragma once
#include <cooperative_groups.h>

#include "mesh.cuh"
#include "reassign.cuh"

namespace cg = cooperative_groups;

namespace acvd {

constexpr int kMaxPasses = 8;          // select + commit passes of one sparse round (upper bound of acvd_params.commit_passes)
constexpr int kPushBuf = 256;          // per-warp staging of work-list entries (one global atomic per ~224 entries)
constexpr int kClusterCap = 384;       // members a cluster may have for its pass to run out of shared memory

// ---------------------------------------------------------------------------------------------------
// member arrays by counting
__global__ void __launch_bounds__(kThreads) k_members_count(int V, int K, const int* __restrict__ cid, int* cnt) {
    const int lane = threadIdx.x & 31;
    for (int v0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; v0 < V; v0 += gridDim.x * blockDim.x) {
        const int v = v0 + lane;
        int c = v < V ? cid[v] : K;
        if (c < 0 || c > K) c = K;
        const unsigned peers = __match_any_sync(0xffffffffu, c);      // consecutive vertices mostly share a cluster
        if (c < K && lane == __ffs(peers) - 1) atomicAdd(cnt + c, __popc(peers));
    }
}
// capacity = twice the size (at least size + 16): a cluster whose array fills up ends the sparse launch and has the arrays
// rebuilt; with half the size as slack the anisotropic configurations (C3) relaunched every ~17 rounds
__global__ void k_members_cap(int K, const int* cnt, int* cap) {   // in place
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c <= K; c += gridDim.x * blockDim.x)
        cap[c] = c < K ? cnt[c] + max(16, cnt[c]) : 0;
}
// fills the arrays (order inside a cluster depends on the atomics: k_cluster_pass sorts it) and leaves the sizes in csize
__global__ void __launch_bounds__(kThreads) k_members_scatter(int V, int K, const int* __restrict__ cid, const int* __restrict__ off,
                                                              int* csize, int* memb, int* pos) {
    const int lane = threadIdx.x & 31;
    for (int v0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; v0 < V; v0 += gridDim.x * blockDim.x) {
        const int v = v0 + lane;
        int c = v < V ? cid[v] : K;
        if (c < 0 || c > K) c = K;
        const unsigned peers = __match_any_sync(0xffffffffu, c);
        const int leader = __ffs(peers) - 1;
        int base = 0;
        if (c < K && lane == leader) base = atomicAdd(csize + c, __popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (c < K) {
            const int slot = off[c] + base + __popc(peers & ((1u << lane) - 1u));
            memb[slot] = v;
            pos[v] = slot;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// union-find on local member indices; `par` may live in shared or global memory (generic pointer)
__device__ __forceinline__ int lcc_find(volatile int* par, int x) {
    while (true) {                           // path halving: a shortcut always points at an ancestor, so racing lanes stay correct
        const int p = par[x];
        if (p == x) return x;
        const int gp = par[p];
        if (gp == p) return p;
        par[x] = gp;
        x = gp;
    }
}
__device__ __forceinline__ void lcc_union(int* par, int x, int y) {
    int rx = lcc_find(par, x), ry = lcc_find(par, y);
    while (rx != ry) {                       // hook the larger root under the smaller one (a failed CAS walks on)
        if (ry < rx) { const int t = rx; rx = ry; ry = t; }
        const int old = atomicCAS(par + ry, ry, rx);
        if (old == ry) return;
        ry = lcc_find(par, old);
        rx = lcc_find(par, rx);
    }
}

struct ClusterPassArgs {
    int V, K;
    const int* __restrict__ off;
    int* memb;
    int* memb_tmp;            // scratch of the same size (clusters above kClusterCap)
    int* pos;
    int* cid;
    const int* __restrict__ csize;
    const int* __restrict__ row_ptr;
    const int* __restrict__ col;
    const double* __restrict__ items;
    double* csum;
    double* cenergy;
    double* ccentroid;
    const int* __restrict__ anchor;
    const float* __restrict__ xyz;
    int* cc_par;              // scratch, one int per member slot (clusters above kClusterCap)
    int* cc_sz;
    unsigned long long* counters;   // [0] clusters with more than one recorded component, [1] vertices reset to NULL
    int do_sort, do_cc, do_stats;
    int apply_resets;         // 0: the connectivity check only counts (multi-GPU: every rank checks its share of the clusters first)
    int k_begin, k_end;       // clusters handled by this launch
    const int* __restrict__ mod_round;   // last round every cluster was modified in
    int cc_since;             // connectivity is checked for clusters modified in round >= cc_since (the others were
                              // connected at the last check and have not changed since)
    EvalCfg cfg;
};

// One warp per cluster.  M: metric of the stored rows; EM: metric whose energy formula is evaluated (QEM's
// unconstrained phase uses the isotropic one).
template <int M, int EM>
__global__ void __launch_bounds__(kThreads) k_cluster_pass(ClusterPassArgs P) {
    constexpr int NPAD = MetricTraits<M>::NPAD;
    __shared__ int s_v[kThreads / 32][kClusterCap];
    __shared__ int s_a[kThreads / 32][kClusterCap];
    __shared__ int s_b[kThreads / 32][kClusterCap];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int warps_per_block = blockDim.x >> 5;
    for (int c = P.k_begin + blockIdx.x * warps_per_block + w; c < P.k_end; c += gridDim.x * warps_per_block) {
        const int b = P.off[c], n = P.csize[c];
        const bool small = n <= kClusterCap;
        int* mv = small ? s_v[w] : P.memb + b;          // the cluster's members, ascending after the sort
        // ---- 1. rank sort of the members (ids are distinct: the ranks are a permutation)
        if (P.do_sort) {
            int* src = small ? s_a[w] : P.memb + b;
            int* dst = small ? s_v[w] : P.memb_tmp + b;
            if (small) { for (int i = lane; i < n; i += 32) src[i] = P.memb[b + i]; __syncwarp(); }
            for (int i = lane; i < n; i += 32) {
                const int x = src[i];
                int r = 0;
                for (int j = 0; j < n; j++) r += (src[j] < x) ? 1 : 0;
                dst[r] = x;
            }
            __syncwarp();
            for (int i = lane; i < n; i += 32) { const int x = dst[i]; P.memb[b + i] = x; P.pos[x] = b + i; }
            __syncwarp();
        } else if (small) {
            for (int i = lane; i < n; i += 32) mv[i] = P.memb[b + i];
            __syncwarp();
        }
        // ---- 2. connected components of the cluster (root = smallest member = the vertex at which the reference's
        //         index-ordered BFS discovers the component, :428-437)
        if (P.do_cc && n > 1 && P.mod_round[c] >= P.cc_since) {
            int* par = small ? s_a[w] : P.cc_par + b;
            int* sz = small ? s_b[w] : P.cc_sz + b;
            // (a) initial forest without atomics: every member points at its smallest same-cluster neighbour with a
            //     smaller id, or at itself.  Every link is a real edge, so a forest with a single root spans the cluster:
            //     connected, nothing else to do -- the common case.
            int n_local_min = 0;
            for (int i0 = 0; i0 < n; i0 += 32) {
                const int i = i0 + lane;
                int best = i;
                if (i < n) {
                    const int v = mv[i];
                    const int beg = P.row_ptr[v], deg = P.row_ptr[v + 1] - beg;
                    int nb[kRingW], pu[kRingW];
#pragma unroll
                    for (int k = 0; k < kRingW; k++) { const int u = k < deg ? P.col[beg + k] : v; nb[k] = u < v ? u : -1; }
#pragma unroll
                    for (int k = 0; k < kRingW; k++) pu[k] = (nb[k] >= 0 && P.cid[nb[k]] == c) ? nb[k] : -1;
#pragma unroll
                    for (int k = 0; k < kRingW; k++) pu[k] = pu[k] >= 0 ? P.pos[pu[k]] - b : i;
#pragma unroll
                    for (int k = 0; k < kRingW; k++) best = min(best, pu[k]);
                    for (int e = beg + kRingW; e < beg + deg; e++) {       // rows longer than kRingW
                        const int u = P.col[e];
                        if (u < v && P.cid[u] == c) best = min(best, P.pos[u] - b);
                    }
                    par[i] = best;
                }
                n_local_min += __popc(__ballot_sync(0xffffffffu, i < n && best == i));
            }
            __syncwarp();
            // (b) several local minima: join the trees over all same-cluster edges (union-find on the local indices)
            if (n_local_min > 1) {
                for (int i = lane; i < n; i += 32) {
                    const int v = mv[i];
                    const int beg = P.row_ptr[v], deg = P.row_ptr[v + 1] - beg;
                    int nb[kRingW], pu[kRingW];
#pragma unroll
                    for (int k = 0; k < kRingW; k++) { const int u = k < deg ? P.col[beg + k] : v; nb[k] = u < v ? u : -1; }
#pragma unroll
                    for (int k = 0; k < kRingW; k++) pu[k] = (nb[k] >= 0 && P.cid[nb[k]] == c) ? nb[k] : -1;
#pragma unroll
                    for (int k = 0; k < kRingW; k++) pu[k] = pu[k] >= 0 ? P.pos[pu[k]] - b : -1;
#pragma unroll 1
                    for (int k = 0; k < kRingW; k++) if (pu[k] >= 0) lcc_union(par, i, pu[k]);
                    for (int e = beg + kRingW; e < beg + deg; e++) {
                        const int u = P.col[e];
                        if (u < v && P.cid[u] == c) lcc_union(par, i, P.pos[u] - b);
                    }
                }
                __syncwarp();
            }
            int n_roots = 1;
            if (n_local_min > 1) {
                n_roots = 0;
                for (int i0 = 0; i0 < n; i0 += 32) {
                    const int i = i0 + lane;
                    int r = -1;
                    if (i < n) r = lcc_find(par, i);
                    __syncwarp();
                    if (i < n) { par[i] = r; sz[i] = 0; }
                    n_roots += __popc(__ballot_sync(0xffffffffu, i < n && r == i));
                }
                __syncwarp();
            }
            if (n_roots > 1) {
                // sizes; an anchored item weighs 1e9 so that its component always wins (:440-447)
                const int anchored = P.anchor ? P.anchor[c] : -1;
                for (int i = lane; i < n; i += 32) atomicAdd(sz + par[i], mv[i] == anchored ? 1000000000 : 1);
                __syncwarp();
                // winner = largest recorded component, first discovered wins ties (:503).  Reference quirk kept: the
                // component discovered at item 0 is never recorded (0 doubles as the "unvisited" sentinel, :463-467),
                // so it is neither counted nor ever reset.
                unsigned long long best = 0;
                int n_comp = 0;
                for (int i0 = 0; i0 < n; i0 += 32) {
                    const int i = i0 + lane;
                    const bool rec = i < n && par[i] == i && mv[i] != 0;
                    if (rec) {
                        const unsigned long long key = ((unsigned long long)(unsigned)sz[i] << 32) | (unsigned)(0xffffffffu - (unsigned)i);
                        best = key > best ? key : best;
                    }
                    n_comp += __popc(__ballot_sync(0xffffffffu, rec));
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { const unsigned long long x = __shfl_xor_sync(0xffffffffu, best, o); best = x > best ? x : best; }
                if (n_comp >= 2) {
                    const int win = (int)(0xffffffffu - (unsigned)(best & 0xffffffffull));
                    unsigned n_reset = 0;
                    for (int i = lane; i < n; i += 32) {
                        const int r = par[i];
                        if (mv[r] != 0 && r != win) { if (P.apply_resets) P.cid[mv[i]] = P.K; n_reset++; }
                    }
                    n_reset = __reduce_add_sync(0xffffffffu, n_reset);
                    if (lane == 0) { atomicAdd(P.counters, 1ull); atomicAdd(P.counters + 1, (unsigned long long)n_reset); }
                }
            }
            __syncwarp();
        }
        // ---- 3. statistics: rows summed in ascending item order, lane-strided partial sums + shuffle tree
        if (P.do_stats) {
            double acc[NPAD];
#pragma unroll
            for (int k = 0; k < NPAD; k++) acc[k] = 0.0;
            for (int i = lane; i < n; i += 32) {
                double it[NPAD];
                load_row_ro<NPAD>(P.items + (int64_t)mv[i] * NPAD, it);
#pragma unroll
                for (int k = 0; k < NPAD; k++) acc[k] += it[k];
            }
#pragma unroll
            for (int k = 0; k < NPAD; k++) acc[k] = warp_sum(acc[k]);
            if (lane == 0) {
                store_row<NPAD>(P.csum + (int64_t)c * NPAD, acc);
                double cen[3], apt[3];
                const double* ap = nullptr;
                if (EM == M_QEM && P.anchor && P.anchor[c] >= 0) {
                    const int av = P.anchor[c];
                    apt[0] = P.xyz[3 * (int64_t)av]; apt[1] = P.xyz[3 * (int64_t)av + 1]; apt[2] = P.xyz[3 * (int64_t)av + 2];
                    ap = apt;
                }
                P.cenergy[c] = cluster_energy<EM>(acc, P.cfg, cen, ap);
                P.ccentroid[3 * c] = cen[0]; P.ccentroid[3 * c + 1] = cen[1]; P.ccentroid[3 * c + 2] = cen[2];
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------
// Per-cluster quadrics of the ACVD post-process (reference DiscreteRemeshing/Examples/ACVD.cxx:237-262): for every item
// of the cluster, the quadric of every input face around it (vtkQuadricTools::AddTriangleQuadric, first 9 coefficients).
// A warp per cluster over its (sorted) members, faces from the vertex -> face incidence, warp-shuffle reduction.
__global__ void __launch_bounds__(kThreads) k_cluster_quadrics(int n_clusters, const int* __restrict__ off, const int* __restrict__ memb,
                                                               const int* __restrict__ csize, const int* __restrict__ vf_ptr,
                                                               const unsigned long long* __restrict__ vf_keys, const float* __restrict__ xyz,
                                                               const int* __restrict__ tri, double* Q9) {
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    for (int c = blockIdx.x * warps_per_block + (threadIdx.x >> 5); c < n_clusters; c += gridDim.x * warps_per_block) {
        const int b = off[c], n = csize[c];
        double q[9];
#pragma unroll
        for (int k = 0; k < 9; k++) q[k] = 0.0;
        for (int i = lane; i < n; i += 32) {
            const int v = memb[b + i];
            for (int j = vf_ptr[v]; j < vf_ptr[v + 1]; j++) {
                Tri3 t;
                load_face(xyz, tri, (int)(vf_keys[j] & 0xffffffffull), t);
                tri_quadric_add(t, q);
            }
        }
#pragma unroll
        for (int k = 0; k < 9; k++) q[k] = warp_sum(q[k]);
        if (lane == 0)
#pragma unroll
            for (int k = 0; k < 9; k++) Q9[9 * (int64_t)c + k] = q[k];
    }
}

// ---------------------------------------------------------------------------------------------------
// sparse rounds
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// warp-staged append to the work list: entries collect in shared memory, one global atomic per flush
struct WarpPush {
    int* buf;          // kPushBuf ints of this warp
    int cnt;           // warp-uniform
    int* work;
    unsigned long long* counter;
    __device__ __forceinline__ void flush(int lane) {
        if (cnt == 0) return;
        int base = 0;
        if (lane == 0) base = (int)atomicAdd(counter, (unsigned long long)cnt);
        base = __shfl_sync(0xffffffffu, base, 0);
        __syncwarp();
        for (int i = lane; i < cnt; i += 32) work[base + i] = buf[i];
        __syncwarp();
        cnt = 0;
    }
    // every lane of the warp calls this (pred false: nothing to add)
    __device__ __forceinline__ void push(bool pred, int val, int lane) {
        const unsigned m = __ballot_sync(0xffffffffu, pred);
        if (pred) buf[cnt + __popc(m & ((1u << lane) - 1u))] = val;
        cnt += __popc(m);
        if (cnt > kPushBuf - 32) flush(lane);
    }
};

// The dirty set of round A.round from the clusters modified in the previous round: their boundary members and the
// foreign neighbours of those (exactly the boundary vertices whose own or an adjacent cluster was modified).  A warp
// per modified cluster; stamp[] keeps a vertex from entering the list twice.
constexpr int kEnumChunks = 4;          // 32-member chunks of one modified cluster handled by different warps

// (modin: written by the previous round of the same launch, so no read-only path)
// One 32-member batch of a modified cluster: the stage-by-stage gathers are written for TWO batches side by side (the two
// tasks a warp takes per iteration), so that the eight dependent levels of a batch (cluster id -> offsets -> member ->
// row -> neighbours -> their clusters -> their stamps -> claims) are paid once per pair: the step is bound by that chain.
struct EnumBatch { int c, b, n, i0; bool on; };

static __device__ __noinline__ void enumerate_modified(const ReassignArgs& A, const int* modin, int n_modin, int* s_buf) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    WarpPush wp{s_buf + (threadIdx.x >> 5) * kPushBuf, 0, A.work, &A.ctr->evaluated};
    const int round = A.round;
    unsigned n_members = 0;
    const int n_tasks = n_modin * kEnumChunks;

    // gathers of two batches, level by level; the claims (atomicExch on the stamps) and the list appends one batch after the other
    auto run_pair = [&](const EnumBatch (&B)[2]) {
        int v[2], beg[2], deg[2];
        bool live[2];
#pragma unroll
        for (int t = 0; t < 2; t++) {
            live[t] = B[t].on && B[t].i0 + lane < B[t].n;
            v[t] = live[t] ? A.mem.memb[B[t].b + B[t].i0 + lane] : 0;
        }
#pragma unroll
        for (int t = 0; t < 2; t++) { beg[t] = live[t] ? A.row_ptr[v[t]] : 0; deg[t] = live[t] ? A.row_ptr[v[t] + 1] - beg[t] : 0; }
        int nb[2][kRingW], cu[2][kRingW], st[2][kRingW];
#pragma unroll
        for (int t = 0; t < 2; t++)
#pragma unroll
            for (int k = 0; k < kRingW; k++) nb[t][k] = k < deg[t] ? A.col[beg[t] + k] : -1;
#pragma unroll
        for (int t = 0; t < 2; t++)
#pragma unroll
            for (int k = 0; k < kRingW; k++) cu[t][k] = nb[t][k] >= 0 ? A.cid[nb[t][k]] : B[t].c;
#pragma unroll
        for (int t = 0; t < 2; t++)
#pragma unroll
            for (int k = 0; k < kRingW; k++) st[t][k] = cu[t][k] != B[t].c ? A.stamp[nb[t][k]] : round;
#pragma unroll
        for (int t = 0; t < 2; t++) {
            if (!B[t].on) continue;                      // (warp-uniform)
            const int c = B[t].c;
            bool bnd = false;
#pragma unroll
            for (int k = 0; k < kRingW; k++) {
                const bool foreign = cu[t][k] != c;
                bnd |= foreign;
                const bool add = foreign && st[t][k] != round && atomicExch(A.stamp + nb[t][k], round) != round;
                wp.push(add, nb[t][k], lane);
            }
            const int extra = __reduce_max_sync(0xffffffffu, deg[t]) - kRingW;     // rows longer than kRingW: slot by slot
            for (int k = 0; k < extra; k++) {
                bool add = false;
                int u = 0;
                if (kRingW + k < deg[t]) {
                    u = A.col[beg[t] + kRingW + k];
                    if (A.cid[u] != c) {
                        bnd = true;
                        add = A.stamp[u] != round && atomicExch(A.stamp + u, round) != round;
                    }
                }
                wp.push(add, u, lane);
            }
            const bool addv = bnd && atomicExch(A.stamp + v[t], round) != round;
            wp.push(addv, v[t], lane);
            // a member that is no boundary vertex any more (its last foreign neighbour joined the cluster) drops the
            // proposal it may still hold, as it does on the tile-filter path: it is not carried over
            if (live[t] && !bnd) A.stamp[v[t]] = round;
            n_members += live[t] ? 1u : 0u;
        }
    };

    for (int task0 = warp; task0 < n_tasks; task0 += 2 * n_warps) {
        EnumBatch B[2];
        int ct[2];
#pragma unroll
        for (int t = 0; t < 2; t++) {
            const int task = task0 + t * n_warps;
            B[t].on = task < n_tasks;
            ct[t] = B[t].on ? modin[task / kEnumChunks] : 0;
            B[t].i0 = B[t].on ? (task % kEnumChunks) * 32 : 0;
        }
#pragma unroll
        for (int t = 0; t < 2; t++) { B[t].c = ct[t]; B[t].b = A.mem.off[ct[t]]; B[t].n = B[t].on ? A.csize[ct[t]] : 0; }
        run_pair(B);
        // clusters with more than kEnumChunks x 32 members (rare): the further batches of the two tasks
        while (true) {
            bool more = false;
#pragma unroll
            for (int t = 0; t < 2; t++) { B[t].i0 += kEnumChunks * 32; B[t].on = B[t].on && B[t].i0 < B[t].n; more |= B[t].on; }
            if (!more) break;
            run_pair(B);
        }
    }
    wp.flush(lane);
    warp_count_add(&A.ctr->pad[1], n_members);       // members visited (bytes model of the round)
}

// live proposals of the previous round whose vertex is not re-evaluated this round compete again with their stored key
__device__ __forceinline__ void carry_sparse(const ReassignArgs& A, int n_prev) {
    const int K = A.K;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_prev; i += gridDim.x * blockDim.x) {
        const int v = A.plist_prev[i];
        const int d = A.prop_dst[v];
        if (d < 0 || A.stamp[v] == A.round) continue;         // committed last round / on this round's work list
        const int a = A.cid[v];
        const unsigned long long key = A.prop_key[v];
        if (a < K) atomicMin(&A.best[a], key);
        atomicMin(&A.best[d], key);
        const int slot = (int)atomicAdd(&A.ctr->proposals, 1ull);
        A.plist[slot] = v;
    }
}

struct SparseCtl {
    RoundCounters* rc;                 // [max_rounds], zeroed by the host: counters of every round of the launch
    unsigned long long* n_mod;         // [max_rounds + 1]: n_mod[i] = clusters modified by round i - 1 ([0]: before the launch)
    unsigned long long* resub;         // [max_rounds * kMaxPasses], zeroed: proposals resubmitted to pass p of round i
    unsigned long long* ts;            // [max_rounds * 4]: %globaltimer after enumerate / evaluate / commits of every round
    int* modlist0; int* modlist1;
    int* plist0; int* plist1;
    unsigned long long* best0; unsigned long long* best1;
    unsigned long long n_prev_props;   // proposals alive before the first round (in the "previous" list of that round)
    int par0;                          // parity (list selector) of the first round
    int max_rounds, passes, has_long_rows;
    long long stop_props;              // convergence event on "live proposals <= stop_props" (first two phases), -1 = off
    int* done;                         // [0] rounds executed
    long long leave_below, leave_above;   // the launch ends after a round that evaluated <= / > this many vertices (-1: off): the
                                          // host continues with the launch shape that fits the work (grid <-> one cluster)
};

// Barrier between the steps of a round.  CL = false: cooperative launch over the whole device (grid barrier).
// CL = true: the launch is ONE thread-block cluster -- the long tail of a phase moves a few hundred vertices per round,
// and what a round costs there is its five or six device-wide barriers, not its work; the hardware cluster barrier
// costs a fraction of a grid barrier.
template <bool CL>
__device__ __forceinline__ void rounds_barrier() {
    if constexpr (CL) {
        __threadfence();                              // the steps exchange data through global memory
        cg::this_cluster().sync();
    } else cg::this_grid().sync();
}

// EM / STRIDE / UM as in k_evaluate / k_commit.  CL = false: cooperative launch, one resident wave of blocks;
// CL = true: one cluster of blocks.  The key tables (best0 / best1: minimum priority key per cluster) must read
// "no key" where a round looks: the grid form clears them wholesale while it enumerates (K entries spread over ~75 k
// threads); the cluster form resets exactly the entries the previous round touched -- the clusters of its proposals
// and the clusters its commits modified -- and expects clean tables at launch (host memset).
template <int EM, int STRIDE, int UM, bool CL>
__global__ void __launch_bounds__(kThreads, 2) k_sparse_rounds(const __grid_constant__ ReassignArgs A0, const __grid_constant__ SparseCtl S) {
    __shared__ int s_buf[(kThreads / 32) * kPushBuf];
    __shared__ ReassignArgs sA[2];      // the round's arguments (per-round fields patched in), double-buffered by round parity
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, n_threads = gridDim.x * blockDim.x;
    const int K = A0.K;
    int it = 0;
    while (it < S.max_rounds) {
        const int par = (S.par0 + it) & 1;
        ReassignArgs& A = sA[it & 1];
        if (threadIdx.x == 0) {
            A = A0;
            A.round = A0.round + it;
            A.ctr = S.rc + it;
            A.plist = par ? S.plist1 : S.plist0;
            A.plist_prev = par ? S.plist0 : S.plist1;
            A.modlist = par ? S.modlist0 : S.modlist1;
            A.n_mod = S.n_mod + it + 1;
            A.best = S.best0;
        }
        __syncthreads();
        const int* modin = par ? S.modlist1 : S.modlist0;
        const int n_modin = (int)S.n_mod[it];
        const int n_prev = it == 0 ? (int)S.n_prev_props : (int)S.rc[it - 1].proposals;
        // ---- enumerate the dirty set (and clear the key table of the first pass)
        if constexpr (CL) {
            if (it > 0) {       // the entries the previous round touched: its surviving proposals' clusters + the clusters it modified
                const ReassignArgs& P = sA[(it - 1) & 1];
                for (int i = tid; i < n_prev; i += n_threads) {
                    const int v = P.plist[i];
                    const int d = P.prop_dst[v];
                    if (d < 0) continue;
                    const int a = P.cid[v];
                    if (a < K) { S.best0[a] = ~0ull; S.best1[a] = ~0ull; }
                    S.best0[d] = ~0ull; S.best1[d] = ~0ull;
                }
                for (int i = tid; i < n_modin; i += n_threads) { const int c = modin[i]; S.best0[c] = ~0ull; S.best1[c] = ~0ull; }
            }
        } else {
            for (int i = tid; i < K; i += n_threads) S.best0[i] = ~0ull;
        }
        enumerate_modified(A, modin, n_modin, s_buf);
        rounds_barrier<CL>();
        if (tid == 0) S.ts[4 * it] = global_timer_ns();
        // ---- evaluate it; untouched live proposals compete again
        const int n_work = (int)A.ctr->evaluated;
        carry_sparse(A, n_prev);
        evaluate_list<EM, STRIDE>(A, n_work);
        if (S.has_long_rows) evaluate_long_list<EM, STRIDE>(A, n_work);
        rounds_barrier<CL>();
        if (tid == 0) S.ts[4 * it + 1] = global_timer_ns();
        // ---- select + commit passes: a pass is skipped (and the round ends) when nothing could be resubmitted to it
        const int n_props = (int)A.ctr->proposals;
        for (int pass = 0; pass < S.passes; pass++) {
            if (threadIdx.x == 0) A.best = (pass & 1) ? S.best1 : S.best0;      // (the previous use of A ended at a barrier)
            __syncthreads();
            if (pass > 0) {
                resubmit_list(A, n_props, S.resub + it * kMaxPasses + pass);
                rounds_barrier<CL>();
                if (S.resub[it * kMaxPasses + pass] == 0) break;
            }
            if (!CL && pass + 1 < S.passes) {                // the next pass's key table is idle during this commit
                unsigned long long* other = (pass & 1) ? S.best0 : S.best1;
                for (int i = tid; i < K; i += n_threads) other[i] = ~0ull;
            }
            commit_list<EM, UM>(A, n_props);
            rounds_barrier<CL>();
        }
        if (tid == 0) S.ts[4 * it + 2] = global_timer_ns();
        const unsigned long long mods = A.ctr->mods;
        it++;
        if (mods == 0 || (S.stop_props >= 0 && (long long)n_props <= S.stop_props) || (A.mem.overflow && *A.mem.overflow)) break;
        if ((S.leave_below >= 0 && n_work <= S.leave_below) || (S.leave_above >= 0 && n_work > S.leave_above)) break;
    }
    if (tid == 0) S.done[0] = it;
}

}  // namespace acvd
