// The boundary-item reassignment loop as conflict-free parallel rounds.
//
// Replaces vtkUniformClustering::ProcessOneLoop (reference Common/vtkUniformClustering.h:833-995).
// The reference pops boundary edges from a FIFO and, per edge (I1,I2) with clusters (c1,c2), compares
// E(c1)+E(c2) with the two single-item moves and commits the best immediately (Gauss-Seidel).
// Its per-edge candidate set over all boundary edges is exactly {(v, b): v boundary vertex, b a cluster
// adjacent to v}, so one round here is:
//
//   scan    : a coalesced streaming pass over the CSR finds the boundary vertices whose own or adjacent
//             clusters changed since their last evaluation ("recently modified" rule, :909-920) and
//             compacts them into a work list;
//   evaluate: every work-list vertex evaluates all adjacent clusters b:
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
#include "reassign_types.cuh"

namespace acvd {

// A vertex that changes cluster changes the signature of its own tile and of its neighbours' tiles.
__device__ __forceinline__ void mark_tiles_stale(const ReassignArgs& A, int v) {
    if (!A.track_stale) return;          // dense mode: the signatures are invalid as a whole and will be rebuilt for every tile
    A.tile_stale[v >> 5] = 1;
    for (int e = A.row_ptr[v]; e < A.row_ptr[v + 1]; e++) A.tile_stale[A.col[e] >> 5] = 1;
}

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

// ---------------------------------------------------------------------------------------------------
// k_modbits: one bit per cluster, set when the cluster was modified in the previous round (or force_all).
// K/8 bytes (50 KB at K = 400k): the scan's "recently modified" look-ups hit L1 instead of gathering
// 4-byte stamps from L2.
// When `ctr` is given the kernel also opens the round's counters (one launch instead of three small memsets / copies):
// the proposals of the previous round become the carry count (prev_mode 1) or the carry list is dropped (2), the
// active-tile count and the round counters are zeroed.
__global__ void __launch_bounds__(kThreads) k_modbits(int K, const int* __restrict__ mod_round, int rm1, int force_all,
                                                      unsigned* __restrict__ bits, RoundCounters* ctr = nullptr,
                                                      unsigned long long* round_scalars = nullptr, int prev_mode = 0,
                                                      const int* __restrict__ csize = nullptr, int* __restrict__ cmeta = nullptr) {
    if (ctr && blockIdx.x == 0 && threadIdx.x == 0) {
        if (prev_mode == 1) round_scalars[1] = ctr->proposals;
        else if (prev_mode == 2) round_scalars[1] = 0;
        round_scalars[0] = 0;
        ctr->proposals = 0; ctr->mods = 0; ctr->tests = 0; ctr->evaluated = 0; ctr->boundary = 0;
        ctr->pad[0] = 0; ctr->pad[1] = 0; ctr->pad[2] = 0;
    }
    const int n_words = (K + 31) >> 5;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_words * 32; c += gridDim.x * blockDim.x) {
        bool m = c < K && (force_all || mod_round[c] >= rm1);
        unsigned w = __ballot_sync(0xffffffffu, m);
        if ((threadIdx.x & 31) == 0) bits[c >> 5] = w;
        // size and modified flag in one word: one load per cluster in the dense bulk scan (entry K = the NULL cluster)
        if (cmeta && c <= K) cmeta[c] = c < K ? (csize[c] | (m ? (int)0x80000000 : 0)) : 0;
    }
}

// The frontier scan works on tiles of 32 consecutive vertices (one warp each): k_tile_filter picks the tiles that
// must be looked at, k_scan<W> classifies their vertices (boundary? dirty?), emits the compact work list of
// boundary vertices that must be (re)evaluated, lets the other boundary vertices re-submit their stored
// proposal, and records the tile's cluster signature.
constexpr int kSigSlots = 8;     // cluster ids remembered per 32-vertex tile (32 bytes)

// k_tile_filter: a tile (32 consecutive vertices) must be re-scanned only if one of the clusters in its
// signature -- the clusters of its vertices and of their neighbours, recorded by the last scan of the
// tile -- was modified in the previous round: a vertex's boundary/dirty state and proposal depend on
// those clusters only, and any move that changes the tile's signature modifies a cluster already in it.
// Reads 32 B per tile instead of ~1 KB of CSR: tail rounds cost microseconds, not a full sweep.
// [tile_begin, n_tiles) is the tile range owned by this rank (the whole mesh on one GPU).
__global__ void __launch_bounds__(kThreads) k_tile_filter(int tile_begin, int n_tiles, int K, int force_all, const int4* __restrict__ sig,
                                                          const unsigned* __restrict__ modbits, unsigned char* tile_active,
                                                          int* active_tiles, unsigned long long* n_active) {
    const int lane = threadIdx.x & 31;
    for (int t0 = tile_begin + ((blockIdx.x * blockDim.x + threadIdx.x) & ~31); t0 < n_tiles; t0 += gridDim.x * blockDim.x) {
        const int t = t0 + lane;
        bool act = false;
        if (t < n_tiles) {
            act = force_all != 0;
            if (!act) {
                int4 s0 = __ldg(sig + 2 * (int64_t)t), s1 = __ldg(sig + 2 * (int64_t)t + 1);
                int c[kSigSlots] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                act = c[0] == -2;   // overflowed signature: always rescan
#pragma unroll
                for (int k = 0; k < kSigSlots; k++)
                    act = act || (c[k] >= 0 && c[k] < K && ((modbits[c[k] >> 5] >> (c[k] & 31)) & 1u));
            }
            tile_active[t] = act ? 1 : 0;
        }
        const unsigned m = __ballot_sync(0xffffffffu, act);
        if (m) {
            int base = 0;
            if (lane == 0) base = (int)atomicAdd(n_active, (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (act) active_tiles[base + __popc(m & ((1u << lane) - 1u))] = t;
        }
    }
}

constexpr int kSigHash = 16;     // slots of the per-warp hash set the signature is collected in

// insert `val` into the warp's open-addressing hash set (shared memory, -1 = empty); lanes insert in parallel
__device__ __forceinline__ void sig_hash_insert(int* set, int val, bool& overflow) {
    unsigned h = ((unsigned)val * 2654435761u) >> 28;
#pragma unroll 1
    for (int probe = 0; probe < kSigHash; probe++) {
        const int slot = (h + probe) & (kSigHash - 1);
        int old = reinterpret_cast<volatile int*>(set)[slot];          // usually already there: no atomic
        if (old == val) return;
        if (old == -1) old = atomicCAS(&set[slot], -1, val);
        if (old == -1 || old == val) return;
    }
    overflow = true;
}

// k_scan<W>: one thread per vertex of an active tile, neighbours from the ELL copy of the adjacency (W columns,
// column-major: every load of a warp is one coalesced line; all W loads and the W gathers of neighbour cluster ids
// are independent, so they are in flight together).  Rows longer than W finish from the CSR (rare).
// The tile signature (distinct clusters of the tile's vertices and of their neighbours) is collected in a 16-slot
// hash set per warp in shared memory: every lane inserts its foreign neighbour clusters with atomicCAS, in
// parallel, so the cost does not depend on how many clusters meet in the tile.
// BULK selects at compile time whether the bulk decision is taken in here (keeps the exact-round variant lean).
template <int W, bool BULK>
__global__ void __launch_bounds__(kThreads) k_scan(ReassignArgs A) {
    __shared__ int s_sig[kThreads / 32][kSigHash];
    const int K = A.K, V = A.V;
    const int lane = threadIdx.x & 31;
    const unsigned lane_lt = (1u << lane) - 1u;
    int* set = s_sig[threadIdx.x >> 5];
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    // dense rounds scan the rank's whole tile range; otherwise the list written by k_tile_filter of this round
    const int n_active = A.all_tiles ? (A.tile_end - A.tile_begin) : (int)*A.n_active_tiles;
    const unsigned* __restrict__ modbits = A.modbits;
    const int* __restrict__ ell = A.ell;
    const int64_t vpad = A.vpad;
    unsigned n_bnd = 0, n_fused = 0, n_tests = 0, n_props = 0;
    // dense rounds: every block streams through one contiguous chunk of tiles, so the cluster ids of the mesh
    // rows above / below (needed again a row later) are still in this SM's L1; list rounds: grid-stride
    int ti_begin = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, ti_end = n_active, ti_step = n_warps;
    if (A.all_tiles) {
        const int chunk = (n_active + gridDim.x - 1) / gridDim.x;
        ti_begin = blockIdx.x * chunk + (threadIdx.x >> 5);
        ti_end = min(n_active, (int)(blockIdx.x + 1) * chunk);
        ti_step = blockDim.x >> 5;
    }
    for (int ti = ti_begin; ti < ti_end; ti += ti_step) {
        const int tile = A.all_tiles ? (A.tile_begin + ti) : A.active_tiles[ti];
        const int v = tile * 32 + lane;
        const bool valid = v < V;
        const int a = valid ? A.cid[v] : -1;
        int nb[W];
#pragma unroll
        for (int k = 0; k < W; k++) nb[k] = __ldg(ell + (int64_t)k * vpad + v);      // v < vpad always
        const bool overflow_row = nb[W - 1] == -2;
        if (overflow_row) nb[W - 1] = A.col[A.row_ptr[v] + W - 1];
#pragma unroll
        for (int k = 0; k < W; k++) nb[k] = valid ? A.cid[nb[k]] : a;               // neighbour cluster ids (short rows are padded with v itself)
        // the signature is rebuilt only when a vertex in / next to the tile moved since it was recorded
        const bool rebuild = A.sig_mode == 2 || (A.sig_mode == 0 && (A.force_all || A.tile_stale[tile]));
        bool sig_overflow = false;
        if (rebuild) {
            if (lane < kSigHash) set[lane] = -1;
            __syncwarp();
            // own clusters: one lane per run of equal ids inserts
            const int prev = __shfl_up_sync(0xffffffffu, a, 1);
            if (valid && (lane == 0 || prev != a)) sig_hash_insert(set, a, sig_overflow);
        }
        bool bnd = false, dirty = false;
        int last = a;
        auto visit = [&](int bb) {
            const bool isb = bb != a;
            bnd |= isb;
            if (isb) {
                if (bb < K) dirty |= (modbits[bb >> 5] >> (bb & 31)) & 1u;
                if (rebuild && bb != last) { sig_hash_insert(set, bb, sig_overflow); last = bb; }
            }
        };
#pragma unroll
        for (int k = 0; k < W; k++) visit(nb[k]);
        if (__any_sync(0xffffffffu, overflow_row)) {        // finish long rows from the CSR, warp-uniformly
            const int e0 = overflow_row ? A.row_ptr[v] + W : 0, e1 = overflow_row ? A.row_ptr[v + 1] : 0;
            const int steps = __reduce_max_sync(0xffffffffu, e1 - e0);
            for (int s = 0; s < steps; s++) visit((e0 + s < e1) ? A.cid[A.col[e0 + s]] : a);
        }
        __syncwarp();
        if (rebuild) {   // compact the hash set into the 8-slot signature (32 B, coalesced); more than 8 clusters: overflow mark
            const int mine = lane < kSigHash ? set[lane] : -1;
            const unsigned full = __ballot_sync(0xffffffffu, mine != -1);
            const bool ovf = __any_sync(0xffffffffu, sig_overflow) || __popc(full) > kSigSlots;
            const int rank = __popc(full & lane_lt);
            int* sig = A.tile_sig + (int64_t)tile * kSigSlots;
            if (lane < kSigSlots) sig[lane] = -1;
            __syncwarp();
            if (mine != -1 && rank < kSigSlots) sig[rank] = mine;
            __syncwarp();
            if (ovf && lane == 0) sig[0] = -2;
            if (lane == 0) A.tile_stale[tile] = 0;
        }
        bnd = bnd && valid;
        if (bnd) {
            n_bnd++;
            dirty = dirty || (a < K && ((modbits[a >> 5] >> (a & 31)) & 1u));
        }
        bool work = bnd && dirty;
        if (BULK) {
            // bulk rounds: the decision needs only the neighbour cluster ids already in registers, the vertex
            // position (coalesced) and the centroids of the few clusters involved -- taken here, no second pass.
            // Rows longer than W (rare) go through the work list to k_bulk_evaluate.
            const bool fused = work && !overflow_row;
            int best_b = -1;
            if (fused) {
                n_fused++;
                if (a >= K) {
#pragma unroll
                    for (int k = W - 1; k >= 0; k--) if (nb[k] < K) best_b = nb[k];     // first assigned neighbour cluster
                } else {
                    const double px = A.xyz[3 * (int64_t)v], py = A.xyz[3 * (int64_t)v + 1], pz = A.xyz[3 * (int64_t)v + 2];
                    const bool blocked = A.csize[a] == 1;
                    const double4 ca = *reinterpret_cast<const double4*>(A.bulk_cen + 4 * (int64_t)a);
                    double dx = px - ca.x, dy = py - ca.y, dz = pz - ca.z;
                    double best = dx * dx + dy * dy + dz * dz, w = 0.0;
                    if (A.bulk_stage == 1) {
                        w = __ldg(A.weight + v);
                        best = ca.w / (ca.w - w) * best;
                    }
#pragma unroll
                    for (int k = 0; k < W; k++) {
                        const int b = nb[k];
                        bool fresh = b != a && b < K;
#pragma unroll
                        for (int q = 0; q < k; q++) fresh = fresh && (nb[q] != b);
                        if (!fresh) continue;
                        n_tests++;
                        if (blocked) continue;
                        const double4 cb = *reinterpret_cast<const double4*>(A.bulk_cen + 4 * (int64_t)b);
                        dx = px - cb.x; dy = py - cb.y; dz = pz - cb.z;
                        double d = dx * dx + dy * dy + dz * dz;
                        if (A.bulk_stage == 1) d = cb.w / (cb.w + w) * d;
                        if (d < best) { best = d; best_b = b; }
                    }
                }
                A.prop_dst[v] = best_b;
                if (best_b >= 0 && a < K && A.bulk_count_leave) atomicAdd(&A.bulk_leave[a], 1);
            }
            const unsigned mp = __ballot_sync(0xffffffffu, fused && best_b >= 0);
            if (lane == 0) { A.prop_mask[tile] = mp; n_props += __popc(mp); }
            work = work && overflow_row;
        }
        const unsigned mw = __ballot_sync(0xffffffffu, work);
        if (mw) {
            int basew = 0;
            if (lane == 0) basew = (int)atomicAdd(&A.ctr->evaluated, (unsigned long long)__popc(mw));
            basew = __shfl_sync(0xffffffffu, basew, 0);
            if (work) A.work[basew + __popc(mw & lane_lt)] = v;
        }
        if (!BULK && bnd && !dirty) {   // clusters unchanged since the last evaluation: the stored proposal is still exact
            const int d = A.prop_dst[v];
            if (d >= 0) {
                const unsigned long long key = A.prop_key[v];
                if (a < K) atomicMin(&A.best[a], key);
                atomicMin(&A.best[d], key);
                int slot = (int)atomicAdd(&A.ctr->proposals, 1ull);
                A.plist[slot] = v;
            }
        }
    }
    warp_count_add(&A.ctr->boundary, n_bnd);
    if (BULK) {
        warp_count_add(&A.ctr->pad[0], n_fused);   // vertices decided inside the scan
        warp_count_add(&A.ctr->tests, n_tests);
        warp_count_add(&A.ctr->proposals, n_props);
    }
}

// k_carry: live proposals of the previous round whose tile is not re-scanned this round (none of their
// clusters changed) compete again with their stored key.
__global__ void __launch_bounds__(kThreads) k_carry(ReassignArgs A) {
    const int K = A.K;
    const int n_old = (int)*A.n_prev_props;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_old; i += gridDim.x * blockDim.x) {
        const int v = A.plist_prev[i];
        if (A.tile_active[v >> 5]) continue;                  // the scan handles vertices of active tiles
        const int d = A.prop_dst[v];
        if (d < 0) continue;                                  // committed last round
        const int a = A.cid[v];
        const unsigned long long key = A.prop_key[v];
        if (a < K) atomicMin(&A.best[a], key);
        atomicMin(&A.best[d], key);
        int slot = (int)atomicAdd(&A.ctr->proposals, 1ull);
        A.plist[slot] = v;
    }
}

// k_resubmit: between commit passes of one round.  A proposal whose two clusters were not touched by
// the commits so far is still exact (same sums, same energies, same ring in its source cluster), so it
// competes again; repeating select + commit a few times approaches a maximal independent set of moves
// per round without re-scanning the mesh.
__device__ __forceinline__ void resubmit_list(const ReassignArgs& A, int n_props, unsigned long long* n_resubmitted = nullptr) {
    const int K = A.K;
    unsigned cnt = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_props; i += gridDim.x * blockDim.x) {
        const int v = A.plist[i];
        const int d = A.prop_dst[v];
        if (d < 0) continue;                                  // committed in an earlier pass
        const int a = A.cid[v];
        if (A.mod_round[d] == A.round || (a < K && A.mod_round[a] == A.round)) continue;
        const unsigned long long key = A.prop_key[v];
        if (a < K) atomicMin(&A.best[a], key);
        atomicMin(&A.best[d], key);
        cnt++;
    }
    if (n_resubmitted) warp_count_add(n_resubmitted, cnt);
}
__global__ void __launch_bounds__(kThreads) k_resubmit(ReassignArgs A) { resubmit_list(A, (int)A.ctr->proposals); }

// ---------------------------------------------------------------------------------------------------
// k_evaluate: every work-list vertex evaluates its candidate moves.
// EM: metric whose energy is evaluated; STRIDE: doubles per payload row in memory.
// (QEM's unconstrained phase evaluates the isotropic energy on the first 4 doubles of its rows.)
//
// The heavy part of a test is the energy of the grown destination cluster (for the quadric metrics a 3x3
// eigen-solve).  A warp therefore walks the candidates by *rank*: all lanes evaluate their first distinct adjacent
// cluster together, then their second, ... -- two or three executions of the heavy body per warp instead of one
// per ring slot.  The ring (<= kRingW neighbours) lives in registers; the connexity predicate
// (vtkVerticesProcessing::ConnexityConstraintProblemLocal, DiscreteRemeshing/vtkVerticesProcessing.h:168-237)
// is evaluated with bit operations on the precomputed ring adjacency matrix (k_build_ringadj).  Rows longer than
// kRingW are left to k_evaluate_long.
constexpr int kRingW = 8;

// ringadj[v]: bit 8 i + j is set iff the i-th and the j-th neighbour of v (CSR order) are joined by a mesh edge;
// 0 for rows longer than kRingW (those use connexity_problem).  Static topology, built once per mesh.
__global__ void __launch_bounds__(kThreads) k_build_ringadj(int V, const int* __restrict__ row_ptr, const int* __restrict__ col,
                                                            unsigned long long* ringadj) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
        const int beg = row_ptr[v], deg = row_ptr[v + 1] - beg;
        unsigned long long m = 0;
        if (deg <= kRingW) {
            int ring[kRingW];
#pragma unroll
            for (int k = 0; k < kRingW; k++) ring[k] = k < deg ? col[beg + k] : -1;
            for (int i = 0; i < deg; i++) {
                const int u = ring[i];
                for (int e = row_ptr[u]; e < row_ptr[u + 1]; e++) {
                    const int w = col[e];
#pragma unroll
                    for (int j = 0; j < kRingW; j++)
                        if (ring[j] == w) m |= 1ull << (8 * i + j);
                }
            }
        }
        ringadj[v] = m;
    }
}

// "the ring members of cluster a (bit mask L over the ring slots) are connected in the sub-graph they induce"
__device__ __forceinline__ bool connexity_problem_ring(unsigned L, unsigned long long adj) {
    if ((L & (L - 1)) == 0) return false;           // zero or one member
    unsigned reach = L & (0u - L);
#pragma unroll 1
    for (int iter = 0; iter < kRingW; iter++) {
        unsigned nxt = reach;
#pragma unroll
        for (int i = 0; i < kRingW; i++)
            if ((reach >> i) & 1u) nxt |= (unsigned)(adj >> (8 * i)) & 0xffu;
        nxt &= L;
        if (nxt == reach) break;
        reach = nxt;
    }
    return reach != L;
}

// k_evaluate_long: the work-list vertices whose rows are longer than kRingW, one thread per vertex, per-slot walk.
// Launched only for meshes that have such vertices.
template <int EM, int STRIDE>
static __device__ __noinline__ void evaluate_long_list(const ReassignArgs& A, int n_work) {
    constexpr int NL = MetricTraits<EM>::NPAD;
    const int K = A.K;
    unsigned n_tests = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_work; i += gridDim.x * blockDim.x) {
        const int v = A.work[i];
        const int beg = A.row_ptr[v], end = A.row_ptr[v + 1];
        if (end - beg <= kRingW) continue;               // k_evaluate's
        const int a = A.cid[v];
        int best_b = -1;
        double best_delta = 0.0, best_ea = 0.0, best_eb = 0.0;
        unsigned long long key = 0;
        if (a >= K) {
            // NULL cluster: adopt the first assigned neighbour cluster, frozen or not -- the reference adopts
            // (:881-907) before it looks at IsClusterFreezed (:909-920); top priority
            for (int e = beg; e < end && best_b < 0; e++) {
                int b = A.cid[A.col[e]];
                if (b < K) best_b = b;
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
                for (int k = 0; k < NL; k++) s[k] -= it[k];
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
                for (int k = 0; k < NL / 2; k++) { double2 u = ps[k]; s[2 * k] = u.x + it[2 * k]; s[2 * k + 1] = u.y + it[2 * k + 1]; }
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
}
template <int EM, int STRIDE>
__global__ void __launch_bounds__(kThreads) k_evaluate_long(ReassignArgs A) { evaluate_long_list<EM, STRIDE>(A, (int)A.ctr->evaluated); }

template <int EM, int STRIDE>
__device__ __forceinline__ void evaluate_list(const ReassignArgs& A, int n_work) {
    constexpr int NL = MetricTraits<EM>::NPAD;   // doubles loaded per row
    const int K = A.K;
    const int lane = threadIdx.x & 31;
    unsigned n_tests = 0;
    // warp-uniform trip count: the candidate loop below votes across the warp
    for (int i0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; i0 < n_work; i0 += gridDim.x * blockDim.x) {
        const int i = i0 + lane;
        const int v = i < n_work ? A.work[i] : 0;
        const int beg = A.row_ptr[v], deg = A.row_ptr[v + 1] - beg;
        const bool active = i < n_work && deg <= kRingW;      // longer rows: k_evaluate_long
        const int a = active ? A.cid[v] : K;
        int best_b = -1;
        double best_delta = 0.0, best_ea = 0.0, best_eb = 0.0;
        unsigned long long key = 0;
        unsigned nt = 0;
        // ---- fast path: ring in registers
        int nbc[kRingW];
#pragma unroll
        for (int k = 0; k < kRingW; k++) nbc[k] = (active && k < deg) ? A.cid[A.col[beg + k]] : a;
        unsigned rem = 0;
        bool blocked = true;
        double it[NL], s[NL];
        double ea_new = 0.0, cur_a = 0.0;
        double anchor_pt[3];
        if (active) {
            if (a >= K) {
                // NULL cluster: adopt the first assigned neighbour cluster, frozen or not (:881-907 come before the
                // frozen test of :909-920); top priority
#pragma unroll
                for (int k = kRingW - 1; k >= 0; k--)
                    if (nbc[k] < K) best_b = nbc[k];
                key = (unsigned long long)(unsigned)v;
            } else if (!(A.frozen && A.frozen[a])) {
                unsigned L = 0;
#pragma unroll
                for (int k = 0; k < kRingW; k++) {
                    const int b = nbc[k];
                    const bool asg = b != a && b < K && !(A.frozen && A.frozen[b]);
                    rem |= (asg ? 1u : 0u) << k;
                    L |= ((k < deg && b == a) ? 1u : 0u) << k;
                }
                blocked = (A.csize[a] == 1) || (A.anchor && A.anchor[a] == v);
                if (!blocked && A.connexity) blocked = connexity_problem_ring(L, A.ringadj[v]);
                cur_a = A.cenergy[a];
                if (!blocked) {
                    load_row_ro<NL>(A.items + (int64_t)v * STRIDE, it);
                    load_row<NL>(A.csum + (int64_t)a * STRIDE, s);
#pragma unroll
                    for (int k = 0; k < NL; k++) s[k] -= it[k];
                    ea_new = cluster_energy<EM>(s, A.cfg, nullptr, anchor_point<EM>(A, a, anchor_pt));
                }
            }
        }
        // candidates by rank (first occurrence of every distinct adjacent cluster, in ring order)
        while (__any_sync(0xffffffffu, rem != 0)) {
            if (rem) {
                const int k0 = __ffs(rem) - 1;
                int b = nbc[0];
#pragma unroll
                for (int k = 1; k < kRingW; k++) b = (k0 == k) ? nbc[k] : b;
#pragma unroll
                for (int k = 0; k < kRingW; k++) rem &= ~((nbc[k] == b ? 1u : 0u) << k);
                nt++;
                if (!blocked) {
                    const double2* ps = reinterpret_cast<const double2*>(A.csum + (int64_t)b * STRIDE);
#pragma unroll
                    for (int k = 0; k < NL / 2; k++) { double2 u = ps[k]; s[2 * k] = u.x + it[2 * k]; s[2 * k + 1] = u.y + it[2 * k + 1]; }
                    const double eb_new = cluster_energy<EM>(s, A.cfg, nullptr, anchor_point<EM>(A, b, anchor_pt));
                    const double tr = ea_new + eb_new;
                    const double cur = cur_a + A.cenergy[b];
                    if (tr < cur) {
                        const double delta = tr - cur;
                        if (best_b < 0 || delta < best_delta) { best_b = b; best_delta = delta; best_ea = ea_new; best_eb = eb_new; }
                    }
                }
            }
        }
        if (active) {
            if (a < K && best_b >= 0)
                key = ((unsigned long long)ordered_float_bits(__double2float_rn(best_delta)) << 32) | (unsigned)v;
            n_tests += nt;
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
    }
    warp_count_add(&A.ctr->tests, n_tests);
}
// n_work = the work list length written by the scan of this round
template <int EM, int STRIDE>
__global__ void __launch_bounds__(kThreads) k_evaluate(ReassignArgs A) { evaluate_list<EM, STRIDE>(A, (int)A.ctr->evaluated); }

// Bookkeeping every committed move owes the sparse rounds (sparse.cuh): the per-cluster member arrays (swap-remove
// from the source, append to the destination) and the list of the clusters modified in this round.  Winners of one
// commit pass touch pairwise disjoint cluster pairs and a cluster is touched at most once per round, so neither
// structure needs atomics beyond the list cursor.  Called BEFORE the sizes are updated.
__device__ __forceinline__ void note_move(const ReassignArgs& A, int v, int a, int d) {
    if (A.mem.memb) {
        if (a < A.K) {
            const int p = A.mem.pos[v];
            const int last = A.mem.memb[A.mem.off[a] + A.csize[a] - 1];
            A.mem.memb[p] = last;
            A.mem.pos[last] = p;
        }
        const int q = A.mem.off[d] + A.csize[d];
        if (q < A.mem.off[d + 1]) { A.mem.memb[q] = v; A.mem.pos[v] = q; }
        else { *A.mem.overflow = 1; A.ctr->pad[2] = 1; }      // array full: the driver rebuilds the arrays before they are read
    }
    if (A.modlist) {
        const int s = (int)atomicAdd(A.n_mod, a < A.K ? 2ull : 1ull);
        A.modlist[s] = d;
        if (a < A.K) A.modlist[s + 1] = a;
    }
}

// UM: metric of the stored sums (all UM::NPAD doubles of a row are updated);
// EM: metric used to re-evaluate an adopting cluster's energy.
template <int EM, int UM>
__device__ __forceinline__ void commit_list(const ReassignArgs& A, int n_props) {
    constexpr int NU = MetricTraits<UM>::NPAD;
    const int K = A.K;
    unsigned n_mods = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_props; i += gridDim.x * blockDim.x) {
        const int v = A.plist[i];
        const int d = A.prop_dst[v];
        if (d < 0) continue;                                  // committed in an earlier pass of this round
        const unsigned long long key = A.prop_key[v];
        const int a = A.cid[v];
        bool win = (A.best[d] == key) && (a >= K || A.best[a] == key);
        if (!win) continue;
        note_move(A, v, a, d);
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
        mark_tiles_stale(A, v);
        n_mods++;
    }
    warp_count_add(&A.ctr->mods, n_mods);
}
// n_props = proposals submitted by the scan / evaluation of this round
template <int EM, int UM>
__global__ void __launch_bounds__(kThreads) k_commit(ReassignArgs A) { commit_list<EM, UM>(A, (int)A.ctr->proposals); }

// ---------------------------------------------------------------------------------------------------
// Bulk rounds (Lloyd criterion) for the phases that the reference ends by "early convergence"
// (Common/vtkUniformClustering.h:773-776) and whose energy is the centroid energy E = sum w |p - c|^2
// (isotropic metric; QEM while unconstrained, vtkQEMetricForClustering.h:277-280).
//
// A move v: a -> b with |p_v - c_b|^2 < |p_v - c_a|^2 is a strictly improving candidate of the
// reference's test (its delta-E is w [W_b/(W_b+w) d_b^2 - W_a/(W_a-w) d_a^2] < 0), and for fixed centroids
// every such move lowers sum w |p - c|^2 independently of the others; recomputing the centroids lowers
// it again.  So all of them commit in the same round, with no per-cluster exclusivity.  The moves the
// reference accepts beyond this criterion (a thin band near the bisectors) are left to the exact rounds
// that follow in the same phase.  Sums are kept in 64-bit fixed point while in bulk mode: integer atomics
// commute, so the result does not depend on the order of the adds; exact double statistics are
// recomputed from the clustering before the exact rounds resume.
struct BulkArgs {
    long long* isum;            // K x 4 fixed-point (S, W)
    double* ccen;               // K x 4: centroid and total weight
    double* cen_energy;         // K: -|S|^2 / W of the fixed-point sums (energy guard of the second bulk stage)
    int* leave_cnt;             // K: vertices that want to leave the cluster this round
    int* join_cnt;              // K: vertices that joined the cluster this round
    double scale;               // fixed-point scale (power of two)
};

__global__ void __launch_bounds__(kThreads) k_bulk_init(int K, int stride, const double* __restrict__ csum, BulkArgs B) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < K; c += gridDim.x * blockDim.x) {
        const double* s = csum + (int64_t)c * stride;
#pragma unroll
        for (int k = 0; k < 4; k++) B.isum[4 * (int64_t)c + k] = __double2ll_rn(s[k] * B.scale);
        B.ccen[4 * c] = s[0] / s[3]; B.ccen[4 * c + 1] = s[1] / s[3]; B.ccen[4 * c + 2] = s[2] / s[3]; B.ccen[4 * c + 3] = s[3];
        B.cen_energy[c] = -(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]) / s[3];
        B.leave_cnt[c] = 0; B.join_cnt[c] = 0;
    }
}

// stage 0 (Lloyd criterion): move to the adjacent cluster whose centroid is strictly closer than the own one.
// stage 1 (exact criterion, for the thin band of moves the first stage leaves): delta-E of the reference's test
// against the round-start sums, w [W_b/(W_b+w) d_b^2 - W_a/(W_a-w) d_a^2] < 0, best candidate.
__global__ void __launch_bounds__(kThreads) k_bulk_evaluate(ReassignArgs A, BulkArgs B, int count_leave, int stage, int stride) {
    const int K = A.K;
    const int n_work = (int)A.ctr->evaluated;
    unsigned n_tests = 0, n_props = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_work; i += gridDim.x * blockDim.x) {
        const int v = A.work[i];
        const int a = A.cid[v];
        const int beg = A.row_ptr[v], end = A.row_ptr[v + 1];
        int best_b = -1;
        if (a >= K) {
            for (int e = beg; e < end && best_b < 0; e++) {
                int b = A.cid[A.col[e]];
                if (b < K) best_b = b;
            }
        } else {
            const double px = A.xyz[3 * (int64_t)v], py = A.xyz[3 * (int64_t)v + 1], pz = A.xyz[3 * (int64_t)v + 2];
            const bool blocked = A.csize[a] == 1;
            const double4 ca = *reinterpret_cast<const double4*>(B.ccen + 4 * (int64_t)a);
            double dx = px - ca.x, dy = py - ca.y, dz = pz - ca.z;
            const double da = dx * dx + dy * dy + dz * dz;
            double w = 0.0, best = da;
            if (stage == 1) {
                w = __ldg(A.weight + v);
                best = ca.w / (ca.w - w) * da;          // what leaving the own cluster gains (per unit weight)
            }
            for (int e = beg; e < end; e++) {
                int b = A.cid[A.col[e]];
                if (b == a || b >= K) continue;
                bool seen = false;
                for (int e2 = beg; e2 < e; e2++) seen |= (A.cid[A.col[e2]] == b);
                if (seen) continue;
                n_tests++;
                if (blocked) continue;
                const double4 cb = *reinterpret_cast<const double4*>(B.ccen + 4 * (int64_t)b);
                dx = px - cb.x; dy = py - cb.y; dz = pz - cb.z;
                double d = dx * dx + dy * dy + dz * dz;
                if (stage == 1) d = cb.w / (cb.w + w) * d;  // what joining b costs (per unit weight)
                if (d < best) { best = d; best_b = b; }
            }
        }
        A.prop_dst[v] = best_b;
        if (best_b >= 0) {
            if (a < K && count_leave) atomicAdd(&B.leave_cnt[a], 1);
            atomicOr(&A.prop_mask[v >> 5], 1u << (v & 31));
            n_props++;
        }
    }
    warp_count_add(&A.ctr->tests, n_tests);
    warp_count_add(&A.ctr->proposals, n_props);
}

// Applies every proposal unless its source cluster would be emptied (then all of that cluster's leavers
// wait: the decision depends only on totals, so it is deterministic).  The proposers are the set bits of the
// tiles' proposal masks: a warp reads 32 masks at a time (coalesced), clears them, and visits the non-empty tiles.
__global__ void __launch_bounds__(kThreads) k_bulk_commit(ReassignArgs A, BulkArgs B, int stride) {
    const int K = A.K;
    const int n_tiles = A.tile_end - A.tile_begin;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    unsigned n_mods = 0;
    for (int base = warp * 32; base < n_tiles; base += n_warps * 32) {
        const int t = A.tile_begin + base + lane;
        const unsigned m = (base + lane < n_tiles) ? A.prop_mask[t] : 0u;
        if (m) A.prop_mask[t] = 0;
        unsigned nz = __ballot_sync(0xffffffffu, m != 0);
        unsigned moved_here = 0;        // lane `src` collects the moved bits of its tile
        while (nz) {
            const int src = __ffs(nz) - 1;
            nz &= nz - 1;
            const unsigned mm = __shfl_sync(0xffffffffu, m, src);
            const bool mine = (mm >> lane) & 1u;
            const int v = (A.tile_begin + base + src) * 32 + lane;
            const int d = mine ? A.prop_dst[v] : -1;
            const int a = mine ? A.cid[v] : -1;
            const bool go = mine && !(a < K && B.leave_cnt[a] >= A.csize[a]);
            if (A.moved_mask) {          // stage 1 keeps what it takes to undo the round (energy guard)
                const unsigned mv = __ballot_sync(0xffffffffu, go);
                if (lane == src) moved_here = mv;
                if (go) A.prop_dst[v] = a;
            }
            if (!go) continue;
            const double* it = A.items + (int64_t)v * stride;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                long long f = __double2ll_rn(__ldg(it + k) * B.scale);
                atomicAdd(reinterpret_cast<unsigned long long*>(&B.isum[4 * (int64_t)d + k]), (unsigned long long)f);
                if (a < K) atomicAdd(reinterpret_cast<unsigned long long*>(&B.isum[4 * (int64_t)a + k]), (unsigned long long)(-f));
            }
            atomicAdd(&B.join_cnt[d], 1);
            A.mod_round[d] = A.round;
            if (a < K) A.mod_round[a] = A.round;
            A.cid[v] = d;
            mark_tiles_stale(A, v);
            n_mods++;
        }
        if (A.moved_mask && base + lane < n_tiles) A.moved_mask[t] = moved_here;
    }
    warp_count_add(&A.ctr->mods, n_mods);
}

// Undo of the last stage-1 bulk round (its energy guard tripped): every moved vertex goes back to the cluster it came
// from (k_bulk_commit left it in prop_dst).  Sums, sizes and centroids need no undo: exact statistics are recomputed
// from the clustering when the bulk rounds end.
__global__ void __launch_bounds__(kThreads) k_bulk_rollback(ReassignArgs A) {
    const int n_tiles = A.tile_end - A.tile_begin;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_tiles * 32; i += gridDim.x * blockDim.x) {
        const int t = A.tile_begin + (i >> 5);
        if ((A.moved_mask[t] >> (i & 31)) & 1u) { const int v = t * 32 + (i & 31); A.cid[v] = A.prop_dst[v]; }
    }
}
// multi-GPU form: the moves are the all-gathered (vertex, destination) records
__global__ void __launch_bounds__(kThreads) k_bulk_rollback_moves(int* cid, const int* __restrict__ prev, const int2* __restrict__ moves, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int v = moves[i].x;
        if (cid[v] == moves[i].y) cid[v] = prev[v];
    }
}

__global__ void __launch_bounds__(kThreads) k_bulk_refresh(int K, int* csize, BulkArgs B) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < K; c += gridDim.x * blockDim.x) {
        const int lv = B.leave_cnt[c], jn = B.join_cnt[c];
        if ((lv | jn) == 0) continue;
        const int sz = csize[c];
        const int left = (lv < sz) ? lv : 0;
        if (left | jn) {
            csize[c] = sz - left + jn;
            const double inv = 1.0 / B.scale;
            const double sx = (double)B.isum[4 * (int64_t)c] * inv, sy = (double)B.isum[4 * (int64_t)c + 1] * inv;
            const double sz = (double)B.isum[4 * (int64_t)c + 2] * inv, w = (double)B.isum[4 * (int64_t)c + 3] * inv;
            B.ccen[4 * c] = sx / w; B.ccen[4 * c + 1] = sy / w; B.ccen[4 * c + 2] = sz / w; B.ccen[4 * c + 3] = w;
            B.cen_energy[c] = -(sx * sx + sy * sy + sz * sz) / w;
        }
        B.leave_cnt[c] = 0; B.join_cnt[c] = 0;
    }
}

// per-vertex bound used to pick the fixed-point scale: max(|S_x|, |S_y|, |S_z|, W)
__global__ void __launch_bounds__(kThreads) k_item_bound(int V, int stride, const double* __restrict__ items, double* out) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
        const double* it = items + (int64_t)v * stride;
        out[v] = fmax(fmax(fabs(it[0]), fabs(it[1])), fmax(fabs(it[2]), fabs(it[3])));
    }
}

// ---------------------------------------------------------------------------------------------------
// Multi-GPU rounds (one process per GPU).  Every rank holds the whole clustering state (cluster ids,
// per-cluster sums/energies/sizes) and owns a contiguous range of 32-vertex tiles: it scans and evaluates
// only its own range.  Per round two exchanges go over NCCL (NVLink 5 / NVSwitch):
//   1. conflict resolution: min-allreduce of the per-cluster priority keys;
//   2. the winners (vertex, destination, the two new energies) are all-gathered -- this *is* the halo
//      exchange of boundary cluster ids, in delta form -- and every rank applies every move to its replica.
// Because all ranks apply the same moves with the same arithmetic, the replicas stay bit-identical and
// the result is identical to the single-GPU run.
struct MoveRec { int v, d; double ea, eb; };

__global__ void __launch_bounds__(kThreads) k_select_winners(ReassignArgs A, MoveRec* moves, unsigned long long* n_moves) {
    const int K = A.K;
    const int n_props = (int)A.ctr->proposals;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_props; i += gridDim.x * blockDim.x) {
        const int v = A.plist[i];
        const int d = A.prop_dst[v];
        if (d < 0) continue;
        const unsigned long long key = A.prop_key[v];
        const int a = A.cid[v];
        if ((A.best[d] == key) && (a >= K || A.best[a] == key)) {
            const double2 pe = A.prop_e[v];
            const int slot = (int)atomicAdd(n_moves, 1ull);
            moves[slot] = MoveRec{v, d, pe.x, pe.y};
        }
    }
}

template <int EM, int UM>
__global__ void __launch_bounds__(kThreads) k_apply_moves(ReassignArgs A, const MoveRec* __restrict__ moves, int n_moves) {
    constexpr int NU = MetricTraits<UM>::NPAD;
    const int K = A.K;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_moves; i += gridDim.x * blockDim.x) {
        const MoveRec m = moves[i];
        const int v = m.v, d = m.d;
        const int a = A.cid[v];
        note_move(A, v, a, d);
        double it[NU], s[NU];
        load_row_ro<NU>(A.items + (int64_t)v * NU, it);
        load_row<NU>(A.csum + (int64_t)d * NU, s);
#pragma unroll
        for (int k = 0; k < NU; k++) s[k] += it[k];
        store_row<NU>(A.csum + (int64_t)d * NU, s);
        if (a >= K) {
            double anchor_pt[3];
            A.cenergy[d] = cluster_energy<EM>(s, A.cfg, nullptr, anchor_point<EM>(A, d, anchor_pt));
        } else {
            A.cenergy[d] = m.eb;
            load_row<NU>(A.csum + (int64_t)a * NU, s);
#pragma unroll
            for (int k = 0; k < NU; k++) s[k] -= it[k];
            store_row<NU>(A.csum + (int64_t)a * NU, s);
            A.cenergy[a] = m.ea;
            A.csize[a] -= 1;
            A.mod_round[a] = A.round;
        }
        A.csize[d] += 1;
        A.mod_round[d] = A.round;
        A.cid[v] = d;
        A.prop_dst[v] = -1;
        mark_tiles_stale(A, v);
    }
}

// multi-GPU bulk rounds: the rank's proposals (set bits of its tiles' masks) -> compact (vertex, destination) records
__global__ void __launch_bounds__(kThreads) k_pack_bulk_moves(ReassignArgs A, int2* moves, unsigned long long* n_moves) {
    const int n_tiles = A.tile_end - A.tile_begin;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int base = warp * 32; base < n_tiles; base += n_warps * 32) {
        const int t = A.tile_begin + base + lane;
        const unsigned m = (base + lane < n_tiles) ? A.prop_mask[t] : 0u;
        if (m) A.prop_mask[t] = 0;
        // exclusive prefix of the popcounts over the 32 tiles, one atomic per warp
        const int cnt = __popc(m);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int x = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += x; }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) continue;
        int slot0 = 0;
        if (lane == 0) slot0 = (int)atomicAdd(n_moves, (unsigned long long)total);
        slot0 = __shfl_sync(0xffffffffu, slot0, 0);
        const int my_off = slot0 + incl - cnt;
        unsigned nz = __ballot_sync(0xffffffffu, m != 0);
        while (nz) {
            const int src = __ffs(nz) - 1;
            nz &= nz - 1;
            const unsigned mm = __shfl_sync(0xffffffffu, m, src);
            const int off = __shfl_sync(0xffffffffu, my_off, src);
            if ((mm >> lane) & 1u) {
                const int v = (A.tile_begin + base + src) * 32 + lane;
                moves[off + __popc(mm & ((1u << lane) - 1u))] = make_int2(v, A.prop_dst[v]);
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads) k_bulk_count(int K, const int* __restrict__ cid, const int2* __restrict__ moves, int n, int* leave_cnt) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int a = cid[moves[i].x];
        if (a < K) atomicAdd(&leave_cnt[a], 1);
    }
}

__global__ void __launch_bounds__(kThreads) k_bulk_apply(ReassignArgs A, BulkArgs B, int stride, const int2* __restrict__ moves, int n) {
    const int K = A.K;
    unsigned n_mods = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int v = moves[i].x, d = moves[i].y;
        const int a = A.cid[v];
        if (a < K && B.leave_cnt[a] >= A.csize[a]) continue;
        const double* it = A.items + (int64_t)v * stride;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            long long f = __double2ll_rn(__ldg(it + k) * B.scale);
            atomicAdd(reinterpret_cast<unsigned long long*>(&B.isum[4 * (int64_t)d + k]), (unsigned long long)f);
            if (a < K) atomicAdd(reinterpret_cast<unsigned long long*>(&B.isum[4 * (int64_t)a + k]), (unsigned long long)(-f));
        }
        atomicAdd(&B.join_cnt[d], 1);
        A.mod_round[d] = A.round;
        if (a < K) A.mod_round[a] = A.round;
        A.cid[v] = d;
        A.prop_dst[v] = a;          // what k_bulk_rollback_moves restores
        mark_tiles_stale(A, v);
        n_mods++;
    }
    warp_count_add(&A.ctr->mods, n_mods);
}

// header exchanged with the moves: [0] local move count, [1] proposals, [2] tests, [3] evaluated, [4] boundary, [5] active tiles
__global__ void k_pack_header(const RoundCounters* ctr, const unsigned long long* round_scalars, const unsigned long long* n_moves,
                              unsigned long long* hdr) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        hdr[0] = *n_moves; hdr[1] = ctr->proposals; hdr[2] = ctr->tests; hdr[3] = ctr->evaluated + ctr->pad[0]; hdr[4] = ctr->boundary;
        hdr[5] = round_scalars[0]; hdr[6] = 0; hdr[7] = 0;
    }
}

}  // namespace acvd
