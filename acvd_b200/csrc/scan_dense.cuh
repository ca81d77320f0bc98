// Dense bulk scan: the frontier scan + bulk decision of a round in which (almost) every tile is active, as a
// TMA-staged streaming kernel (sm_100a).
//
// Same per-vertex work as k_scan<W, true> (reassign.cuh) -- the candidate set of the reference's ProcessOneLoop
// (Common/vtkUniformClustering.h:833-995) over the boundary vertices, decided with the bulk criterion -- but the
// four streams a tile needs (cluster ids, the W ELL columns, positions, and in stage 1 the item weights) are
// contiguous over a run of tiles, so they are moved global -> shared by the TMA engine (cp.async.bulk, completion
// on an mbarrier) in groups of 8 tiles, kept S groups ahead of the warps.  What is left on a warp's critical path
// are the gathers that hit L1/L2 (neighbour cluster ids, modified bits, centroids).
//   * every warp draws the next staged tile from a block-wide ticket counter, so a slow tile never holds the
//     other warps back;
//   * the warp that finishes the last tile of a group refills that stage itself (no producer warp, no "empty"
//     barrier to poll);
//   * the scan writes one 32-bit proposal mask per tile instead of appending to a list: no atomics on the path,
//     and the order of the proposals is fixed.
#pragma once
#include "reassign_types.cuh"

namespace acvd {

constexpr int kDenseWarps = 8;                            // warps per block = tiles per staged group
constexpr int kDenseThreads = 32 * kDenseWarps;
constexpr int kDenseGroupV = 32 * kDenseWarps;            // vertices per group

__host__ __device__ constexpr int dense_stage_bytes(int W) { return kDenseGroupV * (4 + 4 * W + 12 + 8); }
__host__ __device__ constexpr int dense_smem_bytes(int W, int S) { return 128 + S * dense_stage_bytes(W); }

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// blocks until the phase with the given parity has completed; the suspend-time hint lets the hardware park the
// warp instead of spinning through the issue slots the consumers need
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"(0x989680u)
            : "memory");
    } while (!done);
}
// one contiguous global -> shared copy by the TMA engine; `bytes` and both addresses are multiples of 16
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy)
                 : "memory");
}

template <int W>
__device__ __forceinline__ int pick_slot(const int (&nb)[W], int k0) {
    int b = nb[0];
#pragma unroll
    for (int k = 1; k < W; k++) b = (k0 == k) ? nb[k] : b;
    return b;
}

// One block owns a contiguous run of tiles, staged in groups of 8.  S = stages (groups in flight per block),
// MINB = resident blocks per SM the register budget is sized for.
// Requires: cid, xyz and weight allocated up to vpad vertices (whole tiles are copied).
template <int W, int S, int MINB>
__global__ void __launch_bounds__(kDenseThreads, MINB) k_scan_bulk_dense(ReassignArgs A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int GV = kDenseGroupV;
    constexpr int STAGE = dense_stage_bytes(W);
    constexpr int OFF_ELL = 4 * GV, OFF_XYZ = OFF_ELL + 4 * W * GV, OFF_WGT = OFF_XYZ + 12 * GV;
    // control words: full[s] mbarriers at +8 s, per-stage done counters at +64 + 4 s, ticket at +112
    const uint32_t bar0 = smem_addr(smem_raw);
    int* done_cnt = reinterpret_cast<int*>(smem_raw + 64);
    int* ticket = reinterpret_cast<int*>(smem_raw + 112);
    const int lane = threadIdx.x & 31;
    const int n_tiles = A.tile_end - A.tile_begin;
    int chunk = (n_tiles + gridDim.x - 1) / gridDim.x;
    chunk = (chunk + kDenseWarps - 1) / kDenseWarps * kDenseWarps;
    const int t0 = min(A.tile_end, A.tile_begin + (int)blockIdx.x * chunk);
    const int t1 = min(A.tile_end, t0 + chunk);
    const int n_groups = (t1 - t0 + kDenseWarps - 1) / kDenseWarps;
    const bool stage1 = A.bulk_stage == 1;

    // stage group g (tiles t0 + 8 g ...) into its slot: one elected thread arms the barrier and issues the copies
    auto issue_group = [&](int g) {
        uint64_t pol_stream, pol_keep;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
        asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_keep));
        const int s = g % S;
        const int tile = t0 + g * kDenseWarps;
        const uint32_t nv = 32u * (uint32_t)min(kDenseWarps, t1 - tile);
        const int64_t v0 = (int64_t)tile * 32;
        const uint32_t full = bar0 + 8 * s;
        const uint32_t dst = bar0 + 128 + s * STAGE;
        mbar_expect_tx(full, nv * (4u + 4u * W + 12u + (stage1 ? 8u : 0u)));
        bulk_g2s(dst, A.cid + v0, 4 * nv, full, pol_keep);
#pragma unroll
        for (int k = 0; k < W; k++) bulk_g2s(dst + OFF_ELL + 4 * GV * k, A.ell + (int64_t)k * A.vpad + v0, 4 * nv, full, pol_stream);
        bulk_g2s(dst + OFF_XYZ, A.xyz + 3 * v0, 12 * nv, full, pol_stream);
        if (stage1) bulk_g2s(dst + OFF_WGT, A.weight + v0, 8 * nv, full, pol_stream);
    };

    if (threadIdx.x == 0) {
        *ticket = 0;
        for (int s = 0; s < S; s++) { mbar_init(bar0 + 8 * s, 1); done_cnt[s] = 0; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int g = 0; g < S && g < n_groups; g++) issue_group(g);
    }
    __syncthreads();

    const int K = A.K, V = A.V;
    const unsigned* __restrict__ modbits = A.modbits;
    const unsigned lane_lt = (1u << lane) - 1u;
    const bool all_dirty = A.force_all != 0;
    unsigned n_bnd = 0, n_fused = 0, n_tests = 0, n_props = 0;
    const int n_tickets = n_groups * kDenseWarps;
    while (true) {
        int n = 0;
        if (lane == 0) n = atomicAdd(ticket, 1);
        n = __shfl_sync(0xffffffffu, n, 0);
        if (n >= n_tickets) break;
        const int g = n / kDenseWarps, slot = n % kDenseWarps;
        const int s = g % S, it = g / S;
        mbar_wait(bar0 + 8 * s, it & 1);
        const int tile = t0 + n;
        if (tile < t1) {
            const unsigned char* st = smem_raw + 128 + s * STAGE;
            const int* s_cid = reinterpret_cast<const int*>(st);
            const int idx = slot * 32 + lane;
            const int v = tile * 32 + lane;
            const bool valid = v < V;
            const int a = valid ? s_cid[idx] : -1;
            // own cluster's size, centroid and modified bit are needed by every boundary vertex: issued now, they
            // travel together with the neighbour gathers instead of after them
            const int ac = (unsigned)a < (unsigned)K ? a : 0;
            const int a_size = __ldg(A.csize + ac);
            const double4 ca = *reinterpret_cast<const double4*>(A.bulk_cen + 4 * (int64_t)ac);
            const unsigned a_mod = all_dirty ? 1u : (modbits[ac >> 5] >> (ac & 31)) & 1u;
            int nb[W];
#pragma unroll
            for (int k = 0; k < W; k++) nb[k] = reinterpret_cast<const int*>(st + OFF_ELL)[k * GV + idx];
            const bool overflow_row = valid && nb[W - 1] == -2;
            const bool any_overflow = __any_sync(0xffffffffu, nb[W - 1] == -2);
            if (any_overflow) {
                if (overflow_row) nb[W - 1] = A.col[A.row_ptr[v] + W - 1];
                else if (nb[W - 1] == -2) nb[W - 1] = 0;
            }
            // neighbour cluster ids: the ELL pads short rows with the vertex itself, so every slot is a valid index
#pragma unroll
            for (int k = 0; k < W; k++) nb[k] = __ldg(A.cid + nb[k]);
            // rem = slots holding a foreign assigned cluster (the candidates); boundary = any foreign neighbour
            unsigned rem = 0;
            bool bnd = false;
#pragma unroll
            for (int k = 0; k < W; k++) {
                const bool isb = nb[k] != a;
                bnd |= isb;
                rem |= ((isb && (unsigned)nb[k] < (unsigned)K) ? 1u : 0u) << k;
            }
            if (!valid) { rem = 0; bnd = false; }
            // "recently modified" rule (:909-920): own cluster now, candidate clusters inside the candidate loop
            unsigned dirty = all_dirty ? 1u : 0u;
            if (any_overflow) {        // finish long rows from the CSR, warp-uniformly (their decision is k_bulk_evaluate's)
                const int e0 = overflow_row ? A.row_ptr[v] + W : 0, e1 = overflow_row ? A.row_ptr[v + 1] : 0;
                const int steps = __reduce_max_sync(0xffffffffu, e1 - e0);
                for (int q = 0; q < steps; q++) {
                    const int bb = (e0 + q < e1) ? A.cid[A.col[e0 + q]] : a;
                    const bool isb = bb != a;
                    bnd |= isb;
                    if (isb && bb < K) dirty |= (modbits[bb >> 5] >> (bb & 31)) & 1u;
                }
            }
            bnd = bnd && valid;
            n_bnd += bnd ? 1u : 0u;
            const bool cand = bnd && !overflow_row;      // decided here if dirty
            int best_b = -1;
            if (__any_sync(0xffffffffu, bnd)) {
                double px = 0, py = 0, pz = 0, best = 0, w = 0;
                bool blocked = true;
                unsigned ntest = 0;
                if (bnd && a < K) dirty |= a_mod;
                if (!cand) {
                    if (overflow_row && !all_dirty) {    // long row: only the dirty flag of the first W slots is still missing
                        while (rem) {
                            const int b = pick_slot<W>(nb, __ffs(rem) - 1);
                            rem &= rem - 1;
                            dirty |= (modbits[b >> 5] >> (b & 31)) & 1u;
                        }
                    }
                    rem = 0;
                } else if (a >= K) {   // NULL cluster: adopt the first assigned neighbour cluster; dirty if any neighbour cluster is
                    if (rem) best_b = pick_slot<W>(nb, __ffs(rem) - 1);
                    if (!all_dirty) {
                        while (rem) {
                            const int b = pick_slot<W>(nb, __ffs(rem) - 1);
                            rem &= rem - 1;
                            dirty |= (modbits[b >> 5] >> (b & 31)) & 1u;
                        }
                    }
                    rem = 0;
                } else {
                    const float* xs = reinterpret_cast<const float*>(st + OFF_XYZ) + 3 * idx;
                    px = xs[0]; py = xs[1]; pz = xs[2];
                    blocked = a_size == 1;
                    const double dx = px - ca.x, dy = py - ca.y, dz = pz - ca.z;
                    best = dx * dx + dy * dy + dz * dz;
                    if (stage1) {
                        w = reinterpret_cast<const double*>(st + OFF_WGT)[idx];
                        best = ca.w / (ca.w - w) * best;
                    }
                }
                // candidates in slot order (first occurrence of every distinct cluster)
                while (__any_sync(0xffffffffu, rem != 0)) {
                    if (rem) {
                        const int b = pick_slot<W>(nb, __ffs(rem) - 1);
#pragma unroll
                        for (int k = 0; k < W; k++) rem &= ~((nb[k] == b ? 1u : 0u) << k);
                        ntest++;
                        dirty |= (modbits[b >> 5] >> (b & 31)) & 1u;
                        if (!blocked) {
                            const double4 cb = *reinterpret_cast<const double4*>(A.bulk_cen + 4 * (int64_t)b);
                            const double dx = px - cb.x, dy = py - cb.y, dz = pz - cb.z;
                            double d = dx * dx + dy * dy + dz * dz;
                            if (stage1) d = cb.w / (cb.w + w) * d;
                            if (d < best) { best = d; best_b = b; }
                        }
                    }
                }
                // a vertex none of whose clusters changed keeps its earlier outcome: it is neither counted nor proposed
                if (!(cand && dirty)) best_b = -1;
                else { n_fused++; n_tests += ntest; }
                if (best_b >= 0) {
                    A.prop_dst[v] = best_b;
                    // plain RED: the compiler's warp-aggregation loop around atomicAdd costs more than the atomics it saves
                    if (a < K && A.bulk_count_leave) asm volatile("red.global.add.s32 [%0], 1;" ::"l"(A.bulk_leave + a) : "memory");
                }
            }
            const unsigned mp = __ballot_sync(0xffffffffu, best_b >= 0);
            if (lane == 0) { A.prop_mask[tile] = mp; n_props += __popc(mp); }
            // rows longer than W (rare) are decided by k_bulk_evaluate from the work list
            if (any_overflow) {
                const bool ow = bnd && overflow_row && dirty != 0;
                const unsigned mw = __ballot_sync(0xffffffffu, ow);
                if (mw) {
                    int basew = 0;
                    if (lane == 0) basew = (int)atomicAdd(&A.ctr->evaluated, (unsigned long long)__popc(mw));
                    basew = __shfl_sync(0xffffffffu, basew, 0);
                    if (ow) A.work[basew + __popc(mw & lane_lt)] = v;
                }
            }
        }
        // the warp that finishes the last tile of the group refills the stage with the group S ahead
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            if (atomicAdd(&done_cnt[s], 1) == kDenseWarps - 1) {
                done_cnt[s] = 0;
                if (g + S < n_groups) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    issue_group(g + S);
                }
            }
        }
    }
    warp_count_add(&A.ctr->boundary, n_bnd);
    warp_count_add(&A.ctr->pad[0], n_fused);
    warp_count_add(&A.ctr->tests, n_tests);
    warp_count_add(&A.ctr->proposals, n_props);
}

// ---------------------------------------------------------------------------------------------------------------
// Third generation (the second -- static tile assignment, "empty" mbarrier hand-back, two tiles in flight per warp by software
// pipelining -- measured 950 us against 706 us and was removed): an instruction diet for the common tile.
//   * static tile assignment (warp w owns slot w of every staged group), stage hand-back through an "empty" mbarrier;
//   * candidates without a slot loop: on a triangulated surface a boundary vertex sees one or two foreign clusters,
//     which are the minimum and the maximum of its neighbours' cluster ids (one of them may be its own).  Both
//     centroids are requested together and compared in straight-line code -- no pick/dedup loop, no dependent
//     iterations.  An exact tie between the two is resolved by first occurrence in slot order, as the loop would;
//   * one 32-bit word per cluster carries its size and its "modified in the previous round" flag (cmeta, written by
//     k_modbits): one load per cluster instead of a size load plus a bit look-up;
//   * tiles that hold anything else -- a third foreign cluster around some vertex, the NULL cluster, a row longer
//     than W, the partial last tile -- take the generic path (dense_tile_generic: the loop of the first generation,
//     out of line so that its registers do not weigh on the common path).
// Decisions, counters and proposal masks are those of k_scan_bulk_dense / k_scan<W, true>.
struct DenseCounters { unsigned bnd, fused, tests, props; };

template <int W>
__device__ __noinline__ void dense_tile_generic(const ReassignArgs& A, const unsigned char* st, int slot, int tile, DenseCounters& C) {
    constexpr int GV = kDenseGroupV;
    constexpr int OFF_ELL = 4 * GV, OFF_XYZ = OFF_ELL + 4 * W * GV, OFF_WGT = OFF_XYZ + 12 * GV;
    const int lane = threadIdx.x & 31;
    const int K = A.K, V = A.V;
    const unsigned* __restrict__ modbits = A.modbits;
    const unsigned lane_lt = (1u << lane) - 1u;
    const bool all_dirty = A.force_all != 0;
    const bool stage1 = A.bulk_stage == 1;
    const int* s_cid = reinterpret_cast<const int*>(st);
    const int idx = slot * 32 + lane;
    const int v = tile * 32 + lane;
    const bool valid = v < V;
    const int a = valid ? s_cid[idx] : -1;
    const int ac = (unsigned)a < (unsigned)K ? a : 0;
    const int a_size = __ldg(A.csize + ac);
    const double4 ca = *reinterpret_cast<const double4*>(A.bulk_cen + 4 * (int64_t)ac);
    const unsigned a_mod = all_dirty ? 1u : (modbits[ac >> 5] >> (ac & 31)) & 1u;
    int nb[W];
#pragma unroll
    for (int k = 0; k < W; k++) nb[k] = reinterpret_cast<const int*>(st + OFF_ELL)[k * GV + idx];
    const bool overflow_row = valid && nb[W - 1] == -2;
    const bool any_overflow = __any_sync(0xffffffffu, nb[W - 1] == -2);
    if (any_overflow) {
        if (overflow_row) nb[W - 1] = A.col[A.row_ptr[v] + W - 1];
        else if (nb[W - 1] == -2) nb[W - 1] = 0;
    }
#pragma unroll
    for (int k = 0; k < W; k++) nb[k] = __ldg(A.cid + nb[k]);
    unsigned rem = 0;
    bool bnd = false;
#pragma unroll
    for (int k = 0; k < W; k++) {
        const bool isb = nb[k] != a;
        bnd |= isb;
        rem |= ((isb && (unsigned)nb[k] < (unsigned)K) ? 1u : 0u) << k;
    }
    if (!valid) { rem = 0; bnd = false; }
    unsigned dirty = all_dirty ? 1u : 0u;
    if (any_overflow) {        // finish long rows from the CSR, warp-uniformly (their decision is k_bulk_evaluate's)
        const int e0 = overflow_row ? A.row_ptr[v] + W : 0, e1 = overflow_row ? A.row_ptr[v + 1] : 0;
        const int steps = __reduce_max_sync(0xffffffffu, e1 - e0);
        for (int q = 0; q < steps; q++) {
            const int bb = (e0 + q < e1) ? A.cid[A.col[e0 + q]] : a;
            const bool isb = bb != a;
            bnd |= isb;
            if (isb && bb < K) dirty |= (modbits[bb >> 5] >> (bb & 31)) & 1u;
        }
    }
    bnd = bnd && valid;
    C.bnd += bnd ? 1u : 0u;
    const bool cand = bnd && !overflow_row;      // decided here if dirty
    int best_b = -1;
    if (__any_sync(0xffffffffu, bnd)) {
        double px = 0, py = 0, pz = 0, best = 0, w = 0;
        bool blocked = true;
        unsigned ntest = 0;
        if (bnd && a < K) dirty |= a_mod;
        if (!cand) {
            if (overflow_row && !all_dirty) {    // long row: only the dirty flag of the first W slots is still missing
                while (rem) {
                    const int b = pick_slot<W>(nb, __ffs(rem) - 1);
                    rem &= rem - 1;
                    dirty |= (modbits[b >> 5] >> (b & 31)) & 1u;
                }
            }
            rem = 0;
        } else if (a >= K) {   // NULL cluster: adopt the first assigned neighbour cluster; dirty if any neighbour cluster is
            if (rem) best_b = pick_slot<W>(nb, __ffs(rem) - 1);
            if (!all_dirty) {
                while (rem) {
                    const int b = pick_slot<W>(nb, __ffs(rem) - 1);
                    rem &= rem - 1;
                    dirty |= (modbits[b >> 5] >> (b & 31)) & 1u;
                }
            }
            rem = 0;
        } else {
            const float* xs = reinterpret_cast<const float*>(st + OFF_XYZ) + 3 * idx;
            px = xs[0]; py = xs[1]; pz = xs[2];
            blocked = a_size == 1;
            const double dx = px - ca.x, dy = py - ca.y, dz = pz - ca.z;
            best = dx * dx + dy * dy + dz * dz;
            if (stage1) {
                w = reinterpret_cast<const double*>(st + OFF_WGT)[idx];
                best = ca.w / (ca.w - w) * best;
            }
        }
        // candidates in slot order (first occurrence of every distinct cluster)
        while (__any_sync(0xffffffffu, rem != 0)) {
            if (rem) {
                const int b = pick_slot<W>(nb, __ffs(rem) - 1);
#pragma unroll
                for (int k = 0; k < W; k++) rem &= ~((nb[k] == b ? 1u : 0u) << k);
                ntest++;
                dirty |= (modbits[b >> 5] >> (b & 31)) & 1u;
                if (!blocked) {
                    const double4 cb = *reinterpret_cast<const double4*>(A.bulk_cen + 4 * (int64_t)b);
                    const double dx = px - cb.x, dy = py - cb.y, dz = pz - cb.z;
                    double d = dx * dx + dy * dy + dz * dz;
                    if (stage1) d = cb.w / (cb.w + w) * d;
                    if (d < best) { best = d; best_b = b; }
                }
            }
        }
        // a vertex none of whose clusters changed keeps its earlier outcome: it is neither counted nor proposed
        if (!(cand && dirty)) best_b = -1;
        else { C.fused++; C.tests += ntest; }
        if (best_b >= 0) {
            A.prop_dst[v] = best_b;
            if (a < K && A.bulk_count_leave) asm volatile("red.global.add.s32 [%0], 1;" ::"l"(A.bulk_leave + a) : "memory");
        }
    }
    const unsigned mp = __ballot_sync(0xffffffffu, best_b >= 0);
    if (lane == 0) { A.prop_mask[tile] = mp; C.props += __popc(mp); }
    // rows longer than W (rare) are decided by k_bulk_evaluate from the work list
    if (any_overflow) {
        const bool ow = bnd && overflow_row && dirty != 0;
        const unsigned mw = __ballot_sync(0xffffffffu, ow);
        if (mw) {
            int basew = 0;
            if (lane == 0) basew = (int)atomicAdd(&A.ctr->evaluated, (unsigned long long)__popc(mw));
            basew = __shfl_sync(0xffffffffu, basew, 0);
            if (ow) A.work[basew + __popc(mw & lane_lt)] = v;
        }
    }
}

template <int W, int S, int MINB, bool STAGE1, bool STATIC, int PF, int DBG = 0>
__global__ void __launch_bounds__(kDenseThreads, MINB) k_scan_bulk_dense3(const __grid_constant__ ReassignArgs A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int GV = kDenseGroupV;
    constexpr int STAGE = dense_stage_bytes(W);
    constexpr int OFF_ELL = 4 * GV, OFF_XYZ = OFF_ELL + 4 * W * GV, OFF_WGT = OFF_XYZ + 12 * GV;
    // control words: full[s] mbarriers at +8 s; STATIC: empty[s] mbarriers at +64 + 8 s; else per-stage done counters at
    // +64 + 4 s and the ticket at +112
    const uint32_t bar0 = smem_addr(smem_raw);
    int* done_cnt = reinterpret_cast<int*>(smem_raw + 64);
    int* ticket = reinterpret_cast<int*>(smem_raw + 112);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_tiles = A.tile_end - A.tile_begin;
    int chunk = (n_tiles + gridDim.x - 1) / gridDim.x;
    chunk = (chunk + kDenseWarps - 1) / kDenseWarps * kDenseWarps;
    const int t0 = min(A.tile_end, A.tile_begin + (int)blockIdx.x * chunk);
    const int t1 = min(A.tile_end, t0 + chunk);
    const int n_groups = (t1 - t0 + kDenseWarps - 1) / kDenseWarps;
    constexpr bool stage1 = STAGE1;

    auto issue_group = [&](int g) {
        uint64_t pol_stream, pol_keep;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
        asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_keep));
        const int s = g % S;
        const int tile = t0 + g * kDenseWarps;
        const uint32_t nv = 32u * (uint32_t)min(kDenseWarps, t1 - tile);
        const int64_t v0 = (int64_t)tile * 32;
        const uint32_t full = bar0 + 8 * s;
        const uint32_t dst = bar0 + 128 + s * STAGE;
        mbar_expect_tx(full, nv * (4u + 4u * W + 12u + (stage1 ? 8u : 0u)));
        bulk_g2s(dst, A.cid + v0, 4 * nv, full, pol_keep);
#pragma unroll
        for (int k = 0; k < W; k++) bulk_g2s(dst + OFF_ELL + 4 * GV * k, A.ell + (int64_t)k * A.vpad + v0, 4 * nv, full, pol_stream);
        bulk_g2s(dst + OFF_XYZ, A.xyz + 3 * v0, 12 * nv, full, pol_stream);
        if (stage1) bulk_g2s(dst + OFF_WGT, A.weight + v0, 8 * nv, full, pol_stream);
    };

    if (threadIdx.x == 0) {
        if (!STATIC) *ticket = 0;
        for (int s = 0; s < S; s++) {
            mbar_init(bar0 + 8 * s, 1);
            if (STATIC) mbar_init(bar0 + 64 + 8 * s, kDenseWarps); else done_cnt[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int g = 0; g < S && g < n_groups; g++) issue_group(g);
    }
    __syncthreads();

    const int K = A.K;
    const int* __restrict__ cmeta = A.cmeta;
    const int dirty_all = A.force_all != 0 ? (int)0x80000000 : 0;
    const int first_partial_tile = A.V >> 5;        // the tile that holds vertices >= V (if V is not a multiple of 32)
    unsigned n_bnd = 0, n_fused = 0, n_tests = 0, n_props = 0;
    const int n_tickets = n_groups * kDenseWarps;
    int n = STATIC ? warp - kDenseWarps : 0;
    while (true) {
        if (STATIC) n += kDenseWarps;
        else {
            if (lane == 0) n = atomicAdd(ticket, 1);
            n = __shfl_sync(0xffffffffu, n, 0);
        }
        if (n >= n_tickets) break;
        const int g = n / kDenseWarps, slot = n % kDenseWarps;
        const int s = g % S, ph = (g / S) & 1;
        mbar_wait(bar0 + 8 * s, ph);
        const int tile = t0 + n;
        if (tile < t1) {
            const unsigned char* st = smem_raw + 128 + s * STAGE;
            const int idx = slot * 32 + lane;
            const int a = reinterpret_cast<const int*>(st)[idx];
            int nb[W];
#pragma unroll
            for (int k = 0; k < W; k++) nb[k] = reinterpret_cast<const int*>(st + OFF_ELL)[k * GV + idx];
            const bool long_row = nb[W - 1] < 0;            // -2 marks a row longer than W
            nb[W - 1] = max(nb[W - 1], 0);
            if (DBG != 2) {
#pragma unroll
                for (int k = 0; k < W; k++) nb[k] = __ldg(A.cid + nb[k]);
            }
            if (PF > 0 && g + 1 < n_groups) {
                // While the gathers are in flight: send for what the same slot of the NEXT group will gather (its neighbours'
                // cluster ids; PF > 1: its own cluster's word and centroid), so that whichever warp draws that tile finds
                // them in L1 -- the kernel is bound by the chain of L2 round trips of a tile, not by issue slots or HBM.
                // Only if that group has already landed; indices are range-checked (a stage may be refilled under us).
                const int s2 = (g + 1) % S;
                uint32_t ready;
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ready) : "r"(bar0 + 8 * s2), "r"((uint32_t)(((g + 1) / S) & 1)) : "memory");
                if (ready) {
                    const unsigned char* st2 = smem_raw + 128 + s2 * STAGE;
#pragma unroll
                    for (int k = 0; k < W; k++) {
                        const unsigned u = reinterpret_cast<const unsigned*>(st2 + OFF_ELL)[k * GV + idx];
                        if (u < (unsigned)A.V) asm volatile("prefetch.global.L1 [%0];" ::"l"(A.cid + u));
                    }
                    if (PF > 1) {
                        const unsigned a2 = reinterpret_cast<const unsigned*>(st2)[idx];
                        if (a2 < (unsigned)K) {
                            asm volatile("prefetch.global.L1 [%0];" ::"l"(cmeta + a2));
                            asm volatile("prefetch.global.L1 [%0];" ::"l"(A.bulk_cen + 4 * (int64_t)a2));
                        }
                    }
                }
            }
            int m1 = nb[0], M1 = nb[0];
#pragma unroll
            for (int k = 1; k < W; k++) { m1 = min(m1, nb[k]); M1 = max(M1, nb[k]); }
            if (DBG != 0) {      // diagnostic variants (acvd_bench_kernel only): 1 = no decision, 2 = no gathers either
                const bool bnd = (m1 != a) || (M1 != a);
                n_bnd += bnd ? 1u : 0u;
                if (lane == 0) A.prop_mask[tile] = 0;
            } else if (__any_sync(0xffffffffu, long_row || max(M1, a) >= K) || tile >= first_partial_tile) {
                // the NULL cluster, a row longer than W or the partial last tile: the generic tile code, out of line
                DenseCounters T{0u, 0u, 0u, 0u};
                dense_tile_generic<W>(A, st, slot, tile, T);
                n_bnd += T.bnd; n_fused += T.fused; n_tests += T.tests; n_props += T.props;
            } else {
                const bool bnd = (m1 != a) || (M1 != a);
                n_bnd += bnd ? 1u : 0u;
                int best_b = -1;
                if (bnd) {
                    // a neighbour cluster strictly between the two extremes that is not the own one: a third candidate
                    const int lo1 = m1 + 1;
                    const unsigned span = (unsigned)(M1 - lo1);
                    bool third = false;
#pragma unroll
                    for (int k = 0; k < W; k++) third |= ((unsigned)(nb[k] - lo1) < span) && (nb[k] != a);
                    const int meta_a = __ldg(cmeta + a);
                    const double* pa = A.bulk_cen + 4 * (int64_t)a;
                    const float* xs = reinterpret_cast<const float*>(st + OFF_XYZ) + 3 * idx;
                    if (!third) {
                        const int c1 = (m1 != a) ? m1 : M1;
                        const bool two = (m1 != a) && (M1 != a) && (m1 != M1);
                        const int c2 = M1;
                        // the clusters' words (size, modified flag) and the first two centroids are requested together; the
                        // third centroid is sent for (to L1) now and read once the first comparison is done
                        const int meta1 = __ldg(cmeta + c1);
                        const int meta2 = two ? __ldg(cmeta + c2) : 0;
                        const double* p1 = A.bulk_cen + 4 * (int64_t)c1;
                        const double* p2 = A.bulk_cen + 4 * (int64_t)c2;
                        if (two) asm volatile("prefetch.global.L1 [%0];" ::"l"(p2));
                        if ((dirty_all | meta_a | meta1 | meta2) < 0) {          // own or an adjacent cluster modified (:909-920)
                            n_fused++;
                            n_tests += two ? 2u : 1u;
                            if ((meta_a & 0x7fffffff) != 1) {                    // a cluster is never emptied
                                const double4 ca = *reinterpret_cast<const double4*>(pa);
                                const double4 cb1 = *reinterpret_cast<const double4*>(p1);
                                const double px = xs[0], py = xs[1], pz = xs[2];
                                double dx = px - ca.x, dy = py - ca.y, dz = pz - ca.z;
                                double best = dx * dx + dy * dy + dz * dz;
                                double w = 0;
                                if (stage1) {
                                    w = reinterpret_cast<const double*>(st + OFF_WGT)[idx];
                                    best = ca.w / (ca.w - w) * best;
                                }
                                dx = px - cb1.x; dy = py - cb1.y; dz = pz - cb1.z;
                                double d1 = dx * dx + dy * dy + dz * dz;
                                if (stage1) d1 = cb1.w / (cb1.w + w) * d1;
                                if (two) {
                                    const double4 cb2 = *reinterpret_cast<const double4*>(p2);
                                    dx = px - cb2.x; dy = py - cb2.y; dz = pz - cb2.z;
                                    double d2 = dx * dx + dy * dy + dz * dz;
                                    if (stage1) d2 = cb2.w / (cb2.w + w) * d2;
                                    // candidates are taken in slot order with a strict comparison: on an exact tie the one
                                    // that occurs first among the neighbours wins
                                    bool first_is_c1 = true;
                                    if (d1 == d2) {
#pragma unroll
                                        for (int k = W - 1; k >= 0; k--) { if (nb[k] == c1) first_is_c1 = true; else if (nb[k] == c2) first_is_c1 = false; }
                                    }
                                    const bool take2 = first_is_c1 ? (d2 < d1) : (d2 <= d1);
                                    const double dm = take2 ? d2 : d1;
                                    if (dm < best) best_b = take2 ? c2 : c1;
                                } else if (d1 < best) best_b = c1;
                            }
                        }
                    } else {
                        // three or more foreign clusters around the vertex (rare): the slot-order walk, for these lanes only
                        unsigned rem = 0;
#pragma unroll
                        for (int k = 0; k < W; k++) rem |= (nb[k] != a ? 1u : 0u) << k;
                        int dirty = dirty_all | meta_a;
                        const bool blocked = (meta_a & 0x7fffffff) == 1;
                        const double px = xs[0], py = xs[1], pz = xs[2];
                        double best = 0, w = 0;
                        if (!blocked) {
                            const double4 ca = *reinterpret_cast<const double4*>(pa);
                            const double dx = px - ca.x, dy = py - ca.y, dz = pz - ca.z;
                            best = dx * dx + dy * dy + dz * dz;
                            if (stage1) {
                                w = reinterpret_cast<const double*>(st + OFF_WGT)[idx];
                                best = ca.w / (ca.w - w) * best;
                            }
                        }
                        unsigned ntest = 0;
                        while (rem) {
                            const int b = pick_slot<W>(nb, __ffs(rem) - 1);
#pragma unroll
                            for (int k = 0; k < W; k++) rem &= ~((nb[k] == b ? 1u : 0u) << k);
                            ntest++;
                            dirty |= __ldg(cmeta + b);
                            if (!blocked) {
                                const double4 cb = *reinterpret_cast<const double4*>(A.bulk_cen + 4 * (int64_t)b);
                                const double dx = px - cb.x, dy = py - cb.y, dz = pz - cb.z;
                                double d = dx * dx + dy * dy + dz * dz;
                                if (stage1) d = cb.w / (cb.w + w) * d;
                                if (d < best) { best = d; best_b = b; }
                            }
                        }
                        if (dirty < 0) { n_fused++; n_tests += ntest; } else best_b = -1;
                    }
                    if (best_b >= 0) {
                        A.prop_dst[tile * 32 + lane] = best_b;
                        if (A.bulk_count_leave) asm volatile("red.global.add.s32 [%0], 1;" ::"l"(A.bulk_leave + a) : "memory");
                    }
                }
                const unsigned mp = __ballot_sync(0xffffffffu, best_b >= 0);
                if (lane == 0) { A.prop_mask[tile] = mp; n_props += __popc(mp); }
            }
        }
        __syncwarp();
        if (STATIC) {
            // hand the stage back: the 8 warps arrive, warp (g mod 8) waits for all of them and refills it with group g + S
            if (lane == 0) mbar_arrive(bar0 + 64 + 8 * s);
            if (g + S < n_groups && warp == (g % kDenseWarps)) {
                if (lane == 0) {
                    mbar_wait(bar0 + 64 + 8 * s, ph);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    issue_group(g + S);
                }
                __syncwarp();
            }
        } else if (lane == 0) {
            // the warp that finishes the last tile of the group refills the stage with the group S ahead
            __threadfence_block();
            if (atomicAdd(&done_cnt[s], 1) == kDenseWarps - 1) {
                done_cnt[s] = 0;
                if (g + S < n_groups) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    issue_group(g + S);
                }
            }
        }
    }
    warp_count_add(&A.ctr->boundary, n_bnd);
    warp_count_add(&A.ctr->pad[0], n_fused);
    warp_count_add(&A.ctr->tests, n_tests);
    warp_count_add(&A.ctr->proposals, n_props);
}


// ---------------------------------------------------------------------------------------------------------------
// Split form of the dense bulk scan.  The fused kernels above are bound by the latency chain of a tile (stream ->
// neighbour cluster ids -> centroids -> decision) at the 32 warps per SM their registers allow: measured on C4 (one
// launch, 40 M vertices) streaming alone runs at the HBM peak (237 us), the gathers add 184 us and the decision 290 us,
// and neither a leaner decision nor prefetching moves the total.  So the two halves get the shape each one wants:
//
//   k_scan_classify  the frontier scan proper: streams cluster ids + ELL columns (28 B per vertex, TMA-staged as above),
//                    gathers the neighbours' cluster ids, finds the boundary vertices and their (one or two) distinct
//                    foreign clusters in slot order, applies the "recently modified" rule (:909-920) through the 50 KB
//                    bitmap (L1-resident) and appends one 16-byte record (vertex, own cluster, one or two candidates) per
//                    DIRTY boundary vertex to the block's segment of a list.  Few registers, no fp64.
//   k_bulk_decide    the bulk decision over the list: every lane is a vertex that needs one, positions / centroids are
//                    requested together, candidates compared in slot order -- no divergence on "is this a boundary
//                    vertex", and the work shrinks with the dirty set (15.8 M -> 3 M vertices over the C4 bulk rounds).
//
// Vertices with more than two foreign clusters, the NULL cluster in their ring, a row longer than W or the NULL cluster as
// own cluster (all rare) go to the work list of k_bulk_evaluate, as rows longer than W do in the fused kernels.  Decisions, counters and proposal
// masks are those of the fused kernels and of k_scan<W, true>.
__host__ __device__ constexpr int classify_stage_bytes(int W, int VPL) { return kDenseGroupV * VPL * (4 + 4 * W); }
__host__ __device__ constexpr int classify_smem_bytes(int W, int S, int VPL) { return 128 + S * classify_stage_bytes(W, VPL); }
// tiles per scanning block (a multiple of the 8 VPL tiles of a staged group): shared by the two kernels
__host__ __device__ inline int classify_chunk(int n_tiles, int grid, int vpl) {
    int chunk = (n_tiles + grid - 1) / grid;
    const int q = kDenseWarps * vpl;
    return (chunk + q - 1) / q * q;
}

constexpr int kClassifyPush = 256;       // per-warp staging of work-list entries (work-list form of k_scan_classify)

__device__ __forceinline__ unsigned mod_bit(const unsigned* __restrict__ modbits, int c) { return (__ldg(modbits + (c >> 5)) >> (c & 31)) & 1u; }

// VPL = 32-vertex tiles per warp and ticket: the gathers of all of them are in flight together (the kernel is bound by
// the latency chain of a tile, so memory-level parallelism per warp is what counts), and the per-ticket overhead is shared.
// WORKLIST = true: the opening round of an exact phase (everything is dirty, no stored proposal survives): the same scan, but
// every boundary vertex goes to the work list of k_evaluate instead of the candidate list (replaces k_scan<W, false> there).
template <int W, int S, int MINB, int VPL, bool WORKLIST = false>
__global__ void __launch_bounds__(kDenseThreads, MINB) k_scan_classify(const __grid_constant__ ReassignArgs A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int GV = kDenseGroupV * VPL;                    // vertices per staged group
    constexpr int GT = kDenseWarps * VPL;                     // tiles per staged group
    constexpr int STAGE = classify_stage_bytes(W, VPL);
    constexpr int OFF_ELL = 4 * GV;
    // control words: full[s] mbarriers at +8 s, per-stage done counters at +64 + 4 s, ticket at +112, list cursor at +116
    const uint32_t bar0 = smem_addr(smem_raw);
    int* done_cnt = reinterpret_cast<int*>(smem_raw + 64);
    int* ticket = reinterpret_cast<int*>(smem_raw + 112);
    int* cursor = reinterpret_cast<int*>(smem_raw + 116);
    const int lane = threadIdx.x & 31;
    const int n_tiles = A.tile_end - A.tile_begin;
    const int chunk = classify_chunk(n_tiles, gridDim.x, VPL);
    const int t0 = min(A.tile_end, A.tile_begin + (int)blockIdx.x * chunk);
    const int t1 = min(A.tile_end, t0 + chunk);
    const int n_groups = (t1 - t0 + GT - 1) / GT;
    int4* seg = A.blist + (int64_t)(t0 - A.tile_begin) * 32;

    auto issue_group = [&](int g) {
        uint64_t pol_stream, pol_keep;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
        asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_keep));
        const int s = g % S;
        const int tile = t0 + g * GT;
        const uint32_t nv = 32u * (uint32_t)min(GT, t1 - tile);
        const int64_t v0 = (int64_t)tile * 32;
        const uint32_t full = bar0 + 8 * s;
        const uint32_t dst = bar0 + 128 + s * STAGE;
        mbar_expect_tx(full, nv * (4u + 4u * W));
        bulk_g2s(dst, A.cid + v0, 4 * nv, full, pol_keep);
#pragma unroll
        for (int k = 0; k < W; k++) bulk_g2s(dst + OFF_ELL + 4 * GV * k, A.ell + (int64_t)k * A.vpad + v0, 4 * nv, full, pol_stream);
    };

    if (threadIdx.x == 0) {
        *ticket = 0; *cursor = 0;
        for (int s = 0; s < S; s++) { mbar_init(bar0 + 8 * s, 1); done_cnt[s] = 0; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int g = 0; g < S && g < n_groups; g++) issue_group(g);
    }
    __syncthreads();

    const int K = A.K, V = A.V;
    const unsigned* __restrict__ modbits = A.modbits;
    const unsigned lane_lt = (1u << lane) - 1u;
    const bool all_dirty = A.force_all != 0;
    unsigned n_bnd = 0, n_listed = 0;
    const int n_tickets = n_groups * kDenseWarps;
    // work-list form: a third of the vertices go to the list; a global atomic per tile (1.25 M on one address at C4) would
    // bound the kernel, so every warp stages its entries in shared memory and reserves space once per ~200 entries
    __shared__ int s_push[WORKLIST ? kDenseWarps * kClassifyPush : 1];
    int* wbuf = s_push + (WORKLIST ? (threadIdx.x >> 5) * kClassifyPush : 0);
    int wcnt = 0;                                   // warp-uniform
    auto flush_work = [&]() {
        if (wcnt == 0) return;
        int base = 0;
        if (lane == 0) base = (int)atomicAdd(&A.ctr->evaluated, (unsigned long long)wcnt);
        base = __shfl_sync(0xffffffffu, base, 0);
        __syncwarp();
        for (int i = lane; i < wcnt; i += 32) A.work[base + i] = wbuf[i];
        __syncwarp();
        wcnt = 0;
    };
    while (true) {
        int n = 0;
        if (lane == 0) n = atomicAdd(ticket, 1);
        n = __shfl_sync(0xffffffffu, n, 0);
        if (n >= n_tickets) break;
        const int g = n / kDenseWarps, slot = n % kDenseWarps;
        const int s = g % S;
        mbar_wait(bar0 + 8 * s, (g / S) & 1);
        const unsigned char* st = smem_raw + 128 + s * STAGE;
        // all the tiles of the ticket: ids from the staged group, then every gather in flight before the first is used
        int av[VPL], nbv[VPL][W];
#pragma unroll
        for (int j = 0; j < VPL; j++) {
            const int idx = (slot * VPL + j) * 32 + lane;
            av[j] = reinterpret_cast<const int*>(st)[idx];
#pragma unroll
            for (int k = 0; k < W; k++) nbv[j][k] = reinterpret_cast<const int*>(st + OFF_ELL)[k * GV + idx];
        }
        bool longv[VPL];
#pragma unroll
        for (int j = 0; j < VPL; j++) {
            const bool in_range = t0 + n * VPL + j < t1;    // (the last group of the block may be partly filled: stale indices)
            longv[j] = nbv[j][W - 1] < 0;                   // -2 marks a row longer than W
            nbv[j][W - 1] = max(nbv[j][W - 1], 0);
#pragma unroll
            for (int k = 0; k < W; k++) nbv[j][k] = in_range ? __ldg(A.cid + nbv[j][k]) : 0;     // short rows are padded with the vertex itself
        }
        // own cluster's modified bit: asked for now, it travels with the gathers
        unsigned amod[VPL];
#pragma unroll
        for (int j = 0; j < VPL; j++) amod[j] = all_dirty ? 1u : mod_bit(modbits, min(max(av[j], 0), K - 1));
#pragma unroll
        for (int j = 0; j < VPL; j++) {
            const int tile = t0 + n * VPL + j;
            if (tile >= t1) break;
            const int v = tile * 32 + lane;
            const bool valid = v < V;
            const int a = av[j];
            int nb[W];
#pragma unroll
            for (int k = 0; k < W; k++) nb[k] = nbv[j][k];
            const bool long_row = valid && longv[j];
            // extremes of the FOREIGN neighbour clusters: one foreign cluster -> equal, two -> nothing strictly between them
            int mf = 0x7fffffff, Mf = -1;
#pragma unroll
            for (int k = 0; k < W; k++) {
                const bool f = nb[k] != a;
                mf = min(mf, f ? nb[k] : 0x7fffffff);
                Mf = max(Mf, f ? nb[k] : -1);
            }
            bool bnd = valid && Mf >= 0;
            int c1 = -1, c2 = -1;
            bool listed = false, to_work = false;
            if (long_row || (valid && a >= K)) {
                // own cluster NULL or a row longer than W: boundary / dirty over the whole CSR row, decided by k_bulk_evaluate
                const int e0 = A.row_ptr[v], e1 = A.row_ptr[v + 1];
                unsigned dirty = all_dirty ? 1u : 0u;
                bnd = false;
                for (int e = e0; e < e1; e++) {
                    const int b = A.cid[A.col[e]];
                    if (b != a) { bnd = true; if (b < K) dirty |= mod_bit(modbits, b); }
                }
                if (a < K) dirty |= mod_bit(modbits, a);
                to_work = bnd && dirty != 0;
            } else if (bnd) {
                // a foreign cluster strictly between the two extremes, or the NULL cluster around: more than two candidates
                // or none -- rare, left to k_bulk_evaluate (dirty over all the distinct foreign clusters)
                const unsigned span = (unsigned)(Mf - mf - 1);
                bool odd = Mf >= K;
#pragma unroll
                for (int k = 0; k < W; k++) odd |= ((unsigned)(nb[k] - mf - 1) < span) && (nb[k] != a);
                unsigned dirty = amod[j];
                if (!odd) {
                    // one or two foreign clusters, first occurrence first
                    int first = nb[W - 1];
#pragma unroll
                    for (int k = W - 2; k >= 0; k--) first = (nb[k] != a) ? nb[k] : first;
                    c1 = first;
                    const bool two = mf != Mf;
                    if (two) c2 = (first == mf) ? Mf : mf;
                    if (!all_dirty) { dirty |= mod_bit(modbits, c1); if (two) dirty |= mod_bit(modbits, c2); }
                    listed = dirty != 0;
                } else {
                    if (!all_dirty) {
#pragma unroll
                        for (int k = 0; k < W; k++) if (nb[k] != a && nb[k] < K) dirty |= mod_bit(modbits, nb[k]);
                    }
                    to_work = dirty != 0;
                }
            }
            n_bnd += bnd ? 1u : 0u;
            if (WORKLIST) { to_work = to_work || listed; listed = false; }
            const unsigned ml = WORKLIST ? 0u : __ballot_sync(0xffffffffu, listed);
            if (ml) {
                int base = 0;
                if (lane == 0) base = atomicAdd(cursor, __popc(ml));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (listed) seg[base + __popc(ml & lane_lt)] = make_int4(v, a, c1, c2);
                n_listed += listed ? 1u : 0u;
            }
            const unsigned mw = __ballot_sync(0xffffffffu, to_work);
            if (mw) {
                if (WORKLIST) {
                    if (to_work) wbuf[wcnt + __popc(mw & lane_lt)] = v;
                    wcnt += __popc(mw);
                    if (wcnt > kClassifyPush - 32) flush_work();
                } else {
                    int basew = 0;
                    if (lane == 0) basew = (int)atomicAdd(&A.ctr->evaluated, (unsigned long long)__popc(mw));
                    basew = __shfl_sync(0xffffffffu, basew, 0);
                    if (to_work) A.work[basew + __popc(mw & lane_lt)] = v;
                }
            }
        }
        // the warp that finishes the last ticket of the group refills the stage with the group S ahead
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            if (atomicAdd(&done_cnt[s], 1) == kDenseWarps - 1) {
                done_cnt[s] = 0;
                if (g + S < n_groups) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    issue_group(g + S);
                }
            }
        }
    }
    if (WORKLIST) flush_work();
    warp_count_add(&A.ctr->boundary, n_bnd);
    if (!WORKLIST) {
        warp_count_add(&A.ctr->pad[0], n_listed);
        __syncthreads();
        if (threadIdx.x == 0) A.blist_cnt[blockIdx.x] = *cursor;
    }
}

// The bulk decision over the list k_scan_classify wrote with `grid_a` blocks (one segment of `chunk` tiles each).
// `split` blocks share a segment (the kernel is bound by the latency of its gathers: it wants every warp slot of the SM).
template <bool STAGE1>
__global__ void __launch_bounds__(256) k_bulk_decide(const __grid_constant__ ReassignArgs A, int grid_a, int chunk, int split) {
    const int* __restrict__ cmeta = A.cmeta;
    unsigned n_tests = 0, n_props = 0;
    for (int job = blockIdx.x; job < grid_a * split; job += gridDim.x) {
        const int sg = job / split, part = job - sg * split;
        const int t0 = min(A.tile_end, A.tile_begin + sg * chunk);
        const int4* seg = A.blist + (int64_t)(t0 - A.tile_begin) * 32;
        const int n = A.blist_cnt[sg];
        for (int i = part * blockDim.x + threadIdx.x; i < n; i += split * blockDim.x) {
            const int4 e = seg[i];                                              // (vertex, own cluster, first candidate, second or -1)
            const int v = e.x, a = e.y;
            const bool two = e.w >= 0;
            n_tests += two ? 2u : 1u;
            int best_b = -1;
            // everything the decision reads depends on the record only: one level of latency
            const int meta_a = __ldg(cmeta + a);
            const double4 ca = *reinterpret_cast<const double4*>(A.bulk_cen + 4 * (int64_t)a);
            const double4 cb1 = *reinterpret_cast<const double4*>(A.bulk_cen + 4 * (int64_t)e.z);
            const double4 cb2 = *reinterpret_cast<const double4*>(A.bulk_cen + 4 * (int64_t)(two ? e.w : e.z));
            const double px = A.xyz[3 * (int64_t)v], py = A.xyz[3 * (int64_t)v + 1], pz = A.xyz[3 * (int64_t)v + 2];
            double w = 0;
            if (STAGE1) w = __ldg(A.weight + v);
            if ((meta_a & 0x7fffffff) != 1) {                                   // a cluster is never emptied
                double dx = px - ca.x, dy = py - ca.y, dz = pz - ca.z;
                double best = dx * dx + dy * dy + dz * dz;
                if (STAGE1) best = ca.w / (ca.w - w) * best;
                dx = px - cb1.x; dy = py - cb1.y; dz = pz - cb1.z;
                double d = dx * dx + dy * dy + dz * dz;
                if (STAGE1) d = cb1.w / (cb1.w + w) * d;
                if (d < best) { best = d; best_b = e.z; }
                if (two) {
                    dx = px - cb2.x; dy = py - cb2.y; dz = pz - cb2.z;
                    d = dx * dx + dy * dy + dz * dz;
                    if (STAGE1) d = cb2.w / (cb2.w + w) * d;
                    if (d < best) { best = d; best_b = e.w; }
                }
            }
            if (best_b >= 0) {
                A.prop_dst[v] = best_b;
                if (A.bulk_count_leave) asm volatile("red.global.add.s32 [%0], 1;" ::"l"(A.bulk_leave + a) : "memory");
                atomicOr(&A.prop_mask[v >> 5], 1u << (v & 31));
                n_props++;
            }
        }
    }
    warp_count_add(&A.ctr->tests, n_tests);
    warp_count_add(&A.ctr->proposals, n_props);
}

}  // namespace acvd
