// Multi-GPU plumbing and round drivers (included by acvd_capi.cu so the kernels are shared).
// One process per GPU; NCCL over NVLink 5 / NVSwitch, resolved at run time (nccl_dyn.hpp).
#pragma once
#include <cstring>

#include "ctx.cuh"
#include "nccl_dyn.hpp"

static_assert(sizeof(ncclUniqueId) <= ACVD_NCCL_ID_BYTES, "ACVD_NCCL_ID_BYTES too small");

extern "C" int acvd_dist_unique_id(void* id_out) {
    if (!id_out) return ACVD_EINVAL;
    if (!nccl().load()) return ACVD_ENCCL;
    ncclUniqueId id;
    if (nccl().GetUniqueId(&id) != ncclSuccess) return ACVD_ENCCL;
    memset(id_out, 0, ACVD_NCCL_ID_BYTES);
    memcpy(id_out, &id, sizeof id);
    return ACVD_OK;
}

extern "C" int acvd_dist_init(acvd_ctx* c, int rank, int world, const void* id_bytes) {
    if (!c || !id_bytes || world < 1 || rank < 0 || rank >= world) return fail(c, ACVD_EINVAL, "acvd_dist_init: bad arguments");
    if (cudaSetDevice(c->device) != cudaSuccess) return fail(c, ACVD_ECUDA, "acvd_dist_init: cudaSetDevice failed");
    if (!nccl().load()) return fail(c, ACVD_ENCCL, "acvd_dist_init: " + nccl().error);
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof id);
    if (c->comm) { nccl().CommDestroy(c->comm); c->comm = nullptr; }
    ncclResult_t r = nccl().CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) return fail(c, ACVD_ENCCL, std::string("ncclCommInitRank: ") + nccl().GetErrorString(r));
    c->rank = rank;
    c->world = world;
    if (c->h_hdr) cudaFreeHost(c->h_hdr);
    if (cudaMallocHost(&c->h_hdr, (size_t)world * 8 * sizeof(unsigned long long)) != cudaSuccess)
        return fail(c, ACVD_ECUDA, "acvd_dist_init: pinned allocation failed");
    try {
        c->hdr_local.alloc(8); c->hdr_all.alloc((size_t)world * 8); c->n_moves.alloc(1);
    } catch (const CudaError&) { return fail(c, ACVD_ECUDA, "acvd_dist_init: device allocation failed"); }
    return ACVD_OK;
}

// The partition of the work across ranks, as pure host arithmetic (no device needed: the CPU tests call it through the
// C ABI): rank r scans and evaluates the vertices of the 32-vertex tiles [out[0], out[1]) -- a contiguous vertex range --
// uploads the points / faces [out[4], out[5]) / [out[6], out[7]) of the mesh, and runs the cluster pass (statistics,
// connectivity) on the clusters [out[2], out[3]) (equal chunks: the results are all-gathered in place).
extern "C" int acvd_dist_partition(int64_t V, int64_t F, int32_t K, int32_t rank, int32_t world, int64_t* out) {
    if (!out || world < 1 || rank < 0 || rank >= world || V < 0 || F < 0 || K < 0) return ACVD_EINVAL;
    const int64_t n_tiles = (V + 31) / 32;
    out[0] = n_tiles * rank / world;
    out[1] = n_tiles * (rank + 1) / world;
    const int64_t chunk = ((int64_t)K + world - 1) / world;
    out[2] = std::min<int64_t>(K, chunk * rank);
    out[3] = std::min<int64_t>(K, out[2] + chunk);
    out[4] = V * rank / world; out[5] = V * (rank + 1) / world;
    out[6] = F * rank / world; out[7] = F * (rank + 1) / world;
    return ACVD_OK;
}

// tile range owned by this rank
static void dist_tile_range(const acvd_ctx* c, int& t0, int& t1) {
    int64_t r[8];
    acvd_dist_partition(c->V, c->F, c->K, c->rank, c->world, r);
    t0 = (int)r[0]; t1 = (int)r[1];
}

// All-gather of variable-length records: the 64-byte headers first (they carry the local count and the
// round's counters, so this is also the one host synchronisation of the round), then one broadcast per
// rank, grouped.  Returns the total record count; RoundResult gets the counters summed over ranks.
static int64_t dist_gather_moves(acvd_ctx* c, size_t rec_bytes, RoundResult& r) {
    const int W = c->world;
    k_pack_header<<<1, 32, 0, c->stream>>>(c->ctr.p, c->round_scalars.p, c->n_moves.p, c->hdr_local.p);
    ACVD_LAUNCH_CHECK();
    ACVD_NCCL(nccl().AllGather(c->hdr_local.p, c->hdr_all.p, 8, ncclUint64, c->comm, c->stream));
    ACVD_CUDA(cudaMemcpyAsync(c->h_hdr, c->hdr_all.p, (size_t)W * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    int64_t total = 0;
    r.proposals = r.tests = r.evaluated = r.boundary = r.active_tiles = 0;
    for (int i = 0; i < W; i++) {
        const unsigned long long* h = c->h_hdr + 8 * i;
        total += (int64_t)h[0];
        r.proposals += h[1]; r.tests += h[2]; r.evaluated += h[3]; r.boundary += h[4]; r.active_tiles += h[5];
    }
    if (total == 0) return 0;
    c->moves_all.alloc((size_t)total * rec_bytes);
    ACVD_NCCL(nccl().GroupStart());
    int64_t off = 0;
    for (int i = 0; i < W; i++) {
        const int64_t n = (int64_t)c->h_hdr[8 * i];
        if (n > 0)
            ACVD_NCCL(nccl().Broadcast(c->moves_local.p, c->moves_all.p + off * rec_bytes, (size_t)n * rec_bytes, ncclChar, i, c->comm, c->stream));
        off += n;
    }
    ACVD_NCCL(nccl().GroupEnd());
    return total;
}

// in-place NCCL all-gather of the per-cluster statistics every rank computed for its cluster chunk (sums, energies,
// representative points): the "allreduce of per-cluster statistics" of the vertex-range design, without changing a bit
// of what one GPU computes (each cluster is summed by exactly one rank, in item order)
static void dist_allgather_stats(acvd_ctx* c) {
    const size_t chunk = (size_t)dist_cluster_chunk(c), npad = (size_t)payload_npad(c->metric);
    ACVD_NCCL(nccl().GroupStart());
    ACVD_NCCL(nccl().AllGather(c->csum.p + (size_t)c->rank * chunk * npad, c->csum.p, chunk * npad, ncclDouble, c->comm, c->stream));
    ACVD_NCCL(nccl().AllGather(c->cenergy.p + (size_t)c->rank * chunk, c->cenergy.p, chunk, ncclDouble, c->comm, c->stream));
    ACVD_NCCL(nccl().AllGather(c->ccentroid.p + 3 * (size_t)c->rank * chunk, c->ccentroid.p, 3 * chunk, ncclDouble, c->comm, c->stream));
    ACVD_NCCL(nccl().GroupEnd());
}
static void dist_allreduce_counters(acvd_ctx* c, unsigned long long* d, int n) {
    ACVD_NCCL(nccl().AllReduce(d, d, (size_t)n, ncclUint64, ncclSum, c->comm, c->stream));
}

// every rank uploads its slice of a host array; grouped broadcasts complete the device copy on all ranks
static void dist_sliced_upload(acvd_ctx* c, void* d, const void* h, size_t n_items, size_t item_bytes) {
    const int W = c->world;
    auto lo = [&](int r) { return n_items * (size_t)r / (size_t)W; };      // the ranges of acvd_dist_partition
    const size_t b0 = lo(c->rank) * item_bytes, b1 = lo(c->rank + 1) * item_bytes;
    if (b1 > b0) ACVD_CUDA(cudaMemcpyAsync((char*)d + b0, (const char*)h + b0, b1 - b0, cudaMemcpyHostToDevice, c->stream));
    ACVD_NCCL(nccl().GroupStart());
    for (int r = 0; r < W; r++) {
        const size_t r0 = lo(r) * item_bytes, r1 = lo(r + 1) * item_bytes;
        if (r1 > r0) ACVD_NCCL(nccl().Broadcast((char*)d + r0, (char*)d + r0, r1 - r0, ncclChar, r, c->comm, c->stream));
    }
    ACVD_NCCL(nccl().GroupEnd());
}

// When a phase leaves the exchanged rounds for its replicated tail, every rank needs every live proposal (they are
// stored by the rank that owns the vertex): pack, all-gather, install.  Returns the number of live proposals.
struct PropRec { int v, d; unsigned long long key; double ea, eb; };
__global__ void __launch_bounds__(kThreads) k_pack_proposals(ReassignArgs A, PropRec* out, unsigned long long* n_out) {
    const int n_props = (int)A.ctr->proposals;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_props; i += gridDim.x * blockDim.x) {
        const int v = A.plist[i];
        const int d = A.prop_dst[v];
        if (d < 0) continue;
        const double2 e = A.prop_e[v];
        out[(int)atomicAdd(n_out, 1ull)] = PropRec{v, d, A.prop_key[v], e.x, e.y};
    }
}
__global__ void __launch_bounds__(kThreads) k_install_proposals(ReassignArgs A, const PropRec* __restrict__ in, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const PropRec r = in[i];
        A.prop_dst[r.v] = r.d; A.prop_key[r.v] = r.key; A.prop_e[r.v] = make_double2(r.ea, r.eb);
        A.plist[i] = r.v;
    }
}
static int64_t dist_sync_proposals(acvd_ctx* c) {
    EvalCfg cfg = make_cfg(1, 3, 0);
    ReassignArgs A = make_args(c, cfg, 0, 0);          // A.plist = the list the last round wrote
    c->moves_local.alloc(((size_t)c->V / c->world + 4096) * sizeof(PropRec));
    ACVD_CUDA(cudaMemsetAsync(c->n_moves.p, 0, sizeof(unsigned long long), c->stream));
    k_pack_proposals<<<kNumSMs * 4, kThreads, 0, c->stream>>>(A, reinterpret_cast<PropRec*>(c->moves_local.p), c->n_moves.p);
    ACVD_LAUNCH_CHECK();
    RoundResult dummy;
    memset(&dummy, 0, sizeof dummy);
    const int64_t total = dist_gather_moves(c, sizeof(PropRec), dummy);
    ACVD_CUDA(cudaMemsetAsync(c->prop_dst.p, 0xff, (size_t)c->V * sizeof(int), c->stream));
    if (total > 0) {
        k_install_proposals<<<kNumSMs * 4, kThreads, 0, c->stream>>>(A, reinterpret_cast<const PropRec*>(c->moves_all.p), (int)total);
        ACVD_LAUNCH_CHECK();
    }
    return total;
}

// exact round on `world` GPUs
static RoundResult run_round_dist(acvd_ctx* c, const EvalCfg& cfg, int connexity, int force_all, bool as_iso) {
    if (force_all) ACVD_CUDA(cudaMemsetAsync(c->round_scalars.p + 1, 0, sizeof(unsigned long long), c->stream));
    else ACVD_CUDA(cudaMemcpyAsync(c->round_scalars.p + 1, &c->ctr.p->proposals, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, c->stream));
    c->plist_cur ^= 1;
    c->members_valid = false;                              // rebuilt when the phase enters its replicated tail
    ReassignArgs A = make_args(c, cfg, connexity, force_all);
    // the clusters this round modifies (every rank applies every move), for the sparse rounds of the replicated tail
    c->mod_par ^= 1;
    A.modlist = c->mod_par ? c->modlist1.p : c->modlist0.p;
    A.n_mod = c->sp_nmod.p;
    ACVD_CUDA(cudaMemsetAsync(c->sp_nmod.p, 0, sizeof(unsigned long long), c->stream));
    c->modlist_valid = true;
    ACVD_CUDA(cudaMemsetAsync(c->best.p, 0xff, (size_t)c->K * sizeof(unsigned long long), c->stream));
    ACVD_CUDA(cudaMemsetAsync(c->ctr.p, 0, sizeof(RoundCounters), c->stream));
    ACVD_CUDA(cudaMemsetAsync(c->round_scalars.p, 0, sizeof(unsigned long long), c->stream));
    ACVD_CUDA(cudaMemsetAsync(c->n_moves.p, 0, sizeof(unsigned long long), c->stream));
    k_modbits<<<grid_for(c->K), kThreads, 0, c->stream>>>(c->K, c->mod_round.p, c->round - 1, force_all, c->modbits.p);
    ACVD_LAUNCH_CHECK();
    int t0, t1;
    dist_tile_range(c, t0, t1);
    const int own_tiles = t1 - t0;
    const int gs = grid_for((int64_t)own_tiles * 32, kThreads, 8), ge = kNumSMs * 8, gc = kNumSMs * 4;
    const bool filtered = plan_scan(c, A, force_all, t0, t1);
    ACVD_CUDA(cudaEventRecord(c->ev[0], c->stream));
    if (filtered) {
        k_tile_filter<<<grid_for(own_tiles), kThreads, 0, c->stream>>>(t0, t1, c->K, 0, reinterpret_cast<const int4*>(c->tile_sig.p),
                                                                      c->modbits.p, c->tile_active.p, c->active_tiles.p, c->round_scalars.p);
        ACVD_LAUNCH_CHECK();
    }
    launch_scan(c, A, gs);
    ACVD_LAUNCH_CHECK();
    if (filtered) {
        k_carry<<<gc, kThreads, 0, c->stream>>>(A);
        ACVD_LAUNCH_CHECK();
    }
    ACVD_CUDA(cudaEventRecord(c->ev[3], c->stream));
    launch_evaluate(c, A, as_iso, ge);
    ACVD_CUDA(cudaEventRecord(c->ev[1], c->stream));
    c->moves_local.alloc(((size_t)c->K / 2 + 64) * sizeof(MoveRec));
    RoundResult r;
    memset(&r, 0, sizeof r);
    int64_t total = 0;
    for (int pass = 0; pass < c->commit_passes; pass++) {
        if (pass > 0) {
            ACVD_CUDA(cudaMemsetAsync(c->best.p, 0xff, (size_t)c->K * sizeof(unsigned long long), c->stream));
            ACVD_CUDA(cudaMemsetAsync(c->n_moves.p, 0, sizeof(unsigned long long), c->stream));
            k_resubmit<<<gc, kThreads, 0, c->stream>>>(A);
            ACVD_LAUNCH_CHECK();
        }
        // 1. conflict resolution across ranks: the minimum key per cluster
        ACVD_NCCL(nccl().AllReduce(c->best.p, c->best.p, (size_t)c->K, ncclUint64, ncclMin, c->comm, c->stream));
        // 2. winners of this rank -> all ranks
        k_select_winners<<<gc, kThreads, 0, c->stream>>>(A, reinterpret_cast<MoveRec*>(c->moves_local.p), c->n_moves.p);
        ACVD_LAUNCH_CHECK();
        RoundResult rp;
        memset(&rp, 0, sizeof rp);
        const int64_t n = dist_gather_moves(c, sizeof(MoveRec), rp);
        if (pass == 0) r = rp;
        total += n;
        if (n > 0) {
            const MoveRec* mv = reinterpret_cast<const MoveRec*>(c->moves_all.p);
            switch (c->metric) {
                case M_ISO: k_apply_moves<M_ISO, M_ISO><<<gc, kThreads, 0, c->stream>>>(A, mv, (int)n); break;
                case M_QEM:
                    if (as_iso) k_apply_moves<M_ISO, M_QEM><<<gc, kThreads, 0, c->stream>>>(A, mv, (int)n);
                    else k_apply_moves<M_QEM, M_QEM><<<gc, kThreads, 0, c->stream>>>(A, mv, (int)n);
                    break;
                case M_ANISO: k_apply_moves<M_ANISO, M_ANISO><<<gc, kThreads, 0, c->stream>>>(A, mv, (int)n); break;
                default: k_apply_moves<M_ANISOQ, M_ANISOQ><<<gc, kThreads, 0, c->stream>>>(A, mv, (int)n); break;
            }
            ACVD_LAUNCH_CHECK();
        }
        if (n == 0) break;   // nothing won: later passes cannot win either
    }
    ACVD_CUDA(cudaEventRecord(c->ev[2], c->stream));
    ACVD_CUDA(cudaEventSynchronize(c->ev[2]));
    r.mods = (unsigned long long)total;
    ACVD_CUDA(cudaEventElapsedTime(&r.ms_scan, c->ev[0], c->ev[3]));
    ACVD_CUDA(cudaEventElapsedTime(&r.ms_eval, c->ev[3], c->ev[1]));
    ACVD_CUDA(cudaEventElapsedTime(&r.ms_commit, c->ev[1], c->ev[2]));
    c->round++;
    update_density(c, r);
    return r;
}

// ---- one-collective exchange of the bulk rounds.  Every rank sends ONE fixed-size segment -- a 64-byte header (its move
// count and the round's counters) followed by room for `cap` (vertex, destination) records -- in a single ncclAllGather;
// the kernels that apply the moves read the counts from the gathered headers on the device, so the round needs no host
// round trip between packing and applying (the two-step form -- headers, host sync, one broadcast per rank -- costs more
// than the scan of a rank's share of the mesh).  `cap` follows the previous round's largest count with a margin; a rank
// that would overflow it says so in its header, every rank then skips the round's apply (same headers, same decision)
// and the host repeats the exchange in the two-step form.
struct SegMoves { const unsigned char* base; long long seg_bytes; long long cap; int world; };
__device__ __forceinline__ long long seg_count(const SegMoves& S, int r) {
    return (long long)*reinterpret_cast<const unsigned long long*>(S.base + (size_t)r * S.seg_bytes);
}
__device__ __forceinline__ const int2* seg_records(const SegMoves& S, int r) {
    return reinterpret_cast<const int2*>(S.base + (size_t)r * S.seg_bytes + 64);
}
__device__ __forceinline__ bool seg_overflowed(const SegMoves& S) {
    bool o = false;
    for (int r = 0; r < S.world; r++) o |= seg_count(S, r) > S.cap;
    return o;
}

// k_pack_bulk_moves with a bound: records beyond `cap` are dropped (the count still says how many there were) and the
// proposal masks are left alone, so that an overflowing round can be packed again
__global__ void __launch_bounds__(kThreads) k_pack_bulk_moves_cap(ReassignArgs A, int2* moves, unsigned long long* n_moves, long long cap) {
    const int n_tiles = A.tile_end - A.tile_begin;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int base = warp * 32; base < n_tiles; base += n_warps * 32) {
        const int t = A.tile_begin + base + lane;
        const unsigned m = (base + lane < n_tiles) ? A.prop_mask[t] : 0u;
        const int cnt = __popc(m);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int x = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += x; }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) continue;
        long long slot0 = 0;
        if (lane == 0) slot0 = (long long)atomicAdd(n_moves, (unsigned long long)total);
        slot0 = __shfl_sync(0xffffffffu, slot0, 0);
        const long long my_off = slot0 + incl - cnt;
        unsigned nz = __ballot_sync(0xffffffffu, m != 0);
        while (nz) {
            const int src = __ffs(nz) - 1;
            nz &= nz - 1;
            const unsigned mm = __shfl_sync(0xffffffffu, m, src);
            const long long off = __shfl_sync(0xffffffffu, my_off, src);
            if ((mm >> lane) & 1u) {
                const int v = (A.tile_begin + base + src) * 32 + lane;
                const long long q = off + __popc(mm & ((1u << lane) - 1u));
                if (q < cap) moves[q] = make_int2(v, A.prop_dst[v]);
            }
        }
    }
}
__global__ void __launch_bounds__(kThreads) k_bulk_count_seg(int K, const int* __restrict__ cid, SegMoves S, int* leave_cnt) {
    if (seg_overflowed(S)) return;
    for (int r = 0; r < S.world; r++) {
        const int n = (int)seg_count(S, r);
        const int2* moves = seg_records(S, r);
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            const int a = cid[moves[i].x];
            if (a < K) atomicAdd(&leave_cnt[a], 1);
        }
    }
}
__global__ void __launch_bounds__(kThreads) k_bulk_apply_seg(ReassignArgs A, BulkArgs B, int stride, SegMoves S) {
    if (seg_overflowed(S)) return;
    const int K = A.K;
    unsigned n_mods = 0;
    for (int r = 0; r < S.world; r++) {
        const int n = (int)seg_count(S, r);
        const int2* moves = seg_records(S, r);
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
            A.prop_dst[v] = a;          // what the rollback restores
            mark_tiles_stale(A, v);
            n_mods++;
        }
    }
    warp_count_add(&A.ctr->mods, n_mods);
}
__global__ void __launch_bounds__(kThreads) k_bulk_rollback_seg(int* cid, const int* __restrict__ prev, SegMoves S) {
    for (int r = 0; r < S.world; r++) {
        const int n = (int)seg_count(S, r);
        const int2* moves = seg_records(S, r);
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            const int v = moves[i].x;
            if (cid[v] == moves[i].y) cid[v] = prev[v];
        }
    }
}
static bool dist_one_collective() { return getenv("ACVD_DIST_TWO_STEP") == nullptr; }    // A/B knob (read per round)

// bulk (Lloyd-criterion) round on `world` GPUs: local scan + evaluate, all-gather of the (vertex, destination)
// pairs, then every rank counts leavers and applies all moves (integer sums: order-independent)
static RoundResult run_bulk_round_dist(acvd_ctx* c, int force_all, int stage) {
    EvalCfg cfg = make_cfg(0, 0, 0);
    c->plist_cur = 0;
    c->members_valid = false; c->modlist_valid = false;
    ReassignArgs A = make_args(c, cfg, 0, force_all);
    A.bulk = 1; A.bulk_stage = stage; A.bulk_count_leave = 0;
    BulkArgs B = make_bulk_args(c);
    // (the round's counters are opened by k_modbits: one launch instead of three)
    k_modbits<<<grid_for(c->K), kThreads, 0, c->stream>>>(c->K, c->mod_round.p, c->round - 1, force_all, c->modbits.p, c->ctr.p,
                                                          c->round_scalars.p, 2, c->csize.p, c->cmeta.p);
    ACVD_LAUNCH_CHECK();
    int t0, t1;
    dist_tile_range(c, t0, t1);
    const int own_tiles = t1 - t0;
    const int gs = grid_for((int64_t)own_tiles * 32, kThreads, 8), ge = kNumSMs * 8, gc = kNumSMs * 4;
    const bool filtered = plan_scan(c, A, force_all, t0, t1);
    ACVD_CUDA(cudaEventRecord(c->ev[0], c->stream));
    if (filtered) {
        k_tile_filter<<<grid_for(own_tiles), kThreads, 0, c->stream>>>(t0, t1, c->K, 0, reinterpret_cast<const int4*>(c->tile_sig.p),
                                                                      c->modbits.p, c->tile_active.p, c->active_tiles.p, c->round_scalars.p);
        ACVD_LAUNCH_CHECK();
    }
    launch_scan(c, A, gs);
    ACVD_LAUNCH_CHECK();
    ACVD_CUDA(cudaEventRecord(c->ev[3], c->stream));
    k_bulk_evaluate<<<ge, kThreads, 0, c->stream>>>(A, B, 0, stage, payload_npad(c->metric));
    ACVD_LAUNCH_CHECK();
    ACVD_CUDA(cudaEventRecord(c->ev[1], c->stream));
    RoundResult r;
    memset(&r, 0, sizeof r);
    const int W = c->world;
    c->last_bulk_seg = false;
    bool exchanged = false;
    if (!force_all && c->bulk_cap_next > 0 && dist_one_collective()) {
        // ---- one collective: fixed-size segments, counts read on the device
        const long long cap = c->bulk_cap_next;
        const size_t seg = 64 + (size_t)cap * sizeof(int2);
        c->moves_local.alloc(seg); c->moves_all.alloc(seg * (size_t)W);
        ACVD_CUDA(cudaMemsetAsync(c->n_moves.p, 0, sizeof(unsigned long long), c->stream));
        k_pack_bulk_moves_cap<<<gc, kThreads, 0, c->stream>>>(A, reinterpret_cast<int2*>(c->moves_local.p + 64), c->n_moves.p, cap);
        ACVD_LAUNCH_CHECK();
        k_pack_header<<<1, 32, 0, c->stream>>>(c->ctr.p, c->round_scalars.p, c->n_moves.p, reinterpret_cast<unsigned long long*>(c->moves_local.p));
        ACVD_LAUNCH_CHECK();
        ACVD_NCCL(nccl().AllGather(c->moves_local.p, c->moves_all.p, seg, ncclChar, c->comm, c->stream));
        SegMoves S{reinterpret_cast<const unsigned char*>(c->moves_all.p), (long long)seg, cap, W};
        k_bulk_count_seg<<<gc, kThreads, 0, c->stream>>>(c->K, c->cid.p, S, c->leave_cnt.p);
        ACVD_LAUNCH_CHECK();
        k_bulk_apply_seg<<<gc, kThreads, 0, c->stream>>>(A, B, payload_npad(c->metric), S);
        ACVD_LAUNCH_CHECK();
        k_bulk_refresh<<<grid_for(c->K), kThreads, 0, c->stream>>>(c->K, c->csize.p, B);
        ACVD_LAUNCH_CHECK();
        ACVD_CUDA(cudaEventRecord(c->ev[2], c->stream));
        if (stage == 1) bulk_energy_enqueue(c);          // the energy guard's sum travels with the counters
        ACVD_CUDA(cudaMemcpy2DAsync(c->h_hdr, 64, c->moves_all.p, seg, 64, (size_t)W, cudaMemcpyDeviceToHost, c->stream));
        ACVD_CUDA(cudaMemcpyAsync(c->h_ctr, c->ctr.p, sizeof(RoundCounters), cudaMemcpyDeviceToHost, c->stream));
        ACVD_CUDA(cudaStreamSynchronize(c->stream));
        long long mx = 0, total = 0;
        for (int i = 0; i < W; i++) {
            const unsigned long long* h = c->h_hdr + 8 * i;
            mx = std::max<long long>(mx, (long long)h[0]); total += (long long)h[0];
            r.proposals += h[1]; r.tests += h[2]; r.evaluated += h[3]; r.boundary += h[4]; r.active_tiles += h[5];
        }
        c->bulk_cap_next = std::max<long long>(4096, mx + mx / 2 + 1024);
        if (mx > cap) c->bulk_energy_pending = false;       // the round is repeated below: its energy is summed again
        if (mx <= cap) {
            exchanged = true;
            c->last_bulk_seg = true; c->last_bulk_total = total; c->last_seg_bytes = (long long)seg; c->last_seg_cap = cap;
            // the proposal masks of the rank's range were left in place for a possible second packing
            ACVD_CUDA(cudaMemsetAsync(c->prop_mask.p + t0, 0, (size_t)own_tiles * sizeof(unsigned), c->stream));
            r.mods = c->h_ctr->mods;
            ACVD_CUDA(cudaEventElapsedTime(&r.ms_scan, c->ev[0], c->ev[3]));
            ACVD_CUDA(cudaEventElapsedTime(&r.ms_eval, c->ev[3], c->ev[1]));
            ACVD_CUDA(cudaEventElapsedTime(&r.ms_commit, c->ev[1], c->ev[2]));
            c->round++;
            c->stats_valid = false;
            update_density(c, r);
            return r;
        }
        // some rank had more moves than the segment holds: nothing was applied anywhere; repeat in the two-step form
        memset(&r, 0, sizeof r);
    }
    (void)exchanged;
    c->moves_local.alloc(((size_t)(own_tiles) * 32 + 64) * sizeof(int2));
    ACVD_CUDA(cudaMemsetAsync(c->n_moves.p, 0, sizeof(unsigned long long), c->stream));
    k_pack_bulk_moves<<<gc, kThreads, 0, c->stream>>>(A, reinterpret_cast<int2*>(c->moves_local.p), c->n_moves.p);
    ACVD_LAUNCH_CHECK();
    const int64_t total = dist_gather_moves(c, sizeof(int2), r);
    c->last_bulk_total = total;
    {   // the next round's segment capacity from this round's largest per-rank count
        long long mx = 0;
        for (int i = 0; i < W; i++) mx = std::max<long long>(mx, (long long)c->h_hdr[8 * i]);
        c->bulk_cap_next = std::max<long long>(4096, mx + mx / 2 + 1024);
    }
    if (total > 0) {
        const int2* mv = reinterpret_cast<const int2*>(c->moves_all.p);
        k_bulk_count<<<gc, kThreads, 0, c->stream>>>(c->K, c->cid.p, mv, (int)total, c->leave_cnt.p);
        ACVD_LAUNCH_CHECK();
        k_bulk_apply<<<gc, kThreads, 0, c->stream>>>(A, B, payload_npad(c->metric), mv, (int)total);
        ACVD_LAUNCH_CHECK();
        k_bulk_refresh<<<grid_for(c->K), kThreads, 0, c->stream>>>(c->K, c->csize.p, B);
        ACVD_LAUNCH_CHECK();
    }
    ACVD_CUDA(cudaEventRecord(c->ev[2], c->stream));
    ACVD_CUDA(cudaMemcpyAsync(c->h_ctr, c->ctr.p, sizeof(RoundCounters), cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    r.mods = c->h_ctr->mods;   // accepted moves (identical on every rank)
    ACVD_CUDA(cudaEventElapsedTime(&r.ms_scan, c->ev[0], c->ev[3]));
    ACVD_CUDA(cudaEventElapsedTime(&r.ms_eval, c->ev[3], c->ev[1]));
    ACVD_CUDA(cudaEventElapsedTime(&r.ms_commit, c->ev[1], c->ev[2]));
    c->round++;
    c->stats_valid = false;
    update_density(c, r);
    return r;
}

// undo of the last bulk round's moves (stage-1 energy guard), whichever way they were exchanged
static void dist_bulk_rollback(acvd_ctx* c) {
    if (c->last_bulk_seg) {
        SegMoves S{reinterpret_cast<const unsigned char*>(c->moves_all.p), c->last_seg_bytes, c->last_seg_cap, c->world};
        k_bulk_rollback_seg<<<kNumSMs * 4, kThreads, 0, c->stream>>>(c->cid.p, c->prop_dst.p, S);
    } else {
        k_bulk_rollback_moves<<<kNumSMs * 4, kThreads, 0, c->stream>>>(c->cid.p, c->prop_dst.p, reinterpret_cast<const int2*>(c->moves_all.p),
                                                                     (int)c->last_bulk_total);
    }
    ACVD_LAUNCH_CHECK();
}
