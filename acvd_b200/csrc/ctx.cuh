// Context object behind the opaque acvd_ctx handle of include/acvd_b200.h, shared by the translation units.
#pragma once
#include <nccl.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/acvd_b200.h"
#include "common.cuh"
#include "nccl_dyn.hpp"
#include "reassign_types.cuh"

using namespace acvd;

struct NcclError { ncclResult_t code; const char* what; };
#define ACVD_NCCL(expr)                                              \
    do {                                                             \
        ncclResult_t _r = (expr);                                    \
        if (_r != ncclSuccess) throw NcclError{_r, #expr};           \
    } while (0)

struct acvd_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;       // uploads that overlap the mesh build (acvd_set_mesh)
    cudaEvent_t copy_ev[2] = {nullptr, nullptr};
    std::string err;
    // mesh
    int V = 0, F = 0;
    int64_t nnz = 0;   // 2E
    DevBuf<float> xyz;
    DevBuf<int> tri, row_ptr, col, vf_ptr, ell;
    int ell_w = 0;                    // ELL width (6 or 8)
    int max_deg = 0;                  // longest adjacency row
    int64_t vpad = 0;
    DevBuf<unsigned long long> vf_keys, ringadj;
    // last subdivision of this mesh (acvd_subdivide), kept until the next acvd_set_mesh
    int sub_V = 0, sub_F = 0;
    DevBuf<float> sub_xyz;
    DevBuf<int> sub_tri, sub_parent1, sub_parent2;
    // items
    int metric = -1;
    DevBuf<double> area, weight, items;
    bool have_items = false;
    // clusters
    int K = 0;
    DevBuf<int> cid, cid_saved, csize, mod_round, anchor;
    int64_t launches = 0;             // kernels launched (reported per call)
    DevBuf<unsigned char> frozen;
    bool has_frozen = false, has_anchor = false;
    std::vector<int64_t> fixed;
    DevBuf<double> csum, cenergy, ccentroid;
    bool stats_valid = false;
    // reassignment scratch
    DevBuf<unsigned long long> best, prop_key;
    DevBuf<int> prop_dst, plist, plist_b, work, tile_sig, active_tiles;
    DevBuf<unsigned char> tile_active, tile_stale;
    DevBuf<unsigned> moved_mask;                // stage-1 bulk rounds: vertices the last commit moved (rollback)
    bool last_bulk_seg = false;                 // ... laid out as fixed-size segments (one-collective exchange)
    long long last_seg_bytes = 0, last_seg_cap = 0, bulk_cap_next = 0;   // segment geometry; capacity of the next round's segments
    int64_t last_bulk_total = 0;                // multi-GPU: move records of the last bulk round (in moves_all)
    DevBuf<unsigned> prop_mask;                 // bulk rounds: proposing vertices, one bit per vertex
    DevBuf<unsigned long long> round_scalars;   // [0] active tiles, [1] proposals of the previous round
    int plist_cur = 0;
    bool sig_valid = false;           // tile signatures describe the current clustering
    bool dense_next = true;           // next round scans all tiles (activity was high)
    int last_all_tiles = 0, last_tile_count = 0, last_bulk = 0;
    bool last_dense_kernel = false;   // the last scan launch was k_scan_bulk_dense
    // per-cluster member arrays and sparse rounds (sparse.cuh)
    DevBuf<int> memb_off, memb_cap, memb, memb_tmp, memb_pos, cc_par, cc_sz, memb_overflow, stamp, modlist0, modlist1, sp_done;
    DevBuf<unsigned long long> best2, sp_nmod, sp_resub, sp_ts;
    DevBuf<RoundCounters> sp_rc;
    bool members_valid = false;
    bool bulk_energy_pending = false;  // the energy sum of the last stage-1 round was enqueued with its counters (h_scalars[7])
    bool last_sparse_cluster = false;  // the last sparse launch was the one-cluster form
    int mod_par = 0;                  // which modlist the last round wrote (0 / 1)
    bool modlist_valid = false;       // ... and whether it describes the last round completely
    RoundCounters* h_sp_rc = nullptr; // pinned mirrors of the per-round records of one sparse launch
    unsigned long long* h_sp_ts = nullptr;
    // bulk (Lloyd-criterion) rounds
    DevBuf<long long> isum;
    DevBuf<double> bulk_cen, bulk_energy, bulk_energy_sum;
    DevBuf<int> leave_cnt, join_cnt;
    double fx_scale = 0.0;
    DevBuf<double2> prop_e;
    DevBuf<unsigned> modbits;
    DevBuf<int4> blist;           // split dense bulk scan: candidate list (vpad entries), segment counts
    DevBuf<int> blist_cnt;
    DevBuf<int> cmeta;            // K + 1: size | modified << 31, for the dense bulk scan
    DevBuf<RoundCounters> ctr;
    int commit_passes = 4;            // select+commit passes per round (ACVD_COMMIT_PASSES)
    RoundCounters* h_ctr = nullptr;   // pinned, kRoundSlots entries (rounds launched back to back report into separate slots)
    int round = 1;
    int cc_since = 0;             // CleanClustering: clusters not modified since this round were connected at the last check
    cudaEvent_t ev[4 * 8] = {};       // 4 events per round slot
    // generic scratch
    DevBuf<char> cub_temp;
    DevBuf<int> sort_k0, sort_k1, sort_v0, sort_v1, seg, label, comp_size, n_comp, n_roots, null_list, pick, fill_lvl, fill_seq;
    DevBuf<unsigned long long> winner, scalars;   // scalars: small device counters
    unsigned long long* h_scalars = nullptr;      // pinned, 8 entries + one active-tile count per round slot
    std::vector<double> energy_log;
    std::vector<double> energy_time;   // seconds since the start of acvd_minimize at which each entry of energy_log was taken
    int stats_constrained = 1, stats_qlevel = 3;
    // multi-GPU (dist.cuh)
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    bool replicated_tail = false;             // multi-GPU: the tail of a phase runs redundantly on every rank, no exchange
    DevBuf<char> moves_local, moves_all;      // move records of this rank / of all ranks
    DevBuf<unsigned long long> hdr_local, hdr_all, n_moves;
    unsigned long long* h_hdr = nullptr;      // pinned: world x 8
};

inline std::string g_create_error;

inline int fail(acvd_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}

struct AllocScope {   // routes DevBuf allocations of this API call to the context's stream-ordered pool
    AllocScope(cudaStream_t s) { g_alloc_stream = s; g_alloc_async = true; }
    ~AllocScope() { g_alloc_stream = nullptr; g_alloc_async = false; }
};

#define ACVD_API_BEGIN(ctx)                                                          \
    if (!(ctx)) return fail(nullptr, ACVD_EINVAL, "null context");                   \
    try {                                                                            \
        ACVD_CUDA(cudaSetDevice((ctx)->device));                                     \
        AllocScope _alloc_scope((ctx)->stream);
#define ACVD_API_END(ctx)                                                            \
    }                                                                                \
    catch (const CudaError& e) {                                                     \
        char buf[512];                                                               \
        snprintf(buf, sizeof buf, "CUDA error %s (%s) at %s:%d: %s", cudaGetErrorName(e.code), \
                 cudaGetErrorString(e.code), e.file, e.line, e.what);                \
        cudaGetLastError();                                                          \
        return fail((ctx), e.code == cudaErrorMemoryAllocation ? ACVD_ENOMEM : ACVD_ECUDA, buf); \
    }                                                                                \
    catch (const NcclError& e) {                                                     \
        std::string m = std::string("NCCL error in ") + e.what + ": " +             \
                        (nccl().GetErrorString ? nccl().GetErrorString(e.code) : "?"); \
        return fail((ctx), ACVD_ENCCL, m);                                           \
    }                                                                                \
    catch (const std::exception& e) { return fail((ctx), ACVD_EINVAL, e.what()); }   \
    catch (...) { return fail((ctx), ACVD_EINVAL, "unknown error"); }                \
    return ACVD_OK;

