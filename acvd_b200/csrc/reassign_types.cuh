// Argument / counter structs of the reassignment kernels (shared with the host-side context).
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

// Per-cluster member arrays (sparse.cuh): cluster c owns the slots [off[c], off[c + 1]) of memb, the first csize[c]
// of them are its vertices; pos[v] is the slot of v.
struct Members {
    const int* __restrict__ off;
    int* memb;
    int* pos;
    int* overflow;
};

struct ReassignArgs {
    int V, K;
    const int* __restrict__ row_ptr;
    const int* __restrict__ col;
    const unsigned long long* __restrict__ ringadj;   // V: 8x8 adjacency matrix of every vertex ring (rows <= 8)
    const int* __restrict__ ell;        // column-major padded adjacency (W columns of vpad entries)
    long long vpad;
    int* cid;
    const double* __restrict__ items;   // V x stride
    double* csum;                       // K x stride
    double* cenergy;                    // K
    int* csize;                         // K
    int* mod_round;                     // K: last round a cluster was modified
    unsigned* modbits;                  // ceil(K/32) words: cluster modified in the previous round
    const int* cmeta;                   // K + 1: cluster size, bit 31 = modified in the previous round (bulk rounds; written by k_modbits)
    const unsigned char* __restrict__ frozen;   // K or null
    const int* __restrict__ anchor;     // K or null (QEM fixed clusters)
    const float* __restrict__ xyz;      // V x 3 (anchor coordinates)
    unsigned long long* best;           // K: min priority key per cluster this round
    int* prop_dst;                      // V: proposed destination or -1
    unsigned long long* prop_key;       // V
    double2* prop_e;                    // V: (E(a - v), E(b + v)) of the proposal
    int* plist;                         // compact list of proposing vertices this round (exact rounds)
    unsigned* prop_mask;                // n_tiles: proposing vertices of a bulk round, one bit per vertex of the tile
    unsigned* moved_mask;               // n_tiles or null: vertices k_bulk_commit moved (kept by stage 1 for its rollback)
    const double* __restrict__ weight;  // V: item weights (= items[v][3]), dense copy for the bulk rounds
    int* work;                          // compact list of boundary vertices to (re)evaluate this round
    const int* plist_prev;              // proposing vertices of the previous round
    const unsigned long long* n_prev_props;   // their count
    int* tile_sig;                      // n_tiles x 8 cluster-id signature of every 32-vertex tile
    unsigned char* tile_active;         // n_tiles: tile is re-scanned this round
    unsigned char* tile_stale;          // n_tiles: a vertex in / next to the tile moved since its signature was built
    int* active_tiles;                  // compact list of active tiles
    unsigned long long* n_active_tiles;
    RoundCounters* ctr;
    Members mem;                        // memb == null: not maintained
    int* modlist;                       // clusters modified in this round (written by the commits), or null
    unsigned long long* n_mod;
    int* stamp;                         // V: round in which the vertex entered the work list of a sparse round
    int round;
    int force_all;                      // SetAllClustersToModified (:717-722)
    int bulk;                           // bulk round: no stored proposals; the decision is taken inside k_scan
    int bulk_stage;                     // 0: Lloyd criterion, 1: delta-E criterion against the round-start sums
    int bulk_count_leave;               // count leavers per cluster here (single GPU) or after the all-gather
    const double* bulk_cen;             // K x 4 centroid + weight
    int4* blist;                        // split dense bulk scan: (vertex, up to three candidate clusters in slot order) per dirty boundary vertex,
    int* blist_cnt;                     //   one segment per scanning block (segment b starts at its first tile x 32), and the segment lengths
    int* bulk_leave;                    // K
    int item_stride;                    // doubles per item row
    int all_tiles;                      // dense round: scan tiles [tile_begin, tile_end) directly, no filter / list
    int tile_begin, tile_end;
    int sig_mode;                       // 0: rebuild stale signatures, 1: leave signatures alone, 2: rebuild all
    int track_stale;                    // moves mark the tiles around them stale (only while the signatures are valid)
    int connexity;
    EvalCfg cfg;
};

}  // namespace acvd
