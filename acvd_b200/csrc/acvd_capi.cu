// acvd_b200 C ABI (include/acvd_b200.h): context, host-side convergence driver, kernel launches.
// sm_100a only; there is no CPU path — every entry point needs a live CUDA context.
#include "../../include/acvd_b200.h"

#include <cub/cub.cuh>

#include <algorithm>
#include <array>
#include <chrono>
#include <functional>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "cleanup.cuh"
#include "curvature.cuh"
#include "fill.cuh"
#include "manifold.cuh"
#include "common.cuh"
#include "ctx.cuh"
#include "host_sampling.hpp"
#include "mesh.cuh"
#include "metric.cuh"
#include "reassign.cuh"
#include "scan_dense.cuh"
#include "sparse.cuh"

// blocks per SM requested for the scan grid (grid-stride; more blocks than resident ones balance the tail)
static int scan_bps() { static int v = getenv("ACVD_SCAN_BPS") ? atoi(getenv("ACVD_SCAN_BPS")) : 8; return v; }
#define ACVD_SCAN_BPS scan_bps()

// every launch site of our own kernels has the context in scope as `c`: count them for the report
#undef ACVD_LAUNCH_CHECK
#define ACVD_LAUNCH_CHECK()              \
    do {                                 \
        c->launches++;                   \
        ACVD_CUDA(cudaGetLastError());   \
    } while (0)

// ACVD_TRACE=1: wall-clock trace of the host driver's stages on stderr (the reference's ConsoleOutput>1
// per-loop lines are the analogue, Common/vtkUniformClustering.h:752-760)
constexpr int64_t kReplicatedTailEvaluated = 400000;   // multi-GPU: below this many evaluated vertices per round the phase goes replicated
constexpr int kRoundSlots = 8;            // exact rounds that may be in flight between two host synchronisations
constexpr int kBulkMinVertices = 500000;    // acvd_params.bulk_rounds == 0 (automatic): bulk rounds on from this mesh size
constexpr int kTailBatch = 4;             // rounds launched back to back in the long tail of the last phases
constexpr int kSparseChunkAlloc = 512;    // rounds of one sparse launch (kSparseChunk)

static bool trace_on() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("ACVD_TRACE"); on = (e && *e && *e != '0') ? 1 : 0; }
    return on == 1;
}
struct TraceScope {
    const char* name; acvd_ctx* c; std::chrono::steady_clock::time_point t0;
    TraceScope(acvd_ctx* ctx, const char* n) : name(n), c(ctx) { if (trace_on()) { cudaStreamSynchronize(c->stream); t0 = std::chrono::steady_clock::now(); } }
    ~TraceScope() {
        if (!trace_on()) return;
        cudaStreamSynchronize(c->stream);
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        fprintf(stderr, "[acvd trace] %-28s %10.3f ms\n", name, ms);
    }
};

struct EventPair {   // RAII: the events do not leak when a throw leaves the API call
    cudaEvent_t a = nullptr, b = nullptr;
    EventPair() { ACVD_CUDA(cudaEventCreate(&a)); if (cudaEventCreate(&b) != cudaSuccess) { cudaEventDestroy(a); throw CudaError{cudaErrorUnknown, "cudaEventCreate", __FILE__, __LINE__}; } }
    ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    EventPair(const EventPair&) = delete;
    EventPair& operator=(const EventPair&) = delete;
};

__global__ void k_check_range(int64_t n, const int* __restrict__ a, int hi, unsigned long long* bad) {
    unsigned cnt = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) cnt += (a[i] < 0 || a[i] >= hi) ? 1u : 0u;
    acvd::warp_count_add(bad, cnt);
}

static void* cub_temp(acvd_ctx* c, size_t bytes) {
    c->cub_temp.alloc(bytes + 16);
    return c->cub_temp.p;
}

static int bits_for(uint64_t maxval) { int b = 1; while (b < 64 && (maxval >> b)) b++; return b; }

// sort 64-bit keys in place (result guaranteed in `keys`), returns nothing; n may be large
static void sort_keys64(acvd_ctx* c, unsigned long long* keys, unsigned long long* alt, int64_t n, int end_bit) {
    cub::DoubleBuffer<unsigned long long> db(keys, alt);
    size_t tb = 0;
    ACVD_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, db, n, 0, end_bit, c->stream));
    void* t = cub_temp(c, tb);
    ACVD_CUDA(cub::DeviceRadixSort::SortKeys(t, tb, db, n, 0, end_bit, c->stream));
    if (db.Current() != keys)
        ACVD_CUDA(cudaMemcpyAsync(keys, db.Current(), n * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, c->stream));
}

// ---------------------------------------------------------------------------------------------
extern "C" int acvd_abi_version(void) { return ACVD_B200_ABI_VERSION; }
extern "C" int acvd_payload_size(int metric) { return (metric < 0 || metric > 3) ? ACVD_EINVAL : payload_np(metric); }

extern "C" const char* acvd_last_error(acvd_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int acvd_create(acvd_ctx** out, int device) {
    if (!out) return fail(nullptr, ACVD_EINVAL, "null out pointer");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(nullptr, ACVD_ENODEVICE, "no CUDA device: acvd_b200 has no CPU path");
    }
    acvd_ctx* c = new acvd_ctx();
    try {
        if (device < 0) ACVD_CUDA(cudaGetDevice(&device));
        if (device >= n) { delete c; return fail(nullptr, ACVD_EINVAL, "device index out of range"); }
        c->device = device;
        ACVD_CUDA(cudaSetDevice(device));
        ACVD_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        ACVD_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (auto& e : c->copy_ev) ACVD_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        {   // keep freed blocks in the pool (the default threshold 0 returns them to the driver at every synchronisation)
            cudaMemPool_t pool;
            ACVD_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
            uint64_t keep = ~0ull;
            ACVD_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        }
        AllocScope alloc_scope(c->stream);
        for (auto& ev : c->ev) ACVD_CUDA(cudaEventCreate(&ev));
        ACVD_CUDA(cudaMallocHost(&c->h_ctr, kRoundSlots * sizeof(RoundCounters)));
        ACVD_CUDA(cudaMallocHost(&c->h_scalars, (8 + kRoundSlots) * sizeof(unsigned long long)));
        c->ctr.alloc(1);
        c->scalars.alloc(8);
    } catch (const CudaError& err) {
        std::string m = std::string("CUDA error in acvd_create: ") + cudaGetErrorString(err.code);
        delete c;
        return fail(nullptr, ACVD_ECUDA, m);
    }
    *out = c;
    return ACVD_OK;
}

extern "C" int acvd_destroy(acvd_ctx* c) {
    if (!c) return ACVD_OK;
    cudaSetDevice(c->device);
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    for (auto& e : c->copy_ev) if (e) cudaEventDestroy(e);
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    for (auto& ev : c->ev) if (ev) cudaEventDestroy(ev);
    if (c->h_ctr) cudaFreeHost(c->h_ctr);
    if (c->h_scalars) cudaFreeHost(c->h_scalars);
    if (c->h_hdr) cudaFreeHost(c->h_hdr);
    if (c->comm && nccl().load()) nccl().CommDestroy(c->comm);
    delete c;
    return ACVD_OK;
}

// ---------------------------------------------------------------------------------------------
// mesh
static void dist_sliced_upload(acvd_ctx* c, void* d, const void* h, size_t n_items, size_t item_bytes);   // dist.cuh

extern "C" int acvd_set_mesh(acvd_ctx* c, int32_t V, int32_t F, const float* xyz, const int32_t* tri) {
    ACVD_API_BEGIN(c)
    if (V <= 0 || F <= 0 || !xyz || !tri) throw std::runtime_error("acvd_set_mesh: bad arguments");
    c->V = V; c->F = F;
    // a new mesh invalidates everything derived from the old one: items, clusters (acvd_set_num_clusters is required
    // again: the per-vertex buffers are sized by it), saved clustering, tile signatures, fixed-point scale
    c->have_items = false; c->stats_valid = false; c->sub_V = c->sub_F = 0;
    c->K = 0; c->sig_valid = false; c->fx_scale = 0.0; c->cid_saved.release(); c->has_frozen = c->has_anchor = false; c->fixed.clear();
    c->vpad = (((int64_t)V + 31) / 32) * 32;      // per-vertex streams are padded to whole 32-vertex tiles (TMA copies whole tiles)
    c->xyz.alloc(3 * (size_t)c->vpad);
    c->tri.alloc(3 * (size_t)F);
    struct CopyGuard {          // the caller's point array is not read after this call returns, whichever way it returns
        cudaStream_t s; bool armed = false;
        ~CopyGuard() { if (armed) cudaStreamSynchronize(s); }
    } xyz_guard{c->copy_stream};
    {
        TraceScope ts(c, "set_mesh: upload");
        if (c->world > 1) {   // every rank uploads its vertex / face range over its own PCIe link; NVLink completes the copies
            dist_sliced_upload(c, c->xyz.p, xyz, (size_t)V, 3 * sizeof(float));
            dist_sliced_upload(c, c->tri.p, tri, (size_t)F, 3 * sizeof(int));
        } else {
            // the faces first (everything below is built from them); the points follow on the copy stream and travel
            // while the adjacency is built (nothing in this call reads them)
            ACVD_CUDA(cudaMemcpyAsync(c->tri.p, tri, 3 * (size_t)F * sizeof(int), cudaMemcpyHostToDevice, c->stream));
            ACVD_CUDA(cudaEventRecord(c->copy_ev[0], c->stream));
            ACVD_CUDA(cudaStreamWaitEvent(c->copy_stream, c->copy_ev[0], 0));
            xyz_guard.armed = true;
            ACVD_CUDA(cudaMemcpyAsync(c->xyz.p, xyz, 3 * (size_t)V * sizeof(float), cudaMemcpyHostToDevice, c->copy_stream));
            ACVD_CUDA(cudaEventRecord(c->copy_ev[1], c->copy_stream));
        }
        // vertex indices outside [0, V) would index the counting build out of bounds: reject them here
        ACVD_CUDA(cudaMemsetAsync(c->scalars.p, 0, sizeof(unsigned long long), c->stream));
        k_check_range<<<grid_for(3 * (int64_t)F), kThreads, 0, c->stream>>>(3 * (int64_t)F, c->tri.p, V, c->scalars.p);
        ACVD_LAUNCH_CHECK();
        ACVD_CUDA(cudaMemcpyAsync(c->h_scalars, c->scalars.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        ACVD_CUDA(cudaStreamSynchronize(c->stream));
        if (c->h_scalars[0] != 0) { c->V = c->F = 0; throw std::runtime_error("acvd_set_mesh: triangle vertex index out of range"); }
    }
    // --- CSR adjacency and vertex -> face incidence (faces ascending per vertex), by counting: see mesh.cuh
    {
        TraceScope ts(c, "set_mesh: csr + incidence");
        const int64_t n3 = 3 * (int64_t)F;
        DevBuf<int> cnt, cursor, he, deg;
        DevBuf<int4> rec;
        cnt.alloc((size_t)V + 1); cursor.alloc(V); he.alloc((size_t)(2 * n3)); deg.alloc((size_t)V + 1); rec.alloc((size_t)n3);
        c->vf_ptr.alloc((size_t)V + 1); c->vf_keys.alloc((size_t)n3); c->row_ptr.alloc((size_t)V + 1);
        ACVD_CUDA(cudaMemsetAsync(cnt.p, 0, ((size_t)V + 1) * sizeof(int), c->stream));
        ACVD_CUDA(cudaMemsetAsync(cursor.p, 0, (size_t)V * sizeof(int), c->stream));
        ACVD_CUDA(cudaMemsetAsync(deg.p, 0, ((size_t)V + 1) * sizeof(int), c->stream));
        k_count_corners<<<grid_for(F), kThreads, 0, c->stream>>>(F, c->tri.p, cnt.p);
        ACVD_LAUNCH_CHECK();
        size_t tb = 0;
        ACVD_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt.p, c->vf_ptr.p, V + 1, c->stream));
        void* t = cub_temp(c, tb);
        ACVD_CUDA(cub::DeviceScan::ExclusiveSum(t, tb, cnt.p, c->vf_ptr.p, V + 1, c->stream));
        k_scatter_corners<<<grid_for(F), kThreads, 0, c->stream>>>(F, c->tri.p, c->vf_ptr.p, cursor.p, rec.p);
        ACVD_LAUNCH_CHECK();
        k_sort_rows<<<grid_for(V), kThreads, 0, c->stream>>>(V, c->vf_ptr.p, rec.p, c->vf_keys.p, he.p, deg.p);
        ACVD_LAUNCH_CHECK();
        ACVD_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, deg.p, c->row_ptr.p, V + 1, c->stream));
        t = cub_temp(c, tb);
        ACVD_CUDA(cub::DeviceScan::ExclusiveSum(t, tb, deg.p, c->row_ptr.p, V + 1, c->stream));
        int nu = 0;
        ACVD_CUDA(cudaMemcpyAsync(&nu, c->row_ptr.p + V, sizeof nu, cudaMemcpyDeviceToHost, c->stream));
        ACVD_CUDA(cudaStreamSynchronize(c->stream));
        if (nu < 0) throw std::runtime_error("acvd_set_mesh: more than 2^31 adjacency entries");
        c->nnz = nu;
        c->col.alloc((size_t)std::max(nu, 1));
        k_compact_rows<<<grid_for(V), kThreads, 0, c->stream>>>(V, c->vf_ptr.p, c->row_ptr.p, he.p, c->col.p);
        ACVD_LAUNCH_CHECK();
        ACVD_CUDA(cudaStreamSynchronize(c->stream));
    }
    // --- ELL copy for the frontier scan (width 6 when no vertex has more neighbours, else 8 + CSR overflow)
    {
        TraceScope ts(c, "set_mesh: ell");
        int* d_max = reinterpret_cast<int*>(c->scalars.p);
        ACVD_CUDA(cudaMemsetAsync(d_max, 0, sizeof(int), c->stream));
        k_max_degree<<<grid_for(V), kThreads, 0, c->stream>>>(V, c->row_ptr.p, d_max);
        ACVD_LAUNCH_CHECK();
        int max_deg = 0;
        ACVD_CUDA(cudaMemcpyAsync(&max_deg, d_max, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        ACVD_CUDA(cudaStreamSynchronize(c->stream));
        c->ell_w = max_deg <= 6 ? 6 : 8;
        c->max_deg = max_deg;
        c->ell.alloc((size_t)c->ell_w * c->vpad);
        k_build_ell<<<grid_for(c->vpad), kThreads, 0, c->stream>>>(V, c->vpad, c->ell_w, c->row_ptr.p, c->col.p, c->ell.p);
        ACVD_LAUNCH_CHECK();
    }
    // --- ring adjacency matrices for the connexity predicate of k_evaluate
    TraceScope ts_rest(c, "set_mesh: ringadj");
    c->ringadj.alloc((size_t)V);
    k_build_ringadj<<<grid_for(V), kThreads, 0, c->stream>>>(V, c->row_ptr.p, c->col.p, c->ringadj.p);
    ACVD_LAUNCH_CHECK();
    if (xyz_guard.armed) ACVD_CUDA(cudaStreamWaitEvent(c->stream, c->copy_ev[1], 0));      // later work on the stream sees the points
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    ACVD_API_END(c)
}

extern "C" int acvd_get_num_edges(acvd_ctx* c, int64_t* E) {
    if (!c || !E) return fail(c, ACVD_EINVAL, "null argument");
    *E = c->nnz / 2;
    return ACVD_OK;
}

extern "C" int acvd_get_csr(acvd_ctx* c, int32_t* row_ptr, int32_t* col) {
    ACVD_API_BEGIN(c)
    if (!c->V) throw std::runtime_error("acvd_get_csr: no mesh");
    if (row_ptr) ACVD_CUDA(cudaMemcpy(row_ptr, c->row_ptr.p, ((size_t)c->V + 1) * sizeof(int), cudaMemcpyDeviceToHost));
    if (col) ACVD_CUDA(cudaMemcpy(col, c->col.p, (size_t)c->nnz * sizeof(int), cudaMemcpyDeviceToHost));
    ACVD_API_END(c)
}

// ---------------------------------------------------------------------------------------------
// 1 -> 4 subdivision of the context's mesh (kept on the device until fetched)
static void exclusive_sum(acvd_ctx* c, const int* in, int* out, int64_t n);
static void inclusive_sum(acvd_ctx* c, const int* in, int* out, int64_t n) {
    size_t tb = 0;
    ACVD_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb, in, out, n, c->stream));
    void* t = cub_temp(c, tb);
    ACVD_CUDA(cub::DeviceScan::InclusiveSum(t, tb, in, out, n, c->stream));
}

// The reference's edge ids of a triangle list on the device: edges are numbered in the order AddEdge first sees them
// over the faces (Common/vtkSurfaceBase.cxx:1166-1221).  Sort of the 3F undirected half-edges (stable: a run starts with
// the first occurrence of its edge), runs ranked by their first slot.
struct EdgeTable {
    int E = 0;
    DevBuf<int> first_sorted;     // E: half-edge slot 3f + k of the first occurrence of edge e
    DevBuf<int> edge_of_slot;     // 3F: edge id of every half-edge slot (undefined for inactive faces / self loops)
};
static void build_edge_table(acvd_ctx* c, const int* tri, int F, EdgeTable& T) {
    const int64_t n = 3 * (int64_t)F;
    if (n >= ((int64_t)1 << 31)) throw std::runtime_error("edge table: mesh too large");
    DevBuf<unsigned long long> keys, keys_alt;
    DevBuf<int> slots, slots_alt, head, run_incl, first_slot, run_id, run_sorted, edge_of_run;
    keys.alloc(n); keys_alt.alloc(n); slots.alloc(n); slots_alt.alloc(n); head.alloc(n); run_incl.alloc(n);
    k_sub_edge_keys<<<grid_for(F), kThreads, 0, c->stream>>>(F, tri, keys.p, slots.p);
    ACVD_LAUNCH_CHECK();
    {
        cub::DoubleBuffer<unsigned long long> dk(keys.p, keys_alt.p);
        cub::DoubleBuffer<int> dv(slots.p, slots_alt.p);
        size_t tb = 0;
        ACVD_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, n, 0, 64, c->stream));
        void* t = cub_temp(c, tb);
        ACVD_CUDA(cub::DeviceRadixSort::SortPairs(t, tb, dk, dv, n, 0, 64, c->stream));
        if (dk.Current() != keys.p) ACVD_CUDA(cudaMemcpyAsync(keys.p, dk.Current(), n * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, c->stream));
        if (dv.Current() != slots.p) ACVD_CUDA(cudaMemcpyAsync(slots.p, dv.Current(), n * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
    }
    k_sub_heads<<<grid_for(n), kThreads, 0, c->stream>>>(n, keys.p, head.p);
    ACVD_LAUNCH_CHECK();
    inclusive_sum(c, head.p, run_incl.p, n);
    int E = 0;
    ACVD_CUDA(cudaMemcpyAsync(&E, run_incl.p + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    T.E = E;
    first_slot.alloc(std::max(E, 1)); run_id.alloc(std::max(E, 1)); T.first_sorted.alloc(std::max(E, 1)); run_sorted.alloc(std::max(E, 1));
    edge_of_run.alloc(std::max(E, 1)); T.edge_of_slot.alloc(n);
    k_sub_first_slots<<<grid_for(n), kThreads, 0, c->stream>>>(n, head.p, run_incl.p, slots.p, first_slot.p, run_id.p);
    ACVD_LAUNCH_CHECK();
    if (E > 0) {
        size_t tb = 0;
        ACVD_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, first_slot.p, T.first_sorted.p, run_id.p, run_sorted.p, E, 0, 32, c->stream));
        void* t = cub_temp(c, tb);
        ACVD_CUDA(cub::DeviceRadixSort::SortPairs(t, tb, first_slot.p, T.first_sorted.p, run_id.p, run_sorted.p, E, 0, 32, c->stream));
        k_edge_of_run<<<grid_for(E), kThreads, 0, c->stream>>>(E, run_sorted.p, edge_of_run.p);
        ACVD_LAUNCH_CHECK();
        k_sub_slot_edges<<<grid_for(n), kThreads, 0, c->stream>>>(n, keys.p, run_incl.p, slots.p, edge_of_run.p, T.edge_of_slot.p);
        ACVD_LAUNCH_CHECK();
    }
}

extern "C" int acvd_subdivide(acvd_ctx* c, int32_t* n_vertices, int32_t* n_faces) {
    ACVD_API_BEGIN(c)
    if (!c->V || !n_vertices || !n_faces) throw std::runtime_error("acvd_subdivide: set the mesh first");
    const int V = c->V, F = c->F;
    EdgeTable T;
    build_edge_table(c, c->tri.p, F, T);
    const int E = T.E;
    if ((int64_t)V + E >= ((int64_t)1 << 31)) throw std::runtime_error("acvd_subdivide: too many vertices");
    DevBuf<int> flag, rank_incl, dummy;
    const int Vn = V + E;
    c->sub_xyz.alloc(3 * (size_t)Vn); c->sub_parent1.alloc(Vn); c->sub_parent2.alloc(Vn);
    k_sub_old_points<<<grid_for(V), kThreads, 0, c->stream>>>(V, c->xyz.p, c->sub_xyz.p, c->sub_parent1.p, c->sub_parent2.p);
    ACVD_LAUNCH_CHECK();
    if (E > 0) {
        dummy.alloc(E);
        k_sub_edges<<<grid_for(E), kThreads, 0, c->stream>>>(E, V, T.first_sorted.p, c->tri.p, c->xyz.p, c->sub_xyz.p,
                                                              c->sub_parent1.p, c->sub_parent2.p);
        ACVD_LAUNCH_CHECK();
    }
    flag.alloc(F); rank_incl.alloc(F);
    k_sub_face_flags<<<grid_for(F), kThreads, 0, c->stream>>>(F, c->tri.p, flag.p);
    ACVD_LAUNCH_CHECK();
    inclusive_sum(c, flag.p, rank_incl.p, F);
    int n_kept = 0;
    ACVD_CUDA(cudaMemcpyAsync(&n_kept, rank_incl.p + (F - 1), sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    if (4 * (int64_t)n_kept >= ((int64_t)1 << 31) / 3) throw std::runtime_error("acvd_subdivide: too many faces");
    c->sub_tri.alloc(12 * (size_t)std::max(n_kept, 1));
    k_sub_faces<<<grid_for(F), kThreads, 0, c->stream>>>(F, V, c->tri.p, flag.p, rank_incl.p, T.edge_of_slot.p, c->sub_tri.p);
    ACVD_LAUNCH_CHECK();
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    c->sub_V = Vn; c->sub_F = 4 * n_kept;
    *n_vertices = Vn; *n_faces = 4 * n_kept;
    ACVD_API_END(c)
}

// vtkSurface::SplitLongEdges (Common/vtkSurface.cxx:444-604): passes of mark / cut on the device (mesh.cuh).  The result
// stays on the device (fetch with acvd_get_subdivision: parents = the end points of the cut edge, itself for old points).
extern "C" int acvd_split_long_edges(acvd_ctx* c, double ratio, int32_t* n_vertices, int32_t* n_faces, int32_t* n_passes) {
    ACVD_API_BEGIN(c)
    if (!c->V || !n_vertices || !n_faces || !(ratio > 0)) throw std::runtime_error("acvd_split_long_edges: bad arguments");
    int V = c->V, F = c->F;
    // working copies that grow pass by pass (points are appended in place; faces ping-pong)
    DevBuf<float> xyz;
    DevBuf<int> tri_a, tri_b, par1, par2, mark, rank_excl, n_child, child_excl;
    DevBuf<double> len, d_sum;
    size_t cap_v = (size_t)V + (size_t)V / 2 + 1024;
    xyz.alloc(3 * cap_v); par1.alloc(cap_v); par2.alloc(cap_v); tri_a.alloc(3 * (size_t)F);
    ACVD_CUDA(cudaMemcpyAsync(xyz.p, c->xyz.p, 3 * (size_t)V * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
    ACVD_CUDA(cudaMemcpyAsync(tri_a.p, c->tri.p, 3 * (size_t)F * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
    k_iota<<<grid_for(V), kThreads, 0, c->stream>>>(V, par1.p);
    k_iota<<<grid_for(V), kThreads, 0, c->stream>>>(V, par2.p);
    ACVD_LAUNCH_CHECK();
    int* tri = tri_a.p;
    DevBuf<int>* other = &tri_b;
    double threshold = 0;
    int passes = 0;
    d_sum.alloc(1);
    for (;; passes++) {
        if (passes >= 64) throw std::runtime_error("acvd_split_long_edges: more than 64 passes");
        EdgeTable T;
        build_edge_table(c, tri, F, T);
        const int E = T.E;
        if (E == 0) break;
        len.alloc(E); mark.alloc(E); rank_excl.alloc((size_t)E + 1);
        k_split_lengths<<<grid_for(E), kThreads, 0, c->stream>>>(E, T.first_sorted.p, tri, xyz.p, len.p);
        ACVD_LAUNCH_CHECK();
        if (passes == 0) {      // the threshold is fixed by the mesh as given (:455-462)
            size_t tb = 0;
            ACVD_CUDA(cub::DeviceReduce::Sum(nullptr, tb, len.p, d_sum.p, E, c->stream));
            void* t = cub_temp(c, tb);
            ACVD_CUDA(cub::DeviceReduce::Sum(t, tb, len.p, d_sum.p, E, c->stream));
            double total = 0;
            ACVD_CUDA(cudaMemcpyAsync(&total, d_sum.p, sizeof total, cudaMemcpyDeviceToHost, c->stream));
            ACVD_CUDA(cudaStreamSynchronize(c->stream));
            threshold = ratio * total / (double)E;
        }
        k_split_mark<<<grid_for(E), kThreads, 0, c->stream>>>(E, len.p, threshold, mark.p);
        ACVD_LAUNCH_CHECK();
        exclusive_sum(c, mark.p, rank_excl.p, E);
        int last_rank = 0, last_mark = 0;
        ACVD_CUDA(cudaMemcpyAsync(&last_rank, rank_excl.p + (E - 1), sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        ACVD_CUDA(cudaMemcpyAsync(&last_mark, mark.p + (E - 1), sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        ACVD_CUDA(cudaStreamSynchronize(c->stream));
        const int n_cut = last_rank + last_mark;
        if (n_cut == 0) break;
        if ((int64_t)V + n_cut >= ((int64_t)1 << 31)) throw std::runtime_error("acvd_split_long_edges: too many vertices");
        if ((size_t)V + n_cut > cap_v) {      // grow the point arrays
            const size_t cap_new = (size_t)V + n_cut + ((size_t)V + n_cut) / 2;
            DevBuf<float> x2; DevBuf<int> p1, p2;
            x2.alloc(3 * cap_new); p1.alloc(cap_new); p2.alloc(cap_new);
            ACVD_CUDA(cudaMemcpyAsync(x2.p, xyz.p, 3 * (size_t)V * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
            ACVD_CUDA(cudaMemcpyAsync(p1.p, par1.p, (size_t)V * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
            ACVD_CUDA(cudaMemcpyAsync(p2.p, par2.p, (size_t)V * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
            std::swap(xyz.p, x2.p); std::swap(xyz.n, x2.n); std::swap(par1.p, p1.p); std::swap(par1.n, p1.n); std::swap(par2.p, p2.p); std::swap(par2.n, p2.n);
            cap_v = cap_new;
        }
        k_split_points<<<grid_for(E), kThreads, 0, c->stream>>>(E, V, mark.p, rank_excl.p, T.first_sorted.p, tri, xyz.p, par1.p, par2.p);
        ACVD_LAUNCH_CHECK();
        n_child.alloc(F); child_excl.alloc((size_t)F + 1);
        k_split_count<<<grid_for(F), kThreads, 0, c->stream>>>(F, V, tri, T.edge_of_slot.p, mark.p, rank_excl.p, n_child.p);
        ACVD_LAUNCH_CHECK();
        exclusive_sum(c, n_child.p, child_excl.p, F);
        int last_off = 0, last_n = 0;
        ACVD_CUDA(cudaMemcpyAsync(&last_off, child_excl.p + (F - 1), sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        ACVD_CUDA(cudaMemcpyAsync(&last_n, n_child.p + (F - 1), sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        ACVD_CUDA(cudaStreamSynchronize(c->stream));
        const int64_t F_new = (int64_t)last_off + last_n;
        if (3 * F_new >= ((int64_t)1 << 31)) throw std::runtime_error("acvd_split_long_edges: too many faces");
        other->alloc(3 * (size_t)F_new);
        k_split_emit<<<grid_for(F), kThreads, 0, c->stream>>>(F, V, tri, T.edge_of_slot.p, mark.p, rank_excl.p, child_excl.p, other->p);
        ACVD_LAUNCH_CHECK();
        ACVD_CUDA(cudaStreamSynchronize(c->stream));
        tri = other->p;
        other = (other == &tri_b) ? &tri_a : &tri_b;
        V += n_cut; F = (int)F_new;
    }
    c->sub_xyz.alloc(3 * (size_t)V); c->sub_tri.alloc(3 * (size_t)F); c->sub_parent1.alloc(V); c->sub_parent2.alloc(V);
    ACVD_CUDA(cudaMemcpyAsync(c->sub_xyz.p, xyz.p, 3 * (size_t)V * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
    ACVD_CUDA(cudaMemcpyAsync(c->sub_tri.p, tri, 3 * (size_t)F * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
    ACVD_CUDA(cudaMemcpyAsync(c->sub_parent1.p, par1.p, (size_t)V * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
    ACVD_CUDA(cudaMemcpyAsync(c->sub_parent2.p, par2.p, (size_t)V * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    c->sub_V = V; c->sub_F = F;
    *n_vertices = V; *n_faces = F;
    if (n_passes) *n_passes = passes;
    ACVD_API_END(c)
}

extern "C" int acvd_get_subdivision(acvd_ctx* c, float* xyz, int32_t* tri, int32_t* parent1, int32_t* parent2) {
    ACVD_API_BEGIN(c)
    if (!c->sub_V) throw std::runtime_error("acvd_get_subdivision: call acvd_subdivide first");
    if (xyz) ACVD_CUDA(cudaMemcpyAsync(xyz, c->sub_xyz.p, 3 * (size_t)c->sub_V * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (tri) ACVD_CUDA(cudaMemcpyAsync(tri, c->sub_tri.p, 3 * (size_t)c->sub_F * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (parent1) ACVD_CUDA(cudaMemcpyAsync(parent1, c->sub_parent1.p, (size_t)c->sub_V * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (parent2) ACVD_CUDA(cudaMemcpyAsync(parent2, c->sub_parent2.p, (size_t)c->sub_V * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    ACVD_API_END(c)
}

// ---------------------------------------------------------------------------------------------
// curvature (the inputs of the gradation > 0 runs)
extern "C" int acvd_curvature(acvd_ctx* c, int32_t ring_size, double* indicator, float* info6) {
    ACVD_API_BEGIN(c)
    if (!c->V || !indicator) throw std::runtime_error("acvd_curvature: set the mesh first");
    if (ring_size < 1) throw std::runtime_error("acvd_curvature: ring_size must be >= 1");
    const int V = c->V;
    DevBuf<double> d_ind;
    DevBuf<float> d_info;
    DevBuf<int> d_big, d_scratch;
    d_ind.alloc(V); d_big.alloc(V);
    if (info6) d_info.alloc(6 * (size_t)V);
    ACVD_CUDA(cudaMemsetAsync(c->scalars.p, 0, 8 * sizeof(unsigned long long), c->stream));
    CurvMesh M{V, c->row_ptr.p, c->col.p, c->vf_ptr.p, c->vf_keys.p, c->xyz.p, c->tri.p};
    k_curvature<<<grid_for(V, 128, 16), 128, 0, c->stream>>>(M, ring_size, d_ind.p, info6 ? d_info.p : nullptr, d_big.p, c->scalars.p);
    ACVD_LAUNCH_CHECK();
    ACVD_CUDA(cudaMemcpyAsync(c->h_scalars, c->scalars.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    const int64_t n_big = (int64_t)c->h_scalars[0];
    if (n_big > 0) {   // neighbourhoods that did not fit the local list
        d_scratch.alloc((size_t)n_big * kCurvGlobalCap);
        int* d_failed = reinterpret_cast<int*>(c->scalars.p + 1);
        k_curvature_big<<<grid_for(n_big, 128, 16), 128, 0, c->stream>>>(M, ring_size, d_ind.p, info6 ? d_info.p : nullptr, d_big.p, (int)n_big,
                                                                         d_scratch.p, d_failed);
        ACVD_LAUNCH_CHECK();
        ACVD_CUDA(cudaMemcpyAsync(c->h_scalars + 1, c->scalars.p + 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        ACVD_CUDA(cudaStreamSynchronize(c->stream));
        if (c->h_scalars[1] != 0) throw std::runtime_error("acvd_curvature: a vertex neighbourhood exceeds 8192 vertices");
    }
    ACVD_CUDA(cudaMemcpyAsync(indicator, d_ind.p, (size_t)V * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (info6) ACVD_CUDA(cudaMemcpyAsync(info6, d_info.p, 6 * (size_t)V * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    ACVD_API_END(c)
}

// ---------------------------------------------------------------------------------------------
// items
template <int M>
static void compose_items(acvd_ctx* c, const double* d_sum, double ratio, const float* d_pd) {
    k_compose_items<M><<<grid_for(c->V), kThreads, 0, c->stream>>>(c->V, d_sum, ratio, c->weight.p, c->area.p, c->xyz.p,
                                                                  c->tri.p, c->vf_ptr.p, c->vf_keys.p, d_pd, c->items.p);
    ACVD_LAUNCH_CHECK();
}

static void compute_areas(acvd_ctx* c) {
    c->area.alloc(c->V);
    k_vertex_area<<<grid_for(c->V), kThreads, 0, c->stream>>>(c->V, c->vf_ptr.p, c->vf_keys.p, c->xyz.p, c->tri.p, c->area.p);
    ACVD_LAUNCH_CHECK();
}

extern "C" int acvd_build_items(acvd_ctx* c, int metric, double gradation, const double* custom, const float* pd) {
    ACVD_API_BEGIN(c)
    if (!c->V) throw std::runtime_error("acvd_build_items: set the mesh first");
    if (metric < 0 || metric > 3) throw std::runtime_error("acvd_build_items: unknown metric");
    const int V = c->V;
    const bool aniso = (metric == M_ANISO || metric == M_ANISOQ);
    if (aniso && !pd) throw std::runtime_error("acvd_build_items: anisotropic metrics need principal directions");
    // which runs read the indicator: iso when given (vtkIsotropicMetric...:256-259), qem when gradation > 0 (:333-334),
    // anisotropic when gradation != 0 (:384-392)
    bool use_custom = custom && (metric == M_ISO ? true : (metric == M_QEM ? gradation > 0 : gradation != 0));
    if (!custom && ((metric == M_QEM && gradation > 0) || (aniso && gradation != 0)))
        throw std::runtime_error("acvd_build_items: gradation needs custom_weights (curvature indicator)");
    c->metric = metric;
    c->fx_scale = 0.0;                         // the bulk rounds' fixed-point scale depends on the items
    c->weight.alloc((size_t)c->vpad);
    c->items.alloc((size_t)V * payload_npad(metric));
    compute_areas(c);
    DevBuf<double> d_custom, d_sum;
    DevBuf<float> d_pd;
    if (use_custom) {
        d_custom.alloc(V);
        ACVD_CUDA(cudaMemcpyAsync(d_custom.p, custom, (size_t)V * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    }
    if (aniso) {
        d_pd.alloc(6 * (size_t)V);
        ACVD_CUDA(cudaMemcpyAsync(d_pd.p, pd, 6 * (size_t)V * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    }
    k_raw_weight<<<grid_for(V), kThreads, 0, c->stream>>>(V, c->area.p, d_custom.p, gradation, use_custom ? 1 : 0, aniso ? 1 : 0, c->weight.p);
    ACVD_LAUNCH_CHECK();
    d_sum.alloc(1);
    size_t tb = 0;
    ACVD_CUDA(cub::DeviceReduce::Sum(nullptr, tb, c->weight.p, d_sum.p, V, c->stream));
    void* t = cub_temp(c, tb);
    ACVD_CUDA(cub::DeviceReduce::Sum(t, tb, c->weight.p, d_sum.p, V, c->stream));
    const double ratio = (metric == M_QEM) ? 1e4 : 1e5;   // vtkQEMetricForClustering.h:338 vs 1e5 elsewhere
    switch (metric) {
        case M_ISO: compose_items<M_ISO>(c, d_sum.p, ratio, nullptr); break;
        case M_QEM: compose_items<M_QEM>(c, d_sum.p, ratio, nullptr); break;
        case M_ANISO: compose_items<M_ANISO>(c, d_sum.p, ratio, d_pd.p); break;
        default: compose_items<M_ANISOQ>(c, d_sum.p, ratio, d_pd.p); break;
    }
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    c->have_items = true; c->stats_valid = false;
    ACVD_API_END(c)
}

__global__ void k_extract_weight(int V, int npad, const double* __restrict__ items, double* w) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) w[v] = items[(int64_t)v * npad + 3];
}

extern "C" int acvd_set_items(acvd_ctx* c, int metric, const double* payload) {
    ACVD_API_BEGIN(c)
    if (!c->V || !payload || metric < 0 || metric > 3) throw std::runtime_error("acvd_set_items: bad arguments");
    const int V = c->V, np = payload_np(metric), npad = payload_npad(metric);
    c->metric = metric;
    c->fx_scale = 0.0;
    c->items.alloc((size_t)V * npad);
    c->weight.alloc((size_t)c->vpad);
    DevBuf<double> tmp;
    tmp.alloc((size_t)V * np);
    ACVD_CUDA(cudaMemcpyAsync(tmp.p, payload, (size_t)V * np * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    k_pack_rows<<<grid_for((int64_t)V * npad), kThreads, 0, c->stream>>>(V, np, npad, tmp.p, c->items.p, 1);
    ACVD_LAUNCH_CHECK();
    k_extract_weight<<<grid_for(V), kThreads, 0, c->stream>>>(V, npad, c->items.p, c->weight.p);
    ACVD_LAUNCH_CHECK();
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    c->have_items = true; c->stats_valid = false;
    ACVD_API_END(c)
}

extern "C" int acvd_get_items(acvd_ctx* c, double* payload) {
    ACVD_API_BEGIN(c)
    if (!c->have_items || !payload) throw std::runtime_error("acvd_get_items: no items");
    const int V = c->V, np = payload_np(c->metric), npad = payload_npad(c->metric);
    DevBuf<double> tmp;
    tmp.alloc((size_t)V * np);
    k_pack_rows<<<grid_for((int64_t)V * npad), kThreads, 0, c->stream>>>(V, np, npad, c->items.p, tmp.p, 0);
    ACVD_LAUNCH_CHECK();
    ACVD_CUDA(cudaMemcpyAsync(payload, tmp.p, (size_t)V * np * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    ACVD_API_END(c)
}

extern "C" int acvd_get_vertex_areas(acvd_ctx* c, double* areas) {
    ACVD_API_BEGIN(c)
    if (!c->V || !areas) throw std::runtime_error("acvd_get_vertex_areas: no mesh");
    compute_areas(c);
    ACVD_CUDA(cudaMemcpyAsync(areas, c->area.p, (size_t)c->V * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    ACVD_API_END(c)
}

// ---------------------------------------------------------------------------------------------
// clusters
static void set_num_clusters_impl(acvd_ctx* c, int32_t K) {
    if (K <= 0) throw std::runtime_error("acvd_set_num_clusters: K must be positive");
    if (!c->have_items) throw std::runtime_error("acvd_set_num_clusters: build or set the items first");
    const int V = c->V, npad = payload_npad(c->metric);
    c->K = K;
    c->cid.alloc((size_t)c->vpad);
    c->csize.alloc(K); c->mod_round.alloc(K); c->anchor.alloc(K); c->frozen.alloc(K);
    const size_t Kp = (size_t)K + 64;   // room for the equal-chunk in-place all-gather of the statistics (world <= 64)
    c->csum.alloc(Kp * npad); c->cenergy.alloc(Kp); c->ccentroid.alloc(3 * Kp);
    c->isum.alloc(4 * (size_t)K); c->bulk_cen.alloc(4 * (size_t)K + 4); c->cmeta.alloc((size_t)K + 1); c->bulk_energy.alloc(K); c->bulk_energy_sum.alloc(1); c->leave_cnt.alloc(K); c->join_cnt.alloc(K);
    c->best.alloc(K); c->modbits.alloc((size_t)(K + 31) / 32 + 1); c->prop_key.alloc(V); c->prop_dst.alloc(V); c->plist.alloc(V); c->plist_b.alloc(V); c->work.alloc(V);
    {
        const size_t n_tiles = ((size_t)V + 31) / 32;
        c->tile_sig.alloc(n_tiles * kSigSlots); c->tile_active.alloc(n_tiles); c->tile_stale.alloc(n_tiles); c->prop_mask.alloc(n_tiles);
        c->moved_mask.alloc(n_tiles);
        ACVD_CUDA(cudaMemsetAsync(c->prop_mask.p, 0, n_tiles * sizeof(unsigned), c->stream));
        ACVD_CUDA(cudaMemsetAsync(c->tile_stale.p, 1, n_tiles, c->stream)); c->active_tiles.alloc(n_tiles); c->round_scalars.alloc(2);
        ACVD_CUDA(cudaMemsetAsync(c->round_scalars.p, 0, 2 * sizeof(unsigned long long), c->stream));
    } c->prop_e.alloc(V);
    c->best2.alloc(K); c->stamp.alloc(V); c->modlist0.alloc((size_t)K + 64); c->modlist1.alloc((size_t)K + 64);
    c->sp_nmod.alloc(kSparseChunkAlloc + 2); c->memb_overflow.alloc(1);
    ACVD_CUDA(cudaMemsetAsync(c->stamp.p, 0, (size_t)V * sizeof(int), c->stream));
    ACVD_CUDA(cudaMemsetAsync(c->memb_overflow.p, 0, sizeof(int), c->stream));
    c->members_valid = false; c->modlist_valid = false; c->mod_par = 0;
    ACVD_CUDA(cudaMemsetAsync(c->prop_dst.p, 0xff, (size_t)V * sizeof(int), c->stream));
    ACVD_CUDA(cudaMemsetAsync(c->mod_round.p, 0, (size_t)K * sizeof(int), c->stream));
    ACVD_CUDA(cudaMemsetAsync(c->anchor.p, 0xff, (size_t)K * sizeof(int), c->stream));
    ACVD_CUDA(cudaMemsetAsync(c->frozen.p, 0, (size_t)K, c->stream));
    ACVD_CUDA(cudaMemsetAsync(c->csize.p, 0, (size_t)K * sizeof(int), c->stream));
    k_fill<<<grid_for(c->vpad), kThreads, 0, c->stream>>>((int)c->vpad, K, c->cid.p);     // every vertex in the NULL cluster
    ACVD_LAUNCH_CHECK();
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    c->has_frozen = c->has_anchor = false;
    c->fixed.clear();
    c->round = 1; c->cc_since = 0;
    c->stats_valid = false;
}

extern "C" int acvd_set_num_clusters(acvd_ctx* c, int32_t K) {
    ACVD_API_BEGIN(c)
    set_num_clusters_impl(c, K);
    ACVD_API_END(c)
}

extern "C" int acvd_set_clustering(acvd_ctx* c, const int32_t* cl) {
    ACVD_API_BEGIN(c)
    c->cc_since = 0;      // the clustering changes outside the rounds: the next CleanClustering checks every cluster
    if (!c->K || !cl) throw std::runtime_error("acvd_set_clustering: set the number of clusters first");
    ACVD_CUDA(cudaMemcpyAsync(c->cid.p, cl, (size_t)c->V * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    // ids outside [0, K) are "not assigned" for the reference (:560-561): one NULL value, K, inside the library
    k_normalise_null<<<grid_for(c->V), kThreads, 0, c->stream>>>(c->V, c->K, c->cid.p);
    ACVD_LAUNCH_CHECK();
    ACVD_CUDA(cudaMemsetAsync(c->prop_dst.p, 0xff, (size_t)c->V * sizeof(int), c->stream));
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    c->stats_valid = false; c->sig_valid = false; c->members_valid = false; c->modlist_valid = false;
    ACVD_API_END(c)
}

extern "C" int acvd_get_clustering(acvd_ctx* c, int32_t* cl) {
    ACVD_API_BEGIN(c)
    if (!c->K || !cl) throw std::runtime_error("acvd_get_clustering: no clustering");
    ACVD_CUDA(cudaMemcpyAsync(cl, c->cid.p, (size_t)c->V * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    ACVD_API_END(c)
}

extern "C" int acvd_save_clustering(acvd_ctx* c) {
    ACVD_API_BEGIN(c)
    if (!c->K) throw std::runtime_error("acvd_save_clustering: no clustering");
    c->cid_saved.alloc(c->V);
    ACVD_CUDA(cudaMemcpyAsync(c->cid_saved.p, c->cid.p, (size_t)c->V * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    ACVD_API_END(c)
}

extern "C" int acvd_restore_clustering(acvd_ctx* c) {
    ACVD_API_BEGIN(c)
    c->cc_since = 0;      // the clustering changes outside the rounds: the next CleanClustering checks every cluster
    if (!c->K || !c->cid_saved.p) throw std::runtime_error("acvd_restore_clustering: nothing saved");
    ACVD_CUDA(cudaMemcpyAsync(c->cid.p, c->cid_saved.p, (size_t)c->V * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
    ACVD_CUDA(cudaMemsetAsync(c->prop_dst.p, 0xff, (size_t)c->V * sizeof(int), c->stream));
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    c->stats_valid = false; c->sig_valid = false; c->members_valid = false; c->modlist_valid = false;
    ACVD_API_END(c)
}

extern "C" int acvd_set_frozen(acvd_ctx* c, const uint8_t* frozen) {
    ACVD_API_BEGIN(c)
    if (!c->K) throw std::runtime_error("acvd_set_frozen: set the number of clusters first");
    if (frozen) {
        ACVD_CUDA(cudaMemcpy(c->frozen.p, frozen, (size_t)c->K, cudaMemcpyHostToDevice));
        c->has_frozen = std::any_of(frozen, frozen + c->K, [](uint8_t f) { return f != 0; });
    } else {
        ACVD_CUDA(cudaMemset(c->frozen.p, 0, (size_t)c->K));
        c->has_frozen = false;
    }
    ACVD_API_END(c)
}

extern "C" int acvd_get_frozen(acvd_ctx* c, uint8_t* frozen) {
    ACVD_API_BEGIN(c)
    if (!c->K || !frozen) throw std::runtime_error("acvd_get_frozen: no clustering");
    ACVD_CUDA(cudaMemcpy(frozen, c->frozen.p, (size_t)c->K, cudaMemcpyDeviceToHost));
    ACVD_API_END(c)
}

extern "C" int acvd_set_fixed_clusters(acvd_ctx* c, const int64_t* items, int32_t n) {
    ACVD_API_BEGIN(c)
    if (!c->K || n < 0 || n > c->K || (n && !items)) throw std::runtime_error("acvd_set_fixed_clusters: bad arguments");
    std::vector<int> a(c->K, -1);
    c->fixed.assign(items, items + n);
    for (int i = 0; i < n; i++) {
        if (items[i] < 0 || items[i] >= c->V) throw std::runtime_error("acvd_set_fixed_clusters: item out of range");
        a[i] = (int)items[i];
    }
    ACVD_CUDA(cudaMemcpy(c->anchor.p, a.data(), (size_t)c->K * sizeof(int), cudaMemcpyHostToDevice));
    c->has_anchor = n > 0;
    c->stats_valid = false;
    ACVD_API_END(c)
}

extern "C" int acvd_initial_sampling(acvd_ctx* c) {
    ACVD_API_BEGIN(c)
    c->cc_since = 0;      // the clustering changes outside the rounds: the next CleanClustering checks every cluster
    if (!c->K || !c->have_items) throw std::runtime_error("acvd_initial_sampling: need items and a cluster count");
    // the rings in the reference's order come from the device (k_ring_order: neighbours sorted by the first half-edge
    // slot of their edge), so the host does not rebuild the edge table; the region growing itself is sequential
    const int V = c->V;
    DevBuf<int> ring_col;
    ring_col.alloc((size_t)std::max<int64_t>(c->nnz, 1));
    FillMesh M{V, c->K, c->row_ptr.p, c->col.p, c->vf_ptr.p, c->vf_keys.p, c->tri.p};
    k_ring_order<<<grid_for(V), kThreads, 0, c->stream>>>(M, ring_col.p);
    ACVD_LAUNCH_CHECK();
    std::vector<double> w((size_t)V);
    std::vector<int> h_ptr((size_t)V + 1), h_nbr((size_t)c->nnz);
    ACVD_CUDA(cudaMemcpyAsync(w.data(), c->weight.p, (size_t)V * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaMemcpyAsync(h_ptr.data(), c->row_ptr.p, ((size_t)V + 1) * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaMemcpyAsync(h_nbr.data(), ring_col.p, (size_t)c->nnz * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    std::vector<int> out;
    initial_random_sampling(V, c->K, FlatRings{h_ptr.data(), h_nbr.data()}, w.data(), c->fixed, out);
    ACVD_CUDA(cudaMemcpy(c->cid.p, out.data(), (size_t)c->V * sizeof(int), cudaMemcpyHostToDevice));
    ACVD_CUDA(cudaMemset(c->prop_dst.p, 0xff, (size_t)c->V * sizeof(int)));
    c->stats_valid = false; c->sig_valid = false; c->members_valid = false; c->modlist_valid = false;
    ACVD_API_END(c)
}

// ---------------------------------------------------------------------------------------------
// statistics / clean / fill
static EvalCfg make_cfg(int constrained, int qlevel, double thr) {
    EvalCfg cfg;
    cfg.constrained = constrained; cfg.qlevel = qlevel; cfg.thr = thr > 0 ? thr : 1e-3;
    return cfg;
}

// QEM in its unconstrained phase without anchors evaluates the isotropic energy
static bool qem_as_iso(const acvd_ctx* c, int constrained, int qlevel) {
    return c->metric == M_QEM && !c->has_anchor && (!constrained || !qlevel);
}

// ---- per-cluster member arrays (sparse.cuh): counting build, no sort
static void members_build(acvd_ctx* c) {
    const int V = c->V, K = c->K;
    c->memb_off.alloc((size_t)K + 2); c->memb_cap.alloc((size_t)K + 2); c->memb_pos.alloc(V);
    ACVD_CUDA(cudaMemsetAsync(c->memb_cap.p, 0, ((size_t)K + 2) * sizeof(int), c->stream));
    k_members_count<<<grid_for(V), kThreads, 0, c->stream>>>(V, K, c->cid.p, c->memb_cap.p);
    ACVD_LAUNCH_CHECK();
    k_members_cap<<<grid_for(K + 1), kThreads, 0, c->stream>>>(K, c->memb_cap.p, c->memb_cap.p);
    ACVD_LAUNCH_CHECK();
    size_t tb = 0;
    ACVD_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, c->memb_cap.p, c->memb_off.p, K + 1, c->stream));
    void* t = cub_temp(c, tb);
    ACVD_CUDA(cub::DeviceScan::ExclusiveSum(t, tb, c->memb_cap.p, c->memb_off.p, K + 1, c->stream));
    // total slots <= 2 V + 16 K: allocate the bound, no host round trip
    const size_t slots = 2 * (size_t)V + 16 * (size_t)K + 64;
    c->memb.alloc(slots); c->memb_tmp.alloc(slots); c->cc_par.alloc(slots); c->cc_sz.alloc(slots);
    ACVD_CUDA(cudaMemsetAsync(c->csize.p, 0, (size_t)K * sizeof(int), c->stream));
    k_members_scatter<<<grid_for(V), kThreads, 0, c->stream>>>(V, K, c->cid.p, c->memb_off.p, c->csize.p, c->memb.p, c->memb_pos.p);
    ACVD_LAUNCH_CHECK();
    ACVD_CUDA(cudaMemsetAsync(c->memb_overflow.p, 0, sizeof(int), c->stream));
    c->members_valid = true;
}

// one pass over the clusters [k_begin, k_end): sort members [+ connected components / cleaning] [+ statistics]
static void cluster_pass(acvd_ctx* c, bool do_cc, bool do_stats, int constrained, int qlevel, double thr, int k_begin = 0, int k_end = -1,
                         bool apply_resets = true) {
    if (k_end < 0) k_end = c->K;
    ClusterPassArgs P;
    P.V = c->V; P.K = c->K; P.off = c->memb_off.p; P.memb = c->memb.p; P.memb_tmp = c->memb_tmp.p; P.pos = c->memb_pos.p;
    P.cid = c->cid.p; P.csize = c->csize.p; P.row_ptr = c->row_ptr.p; P.col = c->col.p; P.items = c->items.p;
    P.csum = c->csum.p; P.cenergy = c->cenergy.p; P.ccentroid = c->ccentroid.p;
    P.anchor = c->has_anchor ? c->anchor.p : nullptr; P.xyz = c->xyz.p;
    P.cc_par = c->cc_par.p; P.cc_sz = c->cc_sz.p; P.counters = c->scalars.p + 1;
    P.do_sort = 1; P.do_cc = do_cc ? 1 : 0; P.do_stats = do_stats ? 1 : 0;
    P.apply_resets = apply_resets ? 1 : 0; P.k_begin = k_begin; P.k_end = k_end;
    P.mod_round = c->mod_round.p; P.cc_since = getenv("ACVD_CC_ALL") ? 0 : c->cc_since;
    P.cfg = make_cfg(constrained, qlevel, thr);
    const int blocks = grid_for((int64_t)std::max(1, k_end - k_begin) * 32);
#define PASS(MM, EE) k_cluster_pass<MM, EE><<<blocks, kThreads, 0, c->stream>>>(P)
    switch (c->metric) {
        case M_ISO: PASS(M_ISO, M_ISO); break;
        case M_QEM:
            if (qem_as_iso(c, constrained, qlevel)) PASS(M_QEM, M_ISO);   // same formula as the rounds use in this phase
            else PASS(M_QEM, M_QEM);
            break;
        case M_ANISO: PASS(M_ANISO, M_ANISO); break;
        default: PASS(M_ANISOQ, M_ANISOQ); break;
    }
#undef PASS
    ACVD_LAUNCH_CHECK();
    if (do_stats) { c->stats_valid = true; c->stats_constrained = constrained; c->stats_qlevel = qlevel; }
}

// multi-GPU: the cluster range this rank runs the cluster pass on (equal chunks: the results are all-gathered in place)
static int dist_cluster_chunk(const acvd_ctx* c) { return (c->K + c->world - 1) / c->world; }   // = acvd_dist_partition out[2..3]
static void dist_allgather_stats(acvd_ctx* c);     // dist.cuh

// ReComputeStatistics (:376-403) + ReComputeClustersSize (:353-373)
static void recompute_statistics(acvd_ctx* c, int constrained, int qlevel, double thr) {
    TraceScope ts(c, "recompute_statistics");
    members_build(c);
    if (c->world > 1) {     // every rank accumulates its share of the clusters; NCCL all-gather of the per-cluster statistics
        const int chunk = dist_cluster_chunk(c), k0 = std::min(c->K, c->rank * chunk), k1 = std::min(c->K, k0 + chunk);
        cluster_pass(c, false, true, constrained, qlevel, thr, k0, k1);
        dist_allgather_stats(c);
    } else cluster_pass(c, false, true, constrained, qlevel, thr);
}

extern "C" int acvd_recompute_statistics(acvd_ctx* c, int constrained, int qlevel) {
    ACVD_API_BEGIN(c)
    if (!c->K) throw std::runtime_error("acvd_recompute_statistics: no clustering");
    recompute_statistics(c, constrained, qlevel, 0);
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    ACVD_API_END(c)
}

// CleanClustering (:406-549) inside the cluster pass.  With `with_stats` the same pass also leaves fresh statistics,
// valid when nothing had to be cleaned (n_reset == 0: the caller checks stats_valid).
static void dist_allreduce_counters(acvd_ctx* c, unsigned long long* d, int n);   // dist.cuh
static int clean_clustering(acvd_ctx* c, bool with_stats = false, int constrained = 1, int qlevel = 3, double thr = 0) {
    TraceScope ts(c, "clean_clustering");
    if (!c->members_valid) members_build(c);
    ACVD_CUDA(cudaMemsetAsync(c->scalars.p, 0, 8 * sizeof(unsigned long long), c->stream));
    bool sharded = c->world > 1;
    if (sharded) {
        // every rank checks (and accumulates) its share of the clusters without touching the clustering; the counts are
        // summed over the ranks.  Only when some cluster really is disconnected does every rank run the whole pass.
        const int chunk = dist_cluster_chunk(c), k0 = std::min(c->K, c->rank * chunk), k1 = std::min(c->K, k0 + chunk);
        cluster_pass(c, true, with_stats, constrained, qlevel, thr, k0, k1, false);
        dist_allreduce_counters(c, c->scalars.p + 1, 2);
    } else cluster_pass(c, true, with_stats, constrained, qlevel, thr);
    ACVD_CUDA(cudaMemcpyAsync(c->h_scalars, c->scalars.p, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    int disc = (int)c->h_scalars[1], n_reset = (int)c->h_scalars[2];
    if (sharded) {
        if (n_reset > 0) {
            ACVD_CUDA(cudaMemsetAsync(c->scalars.p, 0, 8 * sizeof(unsigned long long), c->stream));
            cluster_pass(c, true, false, constrained, qlevel, thr);
            ACVD_CUDA(cudaMemcpyAsync(c->h_scalars, c->scalars.p, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
            ACVD_CUDA(cudaStreamSynchronize(c->stream));
            disc = (int)c->h_scalars[1]; n_reset = (int)c->h_scalars[2];
            c->stats_valid = false;
        } else if (with_stats) dist_allgather_stats(c);
    }
    if (n_reset > 0) { c->stats_valid = false; c->members_valid = false; c->sig_valid = false; }
    if (trace_on()) fprintf(stderr, "[acvd trace]   disconnected %d, reset %d\n", disc, n_reset);
    c->cc_since = c->round;      // every cluster is connected now: the next check looks at the clusters modified from here on
    return disc;
}

// FillHolesInClustering, order-exact (fill.cuh).  `connexity` = the engine's ConnexityConstraint at the time of the call.
static void fill_holes(acvd_ctx* c, int connexity) {
    TraceScope ts(c, "fill_holes");
    const int V = c->V, K = c->K;
    c->null_list.alloc(V);
    ACVD_CUDA(cudaMemsetAsync(c->scalars.p, 0, 8 * sizeof(unsigned long long), c->stream));
    k_collect_null<<<grid_for(V), kThreads, 0, c->stream>>>(V, K, c->cid.p, c->null_list.p, c->scalars.p);
    ACVD_LAUNCH_CHECK();
    ACVD_CUDA(cudaMemcpyAsync(c->h_scalars, c->scalars.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    const int n = (int)c->h_scalars[0];
    if (n == 0) return;
    if (trace_on()) fprintf(stderr, "[acvd trace]   fill: %d NULL vertices, connexity %d\n", n, connexity);
    c->stats_valid = false; c->members_valid = false; c->sig_valid = false;
    FillMesh M{V, K, c->row_ptr.p, c->col.p, c->vf_ptr.p, c->vf_keys.p, c->tri.p};
    size_t tb = 0;
    if (connexity && n <= kFillSequentialCap) {
        // ---- replay of the reference's FIFO by one thread on the slot-sorted initial edges
        DevBuf<unsigned> slot0, slot1;
        DevBuf<int> idx0, idx1;
        DevBuf<int2> e0, q;
        const size_t cap_init = (size_t)n * (size_t)std::max(c->max_deg, 1);
        slot0.alloc(cap_init); slot1.alloc(cap_init); idx0.alloc(cap_init); idx1.alloc(cap_init); e0.alloc(cap_init);
        ACVD_CUDA(cudaMemsetAsync(c->scalars.p, 0, 8 * sizeof(unsigned long long), c->stream));
        k_fill_initial_edges<<<grid_for(n), kThreads, 0, c->stream>>>(M, n, c->null_list.p, c->cid.p, slot0.p, e0.p, c->scalars.p);
        ACVD_LAUNCH_CHECK();
        ACVD_CUDA(cudaMemcpyAsync(c->h_scalars, c->scalars.p, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        ACVD_CUDA(cudaStreamSynchronize(c->stream));
        const int n_init = (int)c->h_scalars[0];
        const long long cap = (long long)n_init + (long long)c->h_scalars[1];
        if (n_init == 0) return;   // unreachable holes stay NULL, as in the reference (:619-631)
        k_iota<<<grid_for(n_init), kThreads, 0, c->stream>>>(n_init, idx0.p);
        ACVD_LAUNCH_CHECK();
        ACVD_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, slot0.p, slot1.p, idx0.p, idx1.p, n_init, 0, 32, c->stream));
        void* t = cub_temp(c, tb);
        ACVD_CUDA(cub::DeviceRadixSort::SortPairs(t, tb, slot0.p, slot1.p, idx0.p, idx1.p, n_init, 0, 32, c->stream));
        q.alloc((size_t)cap);
        k_gather_int2<<<grid_for(n_init), kThreads, 0, c->stream>>>(n_init, idx1.p, e0.p, q.p);
        ACVD_LAUNCH_CHECK();
        k_fill_sequential<<<1, 32, 0, c->stream>>>(M, n_init, cap, q.p, c->cid.p, connexity, c->scalars.p + 4);
        ACVD_LAUNCH_CHECK();
        ACVD_CUDA(cudaMemcpyAsync(c->h_scalars + 4, c->scalars.p + 4, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        ACVD_CUDA(cudaStreamSynchronize(c->stream));
        if (c->h_scalars[5] == 0) return;
        // a ring longer than kMaxRing among the NULL vertices: finish with the level-synchronous passes below
    }
    // ---- level-synchronous passes (exact while the connexity guard is off)
    // the list order depends on atomics; sort it so nothing below depends on it
    c->sort_k0.alloc(V);
    ACVD_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, c->null_list.p, c->sort_k0.p, n, 0, 32, c->stream));
    void* t = cub_temp(c, tb);
    ACVD_CUDA(cub::DeviceRadixSort::SortKeys(t, tb, c->null_list.p, c->sort_k0.p, n, 0, 32, c->stream));
    ACVD_CUDA(cudaMemcpyAsync(c->null_list.p, c->sort_k0.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
    c->pick.alloc(n); c->fill_lvl.alloc(V); c->fill_seq.alloc(V);
    DevBuf<unsigned long long> key, fkey0, fkey1;
    DevBuf<int> fv0, fv1;
    key.alloc(n); fkey0.alloc(n); fkey1.alloc(n); fv0.alloc(n); fv1.alloc(n);
    ACVD_CUDA(cudaMemsetAsync(c->fill_lvl.p, 0, (size_t)V * sizeof(int), c->stream));
    int64_t remaining = n;
    int base = 0;
    for (int level = 1; remaining > 0; level++) {
        ACVD_CUDA(cudaMemsetAsync(c->scalars.p + 3, 0, sizeof(unsigned long long), c->stream));
        // with the guard on every pass looks at all assigned neighbours again (a refused vertex may pass later)
        k_fill_level_pick<<<grid_for(n), kThreads, 0, c->stream>>>(M, n, c->null_list.p, c->cid.p, c->fill_lvl.p, c->fill_seq.p,
                                                                   connexity ? 1 : level, connexity, key.p, c->pick.p);
        ACVD_LAUNCH_CHECK();
        k_fill_level_apply<<<grid_for(n), kThreads, 0, c->stream>>>(n, c->null_list.p, key.p, c->pick.p, level, c->cid.p, c->fill_lvl.p,
                                                                    fkey0.p, fv0.p, c->scalars.p + 3);
        ACVD_LAUNCH_CHECK();
        ACVD_CUDA(cudaMemcpyAsync(c->h_scalars + 3, c->scalars.p + 3, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        ACVD_CUDA(cudaStreamSynchronize(c->stream));
        const int64_t filled = (int64_t)c->h_scalars[3];
        if (filled == 0) break;   // unreachable holes stay NULL, as in the reference (:619-631)
        remaining -= filled;
        if (remaining > 0 && !connexity) {   // adoption sequence numbers of this level order the next level's ties
            ACVD_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, fkey0.p, fkey1.p, fv0.p, fv1.p, (int)filled, 0, 64, c->stream));
            t = cub_temp(c, tb);
            ACVD_CUDA(cub::DeviceRadixSort::SortPairs(t, tb, fkey0.p, fkey1.p, fv0.p, fv1.p, (int)filled, 0, 64, c->stream));
            k_fill_level_seq<<<grid_for(filled), kThreads, 0, c->stream>>>((int)filled, fv1.p, base, c->fill_seq.p);
            ACVD_LAUNCH_CHECK();
            base += (int)filled;
        }
    }
}

extern "C" int acvd_clean_clustering(acvd_ctx* c, int32_t* disconnected) {
    ACVD_API_BEGIN(c)
    if (!c->K) throw std::runtime_error("acvd_clean_clustering: no clustering");
    int d = clean_clustering(c);
    if (disconnected) *disconnected = d;
    ACVD_API_END(c)
}

extern "C" int acvd_fill_holes(acvd_ctx* c, int connexity) {
    ACVD_API_BEGIN(c)
    if (!c->K) throw std::runtime_error("acvd_fill_holes: no clustering");
    fill_holes(c, connexity);
    ACVD_API_END(c)
}

// ---------------------------------------------------------------------------------------------
// reassignment rounds
struct RoundResult { unsigned long long proposals, mods, tests, evaluated, boundary, active_tiles, members; float ms_scan, ms_eval, ms_commit; bool overflow, sparse; };

static ReassignArgs make_args(acvd_ctx* c, const EvalCfg& cfg, int connexity, int force_all) {
    ReassignArgs A;
    A.V = c->V; A.K = c->K;
    A.row_ptr = c->row_ptr.p; A.col = c->col.p; A.ell = c->ell.p; A.ringadj = c->ringadj.p; A.vpad = c->vpad; A.cid = c->cid.p;
    A.items = c->items.p; A.csum = c->csum.p; A.cenergy = c->cenergy.p; A.csize = c->csize.p;
    A.mod_round = c->mod_round.p;
    A.modbits = c->modbits.p; A.cmeta = c->cmeta.p; A.blist = c->blist.p; A.blist_cnt = c->blist_cnt.p;
    A.frozen = c->has_frozen ? c->frozen.p : nullptr;
    A.anchor = c->has_anchor ? c->anchor.p : nullptr;
    A.xyz = c->xyz.p;
    A.best = c->best.p; A.prop_dst = c->prop_dst.p; A.prop_key = c->prop_key.p; A.prop_e = c->prop_e.p;
    A.prop_mask = c->prop_mask.p; A.moved_mask = nullptr; A.weight = c->weight.p;
    A.plist = c->plist_cur ? c->plist_b.p : c->plist.p;
    A.plist_prev = c->plist_cur ? c->plist.p : c->plist_b.p;
    A.n_prev_props = c->round_scalars.p + 1;
    A.tile_sig = c->tile_sig.p; A.tile_active = c->tile_active.p; A.tile_stale = c->tile_stale.p; A.active_tiles = c->active_tiles.p;
    A.n_active_tiles = c->round_scalars.p;
    A.work = c->work.p; A.ctr = c->ctr.p;
    A.round = c->round; A.force_all = force_all; A.bulk = 0; A.connexity = connexity; A.cfg = cfg;
    A.bulk_stage = 0; A.bulk_count_leave = 0; A.bulk_cen = c->bulk_cen.p; A.bulk_leave = c->leave_cnt.p;
    A.item_stride = payload_npad(c->metric);
    A.all_tiles = 0; A.tile_begin = 0; A.tile_end = (c->V + 31) / 32; A.sig_mode = 0; A.track_stale = 1;
    if (c->members_valid) A.mem = Members{c->memb_off.p, c->memb.p, c->memb_pos.p, c->memb_overflow.p};
    else A.mem = Members{nullptr, nullptr, nullptr, nullptr};
    A.modlist = nullptr; A.n_mod = nullptr; A.stamp = c->stamp.p;
    return A;
}

// dense bulk rounds: the TMA-staged streaming scan (scan_dense.cuh), one wave of MINB blocks per SM
template <int W, int S, int MINB>
static void launch_scan_bulk_dense(acvd_ctx* c, const ReassignArgs& A) {
    static bool configured[64] = {};           // the attribute is per device
    if (c->device >= 64 || !configured[c->device]) {
        ACVD_CUDA(cudaFuncSetAttribute(k_scan_bulk_dense<W, S, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, dense_smem_bytes(W, S)));
        if (c->device < 64) configured[c->device] = true;
    }
    const int n_tiles = A.tile_end - A.tile_begin;
    const int grid = std::max(1, std::min(kNumSMs * MINB, (n_tiles + kDenseWarps - 1) / kDenseWarps));
    k_scan_bulk_dense<W, S, MINB><<<grid, kDenseThreads, dense_smem_bytes(W, S), c->stream>>>(A);
}
template <int W, int S, int MINB, bool STATIC, int PF, int DBG = 0>
static void launch_scan_bulk_dense3(acvd_ctx* c, const ReassignArgs& A) {
    static bool configured[64] = {};
    if (c->device >= 64 || !configured[c->device]) {
        ACVD_CUDA(cudaFuncSetAttribute(k_scan_bulk_dense3<W, S, MINB, false, STATIC, PF, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, dense_smem_bytes(W, S)));
        ACVD_CUDA(cudaFuncSetAttribute(k_scan_bulk_dense3<W, S, MINB, true, STATIC, PF, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, dense_smem_bytes(W, S)));
        if (c->device < 64) configured[c->device] = true;
    }
    const int n_tiles = A.tile_end - A.tile_begin;
    const int grid = std::max(1, std::min(kNumSMs * MINB, (n_tiles + kDenseWarps - 1) / kDenseWarps));
    if (A.bulk_stage == 1) k_scan_bulk_dense3<W, S, MINB, true, STATIC, PF, DBG><<<grid, kDenseThreads, dense_smem_bytes(W, S), c->stream>>>(A);
    else k_scan_bulk_dense3<W, S, MINB, false, STATIC, PF, DBG><<<grid, kDenseThreads, dense_smem_bytes(W, S), c->stream>>>(A);
}
// split dense scan: k_scan_classify (frontier scan -> candidate list) + k_bulk_decide (decision over the list)
template <int W, int S, int MINB, int VPL>
static void launch_scan_split(acvd_ctx* c, const ReassignArgs& A, int decide_bps) {
    static bool configured[64] = {};
    if (c->device >= 64 || !configured[c->device]) {
        ACVD_CUDA(cudaFuncSetAttribute(k_scan_classify<W, S, MINB, VPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, classify_smem_bytes(W, S, VPL)));
        if (c->device < 64) configured[c->device] = true;
    }
    const int n_tiles = A.tile_end - A.tile_begin;
    const int grid = std::max(1, std::min(kNumSMs * MINB, (n_tiles + kDenseWarps * VPL - 1) / (kDenseWarps * VPL)));
    k_scan_classify<W, S, MINB, VPL><<<grid, kDenseThreads, classify_smem_bytes(W, S, VPL), c->stream>>>(A);
    ACVD_LAUNCH_CHECK();
    const int chunk = classify_chunk(n_tiles, grid, VPL);
    const int split = std::max(1, (kNumSMs * decide_bps + grid - 1) / grid);       // blocks per segment: decide_bps resident blocks per SM
    const int gd = grid * split;
    if (A.bulk_stage == 1) k_bulk_decide<true><<<gd, 256, 0, c->stream>>>(A, grid, chunk, split);
    else k_bulk_decide<false><<<gd, 256, 0, c->stream>>>(A, grid, chunk, split);
    c->launches += 1;
}
// opening round of an exact phase: the streaming scan builds k_evaluate's work list (every boundary vertex)
template <int W>
static void launch_scan_opening(acvd_ctx* c, const ReassignArgs& A) {
    constexpr int S = 2, MINB = 4, VPL = 2;
    static bool configured[64] = {};
    if (c->device >= 64 || !configured[c->device]) {
        ACVD_CUDA(cudaFuncSetAttribute(k_scan_classify<W, S, MINB, VPL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, classify_smem_bytes(W, S, VPL)));
        if (c->device < 64) configured[c->device] = true;
    }
    const int n_tiles = A.tile_end - A.tile_begin;
    const int grid = std::max(1, std::min(kNumSMs * MINB, (n_tiles + kDenseWarps * VPL - 1) / (kDenseWarps * VPL)));
    k_scan_classify<W, S, MINB, VPL, true><<<grid, kDenseThreads, classify_smem_bytes(W, S, VPL), c->stream>>>(A);
}
// variants kept for the kernel micro-benchmark (acvd_bench_kernel) and the A/B tests: 0 / 1 the fused first generation
// (3 / 2 stages), 20 / 22 / 25 the third (min/max candidates; + prefetch; + static assignment), 30 / 31 diagnostic (no
// decision / no gathers either), 40-46 the split form (classify + decide; 1, 2, 2, 4 tiles per ticket)
constexpr int kDenseDefaultVariant = 42;   // split form, 2 tiles per ticket, 2 stages, 4 blocks per SM (C4: 558 us per launch over a run vs 746 us for the fused kernel)
static int dense_variant() { const char* e = getenv("ACVD_DENSE_VARIANT"); return e ? atoi(e) : kDenseDefaultVariant; }   // read per launch (tests switch it)
template <int W>
static void launch_scan_bulk_dense_variant(acvd_ctx* c, const ReassignArgs& A, int variant) {
    switch (variant) {
        case 1: launch_scan_bulk_dense<W, 2, 4>(c, A); break;
        case 20: launch_scan_bulk_dense3<W, 3, 4, false, 0>(c, A); break;
        case 22: launch_scan_bulk_dense3<W, 3, 4, false, 2>(c, A); break;
        case 25: launch_scan_bulk_dense3<W, 3, 4, true, 2>(c, A); break;
        case 30: launch_scan_bulk_dense3<W, 3, 4, false, 0, 1>(c, A); break;
        case 31: launch_scan_bulk_dense3<W, 3, 4, false, 0, 2>(c, A); break;
        case 40: launch_scan_split<W, 3, 6, 1>(c, A, 8); break;
        case 42: launch_scan_split<W, 2, 4, 2>(c, A, 8); break;
        case 44: launch_scan_split<W, 2, 5, 2>(c, A, 8); break;
        case 46: launch_scan_split<W, 2, 4, 4>(c, A, 8); break;
        default: launch_scan_bulk_dense<W, 3, 4>(c, A); break;
    }
}

static void launch_scan(acvd_ctx* c, const ReassignArgs& A, int grid) {
    c->last_dense_kernel = false;
    if (A.bulk && A.all_tiles && A.sig_mode == 1 && !getenv("ACVD_NO_DENSE_SCAN")) {
        c->last_dense_kernel = true;
        if (c->ell_w == 6) launch_scan_bulk_dense_variant<6>(c, A, dense_variant()); else launch_scan_bulk_dense_variant<8>(c, A, dense_variant());
        return;
    }
    if (A.bulk) {
        if (c->ell_w == 6) k_scan<6, true><<<grid, kThreads, 0, c->stream>>>(A); else k_scan<8, true><<<grid, kThreads, 0, c->stream>>>(A);
    } else if (A.all_tiles && A.sig_mode == 1 && A.force_all && !getenv("ACVD_NO_DENSE_SCAN")) {
        if (c->ell_w == 6) launch_scan_opening<6>(c, A); else launch_scan_opening<8>(c, A);
    } else {
        if (c->ell_w == 6) k_scan<6, false><<<grid, kThreads, 0, c->stream>>>(A); else k_scan<8, false><<<grid, kThreads, 0, c->stream>>>(A);
    }
}

// Scan mode of a round.  Dense rounds (the previous round found most tiles active, or a phase starts) scan the
// whole tile range and leave the signatures alone; the first sparse round after them rebuilds every signature
// while scanning everything; after that the tile filter is used.  Returns true when k_tile_filter must run.
static bool plan_scan(acvd_ctx* c, ReassignArgs& A, int force_all, int t0, int t1) {
    A.tile_begin = t0; A.tile_end = t1;
    const bool dense = force_all || c->dense_next;
    c->last_tile_count = t1 - t0;
    c->last_bulk = A.bulk;
    if (dense) { A.all_tiles = 1; A.sig_mode = 1; A.track_stale = 0; c->sig_valid = false; c->last_all_tiles = 1; return false; }
    if (!c->sig_valid) { A.all_tiles = 1; A.sig_mode = 2; A.track_stale = 1; c->sig_valid = true; c->last_all_tiles = 1; return false; }
    A.all_tiles = 0; A.sig_mode = 0; A.track_stale = 1; c->last_all_tiles = 0;
    return true;
}
// after the round: decide the next round's mode from how much of the mesh was active (counters summed over ranks,
// so every rank takes the same decision)
static void update_density(acvd_ctx* c, RoundResult& r) {
    // break-even points measured on C4: a TMA-staged dense bulk scan (0.71 ms) costs what the list-based scan costs on
    // 45 % of the tiles, and leaving dense mode costs one signature-rebuilding pass (1.6 ms), so bulk rounds stay dense
    // until little is left; the exact rounds use the list-based kernel in both modes, where only the tile filter is saved
    const double to_sparse = c->last_bulk ? 0.12 : 0.5, to_dense = c->last_bulk ? 0.45 : 0.6;
    if (c->last_all_tiles) {
        r.active_tiles = (unsigned long long)(((int64_t)c->V + 31) / 32);
        c->dense_next = r.boundary > 0 && (double)r.evaluated > to_sparse * (double)r.boundary;
    } else c->dense_next = (double)r.active_tiles > to_dense * (double)(((int64_t)c->V + 31) / 32);
}

// candidate evaluation of the work list: ring-in-registers kernel, plus the per-slot kernel when the mesh has rows
// longer than kRingW
static void launch_evaluate(acvd_ctx* c, const ReassignArgs& A, bool as_iso, int ge) {
#define ACVD_EVAL(KERNEL)                                                                          \
    switch (c->metric) {                                                                           \
        case M_ISO: KERNEL<M_ISO, 4><<<ge, kThreads, 0, c->stream>>>(A); break;                    \
        case M_QEM:                                                                                \
            if (as_iso) KERNEL<M_ISO, 14><<<ge, kThreads, 0, c->stream>>>(A);                      \
            else KERNEL<M_QEM, 14><<<ge, kThreads, 0, c->stream>>>(A);                             \
            break;                                                                                 \
        case M_ANISO: KERNEL<M_ANISO, 14><<<ge, kThreads, 0, c->stream>>>(A); break;               \
        default: KERNEL<M_ANISOQ, 22><<<ge, kThreads, 0, c->stream>>>(A); break;                   \
    }
    ACVD_EVAL(k_evaluate)
    ACVD_LAUNCH_CHECK();
    if (c->max_deg > kRingW) {
        ACVD_EVAL(k_evaluate_long)
        ACVD_LAUNCH_CHECK();
    }
#undef ACVD_EVAL
}

static void launch_round(acvd_ctx* c, const EvalCfg& cfg, int connexity, int force_all, bool as_iso, int slot = 0) {
    cudaEvent_t* ev = c->ev + 4 * slot;
    // proposals of the previous round become the carry list (none survive a phase start: everything is dirty); the
    // counters of the round are opened by k_modbits
    c->plist_cur ^= 1;
    ReassignArgs A = make_args(c, cfg, connexity, force_all);
    // the clusters this round modifies, for a sparse launch that may follow (sparse.cuh)
    c->mod_par ^= 1;
    A.modlist = c->mod_par ? c->modlist1.p : c->modlist0.p;
    A.n_mod = c->sp_nmod.p;
    ACVD_CUDA(cudaMemsetAsync(c->sp_nmod.p, 0, sizeof(unsigned long long), c->stream));
    c->modlist_valid = true;
    ACVD_CUDA(cudaMemsetAsync(c->best.p, 0xff, (size_t)c->K * sizeof(unsigned long long), c->stream));
    const int gs = grid_for((int64_t)c->V, kThreads, ACVD_SCAN_BPS), ge = kNumSMs * 8, gc = kNumSMs * 4;
    k_modbits<<<grid_for(c->K), kThreads, 0, c->stream>>>(c->K, c->mod_round.p, c->round - 1, force_all, c->modbits.p, c->ctr.p,
                                                          c->round_scalars.p, force_all ? 2 : 1);
    ACVD_LAUNCH_CHECK();
    const int n_tiles = (c->V + 31) / 32;
    const bool filtered = plan_scan(c, A, force_all, 0, n_tiles);
    ACVD_CUDA(cudaEventRecord(ev[0], c->stream));
    if (filtered) {
        k_tile_filter<<<grid_for(n_tiles), kThreads, 0, c->stream>>>(0, n_tiles, c->K, 0, reinterpret_cast<const int4*>(c->tile_sig.p),
                                                                    c->modbits.p, c->tile_active.p, c->active_tiles.p, c->round_scalars.p);
        ACVD_LAUNCH_CHECK();
    }
    launch_scan(c, A, gs);
    ACVD_LAUNCH_CHECK();
    if (filtered) {   // live proposals in tiles that were not scanned compete again
        k_carry<<<gc, kThreads, 0, c->stream>>>(A);
        ACVD_LAUNCH_CHECK();
    }
    ACVD_CUDA(cudaEventRecord(ev[3], c->stream));
    launch_evaluate(c, A, as_iso, ge);
    ACVD_CUDA(cudaEventRecord(ev[1], c->stream));
    for (int pass = 0; pass < c->commit_passes; pass++) {
        if (pass > 0) {
            ACVD_CUDA(cudaMemsetAsync(c->best.p, 0xff, (size_t)c->K * sizeof(unsigned long long), c->stream));
            k_resubmit<<<gc, kThreads, 0, c->stream>>>(A);
            ACVD_LAUNCH_CHECK();
        }
        switch (c->metric) {
            case M_ISO: k_commit<M_ISO, M_ISO><<<gc, kThreads, 0, c->stream>>>(A); break;
            case M_QEM:
                if (as_iso) k_commit<M_ISO, M_QEM><<<gc, kThreads, 0, c->stream>>>(A);
                else k_commit<M_QEM, M_QEM><<<gc, kThreads, 0, c->stream>>>(A);
                break;
            case M_ANISO: k_commit<M_ANISO, M_ANISO><<<gc, kThreads, 0, c->stream>>>(A); break;
            default: k_commit<M_ANISOQ, M_ANISOQ><<<gc, kThreads, 0, c->stream>>>(A); break;
        }
        ACVD_LAUNCH_CHECK();
    }
    ACVD_CUDA(cudaEventRecord(ev[2], c->stream));
    ACVD_CUDA(cudaMemcpyAsync(c->h_ctr + slot, c->ctr.p, sizeof(RoundCounters), cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaMemcpyAsync(c->h_scalars + 8 + slot, c->round_scalars.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    c->round++;
}

// fixed-point scale of the bulk rounds: a power of two such that no cluster sum can overflow 2^62
static void ensure_fx_scale(acvd_ctx* c) {
    if (c->fx_scale > 0) return;
    const int V = c->V, stride = payload_npad(c->metric);
    DevBuf<double> bound, total;
    bound.alloc(V); total.alloc(1);
    k_item_bound<<<grid_for(V), kThreads, 0, c->stream>>>(V, stride, c->items.p, bound.p);
    ACVD_LAUNCH_CHECK();
    size_t tb = 0;
    ACVD_CUDA(cub::DeviceReduce::Sum(nullptr, tb, bound.p, total.p, V, c->stream));
    void* t = cub_temp(c, tb);
    ACVD_CUDA(cub::DeviceReduce::Sum(t, tb, bound.p, total.p, V, c->stream));
    double T = 0;
    ACVD_CUDA(cudaMemcpyAsync(&T, total.p, sizeof T, cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    if (!(T > 0) || !std::isfinite(T)) { c->fx_scale = -1; return; }   // degenerate weights: bulk rounds off
    int e = 0;
    std::frexp(T, &e);                      // T = m * 2^e, m in [0.5, 1)
    c->fx_scale = std::ldexp(1.0, 61 - e);  // T * scale < 2^61
}

static BulkArgs make_bulk_args(acvd_ctx* c) {
    BulkArgs B;
    B.isum = c->isum.p; B.ccen = c->bulk_cen.p; B.cen_energy = c->bulk_energy.p; B.leave_cnt = c->leave_cnt.p; B.join_cnt = c->join_cnt.p;
    B.scale = c->fx_scale;
    return B;
}

static void bulk_init(acvd_ctx* c) {
    c->blist.alloc((size_t)c->vpad); c->blist_cnt.alloc(4096);          // candidate list of the split dense scan (one record per vertex at most)
    k_bulk_init<<<grid_for(c->K), kThreads, 0, c->stream>>>(c->K, payload_npad(c->metric), c->csum.p, make_bulk_args(c));
    ACVD_LAUNCH_CHECK();
}

// the same sum, enqueued behind a stage-1 round: its value arrives with the round's counters (one host synchronisation per
// round instead of two); bulk_energy_fetch() after the synchronisation
static void bulk_energy_enqueue(acvd_ctx* c) {
    size_t tb = 0;
    ACVD_CUDA(cub::DeviceReduce::Sum(nullptr, tb, c->bulk_energy.p, c->bulk_energy_sum.p, c->K, c->stream));
    void* t = cub_temp(c, tb);
    ACVD_CUDA(cub::DeviceReduce::Sum(t, tb, c->bulk_energy.p, c->bulk_energy_sum.p, c->K, c->stream));
    ACVD_CUDA(cudaMemcpyAsync(c->h_scalars + 7, c->bulk_energy_sum.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    c->bulk_energy_pending = true;
}
static double bulk_energy(acvd_ctx* c);
static double bulk_energy_fetch(acvd_ctx* c) {       // after the round's synchronisation
    if (!c->bulk_energy_pending) return bulk_energy(c);
    c->bulk_energy_pending = false;
    double e;
    memcpy(&e, c->h_scalars + 7, sizeof e);
    return e;
}
// deterministic sum of the per-cluster centroid energies kept by the bulk rounds
static double bulk_energy(acvd_ctx* c) {
    size_t tb = 0;
    ACVD_CUDA(cub::DeviceReduce::Sum(nullptr, tb, c->bulk_energy.p, c->bulk_energy_sum.p, c->K, c->stream));
    void* t = cub_temp(c, tb);
    ACVD_CUDA(cub::DeviceReduce::Sum(t, tb, c->bulk_energy.p, c->bulk_energy_sum.p, c->K, c->stream));
    double e = 0;
    ACVD_CUDA(cudaMemcpyAsync(&e, c->bulk_energy_sum.p, sizeof e, cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    return e;
}

// one bulk (Lloyd-criterion) round: scan -> evaluate against the centroids -> commit all -> refresh centroids
static void launch_bulk_round(acvd_ctx* c, int force_all, int stage) {
    EvalCfg cfg = make_cfg(0, 0, 0);
    c->plist_cur = 0;
    c->members_valid = false; c->modlist_valid = false;     // bulk commits do not maintain the member arrays
    ReassignArgs A = make_args(c, cfg, 0, force_all);
    A.bulk = 1; A.bulk_stage = stage; A.bulk_count_leave = 1;
    if (stage == 1) A.moved_mask = c->moved_mask.p;      // stage 1 can be undone (energy guard)
    BulkArgs B = make_bulk_args(c);
    k_modbits<<<grid_for(c->K), kThreads, 0, c->stream>>>(c->K, c->mod_round.p, c->round - 1, force_all, c->modbits.p, c->ctr.p,
                                                          c->round_scalars.p, 2, c->csize.p, c->cmeta.p);
    ACVD_LAUNCH_CHECK();
    const int gs = grid_for((int64_t)c->V, kThreads, ACVD_SCAN_BPS), ge = kNumSMs * 2;
    const int n_tiles = (c->V + 31) / 32;
    const bool filtered = plan_scan(c, A, force_all, 0, n_tiles);
    ACVD_CUDA(cudaEventRecord(c->ev[0], c->stream));
    if (filtered) {
        k_tile_filter<<<grid_for(n_tiles), kThreads, 0, c->stream>>>(0, n_tiles, c->K, 0, reinterpret_cast<const int4*>(c->tile_sig.p),
                                                                    c->modbits.p, c->tile_active.p, c->active_tiles.p, c->round_scalars.p);
        ACVD_LAUNCH_CHECK();
    }
    launch_scan(c, A, gs);
    ACVD_LAUNCH_CHECK();
    ACVD_CUDA(cudaEventRecord(c->ev[3], c->stream));
    k_bulk_evaluate<<<ge, kThreads, 0, c->stream>>>(A, B, 1, stage, payload_npad(c->metric));
    ACVD_LAUNCH_CHECK();
    ACVD_CUDA(cudaEventRecord(c->ev[1], c->stream));
    k_bulk_commit<<<grid_for((int64_t)n_tiles), kThreads, 0, c->stream>>>(A, B, payload_npad(c->metric));
    ACVD_LAUNCH_CHECK();
    k_bulk_refresh<<<grid_for(c->K), kThreads, 0, c->stream>>>(c->K, c->csize.p, B);
    ACVD_LAUNCH_CHECK();
    ACVD_CUDA(cudaEventRecord(c->ev[2], c->stream));
    if (stage == 1) bulk_energy_enqueue(c);          // the energy guard's sum travels with the counters
    ACVD_CUDA(cudaMemcpyAsync(c->h_ctr, c->ctr.p, sizeof(RoundCounters), cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaMemcpyAsync(c->h_scalars + 8, c->round_scalars.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    c->round++;
    c->stats_valid = false;
}

// the stage-1 energy guard tripped: the last bulk round is undone (cluster ids only; the statistics are recomputed
// from the clustering when the bulk rounds end)
static void dist_bulk_rollback(acvd_ctx* c);   // dist.cuh
static void bulk_rollback(acvd_ctx* c) {
    if (c->world > 1) {
        if (c->last_bulk_total > 0) dist_bulk_rollback(c);
    } else {
        EvalCfg cfg = make_cfg(0, 0, 0);
        ReassignArgs A = make_args(c, cfg, 0, 0);
        A.moved_mask = c->moved_mask.p;
        k_bulk_rollback<<<grid_for((int64_t)c->V), kThreads, 0, c->stream>>>(A);
        ACVD_LAUNCH_CHECK();
    }
    c->stats_valid = false; c->sig_valid = false;
}

static RoundResult finish_round(acvd_ctx* c, int slot = 0, bool last_of_batch = true) {
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    const RoundCounters* h = c->h_ctr + slot;
    cudaEvent_t* ev = c->ev + 4 * slot;
    RoundResult r;
    r.proposals = h->proposals; r.mods = h->mods; r.tests = h->tests;
    r.evaluated = h->evaluated + h->pad[0]; r.boundary = h->boundary; r.active_tiles = c->h_scalars[8 + slot];
    r.members = 0; r.overflow = h->pad[2] != 0; r.sparse = false;
    if (r.overflow) c->members_valid = false;
    ACVD_CUDA(cudaEventElapsedTime(&r.ms_scan, ev[0], ev[3]));
    ACVD_CUDA(cudaEventElapsedTime(&r.ms_eval, ev[3], ev[1]));
    ACVD_CUDA(cudaEventElapsedTime(&r.ms_commit, ev[1], ev[2]));
    if (last_of_batch) update_density(c, r);
    return r;
}

// ---- sparse rounds (sparse.cuh): one cooperative launch runs up to `max_rounds` rounds; returns the rounds executed
constexpr int kSparseChunk = 512;
static bool sparse_enabled() { static int v = getenv("ACVD_NO_SPARSE") ? 0 : 1; return v == 1; }

// Shape of a sparse launch: the whole device (cooperative grid) while a round has real work, ONE thread-block cluster once
// a round evaluates a few thousand vertices or fewer (its cost is then the barriers between its steps).
constexpr int kSparseClusterBlocks = 8;             // portable cluster size: 8 x 256 threads
constexpr int64_t kSparseClusterBelow = 3000;       // evaluated vertices per round at or below which the cluster form takes over
constexpr int64_t kSparseClusterAbove = 12000;      // ... and above which the grid form takes over again
static bool sparse_cluster_enabled() { return getenv("ACVD_NO_SPARSE_CLUSTER") == nullptr; }   // read per launch (tests switch it)

template <int EM, int STRIDE, int UM>
static void launch_sparse(acvd_ctx* c, ReassignArgs& A, SparseCtl& S, bool cluster) {
    if (cluster) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = dim3(kSparseClusterBlocks); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = 0; cfg.stream = c->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = kSparseClusterBlocks; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        ACVD_CUDA(cudaLaunchKernelEx(&cfg, k_sparse_rounds<EM, STRIDE, UM, true>, A, S));
        return;
    }
    static int blocks_per_sm[64] = {};
    int& bps = blocks_per_sm[c->device < 64 ? c->device : 0];
    if (bps == 0) {
        ACVD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_sparse_rounds<EM, STRIDE, UM, false>, kThreads, 0));
        if (bps < 1) throw std::runtime_error("k_sparse_rounds does not fit an SM");
        if (bps > 4) bps = 4;
    }
    int n_sm = kNumSMs;
    ACVD_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, c->device));
    void* args[] = {&A, &S};
    ACVD_CUDA(cudaLaunchCooperativeKernel((void*)k_sparse_rounds<EM, STRIDE, UM, false>, dim3(n_sm * bps), dim3(kThreads), args, 0, c->stream));
}

static int run_sparse_rounds(acvd_ctx* c, const EvalCfg& cfg, int connexity, bool as_iso, int max_rounds, long long stop_props,
                             int64_t n_prev_props, int64_t last_evaluated, std::vector<RoundResult>& out) {
    const int K = c->K;
    max_rounds = std::max(1, std::min(max_rounds, kSparseChunk));
    c->sp_rc.alloc(kSparseChunk); c->sp_resub.alloc((size_t)kSparseChunk * kMaxPasses); c->sp_ts.alloc((size_t)kSparseChunk * 4);
    c->sp_done.alloc(4);
    if (!c->h_sp_rc) {
        ACVD_CUDA(cudaMallocHost(&c->h_sp_rc, kSparseChunk * sizeof(RoundCounters)));
        ACVD_CUDA(cudaMallocHost(&c->h_sp_ts, (size_t)kSparseChunk * 4 * sizeof(unsigned long long) + 64));
    }
    ReassignArgs A = make_args(c, cfg, connexity, 0);
    A.track_stale = 0;                    // tile signatures are not maintained by the sparse rounds
    c->sig_valid = false;
    SparseCtl S;
    S.rc = c->sp_rc.p; S.n_mod = c->sp_nmod.p; S.resub = c->sp_resub.p; S.ts = c->sp_ts.p;
    S.modlist0 = c->mod_par ? c->modlist1.p : c->modlist0.p;        // written by the previous round
    S.modlist1 = c->mod_par ? c->modlist0.p : c->modlist1.p;
    S.plist1 = c->plist_cur ? c->plist_b.p : c->plist.p;            // proposals of the previous round
    S.plist0 = c->plist_cur ? c->plist.p : c->plist_b.p;
    S.best0 = c->best.p; S.best1 = c->best2.p;
    S.n_prev_props = (unsigned long long)std::max<int64_t>(0, n_prev_props);
    S.par0 = 0; S.max_rounds = max_rounds; S.passes = std::min(c->commit_passes, kMaxPasses);
    S.has_long_rows = c->max_deg > kRingW ? 1 : 0;
    S.stop_props = stop_props;
    S.done = c->sp_done.p;
    // launch shape from the work of the previous round; the launch ends when the other shape fits better
    const bool cluster = sparse_cluster_enabled() && S.passes <= 2 && max_rounds > 1 && last_evaluated <= kSparseClusterBelow;
    S.leave_below = (!cluster && sparse_cluster_enabled() && S.passes <= 2 && max_rounds > 1) ? kSparseClusterBelow : -1;
    S.leave_above = cluster ? kSparseClusterAbove : -1;
    if (cluster) {      // the cluster form resets only the key-table entries a round touched: it starts from clean tables
        ACVD_CUDA(cudaMemsetAsync(c->best.p, 0xff, (size_t)K * sizeof(unsigned long long), c->stream));
        ACVD_CUDA(cudaMemsetAsync(c->best2.p, 0xff, (size_t)K * sizeof(unsigned long long), c->stream));
    }
    ACVD_CUDA(cudaMemsetAsync(c->sp_rc.p, 0, (size_t)max_rounds * sizeof(RoundCounters), c->stream));
    ACVD_CUDA(cudaMemsetAsync(c->sp_nmod.p + 1, 0, (size_t)max_rounds * sizeof(unsigned long long), c->stream));
    ACVD_CUDA(cudaMemsetAsync(c->sp_resub.p, 0, (size_t)max_rounds * kMaxPasses * sizeof(unsigned long long), c->stream));
    ACVD_CUDA(cudaMemsetAsync(c->sp_done.p, 0, 4 * sizeof(int), c->stream));
    ACVD_CUDA(cudaEventRecord(c->ev[0], c->stream));
    switch (c->metric) {
        case M_ISO: launch_sparse<M_ISO, 4, M_ISO>(c, A, S, cluster); break;
        case M_QEM:
            if (as_iso) launch_sparse<M_ISO, 14, M_QEM>(c, A, S, cluster);
            else launch_sparse<M_QEM, 14, M_QEM>(c, A, S, cluster);
            break;
        case M_ANISO: launch_sparse<M_ANISO, 14, M_ANISO>(c, A, S, cluster); break;
        default: launch_sparse<M_ANISOQ, 22, M_ANISOQ>(c, A, S, cluster); break;
    }
    ACVD_LAUNCH_CHECK();
    ACVD_CUDA(cudaEventRecord(c->ev[1], c->stream));
    int n_done = 0;
    ACVD_CUDA(cudaMemcpyAsync(c->h_sp_rc, c->sp_rc.p, (size_t)max_rounds * sizeof(RoundCounters), cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaMemcpyAsync(c->h_sp_ts, c->sp_ts.p, (size_t)max_rounds * 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaMemcpyAsync(c->h_sp_ts + (size_t)kSparseChunk * 4, c->sp_done.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    n_done = *reinterpret_cast<int*>(c->h_sp_ts + (size_t)kSparseChunk * 4);
    if (n_done < 1 || n_done > max_rounds) throw std::runtime_error("sparse rounds: bad round count from the device");
    float ms_total = 0;
    ACVD_CUDA(cudaEventElapsedTime(&ms_total, c->ev[0], c->ev[1]));
    // the list of the clusters the last round modified becomes the input of the next launch
    if (n_done & 1) { c->mod_par ^= 1; c->plist_cur ^= 1; }
    ACVD_CUDA(cudaMemcpyAsync(c->sp_nmod.p, c->sp_nmod.p + n_done, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, c->stream));
    c->round += n_done;
    out.clear();
    double ns_sum = 0;
    for (int i = 0; i < n_done; i++) {
        const RoundCounters& h = c->h_sp_rc[i];
        const unsigned long long* ts = c->h_sp_ts + 4 * i;
        RoundResult r;
        r.proposals = h.proposals; r.mods = h.mods; r.tests = h.tests; r.evaluated = h.evaluated; r.boundary = h.evaluated;
        r.active_tiles = 0; r.members = h.pad[1]; r.overflow = h.pad[2] != 0; r.sparse = true;
        const unsigned long long t_begin = i > 0 ? c->h_sp_ts[4 * (i - 1) + 2] : ts[0];
        r.ms_scan = (float)((double)(ts[0] - t_begin) * 1e-6);
        r.ms_eval = (float)((double)(ts[1] - ts[0]) * 1e-6);
        r.ms_commit = (float)((double)(ts[2] - ts[1]) * 1e-6);
        ns_sum += (double)(ts[2] - t_begin);
        out.push_back(r);
    }
    if (getenv("ACVD_SPARSE_TRACE")) {
        fprintf(stderr, "[acvd sparse] launch of %d rounds: %.1f us by events, %.1f us between stamps\n", n_done, 1e3 * ms_total, ns_sum * 1e-3);
        for (int i = 0; i < n_done; i++)
            fprintf(stderr, "[acvd sparse]   round %4d members %8llu evaluated %8llu tests %8llu proposals %7llu mods %6llu  enum %.1f eval %.1f commit %.1f us\n",
                    c->round - n_done + i, out[i].members, out[i].evaluated, out[i].tests, out[i].proposals, out[i].mods,
                    1e3 * out[i].ms_scan, 1e3 * out[i].ms_eval, 1e3 * out[i].ms_commit);
    }
    // the first enumerate is not bracketed by device time stamps: give it what the events saw beyond the stamps
    if (!out.empty()) out[0].ms_scan += std::max(0.0f, ms_total - (float)(ns_sum * 1e-6));
    if (out.back().overflow) c->members_valid = false;
    c->last_all_tiles = 0; c->last_bulk = 0; c->last_dense_kernel = false; c->dense_next = false;
    c->launches += 1;
    c->last_sparse_cluster = cluster;
    return n_done;
}

#include "dist.cuh"

static double global_energy(acvd_ctx* c) {
    // ComputeGlobalEnergy sums in long double on the host (:1319-1346); K doubles is a small copy
    std::vector<double> e(c->K);
    ACVD_CUDA(cudaMemcpyAsync(e.data(), c->cenergy.p, (size_t)c->K * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    long double s = 0;
    for (double x : e) s += (long double)x;
    return (double)s;
}

// algorithmic bytes per launch (DESIGN.md "bytes model"):
//   scan     tile filter: 32 B signature + 1 B flag per tile (+4 per active tile listed); every vertex of an
//            active tile: row_ptr 4 + cid 4 + deg * (col 4 + neighbour cid 4) (deg = mesh mean), signature
//            write-back 32 B per active tile, + 4 per work-list entry written
//   evaluate per work-list vertex: list entry 4 + its CSR row again (8 + 8 deg) + item row + source cluster
//            row + (size, energy) 12; per test: destination row + energy 8; per proposal: write-back 28
//            (dst 4, key 8, energies 16) + 4 list entry
static int64_t scan_bytes(const acvd_ctx* c, const RoundResult& r) {
    const int64_t n_tiles = ((int64_t)c->V + 31) / 32;
    const double deg = c->V ? (double)c->nnz / c->V : 0.0;
    return 33 * n_tiles + (int64_t)r.active_tiles * (4 + 32 + (int64_t)(32.0 * (8.0 + 8.0 * deg))) + 4 * (int64_t)r.evaluated;
}
// sparse round: per visited member its list entry 4 + row_ptr 8 + deg * (col 4 + cid 4); per work-list entry stamp 4 + entry 4
static int64_t sparse_scan_bytes(const acvd_ctx* c, const RoundResult& r) {
    const double deg = c->V ? (double)c->nnz / c->V : 0.0;
    return (int64_t)((double)r.members * (12.0 + 8.0 * deg)) + 8 * (int64_t)r.evaluated;
}
static int64_t eval_bytes(const acvd_ctx* c, const RoundResult& r, bool as_iso) {
    const int64_t nl = 8 * (int64_t)(as_iso ? 4 : payload_npad(c->metric));
    const double deg = c->V ? (double)c->nnz / c->V : 0.0;
    return (int64_t)((double)r.evaluated * (12.0 + 8.0 * deg)) + (int64_t)r.evaluated * (2 * nl + 12) +
           (int64_t)r.tests * (nl + 8) + (int64_t)r.proposals * 32;
}

// bulk evaluate: list entry 4 + CSR row (8 + 8 deg) + point 12 + own centroid 24 + size 4 per work-list vertex;
// 24 (centroid) per test; 8 per proposal written + commit: item 32 + 2 x 32 fixed-point sums
// (bulk rounds take the decision inside k_scan: per decided vertex position 12 + own centroid 32 + size 4,
//  32 per test, 8 per proposal written; these bytes are added to the scan's)
static int64_t bulk_eval_bytes(const acvd_ctx* c, const RoundResult& r) {
    (void)c;
    return (int64_t)r.evaluated * 48 + (int64_t)r.tests * 32 + (int64_t)r.proposals * 8;
}

extern "C" int acvd_reassign_round(acvd_ctx* c, int constrained, int qlevel, int connexity, int64_t* proposals,
                                   int64_t* mods, int64_t* tests) {
    ACVD_API_BEGIN(c)
    if (!c->K) throw std::runtime_error("acvd_reassign_round: no clustering");
    bool as_iso = qem_as_iso(c, constrained, qlevel);
    if (!c->stats_valid || c->stats_constrained != constrained || c->stats_qlevel != qlevel)
        recompute_statistics(c, constrained, qlevel, 0);
    EvalCfg cfg = make_cfg(constrained, qlevel, 0);
    launch_round(c, cfg, connexity, 1, as_iso);
    RoundResult r = finish_round(c);
    if (proposals) *proposals = (int64_t)r.proposals;
    if (mods) *mods = (int64_t)r.mods;
    if (tests) *tests = (int64_t)r.tests;
    ACVD_API_END(c)
}

extern "C" int acvd_minimize(acvd_ctx* c, const acvd_params* pin, acvd_report* rep) {
    ACVD_API_BEGIN(c)
    if (!c->K || !c->have_items) throw std::runtime_error("acvd_minimize: need mesh, items and clustering");
    acvd_params p;
    memset(&p, 0, sizeof p);
    if (pin) p = *pin;
    if (p.max_loops <= 0) p.max_loops = 5000000;
    if (p.max_convergences <= 0) p.max_convergences = 1000000000;
    if (p.early_stop_div <= 0) p.early_stop_div = 1000;
    if (p.quadrics_level < 0) p.quadrics_level = 0;
    if (pin == nullptr) p.quadrics_level = 3;
    const double thr = p.sv_threshold > 0 ? p.sv_threshold : 1e-3;
    auto t0 = std::chrono::steady_clock::now();
    acvd_report R;
    memset(&R, 0, sizeof R);
    c->energy_log.clear(); c->energy_time.clear();

    // SetConstrainedClustering(0) when UnconstrainedInitialization (:683-688); only QEM honours it
    int constrained = (c->metric == M_QEM && p.unconstrained_init) ? 0 : 1;
    int connexity = p.connexity;
    const int qlevel = p.quadrics_level;
    EventPair clean_ev, total_ev;          // destroyed on every exit path
    cudaEvent_t ec0 = clean_ev.a, ec1 = clean_ev.b, et0 = total_ev.a, et1 = total_ev.b;
    ACVD_CUDA(cudaEventRecord(et0, c->stream));
    const int64_t launches0 = c->launches;
    auto timed_clean = [&](auto&& fn) {
        ACVD_CUDA(cudaEventRecord(ec0, c->stream));
        fn();
        ACVD_CUDA(cudaEventRecord(ec1, c->stream));
        ACVD_CUDA(cudaEventSynchronize(ec1));
        float ms = 0;
        ACVD_CUDA(cudaEventElapsedTime(&ms, ec0, ec1));
        R.ms_clean += ms;
    };
    // prime: FillHoles, ReComputeStatistics, SetAllClustersToModified (:727-730)
    timed_clean([&] { fill_holes(c, connexity); recompute_statistics(c, constrained, qlevel, thr); });
    int force_all = 1;
    int64_t last_proposals = -1;
    unsigned long long last_evaluated = 1ull << 62;   // dirty vertices of the previous round (picks the shape of a sparse launch)
    bool reeval_all = false;   // first replicated round of a multi-GPU tail
    int nconv = 0;
    // earlyStopItems = items of non-frozen clusters (:733-736)
    int64_t early_items = c->V;
    if (c->has_frozen) {
        std::vector<int> sz(c->K);
        std::vector<unsigned char> fr(c->K);
        ACVD_CUDA(cudaMemcpy(sz.data(), c->csize.p, (size_t)c->K * sizeof(int), cudaMemcpyDeviceToHost));
        ACVD_CUDA(cudaMemcpy(fr.data(), c->frozen.p, (size_t)c->K, cudaMemcpyDeviceToHost));
        early_items = 0;
        for (int i = 0; i < c->K; i++) if (!fr[i]) early_items += sz[i];
    }
    int64_t loops = 0;
    // Bulk rounds pay when a scan of the mesh costs more than the fixed cost of a round driven from the host (five launches and
    // a poll, ~0.1 ms): below kBulkMinVertices the persistent sparse-round kernel takes the phase from its opening round
    // (C1, 164 k vertices: 104 bulk rounds of ~95 us for ~15 us of work each).  bulk_rounds > 0 forces them on.
    const int bulk_cap = p.bulk_rounds < 0 ? 0 : (p.bulk_rounds == 0 ? (c->V >= kBulkMinVertices ? 1000 : 0) : p.bulk_rounds);
    const int env_passes = getenv("ACVD_COMMIT_PASSES") ? std::max(1, atoi(getenv("ACVD_COMMIT_PASSES"))) : 0;
    c->commit_passes = std::min(kMaxPasses, p.commit_passes > 0 ? p.commit_passes : (env_passes ? env_passes : 2));   // one default for every world size: N GPUs give the 1-GPU clustering
    while (true) {
        EvalCfg cfg = make_cfg(constrained, qlevel, thr);
        const bool as_iso = qem_as_iso(c, constrained, qlevel);
        // ---- bulk (Lloyd-criterion) rounds open the phases that end by early convergence
        if (force_all && bulk_cap > 0 && nconv <= 1 && !connexity && !c->has_frozen && !c->has_anchor &&
            (c->metric == M_ISO || as_iso)) {
            ensure_fx_scale(c);
            if (c->fx_scale > 0) {
                bulk_init(c);
                int fa = 1;
                // stage 0: Lloyd criterion (monotone by construction); stage 1: the reference's delta-E criterion on
                // the thin band of moves stage 0 leaves, guarded by the energy of the fixed-point sums
                int stage = 0;
                double e_prev = 0;
                for (int b = 0; b < bulk_cap && loops < p.max_loops; b++) {
                    RoundResult r;
                    if (c->world > 1) r = run_bulk_round_dist(c, fa, stage);
                    else { launch_bulk_round(c, fa, stage); r = finish_round(c); }
                    fa = 0;
                    loops++;
                    R.rounds++; R.bulk_rounds++; R.tests += (int64_t)r.tests; R.modifications += (int64_t)r.mods;
                    R.proposals += (int64_t)r.proposals; R.evaluated += (int64_t)r.evaluated;
                    R.ms_scan += r.ms_scan; R.ms_evaluate += r.ms_eval; R.ms_commit += r.ms_commit;
                    R.round_launches++; R.scan_bytes += scan_bytes(c, r) + bulk_eval_bytes(c, r);
                    if (c->last_dense_kernel) {   // this round's scan was the TMA-staged dense kernel: the dominant kernel, reported on its own
                        // counters are summed over the ranks, so are the vertices (every rank scans its share of the tiles)
                        const int64_t nt = ((int64_t)c->V + 31) / 32, nv = nt * 32;
                        R.dense_scan_launches++; R.ms_dense_scan += r.ms_scan; R.dense_scan_vertices += nv;
                        R.dense_scan_bytes += (int64_t)((double)nv * (8.0 + 8.0 * (double)c->nnz / c->V)) + 4 * nt + bulk_eval_bytes(c, r);
                    }
                    if (p.log_energy) {   // exact energy of the current clustering (test/trace path only)
                        recompute_statistics(c, constrained, qlevel, thr);
                        c->energy_log.push_back(global_energy(c));
                        c->energy_time.push_back(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
                    }
                    if (trace_on())
                        fprintf(stderr, "[acvd trace] bulk  %5lld conv %d tiles %8llu boundary %9llu evaluated %9llu tests %9llu proposals %9llu mods %8llu  scan %.0f eval %.0f commit %.0f us\n",
                                (long long)loops, nconv, r.active_tiles, r.boundary, r.evaluated, r.tests, r.proposals, r.mods,
                                1e3 * r.ms_scan, 1e3 * r.ms_eval, 1e3 * r.ms_commit);
                    const bool dry = (int64_t)r.proposals <= early_items / p.early_stop_div;
                    if (stage == 0) {
                        if (dry) { stage = 1; fa = 1; e_prev = bulk_energy(c); }
                    } else {
                        // stage 1 applies its moves simultaneously against round-start sums: no monotonicity guarantee,
                        // hence the guard -- a round that raised the energy is undone, and the exact rounds take over
                        const double e = bulk_energy_fetch(c);
                        if (e > e_prev) {
                            bulk_rollback(c);
                            R.modifications -= (int64_t)r.mods; R.bulk_rollbacks++;
                            if (p.log_energy && !c->energy_log.empty()) { c->energy_log.pop_back(); c->energy_time.pop_back(); }
                            break;
                        }
                        if (dry) break;
                        e_prev = e;
                    }
                }
                // proposals of the bulk rounds are not proposals of the exact rounds
                ACVD_CUDA(cudaMemsetAsync(c->prop_dst.p, 0xff, (size_t)c->V * sizeof(int), c->stream));
                timed_clean([&] { recompute_statistics(c, constrained, qlevel, thr); });
            }
        }
        // Across GPUs the work of a round is split and two exchanges keep the replicas in step.  Once a phase is in
        // its long tail (a few thousand live proposals or fewer) the exchanges cost more than the work: every rank
        // then runs the remaining rounds of the phase redundantly on its own replica -- same deterministic code on
        // the same state, so the replicas stay identical without any communication.  The first such round
        // re-evaluates every boundary vertex, because stored proposals are only known to the rank that owns them.
        RoundResult r;
        memset(&r, 0, sizeof r);
        const unsigned long long r_prev_evaluated = last_evaluated;
        int64_t synced_proposals = -1;             // multi-GPU: live proposals installed on every rank at the switch to the replicated tail
        if (force_all) c->replicated_tail = false;
        auto account = [&](const RoundResult& q) {
            loops++;
            R.rounds++; R.tests += (int64_t)q.tests; R.modifications += (int64_t)q.mods; R.proposals += (int64_t)q.proposals;
            R.ms_scan += q.ms_scan; R.ms_evaluate += q.ms_eval; R.ms_commit += q.ms_commit;
            R.round_launches++; R.evaluate_bytes += eval_bytes(c, q, as_iso); R.evaluated += (int64_t)q.evaluated;
            if (q.sparse) {
                R.sparse_rounds++; R.ms_sparse += q.ms_scan + q.ms_eval + q.ms_commit;
                if (c->last_sparse_cluster) R.sparse_cluster_rounds++;
                R.scan_bytes += sparse_scan_bytes(c, q);
            } else R.scan_bytes += scan_bytes(c, q);
            if (p.log_energy) {
                c->energy_log.push_back(global_energy(c));
                c->energy_time.push_back(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
            }
            if (trace_on())
                fprintf(stderr, "[acvd trace] %s %5lld conv %d tiles %8llu boundary %9llu evaluated %9llu tests %9llu proposals %9llu mods %8llu  scan %.0f eval %.0f commit %.0f us\n",
                        q.sparse ? "sprnd" : "round", (long long)loops, nconv, q.active_tiles, q.boundary, q.evaluated, q.tests, q.proposals, q.mods,
                        1e3 * q.ms_scan, 1e3 * q.ms_eval, 1e3 * q.ms_commit);
        };
        // Sparse rounds (sparse.cuh): after the opening round of a phase the rounds run inside one persistent cooperative
        // kernel that enumerates the dirty vertices from the member arrays of the modified clusters and decides
        // convergence on the device.  Across GPUs they are used once the phase is in its replicated tail.
        const bool sparse_ok = sparse_enabled() && p.sparse_rounds >= 0 && !force_all && !reeval_all && c->modlist_valid && last_proposals >= 0 &&
                               (c->world == 1 || c->replicated_tail);
        if (sparse_ok && !c->members_valid) members_build(c);       // an append overflowed a member array: rebuild (no sort needed)
        // In the long tail of the last phases (connexity on: the phase can only end with a round without moves) a few
        // rounds of the tile-filter path are launched back to back and their counters read together (ACVD_NO_SPARSE=1).
        int batch = 1;
        if (!sparse_ok && (c->world == 1 || c->replicated_tail) && nconv >= 2 && !force_all && !reeval_all && !p.log_energy && !trace_on() &&
            last_proposals >= 0 && !c->last_all_tiles)    // sparse rounds only: a dense round re-plans the scan mode after every round
            batch = (int)std::min<int64_t>(p.rounds_per_sync > 0 ? std::min(p.rounds_per_sync, kRoundSlots) : kTailBatch,
                                           std::max<int64_t>(1, p.max_loops - loops));
        if (sparse_ok) {
            const bool per_round = p.log_energy || trace_on();     // the host wants to look at every round
            const int64_t budget = std::max<int64_t>(1, p.max_loops - loops + 1);
            const long long stop_props = nconv <= 1 ? (long long)(early_items / p.early_stop_div) : -1;
            std::vector<RoundResult> rr;
            const int n = run_sparse_rounds(c, cfg, connexity, as_iso, per_round ? 1 : (int)std::min<int64_t>(budget, kSparseChunk),
                                            stop_props, last_proposals, (int64_t)r_prev_evaluated, rr);
            for (int j = 0; j + 1 < n; j++) account(rr[j]);
            r = rr[n - 1];
        } else if (c->world > 1 && !c->replicated_tail) {
            r = run_round_dist(c, cfg, connexity, force_all, as_iso);
            // Once a round evaluates few enough vertices the exchanges cost more than the work they split: the rest of the
            // phase runs replicated (sparse rounds, no communication).  Every rank first receives every live proposal.
            if ((int64_t)r.evaluated <= kReplicatedTailEvaluated && r.mods > 0) {
                synced_proposals = dist_sync_proposals(c);
                c->replicated_tail = true;
                c->sig_valid = false;            // tile signatures were kept for the rank's own range only
            }
        } else {
            for (int j = 0; j < batch; j++) launch_round(c, cfg, connexity, (j == 0 && (force_all || reeval_all)) ? 1 : 0, as_iso, j);
            if (batch > 1) c->modlist_valid = false;      // the list holds the last round of the batch only if all of them ran to the end
            for (int j = 0; j < batch; j++) {
                r = finish_round(c, j, j == batch - 1);
                if (j == batch - 1 || r.mods == 0) break;
                account(r);   // an intermediate round of the batch
            }
            reeval_all = false;
        }
        last_proposals = synced_proposals >= 0 ? synced_proposals : (int64_t)r.proposals;
        last_evaluated = r.evaluated;
        force_all = 0;
        account(r);
        const int64_t mods = (int64_t)r.mods;
        // convergence event (:773-776); a round commits a conflict-free subset, so the analogue of the
        // reference's "modifications in one sweep" is the number of live improving proposals
        const bool event = (mods == 0) || (loops > p.max_loops) ||
                           ((int64_t)r.proposals <= early_items / p.early_stop_div && nconv <= 1);
        if (!event) continue;
        if (c->metric == M_QEM && p.unconstrained_init && nconv == 0) constrained = 1;   // (:778-781)
        if (nconv >= 1) connexity = 1;                                                  // (:790)
        nconv++;
        R.convergences++;
        int disc = 0;
        // one pass over the clusters checks connectivity AND leaves fresh statistics (valid unless something was cleaned / filled)
        timed_clean([&] { disc = clean_clustering(c, true, constrained, qlevel, thr); fill_holes(c, connexity); });
        R.disconnected = disc;
        const bool done = (disc == 0 && mods == 0) || loops >= p.max_loops || nconv >= p.max_convergences;
        // the reference leaves the incremental sums in place on exit; we always export fresh statistics
        if (!c->stats_valid || c->stats_constrained != constrained || c->stats_qlevel != qlevel)
            timed_clean([&] { recompute_statistics(c, constrained, qlevel, thr); });
        if (done) break;
        force_all = 1;
    }
    R.energy = global_energy(c);
    ACVD_CUDA(cudaEventRecord(et1, c->stream));
    ACVD_CUDA(cudaEventSynchronize(et1));
    float ms_dev = 0;
    ACVD_CUDA(cudaEventElapsedTime(&ms_dev, et0, et1));
    R.ms_device = ms_dev;
    R.kernel_launches = c->launches - launches0;
    R.ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (rep) *rep = R;
    ACVD_API_END(c)
}

extern "C" int acvd_get_cluster_stats(acvd_ctx* c, double* sums, double* centroid, double* energy, int32_t* sizes) {
    ACVD_API_BEGIN(c)
    if (!c->K) throw std::runtime_error("acvd_get_cluster_stats: no clustering");
    if (!c->stats_valid) recompute_statistics(c, c->stats_constrained, c->stats_qlevel, 0);
    const int K = c->K, np = payload_np(c->metric), npad = payload_npad(c->metric);
    if (sums) {
        DevBuf<double> tmp;
        tmp.alloc((size_t)K * np);
        k_pack_rows<<<grid_for((int64_t)K * npad), kThreads, 0, c->stream>>>(K, np, npad, c->csum.p, tmp.p, 0);
        ACVD_LAUNCH_CHECK();
        ACVD_CUDA(cudaMemcpyAsync(sums, tmp.p, (size_t)K * np * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        ACVD_CUDA(cudaStreamSynchronize(c->stream));
    }
    if (centroid) ACVD_CUDA(cudaMemcpyAsync(centroid, c->ccentroid.p, 3 * (size_t)K * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (energy) ACVD_CUDA(cudaMemcpyAsync(energy, c->cenergy.p, (size_t)K * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (sizes) ACVD_CUDA(cudaMemcpyAsync(sizes, c->csize.p, (size_t)K * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    ACVD_API_END(c)
}

extern "C" int acvd_global_energy(acvd_ctx* c, double* energy) {
    ACVD_API_BEGIN(c)
    if (!c->K || !energy) throw std::runtime_error("acvd_global_energy: no clustering");
    if (!c->stats_valid) recompute_statistics(c, c->stats_constrained, c->stats_qlevel, 0);
    *energy = global_energy(c);
    ACVD_API_END(c)
}

extern "C" int acvd_get_energy_log(acvd_ctx* c, double* out, int32_t cap, int32_t* n) {
    if (!c) return fail(nullptr, ACVD_EINVAL, "null context");
    if (n) *n = (int32_t)c->energy_log.size();
    if (out) for (int i = 0; i < cap && i < (int)c->energy_log.size(); i++) out[i] = c->energy_log[i];
    return ACVD_OK;
}

extern "C" int acvd_get_energy_times(acvd_ctx* c, double* out, int32_t cap, int32_t* n) {
    if (!c) return fail(nullptr, ACVD_EINVAL, "null context");
    if (n) *n = (int32_t)c->energy_time.size();
    if (out) for (int i = 0; i < cap && i < (int)c->energy_time.size(); i++) out[i] = c->energy_time[i];
    return ACVD_OK;
}

extern "C" int acvd_representative_points(acvd_ctx* c, int32_t n, const double* Q9, double* P3, int32_t max_sv,
                                          double thr, int32_t* rank_def) {
    ACVD_API_BEGIN(c)
    if (n < 0 || (n && (!Q9 || !P3))) throw std::runtime_error("acvd_representative_points: bad arguments");
    if (n == 0) return ACVD_OK;
    DevBuf<double> q, p;
    DevBuf<int> rd;
    q.alloc(9 * (size_t)n); p.alloc(3 * (size_t)n); rd.alloc(n);
    ACVD_CUDA(cudaMemcpyAsync(q.p, Q9, 9 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    ACVD_CUDA(cudaMemcpyAsync(p.p, P3, 3 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    k_representative_points<<<grid_for(n), kThreads, 0, c->stream>>>(n, q.p, p.p, max_sv, thr > 0 ? thr : 1e-3, rd.p);
    ACVD_LAUNCH_CHECK();
    ACVD_CUDA(cudaMemcpyAsync(P3, p.p, 3 * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (rank_def) ACVD_CUDA(cudaMemcpyAsync(rank_def, rd.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    ACVD_API_END(c)
}

// ACVD's quadric post-process, accumulation part (ACVD.cxx:237-262)
extern "C" int acvd_cluster_quadrics(acvd_ctx* c, int32_t n_clusters, double* Q9) {
    ACVD_API_BEGIN(c)
    if (!c->K || !Q9 || n_clusters < 0 || n_clusters > c->K) throw std::runtime_error("acvd_cluster_quadrics: bad arguments");
    if (n_clusters == 0) return ACVD_OK;
    members_build(c);
    cluster_pass(c, false, false, 1, 3, 0);          // sort only: the summation order is the item order
    DevBuf<double> d_q;
    d_q.alloc(9 * (size_t)n_clusters);
    k_cluster_quadrics<<<grid_for((int64_t)n_clusters * 32), kThreads, 0, c->stream>>>(n_clusters, c->memb_off.p, c->memb.p, c->csize.p, c->vf_ptr.p,
                                                                                       c->vf_keys.p, c->xyz.p, c->tri.p, d_q.p);
    ACVD_LAUNCH_CHECK();
    ACVD_CUDA(cudaMemcpyAsync(Q9, d_q.p, 9 * (size_t)n_clusters * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    ACVD_API_END(c)
}

// the device connexity predicate on caller-given pairs (parity hook for ConnexityConstraintProblemLocal)
extern "C" int acvd_connexity_problem(acvd_ctx* c, int32_t n, const int32_t* items, const int32_t* clusters, int32_t mode, uint8_t* out) {
    ACVD_API_BEGIN(c)
    if (!c->K || n < 0 || (n && (!items || !clusters || !out))) throw std::runtime_error("acvd_connexity_problem: bad arguments");
    if (n == 0) return ACVD_OK;
    for (int i = 0; i < n; i++) if (items[i] < 0 || items[i] >= c->V) throw std::runtime_error("acvd_connexity_problem: item out of range");
    DevBuf<int> d_it, d_cl;
    DevBuf<unsigned char> d_out;
    d_it.alloc(n); d_cl.alloc(n); d_out.alloc(n);
    ACVD_CUDA(cudaMemcpyAsync(d_it.p, items, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    ACVD_CUDA(cudaMemcpyAsync(d_cl.p, clusters, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    k_connexity_query<<<grid_for(n), kThreads, 0, c->stream>>>(n, d_it.p, d_cl.p, c->row_ptr.p, c->col.p, c->cid.p, c->ringadj.p, mode, d_out.p);
    ACVD_LAUNCH_CHECK();
    ACVD_CUDA(cudaMemcpyAsync(out, d_out.p, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    ACVD_API_END(c)
}

// ---------------------------------------------------------------------------------------------
// kernel micro-benchmark: times `reps` back-to-back launches of one kernel on the current state (CUDA events on the
// library stream).  kernel 0: dense bulk scan, `variant` = (stages, blocks/SM) variant, -1 = the list-based k_scan;
// stage = bulk stage (0 Lloyd, 1 delta-E).  The state must come from a few bulk rounds (acvd_minimize with
// max_loops set): proposals are overwritten, nothing is committed.
extern "C" int acvd_bench_kernel(acvd_ctx* c, int kernel, int variant, int stage, int reps, float* ms_per_launch) {
    ACVD_API_BEGIN(c)
    if (!c->K || !c->have_items || kernel != 0 || reps <= 0 || !ms_per_launch) throw std::runtime_error("acvd_bench_kernel: bad arguments");
    if (c->metric != M_ISO && c->metric != M_QEM) throw std::runtime_error("acvd_bench_kernel: bulk scan needs the isotropic or QEM metric");
    ensure_fx_scale(c);
    if (!c->stats_valid) recompute_statistics(c, 0, 0, 0);
    bulk_init(c);
    EvalCfg cfg = make_cfg(0, 0, 0);
    ReassignArgs A = make_args(c, cfg, 0, 0);
    A.bulk = 1; A.bulk_stage = stage; A.bulk_count_leave = 1;
    A.all_tiles = 1; A.sig_mode = 1; A.tile_begin = 0; A.tile_end = (c->V + 31) / 32;
    k_modbits<<<grid_for(c->K), kThreads, 0, c->stream>>>(c->K, c->mod_round.p, c->round - 1, 0, c->modbits.p, nullptr, nullptr, 0,
                                                          c->csize.p, c->cmeta.p);
    ACVD_LAUNCH_CHECK();
    const int gs = grid_for((int64_t)c->V, kThreads, ACVD_SCAN_BPS);
    EventPair evp;
    cudaEvent_t e0 = evp.a, e1 = evp.b;
    for (int r = -2; r < reps; r++) {          // two warm-up launches
        if (r == 0) ACVD_CUDA(cudaEventRecord(e0, c->stream));
        ACVD_CUDA(cudaMemsetAsync(c->ctr.p, 0, sizeof(RoundCounters), c->stream));
        if (variant < 0) {
            if (c->ell_w == 6) k_scan<6, true><<<gs, kThreads, 0, c->stream>>>(A); else k_scan<8, true><<<gs, kThreads, 0, c->stream>>>(A);
        } else if (c->ell_w == 6) launch_scan_bulk_dense_variant<6>(c, A, variant);
        else launch_scan_bulk_dense_variant<8>(c, A, variant);
        ACVD_LAUNCH_CHECK();
    }
    ACVD_CUDA(cudaEventRecord(e1, c->stream));
    ACVD_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    ACVD_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *ms_per_launch = ms / reps;
    ACVD_CUDA(cudaMemsetAsync(c->prop_mask.p, 0, (size_t)((c->V + 31) / 32) * sizeof(unsigned), c->stream));
    ACVD_CUDA(cudaMemsetAsync(c->leave_cnt.p, 0, (size_t)c->K * sizeof(int), c->stream));
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    c->stats_valid = false;
    ACVD_API_END(c)
}

// ---------------------------------------------------------------------------------------------
// integer stages
extern "C" int acvd_boundary_flags(acvd_ctx* c, uint8_t* flags) {
    ACVD_API_BEGIN(c)
    if (!c->K || !flags) throw std::runtime_error("acvd_boundary_flags: no clustering");
    DevBuf<unsigned char> d;
    d.alloc(c->V);
    k_boundary_flags<<<grid_for(c->V), kThreads, 0, c->stream>>>(c->V, c->row_ptr.p, c->col.p, c->cid.p, d.p);
    ACVD_LAUNCH_CHECK();
    ACVD_CUDA(cudaMemcpyAsync(flags, d.p, (size_t)c->V, cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    ACVD_API_END(c)
}

// sorted unique (lo << 32 | hi) pairs of clusters joined by a mesh edge, left on the device; returns the count
static int64_t cluster_adjacency_device(acvd_ctx* c, DevBuf<unsigned long long>& uniq) {
    const int64_t n = c->nnz;
    DevBuf<unsigned long long> keys, alt;
    DevBuf<int64_t> d_num;
    keys.alloc(n); alt.alloc(n); uniq.alloc(n); d_num.alloc(1);
    k_adjacency_keys<<<grid_for(c->V), kThreads, 0, c->stream>>>(c->V, c->K, c->row_ptr.p, c->col.p, c->cid.p, keys.p);
    ACVD_LAUNCH_CHECK();
    sort_keys64(c, keys.p, alt.p, n, 64);
    size_t tb = 0;
    ACVD_CUDA(cub::DeviceSelect::Unique(nullptr, tb, keys.p, uniq.p, d_num.p, n, c->stream));
    void* t = cub_temp(c, tb);
    ACVD_CUDA(cub::DeviceSelect::Unique(t, tb, keys.p, uniq.p, d_num.p, n, c->stream));
    int64_t nu = 0;
    ACVD_CUDA(cudaMemcpyAsync(&nu, d_num.p, sizeof nu, cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    if (nu > 0) {
        unsigned long long last = 0;
        ACVD_CUDA(cudaMemcpy(&last, uniq.p + (nu - 1), sizeof last, cudaMemcpyDeviceToHost));
        if (last == ~0ull) nu--;
    }
    return nu;
}

extern "C" int acvd_cluster_adjacency(acvd_ctx* c, int64_t* out, int64_t cap, int64_t* n_out) {
    ACVD_API_BEGIN(c)
    if (!c->K || !n_out) throw std::runtime_error("acvd_cluster_adjacency: bad arguments");
    DevBuf<unsigned long long> uniq;
    const int64_t nu = cluster_adjacency_device(c, uniq);
    *n_out = nu;
    if (out && nu > 0) ACVD_CUDA(cudaMemcpy(out, uniq.p, (size_t)std::min(nu, cap) * sizeof(int64_t), cudaMemcpyDeviceToHost));
    ACVD_API_END(c)
}

// Dual triangles in first-occurrence order over the input faces, left on the device (d_out: 3 ints per triangle).
// A face qualifies when its three clusters are distinct and assigned; faces are ordered by their sorted cluster triple
// with two stable radix sorts -- (mid, hi) as one 64-bit key, then lo -- so equal triples keep ascending face order and
// no cluster-count limit applies; the first face of every run is the triple's first occurrence.
static int dual_triangles_device(acvd_ctx* c, DevBuf<int>& d_out) {
    const int F = c->F;
    DevBuf<unsigned long long> k0, k1;
    DevBuf<unsigned> lo, lo_perm, lo_sorted;
    DevBuf<int> f0, f1, f2, first, first_sorted;
    k0.alloc(F); k1.alloc(F); lo.alloc(F); lo_perm.alloc(F); lo_sorted.alloc(F);
    f0.alloc(F); f1.alloc(F); f2.alloc(F); first.alloc(F); first_sorted.alloc(F);
    k_dual_keys<<<grid_for(F), kThreads, 0, c->stream>>>(F, c->K, c->tri.p, c->cid.p, k0.p, lo.p, f0.p);
    ACVD_LAUNCH_CHECK();
    size_t tb = 0;
    ACVD_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, k0.p, k1.p, f0.p, f1.p, F, 0, 64, c->stream));
    void* t = cub_temp(c, tb);
    ACVD_CUDA(cub::DeviceRadixSort::SortPairs(t, tb, k0.p, k1.p, f0.p, f1.p, F, 0, 64, c->stream));
    k_gather_u32<<<grid_for(F), kThreads, 0, c->stream>>>(F, f1.p, lo.p, lo_perm.p);
    ACVD_LAUNCH_CHECK();
    ACVD_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, lo_perm.p, lo_sorted.p, f1.p, f2.p, F, 0, 32, c->stream));
    t = cub_temp(c, tb);
    ACVD_CUDA(cub::DeviceRadixSort::SortPairs(t, tb, lo_perm.p, lo_sorted.p, f1.p, f2.p, F, 0, 32, c->stream));
    ACVD_CUDA(cudaMemsetAsync(c->scalars.p, 0, sizeof(unsigned long long), c->stream));
    k_dual_first<<<grid_for(F), kThreads, 0, c->stream>>>(F, c->K, f2.p, c->tri.p, c->cid.p, first.p, c->scalars.p);
    ACVD_LAUNCH_CHECK();
    // ascending first-face ids = first-occurrence order over the input faces; sentinels sort to the end
    ACVD_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, first.p, first_sorted.p, F, 0, 32, c->stream));
    t = cub_temp(c, tb);
    ACVD_CUDA(cub::DeviceRadixSort::SortKeys(t, tb, first.p, first_sorted.p, F, 0, 32, c->stream));
    ACVD_CUDA(cudaMemcpyAsync(c->h_scalars, c->scalars.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    const int n = (int)c->h_scalars[0];
    d_out.alloc(3 * (size_t)std::max(n, 1));
    if (n > 0) {
        k_dual_emit<<<grid_for(n), kThreads, 0, c->stream>>>(n, first_sorted.p, c->tri.p, c->cid.p, d_out.p);
        ACVD_LAUNCH_CHECK();
    }
    return n;
}

extern "C" int acvd_dual_triangles(acvd_ctx* c, int32_t* out, int64_t cap, int64_t* n_out) {
    ACVD_API_BEGIN(c)
    if (!c->K || !n_out) throw std::runtime_error("acvd_dual_triangles: bad arguments");
    DevBuf<int> d_out;
    const int n = dual_triangles_device(c, d_out);
    *n_out = n;
    if (out && n > 0) {
        const int m = (int)std::min<int64_t>(n, cap);
        ACVD_CUDA(cudaMemcpyAsync(out, d_out.p, 3 * (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        ACVD_CUDA(cudaStreamSynchronize(c->stream));
    }
    ACVD_API_END(c)
}

// ---------------------------------------------------------------------------------------------
// manifoldness (the -m 1 loop): vtkSurfaceBase::IsVertexManifold on the device (manifold.cuh)
static void exclusive_sum(acvd_ctx* c, const int* in, int* out, int64_t n) {
    size_t tb = 0;
    ACVD_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, n, c->stream));
    void* t = cub_temp(c, tb);
    ACVD_CUDA(cub::DeviceScan::ExclusiveSum(t, tb, in, out, n, c->stream));
}

extern "C" int acvd_input_manifold_flags(acvd_ctx* c, uint8_t* flags) {
    ACVD_API_BEGIN(c)
    if (!c->V || !flags) throw std::runtime_error("acvd_input_manifold_flags: set the mesh first");
    DevBuf<unsigned char> d;
    d.alloc(c->V);
    FanMesh M{c->V, c->row_ptr.p, c->col.p, c->vf_ptr.p, nullptr, c->vf_keys.p, c->tri.p};
    k_vertex_manifold<<<grid_for(c->V, 128, 16), 128, 0, c->stream>>>(M, d.p);
    ACVD_LAUNCH_CHECK();
    ACVD_CUDA(cudaMemcpyAsync(flags, d.p, (size_t)c->V, cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    ACVD_API_END(c)
}

// the dual (output) mesh of the current clustering with its incidence, on the device
struct OutputMesh {
    DevBuf<int> tri, f_ptr, f_ids, e_ptr, e_nb;
    int n_tri = 0;
    int64_t n_pairs = 0;
};

static void build_output_mesh(acvd_ctx* c, int force_manifold_edges, OutputMesh& O) {
    const int K = c->K;
    DevBuf<int> f_cnt, f_cur, e_cnt, e_cur;
    const int n_tri = O.n_tri = dual_triangles_device(c, O.tri);
    f_cnt.alloc((size_t)K + 1); O.f_ptr.alloc((size_t)K + 1); f_cur.alloc(K); O.f_ids.alloc(3 * (size_t)std::max(n_tri, 1));
    ACVD_CUDA(cudaMemsetAsync(f_cnt.p, 0, ((size_t)K + 1) * sizeof(int), c->stream));
    ACVD_CUDA(cudaMemsetAsync(f_cur.p, 0, (size_t)K * sizeof(int), c->stream));
    if (n_tri > 0) { k_count_tri_corners<<<grid_for(3 * (int64_t)n_tri), kThreads, 0, c->stream>>>(n_tri, O.tri.p, f_cnt.p); ACVD_LAUNCH_CHECK(); }
    exclusive_sum(c, f_cnt.p, O.f_ptr.p, K + 1);
    if (n_tri > 0) { k_scatter_tri_corners<<<grid_for(3 * (int64_t)n_tri), kThreads, 0, c->stream>>>(n_tri, O.tri.p, O.f_ptr.p, f_cur.p, O.f_ids.p); ACVD_LAUNCH_CHECK(); }
    // edges: with ForceManifold every pair of clusters joined by a mesh edge (:1114-1133), else the triangles' edges
    DevBuf<unsigned long long> pairs;
    int64_t n_pairs = 0;
    if (force_manifold_edges) n_pairs = cluster_adjacency_device(c, pairs);
    else if (n_tri > 0) {
        const int64_t n3 = 3 * (int64_t)n_tri;
        DevBuf<unsigned long long> keys, alt;
        DevBuf<int64_t> d_num;
        keys.alloc(n3); alt.alloc(n3); pairs.alloc(n3); d_num.alloc(1);
        k_tri_edge_keys<<<grid_for(n3), kThreads, 0, c->stream>>>(n_tri, O.tri.p, keys.p);
        ACVD_LAUNCH_CHECK();
        sort_keys64(c, keys.p, alt.p, n3, 64);
        size_t tb = 0;
        ACVD_CUDA(cub::DeviceSelect::Unique(nullptr, tb, keys.p, pairs.p, d_num.p, n3, c->stream));
        void* t = cub_temp(c, tb);
        ACVD_CUDA(cub::DeviceSelect::Unique(t, tb, keys.p, pairs.p, d_num.p, n3, c->stream));
        ACVD_CUDA(cudaMemcpyAsync(&n_pairs, d_num.p, sizeof n_pairs, cudaMemcpyDeviceToHost, c->stream));
        ACVD_CUDA(cudaStreamSynchronize(c->stream));
    }
    O.n_pairs = n_pairs;
    e_cnt.alloc((size_t)K + 1); O.e_ptr.alloc((size_t)K + 1); e_cur.alloc(K); O.e_nb.alloc(2 * (size_t)std::max<int64_t>(n_pairs, 1));
    ACVD_CUDA(cudaMemsetAsync(e_cnt.p, 0, ((size_t)K + 1) * sizeof(int), c->stream));
    ACVD_CUDA(cudaMemsetAsync(e_cur.p, 0, (size_t)K * sizeof(int), c->stream));
    if (n_pairs > 0) { k_count_pair_ends<<<grid_for(n_pairs), kThreads, 0, c->stream>>>(n_pairs, pairs.p, e_cnt.p); ACVD_LAUNCH_CHECK(); }
    exclusive_sum(c, e_cnt.p, O.e_ptr.p, K + 1);
    if (n_pairs > 0) { k_scatter_pair_ends<<<grid_for(n_pairs), kThreads, 0, c->stream>>>(n_pairs, pairs.p, O.e_ptr.p, e_cur.p, O.e_nb.p); ACVD_LAUNCH_CHECK(); }
}

static void output_manifold_flags_device(acvd_ctx* c, const OutputMesh& O, DevBuf<unsigned char>& d) {
    d.alloc(c->K);
    FanMesh M{c->K, O.e_ptr.p, O.e_nb.p, O.f_ptr.p, O.f_ids.p, nullptr, O.tri.p};
    k_vertex_manifold<<<grid_for(c->K, 128, 16), 128, 0, c->stream>>>(M, d.p);
    ACVD_LAUNCH_CHECK();
}

extern "C" int acvd_output_manifold_flags(acvd_ctx* c, int32_t force_manifold_edges, uint8_t* flags) {
    ACVD_API_BEGIN(c)
    if (!c->K || !flags) throw std::runtime_error("acvd_output_manifold_flags: no clustering");
    OutputMesh O;
    build_output_mesh(c, force_manifold_edges, O);
    DevBuf<unsigned char> d;
    output_manifold_flags_device(c, O, d);
    ACVD_CUDA(cudaMemcpyAsync(flags, d.p, (size_t)c->K, cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    ACVD_API_END(c)
}

// the host form of the predicate for the vertices the kernel leaves open (more than kMaxRing edges)
static bool host_vertex_manifold(const std::vector<int>& nb, const std::vector<std::array<int, 3>>& faces, int v) {
    const int ne = (int)nb.size();
    if (ne < 2) return false;
    std::vector<int> cnt((size_t)ne, 0), par((size_t)ne);
    for (int j = 0; j < ne; j++) par[(size_t)j] = j;
    auto slot = [&](int u) { for (int j = 0; j < ne; j++) if (nb[(size_t)j] == u) return j; return -1; };
    std::function<int(int)> find = [&](int x) { while (par[(size_t)x] != x) { par[(size_t)x] = par[(size_t)par[(size_t)x]]; x = par[(size_t)x]; } return x; };
    for (const auto& t : faces) {
        int o[2], m = 0;
        for (int k = 0; k < 3; k++) if (t[(size_t)k] != v) { if (m < 2) o[m] = t[(size_t)k]; m++; }
        if (m != 2 || o[0] == o[1]) continue;
        const int ja = slot(o[0]), jb = slot(o[1]);
        if (ja < 0 || jb < 0) return false;
        cnt[(size_t)ja]++; cnt[(size_t)jb]++;
        const int ra = find(ja), rb = find(jb);
        if (ra != rb) par[(size_t)std::max(ra, rb)] = std::min(ra, rb);
    }
    const int r0 = find(0);
    for (int j = 0; j < ne; j++) if (cnt[(size_t)j] != 2 || find(j) != r0) return false;
    return true;
}

template <typename T>
static std::vector<T> download(acvd_ctx* c, const T* d, size_t n) {
    std::vector<T> h(n);
    if (n) ACVD_CUDA(cudaMemcpyAsync(h.data(), d, n * sizeof(T), cudaMemcpyDeviceToHost, c->stream));
    ACVD_CUDA(cudaStreamSynchronize(c->stream));
    return h;
}

// DetectNonManifoldOutputVertices (DiscreteRemeshing/vtkDiscreteRemeshing.h:166-383), one step of the -m 1 loop.
// The manifold tests run on the device; the bookkeeping the reference does on its per-cluster item lists (a handful of
// clusters) is done here on the host side of the library.  On return every cluster is frozen except the offending ones
// and their output neighbours, one new cluster per issue has been appended (seeded with the first item of the offending
// cluster or, for a one-item cluster, the first ring neighbour whose cluster has more than one item), the clustering
// and the cluster tables are those of the grown cluster count.
extern "C" int acvd_detect_non_manifold(acvd_ctx* c, int32_t force_manifold_edges, int32_t* n_issues, int32_t* new_num_clusters) {
    ACVD_API_BEGIN(c)
    c->cc_since = 0;      // the clustering changes outside the rounds: the next CleanClustering checks every cluster
    if (!c->K || !n_issues || !new_num_clusters) throw std::runtime_error("acvd_detect_non_manifold: bad arguments");
    const int V = c->V, K0 = c->K;
    OutputMesh O;
    build_output_mesh(c, force_manifold_edges, O);
    DevBuf<unsigned char> d_flags;
    output_manifold_flags_device(c, O, d_flags);
    std::vector<unsigned char> oflag = download(c, d_flags.p, (size_t)K0);
    std::vector<int> e_ptr = download(c, O.e_ptr.p, (size_t)K0 + 1), f_ptr = download(c, O.f_ptr.p, (size_t)K0 + 1);
    std::vector<int> e_nb, f_ids, otri;         // fetched only when something is flagged
    std::vector<int> cl;
    std::vector<unsigned char> frozen((size_t)K0, 1);
    std::vector<int> suspects;
    for (int k = 0; k < K0; k++) if (oflag[(size_t)k] != 1) suspects.push_back(k);
    std::vector<int> issues;
    std::vector<std::vector<int>> items;         // per suspect cluster (and later per touched cluster): its items, ascending
    std::vector<int> size_of;
    if (!suspects.empty()) {
        e_nb = download(c, O.e_nb.p, (size_t)e_ptr[(size_t)K0]);
        f_ids = download(c, O.f_ids.p, (size_t)f_ptr[(size_t)K0]);
        otri = download(c, O.tri.p, 3 * (size_t)O.n_tri);
        cl = download(c, c->cid.p, (size_t)V);
        size_of.assign((size_t)K0, 0);
        for (int v = 0; v < V; v++) if (cl[(size_t)v] >= 0 && cl[(size_t)v] < K0) size_of[(size_t)cl[(size_t)v]]++;
        // input-vertex manifoldness on the device, only now that it is needed
        DevBuf<unsigned char> d_in;
        d_in.alloc(V);
        FanMesh M{V, c->row_ptr.p, c->col.p, c->vf_ptr.p, nullptr, c->vf_keys.p, c->tri.p};
        k_vertex_manifold<<<grid_for(V, 128, 16), 128, 0, c->stream>>>(M, d_in.p);
        ACVD_LAUNCH_CHECK();
        std::vector<unsigned char> iflag = download(c, d_in.p, (size_t)V);
        std::vector<char> is_suspect((size_t)K0, 0);
        for (int k : suspects) is_suspect[(size_t)k] = 1;
        std::vector<std::vector<int>> sus_items((size_t)K0);
        for (int v = 0; v < V; v++) { const int k = cl[(size_t)v]; if (k >= 0 && k < K0 && is_suspect[(size_t)k]) sus_items[(size_t)k].push_back(v); }
        auto input_manifold = [&](int v) {
            if (iflag[(size_t)v] != 2) return iflag[(size_t)v] == 1;
            int rp[2];
            ACVD_CUDA(cudaMemcpy(rp, c->row_ptr.p + v, 2 * sizeof(int), cudaMemcpyDeviceToHost));
            std::vector<int> nb = download(c, c->col.p + rp[0], (size_t)(rp[1] - rp[0]));
            int fp[2];
            ACVD_CUDA(cudaMemcpy(fp, c->vf_ptr.p + v, 2 * sizeof(int), cudaMemcpyDeviceToHost));
            std::vector<unsigned long long> fk = download(c, c->vf_keys.p + fp[0], (size_t)(fp[1] - fp[0]));
            std::vector<std::array<int, 3>> faces;
            for (auto k : fk) { int t[3]; ACVD_CUDA(cudaMemcpy(t, c->tri.p + 3 * (size_t)(k & 0xffffffffull), 3 * sizeof(int), cudaMemcpyDeviceToHost)); faces.push_back({t[0], t[1], t[2]}); }
            return host_vertex_manifold(nb, faces, v);
        };
        for (int k : suspects) {
            bool manifold = oflag[(size_t)k] == 1;
            if (oflag[(size_t)k] == 2) {
                std::vector<int> nb(e_nb.begin() + e_ptr[(size_t)k], e_nb.begin() + e_ptr[(size_t)k + 1]);
                std::vector<std::array<int, 3>> faces;
                for (int i = f_ptr[(size_t)k]; i < f_ptr[(size_t)k + 1]; i++) { const int f = f_ids[(size_t)i]; faces.push_back({otri[3 * (size_t)f], otri[3 * (size_t)f + 1], otri[3 * (size_t)f + 2]}); }
                manifold = host_vertex_manifold(nb, faces, k);
            }
            if (manifold) continue;
            if (sus_items[(size_t)k].empty()) continue;                 // ".... but empty. Skipping"
            bool problem = true;
            for (int it : sus_items[(size_t)k]) if (!input_manifold(it)) { problem = false; break; }   // the input has the issue too
            if (!problem) continue;
            issues.push_back(k);
            frozen[(size_t)k] = 0;
            for (int i = e_ptr[(size_t)k]; i < e_ptr[(size_t)k + 1]; i++) frozen[(size_t)e_nb[(size_t)i]] = 0;
        }
        // ---- one new cluster per issue (:262-372)
        int K = K0;
        // ring of an item in the reference's order (edge creation order = ascending first half-edge slot)
        auto ring_of = [&](int v) {
            int rp[2], fp[2];
            ACVD_CUDA(cudaMemcpy(rp, c->row_ptr.p + v, 2 * sizeof(int), cudaMemcpyDeviceToHost));
            ACVD_CUDA(cudaMemcpy(fp, c->vf_ptr.p + v, 2 * sizeof(int), cudaMemcpyDeviceToHost));
            std::vector<int> nb = download(c, c->col.p + rp[0], (size_t)(rp[1] - rp[0]));
            std::vector<unsigned long long> fk = download(c, c->vf_keys.p + fp[0], (size_t)(fp[1] - fp[0]));
            std::vector<std::pair<unsigned, int>> order;
            std::vector<std::array<int, 4>> faces;
            for (auto k : fk) { const int f = (int)(k & 0xffffffffull); int t[3]; ACVD_CUDA(cudaMemcpy(t, c->tri.p + 3 * (size_t)f, 3 * sizeof(int), cudaMemcpyDeviceToHost)); faces.push_back({t[0], t[1], t[2], f}); }
            for (int u : nb) {
                unsigned best = 0xffffffffu;
                for (const auto& t : faces) {
                    const int ia = t[0] == v ? 0 : (t[1] == v ? 1 : 2);
                    const int ib = t[0] == u ? 0 : (t[1] == u ? 1 : (t[2] == u ? 2 : -1));
                    if (ib < 0 || ib == ia) continue;
                    const int lo = std::min(ia, ib), hi = std::max(ia, ib);
                    const int side = (lo == 0 && hi == 1) ? 0 : (lo == 1 ? 1 : 2);
                    best = std::min(best, 3u * (unsigned)t[3] + (unsigned)side);
                }
                order.push_back({best, u});
            }
            std::sort(order.begin(), order.end());
            std::vector<int> out;
            for (auto& p : order) out.push_back(p.second);
            return out;
        };
        // unassigned items carry the NULL id = the cluster count, which is about to grow: park them at -1 meanwhile
        // (the reference leaves them at the old count, where they would silently join the first appended cluster)
        if (!issues.empty()) for (int v = 0; v < V; v++) if (cl[(size_t)v] == K0) cl[(size_t)v] = -1;
        for (int k : issues) {
            const int fresh = K;
            auto& mine = sus_items[(size_t)k];
            bool found = false;
            if (mine.size() > 1) {
                const int it = mine.front();
                cl[(size_t)it] = fresh;
                mine.erase(mine.begin());
                size_of[(size_t)k]--;
                found = true;
            } else if (mine.size() == 1) {
                for (int u : ring_of(mine.front())) {
                    const int cu = cl[(size_t)u];
                    if (cu < 0 || cu >= K0) continue;                  // NULL, or a cluster created by this very pass (one item)
                    if (size_of[(size_t)cu] > 1) {
                        cl[(size_t)u] = fresh;
                        size_of[(size_t)cu]--;
                        if (is_suspect[(size_t)cu]) { auto& o = sus_items[(size_t)cu]; o.erase(std::find(o.begin(), o.end(), u)); }
                        found = true;
                        break;
                    }
                }
            }
            if (found) { K++; frozen.push_back(0); }
        }
        if (!issues.empty()) for (int v = 0; v < V; v++) if (cl[(size_t)v] < 0) cl[(size_t)v] = K;
        *new_num_clusters = K;
    } else *new_num_clusters = K0;
    *n_issues = (int32_t)issues.size();
    // ---- apply: frozen flags, and (if clusters were appended) the grown tables with the edited clustering
    const int K = *new_num_clusters;
    if (K != K0) {
        std::vector<int64_t> fixed = c->fixed;
        set_num_clusters_impl(c, K);
        if (!fixed.empty()) {
            std::vector<int> a((size_t)K, -1);
            c->fixed = fixed;
            for (size_t i = 0; i < fixed.size(); i++) a[i] = (int)fixed[i];
            ACVD_CUDA(cudaMemcpy(c->anchor.p, a.data(), (size_t)K * sizeof(int), cudaMemcpyHostToDevice));
            c->has_anchor = true;
        }
        ACVD_CUDA(cudaMemcpy(c->cid.p, cl.data(), (size_t)V * sizeof(int), cudaMemcpyHostToDevice));
    }
    ACVD_CUDA(cudaMemcpy(c->frozen.p, frozen.data(), (size_t)K, cudaMemcpyHostToDevice));
    c->has_frozen = true;
    c->stats_valid = false; c->sig_valid = false; c->members_valid = false; c->modlist_valid = false;
    ACVD_API_END(c)
}

// ---------------------------------------------------------------------------------------------
// multi-GPU plumbing and round drivers: dist.cuh (included above)
