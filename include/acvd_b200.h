/*
 * acvd_b200 — C ABI of the B200-native ACVD clustering hot path.
 *
 * The reference (valette/ACVD) has no FFI; its extension points for this path are the
 * virtual hooks of the clustering engine and the duck-typed Metric template parameter
 * (reference Common/vtkUniformClustering.h:139-142, 212, 261-267;
 *  DiscreteRemeshing/vtkSurfaceClustering.h:44-52 selects the engine at compile time).
 * Each entry point below names the reference member it replaces.  The host classes in
 * acvd_b200/csrc/host (vtkIsotropicDiscreteRemeshing & co.) call only these functions.
 *
 * Conventions: plain pointers and sizes, caller owns every host pointer, the library
 * copies in/out and retains nothing after a call returns.  Every function returns 0 on
 * success or a negative ACVD_E* code; acvd_last_error() gives the message.  No C++
 * exception crosses this boundary.  There is NO CPU fallback: without a CUDA device
 * acvd_create fails with ACVD_ENODEVICE.
 */
#ifndef ACVD_B200_H
#define ACVD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACVD_B200_ABI_VERSION 4

typedef struct acvd_ctx acvd_ctx;

enum {
    ACVD_OK = 0,
    ACVD_EINVAL = -1,     /* bad argument / call order */
    ACVD_ENODEVICE = -2,  /* no CUDA device: the product has no CPU path */
    ACVD_ECUDA = -3,      /* CUDA runtime error (message in acvd_last_error) */
    ACVD_ENCCL = -4,      /* NCCL error */
    ACVD_ENOMEM = -5
};

/* Metric kinds == the four Metric classes of the reference:
 * vtkIsotropicMetricForClustering, vtkQEMetricForClustering,
 * vtkAnisotropicMetricForClustering, vtkQuadricAnisotropicMetricForClustering. */
enum { ACVD_METRIC_ISO = 0, ACVD_METRIC_QEM = 1, ACVD_METRIC_ANISO = 2, ACVD_METRIC_ANISOQ = 3 };

/* Number of accumulated doubles per item/cluster: S[3] W | +Q[9] | +T[6] X[3] | +T X Q  (4/13/13/22). */
int acvd_payload_size(int metric);

/* ---- lifetime ------------------------------------------------------------------------- */
/* device < 0: use the current CUDA device. */
int acvd_create(acvd_ctx** out, int device);
int acvd_destroy(acvd_ctx* ctx);
const char* acvd_last_error(acvd_ctx* ctx); /* ctx may be NULL: last create error */
int acvd_abi_version(void);

/* ---- mesh: replaces vtkSurfaceClustering::SetInput (DiscreteRemeshing/vtkSurfaceClustering.h:179)
 * and the adapter accessors GetEdgeItems/GetItemNeighbours/GetItemEdges
 * (DiscreteRemeshing/vtkVerticesProcessing.h:137-157): uploads float32 points + int32 triangles
 * and builds the CSR vertex adjacency (int32 row_ptr[V+1], col[2E], rows sorted) on the device. */
int acvd_set_mesh(acvd_ctx* ctx, int32_t V, int32_t F, const float* xyz, const int32_t* tri);
int acvd_get_num_edges(acvd_ctx* ctx, int64_t* E);
int acvd_get_csr(acvd_ctx* ctx, int32_t* row_ptr /*V+1*/, int32_t* col /*2E*/);

/* vtkSurface::Subdivide (Common/vtkSurface.cxx:605-677), the pre-pass of vtkDiscreteRemeshing::CheckSubsamplingRatio
 * (DiscreteRemeshing/vtkDiscreteRemeshing.h:841-875, option -s): 1 -> 4 split of the context's mesh on the device.
 * Old points first, then one midpoint per edge in the reference's edge-id order (first seen over the faces),
 * faces (V1,V4,V6) (V4,V2,V5) (V5,V3,V6) (V4,V5,V6) per input face.  The result stays on the device until fetched;
 * parent1/parent2[v] are the edge endpoints of a midpoint (v itself for old points), as the reference keeps them for
 * interpolating the curvature indicator (:733-745).  Any pointer of acvd_get_subdivision may be NULL. */
int acvd_subdivide(acvd_ctx* ctx, int32_t* n_vertices, int32_t* n_faces);
/* vtkSurface::SplitLongEdges (Common/vtkSurface.cxx:444-604, option -l): threshold = ratio x mean edge length of the
 * context's mesh; passes of "cut every edge above the threshold at its midpoint, replace the triangles by the Split2 /
 * Split3 / 1 -> 4 patterns" until no edge is above it.  The result stays on the device like a subdivision (fetch it with
 * acvd_get_subdivision; parents are the end points of the cut edge).  New points follow the old ones pass by pass, in the
 * edge-id order of the pass; faces keep their order, children in the pattern's order. */
int acvd_split_long_edges(acvd_ctx* ctx, double ratio, int32_t* n_vertices, int32_t* n_faces, int32_t* n_passes /*may be NULL*/);
int acvd_get_subdivision(acvd_ctx* ctx, float* xyz /*3 nv*/, int32_t* tri /*3 nf*/, int32_t* parent1 /*nv*/, int32_t* parent2 /*nv*/);

/* vtkCurvatureMeasure with ComputationMethod 1 (polynomial fitting), ElementsType 1 (vertices) and the n-ring
 * neighbourhood (Common/vtkCurvatureMeasure.cxx:188-508, 625-718; defaults :1175-1196: ring_size 3), as called by
 * vtkDiscreteRemeshing::SamplingPreProcessing (DiscreteRemeshing/vtkDiscreteRemeshing.h:640-653).
 * indicator[V] = sqrt(k1^2 + k2^2) (the CustomWeights of acvd_build_items); info6[6 V] = (sqrt|k_a| d_a, sqrt|k_b| d_b),
 * larger |k| first, float32 (the PrincipalDirections of the anisotropic metrics), may be NULL. */
int acvd_curvature(acvd_ctx* ctx, int32_t ring_size, double* indicator /*V*/, float* info6 /*6V or NULL*/);

/* ---- items: replaces Metric::BuildMetric (vtkIsotropicMetricForClustering.h:214-271,
 * vtkQEMetricForClustering.h:297-345, vtk(Quadric)AnisotropicMetricForClustering.h BuildMetric):
 * vertex areas, weights = area * indicator^gradation clamped to [mean/R, mean*R], Value = w*p,
 * per-vertex quadrics / tensors, all computed on the device.
 * custom_weights: V doubles (curvature indicator) or NULL; principal_dirs: 6V floats or NULL. */
int acvd_build_items(acvd_ctx* ctx, int metric, double gradation, const double* custom_weights,
                     const float* principal_dirs);
/* Alternative: caller-provided payload, V x acvd_payload_size(metric) doubles, row-major. */
int acvd_set_items(acvd_ctx* ctx, int metric, const double* payload);
int acvd_get_items(acvd_ctx* ctx, double* payload);
int acvd_get_vertex_areas(acvd_ctx* ctx, double* areas /*V*/);

/* ---- clusters: SetNumberOfClusters (Common/vtkUniformClustering.h:62-69), Clustering array
 * (:201-202), IsClusterFreezed (:328-329), FixedClusters / Cluster::AnchorItem
 * (:331-332, vtkQEMetricForClustering.h:127-129).  Cluster id K is the NULL cluster. */
int acvd_set_num_clusters(acvd_ctx* ctx, int32_t K);
int acvd_set_clustering(acvd_ctx* ctx, const int32_t* clustering /*V*/);
int acvd_get_clustering(acvd_ctx* ctx, int32_t* clustering /*V*/);
/* device-resident copy of the current clustering / restore it (re-running from the same start
 * without a host round trip; SetInitialClustering analogue, vtkUniformClustering.h:337-342) */
int acvd_save_clustering(acvd_ctx* ctx);
int acvd_restore_clustering(acvd_ctx* ctx);
int acvd_set_frozen(acvd_ctx* ctx, const uint8_t* frozen /*K, NULL clears*/);
int acvd_get_frozen(acvd_ctx* ctx, uint8_t* frozen /*K*/);
int acvd_set_fixed_clusters(acvd_ctx* ctx, const int64_t* anchor_items, int32_t n);

/* ---- initial sampling: ComputeInitialRandomSampling (Common/vtkUniformClustering.h:1178-1316).
 * Sequential by construction (mt19937 seed 0 + weight-bounded BFS in the reference's ring order);
 * runs on the host inside this library from the uploaded mesh and the device-built weights,
 * then uploads the result as the current clustering.  Not part of the timed clustering
 * (the reference's own timer starts after it, vtkUniformClustering.h:690). */
int acvd_initial_sampling(acvd_ctx* ctx);

/* ---- the hot path ----------------------------------------------------------------------- */
typedef struct acvd_params {
    int32_t unconstrained_init;   /* UnconstrainedInitialization (vtkUniformClustering.h:322-323) */
    int32_t quadrics_level;       /* QuadricsOptimizationLevel, default 3 (vtkQEMetricForClustering.h:368) */
    int32_t connexity;            /* initial ConnexityConstraint (0; the -m loop resets it to 0) */
    int32_t max_loops;            /* MaxNumberOfLoops, <=0 -> 5000000 */
    int32_t max_convergences;     /* MaxNumberOfConvergences, <=0 -> 1000000000 */
    int32_t early_stop_div;       /* <=0 -> 1000 (vtkUniformClustering.h:775) */
    int32_t log_energy;           /* keep a per-round global-energy trace (energy.txt analogue) */
    int32_t rounds_per_sync;      /* rounds launched back to back between host polls in the tail of the last phases, <=0 -> 4, max 8 */
    double sv_threshold;          /* <=0 -> 1e-3 (Common/vtkQuadricTools.h:36) */
    int32_t bulk_rounds;          /* early phases: 0 -> automatic (bulk Lloyd-criterion rounds for meshes of >= 500 k vertices, cap 1000), <0 -> off, >0 -> on with this cap */
    int32_t commit_passes;        /* select+commit passes per exact round; <=0 -> 2 (the same default on any number of GPUs), max 8 */
    int32_t sparse_rounds;        /* 0 -> rounds after a phase's opening round run in the persistent sparse-round kernel
                                     (dirty set enumerated from the modified clusters' member arrays); <0 -> tile-filter
                                     path for every round (same decisions: the A/B partner of the tests) */
} acvd_params;

typedef struct acvd_report {
    int64_t rounds;          /* reassignment rounds (the parallel analogue of NumberOfLoops) */
    int64_t convergences;    /* convergence events */
    int64_t tests;           /* vertex tests: evaluated (item, target cluster) candidates incl. blocked */
    int64_t modifications;   /* committed moves */
    int64_t proposals;       /* improving candidates submitted to conflict resolution */
    int64_t disconnected;    /* clusters split by the last CleanClustering */
    double energy;           /* ComputeGlobalEnergy (vtkUniformClustering.h:1319-1346) */
    double ms_total;         /* wall time of the call */
    double ms_scan;          /* device time in the frontier-scan kernel (k_scan) */
    double ms_evaluate;      /* device time in the candidate evaluation kernel (k_evaluate) */
    double ms_commit;        /* device time in conflict resolution + commit */
    double ms_clean;         /* device time in stats / clean / fill */
    int64_t round_launches;  /* launches of each of k_scan / k_evaluate / k_commit */
    int64_t scan_bytes;      /* algorithmic bytes moved by k_scan (SURVEY §8d model) */
    int64_t evaluate_bytes;  /* algorithmic bytes moved by k_evaluate */
    int64_t evaluated;       /* work-list vertices evaluated */
    double ms_device;        /* CUDA-event time of the whole call on the context's stream */
    int64_t kernel_launches; /* kernels this call launched */
    int64_t bulk_rounds;     /* rounds run with the bulk (Lloyd-criterion) commit */
    int64_t dense_scan_launches; /* launches of the TMA-staged dense bulk scan (k_scan_bulk_dense), the dominant kernel */
    double ms_dense_scan;        /* device time in those launches */
    int64_t dense_scan_bytes;    /* algorithmic bytes they moved (SURVEY 8d model: 8 + 8 deg per vertex + the tests' operands) */
    int64_t dense_scan_vertices; /* vertices they scanned */
    int64_t bulk_rollbacks;      /* stage-1 bulk rounds undone by the energy guard (0 or 1 per phase) */
    int64_t sparse_rounds;       /* exact rounds run inside the persistent sparse-round kernel (k_sparse_rounds) */
    double ms_sparse;            /* device time of those rounds (enumerate + evaluate + commit, %globaltimer inside the kernel) */
    int64_t sparse_cluster_rounds; /* of those, rounds run by the one-cluster form of the kernel (a few thousand dirty vertices or fewer) */
} acvd_report;

/* MinimizeEnergy (Common/vtkUniformClustering.h:725-830) with ProcessOneLoop (:833-995) replaced by
 * conflict-free parallel rounds.  Phases, convergence events, CleanClustering / FillHoles /
 * ReComputeStatistics follow the reference schedule. */
int acvd_minimize(acvd_ctx* ctx, const acvd_params* params, acvd_report* report);

/* ReComputeStatistics (:376-403): per-cluster sums, sizes, centroid, energy from the clustering. */
int acvd_recompute_statistics(acvd_ctx* ctx, int constrained, int quadrics_level);
/* CleanClustering (:406-549) / FillHolesInClustering (:552-633); *disconnected = clusters cleaned. */
int acvd_clean_clustering(acvd_ctx* ctx, int32_t* disconnected);
/* connexity = the engine's ConnexityConstraint at the time of the call (the :606-607 guard); the adoption order is the
 * reference's FIFO order (edge ids = first-seen order over the faces), so the result is the reference's bit for bit */
int acvd_fill_holes(acvd_ctx* ctx, int connexity);
/* vtkVerticesProcessing::ConnexityConstraintProblemLocal (DiscreteRemeshing/vtkVerticesProcessing.h:168-237) as the
 * kernels evaluate it, on n caller-given (item, cluster) pairs against the current clustering: out[i] = 1 when taking
 * items[i] out of clusters[i] would disconnect its ring neighbours in that cluster.  mode 0 = the product's choice (ring
 * bit matrix for rows <= 8, generic walk for longer rows), 1 = generic walk for every row.  Parity hook for the tests. */
int acvd_connexity_problem(acvd_ctx* ctx, int32_t n, const int32_t* items, const int32_t* clusters, int32_t mode, uint8_t* out /*n*/);
/* One reassignment round on the current state (ProcessOneLoop analogue); outputs may be NULL. */
int acvd_reassign_round(acvd_ctx* ctx, int constrained, int quadrics_level, int connexity,
                        int64_t* proposals, int64_t* modifications, int64_t* tests);
/* any pointer may be NULL; sums is K x acvd_payload_size(metric) */
int acvd_get_cluster_stats(acvd_ctx* ctx, double* sums, double* centroid /*3K*/, double* energy /*K*/,
                           int32_t* sizes /*K*/);
int acvd_global_energy(acvd_ctx* ctx, double* energy);
int acvd_get_energy_log(acvd_ctx* ctx, double* out, int32_t cap, int32_t* n);
/* seconds since the start of the acvd_minimize call at which each entry of the energy log was taken: the time column of
 * energy.txt (the reference's timer, Common/vtkUniformClustering.h:690-706, stamps every loop) */
int acvd_get_energy_times(acvd_ctx* ctx, double* out, int32_t cap, int32_t* n);

/* vtkQuadricTools::ComputeRepresentativePoint (Common/vtkQuadricTools.cxx:168-177), batched:
 * n quadrics (9 doubles each) and points (3 doubles each, updated in place), rank deficiency out. */
int acvd_representative_points(acvd_ctx* ctx, int32_t n, const double* quadrics9, double* points3,
                               int32_t max_sv, double sv_threshold, int32_t* rank_deficiency);

/* Accumulation step of ACVD's quadric post-process (DiscreteRemeshing/Examples/ACVD.cxx:237-262): for the first
 * n_clusters clusters, the sum over the cluster's items of the quadrics of the input faces around each item
 * (vtkQuadricTools::AddTriangleQuadric, 9 coefficients).  Followed on the host side by acvd_representative_points. */
int acvd_cluster_quadrics(acvd_ctx* ctx, int32_t n_clusters, double* quadrics9 /*9 n_clusters*/);

/* ---- integer stages after the hot path (vtkDiscreteRemeshing.h:1003-1133) ---------------- */
/* boundary flag per vertex (has a neighbour in another cluster) */
int acvd_boundary_flags(acvd_ctx* ctx, uint8_t* flags /*V*/);
/* sorted unique (lo<<32|hi) cluster pairs joined by a mesh edge; pass out=NULL to query the count */
int acvd_cluster_adjacency(acvd_ctx* ctx, int64_t* out, int64_t cap, int64_t* n);
/* dual-mesh triangles in first-occurrence order over input faces; out=NULL queries the count */
int acvd_dual_triangles(acvd_ctx* ctx, int32_t* out /*3*cap*/, int64_t cap, int64_t* n);

/* vtkSurfaceBase::IsVertexManifold (Common/vtkSurfaceBase.cxx:259-317) as DetectNonManifoldOutputVertices uses it
 * (DiscreteRemeshing/vtkDiscreteRemeshing.h:166-383, the -m 1 loop): per input vertex, and per output vertex of the dual
 * mesh of the current clustering (with force_manifold_edges the dual edges between adjacent clusters that share no
 * triangle are part of the mesh, :1114-1133).  flags: 1 manifold, 0 not, 2 = more than 64 edges (caller's fallback). */
int acvd_input_manifold_flags(acvd_ctx* ctx, uint8_t* flags /*V*/);
int acvd_output_manifold_flags(acvd_ctx* ctx, int32_t force_manifold_edges, uint8_t* flags /*K*/);
/* vtkDiscreteRemeshing::DetectNonManifoldOutputVertices (DiscreteRemeshing/vtkDiscreteRemeshing.h:166-383), one step of
 * the -m 1 loop on the current clustering: every cluster is frozen except the non-manifold output vertices (whose items
 * are all manifold input vertices) and their output neighbours; one new cluster is appended per issue, seeded with the
 * first item of the offending cluster or, for a one-item cluster, with its first ring neighbour whose cluster has more
 * than one item.  The context's cluster count, clustering and frozen flags are updated (fetch them with
 * acvd_get_clustering); the caller then re-enters acvd_minimize with connexity 0 (:942-943). */
int acvd_detect_non_manifold(acvd_ctx* ctx, int32_t force_manifold_edges, int32_t* n_issues, int32_t* new_num_clusters);

/* ---- measurement hook ---------------------------------------------------------------------- */
/* Times `reps` back-to-back launches of one kernel on the current state with CUDA events on the library's
 * stream (scripts/bench_kernels.py).  kernel 0 = dense bulk scan of the reassignment loop; variant >= 0 selects a
 * (stages, blocks/SM) instantiation of the TMA-staged kernel, -1 the list-based scan; stage = bulk stage. */
int acvd_bench_kernel(acvd_ctx* ctx, int kernel, int variant, int stage, int reps, float* ms_per_launch);

/* ---- multi-GPU (one process per GPU; vertex-range partition) ------------------ */
#define ACVD_NCCL_ID_BYTES 128
int acvd_dist_unique_id(void* id_out /*ACVD_NCCL_ID_BYTES*/);
int acvd_dist_init(acvd_ctx* ctx, int rank, int world, const void* id /*ACVD_NCCL_ID_BYTES*/);
/* The partition the library uses across ranks (pure host arithmetic, callable without a device): out[0..1] = the
 * 32-vertex tiles (a contiguous vertex range) rank scans and evaluates, out[2..3] = the clusters it runs the cluster
 * pass (statistics, connectivity) on, out[4..5] / out[6..7] = the points / faces of the mesh it uploads. */
int acvd_dist_partition(int64_t V, int64_t F, int32_t K, int32_t rank, int32_t world, int64_t* out /*8*/);

#ifdef __cplusplus
}
#endif
#endif /* ACVD_B200_H */
