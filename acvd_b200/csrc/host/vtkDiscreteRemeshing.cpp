// Host side of the remeshing front-ends (see vtkDiscreteRemeshing.h).  The clustering runs on the GPU
// through the C ABI; this file holds the reference's orchestration around it.
#include "vtkDiscreteRemeshing.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <iostream>

using std::cout;
using std::endl;

vtkDiscreteRemeshingB200::vtkDiscreteRemeshingB200(int metric_kind) {
    Metric.Kind = metric_kind;
    Clustering = vtkIntArray::New();
    memset(&Report, 0, sizeof Report);
}

vtkDiscreteRemeshingB200::~vtkDiscreteRemeshingB200() {
    if (Ctx) acvd_destroy(Ctx);
    if (Output) Output->Delete();
    if (OriginalInput && Input) Input->Delete();   // the subdivided level we created
    Clustering->Delete();
}

bool vtkDiscreteRemeshingB200::Check(int rc, const char* what) {
    if (rc == ACVD_OK) return true;
    // the reference reports problems on the console and returns (vtkUniformClustering.h:659-664)
    cout << "ERROR in " << what << " : " << acvd_last_error(Ctx) << endl;
    return false;
}

void vtkDiscreteRemeshingB200::SetInput(vtkSurface* s) {
    Input = s;
}

void vtkDiscreteRemeshingB200::SetInputDensityFile(const char* f) {
    cout << "WARNING : custom density volume " << (f ? f : "") << " ignored (-cd needs vtkImageData, not part of this build)" << endl;
}

void vtkDiscreteRemeshingB200::SetNumberOfClusters(int n) {
    NumberOfClusters = n;
    Clusters.assign((size_t)std::max(n, 0), vtkClusterInfo{{0, 0, 0}, 0, 0, -1});
}

// subdivide the input until it has at least SubsamplingThreshold x NumberOfClusters vertices (:841-875)
void vtkDiscreteRemeshingB200::CheckSubsamplingRatio() {
    while (Input->GetNumberOfPoints() < (vtkIdType)SubsamplingThreshold * NumberOfClusters) {
        if (ConsoleOutput) cout << "Subdividing mesh" << endl;
        // vtkSurface::Subdivide on the device (acvd_subdivide: same vertex / face numbering as the host method)
        if (!Ctx && !Check(acvd_create(&Ctx, Device), "acvd_create")) return;
        if (!Check(acvd_set_mesh(Ctx, (int32_t)Input->GetNumberOfPoints(), (int32_t)Input->GetNumberOfCells(), Input->Points(), Input->Triangles()), "acvd_set_mesh")) return;
        int32_t nv2 = 0, nf2 = 0;
        if (!Check(acvd_subdivide(Ctx, &nv2, &nf2), "acvd_subdivide")) return;
        std::vector<float> xyz2(3 * (size_t)nv2);
        std::vector<int> tri2(3 * (size_t)nf2);
        // parents of the new level expressed in vertices of the level below; old vertices are their own parents
        Parent1.assign((size_t)nv2, 0); Parent2.assign((size_t)nv2, 0);
        if (!Check(acvd_get_subdivision(Ctx, xyz2.data(), tri2.data(), Parent1.data(), Parent2.data()), "acvd_get_subdivision")) return;
        vtkSurface* next = vtkSurface::New();
        next->CreateFromArrays(nv2, xyz2.data(), nf2, tri2.data());
        if (!CustomIndicator.empty()) {
            // the curvature indicator is interpolated linearly to the midpoints (:733-745)
            const size_t n_old = (size_t)Input->GetNumberOfPoints(), n_new = (size_t)next->GetNumberOfPoints();
            std::vector<double> ind(n_new);
            for (size_t i = 0; i < n_old; i++) ind[i] = CustomIndicator[i];
            for (size_t i = n_old; i < n_new; i++) ind[i] = 0.5 * (CustomIndicator[(size_t)Parent1[i]] + CustomIndicator[(size_t)Parent2[i]]);
            CustomIndicator.swap(ind);
        }
        if (!PrincipalDirections.empty()) {   // same linear interpolation for the (sqrt|k| d) pairs
            const size_t n_old = (size_t)Input->GetNumberOfPoints(), n_new = (size_t)next->GetNumberOfPoints();
            std::vector<float> pd(6 * n_new);
            for (size_t i = 0; i < 6 * n_old; i++) pd[i] = PrincipalDirections[i];
            for (size_t i = n_old; i < n_new; i++)
                for (int d = 0; d < 6; d++)
                    pd[6 * i + d] = 0.5f * (PrincipalDirections[6 * (size_t)Parent1[i] + d] + PrincipalDirections[6 * (size_t)Parent2[i] + d]);
            PrincipalDirections.swap(pd);
        }
        if (!OriginalInput) OriginalInput = Input; else Input->Delete();
        Input = next;
        NumberOfSubdivisionsBeforeClustering++;
    }
}

// Curvature indicator sqrt(k1^2 + k2^2) and principal directions when the metric needs them and the caller supplied
// none: vtkCurvatureMeasure with polynomial fitting over the 3-ring (DiscreteRemeshing/vtkDiscreteRemeshing.h:640-653),
// computed on the device (acvd_curvature) on the mesh as given, i.e. before any subdivision; CheckSubsamplingRatio then
// interpolates it to the midpoints (:733-745).
void vtkDiscreteRemeshingB200::SamplingPreProcessing() {
    const bool need_ind = Metric.IsCurvatureIndicatorNeeded() && CustomIndicator.empty();
    const bool need_pd = Metric.IsPrincipalDirectionsNeeded() && PrincipalDirections.empty();
    if (!need_ind && !need_pd) return;
    const vtkIdType nv = Input->GetNumberOfPoints(), nf = Input->GetNumberOfCells();
    if (!Ctx && !Check(acvd_create(&Ctx, Device), "acvd_create")) {
        cout << "ERROR : " << acvd_last_error(nullptr) << endl;
        return;
    }
    cout << "Computing Curvature ........" << endl;
    if (!Check(acvd_set_mesh(Ctx, (int32_t)nv, (int32_t)nf, Input->Points(), Input->Triangles()), "acvd_set_mesh")) return;
    std::vector<double> ind((size_t)nv);
    std::vector<float> pd(need_pd ? 6 * (size_t)nv : 0);
    if (!Check(acvd_curvature(Ctx, 3, ind.data(), need_pd ? pd.data() : nullptr), "acvd_curvature")) return;
    if (need_ind) CustomIndicator.swap(ind);
    if (need_pd) PrincipalDirections.swap(pd);
}

void vtkDiscreteRemeshingB200::FetchClusters() {
    const int K = NumberOfClusters;
    std::vector<double> cen(3 * (size_t)K), en((size_t)K);
    std::vector<int> sz((size_t)K);
    if (!Check(acvd_get_cluster_stats(Ctx, nullptr, cen.data(), en.data(), sz.data()), "acvd_get_cluster_stats")) return;
    Clusters.resize((size_t)K);
    for (int i = 0; i < K; i++) {
        for (int d = 0; d < 3; d++) Clusters[(size_t)i].Centroid[d] = cen[3 * (size_t)i + d];
        Clusters[(size_t)i].Energy = en[(size_t)i];
        Clusters[(size_t)i].Size = sz[(size_t)i];
        Clusters[(size_t)i].AnchorItem = (size_t)i < FixedClusters.size() ? FixedClusters[(size_t)i] : -1;
    }
    Clustering->SetNumberOfValues(Input->GetNumberOfPoints());
    Check(acvd_get_clustering(Ctx, Clustering->GetPointer(0)), "acvd_get_clustering");
}

void vtkDiscreteRemeshingB200::MinimizeEnergy(int connexity) {
    acvd_params p;
    memset(&p, 0, sizeof p);
    p.unconstrained_init = UnconstrainedInitialization;
    p.quadrics_level = Metric.QuadricsOptimizationLevel;
    p.connexity = connexity;
    p.max_loops = MaxNumberOfLoops;
    p.max_convergences = MaxNumberOfConvergences;
    p.log_energy = WriteEnergyLog;
    if (!Check(acvd_minimize(Ctx, &p, &Report), "acvd_minimize")) return;
    if (WriteEnergyLog) {
        // energy.txt: "loop seconds energy" per loop + final energy (vtkUniformClustering.h:677-699, 1384-1395);
        // a loop here is a reassignment round, stamped when its energy was taken (acvd_get_energy_times)
        int32_t n = 0;
        acvd_get_energy_log(Ctx, nullptr, 0, &n);
        std::vector<double> log((size_t)n), when((size_t)n);
        acvd_get_energy_log(Ctx, log.data(), n, &n);
        acvd_get_energy_times(Ctx, when.data(), n, &n);
        std::ofstream f(OutputDirectory + "energy.txt", std::ofstream::out | std::ofstream::trunc);
        for (int i = 0; i < n; i++)
            f << i << " " << when[(size_t)i] << " " << std::setprecision(15) << log[(size_t)i] << std::setprecision(6) << endl;
        f << "Final Energy :" << std::setprecision(15) << Report.energy << endl;
    }
    FetchClusters();
}

int vtkDiscreteRemeshingB200::ProcessClustering() {
    if (NumberOfClusters == 0) {
        cout << "Problem!!! must set NumberOfClusters	to more	than zero!" << endl;
        return 0;
    }
    if (!Ctx && !Check(acvd_create(&Ctx, Device), "acvd_create")) {
        cout << "ERROR : " << acvd_last_error(nullptr) << endl;
        return 0;
    }
    const vtkIdType nv = Input->GetNumberOfPoints(), nf = Input->GetNumberOfCells();
    if (!Check(acvd_set_mesh(Ctx, (int32_t)nv, (int32_t)nf, Input->Points(), Input->Triangles()), "acvd_set_mesh")) return 0;
    const double* ind = CustomIndicator.size() == (size_t)nv ? CustomIndicator.data() : nullptr;
    const float* pd = PrincipalDirections.size() == 6 * (size_t)nv ? PrincipalDirections.data() : nullptr;
    if (Metric.IsPrincipalDirectionsNeeded() && !pd) {
        cout << "ERROR : the anisotropic metric needs principal directions" << endl;
        return 0;
    }
    if (!Check(acvd_build_items(Ctx, Metric.Kind, Metric.Gradation, ind, pd), "acvd_build_items")) return 0;
    if (!Check(acvd_set_num_clusters(Ctx, NumberOfClusters), "acvd_set_num_clusters")) return 0;
    if (!FixedClusters.empty()) {
        std::vector<int64_t> fx(FixedClusters.begin(), FixedClusters.end());
        if (!Check(acvd_set_fixed_clusters(Ctx, fx.data(), (int32_t)fx.size()), "acvd_set_fixed_clusters")) return 0;
    }
    if (InitialClustering.size() == (size_t)nv) {   // SetInitialClustering -> InitialSamplingType 2 (:337-342)
        if (!Check(acvd_set_clustering(Ctx, InitialClustering.data()), "acvd_set_clustering")) return 0;
    } else if (!Check(acvd_initial_sampling(Ctx), "acvd_initial_sampling")) return 0;
    if (ConsoleOutput) cout << "Clustering......" << endl;
    if (UnconstrainedInitialization && ConsoleOutput) cout << "Performing unconstrained initialization" << endl;
    MinimizeEnergy(0);
    if (ConsoleOutput) {
        cout << "The clustering took :" << Report.ms_total * 1e-3 << " seconds." << endl;
        cout << "Number of loops:	" << Report.rounds << endl;
    }
    return 1;
}

// Output vertices = cluster representative points; one triangle per input triangle whose three vertices
// sit in three different clusters, first occurrence only (:1003-1100).  With ForceManifold the dual
// edges between adjacent clusters that share no triangle are added too (:1114-1133).
void vtkDiscreteRemeshingB200::BuildDelaunayTriangulation() {
    if (Output) Output->Delete();
    Output = vtkSurface::New();
    const int K = NumberOfClusters;
    int valid = 0;
    for (int i = 0; i < K; i++) if (Clusters[(size_t)i].Size > 0) { valid = i; break; }
    Output->xyz.reserve(3 * (size_t)K);
    for (int i = 0; i < K; i++) {
        const double* c = Clusters[(size_t)(Clusters[(size_t)i].Size == 0 ? valid : i)].Centroid;
        Output->AddVertex(c[0], c[1], c[2]);
    }
    int64_t n = 0;
    if (!Check(acvd_dual_triangles(Ctx, nullptr, 0, &n), "acvd_dual_triangles")) return;
    Output->tri.assign(3 * (size_t)n, 0);
    if (n > 0 && !Check(acvd_dual_triangles(Ctx, Output->tri.data(), n, &n), "acvd_dual_triangles")) return;
    if (ForceManifold) {
        int64_t na = 0;
        if (!Check(acvd_cluster_adjacency(Ctx, nullptr, 0, &na), "acvd_cluster_adjacency")) return;
        std::vector<int64_t> pairs((size_t)na);
        if (na > 0 && !Check(acvd_cluster_adjacency(Ctx, pairs.data(), na, &na), "acvd_cluster_adjacency")) return;
        for (int64_t i = 0; i < na; i++) {
            const vtkIdType c1 = pairs[(size_t)i] >> 32, c2 = pairs[(size_t)i] & 0xffffffffll;
            if (Output->IsEdge(c1, c2) < 0) Output->AddEdge(c1, c2);
        }
    }
}

// :166-383.  Every cluster is frozen; a non-manifold output vertex whose input vertices are all manifold
// is a topology issue: it and its output neighbours are unfrozen and one new cluster is seeded next to it
// with an item taken from it (or, if it has a single item, from a neighbouring cluster).  The manifold tests
// (vtkSurfaceBase::IsVertexManifold on the dual mesh and on the input) run on the device: acvd_detect_non_manifold.
int vtkDiscreteRemeshingB200::DetectNonManifoldOutputVertices() {
    cout << "Starting detection of non-manifold vertices" << endl;
    int32_t issues = 0, newK = NumberOfClusters;
    if (!Check(acvd_detect_non_manifold(Ctx, 1, &issues, &newK), "acvd_detect_non_manifold")) return 0;
    if (issues) {
        if (!Check(acvd_get_clustering(Ctx, Clustering->GetPointer(0)), "acvd_get_clustering")) return 0;
        if (newK != NumberOfClusters) {
            cout << newK - NumberOfClusters << " clusters added next to non-manifold output vertices" << endl;
            NumberOfClusters = newK;
            Clusters.resize((size_t)newK, vtkClusterInfo{{0, 0, 0}, 0, 0, -1});
        }
    }
    return issues;
}

void vtkDiscreteRemeshingB200::Remesh() {
    if (!Input) { cout << "ERROR : no input mesh" << endl; return; }
    SamplingPreProcessing();      // on the mesh as given (the reference measures curvature before subdividing, :641-644)
    CheckSubsamplingRatio();
    if (ConsoleOutput)
        cout << "Input mesh: " << Input->GetNumberOfPoints() << " vertices	and	" << Input->GetNumberOfCells() << " faces" << endl;
    bool compute = true;
    if (FileLoadSaveOption) {
        int c = 1;
        cout << "Do you want to compute the clustering? (0:NO	1:Yes) ";
        std::cin >> c;
        compute = c == 1;
    }
    if (compute) {
        if (!ProcessClustering()) return;
        if (FileLoadSaveOption == 1) {
            std::ofstream o("clustering.dat", std::ofstream::out | std::ofstream::trunc | std::ios::binary);
            o.write((const char*)Clustering->GetPointer(0), (std::streamsize)(sizeof(int) * (size_t)Input->GetNumberOfPoints()));
        }
    } else {
        // load clustering.dat and recompute the statistics only (:907-931)
        InitialClustering.assign((size_t)Input->GetNumberOfPoints(), 0);
        std::ifstream in("clustering.dat", std::ios::binary);
        in.read((char*)InitialClustering.data(), (std::streamsize)(sizeof(int) * InitialClustering.size()));
        if (!Ctx && !Check(acvd_create(&Ctx, Device), "acvd_create")) return;
        const double* ind = CustomIndicator.empty() ? nullptr : CustomIndicator.data();
        const float* pd = PrincipalDirections.empty() ? nullptr : PrincipalDirections.data();
        if (!Check(acvd_set_mesh(Ctx, (int32_t)Input->GetNumberOfPoints(), (int32_t)Input->GetNumberOfCells(), Input->Points(), Input->Triangles()), "acvd_set_mesh")) return;
        if (!Check(acvd_build_items(Ctx, Metric.Kind, Metric.Gradation, ind, pd), "acvd_build_items")) return;
        if (!Check(acvd_set_num_clusters(Ctx, NumberOfClusters), "acvd_set_num_clusters")) return;
        if (!Check(acvd_set_clustering(Ctx, InitialClustering.data()), "acvd_set_clustering")) return;
        if (!Check(acvd_recompute_statistics(Ctx, 1, Metric.QuadricsOptimizationLevel), "acvd_recompute_statistics")) return;
        FetchClusters();
    }
    BuildDelaunayTriangulation();
    if (ConsoleOutput) Output->DisplayMeshProperties();
    if (ForceManifold) {
        while (int issues = DetectNonManifoldOutputVertices()) {
            cout << issues << " topology issues, restarting minimization" << endl;
            // the library holds the grown cluster tables, the edited clustering and the frozen flags
            MinimizeEnergy(0);   // ConnexityConstraint = 0 (:942)
            BuildDelaunayTriangulation();
        }
        if (ConsoleOutput) Output->DisplayMeshProperties();
    }
    if (BoundaryFixing) cout << "Boundary fixing (-b 1) is not part of this build; the synthetic workloads are closed surfaces" << endl;
}
