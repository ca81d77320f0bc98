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
        vtkIntArray *p1 = vtkIntArray::New(), *p2 = vtkIntArray::New();
        vtkSurface* next = Input->Subdivide(p1, p2);
        // parents of the new level expressed in vertices of the level below; old vertices are their own parents
        Parent1 = p1->v; Parent2 = p2->v;
        p1->Delete(); p2->Delete();
        if (!CustomIndicator.empty()) {
            // the curvature indicator is interpolated linearly to the midpoints (:733-745)
            const size_t n_old = (size_t)Input->GetNumberOfPoints(), n_new = (size_t)next->GetNumberOfPoints();
            std::vector<double> ind(n_new);
            for (size_t i = 0; i < n_old; i++) ind[i] = CustomIndicator[i];
            for (size_t i = n_old; i < n_new; i++) ind[i] = 0.5 * (CustomIndicator[(size_t)Parent1[i]] + CustomIndicator[(size_t)Parent2[i]]);
            CustomIndicator.swap(ind);
        }
        if (!OriginalInput) OriginalInput = Input; else Input->Delete();
        Input = next;
        NumberOfSubdivisionsBeforeClustering++;
    }
}

// Curvature indicator sqrt(k1^2 + k2^2) and principal directions when the metric needs them and the caller
// supplied none.  The reference fits a polynomial patch over the 3-ring (vtkCurvatureMeasure, SURVEY §8f-2, not
// rebuilt yet); this build fits the second fundamental form to the normal curvatures of the 1-ring edges.
// Runs that rely on it say so on the console.
void vtkDiscreteRemeshingB200::SamplingPreProcessing() {
    const bool need_ind = Metric.IsCurvatureIndicatorNeeded() && CustomIndicator.empty();
    const bool need_pd = Metric.IsPrincipalDirectionsNeeded() && PrincipalDirections.empty();
    if (!need_ind && !need_pd) return;
    const vtkIdType nv = Input->GetNumberOfPoints(), nf = Input->GetNumberOfCells();
    cout << "Curvature: discrete per-vertex estimate (normal-curvature tensor over the 1-ring); vtkCurvatureMeasure's "
            "polynomial fitting over the 3-ring is not part of this build" << endl;
    const float* X = Input->Points();
    const int* T = Input->Triangles();
    // area-weighted vertex normals
    std::vector<double> nrm((size_t)nv * 3, 0.0);
    for (vtkIdType f = 0; f < nf; f++) {
        const int v[3] = {T[3 * f], T[3 * f + 1], T[3 * f + 2]};
        double p[3][3];
        for (int k = 0; k < 3; k++) for (int d = 0; d < 3; d++) p[k][d] = X[3 * (size_t)v[k] + d];
        const double e1[3] = {p[1][0] - p[0][0], p[1][1] - p[0][1], p[1][2] - p[0][2]};
        const double e2[3] = {p[2][0] - p[0][0], p[2][1] - p[0][1], p[2][2] - p[0][2]};
        const double cr[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
        for (int k = 0; k < 3; k++) for (int d = 0; d < 3; d++) nrm[3 * (size_t)v[k] + d] += cr[d];
    }
    if (need_ind) CustomIndicator.assign((size_t)nv, 0.0);
    if (need_pd) PrincipalDirections.assign((size_t)nv * 6, 0.0f);
    vtkIdList* nb = vtkIdList::New();
    for (vtkIdType v = 0; v < nv; v++) {
        double* n = &nrm[3 * (size_t)v];
        const double nl = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        if (nl <= 0) continue;
        for (int d = 0; d < 3; d++) n[d] /= nl;
        // tangent frame
        double a[3] = {std::fabs(n[0]) < 0.9 ? 1.0 : 0.0, std::fabs(n[0]) < 0.9 ? 0.0 : 1.0, 0.0};
        double t1[3] = {a[1] * n[2] - a[2] * n[1], a[2] * n[0] - a[0] * n[2], a[0] * n[1] - a[1] * n[0]};
        const double t1l = std::sqrt(t1[0] * t1[0] + t1[1] * t1[1] + t1[2] * t1[2]);
        for (int d = 0; d < 3; d++) t1[d] /= t1l;
        const double t2[3] = {n[1] * t1[2] - n[2] * t1[1], n[2] * t1[0] - n[0] * t1[2], n[0] * t1[1] - n[1] * t1[0]};
        // least squares for the second fundamental form: kn(theta) = A c^2 + 2 B c s + C s^2
        double M[3][3] = {{0}}, rhs[3] = {0, 0, 0};
        Input->GetVertexNeighbours(v, nb);
        for (vtkIdType j = 0; j < nb->GetNumberOfIds(); j++) {
            const vtkIdType u = nb->GetId(j);
            double e[3];
            for (int d = 0; d < 3; d++) e[d] = (double)X[3 * (size_t)u + d] - (double)X[3 * (size_t)v + d];
            const double l2 = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
            if (l2 <= 0) continue;
            const double kn = -2.0 * (n[0] * e[0] + n[1] * e[1] + n[2] * e[2]) / l2;   // outward normal: convex -> positive
            double c = e[0] * t1[0] + e[1] * t1[1] + e[2] * t1[2], sn = e[0] * t2[0] + e[1] * t2[1] + e[2] * t2[2];
            const double cl = std::sqrt(c * c + sn * sn);
            if (cl <= 0) continue;
            c /= cl; sn /= cl;
            const double row[3] = {c * c, 2 * c * sn, sn * sn};
            for (int r = 0; r < 3; r++) { rhs[r] += row[r] * kn; for (int q = 0; q < 3; q++) M[r][q] += row[r] * row[q]; }
        }
        // solve the 3x3 normal equations (Cramer)
        auto det3 = [](double m[3][3]) {
            return m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
                   m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
        };
        const double D = det3(M);
        if (std::fabs(D) < 1e-300) continue;
        double sol[3];
        for (int k = 0; k < 3; k++) {
            double Mk[3][3];
            for (int r = 0; r < 3; r++) for (int q = 0; q < 3; q++) Mk[r][q] = (q == k) ? rhs[r] : M[r][q];
            sol[k] = det3(Mk) / D;
        }
        const double A = sol[0], B = sol[1], C = sol[2];
        const double tr = A + C, df = std::sqrt((A - C) * (A - C) * 0.25 + B * B);
        double k1 = 0.5 * tr + df, k2 = 0.5 * tr - df;
        double ang = 0.5 * std::atan2(2 * B, A - C);
        double d1[3], d2[3];
        for (int d = 0; d < 3; d++) { d1[d] = std::cos(ang) * t1[d] + std::sin(ang) * t2[d]; d2[d] = -std::sin(ang) * t1[d] + std::cos(ang) * t2[d]; }
        if (std::fabs(k2) > std::fabs(k1)) { std::swap(k1, k2); for (int d = 0; d < 3; d++) std::swap(d1[d], d2[d]); }
        if (need_ind) CustomIndicator[(size_t)v] = std::sqrt(k1 * k1 + k2 * k2);
        if (need_pd) {
            // (sqrt|ka| da, sqrt|kb| db), larger |k| first, float32 (vtkCurvatureMeasure.cxx:445-484, :742)
            const double s1 = std::sqrt(std::fabs(k1)), s2 = std::sqrt(std::fabs(k2));
            for (int d = 0; d < 3; d++) { PrincipalDirections[6 * (size_t)v + d] = (float)(s1 * d1[d]); PrincipalDirections[6 * (size_t)v + 3 + d] = (float)(s2 * d2[d]); }
        }
    }
    nb->Delete();
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
        // a loop here is a reassignment round and the time column is the round's share of the total
        int32_t n = 0;
        acvd_get_energy_log(Ctx, nullptr, 0, &n);
        std::vector<double> log((size_t)n);
        acvd_get_energy_log(Ctx, log.data(), n, &n);
        std::ofstream f(OutputDirectory + "energy.txt", std::ofstream::out | std::ofstream::trunc);
        for (int i = 0; i < n; i++)
            f << i << " " << Report.ms_total * 1e-3 * (i + 1) / std::max(1, n) << " " << std::setprecision(15) << log[(size_t)i] << std::setprecision(6) << endl;
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
// with an item taken from it (or, if it has a single item, from a neighbouring cluster).
int vtkDiscreteRemeshingB200::DetectNonManifoldOutputVertices() {
    cout << "Starting detection of non-manifold vertices" << endl;
    const vtkIdType nv = Input->GetNumberOfPoints();
    int K = NumberOfClusters;
    std::vector<std::vector<int>> items((size_t)K);
    int* cl = Clustering->GetPointer(0);
    int wrong = 0;
    for (vtkIdType i = 0; i < nv; i++) {
        if (cl[i] < 0 || cl[i] >= K) { wrong++; cl[i] = -1; }   // NULL items: re-labelled with the new NULL id below
        else items[(size_t)cl[i]].push_back((int)i);
    }
    if (wrong) cout << wrong << " uncorrectly associated items" << endl;
    Frozen.assign((size_t)K, 1);
    std::vector<int> issues;
    vtkIdList* nb = vtkIdList::New();
    for (int c = 0; c < K; c++) {
        if (Output->IsVertexManifold(c)) continue;
        cout << "Cluster " << c << " is non manifold" << endl;
        if (items[(size_t)c].empty()) { cout << ".... but empty. Skipping" << endl; continue; }
        cout << items[(size_t)c].size() << " items inside" << endl;
        bool problem = true;
        for (int it : items[(size_t)c]) if (!Input->IsVertexManifold(it)) {
            problem = false;
            cout << "discarding this topology issue as the input mesh also has a topology issue here" << endl;
            break;
        }
        if (!problem) continue;
        issues.push_back(c);
        Frozen[(size_t)c] = 0;
        Output->GetVertexNeighbours(c, nb);
        for (vtkIdType i = 0; i < nb->GetNumberOfIds(); i++) Frozen[(size_t)nb->GetId(i)] = 0;
    }
    for (int c : issues) {
        const int fresh = K;
        bool placed = false;
        auto& mine = items[(size_t)c];
        if (mine.size() > 1) {
            cl[mine.front()] = fresh;
            items.push_back({mine.front()});
            mine.erase(mine.begin());
            placed = true;
        } else {
            Input->GetVertexNeighbours(mine.front(), nb);
            for (vtkIdType j = 0; j < nb->GetNumberOfIds() && !placed; j++) {
                const int u = (int)nb->GetId(j), cu = cl[u];
                if (cu < 0 || cu >= (int)items.size() || items[(size_t)cu].size() <= 1) continue;
                cl[u] = fresh;
                auto& other = items[(size_t)cu];
                other.erase(std::find(other.begin(), other.end(), u));
                items.push_back({u});
                placed = true;
            }
            if (!placed) cout << "Could not find a place to add cluster " << fresh << " near cluster " << c << endl;
        }
        if (placed) { K++; Frozen.push_back(0); }
    }
    nb->Delete();
    // unassigned items carry the NULL id, which is the (possibly grown) cluster count
    for (vtkIdType i = 0; i < nv; i++) if (cl[i] < 0) cl[i] = K;
    if (K != NumberOfClusters) {
        NumberOfClusters = K;
        Clusters.resize((size_t)K, vtkClusterInfo{{0, 0, 0}, 0, 0, -1});
    }
    return (int)issues.size();
}

void vtkDiscreteRemeshingB200::Remesh() {
    if (!Input) { cout << "ERROR : no input mesh" << endl; return; }
    CheckSubsamplingRatio();
    SamplingPreProcessing();
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
            // the cluster count may have grown: re-create the cluster tables, keep the clustering, freeze the rest
            std::vector<int> keep(Clustering->v);
            if (!Check(acvd_set_num_clusters(Ctx, NumberOfClusters), "acvd_set_num_clusters")) return;
            if (!FixedClusters.empty()) {
                std::vector<int64_t> fx(FixedClusters.begin(), FixedClusters.end());
                Check(acvd_set_fixed_clusters(Ctx, fx.data(), (int32_t)fx.size()), "acvd_set_fixed_clusters");
            }
            if (!Check(acvd_set_clustering(Ctx, keep.data()), "acvd_set_clustering")) return;
            if (!Check(acvd_set_frozen(Ctx, Frozen.data()), "acvd_set_frozen")) return;
            MinimizeEnergy(0);   // ConnexityConstraint = 0 (:942)
            BuildDelaunayTriangulation();
        }
        if (ConsoleOutput) Output->DisplayMeshProperties();
    }
    if (BoundaryFixing) cout << "Boundary fixing (-b 1) is not part of this build; the synthetic workloads are closed surfaces" << endl;
}
