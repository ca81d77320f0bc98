// Host-side mirror of the reference's remeshing classes for the three front-ends:
//   vtkIsotropicDiscreteRemeshing   (ACVD)                 DiscreteRemeshing/vtkIsotropicDiscreteRemeshing.h:42,55
//   vtkQIsotropicDiscreteRemeshing  (ACVDQ)                DiscreteRemeshing/vtkIsotropicDiscreteRemeshing.h:79-80
//   vtkAnisotropicDiscreteRemeshing (AnisotropicRemeshingQ) DiscreteRemeshing/vtkAnisotropicDiscreteRemeshing.h:40,49
// Same method names, argument meaning and error behaviour (console message + return, no exceptions) as
// vtkDiscreteRemeshing<Metric> / vtkUniformClustering<Metric>
// (DiscreteRemeshing/vtkDiscreteRemeshing.h:61-83, Common/vtkUniformClustering.h:62-134).
// The clustering engine behind them is the CUDA library: every call into the hot path goes through the
// C ABI of include/acvd_b200.h, at the three points where the reference calls its engine
// (ProcessClustering :897, MinimizeEnergy in the -m loop :943, ReComputeStatistics :930).
#pragma once
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "../../../include/acvd_b200.h"
#include "vtkSurface.h"

// the part of the Metric concept the front-ends reach through GetMetric()
class vtkMetricOptions {
public:
    void SetGradation(double g) { Gradation = g; }
    double GetGradation() const { return Gradation; }
    void SetQuadricsOptimizationLevel(int l) { QuadricsOptimizationLevel = l; }
    int GetQuadricsOptimizationLevel() const { return QuadricsOptimizationLevel; }
    int IsCurvatureIndicatorNeeded() const { return (Kind >= ACVD_METRIC_ANISO) ? 1 : (Gradation > 0 ? 1 : 0); }
    int IsPrincipalDirectionsNeeded() const { return Kind >= ACVD_METRIC_ANISO ? 1 : 0; }
    int Kind = ACVD_METRIC_ISO;
    double Gradation = 0;
    int QuadricsOptimizationLevel = 3;
};

struct vtkClusterInfo { double Centroid[3]; double Energy; int Size; vtkIdType AnchorItem; };

class vtkDiscreteRemeshingB200 {
public:
    void Delete() { delete this; }

    // ---- reference API
    void SetInput(vtkSurface* s);
    vtkSurface* GetInput() { return Input; }
    void SetNumberOfClusters(int n);
    int GetNumberOfClusters() const { return NumberOfClusters; }
    void SetConsoleOutput(int v) { ConsoleOutput = v; }
    void SetSubsamplingThreshold(int v) { SubsamplingThreshold = v; }
    void SetForceManifold(bool v) { ForceManifold = v; }
    void SetBoundaryFixing(int v) { BoundaryFixing = v; }
    void SetFileLoadSaveOption(int v) { FileLoadSaveOption = v; }
    void SetDisplay(int v) { Display = v; }                    // accepted and ignored (no rendering in this build)
    void SetAnchorRenderWindow(void*) {}
    void SetOutputDirectory(char* d) { OutputDirectory = d ? d : ""; }
    void SetWriteToGlobalEnergyLog(int v) { WriteEnergyLog = v; }
    void SetUnconstrainedInitialization(int v) { UnconstrainedInitialization = v; }
    void SetFixedClusters(vtkIdList* l) { FixedClusters = l ? l->ids : std::vector<vtkIdType>(); }
    void SetInitialClustering(vtkIntArray* a) { InitialClustering = a ? a->v : std::vector<int>(); }
    void SetMaxNumberOfLoops(int v) { MaxNumberOfLoops = v; }
    void SetMaxNumberOfConvergences(int v) { MaxNumberOfConvergences = v; }
    void SetNumberOfThreads(int) {}                            // P variants only: the GPU engine ignores it
    void SetPoolingRatio(int) {}
    void SetInputDensityFile(const char*);                     // -cd: custom density volumes are not supported by this build
    void SetMaxCustomDensity(double) {}
    void SetMinCustomDensity(double) {}
    void SetCustomDensityMultiplicationFactor(double) {}
    vtkMetricOptions* GetMetric() { return &Metric; }
    vtkClusterInfo* GetCluster(int i) { return &Clusters[(size_t)i]; }
    int GetClusterSize(int i) { return Clusters[(size_t)i].Size; }
    vtkIntArray* GetClustering() { return Clustering; }
    int GetNumberOfItems() { return Input ? (int)Input->GetNumberOfPoints() : 0; }
    int GetClusteringType() const { return 1; }                // items are vertices (vtkVerticesProcessing)
    vtkSurface* GetOutput() { return Output; }
    void Remesh();                                             // vtkDiscreteRemeshing.h:877-954

    // ---- extensions of this build
    // curvature indicator / principal directions supplied by the caller instead of vtkCurvatureMeasure
    void SetCurvatureIndicator(const double* values, vtkIdType n) { CustomIndicator.assign(values, values + n); }
    void SetPrincipalDirections(const float* values, vtkIdType n6) { PrincipalDirections.assign(values, values + n6); }
    void SetDevice(int d) { Device = d; }
    const acvd_report& GetReport() const { return Report; }
    acvd_ctx* GetContext() { return Ctx; }                      // the engine's context (mesh + final clustering stay resident)

protected:
    explicit vtkDiscreteRemeshingB200(int metric_kind);
    virtual ~vtkDiscreteRemeshingB200();

    void CheckSubsamplingRatio();                              // :841-875
    void SamplingPreProcessing();                              // :583-838 (curvature indicator)
    int ProcessClustering();                                   // vtkUniformClustering.h:655-714
    void MinimizeEnergy(int connexity);                        // re-entry of the -m loop (:943)
    void FetchClusters();
    void BuildDelaunayTriangulation();                         // :1003-1133
    int DetectNonManifoldOutputVertices();                     // :166-383
    bool Check(int rc, const char* what);

    vtkSurface* Input = nullptr;
    vtkSurface* OriginalInput = nullptr;
    vtkSurface* Output = nullptr;
    vtkIntArray* Clustering = nullptr;
    std::vector<vtkClusterInfo> Clusters;
    std::vector<unsigned char> Frozen;
    vtkMetricOptions Metric;
    acvd_ctx* Ctx = nullptr;
    acvd_report Report;
    int Device = -1;
    int NumberOfClusters = 0;
    int ConsoleOutput = 0;
    int SubsamplingThreshold = 10;
    bool ForceManifold = false;
    int BoundaryFixing = 0;
    int FileLoadSaveOption = 0;
    int Display = 0;
    int WriteEnergyLog = 0;
    int UnconstrainedInitialization = 0;
    int MaxNumberOfLoops = 5000000;
    int MaxNumberOfConvergences = 1000000000;
    int NumberOfSubdivisionsBeforeClustering = 0;
    std::string OutputDirectory;
    std::vector<vtkIdType> FixedClusters;
    std::vector<int> InitialClustering;
    std::vector<double> CustomIndicator;
    std::vector<float> PrincipalDirections;
    std::vector<int> Parent1, Parent2;
};

class vtkIsotropicDiscreteRemeshing : public vtkDiscreteRemeshingB200 {
public:
    static vtkIsotropicDiscreteRemeshing* New() { return new vtkIsotropicDiscreteRemeshing; }
protected:
    vtkIsotropicDiscreteRemeshing() : vtkDiscreteRemeshingB200(ACVD_METRIC_ISO) {}
};

class vtkQIsotropicDiscreteRemeshing : public vtkDiscreteRemeshingB200 {
public:
    static vtkQIsotropicDiscreteRemeshing* New() { return new vtkQIsotropicDiscreteRemeshing; }
protected:
    vtkQIsotropicDiscreteRemeshing() : vtkDiscreteRemeshingB200(ACVD_METRIC_QEM) {}
};

// the reference instantiates vtkVerticesProcessing<vtkDiscreteRemeshing<vtkQuadricAnisotropicMetricForClustering>>
// for AnisotropicRemeshingQ (Examples/AnisotropicRemeshingQ.cxx:80-84)
class vtkAnisotropicDiscreteRemeshing : public vtkDiscreteRemeshingB200 {
public:
    static vtkAnisotropicDiscreteRemeshing* New() { return new vtkAnisotropicDiscreteRemeshing; }
protected:
    vtkAnisotropicDiscreteRemeshing() : vtkDiscreteRemeshingB200(ACVD_METRIC_ANISOQ) {}
};
