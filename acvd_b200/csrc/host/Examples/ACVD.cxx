// ACVD front-end on the B200 engine: same command line as the reference's
// DiscreteRemeshing/Examples/ACVD.cxx (file nvertices gradation [-key value]...), same output files
// (smooth_<outputfile> before the quadric post-process, then <outputfile>, default simplification.ply).
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "../vtkDiscreteRemeshing.h"

using std::cout;
using std::endl;

int main(int argc, char* argv[]) {
    int Display = 0, NumberOfSamples = 0, SubsamplingThreshold = 10;
    int QuadricsOptimizationLevel = 1;   // as in the reference (ACVD.cxx:66), although its usage text says 3
    double Gradation = 0;
    char* OutputDirectory = 0;
    const char* outputfile = "simplification.ply";
    if (argc > 3) {
        cout << "load : " << argv[1] << endl;
        NumberOfSamples = atoi(argv[2]);
        Gradation = atof(argv[3]);
    } else {
        cout << "Usage : ACVD file nvertices gradation [options]" << endl;
        cout << "nvertices is the desired number of vertices" << endl;
        cout << "gradation defines the influence of local curvature (0=uniform meshing)" << endl;
        cout << endl << "Optionnal arguments : " << endl;
        cout << "-b 0/1 : sets mesh boundary fixing off/on (default : 0)" << endl;
        cout << "-s threshold : defines the subsampling threshold i.e. the input mesh will be subdivided until inputVertices > outputVertices * ratio" << endl;
        cout << "-o directory : sets the output directory " << endl;
        cout << "-of file : sets the output file name " << endl;
        cout << "-d 0/1/2 : enables display (accepted and ignored by this build)" << endl;
        cout << "-l ratio : split the edges longer than ( averageLength * ratio )" << endl;
        cout << "-q 0/1/2 : set the number of eigenvalues for quadrics post-processing (default : 3)" << endl;
        cout << "-m 0/1 : enforce a manifold output ON/OFF (default : 0)" << endl;
        cout << "-w 0/1 : write the energy log energy.txt" << endl;
        cout << "-dev n : CUDA device index (this build)" << endl;
        return 0;
    }
    vtkSurface* Mesh = vtkSurface::New();
    Mesh->CreateFromFile(argv[1]);
    Mesh->DisplayMeshProperties();
    vtkIsotropicDiscreteRemeshing* Remesh = vtkIsotropicDiscreteRemeshing::New();
    // optional arguments: key/value pairs, unknown keys ignored, always stepping by two (ACVD.cxx:113-192)
    for (int i = 4; i + 1 < argc; i += 2) {
        char* key = argv[i];
        char* value = argv[i + 1];
        if (strcmp(key, "-m") == 0) { Remesh->SetForceManifold(atoi(value)); cout << "Force Manifold=" << atoi(value) << endl; }
        if (strcmp(key, "-s") == 0) { SubsamplingThreshold = atoi(value); cout << "Subsampling Threshold=" << SubsamplingThreshold << endl; }
        if (strcmp(key, "-d") == 0) { Display = atoi(value); cout << "Display=" << Display << endl; }
        if (strcmp(key, "-np") == 0) { cout << "Number of threads=" << atoi(value) << " (ignored: GPU engine)" << endl; }
        if (strcmp(key, "-o") == 0) { OutputDirectory = value; cout << "OutputDirectory: " << OutputDirectory << endl; Remesh->SetOutputDirectory(value); }
        if (strcmp(key, "-of") == 0) { outputfile = value; cout << "Output file name: " << outputfile << endl; }
        if (strcmp(key, "-l") == 0) {
            cout << "Splitting edges longer than " << atof(value) << " times the average edge length" << endl;
            Mesh->SplitLongEdges(atof(value));
        }
        if (strcmp(key, "-w") == 0) { cout << "Setting writing energy log file to " << atoi(value) << endl; Remesh->SetWriteToGlobalEnergyLog(atoi(value)); }
        if (strcmp(key, "-q") == 0) { cout << "Setting number of eigenvalues for quadrics to " << atoi(value) << endl; QuadricsOptimizationLevel = atoi(value); }
        if (strcmp(key, "-cd") == 0) { cout << "Setting custom file for density info : " << value << endl; Remesh->SetInputDensityFile(value); }
        if (strcmp(key, "-cmax") == 0) { cout << "Setting maximum custom density to : " << value << endl; Remesh->SetMaxCustomDensity(atof(value)); }
        if (strcmp(key, "-cmin") == 0) { cout << "Setting minimum custom density to : " << value << endl; Remesh->SetMinCustomDensity(atof(value)); }
        if (strcmp(key, "-cf") == 0) { cout << "Setting custom density multiplication factor to : " << value << endl; Remesh->SetCustomDensityMultiplicationFactor(atof(value)); }
        if (strcmp(key, "-b") == 0) { cout << "Setting boundary fixing to : " << value << endl; Remesh->SetBoundaryFixing(atoi(value)); }
        if (strcmp(key, "-dev") == 0) Remesh->SetDevice(atoi(value));
    }
    Remesh->SetInput(Mesh);
    Remesh->SetFileLoadSaveOption(0);
    Remesh->SetNumberOfClusters(NumberOfSamples);
    Remesh->SetConsoleOutput(2);
    Remesh->SetSubsamplingThreshold(SubsamplingThreshold);
    Remesh->GetMetric()->SetGradation(Gradation);
    Remesh->SetDisplay(Display);
    Remesh->Remesh();
    if (!Remesh->GetOutput()) return 1;

    const std::string dir = OutputDirectory ? OutputDirectory : "";
    if (QuadricsOptimizationLevel != 0) {
        // quadric post-process (ACVD.cxx:217-276): per cluster, sum the quadrics of the input faces around its
        // vertices and move the output vertex to the representative point of that quadric
        vtkIntArray* Clustering = Remesh->GetClustering();
        Remesh->GetOutput()->WriteToFile((dir + "smooth_" + outputfile).c_str());
        // accumulation on the device (acvd_cluster_quadrics: a warp per cluster over its items' faces), then the batched
        // representative-point solve (vtkQuadricTools::ComputeRepresentativePoint); as in the reference the arrays are
        // sized by NumberOfSamples, so items of clusters appended by the -m loop count as misclassed (ACVD.cxx:237-249)
        const int nq = std::min(NumberOfSamples, Remesh->GetNumberOfClusters());
        std::vector<double> Q((size_t)NumberOfSamples * 9, 0.0);
        int misclassed = 0;
        for (int i = 0; i < Remesh->GetNumberOfItems(); i++) {
            const int c = Clustering->GetValue(i);
            if (c < 0 || c >= NumberOfSamples) misclassed++;
        }
        if (misclassed) cout << misclassed << " Items with wrong cluster association" << endl;
        std::vector<double> P((size_t)NumberOfSamples * 3);
        for (int i = 0; i < NumberOfSamples; i++) Remesh->GetOutput()->GetPoint(i, &P[(size_t)i * 3]);
        acvd_ctx* ctx = Remesh->GetContext();
        if (!ctx || acvd_cluster_quadrics(ctx, nq, Q.data()) != ACVD_OK ||
            acvd_representative_points(ctx, NumberOfSamples, Q.data(), P.data(), QuadricsOptimizationLevel, 1e-3, nullptr) != ACVD_OK)
            cout << "ERROR : " << acvd_last_error(ctx) << endl;
        for (int i = 0; i < NumberOfSamples; i++) Remesh->GetOutput()->SetPointCoordinates(i, &P[(size_t)i * 3]);
        cout << "After Quadrics Post-processing : " << endl;
        Remesh->GetOutput()->DisplayMeshProperties();
    }
    std::string real = OutputDirectory ? std::string(OutputDirectory) + "/" : "";
    real += outputfile;
    Remesh->GetOutput()->WriteToFile(real.c_str());
    Remesh->Delete();
    Mesh->Delete();
    return 0;
}
