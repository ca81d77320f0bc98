// AnisotropicRemeshingQ front-end on the B200 engine: same command line as the reference's
// DiscreteRemeshing/Examples/AnisotropicRemeshingQ.cxx (no -m / -of / -w there either), output Remeshing.ply.
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>

#include "../vtkDiscreteRemeshing.h"

using std::cin;
using std::cout;
using std::endl;

int main(int argc, char* argv[]) {
    int Display = 0, NumberOfSamples = 200, SubsamplingThreshold = 10;
    double Gradation = 0;
    char* OutputDirectory = 0;
    if (argc <= 1) {
        cout << "Usage : AnisotropicRemeshingQ file nvertices gradation [options]" << endl;
        cout << "nvertices is the desired number of vertices" << endl;
        cout << "gradation defines the influence of local curvature (0=uniform meshing)" << endl;
        cout << endl << "Optionnal arguments : " << endl;
        cout << "-b 0/1 : sets mesh boundary fixing off/on (default : 0)" << endl;
        cout << "-d 0/1/2 : enables display (accepted and ignored by this build)" << endl;
        cout << "-l ratio : split the edges longer than ( averageLength * ratio )" << endl;
        cout << "-q 1/2/3 : qets number of eigenvalues used for quadric-based vertex relocation to 0/1/2 (default : 3)" << endl;
        return 0;
    }
    cout << "load : " << argv[1] << endl;
    vtkSurface* Mesh = vtkSurface::New();
    vtkAnisotropicDiscreteRemeshing* Remesh = vtkAnisotropicDiscreteRemeshing::New();
    Mesh->CreateFromFile(argv[1]);
    Mesh->DisplayMeshProperties();
    if (argc > 2) NumberOfSamples = atoi(argv[2]);
    else { cout << "Number of vertices ? "; cin >> NumberOfSamples; }
    if (argc > 3) Gradation = atof(argv[3]);
    else { cout << "Gradation ? "; cin >> Gradation; }
    cout << argc << " Arguments" << endl;
    for (int i = 4; i + 1 < argc; i += 2) {
        if (strcmp(argv[i], "-s") == 0) { SubsamplingThreshold = atoi(argv[i + 1]); cout << "Subsampling Threshold=" << SubsamplingThreshold << endl; }
        if (strcmp(argv[i], "-d") == 0) { Display = atoi(argv[i + 1]); cout << "Display=" << Display << endl; }
        if (strcmp(argv[i], "-np") == 0) cout << "Number of threads=" << atoi(argv[i + 1]) << " (ignored: GPU engine)" << endl;
        if (strcmp(argv[i], "-o") == 0) { OutputDirectory = argv[i + 1]; cout << "OutputDirectory: " << OutputDirectory << endl; }
        if (strcmp(argv[i], "-l") == 0) {
            Mesh->SplitLongEdges(atof(argv[i + 1]));
            cout << "Splitting edges longer than " << atof(argv[i + 1]) << " times the average edge length" << endl;
        }
        if (strcmp(argv[i], "-q") == 0) {
            cout << "Setting number of eigenvalues for quadrics to " << atoi(argv[i + 1]) << endl;
            Remesh->GetMetric()->SetQuadricsOptimizationLevel(atoi(argv[i + 1]));
        }
        if (strcmp(argv[i], "-b") == 0) { cout << "Setting boundary fixing to : " << argv[i + 1] << endl; Remesh->SetBoundaryFixing(atoi(argv[i + 1])); }
        if (strcmp(argv[i], "-dev") == 0) Remesh->SetDevice(atoi(argv[i + 1]));
    }
    Remesh->SetInput(Mesh);
    Remesh->SetNumberOfClusters(NumberOfSamples);
    Remesh->SetConsoleOutput(2);
    Remesh->SetSubsamplingThreshold(SubsamplingThreshold);
    Remesh->GetMetric()->SetGradation(Gradation);
    Remesh->SetDisplay(Display);
    Remesh->Remesh();
    if (!Remesh->GetOutput()) return 1;
    std::string real = OutputDirectory ? std::string(OutputDirectory) + "Remeshing.ply" : "Remeshing.ply";
    if (OutputDirectory) cout << "OutputDirectory: " << OutputDirectory << endl;
    Remesh->GetOutput()->WriteToFile(real.c_str());
    Remesh->Delete();
    Mesh->Delete();
    return 0;
}
