// ACVDQ front-end on the B200 engine: same command line as the reference's
// DiscreteRemeshing/Examples/ACVDQ.cxx, output simplification.ply (binary PLY).
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>

#include "../vtkDiscreteRemeshing.h"

using std::cin;
using std::cout;
using std::endl;

int main(int argc, char* argv[]) {
    int display = 0, numberOfSamples = 500, subsamplingThreshold = 10;
    double gradation = 0;
    char* outputDirectory = 0;
    const char* outputfile = "simplification.ply";
    vtkIdList* fixedVertices = 0;
    if (argc <= 1) {
        cout << "Usage : ACVDQ file nvertices gradation [options]" << endl;
        cout << "nvertices is the desired number of vertices" << endl;
        cout << "gradation defines the influence of local curvature (0=uniform meshing)" << endl;
        cout << endl << "Optionnal arguments : " << endl;
        cout << "-b 0/1 : sets mesh boundary fixing off/on (default : 0)" << endl;
        cout << "-s threshold : defines the subsampling threshold i.e. the input mesh will be subdivided until its number ";
        cout << " of vertices is above nvertices*threshold (default=10)" << endl;
        cout << "-d 0/1/2 : enables display (accepted and ignored by this build)" << endl;
        cout << "-l ratio : split the edges longer than ( averageLength * ratio )" << endl;
        cout << "-q 1/2/3 : qets number of eigenvalues used for quadric-based vertex relocation to 0/1/2 (default : 3)" << endl;
        cout << "-m 0/1 : enforce a manifold output ON/OFF (default : 0)" << endl;
        cout << "-fv file / -ft file : fixed vertices / triangles" << endl;
        cout << "-dev n : CUDA device index (this build)" << endl;
        return 0;
    }
    cout << "load : " << argv[1] << endl;
    vtkSurface* mesh = vtkSurface::New();
    vtkQIsotropicDiscreteRemeshing* remesh = vtkQIsotropicDiscreteRemeshing::New();
    mesh->CreateFromFile(argv[1]);
    mesh->DisplayMeshProperties();
    if (argc > 2) numberOfSamples = atoi(argv[2]);
    else { cout << "Number of vertices ? "; cin >> numberOfSamples; }
    if (argc > 3) gradation = atof(argv[3]);
    else { cout << "Gradation ? "; cin >> gradation; }
    for (int i = 4; i + 1 < argc; i += 2) {
        char* key = argv[i];
        char* value = argv[i + 1];
        if (strcmp(key, "-m") == 0) { remesh->SetForceManifold(atoi(value)); cout << "Force Manifold=" << atoi(value) << endl; }
        else if (strcmp(key, "-s") == 0) { subsamplingThreshold = atoi(value); cout << "Subsampling Threshold=" << subsamplingThreshold << endl; }
        else if (strcmp(key, "-d") == 0) { display = atoi(value); cout << "Display=" << display << endl; }
        if (strcmp(key, "-np") == 0 || strcmp(key, "-p") == 0) cout << key << " " << value << " ignored (GPU engine)" << endl;
        if (strcmp(key, "-o") == 0) { outputDirectory = value; cout << "OutputDirectory: " << outputDirectory << endl; remesh->SetOutputDirectory(value); }
        else if (strcmp(key, "-of") == 0) { outputfile = value; cout << "Output file name: " << outputfile << endl; }
        else if (strcmp(key, "-l") == 0) {
            cout << "Splitting edges longer than " << atof(value) << " times the average edge length" << endl;
            mesh->SplitLongEdges(atof(value));
        } else if (strcmp(key, "-w") == 0) { cout << "Setting writing energy log file to " << atoi(value) << endl; remesh->SetWriteToGlobalEnergyLog(atoi(value)); }
        if (strcmp(key, "-q") == 0) { cout << "Setting number of eigenvalues for quadrics to " << atoi(value) << endl; remesh->GetMetric()->SetQuadricsOptimizationLevel(atoi(value)); }
        else if (strcmp(key, "-cd") == 0) remesh->SetInputDensityFile(value);
        else if (strcmp(key, "-cmax") == 0) remesh->SetMaxCustomDensity(atof(value));
        else if (strcmp(key, "-cmin") == 0) remesh->SetMinCustomDensity(atof(value));
        if (strcmp(key, "-cf") == 0) remesh->SetCustomDensityMultiplicationFactor(atof(value));
        else if (strcmp(key, "-b") == 0) { cout << "Setting boundary fixing to : " << value << endl; remesh->SetBoundaryFixing(atoi(value)); }
        else if (strcmp(key, "-fv") == 0) {
            std::ifstream input(value);
            int id;
            fixedVertices = vtkIdList::New();
            while (input >> id) fixedVertices->InsertNextId(id);
        } else if (strcmp(key, "-ft") == 0) {
            std::ifstream input(value);
            std::vector<char> fixed((size_t)mesh->GetNumberOfPoints(), 0);
            fixedVertices = vtkIdList::New();
            int id, n = 0;
            while (input >> id) {
                n++;
                vtkIdType v1, v2, v3;
                mesh->GetFaceVertices(id, v1, v2, v3);
                fixed[(size_t)v1] = fixed[(size_t)v2] = fixed[(size_t)v3] = 1;
            }
            for (vtkIdType i = 0; i < mesh->GetNumberOfPoints(); i++) if (fixed[(size_t)i]) fixedVertices->InsertNextId(i);
            cout << "Added " << n << " constraints on triangles" << endl;
        }
        if (strcmp(key, "-dev") == 0) remesh->SetDevice(atoi(value));
    }
    remesh->SetInput(mesh);
    remesh->SetFileLoadSaveOption(0);
    remesh->SetConsoleOutput(2);
    remesh->SetSubsamplingThreshold(subsamplingThreshold);
    remesh->GetMetric()->SetGradation(gradation);
    remesh->SetDisplay(display);
    remesh->SetUnconstrainedInitialization(1);
    if (fixedVertices) {
        remesh->SetFixedClusters(fixedVertices);
        remesh->SetNumberOfClusters(numberOfSamples + (int)fixedVertices->GetNumberOfIds());
        cout << "Read " << fixedVertices->GetNumberOfIds() << " fixed Ids" << endl;
        for (vtkIdType i = 0; i < fixedVertices->GetNumberOfIds(); i++) remesh->GetCluster((int)i)->AnchorItem = fixedVertices->GetId(i);
    } else remesh->SetNumberOfClusters(numberOfSamples);
    remesh->Remesh();
    if (!remesh->GetOutput()) return 1;
    if (fixedVertices) {   // the anchored output vertices must sit exactly on their input vertices (ACVDQ.cxx:341-366)
        vtkSurface* mesh2 = remesh->GetOutput();
        for (vtkIdType i = 0; i < fixedVertices->GetNumberOfIds(); i++) {
            double c1[3], c2[3];
            const vtkIdType v = fixedVertices->GetId(i);
            remesh->GetInput()->GetPointCoordinates(v, c1);
            mesh2->GetPointCoordinates(i, c2);
            for (int j = 0; j < 3; j++) {
                if (c1[j] == c2[j]) continue;
                cout << "Error, vertex " << v << " has been lost" << endl;
                exit(1);
            }
        }
        cout << "Constraints on vertices have been checked" << endl;
        fixedVertices->Delete();
    }
    std::string realFile;
    if (outputDirectory) realFile += outputDirectory;
    realFile += outputfile;
    remesh->GetOutput()->WriteToFile(realFile.c_str());
    remesh->Delete();
    mesh->Delete();
    return 0;
}
