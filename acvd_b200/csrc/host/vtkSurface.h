// VTK-free stand-in for the parts of vtkSurface (reference Common/vtkSurface.{h,cxx},
// Common/vtkSurfaceBase.{h,cxx}) that the ACVD front-ends touch: points (float32, as a default
// vtkPoints), triangles, reference counting, file IO (PLY / OBJ / OFF in, binary PLY out), 1->4
// subdivision, mesh statistics, vertex manifoldness.  The clustering itself never walks this
// structure: it goes through the C ABI (include/acvd_b200.h) onto the GPU.
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <vector>

typedef long long vtkIdType;

// minimal id list with the few vtkIdList calls the front-ends use
class vtkIdList {
public:
    static vtkIdList* New() { return new vtkIdList; }
    void Delete() { delete this; }
    vtkIdType GetNumberOfIds() const { return (vtkIdType)ids.size(); }
    vtkIdType GetId(vtkIdType i) const { return ids[(size_t)i]; }
    void InsertNextId(vtkIdType v) { ids.push_back(v); }
    void Reset() { ids.clear(); }
    std::vector<vtkIdType> ids;
};

class vtkIntArray {
public:
    static vtkIntArray* New() { return new vtkIntArray; }
    void Delete() { delete this; }
    void SetNumberOfValues(vtkIdType n) { v.assign((size_t)n, 0); }
    vtkIdType GetNumberOfTuples() const { return (vtkIdType)v.size(); }
    int GetValue(vtkIdType i) const { return v[(size_t)i]; }
    void SetValue(vtkIdType i, int x) { v[(size_t)i] = x; }
    int* GetPointer(vtkIdType i) { return v.data() + i; }
    std::vector<int> v;
};

class vtkSurface {
public:
    static vtkSurface* New() { return new vtkSurface; }
    void Register(void* = nullptr) { refs++; }
    void UnRegister(void* = nullptr) { if (--refs <= 0) delete this; }
    void Delete() { UnRegister(); }

    // ---- construction
    void CreateFromFile(const char* path);                     // vtkSurface.cxx:2017 (by extension: ply, obj, off)
    void CreateFromArrays(vtkIdType nv, const float* xyz, vtkIdType nf, const int* tri);
    vtkIdType AddVertex(double x, double y, double z);
    vtkIdType AddVertex(const double* p) { return AddVertex(p[0], p[1], p[2]); }
    vtkIdType AddFace(vtkIdType v1, vtkIdType v2, vtkIdType v3);
    vtkIdType IsFace(vtkIdType v1, vtkIdType v2, vtkIdType v3);  // face with this vertex set, or -1
    vtkIdType AddEdge(vtkIdType v1, vtkIdType v2);             // face-less edge (ForceManifold dual edges)
    vtkIdType IsEdge(vtkIdType v1, vtkIdType v2);

    // ---- queries
    vtkIdType GetNumberOfPoints() const { return (vtkIdType)(xyz.size() / 3); }
    vtkIdType GetNumberOfCells() const { return (vtkIdType)(tri.size() / 3); }
    vtkIdType GetNumberOfEdges();
    void GetPoint(vtkIdType v, double* p) const { p[0] = xyz[3 * v]; p[1] = xyz[3 * v + 1]; p[2] = xyz[3 * v + 2]; }
    void GetPointCoordinates(vtkIdType v, double* p) const { GetPoint(v, p); }
    void SetPointCoordinates(vtkIdType v, const double* p) { xyz[3 * v] = (float)p[0]; xyz[3 * v + 1] = (float)p[1]; xyz[3 * v + 2] = (float)p[2]; }
    void GetFaceVertices(vtkIdType f, vtkIdType& a, vtkIdType& b, vtkIdType& c) const { a = tri[3 * f]; b = tri[3 * f + 1]; c = tri[3 * f + 2]; }
    void GetVertexNeighbourFaces(vtkIdType v, vtkIdList* out);
    void GetVertexNeighbours(vtkIdType v, vtkIdList* out);
    bool IsVertexManifold(vtkIdType v);                        // vtkSurfaceBase.cxx:259-317
    double GetFaceArea(vtkIdType f) const;
    void GetBounds(double b[6]) const;

    // ---- processing
    vtkSurface* Subdivide(vtkIntArray* parent1 = nullptr, vtkIntArray* parent2 = nullptr);   // vtkSurface.cxx:605-677
    void SplitLongEdges(double ratio);                         // vtkSurface.cxx:444-604, on the device (acvd_split_long_edges)
    void DisplayMeshProperties();                              // vtkSurface.cxx:762-873 (subset)
    void GetMeshProperties(vtkIdType& nonManifoldEdges, vtkIdType& boundaryEdges, vtkIdType& components);

    // ---- output
    void WriteToFile(const char* path);                        // binary little-endian PLY (or .obj / .off by extension)

    const float* Points() const { return xyz.data(); }
    const int* Triangles() const { return tri.data(); }

    std::vector<float> xyz;
    std::vector<int> tri;

private:
    vtkSurface() {}
    ~vtkSurface() {}
    int refs = 1;
    // lazily built topology
    void BuildTopology();
    void Invalidate() { topo_valid = false; }
    bool topo_valid = false;
    std::vector<int> vf_ptr, vf;                     // vertex -> faces
    std::vector<std::array<int, 2>> edges;           // undirected, (lo, hi)
    std::vector<int> edge_nfaces;                    // faces per edge
    std::vector<int> ve_ptr, ve;                     // vertex -> edges
    std::vector<std::array<int, 2>> extra_edges;     // face-less edges added with AddEdge
};
