// VTK-free mesh container for the ACVD front-ends (see vtkSurface.h).
#include "vtkSurface.h"

#include "../../../include/acvd_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <unordered_map>

namespace {

std::string lower_ext(const std::string& path) {
    size_t p = path.find_last_of('.');
    std::string e = p == std::string::npos ? "" : path.substr(p + 1);
    for (auto& c : e) c = (char)tolower(c);
    return e;
}

inline uint64_t edge_key(int a, int b) {
    uint32_t lo = (uint32_t)std::min(a, b), hi = (uint32_t)std::max(a, b);
    return ((uint64_t)lo << 32) | hi;
}

struct PlyProp { std::string type, name; bool is_list = false; std::string count_type, item_type; };
struct PlyElem { std::string name; long long count = 0; std::vector<PlyProp> props; };

size_t ply_size(const std::string& t) {
    if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
    if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
    if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32") return 4;
    if (t == "double" || t == "float64" || t == "int64" || t == "uint64") return 8;
    return 0;
}

double ply_read_bin(std::istream& in, const std::string& t) {
    char buf[8] = {0};
    in.read(buf, (std::streamsize)ply_size(t));
    if (t == "char" || t == "int8") return (double)*(int8_t*)buf;
    if (t == "uchar" || t == "uint8") return (double)*(uint8_t*)buf;
    if (t == "short" || t == "int16") return (double)*(int16_t*)buf;
    if (t == "ushort" || t == "uint16") return (double)*(uint16_t*)buf;
    if (t == "int" || t == "int32") return (double)*(int32_t*)buf;
    if (t == "uint" || t == "uint32") return (double)*(uint32_t*)buf;
    if (t == "float" || t == "float32") return (double)*(float*)buf;
    if (t == "double" || t == "float64") return *(double*)buf;
    if (t == "int64") return (double)*(int64_t*)buf;
    return (double)*(uint64_t*)buf;
}

}  // namespace

// ------------------------------------------------------------------------------------------ IO
void vtkSurface::CreateFromArrays(vtkIdType nv, const float* p, vtkIdType nf, const int* t) {
    xyz.assign(p, p + 3 * nv);
    tri.assign(t, t + 3 * nf);
    extra_edges.clear();
    Invalidate();
}

void vtkSurface::CreateFromFile(const char* path) {
    xyz.clear(); tri.clear(); extra_edges.clear(); Invalidate();
    const std::string ext = lower_ext(path);
    std::ifstream in(path, std::ios::binary);
    if (!in) { std::cout << "ERROR : could not open " << path << std::endl; exit(1); }
    auto add_poly = [&](const std::vector<int>& idx) {
        for (size_t k = 1; k + 1 < idx.size(); k++) { tri.push_back(idx[0]); tri.push_back(idx[k]); tri.push_back(idx[k + 1]); }
    };
    if (ext == "ply") {
        std::string line, fmt;
        std::vector<PlyElem> elems;
        std::getline(in, line);
        while (std::getline(in, line)) {
            if (!line.empty() && line.back() == '\r') line.pop_back();
            std::istringstream ls(line);
            std::string tok; ls >> tok;
            if (tok == "format") ls >> fmt;
            else if (tok == "element") { PlyElem e; ls >> e.name >> e.count; elems.push_back(e); }
            else if (tok == "property") {
                PlyProp p; ls >> p.type;
                if (p.type == "list") { p.is_list = true; ls >> p.count_type >> p.item_type >> p.name; }
                else ls >> p.name;
                if (!elems.empty()) elems.back().props.push_back(p);
            } else if (tok == "end_header") break;
        }
        const bool ascii = fmt == "ascii";
        if (!ascii && fmt != "binary_little_endian") { std::cout << "ERROR : unsupported PLY format " << fmt << std::endl; exit(1); }
        for (auto& e : elems) {
            for (long long i = 0; i < e.count; i++) {
                double vx = 0, vy = 0, vz = 0;
                std::vector<int> idx;
                for (auto& p : e.props) {
                    if (p.is_list) {
                        long long n;
                        if (ascii) in >> n; else n = (long long)ply_read_bin(in, p.count_type);
                        std::vector<int> tmp((size_t)n);
                        for (long long k = 0; k < n; k++) { double d; if (ascii) in >> d; else d = ply_read_bin(in, p.item_type); tmp[(size_t)k] = (int)d; }
                        if (e.name == "face" && (p.name == "vertex_indices" || p.name == "vertex_index")) idx = tmp;
                    } else {
                        double d; if (ascii) in >> d; else d = ply_read_bin(in, p.type);
                        if (p.name == "x") vx = d; else if (p.name == "y") vy = d; else if (p.name == "z") vz = d;
                    }
                }
                if (e.name == "vertex") { xyz.push_back((float)vx); xyz.push_back((float)vy); xyz.push_back((float)vz); }
                else if (e.name == "face") add_poly(idx);
            }
        }
    } else if (ext == "obj") {
        std::string line;
        while (std::getline(in, line)) {
            std::istringstream ls(line);
            std::string tok; ls >> tok;
            if (tok == "v") { double x, y, z; ls >> x >> y >> z; xyz.push_back((float)x); xyz.push_back((float)y); xyz.push_back((float)z); }
            else if (tok == "f") {
                std::vector<int> idx; std::string w;
                const int nv = (int)(xyz.size() / 3);
                while (ls >> w) { int id = atoi(w.c_str()); idx.push_back(id > 0 ? id - 1 : nv + id); }
                add_poly(idx);
            }
        }
    } else if (ext == "off") {
        std::string hdr; in >> hdr;
        long long nv, nf, ne; in >> nv >> nf >> ne;
        for (long long i = 0; i < nv; i++) { double x, y, z; in >> x >> y >> z; xyz.push_back((float)x); xyz.push_back((float)y); xyz.push_back((float)z); }
        for (long long i = 0; i < nf; i++) { int n; in >> n; std::vector<int> idx((size_t)n); for (auto& k : idx) in >> k; add_poly(idx); }
    } else {
        std::cout << "ERROR : unsupported mesh format ." << ext << " (this VTK-free build reads .ply .obj .off)" << std::endl;
        exit(1);
    }
}

void vtkSurface::WriteToFile(const char* path) {
    const std::string ext = lower_ext(path);
    const long long nv = GetNumberOfPoints(), nf = GetNumberOfCells();
    if (ext == "obj") {
        std::ofstream o(path);
        o.precision(9);
        for (long long i = 0; i < nv; i++) o << "v " << xyz[3 * i] << " " << xyz[3 * i + 1] << " " << xyz[3 * i + 2] << "\n";
        for (long long i = 0; i < nf; i++) o << "f " << tri[3 * i] + 1 << " " << tri[3 * i + 1] + 1 << " " << tri[3 * i + 2] + 1 << "\n";
        return;
    }
    if (ext == "off") {
        std::ofstream o(path);
        o.precision(9);
        o << "OFF\n" << nv << " " << nf << " 0\n";
        for (long long i = 0; i < nv; i++) o << xyz[3 * i] << " " << xyz[3 * i + 1] << " " << xyz[3 * i + 2] << "\n";
        for (long long i = 0; i < nf; i++) o << "3 " << tri[3 * i] << " " << tri[3 * i + 1] << " " << tri[3 * i + 2] << "\n";
        return;
    }
    // binary little-endian PLY: float32 xyz, uchar + int32 face lists (vtkPLYWriter's default layout)
    std::ofstream o(path, std::ios::binary);
    o << "ply\nformat binary_little_endian 1.0\ncomment acvd_b200\nelement vertex " << nv
      << "\nproperty float x\nproperty float y\nproperty float z\nelement face " << nf
      << "\nproperty list uchar int vertex_indices\nend_header\n";
    o.write((const char*)xyz.data(), (std::streamsize)(xyz.size() * sizeof(float)));
    for (long long i = 0; i < nf; i++) {
        unsigned char n = 3;
        o.write((const char*)&n, 1);
        o.write((const char*)&tri[3 * i], 3 * sizeof(int));
    }
}

// ------------------------------------------------------------------------------------------ construction
vtkIdType vtkSurface::AddVertex(double x, double y, double z) {
    xyz.push_back((float)x); xyz.push_back((float)y); xyz.push_back((float)z);
    Invalidate();
    return GetNumberOfPoints() - 1;
}
vtkIdType vtkSurface::AddFace(vtkIdType a, vtkIdType b, vtkIdType c) {
    tri.push_back((int)a); tri.push_back((int)b); tri.push_back((int)c);
    Invalidate();
    return GetNumberOfCells() - 1;
}
vtkIdType vtkSurface::IsFace(vtkIdType a, vtkIdType b, vtkIdType c) {
    BuildTopology();
    for (int i = vf_ptr[a]; i < vf_ptr[a + 1]; i++) {
        const int* t = &tri[3 * (size_t)vf[i]];
        int hit = 0;
        for (int k = 0; k < 3; k++) hit += (t[k] == a) + (t[k] == b) + (t[k] == c);
        if (hit == 3 && a != b && b != c && a != c) return vf[i];
    }
    return -1;
}
vtkIdType vtkSurface::IsEdge(vtkIdType a, vtkIdType b) {
    BuildTopology();
    for (int i = ve_ptr[a]; i < ve_ptr[a + 1]; i++) {
        const auto& e = edges[(size_t)ve[i]];
        if ((e[0] == a && e[1] == b) || (e[0] == b && e[1] == a)) return ve[i];
    }
    return -1;
}
vtkIdType vtkSurface::AddEdge(vtkIdType a, vtkIdType b) {
    extra_edges.push_back({(int)std::min(a, b), (int)std::max(a, b)});
    Invalidate();
    return -1;
}

// Edges are numbered in first-seen order over the faces, as the reference's AddEdge does
// (Common/vtkSurfaceBase.cxx:1166-1221); Subdivide relies on that order for the midpoint ids.
void vtkSurface::BuildTopology() {
    if (topo_valid) return;
    const int nv = (int)GetNumberOfPoints(), nf = (int)GetNumberOfCells();
    vf_ptr.assign((size_t)nv + 1, 0);
    for (int f = 0; f < nf; f++) for (int k = 0; k < 3; k++) vf_ptr[(size_t)tri[3 * (size_t)f + k] + 1]++;
    for (int v = 0; v < nv; v++) vf_ptr[v + 1] += vf_ptr[v];
    vf.assign((size_t)vf_ptr[nv], 0);
    {
        std::vector<int> cur(vf_ptr.begin(), vf_ptr.end() - 1);
        for (int f = 0; f < nf; f++) for (int k = 0; k < 3; k++) vf[(size_t)cur[tri[3 * (size_t)f + k]]++] = f;
    }
    edges.clear(); edge_nfaces.clear();
    std::unordered_map<uint64_t, int> idx;
    idx.reserve((size_t)nf * 2);
    auto touch = [&](int a, int b, int faces) {
        if (a == b) return;
        auto it = idx.find(edge_key(a, b));
        if (it == idx.end()) { idx.emplace(edge_key(a, b), (int)edges.size()); edges.push_back({a, b}); edge_nfaces.push_back(faces); }
        else edge_nfaces[(size_t)it->second] += faces;
    };
    for (int f = 0; f < nf; f++) {
        const int* t = &tri[3 * (size_t)f];
        if (t[0] == t[1]) continue;
        for (int k = 0; k < 3; k++) touch(t[k], t[(k + 1) % 3], 1);
    }
    for (auto& e : extra_edges) touch(e[0], e[1], 0);
    ve_ptr.assign((size_t)nv + 1, 0);
    for (auto& e : edges) { ve_ptr[(size_t)e[0] + 1]++; ve_ptr[(size_t)e[1] + 1]++; }
    for (int v = 0; v < nv; v++) ve_ptr[v + 1] += ve_ptr[v];
    ve.assign((size_t)ve_ptr[nv], 0);
    {
        std::vector<int> cur(ve_ptr.begin(), ve_ptr.end() - 1);
        for (int e = 0; e < (int)edges.size(); e++) { ve[(size_t)cur[edges[e][0]]++] = e; ve[(size_t)cur[edges[e][1]]++] = e; }
    }
    topo_valid = true;
}

vtkIdType vtkSurface::GetNumberOfEdges() { BuildTopology(); return (vtkIdType)edges.size(); }

void vtkSurface::GetVertexNeighbourFaces(vtkIdType v, vtkIdList* out) {
    BuildTopology();
    out->Reset();
    for (int i = vf_ptr[v]; i < vf_ptr[v + 1]; i++) out->InsertNextId(vf[i]);
}
void vtkSurface::GetVertexNeighbours(vtkIdType v, vtkIdList* out) {
    BuildTopology();
    out->Reset();
    for (int i = ve_ptr[v]; i < ve_ptr[v + 1]; i++) { const auto& e = edges[(size_t)ve[i]]; out->InsertNextId(e[0] == v ? e[1] : e[0]); }
}

// A vertex is manifold when it has at least two edges, every one of them carries exactly two faces (the reference's
// IsEdgeManifold, vtkSurfaceBase.h:521-528, refuses boundary edges too) and its incident faces form one closed fan that
// reaches every incident edge (same predicate as the fan walk at reference Common/vtkSurfaceBase.cxx:259-317).
bool vtkSurface::IsVertexManifold(vtkIdType v) {
    BuildTopology();
    const int ne = ve_ptr[v + 1] - ve_ptr[v];
    if (ne < 2) return false;
    std::vector<int> nb((size_t)ne);
    for (int i = 0; i < ne; i++) {
        const int e = ve[(size_t)ve_ptr[v] + i];
        if (edge_nfaces[(size_t)e] != 2) return false;
        nb[(size_t)i] = edges[(size_t)e][0] == v ? edges[(size_t)e][1] : edges[(size_t)e][0];
    }
    // link graph: neighbours a, b are joined when the face (v, a, b) exists
    std::vector<std::vector<int>> adj((size_t)ne);
    auto slot = [&](int u) { for (int i = 0; i < ne; i++) if (nb[(size_t)i] == u) return i; return -1; };
    for (int i = vf_ptr[v]; i < vf_ptr[v + 1]; i++) {
        const int* t = &tri[3 * (size_t)vf[i]];
        int o[2], n = 0;
        for (int k = 0; k < 3; k++) if (t[k] != v && n < 2) o[n++] = t[k];
        if (n != 2) continue;
        int a = slot(o[0]), b = slot(o[1]);
        if (a < 0 || b < 0) continue;
        adj[(size_t)a].push_back(b); adj[(size_t)b].push_back(a);
    }
    std::vector<char> seen((size_t)ne, 0);
    std::vector<int> stack{0};
    seen[0] = 1;
    int reached = 1;
    while (!stack.empty()) {
        int a = stack.back(); stack.pop_back();
        for (int b : adj[(size_t)a]) if (!seen[(size_t)b]) { seen[(size_t)b] = 1; reached++; stack.push_back(b); }
    }
    return reached == ne;
}

double vtkSurface::GetFaceArea(vtkIdType f) const {
    const int* t = &tri[3 * (size_t)f];
    double a[3], b[3], c[3];
    GetPoint(t[0], a); GetPoint(t[1], b); GetPoint(t[2], c);
    double u[3] = {c[0] - b[0], c[1] - b[1], c[2] - b[2]}, w[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
    double n[3] = {u[1] * w[2] - u[2] * w[1], u[2] * w[0] - u[0] * w[2], u[0] * w[1] - u[1] * w[0]};
    return 0.5 * std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
}

void vtkSurface::GetBounds(double b[6]) const {
    b[0] = b[2] = b[4] = 1e300; b[1] = b[3] = b[5] = -1e300;
    for (size_t i = 0; i + 2 < xyz.size(); i += 3)
        for (int k = 0; k < 3; k++) { b[2 * k] = std::min(b[2 * k], (double)xyz[i + k]); b[2 * k + 1] = std::max(b[2 * k + 1], (double)xyz[i + k]); }
}

// 1 -> 4 subdivision: old points first, then one midpoint per edge in edge-id order, then per old face
// (V1,V4,V6) (V4,V2,V5) (V5,V3,V6) (V4,V5,V6)  (reference Common/vtkSurface.cxx:605-677)
vtkSurface* vtkSurface::Subdivide(vtkIntArray* parent1, vtkIntArray* parent2) {
    BuildTopology();
    const int nv = (int)GetNumberOfPoints(), nf = (int)GetNumberOfCells(), ne = (int)edges.size();
    vtkSurface* out = vtkSurface::New();
    out->xyz.reserve(3 * ((size_t)nv + ne));
    out->xyz = xyz;
    if (parent1) { parent1->v.resize((size_t)nv + ne); parent2->v.resize((size_t)nv + ne); }
    for (int e = 0; e < ne; e++) {
        const int a = edges[(size_t)e][0], b = edges[(size_t)e][1];
        for (int k = 0; k < 3; k++) out->xyz.push_back((float)(0.5 * ((double)xyz[3 * (size_t)a + k] + (double)xyz[3 * (size_t)b + k])));
        if (parent1) { parent1->v[(size_t)nv + e] = a; parent2->v[(size_t)nv + e] = b; }
    }
    std::unordered_map<uint64_t, int> idx;
    idx.reserve((size_t)ne * 2);
    for (int e = 0; e < ne; e++) idx.emplace(edge_key(edges[(size_t)e][0], edges[(size_t)e][1]), e);
    out->tri.reserve(12 * (size_t)nf);
    for (int f = 0; f < nf; f++) {
        const int* t = &tri[3 * (size_t)f];
        if (t[0] == t[1] || t[1] == t[2] || t[0] == t[2]) continue;
        const int v4 = nv + idx[edge_key(t[0], t[1])], v5 = nv + idx[edge_key(t[1], t[2])], v6 = nv + idx[edge_key(t[2], t[0])];
        const int q[12] = {t[0], v4, v6, v4, t[1], v5, v5, t[2], v6, v4, v5, v6};
        out->tri.insert(out->tri.end(), q, q + 12);
    }
    return out;
}

// vtkSurface::SplitLongEdges (reference Common/vtkSurface.cxx:444-604): threshold = ratio x mean edge length of the mesh
// as given, then passes of "cut every edge above it at its midpoint, replace the triangles by the Split2 / Split3 / 1 -> 4
// patterns".  Runs on the device through the C ABI (acvd_split_long_edges); new points follow the old ones.
void vtkSurface::SplitLongEdges(double ratio) {
    acvd_ctx* ctx = nullptr;
    if (acvd_create(&ctx, -1) != ACVD_OK) { std::cout << "ERROR : " << acvd_last_error(nullptr) << std::endl; return; }
    int32_t nv = 0, nf = 0, passes = 0;
    if (acvd_set_mesh(ctx, (int32_t)GetNumberOfPoints(), (int32_t)GetNumberOfCells(), Points(), Triangles()) != ACVD_OK ||
        acvd_split_long_edges(ctx, ratio, &nv, &nf, &passes) != ACVD_OK) {
        std::cout << "ERROR : " << acvd_last_error(ctx) << std::endl;
        acvd_destroy(ctx);
        return;
    }
    std::vector<float> p(3 * (size_t)nv);
    std::vector<int> t(3 * (size_t)nf);
    if (acvd_get_subdivision(ctx, p.data(), t.data(), nullptr, nullptr) == ACVD_OK) { xyz.swap(p); tri.swap(t); Invalidate(); }
    else std::cout << "ERROR : " << acvd_last_error(ctx) << std::endl;
    acvd_destroy(ctx);
}

void vtkSurface::GetMeshProperties(vtkIdType& non_manifold, vtkIdType& boundary, vtkIdType& components) {
    BuildTopology();
    non_manifold = boundary = 0;
    for (int n : edge_nfaces) { if (n > 2) non_manifold++; if (n == 1) boundary++; }
    const int nv = (int)GetNumberOfPoints();
    std::vector<int> label((size_t)nv, -1);
    components = 0;
    std::vector<int> stack;
    for (int s = 0; s < nv; s++) {
        if (label[(size_t)s] >= 0 || ve_ptr[s + 1] == ve_ptr[s]) continue;
        label[(size_t)s] = (int)components; stack.push_back(s);
        while (!stack.empty()) {
            int v = stack.back(); stack.pop_back();
            for (int i = ve_ptr[v]; i < ve_ptr[v + 1]; i++) {
                const auto& e = edges[(size_t)ve[i]];
                int u = e[0] == v ? e[1] : e[0];
                if (label[(size_t)u] < 0) { label[(size_t)u] = (int)components; stack.push_back(u); }
            }
        }
        components++;
    }
}

void vtkSurface::DisplayMeshProperties() {
    vtkIdType nm, bd, cc;
    GetMeshProperties(nm, bd, cc);
    std::cout << "*****************************************************************************" << std::endl;
    std::cout << "Mesh with " << GetNumberOfCells() << " polygons, " << GetNumberOfPoints() << " points, "
              << GetNumberOfEdges() << " edges" << std::endl;
    double b[6]; GetBounds(b);
    std::cout << "Bounding Box: [" << b[0] << ", " << b[2] << ", " << b[4] << "]  [" << b[1] << ", " << b[3] << ", " << b[5] << "]" << std::endl;
    std::cout << nm << " non-manifold edges, " << bd << " boundary edges, " << cc << " connected components" << std::endl;
    std::cout << "*****************************************************************************" << std::endl;
}
