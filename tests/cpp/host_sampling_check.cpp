// CPU check of the product-side initial sampling (acvd_b200/csrc/host_sampling.hpp): the flat-ring version the library runs
// (prefetching, FIFO without already-assigned entries) must give the clustering of the plain restatement of
// ComputeInitialRandomSampling (Common/vtkUniformClustering.h:1178-1316) on the same rings and weights.
#include <cstdio>
#include <cstdlib>
#include "host_sampling.hpp"
using namespace acvd;

static void grid_mesh(int nx, int ny, std::vector<int>& tri) {      // open nx x ny grid, two triangles per cell
    tri.clear();
    for (int y = 0; y + 1 < ny; y++)
        for (int x = 0; x + 1 < nx; x++) {
            const int a = y * nx + x, b = a + 1, c = a + nx, d = c + 1;
            tri.insert(tri.end(), {a, b, d, a, d, c});
        }
}

int main() {
    int bad = 0;
    const int cases[][3] = {{40, 30, 12}, {64, 64, 200}, {17, 23, 391}, {50, 50, 2400}, {9, 9, 1}};   // nx, ny, K (incl. K close to V: the steal pass)
    for (const auto& cs : cases) {
        const int nx = cs[0], ny = cs[1], K = cs[2], V = nx * ny;
        std::vector<int> tri;
        grid_mesh(nx, ny, tri);
        HostRings H;
        H.build(V, (int)tri.size() / 3, tri.data());
        std::vector<int> ptr(V + 1, 0), nbr;
        for (int v = 0; v < V; v++) {
            ptr[v + 1] = ptr[v] + H.len[v];
            for (int k = 0; k < H.len[v]; k++) nbr.push_back(H.nbr[H.ptr[v] + k]);
        }
        std::vector<double> w(V);
        for (int v = 0; v < V; v++) w[v] = 1.0 + 0.75 * (((unsigned)v * 2654435761u >> 20) & 0xff) / 255.0;
        for (int with_fixed = 0; with_fixed < 2; with_fixed++) {
            std::vector<int64_t> fixed;
            if (with_fixed && K >= 3) fixed = {0, (int64_t)V / 2, (int64_t)V - 1};
            std::vector<int> a, b;
            initial_random_sampling(V, K, H, w.data(), fixed, a);
            initial_random_sampling(V, K, FlatRings{ptr.data(), nbr.data()}, w.data(), fixed, b);
            if (a != b) { printf("MISMATCH nx=%d ny=%d K=%d fixed=%d\n", nx, ny, K, with_fixed); bad++; }
        }
    }
    printf(bad ? "FAIL\n" : "OK\n");
    return bad ? 1 : 0;
}
