// Host-side initial sampling for the product path.
//
// ComputeInitialRandomSampling (reference Common/vtkUniformClustering.h:1178-1316) is sequential by
// construction: a std::mt19937(0)-driven pseudo-shuffle followed by weight-bounded breadth-first region
// growing in the mesh's ring order.  SURVEY §8e keeps it on the host ("run once on host and scatter").
// The ring order it walks is the reference's: edges are numbered in first-seen order over the faces and
// appended to the rings of both endpoints (Common/vtkSurfaceBase.cxx:1166-1221, 1057-1068).
#pragma once
#include <cstdint>
#include <queue>
#include <random>
#include <vector>

namespace acvd {

// Vertex rings in edge-creation order, flattened: neighbours of v are nbr[ptr[v] .. ptr[v] + len[v]).
struct HostRings {
    std::vector<int64_t> ptr;
    std::vector<int> len, nbr;

    void build(int V, int F, const int* tri) {
        std::vector<int> cap(V, 0);
        for (int64_t i = 0; i < 3 * (int64_t)F; i++) cap[tri[i]] += 2;
        ptr.assign((size_t)V + 1, 0);
        for (int v = 0; v < V; v++) ptr[v + 1] = ptr[v] + cap[v];
        nbr.assign((size_t)ptr[V], -1);
        len.assign(V, 0);
        for (int f = 0; f < F; f++) {
            const int* t = tri + 3 * (int64_t)f;
            if (t[0] == t[1]) continue;
            for (int k = 0; k < 3; k++) {
                int a = t[k], b = t[(k + 1) % 3];
                if (a == b) continue;
                bool known = false;
                const int* ra = &nbr[ptr[a]];
                for (int i = len[a] - 1; i >= 0; i--) if (ra[i] == b) { known = true; break; }
                if (known) continue;
                nbr[ptr[a] + len[a]++] = b;
                nbr[ptr[b] + len[b]++] = a;
            }
        }
    }
};

// Rings handed over ready-made (CSR layout, neighbours in the reference's ring order: k_ring_order on the device).
struct FlatRings {
    const int* ptr;   // V + 1
    const int* nbr;
};

inline void initial_random_sampling(int V, int K, const FlatRings& R, const double* weight,
                                    const std::vector<int64_t>& fixed, std::vector<int>& out);

// Returns the sampling (cluster id per vertex, K = unassigned never remains unless the mesh is disconnected).
inline void initial_random_sampling(int V, int K, const HostRings& R, const double* weight,
                                    const std::vector<int64_t>& fixed, std::vector<int>& out) {
    out.assign(V, K);
    int offset = 0;
    for (; offset < (int)fixed.size(); offset++) out[fixed[offset]] = offset;
    std::vector<int> order(V);
    for (int i = 0; i < V; i++) order[i] = i;
    std::mt19937 rng;
    rng.seed(0);
    const int n = V;
    for (int i = n - 1; i > 0; --i) std::swap(order[i], order[rng() % n]);   // as in the reference: not Fisher-Yates
    double total = 0;
    for (int i = 0; i < V; i++) total += weight[i];
    const double target = total / (double)K;
    int items_left = V, regions_left = K - offset, cursor = 0;
    std::queue<int> q;
    while (items_left > 0 && regions_left > 0) {
        while (cursor < V && out[order[cursor]] != K) cursor++;
        if (cursor >= V) break;
        std::queue<int>().swap(q);
        q.push(order[cursor]);
        double acc = 0;
        regions_left--;
        const int id = regions_left + offset;
        while (!q.empty()) {
            int it = q.front();
            q.pop();
            if (out[it] != K) continue;
            out[it] = id;
            acc += weight[it];
            items_left--;
            const int* r = &R.nbr[R.ptr[it]];
            for (int k = 0; k < R.len[it]; k++) q.push(r[k]);
            if (acc > target) break;
        }
    }
    if (regions_left == 0) return;
    // not enough seeds reached: steal single items from clusters of size > 1 (:1269-1311)
    std::vector<int> sizes(K, 0);
    for (int i = 0; i < V; i++) { order[i] = i; if (out[i] != K) sizes[out[i]]++; }
    for (int i = n - 1; i > 0; --i) std::swap(order[i], order[rng() % n]);
    cursor = 0;
    while (regions_left) {
        int it = -1;
        while (cursor < V) {
            it = order[cursor++];
            int c = out[it];
            if (c == K) break;
            if (sizes[c] == 1) continue;
            out[it] = regions_left + offset;
            sizes[c]--;
            if (regions_left + offset < K) sizes[regions_left + offset]++;
            break;
        }
        if (it < 0 || cursor > V) return;
        regions_left--;
        out[it] = regions_left + offset;
    }
}

// Same algorithm on flat rings.  The FIFO of a region is a small reused vector; a neighbour that is already assigned when
// it would be pushed is skipped -- the reference pushes it and drops it when popped (:1238), and assignments are never
// undone inside this loop, so the order of the entries that matter is unchanged.
inline void initial_random_sampling(int V, int K, const FlatRings& R, const double* weight,
                                    const std::vector<int64_t>& fixed, std::vector<int>& out) {
    out.assign(V, K);
    int offset = 0;
    for (; offset < (int)fixed.size(); offset++) out[fixed[offset]] = offset;
    std::vector<int> order(V);
    for (int i = 0; i < V; i++) order[i] = i;
    std::mt19937 rng;
    rng.seed(0);
    const int n = V;
    for (int i = n - 1; i > 0; --i) std::swap(order[i], order[rng() % n]);   // as in the reference: not Fisher-Yates
    double total = 0;
    for (int i = 0; i < V; i++) total += weight[i];
    const double target = total / (double)K;
    int items_left = V, regions_left = K - offset, cursor = 0;
    std::vector<int> q;
    q.reserve(4096);
    while (items_left > 0 && regions_left > 0) {
        while (cursor < V && out[order[cursor]] != K) cursor++;
        if (cursor >= V) break;
        q.clear();
        q.push_back(order[cursor]);
        size_t head = 0;
        double acc = 0;
        regions_left--;
        const int id = regions_left + offset;
        while (head < q.size()) {
            const int it = q[head++];
            if (out[it] != K) continue;
            out[it] = id;
            acc += weight[it];
            items_left--;
            for (int k = R.ptr[it]; k < R.ptr[it + 1]; k++) { const int u = R.nbr[k]; if (out[u] == K) q.push_back(u); }
            if (acc > target) break;
        }
    }
    if (regions_left == 0) return;
    // not enough seeds reached: steal single items from clusters of size > 1 (:1269-1311)
    std::vector<int> sizes(K, 0);
    for (int i = 0; i < V; i++) { order[i] = i; if (out[i] != K) sizes[out[i]]++; }
    for (int i = n - 1; i > 0; --i) std::swap(order[i], order[rng() % n]);
    cursor = 0;
    while (regions_left) {
        int it = -1;
        while (cursor < V) {
            it = order[cursor++];
            int c = out[it];
            if (c == K) break;
            if (sizes[c] == 1) continue;
            out[it] = regions_left + offset;
            sizes[c]--;
            if (regions_left + offset < K) sizes[regions_left + offset]++;
            break;
        }
        if (it < 0 || cursor > V) return;
        regions_left--;
        out[it] = regions_left + offset;
    }
}

}  // namespace acvd
