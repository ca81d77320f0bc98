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
// The walk is bound by cache misses (a region starts at a random vertex; the pseudo-shuffle swaps with a random slot), so
// the addresses that are known ahead of time are prefetched: the swap partners of the shuffle (the generator does not
// depend on the array), and the ring / weight of every vertex as it enters the FIFO.
inline void pseudo_shuffle(std::vector<int>& order, std::mt19937& rng) {
    const int n = (int)order.size();
    constexpr int B = 1024, D = 24;                 // generator block, prefetch distance
    uint32_t js[B];
    int i = n - 1;
    while (i > 0) {                                 // for (i = n - 1; i > 0; --i) swap(order[i], order[rng() % n]): not Fisher-Yates, as in the reference
        const int cnt = i < B ? i : B;
        for (int k = 0; k < cnt; k++) js[k] = (uint32_t)(rng() % (uint32_t)n);
        for (int k = 0; k < D && k < cnt; k++) __builtin_prefetch(&order[js[k]], 1);
        for (int k = 0; k < cnt; k++) {
            if (k + D < cnt) __builtin_prefetch(&order[js[k + D]], 1);
            std::swap(order[i - k], order[js[k]]);
        }
        i -= cnt;
    }
}

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
    pseudo_shuffle(order, rng);
    double total = 0;
    for (int i = 0; i < V; i++) total += weight[i];
    const double target = total / (double)K;
    int items_left = V, regions_left = K - offset, cursor = 0;
    std::vector<int> q;
    q.reserve(4096);
    // software pipeline over the FIFO: the neighbour row of the entry PD1 ahead is requested, and the cluster ids / weights
    // of the neighbours of the entry PD2 ahead (whose row has arrived by then)
    constexpr int PD1 = 8, PD2 = 4;
    while (items_left > 0 && regions_left > 0) {
        while (cursor < V && out[order[cursor]] != K) {
            if (cursor + 32 < V) __builtin_prefetch(&out[order[cursor + 32]]);
            cursor++;
        }
        if (cursor >= V) break;
        q.clear();
        q.push_back(order[cursor]);
        size_t head = 0;
        double acc = 0;
        regions_left--;
        const int id = regions_left + offset;
        while (head < q.size()) {
            if (head + PD1 < q.size()) { const int f = q[head + PD1]; __builtin_prefetch(&R.nbr[R.ptr[f]]); __builtin_prefetch(&R.nbr[R.ptr[f + 1] - 1]); }
            if (head + PD2 < q.size()) {
                const int f = q[head + PD2];
                __builtin_prefetch(&weight[f]);
                for (int k = R.ptr[f]; k < R.ptr[f + 1]; k++) __builtin_prefetch(&out[R.nbr[k]]);
            }
            const int it = q[head++];
            if (out[it] != K) continue;
            out[it] = id;
            acc += weight[it];
            items_left--;
            for (int k = R.ptr[it]; k < R.ptr[it + 1]; k++) {
                const int u = R.nbr[k];
                if (out[u] == K) { q.push_back(u); __builtin_prefetch(&R.ptr[u]); }
            }
            if (acc > target) break;
        }
    }
    if (regions_left == 0) return;
    // not enough seeds reached: steal single items from clusters of size > 1 (:1269-1311)
    std::vector<int> sizes(K, 0);
    for (int i = 0; i < V; i++) { order[i] = i; if (out[i] != K) sizes[out[i]]++; }
    pseudo_shuffle(order, rng);
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
