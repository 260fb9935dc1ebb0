// CPU oracle -- TEST INFRASTRUCTURE ONLY (never linked or imported by the product).
// Restatement of the reference's honeycomb maze generator:
//   src/libs/mazes/src/honeycombmaze.cpp:1-84     (graph of hexagonal cells, cell centres, border segments, bounds)
//   src/libs/mazes/src/maze.cpp:9-39              (InitialiseGraph, GenerateMaze, RemoveBorders)
//   src/libs/mazes/src/kruskal.cpp:6-31           (random spanning tree: shuffled edges + union-find)
// Pinned against the real library compiled from /root/reference (oracle/_ref, tests/test_ref_shim.py).
//
// DEVIATION: upstream seeds the spanning-tree generator from std::random_device (spanningtreealgorithm.cpp:3-5), so the
// reference's mazes are NOT reproducible from the env seed.  Here the generator is seeded explicitly by the caller (the env
// derives it from its episode seed without consuming its own stream), which keeps every other draw of the env in place.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <numeric>
#include <random>
#include <utility>
#include <vector>

namespace orc {

struct HoneyCombMaze {
    struct Adj { int cell; std::array<double, 4> border; };  // neighbour (-1: outside) + border segment x1,y1,x2,y2
    int size = 0, vertices = 0;
    std::vector<std::vector<Adj>> adjacency;
    std::vector<std::pair<double, double>> cellCenters;

    explicit HoneyCombMaze(int n) : size(n), vertices(3 * n * (n - 1) + 1) {}

    std::pair<int, int> vExtent(int u) const { return u < 0 ? std::make_pair(-size - u + 1, size - 1) : std::make_pair(-size + 1, size - 1 - u); }
    bool isValidNode(int u, int v) const {
        if (u <= -size || u >= size) return false;
        const auto e = vExtent(u);
        return v >= e.first && v <= e.second;
    }
    int vertexIndex(int u, int v) const {
        if (u <= 0) return ((3 * size + u) * (size + u - 1)) / 2 + v;
        return (3 * size * (size - 1) + (4 * size - u - 1) * u) / 2 + v;
    }
    static std::pair<double, double> center(int u, int v) {
        const double dxu = std::sqrt(3) / 2, dyu = 1.5, dxv = std::sqrt(3), dyv = 0;
        return {dxu * u + dxv * v, dyu * u + dyv * v};
    }
    static std::array<double, 4> edge(int u, int v, int e) {
        const double dxu = std::sqrt(3) / 2, dyu = 1.5, dxv = std::sqrt(3), dyv = 0;
        const double cx = dxu * u + dxv * v, cy = dyu * u + dyv * v;
        const double theta1 = (e - 2.5) * M_PI / 3, theta2 = theta1 + M_PI / 3;
        return {cx + std::cos(theta1), cy + std::sin(theta1), cx + std::cos(theta2), cy + std::sin(theta2)};
    }
    void initialiseGraph() {
        static const int neigh[6][2] = {{-1, 0}, {-1, 1}, {0, 1}, {1, 0}, {1, -1}, {0, -1}};
        adjacency.assign(size_t(vertices), {});
        cellCenters.assign(size_t(vertices), {0.0, 0.0});
        for (int u = -size + 1; u < size; ++u) {
            const auto ve = vExtent(u);
            for (int v = ve.first; v <= ve.second; ++v) {
                const int node = vertexIndex(u, v);
                cellCenters[size_t(node)] = center(u, v);
                for (int n = 0; n < 6; ++n) {
                    const int uu = u + neigh[n][0], vv = v + neigh[n][1];
                    if (isValidNode(uu, vv)) {
                        const int nnode = vertexIndex(uu, vv);
                        if (nnode > node) continue;
                        const auto b = edge(u, v, n);
                        adjacency[size_t(node)].push_back({nnode, b});
                        adjacency[size_t(nnode)].push_back({node, b});
                    } else {
                        adjacency[size_t(node)].push_back({-1, edge(u, v, n)});  // bordersForEntranceAndExit == true
                    }
                }
            }
        }
    }
    // Kruskal::SpanningTree + Maze::RemoveBorders
    void generate(std::mt19937 &generator) {
        std::vector<std::pair<int, int>> edges;
        for (int i = 0; i < vertices; ++i)
            for (const auto &e : adjacency[size_t(i)])
                if (e.cell > i) edges.push_back({i, e.cell});
        std::shuffle(edges.begin(), edges.end(), generator);
        std::vector<int> parent(size_t(vertices), 0);
        std::iota(parent.begin(), parent.end(), 0);
        auto find = [&](int u) { int r = u; while (parent[size_t(r)] != r) r = parent[size_t(r)]; while (parent[size_t(u)] != r) { const int nx = parent[size_t(u)]; parent[size_t(u)] = r; u = nx; } return r; };
        for (const auto &e : edges) {
            const int u = find(e.first), v = find(e.second);
            if (u == v) continue;
            parent[size_t(u)] = v;
            auto erase1 = [&](int a, int b) {
                auto &l = adjacency[size_t(a)];
                for (size_t i = 0; i < l.size(); ++i)
                    if (l[i].cell == b) { l.erase(l.begin() + long(i)); break; }
            };
            erase1(e.first, e.second);
            erase1(e.second, e.first);
        }
    }
    std::array<double, 4> coordinateBounds() const {
        const double xlim = std::sqrt(3) * (size - 0.5), ylim = 1.5 * size - 0.5;
        return {-xlim, -ylim, xlim, ylim};
    }
};

}  // namespace orc
