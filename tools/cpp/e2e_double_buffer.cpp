// Host-delivered throughput of a double-buffered consumer through the C ABI alone (no Python in the loop): G engines of total/G
// envs each, engine g's next step is begun right after its previous result was read (mv_step_begin / mv_step_end).
//   g++ -std=c++17 -O2 -Iinclude -o /tmp/e2e_db tools/cpp/e2e_double_buffer.cpp -Lmegaverse_b200 -lmegaverse_b200 -Wl,-rpath,$PWD/megaverse_b200
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "megaverse_b200.h"

int main(int argc, char **argv) {
    const char *scenario = argc > 1 ? argv[1] : "TowerBuilding";
    const int total = argc > 2 ? atoi(argv[2]) : 256, A = argc > 3 ? atoi(argv[3]) : 1, K = argc > 4 ? atoi(argv[4]) : 1500;
    const int configs[][2] = {{1, 1}, {1, 0}, {2, 1}, {2, 0}, {4, 0}, {4, 1}, {8, 0}};
    for (auto &cfg : configs) {
        const int G = cfg[0], zeroCopy = cfg[1], E = total / G, N = E * A;
        std::vector<mv_handle> hs(static_cast<size_t>(G));
        for (int g = 0; g < G; ++g) {
            if (mv_create(scenario, 128, 72, E, A, 8 / G > 0 ? 8 / G : 1, 0, nullptr, nullptr, 0, &hs[size_t(g)]) != MV_OK) { std::fprintf(stderr, "create failed: %s\n", mv_last_error(nullptr)); return 3; }
            mv_set_option(hs[size_t(g)], "zero_copy", zeroCopy);
            for (int e = 0; e < E; ++e) mv_seed_env(hs[size_t(g)], e, 42 + g * E + e);
            mv_reset(hs[size_t(g)]);
        }
        std::mt19937 rng(1);
        std::uniform_int_distribution<> bit(0, 10);
        std::vector<int32_t> masks(static_cast<size_t>(N));
        unsigned long long sink = 0;
        auto begin = [&](int g) {
            for (auto &m : masks) m = 1 << bit(rng);
            mv_set_actions(hs[size_t(g)], masks.data());
            if (mv_step_begin(hs[size_t(g)]) != MV_OK) { std::fprintf(stderr, "begin: %s\n", mv_last_error(hs[size_t(g)])); exit(4); }
        };
        auto end = [&](int g) {
            if (mv_step_end(hs[size_t(g)]) != MV_OK) { std::fprintf(stderr, "end: %s\n", mv_last_error(hs[size_t(g)])); exit(4); }
            const uint8_t *obs, *dones; const float *rew;
            mv_obs_host(hs[size_t(g)], &obs); mv_dones(hs[size_t(g)], &dones); mv_rewards(hs[size_t(g)], &rew);
            sink += obs[0] + obs[size_t(N) * 128 * 72 * 4 - 1] + dones[0] + (rew[0] != 0);
        };
        auto loop = [&](int n) {
            for (int g = 0; g < G; ++g) begin(g);
            for (int t = 1; t < n; ++t)
                for (int g = 0; g < G; ++g) { end(g); begin(g); }
            for (int g = 0; g < G; ++g) end(g);
        };
        loop(100);
        const auto t0 = std::chrono::steady_clock::now();
        loop(K);
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::printf("%s %dx%d: groups=%d zero_copy=%d  %7.1f us per %d obs  %.3f M obs/s (host-delivered, C ABI)  [%llu]\n", scenario, total, A, G, zeroCopy,
                    dt / K * 1e6, total * A, double(total) * A * K / dt / 1e6, sink);
        std::fflush(stdout);
        for (auto h : hs) mv_close(h);
    }
    return 0;
}
