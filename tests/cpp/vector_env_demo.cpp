// Uses the C++ facade the way megaverse_test_app.cpp uses VectorEnv: random single-bit actions, prints a checksum line.
// Exit code 0 = ran; 3 = construction failed (e.g. no CUDA device: the product has no CPU fallback).
#include <cstdio>
#include <random>

#include "megaverse_b200_vector_env.hpp"

int main(int argc, char **argv) {
    const int numEnvs = argc > 1 ? atoi(argv[1]) : 4, numAgents = argc > 2 ? atoi(argv[2]) : 2, steps = argc > 3 ? atoi(argv[3]) : 50;
    try {
        megaverse_b200::VectorEnv venv("TowerBuilding", numEnvs, numAgents, 2);
        for (int e = 0; e < numEnvs; ++e) venv.seedEnv(e, 42 + e);
        venv.reset();
        std::mt19937 rng(42);
        std::uniform_int_distribution<> bit(0, 10);
        unsigned long long checksum = 0;
        double reward = 0;
        int dones = 0;
        for (int t = 0; t < steps; ++t) {
            for (int e = 0; e < numEnvs; ++e)
                for (int a = 0; a < numAgents; ++a) venv.setAction(e, a, 1 << bit(rng));
            venv.step();
            for (int e = 0; e < numEnvs; ++e) {
                dones += venv.done[size_t(e)] ? 1 : 0;
                for (int a = 0; a < numAgents; ++a) {
                    reward += venv.getLastReward(e, a);
                    const uint8_t *o = venv.getObservation(e, a);
                    for (int i = 0; i < 128 * 72 * 4; i += 97) checksum = checksum * 1315423911ull + o[i];
                }
            }
        }
        std::printf("vector_env_demo ok: %d envs x %d agents, %d steps, dones %d, reward %.3f, checksum %llu\n", numEnvs, numAgents, steps, dones, reward, checksum);
        venv.close();
    } catch (const std::exception &ex) {
        std::fprintf(stderr, "%s\n", ex.what());
        return 3;
    }
    return 0;
}
