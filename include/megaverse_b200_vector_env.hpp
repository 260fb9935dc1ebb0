// C++ facade over the C ABI with the shape of the reference's VectorEnv + EnvRenderer pair
//   src/libs/env/include/env/vector_env.hpp:14-57   (step / reset / close, public `done` and `trueObjectives`)
//   src/libs/env/include/env/env_renderer.hpp:12-32 (getObservation(envIdx, agentIdx) -> const uint8_t*)
//   src/libs/env/include/env/env.hpp (setAction, getLastReward) for the per-agent accessors the bindings use
// Header only; link against megaverse_b200/libmegaverse_b200.so.  Errors are exceptions (the reference exit()s).
#pragma once
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "megaverse_b200.h"

namespace megaverse_b200 {

class VectorEnv {
public:
    VectorEnv(const std::string &scenario, int numEnvs, int numAgentsPerEnv, int numThreads, int w = 128, int h = 72,
              const std::map<std::string, float> &floatParams = {}, int device = 0)
        : numEnvs_(numEnvs), numAgents_(numAgentsPerEnv), w_(w), h_(h), done(size_t(numEnvs), false),
          trueObjectives(size_t(numEnvs), std::vector<float>(size_t(numAgentsPerEnv), 0.0f)), masks_(size_t(numEnvs) * size_t(numAgentsPerEnv), 0) {
        std::vector<const char *> keys;
        std::vector<float> vals;
        for (auto &kv : floatParams) { keys.push_back(kv.first.c_str()); vals.push_back(kv.second); }
        if (mv_create(scenario.c_str(), w, h, numEnvs, numAgentsPerEnv, numThreads, device, keys.data(), vals.data(), int(keys.size()), &h__) != MV_OK)
            throw std::runtime_error(std::string("megaverse_b200::VectorEnv: ") + mv_last_error(nullptr));
    }
    ~VectorEnv() { close(); }
    VectorEnv(const VectorEnv &) = delete;
    VectorEnv &operator=(const VectorEnv &) = delete;

    void seed(int s) { check(mv_seed(h__, s)); }
    void seedEnv(int envIdx, int s) { check(mv_seed_env(h__, envIdx, s)); }  // Env::seed

    void reset() {  // VectorEnv::reset (vector_env.cpp:110-120)
        check(mv_reset(h__));
        refresh();
    }
    void setAction(int envIdx, int agentIdx, int actionMask) { masks_.at(size_t(envIdx) * size_t(numAgents_) + size_t(agentIdx)) = actionMask; }  // Env::setAction
    void step() {  // VectorEnv::step (vector_env.cpp:89-108): all envs, finished ones already reset, observations rendered
        check(mv_set_actions(h__, masks_.data()));
        check(mv_step(h__));
        std::fill(masks_.begin(), masks_.end(), 0);  // env.cpp:140-142
        refresh();
    }
    void close() {
        if (h__) { mv_close(h__); h__ = nullptr; }
    }

    const uint8_t *getObservation(int envIdx, int agentIdx) const {  // EnvRenderer::getObservation: uint8[h][w][4], valid until the next step
        const uint8_t *o = nullptr;
        check(mv_obs_host(h__, &o));
        return o + (size_t(envIdx) * size_t(numAgents_) + size_t(agentIdx)) * size_t(w_) * size_t(h_) * 4;
    }
    float getLastReward(int envIdx, int agentIdx) const {
        const float *r = nullptr;
        check(mv_rewards(h__, &r));
        return r[size_t(envIdx) * size_t(numAgents_) + size_t(agentIdx)];
    }
    int numEnvs() const { return numEnvs_; }
    int numAgents() const { return numAgents_; }
    mv_handle handle() const { return h__; }

private:
    void check(int rc) const {
        if (rc != MV_OK) throw std::runtime_error(std::string("megaverse_b200::VectorEnv: ") + (h__ ? mv_last_error(h__) : "closed"));
    }
    void refresh() {
        const uint8_t *d = nullptr;
        const float *t = nullptr;
        check(mv_dones(h__, &d));
        check(mv_true_objectives(h__, &t));
        for (int e = 0; e < numEnvs_; ++e) {
            done[size_t(e)] = d[e] != 0;
            if (d[e])  // captured at the step the episode ended, before the reset (vector_env.cpp:94-99)
                for (int a = 0; a < numAgents_; ++a) trueObjectives[size_t(e)][size_t(a)] = t[size_t(e) * size_t(numAgents_) + size_t(a)];
        }
    }

    mv_handle h__ = nullptr;
    int numEnvs_, numAgents_, w_, h_;

public:
    std::vector<bool> done;                          // VectorEnv::done
    std::vector<std::vector<float>> trueObjectives;  // VectorEnv::trueObjectives

private:
    std::vector<int32_t> masks_;
};

}  // namespace megaverse_b200
