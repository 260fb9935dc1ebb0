// pybind11 module `megaverse_b200.extension.megaverse`: the reference's MegaverseGym class surface
// (src/libs/bindings/megaverse.cpp:267-292 -- same method names, argument meaning and lifetime rules), implemented as a
// thin C++ host layer over the C ABI (include/megaverse_b200.h).  No CUDA or torch types cross this boundary.
//
// Differences that are deliberate and documented in INTEGRATION.md:
//   * use_vulkan is accepted and ignored: there is one backend (CUDA); without a GPU construction raises RuntimeError
//     (the reference would exit(-1) through TLOG(FATAL)).
//   * extra batched methods (set_actions_batch, get_observations, get_dones, ...) next to the per-agent ones, because the
//     reference's 2N+E+2 pybind round trips per step would cap throughput far below the kernels.
//   * the GIL is released around reset()/step().
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "megaverse_b200.h"

namespace py = pybind11;

namespace {

int g_logLevel = 2;
void setMegaverseLogLevel(int level) { g_logLevel = level; }  // tiny_logger.hpp:63-68; the engine itself does not log

class MegaverseGym {
public:
    MegaverseGym(const std::string &scenario, int w, int h, int numEnvs, int numAgentsPerEnv, int numSimulationThreads, bool useVulkan,
                 const std::map<std::string, float> &floatParams)
        : numEnvs_(numEnvs), numAgentsPerEnv_(numAgentsPerEnv), w_(w), h_(h) {
        (void)useVulkan;
        std::vector<const char *> keys;
        std::vector<float> vals;
        for (auto &kv : floatParams) { keys.push_back(kv.first.c_str()); vals.push_back(kv.second); }
        // the reference's constructor has no device argument (one process per GPU, CUDA_VISIBLE_DEVICES): MEGAVERSE_B200_DEVICE picks the
        // ordinal for processes that see several GPUs
        int device = 0;
        if (const char *dv = std::getenv("MEGAVERSE_B200_DEVICE")) device = std::atoi(dv);
        const int rc = mv_create(scenario.c_str(), w, h, numEnvs, numAgentsPerEnv, numSimulationThreads, device, keys.data(), vals.data(), int(keys.size()), &h__);
        if (rc != MV_OK) throw std::runtime_error(std::string("MegaverseGym: ") + mv_last_error(nullptr));
        masks_.assign(size_t(numEnvs) * numAgentsPerEnv, 0);
    }
    ~MegaverseGym() { close(); }

    void check(int rc) const {
        if (rc == MV_OK) return;
        const std::string msg = h__ ? mv_last_error(h__) : "closed";
        if (rc == MV_ERR_ARG) throw std::out_of_range(msg);
        throw std::runtime_error(msg);
    }
    void alive() const { if (!h__) throw std::runtime_error("MegaverseGym is closed"); }

    void seed(int seedValue) { alive(); check(mv_seed(h__, seedValue)); }
    int numAgents() const { return numAgentsPerEnv_; }
    std::vector<int> actionSpaceSizes() const { return {3, 3, 3, 2, 2, 3}; }  // Env::actionSpaceSizes, env.cpp:33

    void reset() {
        alive();
        int rc;
        { py::gil_scoped_release nogil; rc = mv_reset(h__); }
        check(rc);
    }
    void setActions(int envIdx, int agentIdx, std::vector<int> actions) {
        alive();
        if (envIdx < 0 || envIdx >= numEnvs_ || agentIdx < 0 || agentIdx >= numAgentsPerEnv_) throw std::out_of_range("set_actions: bad env/agent index");
        int32_t heads[6] = {0, 0, 0, 0, 0, 0};
        for (size_t i = 0; i < actions.size() && i < 6; ++i) heads[i] = actions[i];
        masks_[size_t(envIdx) * numAgentsPerEnv_ + agentIdx] = mv_encode_action(heads);
    }
    // batched extras: masks int32[N] (already encoded) or heads int32[N,6]
    void setActionsBatch(py::array_t<int32_t, py::array::c_style | py::array::forcecast> a) {
        alive();
        const size_t N = masks_.size();
        if (a.ndim() == 1 && size_t(a.shape(0)) == N) {
            std::memcpy(masks_.data(), a.data(), sizeof(int32_t) * N);
        } else if (a.ndim() == 2 && size_t(a.shape(0)) == N && a.shape(1) == 6) {
            for (size_t i = 0; i < N; ++i) masks_[i] = mv_encode_action(a.data() + i * 6);
        } else throw std::invalid_argument("set_actions_batch expects int32[N] masks or int32[N,6] heads");
    }
    void step() {
        alive();
        int rc;
        {
            py::gil_scoped_release nogil;
            rc = mv_set_actions(h__, masks_.data());
            if (rc == MV_OK) rc = mv_step(h__);
        }
        std::fill(masks_.begin(), masks_.end(), 0);  // env.cpp:140-142
        check(rc);
    }
    bool isDone(int envIdx) {
        alive();
        const uint8_t *d;
        check(mv_dones(h__, &d));
        if (envIdx < 0 || envIdx >= numEnvs_) throw std::out_of_range("is_done: bad env index");
        return d[envIdx] != 0;
    }
    std::vector<float> getLastRewards() {
        alive();
        const float *r;
        check(mv_rewards(h__, &r));
        return std::vector<float>(r, r + masks_.size());
    }
    py::array_t<uint8_t> getObservation(int envIdx, int agentIdx) {
        alive();
        const uint8_t *o;
        check(mv_obs_host(h__, &o));
        if (envIdx < 0 || envIdx >= numEnvs_ || agentIdx < 0 || agentIdx >= numAgentsPerEnv_) throw std::out_of_range("get_observation: bad env/agent index");
        const size_t view = size_t(envIdx) * numAgentsPerEnv_ + agentIdx;
        return py::array_t<uint8_t>({h_, w_, 4}, o + view * size_t(w_) * h_ * 4, py::none{});  // numpy object does not own memory
    }
    py::array_t<uint8_t> getObservations() {
        alive();
        const uint8_t *o;
        check(mv_obs_host(h__, &o));
        return py::array_t<uint8_t>({int(masks_.size()), h_, w_, 4}, o, py::none{});
    }
    py::array_t<float> getRewardsArray() {
        alive();
        const float *r;
        check(mv_rewards(h__, &r));
        return py::array_t<float>({int(masks_.size())}, r, py::none{});
    }
    py::array_t<uint8_t> getDones() {
        alive();
        const uint8_t *d;
        check(mv_dones(h__, &d));
        return py::array_t<uint8_t>({numEnvs_}, d, py::none{});
    }
    py::array_t<float> getTrueObjectives() {
        alive();
        const float *t;
        check(mv_true_objectives(h__, &t));
        return py::array_t<float>({int(masks_.size())}, t, py::none{});
    }
    float trueObjective(int envIdx, int agentIdx) const {
        alive();
        const float *t;
        check(mv_true_objectives(h__, &t));
        if (envIdx < 0 || envIdx >= numEnvs_ || agentIdx < 0 || agentIdx >= numAgentsPerEnv_) throw std::out_of_range("true_objective: bad env/agent index");
        return t[size_t(envIdx) * numAgentsPerEnv_ + agentIdx];
    }
    // hi-res rendering (megaverse.cpp:148-177,199-203): the same rasteriser at renderW x renderH; the overview camera / viewer is
    // out of scope (a reference build without WITH_GUI does nothing there either)
    void setRenderResolution(int hiresW, int hiresH) { renderW_ = hiresW; renderH_ = hiresH; }
    void drawHires() {
        alive();
        py::gil_scoped_release nogil;
        check(mv_draw_hires(h__, renderW_, renderH_, &hires_));
    }
    void drawOverview() {}
    py::array_t<uint8_t> getHiresObservation(int envIdx, int agentIdx) {
        alive();
        if (!hires_) throw std::runtime_error("get_hires_observation before draw_hires");
        if (envIdx < 0 || envIdx >= numEnvs_ || agentIdx < 0 || agentIdx >= numAgentsPerEnv_) throw std::out_of_range("get_hires_observation: bad env/agent index");
        const size_t view = size_t(envIdx) * numAgentsPerEnv_ + agentIdx;
        return py::array_t<uint8_t>({renderH_, renderW_, 4}, hires_ + view * size_t(renderW_) * renderH_ * 4, py::none{});  // does not own memory
    }

    std::map<std::string, float> getRewardShaping(int envIdx, int agentIdx) {
        alive();
        const char *keys[32];
        float vals[32];
        int n = 0;
        check(mv_get_reward_shaping(h__, envIdx, agentIdx, keys, vals, 32, &n));
        std::map<std::string, float> m;
        for (int i = 0; i < n && i < 32; ++i) m[keys[i]] = vals[i];
        return m;
    }
    void setRewardShaping(int envIdx, int agentIdx, const std::map<std::string, float> &rs) {
        alive();
        std::vector<const char *> keys;
        std::vector<float> vals;
        for (auto &kv : rs) { keys.push_back(kv.first.c_str()); vals.push_back(kv.second); }
        check(mv_set_reward_shaping(h__, envIdx, agentIdx, keys.data(), vals.data(), int(keys.size())));
    }
    uintptr_t obsDevicePtr() { alive(); uint8_t *p; check(mv_obs_device(h__, &p)); return reinterpret_cast<uintptr_t>(p); }
    int faults() { alive(); int32_t f; check(mv_faults(h__, &f)); return f; }
    int faultWord() { alive(); int32_t f; check(mv_fault_word(h__, &f)); return f; }
    void setOption(const std::string &key, int value) { alive(); check(mv_set_option(h__, key.c_str(), value)); }
    int levelsSkipped() { alive(); return mv_levels_skipped(h__); }

    void close() {
        if (h__) { mv_close(h__); h__ = nullptr; }
    }

private:
    mv_handle h__ = nullptr;
    int numEnvs_, numAgentsPerEnv_, w_, h_;
    int renderW_ = 768, renderH_ = 432;
    const uint8_t *hires_ = nullptr;
    std::vector<int32_t> masks_;
};

}  // namespace

PYBIND11_MODULE(megaverse, m) {
    m.doc() = "megaverse_b200 Python bindings (MegaverseGym surface of the reference)";
    m.def("set_megaverse_log_level", &setMegaverseLogLevel, "Megaverse Log Level (0 to disable all logs, 2 for warnings");
    py::class_<MegaverseGym>(m, "MegaverseGym")
        .def(py::init<const std::string &, int, int, int, int, int, bool, const std::map<std::string, float> &>())
        .def("num_agents", &MegaverseGym::numAgents)
        .def("action_space_sizes", &MegaverseGym::actionSpaceSizes)
        .def("seed", &MegaverseGym::seed)
        .def("reset", &MegaverseGym::reset)
        .def("set_actions", &MegaverseGym::setActions)
        .def("step", &MegaverseGym::step)
        .def("is_done", &MegaverseGym::isDone)
        .def("get_observation", &MegaverseGym::getObservation)
        .def("get_last_rewards", &MegaverseGym::getLastRewards)
        .def("true_objective", &MegaverseGym::trueObjective)
        .def("set_render_resolution", &MegaverseGym::setRenderResolution)
        .def("draw_hires", &MegaverseGym::drawHires)
        .def("draw_overview", &MegaverseGym::drawOverview)
        .def("get_hires_observation", &MegaverseGym::getHiresObservation)
        .def("get_reward_shaping", &MegaverseGym::getRewardShaping)
        .def("set_reward_shaping", &MegaverseGym::setRewardShaping)
        .def("close", &MegaverseGym::close)
        // batched / device extras
        .def("set_actions_batch", &MegaverseGym::setActionsBatch)
        .def("get_observations", &MegaverseGym::getObservations)
        .def("get_rewards", &MegaverseGym::getRewardsArray)
        .def("get_dones", &MegaverseGym::getDones)
        .def("get_true_objectives", &MegaverseGym::getTrueObjectives)
        .def("obs_device_ptr", &MegaverseGym::obsDevicePtr)
        .def("faults", &MegaverseGym::faults)
        .def("fault_word", &MegaverseGym::faultWord)
        .def("set_option", &MegaverseGym::setOption)
        .def("levels_skipped", &MegaverseGym::levelsSkipped);
}
