/* megaverse_b200 -- thin C ABI of the B200 batched voxel-world step + render engine.
 *
 * Drop-in boundary for the reference's per-step hot path.  Each entry point replaces what the reference's pybind11
 * class MegaverseGym (src/libs/bindings/megaverse.cpp:36-263) reaches through VectorEnv (src/libs/env/src/vector_env.cpp)
 * and EnvRenderer (src/libs/env/include/env/env_renderer.hpp:12-32).  Plain pointers and sizes only; no torch / pybind
 * types.  Every call returns 0 on success or a negative MV_ERR_* code; mv_last_error() gives the message.  The engine
 * never calls exit() (the reference's TLOG(FATAL) does, src/libs/util/src/tiny_logger.cpp:109-113).
 *
 * Threading: calls on one handle must be serialised by the caller (the reference holds the GIL for every call).
 * Observation memory is owned by the engine and reused every step, exactly like the reference's renderer buffer
 * (megaverse.cpp:139-143): pointers stay valid until mv_close, contents until the next mv_step / mv_reset.
 */
#ifndef MEGAVERSE_B200_H
#define MEGAVERSE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mv_engine *mv_handle;

#define MV_OK 0
#define MV_ERR_ARG -1          /* bad argument (unknown scenario, bad sizes, unknown reward key -> std::out_of_range in the reference) */
#define MV_ERR_CUDA -2         /* CUDA runtime failure, or no CUDA device: the product has NO CPU fallback */
#define MV_ERR_CAPACITY -3     /* a generated level exceeds the engine's fixed capacities */
#define MV_ERR_STATE -4        /* call order (e.g. step before reset) */

/* MegaverseGym::MegaverseGym (megaverse.cpp:38-58).  num_threads = host level-generation workers (the reference's
 * numSimulationThreads drove Bullet on the CPU, vector_env.cpp:6-40).  device = CUDA ordinal. */
int mv_create(const char *scenario, int w, int h, int num_envs, int num_agents_per_env, int num_threads, int device,
              const char *const *param_keys, const float *param_vals, int nparams, mv_handle *out);
/* message of the last failed call; pass NULL for a failed mv_create */
const char *mv_last_error(mv_handle h);

/* MegaverseGym::seed (megaverse.cpp:60-69): master mt19937 -> one randRange(0,1<<30) per env */
int mv_seed(mv_handle h, int seed);
/* Env::seed for one env (megaverse_test_app.cpp:250-254 seeds env i with 42+i) */
int mv_seed_env(mv_handle h, int env, int seed);

/* MegaverseGym::reset -> VectorEnv::reset (vector_env.cpp:110-120): new episode in every env + first render */
int mv_reset(mv_handle h);

/* MegaverseGym::setActions for all agents at once: masks[env*A + agent] = Action bit mask (env.hpp:22-42).
 * mv_encode_action converts one 6-tuple of Discrete heads {3,3,3,2,2,3} exactly like megaverse.cpp:100-116. */
int mv_set_actions(mv_handle h, const int32_t *masks);
int32_t mv_encode_action(const int32_t *heads6);

/* MegaverseGym::step -> VectorEnv::step (vector_env.cpp:89-108): physics + scenario logic, reset of finished envs, render.
 * Synchronous: on return observations / rewards / dones are in host memory. */
int mv_step(mv_handle h);
/* mv_step in two halves, for a double-buffered consumer (two engines of half the envs each, the arrangement Sample Factory's
 * sampler uses with two env groups per worker: while the policy looks at group A, group B steps).  mv_step_begin uploads the
 * action masks and enqueues the kernels and the device->host copies, then returns; mv_step_end waits for them and does the
 * episode bookkeeping -- on return the host buffers are valid exactly as after mv_step.  With option "zero_copy" 0 the
 * observation tensor travels by the copy engine, which overlaps with the other engine's kernels.  Any other call that needs a
 * finished step (mv_reset, mv_step_device, mv_fetch_obs) ends an outstanding begin first; a second begin is MV_ERR_STATE. */
int mv_step_begin(mv_handle h);
int mv_step_end(mv_handle h);

/* MegaverseGym::getObservation (megaverse.cpp:139-143): uint8[N][h][w][4] RGBA, view index env*A+agent, host memory */
int mv_obs_host(mv_handle h, const uint8_t **out);
/* float32[N][h][w] view-space depth (V4R depth output definition), only when option "depth" is 1 */
int mv_depth_host(mv_handle h, const float **out);
/* MegaverseGym::getLastRewards (megaverse.cpp:128-137): float[N] */
int mv_rewards(mv_handle h, const float **out);
/* VectorEnv::done (vector_env.hpp): uint8[num_envs] */
int mv_dones(mv_handle h, const uint8_t **out);
/* VectorEnv::trueObjectives / MegaverseGym::trueObjective (megaverse.cpp:209-212): float[N] */
int mv_true_objectives(mv_handle h, const float **out);

/* MegaverseGym::getRewardShaping / setRewardShaping (megaverse.cpp:214-222).  get: fills up to cap entries, returns the
 * number of keys in *n.  Key strings are owned by the engine. */
int mv_get_reward_shaping(mv_handle h, int env, int agent, const char **keys, float *vals, int cap, int *n);
int mv_set_reward_shaping(mv_handle h, int env, int agent, const char *const *keys, const float *vals, int n);

/* options: "depth" (0/1, before first reset), "obs_to_host" (0/1: whether mv_step delivers the observation tensor to host
 * memory; 1 by default), "zero_copy" (0/1, default 1: host-facing steps let the rasteriser store rows straight into the pinned host buffer, the
 * PCIe writes overlapping the drawing -- the HBM tensor is then stale and mv_obs_device refuses it until the next mv_step_device; 0:
 * rasterise into HBM in "host_slices" launches (0 = by size) whose downloads run on the copy engine while the next slice is drawn),
 * "host_progressive" (0..16, default 0: host-facing steps rasterise into HBM in ONE launch and the copy engine follows it slice by slice of
 * whole envs -- the rasteriser counts finished work items per slice, the copy stream waits on the counters; measured slower than the
 * zero-copy stores wherever single views are expensive),
 * "fast_shading" (0/1, default 1: +-1 LSB fragment maths),
 * "tri_cap" (32..1022, default 368: the largest that leaves two CTAs per SM; triangles one raster CTA keeps in shared memory; a view with more is drawn in several batches,
 * results do not depend on it), "raster_bands" (row bands a view is cut into, one work item of the persistent raster grid each;
 * chosen from the number of views by default, results do not depend on it),
 * "raster_grid" (CTAs of the persistent raster grid, 0 = as many as the GPU holds (default): engines that share one GPU -- one per scenario
 * of a mixed batch -- take a share each so that their grids run side by side instead of queueing behind each other),
 * "raster_sched" (0/1/2, default 1: launches with more work items than raster CTAs draw the envs in the order of what their views cost in
 * the previous step, most expensive first, and the step kernel steps them in that order; 0 natural order, 2 always; results do not depend
 * on it),
 * "static_cap" (before the first reset: initial size of the per-level static-box arrays, default 768; they grow whenever a
 * generated level has more boxes -- the reference has no bound, component_voxel_grid.hpp:108-187),
 * "skip_unfit_levels" (0/1, default 0: a generated level that exceeds one of the remaining fixed capacities -- movable objects, reward
 * objects, terrain slabs, the dense grid -- makes mv_step / mv_reset fail with MV_ERR_CAPACITY, and keeps failing: the env would
 * otherwise leave the reference's level sequence; with 1 the env takes the next level of its stream instead and mv_levels_skipped
 * counts it),
 * "overlap" (0/1, default 1: the raster kernel is a programmatic dependent launch of the step kernel and synchronises per env;
 * 0 serialises the kernels so that mv_last_kernel_ms can time them separately) */
int mv_set_option(mv_handle h, const char *key, int value);

/* Device-resident path: the caller's consumer reads the tensors in HBM.
 * mv_step_device: like mv_step but takes the action masks from DEVICE memory (NULL = the engine's own buffer, see
 * mv_actions_device) and leaves the observation tensor on the device; rewards/dones still land on the host. */
int mv_step_device(mv_handle h, const int32_t *d_masks);
/* Redirect the HBM output of the rasteriser into caller-owned device memory: d_obs = uint8[N][h][w][4] (and d_depth = float[N][h][w]
 * when option depth is on; NULL keeps the engine's own).  Lets several engines -- one per scenario of a multi-task batch, reference
 * megaverse_env.py:27-39 -- write into slices of ONE contiguous tensor that a consumer or an NCCL gather reads without a staging copy.
 * NULL, NULL restores the engine's own buffers.  The pointer takes effect with the next step; steps already enqueued keep writing the
 * previous buffer (no synchronisation: a consumer that double-buffers its tensor -- e.g. to gather step t over NCCL while step t+1 is
 * drawn -- switches every step); mv_sync before freeing a buffer. */
int mv_set_obs_buffer(mv_handle h, uint8_t *d_obs, float *d_depth);
/* mv_step_device is ASYNCHRONOUS: it returns after enqueueing the step on the engine stream (device tensors are valid in
 * stream order).  mv_sync waits for everything enqueued and publishes the last step's rewards/dones/true objectives to the
 * host pointers.  Episode bookkeeping lags the device by two steps on this path, so it needs episodes of >= 4 steps
 * (always true with the scenarios' own parameters); a violation raises MV_FAULT_LEVEL_NOT_READY in mv_faults. */
/* MegaverseGym::drawHires + getHiresObservation (megaverse.cpp:154-177,199-203): renders every agent view once more at w x h
 * (multiples of 32 x 4, e.g. the reference's 768 x 432) from the state of the last step; *out = uint8[N][h][w][4], engine-owned,
 * valid until the next mv_draw_hires / mv_close */
int mv_draw_hires(mv_handle h, int w, int hgt, const uint8_t **out);
int mv_sync(mv_handle h);
/* after mv_step_device steps: waits, then copies the device obs (and depth, when enabled) into the host buffers that
 * mv_obs_host / mv_depth_host return */
int mv_fetch_obs(mv_handle h);
int mv_actions_device(mv_handle h, int32_t **d_masks);
int mv_obs_device(mv_handle h, uint8_t **d_obs);
int mv_depth_device(mv_handle h, float **d_depth);
int mv_rewards_device(mv_handle h, float **d_rewards);
int mv_dones_device(mv_handle h, uint8_t **d_dones);
/* the CUDA stream (cudaStream_t) all engine work is ordered on */
int mv_stream(mv_handle h, void **stream);

/* sticky per-env fault bits ORed over all envs (MV_FAULT_* in mv_types.h); 0 = healthy */
int mv_faults(mv_handle h, int32_t *out);
/* the same bits without a device round trip: the step kernel ORs every fault it raises into a pinned host word (sticky).  Valid for
 * the steps whose results the host has (after mv_step / mv_step_end / mv_sync).  Non-zero means physics or level state left the
 * envelope the engine guarantees (MV_FAULT_* in csrc/mv_types.h): the Python MegaverseEnv raises on it. */
int mv_fault_word(mv_handle h, int32_t *out);
/* number of kernels the engine launched since creation */
int mv_kernel_launches(mv_handle h, int64_t *out);
/* device time of the last step's kernels in milliseconds: [0] step kernel, [1] raster kernel (CUDA events) */
int mv_last_kernel_ms(mv_handle h, float *out2);

/* MegaverseGym::close (megaverse.cpp:224-243) */
int mv_close(mv_handle h);

/* ---- introspection for parity tests (same layouts as the oracle's orc_get_* in oracle/orc_api.cpp) ---- */
int mv_debug_get_level(mv_handle h, int env, int32_t *out, int cap);
int mv_debug_get_state(mv_handle h, int env, float *out, int cap);
int mv_debug_get_voxels(mv_handle h, int env, int32_t *out, int cap);
int mv_debug_get_instances(mv_handle h, int env, float *out, int cap);
int mv_debug_get_view(mv_handle h, int env, int agent, float *out16);
/* render caller-supplied instances (18 floats each: mesh, colour, 16 model) with one view matrix through the CUDA
 * rasteriser: rgba uint8[h][w][4], depth float[h][w] or NULL.  Host pointers. */
int mv_debug_render_instances(const float *view16, const float *inst18, int n, int w, int h, uint8_t *rgba, float *depth);
/* host-only (no CUDA needed): run the level generator for the env RNG stream seeded with env_seed and dump the level of
 * episode `episode` (mv_debug_get_level layout, followed by each agent's 9 spawn-basis floats as bit patterns) */
int mv_debug_generate_level(const char *scenario, int num_agents, int env_seed, int episode, const char *const *param_keys,
                            const float *param_vals, int nparams, int32_t *out, int cap);
/* libstdc++ unordered_set iteration-order emulation (bzset.h): ops[i] = {op(0 insert,1 erase,2 clear), x, y, z};
 * writes the final iteration order as xyz triples, returns the element count */
/* per-env cycle stamps of the step kernel's phases: out = uint32[E][16] (0 staged, 1 actions, 2 candidate list, 3 controllers,
 * 4 transforms, 5 scenario, 6 outputs/reset, 7 instance list, 8 commit, 12 candidate count); enable=1 arms it, 0 frees it */
int mv_debug_step_profile(mv_handle h, uint32_t *out, int enable);
/* rasteriser launch shape: out4 = {persistent grid size, CTAs per SM, dynamic shared memory per CTA in bytes, row bands per view} */
int mv_debug_raster_config(mv_handle h, int32_t *out4);
/* current size of the per-level static-box arrays (option "static_cap" at start, grows on demand) */
int mv_debug_static_cap(mv_handle h);
/* rasteriser work counters since the last enable: out16 = {work items, instances read, instances with visible items, items (box faces /
 * mesh triangles set up), items clipped at the near / far plane, triangles drawn, batches, -, then thread-0 cycle sums: head (work
 * claim, env stamp, view matrix), TMA waits, instance passes, item passes, final tile pass, whole work item, -, -}; enable=1 arms /
 * clears, 0 frees */
int mv_debug_raster_stats(mv_handle h, unsigned long long *out16, int enable);
/* host-only: colour tables of the generators + the rasteriser's palette (tests pin them against the reference's env/const.hpp) */
int mv_debug_color_tables(uint32_t *out, int cap);
/* host-only: default reward shaping ("R key=hexbits") and default float parameters ("P key=hexbits") of a scenario, one per line */
int mv_debug_defaults(const char *scenario, char *out, int cap);
/* host-only: number of levels among the first `episodes` of the env stream seeded env_seed that do not fit the engine's capacities */
int mv_debug_count_unfit_levels(const char *scenario, int num_agents, int env_seed, int episodes, const char *const *keys, const float *vals, int nparams);
/* levels replaced so far under option "skip_unfit_levels" (0 unless that option is set) */
int mv_levels_skipped(mv_handle h);
int mv_debug_bzset(const int32_t *ops, int nops, int32_t *out_xyz, int cap);

#ifdef __cplusplus
}
#endif
#endif
