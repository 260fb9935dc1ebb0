"""MegaverseEnv: the reference's Python env class (megaverse/megaverse_env.py:42-201) over the B200 engine.

Same constructor, attributes and methods, so Sample-Factory's Wrapper (megaverse_rl/megaverse_utils.py:48-88) drives it
unchanged.  The per-agent pybind loops of the reference (megaverse_env.py:121-162) are replaced by one batched call each;
the values returned are the same lists."""
import numpy as np

from .gym_shim import Box, Discrete, Env, Tuple

# the product fails loudly when its native extension is missing
from .extension.megaverse import MegaverseGym, set_megaverse_log_level  # noqa: E402

MEGAVERSE8 = ['TowerBuilding', 'ObstaclesEasy', 'ObstaclesHard', 'Collect', 'Sokoban', 'HexMemory', 'HexExplore', 'Rearrange']
OBSTACLES_MULTITASK = ['ObstaclesWalls', 'ObstaclesSteps', 'ObstaclesLava', 'ObstaclesEasy', 'ObstaclesHard']


def make_env_multitask(multitask_name, task_idx, num_envs, num_agents_per_env, num_simulation_threads, use_vulkan=False, params=None):
    assert 'multitask' in multitask_name
    if multitask_name.endswith('megaverse8'):
        tasks = MEGAVERSE8
    elif multitask_name.endswith('obstacles'):
        tasks = OBSTACLES_MULTITASK
    else:
        raise NotImplementedError()
    scenario = tasks[task_idx % len(tasks)]
    return MegaverseEnv(scenario, num_envs, num_agents_per_env, num_simulation_threads, use_vulkan, params)


FAULT_NAMES = {1: "LEVEL_NOT_READY", 2: "TRI_OVERFLOW", 4: "GRID_RANGE", 8: "NAN", 16: "ENVELOPE", 32: "CAND_OVERFLOW"}  # csrc/mv_types.h


class MegaverseFault(RuntimeError):
    """the engine left the envelope in which its results equal the reference's (a capacity of the collision or level code was exceeded,
    a NaN position, a level that arrived late): frames, rewards and dones from here on are not trustworthy"""


class MegaverseEnv(Env):
    # The number of static boxes of a level is unbounded, as in the reference; the remaining fixed capacities (movable objects, reward
    # objects, terrain slabs) were never exceeded on hundreds of thousands of generated levels.  Should one be, the strict default makes
    # step() / reset() fail instead of leaving the reference's level sequence; True takes the env's next level instead and counts it
    # (`levels_skipped()`, also reported in the infos).
    SKIP_UNFIT_LEVELS = False

    def __init__(self, scenario_name, num_envs, num_agents_per_env, num_simulation_threads, use_vulkan=False, params=None):
        scenario_name = scenario_name.casefold()
        self.scenario_name = scenario_name
        self.is_multiagent = True
        set_megaverse_log_level(2)
        self.img_w = 128
        self.img_h = 72
        self.channels = 3
        self.use_vulkan = use_vulkan
        self.num_agents = num_envs * num_agents_per_env
        self.num_envs = num_envs
        self.num_agents_per_env = num_agents_per_env

        float_params = {}
        if params is not None:
            for k, v in params.items():
                if isinstance(v, float):
                    float_params[k] = v
                else:
                    raise Exception('Params of type %r not supported', type(v))

        self.env = MegaverseGym(self.scenario_name, self.img_w, self.img_h, num_envs, num_agents_per_env, num_simulation_threads, use_vulkan, float_params)
        if self.SKIP_UNFIT_LEVELS:
            self.env.set_option("skip_unfit_levels", 1)
        self.default_shaping_scheme = self.env.get_reward_shaping(0, 0)
        self.action_space = self.generate_action_space(self.env.action_space_sizes())
        self.observation_space = Box(0, 255, (self.channels, self.img_h, self.img_w), dtype=np.uint8)

    @staticmethod
    def generate_action_space(action_space_sizes):
        return Tuple([Discrete(sz) for sz in action_space_sizes])

    def seed(self, seed=None):
        if seed is None:
            return
        assert isinstance(seed, int), 'Expect seed to be an integer'
        self.env.seed(seed)

    def observations(self):
        # one [N,h,w,4] view of engine memory -> per-agent CHW views, exactly what the reference's loop produces
        obs = self.env.get_observations()
        chw = np.transpose(obs[:, :, :, :3], (0, 3, 1, 2))
        return [chw[i] for i in range(self.num_agents)]

    def check_faults(self):
        """raise if the engine latched a fault bit (one pinned-memory read, no device round trip)"""
        word = self.env.fault_word()
        if word:
            names = [n for b, n in FAULT_NAMES.items() if word & b]
            raise MegaverseFault("megaverse_b200 engine fault bits 0x%x (%s)" % (word, ", ".join(names)))

    def reset(self):
        self.env.reset()
        self.check_faults()
        return self.observations()

    def step(self, actions):
        self.env.set_actions_batch(np.asarray(actions, dtype=np.int32).reshape(self.num_agents, 6))
        self.env.step()
        self.check_faults()

        env_dones = self.env.get_dones()
        dones, infos = [], []
        for env_i in range(self.num_envs):
            done = bool(env_dones[env_i])
            dones.extend([done for _ in range(self.num_agents_per_env)])
            if done:
                infos.extend([dict(true_reward=float(self.env.true_objective(env_i, j))) for j in range(self.num_agents_per_env)])
            else:
                infos.extend([{} for _ in range(self.num_agents_per_env)])

        if self.SKIP_UNFIT_LEVELS:
            skipped = self.env.levels_skipped()
            if skipped:
                for info in infos:
                    info['levels_skipped'] = skipped
        rewards = self.env.get_last_rewards()
        obs = self.observations()
        return obs, rewards, dones, infos

    def convert_obs(self, obs):
        return obs[:, :, ::-1]  # RGB -> BGR for display; rows are already top-down (the Vulkan path of the reference)

    def render(self, mode='human'):
        self.env.draw_overview()
        self.env.draw_hires()
        rows = []
        for env_i in range(self.num_envs):
            obs = [self.convert_obs(self.env.get_hires_observation(env_i, i)[:, :, :3]) for i in range(self.num_agents_per_env)]
            rows.append(np.concatenate(obs, axis=1))
        obs_final = np.concatenate(rows, axis=0)
        if mode == 'human':
            try:
                import cv2
                cv2.imshow(f'agent_{id(self)}', obs_final)
                cv2.waitKey(1)
            except Exception:  # noqa: BLE001  (headless boxes)
                pass
        return obs_final

    def get_default_reward_shaping(self):
        return self.default_shaping_scheme

    def get_current_reward_shaping(self, actor_idx: int):
        env_idx = actor_idx // self.num_agents_per_env
        agent_idx = actor_idx % self.num_agents_per_env
        return self.env.get_reward_shaping(env_idx, agent_idx)

    def set_reward_shaping(self, reward_shaping: dict, actor_idx: int):
        env_idx = actor_idx // self.num_agents_per_env
        agent_idx = actor_idx % self.num_agents_per_env
        return self.env.set_reward_shaping(env_idx, agent_idx, reward_shaping)

    def levels_skipped(self):
        """(extension) levels replaced because they exceeded an engine capacity, see __init__"""
        return self.env.levels_skipped()

    def close(self):
        if self.env:
            self.env.close()
