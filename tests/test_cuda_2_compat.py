"""GPU tests of the reference-facing single-env surface (crowdnav_b200.compat): the flow of crowd_nav/test.py:52-109
(gym.make('CrowdSim-v0') / configure / Robot / ORCA / Explorer.run_k_episodes) against the reference's recorded results."""
import logging

import numpy as np
import pytest
import torch

from util import load_golden, scene_arrays

pytestmark = pytest.mark.gpu


def _make(human_num=5, test_sim='circle_crossing', robot_visible=False):
    import crowdnav_b200.compat as compat
    from crowdnav_b200.batched import default_config
    compat.install()
    import gym
    from crowd_sim.envs.utils.robot import Robot
    from crowd_sim.envs.policy.orca import ORCA
    cfg = default_config(human_num=human_num, test_sim=test_sim, robot_visible=robot_visible)
    env = gym.make('CrowdSim-v0')
    env.configure(cfg)
    robot = Robot(cfg, 'robot')
    policy = ORCA()
    robot.set_policy(policy)
    env.set_robot(robot)
    policy.set_phase('test'); policy.set_device(torch.device('cuda:0')); policy.set_env(env)
    return env, robot


def test_reference_test_py_flow_reproduces_first_cases():
    from crowdnav_b200.explorer import summarize
    d = load_golden('suite_circle5_invisible')
    k = 16
    env, robot = _make()
    from crowd_nav.utils.explorer import Explorer
    explorer = Explorer(env, robot, torch.device('cuda:0'), gamma=0.9)
    lines = []
    handler = logging.Handler(); handler.emit = lambda rec: lines.append(rec.getMessage())
    root = logging.getLogger(); root.addHandler(handler); old = root.level; root.setLevel(logging.INFO)
    try:
        explorer.run_k_episodes(k, 'test', print_failure=True)
    finally:
        root.removeHandler(handler); root.setLevel(old)
    rows = torch.tensor([[c['info'], c['steps'], 25.0 if c['info'] == 4 else float(c['global_time']), float(c['return']),
                          c['too_close'], float(c['min_dist_sum'])] for c in d['cases'][:k]], dtype=torch.float64)
    expect = []
    summarize(rows, k, 'test', 25, 0.25, print_failure=True, log=expect.append)
    assert [l for l in lines if 'human number' not in l and 'andomize' not in l and 'simulation' not in l and 'width' not in l] == expect
    assert env.case_counter['test'] == k


def test_step_and_lookahead_semantics():
    d = load_golden('traj_circle5_invisible')['trajectories']['3']
    env, robot = _make()
    from crowd_sim.envs.utils.info import Danger, Nothing, Collision, ReachGoal, Timeout
    ob = env.reset('test', 3)
    r0, h0 = scene_arrays(d[0]['pre'])
    assert abs(env.humans[0].px - h0[0, 0]) < 1e-12 and len(ob) == 5 and env.global_time == 0
    for t, s in enumerate(d):
        action = robot.act(ob)
        # lookahead must not mutate anything and must agree with the real step that follows
        before = (robot.px, robot.py, env.global_time, [h.px for h in env.humans])
        ob_l, rew_l, done_l, info_l = env.onestep_lookahead(action)
        assert before == (robot.px, robot.py, env.global_time, [h.px for h in env.humans])
        ob, reward, done, info = env.step(action)
        assert (rew_l, done_l, type(info_l)) == (reward, done, type(info))
        assert [o.px for o in ob_l] == [o.px for o in ob] and [o.vy for o in ob_l] == [o.vy for o in ob]
        assert abs(action.vx - float(s['action'][0])) < 1e-6 and abs(action.vy - float(s['action'][1])) < 1e-6
        assert abs(reward - float(s['reward'])) < 1e-9 and done == s['done']
        assert {0: Nothing, 1: Danger, 2: ReachGoal, 3: Collision, 4: Timeout}[s['info']] is type(info)
        rp, hp = scene_arrays(s['post'])
        assert abs(robot.px - rp[0]) < 1e-5 and abs(robot.py - rp[1]) < 1e-5
        assert np.abs(np.array([[h.px, h.py] for h in env.humans]) - hp[:, :2]).max() < 1e-5
    assert done and str(info) == 'Reaching goal' and abs(env.global_time - 0.25 * len(d)) < 1e-12


def test_get_human_times_through_the_single_env_surface():
    """env.get_human_times() (crowd_sim.py:209-249) after an episode the ORCA robot finished at its goal: the reference's
    recorded arrival times; raises like the reference while the robot is not at its goal."""
    r = [x for x in load_golden('human_times')['rows'] if x['tag'] == 'circle5'][0]
    env, robot = _make()
    ob = env.reset('test', r['case'])
    with pytest.raises(ValueError):
        env.get_human_times()
    done = False
    while not done:
        ob, reward, done, info = env.step(robot.act(ob))
    assert str(info) == 'Reaching goal'
    assert [float(t) for t in env.human_times] == [float(t) for t in r['human_times_before']]
    times = env.get_human_times()
    assert times == [float(t) for t in r['human_times']] and env.global_time == float(r['global_time_after'])
    assert abs(robot.px - float(r['final_robot'][0])) < 1e-6 and abs(env.humans[2].py - float(r['final_humans'][2][1])) < 1e-6


def test_debug_scene_minus_one():
    env, robot = _make()
    ob = env.reset('test', -1)          # crowd_sim.py:286-292
    assert len(ob) == 3 and env.human_num == 3
    assert [(h.px, h.py, h.gx, h.gy) for h in env.humans] == [(0.0, -6.0, 0.0, 5.0), (-5.0, -5.0, -5.0, 5.0), (5.0, -5.0, 5.0, 5.0)]
    for _ in range(5):
        ob, reward, done, info = env.step(robot.act(ob))
        assert env.humans[1].px == -env.humans[2].px and robot.px == 0.0     # mirror symmetry is preserved exactly


def test_mixed_rule_episodes_match_reference():
    """test_sim = 'mixed' through the single-env surface: per case the number of humans, terminal class, step count and
    final robot position of the reference's recorded episodes (the list of humans shrinks / grows per episode like the
    reference's; here that does not raise -- the reference's own step() does)."""
    cases = load_golden('suite_mixed5_invisible')['cases'][:14]
    env, robot = _make(test_sim='mixed')
    code = {'ReachGoal': 2, 'Collision': 3, 'Timeout': 4}
    for c in cases:
        ob = env.reset('test', c['case'])
        assert len(ob) == len(env.humans) == len(c['init']['humans']), c['case']
        _, h0 = scene_arrays(c['init'])
        assert np.abs(np.array([[h.px, h.py, h.gx, h.gy] for h in env.humans]) - h0[:, [0, 1, 4, 5]]).max() < 1e-12
        done, steps = False, 0
        while not done:
            ob, reward, done, info = env.step(robot.act(ob))
            steps += 1
        assert code[type(info).__name__] == c['info'] and steps == c['steps'], c['case']
        r, _ = scene_arrays(c['final'])
        assert abs(robot.px - r[0]) < 1e-9 and abs(robot.py - r[1]) < 1e-9


def test_single_human_training_scenes(oracle):
    """crowd_sim.py:266-267,277-279: a policy with multiagent_training = False (CADRL) gets ONE-human circle-crossing
    scenes in the train / val phases (seed = 2000 + case / 0 + case), the full crowd in the test phase."""
    env, robot = _make(human_num=5, test_sim='square_crossing')
    robot.policy.multiagent_training = False
    for phase, offset, case in (('train', 2000, 7), ('val', 0, 3)):
        ob = env.reset(phase, case)
        assert len(ob) == 1 and len(env.humans) == 1 and env.train_val_sim == 'circle_crossing'
        host = oracle.HostState(1, 1)
        oracle.reset(host, [offset + case], 'circle_crossing')
        assert abs(env.humans[0].px - host.h_pos[0, 0, 0]) < 1e-12 and abs(env.humans[0].py - host.h_pos[0, 0, 1]) < 1e-12
        ob, reward, done, info = env.step(robot.act(ob))
        assert len(ob) == 1
    ob = env.reset('test', 0)
    assert len(ob) == 5 and len(env.humans) == 5
