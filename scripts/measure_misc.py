#!/usr/bin/env python
"""Secondary measurements for DESIGN.md (run under gpurun): BASELINE configs 1, 3 and 4 on one GPU."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from crowdnav_b200.batched import BatchedCrowdSim, default_config
from crowdnav_b200.explorer import BatchedExplorer
from crowdnav_b200.policy import make_sarl

out = {}


def graph_time(fn, S=12, reps=5, env=None):
    """us per call; with `env` the state is restored before every timed replay so that all replays see the same
    mid-episode scenes (the step kernel's cost depends on how much the crowd interacts)."""
    snap = [getattr(env.state, f).clone() for f in env.state.FIELDS] if env is not None else None
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(S):
            fn()
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        if snap is not None:
            for f, t in zip(env.state.FIELDS, snap):
                getattr(env.state, f).copy_(t)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / S)
    return best * 1e3


# config 4: 4096 envs x 20 humans, square crossing (generic kernel)
for B in (4096, 65536):
    env = BatchedCrowdSim(B); env.configure(default_config(human_num=20, test_sim='square_crossing', train_val_sim='square_crossing'))
    env.set_robot_policy('orca'); env.reset_seeds(torch.arange(B) + 2000, rule='square_crossing')
    for _ in range(8): env.step()
    us = graph_time(env.step, env=env)
    out['cfg4_step_N20_B%d' % B] = {'us_per_launch': round(us, 2), 'env_steps_per_s': round(B / us * 1e6), 'GBps': round(B * 2074 / us / 1e3, 1)}
    del env

# config 3: SARL rollout pieces at 4096 envs x 5 humans
B = 4096
env = BatchedCrowdSim(B); env.configure(default_config(human_num=5)); env.set_robot_policy('external_xy')
env.reset_seeds(torch.arange(B) + 2000)
pol = make_sarl(seed=0); pol.set_device(env.device)
for _ in range(3): a = pol.act_batch(env)
st = torch.empty((B, 81, 5, 13), dtype=torch.float32, device=env.device); rw = torch.empty((B, 81), dtype=torch.float64, device=env.device)
us_look = graph_time(lambda: env.lookahead_pack(pol.actions, out_states=st, out_reward=rw), S=10)
out['cfg3_lookahead_pack'] = {'us_per_launch': round(us_look, 1), 'written_GBps': round(B * 81 * (5 * 13 * 4 + 8) / us_look / 1e3, 1)}
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): a = pol.act_batch(env)
torch.cuda.synchronize(); out['cfg3_sarl_decision_ms'] = round((time.perf_counter() - t0) / 20 * 1e3, 3)
ex = BatchedExplorer(env, pol, gamma=0.9)
torch.cuda.synchronize(); t0 = time.perf_counter()
stats = ex.run_k_episodes(4096, 'train')
torch.cuda.synchronize(); dt = time.perf_counter() - t0
out['cfg3_sarl_rollout'] = {'episodes': 4096, 'env_steps': stats['env_steps'], 'seconds': round(dt, 3), 'env_steps_per_s': round(stats['env_steps'] / dt),
                            'rates': [stats['success_rate'], stats['collision_rate'], stats['timeout_rate']]}

# config 1 on the GPU: the 500 test cases with the ORCA robot, wall clock of run_k_episodes (includes host loop + sync)
env = BatchedCrowdSim(512); env.configure(default_config(human_num=5))
ex = BatchedExplorer(env, 'orca', gamma=0.9)
ex.run_k_episodes(500, 'test'); env.case_counter['test'] = 0
torch.cuda.synchronize(); t0 = time.perf_counter()
stats = ex.run_k_episodes(500, 'test')
torch.cuda.synchronize(); dt = time.perf_counter() - t0
out['cfg1_500_test_cases'] = {'seconds': round(dt, 4), 'env_steps': stats['env_steps'], 'success': stats['success'], 'collision': stats['collision'], 'timeout': stats['timeout']}

# config 1 through the SINGLE-ENV drop-in surface: the literal flow of crowd_nav/test.py --policy orca (gym.make, Robot, ORCA,
# Explorer.run_k_episodes over the 500 test cases): one crowdsim_orca_act + one crowdsim_step launch and one read-back per step
import crowdnav_b200.compat as compat
compat.install()
import gym, logging
from crowd_sim.envs.utils.robot import Robot
from crowd_sim.envs.policy.orca import ORCA
from crowd_nav.utils.explorer import Explorer
cfg = default_config(human_num=5)
env1 = gym.make('CrowdSim-v0'); env1.configure(cfg)
robot = Robot(cfg, 'robot'); pol1 = ORCA(); robot.set_policy(pol1); env1.set_robot(robot)
pol1.set_phase('test'); pol1.set_device(torch.device('cuda:0')); pol1.set_env(env1)
ex1 = Explorer(env1, robot, torch.device('cuda:0'), gamma=0.9)
lines = []
h = logging.Handler(); h.emit = lambda rec: lines.append(rec.getMessage()); logging.getLogger().addHandler(h); logging.getLogger().setLevel(logging.INFO)
t0 = time.perf_counter()
ex1.run_k_episodes(500, 'test')
dt = time.perf_counter() - t0
out['cfg1_compat_test_py_flow'] = {'seconds': round(dt, 2), 'env_steps': 15190, 'env_steps_per_s': round(15190 / dt), 'log': [l for l in lines if 'success rate' in l][:1],
                                   'note': 'reference Python on the C rvo2 shim in the build container: 3.2 s (4.8 k env-steps/s)'}
print(json.dumps(out, indent=1))
