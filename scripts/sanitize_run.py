#!/usr/bin/env python
"""Small run through every kernel of the library for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):
   compute-sanitizer --tool racecheck python scripts/sanitize_run.py
Covers: small-crowd step kernel (per-warp and per-block lp3 queue), multi-step kernel with auto-reset and a CONCURRENT scene
prefetch on a side stream (the release / acquire slot hand-over), crowd kernel (N = 12), generic kernel, scene generation,
lookahead pack / humans / onestep_lookahead, occupancy maps, human_times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from crowdnav_b200 import _abi
from crowdnav_b200.batched import BatchedCrowdSim, default_config
lib = _abi.load()


def make(B, N, policy='orca', rule='circle_crossing'):
    env = BatchedCrowdSim(B); env.configure(default_config(human_num=N, test_sim=rule, train_val_sim=rule)); env.set_robot_policy(policy)
    return env

# tight scenes (overlaps -> lp3): squeeze the circle
for B, N in ((300, 5), (3000, 5), (200, 3), (150, 12)):          # 3000 envs x 6 lanes: per-block lp3 queue; 12 humans: crowd kernel
    env = make(B, N)
    env.circle_radius = 1.5 if N <= 5 else 4.0
    env.reset_seeds(torch.arange(B) + 2000)
    for _ in range(10):
        env.step()
lib.crowdsim_debug_force_generic(1)
env = make(200, 5); env.reset_seeds(torch.arange(200) + 2000)
for _ in range(6):
    env.step()
lib.crowdsim_debug_force_generic(0)

# multi-step kernel + auto-reset + concurrent generator
env = make(256, 5)
ep = env.track_episodes(4000); env.set_case_queue(0, 4000, 'train'); env.enable_autoreset(); env.reset_seeds(use_queue=True); env.prefetch()
side = torch.cuda.Stream()
for it in range(40):
    with torch.cuda.stream(side):
        env.prefetch()                      # NOT ordered against the steps
    env.step_n(4)
torch.cuda.synchronize()
print('episodes finished', int((ep.res_steps > 0).sum()))

# value-network support + lookahead + human times
env = make(64, 5, policy='external_xy'); env.reset_seeds(torch.arange(64) + 1000)
acts = torch.tensor([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]], dtype=torch.float64, device=env.device)
env.lookahead_pack(acts); env.lookahead_humans(); env.pack_joint(); env.occupancy_maps()
env.onestep_lookahead(torch.zeros((64, 2), dtype=torch.float64, device=env.device))
env.human_times(max_steps=60)
env = make(16, 20, policy='external_xy', rule='square_crossing'); env.reset_seeds(torch.arange(16) + 1000, rule='square_crossing')
env.lookahead_pack(acts); env.human_times(max_steps=20)
torch.cuda.synchronize()
print('sanitize run done, launches:', lib.crowdsim_launch_count())
