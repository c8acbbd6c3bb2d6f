#!/usr/bin/env python
"""The bench's launch sequence with EAGER launches (no CUDA graph), for ncu: the same kernels with the same arguments as
bench.py's timed region -- per batch one crowdsim_step_n launch of C env-steps and one scene-prefetch launch, batches rotating.
ncu serialises launches anyway, so compare SHARES of the GPU time, not absolutes. python scripts/eager_loop.py [N] [rule] [C]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from crowdnav_b200.batched import BatchedCrowdSim, default_config

N = int(sys.argv[1]) if len(sys.argv) > 1 else 5
rule = sys.argv[2] if len(sys.argv) > 2 else 'circle_crossing'
C = int(sys.argv[3]) if len(sys.argv) > 3 else 16
B, pools, rounds = 4096, (64 if N <= 5 else 16), 28      # (profiles/r02_launch_list_* and r02_step_n_4096envs_* were captured with C = 8)
envs = []
for p in range(pools):
    env = BatchedCrowdSim(B); env.configure(default_config(human_num=N, test_sim=rule, train_val_sim=rule)); env.set_robot_policy('orca')
    env.k_total = 64 * B
    env.track_episodes(env.k_total, gamma=0.9); env.set_case_queue(p * env.k_total, env.k_total, 'train')
    env.enable_autoreset(rule); env.reset_seeds(rule=rule, use_queue=True); env.prefetch(); envs.append(env)
torch.cuda.synchronize()
for r in range(rounds):                       # rounds 24.. are in steady state (>= 192 steps per batch): profile those
    for env in envs:
        env.step_n(C) if C > 1 else env.step()
        env.prefetch()
torch.cuda.synchronize()
print('done', rounds * pools)
