#!/usr/bin/env python
"""The bench's step / scene-prefetch sequence with EAGER launches (no CUDA graph), for ncu: same kernels, same arguments,
same rotating batches as bench.py's timed region; ncu serialises launches anyway, so compare SHARES, not absolutes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from crowdnav_b200.batched import BatchedCrowdSim, default_config

N = int(sys.argv[1]) if len(sys.argv) > 1 else 5
rule = sys.argv[2] if len(sys.argv) > 2 else 'circle_crossing'
B, pools, K, PE = 4096, (64 if N <= 5 else 16), 1600, 4     # PE: prefetch on every 4th visit of a batch, like bench.py
envs = []
for p in range(pools):
    env = BatchedCrowdSim(B); env.configure(default_config(human_num=N, test_sim=rule, train_val_sim=rule)); env.set_robot_policy('orca')
    env.k_total = 8 * B
    env.track_episodes(env.k_total, gamma=0.9); env.set_case_queue(p * env.k_total, env.k_total, 'train')
    env.enable_autoreset(rule); env.reset_seeds(rule=rule, use_queue=True); env.prefetch(); envs.append(env)
torch.cuda.synchronize()
for t in range(K):
    env = envs[t % pools]
    env.step()
    if (t // pools) % PE == 0:
        env.prefetch()
torch.cuda.synchronize()
print('done', K)
