#!/usr/bin/env python
"""BASELINE config 4 stepped eagerly for ncu: python scripts/eager_crowd.py [N] [envs] (profile launch 12: mid-episode)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from crowdnav_b200.batched import BatchedCrowdSim, default_config
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
env = BatchedCrowdSim(B); env.configure(default_config(human_num=N, test_sim='square_crossing', train_val_sim='square_crossing')); env.set_robot_policy('orca')
env.reset_seeds(torch.arange(B, dtype=torch.int64) + 5000, rule='square_crossing')
for _ in range(16):
    env.step()
torch.cuda.synchronize()
print('done')
