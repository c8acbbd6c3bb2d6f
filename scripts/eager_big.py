#!/usr/bin/env python
"""One big batch stepped eagerly (for ncu at full occupancy): python scripts/eager_big.py [envs]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from crowdnav_b200.batched import BatchedCrowdSim, default_config
B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
env = BatchedCrowdSim(B); env.configure(default_config(human_num=5)); env.set_robot_policy('orca')
env.reset_seeds(torch.arange(B, dtype=torch.int64) % (2 ** 31) + 5000)
for _ in range(24):          # launches 12.. are mid-episode (the crowd interacts): profile those (ncu -s 14)
    env.step()
torch.cuda.synchronize()
print('done')
