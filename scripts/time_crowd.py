#!/usr/bin/env python
"""BASELINE config 4 (4096 envs x 20 humans, square_crossing) step kernel: us per launch, mid-episode, crowd kernel vs the
generic kernel of round 1 (crowdsim_debug_force_generic): python scripts/time_crowd.py [N] [envs]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from crowdnav_b200 import _abi
from crowdnav_b200.batched import BatchedCrowdSim, default_config
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
lib = _abi.load()
for B in ([int(sys.argv[2])] if len(sys.argv) > 2 else [4096, 65536]):
    for generic in (1, 0):
        lib.crowdsim_debug_force_generic(generic)
        env = BatchedCrowdSim(B); env.configure(default_config(human_num=N, test_sim='square_crossing', train_val_sim='square_crossing')); env.set_robot_policy('orca')
        env.reset_seeds(torch.arange(B, dtype=torch.int64) + 5000, rule='square_crossing')
        for _ in range(10):
            env.step()
        snap = {f: getattr(env.state, f).clone() for f in env.state.FIELDS}
        st = torch.cuda.Stream()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(8):
                env.step()
        best = 1e9
        for r in range(5):
            for f, t in snap.items():
                getattr(env.state, f).copy_(t)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(st):
                e0.record(); g.replay(); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 8 * 1e3)
        bytes_ = B * (8 * (19 + 12 * N) + 2)
        print('N=%d B=%d %s: %.2f us per launch, %.1f M env-steps/s, %.1f GB/s algorithmic = %.4f of 6560.6' % (
            N, B, 'generic kernel (round 1)' if generic else 'crowd kernel (step_mid)', best, B / best, bytes_ / best / 1e3, bytes_ / best / 1e3 / 6560.6))
lib.crowdsim_debug_force_generic(0)
