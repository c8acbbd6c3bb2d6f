#!/usr/bin/env python
"""Sweep the small-crowd step kernel's packing (envs per warp) and the generic kernel over batch sizes (run under gpurun).
Each configuration: a CUDA graph of S back-to-back step launches on one batch (state larger than L2 only for the big
batches; the point here is kernel latency/throughput, the bench does the L2-honest measurement), timed with CUDA events.
Prints one JSON line per configuration and a summary table."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from crowdnav_b200 import _abi  # noqa: E402
from crowdnav_b200.batched import BatchedCrowdSim, default_config  # noqa: E402


def time_cfg(B, N, epw, generic, S=20, reps=5):
    lib = _abi.load()
    lib.crowdsim_debug_force_generic(1 if generic else 0)
    lib.crowdsim_debug_force_epw(epw)
    env = BatchedCrowdSim(B)
    env.configure(default_config(human_num=N))
    env.set_robot_policy('orca')
    env.reset_seeds(torch.arange(B, dtype=torch.int64) + 2000)
    for _ in range(3):
        env.step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(S):
            env.step()
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        env.reset_seeds(torch.arange(B, dtype=torch.int64) + 2000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / S)
    lib.crowdsim_debug_force_generic(0); lib.crowdsim_debug_force_epw(0)
    return best * 1e3   # us per launch


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    opts = {1: [1, 4, 16], 2: [1, 3, 10], 3: [1, 2, 8], 4: [1, 2, 6], 5: [1, 2, 3, 5]}[N]
    rows = []
    for B in (1024, 4096, 16384, 65536, 262144, 1048576):
        for epw in opts + ['generic']:
            us = time_cfg(B, N, 0 if epw == 'generic' else epw, epw == 'generic')
            row = {'B': B, 'N': N, 'epw': epw, 'us_per_launch': round(us, 2), 'env_steps_per_s': round(B / us * 1e6)}
            print(json.dumps(row), flush=True)
            rows.append(row)
    print('%8s ' % 'B' + ' '.join('%12s' % str(o) for o in opts + ['generic']))
    for B in sorted({r['B'] for r in rows}):
        print('%8d ' % B + ' '.join('%10.1fus' % next(r['us_per_launch'] for r in rows if r['B'] == B and r['epw'] == o) for o in opts + ['generic']))


if __name__ == '__main__':
    main()
