#!/usr/bin/env python
"""Where does the time of a bench step go? Times CUDA graphs of the rotating-pool loop in several configurations."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from crowdnav_b200.batched import BatchedCrowdSim, default_config

B, N, pools, K = 4096, 5, 64, 256
dev = torch.device('cuda:0')
envs = []
for p in range(pools):
    env = BatchedCrowdSim(B, device=dev); env.configure(default_config(human_num=N)); env.set_robot_policy('orca')
    env.track_episodes(1); env.episodes.ep_case.fill_(-1)
    env.enable_autoreset('circle_crossing', seed_stride=pools * B)
    env.reset_seeds(torch.arange(B, dtype=torch.int64) + 2000 + p * B, seed_stride=pools * B); env.prefetch()
    envs.append(env)
torch.cuda.synchronize()
main = torch.cuda.Stream(); sides = [torch.cuda.Stream() for _ in range(4)]


def run(name, mode, nsides=4):
    g = torch.cuda.CUDAGraph(); pending = {}
    with torch.cuda.graph(g, stream=main):
        for t in range(K):
            p = t % pools; env = envs[p]
            if p in pending: main.wait_event(pending.pop(p))
            if mode == 'step_noar':
                ep, ar = env.episodes, env.autoreset; env.episodes = None; env.autoreset = None; env.step(); env.episodes, env.autoreset = ep, ar
            else:
                env.step()
            if mode == 'same_stream':
                env.prefetch()
            elif mode == 'side':
                ev = torch.cuda.Event(); ev.record(main); sd = sides[p % nsides]; sd.wait_event(ev)
                with torch.cuda.stream(sd):
                    env.prefetch(); d = torch.cuda.Event(); d.record(sd)
                pending[p] = d
            elif mode == 'reset_same_stream':
                env.reset_seeds(mask=env.done, seed_stride=pools * B)
        for ev in pending.values(): main.wait_event(ev)
    with torch.cuda.stream(main):
        g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(main):
            e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1) / K * 1e3)
    print('%-40s %7.2f us/step' % (name, best), flush=True)


run('step only, no ep/ar', 'step_noar')
run('step with ep+ar, no prefetch (envs park)', 'step_ar')
for env in envs:
    env.reset_seeds(seed_stride=pools * B); env.autoreset.n_state.zero_(); env.autoreset.want.zero_(); env.prefetch()
run('step + prefetch, same stream', 'same_stream')
run('step + prefetch on 4 side streams', 'side', 4)
run('step + prefetch on 1 side stream', 'side', 1)
ar_save = [e.autoreset for e in envs]
for e in envs: e.autoreset = None
run('step + masked reset, same stream (old)', 'reset_same_stream')
