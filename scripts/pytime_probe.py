import os, sys, torch
sys.path.insert(0, '/root/repo')
from crowdnav_b200.batched import BatchedCrowdSim, default_config
B=4096
env = BatchedCrowdSim(B); env.configure(default_config(human_num=5)); env.set_robot_policy('orca')
env.reset_seeds(torch.arange(B, dtype=torch.int64) + 2000)
env.enable_autoreset('circle_crossing', seed_stride=B); env.prefetch()
for _ in range(60): env.step(); env.prefetch()
torch.cuda.synchronize()
for S in (50,):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(S): env.step(); env.prefetch()
    g.replay(); torch.cuda.synchronize()
    best=1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); best=min(best, e0.elapsed_time(e1)/S*1e3)
    print('python/torch graph of %d x (step + prefetch, same stream, auto-reset keeps states mid-episode): %.2f us per pair' % (S, best))
ar = env.autoreset; env.autoreset = None
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(8): env.step()
g.replay(); torch.cuda.synchronize(); env.autoreset = ar
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
print('step only on mid-episode states (8 launches): %.2f us per launch' % (e0.elapsed_time(e1) / 8 * 1e3))
