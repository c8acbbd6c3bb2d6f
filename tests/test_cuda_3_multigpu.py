"""Multi-GPU path on hardware (BASELINE config 5: envs sharded over the GPUs of one box, one NCCL gather of the episode rows):
two ranks, one GPU each, run the 500 test cases through BatchedExplorer(rank, world=2) over NCCL; the gathered rows on rank 0
must equal the rows a single GPU produces. Needs >= 2 GPUs (gpurun --gpus 2); skipped otherwise."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from crowdnav_b200.batched import BatchedCrowdSim, default_config
from crowdnav_b200.explorer import BatchedExplorer
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
env = BatchedCrowdSim(128, device='cuda:%%d' %% local)
env.configure(default_config(human_num=5))
ex = BatchedExplorer(env, 'orca', gamma=0.9, rank=rank, world=world)
stats = ex.run_k_episodes(500, 'test')
if rank == 0:
    torch.save({'rows': ex.last_rows.cpu(), 'stats': stats}, sys.argv[1])
dist.destroy_process_group()
''' % ROOT


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_two_rank_nccl_gather_matches_single_gpu(tmp_path, cuda_env):
    from crowdnav_b200.explorer import BatchedExplorer
    env = cuda_env(128, 5)
    single = BatchedExplorer(env, 'orca', gamma=0.9)
    stats1 = single.run_k_episodes(500, 'test')
    rows1 = single.last_rows.cpu()
    script, out = str(tmp_path / 'worker.py'), str(tmp_path / 'rows.pt')
    open(script, 'w').write(WORKER)
    port = 29600 + os.getpid() % 1000
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                        '--master-port', str(port), script, out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    got = torch.load(out)
    assert torch.equal(got['rows'], rows1)                         # every per-case row: class, steps, time, return, danger stats
    assert (got['stats']['success'], got['stats']['collision'], got['stats']['timeout']) == (213, 284, 3) == (stats1['success'], stats1['collision'], stats1['timeout'])
    assert got['stats']['timeout_cases'] == [118, 168, 224] and got['stats']['env_steps'] == 15190
