import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a real B200 (run with -m gpu under gpurun)')


@pytest.fixture(scope='session')
def oracle():
    """CPU oracle front-end (oracle/pyoracle.py); builds oracle/_build/*.so with gcc if missing."""
    import build as oracle_build
    oracle_build.build()
    import pyoracle
    return pyoracle


@pytest.fixture(scope='session')
def cuda_env():
    """Factory for BatchedCrowdSim on cuda:0; fails loudly if the CUDA library is missing."""
    import torch
    assert torch.cuda.is_available(), 'gpu-marked test without a GPU'
    from crowdnav_b200 import _abi, build as cuda_build
    cuda_build.build()              # (re)build in-tree with nvcc if the library is missing or stale; never a CPU fallback
    _abi.load()
    from crowdnav_b200.batched import BatchedCrowdSim, default_config

    def make(B, N=5, test_sim='circle_crossing', robot_visible=False, robot_policy='orca', randomize=False):
        env = BatchedCrowdSim(B)
        env.configure(default_config(human_num=N, test_sim=test_sim, robot_visible=robot_visible,
                                     randomize_attributes=randomize))
        env.set_robot_policy(robot_policy)
        return env
    return make
