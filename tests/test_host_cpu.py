"""CPU tests of the host-side logic: BatchedExplorer's reductions / log lines / sharding / the one collective (gloo,
world_size 2), and the value-network ports (weights, action space, greedy decision) against reference fixtures."""
import os
import sys

import numpy as np
import pytest
import torch

from util import SUITES, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rows_from_golden(cases):
    rows = [[c['info'], c['steps'], 25.0 if c['info'] == 4 else float(c['global_time']), float(c['return']),
             c['too_close'], float(c['min_dist_sum'])] for c in cases]
    return torch.tensor(rows, dtype=torch.float64)


@pytest.mark.parametrize('name', [n for n in sorted(SUITES) if n != 'circle5_random_attr'])
def test_summarize_emits_reference_log_lines(name):
    """explorer.py:74-90: from per-case rows, the exact lines the reference's own Explorer printed for the same cases."""
    from crowdnav_b200.explorer import summarize
    d = load_golden('suite_' + name)
    lines = []
    stats = summarize(_rows_from_golden(d['cases']), len(d['cases']), 'test', 25, 0.25, print_failure=True, log=lines.append)
    assert lines == d['log_lines']
    assert stats['success'] == d['counts']['success'] and stats['collision'] == d['counts']['collision']
    assert stats['env_steps'] == d['total_env_steps']


def test_shard_range_partitions():
    from crowdnav_b200.explorer import shard_range
    for k in (1, 7, 500, 131072, 131075):
        for world in (1, 2, 3, 8):
            spans = [shard_range(k, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(n for _, n in spans) == k
            for (s0, n0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + n0 == s1
            assert max(n for _, n in spans) - min(n for _, n in spans) <= 1


def _gloo_worker(rank, world, port, k, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from crowdnav_b200.explorer import shard_range, gather_results, summarize
    from util import load_golden as lg
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    cases = lg('suite_circle5_invisible')['cases'][:k]
    rows = _rows_from_golden(cases)
    start, n = shard_range(k, rank, world)
    full = gather_results(rows[start:start + n].clone(), k, rank, world)
    assert torch.equal(full, rows)
    if rank == 0:
        lines = []
        summarize(full, k, 'test', 25, 0.25, print_failure=True, log=lines.append)
        with open(os.path.join(out_dir, 'lines.txt'), 'w') as f:
            f.write('\n'.join(lines))
    dist.destroy_process_group()


@pytest.mark.parametrize('k', [500, 333])
def test_two_rank_gather_gloo(tmp_path, k):
    """The N > 1 path on CPU: 2 processes, contiguous case shards (uneven for k = 333), one all_gather of the result rows,
    rank 0 prints the same lines as a single process."""
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() + k) % 2000
    mp.spawn(_gloo_worker, args=(2, port, k, str(tmp_path)), nprocs=2, join=True)
    from crowdnav_b200.explorer import summarize
    cases = load_golden('suite_circle5_invisible')['cases'][:k]
    lines = []
    summarize(_rows_from_golden(cases), k, 'test', 25, 0.25, print_failure=True, log=lines.append)
    assert open(os.path.join(str(tmp_path), 'lines.txt')).read().split('\n') == lines


def test_action_space_matches_reference():
    from crowdnav_b200.policy import build_action_space
    d = load_golden('rotate_lookahead')
    ref = np.array([[float(x) for x in a] for a in d['action_space']])
    assert np.array_equal(build_action_space(1.0), ref)
    assert ref.shape == (81, 2)


def test_sarl_network_port_matches_reference_values():
    """Same construction order => same seed-0 initial weights as the reference's ValueNetwork (sarl.py:9-27); values of
    the reference's own rotated lookahead states agree to float32 round-off, and so does the greedy decision."""
    from crowdnav_b200.policy import make_sarl
    d = load_golden('rotate_lookahead')
    pol = make_sarl(gamma=d['gamma'], seed=d['sarl_seed'])
    model = pol.get_model()
    disc = pow(d['gamma'], 0.25 * 1.0)
    worst = 0.0
    for row in d['rows']:
        states = torch.tensor([[[float(v) for v in r] for r in la['rotated']] for la in row['lookahead']], dtype=torch.float32)
        with torch.no_grad():
            v = model(states)[:, 0].double().numpy()
        ref_v = np.array([float(la['value']) for la in row['lookahead']])
        worst = max(worst, float(np.abs(v - ref_v).max()))
        total = np.array([float(la['reward']) for la in row['lookahead']]) + disc * v
        best = int(np.argmax(total))
        top2 = np.sort(total)[-2:]
        if top2[1] - top2[0] > 1e-5:
            assert [float(x) for x in row['lookahead'][best]['action']] == [float(x) for x in row['sarl_action']]
    assert worst < 1e-6


def test_sarl_state_dict_keys_match_reference_layout():
    from crowdnav_b200.policy import SARLValueNetwork, CADRLValueNetwork
    keys = set(SARLValueNetwork().state_dict().keys())
    assert {'mlp1.0.weight', 'mlp1.2.weight', 'mlp2.0.weight', 'mlp2.2.weight', 'attention.0.weight', 'attention.2.weight',
            'attention.4.weight', 'mlp3.0.weight', 'mlp3.6.weight'} <= keys
    assert SARLValueNetwork().mlp1[0].in_features == 13 and SARLValueNetwork().attention[0].in_features == 200
    assert SARLValueNetwork().mlp3[0].in_features == 56
    assert 'value_network.0.weight' in CADRLValueNetwork().state_dict()


def test_compat_types_and_module_aliases():
    """crowdnav_b200.compat.install(): the reference's import paths resolve; value types keep the reference's contract."""
    import crowdnav_b200.compat as compat
    compat.install(force_gym_shim=True)
    import gym
    from crowd_sim.envs.utils.state import FullState, ObservableState, JointState
    from crowd_sim.envs.utils.action import ActionXY, ActionRot
    from crowd_sim.envs.utils.info import Timeout, ReachGoal, Danger, Collision, Nothing
    from crowd_sim.envs.policy.policy_factory import policy_factory
    from crowd_sim.envs.utils.robot import Robot  # noqa: F401
    from crowd_nav.utils.explorer import Explorer, average  # noqa: F401
    fs = FullState(1, 2, 3, 4, 0.3, 5, 6, 1.0, 0.5); ob = ObservableState(7, 8, 9, 10, 0.4)
    assert fs + ob == (1, 2, 3, 4, 0.3, 5, 6, 1.0, 0.5, 7, 8, 9, 10, 0.4)          # state.py:17-18,36-37 -> the 14-tuple
    assert not isinstance(fs, ObservableState) and fs.position == (1, 2) and fs.goal_position == (5, 6) and ob.velocity == (9, 10)
    JointState(fs, [ob])
    with pytest.raises(AssertionError):
        JointState(ob, [ob])
    assert (str(Timeout()), str(ReachGoal()), str(Danger(0.1)), str(Collision()), str(Nothing())) == \
        ('Timeout', 'Reaching goal', 'Too close', 'Collision', '')
    assert Danger(0.05).min_dist == 0.05 and ActionXY(1, 2).vx == 1 and ActionRot(1, 2).r == 2
    assert set(policy_factory) == {'linear', 'orca', 'none'} and policy_factory['none']() is None
    assert average([]) == 0 and average([1, 2]) == 1.5
    assert type(gym.make('CrowdSim-v0')).__name__ == 'CrowdSim'
