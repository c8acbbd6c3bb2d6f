"""CPU tests of the oracle itself: pins against the committed golden fixtures (produced by the reference's own
Python in the build container, oracle/gen_golden.py), SURVEY.md Appendix B digests and analytic known answers."""
import hashlib
import math
import os
import sys

import numpy as np
import pytest

from util import SUITES, load_golden, scene_arrays, fill_host_state

# SURVEY.md Appendix B (independent throw-away restatement by the surveyor) -- the only external pin there is.
APPENDIX_B = {
    'circle5_invisible': dict(counts=(213, 284, 3), timeouts=[118, 168, 224], steps=15190, sha='ab25dfb557239780'),
    'square5_invisible': dict(counts=(369, 129, 2), timeouts=[192, 472], steps=15740),
    'square20_invisible': dict(counts=(20, 79, 1), timeouts=[61], steps=2408),
    'circle5_visible': dict(counts=(500, 0, 0), timeouts=[], steps=20037),
}


def test_mt19937_matches_numpy(oracle):
    for seed in (0, 1000, 1499, 2000, 2 ** 32 - 2001):
        np.random.seed(seed)
        ref = np.array([np.random.random() for _ in range(1500)])     # crosses the 624-word twist boundary twice
        assert (oracle.mt19937_doubles(seed, 1500) == ref).all()
    np.random.seed(1000)
    assert np.random.random() == 0.6535895854646095 and np.random.random() == 0.11500694312440574   # SURVEY 8c KAT


@pytest.mark.parametrize('name', sorted(APPENDIX_B))
def test_golden_matches_survey_appendix_b(name):
    d = load_golden('suite_' + name)
    exp = APPENDIX_B[name]
    assert (d['counts']['success'], d['counts']['collision'], d['counts']['timeout']) == exp['counts']
    assert [c['case'] for c in d['cases'] if c['info'] == 4] == exp['timeouts']
    assert d['total_env_steps'] == exp['steps']
    if 'sha' in exp:
        coll = ' '.join(str(c['case']) for c in d['cases'] if c['info'] == 3)
        assert hashlib.sha256(coll.encode()).hexdigest()[:16] == exp['sha']
    if name == 'circle5_invisible':
        assert (d['cases'][0]['info'], d['cases'][0]['steps']) == (3, 24)      # case 0: Collision at step 24
        assert (d['cases'][3]['info'], d['cases'][3]['steps']) == (2, 34)      # case 3: ReachGoal at step 34
        assert d['log_lines'][0] == 'TEST  has success rate: 0.43, collision rate: 0.57, nav time: 10.86, total reward: -0.0220'
        assert d['log_lines'][1] == 'Frequency of being in danger: 0.30 and average min separate distance in danger: 0.08'


@pytest.mark.parametrize('name', sorted(SUITES))
def test_batched_c_oracle_reproduces_reference_python(oracle, name):
    """The plain-C batched restatement (reset + step + bookkeeping) is bit-identical to the reference's Python loop."""
    N, rule, vis, rand = SUITES[name]
    d = load_golden('suite_' + name)
    cases = d['cases']
    prm = oracle.default_params(robot_visible=vis)
    ep, st = oracle.run_episodes(prm, N, [1000 + c['case'] for c in cases], rule, randomize_attributes=rand)
    for i, c in enumerate(cases):
        assert ep.res_info[i] == c['info'] and ep.res_steps[i] == c['steps'], c['case']
        assert ep.res_time[i] == (25.0 if c['info'] == 4 else float(c['global_time']))
        assert ep.res_return[i] == float(c['return'])
        assert ep.res_too_close[i] == c['too_close']
        assert ep.res_min_dist_sum[i] == float(c['min_dist_sum'])
        r, h = scene_arrays(c['final'], N)
        assert (ep.res_final_rpos[i] == r[:2]).all()
        assert (st.h_pos[i] == h[:, :2]).all() and (st.h_vel[i] == h[:, 2:4]).all()


def test_reset_scenes_match_reference(oracle):
    d = load_golden('reset_scenes')
    for name, blk in d.items():
        kw = blk['config']
        rows = blk['rows']
        N = kw['human_num']
        st = oracle.HostState(len(rows), N)
        oracle.reset(st, [r['seed'] for r in rows], kw['test_sim'], randomize_attributes=kw.get('randomize', False))
        for e, row in enumerate(rows):
            r, h = scene_arrays(row['scene'], N)
            assert (st.r_pos[e] == r[0:2]).all() and (st.r_goal[e] == r[4:6]).all() and st.r_theta[e] == r[8], name
            assert (st.h_pos[e] == h[:, 0:2]).all() and (st.h_goal[e] == h[:, 4:6]).all(), (name, row['case'])
            assert (st.h_attr[e] == h[:, 6:8]).all(), name


def test_trajectory_steps_match_reference(oracle):
    """Every recorded step of the golden trajectories: pre-state -> one oracle step == recorded post-state."""
    for name in ('circle5_invisible', 'square5_invisible', 'square20_invisible', 'circle5_visible', 'mixed5_invisible'):
        N, rule, vis, _ = SUITES[name]
        d = load_golden('traj_' + name)
        prm = oracle.default_params(robot_visible=vis)
        for case, steps in d['trajectories'].items():
            st = fill_host_state(oracle, [s['pre'] for s in steps], N)
            st.g_time[:] = [float(s['global_time']) - 0.25 for s in steps]
            io = oracle.HostStepIO(len(steps))
            oracle.step(prm, st, io)
            for e, s in enumerate(steps):
                r, h = scene_arrays(s['post'], N)
                assert (io.action_out[e] == [float(x) for x in s['action']]).all()
                assert io.reward[e] == float(s['reward']) and io.done[e] == s['done'] and io.info[e] == s['info']
                if s['dmin'] is not None:
                    assert io.dmin[e] == float(s['dmin'])
                assert (st.r_pos[e] == r[0:2]).all() and (st.h_pos[e] == h[:, 0:2]).all() and (st.h_vel[e] == h[:, 2:4]).all()


def _solve_alone(oracle, pos, goal, v_pref=1.0):
    st = oracle.HostState(1, 0)
    st.r_pos[0] = pos; st.r_goal[0] = goal; st.r_attr[0] = (0.3, v_pref)
    return oracle.orca_act(oracle.default_params(), st)[0]


def test_orca_known_answers(oracle):
    # no neighbours: new velocity = preferred velocity (goal direction, unit speed cap) -- SURVEY 8c (3)
    v = _solve_alone(oracle, (0.0, -4.0), (0.0, 4.0))
    assert v[0] == 0.0 and v[1] == 1.0
    v = _solve_alone(oracle, (0.0, 0.0), (0.3, 0.4))           # closer than 1 m: pref = goal - pos (not normalised)
    assert v[0] == float(np.float32(0.3)) and v[1] == float(np.float32(0.4))
    v = _solve_alone(oracle, (0.0, 0.0), (3.0, 4.0), v_pref=0.5)   # pref has unit length, clipped to maxSpeed = v_pref
    assert abs(v[0] - 0.3) < 1e-6 and abs(v[1] - 0.4) < 1e-6
    # one static human behind the robot: its ORCA half-plane does not cut the preferred velocity
    st = oracle.HostState(1, 1)
    st.r_pos[0] = (0, -4); st.r_goal[0] = (0, 4); st.r_attr[0] = (0.3, 1.0)
    st.h_pos[0, 0] = (3.0, -8.0); st.h_goal[0, 0] = (3.0, -8.0); st.h_attr[0, 0] = (0.3, 1.0)
    v = oracle.orca_act(oracle.default_params(), st)[0]
    assert v[0] == 0.0 and v[1] == 1.0
    # head-on human 1 m ahead, both at rest: robot must deviate, speed stays <= 1, result symmetric under mirroring x
    st.h_pos[0, 0] = (0.0, -3.0)
    v1 = oracle.orca_act(oracle.default_params(), st)[0]
    assert math.hypot(*v1) <= 1.0 + 1e-6 and v1[1] < 1.0
    st.h_pos[0, 0] = (0.2, -3.0)
    va = oracle.orca_act(oracle.default_params(), st)[0]
    st.h_pos[0, 0] = (-0.2, -3.0)
    vb = oracle.orca_act(oracle.default_params(), st)[0]
    assert va[0] == -vb[0] and va[1] == vb[1]


def test_debug_scene_symmetry(oracle):
    """crowd_sim.py:286-292 test_case=-1: three hand-placed humans, mirror symmetric about x = 0."""
    st = oracle.HostState(1, 3)
    st.r_pos[0] = (0, -4); st.r_goal[0] = (0, 4); st.r_attr[0] = (0.3, 1.0); st.r_theta[0] = math.pi / 2
    for i, (p, g) in enumerate([((0, -6), (0, 5)), ((-5, -5), (-5, 5)), ((5, -5), (5, 5))]):
        st.h_pos[0, i] = p; st.h_goal[0, i] = g; st.h_attr[0, i] = (0.3, 1.0)
    io = oracle.HostStepIO(1)
    prm = oracle.default_params()
    for _ in range(10):
        oracle.step(prm, st, io)
        assert st.h_pos[0, 1, 0] == -st.h_pos[0, 2, 0] and st.h_pos[0, 1, 1] == st.h_pos[0, 2, 1]
        assert st.r_pos[0, 0] == 0.0


def test_kdtree_sim_equals_bruteforce(oracle):
    """rvo2 shim (kd-tree, leaf size 10, full doStep) vs the brute-force neighbour scan used by the batched oracle
    and the CUDA kernels, on 21-agent scenes where the tree really splits."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle', 'shims'))
    import rvo2
    rng = np.random.RandomState(7)
    N = 20
    for trial in range(30):
        st = oracle.HostState(1, N)
        st.h_pos[0] = rng.uniform(-5, 5, (N, 2)); st.h_vel[0] = rng.uniform(-1, 1, (N, 2)).astype(np.float32)
        st.h_goal[0] = rng.uniform(-5, 5, (N, 2)); st.h_attr[0] = (0.3, 1.0)
        st.r_pos[0] = (0, -4); st.r_goal[0] = (0, 4); st.r_attr[0] = (0.3, 1.0)
        act = oracle.orca_act(oracle.default_params(), st)[0]
        sim = rvo2.PyRVOSimulator(0.25, 10, 10, 5, 5, 0.3, 1)
        sim.addAgent(tuple(st.r_pos[0]), 10, 10, 5, 5, 0.3 + 0.01 + 0, 1.0, (0.0, 0.0))
        for i in range(N):
            sim.addAgent(tuple(st.h_pos[0, i]), 10, 10, 5, 5, 0.3 + 0.01 + 0, 1, tuple(st.h_vel[0, i]))
        for rep in range(3):          # repeat: the kd-tree's agent permutation persists between doStep calls
            sim.setAgentPosition(0, tuple(st.r_pos[0])); sim.setAgentVelocity(0, (0.0, 0.0))
            for i in range(N):
                sim.setAgentPosition(i + 1, tuple(st.h_pos[0, i])); sim.setAgentVelocity(i + 1, tuple(st.h_vel[0, i]))
                sim.setAgentPrefVelocity(i + 1, (0, 0))
            sim.setAgentPrefVelocity(0, (0.0, 1.0))
            sim.doStep()
            assert sim.getAgentVelocity(0) == (act[0], act[1])


def test_oracle_workload_statistics(oracle):
    """Workload shape quoted in DESIGN.md: lines per solve and LP3 share on cfg1 (4.17 lines, 4.58 %)."""
    d = load_golden('suite_circle5_invisible')
    oracle.lib().oracle_clear_stats()
    ep, _ = oracle.run_episodes(oracle.default_params(), 5, [1000 + c['case'] for c in d['cases']])
    solves, lines, _, lp3 = oracle.get_stats()
    assert solves == 6 * 15190
    assert abs(lines / solves - 4.17) < 0.01
    assert abs(lp3 / solves - 0.0458) < 0.001


@pytest.mark.parametrize('slots', [7, 64, 500])
def test_autoreset_case_queue_reproduces_suite(oracle, slots):
    """Auto-reset protocol + shared case queue (include/crowdsim_b200.h): 500 test cases streamed through `slots` env
    slots (prefetch -> install on termination) give, per case, exactly the reference's episode."""
    N = 5
    cases = load_golden('suite_circle5_invisible')['cases']
    k = len(cases)
    prm = oracle.default_params()
    st = oracle.HostState(slots, N); io = oracle.HostStepIO(slots); ep = oracle.HostEpisodes(slots, k)
    ar = oracle.HostAutoReset(slots, N)
    counter = np.zeros(1, dtype=np.int32)
    q = dict(case_counter=counter, case_total=k, seed_base=1000)
    oracle.reset(st, None, ep=ep, **q)                       # first `slots` cases straight into the live state
    oracle.prefetch(ar, slots, N, **q)
    for it in range(100000):
        if not st.active.any() and not ar.want.any():
            break
        oracle.step(prm, st, io, ep, ar)
        oracle.prefetch(ar, slots, N, **q)
    assert int(counter[0]) >= k and not st.active.any()
    for i, c in enumerate(cases):
        assert ep.res_info[i] == c['info'] and ep.res_steps[i] == c['steps'], c['case']
        assert ep.res_return[i] == float(c['return'])
        r, _ = scene_arrays(c['final'])
        assert (ep.res_final_rpos[i] == r[:2]).all()


@pytest.mark.parametrize('name', ['circle5_invisible', 'square5_invisible', 'circle5_visible'])
def test_python_loop_restatement_reproduces_reference(name):
    """oracle/pyloop.py (the reference's loop structure restated in Python on the rvo2 shim) against the golden suites:
    a third independent restatement, also used as the reference-shaped CPU timing in bench.py --impl reference."""
    import pyloop
    N, rule, vis, _ = SUITES[name]
    cases = load_golden('suite_' + name)['cases'][:40]
    for c in cases:
        info, steps, t, rxy = pyloop.run_episode(1000 + c['case'], N, rule, bool(vis))
        assert (info, steps) == (c['info'], c['steps']), c['case']
        assert t == float(c['global_time'])
        r, _ = scene_arrays(c['final'])
        assert rxy == (r[0], r[1])


@pytest.mark.parametrize('N', [1, 2, 5, 9, 10, 12])
def test_rvo2_shim_vs_batched_oracle_random_crowds(oracle, N):
    """Two code paths of the oracle against each other on tight random crowds (overlapping agents -> the one-time-step
    branch, infeasible LPs -> lp3): the PyRVOSimulator shim (full doStep of every agent, kd-tree) and the batched env
    oracle's per-agent solve must give the robot the same velocity, bit for bit."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle', 'shims'))
    import rvo2
    rng = np.random.RandomState(100 + N)
    for trial in range(60):
        spread = rng.choice([0.8, 2.0, 5.0])
        st = oracle.HostState(1, N)
        st.h_pos[0] = rng.uniform(-spread, spread, (N, 2)); st.h_vel[0] = rng.uniform(-1, 1, (N, 2)).astype(np.float32)
        st.h_attr[0, :, 0] = rng.uniform(0.2, 0.5, N); st.h_attr[0, :, 1] = 1.0
        st.r_pos[0] = rng.uniform(-spread, spread, 2); st.r_vel[0] = rng.uniform(-1, 1, 2).astype(np.float32)
        st.r_goal[0] = rng.uniform(-6, 6, 2); st.r_attr[0] = (rng.uniform(0.2, 0.5), rng.uniform(0.5, 1.5))
        act = oracle.orca_act(oracle.default_params(), st)[0]
        sim = rvo2.PyRVOSimulator(0.25, 10, 10, 5, 5, 0.3, 1)
        sim.addAgent(tuple(st.r_pos[0]), 10, 10, 5, 5, st.r_attr[0, 0] + 0.01 + 0, st.r_attr[0, 1], tuple(st.r_vel[0]))
        for i in range(N):
            sim.addAgent(tuple(st.h_pos[0, i]), 10, 10, 5, 5, st.h_attr[0, i, 0] + 0.01 + 0, 1, tuple(st.h_vel[0, i]))
            sim.setAgentPrefVelocity(i + 1, (0, 0))
        g = st.r_goal[0] - st.r_pos[0]
        speed = np.linalg.norm(g)
        sim.setAgentPrefVelocity(0, tuple(g / speed if speed > 1 else g))
        sim.doStep()
        assert sim.getAgentVelocity(0) == (act[0], act[1]), (N, trial)


def test_occupancy_maps_match_reference(oracle):
    """oracle.occupancy_maps vs the reference's MultiHumanRL.build_occupancy_maps (multi_human_rl.py:109-163) on recorded
    scenes, lookahead states and random crowds, 3 grid configurations x 3 channel modes. The occupancy pattern must be
    identical; mean velocities agree to float32 rounding (the reference sums Python floats, then casts to float32)."""
    rows = load_golden('occupancy_maps')['rows']
    assert len(rows) > 100
    for r in rows:
        h = np.array([[float(v) for v in hh] for hh in r['humans']])
        ref = np.array([[float(v) for v in m] for m in r['maps']], dtype=np.float32)
        got = oracle.occupancy_maps(h[None, :, 0:2], h[None, :, 2:4], r['cell_num'], float(r['cell_size']), r['channels'])[0]
        assert got.shape == ref.shape, r['tag']
        assert np.array_equal(got != 0, ref != 0), r['tag']
        assert np.abs(got - ref).max() <= 1e-6, r['tag']
    assert any(np.array(r['maps'], dtype=np.float64).any() for r in rows)


def test_run_passes_equals_step_plus_reset(oracle):
    """oracle.run_passes (bench.py's CPU arm: n lockstep passes inside one OpenMP region, finished envs re-generated from
    their per-slot seeds) is bit-identical to n x (step; reset(mask = done)), for any thread count."""
    B, N, n = 300, 5, 60
    prm = oracle.default_params()
    def fresh():
        st = oracle.HostState(B, N); io = oracle.HostStepIO(B)
        seeds = (np.arange(B) + 2000).astype(np.uint32)
        oracle.reset(st, seeds, 'circle_crossing', seed_stride=B)
        return st, io, seeds
    a_st, a_io, a_seeds = fresh()
    finished = 0
    for _ in range(n):
        oracle.step(prm, a_st, a_io)
        finished += int(a_io.done.sum())
        oracle.reset(a_st, a_seeds, 'circle_crossing', mask=a_io.done, seed_stride=B)
    assert finished > B                                      # every env finished at least one episode on average
    for threads in (1, 3):
        oracle.set_threads(threads)
        b_st, b_io, b_seeds = fresh()
        oracle.run_passes(prm, b_st, b_io, b_seeds, n // 2, 'circle_crossing', seed_stride=B)
        oracle.run_passes(prm, b_st, b_io, b_seeds, n - n // 2, 'circle_crossing', seed_stride=B)
        for f in ('h_pos', 'h_vel', 'h_goal', 'r_pos', 'r_vel', 'g_time'):
            assert np.array_equal(getattr(a_st, f), getattr(b_st, f)), (threads, f)
        assert np.array_equal(a_seeds, b_seeds) and np.array_equal(a_io.info, b_io.info) and np.array_equal(a_io.reward, b_io.reward)
    oracle.set_threads(os.cpu_count() or 1)
