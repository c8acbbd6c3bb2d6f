"""GPU parity tests: the CUDA library (through its C ABI, crowdnav_b200/_abi.py) against the CPU oracle and the
committed golden fixtures of the reference. Bar: bit-exact for flags / integer fields AND (in practice) for every
float64 state array, because both sides evaluate the same individually-rounded operations; the only tolerated
differences are CUDA's double cos/sin in scenario generation (<= 4 ulp on initial coordinates) and the float32
atan2f/cosf/sinf of the rotate rows (1e-5)."""
import numpy as np
import pytest
import torch

from util import SUITES, load_golden, scene_arrays, fill_host_state

pytestmark = pytest.mark.gpu

STATE_FIELDS = ('h_pos', 'h_vel', 'h_goal', 'h_attr', 'r_pos', 'r_vel', 'r_goal', 'r_attr', 'g_time')


def _assert_state_equal(env, host, fields=STATE_FIELDS, what=''):
    dev = env.state.to_host()
    for f in fields:
        a, b = dev[f], getattr(host, f)
        assert np.array_equal(a, b), '%s: field %s differs in %d entries (max abs %.3g)' % (
            what, f, int((a != b).sum()), float(np.abs(a - b).max()))


def _assert_io_equal(env, io, what=''):
    assert np.array_equal(env.done.cpu().numpy(), io.done), what + ' done'
    assert np.array_equal(env.info.cpu().numpy(), io.info), what + ' info'
    assert np.array_equal(env.reward.cpu().numpy(), io.reward), what + ' reward'
    assert np.array_equal(env.dmin.cpu().numpy(), io.dmin), what + ' dmin'
    assert np.array_equal(env.action_out.cpu().numpy(), io.action_out), what + ' action_out'


def _random_host_state(oracle, B, N, seed, spread=4.5):
    rng = np.random.RandomState(seed)
    st = oracle.HostState(B, N)
    st.h_pos[...] = rng.uniform(-spread, spread, (B, N, 2))
    st.h_vel[...] = rng.uniform(-1, 1, (B, N, 2)).astype(np.float32)          # velocities are float32-valued actions
    st.h_goal[...] = rng.uniform(-spread, spread, (B, N, 2))
    st.h_attr[..., 0] = rng.uniform(0.2, 0.5, (B, N)); st.h_attr[..., 1] = rng.uniform(0.5, 1.5, (B, N))
    st.r_pos[...] = rng.uniform(-spread, spread, (B, 2)); st.r_vel[...] = rng.uniform(-1, 1, (B, 2)).astype(np.float32)
    st.r_goal[...] = rng.uniform(-spread, spread, (B, 2))
    st.r_attr[:, 0] = rng.uniform(0.2, 0.5, B); st.r_attr[:, 1] = rng.uniform(0.5, 1.5, B)
    st.r_theta[...] = rng.uniform(0, 2 * np.pi, B)
    st.g_time[...] = 0.25 * rng.randint(0, 99, B)
    return st


@pytest.mark.parametrize('name', ['circle5_invisible', 'square5_invisible', 'square20_invisible', 'circle5_visible', 'mixed5_invisible'])
def test_step_reproduces_golden_trajectories(cuda_env, oracle, name):
    """Each recorded reference step (pre-state, action, reward, info, post-state) is reproduced bit-exactly."""
    N, rule, vis, _ = SUITES[name]
    d = load_golden('traj_' + name)
    for case, steps in d['trajectories'].items():
        host = fill_host_state(oracle, [s['pre'] for s in steps], N)
        host.g_time[:] = [float(s['global_time']) - 0.25 for s in steps]
        env = cuda_env(len(steps), N, rule, robot_visible=bool(vis))
        env.state.load_host(host)
        env.step()
        torch.cuda.synchronize()
        dev = env.state.to_host()
        for e, s in enumerate(steps):
            r, h = scene_arrays(s['post'], N)
            assert (env.action_out[e].cpu().numpy() == [float(x) for x in s['action']]).all(), (name, case, e)
            assert float(env.reward[e]) == float(s['reward']) and int(env.done[e]) == int(s['done']) and int(env.info[e]) == s['info']
            if s['dmin'] is not None:
                assert float(env.dmin[e]) == float(s['dmin'])
            assert (dev['r_pos'][e] == r[0:2]).all() and (dev['h_pos'][e] == h[:, 0:2]).all() and (dev['h_vel'][e] == h[:, 2:4]).all()


@pytest.fixture(autouse=True)
def _default_kernel_routing():
    """Every test starts with the default routing (N <= 5 -> warp-cooperative kernel)."""
    from crowdnav_b200 import _abi
    _abi.load().crowdsim_debug_force_generic(0)
    yield
    _abi.load().crowdsim_debug_force_generic(0)


@pytest.mark.parametrize('N,vis,policy,generic', [
    (5, 0, 'orca', 0), (5, 1, 'orca', 0), (5, 0, 'external_xy', 0), (5, 1, 'external_xy', 0), (4, 1, 'orca', 0), (3, 0, 'orca', 0),
    (2, 1, 'external_xy', 0), (2, 0, 'orca', 0), (1, 0, 'orca', 0), (1, 1, 'orca', 0), (5, 0, 'external_rot', 0),
    (5, 0, 'orca', 1), (5, 1, 'orca', 1), (1, 0, 'orca', 1), (2, 1, 'external_xy', 1), (5, 0, 'external_rot', 1),
    (10, 1, 'orca', 0), (11, 0, 'orca', 0), (20, 0, 'orca', 0), (20, 1, 'external_xy', 0), (33, 1, 'orca', 0), (63, 1, 'orca', 0),
    (0, 0, 'orca', 0), (6, 1, 'orca', 0), (7, 0, 'orca', 0), (9, 1, 'external_xy', 0), (20, 1, 'orca', 0),
    (20, 0, 'orca', 1), (11, 1, 'orca', 1), (6, 0, 'external_xy', 1)])
def test_step_random_scenes_bit_exact(cuda_env, oracle, N, vis, policy, generic):
    """Dense random scenes (many overlapping agents -> collision branch, lp3 fallback, 10-of-N truncation),
    8 consecutive steps, every state/output array compared for equality with the CPU oracle. Routing: N <= 5 small-crowd
    kernel, N > 5 crowd kernel (step_mid.cuh); generic = 1 forces the round-1 generic kernel (A/B partner of both)."""
    B = 1500 if N <= 20 else 300
    host = _random_host_state(oracle, B, N, seed=100 + N)
    env = cuda_env(B, N, robot_visible=bool(vis), robot_policy=policy)
    env.state.load_host(host)
    from crowdnav_b200 import _abi
    _abi.load().crowdsim_debug_force_generic(generic)
    prm = oracle.default_params(robot_visible=vis, robot_policy={'orca': _abi.ROBOT_ORCA, 'external_xy': _abi.ROBOT_EXTERNAL_XY,
                                                                 'external_rot': _abi.ROBOT_EXTERNAL_ROT}[policy])
    io = oracle.HostStepIO(B)
    rng = np.random.RandomState(5)
    for t in range(8):
        if policy == 'external_rot':
            io.action[:, 0] = rng.uniform(0, 1, B); io.action[:, 1] = rng.uniform(-0.8, 0.8, B)
        else:
            io.action[...] = rng.uniform(-1, 1, (B, 2))
        act = torch.from_numpy(io.action).to(env.device)
        env.step(None if policy == 'orca' else act)
        oracle.step(prm, host, io)
        torch.cuda.synchronize()
        if policy == 'external_rot':
            # CUDA's double cos/sin are not glibc's: compare with a tolerance, then resynchronise the states
            dev = env.state.to_host()
            for f in ('h_pos', 'h_vel'):
                assert np.array_equal(dev[f], getattr(host, f))
            assert np.allclose(dev['r_pos'], host.r_pos, rtol=0, atol=1e-12) and np.allclose(dev['r_theta'], host.r_theta, rtol=0, atol=1e-12)
            assert np.array_equal(env.info.cpu().numpy(), io.info)
            env.state.load_host(host)
        else:
            _assert_state_equal(env, host, what='N=%d step %d' % (N, t))
            _assert_io_equal(env, io, what='N=%d step %d' % (N, t))


@pytest.mark.parametrize('N,vis', [(5, 0), (5, 1), (3, 1), (2, 0)])
def test_step_random_scenes_full_chip_grid(cuda_env, oracle, N, vis):
    """Launches of more than 3 blocks per SM take the block-compacted linearProgram3 queue of the small-crowd kernel
    (smaller ones the per-warp queue, step_kernel.cu: launch): same dense random scenes, 20 000 envs, 6 steps, equality."""
    B = 20000
    host = _random_host_state(oracle, B, N, seed=900 + N)
    env = cuda_env(B, N, robot_visible=bool(vis), robot_policy='orca')
    env.state.load_host(host)
    prm = oracle.default_params(robot_visible=vis)
    io = oracle.HostStepIO(B)
    for t in range(6):
        env.step()
        oracle.step(prm, host, io)
        torch.cuda.synchronize()
        _assert_state_equal(env, host, what='N=%d step %d' % (N, t))
        _assert_io_equal(env, io, what='N=%d step %d' % (N, t))


@pytest.mark.parametrize('generic', [0, 1])
@pytest.mark.parametrize('name', sorted(SUITES))
def test_full_suites_from_reference_scenes(cuda_env, oracle, name, generic):
    """Whole episodes on the GPU from the reference's own initial scenes: terminal class, step count, time, discounted
    return, danger statistics and final positions identical to the reference's Python for every test case."""
    N, rule, vis, rand = SUITES[name]
    cases = load_golden('suite_' + name)['cases']
    B = len(cases)
    host = fill_host_state(oracle, [c['init'] for c in cases], N)
    from crowdnav_b200 import _abi
    _abi.load().crowdsim_debug_force_generic(generic)
    env = cuda_env(B, N, rule, robot_visible=bool(vis))
    ep = env.track_episodes(B)
    env.state.load_host(host)
    ep.ep_case.copy_(torch.arange(B, dtype=torch.int32))
    for _ in range(110):
        env.step()
    torch.cuda.synchronize()
    assert int(env.state.active.sum()) == 0
    info = ep.res_info.cpu().numpy(); steps = ep.res_steps.cpu().numpy(); t = ep.res_time.cpu().numpy()
    ret = ep.res_return.cpu().numpy(); tc = ep.res_too_close.cpu().numpy(); mds = ep.res_min_dist_sum.cpu().numpy()
    frp = ep.res_final_rpos.cpu().numpy(); hp = env.state.h_pos.cpu().numpy()
    for i, c in enumerate(cases):
        assert info[i] == c['info'] and steps[i] == c['steps'], (name, c['case'])
        assert t[i] == (25.0 if c['info'] == 4 else float(c['global_time']))
        assert ret[i] == float(c['return']) and tc[i] == c['too_close'] and mds[i] == float(c['min_dist_sum'])
        r, h = scene_arrays(c['final'], N)
        assert (frp[i] == r[:2]).all() and (hp[i] == h[:, :2]).all()


@pytest.mark.parametrize('name', ['circle5_invisible', 'square5_invisible', 'square20_invisible', 'circle5_visible', 'mixed5_invisible'])
def test_full_suites_device_reset(cuda_env, name):
    """Same, but with scenes generated ON DEVICE from the case seeds (crowdsim_reset). Flags bit-exact, positions
    within 1e-5 (north_star bar; CUDA cos/sin may move initial coordinates by an ulp)."""
    N, rule, vis, _ = SUITES[name]
    cases = load_golden('suite_' + name)['cases']
    B = len(cases)
    env = cuda_env(B, N, rule, robot_visible=bool(vis))
    ep = env.track_episodes(B)
    env.reset('test', cases=[c['case'] for c in cases])
    ep.ep_case.copy_(torch.arange(B, dtype=torch.int32))
    for _ in range(110):
        env.step()
    torch.cuda.synchronize()
    info = ep.res_info.cpu().numpy(); steps = ep.res_steps.cpu().numpy(); frp = ep.res_final_rpos.cpu().numpy()
    assert [int(x) for x in info] == [c['info'] for c in cases]
    assert [int(x) for x in steps] == [c['steps'] for c in cases]
    fr = np.array([scene_arrays(c['final'])[0][:2] for c in cases])
    assert np.abs(frp - fr).max() < 1e-5


def test_reset_matches_oracle(cuda_env, oracle):
    worst = 0
    for N, rule, rand in [(5, 'circle_crossing', False), (5, 'square_crossing', False), (20, 'square_crossing', False),
                          (5, 'circle_crossing', True), (10, 'circle_crossing', False), (5, 'square_crossing', True),
                          (5, 'mixed', False), (5, 'mixed', True), (7, 'mixed', False)]:
        # NB: keep the packing feasible -- 10 humans with random radii up to 0.5 cannot all keep 1.2 m from each
        # other's starts AND goals on the r = 4 circle; the reference's rejection sampling would spin forever too.
        B = 2000
        seeds = np.concatenate([np.arange(1000, 1500), np.arange(2000, 3000), [0, 1, 99, 4294967295, 4294965295],
                                np.random.RandomState(1).randint(0, 2 ** 32, B - 1505, dtype=np.uint64)]).astype(np.uint64)
        host = oracle.HostState(B, N)
        oracle.reset(host, seeds.astype(np.uint32), rule, randomize_attributes=rand)
        env = cuda_env(B, N, rule, randomize=rand)
        env.reset_seeds(torch.from_numpy(seeds.astype(np.int64)), rule=rule)
        torch.cuda.synchronize()
        dev = env.state.to_host()
        for f in ('h_attr', 'r_pos', 'r_goal', 'r_attr', 'r_vel', 'h_vel', 'g_time', 'r_theta'):
            assert np.array_equal(dev[f], getattr(host, f)), f
        for f in ('h_pos', 'h_goal'):
            # px = 4*cos(angle) + noise: CUDA's cos/sin are within 1-2 ulp of glibc's, i.e. <= ~2e-15 absolute on
            # |4 cos| <= 4 (cancellation against the noise term makes a relative/ulp bound meaningless)
            d = np.abs(dev[f] - getattr(host, f)).max()
            worst = max(worst, float(d))
            assert d <= 4e-15, (N, rule, f, d)
            frac_exact = float((dev[f] == getattr(host, f)).mean())
            print('%s N=%d %s: %.1f%% of coordinates bit-identical, max abs diff %.2e' % (rule, N, f, 100 * frac_exact, d))
        if rule == 'square_crossing':        # no cos/sin on this path: bit-exact
            assert np.array_equal(dev['h_pos'], host.h_pos) and np.array_equal(dev['h_goal'], host.h_goal)
    print('worst abs difference of initial coordinates:', worst)


def test_mixed_rule_counts(cuda_env):
    """Rule `mixed`: per-scene human count follows the reference's distribution (crowd_sim.py:105-106; a static scene with
    zero humans holds one dummy) -- checked against the fixture's counts for the same seeds."""
    cases = load_golden('suite_mixed5_invisible')['cases']
    env = cuda_env(len(cases), 5, 'mixed')
    env.reset('test', cases=[c['case'] for c in cases])
    torch.cuda.synchronize()
    assert env.human_counts().cpu().tolist() == [len(c['init']['humans']) for c in cases]


def test_reset_mask_and_active(cuda_env, oracle):
    B, N = 300, 5
    env = cuda_env(B, N)
    env.reset_seeds(torch.arange(B) + 2000)
    before = env.state.to_host()
    mask = (np.arange(B) % 3 == 0).astype(np.uint8)
    env.state.active.zero_()
    env.reset_seeds(torch.arange(B) + 5000, mask=torch.from_numpy(mask))
    after = env.state.to_host()
    assert np.array_equal(after['h_pos'][mask == 0], before['h_pos'][mask == 0])
    assert not np.array_equal(after['h_pos'][mask == 1], before['h_pos'][mask == 1])
    assert np.array_equal(after['active'], mask)
    # frozen envs are not touched by a step
    snap = env.state.to_host()
    env.step()
    torch.cuda.synchronize()
    now = env.state.to_host()
    for f in STATE_FIELDS:
        assert np.array_equal(now[f][mask == 0], snap[f][mask == 0]), f
    assert (now['g_time'][mask == 1] == 0.25).all()


def test_orca_act_matches_oracle(cuda_env, oracle):
    for N, vis in [(5, 0), (20, 1)]:
        B = 1000
        host = _random_host_state(oracle, B, N, seed=3)
        env = cuda_env(B, N, robot_visible=bool(vis), robot_policy='external_xy')
        env.state.load_host(host)
        snap = env.state.to_host()
        act = env.orca_act().cpu().numpy()
        ref = oracle.orca_act(oracle.default_params(robot_visible=vis), host)
        assert np.array_equal(act, ref)
        now = env.state.to_host()
        for f in STATE_FIELDS:
            assert np.array_equal(now[f], snap[f])


def test_pack_and_lookahead_match_oracle_and_reference(cuda_env, oracle):
    d = load_golden('rotate_lookahead')
    actions = np.array([[float(x) for x in a] for a in d['action_space']])
    rows = d['rows']
    N = 5
    host = fill_host_state(oracle, [r['scene'] for r in rows], N)
    host.g_time[:] = [float(r['global_time']) for r in rows]
    env = cuda_env(len(rows), N, robot_policy='external_xy')
    env.state.load_host(host)
    packed = env.pack_joint().cpu().numpy()
    states, reward = env.lookahead_pack(torch.from_numpy(actions).to(env.device))
    states = states.cpu().numpy(); reward = reward.cpu().numpy()
    o_packed = oracle.pack_joint(host)
    o_states, o_reward = oracle.lookahead_pack(oracle.default_params(robot_policy=0), host, actions)
    assert np.array_equal(reward, o_reward)
    assert np.abs(packed - o_packed).max() < 1e-5 and np.abs(states - o_states).max() < 1e-5
    for e, r in enumerate(rows):      # the reference's own torch rotate / onestep_lookahead outputs
        ref_cur = np.array([[float(v) for v in row] for row in r['rotated_current']], dtype=np.float32)
        assert np.abs(packed[e] - ref_cur).max() < 1e-5
        for k, la in enumerate(r['lookahead']):
            assert reward[e, k] == float(la['reward'])
            ref = np.array([[float(v) for v in row] for row in la['rotated']], dtype=np.float32)
            assert np.abs(states[e, k] - ref).max() < 2e-5, (e, k)
    # bigger random batch, N = 20 (tile path with many rows)
    host = _random_host_state(oracle, 257, 20, seed=9)
    env = cuda_env(257, 20, robot_visible=True, robot_policy='external_xy')
    env.state.load_host(host)
    states, reward = env.lookahead_pack(torch.from_numpy(actions).to(env.device))
    o_states, o_reward = oracle.lookahead_pack(oracle.default_params(robot_visible=1, robot_policy=0), host, actions)
    assert np.array_equal(reward.cpu().numpy(), o_reward)
    assert np.abs(states.cpu().numpy() - o_states).max() < 1e-4


def test_large_batch_linearity(cuda_env, oracle):
    """BASELINE.json sizes (4096 and 131072/8 = 16384 envs): a batch built from 500 distinct scenes tiled must give,
    for every copy, exactly the result of the 500-env batch (envs are independent -> size-independent property)."""
    N = 5
    cases = load_golden('suite_circle5_invisible')['cases']
    base = fill_host_state(oracle, [c['init'] for c in cases], N)
    small = cuda_env(500, N); small.state.load_host(base)
    for B in (4096, 16384):
        env = cuda_env(B, N)
        idx = torch.arange(B, device=env.device) % 500
        for f in env.state.FIELDS:
            getattr(env.state, f).copy_(getattr(small.state, f)[idx])
        s2 = cuda_env(500, N); s2.state.load_host(base)
        for _ in range(20):
            env.step(); s2.step()
        torch.cuda.synchronize()
        for f in ('h_pos', 'h_vel', 'r_pos', 'g_time'):
            assert torch.equal(getattr(env.state, f), getattr(s2.state, f)[idx]), (B, f)
        assert torch.equal(env.info, s2.info[idx]) and torch.equal(env.reward, s2.reward[idx])


def test_autoreset_install_bit_exact(cuda_env, oracle):
    """Consumer side of the auto-reset protocol: with identical prefetched scenes (generated by the oracle, copied into
    the device slots after every step) GPU and oracle stay bit-identical through hundreds of episode boundaries,
    including parked envs (slot not ready) and queue exhaustion."""
    for N, generic in [(5, 0), (5, 1), (12, 0)]:
        from crowdnav_b200 import _abi
        _abi.load().crowdsim_debug_force_generic(generic)
        B, k = 96, 700
        prm = oracle.default_params()
        host = oracle.HostState(B, N); io = oracle.HostStepIO(B); hep = oracle.HostEpisodes(B, k); har = oracle.HostAutoReset(B, N)
        counter = np.zeros(1, dtype=np.int32)
        q = dict(case_counter=counter, case_total=k, seed_base=2000)
        oracle.reset(host, None, ep=hep, **q)
        env = cuda_env(B, N)
        ep = env.track_episodes(k)
        env.enable_autoreset()
        env.state.load_host(host)
        ep.ep_case.copy_(torch.from_numpy(hep.ep_case))
        it = 0
        while (host.active.any() or har.want.any()) and it < 3000:
            if it % 3 == 0:                                  # prefetch only every third step: some envs find an EMPTY slot and park
                oracle.prefetch(har, B, N, **q)
            env.autoreset.load_host(har)
            env.step()
            oracle.step(prm, host, io, hep, har)
            torch.cuda.synchronize()
            d = env.autoreset.to_host()
            assert np.array_equal(d['n_state'], har.n_state) and np.array_equal(d['want'], har.want), it
            assert np.array_equal(env.state.active.cpu().numpy(), host.active), it
            if it % 25 == 0:
                _assert_state_equal(env, host, what='autoreset N=%d it=%d' % (N, it))
            it += 1
        assert int(counter[0]) >= k
        _assert_state_equal(env, host, what='autoreset final')
        for f in ('res_info', 'res_steps', 'res_time', 'res_return', 'res_too_close', 'res_min_dist_sum', 'res_final_rpos'):
            assert np.array_equal(getattr(ep, f).cpu().numpy(), getattr(hep, f)), f


@pytest.mark.parametrize('slots,suite', [(64, 'circle5_invisible'), (500, 'circle5_invisible'), (48, 'mixed5_invisible')])
def test_autoreset_device_prefetch_reproduces_suite(cuda_env, slots, suite):
    """Full device pipeline: scenes prefetched ON DEVICE from the shared case queue (side stream), installed by the step
    kernel. All test cases of the suite through `slots` slots: terminal class and step count exact, final position
    within 1e-5. (`mixed`: scenes with 1..5 humans, the other slots parked, stream through the same pipeline.)"""
    N, rule, _, _ = SUITES[suite]
    cases = load_golden('suite_' + suite)['cases']
    k = len(cases)
    env = cuda_env(slots, N, rule)
    ep = env.track_episodes(k)
    env.set_case_queue(0, k, 'test')
    env.enable_autoreset(rule)
    env.reset_seeds(rule=rule, use_queue=True)
    side = torch.cuda.Stream()
    for it in range(4000):
        if it % 2 == 0:
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                env.prefetch()                                # overlaps with the following steps
        env.step()
        if it % 64 == 63 and int(env.state.active.sum()) == 0 and int(env.autoreset.want.sum()) == 0:
            break
    torch.cuda.synchronize()
    assert int(env.state.active.sum()) == 0
    info = ep.res_info.cpu().numpy(); steps = ep.res_steps.cpu().numpy(); frp = ep.res_final_rpos.cpu().numpy()
    assert [int(x) for x in info] == [c['info'] for c in cases]
    assert [int(x) for x in steps] == [c['steps'] for c in cases]
    fr = np.array([scene_arrays(c['final'])[0][:2] for c in cases])
    assert np.abs(frp - fr).max() < 1e-5


@pytest.mark.parametrize('N,vis', [(5, 0), (5, 1), (4, 1), (3, 0), (2, 1), (1, 0)])
def test_step_n_equals_n_single_steps(cuda_env, oracle, N, vis):
    """crowdsim_step_n (one launch, state in registers across the steps) against n x oracle step on dense random scenes:
    every state / output array equal after launches of 2, 3, 8 and 16 steps (outputs = those of the last step)."""
    B = 1500
    host = _random_host_state(oracle, B, N, seed=300 + N)
    env = cuda_env(B, N, robot_visible=bool(vis), robot_policy='orca')
    env.state.load_host(host)
    prm = oracle.default_params(robot_visible=vis)
    io = oracle.HostStepIO(B)
    for n in (2, 3, 8, 16):
        env.step_n(n)
        for _ in range(n):
            oracle.step(prm, host, io)
        torch.cuda.synchronize()
        _assert_state_equal(env, host, what='step_n N=%d n=%d' % (N, n))
        _assert_io_equal(env, io, what='step_n N=%d n=%d' % (N, n))


@pytest.mark.parametrize('name', ['circle5_invisible', 'circle5_visible', 'square5_invisible', 'mixed5_invisible'])
def test_step_n_full_suites_from_reference_scenes(cuda_env, oracle, name):
    """Whole reference episodes through crowdsim_step_n with episode bookkeeping (envs freeze when their episode ends,
    inside the launch): result rows identical to the reference's Python for every test case."""
    N, rule, vis, rand = SUITES[name]
    cases = load_golden('suite_' + name)['cases']
    B = len(cases)
    host = fill_host_state(oracle, [c['init'] for c in cases], N)
    env = cuda_env(B, N, rule, robot_visible=bool(vis))
    ep = env.track_episodes(B)
    env.state.load_host(host)
    ep.ep_case.copy_(torch.arange(B, dtype=torch.int32))
    for n in (1, 2, 5, 16, 16, 16, 16, 16, 16, 16):
        env.step_n(n)
    torch.cuda.synchronize()
    assert int(env.state.active.sum()) == 0
    info = ep.res_info.cpu().numpy(); steps = ep.res_steps.cpu().numpy(); t = ep.res_time.cpu().numpy()
    ret = ep.res_return.cpu().numpy(); tc = ep.res_too_close.cpu().numpy(); mds = ep.res_min_dist_sum.cpu().numpy()
    frp = ep.res_final_rpos.cpu().numpy(); hp = env.state.h_pos.cpu().numpy()
    for i, c in enumerate(cases):
        assert info[i] == c['info'] and steps[i] == c['steps'], (name, c['case'])
        assert t[i] == (25.0 if c['info'] == 4 else float(c['global_time']))
        assert ret[i] == float(c['return']) and tc[i] == c['too_close'] and mds[i] == float(c['min_dist_sum'])
        r, h = scene_arrays(c['final'], N)
        assert (frp[i] == r[:2]).all() and (hp[i] == h[:, :2]).all()


@pytest.mark.parametrize('N,n', [(5, 4), (5, 7), (3, 5)])
def test_step_n_autoreset_bit_exact(cuda_env, oracle, N, n):
    """crowdsim_step_n with the auto-reset protocol: scenes prefetched by the oracle before every launch, installed inside
    the launch when an episode ends (a second termination in the same launch parks), result rows, slot flags and the whole
    state equal to n x oracle step through hundreds of episode boundaries and the exhaustion of the case queue."""
    B, k = 96, 600
    prm = oracle.default_params()
    host = oracle.HostState(B, N); io = oracle.HostStepIO(B); hep = oracle.HostEpisodes(B, k); har = oracle.HostAutoReset(B, N)
    counter = np.zeros(1, dtype=np.int32)
    q = dict(case_counter=counter, case_total=k, seed_base=2000)
    oracle.reset(host, None, ep=hep, **q)
    env = cuda_env(B, N)
    ep = env.track_episodes(k)
    env.enable_autoreset()
    env.state.load_host(host)
    ep.ep_case.copy_(torch.from_numpy(hep.ep_case))
    it = 0
    while (host.active.any() or har.want.any()) and it < 1500:
        if it % 2 == 0:                                      # no refill before every other launch: more envs park
            oracle.prefetch(har, B, N, **q)
        env.autoreset.load_host(har)
        env.step_n(n)
        for _ in range(n):
            oracle.step(prm, host, io, hep, har)
        torch.cuda.synchronize()
        d = env.autoreset.to_host()
        assert np.array_equal(d['n_state'], har.n_state) and np.array_equal(d['want'], har.want), it
        assert np.array_equal(env.state.active.cpu().numpy(), host.active), it
        if it % 10 == 0:
            _assert_state_equal(env, host, what='step_n autoreset N=%d it=%d' % (N, it))
            for f in ('ep_case', 'ep_steps', 'ep_return', 'ep_too_close', 'ep_min_dist_sum'):
                assert np.array_equal(getattr(ep, f).cpu().numpy(), getattr(hep, f)), (f, it)
        it += 1
    assert int(counter[0]) >= k
    _assert_state_equal(env, host, what='step_n autoreset final')
    _assert_io_equal(env, io, what='step_n autoreset final')
    for f in ('res_info', 'res_steps', 'res_time', 'res_return', 'res_too_close', 'res_min_dist_sum', 'res_final_rpos'):
        assert np.array_equal(getattr(ep, f).cpu().numpy(), getattr(hep, f)), f


def test_step_n_generic_and_external_fall_back_to_launch_loops(cuda_env, oracle):
    """n_steps > 1 outside the register-resident case (N > 5; external robot action, applied on every step) = n launches
    of the single-step kernels: same results as n oracle steps."""
    for N, policy in ((8, 'orca'), (5, 'external_xy')):
        from crowdnav_b200 import _abi
        B = 400
        host = _random_host_state(oracle, B, N, seed=77 + N)
        env = cuda_env(B, N, robot_policy=policy)
        env.state.load_host(host)
        prm = oracle.default_params(robot_policy=_abi.ROBOT_ORCA if policy == 'orca' else _abi.ROBOT_EXTERNAL_XY)
        io = oracle.HostStepIO(B)
        io.action[...] = np.random.RandomState(2).uniform(-1, 1, (B, 2))
        act = torch.from_numpy(io.action).to(env.device)
        env.step(None if policy == 'orca' else act, n_steps=5)
        for _ in range(5):
            oracle.step(prm, host, io)
        torch.cuda.synchronize()
        _assert_state_equal(env, host, what='step_n fallback N=%d' % N)
        _assert_io_equal(env, io, what='step_n fallback N=%d' % N)


def test_human_times_match_reference(cuda_env, oracle):
    """crowdsim_human_times against the reference's own CrowdSim.get_human_times (tests/golden/human_times: run after ORCA-robot
    episodes that ended at the goal; 5, 10 and 20 humans): arrival times, final global_time and the final positions of all
    agents -- the centralised float32 simulation is reproduced bit for bit; the live state is not touched."""
    rows = load_golden('human_times')['rows']
    assert len(rows) >= 10
    for r in rows:
        N = r['N']
        host = fill_host_state(oracle, [r['scene']], N)
        host.g_time[:] = float(r['global_time'])
        env = cuda_env(1, N, robot_visible=r['robot_visible'])
        env.state.load_host(host)
        before = torch.tensor([[float(t) for t in r['human_times_before']]], dtype=torch.float64)
        ht, gt, fp = env.human_times(before)
        torch.cuda.synchronize()
        assert ht[0].tolist() == [float(t) for t in r['human_times']], (r['tag'], r['case'])
        assert float(gt[0]) == float(r['global_time_after'])
        want = np.array([[float(x) for x in r['final_robot']]] + [[float(x) for x in h] for h in r['final_humans']])
        assert np.array_equal(fp[0].cpu().numpy(), want), (r['tag'], r['case'])
        _assert_state_equal(env, host, what='human_times leaves the state alone')


@pytest.mark.parametrize('N,vis,policy', [(5, 0, 'external_xy'), (5, 1, 'external_xy'), (3, 1, 'external_rot'), (12, 1, 'external_xy')])
def test_onestep_lookahead_is_a_step_without_update(cuda_env, oracle, N, vis, policy):
    """crowdsim_onestep_lookahead = step(action, update=False) (crowd_sim.py:314-315, 414-416): reward / dmin / done / info of
    the step that would happen, the humans' next observable states, and NOTHING mutated -- checked against one oracle step on
    a copy of the state."""
    from crowdnav_b200 import _abi
    B = 600
    host = _random_host_state(oracle, B, N, seed=60 + N)
    env = cuda_env(B, N, robot_visible=bool(vis), robot_policy=policy)
    env.state.load_host(host)
    pol = {'external_xy': _abi.ROBOT_EXTERNAL_XY, 'external_rot': _abi.ROBOT_EXTERNAL_ROT}[policy]
    prm = oracle.default_params(robot_visible=vis, robot_policy=pol)
    io = oracle.HostStepIO(B)
    rng = np.random.RandomState(8)
    if policy == 'external_rot':
        io.action[:, 0] = rng.uniform(0, 1, B); io.action[:, 1] = rng.uniform(-0.8, 0.8, B)
    else:
        io.action[...] = rng.uniform(-1, 1, (B, 2))
    (npos, nvel, _), rew, done, info = env.onestep_lookahead(torch.from_numpy(io.action).to(env.device))
    torch.cuda.synchronize()
    _assert_state_equal(env, host, what='lookahead leaves the state alone')
    import copy
    stepped = copy.deepcopy(host)
    oracle.step(prm, stepped, io)
    assert np.array_equal(npos.cpu().numpy(), stepped.h_pos) and np.array_equal(nvel.cpu().numpy(), stepped.h_vel)
    assert np.array_equal(info.cpu().numpy(), io.info) and np.array_equal(done.cpu().numpy(), io.done)
    if policy == 'external_rot':
        assert np.allclose(rew.cpu().numpy(), io.reward, rtol=0, atol=1e-12)
    else:
        assert np.array_equal(rew.cpu().numpy(), io.reward) and np.array_equal(env.dmin.cpu().numpy(), io.dmin)


def test_lookahead_humans_matches_oracle(cuda_env, oracle):
    """crowdsim_lookahead_humans = the `ob` of env.onestep_lookahead (crowd_sim.py:414-416): bit-exact against one oracle
    step on a copy of the state, small and large crowds, robot visible or not; the live state is untouched."""
    for N, vis in ((5, 0), (5, 1), (2, 1), (20, 0)):
        B = 700
        host = _random_host_state(oracle, B, N, seed=40 + N)
        env = cuda_env(B, N, robot_visible=bool(vis), robot_policy='external_xy')
        env.state.load_host(host)
        npos, nvel = env.lookahead_humans()
        torch.cuda.synchronize()
        o_pos, o_vel = oracle.lookahead_humans(oracle.default_params(robot_visible=vis, robot_policy=0), host)
        assert np.array_equal(npos.cpu().numpy(), o_pos) and np.array_equal(nvel.cpu().numpy(), o_vel), (N, vis)
        _assert_state_equal(env, host, what='state untouched')


def test_occupancy_maps_match_reference_and_oracle(cuda_env, oracle):
    """crowdsim_occupancy_maps vs (a) the reference's own build_occupancy_maps outputs (fixtures), (b) the oracle on random
    batches. Occupancy pattern identical, mean velocities to 1e-6 (float64 trig of CUDA vs glibc differs in the last ulp)."""
    rows = load_golden('occupancy_maps')['rows']
    for r in rows:
        h = np.array([[float(v) for v in hh] for hh in r['humans']])
        ref = np.array([[float(v) for v in m] for m in r['maps']], dtype=np.float32)
        env = cuda_env(1, h.shape[0])
        pos = torch.from_numpy(h[None, :, 0:2].copy()).to(env.device); vel = torch.from_numpy(h[None, :, 2:4].copy()).to(env.device)
        got = env.occupancy_maps(pos, vel, r['cell_num'], float(r['cell_size']), r['channels'])[0].cpu().numpy()
        assert np.array_equal(got != 0, ref != 0), r['tag']
        assert np.abs(got - ref).max() <= 1e-6, r['tag']
    rng = np.random.RandomState(3)
    for N in (2, 5, 20):
        B = 300
        pos = rng.uniform(-2.5, 2.5, (B, N, 2)); vel = rng.uniform(-1, 1, (B, N, 2))
        vel[::7, 0] = 0.0                                            # standing humans
        env = cuda_env(B, N)
        for ch in (1, 2, 3):
            got = env.occupancy_maps(torch.from_numpy(pos).to(env.device), torch.from_numpy(vel).to(env.device), 4, 1.0, ch).cpu().numpy()
            ref = oracle.occupancy_maps(pos, vel, 4, 1.0, ch)
            assert (got != 0).sum() > 0                               # the maps are not empty
            mism = (got != 0) != (ref != 0)
            assert mism.sum() == 0, (N, ch, int(mism.sum()))
            assert np.abs(got - ref).max() <= 1e-6
    env = cuda_env(4, 1)
    with pytest.raises(ValueError):
        env.occupancy_maps()                                          # the reference raises for a single human too
