#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched CrowdSim-v0 step path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--envs 4096] [--humans 5]

Workload (config.workload): BASELINE.json configs[1] = 4096 batched envs x 5 ORCA humans, circle_crossing, ORCA robot,
robot invisible, per GPU. One bench "step" = one lockstep pass of the hot path over one 4096-env batch: the fused step
kernel (6 ORCA solves/env, collision/reward/terminal, integration, episode bookkeeping) plus on-device re-generation
of the scenes of envs whose episode just ended (fresh MT19937 seeds), so every env is live on every step.
Because 4096 envs of state are only 2.6 MB, the bench rotates through POOLS independent batches whose combined state
exceeds the 126 MB L2 ("inputs larger than L2"); the batch touched by a step was last touched POOLS steps ago.
The timed region is ONE CUDA graph of K step launches (+ one scene-prefetch launch per batch on every 4th visit, side
streams); batch p always runs on stream p mod S (--streams, default 16), so the steps of one batch stay ordered while
independent batches overlap on the device. `value` = env-steps PERFORMED (counted by the step kernel) / device time;
`single_stream` = the same with one batch in flight; `e2e` = HostStepper.launch()/wait() over 16 batches with pinned host
buffers in and out on every batch-step; `roofline` = the step kernel alone (single-stream graph, CUDA events).

Printed JSON keys follow the driver contract; see DESIGN.md "Measurement" for definitions. The oracle (oracle/) is
executed here ONLY in the cpu_baseline leg and in --impl reference.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

try:                                             # the metric string is BASELINE.json's, verbatim
    with open(os.path.join(ROOT, 'BASELINE.json')) as _f:
        METRIC = json.load(_f)['metric']
except Exception:
    METRIC = 'env-steps/sec at 5 humans x batched envs; 500-case success/collision parity'
ALG_BYTES = lambda n: 8 * (19 + 12 * n) + 2      # 634 B at N=5, 2074 B at N=20  # noqa: E731


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=25600)
    ap.add_argument('--warmup', type=int, default=2560)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--envs', type=int, default=4096, help='envs per batch per GPU')
    ap.add_argument('--humans', type=int, default=5)
    ap.add_argument('--pools', type=int, default=0, help='independent batches rotated through (0 = enough to exceed L2)')
    ap.add_argument('--rule', default='circle_crossing')
    ap.add_argument('--streams', type=int, default=16, help='independent env batches stepped concurrently (CUDA streams inside the timed graph)')
    ap.add_argument('--prefetch-every', type=int, default=4, help='scene-prefetch launch for a batch on every n-th visit of that batch')
    ap.add_argument('--e2e-batches', type=int, default=16, help='independent env batches kept in flight by the e2e leg')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-python-loop', action='store_true', help='reference arm: skip the reference-shaped Python loop timing')
    ap.add_argument('--no-scale', action='store_true', help='skip the supplementary 1 Mi-env launch measurement')
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler(object):
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe): one long-lived
    `nvidia-smi -lms 20` child whose lines are timestamped on arrival; only samples that fall between mark_start() and
    mark_stop() are used (the child is started earlier so that its start-up cost is outside the window)."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index=0):
        self.rows, self.t0, self.t1 = [], None, None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '20'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        if self.proc is None:
            return
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(',')]))

    def wait_ready(self, timeout=20.0):
        """Block until the child has delivered its first sample (nvidia-smi can take seconds to start on a cold box)."""
        t_end = time.perf_counter() + timeout
        while not self.rows and time.perf_counter() < t_end and self.proc is not None and self.proc.poll() is None:
            time.sleep(0.01)
        return bool(self.rows)

    def mark_start(self):
        self.t0 = time.perf_counter()

    def mark_stop(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()            # the child we started, by handle
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        self.thread.join(timeout=5)
        inside = [r for t, r in self.rows if self.t0 is not None and self.t0 <= t <= (self.t1 or 1e30) and len(r) >= 7]
        used = inside if inside else [r for _, r in self.rows[-3:] if len(r) >= 7]
        num = lambda v: int(float(v)) if v.replace('.', '', 1).isdigit() else None  # noqa: E731
        sm = sorted(x for x in (num(r[0]) for r in used) if x is not None)
        mx = [x for x in (num(r[1]) for r in used) if x is not None]
        pw = [float(r[2]) for r in used if r[2].replace('.', '', 1).isdigit()]
        reasons = set()
        for r in used:
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': sorted(reasons),
                'samples': len(inside), 'power_w_max': max(pw) if pw else None,
                'note': None if inside else 'timed region shorter than the sampling period: nearest samples used'}


# ----------------------------------------------------------------------------------------------------------------------
def pick_threads(po, args):
    """'All the host threads it can use': each thread owns a fixed range of the batch's envs inside one parallel region
    (oracle.run_passes); SMT oversubscription or a busy host can still make fewer threads faster, so calibrate: time short
    runs at cpu_count, /2, /4, ... and keep the best."""
    import numpy as np
    B, N = args.envs, args.humans
    prm = po.default_params()
    st = po.HostState(B, N); io = po.HostStepIO(B)
    seeds = (np.arange(B) + 2000).astype(np.uint32)
    po.reset(st, seeds, args.rule, seed_stride=B)
    best = (0.0, 1)
    n = os.cpu_count() or 1
    cands = sorted({max(1, n >> s) for s in range(0, 6)}, reverse=True)
    for th in cands:
        po.set_threads(th)
        po.run_passes(prm, st, io, seeds, 10, args.rule, seed_stride=B)
        rate = 0.0
        for _ in range(3):                               # best of 3 short trials (>= 60 ms each) per thread count
            t0 = time.perf_counter(); done = 0
            while time.perf_counter() - t0 < 0.06:
                po.run_passes(prm, st, io, seeds, 20, args.rule, seed_stride=B); done += 20
            rate = max(rate, done * B / (time.perf_counter() - t0))
            if rate < 0.25 * best[0]:                    # hopeless thread count (oversubscribed / spinning): do not retry it
                break
        if rate > best[0]:
            best = (rate, th)
    po.set_threads(best[1])
    return best[1]


def cpu_oracle_rate(args, seconds=12.0):
    """Oracle port (plain C restatement of the reference loop) on the host cores, same workload definition:
    lockstep passes over a 4096-env batch with auto-reset. Returns (env-steps/s, threads, sample description)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import pyoracle as po
    nthreads = pick_threads(po, args)
    B, N = args.envs, args.humans
    prm = po.default_params()
    st = po.HostState(B, N); io = po.HostStepIO(B)
    seeds = (np.arange(B) + 2000).astype(np.uint32)
    po.reset(st, seeds, args.rule, seed_stride=B)
    po.run_passes(prm, st, io, seeds, 50, args.rule, seed_stride=B)
    rates, n_total, t_begin = [], 0, time.perf_counter()
    for _ in range(5):                                   # median of 5 segments: the host is shared, single segments are noisy
        t0 = time.perf_counter(); n = 0
        while time.perf_counter() - t0 < seconds / 5:
            po.run_passes(prm, st, io, seeds, 100, args.rule, seed_stride=B); n += 100
        rates.append(B * n / (time.perf_counter() - t0)); n_total += n
    rates.sort()
    return rates[2], nthreads, '%d lockstep passes over a %d-env batch (auto-reset) inside one OpenMP region, %.1f s, median of 5 segments (min %.2e, max %.2e)' % (
        n_total, B, time.perf_counter() - t_begin, rates[0], rates[-1])


WORKLOAD = '%d batched envs x %d ORCA humans, %s, ORCA robot (invisible), auto-reset, per GPU'


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path = oracle port (the reference is Python + an
    absent native rvo2; it cannot travel to the GPU box), all host threads, same config/metric."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import pyoracle as po
    B, N = args.envs, args.humans
    cores = pick_threads(po, args)
    prm = po.default_params()
    st = po.HostState(B, N); io = po.HostStepIO(B)
    seeds = (np.arange(B) + 2000).astype(np.uint32)
    po.reset(st, seeds, args.rule, seed_stride=B)

    po.run_passes(prm, st, io, seeds, args.warmup, args.rule, seed_stride=B)
    t0 = time.perf_counter()
    po.run_passes(prm, st, io, seeds, args.steps, args.rule, seed_stride=B)     # all passes inside one OpenMP parallel region
    dt = time.perf_counter() - t0
    v = B * args.steps / dt
    # the same port driven one call per pass from the interpreter (step; reset of the finished envs), like a host loop would
    t1 = time.perf_counter(); n_calls = 0
    while time.perf_counter() - t1 < 1.0:
        po.step(prm, st, io)
        po.reset(st, seeds, args.rule, mask=io.done, seed_stride=B)
        n_calls += 1
    per_call = B * n_calls / (time.perf_counter() - t1)
    # context: the same path with the reference's STRUCTURE (interpreter-bound Python loop around a native rvo2 step,
    # oracle/pyloop.py) on a bounded sample -- the reference's real files cannot travel to this box
    py = None
    if not args.no_python_loop:
        import pyloop
        one, n1 = pyloop.timed_rate(list(range(1000, 1064)), N, args.rule)
        allc, n2, procs = pyloop.timed_rate_all_cores(list(range(1000, 1000 + 16 * min(os.cpu_count() or 1, 64))), N, args.rule,
                                                      procs=min(os.cpu_count() or 1, 64))
        py = {'one_core_env_steps_per_s': one, 'all_cores_env_steps_per_s': allc, 'processes': procs,
              'sample': '%d + %d env-steps of the seeded test cases; Python loop + C rvo2 shim (oracle/pyloop.py); the reference\'s '
                        'own Python measured 4.8 k env-steps/s/core in the build container (DESIGN.md 6)' % (n1, n2)}
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'env-steps/s', 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64 state + f32 ORCA solver', 'data': 'synthetic',
            'config': {'workload': WORKLOAD % (B, N, args.rule), 'envs_per_gpu': B, 'humans': N,
                       'note': 'CPU arm: one 4096-env batch stepped in lockstep by all host threads (rank 0 only)'},
            'cpu_baseline': {'value': v, 'unit': 'env-steps/s', 'cores': cores, 'kind': 'port',
                             'sample': '%d lockstep passes over a %d-env batch in one C call; C restatement of the reference loop '
                                       '(oracle/crowdsim_oracle.c, OpenMP: every thread owns a range of envs, thread count calibrated, host has %d logical CPUs)' % (args.steps, B, os.cpu_count() or 0),
                             'per_pass_calls_value': per_call},
            'e2e': {'value': v, 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'python_loop': py, 'gpu_launches': 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from crowdnav_b200 import _abi
    from crowdnav_b200.batched import BatchedCrowdSim, default_config

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    sampler = ClockSampler(local)                            # started early: its start-up must be over before the timed region
    lib = _abi.load()
    B, N, K, W = args.envs, args.humans, args.steps, args.warmup
    bytes_per_env = ALG_BYTES(N)
    pools = args.pools or max(2, int(1.3 * 126e6 / (B * bytes_per_env)) + 1)

    envs = []
    for p in range(pools):
        env = BatchedCrowdSim(B, device=dev)
        env.configure(default_config(human_num=N, test_sim=args.rule, train_val_sim=args.rule))
        env.set_robot_policy('orca')
        # every batch streams its own range of train-phase cases (seed = 2000 + case, crowd_sim.py:272-273) through its
        # B slots; per-episode result rows are recorded on device and reduced once at the end (the path's one collective)
        visits = (max(W, 3) + K + 3) // pools + 2
        env.k_total = B * (visits // 5 + 3)
        env.track_episodes(env.k_total, gamma=0.9)
        env.set_case_queue((rank * pools + p) * env.k_total, env.k_total, 'train')
        env.enable_autoreset(args.rule)
        env.reset_seeds(rule=args.rule, use_queue=True)
        env.prefetch()
        envs.append(env)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # One bench step = the fused step kernel on the main stream (it also installs the prefetched next scene of every env
    # whose episode just ended) + one scene-prefetch kernel for that batch on a side stream (off the critical path: it
    # only has to finish before the same batch is stepped again). The whole timed region is ONE CUDA graph of K steps:
    # Python/ctypes launch overhead (~20 us per call) would otherwise dominate a 4096-env step.
    main = torch.cuda.Stream(device=dev)
    LANES = max(1, min(args.streams, pools))
    lanes = [torch.cuda.Stream(device=dev) for _ in range(LANES)]
    sides = [torch.cuda.Stream(device=dev) for _ in range(max(4, LANES))]
    it = 0
    n_prefetch = 0

    def capture(n_steps, with_prefetch=True, step_fn=None, n_lanes=None):
        """One CUDA graph of n_steps bench steps. Batch p is always stepped on lane (stream) p % n_lanes: the steps of one
        batch stay ordered, steps of different (independent) batches may overlap on the device when n_lanes > 1."""
        nonlocal it, n_prefetch
        n_prefetch = 0
        n_lanes = n_lanes or LANES
        g = torch.cuda.CUDAGraph()
        pending = {}                                   # pool -> event of its last prefetch inside this capture
        with torch.cuda.graph(g, stream=main):
            fork = torch.cuda.Event(); fork.record(main)
            for ls in lanes[:n_lanes]:
                ls.wait_event(fork)
            for _ in range(n_steps):
                p = it % pools; env = envs[p]; it += 1
                ls = lanes[p % n_lanes]
                with torch.cuda.stream(ls):
                    if p in pending:
                        ls.wait_event(pending.pop(p))
                    (step_fn or (lambda e: e.step()))(env)
                    if with_prefetch and ((it - 1) // pools) % args.prefetch_every == 0:
                        n_prefetch += 1
                        ev = torch.cuda.Event(); ev.record(ls)
                        sd = sides[p % len(sides)]
                        sd.wait_event(ev)
                        with torch.cuda.stream(sd):
                            env.prefetch()
                            done = torch.cuda.Event(); done.record(sd)
                        pending[p] = done
            for ev in pending.values():                # join the side branches and the lanes
                main.wait_event(ev)
            for ls in lanes[:n_lanes]:
                ev = torch.cuda.Event(); ev.record(ls); main.wait_event(ev)
        return g
    with torch.cuda.stream(main):
        for _ in range(3):
            envs[it % pools].step(); envs[it % pools].prefetch(); it += 1
    torch.cuda.synchronize()
    g_warm = capture(max(W, 3))
    g_timed = capture(K)
    timed_prefetches = n_prefetch

    def env_steps_done():
        """env-steps actually performed so far on this rank (finished episodes + episodes in progress): an env whose next
        scene is not ready when its episode ends is parked until the refill arrives and performs no env-step meanwhile."""
        tot = 0
        for env in envs:
            n = int(min(env._case_counter.item(), env.k_total))
            tot += int(env.episodes.res_steps[:n].sum().item()) + int((env.episodes.ep_steps * env.state.active.to(torch.int32)).sum().item())
        return tot
    with torch.cuda.stream(main):
        g_warm.replay()
    barrier()
    steps_before = env_steps_done()
    sampler.wait_ready()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.mark_start()
    with torch.cuda.stream(main):
        e0.record()
        g_timed.replay()
        e1.record()
    barrier()
    sampler.mark_stop()
    launches = K + timed_prefetches
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    live_steps = env_steps_done() - steps_before             # counted on device by the step kernel itself
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    cnt = torch.tensor([live_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms_max = float(t.item())
    live_total = int(cnt.item())
    value = live_total / (ms_max * 1e-3)                     # == world * B * K / time unless envs were parked

    # ---- the path's single collective: gather of episode statistics (terminal-class counts + env-steps of all finished
    # episodes of every rank) to rank 0 ----
    def episode_summary():
        tot = torch.zeros(5, dtype=torch.float64, device=dev)
        for env in envs:
            n = int(min(env._case_counter.item(), env.k_total))
            info = env.episodes.res_info[:n]; steps = env.episodes.res_steps[:n]
            fin = steps > 0
            for j, code in enumerate((_abi.INFO_REACHGOAL, _abi.INFO_COLLISION, _abi.INFO_TIMEOUT)):
                tot[j] += ((info == code) & fin).sum()
            tot[3] += steps[fin].sum(); tot[4] += fin.sum()
        return tot
    summ = episode_summary()
    if world > 1:
        gathered = [torch.empty_like(summ) for _ in range(world)]
        dist.all_gather(gathered, summ)
        summ = torch.stack(gathered).sum(dim=0)
    summ = summ.tolist()
    episodes = {'finished': int(summ[4]), 'success_rate': summ[0] / max(summ[4], 1), 'collision_rate': summ[1] / max(summ[4], 1),
                'timeout_rate': summ[2] / max(summ[4], 1), 'mean_steps': summ[3] / max(summ[4], 1),
                'note': 'all episodes finished so far on all ranks (gathered with one NCCL all_gather when n_gpus > 1); reference '
                        '500-case test suite: 0.43 / 0.57 / 0.006'}

    # ---- the same bench step with ONE batch in flight (every step launch waits for the previous one): reported next to
    # `value`, which keeps `--streams` independent batches in flight ----
    single = None
    if LANES > 1:
        Ks = max(pools, K // 4)
        g_single = capture(Ks, n_lanes=1)
        barrier()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(main):
            q0.record(); g_single.replay(); q1.record()
        barrier()
        tq = torch.tensor([q0.elapsed_time(q1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tq, op=dist.ReduceOp.MAX)
        single = {'value': world * B * Ks / (float(tq.item()) * 1e-3), 'unit': 'env-steps/s', 'steps': Ks, 'ms_per_step': float(tq.item()) / Ks,
                  'note': 'one 4096-env batch in flight at a time (single stream), nominal env-step count'}
        del g_single

    # ---- roofline of the dominant kernel (the step kernel): a graph of one step launch per pool, no resets in between,
    # replayed R times; CUDA events on the launching stream; per-launch duration = elapsed / (R * pools). The pools
    # rotate, so each launch reads its state from HBM, not L2. ----
    def step_only(env):
        ep, ar = env.episodes, env.autoreset         # no bookkeeping / auto-reset: finished envs keep stepping (same work)
        env.episodes = None; env.autoreset = None
        env.step()
        env.episodes, env.autoreset = ep, ar
    g_step = capture(pools, with_prefetch=False, step_fn=step_only, n_lanes=1)
    with torch.cuda.stream(main):
        g_step.replay()
    torch.cuda.synchronize()
    R = max(3, min(20, 1200 // pools))
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(main):
        k0.record()
        for _ in range(R):
            g_step.replay()
        k1.record()
    torch.cuda.synchronize()
    k_avg = k0.elapsed_time(k1) / (R * pools)
    peak, peak_src = load_peaks()
    achieved = B * bytes_per_env / (k_avg * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'step_traffic.json')
    if os.path.exists(tpath) and N == 5 and B == 4096:
        tj = json.load(open(tpath))
        traffic = tj['dram_bytes_read'] + tj['dram_bytes_write']          # from the committed ncu --set full capture, per launch
    roofline = {'bound': 'hbm', 'kernel': 'cs::step_flat_kernel' if N <= 5 else 'cs::step_kernel', 'achieved': achieved, 'peak': peak,
                'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                'algorithmic_bytes_per_launch': B * bytes_per_env, 'avg_launch_us': 1e3 * k_avg,
                'how': 'CUDA events around %d replays of a single-stream graph of %d back-to-back step launches (one per rotating batch)' % (R, pools),
                'timed_region_GBps': (live_total / world) * bytes_per_env / (ms_max * 1e-3) / 1e9,
                'timed_region_frac': (live_total / world) * bytes_per_env / (ms_max * 1e-3) / 1e9 / peak,
                'timed_region_note': 'algorithmic bytes of all steps of the timed region / its duration, with %d independent batches in flight' % LANES}

    # ---- supplementary: the same step kernel when the batch fills the chip (1 Mi envs in ONE launch, state = 665 MB) ----
    scale = None
    if rank == 0 and not args.no_scale:
        Bs = 1 << 20
        big = BatchedCrowdSim(Bs, device=dev)
        big.configure(default_config(human_num=N, test_sim=args.rule, train_val_sim=args.rule))
        big.set_robot_policy('orca')
        big.reset_seeds(torch.arange(Bs, dtype=torch.int64) % (2 ** 31) + 5000, rule=args.rule)
        with torch.cuda.stream(main):
            for _ in range(12):
                big.step()                                   # into the episodes (agents meet around step 12-20)
            gb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gb, stream=main):
            for _ in range(8):
                big.step()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(main):
            s0.record(); gb.replay(); s1.record()
        torch.cuda.synchronize()
        us = s0.elapsed_time(s1) / 8 * 1e3
        scale = {'envs_per_launch': Bs, 'us_per_launch': us, 'env_steps_per_s': Bs / us * 1e6, 'achieved_GBps': Bs * bytes_per_env / us / 1e3,
                 'roofline_frac': Bs * bytes_per_env / us / 1e3 / peak, 'note': 'step kernel only, no resets; shows the issue-bound regime when the chip is full'}
        del big, gb

    # ---- e2e: the public host-facing API (HostStepper.step): pinned HOST buffers in and out every step. The robot is
    # driven from the host like the reference's Explorer loop does it: action up, obs/reward/done/info (+ the robot's
    # next ORCA decision) down, host waits for the results before the next step. ----
    import numpy as np
    from crowdnav_b200.batched import HostStepper
    P = max(1, min(pools, args.e2e_batches))
    steppers = []
    for env in envs[:P]:
        env.reset_seeds(rule=args.rule, use_queue=True)      # fresh scenes (the step-only pass ran past terminal states)
        env.set_robot_policy('external_xy')
        steppers.append(HostStepper(env, next_orca_action=True))
    for st in steppers:
        st.step()
        for _ in range(5):
            st.h_action.copy_(st.h_next_action); st.step()
    ke = min(K, 400)

    def e2e_rate(group):
        """Round-robin over the batches in `group`; each visit = wait for the batch's previous step (results in host
        memory), host-side "policy" (apply the decision the device computed), enqueue its next step."""
        for st in group:
            st.launch()
        barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            for st in group:
                st.wait()
                np.copyto(st.np_action, st.np_next_action)
                st.launch()
        for st in group:
            st.wait()
        dt_ = time.perf_counter() - t0
        t = torch.tensor([dt_], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return world * B * ke * len(group) / float(t.item())

    e2e_single = e2e_rate(steppers[:1])                      # one batch, host blocks on every step (the reference's loop shape)
    e2e_value = e2e_rate(steppers) if P > 1 else e2e_single  # P independent batches in flight
    stepper = steppers[0]
    h2d, d2h = stepper.h2d_bytes, stepper.d2h_bytes
    launches_note = 'timed region: %d step + %d scene-prefetch kernel launches (one CUDA graph, prefetch on side streams, every %d-th visit of a batch)' % (K, timed_prefetches, args.prefetch_every)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, sample = cpu_oracle_rate(args)
        cpu = {'value': v, 'unit': 'env-steps/s', 'cores': cores, 'kind': 'port',
               'sample': sample + '; C restatement of the reference loop (oracle/crowdsim_oracle.c), OpenMP over envs'}

    # ---- second half of the metric ("500-case success/collision parity"): BASELINE config 1's 500 test cases (seeds 1000..1499,
    # ORCA robot) through BatchedExplorer on this GPU, against the reference's recorded outcome (tests/golden).
    # Outside every timed region; a failure here is reported, it does not take the throughput line down.
    parity = None
    if rank == 0:
        try:
            from crowdnav_b200.explorer import BatchedExplorer
            penv = BatchedCrowdSim(512, device=dev)
            penv.configure(default_config(human_num=5))
            st500 = BatchedExplorer(penv, 'orca', gamma=0.9).run_k_episodes(500, 'test')
            ref500 = {'success': 213, 'collision': 284, 'timeout': 3, 'timeout_cases': [118, 168, 224], 'env_steps': 15190}
            got500 = {k_: st500[k_] for k_ in ref500}
            parity = {'cases': 500, 'ours': got500, 'reference': ref500, 'match': got500 == ref500,
                      'note': 'test.py --policy orca flow (5 humans, circle_crossing, invisible robot); per-case bit-exact parity is in tests/'}
        except Exception as ex:                              # noqa: BLE001
            parity = {'error': repr(ex)}

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': 'env-steps/s', 'n_gpus': world, 'steps': K, 'warmup': max(W, 3),
                'ms_per_step': ms_max / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': 'f64 state + f32 ORCA solver', 'data': 'synthetic',
                'config': {'workload': WORKLOAD % (B, N, args.rule),
                           'envs_per_gpu': B, 'humans': N, 'l2': 'inputs larger than L2: %d rotating independent batches = %.0f MB of state' % (pools, pools * B * bytes_per_env / 1e6),
                           'batches_in_flight': LANES,
                           'parallelism': 'independent envs sharded over %d GPU(s), no data-path collective' % world},
                'env_steps': {'performed': live_total, 'nominal': world * B * K,
                              'note': 'value = performed / time; performed is counted by the step kernel (episode step counters), nominal = envs x steps; they differ only if envs waited for a scene refill'},
                'clocks': clocks, 'gpu_launches': int(launches), 'gpu_launches_note': launches_note,
                'e2e': {'value': e2e_value, 'unit': 'env-steps/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                        'steps': ke, 'batches_in_flight': P, 'single_batch_blocking': e2e_single,
                        'note': 'HostStepper.launch()/wait() round-robin over %d independent 4096-env batches: per batch-step a pinned host action buffer goes up and obs/reward/done/info/next ORCA action come down (byte counts are per batch-step), the host waits for a batch\'s results before it feeds that batch again; single_batch_blocking = one batch, host blocks on every step' % P},
                'single_stream': single, 'parity_500_cases': parity, 'episodes': episodes, 'roofline': roofline, 'scale': scale, 'cpu_baseline': cpu}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
