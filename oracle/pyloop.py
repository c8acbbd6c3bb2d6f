"""oracle/pyloop.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The reference's per-episode loop restated in plain Python with the reference's own structure: one `rvo2.PyRVOSimulator`
per agent that is refilled and stepped every env-step (crowd_sim/envs/policy/orca.py:82-132), a float64 Python env
around it (crowd_sim/envs/crowd_sim.py:251-312 reset, :317-420 step) and the Explorer's episode loop
(crowd_nav/utils/explorer.py:35-72). `rvo2` is the oracle shim (oracle/shims/rvo2 -> oracle/rvo2_sim.c).

Purpose: (1) a third, structurally independent restatement that is checked against the golden fixtures
(tests/test_oracle_cpu.py), (2) a CPU timing that has the *shape* of the reference (interpreter-bound Python around a
native solver) for bench.py --impl reference, next to the much faster plain-C port (oracle/crowdsim_oracle.c); the
reference's real files cannot travel to the GPU box.
"""
import os
import sys

import numpy as np
from numpy.linalg import norm

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'shims'))
sys.path.insert(0, HERE)
import build as _build  # noqa: E402

_build.build()
import rvo2  # noqa: E402  (oracle/shims/rvo2)

PARAMS = (10, 10, 5, 5)           # orca.py:61-64 neighbor_dist, max_neighbors, time_horizon, time_horizon_obst
DT, TIME_LIMIT = 0.25, 25
RADIUS, V_PREF, DISCOMFORT, CIRCLE_R, SQUARE_W = 0.3, 1.0, 0.2, 4.0, 10.0


class _Agent(object):
    __slots__ = ('px', 'py', 'vx', 'vy', 'gx', 'gy', 'radius', 'v_pref', 'sim')

    def __init__(self, px, py, gx, gy):
        self.px, self.py, self.gx, self.gy, self.vx, self.vy = px, py, gx, gy, 0.0, 0.0
        self.radius, self.v_pref, self.sim = RADIUS, V_PREF, None

    def orca(self, others, safety=0):
        """orca.py:82-132 for this agent observing `others` (list of _Agent)."""
        if self.sim is not None and self.sim.getNumAgents() != len(others) + 1:
            self.sim = None
        if self.sim is None:
            self.sim = rvo2.PyRVOSimulator(DT, *PARAMS, RADIUS, 1)
            self.sim.addAgent((self.px, self.py), *PARAMS, self.radius + 0.01 + safety, self.v_pref, (self.vx, self.vy))
            for o in others:
                self.sim.addAgent((o.px, o.py), *PARAMS, o.radius + 0.01 + safety, 1, (o.vx, o.vy))
        else:
            self.sim.setAgentPosition(0, (self.px, self.py)); self.sim.setAgentVelocity(0, (self.vx, self.vy))
            for i, o in enumerate(others):
                self.sim.setAgentPosition(i + 1, (o.px, o.py)); self.sim.setAgentVelocity(i + 1, (o.vx, o.vy))
        velocity = np.array((self.gx - self.px, self.gy - self.py))
        speed = np.linalg.norm(velocity)
        pref = velocity / speed if speed > 1 else velocity
        self.sim.setAgentPrefVelocity(0, tuple(pref))
        for i in range(len(others)):
            self.sim.setAgentPrefVelocity(i + 1, (0, 0))
        self.sim.doStep()
        return self.sim.getAgentVelocity(0)


def _segment_dist0(x1, y1, x2, y2):
    """utils.py:4-26 with the query point at the origin."""
    px, py = x2 - x1, y2 - y1
    if px == 0 and py == 0:
        return norm((-x1, -y1))
    u = ((0 - x1) * px + (0 - y1) * py) / (px * px + py * py)
    u = 1 if u > 1 else (0 if u < 0 else u)
    return norm((x1 + u * px, y1 + u * py))


def make_scene(seed, n, rule='circle_crossing'):
    """crowd_sim.py:251-312 + :155-207 (no attribute randomisation)."""
    np.random.seed(seed)
    robot = _Agent(0.0, -CIRCLE_R, 0.0, CIRCLE_R)
    humans = []
    for _ in range(n):
        if rule == 'circle_crossing':
            while True:
                angle = np.random.random() * np.pi * 2
                px = CIRCLE_R * np.cos(angle) + (np.random.random() - 0.5) * V_PREF
                py_noise = (np.random.random() - 0.5) * V_PREF
                py = CIRCLE_R * np.sin(angle) + py_noise
                if not any(norm((px - a.px, py - a.py)) < 2 * RADIUS + DISCOMFORT or norm((px - a.gx, py - a.gy)) < 2 * RADIUS + DISCOMFORT
                           for a in [robot] + humans):
                    break
            humans.append(_Agent(px, py, -px, -py))
        else:
            sign = -1 if np.random.random() > 0.5 else 1
            while True:
                px = np.random.random() * SQUARE_W * 0.5 * sign
                py = (np.random.random() - 0.5) * SQUARE_W
                if not any(norm((px - a.px, py - a.py)) < 2 * RADIUS + DISCOMFORT for a in [robot] + humans):
                    break
            while True:
                gx = np.random.random() * SQUARE_W * 0.5 * -sign
                gy = (np.random.random() - 0.5) * SQUARE_W
                if not any(norm((gx - a.gx, gy - a.gy)) < 2 * RADIUS + DISCOMFORT for a in [robot] + humans):
                    break
            humans.append(_Agent(px, py, gx, gy))
    return robot, humans


def run_episode(seed, n=5, rule='circle_crossing', robot_visible=False):
    """One episode with an ORCA robot (test.py --policy orca). Returns (info code, steps, global_time, robot xy)."""
    robot, humans = make_scene(seed, n, rule)
    t, steps = 0.0, 0
    while True:
        ax, ay = robot.orca(humans)                                                  # explorer.py:42
        acts = [h.orca([o for o in humans if o is not h] + ([robot] if robot_visible else [])) for h in humans]
        dmin, collision = float('inf'), False
        for h in humans:                                                             # crowd_sim.py:331-351
            px, py = h.px - robot.px, h.py - robot.py
            vx, vy = h.vx - ax, h.vy - ay
            c = _segment_dist0(px, py, px + vx * DT, py + vy * DT) - h.radius - robot.radius
            if c < 0:
                collision = True
                break
            elif c < dmin:
                dmin = c
        reach = norm(np.array((robot.px + ax * DT, robot.py + ay * DT)) - np.array((robot.gx, robot.gy))) < robot.radius
        if t >= TIME_LIMIT - 1:
            info = 4
        elif collision:
            info = 3
        elif reach:
            info = 2
        else:
            info = 1 if dmin < DISCOMFORT else 0
        robot.px, robot.py, robot.vx, robot.vy = robot.px + ax * DT, robot.py + ay * DT, ax, ay
        for h, (hx, hy) in zip(humans, acts):
            h.px, h.py, h.vx, h.vy = h.px + hx * DT, h.py + hy * DT, hx, hy
        t += DT
        steps += 1
        if info >= 2:
            return info, steps, t, (robot.px, robot.py)


def timed_rate(seeds, n=5, rule='circle_crossing'):
    import time
    t0 = time.perf_counter()
    steps = sum(run_episode(s, n, rule)[1] for s in seeds)
    return steps / (time.perf_counter() - t0), steps


def _worker(args):
    seeds, n, rule = args
    return sum(run_episode(s, n, rule)[1] for s in seeds)


def timed_rate_all_cores(seeds, n=5, rule='circle_crossing', procs=None):
    """The reference is single-threaded; 'all cores' = independent processes over disjoint case ranges."""
    import multiprocessing as mp
    import time
    procs = procs or os.cpu_count() or 1
    chunks = [seeds[i::procs] for i in range(procs)]
    with mp.get_context('fork').Pool(procs) as pool:
        pool.map(_worker, [([1000], n, rule)] * procs)          # warm the workers
        t0 = time.perf_counter()
        steps = sum(pool.map(_worker, [(c, n, rule) for c in chunks]))
        dt = time.perf_counter() - t0
    return steps / dt, steps, procs


if __name__ == '__main__':
    r, s = timed_rate(list(range(1000, 1064)))
    print('python loop + C rvo2 shim: %.0f env-steps/s on one core (%d env-steps)' % (r, s))
