"""Explorer for the single-environment surface (crowd_nav/utils/explorer.py:7-132): same constructor and
run_k_episodes / update_memory signatures, statistics and log lines. It drives env.reset / robot.act / env.step one
episode at a time, reduces every episode to the same result row the batched step kernel writes on device
(terminal class, steps, time, discounted return, danger count, sum of danger distances) and hands the rows to the one
reducer both explorers share (crowdnav_b200.explorer.summarize). For throughput use
crowdnav_b200.explorer.BatchedExplorer, which runs the same k episodes through thousands of device-resident slots."""
import copy

import torch

from .. import _abi
from ..explorer import average, summarize  # noqa: F401  (average: module-level name of the reference explorer.py:128-132)
from .statetypes import Collision, Danger, ReachGoal, Timeout

_TERMINAL_CODE = {ReachGoal: _abi.INFO_REACHGOAL, Collision: _abi.INFO_COLLISION, Timeout: _abi.INFO_TIMEOUT}


class Explorer(object):
    def __init__(self, env, robot, device, memory=None, gamma=None, target_policy=None):
        self.env, self.robot, self.device = env, robot, device
        self.memory, self.gamma, self.target_policy = memory, gamma, target_policy
        self.target_model = None

    def update_target_model(self, target_model):
        self.target_model = copy.deepcopy(target_model)

    def _rollout(self, phase):
        """One episode: (terminal info object, decision states, rewards, danger count, sum of danger distances)."""
        ob = self.env.reset(phase)
        states, rewards, n_danger, danger_sum = [], [], 0, 0.0
        while True:
            action = self.robot.act(ob)
            ob, reward, done, info = self.env.step(action)
            states.append(self.robot.policy.last_state)
            rewards.append(reward)
            if isinstance(info, Danger):
                n_danger += 1
                danger_sum += info.min_dist
            if done:
                return info, states, rewards, n_danger, danger_sum

    def run_k_episodes(self, k, phase, update_memory=False, imitation_learning=False, episode=None, print_failure=False):
        self.robot.policy.set_phase(phase)
        rows = []
        for _ in range(k):
            info, states, rewards, n_danger, danger_sum = self._rollout(phase)
            # read AFTER the rollout: env.reset() is what assigns robot.time_step (crowd_sim.py:296-298; the reference
            # reads it at explorer.py:71-72). exponent = (t * dt) * v_pref, in the reference's association
            dt, v_pref = self.robot.time_step, self.robot.v_pref
            code = _TERMINAL_CODE.get(type(info))
            if code is None:
                raise ValueError('Invalid end signal from environment')
            if update_memory and code != _abi.INFO_TIMEOUT:            # explorer.py:66-69: successes and collisions only
                self.update_memory(states, None, rewards, imitation_learning)
            ret = sum([pow(self.gamma, t * dt * v_pref) * r for t, r in enumerate(rewards)])
            t_end = self.env.time_limit if code == _abi.INFO_TIMEOUT else self.env.global_time
            rows.append([code, len(rewards), t_end, ret, n_danger, danger_sum])
        summarize(torch.tensor(rows, dtype=torch.float64), k, phase, self.env.time_limit, self.robot.time_step,
                  episode=episode, print_failure=print_failure)

    def update_memory(self, states, actions, rewards, imitation_learning=False):
        """explorer.py:92-125: one (state, value) pair per decision. Imitation learning stores the discounted
        return-to-go of the demonstration, RL the one-step bootstrap from the target network (terminal step: the reward)."""
        if self.memory is None or self.gamma is None:
            raise ValueError('Memory or gamma value is not set!')
        dt, v_pref = self.robot.time_step, self.robot.v_pref
        last = len(states) - 1
        for i, state in enumerate(states):
            if imitation_learning:
                state = self.target_policy.transform(state)
                value = sum([pow(self.gamma, max(t - i, 0) * dt * v_pref) * r * (1 if t >= i else 0)
                             for t, r in enumerate(rewards)])
            elif i == last:
                value = rewards[i]
            else:
                bootstrap = self.target_model(states[i + 1].unsqueeze(0)).data.item()
                value = rewards[i] + pow(self.gamma, dt * v_pref) * bootstrap
            self.memory.push((state, torch.Tensor([value]).to(self.device)))
