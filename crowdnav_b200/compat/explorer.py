"""Explorer for the single-environment surface (crowd_nav/utils/explorer.py:7-132): same constructor, run_k_episodes
signature, statistics and log lines; drives env.reset / robot.act / env.step one episode at a time. For throughput use
crowdnav_b200.explorer.BatchedExplorer, which runs the same k episodes through thousands of device-resident slots."""
import copy
import logging

import torch

from .statetypes import Collision, Danger, ReachGoal, Timeout
from ..explorer import average


class Explorer(object):
    def __init__(self, env, robot, device, memory=None, gamma=None, target_policy=None):
        self.env, self.robot, self.device = env, robot, device
        self.memory, self.gamma, self.target_policy = memory, gamma, target_policy
        self.target_model = None

    def update_target_model(self, target_model):
        self.target_model = copy.deepcopy(target_model)

    def run_k_episodes(self, k, phase, update_memory=False, imitation_learning=False, episode=None, print_failure=False):
        self.robot.policy.set_phase(phase)
        times = {ReachGoal: [], Collision: [], Timeout: []}
        cases = {Collision: [], Timeout: []}
        too_close, min_dist, returns = 0, [], []
        for i in range(k):
            ob = self.env.reset(phase)
            done, states, actions, rewards = False, [], [], []
            while not done:
                action = self.robot.act(ob)
                ob, reward, done, info = self.env.step(action)
                states.append(self.robot.policy.last_state); actions.append(action); rewards.append(reward)
                if isinstance(info, Danger):
                    too_close += 1
                    min_dist.append(info.min_dist)
            kind = type(info)
            if kind not in times:
                raise ValueError('Invalid end signal from environment')
            times[kind].append(self.env.time_limit if kind is Timeout else self.env.global_time)
            if kind in cases:
                cases[kind].append(i)
            if update_memory and kind in (ReachGoal, Collision):
                self.update_memory(states, actions, rewards, imitation_learning)
            returns.append(sum([pow(self.gamma, t * self.robot.time_step * self.robot.v_pref) * r for t, r in enumerate(rewards)]))
        n_ok, n_col, n_to = len(times[ReachGoal]), len(times[Collision]), len(times[Timeout])
        assert n_ok + n_col + n_to == k
        nav_time = sum(times[ReachGoal]) / n_ok if n_ok else self.env.time_limit
        extra = '' if episode is None else 'in episode {} '.format(episode)
        logging.info('{:<5} {}has success rate: {:.2f}, collision rate: {:.2f}, nav time: {:.2f}, total reward: {:.4f}'.
                     format(phase.upper(), extra, n_ok / k, n_col / k, nav_time, average(returns)))
        if phase in ['val', 'test']:
            num_step = sum(times[ReachGoal] + times[Collision] + times[Timeout]) / self.robot.time_step
            logging.info('Frequency of being in danger: %.2f and average min separate distance in danger: %.2f',
                         too_close / num_step, average(min_dist))
        if print_failure:
            logging.info('Collision cases: ' + ' '.join([str(x) for x in cases[Collision]]))
            logging.info('Timeout cases: ' + ' '.join([str(x) for x in cases[Timeout]]))

    def update_memory(self, states, actions, rewards, imitation_learning=False):
        """explorer.py:92-125: (state, value) pairs; IL uses the discounted return-to-go, RL the target network."""
        if self.memory is None or self.gamma is None:
            raise ValueError('Memory or gamma value is not set!')
        step_discount = pow(self.gamma, self.robot.time_step * self.robot.v_pref)
        for i, state in enumerate(states):
            if imitation_learning:
                state = self.target_policy.transform(state)
                value = sum([pow(self.gamma, max(t - i, 0) * self.robot.time_step * self.robot.v_pref) * r * (1 if t >= i else 0)
                             for t, r in enumerate(rewards)])
            elif i == len(states) - 1:
                value = rewards[i]
            else:
                value = rewards[i] + step_discount * self.target_model(states[i + 1].unsqueeze(0)).data.item()
            self.memory.push((state, torch.Tensor([value]).to(self.device)))
