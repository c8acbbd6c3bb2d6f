"""Agent / Human / Robot (crowd_sim/envs/utils/agent.py:10-138, human.py:5-17, robot.py:5-14): the physical attributes
and the act() protocol. In the CUDA-backed CrowdSim the agents' kinematic state lives on the device; these objects are
host-side mirrors refreshed after every env.reset / env.step (attribute names are the contract: px, py, gx, gy, vx, vy,
theta, radius, v_pref, visible, policy, kinematics, sensor, time_step)."""
import numpy as np

from .policies import policy_factory
from .statetypes import ActionXY, ActionRot, FullState, JointState, ObservableState


class Agent(object):
    def __init__(self, config, section):
        self.visible = config.getboolean(section, 'visible')
        self.v_pref = config.getfloat(section, 'v_pref')
        self.radius = config.getfloat(section, 'radius')
        self.policy = policy_factory[config.get(section, 'policy')]()
        self.sensor = config.get(section, 'sensor')
        self.kinematics = self.policy.kinematics if self.policy is not None else None
        self.px = self.py = self.gx = self.gy = self.vx = self.vy = self.theta = None
        self.time_step = None

    def print_info(self):
        import logging
        logging.info('Agent is {} and has {} kinematic constraint'.format(
            'visible' if self.visible else 'invisible', self.kinematics))

    def set_policy(self, policy):
        self.policy = policy
        self.kinematics = policy.kinematics

    def sample_random_attributes(self):
        self.v_pref = np.random.uniform(0.5, 1.5)
        self.radius = np.random.uniform(0.3, 0.5)

    def set(self, px, py, gx, gy, vx, vy, theta, radius=None, v_pref=None):
        self.px, self.py, self.gx, self.gy, self.vx, self.vy, self.theta = px, py, gx, gy, vx, vy, theta
        if radius is not None:
            self.radius = radius
        if v_pref is not None:
            self.v_pref = v_pref

    def get_observable_state(self):
        return ObservableState(self.px, self.py, self.vx, self.vy, self.radius)

    def get_next_observable_state(self, action):
        self.check_validity(action)
        px, py = self.compute_position(action, self.time_step)
        if self.kinematics == 'holonomic':
            vx, vy = action.vx, action.vy
        else:
            th = self.theta + action.r
            vx, vy = action.v * np.cos(th), action.v * np.sin(th)
        return ObservableState(px, py, vx, vy, self.radius)

    def get_full_state(self):
        return FullState(self.px, self.py, self.vx, self.vy, self.radius, self.gx, self.gy, self.v_pref, self.theta)

    def get_position(self):
        return self.px, self.py

    def set_position(self, position):
        self.px, self.py = position[0], position[1]

    def get_goal_position(self):
        return self.gx, self.gy

    def get_velocity(self):
        return self.vx, self.vy

    def set_velocity(self, velocity):
        self.vx, self.vy = velocity[0], velocity[1]

    def act(self, ob):
        raise NotImplementedError

    def check_validity(self, action):
        assert isinstance(action, ActionXY if self.kinematics == 'holonomic' else ActionRot)

    def compute_position(self, action, delta_t):
        self.check_validity(action)
        if self.kinematics == 'holonomic':
            return self.px + action.vx * delta_t, self.py + action.vy * delta_t
        th = self.theta + action.r
        return self.px + np.cos(th) * action.v * delta_t, self.py + np.sin(th) * action.v * delta_t

    def step(self, action):
        self.check_validity(action)
        self.px, self.py = self.compute_position(action, self.time_step)
        if self.kinematics == 'holonomic':
            self.vx, self.vy = action.vx, action.vy
        else:
            self.theta = (self.theta + action.r) % (2 * np.pi)
            self.vx, self.vy = action.v * np.cos(self.theta), action.v * np.sin(self.theta)

    def reached_destination(self):
        return np.linalg.norm(np.array(self.get_position()) - np.array(self.get_goal_position())) < self.radius


class Human(Agent):
    def act(self, ob):
        return self.policy.predict(JointState(self.get_full_state(), ob))


class Robot(Agent):
    def act(self, ob):
        if self.policy is None:
            raise AttributeError('Policy attribute has to be set!')
        return self.policy.predict(JointState(self.get_full_state(), ob))
