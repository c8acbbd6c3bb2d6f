"""Agent / Human / Robot (crowd_sim/envs/utils/agent.py:10-138, human.py:5-17, robot.py:5-14): the physical attributes
and the act() protocol. In the CUDA-backed CrowdSim the agents' kinematic state lives on the device; these objects are
host-side mirrors refreshed after every env.reset / env.step. The attribute and method names are the contract the
reference's policies and drivers rely on (px, py, gx, gy, vx, vy, theta, radius, v_pref, visible, policy, kinematics,
sensor, time_step; set / get_* / compute_position / step / reached_destination / act)."""
import logging
import math

import numpy as np

from .policies import policy_factory
from .statetypes import ActionXY, ActionRot, FullState, JointState, ObservableState

_KINEMATIC_FIELDS = ('px', 'py', 'gx', 'gy', 'vx', 'vy', 'theta')


def _pair(first, second):
    """Accessor pair for a two-vector stored as two scalar attributes: (getter returning a tuple, setter from a sequence)."""
    def getter(self):
        return getattr(self, first), getattr(self, second)

    def setter(self, value):
        setattr(self, first, value[0]); setattr(self, second, value[1])
    return getter, setter


class Agent(object):
    def __init__(self, config, section):
        read = {'visible': config.getboolean, 'v_pref': config.getfloat, 'radius': config.getfloat, 'sensor': config.get}
        for name, fn in read.items():
            setattr(self, name, fn(section, name))
        self.policy = policy_factory[config.get(section, 'policy')]()
        self.kinematics = getattr(self.policy, 'kinematics', None)
        for name in _KINEMATIC_FIELDS:
            setattr(self, name, None)
        self.time_step = None

    get_position, set_position = _pair('px', 'py')
    get_velocity, set_velocity = _pair('vx', 'vy')
    get_goal_position = _pair('gx', 'gy')[0]

    def print_info(self):
        logging.info('Agent is {} and has {} kinematic constraint'.format('visible' if self.visible else 'invisible',
                                                                          self.kinematics))

    def set_policy(self, policy):
        self.policy, self.kinematics = policy, policy.kinematics

    def sample_random_attributes(self):
        """agent.py:39-45: v_pref first, then radius (the order fixes the random stream)."""
        self.v_pref = np.random.uniform(0.5, 1.5)
        self.radius = np.random.uniform(0.3, 0.5)

    def set(self, px, py, gx, gy, vx, vy, theta, radius=None, v_pref=None):
        for name, value in zip(_KINEMATIC_FIELDS, (px, py, gx, gy, vx, vy, theta)):
            setattr(self, name, value)
        self.radius = self.radius if radius is None else radius
        self.v_pref = self.v_pref if v_pref is None else v_pref

    # ---- kinematics (agent.py:63-74, 110-135) ----
    def _holonomic(self):
        return self.kinematics == 'holonomic'

    def check_validity(self, action):
        assert isinstance(action, ActionXY if self._holonomic() else ActionRot)

    def _world_velocity(self, action, heading):
        """Velocity in the world frame: the action itself (holonomic) or speed along `heading` (unicycle)."""
        if self._holonomic():
            return action.vx, action.vy
        return action.v * np.cos(heading), action.v * np.sin(heading)      # numpy's cos/sin like the reference (not libm's)

    def compute_position(self, action, delta_t):
        self.check_validity(action)
        if self._holonomic():
            return self.px + action.vx * delta_t, self.py + action.vy * delta_t
        heading = self.theta + action.r
        return self.px + np.cos(heading) * action.v * delta_t, self.py + np.sin(heading) * action.v * delta_t

    def get_next_observable_state(self, action):
        nx, ny = self.compute_position(action, self.time_step)
        wx, wy = self._world_velocity(action, None if self._holonomic() else self.theta + action.r)
        return ObservableState(nx, ny, wx, wy, self.radius)

    def step(self, action):
        self.px, self.py = self.compute_position(action, self.time_step)
        if not self._holonomic():
            self.theta = (self.theta + action.r) % (2 * math.pi)
        self.vx, self.vy = self._world_velocity(action, self.theta)

    def reached_destination(self):
        return np.linalg.norm((self.px - self.gx, self.py - self.gy)) < self.radius

    # ---- state tuples ----
    def get_observable_state(self):
        return ObservableState(self.px, self.py, self.vx, self.vy, self.radius)

    def get_full_state(self):
        return FullState(self.px, self.py, self.vx, self.vy, self.radius, self.gx, self.gy, self.v_pref, self.theta)

    def act(self, ob):
        raise NotImplementedError


class Human(Agent):
    def act(self, ob):
        return self.policy.predict(JointState(self.get_full_state(), ob))


class Robot(Agent):
    def act(self, ob):
        if self.policy is None:
            raise AttributeError('Policy attribute has to be set!')
        return self.policy.predict(JointState(self.get_full_state(), ob))
