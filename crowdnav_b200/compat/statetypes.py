"""Value types of the reference's agent/state model, restated (names and field order are the contract):
  ActionXY / ActionRot        crowd_sim/envs/utils/action.py:3-4
  FullState / ObservableState / JointState   crowd_sim/envs/utils/state.py:1-50  (the 9- and 5-tuples, `+` = tuple concat)
  Timeout / ReachGoal / Danger / Collision / Nothing   crowd_sim/envs/utils/info.py:1-38 (isinstance-tested by Explorer,
                                                         str()-printed by test.py:104)
"""
from collections import namedtuple

ActionXY = namedtuple('ActionXY', ['vx', 'vy'])
ActionRot = namedtuple('ActionRot', ['v', 'r'])


class _TupleState(object):
    """`tuple + state` appends the state's fields (state.py:17-18,36-37); str() prints them space-separated."""
    FIELDS = ()

    def _tuple(self):
        return tuple(getattr(self, f) for f in self.FIELDS)

    def __add__(self, other):
        return other + self._tuple()

    def __str__(self):
        return ' '.join(str(x) for x in self._tuple())


class ObservableState(_TupleState):
    FIELDS = ('px', 'py', 'vx', 'vy', 'radius')

    def __init__(self, px, py, vx, vy, radius):
        self.px, self.py, self.vx, self.vy, self.radius = px, py, vx, vy, radius
        self.position = (px, py)
        self.velocity = (vx, vy)


class FullState(_TupleState):          # deliberately NOT a subclass of ObservableState (cadrl.py:105-129 dispatches on the type)
    FIELDS = ('px', 'py', 'vx', 'vy', 'radius', 'gx', 'gy', 'v_pref', 'theta')

    def __init__(self, px, py, vx, vy, radius, gx, gy, v_pref, theta):
        self.px, self.py, self.vx, self.vy, self.radius = px, py, vx, vy, radius
        self.gx, self.gy, self.v_pref, self.theta = gx, gy, v_pref, theta
        self.position = (px, py)
        self.goal_position = (gx, gy)
        self.velocity = (vx, vy)


class JointState(object):
    def __init__(self, self_state, human_states):
        assert isinstance(self_state, FullState)
        for h in human_states:
            assert isinstance(h, ObservableState)
        self.self_state = self_state
        self.human_states = human_states


class Timeout(object):
    def __str__(self):
        return 'Timeout'


class ReachGoal(object):
    def __str__(self):
        return 'Reaching goal'


class Danger(object):
    def __init__(self, min_dist):
        self.min_dist = min_dist

    def __str__(self):
        return 'Too close'


class Collision(object):
    def __str__(self):
        return 'Collision'


class Nothing(object):
    def __str__(self):
        return ''


INFO_BY_CODE = {0: Nothing, 2: ReachGoal, 3: Collision, 4: Timeout}


def info_from_code(code, dmin):
    return Danger(dmin) if code == 1 else INFO_BY_CODE[code]()
