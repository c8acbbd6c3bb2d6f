"""`rvo2` stand-in (TEST INFRASTRUCTURE, not product code): the Python-RVO2 `PyRVOSimulator`
subset used by the reference (orca.py:95-129, crowd_sim.py:221-245), backed through ctypes by
oracle/_build/librvo2_oracle.so (oracle/rvo2_sim.c + oracle/rvo2_f32.h).
Python floats are cast to C float at this boundary exactly like the Cython wrapper
(SURVEY.md Appendix A.5); results come back as Python floats holding float32 values."""
import ctypes
import os

_here = os.path.dirname(os.path.abspath(__file__))
_so = os.path.join(_here, '..', '..', '_build', 'librvo2_oracle.so')
if not os.path.exists(_so):
    raise ImportError('oracle library missing: run `python oracle/build.py` (%s)' % _so)
_lib = ctypes.CDLL(_so)
_d, _l, _p = ctypes.c_double, ctypes.c_long, ctypes.c_void_p
_lib.rvo_sim_create.restype = _p
_lib.rvo_sim_create.argtypes = [_d, _d, _l, _d, _d, _d, _d, _d, _d]
_lib.rvo_sim_destroy.argtypes = [_p]
_lib.rvo_sim_add_agent.restype = _l
_lib.rvo_sim_add_agent.argtypes = [_p, _d, _d, _d, _l, _d, _d, _d, _d, _d, _d]
_lib.rvo_sim_add_agent_default.restype = _l
_lib.rvo_sim_add_agent_default.argtypes = [_p, _d, _d]
_lib.rvo_sim_num_agents.restype = _l
_lib.rvo_sim_num_agents.argtypes = [_p]
_lib.rvo_sim_global_time.restype = _d
_lib.rvo_sim_global_time.argtypes = [_p]
for _n in ('rvo_sim_set_position', 'rvo_sim_set_velocity', 'rvo_sim_set_pref_velocity'):
    getattr(_lib, _n).restype = ctypes.c_int
    getattr(_lib, _n).argtypes = [_p, _l, _d, _d]
for _n in ('rvo_sim_get_position', 'rvo_sim_get_velocity'):
    getattr(_lib, _n).restype = ctypes.c_int
    getattr(_lib, _n).argtypes = [_p, _l, ctypes.POINTER(_d), ctypes.POINTER(_d)]
_lib.rvo_sim_do_step.argtypes = [_p]
_lib.rvo_sim_get_stats.argtypes = [_p, ctypes.POINTER(_l)]


class PyRVOSimulator(object):
    def __init__(self, timeStep, neighborDist, maxNeighbors, timeHorizon, timeHorizonObst, radius, maxSpeed,
                 velocity=(0, 0)):
        self._h = _lib.rvo_sim_create(timeStep, neighborDist, int(maxNeighbors), timeHorizon, timeHorizonObst,
                                      radius, maxSpeed, velocity[0], velocity[1])

    def __del__(self):
        h, self._h = getattr(self, '_h', None), None
        if h:
            _lib.rvo_sim_destroy(h)

    def addAgent(self, pos, neighborDist=None, maxNeighbors=None, timeHorizon=None, timeHorizonObst=None,
                 radius=None, maxSpeed=None, velocity=None):
        args = (neighborDist, maxNeighbors, timeHorizon, timeHorizonObst, radius, maxSpeed, velocity)
        if all(a is None for a in args):
            return _lib.rvo_sim_add_agent_default(self._h, pos[0], pos[1])
        if any(a is None for a in args):
            raise ValueError('Either pass only the position, or pass all parameters.')
        return _lib.rvo_sim_add_agent(self._h, pos[0], pos[1], neighborDist, int(maxNeighbors), timeHorizon,
                                      timeHorizonObst, radius, maxSpeed, velocity[0], velocity[1])

    def getNumAgents(self):
        return _lib.rvo_sim_num_agents(self._h)

    def getGlobalTime(self):
        return _lib.rvo_sim_global_time(self._h)

    def _chk(self, rc, i):
        if rc != 0:
            raise IndexError('agent index %r out of range' % (i,))

    def setAgentPosition(self, i, pos):
        self._chk(_lib.rvo_sim_set_position(self._h, i, pos[0], pos[1]), i)

    def setAgentVelocity(self, i, vel):
        self._chk(_lib.rvo_sim_set_velocity(self._h, i, vel[0], vel[1]), i)

    def setAgentPrefVelocity(self, i, vel):
        self._chk(_lib.rvo_sim_set_pref_velocity(self._h, i, vel[0], vel[1]), i)

    def getAgentPosition(self, i):
        x, y = _d(), _d()
        self._chk(_lib.rvo_sim_get_position(self._h, i, ctypes.byref(x), ctypes.byref(y)), i)
        return (x.value, y.value)

    def getAgentVelocity(self, i):
        x, y = _d(), _d()
        self._chk(_lib.rvo_sim_get_velocity(self._h, i, ctypes.byref(x), ctypes.byref(y)), i)
        return (x.value, y.value)

    def doStep(self):
        _lib.rvo_sim_do_step(self._h)

    def stats(self):
        out = (_l * 4)()
        _lib.rvo_sim_get_stats(self._h, out)
        return tuple(out)
