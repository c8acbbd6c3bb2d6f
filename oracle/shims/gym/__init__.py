"""Minimal `gym` stand-in (TEST INFRASTRUCTURE). Just enough for the reference's
`/root/reference/crowd_sim/__init__.py:1-6` (register) and `crowd_nav/test.py:64` (gym.make)
to run unmodified in a container without gym. Not product code."""
import importlib


class Env(object):
    metadata = {}

    def step(self, action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    def render(self, mode='human'):
        raise NotImplementedError


_registry = {}


def register(id, entry_point=None, **kwargs):
    _registry[id] = (entry_point, kwargs)


def make(id, **kwargs):
    entry_point, reg_kwargs = _registry[id]
    if callable(entry_point):
        cls = entry_point
    else:
        mod_name, cls_name = entry_point.split(':')
        cls = getattr(importlib.import_module(mod_name), cls_name)
    kw = dict(reg_kwargs)
    kw.update(kwargs)
    return cls(**kw)
