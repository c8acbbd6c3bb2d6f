"""Empty `matplotlib` stand-in (TEST INFRASTRUCTURE): the reference imports
matplotlib.lines / matplotlib.patches at module import (crowd_sim.py:3,6); rendering is out of scope."""
