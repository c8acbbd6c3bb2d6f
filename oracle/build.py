#!/usr/bin/env python
"""Build the CPU oracle libraries (TEST INFRASTRUCTURE) into oracle/_build/.

  librvo2_oracle.so      oracle/rvo2_sim.c        -> `rvo2.PyRVOSimulator` shim (oracle/shims/rvo2)
  libcrowdsim_oracle.so  oracle/crowdsim_oracle.c -> batched CPU restatement with the C-ABI's array layout

Flags: -O2 -ffp-contract=off, no -ffast-math, no -march=native: every float32/float64 operation is rounded
individually (x86-64 SSE2), which is what a stock x86-64 build of RVO2 / CPython does.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_build')
CFLAGS = ['-O2', '-ffp-contract=off', '-fno-fast-math', '-fPIC', '-shared', '-std=gnu99', '-Wall', '-Wno-unused-function']


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=False):
    os.makedirs(OUT, exist_ok=True)
    hdrs = [os.path.join(HERE, 'rvo2_f32.h'), os.path.join(HERE, '..', 'include', 'crowdsim_b200.h'), __file__]
    jobs = [
        ('librvo2_oracle.so', ['rvo2_sim.c'], []),
        ('libcrowdsim_oracle.so', ['crowdsim_oracle.c'], ['-fopenmp']),
    ]
    for name, srcs, extra in jobs:
        target = os.path.join(OUT, name)
        srcs = [os.path.join(HERE, s) for s in srcs]
        if force or _stale(target, srcs + hdrs):
            cmd = ['gcc'] + CFLAGS + extra + srcs + ['-o', target, '-lm']
            if verbose:
                print(' '.join(cmd))
            subprocess.check_call(cmd)
    return OUT


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose=True)
