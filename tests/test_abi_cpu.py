"""CPU-only checks of the drop-in boundary: libcrowdsim_b200.so builds for sm_100a, loads without a GPU, exports
every symbol include/crowdsim_b200.h declares, its structs have the layout the ctypes mirror assumes, argument
validation returns the documented error codes before any CUDA call, and the solver was compiled without FMA
contraction (numerics contract of orca_device.cuh)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'crowdsim_b200.h')


@pytest.fixture(scope='module')
def lib():
    from crowdnav_b200 import build, _abi
    build.build()
    return _abi.load()


def test_exports_every_declared_symbol(lib):
    from crowdnav_b200 import _abi
    src = open(HEADER).read()
    declared = set(re.findall(r'^\s*(?:int|void|unsigned long long)\s+(crowdsim_\w+)\s*\(', src, re.M))
    assert declared == set(_abi.EXPORTS), declared ^ set(_abi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.crowdsim_abi_version() == _abi.ABI_VERSION
    assert lib.crowdsim_launch_count() == 0


def test_struct_layout_matches_header(tmp_path):
    from crowdnav_b200 import _abi
    pairs = [('crowdsim_params', _abi.Params), ('crowdsim_state', _abi.State), ('crowdsim_step_io', _abi.StepIO),
             ('crowdsim_episodes', _abi.Episodes), ('crowdsim_reset_args', _abi.ResetArgs),
             ('crowdsim_autoreset', _abi.AutoReset)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "%s"' % HEADER, 'int main(void){']
    for cname, ct in pairs:
        lines.append('printf("%s %%zu", sizeof(%s));' % (cname, cname))
        for fname, _ in ct._fields_:
            lines.append('printf(" %%zu", offsetof(%s, %s));' % (cname, fname))
        lines.append('printf("\\n");')
    lines.append('return 0;}')
    c = tmp_path / 'layout.c'
    c.write_text('\n'.join(lines))
    exe = tmp_path / 'layout'
    subprocess.check_call(['gcc', str(c), '-o', str(exe)])
    out = subprocess.check_output([str(exe)]).decode().splitlines()
    for (cname, ct), line in zip(pairs, out):
        parts = line.split()
        assert parts[0] == cname
        assert int(parts[1]) == C.sizeof(ct), cname
        for (fname, _), off in zip(ct._fields_, parts[2:]):
            assert getattr(ct, fname).offset == int(off), (cname, fname)


def test_argument_validation_without_gpu(lib):
    from crowdnav_b200 import _abi
    prm = _abi.Params(0.25, 25.0, 1.0, -0.25, 0.2, 0.5, 10.0, 5.0, 10, 0.0, 0.0, 0, _abi.ROBOT_ORCA)
    st, io = _abi.State(), _abi.StepIO()
    assert lib.crowdsim_step(None, 1, 5, C.byref(st), C.byref(io), None, None, None) == -1
    assert lib.crowdsim_step(C.byref(prm), 1, 5, C.byref(st), C.byref(io), None, None, None) == -1        # NULL arrays
    assert lib.crowdsim_step(C.byref(prm), 1, _abi.MAX_HUMANS + 1, C.byref(st), C.byref(io), None, None, None) == -2
    prm.max_neighbors = _abi.MAX_NEIGHBORS + 1
    assert lib.crowdsim_step(C.byref(prm), 1, 5, C.byref(st), C.byref(io), None, None, None) == -2
    prm.max_neighbors = 10
    assert lib.crowdsim_reset(None, 1, 5, C.byref(st), None, None) == -1
    assert lib.crowdsim_prefetch_scenes(None, 1, 5, None, None) == -1
    assert lib.crowdsim_pack_joint(1, 5, C.byref(st), 0, None, None) == -1
    assert lib.crowdsim_lookahead_pack(C.byref(prm), 1, 5, C.byref(st), None, 81, 0, None, None, None) == -1
    assert lib.crowdsim_orca_act(C.byref(prm), 1, 5, C.byref(st), None, None) == -1
    assert lib.crowdsim_graph_launch(None, None, None) == -1 and lib.crowdsim_event_wait(None) == -1
    assert lib.crowdsim_launch_count() == 0          # nothing was launched by rejected calls


def test_missing_library_fails_loudly(monkeypatch):
    from crowdnav_b200 import _abi
    monkeypatch.setattr(_abi, '_lib', None)
    monkeypatch.setattr(_abi, 'LIB_PATH', '/nonexistent/libcrowdsim_b200.so')
    with pytest.raises(_abi.CudaLibraryMissing):
        _abi.load()


def test_no_cpu_fallback_in_batched_env(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from crowdnav_b200.batched import BatchedCrowdSim
    with pytest.raises(RuntimeError):
        BatchedCrowdSim(4)


def test_solver_compiled_without_fma_contraction(tmp_path):
    """--fmad=false: the PTX of the step kernel contains no float32 fma at all (IEEE div/sqrt are PTX ops)."""
    from crowdnav_b200 import build
    ptx = tmp_path / 'step.ptx'
    flags = [f for f in build.NVCC_FLAGS if f not in ('-shared', '-Xcompiler', '-fPIC', '-cudart', 'shared', '-lineinfo')]
    subprocess.check_call([build._nvcc()] + flags + ['-ptx', os.path.join(build.CSRC, 'step_kernel.cu'), '-o', str(ptx)])
    text = ptx.read_text()
    assert '--fmad=false' in build.NVCC_FLAGS
    assert len(re.findall(r'\bfma\.rn\.f32\b', text)) == 0
    assert len(re.findall(r'\bmad\.f32\b', text)) == 0
    assert 'div.rn.f32' in text and 'sqrt.rn.f32' in text


def test_product_does_not_import_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may import, link or execute anything under oracle/."""
    import ast
    pkg = os.path.join(ROOT, 'crowdnav_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            path = os.path.join(dirpath, f)
            if f.endswith('.py'):
                tree = ast.parse(open(path).read())
                for node in ast.walk(tree):
                    if isinstance(node, ast.Import):
                        assert not any('oracle' in a.name for a in node.names), path
                    elif isinstance(node, ast.ImportFrom):
                        assert 'oracle' not in (node.module or ''), path
                    elif isinstance(node, ast.Constant) and isinstance(node.value, str) and not isinstance(getattr(node, 'parent', None), ast.Expr):
                        assert 'libcrowdsim_oracle' not in node.value and 'librvo2_oracle' not in node.value, path
            elif f.endswith(('.cu', '.cuh', '.h')):
                for line in open(path):
                    if line.lstrip().startswith('#include'):
                        assert 'oracle' not in line, path
