"""Wall-normal profiles and friction Reynolds number (calcAvgChan init.cpp:150-208, printRes :210-256).

CPU: the oracle's restatement against a vectorised numpy formulation of the same definitions.
GPU: libcudns (on-device reductions) against the oracle, in a process of its own (tools/check_diagnostics.py; its output of the
hardware run is kept under profiles/)."""
import numpy as np
import pytest

import cudanavierstokes_b200 as cd
import oracle_binding as ob
from common import CONFIGS, apply_cfg

COEFF_F = {1: [-0.5], 2: [1 / 12, -2 / 3], 3: [-1 / 60, 3 / 20, -3 / 4], 4: [1 / 280, -4 / 105, 1 / 5, -4 / 5]}   # globals.h:69-82


def _channel_oracle(name="chan_s3v2", **over):
    op = apply_cfg(ob.params_tgv(16, 3), dict(CONFIGS[name], **over))
    o = ob.Oracle(op); o.init_channel()
    return op, o


def _numpy_profiles(st):
    r, u, v, w, e = st                                    # [mz][my][mx]
    ax = (0, 1)
    rm = r.mean(axis=ax); um = (r * u).mean(axis=ax) / rm; vm = (r * v).mean(axis=ax) / rm; wm = (r * w).mean(axis=ax) / rm
    em = e.mean(axis=ax)
    return np.stack([rm, um, vm, wm, em, ((r - rm) ** 2).mean(axis=ax), ((u - um) ** 2).mean(axis=ax), ((v - vm) ** 2).mean(axis=ax),
                     ((w - wm) ** 2).mean(axis=ax), ((e - em) ** 2).mean(axis=ax)])


def _numpy_retau(st, s, dx, xp0, Re):
    r, u, v, w, e = st
    cF = COEFF_F[s]
    dudx = sum(cF[i] * (w[:, :, s - i - 1] - w[:, :, s - i]) for i in range(s)) / dx * xp0
    muw = 1.0 / Re
    return (np.sqrt(muw * np.abs(dudx) / r[:, :, 0]) * r[:, :, 0] / muw).mean()


@pytest.mark.parametrize("name", ["chan_s3v2", "chan_s2v2"])
def test_oracle_profiles_and_retau_match_numpy(name):
    op, o = _channel_oracle(name)
    o.run(2)
    st = o.state()
    got = o.profiles(); ref = _numpy_profiles(st)
    for a, b in zip(got, ref):
        assert np.abs(a - b).max() <= 1e-13 * max(np.abs(b).max(), 1e-30) + 1e-30
    rt = _numpy_retau(st, op.stencilSize, o.dx, o.xp[0], op.Re)
    assert abs(o.retau() - rt) <= 1e-12 * rt


@pytest.mark.gpu
def test_device_profiles_and_retau_match_oracle():
    """channel (two stencil pairs, after 3 steps) and a ragged periodic box: tools/check_diagnostics.py in its own process"""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "check_diagnostics.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
