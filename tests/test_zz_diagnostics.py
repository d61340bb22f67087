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


# ---- mean square vorticity (libcudns's par2 of an unforced periodic box: the dissipation history of a Taylor-Green run) ----------
COEFF_VF = {1: [0.5], 2: [2 / 3, -1 / 12], 3: [3 / 4, -3 / 20, 1 / 60], 4: [4 / 5, -1 / 5, 4 / 105, -1 / 280]}   # a_l of globals.h:84-95


def _numpy_enstrophy(st, v, L):
    """<|curl u|^2> with the periodic central differences of order 2v (np.roll), plain volume mean (uniform grid)"""
    _, u, vv, w, _ = st                                   # [mz][my][mx]
    def d(f, axis, n, length):
        h = length / n
        return sum(a * (np.roll(f, -(l + 1), axis) - np.roll(f, l + 1, axis)) for l, a in enumerate(COEFF_VF[v])) / h
    mz, my, mx = u.shape
    dx = lambda f: d(f, 2, mx, L[0]); dy = lambda f: d(f, 1, my, L[1]); dz = lambda f: d(f, 0, mz, L[2])
    ox, oy, oz = dy(w) - dz(vv), dz(u) - dx(w), dx(vv) - dy(u)
    return (ox * ox + oy * oy + oz * oz).mean()


@pytest.mark.parametrize("s,v", [(3, 3), (4, 4), (3, 2)])
def test_oracle_enstrophy_matches_numpy_and_the_taylor_green_value(s, v):
    op = ob.params_tgv(32, s, stencilVisc=v)
    o = ob.Oracle(op); o.init_chit()
    e0 = o.enstrophy()
    assert abs(e0 - _numpy_enstrophy(o.state(), v, (op.Lx, op.Ly, op.Lz))) <= 1e-13 * e0
    assert abs(e0 - 0.75) <= 2e-3 * 0.75 ** v             # analytic <w.w> of the Taylor-Green start is 3/4; truncation error of order 2v
    o.run(3)
    e3 = o.enstrophy()
    assert abs(e3 - _numpy_enstrophy(o.state(), v, (op.Lx, op.Ly, op.Lz))) <= 1e-13 * e3


@pytest.mark.gpu
@pytest.mark.parametrize("s,v,shape", [(4, 4, (64, 32, 40)), (3, 2, (40, 20, 24)), (2, 2, (32, 32, 32))])
def test_device_enstrophy_and_par2_history_match_oracle(s, v, shape):
    from common import make_pair, smooth_random_state
    op = ob.params_tgv(32, s, stencilVisc=v, mx=shape[0], my=shape[1], mz=shape[2], checkBulk=2)
    o, sol, grid = make_pair(op)
    st = smooth_random_state(o); o.set_state(st); sol.set_state(st)
    e_o, e_s = o.enstrophy(), sol.enstrophy()
    assert abs(e_s - e_o) <= 1e-12 * e_o
    # par2 stays unwritten by default (reference: calc_stress.cu:191-197) ...
    t, p1, p2 = sol.advance(2)
    assert np.isnan(p2).all()
    sol.close()
    # ... and carries <w.w> of the state at the start of every checkBulk-th step when the extension is switched on
    o2, sol2, grid = make_pair(op)
    sol2.close()
    cp = cd.Params.from_buffer_copy(sol2.p); cp.par2_enstrophy = 1
    sol2 = cd.Solver(cp, grid)
    o2.set_state(st); sol2.set_state(st)
    dt = 0.5 * o2.calc_dt(); o2.set_dt(dt); sol2.set_dt(dt)           # fixed: the oracle is stepped one step per call (cadence quirk Q11)
    t, p1, p2 = sol2.advance(4)
    ref = []
    for n in range(4):
        ref.append(o2.enstrophy() if n % 2 == 0 else np.nan)
        o2.run(1)
    for n in range(4):
        if n % 2 == 0:
            assert abs(p2[n] - ref[n]) <= 1e-11 * ref[n], (n, p2[n], ref[n])
        else:
            assert np.isnan(p2[n])
    sol2.close()
