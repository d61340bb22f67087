"""CPU tests that PIN the oracle: against the outputs of the reference's own GPU binary (tests/golden/ref_*.npz,
produced on a B200 from unmodified reference sources by oracle/refbuild/), against the independent anchor values of
SURVEY.md Appendix C, and against analytic known answers."""
import numpy as np
import pytest

import oracle_binding as ob
from common import CONFIGS, apply_cfg, blasius_profiles, conserved, load_golden, relerr


def _oracle_for(name):
    cfg = CONFIGS[name]
    o = ob.Oracle(apply_cfg(ob.params_tgv(24, 3), cfg))
    if cfg["case"] == "tgv":
        o.init_chit()
    elif cfg["case"] == "channel":
        o.init_channel()
    else:
        x, r, u, w, e = blasius_profiles()
        o.set_sponge_from_profiles(x[1:], r[1:], u[1:], w[1:], e[1:])   # quirk Q9: the reference skips the first knot
    return o, cfg


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_oracle_matches_reference_gpu_binary(name):
    o, cfg = _oracle_for(name)
    g = load_golden(name)
    # initial condition + grid: bit-exact (same host arithmetic as src/init.cpp / src/sponge.cu)
    assert np.array_equal(o.x, g["x"])
    for a, b in zip(o.state(), g["file0"]):
        assert np.array_equal(a, b)
    t, p1, p2 = o.run(cfg["nsteps"])
    s1 = conserved(o.state())
    ref1 = conserved(list(g["file1"]))
    o.run(cfg["nsteps"])
    s2 = conserved(o.state()); ref2 = conserved(list(g["file2"]))
    lim = 2e-10 if cfg["case"] == "blayer" else 1e-12     # BL spanwise momentum is a 1e-6-sized response
    for c in range(5):
        den = max(np.abs(ref2[c]).max(), 1e-300)
        assert np.abs(s1[c][::2, ::2, ::2] - ref1[c]).max() / den < lim, (c, "file1")
        assert np.abs(s2[c] - ref2[c]).max() / den < lim, (c, "file2")
    # dt / dpdz printed by the reference at its last CFL refresh (7 significant digits)
    assert abs(o.dt - g["dt_dpdz"][-1, 0]) <= 5e-7 * o.dt
    if cfg["forcing"]:
        assert abs(o.dpdz - g["dt_dpdz"][-1, 1]) <= 5e-7 * abs(o.dpdz)
    # solution.txt row 0: "file*(t+1) time par1 par2 dt" printed with %lf (6 decimals)
    assert abs(p1[0] - g["solution"][0, 2]) < 1e-6


def test_survey_appendix_c_anchors_c1_64():
    """C1 as BASELINE.md section 6 writes it (64^3, s=v=3, low-storage RK3, 100 steps): par1 history, dt at the step-90 refresh and
    samples of the final state from the surveyor's independent numpy restatement (SURVEY.md Appendix C, quirk Q1 on)"""
    o = ob.Oracle(ob.params_tgv(64, 3)); o.init_chit()
    t, p1, _ = o.run(100)
    for i, v in {0: 2.500000000000000e-01, 10: 2.499583342263006e-01, 20: 2.499169187474198e-01, 50: 2.497906456827162e-01,
                 90: 2.496226872797876e-01}.items():
        assert abs(p1[i] - v) < 1e-13, i
    assert abs(o.dt - 4.467409949993332e-03) < 1e-14
    assert abs(o.bulk()[0] - 2.495801472523635e-01) < 1e-13
    st = o.state()
    assert abs(st[0].sum() - 64 ** 3) < 1e-7
    assert abs(np.abs(st[3]).max() - 1.110551585542883e-01) < 1e-13
    assert abs(st[0][0, 0, 0] - 1.005253830519991e+00) < 1e-13 and abs(st[4][3, 5, 7] - 1.790226096680589e+02) < 1e-10
    assert abs(st[4].mean() - 1.786964632771293e+02) < 1e-10


def test_survey_appendix_c_anchors_32():
    """32^3, s=v=3, 20 steps: dt0, dt10 and par1 from the surveyor's independent numpy restatement"""
    o = ob.Oracle(ob.params_tgv(32, 3)); o.init_chit()
    t, p1, _ = o.run(10)
    assert abs(o.dt - 8.936658831646e-03) < 1e-14 and abs(p1[0] - 0.25) < 1e-14
    t, p1, _ = o.run(10)
    assert abs(o.dt - 8.934680637233e-03) < 1e-14
    assert abs(p1[0] - 2.499168281882423e-01) < 1e-13
    assert abs(o.bulk()[0] - 2.498325548148970e-01) < 1e-13


def test_dt0_formula_tgv():
    """dt0 = CFL * dx / (1 + 1/Ma) for the Taylor-Green field (|u|max = 1, c = 1/Ma)  -- SURVEY 7.1(d)"""
    o = ob.Oracle(ob.params_tgv(64, 3)); o.init_chit()
    dx = 2 * np.pi / 64
    assert abs(o.calc_dt() - 4.463954261341933e-03) < 1e-15
    assert abs(o.calc_dt() - 0.5 * dx / (1 + 10.0)) / o.calc_dt() < 2e-2


@pytest.mark.parametrize("s", [1, 2, 3, 4])
def test_operator_order_and_telescoping(s):
    """flux forms approximate -d(fgh)/dx to order 2s and sum to zero over a periodic line (cuda_derivs.h:30-155)"""
    L = ob.lib(); errs = []
    for n in (32, 64):
        x = 2 * np.pi * (np.arange(n) + 0.5) / n; inv = n / (2 * np.pi)
        f = 1 + 0.3 * np.sin(x); g = 0.5 + 0.2 * np.cos(2 * x); h = 2 + 0.1 * np.sin(x + 0.3)
        out = np.zeros(n); d1 = np.zeros(n); d2 = np.zeros(n)
        L.ora_kat_flux_cube(s, n, inv, ob._dp(f), ob._dp(g), ob._dp(h), ob._dp(out))
        exact = -((0.3 * np.cos(x)) * g * h + f * (-0.4 * np.sin(2 * x)) * h + f * g * (0.1 * np.cos(x + 0.3)))
        errs.append(np.abs(out - exact).max())
        assert abs(out.sum()) < 1e-10 * np.abs(out).sum()
        L.ora_kat_d1(s, n, inv, ob._dp(f), ob._dp(d1)); L.ora_kat_d2(s, n, inv * inv, ob._dp(f), ob._dp(d2))
        # derDevShared1x carries the reference's sign convention: coeffF[it]*(f[-l]-f[+l]) = +df/dx
        assert np.abs(d1 - 0.3 * np.cos(x)).max() < 0.3 * (2 * np.pi / n) ** (2 * s) * 2
        assert np.abs(d2 + 0.3 * np.sin(x)).max() < 0.3 * (2 * np.pi / n) ** (2 * s) * 2
    order = np.log2(errs[0] / errs[1])
    assert order > 2 * s - 0.35, order


def test_discrete_conservation_and_quirk_q1():
    """sum(rhs)=0 for rho and rho*u_i on a periodic box; with mu->0 also for rho*E.  With viscosity, the
    reference's y-dissipation quirk (cuda_rhs.cu:175) breaks energy conservation, the corrected form keeps it."""
    o = ob.Oracle(ob.params_tgv(16, 3, Re=1e30)); o.init_chit()
    for k, r in enumerate(o.rhs()):
        assert abs(r.sum()) < 1e-9 * max(np.abs(r).sum(), 1.0), k
    mean = {}
    for q1 in (1, 0):
        o = ob.Oracle(ob.params_tgv(32, 3, quirk_q1=q1)); o.init_chit()
        mean[q1] = abs(o.rhs()[4].mean())
    # corrected form: truncation-level residual (the viscous terms are in expanded, non-telescoping form);
    # reference form: a spurious O(1/Re) energy source, 5 orders of magnitude larger
    assert mean[0] < 1e-9 and mean[1] > 1e4 * mean[0]


def test_rk_temporal_order():
    """fixed dt halving: low-storage and Kutta RK3 converge with order 3, the RK4 extension with order 4"""
    def run(dt, n, **kw):
        o = ob.Oracle(ob.params_tgv(12, 2, **kw)); o.init_chit(); o.set_dt(dt); o.run(n)
        return np.concatenate([a.ravel() for a in conserved(o.state())])
    for kw, expect in ((dict(), 3), (dict(lowStorage=0), 3), (dict(rk4=1), 4)):
        T = 0.02      # dt <= 0.005: well inside the acoustic stability limit of the 12^3 grid
        a, b, c = run(T / 4, 4, **kw), run(T / 8, 8, **kw), run(T / 16, 16, **kw)
        order = np.log2(np.abs(a - b).max() / np.abs(b - c).max())
        assert abs(order - expect) < 0.4, (kw, order)
