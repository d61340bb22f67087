"""Single precision (`myprec float`, src/globals.h:5-6; cudns_params.precision = 1): the float copy of the device side -- stage kernel,
dilatation pass, reductions, solver object -- against the FP64 oracle.

What single precision can deliver here: a float carries 6e-8 of the value.  The state is held to a few units of that per step.  The
instantaneous right-hand side is not: at Ma = 0.1 the pressure is ~1/(gam Ma^2) = 71 while its variation over the box is ~1e-2, so the
pressure gradient formed from neighbouring float values carries ~71 * 6e-8 / (a_l dx) of noise -- 1e-4..1e-3 of max|rhs| whatever
the kernel does (the reference built with myprec float has the same property).  The tolerances below are the measured errors with a
margin of ~3 and are there to catch defects (a wrong coefficient or index shows up at 1e-2 and above), not to certify digits."""
import numpy as np
import pytest

import cudanavierstokes_b200 as cd
import oracle_binding as ob
from common import cons_errs, copy_params, relerr, smooth_random_state

pytestmark = pytest.mark.gpu
EPS32 = float(np.finfo(np.float32).eps)


def make_pair32(op):
    o = ob.Oracle(op)
    cp = copy_params(op, cd.Params()); cp.nranks = 1; cp.rank = 0; cp.device = 0; cp.precision = 1
    grid = cd.init_grid(cp)
    return o, cd.Solver(cp, grid), grid


def test_state_round_trip_is_the_float_rounding_of_the_input():
    op = ob.params_tgv(24, 3, mx=40, my=20, mz=24)
    o, s, grid = make_pair32(op)
    st = smooth_random_state(o); s.set_state(st)
    got = s.get_state()
    for a, b in zip(got, st):
        assert np.array_equal(a, b.astype(np.float32).astype(np.float64))
    s.close()


@pytest.mark.parametrize("sv", [(1, 1), (2, 2), (3, 2), (3, 3), (4, 2), (4, 4)])
def test_rhs_f32_vs_fp64_oracle(sv):
    op = ob.params_tgv(24, sv[0], stencilVisc=sv[1], Ma=0.5)        # Ma = 0.5: pressure variations within 1e-5 of float resolution
    o, s, grid = make_pair32(op)
    st = smooth_random_state(o); o.set_state(st); s.set_state(st)
    a = s.rhs(); b = o.rhs()
    errs = [relerr(x, y) for x, y in zip(a, b)]
    print("rhs f32", sv, ["%.1e" % e for e in errs])
    assert max(errs) < 2e-4, errs
    s.close()


@pytest.mark.parametrize("scheme", ["ls3", "kutta", "rk4"])
@pytest.mark.parametrize("n,s_,nsteps", [(32, 4, 10), (40, 3, 5)])
def test_steps_f32_vs_fp64_oracle(scheme, n, s_, nsteps):
    op = ob.params_tgv(n, s_, lowStorage=int(scheme == "ls3"), rk4=int(scheme == "rk4"))
    o, s, grid = make_pair32(op)
    o.init_chit(); s.set_state(o.state())
    o.run(nsteps); s.advance(nsteps)
    errs = cons_errs(s.get_state(), o.state())
    print("steps f32", scheme, n, s_, nsteps, ["%.1e" % e for e in errs])
    # rho and rho*E: a few float roundings per stage; momentum: the Ma = 0.1 pressure-gradient noise integrated over the steps
    assert errs[0] < 20 * EPS32 and errs[4] < 20 * EPS32, errs
    assert max(errs[1:4]) < 2e-4, errs
    t32 = s.scalars()["time"]
    assert abs(t32 - o.L.ora_get_time(o.h)) < 1e-5 * t32           # dt from the float state, accumulated in double
    s.close()


def test_f32_history_and_diagnostics():
    op = ob.params_tgv(32, 3, checkBulk=2)
    o, s, grid = make_pair32(op)
    o.init_chit(); s.set_state(o.state())
    assert abs(s.bulk()[0] - o.bulk()[0]) < 1e-6 * o.bulk()[0]        # <u.u>: double sum over float values
    assert abs(s.enstrophy() - o.enstrophy()) < 1e-5 * o.enstrophy()
    assert abs(s.calc_dt() - o.calc_dt()) < 1e-6 * o.calc_dt()
    t, p1, p2 = s.advance(4)
    to, p1o, p2o = o.run(4)
    assert np.allclose(p1[::2], p1o[::2], rtol=1e-5) and np.isnan(p1[1::2]).all()
    s.close()


@pytest.mark.parametrize("name", ["chan_s3v2", "chan_s2v2", "bl_s3v2"])
def test_f32_walls_stretched_grid_and_boundary_layer_vs_reference_gpu_binary(name):
    """the set-ups the lean kernel serves (isothermal walls / boundary layer with sponge and blowing-suction, tanh-stretched x,
    T^0.75 and T^1.5 viscosity) in single precision, from the reference's own file 0 to its file 2 (FP64 goldens)"""
    from common import CONFIGS, apply_cfg, blasius_profiles, conserved, load_golden
    cfg = CONFIGS[name]; g = load_golden(name)
    op = apply_cfg(ob.params_tgv(24, 3), cfg)
    cp = apply_cfg(cd.Params(), cfg); cp.gam = 1.4; cp.TwallTop = cp.TwallBot = 1.0; cp.quirk_q1 = 1; cp.nranks = 1; cp.precision = 1
    for k in ("spTopStr", "spTopLen", "spTopExp", "spInlStr", "spInlLen", "spInlExp", "spOutStr", "spOutLen", "spOutExp",
              "kC", "LP", "amp1", "amp2", "omega2"):
        setattr(cp, k, getattr(op, k))
    grid = cd.init_grid(cp)
    s = cd.Solver(cp, grid)
    if cfg["case"] == "blayer":
        x, r, u, w, e = blasius_profiles()
        sx, sz, ref, ic = cd.build_sponge(cp, grid, x[1:], r[1:], u[1:], w[1:])
        s.set_sponge(sx, sz, ref)
    s.set_state(list(g["file0"]))
    s.advance(cfg["nsteps"]); s.advance(cfg["nsteps"])
    got = conserved(s.get_state()); ref = conserved(list(g["file2"]))
    mom = max(np.abs(ref[k]).max() for k in (1, 2, 3))
    errs = [relerr(got[0], ref[0])] + [relerr(got[k], ref[k], floor=mom) for k in (1, 2, 3)] + [relerr(got[4], ref[4])]
    print("f32", name, ["%.1e" % e for e in errs])
    assert errs[0] < 2e-5 and errs[4] < 2e-5 and max(errs[1:4]) < 2e-5, errs       # measured: 4e-6, 4e-6, 5e-6
    sc = s.scalars()
    assert abs(sc["dt"] - g["dt_dpdz"][-1, 0]) <= 1e-4 * sc["dt"]
    if cfg["forcing"]:
        assert abs(sc["dpdz"] - g["dt_dpdz"][-1, 1]) <= 1e-3 * abs(sc["dpdz"])
    s.close()


@pytest.mark.parametrize("case", ["channel_linear_s4v4", "channel_linear_s1v1", "periodic_sqrt_visc_s3v3", "periodic_odd_tiles_s4v2"])
def test_f32_lean_kernel_variants_vs_fp64_oracle(case):
    """the remaining variants of the lean kernel in single precision: 8 staged quantities with walls (linear viscosity law), 9
    without walls (periodic box, mu ~ sqrt(T)), and a periodic box the fifth generation does not take (mx = 2 mod 4 is refused,
    mx = 36: ragged tiles)"""
    if case.startswith("channel"):
        s_, v_ = (4, 4) if case.endswith("s4v4") else (1, 1)
        op = apply_cfg_channel(s_, v_)
        o, s, grid = make_pair32(op)
        o.init_channel(); st = o.state()
    elif case.startswith("periodic_sqrt"):
        op = ob.params_tgv(24, 3, mx=32, my=20, mz=24, viscexp=0.5, Ma=0.5)
        o, s, grid = make_pair32(op)
        st = smooth_random_state(o)
    else:
        op = ob.params_tgv(24, 4, stencilVisc=2, mx=36, my=20, mz=24, viscexp=0.75, Ma=0.5)
        o, s, grid = make_pair32(op)
        st = smooth_random_state(o)
    o.set_state(st); s.set_state(st)
    a = s.rhs(); b = o.rhs()
    rerr = [relerr(x, y, floor=max(np.abs(z).max() for z in b[1:4]) if 1 <= k <= 3 else 0.0) for k, (x, y) in enumerate(zip(a, b))]
    o.run(5); s.advance(5)
    errs = cons_errs(s.get_state(), o.state())
    print("f32 lean", case, "rhs", ["%.1e" % e for e in rerr], "5 steps", ["%.1e" % e for e in errs])
    assert max(rerr) < 2e-5, rerr                                                    # measured: 4e-6
    assert errs[0] < 5e-6 and errs[4] < 5e-6 and max(errs[1:4]) < 1e-5, errs        # measured: 6e-7, 4e-7, 2e-6
    s.close()


def apply_cfg_channel(s_, v_):
    from common import CONFIGS, apply_cfg
    op = apply_cfg(ob.params_tgv(24, 3), dict(CONFIGS["chan_s3v2"], stencilSize=s_, stencilVisc=v_, viscexp=1.0, mx=64, my=24, mz=32))
    return op


def test_f32_refuses_a_row_length_it_cannot_stage():
    """single precision outside the fifth generation's set-ups needs 16-byte rows: mx % 4 == 0"""
    op = ob.params_tgv(24, 3, mx=38, my=20, mz=24, viscexp=0.5)
    cp = copy_params(op, cd.Params()); cp.nranks = 1; cp.precision = 1
    with pytest.raises(cd.CudnsError):
        cd.Solver(cp)
