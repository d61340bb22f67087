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


def test_f32_refuses_what_it_is_not_built_for():
    op = ob.params_channel()
    cp = copy_params(op, cd.Params()); cp.nranks = 1; cp.precision = 1
    cp.mx, cp.my, cp.mz = 32, 24, 24
    with pytest.raises(cd.CudnsError):
        cd.Solver(cp)
