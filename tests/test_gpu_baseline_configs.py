"""GPU parity at the sizes BASELINE.json / BASELINE.md section 6 name (the goldens of tests/golden are 24^3-sized):

* C1 as written: decaying Taylor-Green vortex 64^3, low-storage RK3, 6th order (s = v = 3), FP64, N = 1 / 10 / 100 steps against
  the oracle (<= 1e-12 relative on the conserved variables) and against the anchor values of SURVEY.md Appendix C;
* a 128^3 8th-order step test of the kernel bench.py times: several z chunks AND full tiles, all three schemes;
* size-independent properties at the bench size itself (512^3 would take the CPU oracle minutes): see test_bench_size_properties."""
import numpy as np
import pytest

import cudanavierstokes_b200 as cd
import oracle_binding as ob
from common import cons_errs, conserved, make_pair, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-12


def test_c1_tgv64_rk3_6th_order_1_10_100_steps():
    op = ob.params_tgv(64, 3)                       # C1: s = v = 3, low-storage RK3, Re 1600, Ma 0.1, CFL 0.5, checkCFL = checkBulk = 10
    for n in (1, 10, 100):                          # each from the initial condition, ONE advance call (the dt cadence restarts per call, quirk Q11)
        o, s, grid = make_pair(op)
        o.init_chit(); s.set_state(o.state())
        t0, a1, _ = o.run(n); t1, b1, _ = s.advance(n)
        errs = cons_errs(s.get_state(), o.state())
        assert max(errs) < TOL, (n, errs)
        assert abs(s.scalars()["dt"] - o.dt) <= 1e-13 * o.dt, n
        np.testing.assert_allclose(t1, t0, rtol=0, atol=1e-12)
        # kinetic-energy history par1 = <u.u> at the checkBulk steps (calc_stress.cu:192-196): "identical histories" (north_star)
        for i in range(0, n, 10):
            assert abs(b1[i] - a1[i]) <= 1e-13, (n, i)
        if n < 100:
            s.close()
    # SURVEY Appendix C (independent numpy restatement, quirk Q1 on = reference behaviour): dt and par1 at the refresh steps ...
    anchors = {0: 2.500000000000000e-01, 10: 2.499583342263006e-01, 50: 2.497906456827162e-01, 90: 2.496226872797876e-01}
    for i, v in anchors.items():
        assert abs(b1[i] - v) < 2e-13, (i, b1[i])
    assert abs(s.scalars()["dt"] - 4.467409949993332e-03) < 1e-14            # refreshed at step 90
    # ... and the state after 100 steps
    assert abs(s.bulk()[0] - 2.495801472523635e-01) < 2e-13
    assert abs(o.bulk()[0] - 2.495801472523635e-01) < 2e-13
    st = s.get_state()
    assert abs(st[0].sum() - 64 ** 3) < 1e-7                                   # mass is conserved to round-off
    assert abs(np.abs(st[3]).max() - 1.110551585542883e-01) < 1e-12
    assert abs(st[0][0, 0, 0] - 1.005253830519991e+00) < 1e-12
    assert abs(st[4][3, 5, 7] - 1.790226096680589e+02) < 1e-9
    s.close()


@pytest.mark.parametrize("scheme", ["lowstorage", "kutta", "rk4"])
def test_tgv128_8th_order_steps_vs_oracle(scheme):
    """128^3, s = v = 4: the configuration of the bench kernel at a size with four 32-plane z chunks and only full tiles"""
    op = ob.params_tgv(128, 4, lowStorage=int(scheme == "lowstorage"), rk4=int(scheme == "rk4"))
    o, s, grid = make_pair(op)
    o.init_chit(); s.set_state(o.state())
    a = s.rhs(); b = o.rhs()
    assert max(relerr(x, y) for x, y in zip(a, b)) < 1e-11
    o.run(3); s.advance(3)
    errs = cons_errs(s.get_state(), o.state())
    assert max(errs) < TOL, errs


@pytest.mark.parametrize("scheme", ["lowstorage", "rk4", "lowstorage_f32"])
def test_bench_size_properties(scheme):
    """512 x 512 x 64 slab of the bench grid (full 512 x 512 planes = the bench's tiles and TMA boxes; 64 planes keep the host arrays
    small).  Size-independent properties: (1) mass and momentum sums are conserved by a step to round-off (the split form
    telescopes), (2) a z-independent start stays z-independent although every plane meets a different ring phase / z chunk /
    prologue of the stage kernel, (3) with a fixed dt two calls of one step equal one call of two steps bit for bit."""
    f32 = scheme.endswith("_f32")                       # the single-precision copy of the same kernels: float round-off instead of double
    p = cd.params_tgv(512, 4, mz=64, lowStorage=int(scheme.startswith("lowstorage")), rk4=int(scheme == "rk4"), precision=int(f32))
    p.Lz = 2 * np.pi * 64 / 512                         # same spacing in z as in x and y
    p.nranks = 1
    grid = cd.init_grid(p)
    st0 = cd.init_chit(p, grid)
    # one z period of the Taylor-Green field is 2 pi: restrict the start to a z-independent vortex sheet so that the slab is periodic
    zfix = [np.repeat(a[:1], 64, axis=0) for a in st0]
    s = cd.Solver(p, grid); s.set_state(zfix)
    if f32:
        zfix = [a.astype(np.float32).astype(np.float64) for a in zfix]     # what the device holds
    m0 = [c.sum() for c in conserved(zfix)]
    s.advance(2)
    a = s.get_state()
    m1 = [c.sum() for c in conserved(a)]
    N = 512 * 512 * 64
    assert abs(m1[0] - m0[0]) < (1e-7 if f32 else 1e-13) * abs(m0[0]), (m1[0], m0[0])
    for k in (1, 2, 3):
        # (single precision: ~1e-6 of rounding per point after two steps -- the z-independent, x-y-symmetric start makes the roundings of
        # all symmetric copies of a point identical, so they add up coherently, not like a random walk: bound 1e-7 per point)
        assert abs(m1[k] - m0[k]) < (1e-7 * N if f32 else 1e-11 * N ** 0.5), (k, m1[k], m0[k])
    # z-independent data stay z-independent (every plane goes through a different ring phase / chunk / prologue of the kernel)
    for f in a:
        assert np.abs(f - f[:1]).max() <= (2e-6 if f32 else 1e-13) * max(np.abs(f).max(), 1.0), np.abs(f - f[:1]).max()
    s.close()
    # (3) with a fixed dt, advance(2) == advance(1) + advance(1) bit for bit (buffer rotation, aux-field validity across calls)
    s3 = cd.Solver(p, grid); s3.set_state(zfix); s3.set_dt(1e-3); s3.advance(2); c1 = s3.get_state(); s3.close()
    s4 = cd.Solver(p, grid); s4.set_state(zfix); s4.set_dt(1e-3); s4.advance(1); s4.advance(1); c2 = s4.get_state(); s4.close()
    for x, y in zip(c1, c2):
        assert np.array_equal(x, y)
