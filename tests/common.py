"""Helpers shared by the parity tests: build an oracle and a libcudns solver from the same knobs."""
import numpy as np

import cudanavierstokes_b200 as cd
import oracle_binding as ob
from ref_cases import apply_cfg, blasius_profiles, load_golden, CONFIGS  # noqa: F401

FIELDS = ("r", "u", "v", "w", "e")


def copy_params(src, dst):
    """copy every same-named field between an oracle params struct and a libcudns one"""
    dst_names = {n for n, *_ in dst._fields_}
    for n, *_ in src._fields_:
        if n in dst_names and n != "reserved":
            setattr(dst, n, getattr(src, n))
    return dst


def make_pair(op):
    """op: oracle params -> (Oracle, Solver) sharing the oracle's grid"""
    o = ob.Oracle(op)
    cp = copy_params(op, cd.Params()); cp.nranks = 1; cp.rank = 0; cp.device = 0
    grid = cd.init_grid(cp)
    return o, cd.Solver(cp, grid), grid


def relerr(a, b, floor=0.0):
    """max-norm error relative to max|b| (per conserved variable)"""
    den = max(np.abs(b).max(), floor, 1e-300)
    return np.abs(np.asarray(a) - np.asarray(b)).max() / den


def cons_errs(got, ref):
    """per conserved variable: max-norm error relative to max|ref| -- the three momentum components relative to the largest of them
    (the Taylor-Green start has rho*w = 0: after a few steps it is a dt-sized response whose round-off, measured against ITS OWN
    maximum, says nothing about the accuracy of the momentum vector; BASELINE.md section 6 asks 1e-12 on the conserved variables)"""
    a, b = conserved(got), conserved(ref)
    mom = max(np.abs(b[k]).max() for k in (1, 2, 3))
    return [relerr(a[0], b[0]), relerr(a[1], b[1], floor=mom), relerr(a[2], b[2], floor=mom), relerr(a[3], b[3], floor=mom), relerr(a[4], b[4])]


def conserved(st):
    r, u, v, w, e = st
    return [r, r * u, r * v, r * w, e]


def smooth_random_state(o, seed=1234, amp=0.1):
    """SURVEY 8(d) robustness input: low-wavenumber random rho, T, u_i with rho,T > 0"""
    rng = np.random.default_rng(seed)
    p = o.p
    X = 2 * np.pi * (np.arange(p.mx) + 0.5) / p.mx
    Y = 2 * np.pi * (np.arange(p.my) + 0.5) / p.my
    Z = 2 * np.pi * (np.arange(p.mz) + 0.5) / p.mz
    zz, yy, xx = np.meshgrid(Z, Y, X, indexing="ij")

    def field(mean):
        f = np.full(xx.shape, mean)
        for _ in range(6):
            k = rng.integers(-3, 4, size=3); ph = rng.uniform(0, 2 * np.pi); a = rng.uniform(-1, 1) * amp / 3
            f = f + a * np.cos(k[0] * xx + k[1] * yy + k[2] * zz + ph)
        return f
    Rgas = 1.0 / (p.gam * p.Ma * p.Ma)
    r = field(1.0); T = field(1.0); u = field(0.0); v = field(0.0); w = field(0.0)
    e = r * (Rgas * T / (p.gam - 1.0) + 0.5 * (u * u + v * v + w * w))
    return [r, u, v, w, e]
