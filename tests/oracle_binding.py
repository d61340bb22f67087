"""ctypes binding of the CPU oracle (oracle/libcudns_oracle.so).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg.  Never by the product package.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB = os.path.join(ORACLE_DIR, "libcudns_oracle.so")


class OraParams(C.Structure):
    _fields_ = [
        ("mx", C.c_int), ("my", C.c_int), ("mz", C.c_int),
        ("stencilSize", C.c_int), ("stencilVisc", C.c_int),
        ("Lx", C.c_double), ("Ly", C.c_double), ("Lz", C.c_double),
        ("CFL", C.c_double),
        ("lowStorage", C.c_int), ("boundaryLayer", C.c_int), ("perturbed", C.c_int),
        ("forcing", C.c_int), ("periodicX", C.c_int), ("nonUniformX", C.c_int),
        ("checkCFLcondition", C.c_int), ("checkBulk", C.c_int),
        ("Re", C.c_double), ("Pr", C.c_double), ("Ma", C.c_double), ("viscexp", C.c_double), ("gam", C.c_double),
        ("stretch", C.c_double), ("TwallTop", C.c_double), ("TwallBot", C.c_double),
        ("spTopStr", C.c_double), ("spTopLen", C.c_double), ("spTopExp", C.c_double),
        ("spInlStr", C.c_double), ("spInlLen", C.c_double), ("spInlExp", C.c_double),
        ("spOutStr", C.c_double), ("spOutLen", C.c_double), ("spOutExp", C.c_double),
        ("kC", C.c_int), ("LP", C.c_int),
        ("amp1", C.c_double), ("amp2", C.c_double), ("omega1", C.c_double), ("omega2", C.c_double),
        ("quirk_q1", C.c_int), ("rk4", C.c_int),
    ]


def build(force=False):
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(ORACLE_DIR, "cudns_oracle.c")):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = C.CDLL(_LIB)
        P = C.POINTER(OraParams)
        dp = C.POINTER(C.c_double)
        L.ora_params_tgv.argtypes = [P, C.c_int, C.c_int]
        L.ora_params_channel.argtypes = [P]
        L.ora_params_blayer.argtypes = [P]
        L.ora_create.argtypes = [P]; L.ora_create.restype = C.c_void_p
        L.ora_destroy.argtypes = [C.c_void_p]
        L.ora_set_threads.argtypes = [C.c_int]
        L.ora_get_threads.restype = C.c_int
        for name in ("ora_x", "ora_xp", "ora_xpp", "ora_y", "ora_z", "ora_dxv", "ora_coeffVSx",
                     "ora_r", "ora_u", "ora_v", "ora_w", "ora_e", "ora_spongeX", "ora_spongeZ"):
            f = getattr(L, name); f.argtypes = [C.c_void_p]; f.restype = dp
        L.ora_ref.argtypes = [C.c_void_p, C.c_int]; L.ora_ref.restype = dp
        L.ora_derived.argtypes = [C.c_void_p, C.c_int]; L.ora_derived.restype = dp
        L.ora_dx.argtypes = [C.c_void_p]; L.ora_dx.restype = C.c_double
        for name in ("ora_init_chit", "ora_init_channel", "ora_copy_field_in", "ora_calc_state"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.ora_set_sponge_from_profiles.argtypes = [C.c_void_p, dp, dp, dp, dp, dp, C.c_int, C.c_int]
        L.ora_calc_rhs.argtypes = [C.c_void_p, C.POINTER(dp)]
        L.ora_calc_dt.argtypes = [C.c_void_p]; L.ora_calc_dt.restype = C.c_double
        L.ora_calc_bulk.argtypes = [C.c_void_p, dp, dp]
        L.ora_calc_enstrophy.argtypes = [C.c_void_p]; L.ora_calc_enstrophy.restype = C.c_double
        L.ora_calc_profiles.argtypes = [C.c_void_p, dp]
        L.ora_calc_retau.argtypes = [C.c_void_p]; L.ora_calc_retau.restype = C.c_double
        L.ora_run.argtypes = [C.c_void_p, C.c_int, dp, dp, dp]
        L.ora_post_create.argtypes = [C.c_void_p, C.c_int]; L.ora_post_create.restype = C.c_void_p
        L.ora_post_destroy.argtypes = [C.c_void_p]
        L.ora_post_add_mean.argtypes = [C.c_void_p, C.c_void_p]
        L.ora_post_finish_mean.argtypes = [C.c_void_p]
        L.ora_post_add_fluc.argtypes = [C.c_void_p, C.c_void_p]
        L.ora_post_get.argtypes = [C.c_void_p, dp, dp, dp, dp, dp]
        for name in ("ora_get_dt", "ora_get_dpdz", "ora_get_time"):
            f = getattr(L, name); f.argtypes = [C.c_void_p]; f.restype = C.c_double
        L.ora_set_dt.argtypes = [C.c_void_p, C.c_double]
        L.ora_set_fixed_dt.argtypes = [C.c_void_p, C.c_int]
        L.ora_kat_flux_cube.argtypes = [C.c_int, C.c_int, C.c_double, dp, dp, dp, dp]
        L.ora_kat_flux_quad.argtypes = [C.c_int, C.c_int, C.c_double, dp, dp, dp]
        L.ora_kat_d1.argtypes = [C.c_int, C.c_int, C.c_double, dp, dp]
        L.ora_kat_d2.argtypes = [C.c_int, C.c_int, C.c_double, dp, dp]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def set_threads(n=None):
    """use n host threads (default: every core this process may run on) whatever OMP_NUM_THREADS says; returns the count in effect"""
    if n is None:
        n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().ora_set_threads(int(n))
    return int(lib().ora_get_threads())


def params_tgv(n, stencil, **over):
    p = OraParams(); lib().ora_params_tgv(C.byref(p), n, stencil)
    for k, v in over.items():
        setattr(p, k, v)
    return p


def params_channel(**over):
    p = OraParams(); lib().ora_params_channel(C.byref(p))
    for k, v in over.items():
        setattr(p, k, v)
    return p


def params_blayer(**over):
    p = OraParams(); lib().ora_params_blayer(C.byref(p))
    for k, v in over.items():
        setattr(p, k, v)
    return p


class Oracle:
    """Thin object wrapper; arrays are numpy views on the oracle's own storage, shape (mz,my,mx)."""

    def __init__(self, params):
        self.L = lib()
        self.p = params
        self.h = self.L.ora_create(C.byref(params))
        if not self.h:
            raise ValueError("ora_create failed (bad stencil sizes?)")
        self.shape = (params.mz, params.my, params.mx)
        self.N = params.mx * params.my * params.mz

    def close(self):
        if self.h:
            self.L.ora_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _view(self, ptr, shape):
        n = int(np.prod(shape))
        return np.ctypeslib.as_array(ptr, shape=(n,)).reshape(shape)

    # grid
    @property
    def x(self): return self._view(self.L.ora_x(self.h), (self.p.mx,))
    @property
    def xp(self): return self._view(self.L.ora_xp(self.h), (self.p.mx,))
    @property
    def xpp(self): return self._view(self.L.ora_xpp(self.h), (self.p.mx,))
    @property
    def y(self): return self._view(self.L.ora_y(self.h), (self.p.my,))
    @property
    def z(self): return self._view(self.L.ora_z(self.h), (self.p.mz,))
    @property
    def dxv(self): return self._view(self.L.ora_dxv(self.h), (self.p.mx,))
    @property
    def dx(self): return self.L.ora_dx(self.h)
    # state
    @property
    def r(self): return self._view(self.L.ora_r(self.h), self.shape)
    @property
    def u(self): return self._view(self.L.ora_u(self.h), self.shape)
    @property
    def v(self): return self._view(self.L.ora_v(self.h), self.shape)
    @property
    def w(self): return self._view(self.L.ora_w(self.h), self.shape)
    @property
    def e(self): return self._view(self.L.ora_e(self.h), self.shape)

    def state(self):
        return [a.copy() for a in (self.r, self.u, self.v, self.w, self.e)]

    def set_state(self, arrs):
        for dst, src in zip((self.r, self.u, self.v, self.w, self.e), arrs):
            dst[...] = src
        self.L.ora_copy_field_in(self.h)

    def derived(self, which):
        return self._view(self.L.ora_derived(self.h, which), self.shape)

    def init_chit(self):
        self.L.ora_init_chit(self.h); self.L.ora_copy_field_in(self.h)

    def init_channel(self):
        self.L.ora_init_channel(self.h); self.L.ora_copy_field_in(self.h)

    def set_sponge_from_profiles(self, xIn, rIn, uIn, wIn, eIn, fill_ic=True):
        a = [np.ascontiguousarray(q, dtype=np.float64) for q in (xIn, rIn, uIn, wIn, eIn)]
        self.L.ora_set_sponge_from_profiles(self.h, *[_dp(q) for q in a], len(a[0]), int(fill_ic))
        if fill_ic:
            self.L.ora_copy_field_in(self.h)

    @property
    def spongeX(self): return self._view(self.L.ora_spongeX(self.h), (self.p.mx,))
    @property
    def spongeZ(self): return self._view(self.L.ora_spongeZ(self.h), (self.p.mz,))
    def ref(self, which): return self._view(self.L.ora_ref(self.h, which), (self.p.mz, self.p.mx))

    def rhs(self):
        out = [np.zeros(self.shape) for _ in range(5)]
        arr = (C.POINTER(C.c_double) * 5)(*[_dp(o) for o in out])
        self.L.ora_calc_rhs(self.h, arr)
        return out

    def calc_dt(self): return self.L.ora_calc_dt(self.h)

    def bulk(self):
        a = C.c_double(0.0); b = C.c_double(0.0)
        self.L.ora_calc_bulk(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def profiles(self):
        out = np.zeros((10, self.p.mx))
        self.L.ora_calc_profiles(self.h, _dp(out))
        return out

    def retau(self): return self.L.ora_calc_retau(self.h)

    def enstrophy(self): return self.L.ora_calc_enstrophy(self.h)

    def post_stats(self, snapshots):
        """postproc/post.cpp over a list of states: dict(mean[13][mx], fluc[13][mx], bulk[13], Ret, ut)"""
        P = self.L.ora_post_create(self.h, len(snapshots))
        for st in snapshots:
            self.set_state(st); self.L.ora_post_add_mean(P, self.h)
        self.L.ora_post_finish_mean(P)
        for st in snapshots:
            self.set_state(st); self.L.ora_post_add_fluc(P, self.h)
        mean = np.zeros((13, self.p.mx)); fluc = np.zeros((13, self.p.mx)); bulk = np.zeros(13)
        a = C.c_double(0.0); b = C.c_double(0.0)
        self.L.ora_post_get(P, _dp(mean), _dp(fluc), _dp(bulk), C.byref(a), C.byref(b))
        self.L.ora_post_destroy(P)
        return dict(mean=mean, fluc=fluc, bulk=bulk, Ret=a.value, ut=b.value)

    def run(self, nsteps):
        t = np.zeros(nsteps); p1 = np.full(nsteps, np.nan); p2 = np.full(nsteps, np.nan)
        self.L.ora_run(self.h, nsteps, _dp(t), _dp(p1), _dp(p2))
        return t, p1, p2

    @property
    def dt(self): return self.L.ora_get_dt(self.h)
    @property
    def dpdz(self): return self.L.ora_get_dpdz(self.h)
    def set_dt(self, dt, fixed=True):
        self.L.ora_set_dt(self.h, dt); self.L.ora_set_fixed_dt(self.h, int(fixed))
