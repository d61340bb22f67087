"""cudanavierstokes_b200 -- Python host-side mirror of the libcudns C ABI (include/cudns.h).

The product is the shared library ``libcudns.so`` (hand-written sm_100a CUDA kernels behind a C ABI that
replaces the GPU side of CUDA-DNS: src/main.h:23-42 of the reference).  This module is the thin ctypes
binding used by the tests and bench.py; it keeps the reference's names (setGPUParameters/initSolver ->
``Solver(...)``, copyField -> ``set_state``/``get_state``, solverWrapper's inner loop -> ``advance``).

There is no CPU fallback: importing works without a GPU (so that the symbol table can be checked), but
creating a ``Solver`` raises ``CudnsError`` unless a B200-class device is present.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CUDNS_LIB", os.path.join(_HERE, "libcudns.so"))   # override: tuning builds only
CSRC = os.path.join(_HERE, "csrc")

__all__ = ["Params", "PeerInfo", "Solver", "CudnsError", "lib", "build", "params_tgv", "params_channel", "params_blayer",
           "init_grid", "init_chit", "init_channel", "build_sponge", "write_field", "read_field", "write_xdmf", "blasius_profiles", "stats_write", "EXPORTS"]

# every symbol include/cudns.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "cudns_last_error", "cudns_version", "cudns_params_tgv", "cudns_params_channel", "cudns_params_blayer",
    "cudns_init_grid", "cudns_init_chit", "cudns_init_channel", "cudns_build_sponge", "cudns_write_field",
    "cudns_read_field", "cudns_create", "cudns_destroy", "cudns_memory_report", "cudns_set_state",
    "cudns_get_state", "cudns_set_state_device", "cudns_get_state_device", "cudns_set_sponge", "cudns_advance",
    "cudns_calc_rhs", "cudns_calc_dt", "cudns_calc_bulk", "cudns_get_scalars", "cudns_set_dt",
    "cudns_halo_local_info", "cudns_halo_connect", "cudns_halo_buffers", "cudns_set_allreduce",
    "cudns_set_exchange", "cudns_get_stream", "cudns_get_counters", "cudns_profile_stage",
    "cudns_set_stage_timing", "cudns_get_stage_timing",
    "cudns_write_xdmf", "cudns_write_fields_async", "cudns_io_wait", "cudns_read_fields",
    "cudns_calc_profiles", "cudns_calc_retau", "cudns_blasius_profiles", "cudns_calc_enstrophy",
    "cudns_stats_begin", "cudns_stats_add_mean", "cudns_stats_finish_mean", "cudns_stats_add_fluc", "cudns_stats_get",
    "cudns_stats_write", "cudns_postprocess", "cudns_team_create", "cudns_team_destroy",
]


class CudnsError(RuntimeError):
    pass


class Params(C.Structure):
    """struct cudns_params (include/cudns.h) = the knobs of src/globals.h:14-58, sponge.h, perturbation.h"""
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
        ("nranks", C.c_int), ("rank", C.c_int), ("device", C.c_int),
        ("par2_enstrophy", C.c_int), ("precision", C.c_int), ("reserved", C.c_int * 3),
    ]


class PeerInfo(C.Structure):
    """struct cudns_peer_info (include/cudns.h): what a rank publishes so that its z-slab neighbours can map its state block"""
    _fields_ = [("mem_handle", C.c_ubyte * 64), ("device", C.c_int), ("pid", C.c_int),
                ("local_ptr", C.c_uint64), ("block_bytes", C.c_uint64)]


def build(force=False, verbose=False):
    """compile libcudns.so in-tree for sm_100a (nvcc cross-compiles without a GPU)"""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp", ".h", ".inc", "Makefile"))]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "cudns.h"))
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        out = subprocess.run(["make", "-j8", "-C", CSRC] + (["-B"] if force else []), capture_output=True, text=True)
        if verbose or out.returncode:
            print(out.stdout[-4000:], out.stderr[-4000:])
        if out.returncode:
            raise CudnsError("building libcudns.so failed")
    return LIB_PATH


_lib = None
ALLREDUCE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int, C.c_int)
EXCHANGE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)


def lib():
    """load libcudns.so (fails loudly if it has not been built)"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CudnsError("libcudns.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    dp = C.POINTER(C.c_double)
    PP = C.POINTER(Params)
    H = C.c_void_p
    L.cudns_last_error.restype = C.c_char_p
    L.cudns_version.restype = C.c_char_p
    L.cudns_params_tgv.argtypes = [PP, C.c_int, C.c_int]
    L.cudns_params_channel.argtypes = [PP]
    L.cudns_params_blayer.argtypes = [PP]
    L.cudns_init_grid.argtypes = [PP, dp, dp, dp, dp, dp, dp]
    L.cudns_init_chit.argtypes = [PP, dp, dp, dp, dp, dp, dp, dp, dp]
    L.cudns_init_channel.argtypes = [PP, dp, dp, dp, dp, dp, dp, dp, dp]
    L.cudns_build_sponge.argtypes = [PP, dp, dp, dp, dp, dp, dp, C.c_int, dp, dp, dp, dp, dp, dp, dp, dp]
    L.cudns_write_field.argtypes = [C.c_char_p, C.c_char, C.c_int, dp, C.c_size_t]
    L.cudns_read_field.argtypes = [C.c_char_p, C.c_char, C.c_int, dp, C.c_size_t]
    L.cudns_create.argtypes = [PP, dp, dp, dp, C.POINTER(H)]
    L.cudns_destroy.argtypes = [H]
    L.cudns_memory_report.argtypes = [H, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    L.cudns_set_state.argtypes = [H, dp, dp, dp, dp, dp]
    L.cudns_get_state.argtypes = [H, dp, dp, dp, dp, dp]
    L.cudns_set_state_device.argtypes = [H, C.c_void_p] + [C.c_void_p] * 4
    L.cudns_get_state_device.argtypes = [H, C.c_void_p] + [C.c_void_p] * 4
    L.cudns_set_sponge.argtypes = [H, dp, dp, dp]
    L.cudns_advance.argtypes = [H, C.c_int, dp, dp, dp]
    L.cudns_calc_rhs.argtypes = [H, dp, dp, dp, dp, dp]
    L.cudns_calc_dt.argtypes = [H, dp]
    L.cudns_calc_bulk.argtypes = [H, dp, dp]
    L.cudns_get_scalars.argtypes = [H, dp, dp, dp]
    L.cudns_set_dt.argtypes = [H, C.c_double, C.c_int]
    L.cudns_halo_local_info.argtypes = [H, C.POINTER(PeerInfo)]
    L.cudns_halo_connect.argtypes = [H, C.POINTER(PeerInfo), C.POINTER(PeerInfo)]
    L.cudns_halo_buffers.argtypes = [H, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.cudns_set_allreduce.argtypes = [H, ALLREDUCE_FN, C.c_void_p]
    L.cudns_set_exchange.argtypes = [H, EXCHANGE_FN, C.c_void_p]
    L.cudns_get_stream.argtypes = [H, C.POINTER(C.c_void_p)]
    L.cudns_get_counters.argtypes = [H, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.cudns_profile_stage.argtypes = [H, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.cudns_set_stage_timing.argtypes = [H, C.c_int]
    L.cudns_get_stage_timing.argtypes = [H, dp, dp, dp, C.POINTER(C.c_uint64)]
    L.cudns_write_fields_async.argtypes = [H, C.c_char_p, C.c_int]
    L.cudns_calc_profiles.argtypes = [H, dp]
    L.cudns_blasius_profiles.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, dp, dp, dp, dp, dp]
    L.cudns_calc_retau.argtypes = [H, dp]
    L.cudns_calc_enstrophy.argtypes = [H, dp]
    L.cudns_stats_begin.argtypes = [H, C.c_int]
    for name in ("cudns_stats_add_mean", "cudns_stats_finish_mean", "cudns_stats_add_fluc"):
        getattr(L, name).argtypes = [H]
    L.cudns_stats_get.argtypes = [H, dp, dp, dp, dp, dp]
    L.cudns_stats_write.argtypes = [C.c_char_p, C.c_int, dp, dp, dp, dp, C.c_double, C.c_double]
    L.cudns_postprocess.argtypes = [H, C.c_char_p, C.c_int, C.c_int, dp, C.c_char_p]
    L.cudns_team_create.argtypes = [C.POINTER(H), C.c_int, C.POINTER(C.c_void_p)]
    L.cudns_team_destroy.argtypes = [C.c_void_p]
    L.cudns_io_wait.argtypes = [H, C.POINTER(C.c_uint64)]
    L.cudns_read_fields.argtypes = [H, C.c_char_p, C.c_int]
    L.cudns_write_xdmf.argtypes = [C.c_char_p, C.c_int, dp, C.c_int, dp, C.c_int, dp, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_double, C.c_char_p]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise CudnsError("libcudns error %d: %s" % (rc, lib().cudns_last_error().decode()))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _set(p, over):
    for k, v in over.items():
        setattr(p, k, v)
    return p


def params_tgv(n, stencil, **over):
    p = Params(); _check(lib().cudns_params_tgv(C.byref(p), n, stencil)); return _set(p, over)


def params_channel(**over):
    p = Params(); _check(lib().cudns_params_channel(C.byref(p))); return _set(p, over)


def params_blayer(**over):
    p = Params(); _check(lib().cudns_params_blayer(C.byref(p))); return _set(p, over)


def init_grid(p):
    """initGrid (init.cpp:32-91) -> dict(x, xp, xpp, y, z, dx)"""
    x = np.zeros(p.mx); xp = np.zeros(p.mx); xpp = np.zeros(p.mx); y = np.zeros(p.my); z = np.zeros(p.mz)
    dx = C.c_double(0)
    _check(lib().cudns_init_grid(C.byref(p), _dp(x), _dp(xp), _dp(xpp), _dp(y), _dp(z), C.byref(dx)))
    return dict(x=x, xp=xp, xpp=xpp, y=y, z=z, dx=dx.value)


def _five(p):
    return [np.zeros((p.mz, p.my, p.mx)) for _ in range(5)]


def init_chit(p, grid):
    """initCHIT (init.cpp:126-148): Taylor-Green vortex, global arrays [mz][my][mx]"""
    f = _five(p)
    _check(lib().cudns_init_chit(C.byref(p), _dp(grid["x"]), _dp(grid["y"]), _dp(grid["z"]), *[_dp(a) for a in f]))
    return f


def init_channel(p, grid):
    """initChannel (init.cpp:94-124)"""
    f = _five(p)
    _check(lib().cudns_init_channel(C.byref(p), _dp(grid["x"]), _dp(grid["y"]), _dp(grid["z"]), *[_dp(a) for a in f]))
    return f


def build_sponge(p, grid, xIn, rIn, uIn, wIn, fill_ic=True):
    """calculateSponge host half (sponge.cu:83-195) -> (sigma_x, sigma_z, ref5[5][mz][mx], ic or None)"""
    a = [np.ascontiguousarray(q, dtype=np.float64) for q in (xIn, rIn, uIn, wIn)]
    sx = np.zeros(p.mx); sz = np.zeros(p.mz); ref = np.zeros((5, p.mz, p.mx))
    f = _five(p) if fill_ic else None
    nul = C.POINTER(C.c_double)()
    fa = [_dp(q) for q in f] if fill_ic else [nul] * 5
    _check(lib().cudns_build_sponge(C.byref(p), _dp(grid["x"]), _dp(grid["z"]), *[_dp(q) for q in a], len(a[0]),
                                    _dp(sx), _dp(sz), _dp(ref), *fa))
    return sx, sz, ref, f


def write_field(directory, name, timestep, arr):
    a = np.ascontiguousarray(arr, dtype=np.float64)
    _check(lib().cudns_write_field(directory.encode(), name.encode(), timestep, _dp(a), a.size))


def read_field(directory, name, timestep, shape):
    a = np.zeros(shape)
    _check(lib().cudns_read_field(directory.encode(), name.encode(), timestep, _dp(a), a.size))
    return a


def blasius_profiles(gam=1.4, Ma=0.35, Pr=0.75, n=1000):
    """(x, r, u, w, e) similarity profiles of the boundary-layer inflow (python-utils/selfSimilarSol.py)"""
    out = [np.zeros(n) for _ in range(5)]
    _check(lib().cudns_blasius_profiles(gam, Ma, Pr, n, *[_dp(a) for a in out]))
    return out


def stats_write(outdir, x, mean, fluc, bulk, retau, utau):
    """mean.txt, fluc.txt, bulk.txt in the format of the reference's post-processing tool (Variables::printFile, post.cpp:61-86)"""
    a = [np.ascontiguousarray(q, dtype=np.float64) for q in (x, mean, fluc, bulk)]
    _check(lib().cudns_stats_write(str(outdir).encode(), a[0].size, *[_dp(q) for q in a], float(retau), float(utau)))


def write_xdmf(path, x, y, z, timesteps, dt, names="ruvwe", single_precision=False):
    """XDMF sidecar for the fields/ directory (python-utils/writexmf.py)"""
    x = np.ascontiguousarray(x, dtype=np.float64); y = np.ascontiguousarray(y, dtype=np.float64); z = np.ascontiguousarray(z, dtype=np.float64)
    ts = (C.c_int * len(timesteps))(*[int(t) for t in timesteps])
    _check(lib().cudns_write_xdmf(str(path).encode(), int(single_precision), _dp(x), x.size, _dp(y), y.size, _dp(z), z.size,
                                  ts, len(timesteps), float(dt), names.encode()))


class Solver:
    """One rank's solver = setDevice + setGPUParameters + initSolver of the reference.

    ``p.nranks``/``p.rank`` select a z-slab; host arrays passed to set_state/get_state are the LOCAL slab
    [mz/nranks][my][mx].
    """

    def __init__(self, p, grid=None):
        self.L = lib()
        self.p = p
        self.grid = grid if grid is not None else init_grid(p)
        self.h = C.c_void_p()
        _check(self.L.cudns_create(C.byref(p), _dp(self.grid["x"]), _dp(self.grid["xp"]), _dp(self.grid["xpp"]), C.byref(self.h)))
        self.mzl = p.mz // p.nranks
        self.shape = (self.mzl, p.my, p.mx)
        self._cb = []

    def close(self):
        if getattr(self, "h", None):
            self.L.cudns_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # copyField(0) / copyField(1)
    def set_state(self, arrs):
        a = [np.ascontiguousarray(q, dtype=np.float64) for q in arrs]
        for q in a:
            if q.shape != self.shape:
                raise ValueError("state arrays must have shape %s" % (self.shape,))
        _check(self.L.cudns_set_state(self.h, *[_dp(q) for q in a]))

    def get_state(self):
        out = [np.zeros(self.shape) for _ in range(5)]
        _check(self.L.cudns_get_state(self.h, *[_dp(q) for q in out]))
        return out

    def get_state_into(self, out):
        """copyField(1) into caller-owned (e.g. pinned) float64 arrays of the slab shape"""
        for q in out:
            if q.shape != self.shape or q.dtype != np.float64 or not q.flags["C_CONTIGUOUS"]:
                raise ValueError("output arrays must be C-contiguous float64 of shape %s" % (self.shape,))
        _check(self.L.cudns_get_state(self.h, *[_dp(q) for q in out]))
        return out

    def set_state_device(self, ptrs):
        _check(self.L.cudns_set_state_device(self.h, *[C.c_void_p(int(q)) for q in ptrs]))

    def get_state_device(self, ptrs):
        _check(self.L.cudns_get_state_device(self.h, *[C.c_void_p(int(q)) for q in ptrs]))

    def set_sponge(self, sigma_x, sigma_z, ref5):
        sx = np.ascontiguousarray(sigma_x, dtype=np.float64); sz = np.ascontiguousarray(sigma_z, dtype=np.float64)
        rf = np.ascontiguousarray(ref5, dtype=np.float64)
        if sz.shape != (self.mzl,) or rf.shape != (5, self.mzl, self.p.mx):
            raise ValueError("sponge tables must be local-slab sized")
        _check(self.L.cudns_set_sponge(self.h, _dp(sx), _dp(sz), _dp(rf)))

    # runSimulation[LowStorage]
    def advance(self, nsteps, history=True):
        if history:
            t = np.zeros(nsteps); p1 = np.full(nsteps, np.nan); p2 = np.full(nsteps, np.nan)
            _check(self.L.cudns_advance(self.h, nsteps, _dp(t), _dp(p1), _dp(p2)))
            return t, p1, p2
        nul = C.POINTER(C.c_double)()
        _check(self.L.cudns_advance(self.h, nsteps, nul, nul, nul))
        return None

    def rhs(self):
        out = [np.zeros(self.shape) for _ in range(5)]
        _check(self.L.cudns_calc_rhs(self.h, *[_dp(q) for q in out]))
        return out

    def calc_dt(self):
        d = C.c_double(0); _check(self.L.cudns_calc_dt(self.h, C.byref(d))); return d.value

    def bulk(self):
        a = C.c_double(float("nan")); b = C.c_double(float("nan"))
        _check(self.L.cudns_calc_bulk(self.h, C.byref(a), C.byref(b))); return a.value, b.value

    def scalars(self):
        a = C.c_double(0); b = C.c_double(0); c = C.c_double(0)
        _check(self.L.cudns_get_scalars(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(dt=a.value, dpdz=b.value, time=c.value)

    def set_dt(self, dt, fixed=True):
        _check(self.L.cudns_set_dt(self.h, dt, int(fixed)))

    def memory_report(self):
        a = C.c_size_t(0); b = C.c_size_t(0); c = C.c_size_t(0)
        _check(self.L.cudns_memory_report(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(solver=a.value, free=b.value, total=c.value)

    def counters(self):
        a = C.c_uint64(0); b = C.c_uint64(0)
        _check(self.L.cudns_get_counters(self.h, C.byref(a), C.byref(b)))
        return dict(kernel_launches=a.value, rk_stages=b.value)

    def stream(self):
        s = C.c_void_p(); _check(self.L.cudns_get_stream(self.h, C.byref(s))); return s.value

    def profiles(self):
        """y-z averaged wall-normal profiles (calcAvgChan): rows rho, u~, v~, w~, rho E, then their mean squares"""
        out = np.zeros((10, self.p.mx))
        _check(self.L.cudns_calc_profiles(self.h, _dp(out)))
        return out

    def retau(self):
        v = C.c_double(0.0)
        _check(self.L.cudns_calc_retau(self.h, C.byref(v)))
        return v.value

    def enstrophy(self):
        """mean square vorticity <w.w> of a periodic box (the dissipation measure of a Taylor-Green run; extension)"""
        v = C.c_double(0.0)
        _check(self.L.cudns_calc_enstrophy(self.h, C.byref(v)))
        return v.value

    def post_stats(self, snapshots):
        """postproc/post.cpp over a list of (local-slab) states, reduced on the device: dict(mean[13][mx], fluc[13][mx], bulk[13], Ret, ut)"""
        _check(self.L.cudns_stats_begin(self.h, len(snapshots)))
        for st in snapshots:
            self.set_state(st); _check(self.L.cudns_stats_add_mean(self.h))
        _check(self.L.cudns_stats_finish_mean(self.h))
        for st in snapshots:
            self.set_state(st); _check(self.L.cudns_stats_add_fluc(self.h))
        mean = np.zeros((13, self.p.mx)); fluc = np.zeros((13, self.p.mx)); bulk = np.zeros(13)
        a = C.c_double(0.0); b = C.c_double(0.0)
        _check(self.L.cudns_stats_get(self.h, _dp(mean), _dp(fluc), _dp(bulk), C.byref(a), C.byref(b)))
        return dict(mean=mean, fluc=fluc, bulk=bulk, Ret=a.value, ut=b.value)

    def postprocess(self, directory, first, last, outdir):
        """post.cpp's main(): statistics over fields/<c>.<first..last>.bin -> mean.txt, fluc.txt, bulk.txt in outdir"""
        _check(self.L.cudns_postprocess(self.h, str(directory).encode(), int(first), int(last), _dp(self.grid["x"]), str(outdir).encode()))

    def write_fields_async(self, directory, timestep):
        """snapshot the current state into fields/{r,u,v,w,e}.<timestep>.bin without stalling the step loop"""
        _check(lib().cudns_write_fields_async(self.h, str(directory).encode(), int(timestep)))

    def io_wait(self):
        n = C.c_uint64(0)
        _check(lib().cudns_io_wait(self.h, C.byref(n)))
        return int(n.value)

    def read_fields(self, directory, timestep):
        """restart from fields/{r,u,v,w,e}.<timestep>.bin"""
        _check(lib().cudns_read_fields(self.h, str(directory).encode(), int(timestep)))

    def profile_stage(self, reps=3):
        a = C.c_float(0); b = C.c_float(0); c = C.c_float(0)
        _check(self.L.cudns_profile_stage(self.h, reps, C.byref(a), C.byref(b), C.byref(c)))
        return dict(theta_ms=a.value, rhs_stage_ms=b.value, halo_ms=c.value)

    def stage_timing(self, on=None):
        """on=True/False: switch the per-stage event timing of advance() on (sums reset) / off; on=None: read the sums"""
        if on is not None:
            _check(self.L.cudns_set_stage_timing(self.h, int(bool(on))))
            return None
        a = C.c_double(0); b = C.c_double(0); c = C.c_double(0); n = C.c_uint64(0)
        _check(self.L.cudns_get_stage_timing(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(n)))
        k = max(int(n.value), 1)
        return dict(stages=int(n.value), theta_ms=a.value / k, rhs_stage_ms=b.value / k, halo_ms=c.value / k)

    def halo_buffers(self):
        p = [C.c_void_p() for _ in range(4)]; n = C.c_size_t(0)
        _check(self.L.cudns_halo_buffers(self.h, *[C.byref(q) for q in p], C.byref(n)))
        return [q.value for q in p], n.value

    def halo_local_info(self):
        """bytes of this rank's cudns_peer_info (CUDA IPC handle of the state block), to be sent to the slab neighbours"""
        pi = PeerInfo(); _check(self.L.cudns_halo_local_info(self.h, C.byref(pi)))
        return bytes(pi)

    def halo_connect(self, lower, upper):
        """map the neighbours' state blocks (bytes from their halo_local_info): from now on the stage kernel writes their
        ghost planes directly over NVLink and no pack / exchange / unpack kernels run inside the step loop"""
        lo = PeerInfo.from_buffer_copy(lower); up = PeerInfo.from_buffer_copy(upper)
        _check(self.L.cudns_halo_connect(self.h, C.byref(lo), C.byref(up)))

    def set_allreduce(self, fn):
        """fn(device_ptr:int, n:int, op:int) with op 0 min, 1 sum, 2 max; must act on the solver's stream"""
        cb = ALLREDUCE_FN(lambda user, ptr, n, op: fn(ptr, n, op))
        self._cb.append(cb)
        _check(self.L.cudns_set_allreduce(self.h, cb, None))

    def set_exchange(self, fn):
        """fn(stream:int): move send_lo->lower.recv_hi and send_hi->upper.recv_lo on that stream"""
        cb = EXCHANGE_FN(lambda user, stream: fn(stream))
        self._cb.append(cb)
        _check(self.L.cudns_set_exchange(self.h, cb, None))
