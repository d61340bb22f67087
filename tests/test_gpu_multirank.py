"""GPU tests of the z-slab decomposition (multi-GPU must equal single-GPU).

* two ranks inside ONE process on ONE device, each driven by its own host thread: exercises the slab logic, the stage
  kernel's stores into the neighbour's ghost planes (peer memory = the other solver's block, same device) and the
  device-side epoch hand-shake on any single-GPU box;
* the same through torch.distributed / CUDA IPC / NCCL with one process per GPU when the box has at least two GPUs."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

import cudanavierstokes_b200 as cd
from common import conserved, relerr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _DevMem:
    def __init__(self, ptr, ndoubles):
        self.__cuda_array_interface__ = {"shape": (ndoubles,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def _run_two_ranks(p_of, full, grid, steps, peer, sponge=None, scalars=None, after=None, extras=None):
    """advance the same problem as 2 slabs in two threads; returns the gathered state.  sponge = (sigma_x, sigma_z, ref5) GLOBAL
    tables (each rank takes its slab); scalars: optional list that receives rank 0's dt / dpdz / time after the run; after(solver,
    rank) runs on every rank's thread after the steps (collective calls: both ranks make them in the same order), its results land
    in extras[rank]"""
    import torch
    nr = 2
    sols = [cd.Solver(p_of(nr, r), grid) for r in range(nr)]
    bar = threading.Barrier(nr)
    shared = {"red": [None] * nr}
    bufs = []
    for s in sols:
        ptrs, nbytes = s.halo_buffers()
        bufs.append({k: torch.as_tensor(_DevMem(q, nbytes // 8), device="cuda") for k, q in zip(("send_lo", "send_hi", "recv_lo", "recv_hi"), ptrs)})
    streams = [torch.cuda.ExternalStream(s.stream()) for s in sols]
    errors = []

    def make_callbacks(r):
        lo, up = (r - 1) % nr, (r + 1) % nr

        def exchange(_stream):
            streams[r].synchronize(); bar.wait()                 # every rank's send blocks are packed
            with torch.cuda.stream(streams[r]):
                bufs[r]["recv_hi"].copy_(bufs[up]["send_lo"]); bufs[r]["recv_lo"].copy_(bufs[lo]["send_hi"])
            streams[r].synchronize(); bar.wait()

        def allreduce(ptr, count, op):
            t = torch.as_tensor(_DevMem(ptr, count), device="cuda")
            streams[r].synchronize()
            shared["red"][r] = t.clone(); bar.wait()
            st = torch.stack(shared["red"])
            res = st.min(0).values if op == 0 else st.sum(0) if op == 1 else st.max(0).values
            bar.wait()
            with torch.cuda.stream(streams[r]):
                t.copy_(res)
            streams[r].synchronize()
        return exchange, allreduce

    out = [None] * nr

    def worker(r):
        try:
            ex, ar = make_callbacks(r)
            sols[r].set_exchange(ex); sols[r].set_allreduce(ar)
            if peer:
                infos = [s.halo_local_info() for s in sols]
                sols[r].halo_connect(infos[(r - 1) % nr], infos[(r + 1) % nr])
            bar.wait()
            mzl = sols[r].mzl
            if sponge is not None:
                sx, sz, ref = sponge
                sols[r].set_sponge(sx, sz[r * mzl:(r + 1) * mzl], ref[:, r * mzl:(r + 1) * mzl])
            sols[r].set_state([a[r * mzl:(r + 1) * mzl] for a in full])
            sols[r].advance(steps)
            out[r] = sols[r].get_state()
            if scalars is not None and r == 0:
                scalars.append(sols[r].scalars())
            if after is not None:
                extras[r] = after(sols[r], r)
        except Exception as e:       # noqa: BLE001
            errors.append(e); bar.abort()

    th = [threading.Thread(target=worker, args=(r,)) for r in range(nr)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=150)
    alive = any(t.is_alive() for t in th)
    assert not alive, "multi-rank run hung"
    assert not errors, errors
    for s in sols:
        s.close()
    return [np.concatenate([out[r][f] for r in range(nr)]) for f in range(5)]


@pytest.mark.timeout(240)
@pytest.mark.parametrize("peer", [True, False])
@pytest.mark.parametrize("case", ["tgv_s4v4", "tgv_s3v2_visc07_kutta"])
def test_two_slabs_equal_one(case, peer):
    def p_of(nranks, rank):
        if case == "tgv_s4v4":
            p = cd.params_tgv(32, 4, mz=48)
        else:
            p = cd.params_tgv(32, 3, stencilVisc=2, viscexp=0.7, lowStorage=0, mz=40)
        p.nranks = nranks; p.rank = rank; p.device = 0
        return p
    p1 = p_of(1, 0)
    grid = cd.init_grid(p1)
    full = cd.init_chit(p1, grid)
    ref = cd.Solver(p1, grid); ref.set_state(full); ref.advance(12); single = ref.get_state(); ref.close()
    multi = _run_two_ranks(p_of, full, grid, 12, peer)
    errs = [relerr(a, b) for a, b in zip(conserved(multi), conserved(single))]
    assert max(errs) < 1e-13, errs          # multi-GPU == single-GPU to 1e-13 (observed: identical)


def _golden_params(name, prec=0):
    import oracle_binding as ob
    from common import CONFIGS, apply_cfg
    cfg = CONFIGS[name]
    op = apply_cfg(ob.params_tgv(24, 3), cfg)

    def p_of(nranks, rank):
        cp = apply_cfg(cd.Params(), cfg); cp.gam = 1.4; cp.TwallTop = cp.TwallBot = 1.0; cp.quirk_q1 = 1
        for k in ("spTopStr", "spTopLen", "spTopExp", "spInlStr", "spInlLen", "spInlExp", "spOutStr", "spOutLen", "spOutExp",
                  "kC", "LP", "amp1", "amp2", "omega2"):
            setattr(cp, k, getattr(op, k))
        cp.nranks = nranks; cp.rank = rank; cp.device = 0; cp.precision = prec
        return cp
    return cfg, p_of


@pytest.mark.timeout(300)
@pytest.mark.parametrize("peer,prec", [(True, 0), (False, 0), (True, 1)])
def test_two_slabs_channel_forcing(peer, prec):
    """isothermal-wall channel with the pressure-gradient controller: the bulk integrals (calcPressureGrad calc_stress.cu:98-120) are
    cross-rank SUMs, dt a cross-rank MAX; walls in x, stretched grid, periodic z across the slab seam"""
    cfg, p_of = _golden_params("chan_s3v2", prec)
    p1 = p_of(1, 0)
    grid = cd.init_grid(p1)
    full = cd.init_channel(p1, grid)
    ref = cd.Solver(p1, grid); ref.set_state(full); ref.advance(12); single = ref.get_state(); sc1 = ref.scalars(); ref.close()
    sc2 = []
    multi = _run_two_ranks(p_of, full, grid, 12, peer, scalars=sc2)
    errs = [relerr(a, b) for a, b in zip(conserved(multi), conserved(single))]
    # the forcing integrals are summed per slab and then across ranks: a different order than one device's single sum, so dpdz (and
    # with it the state) agrees to round-off, not bit for bit
    # (single precision: dpdz still agrees to double round-off -- the integrals are double sums --, but one flipped float rounding
    # of dt * dpdz is a 6e-8 difference in the state)
    assert max(errs) < (1e-12 if prec == 0 else 2e-6), errs
    assert abs(sc2[0]["dpdz"] - sc1["dpdz"]) <= (1e-12 if prec == 0 else 1e-6) * abs(sc1["dpdz"])
    assert abs(sc2[0]["dt"] - sc1["dt"]) <= (1e-14 if prec == 0 else 1e-6) * sc1["dt"]


@pytest.mark.timeout(300)
@pytest.mark.parametrize("peer,prec", [(True, 0), (False, 0), (True, 1)])
def test_two_slabs_boundary_layer(peer, prec):
    """spatially developing boundary layer: z is NOT periodic (the global bottom / top slabs have no neighbour there and rebuild the
    extrapolation ghosts on chip, api.cu ghost_targets / handshake), sponges and wall blowing/suction indexed by the GLOBAL plane"""
    from common import blasius_profiles
    cfg, p_of = _golden_params("bl_s3v2", prec)
    p1 = p_of(1, 0)
    grid = cd.init_grid(p1)
    x, r, u, w, e = blasius_profiles()
    sx, sz, refq, ic = cd.build_sponge(p1, grid, x[1:], r[1:], u[1:], w[1:])
    ref = cd.Solver(p1, grid); ref.set_sponge(sx, sz, refq); ref.set_state(ic); ref.advance(12); single = ref.get_state(); ref.close()
    multi = _run_two_ranks(p_of, ic, grid, 12, peer, sponge=(sx, sz, refq))
    errs = [relerr(a, b, floor=1e-30) for a, b in zip(conserved(multi), conserved(single))]
    assert max(errs) < 1e-13, errs


@pytest.mark.timeout(300)
@pytest.mark.parametrize("prec", [0, 1])
def test_two_slabs_enstrophy_and_single_precision(prec):
    """the dissipation measure <w.w> is a cross-rank SUM; precision = 1: the float copy of the device side through the same slab
    logic (ghost stores into the neighbour's block, hand-shake, scalar reductions in double)"""
    def p_of(nranks, rank):
        p = cd.params_tgv(32, 3, mz=48, precision=prec, checkBulk=2, par2_enstrophy=1)
        p.nranks = nranks; p.rank = rank; p.device = 0
        return p
    p1 = p_of(1, 0)
    grid = cd.init_grid(p1)
    full = cd.init_chit(p1, grid)
    ref = cd.Solver(p1, grid); ref.set_state(full); t1, a1, b1 = ref.advance(6); single = ref.get_state(); e1 = ref.enstrophy(); ref.close()
    extras = [None, None]
    multi = _run_two_ranks(p_of, full, grid, 6, True, after=lambda sol, r: sol.enstrophy(), extras=extras)
    errs = [relerr(a, b) for a, b in zip(conserved(multi), conserved(single))]
    assert max(errs) < 1e-13, errs
    assert abs(extras[0] - e1) <= 1e-13 * e1 and extras[0] == extras[1]
    assert np.isfinite(b1[::2]).all() and np.isnan(b1[1::2]).all()         # par2 = <w.w> at the checkBulk steps only


@pytest.mark.timeout(300)
def test_two_slabs_post_statistics():
    """postproc/post.cpp as device reductions across two slabs (partial sums per rank, cross-rank SUM through the callback): equal to
    one slab and to the reference's own tool"""
    import oracle_binding  # noqa: F401  (tests/ on the path)
    from common import load_golden
    from ref_cases import GOLDEN
    cfg, p_of = _golden_params("chan_s3v2")
    g = load_golden("chan_s3v2")
    snaps = [list(g["file0"]), list(g["file2"])]
    p1 = p_of(1, 0)
    grid = cd.init_grid(p1)
    ref = cd.Solver(p1, grid); one = ref.post_stats(snaps); ref.close()

    def stats(sol, r):
        mzl = sol.mzl
        return sol.post_stats([[a[r * mzl:(r + 1) * mzl] for a in st] for st in snaps])
    extras = [None, None]
    _run_two_ranks(p_of, snaps[0], grid, 0, False, after=stats, extras=extras)
    for key in ("mean", "fluc", "bulk"):
        scale = np.abs(one[key]).max(axis=-1, keepdims=True) if key != "bulk" else np.abs(one[key])
        if key == "fluc":
            scale = np.maximum(scale, 1e-3 * np.abs(one["mean"]).max(axis=-1, keepdims=True) ** 2)
        assert (np.abs(extras[0][key] - one[key]) <= 1e-12 * np.maximum(scale, 1e-300)).all(), key
        assert np.array_equal(extras[0][key], extras[1][key])                 # every rank holds the reduced result
    assert abs(extras[0]["Ret"] - one["Ret"]) <= 1e-12 * one["Ret"]
    tool = np.load(os.path.join(GOLDEN, "ref_post_chan_s3v2.npz"))
    assert abs(extras[0]["Ret"] - float(tool["Ret"])) <= 1e-6


@pytest.mark.timeout(300)
@pytest.mark.parametrize("case", ["tgv_ls3", "tgv_rk4_f32"])
def test_team_two_slabs_in_one_process(case):
    """cudns_team_create: the library's own wiring of n solvers driven by n host threads of one process (peer mapping + host-side
    all-reduce / exchange callbacks) -- no torch, no NCCL: equal to one slab"""
    import ctypes as C

    def p_of(nranks, rank):
        p = cd.params_tgv(32, 4, mz=48, precision=int(case.endswith("f32")), lowStorage=int("ls3" in case), rk4=int("rk4" in case))
        p.nranks = nranks; p.rank = rank; p.device = 0
        return p
    p1 = p_of(1, 0)
    grid = cd.init_grid(p1)
    full = cd.init_chit(p1, grid)
    ref = cd.Solver(p1, grid); ref.set_state(full); ref.advance(8); single = ref.get_state(); b1 = ref.bulk()[0]; ref.close()
    sols = [cd.Solver(p_of(2, r), grid) for r in range(2)]
    arr = (C.c_void_p * 2)(*[s.h for s in sols])
    team = C.c_void_p()
    assert cd.lib().cudns_team_create(arr, 2, C.byref(team)) == 0, cd.lib().cudns_last_error()
    out, bulk, errors = [None, None], [None, None], []

    def worker(r):
        try:
            mzl = sols[r].mzl
            sols[r].set_state([a[r * mzl:(r + 1) * mzl] for a in full])
            sols[r].advance(8)
            out[r] = sols[r].get_state(); bulk[r] = sols[r].bulk()[0]
        except Exception as e:       # noqa: BLE001
            errors.append(e)
    th = [threading.Thread(target=worker, args=(r,)) for r in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=120)
    assert not any(t.is_alive() for t in th), "team run hung"
    assert not errors, errors
    for s in sols:
        s.close()
    cd.lib().cudns_team_destroy(team)
    multi = [np.concatenate([out[r][f] for r in range(2)]) for f in range(5)]
    errs = [relerr(a, b) for a, b in zip(conserved(multi), conserved(single))]
    assert max(errs) < 1e-13, errs
    assert bulk[0] == bulk[1] and abs(bulk[0] - b1) <= 1e-13 * b1


@pytest.mark.timeout(300)
def test_driver_two_slabs_reproduce_the_reference_channel_run(tmp_path):
    """cudns_run ngpus=2 (two host threads, two slabs; samedevice=1 puts both on this box's GPU): the channel golden of the reference's
    own GPU binary, fields written by both slabs into the same files"""
    from common import CONFIGS, load_golden
    name = "chan_s3v2"
    cfg = CONFIGS[name]; g = load_golden(name)
    out = tmp_path / "run"
    os.makedirs(out / "fields")
    for c, a in zip("ruvwe", g["file0"]):
        np.ascontiguousarray(a).tofile(out / "fields" / ("%s.0000000.bin" % c))
    keys = ("mx", "my", "mz", "stencilSize", "stencilVisc", "Lx", "Ly", "Lz", "CFL", "checkCFLcondition", "checkBulk", "Re", "Pr", "Ma",
            "viscexp", "stretch", "forcing", "periodicX", "nonUniformX", "lowStorage")
    exe = os.path.join(ROOT, "cudanavierstokes_b200", "cudns_run")
    args = [exe, "case=channel", "restartFile=0", "nfiles=2", "nsteps=%d" % cfg["nsteps"], "outdir=%s" % out, "ngpus=2", "samedevice=1"]
    args += ["%s=%r" % (k, cfg[k]) for k in keys]
    r = subprocess.run(args, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    got = [np.fromfile(out / "fields" / ("%s.0000002.bin" % c)).reshape(g["file2"][0].shape) for c in "ruvwe"]
    errs = [relerr(a, b) for a, b in zip(conserved(got), conserved(list(g["file2"])))]
    assert max(errs) < 5e-12, errs
    sol = np.loadtxt(out / "solution.txt")
    assert np.allclose(sol[:, 1:], g["solution"][:, 1:], rtol=0, atol=2e-6)
    assert "on 2 GPUs" in r.stdout


@pytest.mark.timeout(600)
def test_two_gpus_torchrun_peer_memory():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "mgpu_check.py"), "48", "6", "peer", "tgv"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=540)
    assert out.returncode == 0 and "mgpu_check" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
