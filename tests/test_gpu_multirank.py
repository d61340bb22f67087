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


def _run_two_ranks(p_of, full, grid, steps, peer):
    """advance the same problem as 2 slabs in two threads; returns the gathered state"""
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
            sols[r].set_state([a[r * mzl:(r + 1) * mzl] for a in full])
            sols[r].advance(steps)
            out[r] = sols[r].get_state()
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


@pytest.mark.timeout(600)
def test_two_gpus_torchrun_peer_memory():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "mgpu_check.py"), "48", "6", "peer", "tgv"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=540)
    assert out.returncode == 0 and "mgpu_check" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
