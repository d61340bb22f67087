"""z-slab decomposition plumbing: one process per GPU, torch.distributed for the transport.

Replaces the host half of the reference's halo exchange (updateHalo / updateHaloFive, src/comm.cpp:90-134: pack ->
D2H -> 4 x MPI_Sendrecv -> H2D -> unpack) and its scalar reductions (allReduceToMin / allReduceSum / allReduceArray,
src/comm.cpp:294-335).  The library packs the first / last (s+v) interior planes of the five state fields into two
contiguous device blocks; this module moves them between slab neighbours (NCCL send/recv over NVLink on GPU ranks, gloo
in the CPU tests) on the solver's own stream, so that the exchange is ordered with the kernels around it.

Topology (replaces splitComm, src/comm.cpp:144-203): rank r owns planes [r*mz/n, (r+1)*mz/n); lower = (r-1) mod n,
upper = (r+1) mod n (periodic, like the reference's MPI_Cart_create with periods = 1).
"""
import numpy as np

__all__ = ["neighbours", "slab", "exchange_pairs", "host_exchange", "attach"]

OPS = {0: "min", 1: "sum", 2: "max"}


def neighbours(rank, nranks):
    """(lower, upper) slab neighbours, periodic"""
    return (rank - 1) % nranks, (rank + 1) % nranks


def slab(rank, nranks, mz):
    """[k0, k1) global plane range of this rank"""
    if mz % nranks:
        raise ValueError("mz must be divisible by the number of ranks")
    n = mz // nranks
    return rank * n, (rank + 1) * n


def exchange_pairs(rank, nranks):
    """Ordered list of (kind, peer, buffer, tag): my send_lo block fills the lower neighbour's upper ghost planes
    (its recv_hi), my send_hi block the upper neighbour's lower ghost planes (its recv_lo).  The order is the same on
    every rank, which is what makes the 2-rank case (lower == upper) match up."""
    lo, up = neighbours(rank, nranks)
    return [("send", lo, "send_lo", 0), ("send", up, "send_hi", 1),
            ("recv", up, "recv_hi", 0), ("recv", lo, "recv_lo", 1)]


def host_exchange(bufs, rank, nranks, dist):
    """run the exchange on CPU tensors (gloo); bufs: dict name -> torch tensor.  Used by the CPU tests."""
    reqs = []
    for kind, peer, name, tag in exchange_pairs(rank, nranks):
        if kind == "send":
            reqs.append(dist.isend(bufs[name], dst=peer, tag=tag))
        else:
            reqs.append(dist.irecv(bufs[name], src=peer, tag=tag))
    for r in reqs:
        r.wait()


class _DevMem:
    """a raw device pointer dressed up for torch.as_tensor"""

    def __init__(self, ptr, ndoubles):
        self.__cuda_array_interface__ = {"shape": (ndoubles,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def attach(solver, dist=None, peer=True):
    """wire a Solver (nranks > 1) to torch.distributed.

    * scalar reductions (dt: MIN/MAX, bulk/forcing: SUM): NCCL all_reduce on the solver's stream;
    * halo transport inside the step loop: peer memory (``peer=True``, default) -- the ranks swap the CUDA IPC handles of
      their state blocks once (all_gather_object), after which the stage kernel stores its boundary planes straight into
      the neighbours' ghost planes over NVLink and a device-side flag replaces the exchange; if the handles cannot be
      mapped (or ``peer=False``) the stage falls back to pack -> NCCL send/recv -> unpack on the solver's stream;
    * the one-off ghost fill of cudns_set_state always uses the send/recv path.

    Everything is enqueued on the solver's stream (torch.cuda.ExternalStream), nothing synchronises the host.
    """
    import torch
    import torch.distributed as td
    dist = dist or td
    p = solver.p
    rank, nranks = p.rank, p.nranks
    ptrs, nbytes = solver.halo_buffers()
    n = nbytes // 8
    dev = torch.device("cuda", p.device)
    names = ("send_lo", "send_hi", "recv_lo", "recv_hi")
    bufs = {k: torch.as_tensor(_DevMem(q, n), device=dev) for k, q in zip(names, ptrs)}
    ext = torch.cuda.ExternalStream(solver.stream(), device=dev)
    redop = {0: td.ReduceOp.MIN, 1: td.ReduceOp.SUM, 2: td.ReduceOp.MAX}

    def exchange(stream):
        with torch.cuda.stream(ext):
            ops = []
            for kind, peer, name, _tag in exchange_pairs(rank, nranks):
                ops.append(td.P2POp(td.isend if kind == "send" else td.irecv, bufs[name], peer))
            for r in td.batch_isend_irecv(ops):
                r.wait()          # stream-ordered wait for NCCL work (does not block the host)

    def allreduce(ptr, count, op):
        t = torch.as_tensor(_DevMem(ptr, count), device=dev)
        with torch.cuda.stream(ext):
            dist.all_reduce(t, op=redop[op])

    solver.set_exchange(exchange)
    solver.set_allreduce(allreduce)
    solver._dist_keep = (bufs, ext)
    solver.peer_transport = False
    if peer:
        infos = [None] * nranks
        td.all_gather_object(infos, solver.halo_local_info())
        lo, up = neighbours(rank, nranks)
        ok = 1
        try:
            solver.halo_connect(infos[lo], infos[up])
        except Exception as e:                      # e.g. no peer access between the two devices
            ok = 0; solver._peer_error = str(e)
        flag = torch.tensor([ok], device=dev, dtype=torch.int32)
        td.all_reduce(flag, op=td.ReduceOp.MIN)     # all ranks or none: the hand-shake needs both sides
        if int(flag.item()) != 1:
            raise RuntimeError("peer-memory halo transport unavailable on some rank: %s" % getattr(solver, "_peer_error", "(other rank)"))
        solver.peer_transport = True
    return solver


def split_field(a, nranks):
    """global [mz][my][mx] -> list of slabs"""
    return np.split(np.asarray(a), nranks, axis=0)
