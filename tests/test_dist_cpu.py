"""CPU tests of the z-slab plumbing (cudanavierstokes_b200/dist.py) with two gloo ranks: the pairing of the four halo
blocks, the periodic neighbour topology (the role of splitComm, src/comm.cpp:144-203) and the host-side exchange that
mirrors updateHaloFive (src/comm.cpp:114-134).  No GPU, no compute calls into libcudns."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from cudanavierstokes_b200 import dist as cdist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_topology_and_slabs():
    assert cdist.neighbours(0, 4) == (3, 1) and cdist.neighbours(3, 4) == (2, 0) and cdist.neighbours(0, 1) == (0, 0)
    assert cdist.slab(2, 4, 64) == (32, 48)
    with pytest.raises(ValueError):
        cdist.slab(0, 3, 64)
    a = np.arange(8 * 2 * 2).reshape(8, 2, 2)
    parts = cdist.split_field(a, 4)
    assert len(parts) == 4 and np.array_equal(np.concatenate(parts), a)


@pytest.mark.parametrize("n", [2, 3, 8])
def test_exchange_pairs_match_up(n):
    """every send has exactly one matching receive on the peer, with the same tag, and the blocks land on the right side"""
    sends, recvs = [], []
    for r in range(n):
        for kind, peer, buf, tag in cdist.exchange_pairs(r, n):
            (sends if kind == "send" else recvs).append((r, peer, buf, tag))
    for src, dst, buf, tag in sends:
        want = "recv_hi" if buf == "send_lo" else "recv_lo"     # my low planes are the lower neighbour's upper ghosts
        assert (dst, src, want, tag) in recvs


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import numpy as np, torch, torch.distributed as dist
    from cudanavierstokes_b200 import dist as cdist
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    mz, my, mx, g = 12, 5, 6, 3                                   # global planes, ghost depth
    glob = np.arange(mz * my * mx, dtype=np.float64).reshape(mz, my, mx) * 0.25 + 1.0
    k0, k1 = cdist.slab(rank, world, mz)
    mine = glob[k0:k1]
    bufs = {"send_lo": torch.from_numpy(mine[:g].copy()), "send_hi": torch.from_numpy(mine[-g:].copy()),
            "recv_lo": torch.zeros(g, my, mx, dtype=torch.float64), "recv_hi": torch.zeros(g, my, mx, dtype=torch.float64)}
    cdist.host_exchange(bufs, rank, world, dist)
    lo_expect = glob[[(k0 - g + i) %% mz for i in range(g)]]      # periodic wrap, like MPI_Cart_create(periods=1)
    hi_expect = glob[[(k1 + i) %% mz for i in range(g)]]
    assert np.array_equal(bufs["recv_lo"].numpy(), lo_expect), "lower ghosts"
    assert np.array_equal(bufs["recv_hi"].numpy(), hi_expect), "upper ghosts"
    t = torch.tensor([float(rank + 1)], dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MIN); assert t.item() == 1.0
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


@pytest.mark.parametrize("world", [2, 4])
def test_host_exchange_two_gloo_ranks(tmp_path, world):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("ok") == world
