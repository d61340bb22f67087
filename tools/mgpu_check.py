"""Multi-GPU check (run under torchrun, one rank per GPU): z-slab run == single-GPU run, and per-stage timing.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/mgpu_check.py [n] [steps] [peer|nccl] [case]

Rank 0 also runs the whole grid on its own GPU with nranks = 1 and compares the gathered slabs (SURVEY 8e:
multi-GPU must equal single-GPU; tolerance 1e-13 relative on the conserved variables)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import cudanavierstokes_b200 as cd
from cudanavierstokes_b200 import dist as cdist

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
mode = sys.argv[3] if len(sys.argv) > 3 else "peer"
case = sys.argv[4] if len(sys.argv) > 4 else "tgv"
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def params(nranks, r):
    if case == "tgv":
        p = cd.params_tgv(n, 4)
    elif case == "tgv_s3v2_visc07":
        p = cd.params_tgv(n, 3, stencilVisc=2, viscexp=0.7)
    elif case == "kutta":
        p = cd.params_tgv(n, 4, lowStorage=0)
    elif case == "tgv_f32":
        p = cd.params_tgv(n, 4, precision=1)
    elif case == "rk4_f32":
        p = cd.params_tgv(n, 4, precision=1, lowStorage=0, rk4=1)
    else:
        raise SystemExit("unknown case")
    p.nranks = nranks; p.rank = r; p.device = local
    return p


p = params(world, rank)
grid = cd.init_grid(p)
full = cd.init_chit(params(1, 0), grid)                       # global initial condition (host)
mzl = n // world
slab = [a[rank * mzl:(rank + 1) * mzl] for a in full]
sol = cd.Solver(p, grid)
cdist.attach(sol, peer=(mode == "peer"))
sol.set_state(slab)
sol.advance(steps)
mine = np.stack(sol.get_state())
gathered = [torch.empty((5, mzl, n, n), dtype=torch.float64, device="cuda") for _ in range(world)] if rank == 0 else None
dist.gather(torch.from_numpy(mine).cuda(), gathered, dst=0)
# timing of the step loop (device events on the solver's stream, max over ranks)
stream = torch.cuda.ExternalStream(sol.stream(), device=torch.device("cuda", local))
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0.record(stream); sol.advance(steps, history=False); e1.record(stream)
torch.cuda.synchronize(); dist.barrier()
t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
prof = sol.profile_stage(3)
if rank == 0:
    multi = torch.cat(gathered, dim=1).cpu().numpy()
    ref = cd.Solver(params(1, 0), grid); ref.set_state(full); ref.advance(steps); single = np.stack(ref.get_state()); ref.close()
    cons = lambda s: [s[0], s[0] * s[1], s[0] * s[2], s[0] * s[3], s[4]]
    errs = [float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)) for a, b in zip(cons(multi), cons(single))]
    stages = 4 if case.startswith("rk4") else 3
    print("mgpu_check n=%d ranks=%d mode=%s case=%s steps=%d: max rel diff vs single GPU %s | %.3f ms/step -> %.2f Gpts*stage/s | stage kernels %s peer=%s"
          % (n, world, mode, case, steps, ["%.1e" % e for e in errs], t.item() / steps, n ** 3 * stages * steps / (t.item() * 1e-3) / 1e9, prof,
             getattr(sol, "peer_transport", None)), flush=True)
    assert max(errs) < 1e-13, errs
sol.close()
dist.destroy_process_group()
