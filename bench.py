#!/usr/bin/env python
"""bench.py -- throughput of the right-hand-side + Runge-Kutta path (libcudns, C ABI) on B200.

Metric (BASELINE.json): Mpts*RK-stage/s = grid points x RK stages advanced per second of the step loop (dt reductions,
bulk diagnostics and halo exchange included, H<->D copies and file I/O excluded), plus the achieved HBM GB/s of the
dominant kernel against the measured roofline.  Workload: Taylor-Green vortex, FP64, 8th order (stencilSize =
stencilVisc = 4), default 512^3 (BASELINE config 5 / north_star target; every field is 1 GiB, far larger than the
126 MB L2, so consecutive steps never find their inputs in cache).  One "step" = one time step = 3 RK stages
(low-storage RK3, the reference's default scheme) or 4 (--scheme rk4, the extension north_star names).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n 512] [--scheme ls3|kutta3|rk4] [--impl reference]

N > 1: launched by torch.distributed.run, one rank per GPU, z-slab decomposition of the SAME global grid (strong
scaling); the stage kernel stores its boundary planes straight into the neighbours' ghost planes over NVLink peer memory
(CUDA IPC handles swapped once through torch.distributed), scalar reductions go through NCCL on the solver's stream.

--impl reference: the reference has no CPU implementation; the arm times the CPU oracle
(oracle/, an OpenMP restatement of the reference's algorithm, kind "port") on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STAGES = {"ls3": 3, "kutta3": 3, "rk4": 4}
# algorithmic bytes per point per stage: 2-register scheme 20 words, 3-register 25 words
ALG_BYTES = {"ls3": 160.0, "kutta3": 200.0, "rk4": 200.0}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_traffic(kernel_key):
    """dram bytes per launch of the dominant kernel from the committed ncu summary (profiles/), or None"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel_key)
    except Exception:
        return None


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.dev = dev; self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                      stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            self.p.kill(); out = ""
        sm, mx, pw, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in out.strip().splitlines():
            c = [t.strip() for t in line.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0])); mx.append(float(c[1])); pw.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def scheme_params(cd, n, scheme):
    p = cd.params_tgv(n, 4)
    p.lowStorage = 1 if scheme == "ls3" else 0
    p.rk4 = 1 if scheme == "rk4" else 0
    return p


def tgv_slab(out, grid, p, k0, mzl):
    """Taylor-Green initial condition of planes [k0, k0+mzl) (the formulas of initCHIT, init.cpp:126-148), a few planes at a
    time so that the temporaries stay small next to the slab itself (1024^3 on 8 ranks: 5.4 GB of pinned state per rank)"""
    import numpy as np
    Rgas = float(np.float32(1.0) / np.float64(p.gam * p.Ma * p.Ma))
    fx = 2 * np.pi * grid["x"] / p.Lx; fy = 2 * np.pi * grid["y"] / p.Ly
    sx, cx, c2x = np.sin(fx)[None, None, :], np.cos(fx)[None, None, :], np.cos(2 * fx)[None, None, :]
    sy, cy, c2y = np.sin(fy)[None, :, None], np.cos(fy)[None, :, None], np.cos(2 * fy)[None, :, None]
    uxy, vxy, pxy = sx * cy, -cx * sy, c2x + c2y
    r, u, v, w, e = out
    for a in range(0, mzl, 8):
        b = min(a + 8, mzl)
        fz = 2 * np.pi * grid["z"][k0 + a:k0 + b] / p.Lz
        cz, c2z = np.cos(fz)[:, None, None], np.cos(2 * fz)[:, None, None]
        u[a:b] = uxy * cz
        v[a:b] = vxy * cz
        w[a:b] = 0.0
        press = Rgas + (1.0 / 16.0) * pxy * (c2z + 2.0)
        r[a:b] = press / Rgas
        e[a:b] = press / (p.gam - 1.0) + 0.5 * r[a:b] * (u[a:b] * u[a:b] + v[a:b] * v[a:b])


def cpu_oracle_rate(scheme, budget_s=12.0, n=128):
    """time the CPU oracle (all host threads) on a bounded sample of the workload: same physics/stencil, n^3 grid"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    op = ob.params_tgv(n, 4, lowStorage=int(scheme == "ls3"), rk4=int(scheme == "rk4"))
    o = ob.Oracle(op); o.init_chit()
    t0 = time.perf_counter(); o.run(1); t1 = time.perf_counter() - t0
    steps = max(1, min(40, int(budget_s / max(t1, 1e-3))))
    t0 = time.perf_counter(); o.run(steps); dt = time.perf_counter() - t0
    o.close()
    rate = n ** 3 * STAGES[scheme] * steps / dt / 1e6
    return rate, steps, dt, n


def run_reference(args):
    """--impl reference: the oracle port on the host cores; each "step" = a bounded sample (one time step of the n=128
    sub-problem)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    n = args.ref_n
    scheme = args.scheme
    op = ob.params_tgv(n, 4, lowStorage=int(scheme == "ls3"), rk4=int(scheme == "rk4"))
    o = ob.Oracle(op); o.init_chit()
    for _ in range(args.warmup):
        o.run(1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.run(1)
    dt = time.perf_counter() - t0
    cores = os.cpu_count()
    val = n ** 3 * STAGES[scheme] * args.steps / dt / 1e6
    sample = "TGV %d^3 (same physics, stencil and scheme as the %d^3 workload), 1 time step per bench step" % (n, args.n)
    line = {"impl": "reference", "metric": "Mpts*RK-stage/s", "value": val, "unit": "Mpts*RK-stage/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "tgv%d_s4v4_fp64_%s" % (args.n, scheme), "scheme": scheme, "stencilSize": 4, "stencilVisc": 4},
            "cpu_baseline": {"value": val, "unit": "Mpts*RK-stage/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "Mpts*RK-stage/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n", "--grid", dest="n", type=int, default=512, help="grid is n^3 (use --grid under torchrun: its parser "
                    "treats a bare --n as an abbreviation of its own options)")
    ap.add_argument("--scheme", default="ls3", choices=sorted(STAGES))
    ap.add_argument("--impl", default="cudns", choices=["cudns", "reference"])
    ap.add_argument("--ref-n", type=int, default=128)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cudns" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import cudanavierstokes_b200 as cd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libcudns has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world != args.gpus and rank == 0:
        print("warning: --gpus %d but WORLD_SIZE %d; using WORLD_SIZE" % (args.gpus, world), file=sys.stderr)

    n, scheme = args.n, args.scheme
    stages = STAGES[scheme]
    p = scheme_params(cd, n, scheme)
    p.nranks = world; p.rank = rank; p.device = local
    grid = cd.init_grid(p)
    sol = cd.Solver(p, grid)
    if world > 1:
        from cudanavierstokes_b200 import dist as cdist
        cdist.attach(sol)
    mzl = n // world
    # Taylor-Green initial condition of this rank's slab (the formulas of initCHIT, init.cpp:126-148, vectorised)
    pin = torch.empty((5, mzl, n, n), dtype=torch.float64).pin_memory()
    host = pin.numpy()
    tgv_slab(host, grid, p, rank * mzl, mzl)
    views = [host[f] for f in range(5)]
    sol.set_state(views)
    stream = torch.cuda.ExternalStream(sol.stream(), device=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sol.advance(args.warmup, history=False)
    c0 = sol.counters()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    ev0.record(stream)
    sol.advance(args.steps, history=False)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    c1 = sol.counters()
    if dist is not None:                                           # max over ranks (also: every rank must derive the same `extra`)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # nvidia-smi delivers a sample every ~100 ms: when the timed region was shorter than that (many GPUs, small K), the same
    # workload keeps running untimed until the sampler has seen it for ~0.8 s, so the clocks line describes this load
    extra = 0
    if ms < 800.0:
        extra = int((800.0 - ms) / max(ms / args.steps, 1e-3)) + 1
        sol.advance(extra, history=False)
        barrier()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["covers"] = "timed region" if extra == 0 else "timed region + %d untimed steps of the same workload" % extra
    npts = float(n) ** 3
    value = npts * stages * args.steps / (ms * 1e-3) / 1e6          # Mpts*stage/s, whole job
    launches = int(c1["kernel_launches"] - c0["kernel_launches"])

    # ---- per-kernel device time of one stage (CUDA events on the solver's stream, inside the library)
    prof = sol.profile_stage(5)
    peak, peak_kind = measured_peak()
    alg = ALG_BYTES[scheme] * npts / world                           # bytes per launch of the stage kernel on one rank
    ach = alg / (prof["rhs_stage_ms"] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "fast::stage_kernel (fused RHS + RK stage update + H,T of the new state + halo stores)", "achieved": ach, "peak": peak, "peak_kind": peak_kind,
                "unit": "GB/s", "frac": ach / peak, "traffic": ncu_traffic("rhs_stage_%d" % n),
                "alg_bytes_per_launch": alg, "kernel_ms": prof["rhs_stage_ms"], "theta_ms": prof["theta_ms"],
                "zghost_ms": prof["halo_ms"],
                # whole step (dilatation pass, reductions, hand-shake included), per GPU against one GPU's peak
                "whole_step_achieved": ALG_BYTES[scheme] * value * 1e6 / 1e9 / world, "whole_step_frac": ALG_BYTES[scheme] * value * 1e6 / 1e9 / world / peak}

    # ---- end to end through the C ABI with host buffers: copyField(0) + one step + copyField(1) per step
    e2e = None
    if not args.no_e2e:
        out = torch.empty((5, mzl, n, n), dtype=torch.float64).pin_memory()
        oviews = [out.numpy()[f] for f in range(5)]
        sol.set_state(views); sol.advance(1, history=False); sol.get_state_into(oviews)      # warm
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            sol.set_state(views)
            sol.advance(1, history=False)
            sol.get_state_into(oviews)
        barrier()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
        nbytes = 5 * 8 * int(npts)
        e2e = {"value": npts * stages * args.e2e_steps / dt / 1e6, "unit": "Mpts*RK-stage/s", "steps": args.e2e_steps,
               "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
               "what": "cudns_set_state(pinned host) + cudns_advance(1) + cudns_get_state(pinned host) per step"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        rate, steps, dt, nn = cpu_oracle_rate(scheme)
        cpu = {"value": rate, "unit": "Mpts*RK-stage/s", "cores": os.cpu_count(), "kind": "port",
               "sample": "CPU oracle (OpenMP, all host threads), TGV %d^3 same stencil/scheme, %d steps in %.1f s" % (nn, steps, dt)}
    if rank == 0:
        line = {"metric": "Mpts*RK-stage/s", "value": value, "unit": "Mpts*RK-stage/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "tgv%d_s4v4_fp64_%s" % (n, scheme), "grid": [n, n, n], "scheme": scheme,
                           "stages_per_step": stages, "stencilSize": 4, "stencilVisc": 4, "decomposition": "z-slabs x%d" % world,
                           "halo": ("peer-memory stores from the stage kernel over NVLink (CUDA IPC) + device-side epoch flags"
                                    if world > 1 else "periodic z wrap stored by the stage kernel"),
                           "cache": "inputs larger than L2 (each of the >=22 resident fields is %.2f GiB per GPU)" % (npts * 8 / world / 2 ** 30)},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
                "hbm_gbs_whole_step": ALG_BYTES[scheme] * value * 1e6 / 1e9 / world}
        print(json.dumps(line), flush=True)
    sol.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
