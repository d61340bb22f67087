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
(oracle/, an OpenMP restatement of the reference's algorithm, kind "port") on this box's host cores, every host
thread (OMP_NUM_THREADS is overridden: torch.distributed.run exports 1), on a bounded sample of the SAME workload:
full 512 x 512 planes of the same grid spacing, stencil and scheme, a slab of --ref-planes planes instead of 512.

The bench line also carries: `roofline` of the stage kernel (burst time of isolated launches AND the time of the same
kernel inside the step loop, CUDA events around every stage), `schemes.rk4` = the same measurement for the classical
RK4 step north_star names (200 algorithmic bytes per point-stage), `ref_gpu_baseline` = the reference's own GPU build
(oracle/_ref/perf512_*) run in the same job, `cpu_baseline`, `e2e`, `clocks`.
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


def scheme_params(cd, n, scheme, prec=0):
    p = cd.params_tgv(n, 4)
    p.lowStorage = 1 if scheme == "ls3" else 0
    p.rk4 = 1 if scheme == "rk4" else 0
    p.precision = prec                   # 0: double (the headline), 1: float (`myprec float`, globals.h:5-6)
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


def make_config(n, scheme, world, prec=0):
    """the `config` object of the bench line -- built by one function so that both arms print the same"""
    npts = float(n) ** 3 * (0.5 if prec else 1.0)
    return {"workload": "tgv%d_s4v4_%s_%s" % (n, "fp32" if prec else "fp64", scheme), "grid": [n, n, n], "scheme": scheme, "stages_per_step": STAGES[scheme],
            "stencilSize": 4, "stencilVisc": 4, "decomposition": "z-slabs x%d" % world,
            "halo": ("peer-memory stores from the stage kernel over NVLink (CUDA IPC) + device-side epoch flags"
                     if world > 1 else "periodic z wrap stored by the stage kernel"),
            "cache": "inputs larger than L2 (each of the >=22 resident fields is %.2f GiB per GPU)" % (npts * 8 / world / 2 ** 30)}


def oracle_slab(n, planes, scheme):
    """the CPU oracle on a bounded sample of the n^3 workload: full n x n planes, same grid spacing / stencil / scheme, `planes`
    planes in z (periodic).  Returns (oracle, points)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import math
    import oracle_binding as ob
    threads = ob.set_threads()                       # every core this process may run on, whatever OMP_NUM_THREADS says
    op = ob.params_tgv(n, 4, mz=planes, Lz=2 * math.pi * planes / n, lowStorage=int(scheme == "ls3"), rk4=int(scheme == "rk4"))
    o = ob.Oracle(op); o.init_chit()
    return o, n * n * planes, threads


def cpu_oracle_rate(scheme, n, planes, budget_s=12.0):
    """time the CPU oracle (all host threads) on the slab sample of the workload for about budget_s seconds"""
    o, pts, threads = oracle_slab(n, planes, scheme)
    t0 = time.perf_counter(); o.run(1); t1 = time.perf_counter() - t0
    steps = max(1, min(40, int(budget_s / max(t1, 1e-3))))
    t0 = time.perf_counter(); o.run(steps); dt = time.perf_counter() - t0
    o.close()
    return pts * STAGES[scheme] * steps / dt / 1e6, steps, dt, threads


def ref_gpu_baseline(n, scheme, root=ROOT):
    """the reference's OWN GPU build (oracle/_ref/perf<n>_n<steps>, compiled from the unmodified sources by oracle/refbuild/, sm_100
    recompile) run on this GPU: two run lengths, per-step time from the difference of the whole-run wall times it prints (set-up and
    fields/ I/O cancel).  None when the binaries are not there or the scheme is not the reference's default."""
    import glob, re, shutil, tempfile
    if scheme != "ls3":
        return {"unavailable": "the reference binaries are built for its default scheme (low-storage RK3)"}
    bins = sorted(glob.glob(os.path.join(root, "oracle", "_ref", "perf%d_n*" % n)), key=lambda b: int(b.rsplit("_n", 1)[1]))
    if len(bins) < 2:
        return {"unavailable": "oracle/_ref/perf%d_n* not built (needs /root/reference at build time)" % n}
    runs = []
    for b in (bins[0], bins[-1]):
        d = tempfile.mkdtemp(prefix="refperf.")
        os.makedirs(os.path.join(d, "fields"))
        try:
            out = subprocess.run([b], cwd=d, capture_output=True, text=True, timeout=600).stdout
        except Exception as e:       # noqa: BLE001
            shutil.rmtree(d, ignore_errors=True)
            return {"unavailable": "running %s failed: %s" % (os.path.basename(b), e)}
        shutil.rmtree(d, ignore_errors=True)
        m = re.search(r"The total time is\D*([0-9.eE+-]+)", out)
        if not m:
            return {"unavailable": "no wall time in the output of %s" % os.path.basename(b)}
        runs.append((int(b.rsplit("_n", 1)[1]), float(m.group(1))))
    (n0, t0), (n1, t1) = runs
    ms = (t1 - t0) / (n1 - n0) * 1e3
    return {"value": float(n) ** 3 * 3 / (ms * 1e-3) / 1e6, "unit": "Mpts*RK-stage/s", "ms_per_step": ms,
            "how": "oracle/_ref/%s and %s (unmodified reference kernels, sm_100): (%.2f s - %.2f s) / (%d - %d steps)" %
                   (os.path.basename(bins[-1]), os.path.basename(bins[0]), t1, t0, n1, n0)}


def run_reference(args):
    """--impl reference: the oracle port on the host cores; each "step" = one time step of the bounded sample (a slab of full
    n x n planes of the same workload)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n, scheme = args.n, args.scheme
    o, pts, threads = oracle_slab(n, args.ref_planes, scheme)
    for _ in range(args.warmup):
        o.run(1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.run(1)
    dt = time.perf_counter() - t0
    val = pts * STAGES[scheme] * args.steps / dt / 1e6
    sample = ("CPU oracle (OpenMP restatement of the reference, %d host threads) on a %d x %d x %d slab of the %d^3 workload "
              "(same grid spacing, stencil, scheme; periodic in z), 1 time step per bench step" % (threads, n, n, args.ref_planes, n))
    line = {"impl": "reference", "metric": "Mpts*RK-stage/s", "value": val, "unit": "Mpts*RK-stage/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": make_config(n, scheme, args.gpus), "sample_grid": [n, n, args.ref_planes],
            "cpu_baseline": {"value": val, "unit": "Mpts*RK-stage/s", "cores": threads, "kind": "port", "sample": sample},
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
    ap.add_argument("--ref-planes", type=int, default=32, help="planes of the CPU sample slab (full n x n planes each)")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the run of the reference's own GPU build (about 2 minutes)")
    ap.add_argument("--no-schemes", action="store_true", help="skip the additional RK4 measurement")
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

    n = args.n
    npts = float(n) ** 3
    mzl = n // world
    peak, peak_kind = measured_peak()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # Taylor-Green initial condition of this rank's slab (the formulas of initCHIT, init.cpp:126-148, vectorised), pinned host memory
    p0 = scheme_params(cd, n, args.scheme)
    p0.nranks = world; p0.rank = rank; p0.device = local
    grid = cd.init_grid(p0)
    pin = torch.empty((5, mzl, n, n), dtype=torch.float64).pin_memory()
    host = pin.numpy()
    tgv_slab(host, grid, p0, rank * mzl, mzl)
    views = [host[f] for f in range(5)]

    def measure(scheme, steps, warmup, with_clocks, prec=0):
        """one solver of `scheme`: K timed steps between CUDA events on the solver's stream (max over ranks), then the same K steps
        again with CUDA events around every stage (kernel times under the sustained load of the step loop), then isolated launches"""
        stages = STAGES[scheme]
        p = scheme_params(cd, n, scheme, prec)
        p.nranks = world; p.rank = rank; p.device = local
        sol = cd.Solver(p, grid)
        if world > 1:
            from cudanavierstokes_b200 import dist as cdist
            cdist.attach(sol)
        sol.set_state(views)
        stream = torch.cuda.ExternalStream(sol.stream(), device=torch.device("cuda", local))
        sol.advance(warmup, history=False)
        c0 = sol.counters()
        ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(local) if with_clocks else None
        barrier()
        if sampler is not None and rank == 0:
            sampler.start()
        ev0.record(stream)
        sol.advance(steps, history=False)
        ev1.record(stream)
        barrier()
        ms = max_over_ranks(ev0.elapsed_time(ev1))
        c1 = sol.counters()
        # the same steps once more, back to back, with events around the dilatation pass / stage kernel / hand-shake of every stage:
        # the kernels' durations under the conditions of the timed region (clocks under the power cap); nvidia-smi also needs
        # ~0.8 s of this load to deliver samples when the timed region was short (every rank derives the same count from `ms`)
        extra = max(steps, int(800.0 / max(ms / steps, 1e-3)) + 1 if ms < 800.0 else steps)
        sol.stage_timing(True)
        done = 0
        while done < extra:
            m = min(extra - done, 1024 // stages)
            sol.advance(m, history=False); done += m
        sust = sol.stage_timing()
        sol.stage_timing(False)
        barrier()
        clocks = sampler.stop() if (sampler is not None and rank == 0) else None
        if clocks is not None:
            clocks["covers"] = "timed region + %d steps of the same workload with per-stage events" % extra
        prof = sol.profile_stage(5)
        value = npts * stages * steps / (ms * 1e-3) / 1e6          # Mpts*stage/s, whole job
        bpp = ALG_BYTES[scheme] * (0.5 if prec else 1.0)           # algorithmic bytes per point-stage (80 B in FP32)
        alg = bpp * npts / world                                    # algorithmic bytes per launch of the stage kernel on one rank
        k_sust = max_over_ranks(sust["rhs_stage_ms"]); k_burst = max_over_ranks(prof["rhs_stage_ms"])
        kernel = ("duo::stage_kernel<4,4> (two points per thread)" if (scheme != "ls3" or prec or os.environ.get("CUDNS_DUO") == "1")
                  else "fast::stage_kernel<4,4,8>") + ": fused RHS + RK stage update + H,T of the new state + halo stores"
        roofline = {"bound": "hbm", "kernel": kernel, "achieved": alg / (k_sust * 1e-3) / 1e9, "peak": peak, "peak_kind": peak_kind,
                    "unit": "GB/s", "frac": alg / (k_sust * 1e-3) / 1e9 / peak,
                    "traffic": ncu_traffic("rhs_stage_%d_%s%s" % (n, scheme, "_f32" if prec else "")) if world == 1 else None,
                    "alg_bytes_per_launch": alg, "alg_bytes_per_point": bpp,
                    "kernel_ms": k_sust, "kernel_ms_how": "mean over %d launches inside the step loop (CUDA events around every stage, sustained clocks)" % sust["stages"],
                    "kernel_ms_burst": k_burst, "frac_burst": alg / (k_burst * 1e-3) / 1e9 / peak,
                    "kernel_ms_burst_how": "5 isolated launches of the scheme's most frequent stage shape (cudns_profile_stage)",
                    "theta_ms": max_over_ranks(sust["theta_ms"]), "theta_ms_burst": prof["theta_ms"], "handshake_ms": max_over_ranks(sust["halo_ms"]),
                    # whole step (dilatation pass, reductions, hand-shake included), per GPU against one GPU's peak
                    "whole_step_achieved": bpp * value * 1e6 / 1e9 / world,
                    "whole_step_frac": bpp * value * 1e6 / 1e9 / world / peak}
        return sol, {"value": value, "ms_per_step": ms / steps, "launches": int(c1["kernel_launches"] - c0["kernel_launches"]),
                     "roofline": roofline, "clocks": clocks}

    scheme = args.scheme
    stages = STAGES[scheme]
    sol, main_res = measure(scheme, args.steps, args.warmup, True)

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
        dt = max_over_ranks(time.perf_counter() - t0)
        nbytes = 5 * 8 * int(npts)
        e2e = {"value": npts * stages * args.e2e_steps / dt / 1e6, "unit": "Mpts*RK-stage/s", "steps": args.e2e_steps,
               "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
               "what": "cudns_set_state(pinned host) + cudns_advance(1) + cudns_get_state(pinned host) per step"}
        del out, oviews
    sol.close()

    # ---- the classical RK4 step north_star names (an extension: the reference has Wray and Kutta RK3 only), same grid and stencil
    schemes = None
    if not args.no_schemes and scheme != "rk4":
        s4, r4 = measure("rk4", args.steps, args.warmup, False)
        s4.close()
        schemes = {"rk4": {"config": make_config(n, "rk4", world), "value": r4["value"], "unit": "Mpts*RK-stage/s", "ms_per_step": r4["ms_per_step"],
                           "stages_per_step": 4, "roofline": r4["roofline"], "gpu_launches": r4["launches"], "dtype": "f64"}}
        # ---- single precision (`myprec float` of the reference, globals.h:5-6): the same step with the float copy of the device side,
        # 80 algorithmic bytes per point-stage.  Not the headline (BASELINE names FP64); reported because it is the configuration in which
        # the path is closest to its HBM roofline
        s32, r32 = measure(scheme, args.steps, args.warmup, False, prec=1)
        s32.close()
        schemes["%s_fp32" % scheme] = {"config": make_config(n, scheme, world, 1), "value": r32["value"], "unit": "Mpts*RK-stage/s",
                                       "ms_per_step": r32["ms_per_step"], "stages_per_step": stages, "roofline": r32["roofline"],
                                       "gpu_launches": r32["launches"], "dtype": "f32"}

    cpu = None
    refgpu = None
    if rank == 0 and world == 1:
        if not args.no_cpu:
            rate, steps, dt, threads = cpu_oracle_rate(scheme, n, args.ref_planes)
            cpu = {"value": rate, "unit": "Mpts*RK-stage/s", "cores": threads, "kind": "port",
                   "sample": "CPU oracle (OpenMP, %d host threads) on a %d x %d x %d slab of the workload, same stencil/scheme, %d steps in %.1f s" %
                             (threads, n, n, args.ref_planes, steps, dt)}
        if not args.no_ref_gpu:
            torch.cuda.empty_cache()
            refgpu = ref_gpu_baseline(n, scheme)
    if rank == 0:
        line = {"metric": "Mpts*RK-stage/s", "value": main_res["value"], "unit": "Mpts*RK-stage/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": make_config(n, scheme, world),
                "roofline": main_res["roofline"], "cpu_baseline": cpu, "ref_gpu_baseline": refgpu, "e2e": e2e,
                "gpu_launches": main_res["launches"], "clocks": main_res["clocks"], "schemes": schemes,
                "hbm_gbs_whole_step": ALG_BYTES[scheme] * main_res["value"] * 1e6 / 1e9 / world}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
