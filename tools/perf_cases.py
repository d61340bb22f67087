"""Throughput of the non-periodic BASELINE configurations at their full sizes (the lean GEN stage kernel):
C3 supersonic isothermal channel 160x192x192 (4th order: s=v=2, and the preset's s=3,v=2), C4 boundary layer 240x64x2048
with sponges and wall blowing/suction (s=3,v=2).  usage: tools/perf_cases.py [steps] [f32|f64] [name filter]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cudanavierstokes_b200 as cd
from ref_cases import CONFIGS, apply_cfg, blasius_profiles

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
prec = int(len(sys.argv) > 2 and sys.argv[2] == "f32")
only = sys.argv[3] if len(sys.argv) > 3 else ""

def run(name, cfg, over):
    if only not in name:
        return
    p = apply_cfg(cd.Params(), dict(cfg, **over)); p.gam = 1.4; p.TwallTop = p.TwallBot = 1.0; p.quirk_q1 = 1; p.nranks = 1; p.precision = prec
    ref = cd.params_blayer() if cfg["case"] == "blayer" else cd.params_channel()
    for k in ("spTopStr", "spTopLen", "spTopExp", "spInlStr", "spInlLen", "spInlExp", "spOutStr", "spOutLen", "spOutExp",
              "kC", "LP", "amp1", "amp2", "omega2"):
        setattr(p, k, getattr(ref, k))
    g = cd.init_grid(p)
    s = cd.Solver(p, g)
    if cfg["case"] == "blayer":
        x, r, u, w, e = blasius_profiles()
        sx, sz, rf, ic = cd.build_sponge(p, g, x[1:], r[1:], u[1:], w[1:])
        s.set_sponge(sx, sz, rf); s.set_state(ic)
    else:
        s.set_state(cd.init_channel(p, g))
    s.advance(3, history=False)
    prof = s.profile_stage(3)
    t0 = time.time(); s.advance(steps, history=False); dt = time.time() - t0
    st = s.get_state()
    N = p.mx * p.my * p.mz
    ok = all(np.isfinite(a).all() for a in st)
    print("perf_case%s %-28s %4dx%4dx%4d s=%d v=%d: %.3f ms/step -> %.2f Gpts*stage/s | theta %.3f ms stage %.3f ms | finite=%s" %
          (" f32" if prec else "", name, p.mx, p.my, p.mz, p.stencilSize, p.stencilVisc, dt / steps * 1e3, 3 * N * steps / dt / 1e9,
           prof["theta_ms"], prof["rhs_stage_ms"], ok), flush=True)
    s.close()

run("channel M=1.5 (4th order)", CONFIGS["chan_s2v2"], dict(mx=160, my=192, mz=192, checkCFLcondition=100, checkBulk=100))
run("channel M=1.5 (preset)", CONFIGS["chan_s3v2"], dict(mx=160, my=192, mz=192, checkCFLcondition=100, checkBulk=100))
run("boundary layer + sponges", CONFIGS["bl_s3v2"], dict(mx=240, my=64, mz=2048, checkCFLcondition=100, checkBulk=100))
