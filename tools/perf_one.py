"""One wall-bounded BASELINE configuration at full size (for ncu): tools/perf_one.py chan2|chan3|bl [f32]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cudanavierstokes_b200 as cd
from ref_cases import CONFIGS, apply_cfg, blasius_profiles

which = sys.argv[1]; prec = int(len(sys.argv) > 2 and sys.argv[2] == "f32")
cfg, over = {"chan2": (CONFIGS["chan_s2v2"], dict(mx=160, my=192, mz=192)), "chan3": (CONFIGS["chan_s3v2"], dict(mx=160, my=192, mz=192)),
             "bl": (CONFIGS["bl_s3v2"], dict(mx=240, my=64, mz=2048))}[which]
p = apply_cfg(cd.Params(), dict(cfg, checkCFLcondition=100, checkBulk=100, **over)); p.gam = 1.4; p.TwallTop = p.TwallBot = 1.0; p.quirk_q1 = 1; p.nranks = 1
p.precision = prec
ref = cd.params_blayer() if cfg["case"] == "blayer" else cd.params_channel()
for k in ("spTopStr", "spTopLen", "spTopExp", "spInlStr", "spInlLen", "spInlExp", "spOutStr", "spOutLen", "spOutExp", "kC", "LP", "amp1", "amp2", "omega2"):
    setattr(p, k, getattr(ref, k))
g = cd.init_grid(p)
s = cd.Solver(p, g)
if cfg["case"] == "blayer":
    x, r, u, w, e = blasius_profiles()
    sx, sz, rf, ic = cd.build_sponge(p, g, x[1:], r[1:], u[1:], w[1:])
    s.set_sponge(sx, sz, rf); s.set_state(ic)
else:
    s.set_state(cd.init_channel(p, g))
s.advance(4, history=False)
s.close()
