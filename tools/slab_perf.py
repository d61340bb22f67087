"""Per-kernel times of one rank's slab of a z-decomposed run on ONE GPU (periodic wrap instead of neighbours): tools/slab_perf.py mx my mz [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudanavierstokes_b200 as cd
mx, my, mz = (int(a) for a in sys.argv[1:4]); reps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
p = cd.params_tgv(mx, 4, stencilVisc=4, mx=mx, my=my, mz=mz); p.nranks = 1
g = cd.init_grid(p); s = cd.Solver(p, g); s.set_state(cd.init_chit(p, g)); s.advance(2, history=False)
r = s.profile_stage(reps)
print("slab %dx%dx%d: theta %.4f ms stage %.4f ms  (CUDNS_THETA_ZCHUNKS=%s CUDNS_ZCHUNKS=%s)" % (mx, my, mz, r["theta_ms"], r["rhs_stage_ms"], os.environ.get("CUDNS_THETA_ZCHUNKS"), os.environ.get("CUDNS_ZCHUNKS")))
s.close()
