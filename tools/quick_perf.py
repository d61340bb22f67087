"""Quick per-kernel timing of one RK stage (cudns_profile_stage) -- development aid, not the bench."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cudanavierstokes_b200 as cd

def run(n, s, v, scheme="ls3", reps=int(os.environ.get("QP_REPS", "5")), prec=0):
    p = cd.params_tgv(n, s, stencilVisc=v, lowStorage=int(scheme == "ls3"), rk4=int(scheme == "rk4")); p.nranks = 1; p.precision = prec
    g = cd.init_grid(p)
    sol = cd.Solver(p, g)
    st = cd.init_chit(p, g)
    sol.set_state(st)
    sol.advance(1, history=False)
    r = sol.profile_stage(reps)
    tot = r["theta_ms"] + r["rhs_stage_ms"] + r["halo_ms"]
    N = n ** 3
    stages = 4 if scheme == "rk4" else 3
    print("%s%s n=%d s=%d v=%d theta %.3f ms rhs_stage %.3f ms zwrap %.3f ms | %.2f Gpts*stage/s  %.0f GB/s algorithmic" %
          (scheme, " f32" if prec else "", n, s, v, r["theta_ms"], r["rhs_stage_ms"], r["halo_ms"], N / tot / 1e6, (80.0 if prec else 160.0) * N / tot / 1e6), flush=True)
    t0 = time.time(); sol.advance(5, history=False); t1 = time.time()
    print("   advance(5): %.3f ms/step wall -> %.2f Gpts*stage/s" % ((t1 - t0) / 5 * 1e3, stages * N * 5 / (t1 - t0) / 1e9), flush=True)
    sol.close()

if __name__ == "__main__":
    for a in sys.argv[1:]:
        t = a.split(",")
        run(int(t[0]), int(t[1]), int(t[2]), t[3] if len(t) > 3 else "ls3", prec=int(len(t) > 4 and t[4] == "f32"))
