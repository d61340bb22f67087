"""Quick per-kernel timing of one RK stage (cudns_profile_stage) -- development aid, not the bench."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cudanavierstokes_b200 as cd

def run(n, s, v, reps=5):
    p = cd.params_tgv(n, s, stencilVisc=v); p.nranks = 1
    g = cd.init_grid(p)
    sol = cd.Solver(p, g)
    st = cd.init_chit(p, g)
    sol.set_state(st)
    sol.advance(1, history=False)
    r = sol.profile_stage(reps)
    tot = r["theta_ms"] + r["rhs_stage_ms"] + r["halo_ms"]
    N = n ** 3
    print("n=%d s=%d v=%d theta %.3f ms rhs_stage %.3f ms zwrap %.3f ms | %.2f Gpts*stage/s  %.0f GB/s algorithmic" %
          (n, s, v, r["theta_ms"], r["rhs_stage_ms"], r["halo_ms"], N / tot / 1e6, 160.0 * N / tot / 1e6), flush=True)
    t0 = time.time(); sol.advance(5, history=False); t1 = time.time()
    print("   advance(5): %.3f ms/step wall -> %.2f Gpts*stage/s" % ((t1 - t0) / 5 * 1e3, 3 * N * 5 / (t1 - t0) / 1e9), flush=True)
    sol.close()

if __name__ == "__main__":
    for a in sys.argv[1:]:
        n, s, v = [int(t) for t in a.split(",")]
        run(n, s, v)
