"""Taylor-Green vortex Re=1600 (BASELINE config 2; DeBonis, AIAA 2013-0382): kinetic-energy history K(t) = <u.u>/2 (the reference's
par1, calc_stress.cu:192-196) and dissipation -dK/dt to t = tend on an n^3 grid, 8th order FP64.
usage: tools/tgv_debonis.py [n=256] [scheme=ls3|rk4] [tend=20] [out.csv]
Published incompressible reference (spectral 512^3, van Rees et al. 2011 / DeBonis 2013): peak dissipation 0.01289 at t = 8.98."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cudanavierstokes_b200 as cd

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
scheme = sys.argv[2] if len(sys.argv) > 2 else "ls3"
tend = float(sys.argv[3]) if len(sys.argv) > 3 else 20.0
out = sys.argv[4] if len(sys.argv) > 4 else "gpurun_out/tgv%d_%s_history.csv" % (n, scheme)
p = cd.params_tgv(n, 4, Pr=0.71, lowStorage=int(scheme == "ls3"), rk4=int(scheme == "rk4"))
g = cd.init_grid(p)
s = cd.Solver(p, g)
s.set_state(cd.init_chit(p, g))
T, K = [], []
t0 = time.time(); steps = 0
chunk = 1000
while True:
    t, p1, _ = s.advance(chunk)
    steps += chunk
    idx = np.arange(0, chunk, p.checkBulk)
    T.extend(t[idx]); K.extend(0.5 * p1[idx])
    if not np.isfinite(p1[idx]).all():
        print("non-finite kinetic energy at step", steps); break
    if t[-1] >= tend:
        break
wall = time.time() - t0
T = np.array(T); K = np.array(K)
eps = -np.gradient(K, T)
ipk = int(np.argmax(eps))
os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
np.savetxt(out, np.c_[T, K, eps], header="t K=<u.u>/2 -dK/dt", fmt="%.10e")
nst = 4 if scheme == "rk4" else 3
print("tgv_debonis n=%d %s: %d steps to t=%.2f in %.1f s (%.2f Gpts*stage/s incl. diagnostics); K(0)=%.6f K(end)=%.6f; "
      "peak dissipation %.5f at t=%.2f (spectral reference 0.01289 at t=8.98)" %
      (n, scheme, steps, T[-1], wall, n ** 3 * nst * steps / wall / 1e9, K[0], K[-1], eps[ipk], T[ipk]))
