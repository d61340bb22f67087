"""On-device wall-normal profiles / friction Reynolds number against the oracle (run as its own process by
tests/test_zz_diagnostics.py so that a fault cannot touch the test session).  exit code 0 = all checks passed."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_binding as ob
from common import CONFIGS, apply_cfg, make_pair, smooth_random_state


def close(got, ref, tol=1e-11):
    return all(np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1e-30) + 1e-28 for a, b in zip(got, ref))


ok = True
for name in ("chan_s3v2", "chan_s2v2"):
    op = apply_cfg(ob.params_tgv(16, 3), CONFIGS[name])
    o, s, grid = make_pair(op)
    o.init_channel(); s.set_state(o.state())
    o.run(3); s.advance(3)
    a = close(s.profiles(), o.profiles()); b = abs(s.retau() - o.retau()) <= 1e-11 * o.retau()
    print("check_diagnostics %s: profiles %s, Re_tau %s (%.6f vs %.6f)" % (name, a, b, s.retau(), o.retau()))
    ok = ok and a and b
op = ob.params_tgv(24, 3, mx=40, my=20, mz=24)
o, s, grid = make_pair(op)
st = smooth_random_state(o); o.set_state(st); s.set_state(st)
a = close(s.profiles(), o.profiles())
print("check_diagnostics ragged periodic box: profiles %s" % a)
ok = ok and a
sys.exit(0 if ok else 1)
