"""Convert the raw outputs of the reference's own GPU binary (gpurun_out/ref/<cfg>/, produced on a
B200 by oracle/refbuild/run_goldens.sh) into compact fixtures tests/golden/ref_<cfg>.npz.

Each fixture holds: the state after file 0 (initial), 1 and 2 (10 and 20 steps) as float64
[5][mz][my][mx] (r,u,v,w,e), the solution.txt rows and the (dt, dpdz) pairs the reference printed at
every CFL refresh.  To keep the repository small, file 1 is stored on a stride-2 subsample.

usage: python tests/golden/make_ref_goldens.py [gpurun_out/ref]
"""
import os, re, sys
import numpy as np

here = os.path.dirname(os.path.abspath(__file__))
root = os.path.dirname(os.path.dirname(here))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "gpurun_out", "ref")
sys.path.insert(0, here)
from ref_configs import CONFIGS

for name, cfg in CONFIGS.items():
    d = os.path.join(src, name)
    if not os.path.isdir(d):
        print("skip", name); continue
    shape = (cfg["mz"], cfg["my"], cfg["mx"])
    out = {}
    for f in (0, 1, 2):
        st = np.stack([np.fromfile(os.path.join(d, "fields", "%s.%07d.bin" % (c, f))).reshape(shape) for c in "ruvwe"])
        out["file%d" % f] = st[:, ::2, ::2, ::2].copy() if f == 1 else st
    sol = np.loadtxt(os.path.join(d, "solution.txt"), ndmin=2)
    out["solution"] = sol
    log = open(os.path.join(d, "stdout.txt")).read()
    out["dt_dpdz"] = np.array([[float(a), float(b)] for a, b in re.findall(r"step number \d+ with (\S+) (\S+)", log)])
    out["x"] = np.fromfile(os.path.join(d, "fields", "x.bin"))
    np.savez_compressed(os.path.join(here, "ref_%s.npz" % name), **out)
    print(name, os.path.getsize(os.path.join(here, "ref_%s.npz" % name)) // 1024, "kB")
