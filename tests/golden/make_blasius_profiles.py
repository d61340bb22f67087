"""Generate blasius1D/{x,r,u,w,e}Prof.bin by IMPORTING the reference's own similarity solver
(/root/reference/python-utils/selfSimilarSol.py) in this container, for the shipped boundary-layer
preset (src/globals.h: Ma=0.35, Pr=0.75, Re=1500, gam=1.4).  The reference imports matplotlib only
for plotting; it is absent here, so an empty stand-in module is registered first.  The outputs
(5 x 1000 float64 = 40 kB) are committed as fixtures; /root/reference is not needed at test time.

usage:  python tests/golden/make_blasius_profiles.py     (from the repo root or anywhere)
"""
import os, sys, types
here = os.path.dirname(os.path.abspath(__file__))
for name in ("matplotlib", "matplotlib.pyplot"):
    m = types.ModuleType(name); m.rc = lambda *a, **k: None; m.rcParams = {}
    sys.modules[name] = m
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.path.insert(0, "/root/reference/python-utils")
from selfSimilarSol import selfSimilarSol
os.makedirs(os.path.join(here, "blasius1D"), exist_ok=True)
os.chdir(here)
selfSimilarSol(Ma=0.35, Pr=0.75, Re=1500, gam=1.4)
print("wrote", sorted(os.listdir(os.path.join(here, "blasius1D"))))
