"""Shared helpers: build oracle / libcudns parameter sets for the reference golden configurations."""
import os
import numpy as np
import sys
HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
sys.path.insert(0, GOLDEN)
from ref_configs import CONFIGS  # noqa: E402

_PARAM_KEYS = ("mx", "my", "mz", "stencilSize", "stencilVisc", "Lx", "Ly", "Lz", "boundaryLayer", "perturbed",
               "forcing", "periodicX", "nonUniformX", "lowStorage", "checkCFLcondition", "checkBulk",
               "Re", "Pr", "Ma", "viscexp", "stretch")


def apply_cfg(p, cfg):
    """copy a CONFIGS entry into a params struct (oracle or libcudns: same field names)"""
    for k in _PARAM_KEYS:
        setattr(p, k, cfg[k])
    p.CFL = float(np.float32(cfg["CFL"]))       # globals.h:28 is a float literal
    p.omega1 = cfg["Re"] * 121.e-6               # perturbation.h:20
    return p


def blasius_profiles():
    d = os.path.join(GOLDEN, "blasius1D")
    return [np.fromfile(os.path.join(d, "%sProf.bin" % c)) for c in "xruwe"]


def load_golden(name):
    return np.load(os.path.join(GOLDEN, "ref_%s.npz" % name))
