"""inputs of the XDMF golden (shared by make_xdmf_golden.py and tests/test_io.py)"""
import numpy as np


def inputs():
    x = np.linspace(0.0, 2.0, 7) ** 1.5
    y = 2 * np.pi * (np.arange(5) + 0.5) / 5
    z = 4 * np.pi * (np.arange(6) + 0.5) / 6
    return x, y, z, np.array([0, 100, 2500]), 1.25e-3, ["r", "u", "v", "w", "e"]
