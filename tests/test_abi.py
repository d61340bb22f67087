"""CPU-side checks of the C ABI boundary: the library loads, exports every symbol include/cudns.h declares,
its host-side helpers (grid, initial conditions, sponge tables, fields/ I/O) agree with the oracle, and it refuses
to run without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

import cudanavierstokes_b200 as cd
import oracle_binding as ob
from common import CONFIGS, apply_cfg, blasius_profiles, copy_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _built():
    cd.build()


def test_header_symbols_all_exported():
    hdr = open(os.path.join(ROOT, "include", "cudns.h")).read()
    declared = set(re.findall(r"\b(cudns_[a-z_0-9]+)\s*\(", hdr)) - {"cudns_allreduce_fn", "cudns_exchange_fn"}
    assert declared == set(cd.EXPORTS), declared ^ set(cd.EXPORTS)
    L = cd.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in L.cudns_version()


def test_params_struct_layout_matches_oracle_prefix():
    """the first 42 fields of cudns_params are the reference's knobs in the same order as the oracle's struct"""
    a = [n for n, *_ in ob.OraParams._fields_]
    b = [n for n, *_ in cd.Params._fields_][:len(a)]
    assert a == b


@pytest.mark.parametrize("case", ["tgv", "channel", "blayer"])
def test_grid_and_initial_conditions_match_oracle(case):
    if case == "tgv":
        op = ob.params_tgv(16, 3)
    elif case == "channel":
        op = apply_cfg(ob.params_tgv(16, 3), CONFIGS["chan_s3v2"])
    else:
        op = apply_cfg(ob.params_tgv(16, 3), dict(CONFIGS["bl_s3v2"], mz=48))
    cp = copy_params(op, cd.Params()); cp.nranks = 1
    o = ob.Oracle(op); g = cd.init_grid(cp)
    for k in ("x", "xp", "xpp", "y", "z"):
        assert np.array_equal(g[k], getattr(o, k)), k
    assert g["dx"] == o.dx
    if case == "tgv":
        o.init_chit(); mine = cd.init_chit(cp, g)
    elif case == "channel":
        o.init_channel(); mine = cd.init_channel(cp, g)
    else:
        x, r, u, w, e = blasius_profiles()
        o.set_sponge_from_profiles(x, r, u, w, e)
        sx, sz, ref, mine = cd.build_sponge(cp, g, x, r, u, w)
        assert np.array_equal(sx, o.spongeX) and np.array_equal(sz, o.spongeZ)
        for q in range(5):
            assert np.array_equal(ref[q], o.ref(q))
    for a, b in zip(mine, o.state()):
        assert np.array_equal(a, b)


def test_fields_io_format(tmp_path):
    """fields/<c>.<%07d>.bin: raw little-endian float64 [mz][my][mx], no header (comm.cpp:218-250)"""
    os.makedirs(tmp_path / "fields")
    a = np.arange(2 * 3 * 4, dtype=np.float64).reshape(2, 3, 4) * 0.5
    cd.write_field(str(tmp_path), "r", 12, a)
    raw = np.fromfile(tmp_path / "fields" / "r.0000012.bin", dtype="<f8")
    assert np.array_equal(raw, a.ravel())
    assert np.array_equal(cd.read_field(str(tmp_path), "r", 12, a.shape), a)
    with pytest.raises(cd.CudnsError):
        cd.read_field(str(tmp_path), "u", 12, a.shape)


def test_parameter_validation():
    for over in (dict(stencilVisc=4, stencilSize=3), dict(stencilSize=5), dict(mx=31), dict(nranks=3), dict(Re=-1.0)):
        p = cd.params_tgv(32, 3, **over)
        with pytest.raises(cd.CudnsError):
            cd.init_grid(p)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(cd.CudnsError, match="no CUDA device|CUDA"):
        cd.Solver(cd.params_tgv(16, 2))
