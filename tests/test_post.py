"""Post-processing statistics (postproc/post.cpp): Reynolds / Favre means, mean squares about them,
volume averages, friction Reynolds number over a series of saved fields.

CPU: the oracle's restatement against the outputs of the reference's OWN tool (tests/golden/ref_post_*.npz, made by
tests/golden/make_post_goldens.py from the fields the reference's GPU binary wrote).
GPU: libcudns (device reductions behind cudns_stats_* / cudns_postprocess) against the oracle, and against the same goldens."""
import os
import numpy as np
import pytest

import cudanavierstokes_b200 as cd
import oracle_binding as ob
from common import CONFIGS, apply_cfg, copy_params
from ref_cases import load_golden, GOLDEN

PRINT_TOL = 1.5e-6            # %le prints 7 significant digits


def _snapshots(name):
    g = load_golden(name)
    return [list(g["file0"]), list(g["file2"])]


def _check_against_tool(res, x, name):
    ref = np.load(os.path.join(GOLDEN, "ref_post_%s.npz" % name))
    for key in ("mean", "fluc"):
        got = np.column_stack([x] + [res[key][n] for n in range(13)])
        err = np.abs(got - ref[key]) / np.maximum(np.abs(ref[key]), 1e-300)
        # a printed 0 or denormal cannot be held to a relative error
        err = np.where(np.abs(ref[key]) < 1e-290, 0.0, err)
        assert err.max() <= PRINT_TOL, (key, np.unravel_index(err.argmax(), err.shape), err.max())
    gotb = np.concatenate([[x[0]], res["bulk"]])
    assert (np.abs(gotb - ref["bulk"][0]) <= PRINT_TOL * np.abs(ref["bulk"][0])).all()
    assert abs(res["Ret"] - float(ref["Ret"])) <= 1e-6 and abs(res["ut"] - float(ref["ut"])) <= 1e-6      # %lf: 6 decimals


@pytest.mark.parametrize("name", ["chan_s3v2", "chan_s2v2"])
def test_oracle_post_matches_the_reference_tool(name):
    op = apply_cfg(ob.params_tgv(16, 3), CONFIGS[name])
    o = ob.Oracle(op)
    res = o.post_stats(_snapshots(name))
    _check_against_tool(res, np.array(o.x), name)


def test_stats_write_reproduces_the_tools_files_byte_for_byte(tmp_path):
    """cudns_stats_write (host code, no GPU) fed with the oracle's numbers prints what the reference's tool printed"""
    name = "chan_s3v2"
    op = apply_cfg(ob.params_tgv(16, 3), CONFIGS[name])
    o = ob.Oracle(op)
    res = o.post_stats(_snapshots(name))
    cd.stats_write(tmp_path, np.array(o.x), res["mean"], res["fluc"], res["bulk"], res["Ret"], res["ut"])
    ref = np.load(os.path.join(GOLDEN, "ref_post_%s.npz" % name))
    for k in ("mean", "fluc", "bulk"):
        got = open(os.path.join(tmp_path, k + ".txt")).read().splitlines()
        want = str(ref[k + "_txt"]).splitlines()
        assert len(got) == len(want)
        assert got[1:16] == want[1:16]                         # legend and separator
        same = sum(a == b for a, b in zip(got, want))
        # the numbers agree to print precision; a last-digit flip of a %le column is possible where the 8th digit is a 5
        assert same >= len(want) - 3, (k, same, len(want))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["chan_s3v2", "chan_s2v2"])
def test_device_post_matches_oracle_and_the_reference_tool(name):
    op = apply_cfg(ob.params_tgv(16, 3), CONFIGS[name])
    o = ob.Oracle(op)
    cp = copy_params(op, cd.Params()); cp.nranks = 1
    sol = cd.Solver(cp)
    snaps = _snapshots(name)
    ref = o.post_stats(snaps)
    got = sol.post_stats(snaps)
    for key in ("mean", "fluc"):
        for n in range(13):
            scale = max(np.abs(ref[key][n]).max(), 1e-300)
            if key == "fluc":               # a variance is held to the square of the quantity's mean (cf. tools/check_diagnostics.py)
                scale = max(scale, 1e-3 * np.abs(ref["mean"][n]).max() ** 2)
            assert np.abs(got[key][n] - ref[key][n]).max() <= 1e-12 * scale, (key, n)
    assert np.abs(got["bulk"] - ref["bulk"]).max() <= 1e-12 * np.abs(ref["bulk"]).max()
    assert abs(got["Ret"] - ref["Ret"]) <= 1e-12 * ref["Ret"] and abs(got["ut"] - ref["ut"]) <= 1e-12 * ref["ut"]
    _check_against_tool(got, sol.grid["x"], name)
    sol.close()


@pytest.mark.gpu
def test_device_postprocess_reads_fields_and_writes_the_three_files(tmp_path):
    """cudns_postprocess = main() of post.cpp: fields/ on disk in, mean.txt / fluc.txt / bulk.txt out"""
    name = "chan_s3v2"
    op = apply_cfg(ob.params_tgv(16, 3), CONFIGS[name])
    cp = copy_params(op, cd.Params()); cp.nranks = 1
    sol = cd.Solver(cp)
    os.makedirs(os.path.join(tmp_path, "fields"))
    for n, st in enumerate(_snapshots(name)):
        for c, a in zip("ruvwe", st):
            cd.write_field(str(tmp_path), c, n + 1, a)
    sol.postprocess(tmp_path, 1, 2, tmp_path)
    ref = np.load(os.path.join(GOLDEN, "ref_post_%s.npz" % name))
    for k in ("mean", "fluc", "bulk"):
        got = open(os.path.join(tmp_path, k + ".txt")).read().splitlines()
        want = str(ref[k + "_txt"]).splitlines()
        assert len(got) == len(want) and got[1:16] == want[1:16]
        assert sum(a == b for a, b in zip(got, want)) >= len(want) - 3
    sol.close()


@pytest.mark.gpu
def test_driver_post_mode_is_the_reference_tool(tmp_path):
    """cudns_run post=<first>:<last> = postproc/post.cpp (argv[1], argv[2]): same three files"""
    import subprocess
    name = "chan_s2v2"
    cfg = CONFIGS[name]
    os.makedirs(os.path.join(tmp_path, "fields"))
    for n, st in enumerate(_snapshots(name)):
        for c, a in zip("ruvwe", st):
            cd.write_field(str(tmp_path), c, n + 1, a)
    keys = ("mx", "my", "mz", "stencilSize", "stencilVisc", "Lx", "Ly", "Lz", "Re", "Pr", "Ma", "viscexp", "stretch", "forcing", "periodicX", "nonUniformX")
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cudanavierstokes_b200", "cudns_run")
    r = subprocess.run([exe, "case=channel", "post=1:2", "outdir=%s" % tmp_path] + ["%s=%r" % (k, cfg[k]) for k in keys],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    ref = np.load(os.path.join(GOLDEN, "ref_post_%s.npz" % name))
    for k in ("mean", "fluc", "bulk"):
        got = open(os.path.join(tmp_path, k + ".txt")).read().splitlines()
        want = str(ref[k + "_txt"]).splitlines()
        assert len(got) == len(want) and sum(a == b for a, b in zip(got, want)) >= len(want) - 3
