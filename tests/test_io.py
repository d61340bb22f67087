"""fields/ output, restart and the XDMF sidecar: host-side pieces without a GPU, the asynchronous
writer against the solver's own state on a GPU."""
import os
import sys

import numpy as np
import pytest

import cudanavierstokes_b200 as cd

HERE = os.path.dirname(os.path.abspath(__file__))


def _golden_inputs():
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from xdmf_inputs import inputs
    return inputs()


def test_xdmf_sidecar_matches_the_reference_writer(tmp_path):
    """byte for byte the file python-utils/writexmf.py writes for the same inputs (tests/golden/xdmf_ref.xmf)"""
    x, y, z, ts, dt, names = _golden_inputs()
    out = tmp_path / "fields.xmf"
    cd.write_xdmf(out, x, y, z, ts, dt, "".join(names))
    assert out.read_bytes() == open(os.path.join(HERE, "golden", "xdmf_ref.xmf"), "rb").read()


def test_field_file_format_roundtrip(tmp_path):
    """fields/<c>.<%07d>.bin: raw float64 [mz][my][mx], no header (init.cpp:13-30, comm.cpp:205-279)"""
    os.makedirs(tmp_path / "fields")
    a = np.arange(4 * 3 * 5, dtype=np.float64).reshape(4, 3, 5) * 0.5
    cd.write_field(str(tmp_path), "r", 42, a)
    raw = np.fromfile(tmp_path / "fields" / "r.0000042.bin", dtype=np.float64)
    assert raw.size == a.size and np.array_equal(raw.reshape(a.shape), a)
    assert np.array_equal(cd.read_field(str(tmp_path), "r", 42, a.shape), a)


@pytest.mark.gpu
def test_async_writer_snapshots_while_the_step_loop_runs(tmp_path):
    """the snapshot is stream-ordered with cudns_advance: files hold the state of the moment of the call although the step
    loop keeps running; a restart from them reproduces the continued run bit for bit"""
    p = cd.params_tgv(32, 3)
    g = cd.init_grid(p)
    s = cd.Solver(p, g)
    s.set_state(cd.init_chit(p, g))
    s.advance(3)
    snap = s.get_state()
    s.write_fields_async(tmp_path, 3)
    s.advance(4)                                   # overlaps the D2H copy and the file writes
    s.write_fields_async(tmp_path, 7)              # waits for snapshot 3 to be on disk, then snapshots step 7
    assert s.io_wait() == 10
    end = s.get_state()
    for c, a in zip("ruvwe", snap):
        got = np.fromfile(tmp_path / "fields" / ("%s.0000003.bin" % c), dtype=np.float64).reshape(a.shape)
        assert np.array_equal(got, a)
    for c, a in zip("ruvwe", end):
        got = np.fromfile(tmp_path / "fields" / ("%s.0000007.bin" % c), dtype=np.float64).reshape(a.shape)
        assert np.array_equal(got, a)
    # restart (initField + copyField(0)) from step 3 and repeat the 4 steps
    r = cd.Solver(p, g)
    r.read_fields(tmp_path, 3)
    for a, b in zip(r.get_state(), snap):
        assert np.array_equal(a, b)
    with pytest.raises(cd.CudnsError):
        r.read_fields(tmp_path, 99)


@pytest.mark.gpu
def test_async_writer_two_slabs_share_one_file(tmp_path):
    """two ranks (two handles in this process) write their z slabs into the same global files at their byte offsets"""
    n = 32
    full = cd.params_tgv(n, 2); gfull = cd.init_grid(full)
    st = cd.init_chit(full, gfull)
    sols = []
    for rk in range(2):
        p = cd.params_tgv(n, 2); p.nranks = 2; p.rank = rk
        s = cd.Solver(p, cd.init_grid(p))
        sols.append(s)
    # fill each rank's slab through the device path that needs no ghost exchange: write the files, then compare
    mzl = n // 2
    for rk, s in enumerate(sols):
        s.set_exchange(lambda stream: None)        # ghosts are irrelevant for an I/O test ...
        s.set_allreduce(lambda ptr, n, op: None)   # ... and so are the cross-slab scalars (the library refuses to run without the callback)
        s.set_state([a[rk * mzl:(rk + 1) * mzl] for a in st])
        s.write_fields_async(tmp_path, 5)
    for s in sols:
        s.io_wait()
    for c, a in zip("ruvwe", st):
        got = np.fromfile(tmp_path / "fields" / ("%s.0000005.bin" % c), dtype=np.float64).reshape(a.shape)
        assert np.array_equal(got, a)
