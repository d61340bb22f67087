"""CPU-checkable pieces of bench.py: the slab initial condition equals initCHIT bit for bit, and the reference arm prints one
JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

import numpy as np

import cudanavierstokes_b200 as cd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_tgv_slab_equals_init_chit():
    import bench
    p = cd.params_tgv(24, 4); g = cd.init_grid(p)
    ref = cd.init_chit(p, g)
    for world in (1, 2, 3):
        mzl = 24 // world
        for rank in range(world):
            out = np.zeros((5, mzl, 24, 24))
            bench.tgv_slab(out, g, p, rank * mzl, mzl)
            for f in range(5):
                assert np.array_equal(out[f], ref[f][rank * mzl:(rank + 1) * mzl])


def test_reference_arm_prints_the_contract_line():
    """a small grid so that the CPU suite stays short: 64 x 64 x 16 slab of a 64^3 workload; OMP_NUM_THREADS=1 in the environment (what
    torch.distributed.run exports) must not leave the arm single-threaded"""
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--grid", "64",
                        "--ref-planes", "16"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "Mpts*RK-stage/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    ncpu = len(os.sched_getaffinity(0))
    assert line["cpu_baseline"]["cores"] == ncpu                  # every core, although OMP_NUM_THREADS said 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    # the config object is the GPU arm's (same workload label and grid); the sample that was actually timed is stated next to it
    import bench
    assert line["config"] == bench.make_config(64, "ls3", 1)
    assert line["sample_grid"] == [64, 64, 16] and "64 x 64 x 16 slab" in line["cpu_baseline"]["sample"]
    # under torchrun only rank 0 works
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--grid", "64",
                        "--ref-planes", "16"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
