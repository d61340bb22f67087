"""Golden outputs of the reference's own post-processing tool (postproc/post.cpp, built by oracle/refbuild/build_post_all.sh into
oracle/_ref/post_<cfg>; a host-only program, it runs in the build container).

Input: the fields the reference's GPU binary wrote for the golden configurations (file0 = initial state and file2 = 20 steps of
tests/golden/ref_<cfg>.npz, both stored at full resolution), written as fields/<c>.0000001.bin and .0000002.bin.  Output: the
numbers of mean.txt, fluc.txt and bulk.txt (14 columns in %le: 7 significant digits) and the header's Re_tau / u_tau (%lf: 6
decimals) and the raw text of the three files (format golden of cudns_stats_write) -> tests/golden/ref_post_<cfg>.npz.

usage: bash oracle/refbuild/build_post_all.sh && python tests/golden/make_post_goldens.py
"""
import os, re, subprocess, sys, tempfile
import numpy as np

here = os.path.dirname(os.path.abspath(__file__))
root = os.path.dirname(os.path.dirname(here))


def parse(path):
    lines = open(path).read().splitlines()
    m = re.match(r"Reynolds number based on utau (\S+) with utau (\S+)", lines[0])
    sep = max(i for i, l in enumerate(lines) if l.startswith("-----"))
    rows = np.array([[float(t) for t in l.split()] for l in lines[sep + 1:] if l.strip()])
    return float(m.group(1)), float(m.group(2)), rows


for name in ("chan_s3v2", "chan_s2v2"):       # the tool builds the channel grid (two-sided tanh, postproc/init.cpp:34-70): channel cases only
    g = np.load(os.path.join(here, "ref_%s.npz" % name))
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "fields")); os.makedirs(os.path.join(d, "run"))
        for n, key in ((1, "file0"), (2, "file2")):
            for c, a in zip("ruvwe", g[key]):
                np.ascontiguousarray(a, dtype=np.float64).tofile(os.path.join(d, "fields", "%s.%07d.bin" % (c, n)))
        subprocess.run([os.path.join(root, "oracle", "_ref", "post_" + name), "1", "2"], cwd=os.path.join(d, "run"), check=True,
                       stdout=subprocess.DEVNULL)
        ret, ut, mean = parse(os.path.join(d, "run", "mean.txt"))
        _, _, fluc = parse(os.path.join(d, "run", "fluc.txt"))
        _, _, bulk = parse(os.path.join(d, "run", "bulk.txt"))
        texts = {k: open(os.path.join(d, "run", k + ".txt")).read() for k in ("mean", "fluc", "bulk")}
    np.savez_compressed(os.path.join(here, "ref_post_%s.npz" % name), mean=mean, fluc=fluc, bulk=bulk, Ret=ret, ut=ut,
                        mean_txt=texts["mean"], fluc_txt=texts["fluc"], bulk_txt=texts["bulk"])
    print(name, mean.shape, fluc.shape, bulk.shape, ret, ut)
