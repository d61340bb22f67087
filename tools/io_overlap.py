"""Does output stall the step loop?  Time K steps alone, K steps with an asynchronous snapshot in flight, and the blocking
alternative (cudns_get_state + cudns_write_field).  usage: tools/io_overlap.py [n] [steps] [dir]"""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import cudanavierstokes_b200 as cd

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
d = sys.argv[3] if len(sys.argv) > 3 else tempfile.mkdtemp()
p = cd.params_tgv(n, 4); g = cd.init_grid(p)
s = cd.Solver(p, g); s.set_state(cd.init_chit(p, g)); s.advance(3, history=False)

def timed(f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3

s.write_fields_async(d, 0); s.io_wait()        # first call allocates the pinned staging buffer (one-off, ~0.4 ms per MB)
alone = timed(lambda: s.advance(K, history=False))
def with_async():
    s.write_fields_async(d, 1); s.advance(K, history=False)
loop_async = timed(with_async)
drain = timed(lambda: s.io_wait())
def blocking():
    st = s.get_state()
    os.makedirs(os.path.join(d, "fields"), exist_ok=True)
    for c, a in zip("ruvwe", st): cd.write_field(d, c, 2, a)
    s.advance(K, history=False)
blk = timed(blocking)
print("io_overlap n=%d K=%d: steps alone %.1f ms | with async snapshot in flight %.1f ms (+%.1f %%), writer drained %.1f ms later | blocking "
      "get_state + write_field + steps %.1f ms" % (n, K, alone, loop_async, 100 * (loop_async / alone - 1), drain, blk))
