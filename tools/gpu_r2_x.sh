#!/bin/bash
mkdir -p gpurun_out
{
for n in 2 4 6 8 12 16 24; do echo "== channel/BL CUDNS_ZCHUNKS=$n"; CUDNS_ZCHUNKS=$n timeout 300 python tools/perf_cases.py 10 2>&1 | cut -c1-150; done
for n in 32 64; do echo "== BL CUDNS_ZCHUNKS=$n"; CUDNS_ZCHUNKS=$n timeout 300 python tools/perf_cases.py 10 2>&1 | grep boundary | cut -c1-150; done
for n in 1 2 4 8 16; do echo "== tgv512 CUDNS_ZCHUNKS=$n"; CUDNS_ZCHUNKS=$n timeout 300 python tools/quick_perf.py 512,4,4 512,4,4,ls3,f32 2>&1 | grep -v advance
  CUDNS_DUO=1 CUDNS_ZCHUNKS=$n timeout 300 python tools/quick_perf.py 512,4,4,rk4 2>&1 | grep -v advance; done
} | tee gpurun_out/r2x_zchunk_sweep.log
