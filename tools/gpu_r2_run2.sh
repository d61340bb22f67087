#!/bin/bash
# 2-GPU session: the C++ host driver with two slabs in one process (no Python, no NCCL)
mkdir -p gpurun_out
{
for g in 1 2; do
  rm -rf /tmp/run$g; timeout 600 cudanavierstokes_b200/cudns_run case=tgv mx=512 my=512 mz=512 stencilSize=4 nsteps=30 nfiles=1 ngpus=$g outdir=/tmp/run$g xdmf=0 2>&1 | grep -E "cudns_run|total time|per time step|file number"
done
cmp /tmp/run1/fields/r.0000001.bin /tmp/run2/fields/r.0000001.bin && cmp /tmp/run1/fields/e.0000001.bin /tmp/run2/fields/e.0000001.bin && echo "fields of the 1-GPU and the 2-GPU run are byte-identical"
timeout 600 cudanavierstokes_b200/cudns_run case=tgv mx=512 my=512 mz=512 stencilSize=4 nsteps=30 nfiles=1 ngpus=2 precision=1 par2_enstrophy=1 outdir=/tmp/run3 xdmf=0 2>&1 | grep -E "cudns_run|total time|file number"
} | tee gpurun_out/r2_driver_2gpu.log
