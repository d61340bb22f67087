#!/bin/bash
# Reference configurations used to pin the oracle (each is a separate compile of the
# reference, because every knob of src/globals.h is a macro).  Small grids so that the
# outputs can be committed as fixtures under tests/golden/.
TGV="Lx=(2.0*M_PI) Ly=(2.0*M_PI) Lz=(2.0*M_PI) mx_tot=24 my_tot=24 mz_tot=24 nsteps=10 nfiles=2 CFL=0.5f \
boundaryLayer=(false) perturbed=(false) forcing=(false) periodicX=(true) nonUniformX=(false) \
checkCFLcondition=10 checkBulk=10 Re=1600.0 Pr=1.0 Ma=0.1 viscexp=1.0 nDivZ=(2)"
declare -A CFG
CFG[tgv24_s3v3_ls]="$TGV stencilSize=3 stencilVisc=3 lowStorage=(true)"
CFG[tgv24_s4v4_kutta]="$TGV stencilSize=4 stencilVisc=4 lowStorage=(false)"
CFG[tgv24_s4v4_ls]="$TGV stencilSize=4 stencilVisc=4 lowStorage=(true)"
CFG[tgv24_s4v2_ls]="$TGV stencilSize=4 stencilVisc=2 lowStorage=(true)"
CFG[tgv24_s2v2_ls]="$TGV stencilSize=2 stencilVisc=2 lowStorage=(true)"
CFG[tgv24_s1v1_ls]="$TGV stencilSize=1 stencilVisc=1 lowStorage=(true)"
# supersonic channel (globals/channel.h) on a small grid, dt/forcing refresh every 5 steps
CHAN="Lx=(2.0) Ly=(2.0*M_PI) Lz=(4.0*M_PI) mx_tot=32 my_tot=24 mz_tot=24 nsteps=10 nfiles=2 CFL=0.75f \
lowStorage=(true) boundaryLayer=(false) perturbed=(false) forcing=(true) periodicX=(false) nonUniformX=(true) \
checkCFLcondition=5 checkBulk=5 Re=2800.0 Pr=0.75 Ma=1.5 viscexp=0.75 stretch=3.0 nDivZ=(2)"
CFG[chan_s3v2]="$CHAN stencilSize=3 stencilVisc=2"
CFG[chan_s2v2]="$CHAN stencilSize=2 stencilVisc=2"
# boundary layer (src/globals.h) on a small grid; perturbation strip kC=110,LP=40 needs mz>130
BL="Lx=(20.0) Ly=(7.0) Lz=(500.0) mx_tot=48 my_tot=16 mz_tot=192 nsteps=10 nfiles=2 CFL=0.75f \
lowStorage=(true) boundaryLayer=(true) perturbed=(true) forcing=(false) periodicX=(false) nonUniformX=(true) \
checkCFLcondition=5 checkBulk=5 Re=1500.0 Pr=0.75 Ma=0.35 viscexp=1.5 stretch=5.0 nDivZ=(8)"
CFG[bl_s3v2]="$BL stencilSize=3 stencilVisc=2"
# ---- performance baselines: "the reference's own GPU build" on BASELINE configs C2/C5.
# nsteps is compile-time and the reference only prints a whole-run wall clock that includes
# set-up and file I/O, so each size is built twice; (t_long - t_short)/(n_long - n_short) is its
# per-step time.  Not goldens: outputs are discarded.
declare -A PERF
P256="Lx=(2.0*M_PI) Ly=(2.0*M_PI) Lz=(2.0*M_PI) mx_tot=256 my_tot=256 mz_tot=256 nfiles=1 CFL=0.5f \
boundaryLayer=(false) perturbed=(false) forcing=(false) periodicX=(true) nonUniformX=(false) lowStorage=(true) \
checkCFLcondition=10 checkBulk=10 Re=1600.0 Pr=1.0 Ma=0.1 viscexp=1.0 nDivZ=(8) stencilSize=4 stencilVisc=4"
PERF[perf256_n10]="$P256 nsteps=10"
PERF[perf256_n40]="$P256 nsteps=40"
P512="${P256//256/512}"
PERF[perf512_n5]="$P512 nsteps=5"
PERF[perf512_n15]="$P512 nsteps=15"
# (a longer run: the difference of two whole-run wall clocks carries ~0.5 s of set-up noise, 10 steps of 0.3 s drown in it)
PERF[perf512_n45]="$P512 nsteps=45"
