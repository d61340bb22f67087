"""Parameter sets of the reference golden runs (mirror of oracle/refbuild/configs.sh)."""
import math

_TGV = dict(mx=24, my=24, mz=24, Lx=2 * math.pi, Ly=2 * math.pi, Lz=2 * math.pi, CFL=0.5,
            boundaryLayer=0, perturbed=0, forcing=0, periodicX=1, nonUniformX=0, lowStorage=1,
            checkCFLcondition=10, checkBulk=10, Re=1600.0, Pr=1.0, Ma=0.1, viscexp=1.0, stretch=5.0,
            nsteps=10, case="tgv")
_CHAN = dict(mx=32, my=24, mz=24, Lx=2.0, Ly=2 * math.pi, Lz=4 * math.pi, CFL=0.75,
             boundaryLayer=0, perturbed=0, forcing=1, periodicX=0, nonUniformX=1, lowStorage=1,
             checkCFLcondition=5, checkBulk=5, Re=2800.0, Pr=0.75, Ma=1.5, viscexp=0.75, stretch=3.0,
             nsteps=10, case="channel")
_BL = dict(mx=48, my=16, mz=192, Lx=20.0, Ly=7.0, Lz=500.0, CFL=0.75,
           boundaryLayer=1, perturbed=1, forcing=0, periodicX=0, nonUniformX=1, lowStorage=1,
           checkCFLcondition=5, checkBulk=5, Re=1500.0, Pr=0.75, Ma=0.35, viscexp=1.5, stretch=5.0,
           nsteps=10, case="blayer")

CONFIGS = {
    "tgv24_s3v3_ls": dict(_TGV, stencilSize=3, stencilVisc=3),
    "tgv24_s4v4_kutta": dict(_TGV, stencilSize=4, stencilVisc=4, lowStorage=0),
    "tgv24_s4v4_ls": dict(_TGV, stencilSize=4, stencilVisc=4),
    "tgv24_s4v2_ls": dict(_TGV, stencilSize=4, stencilVisc=2),
    "tgv24_s2v2_ls": dict(_TGV, stencilSize=2, stencilVisc=2),
    "tgv24_s1v1_ls": dict(_TGV, stencilSize=1, stencilVisc=1),
    "chan_s3v2": dict(_CHAN, stencilSize=3, stencilVisc=2),
    "chan_s2v2": dict(_CHAN, stencilSize=2, stencilVisc=2),
    "bl_s3v2": dict(_BL, stencilSize=3, stencilVisc=2),
}
