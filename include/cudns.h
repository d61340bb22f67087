/*
 * cudns.h -- C ABI of libcudns: a B200-native (sm_100a) replacement for the GPU side of
 * CUDA-DNS (simone-silvestri/CudaNavierStokes): right-hand-side evaluation + Runge-Kutta
 * time advance of the 3-D compressible Navier-Stokes equations, with halo exchange.
 *
 * The reference has no FFI; its seam is the set of free functions that src/main.cpp calls
 * into the .cu files (src/main.h:23-42) plus host globals (src/globals.h:97-105).  Every entry
 * point below names the reference function(s) it replaces (paths relative to the reference
 * repository).  Plain pointers and sizes only; all device memory is owned by the library, all
 * host arrays by the caller.  Every function returns 0 on success or a CUDNS_E* code;
 * cudns_last_error() returns a human-readable message for the calling thread.
 *
 * Array layout (host side, identical to the reference, src/globals.h:60): x fastest,
 * index = i + j*mx + k*mx*my; for a multi-rank run each rank passes its own z-slab
 * [mz_local][my][mx].  State = (r,u,v,w,e) = (rho, u, v, w, rho*E): PRIMITIVE velocities.
 */
#ifndef CUDNS_H_
#define CUDNS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CUDNS_OK          0
#define CUDNS_EINVAL      1   /* bad parameter (reference: printf + exit(1), cuda_utils.cu:55-58,233-297) */
#define CUDNS_ECUDA       2   /* CUDA runtime/driver error (reference: silently ignored, cuda_globals.h:12-22) */
#define CUDNS_ENOMEM      3
#define CUDNS_ESTATE      4   /* call order violation (e.g. advance before set_state) */
#define CUDNS_EUNSUPPORTED 5

/* Every knob of src/globals.h:14-58, src/sponge.h:5-17 and src/perturbation.h:14-21, as runtime
 * values (the reference fixes them at compile time).  stencilSize/stencilVisc select pre-built
 * kernel instantiations. */
typedef struct cudns_params {
    int mx, my, mz;            /* GLOBAL grid mx_tot,my_tot,mz_tot                    globals.h:23-25 */
    int stencilSize;           /* advective half-width s in {1,2,3,4} (order 2s)      globals.h:17 */
    int stencilVisc;           /* viscous half-width  v <= s                          globals.h:18 */
    double Lx, Ly, Lz;         /*                                                     globals.h:20-22 */
    double CFL;                /* globals.h:28 (float literal there: pass (double)(float)value) */
    int lowStorage;            /* 1 Wray low-storage RK3, 0 Kutta RK3                 globals.h:31 */
    int boundaryLayer;         /*                                                     globals.h:32 */
    int perturbed;             /* wall blowing/suction strip                          globals.h:33 */
    int forcing;               /* body force dpdz + controller                        globals.h:34 */
    int periodicX;             /*                                                     globals.h:35 */
    int nonUniformX;           /*                                                     globals.h:36 */
    int checkCFLcondition;     /* dt (and dpdz) refresh cadence in steps              globals.h:38 */
    int checkBulk;             /* bulk diagnostics cadence                            globals.h:39 */
    double Re, Pr, Ma, viscexp, gam;   /*                                             globals.h:41-45 */
    double stretch;            /*                                                     globals.h:50 */
    double TwallTop, TwallBot; /*                                                     globals.h:51-52 */
    double spTopStr, spTopLen, spTopExp;   /* sponge.h:5-7  */
    double spInlStr, spInlLen, spInlExp;   /* sponge.h:10-12 */
    double spOutStr, spOutLen, spOutExp;   /* sponge.h:15-17 */
    int    kC, LP;                          /* perturbation.h:16-17 */
    double amp1, amp2, omega1, omega2;      /* perturbation.h:19-21 (lambda = Ly/2pi is derived) */
    int quirk_q1;              /* 1: replicate src/cuda_rhs.cu:175 (y-dissipation uses dvdz); 0: dwdy */
    int rk4;                   /* 1: classical RK4 (extension; README.md:26 names it, the code lacks it) */
    /* --- decomposition (replaces pRow/pCol, globals.h:14-15, and Communicator, main.h:10-21):
     * z-slabs, one rank per GPU; x and y are never split. */
    int nranks;                /* number of z-slabs (1 = single GPU)                  */
    int rank;                  /* this process' slab index: k in [rank*mz/nranks, (rank+1)*mz/nranks) */
    int device;                /* CUDA device ordinal for this rank (setDevice, cuda_utils.cu:815) */
    int par2_enstrophy;        /* extension, default 0 = reference behaviour (calc_stress.cu:191-197 never writes par2 without forcing):
                                * 1 makes cudns_advance record the mean square vorticity <w.w> of an unforced periodic box in par2 */
    int precision;             /* `myprec` (globals.h:5-6) as a run-time value: 0 = double (default), 1 = float.  Device state, coefficient
                                * tables and all kernel arithmetic run in that precision; host arrays crossing this ABI are double in both
                                * (copyField casts, cuda_utils.cu:317-355), reduced scalars and statistics are accumulated in double.
                                * float serves every set-up double does (walls, stretched x, boundary layer, any viscosity law); outside
                                * the periodic / uniform / linear-viscosity case it needs mx % 4 == 0 (16-byte rows) */
    int reserved[3];
} cudns_params;

typedef struct cudns_solver *cudns_handle;

const char *cudns_last_error(void);
const char *cudns_version(void);

/* presets: python-utils/CompNavierStokes.py:1-7 (Taylor-Green), globals/channel.h, src/globals.h */
int cudns_params_tgv(cudns_params *p, int n, int stencil);
int cudns_params_channel(cudns_params *p);
int cudns_params_blayer(cudns_params *p);

/* ---- host-side set-up helpers (no GPU needed): restate src/init.cpp -------------------- */
/* initGrid (init.cpp:32-91): x[mx], xp[mx] (=1/x'), xpp[mx], y[my], z[mz]; returns host dx in *dx. */
int cudns_init_grid(const cudns_params *p, double *x, double *xp, double *xpp, double *y, double *z, double *dx);
/* initCHIT (init.cpp:126-148) / initChannel (init.cpp:94-124; libc rand(), reference loop order).
 * Arrays are the GLOBAL field [mz][my][mx]. */
int cudns_init_chit(const cudns_params *p, const double *x, const double *y, const double *z,
                    double *r, double *u, double *v, double *w, double *e);
int cudns_init_channel(const cudns_params *p, const double *x, const double *y, const double *z,
                       double *r, double *u, double *v, double *w, double *e);
/* calculateSponge, host half (sponge.cu:83-195): sponge strengths + conservative reference state from
 * the Blasius profile arrays (blasius1D/{x,r,u,w,e}Prof.bin, n entries each), with a correct 0-based
 * natural spline.  sigma_x[mx], sigma_z[mz], ref[5][mx*mz] (index i + k*mx); if r..e are non-NULL the
 * initial field (restartFile<0 branch) is written too. */
int cudns_build_sponge(const cudns_params *p, const double *x, const double *z,
                       const double *xIn, const double *rIn, const double *uIn, const double *wIn, int n,
                       double *sigma_x, double *sigma_z, double *ref5,
                       double *r, double *u, double *v, double *w, double *e);
/* python-utils/selfSimilarSol.py:1-98 without Python: compressible self-similar (Blasius) boundary-layer profiles for
 * cudns_build_sponge -- x = wall distance / delta_99, r, u (wall-normal), w (streamwise), e = T Rgas/(gam-1); n >= 600 entries
 * each (the reference's blasius1D/{x,r,u,w,e}Prof.bin hold n = 1000).  Shooting + RK4 instead of scipy.solve_bvp. */
int cudns_blasius_profiles(double gam, double Ma, double Pr, int n, double *x, double *r, double *u, double *w, double *e);
/* writeField / initField (init.cpp:13-30 -> comm.cpp:205-279): fields/<c>.<%07d>.bin, raw float64,
 * [mz_tot][my_tot][mx_tot], no header.  dir is the directory that contains "fields". */
int cudns_write_field(const char *dir, char name, int timestep, const double *var, size_t count);
int cudns_read_field(const char *dir, char name, int timestep, double *var, size_t count);
/* XDMF 2.0 sidecar describing fields/<c>.<%07d>.bin for ParaView/VisIt (python-utils/writexmf.py:1-77, makexmf.py): rectilinear
 * mesh with inline coordinates, one temporal collection; names = one character per field ("ruvwe"), time of step t = t*dt. */
int cudns_write_xdmf(const char *path, int single_precision, const double *x, int nx, const double *y, int ny, const double *z, int nz,
                     const int *timesteps, int nt, double dt, const char *names);

/* ---- solver life cycle ----------------------------------------------------------------- */
/* setDevice + setGPUParameters + initSolver (cuda_utils.cu:815,49-179,415-519).  x,xp,xpp are the
 * arrays of cudns_init_grid (the caller may substitute its own metric tables). */
int cudns_create(const cudns_params *p, const double *x, const double *xp, const double *xpp, cudns_handle *out);
/* clearSolver (cuda_utils.cu:521-595) */
int cudns_destroy(cudns_handle h);
/* checkGpuMem (cuda_utils.cu:780-813): bytes held by this solver / free / total on its device */
int cudns_memory_report(cudns_handle h, size_t *solver_bytes, size_t *free_bytes, size_t *total_bytes);

/* copyField(0) (cuda_utils.cu:299-333): host slab -> device, ghost fill, time_on_GPU = 0 and the
 * first calcState.  copyField(1) (:334-355): device -> host slab. */
int cudns_set_state(cudns_handle h, const double *r, const double *u, const double *v, const double *w, const double *e);
int cudns_get_state(cudns_handle h, double *r, double *u, double *v, double *w, double *e);
/* same, but the five arrays already live on this solver's device (contiguous [mz_local][my][mx]) */
int cudns_set_state_device(cudns_handle h, const double *d_r, const double *d_u, const double *d_v, const double *d_w, const double *d_e);
int cudns_get_state_device(cudns_handle h, double *d_r, double *d_u, double *d_v, double *d_w, double *d_e);
/* copySpongeToDevice (sponge.cu:43-81): sigma_x[mx], sigma_z[mz_local], ref5 = 5 tables [mx*mz_local] */
int cudns_set_sponge(cudns_handle h, const double *sigma_x, const double *sigma_z, const double *ref5);

/* runSimulationLowStorage / runSimulation (cuda_main.cu:44-186) = one "file" of solverWrapper
 * (cuda_main.cu:267-327): advance nsteps steps entirely on the device.  time/par1/par2 are host
 * arrays of nsteps entries or NULL (par1/par2 only written where istep % checkBulk == 0, like the
 * reference; other entries untouched).  The istep cadence restarts at 0 on every call (quirk Q11). */
int cudns_advance(cudns_handle h, int nsteps, double *time, double *par1, double *par2);

/* calcRHS (cuda_main.cu:15-42): full right-hand side of (rho, rho u, rho v, rho w, rho E) at the
 * current state into five host slabs (test entry point; the product path never materialises it). */
int cudns_calc_rhs(cudns_handle h, double *rhs_r, double *rhs_u, double *rhs_v, double *rhs_w, double *rhs_e);
/* calcTimeStep (calc_stress.cu:122-160) + allReduceToMin (comm.cpp:294) */
int cudns_calc_dt(cudns_handle h, double *dt);
/* calcBulk (calc_stress.cu:162-201) + allReduceSum (comm.cpp:326) */
int cudns_calc_bulk(cudns_handle h, double *par1, double *par2);
/* mean square vorticity <w.w> = volume average (the weights of calcBulk's <u.u>) of |curl u|^2 with the viscous-order differences of
 * derVelX/Y/Z (calc_stress.cu:20-86), summed over the slabs: the dissipation measure of a Taylor-Green run (epsilon = <w.w>/Re in
 * the incompressible limit).  Extension -- the reference has no counterpart; periodic boxes only (CUDNS_EUNSUPPORTED otherwise). */
int cudns_calc_enstrophy(cudns_handle h, double *enstrophy);
/* device scalars dtC, dpdz (cuda_utils.cu:65-75), accumulated time (cuda_main.cu:117-119) */
int cudns_get_scalars(cudns_handle h, double *dt, double *dpdz, double *time);
/* fix dt (tests): fixed != 0 disables the CFL refresh inside cudns_advance */
int cudns_set_dt(cudns_handle h, double dt, int fixed);

/* ---- multi-GPU halo plumbing (replaces updateHalo[Five] comm.cpp:90-134 and
 *      fillBoundaries[Five] cuda_utils.cu:597-778) ------------------------------------------
 * z-slab neighbours exchange (s+v) full padded planes of the 5 state fields once per RK stage.
 * Two transports:
 *  (a) peer memory: each rank publishes a CUDA IPC handle of its state allocation
 *      (cudns_halo_local_info); after cudns_halo_connect() the stage kernel stores its first / last
 *      (s+v) planes straight into the neighbours' ghost planes over NVLink while it computes, and a
 *      device-side epoch flag per neighbour replaces the host synchronisation;
 *  (b) external: the caller moves the bytes (e.g. NCCL send/recv on views of the buffers returned
 *      by cudns_halo_buffers) between cudns_stage_begin/cudns_stage_end.
 * Scalar reductions (dt: MIN, bulk/forcing: SUM) are delegated to a caller-provided callback so the
 * library does not link an MPI or NCCL of its own. */
#define CUDNS_IPC_HANDLE_BYTES 64
typedef struct cudns_peer_info {
    unsigned char mem_handle[CUDNS_IPC_HANDLE_BYTES];   /* cudaIpcMemHandle_t of the solver's state block (+ mailbox) */
    int device;
    int pid;
    uint64_t local_ptr;          /* the block's address in the owning process (used when both ranks share a process) */
    uint64_t block_bytes;        /* size of the block: must be equal on neighbouring ranks */
} cudns_peer_info;
int cudns_halo_local_info(cudns_handle h, cudns_peer_info *mine);
/* lower = rank-1 (periodic), upper = rank+1 (periodic) */
int cudns_halo_connect(cudns_handle h, const cudns_peer_info *lower, const cudns_peer_info *upper);
/* device pointers + byte counts of the contiguous send/recv plane blocks (transport (b)):
 * send_lo/send_hi are the first/last (s+v) interior planes (packed 5 fields), recv_lo/recv_hi the
 * ghost blocks. */
int cudns_halo_buffers(cudns_handle h, void **send_lo, void **send_hi, void **recv_lo, void **recv_hi, size_t *bytes_each);
typedef void (*cudns_allreduce_fn)(void *user, double *vals, int n, int op /*0 min, 1 sum, 2 max*/);
int cudns_set_allreduce(cudns_handle h, cudns_allreduce_fn fn, void *user);
typedef void (*cudns_exchange_fn)(void *user, void *stream /* cudaStream_t */);
/* transport (b): called once per RK stage (and once in set_state) on the solver's stream after the
 * send blocks are packed; must enqueue the exchange on that stream (or synchronise itself). */
int cudns_set_exchange(cudns_handle h, cudns_exchange_fn fn, void *user);
/* Several GPUs of one box from ONE process, without MPI or NCCL: n solvers created with nranks = n, rank = 0..n-1 on n devices, one
 * host thread each (the role of the reference's MPI ranks, src/comm.cpp).  cudns_team_create maps the slabs' state blocks into each
 * other (peer access: transport (a)) and installs host-side all-reduce / exchange callbacks that meet at a barrier; afterwards every
 * collective call (cudns_set_state, cudns_advance, cudns_calc_*, cudns_stats_*, ...) must be made by all n threads in the same order.
 * Destroy the team after the solvers. */
typedef struct cudns_team *cudns_team_handle;
int cudns_team_create(cudns_handle *solvers, int n, cudns_team_handle *out);
int cudns_team_destroy(cudns_team_handle t);
/* the CUDA stream the solver launches on (cudaStream_t), for event timing by the caller */
int cudns_get_stream(cudns_handle h, void **stream);

/* counters: kernels launched by this solver since creation / algorithmic bytes moved */
int cudns_get_counters(cudns_handle h, uint64_t *kernel_launches, uint64_t *rk_stages);
/* per-kernel device time of the last cudns_profile_stage() call, in ms: theta, rhs_stage, halo */
int cudns_profile_stage(cudns_handle h, int reps, float *ms_theta, float *ms_rhs, float *ms_halo);
/* per-kernel device times of the step loop itself (no reference counterpart: the reference only prints whole-run wall time,
 * main.cpp:80-96): cudns_set_stage_timing(h, 1) resets the sums and makes every later cudns_advance bracket the dilatation pass,
 * the stage kernel and the hand-shake of each stage with CUDA events (up to 1024 stages per call); cudns_get_stage_timing returns
 * the summed milliseconds and the number of stages they cover -- the kernel time under sustained load, next to the isolated-launch
 * time of cudns_profile_stage */
int cudns_set_stage_timing(cudns_handle h, int on);
int cudns_get_stage_timing(cudns_handle h, double *theta_ms, double *stage_ms, double *halo_ms, uint64_t *nstages);

/* ---- output and restart that do not stall the step loop ------------------------------------------
 * writeField (init.cpp:23-30) -> saveFileMPI (comm.cpp:205-250): fields/{r,u,v,w,e}.<%07d>.bin of the GLOBAL grid, raw float64
 * [mz_tot][my][mx].  cudns_write_fields_async snapshots the current state on the device (stream-ordered with cudns_advance),
 * returns at once, copies it to pinned host memory on a second stream and lets a writer thread pwrite() this rank's slab at its
 * byte offset (the role of the reference's MPI-IO subarray view).  One snapshot is in flight per handle: the next call waits
 * for the previous one to reach the disk.  dir is the directory that contains "fields" (created if missing). */
int cudns_write_fields_async(cudns_handle h, const char *dir, int timestep);
/* block until every snapshot is on disk; returns (and clears) the first writer error; files_written may be NULL */
int cudns_io_wait(cudns_handle h, uint64_t *files_written);
/* initField (init.cpp:13-21) -> readFileMPI (comm.cpp:252-279) + copyField(0): restart from fields/{r,u,v,w,e}.<%07d>.bin; every
 * rank reads its own slab */
int cudns_read_fields(cudns_handle h, const char *dir, int timestep);

/* ---- on-device diagnostics: no copyField(1) + host loops ------------------------------------------
 * calcAvgChan (init.cpp:150-208): prof[10][mx] (host) = y-z means per wall-normal index of rho, Favre-averaged u, v, w, rho E
 * (rows 0-4) and the mean squares of rho, u, v, w, rho E about them (rows 5-9): the columns 2-11 of the reference's prof.txt. */
int cudns_calc_profiles(cudns_handle h, double *prof);
/* printRes (init.cpp:210-256): average friction Reynolds number of the wall at i = 0 */
int cudns_calc_retau(cudns_handle h, double *retau);

/* ---- post-processing statistics: postproc/post.cpp as device reductions ---------------------------------
 * The reference's tool reads fields/<c>.<first..last>.bin back to the host twice and loops there: Reynolds and Favre means per
 * wall-normal index (addMean, post.cpp:225-257), volume averages (the same with N = 1), friction Reynolds number and velocity
 * (calcRet, :280-326), then the mean squares about those means (addFluc, :201-223).  Here every snapshot is reduced on the device
 * from the solver's current state -- live (between cudns_advance calls) or re-read (cudns_read_fields):
 *     cudns_stats_begin(h, n);  n x { state; cudns_stats_add_mean(h); }  cudns_stats_finish_mean(h);
 *     n x { the same states again; cudns_stats_add_fluc(h); }  cudns_stats_get(h, ...);
 * mean / fluc are [13][mx], bulk is [13], in the column order of Variables::printFile (post.cpp:61-86):
 * rho, uFavre, vFavre, wFavre, u, v, w, eTotal, hFavre, h, T, p, mu.  Slabs add up through the all-reduce callback (every rank gets the
 * result).  Re_tau / u_tau are 0 for periodic-x set-ups and for stencilSize = 4 (the reference's index arithmetic, post.cpp:303-306,
 * leaves its array there). */
int cudns_stats_begin(cudns_handle h, int nsnapshots);
int cudns_stats_add_mean(cudns_handle h);
int cudns_stats_finish_mean(cudns_handle h);
int cudns_stats_add_fluc(cudns_handle h);
int cudns_stats_get(cudns_handle h, double *mean, double *fluc, double *bulk, double *retau, double *utau);
/* Variables::printFile (post.cpp:61-86): mean.txt, fluc.txt, bulk.txt in outdir (NULL or "": the working directory); host only */
int cudns_stats_write(const char *outdir, int mx, const double *x, const double *mean, const double *fluc, const double *bulk, double retau, double utau);
/* main() of post.cpp (:126-199): both passes over fields/<c>.<first..last>.bin under dir, then the three files (written by rank 0) */
int cudns_postprocess(cudns_handle h, const char *dir, int first, int last, const double *x, const char *outdir);

#ifdef __cplusplus
}
#endif
#endif /* CUDNS_H_ */
