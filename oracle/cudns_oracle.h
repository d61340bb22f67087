/*
 * cudns_oracle.h -- CPU oracle for the CUDA-DNS right-hand-side + Runge-Kutta path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * algorithm (simone-silvestri/CudaNavierStokes, src/).  Only tests/, the smoke
 * check in __graft_entry__.py and bench.py's cpu_baseline / --impl reference leg
 * may link or call it.  The product (libcudns.so) never does.
 *
 * Parity status: PINNED against the reference's own GPU binary (built unmodified
 * from /root/reference/src by oracle/refbuild/, run on a B200, outputs committed
 * under tests/golden/ref_*.npz) -- see tests/test_oracle_vs_reference.py.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference/).
 */
#ifndef CUDNS_ORACLE_H_
#define CUDNS_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

/* All knobs of src/globals.h:14-58, src/sponge.h:5-17, src/perturbation.h:14-21
 * as runtime values (the reference has them as compile-time macros). */
typedef struct ora_params {
    int mx, my, mz;            /* mx_tot,my_tot,mz_tot (single rank: pRow=pCol=1)   globals.h:23-25 */
    int stencilSize;           /* advective half-width s (order 2s)                 globals.h:17 */
    int stencilVisc;           /* viscous half-width v <= s                         globals.h:18 */
    double Lx, Ly, Lz;         /* globals.h:20-22 */
    double CFL;                /* globals.h:28 (a float literal in the reference: pass (double)(float)x) */
    int lowStorage;            /* globals.h:31 */
    int boundaryLayer;         /* globals.h:32 */
    int perturbed;             /* globals.h:33 */
    int forcing;               /* globals.h:34 */
    int periodicX;             /* globals.h:35 */
    int nonUniformX;           /* globals.h:36 */
    int checkCFLcondition;     /* globals.h:38 */
    int checkBulk;             /* globals.h:39 */
    double Re, Pr, Ma, viscexp, gam;   /* globals.h:41-45 */
    double stretch;            /* globals.h:50 */
    double TwallTop, TwallBot; /* globals.h:51-52 */
    /* sponge.h:5-17 */
    double spTopStr, spTopLen, spTopExp;
    double spInlStr, spInlLen, spInlExp;
    double spOutStr, spOutLen, spOutExp;
    /* perturbation.h:14-21 */
    int    kC, LP;
    double amp1, amp2, omega1, omega2;     /* lambda = Ly/(2 pi) is derived */
    /* reference quirk Q1 (cuda_rhs.cu:175): y-dissipation uses dvdz where dwdy is meant.
     * 1 = replicate the reference (default for parity), 0 = physically consistent form. */
    int quirk_q1;
    /* time scheme: 0 = what lowStorage selects (reference), 4 = classical RK4 (extension,
     * not in the reference; validated by temporal convergence only). */
    int rk4;
} ora_params;

typedef struct ora_solver ora_solver;

/* fill *p with the Python-wrapper defaults for the Taylor-Green case
 * (python-utils/CompNavierStokes.py:1-7) on an n^3 grid, 2*pi box. */
void ora_params_tgv(ora_params *p, int n, int stencil);
/* globals/channel.h preset */
void ora_params_channel(ora_params *p);
/* src/globals.h as shipped (= globals/boundaryLayer.h with Lz=500,nDivZ=8) */
void ora_params_blayer(ora_params *p);

/* create: runs initGrid (init.cpp:32-91) and the coefficient/metric part of
 * setGPUParameters (cuda_utils.cu:49-139). */
ora_solver *ora_create(const ora_params *p);
void ora_destroy(ora_solver *s);
/* host threads used by the OpenMP loops (CPU-baseline timing) */
void ora_set_threads(int n);
int ora_get_threads(void);

/* grid + metrics access (length mx / my / mz) */
const double *ora_x(const ora_solver *s);
const double *ora_xp(const ora_solver *s);
const double *ora_xpp(const ora_solver *s);
const double *ora_y(const ora_solver *s);
const double *ora_z(const ora_solver *s);
const double *ora_dxv(const ora_solver *s);
/* coeffVSx table [ (2v+1) * mx ] (cuda_utils.cu:107-122) */
const double *ora_coeffVSx(const ora_solver *s);
double ora_dx(const ora_solver *s);   /* the host global dx after initGrid */

/* host state arrays r,u,v,w,e  [k][j][i], x fastest (globals.h:60,101-105) */
double *ora_r(ora_solver *s); double *ora_u(ora_solver *s); double *ora_v(ora_solver *s);
double *ora_w(ora_solver *s); double *ora_e(ora_solver *s);

/* initial conditions */
void ora_init_chit(ora_solver *s);       /* init.cpp:126-148 */
void ora_init_channel(ora_solver *s);    /* init.cpp:94-124 ; uses libc rand() in the reference's loop order */
/* sponge strengths + reference state (sponge.cu:115-129,160-195). profile arrays as read from
 * blasius1D/{x,r,u,w,e}Prof.bin; n = number of entries.  Also fills the IC (restartFile<0).
 * The spline is the reference's 1-based Numerical-Recipes routine called on 0-based arrays
 * (quirk Q9): pass correct_spline=1 for a proper 0-based spline (what libcudns' host side uses). */
void ora_set_sponge_from_profiles(ora_solver *s, const double *xIn, const double *rIn,
                                  const double *uIn, const double *wIn, const double *eIn, int n,
                                  int fill_ic);
/* direct table access (spongeX[mx], spongeZ[mz], rref..eref [k*mx+i], conservative refs) */
double *ora_spongeX(ora_solver *s); double *ora_spongeZ(ora_solver *s);
double *ora_ref(ora_solver *s, int which); /* 0 r,1 ru,2 rv,3 rw,4 e */

/* copyField(0): (cuda_utils.cu:299-333) -- take the host state as the solver state and run
 * the first calcState */
void ora_copy_field_in(ora_solver *s);

/* calcState (cuda_main.cu:218-242) on the current state -> h,t,p,mu,lam */
void ora_calc_state(ora_solver *s);
const double *ora_derived(const ora_solver *s, int which); /* 0 h,1 t,2 p,3 mu,4 lam,5 dil, 6..14 gij[0..8] */

/* calcRHS (cuda_main.cu:15-42): full right-hand side of (rho, rho u, rho v, rho w, rho E)
 * into rhs[5][mx*my*mz]. */
void ora_calc_rhs(ora_solver *s, double *rhs[5]);

/* calcTimeStep (calc_stress.cu:122-160) using the mu of the last calcState; returns dt */
double ora_calc_dt(ora_solver *s);
/* calcBulk (calc_stress.cu:162-201) */
void ora_calc_bulk(ora_solver *s, double *par1, double *par2);
/* mean square vorticity at viscous order (libcudns's par2 of the unforced runs; not a reference quantity) */
double ora_calc_enstrophy(ora_solver *s);
/* calcAvgChan init.cpp:150-208: prof[10][mx] = y-z means (rho, Favre u,v,w, rho E) and mean squares about them */
void ora_calc_profiles(ora_solver *s, double *prof);
/* postproc/post.cpp: statistics over a series of saved fields.  create(nfiles); for every file: load it into the solver
 * (ora state arrays + ora_copy_field_in) and add_mean; finish_mean; for every file again: add_fluc; get.  mean/fluc are [13][mx]
 * in the column order of Variables::printFile (rho uF vF wF u v w e hF h T p mu), bulk[13] the volume averages */
typedef struct ora_post ora_post;
ora_post *ora_post_create(ora_solver *s, int nfiles);
void ora_post_destroy(ora_post *p);
void ora_post_add_mean(ora_post *p, ora_solver *s);
void ora_post_finish_mean(ora_post *p);
void ora_post_add_fluc(ora_post *p, ora_solver *s);
void ora_post_get(const ora_post *p, double *mean, double *fluc, double *bulk, double *Ret, double *ut);
/* printRes init.cpp:210-256: average friction Reynolds number at the wall i = 0 */
double ora_calc_retau(ora_solver *s);

/* runSimulationLowStorage / runSimulation (cuda_main.cu:44-186): advance nsteps steps
 * ("one file").  time/par1/par2 may be NULL or arrays of nsteps entries (par entries are
 * only written at istep % checkBulk == 0, others left untouched). */
void ora_run(ora_solver *s, int nsteps, double *time, double *par1, double *par2);

double ora_get_dt(const ora_solver *s);
double ora_get_dpdz(const ora_solver *s);
double ora_get_time(const ora_solver *s);
void   ora_set_dt(ora_solver *s, double dt);   /* for fixed-dt tests */
void   ora_set_fixed_dt(ora_solver *s, int on); /* 1: never refresh dt inside ora_run */

/* standalone operator entry points for known-answer tests: 1-D periodic line of n points */
void ora_kat_flux_cube(int s, int n, double invd, const double *f, const double *g, const double *h, double *out);
void ora_kat_flux_quad(int s, int n, double invd, const double *f, const double *g, double *out);
void ora_kat_d1(int s, int n, double invd, const double *f, double *out);
void ora_kat_d2(int s, int n, double invd2, const double *f, double *out);

#ifdef __cplusplus
}
#endif
#endif
