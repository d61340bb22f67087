/*
 * cudns_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see cudns_oracle.h).
 *
 * Plain-C restatement of the reference's RHS + Runge-Kutta path.  The structure
 * follows the reference literally (calcState -> derVel -> calcDil -> RHS X/Y/Z ->
 * sponge -> RK update), one pencil at a time with s ghost cells per side exactly
 * like the reference's shared-memory pencils, so that every boundary rule can be
 * copied 1:1.  No attempt is made to be fast beyond an OpenMP loop over pencils.
 *
 * Citations: paths relative to /root/reference/.
 */
#include "cudns_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXS 4
#define MAXD(a,b) ((a) >= (b) ? (a) : (b))   /* main.h:7 */
#define MIND(a,b) ((a) <= (b) ? (a) : (b))   /* main.h:8 */

/* globals.h:69-95 (far-to-near ordering) */
static const double CF1[] = {-1.0/2.0};
static const double CS1[] = {1.0, -2.0};
static const double CF2[] = { 1.0/12.0, -2.0/3.0};
static const double CS2[] = {-1.0/12.0,  4.0/3.0, -5.0/2.0};
static const double CF3[] = {-1.0/60.0,  3.0/20.0, -3.0/4.0};
static const double CS3[] = { 1.0/90.0, -3.0/20.0,  3.0/2.0, -49.0/18.0};
static const double CF4[] = { 1.0/280.0, -4.0/105.0,  1.0/5.0, -4.0/5.0};
static const double CS4[] = {-1.0/560.0,  8.0/315.0, -1.0/5.0,  8.0/5.0,  -205.0/72.0};
static const double *CF_TAB[] = {0, CF1, CF2, CF3, CF4};
static const double *CS_TAB[] = {0, CS1, CS2, CS3, CS4};

struct ora_solver {
    ora_params P;
    int mx, my, mz, s, v;
    size_t N;
    double Rgas, Ec;
    const double *cF, *cS, *cVF, *cVS;
    /* grid */
    double dx;                  /* host global dx (init.cpp:36 / :63) */
    double *x, *xp, *xpp, *y, *z, *dxv;
    double d_dx, d_dy, d_dz, d_d2x, d_d2y, d_d2z;   /* cuda_utils.cu:61-63,88-90 */
    double *coeffVSx;           /* cuda_utils.cu:107-122 */
    /* host state (globals.h:101-105) == device state after copyField(0) */
    double *r, *u, *v_, *w, *e;
    /* derived (cuda_main.h) */
    double *h, *t, *p, *m, *l, *dil, *gij[9];
    double *rhs1[5], *rhs2[5], *rhs3[5], *old[5];
    /* sponge */
    double *spongeX, *spongeZ, *ref[5];
    int have_sponge;
    /* scalars on "device" */
    double dtC, dpdz, time_on_GPU, time_last;
    int fixed_dt;
};

/* ------------------------------------------------------------------ presets */
static void params_common(ora_params *p) {
    memset(p, 0, sizeof(*p));
    p->lowStorage = 1; p->quirk_q1 = 1;
    p->gam = 1.4; p->stretch = 5.0; p->TwallTop = 1.0; p->TwallBot = 1.0;
    p->spTopStr = 1.0; p->spTopLen = 1.0; p->spTopExp = 2.0;   /* sponge.h:5-17 */
    p->spInlStr = 0.5; p->spInlLen = 20.0; p->spInlExp = 2.0;
    p->spOutStr = 0.5; p->spOutLen = 20.0; p->spOutExp = 2.0;
    p->kC = 110; p->LP = 40; p->amp1 = 2e-4; p->amp2 = 3e-5; p->omega2 = 10.0; /* perturbation.h:16-21 */
}

void ora_params_tgv(ora_params *p, int n, int stencil) {   /* CompNavierStokes.py:1-7 + examples.py */
    params_common(p);
    p->mx = p->my = p->mz = n; p->stencilSize = stencil; p->stencilVisc = stencil;
    p->Lx = p->Ly = p->Lz = 2.0*M_PI;
    p->CFL = 0.5; p->periodicX = 1; p->nonUniformX = 0;
    p->checkCFLcondition = 10; p->checkBulk = 10;
    p->Re = 1600.0; p->Pr = 1.0; p->Ma = 0.1; p->viscexp = 1.0;
    p->omega1 = p->Re*121.e-6;
}

void ora_params_channel(ora_params *p) {    /* globals/channel.h:17-52 */
    params_common(p);
    p->mx = 160; p->my = 192; p->mz = 192; p->stencilSize = 3; p->stencilVisc = 2;
    p->Lx = 2.0; p->Ly = 2.0*M_PI; p->Lz = 4.0*M_PI;
    p->CFL = (double)0.75f; p->forcing = 1; p->periodicX = 0; p->nonUniformX = 1;
    p->checkCFLcondition = 100; p->checkBulk = 100;
    p->Re = 2800.0; p->Pr = 0.75; p->Ma = 1.5; p->viscexp = 0.75; p->stretch = 3.0;
    p->omega1 = p->Re*121.e-6;
}

void ora_params_blayer(ora_params *p) {     /* src/globals.h:17-52 */
    params_common(p);
    p->mx = 240; p->my = 64; p->mz = 2048; p->stencilSize = 3; p->stencilVisc = 2;
    p->Lx = 20.0; p->Ly = 7.0; p->Lz = 500.0;
    p->CFL = (double)0.75f; p->boundaryLayer = 1; p->perturbed = 1; p->periodicX = 0; p->nonUniformX = 1;
    p->checkCFLcondition = 100; p->checkBulk = 100;
    p->Re = 1500.0; p->Pr = 0.75; p->Ma = 0.35; p->viscexp = 1.5; p->stretch = 5.0;
    p->omega1 = p->Re*121.e-6;
}

/* ------------------------------------------------------------------ grid */
/* init.cpp:258-277 */
static void derivGrid(const ora_solver *S, double *d2f, double *df, const double *f, double dx) {
    int mx = S->mx, s = S->s; double Lx = S->P.Lx;
    double *fb = (double*)malloc(sizeof(double)*(mx+2*s));
    for (int i = s; i < mx+s; i++) fb[i] = f[i-s];
    for (int i = 0; i < s; i++) {
        fb[i] = -fb[2*s-i-1];
        fb[mx+s+i] = 2*Lx - fb[mx+s-i-1];
    }
    for (int i = 0; i < mx; i++) {
        df[i] = 0.0;
        d2f[i] = S->cS[s]*fb[i+s]/dx/dx;
        for (int it = 0; it < s; it++) {
            df[i]  += S->cF[it]*(fb[i+it]-fb[i+s*2-it])/dx;
            d2f[i] += S->cS[it]*(fb[i+it]+fb[i+s*2-it])/dx/dx;
        }
    }
    free(fb);
}

/* init.cpp:32-91 */
static void initGrid(ora_solver *S) {
    const ora_params *P = &S->P; int mx = S->mx;
    S->dx = P->Lx*(1.0)/(mx);
    double *xn = (double*)malloc(sizeof(double)*(mx+1));
    int denom = mx; double denom2 = 2.0;
    if (P->boundaryLayer) { denom *= 2; denom2 /= 2; }
    for (int i = 0; i < mx+1; i++)
        xn[i] = tanh(P->stretch*((i*1.0)/denom-0.5))/tanh(P->stretch*0.5);
    for (int i = 0; i < mx; i++)
        S->x[i] = P->Lx * (1.0 + (xn[i] + xn[i+1])/2.0)/denom2;
    derivGrid(S, S->xpp, S->xp, S->x, S->dx);
    for (int i = 0; i < mx; i++) S->xp[i] = 1.0/S->xp[i];
    if (!P->nonUniformX) {
        for (int i = 0; i < mx; i++) S->x[i] = P->Lx*(0.5+i*1.0)/(mx);
        S->dx = S->x[1] - S->x[0];
    }
    for (int j = 0; j < S->my; j++) S->y[j] = P->Ly*(0.5+j*1.0)/(S->my);
    for (int k = 0; k < S->mz; k++) S->z[k] = P->Lz*(0.5+k*1.0)/(S->mz);
    free(xn);
}

/* cuda_utils.cu:60-122 */
static void setGPUParameters(ora_solver *S) {
    int mx = S->mx, v = S->v;
    double h_dx = 1.0/(S->dx);
    double h_dy = 1.0/(S->y[1] - S->y[0]);
    double h_dz = 1.0/(S->z[1] - S->z[0]);
    S->dxv[0] = (S->x[1]+S->x[0])/2.0;
    for (int i = 1; i < mx-1; i++) S->dxv[i] = (S->x[i+1]-S->x[i-1])/2.0;
    S->dxv[mx-1] = S->P.Lx - (S->x[mx-1]+S->x[mx-2])/2.0;
    double h_d2x = h_dx*h_dx, h_d2y = h_dy*h_dy, h_d2z = h_dz*h_dz;
    const double *xp = S->xp, *xpp = S->xpp;
    for (int it = 0; it < v; it++)
        for (int i = 0; i < mx; i++)
            S->coeffVSx[i+it*mx] = (S->cVS[it]*(xp[i]*xp[i])*h_d2x - S->cVF[it]*xpp[i]*(xp[i]*xp[i]*xp[i])*h_dx);
    for (int i = 0; i < mx; i++)
        S->coeffVSx[i+v*mx] = S->cVS[v]*(xp[i]*xp[i])*h_d2x;
    for (int it = v+1; it < 2*v+1; it++)
        for (int i = 0; i < mx; i++)
            S->coeffVSx[i+it*mx] = (S->cVS[2*v-it]*(xp[i]*xp[i])*h_d2x + S->cVF[2*v-it]*xpp[i]*(xp[i]*xp[i]*xp[i])*h_dx);
    S->d_dx = h_dx; S->d_dy = h_dy; S->d_dz = h_dz;
    S->d_d2x = h_d2x; S->d_d2y = h_d2y; S->d_d2z = h_d2z;
    S->dpdz = S->P.forcing ? 0.00372 : 0.0;     /* cuda_utils.cu:68-70 */
}

static double *dalloc(size_t n) { double *p = (double*)calloc(n, sizeof(double)); if (!p) { fprintf(stderr,"oracle: out of memory\n"); exit(1);} return p; }

ora_solver *ora_create(const ora_params *p) {
    ora_solver *S = (ora_solver*)calloc(1, sizeof(ora_solver));
    S->P = *p; S->mx = p->mx; S->my = p->my; S->mz = p->mz; S->s = p->stencilSize; S->v = p->stencilVisc;
    if (S->s < 1 || S->s > MAXS || S->v < 1 || S->v > S->s) { free(S); return NULL; }
    S->N = (size_t)p->mx*p->my*p->mz;
    S->Rgas = (1.f/(p->gam*p->Ma*p->Ma));          /* globals.h:48 */
    S->Ec   = ((p->gam - 1.f)*p->Ma*p->Ma);        /* globals.h:47 */
    S->cF = CF_TAB[S->s]; S->cS = CS_TAB[S->s]; S->cVF = CF_TAB[S->v]; S->cVS = CS_TAB[S->v];
    S->x = dalloc(S->mx); S->xp = dalloc(S->mx); S->xpp = dalloc(S->mx); S->dxv = dalloc(S->mx);
    S->y = dalloc(S->my); S->z = dalloc(S->mz);
    S->coeffVSx = dalloc((size_t)S->mx*(2*S->v+1));
    S->r = dalloc(S->N); S->u = dalloc(S->N); S->v_ = dalloc(S->N); S->w = dalloc(S->N); S->e = dalloc(S->N);
    S->h = dalloc(S->N); S->t = dalloc(S->N); S->p = dalloc(S->N); S->m = dalloc(S->N); S->l = dalloc(S->N);
    S->dil = dalloc(S->N);
    for (int i = 0; i < 9; i++) S->gij[i] = dalloc(S->N);
    for (int i = 0; i < 5; i++) { S->rhs1[i] = dalloc(S->N); S->rhs2[i] = dalloc(S->N); }
    if (!p->lowStorage || p->rk4) for (int i = 0; i < 5; i++) { S->rhs3[i] = dalloc(S->N); S->old[i] = dalloc(S->N); }
    S->spongeX = dalloc(S->mx); S->spongeZ = dalloc(S->mz);
    for (int i = 0; i < 5; i++) S->ref[i] = dalloc((size_t)S->mx*S->mz);
    initGrid(S);
    setGPUParameters(S);
    return S;
}

/* host threads of the OpenMP regions (bench.py's CPU baseline: torch.distributed.run exports OMP_NUM_THREADS=1 to its workers) */
void ora_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int ora_get_threads(void) { return omp_get_max_threads(); }

void ora_destroy(ora_solver *S) {
    if (!S) return;
    free(S->x); free(S->xp); free(S->xpp); free(S->dxv); free(S->y); free(S->z); free(S->coeffVSx);
    free(S->r); free(S->u); free(S->v_); free(S->w); free(S->e);
    free(S->h); free(S->t); free(S->p); free(S->m); free(S->l); free(S->dil);
    for (int i = 0; i < 9; i++) free(S->gij[i]);
    for (int i = 0; i < 5; i++) { free(S->rhs1[i]); free(S->rhs2[i]); free(S->rhs3[i]); free(S->old[i]); free(S->ref[i]); }
    free(S->spongeX); free(S->spongeZ);
    free(S);
}

const double *ora_x(const ora_solver *s) { return s->x; }
const double *ora_xp(const ora_solver *s) { return s->xp; }
const double *ora_xpp(const ora_solver *s) { return s->xpp; }
const double *ora_y(const ora_solver *s) { return s->y; }
const double *ora_z(const ora_solver *s) { return s->z; }
const double *ora_dxv(const ora_solver *s) { return s->dxv; }
const double *ora_coeffVSx(const ora_solver *s) { return s->coeffVSx; }
double ora_dx(const ora_solver *s) { return s->dx; }
double *ora_r(ora_solver *s) { return s->r; } double *ora_u(ora_solver *s) { return s->u; }
double *ora_v(ora_solver *s) { return s->v_; } double *ora_w(ora_solver *s) { return s->w; }
double *ora_e(ora_solver *s) { return s->e; }
double *ora_spongeX(ora_solver *s) { s->have_sponge = 1; return s->spongeX; }
double *ora_spongeZ(ora_solver *s) { s->have_sponge = 1; return s->spongeZ; }
double *ora_ref(ora_solver *s, int which) { return s->ref[which]; }
double ora_get_dt(const ora_solver *s) { return s->dtC; }
double ora_get_dpdz(const ora_solver *s) { return s->dpdz; }
double ora_get_time(const ora_solver *s) { return s->time_last; }
void ora_set_dt(ora_solver *s, double dt) { s->dtC = dt; }
void ora_set_fixed_dt(ora_solver *s, int on) { s->fixed_dt = on; }
const double *ora_derived(const ora_solver *s, int which) {
    switch (which) { case 0: return s->h; case 1: return s->t; case 2: return s->p; case 3: return s->m;
                     case 4: return s->l; case 5: return s->dil; default: return s->gij[which-6]; }
}

#define IDX(i,j,k) ((size_t)(k)*mx*my + (size_t)(j)*mx + (i))   /* globals.h:60 */

/* ------------------------------------------------------------------ initial conditions */
/* init.cpp:126-148 */
void ora_init_chit(ora_solver *S) {
    int mx = S->mx, my = S->my, mz = S->mz; const ora_params *P = &S->P;
    double V0 = 1.0, T0 = 1.0, P0 = T0*S->Rgas, R0 = 1.0, gam = P->gam;
    for (int i = 0; i < mx; i++) {
        double fx = 2*M_PI*S->x[i]/P->Lx;
        for (int j = 0; j < my; j++) {
            double fy = 2*M_PI*S->y[j]/P->Ly;
            for (int k = 0; k < mz; k++) {
                double fz = 2*M_PI*S->z[k]/P->Lz;
                size_t g = IDX(i,j,k);
                S->u[g]  =  V0*sin(fx/1.0)*cos(fy/1.0)*cos(fz/1.0);
                S->v_[g] = -V0*cos(fx/1.0)*sin(fy/1.0)*cos(fz/1.0);
                S->w[g]  =  0.0;
                double press = P0 + 1.0/16.0*R0*V0*V0 * (cos(2.0*fx/1.0) + cos(2.0*fy/1.0)) * (cos(2.0*fz/1.0) + 2.0);
                S->r[g] = press/S->Rgas/T0;
                S->e[g] = press/(gam-1.0) + 0.5 * S->r[g] * (pow(S->u[g],2) + pow(S->v_[g],2) + pow(S->w[g],2));
            } } }
}

/* init.cpp:94-124 */
void ora_init_channel(ora_solver *S) {
    int mx = S->mx, my = S->my, mz = S->mz; const ora_params *P = &S->P;
    double T0 = 1.0, P0 = T0*S->Rgas, R0 = 1.0, gam = P->gam;
    double U0 = pow(gam,0.5)*P->Ma;
    srand(1);   /* the reference never seeds: glibc default sequence */
    for (int i = 0; i < mx; i++) {
        for (int j = 0; j < my; j++) {
            for (int k = 0; k < mz; k++) {
                double rr1 = rand()*1.0/(RAND_MAX*1.0) - 0.5;
                double rr2 = rand()*1.0/(RAND_MAX*1.0) - 0.5;
                double rr3 = rand()*1.0/(RAND_MAX*1.0) - 0.5;
                double ufluc = 0.02*rr1, vfluc = 0.02*rr2, wfluc = 0.02*rr3;
                double wmean = 1.5*U0*R0*S->x[i]*(1.0-S->x[i]/P->Lx);
                ufluc = ufluc + 0.05*sin(0.5*M_PI*S->x[i])*cos(2*M_PI*S->y[j]);
                vfluc = vfluc + 0.05*sin(0.5*M_PI*S->x[i])*sin(2*M_PI*S->y[j]);
                size_t g = IDX(i,j,k);
                S->u[g] = ufluc; S->v_[g] = vfluc; S->w[g] = wmean + wfluc;
                S->r[g] = R0;
                S->e[g] = P0/(gam-1.0) + 0.5 * S->r[g] * (pow(S->u[g],2) + pow(S->v_[g],2) + pow(S->w[g],2));
            } } }
}

/* natural cubic spline, 0-based (Numerical Recipes spline/splint, sponge.cu:262-313) */
static void spline0(const double *x, const double *y, int n, double *y2) {
    double *u = (double*)malloc(sizeof(double)*n);
    y2[0] = u[0] = 0.0;
    for (int i = 1; i <= n-2; i++) {
        double sig = (x[i]-x[i-1])/(x[i+1]-x[i-1]);
        double p = sig*y2[i-1]+2.0;
        y2[i] = (sig-1.0)/p;
        u[i] = (y[i+1]-y[i])/(x[i+1]-x[i]) - (y[i]-y[i-1])/(x[i]-x[i-1]);
        u[i] = (6.0*u[i]/(x[i+1]-x[i-1])-sig*u[i-1])/p;
    }
    double qn = 0.0, un = 0.0;
    y2[n-1] = (un-qn*u[n-2])/(qn*y2[n-2]+1.0);
    for (int k = n-2; k >= 0; k--) y2[k] = y2[k]*y2[k+1]+u[k];
    free(u);
}
static double splint0(const double *xa, const double *ya, const double *y2a, int n, double x) {
    int klo = 0, khi = n-1;
    while (khi-klo > 1) { int k = (khi+klo) >> 1; if (xa[k] > x) khi = k; else klo = k; }
    double h = xa[khi]-xa[klo];
    double a = (xa[khi]-x)/h, b = (x-xa[klo])/h;
    return a*ya[klo]+b*ya[khi]+((a*a*a-a)*y2a[klo]+(b*b*b-b)*y2a[khi])*(h*h)/6.0;
}

/* sponge.cu:115-129 (strengths), :160-195 (reference state + IC), :47-56 (conservative refs) */
void ora_set_sponge_from_profiles(ora_solver *S, const double *xIn, const double *rIn,
                                  const double *uIn, const double *wIn, const double *eIn, int n,
                                  int fill_ic) {
    const ora_params *P = &S->P; int mx = S->mx, my = S->my, mz = S->mz; (void)eIn;
    S->have_sponge = 1;
    for (int i = 0; i < mx; i++) {
        S->spongeX[i] = 0.0;
        if ((P->spTopLen > 0.0) && (S->x[i] >= P->Lx - P->spTopLen))
            S->spongeX[i] = P->spTopStr*pow((S->x[i] - (P->Lx-P->spTopLen))/P->spTopLen, P->spTopExp);
    }
    for (int k = 0; k < mz; k++) {
        S->spongeZ[k] = 0.0;
        double fz = S->z[k];
        if ((P->spInlLen > 0.0) && (fz <= P->spInlLen))
            S->spongeZ[k] = P->spInlStr*pow((P->spInlLen-fz)/P->spInlLen, P->spInlExp);
        if ((P->spOutLen > 0.0) && (fz >= (P->Lz-P->spOutLen)))
            S->spongeZ[k] = P->spOutStr*pow((fz - (P->Lz-P->spOutLen))/P->spOutLen, P->spOutExp);
    }
    double *r2 = (double*)malloc(sizeof(double)*n), *u2 = (double*)malloc(sizeof(double)*n), *w2 = (double*)malloc(sizeof(double)*n);
    spline0(xIn, rIn, n, r2); spline0(xIn, uIn, n, u2); spline0(xIn, wIn, n, w2);
    double gam = P->gam;
    for (int k = 0; k < mz; k++)
        for (int i = 0; i < mx; i++) {
            double scale = pow(1 + S->z[k]/P->Re, 0.5);
            double rr = splint0(xIn, rIn, r2, n, S->x[i]/scale);
            double uu = splint0(xIn, uIn, u2, n, S->x[i]/scale);
            uu /= (scale*P->Re);
            double ww = splint0(xIn, wIn, w2, n, S->x[i]/scale);
            double pconst = S->Rgas;
            double ee = pconst/(gam-1.0) + rr*0.5*(uu*uu+ww*ww);
            size_t q = (size_t)i + (size_t)k*mx;
            /* copySpongeToDevice (sponge.cu:58-62): conservative references */
            S->ref[0][q] = rr; S->ref[1][q] = uu*rr; S->ref[2][q] = 0.0; S->ref[3][q] = ww*rr; S->ref[4][q] = ee;
            if (fill_ic)
                for (int j = 0; j < my; j++) {
                    size_t g = IDX(i,j,k);
                    S->r[g] = rr; S->u[g] = uu; S->v_[g] = 0.0; S->w[g] = ww; S->e[g] = ee;
                }
        }
    free(r2); free(u2); free(w2);
}

/* ------------------------------------------------------------------ calcState (cuda_main.cu:218-242) */
void ora_calc_state(ora_solver *S) {
    const ora_params *P = &S->P;
    const double gam = P->gam, Rgas = S->Rgas, Re = P->Re, Pr = P->Pr, Ec = S->Ec, viscexp = P->viscexp;
    const double cvInv = (gam - 1.0)/Rgas;
    #pragma omp parallel for schedule(static)
    for (long gl = 0; gl < (long)S->N; gl++) {
        double invrho = 1.0/S->r[gl];
        double en = S->e[gl]*invrho - 0.5*(S->u[gl]*S->u[gl] + S->v_[gl]*S->v_[gl] + S->w[gl]*S->w[gl]);
        S->t[gl] = cvInv*en;
        S->p[gl] = S->r[gl]*Rgas*S->t[gl];
        S->h[gl] = (S->e[gl] + S->p[gl])*invrho;
        double suth = pow(S->t[gl], viscexp);
        S->m[gl] = suth/Re;
        S->l[gl] = suth/Re/Pr/Ec;
    }
}

void ora_copy_field_in(ora_solver *S) {   /* cuda_utils.cu:317-333 + initDevice :372-383 */
    S->time_on_GPU = 0.0;
    ora_calc_state(S);
}

/* ------------------------------------------------------------------ pencil stencils (cuda_derivs.h) */
typedef struct {
    int s, v, n;        /* ghost width, viscous width, pencil length */
    const double *cF, *cS, *cVF, *cVS;
    double d1, d2;      /* d_dx / d_d2x of this direction */
    int dir;            /* 0 x, 1 y, 2 z */
    int nonuni;         /* nonUniformX && dir==0 */
    const double *xp, *cVSx; int mx;
} pen_ops;

/* fluxQuadShared{x,y,z} cuda_derivs.h:30-53,77-96,120-139 */
static inline double fluxQuad(const pen_ops *o, const double *f, const double *g, int si) {
    double flxp = 0.0, flxm = 0.0; int s = o->s;
    for (int lt = 1; lt < s+1; lt++)
        for (int mt = 0; mt < lt; mt++) {
            flxp -= o->cF[s-lt]*(f[si-mt]+f[si-mt+lt])*(g[si-mt]+g[si-mt+lt]);
            flxm -= o->cF[s-lt]*(f[si-mt-1]+f[si-mt+lt-1])*(g[si-mt-1]+g[si-mt+lt-1]);
        }
    double df = 0.5*o->d1*(flxm - flxp);
    if (o->nonuni) df = df*o->xp[si-s];
    return df;
}
/* fluxCubeShared{x,y,z} cuda_derivs.h:55-75,98-118,141-155 */
static inline double fluxCube(const pen_ops *o, const double *f, const double *g, const double *h, int si) {
    double flxp = 0.0, flxm = 0.0; int s = o->s;
    for (int lt = 1; lt < s+1; lt++)
        for (int mt = 0; mt < lt; mt++) {
            flxp -= o->cF[s-lt]*(f[si-mt]+f[si-mt+lt])*(g[si-mt]+g[si-mt+lt])*(h[si-mt]+h[si-mt+lt]);
            flxm -= o->cF[s-lt]*(f[si-mt-1]+f[si-mt+lt-1])*(g[si-mt-1]+g[si-mt+lt-1])*(h[si-mt-1]+h[si-mt+lt-1]);
        }
    double df = 0.25*o->d1*(flxm - flxp);
    if (o->nonuni) df = df*o->xp[si-s];
    return df;
}
/* derDevShared1{x,y,z} cuda_derivs.h:157-171,234-242,271-279 */
static inline double der1A(const pen_ops *o, const double *f, int si) {
    double df = 0.0; int s = o->s;
    if (o->dir == 0) {
        for (int it = 0; it < s; it++) df += o->cF[it]*(f[si+it-s]-f[si+s-it]);
        df = df*o->d1;
        if (o->nonuni) df = df*o->xp[si-s];
    } else {
        for (int it = 0; it < s; it++) df += o->cF[it]*(f[si+it-s]-f[si+s-it])*o->d1;
    }
    return df;
}
/* derDevSharedV1{x,y,z} cuda_derivs.h:192-205,256-264,293-301 */
static inline double der1V(const pen_ops *o, const double *f, int si) {
    double df = 0.0; int v = o->v;
    if (o->dir == 0) {
        for (int it = 0; it < v; it++) df += o->cVF[it]*(f[si+it-v]-f[si+v-it]);
        df = df*o->d1;
        if (o->nonuni) df = df*o->xp[si-o->s];
    } else {
        for (int it = 0; it < v; it++) df += o->cVF[it]*(f[si+it-v]-f[si+v-it])*o->d1;
    }
    return df;
}
/* derDevSharedV2{x,y,z} cuda_derivs.h:207-226,266-269(y),303-312(z) */
static inline double der2V(const pen_ops *o, const double *f, int si) {
    int v = o->v; double d2f;
    if (o->nonuni) {
        d2f = 0.0;
        for (int it = 0; it < 2*v+1; it++) d2f += o->cVSx[it*o->mx+(si-o->s)]*(f[si+it-v]);
    } else {
        d2f = o->cVS[v]*f[si]*o->d2;
        for (int it = 0; it < v; it++) d2f += o->cVS[it]*(f[si+it-v]+f[si+v-it])*o->d2;
    }
    return d2f;
}

/* ------------------------------------------------------------------ ghost fills (boundary.h) */
/* pencil arrays have n+2s entries; "g" below is the reference's id.i (0..s-1), si = g+s */
static inline void perBC(double *f, int g, int s, int n)        { f[g] = f[g+n]; f[g+s+n] = f[g+s]; }          /* boundary.h:38-51 */
static inline void wallMir(double *f, int g, int s, int n)      { f[g] = f[2*s-g-1]; f[g+s+n] = f[n+s-g-1]; }   /* :116-119 */
static inline void wallVel(double *f, int g, int s, int n)      { f[g] = -f[2*s-g-1]; f[g+s+n] = -f[n+s-g-1]; } /* :111-114 */
static inline void wallExt(double *f, int g, int s, int n, double top, double bot) {                           /* :101-104 */
    f[g] = 2.0*bot - f[2*s-g-1]; f[g+s+n] = 2.0*top - f[n+s-g-1]; }
static inline void topExt(double *f, int g, int s, int n)       { f[g+s+n] = 2.0*f[n+s-1] - f[n+s-g-2]; }       /* :150-152 and :154-156 */
static inline void botExtNode(double *f, int g, int s)          { f[g] = 2.0*f[s] - f[2*s-g]; }                  /* botBCzExt :158-160 */
static inline void botExtCell(double *f, int g, int s, double b){ f[g] = 2.0*b - f[2*s-g-1]; }                   /* botBCxExt :162-164 */
static inline void botMir(double *f, int g, int s)              { f[g] = f[2*s-g-1]; }                          /* botBCxMir :178-180 */

/* perturbation.h:25-53.  Sets ghost g of u (all s ghosts get the same value). */
static inline void perturbU(const ora_solver *S, double *su, int g, int j, int k) {
    const ora_params *P = &S->P;
    int kSt = P->kC - P->LP/2, kEn = P->kC + P->LP/2, ktot = k;
    int alpha, beta, kappa;
    if (ktot >= kSt && ktot <= kEn) {
        if (ktot < P->kC) { kappa = 1; alpha = ktot - kSt; beta = P->kC - kSt; }
        else              { kappa = -1; alpha = kEn - ktot; beta = kEn - P->kC; }
        double ksi = alpha*1.0/beta;
        double gg = (15.1875*ksi*ksi*ksi*ksi*ksi) - (35.4375*ksi*ksi*ksi*ksi) + (20.25*ksi*ksi*ksi);
        double y_glob = (j)/S->d_dy;
        double lambda = P->Ly/(2.0*M_PI);
        su[g] = P->amp1*kappa*gg*sin(P->omega1*S->time_on_GPU) + P->amp2*kappa*gg*sin(P->omega2*S->time_on_GPU)*cos(y_glob/lambda);
    }
}

static inline void mlBound(const ora_solver *S, double *m, double *l, const double *t, int idx) {   /* boundary.h:135-148 */
    double suth = pow(t[idx], S->P.viscexp);
    m[idx] = suth/S->P.Re;
    l[idx] = suth/S->P.Re/S->P.Pr/S->Ec;
}
static inline void rhBound(const ora_solver *S, double *r, double *h, const double *p, const double *t,
                           const double *u, const double *v, const double *w, int idx) {             /* boundary.h:121-133 */
    h[idx] = t[idx]*S->Rgas*S->P.gam/(S->P.gam - 1.0) + 0.5*(u[idx]*u[idx]+v[idx]*v[idx]+w[idx]*w[idx]);
    r[idx] = p[idx]/(S->Rgas*t[idx]);
}

/* BCxderVel boundary_condition_x.h:23-45 */
static void bcx_dervel(const ora_solver *S, double *su, double *sv, double *sw, int j, int k) {
    int s = S->s, n = S->mx;
    for (int g = 0; g < s; g++) {
        if (S->P.periodicX) { perBC(su,g,s,n); perBC(sv,g,s,n); perBC(sw,g,s,n); }
        else if (S->P.boundaryLayer) {
            topExt(su,g,s,n); topExt(sv,g,s,n); topExt(sw,g,s,n);
            botExtCell(su,g,s,0.0); botExtCell(sv,g,s,0.0); botExtCell(sw,g,s,0.0);
            if (S->P.perturbed) perturbU(S, su, g, j, k);
        } else { wallVel(su,g,s,n); wallVel(sv,g,s,n); wallVel(sw,g,s,n); }
    }
}
/* BCxNumber1 boundary_condition_x.h:47-85 */
static void bcx_1(const ora_solver *S, double *su, double *sv, double *sw, double *sp, double *st, double *sm, double *sl, int j, int k) {
    int s = S->s, n = S->mx;
    for (int g = 0; g < s; g++) {
        if (S->P.periodicX) {
            perBC(su,g,s,n); perBC(sv,g,s,n); perBC(sw,g,s,n); perBC(st,g,s,n); perBC(sp,g,s,n); perBC(sm,g,s,n); perBC(sl,g,s,n);
        } else if (S->P.boundaryLayer) {
            topExt(su,g,s,n); topExt(sv,g,s,n); topExt(sw,g,s,n); topExt(sp,g,s,n); topExt(st,g,s,n);
            botMir(sp,g,s); botMir(st,g,s);
            botExtCell(su,g,s,0.0); botExtCell(sv,g,s,0.0); botExtCell(sw,g,s,0.0);
            if (S->P.perturbed) perturbU(S, su, g, j, k);
            mlBound(S, sm, sl, st, g); mlBound(S, sm, sl, st, g+s+n);
        } else {
            wallMir(sp,g,s,n); wallVel(su,g,s,n); wallVel(sv,g,s,n); wallVel(sw,g,s,n);
            wallExt(st,g,s,n,S->P.TwallTop,S->P.TwallBot);
            mlBound(S, sm, sl, st, g); mlBound(S, sm, sl, st, g+s+n);
        }
    }
}
/* BCxNumber2 boundary_condition_x.h:87-101 */
static void bcx_2(const ora_solver *S, double *sd) {
    int s = S->s, n = S->mx;
    for (int g = 0; g < s; g++) {
        if (S->P.periodicX) perBC(sd,g,s,n);
        else if (S->P.boundaryLayer) { topExt(sd,g,s,n); botMir(sd,g,s); }
        else wallMir(sd,g,s,n);
    }
}
/* BCxNumber3 boundary_condition_x.h:103-115 */
static void bcx_3(const ora_solver *S, const double *su, const double *sv, const double *sw, const double *sp, const double *st, double *sr, double *sh) {
    int s = S->s, n = S->mx;
    for (int g = 0; g < s; g++) {
        if (S->P.periodicX) { perBC(sr,g,s,n); perBC(sh,g,s,n); }
        else { rhBound(S,sr,sh,sp,st,su,sv,sw,g); rhBound(S,sr,sh,sp,st,su,sv,sw,g+s+n); }
    }
}
/* z ghost fill of one array: BCzNumber1-4 (boundary_condition_z.h:85-332), nDivZ sub-blocking is a no-op numerically */
static inline void bcz_one(const ora_solver *S, double *f) {
    int s = S->s, n = S->mz;
    for (int g = 0; g < s; g++) {
        if (S->P.boundaryLayer) { topExt(f,g,s,n); botExtNode(f,g,s); }
        else perBC(f,g,s,n);
    }
}
static inline void bcy_one(const ora_solver *S, double *f) {   /* boundary_condition_y.h:19-58 */
    int s = S->s, n = S->my;
    for (int g = 0; g < s; g++) perBC(f,g,s,n);
}

/* ------------------------------------------------------------------ velocity gradients (calc_stress.cu:20-96) */
static void derVel(ora_solver *S) {
    int mx = S->mx, my = S->my, mz = S->mz, s = S->s, v = S->v;
    int nmax = MAXD(mx, MAXD(my, mz)) + 2*s;
    #pragma omp parallel
    {
        double *bu = (double*)malloc(sizeof(double)*nmax*3), *bv = bu+nmax, *bw = bv+nmax;
        pen_ops o; o.s = s; o.v = v; o.cF = S->cF; o.cS = S->cS; o.cVF = S->cVF; o.cVS = S->cVS;
        o.xp = S->xp; o.cVSx = S->coeffVSx; o.mx = mx;
        /* derVelX :20-45 */
        o.dir = 0; o.n = mx; o.d1 = S->d_dx; o.d2 = S->d_d2x; o.nonuni = S->P.nonUniformX;
        #pragma omp for collapse(2) schedule(static)
        for (int k = 0; k < mz; k++) for (int j = 0; j < my; j++) {
            for (int i = 0; i < mx; i++) { size_t g = IDX(i,j,k); bu[i+s] = S->u[g]; bv[i+s] = S->v_[g]; bw[i+s] = S->w[g]; }
            bcx_dervel(S, bu, bv, bw, j, k);
            for (int i = 0; i < mx; i++) { size_t g = IDX(i,j,k);
                S->gij[0][g] = der1V(&o, bu, i+s); S->gij[1][g] = der1V(&o, bv, i+s); S->gij[2][g] = der1V(&o, bw, i+s); }
        }
        /* derVelY :47-54 + derDevV1yL cuda_derivs.h:314-354 (periodic wrap) */
        o.dir = 1; o.n = my; o.d1 = S->d_dy; o.d2 = S->d_d2y; o.nonuni = 0;
        #pragma omp for collapse(2) schedule(static)
        for (int k = 0; k < mz; k++) for (int i = 0; i < mx; i++) {
            for (int j = 0; j < my; j++) { size_t g = IDX(i,j,k); bu[j+s] = S->u[g]; bv[j+s] = S->v_[g]; bw[j+s] = S->w[g]; }
            bcy_one(S, bu); bcy_one(S, bv); bcy_one(S, bw);
            for (int j = 0; j < my; j++) { size_t g = IDX(i,j,k);
                S->gij[3][g] = der1V(&o, bu, j+s); S->gij[4][g] = der1V(&o, bv, j+s); S->gij[5][g] = der1V(&o, bw, j+s); }
        }
        /* derVelZ :56-63 + derDevV1zL cuda_derivs.h:356-386 + BCzderVel boundary_condition_z.h:25-83 */
        o.dir = 2; o.n = mz; o.d1 = S->d_dz; o.d2 = S->d_d2z; o.nonuni = 0;
        #pragma omp for collapse(2) schedule(static)
        for (int j = 0; j < my; j++) for (int i = 0; i < mx; i++) {
            for (int k = 0; k < mz; k++) { size_t g = IDX(i,j,k); bu[k+s] = S->u[g]; bv[k+s] = S->v_[g]; bw[k+s] = S->w[g]; }
            bcz_one(S, bu); bcz_one(S, bv); bcz_one(S, bw);
            for (int k = 0; k < mz; k++) { size_t g = IDX(i,j,k);
                S->gij[6][g] = der1V(&o, bu, k+s); S->gij[7][g] = der1V(&o, bv, k+s); S->gij[8][g] = der1V(&o, bw, k+s); }
        }
        free(bu);
    }
    /* calcDil :87-96 */
    #pragma omp parallel for schedule(static)
    for (long g = 0; g < (long)S->N; g++) S->dil[g] = S->gij[0][g] + S->gij[4][g] + S->gij[8][g];
}

/* ------------------------------------------------------------------ directional RHS (cuda_rhs.cu) */
static void rhs_dir(ora_solver *S, int dir, double *rhs[5]) {
    int mx = S->mx, my = S->my, mz = S->mz, s = S->s;
    int n = dir == 0 ? mx : (dir == 1 ? my : mz);
    int na = dir == 0 ? my : (dir == 1 ? mx : mx);   /* inner of the two outer loops */
    int nb = dir == 0 ? mz : (dir == 1 ? mz : my);
    int np = n + 2*s;
    /* gradient roles per direction (argument lists at cuda_main.cu:33,36,38):
       dself[i] = d u_i / d x_dir  ;  dother[i] = d u_dir / d x_i */
    const double *dself[3], *dother[3];
    for (int i = 0; i < 3; i++) { dself[i] = S->gij[3*dir + i]; dother[i] = S->gij[3*i + dir]; }
    const double *qdiss = (dir == 1 && S->P.quirk_q1) ? S->gij[7] : NULL;   /* Q1: cuda_rhs.cu:175 uses dvdz */
    const double dpdz = S->dpdz;
    #pragma omp parallel
    {
        double *buf = (double*)malloc(sizeof(double)*np*10);
        double *su = buf, *sv = su+np, *sw = sv+np, *st = sw+np, *sp = st+np, *sm = sp+np, *sl = sm+np, *sd = sl+np, *sr = sd+np, *sh = sr+np;
        pen_ops o; o.s = s; o.v = S->v; o.n = n; o.cF = S->cF; o.cS = S->cS; o.cVF = S->cVF; o.cVS = S->cVS;
        o.xp = S->xp; o.cVSx = S->coeffVSx; o.mx = mx; o.dir = dir;
        o.d1 = dir == 0 ? S->d_dx : (dir == 1 ? S->d_dy : S->d_dz);
        o.d2 = dir == 0 ? S->d_d2x : (dir == 1 ? S->d_d2y : S->d_d2z);
        o.nonuni = (dir == 0) && S->P.nonUniformX;
        #pragma omp for collapse(2) schedule(static)
        for (int b = 0; b < nb; b++) for (int a = 0; a < na; a++) {
            #define GIDX(q) (dir == 0 ? IDX(q,a,b) : (dir == 1 ? IDX(a,q,b) : IDX(a,b,q)))
            for (int q = 0; q < n; q++) { size_t g = GIDX(q); int si = q+s;
                su[si] = S->u[g]; sv[si] = S->v_[g]; sw[si] = S->w[g]; st[si] = S->t[g]; sp[si] = S->p[g];
                sm[si] = S->m[g]; sl[si] = S->l[g]; sd[si] = S->dil[g]; sr[si] = S->r[g]; sh[si] = S->h[g]; }
            if (dir == 0) { bcx_1(S, su, sv, sw, sp, st, sm, sl, a, b); bcx_2(S, sd); bcx_3(S, su, sv, sw, sp, st, sr, sh); }
            else if (dir == 1) { bcy_one(S,su); bcy_one(S,sv); bcy_one(S,sw); bcy_one(S,sm); bcy_one(S,sd); bcy_one(S,sp);
                                 bcy_one(S,sl); bcy_one(S,st); bcy_one(S,sr); bcy_one(S,sh); }
            else { bcz_one(S,su); bcz_one(S,sv); bcz_one(S,sw); bcz_one(S,sm); bcz_one(S,sd); bcz_one(S,sp);
                   bcz_one(S,sl); bcz_one(S,st); bcz_one(S,sr); bcz_one(S,sh); }
            const double *sU = dir == 0 ? su : (dir == 1 ? sv : sw);
            for (int q = 0; q < n; q++) { size_t g = GIDX(q); int si = q+s;
                double tmp[3], etmp, rtmp, wrk1, wrk2;
                /* stresses: cuda_rhs.cu:52-54 / :169-171 / :303-305 */
                for (int i = 0; i < 3; i++) tmp[i] = (i == dir) ? (2 * dself[i][g] - 2./3.*sd[si]) : (dself[i][g] + dother[i][g]);
                /* dissipation: :57 / :174 (Q1) / :308 */
                if (qdiss) etmp = sm[si]*(tmp[0]*dself[0][g] + tmp[1]*dself[1][g] + tmp[2]*qdiss[g]);
                else       etmp = sm[si]*(tmp[0]*dself[0][g] + tmp[1]*dself[1][g] + tmp[2]*dself[2][g]);
                wrk2 = der1V(&o, sm, si);                                  /* :60 */
                tmp[0] *= wrk2; tmp[1] *= wrk2; tmp[2] *= wrk2;
                wrk1 = der2V(&o, su, si); tmp[0] = tmp[0] + wrk1*sm[si];   /* :66-71 */
                wrk1 = der2V(&o, sv, si); tmp[1] = tmp[1] + wrk1*sm[si];
                wrk1 = der2V(&o, sw, si); tmp[2] = tmp[2] + wrk1*sm[si];
                etmp = etmp + su[si]*tmp[0] + sv[si]*tmp[1] + sw[si]*tmp[2];   /* :74 */
                if (dir == 0) {
                    /* x: conduction first (:77-81), then dilatation/pressure (:91-94) */
                    wrk1 = der2V(&o, st, si); etmp = etmp + wrk1*sl[si];
                    wrk2 = der1V(&o, sl, si); wrk1 = der1V(&o, st, si); etmp = etmp + wrk1*wrk2;
                    wrk2 = der1V(&o, sd, si); wrk1 = der1A(&o, sp, si);
                    tmp[0] = tmp[0] + sm[si]*wrk2/3.0 - wrk1;
                    etmp   = etmp   + sm[si]*wrk2/3.0*su[si];
                } else {
                    /* y,z: dilatation (:193-195/:327-329), pressure (:205-206/:337-338), conduction (:217-221/:349-353) */
                    wrk2 = der1V(&o, sd, si);
                    tmp[dir] = tmp[dir] + sm[si]*wrk2/3.0;
                    etmp     = etmp     + sm[si]*wrk2/3.0*sU[si];
                    wrk1 = der1A(&o, sp, si); tmp[dir] = tmp[dir] - wrk1;
                    wrk1 = der2V(&o, st, si); etmp = etmp + wrk1*sl[si];
                    wrk2 = der1V(&o, sl, si); wrk1 = der1V(&o, st, si); etmp = etmp + wrk1*wrk2;
                }
                /* advective split-form fluxes :107-121 / :231-245 / :365-379 */
                rtmp = fluxQuad(&o, sr, sU, si);
                tmp[0] = tmp[0] + fluxCube(&o, sr, sU, su, si);
                tmp[1] = tmp[1] + fluxCube(&o, sr, sU, sv, si);
                tmp[2] = tmp[2] + fluxCube(&o, sr, sU, sw, si);
                etmp   = etmp   + fluxCube(&o, sr, sU, sh, si);
                if (dir == 0) { rhs[0][g] = rtmp; rhs[1][g] = tmp[0]; rhs[2][g] = tmp[1]; rhs[3][g] = tmp[2]; rhs[4][g] = etmp; }   /* :123-127 */
                else if (dir == 1) { rhs[0][g] += rtmp; rhs[1][g] += tmp[0]; rhs[2][g] += tmp[1]; rhs[3][g] += tmp[2]; rhs[4][g] += etmp; } /* :254-258 */
                else { rhs[0][g] += rtmp; rhs[1][g] += tmp[0]; rhs[2][g] += tmp[1]; rhs[3][g] += tmp[2] + dpdz; rhs[4][g] += etmp + dpdz*sw[si]; } /* :389-393 */
            }
            #undef GIDX
        }
        free(buf);
    }
}

/* addSponge sponge.cu:31-41 */
static void addSponge(ora_solver *S, double *rhs[5]) {
    int mx = S->mx, my = S->my, mz = S->mz;
    #pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < mz; k++) for (int j = 0; j < my; j++) for (int i = 0; i < mx; i++) {
        size_t g = IDX(i,j,k), q = (size_t)i + (size_t)k*mx;
        double sg = (S->spongeX[i] + S->spongeZ[k]);
        rhs[0][g] += sg * (S->ref[0][q] - S->r[g]);
        rhs[1][g] += sg * (S->ref[1][q] - S->r[g]*S->u[g]);
        rhs[2][g] += sg * (S->ref[2][q] - S->r[g]*S->v_[g]);
        rhs[3][g] += sg * (S->ref[3][q] - S->r[g]*S->w[g]);
        rhs[4][g] += sg * (S->ref[4][q] - S->e[g]);
    }
}

/* calcRHS cuda_main.cu:15-42 */
void ora_calc_rhs(ora_solver *S, double *rhs[5]) {
    ora_calc_state(S);
    derVel(S);
    rhs_dir(S, 0, rhs);
    rhs_dir(S, 1, rhs);
    rhs_dir(S, 2, rhs);
    if (S->P.boundaryLayer) addSponge(S, rhs);
}

/* ------------------------------------------------------------------ dt, bulk, forcing */
/* calcTimeStep + deviceCalcDt calc_stress.cu:122-160, hostReduceToMin cuda_math.cu:308-331 */
double ora_calc_dt(ora_solver *S) {
    int mx = S->mx; const double gam = S->P.gam, CFL = S->P.CFL;
    double mn = 100000;
    #pragma omp parallel for reduction(min:mn) schedule(static)
    for (long g = 0; g < (long)S->N; g++) {
        int i = (int)(g % mx);
        double ien = S->e[g]/S->r[g] - 0.5*(S->u[g]*S->u[g] + S->v_[g]*S->v_[g] + S->w[g]*S->w[g]);
        double sos = pow(gam*(gam-1)*ien, 0.5);
        double dx = S->dxv[i], d2x = dx*dx;
        double dtConvInv = MAXD((fabs(S->u[g]) + sos)/dx, MAXD((fabs(S->v_[g]) + sos)*S->d_dy, (fabs(S->w[g]) + sos)*S->d_dz));
        double dtViscInv = MAXD(S->m[g]/d2x, MAXD(S->m[g]*S->d_d2y, S->m[g]*S->d_d2z));
        double val = CFL/MAXD(dtConvInv, dtViscInv);
        mn = MIND(mn, val);
    }
    return mn;
}

/* integrateThreads / AverageThreads cuda_math.cu:143-201 */
static double volumeIntegral(const ora_solver *S, const double *a, const double *b, int average) {
    int mx = S->mx; double sum = 0.0;
    const double Lx = S->P.Lx, Ly = S->P.Ly, Lz = S->P.Lz;
    #pragma omp parallel for reduction(+:sum) schedule(static)
    for (long g = 0; g < (long)S->N; g++) {
        int i = (int)(g % mx);
        double val = b ? a[g]*b[g] : a[g];
        if (average) sum += val*S->dxv[i]/S->d_dy/S->d_dz/Lx/Ly/Lz;
        else         sum += val*S->dxv[i]/S->d_dy/S->d_dz;
    }
    return sum;
}

/* calcBulk calc_stress.cu:162-201 */
void ora_calc_bulk(ora_solver *S, double *par1, double *par2) {
    if (S->P.forcing) {
        double rbulk = volumeIntegral(S, S->r, NULL, 0);
        double p1 = volumeIntegral(S, S->r, S->w, 0);
        *par1 = p1/rbulk;
        *par2 = volumeIntegral(S, S->e, NULL, 0);
    } else {
        int mx = S->mx; double sum = 0.0;
        const double Lx = S->P.Lx, Ly = S->P.Ly, Lz = S->P.Lz;
        #pragma omp parallel for reduction(+:sum) schedule(static)
        for (long g = 0; g < (long)S->N; g++) {
            int i = (int)(g % mx);
            double sca = S->u[g]*S->u[g] + S->v_[g]*S->v_[g] + S->w[g]*S->w[g];   /* deviceSca cuda_math.cu:26-29 */
            sum += sca*S->dxv[i]/S->d_dy/S->d_dz/Lx/Ly/Lz;
        }
        *par1 = sum;
        /* par2 is never written when forcing=false (quirk Q6) */
    }
}

/* Mean square vorticity <w.w> (volume average with the weights of calcBulk's <u.u>) from the viscous-order velocity gradients of
 * derVelX/Y/Z (calc_stress.cu:20-86).  NOT a reference quantity: the reference never writes par2 without forcing (quirk Q6), so a
 * Taylor-Green run has no dissipation history; this is the definition libcudns offers for it (epsilon = <w.w>/Re for the
 * incompressible limit), restated here so that the device reduction has a checker. */
double ora_calc_enstrophy(ora_solver *S) {
    derVel(S);
    int mx = S->mx; double sum = 0.0;
    const double Lx = S->P.Lx, Ly = S->P.Ly, Lz = S->P.Lz;
    #pragma omp parallel for reduction(+:sum) schedule(static)
    for (long g = 0; g < (long)S->N; g++) {
        int i = (int)(g % mx);
        /* gij[3*d + m] = d u_m / d x_d */
        double wx = S->gij[3*1+2][g] - S->gij[3*2+1][g], wy = S->gij[3*2+0][g] - S->gij[3*0+2][g], wz = S->gij[3*0+1][g] - S->gij[3*1+0][g];
        sum += (wx*wx + wy*wy + wz*wz)*S->dxv[i]/S->d_dy/S->d_dz/Lx/Ly/Lz;
    }
    return sum;
}

/* calcAvgChan init.cpp:150-208: y-z averages per wall-normal index i.  prof[0..4][i] = <rho>, <rho u>/<rho>, <rho v>/<rho>,
 * <rho w>/<rho>, <rho E>; prof[5..9][i] = mean squares of (rho, u, v, w, rho E) about those means (prof.txt columns 2..11) */
void ora_calc_profiles(ora_solver *S, double *prof) {
    const int mx = S->mx, my = S->my, mz = S->mz;
    double *rm = prof, *um = prof + mx, *vm = prof + 2*mx, *wm = prof + 3*mx, *em = prof + 4*mx;
    double *rf = prof + 5*mx, *uf = prof + 6*mx, *vf = prof + 7*mx, *wf = prof + 8*mx, *ef = prof + 9*mx;
    for (int i = 0; i < mx; i++) {
        rm[i] = um[i] = vm[i] = wm[i] = em[i] = 0.0; rf[i] = uf[i] = vf[i] = wf[i] = ef[i] = 0.0;
        for (int k = 0; k < mz; k++)
            for (int j = 0; j < my; j++) {
                size_t g = (size_t)i + (size_t)j*mx + (size_t)k*mx*my;
                rm[i] += S->r[g]/my/mz;
                um[i] += S->r[g]*S->u[g]/my/mz;
                vm[i] += S->r[g]*S->v_[g]/my/mz;
                wm[i] += S->r[g]*S->w[g]/my/mz;
                em[i] += S->e[g]/my/mz;
            }
    }
    for (int i = 0; i < mx; i++) {
        um[i] /= rm[i]; vm[i] /= rm[i]; wm[i] /= rm[i];
        for (int k = 0; k < mz; k++)
            for (int j = 0; j < my; j++) {
                size_t g = (size_t)i + (size_t)j*mx + (size_t)k*mx*my;
                rf[i] += (S->r[g]-rm[i])*(S->r[g]-rm[i])/my/mz;
                uf[i] += (S->u[g]-um[i])*(S->u[g]-um[i])/my/mz;
                vf[i] += (S->v_[g]-vm[i])*(S->v_[g]-vm[i])/my/mz;
                wf[i] += (S->w[g]-wm[i])*(S->w[g]-wm[i])/my/mz;
                ef[i] += (S->e[g]-em[i])*(S->e[g]-em[i])/my/mz;
            }
    }
}

/* printRes init.cpp:210-256: average friction Reynolds number of the wall at i = 0 (one-sided stencil on the anti-mirrored
 * streamwise velocity w, wall temperature 1) */
double ora_calc_retau(ora_solver *S) {
    const int mx = S->mx, my = S->my, mz = S->mz, s = S->s;
    double Ret = 0.0;
    for (int k = 0; k < mz; k++)
        for (int j = 0; j < my; j++) {
            double temw = 1.0;
            double suth = pow(temw, S->P.viscexp);
            double muw = suth/S->P.Re;
            double ub[2*4+1];
            for (int i = s; i < s*2+1; i++) ub[i] = S->w[i-s + j*mx + (size_t)k*my*mx];
            for (int i = 0; i < s; i++)     ub[i] = S->w[s-i-1 + j*mx + (size_t)k*my*mx];
            double dudx = 0;
            for (int i = 0; i < s; i++) dudx += S->cF[i]*(ub[i]-ub[s*2-i])/S->dx;
            dudx *= S->xp[0];
            size_t g0 = (size_t)j*mx + (size_t)k*mx*my;
            double ut = sqrt(muw*fabs(dudx)/S->r[g0]);       /* the reference calls the integer abs(); fabs is what is meant */
            Ret += ut*S->r[g0]/muw;
        }
    return Ret/my/mz;
}

/* ------------------------------------------------------------------ post-processing statistics (postproc/post.cpp)
 * Restatement of the reference's post-processing tool: Reynolds / Favre means and the mean squares about them per wall-normal
 * index, volume averages, friction Reynolds number and friction velocity, over a series of saved fields (post.cpp:126-326).
 * Pinned by running the reference's own tool (oracle/refbuild/build_ref_post.sh -> oracle/_ref/post_*) on the same fields/:
 * tests/golden/ref_post_*.npz, tests/test_post.py.  Row order of the 13 quantities = column order of Variables::printFile
 * (post.cpp:61-86): rho, uFavre, vFavre, wFavre, u, v, w, eTotal, hFavre, h, T, p, mu. */
enum { PQ_R = 0, PQ_UF, PQ_VF, PQ_WF, PQ_U, PQ_V, PQ_W, PQ_E, PQ_HF, PQ_H, PQ_T, PQ_P, PQ_M, PQ_N };
struct ora_post {
    int mx, nfiles;
    double denom;
    double *mean, *fluc;        /* [13][mx] */
    double bulk[PQ_N];
    double Ret, ut;
};
ora_post *ora_post_create(ora_solver *S, int nfiles) {
    ora_post *P = (ora_post*)calloc(1, sizeof(ora_post));
    P->mx = S->mx; P->nfiles = nfiles;
    P->denom = (double)(nfiles*S->my*S->mz);               /* post.cpp:166 (int arithmetic there) */
    P->mean = (double*)calloc((size_t)PQ_N*S->mx, sizeof(double));
    P->fluc = (double*)calloc((size_t)PQ_N*S->mx, sizeof(double));
    return P;
}
void ora_post_destroy(ora_post *P) { if (P) { free(P->mean); free(P->fluc); free(P); } }
/* calcState post.cpp:259-278 at one point */
static void post_state(const ora_solver *S, size_t g, double q[PQ_N]) {
    const double cvInv = (S->P.gam - 1.0)/S->Rgas;
    const double r = S->r[g], u = S->u[g], v = S->v_[g], w = S->w[g], e = S->e[g];
    const double invrho = 1.0/r;
    const double en = e*invrho - 0.5*(u*u + v*v + w*w);
    const double t = cvInv*en, p = r*S->Rgas*t, h = (e + p)*invrho, m = pow(t, S->P.viscexp)/S->P.Re;
    q[PQ_R] = r; q[PQ_UF] = r*u; q[PQ_VF] = r*v; q[PQ_WF] = r*w; q[PQ_U] = u; q[PQ_V] = v; q[PQ_W] = w; q[PQ_E] = e;
    q[PQ_HF] = r*h; q[PQ_H] = h; q[PQ_T] = t; q[PQ_P] = p; q[PQ_M] = m;
}
/* addMean (post.cpp:225-257) for `mean` and `bulk`, calcRet (:280-326), of the solver's current state */
void ora_post_add_mean(ora_post *P, ora_solver *S) {
    const int mx = S->mx, my = S->my, mz = S->mz, s = S->s;
    for (int i = 0; i < mx; i++) {
        const double denom2 = S->P.Lx/S->dxv[i]*P->denom;
        for (int k = 0; k < mz; k++) for (int j = 0; j < my; j++) {
            double q[PQ_N]; post_state(S, (size_t)i + (size_t)mx*(j + (size_t)my*k), q);
            for (int n = 0; n < PQ_N; n++) { P->mean[(size_t)n*mx + i] += q[n]/P->denom; P->bulk[n] += q[n]/denom2; }
        }
    }
    double ut = 0.0, Ret = 0.0;
    double *ub = (double*)malloc(sizeof(double)*(mx + 2*s + 2)), *dudx = (double*)malloc(sizeof(double)*(mx + 2*s + 2));
    for (int k = 0; k < mz; k++) for (int j = 0; j < my; j++) {
        const size_t row = (size_t)mx*(j + (size_t)my*k);
        double q0[PQ_N], q1[PQ_N]; post_state(S, row, q0); post_state(S, row + mx - 1, q1);
        const double rw = 0.5*(q0[PQ_P] + q1[PQ_P])/S->Rgas, muw = 1.0/S->P.Re;
        for (int i = 0; i < mx; i++) ub[i+s+1] = S->w[row + i];
        for (int i = 0; i < s+1; i++) { ub[i] = -S->w[row + s - i]; ub[mx+s+1+i] = -S->w[row + mx - i - 1]; }
        for (int jj = 3; jj < mx+s+2; jj++) {
            dudx[jj] = 0;
            for (int i = 0; i < s; i++) {
                /* s = 4: the reference reads ub[-1] at jj = 3 (post.cpp:303-306, outside its array); taken as 0 here */
                const int a = jj+i-s, b = jj-i+s;
                dudx[jj] += S->cF[i]*((a >= 0 ? ub[a] : 0.0) - ub[b])/S->dx;
            }
        }
        double dudxavg = fabs(dudx[3]) + fabs(dudx[4]) + fabs(dudx[mx+s]) + fabs(dudx[mx+s+1]);
        dudxavg = dudxavg*0.25*S->xp[0];
        const double uttemp = sqrt(muw*dudxavg/rw);
        ut += uttemp; Ret += uttemp*rw/muw;
    }
    free(ub); free(dudx);
    P->Ret += Ret/my/mz; P->ut += ut/my/mz;
}
/* after the last file: post.cpp:180-187 (averages over the files, Favre division) */
void ora_post_finish_mean(ora_post *P) {
    P->Ret /= P->nfiles; P->ut /= P->nfiles;
    for (int i = 0; i < P->mx; i++) {
        const double r = P->mean[(size_t)PQ_R*P->mx + i];
        P->mean[(size_t)PQ_UF*P->mx + i] /= r; P->mean[(size_t)PQ_VF*P->mx + i] /= r; P->mean[(size_t)PQ_WF*P->mx + i] /= r;
        P->mean[(size_t)PQ_HF*P->mx + i] /= r;
    }
    P->bulk[PQ_UF] /= P->bulk[PQ_R]; P->bulk[PQ_VF] /= P->bulk[PQ_R]; P->bulk[PQ_WF] /= P->bulk[PQ_R]; P->bulk[PQ_HF] /= P->bulk[PQ_R];
}
/* addFluc (post.cpp:201-223): mean squares about the Reynolds means (r,u,v,w,e,h,t,p,m) and of u,v,w,h about the Favre means */
void ora_post_add_fluc(ora_post *P, ora_solver *S) {
    const int mx = S->mx, my = S->my, mz = S->mz;
    for (int i = 0; i < mx; i++) for (int k = 0; k < mz; k++) for (int j = 0; j < my; j++) {
        double q[PQ_N]; post_state(S, (size_t)i + (size_t)mx*(j + (size_t)my*k), q);
        q[PQ_UF] = q[PQ_U]; q[PQ_VF] = q[PQ_V]; q[PQ_WF] = q[PQ_W]; q[PQ_HF] = q[PQ_H];
        for (int n = 0; n < PQ_N; n++) { const double d = q[n] - P->mean[(size_t)n*mx + i]; P->fluc[(size_t)n*mx + i] += d*d/P->denom; }
    }
}
void ora_post_get(const ora_post *P, double *mean, double *fluc, double *bulk, double *Ret, double *ut) {
    if (mean) memcpy(mean, P->mean, sizeof(double)*PQ_N*P->mx);
    if (fluc) memcpy(fluc, P->fluc, sizeof(double)*PQ_N*P->mx);
    if (bulk) memcpy(bulk, P->bulk, sizeof(double)*PQ_N);
    if (Ret) *Ret = P->Ret; if (ut) *ut = P->ut;
}

/* calcTimeStepPressGrad cuda_main.cu:249-265, calcPressureGrad calc_stress.cu:98-120 */
static void calcTimeStepPressGrad(ora_solver *S) {
    if (!S->fixed_dt) S->dtC = ora_calc_dt(S);
    if (S->P.forcing) {
        double dpdz_prev = S->dpdz;
        double rbulk = volumeIntegral(S, S->r, NULL, 0);
        double a = volumeIntegral(S, S->r, S->w, 0);
        S->dpdz = 0.99*(dpdz_prev) - 0.5*(a/(rbulk)-1);       /* deviceCalcPress calc_stress.cu:16-18 */
    }
}

/* ------------------------------------------------------------------ time integration */
static const double RK_ALPHA[3] = {0., -17./60., -5./12.};   /* cuda_main.cu:10 */
static const double RK_BETA[3]  = {8./15., 5./12., 3./4.};   /* cuda_main.cu:11 */

/* one low-storage stage: cuda_main.cu:126-139 (and :152-165, :171-184) */
static void ls_stage(ora_solver *S, double *ra[5], double *rb[5], int step) {
    const double dt = S->dtC, al = RK_ALPHA[step], be = RK_BETA[step];
    double *var[5] = {S->r, S->u, S->v_, S->w, S->e};
    #pragma omp parallel for schedule(static)
    for (long g = 0; g < (long)S->N; g++) {
        double rold = S->r[g];
        double q[5] = {rold, S->u[g]*rold, S->v_[g]*rold, S->w[g]*rold, S->e[g]};      /* deviceMul */
        for (int c = 0; c < 5; c++) q[c] = q[c] + (dt)*(al*ra[c][g] + be*rb[c][g]);    /* sumLowStorageRK3 :244-247 */
        var[0][g] = q[0]; var[4][g] = q[4];
        var[1][g] = q[1]/q[0]; var[2][g] = q[2]/q[0]; var[3][g] = q[3]/q[0];            /* deviceDiv */
    }
}

static void step_lowstorage(ora_solver *S) {
    ora_calc_rhs(S, S->rhs1); ls_stage(S, S->rhs1, S->rhs1, 0);
    ora_calc_rhs(S, S->rhs2); ls_stage(S, S->rhs1, S->rhs2, 1);
    ora_calc_rhs(S, S->rhs1); ls_stage(S, S->rhs2, S->rhs1, 2);
}

/* runSimulation cuda_main.cu:57-105 (Kutta RK3) */
static void step_kutta(ora_solver *S) {
    const double dt = S->dtC; long N = (long)S->N;
    double *var[5] = {S->r, S->u, S->v_, S->w, S->e};
    #pragma omp parallel for schedule(static)
    for (long g = 0; g < N; g++) {
        S->old[0][g] = S->r[g]; S->old[4][g] = S->e[g];
        S->old[1][g] = S->r[g]*S->u[g]; S->old[2][g] = S->r[g]*S->v_[g]; S->old[3][g] = S->r[g]*S->w[g];
    }
    ora_calc_rhs(S, S->rhs1);
    #pragma omp parallel for schedule(static)
    for (long g = 0; g < N; g++) {
        var[0][g] = (S->old[0][g] + S->rhs1[0][g]*(dt)/2.0);              /* eulerSum :188-191 */
        var[4][g] = (S->old[4][g] + S->rhs1[4][g]*(dt)/2.0);
        for (int c = 1; c < 4; c++) var[c][g] = (S->old[c][g] + S->rhs1[c][g]*(dt)/2.0)/var[0][g];   /* eulerSumR */
    }
    ora_calc_rhs(S, S->rhs2);
    #pragma omp parallel for schedule(static)
    for (long g = 0; g < N; g++) {
        var[0][g] = S->old[0][g] + (2*S->rhs2[0][g] - S->rhs1[0][g])*(dt);   /* eulerSum3 :198-201 */
        var[4][g] = S->old[4][g] + (2*S->rhs2[4][g] - S->rhs1[4][g])*(dt);
        for (int c = 1; c < 4; c++) var[c][g] = (S->old[c][g] + (2*S->rhs2[c][g] - S->rhs1[c][g])*(dt))/var[0][g];
    }
    ora_calc_rhs(S, S->rhs3);
    #pragma omp parallel for schedule(static)
    for (long g = 0; g < N; g++) {
        var[0][g] = S->old[0][g] + (dt)*(S->rhs1[0][g] + 4*S->rhs2[0][g] + S->rhs3[0][g])/6.;   /* rk3final :208-211 */
        var[4][g] = S->old[4][g] + (dt)*(S->rhs1[4][g] + 4*S->rhs2[4][g] + S->rhs3[4][g])/6.;
        for (int c = 1; c < 4; c++) var[c][g] = (S->old[c][g] + (dt)*(S->rhs1[c][g] + 4*S->rhs2[c][g] + S->rhs3[c][g])/6.)/var[0][g];
    }
}

/* classical RK4 -- EXTENSION, not in the reference (README.md:26 only mentions it).
 * q1 = q0 + dt/2 k1 ; q2 = q0 + dt/2 k2 ; q3 = q0 + dt k3 ; q = q0 + dt/6 (k1+2k2+2k3+k4).
 * rhs3 accumulates k1+2k2+2k3. */
static void step_rk4(ora_solver *S) {
    const double dt = S->dtC; long N = (long)S->N;
    double *var[5] = {S->r, S->u, S->v_, S->w, S->e};
    #pragma omp parallel for schedule(static)
    for (long g = 0; g < N; g++) {
        S->old[0][g] = S->r[g]; S->old[4][g] = S->e[g];
        S->old[1][g] = S->r[g]*S->u[g]; S->old[2][g] = S->r[g]*S->v_[g]; S->old[3][g] = S->r[g]*S->w[g];
    }
    const double a[4] = {0.5, 0.5, 1.0, 0.0}, b[4] = {1.0, 2.0, 2.0, 1.0};
    for (int st = 0; st < 4; st++) {
        ora_calc_rhs(S, S->rhs1);
        #pragma omp parallel for schedule(static)
        for (long g = 0; g < N; g++) {
            double q[5];
            for (int c = 0; c < 5; c++) {
                double acc = (st == 0 ? 0.0 : S->rhs3[c][g]) + b[st]*S->rhs1[c][g];
                S->rhs3[c][g] = acc;
                q[c] = (st < 3) ? S->old[c][g] + (a[st]*dt)*S->rhs1[c][g] : S->old[c][g] + (dt/6.0)*acc;
            }
            var[0][g] = q[0]; var[4][g] = q[4];
            var[1][g] = q[1]/q[0]; var[2][g] = q[2]/q[0]; var[3][g] = q[3]/q[0];
        }
    }
}

/* runSimulationLowStorage / runSimulation step loop: cuda_main.cu:50-55 / :115-120 */
void ora_run(ora_solver *S, int nsteps, double *time, double *par1, double *par2) {
    for (int istep = 0; istep < nsteps; istep++) {
        if (istep % S->P.checkCFLcondition == 0) calcTimeStepPressGrad(S);
        S->time_last = S->time_last + S->dtC;      /* deviceSumOne chain; time starts at 0 (Q6: garbage in the reference) */
        if (time) time[istep] = S->time_last;
        S->time_on_GPU += S->dtC;                  /* deviceAdvanceTime calc_stress.cu:12-14 */
        if (istep % S->P.checkBulk == 0) {
            double p1 = 0.0, p2 = 0.0;
            ora_calc_bulk(S, &p1, &p2);
            if (par1) par1[istep] = p1;
            if (par2 && S->P.forcing) par2[istep] = p2;
        }
        if (S->P.rk4) step_rk4(S);
        else if (S->P.lowStorage) step_lowstorage(S);
        else step_kutta(S);
    }
}

/* ------------------------------------------------------------------ known-answer entry points */
static void kat_setup(pen_ops *o, int s, double invd) {
    memset(o, 0, sizeof(*o)); o->s = s; o->v = s; o->cF = CF_TAB[s]; o->cS = CS_TAB[s]; o->cVF = CF_TAB[s]; o->cVS = CS_TAB[s];
    o->d1 = invd; o->d2 = invd; o->dir = 1; o->nonuni = 0;
}
static double *kat_pad(int s, int n, const double *f) {
    double *b = (double*)malloc(sizeof(double)*(n+2*s));
    for (int i = 0; i < n; i++) b[i+s] = f[i];
    for (int g = 0; g < s; g++) perBC(b, g, s, n);
    return b;
}
void ora_kat_flux_cube(int s, int n, double invd, const double *f, const double *g, const double *h, double *out) {
    pen_ops o; kat_setup(&o, s, invd); double *a = kat_pad(s,n,f), *b = kat_pad(s,n,g), *c = kat_pad(s,n,h);
    for (int i = 0; i < n; i++) out[i] = fluxCube(&o, a, b, c, i+s);
    free(a); free(b); free(c);
}
void ora_kat_flux_quad(int s, int n, double invd, const double *f, const double *g, double *out) {
    pen_ops o; kat_setup(&o, s, invd); double *a = kat_pad(s,n,f), *b = kat_pad(s,n,g);
    for (int i = 0; i < n; i++) out[i] = fluxQuad(&o, a, b, i+s);
    free(a); free(b);
}
void ora_kat_d1(int s, int n, double invd, const double *f, double *out) {
    pen_ops o; kat_setup(&o, s, invd); double *a = kat_pad(s,n,f);
    for (int i = 0; i < n; i++) out[i] = der1A(&o, a, i+s);
    free(a);
}
void ora_kat_d2(int s, int n, double invd2, const double *f, double *out) {
    pen_ops o; kat_setup(&o, s, invd2); double *a = kat_pad(s,n,f);
    for (int i = 0; i < n; i++) out[i] = der2V(&o, a, i+s);
    free(a);
}
