// Host-side set-up helpers of libcudns (no CUDA): grid + metrics, initial conditions, sponge tables,
// fields/ I/O.  These restate src/init.cpp, the host half of src/sponge.cu and src/comm.cpp's file
// format so that a caller of the reference finds the same operator surface.  Paths in comments are
// relative to the reference repository.
#include "cudns_internal.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace cudns_shared {
static thread_local std::string g_err;
void set_error(const std::string &m) { g_err = m; }

// globals.h:69-95, stored far-to-near like the reference
static const double kF1[] = {-1.0 / 2.0};
static const double kS1[] = {1.0, -2.0};
static const double kF2[] = {1.0 / 12.0, -2.0 / 3.0};
static const double kS2[] = {-1.0 / 12.0, 4.0 / 3.0, -5.0 / 2.0};
static const double kF3[] = {-1.0 / 60.0, 3.0 / 20.0, -3.0 / 4.0};
static const double kS3[] = {1.0 / 90.0, -3.0 / 20.0, 3.0 / 2.0, -49.0 / 18.0};
static const double kF4[] = {1.0 / 280.0, -4.0 / 105.0, 1.0 / 5.0, -4.0 / 5.0};
static const double kS4[] = {-1.0 / 560.0, 8.0 / 315.0, -1.0 / 5.0, 8.0 / 5.0, -205.0 / 72.0};
const double *coeff_first(int s) { static const double *t[] = {nullptr, kF1, kF2, kF3, kF4}; return t[s]; }
const double *coeff_second(int s) { static const double *t[] = {nullptr, kS1, kS2, kS3, kS4}; return t[s]; }

int check_params(const cudns_params *p) {
    if (!p) { set_error("params is NULL"); return CUDNS_EINVAL; }
    if (p->mx < 2 || p->my < 2 || p->mz < 2) { set_error("grid extents must be >= 2"); return CUDNS_EINVAL; }
    if (p->stencilSize < 1 || p->stencilSize > 4) { set_error("stencilSize must be in 1..4"); return CUDNS_EINVAL; }
    if (p->stencilVisc < 1 || p->stencilVisc > p->stencilSize) {
        set_error("stencilVisc must satisfy 1 <= stencilVisc <= stencilSize (globals.h:16)"); return CUDNS_EINVAL; }
    if (p->nranks < 1 || p->rank < 0 || p->rank >= p->nranks) { set_error("bad rank/nranks"); return CUDNS_EINVAL; }
    if (p->mz % p->nranks) { set_error("mz must be divisible by nranks"); return CUDNS_EINVAL; }
    if (p->mz / p->nranks < p->stencilSize + p->stencilVisc) { set_error("z-slab thinner than the halo depth s+v"); return CUDNS_EINVAL; }
    if (p->my < p->stencilSize || p->mx < p->stencilSize + 1) { set_error("grid smaller than the stencil"); return CUDNS_EINVAL; }
    if (p->mx % 2) { set_error("mx must be even (16-byte row alignment; the reference needs mx % sPencils == 0)"); return CUDNS_EINVAL; }
    if (!(p->Re > 0) || !(p->Ma > 0) || !(p->Pr > 0) || !(p->gam > 1)) { set_error("Re, Ma, Pr must be > 0 and gam > 1"); return CUDNS_EINVAL; }
    if (p->checkCFLcondition < 1 || p->checkBulk < 1) { set_error("checkCFLcondition / checkBulk must be >= 1"); return CUDNS_EINVAL; }
    if (p->precision != 0 && p->precision != 1) { set_error("precision must be 0 (double) or 1 (float)"); return CUDNS_EINVAL; }
    if (p->boundaryLayer && p->periodicX) { set_error("boundaryLayer requires periodicX = 0"); return CUDNS_EINVAL; }
    return CUDNS_OK;
}
}  // namespace cudns_shared

using namespace cudns_shared;

extern "C" {

const char *cudns_last_error(void) { return cudns::g_err.c_str(); }
const char *cudns_version(void) { return "cudns-b200 0.1 (sm_100a)"; }

static void params_common(cudns_params *p) {
    std::memset(p, 0, sizeof(*p));
    p->lowStorage = 1; p->quirk_q1 = 1; p->nranks = 1;
    p->gam = 1.4; p->stretch = 5.0; p->TwallTop = 1.0; p->TwallBot = 1.0;
    p->spTopStr = 1.0; p->spTopLen = 1.0; p->spTopExp = 2.0;     // sponge.h:5-17
    p->spInlStr = 0.5; p->spInlLen = 20.0; p->spInlExp = 2.0;
    p->spOutStr = 0.5; p->spOutLen = 20.0; p->spOutExp = 2.0;
    p->kC = 110; p->LP = 40; p->amp1 = 2e-4; p->amp2 = 3e-5; p->omega2 = 10.0;   // perturbation.h:16-21
}

int cudns_params_tgv(cudns_params *p, int n, int stencil) {   // python-utils/CompNavierStokes.py:1-7
    if (!p) return CUDNS_EINVAL;
    params_common(p);
    p->mx = p->my = p->mz = n; p->stencilSize = stencil; p->stencilVisc = stencil;
    p->Lx = p->Ly = p->Lz = 2.0 * M_PI;
    p->CFL = 0.5; p->periodicX = 1;
    p->checkCFLcondition = 10; p->checkBulk = 10;
    p->Re = 1600.0; p->Pr = 1.0; p->Ma = 0.1; p->viscexp = 1.0;
    p->omega1 = p->Re * 121.e-6;
    return CUDNS_OK;
}
int cudns_params_channel(cudns_params *p) {   // globals/channel.h:17-52
    if (!p) return CUDNS_EINVAL;
    params_common(p);
    p->mx = 160; p->my = 192; p->mz = 192; p->stencilSize = 3; p->stencilVisc = 2;
    p->Lx = 2.0; p->Ly = 2.0 * M_PI; p->Lz = 4.0 * M_PI;
    p->CFL = (double)0.75f; p->forcing = 1; p->nonUniformX = 1;
    p->checkCFLcondition = 100; p->checkBulk = 100;
    p->Re = 2800.0; p->Pr = 0.75; p->Ma = 1.5; p->viscexp = 0.75; p->stretch = 3.0;
    p->omega1 = p->Re * 121.e-6;
    return CUDNS_OK;
}
int cudns_params_blayer(cudns_params *p) {    // src/globals.h:17-52
    if (!p) return CUDNS_EINVAL;
    params_common(p);
    p->mx = 240; p->my = 64; p->mz = 2048; p->stencilSize = 3; p->stencilVisc = 2;
    p->Lx = 20.0; p->Ly = 7.0; p->Lz = 500.0;
    p->CFL = (double)0.75f; p->boundaryLayer = 1; p->perturbed = 1; p->nonUniformX = 1;
    p->checkCFLcondition = 100; p->checkBulk = 100;
    p->Re = 1500.0; p->Pr = 0.75; p->Ma = 0.35; p->viscexp = 1.5; p->stretch = 5.0;
    p->omega1 = p->Re * 121.e-6;
    return CUDNS_OK;
}

// initGrid + derivGrid, init.cpp:32-91,258-277
int cudns_init_grid(const cudns_params *p, double *x, double *xp, double *xpp, double *y, double *z, double *dx_out) {
    int rc = check_params(p); if (rc) return rc;
    if (!x || !xp || !xpp || !y || !z) { set_error("NULL grid array"); return CUDNS_EINVAL; }
    const int mx = p->mx, s = p->stencilSize;
    const double *cF = coeff_first(s), *cS = coeff_second(s);
    double dx = p->Lx * (1.0) / (mx);
    std::vector<double> xn(mx + 1);
    int denom = mx; double denom2 = 2.0;
    if (p->boundaryLayer) { denom *= 2; denom2 /= 2; }     // one-sided clustering
    for (int i = 0; i < mx + 1; i++) xn[i] = std::tanh(p->stretch * ((i * 1.0) / denom - 0.5)) / std::tanh(p->stretch * 0.5);
    for (int i = 0; i < mx; i++) x[i] = p->Lx * (1.0 + (xn[i] + xn[i + 1]) / 2.0) / denom2;
    {   // metric derivatives with odd / (2Lx - .) reflections
        std::vector<double> fb(mx + 2 * s);
        for (int i = s; i < mx + s; i++) fb[i] = x[i - s];
        for (int i = 0; i < s; i++) { fb[i] = -fb[2 * s - i - 1]; fb[mx + s + i] = 2 * p->Lx - fb[mx + s - i - 1]; }
        for (int i = 0; i < mx; i++) {
            xp[i] = 0.0; xpp[i] = cS[s] * fb[i + s] / dx / dx;
            for (int it = 0; it < s; it++) {
                xp[i] += cF[it] * (fb[i + it] - fb[i + s * 2 - it]) / dx;
                xpp[i] += cS[it] * (fb[i + it] + fb[i + s * 2 - it]) / dx / dx;
            }
        }
    }
    for (int i = 0; i < mx; i++) xp[i] = 1.0 / xp[i];
    if (!p->nonUniformX) {
        for (int i = 0; i < mx; i++) x[i] = p->Lx * (0.5 + i * 1.0) / (mx);
        dx = x[1] - x[0];
    }
    for (int j = 0; j < p->my; j++) y[j] = p->Ly * (0.5 + j * 1.0) / (p->my);
    for (int k = 0; k < p->mz; k++) z[k] = p->Lz * (0.5 + k * 1.0) / (p->mz);
    if (dx_out) *dx_out = dx;
    return CUDNS_OK;
}

#define GIDX(i, j, k) ((size_t)(k) * mx * my + (size_t)(j) * mx + (size_t)(i))

// initCHIT, init.cpp:126-148 (Taylor-Green vortex)
int cudns_init_chit(const cudns_params *p, const double *x, const double *y, const double *z,
                    double *r, double *u, double *v, double *w, double *e) {
    int rc = check_params(p); if (rc) return rc;
    const int mx = p->mx, my = p->my, mz = p->mz;
    const double Rgas = (1.f / (p->gam * p->Ma * p->Ma));
    const double V0 = 1.0, T0 = 1.0, P0 = T0 * Rgas, R0 = 1.0;
    for (int i = 0; i < mx; i++) {
        double fx = 2 * M_PI * x[i] / p->Lx;
        for (int j = 0; j < my; j++) {
            double fy = 2 * M_PI * y[j] / p->Ly;
            for (int k = 0; k < mz; k++) {
                double fz = 2 * M_PI * z[k] / p->Lz;
                size_t g = GIDX(i, j, k);
                u[g] = V0 * std::sin(fx / 1.0) * std::cos(fy / 1.0) * std::cos(fz / 1.0);
                v[g] = -V0 * std::cos(fx / 1.0) * std::sin(fy / 1.0) * std::cos(fz / 1.0);
                w[g] = 0.0;
                double press = P0 + 1.0 / 16.0 * R0 * V0 * V0 * (std::cos(2.0 * fx / 1.0) + std::cos(2.0 * fy / 1.0)) * (std::cos(2.0 * fz / 1.0) + 2.0);
                r[g] = press / Rgas / T0;
                e[g] = press / (p->gam - 1.0) + 0.5 * r[g] * (std::pow(u[g], 2) + std::pow(v[g], 2) + std::pow(w[g], 2));
            }
        }
    }
    return CUDNS_OK;
}

// initChannel, init.cpp:94-124.  The reference never seeds rand(): glibc's default sequence (srand(1)).
int cudns_init_channel(const cudns_params *p, const double *x, const double *y, const double *z,
                       double *r, double *u, double *v, double *w, double *e) {
    int rc = check_params(p); if (rc) return rc;
    (void)z;
    const int mx = p->mx, my = p->my, mz = p->mz;
    const double Rgas = (1.f / (p->gam * p->Ma * p->Ma));
    const double T0 = 1.0, P0 = T0 * Rgas, R0 = 1.0;
    const double U0 = std::pow(p->gam, 0.5) * p->Ma;
    srand(1);
    for (int i = 0; i < mx; i++)
        for (int j = 0; j < my; j++)
            for (int k = 0; k < mz; k++) {
                double rr1 = rand() * 1.0 / (RAND_MAX * 1.0) - 0.5;
                double rr2 = rand() * 1.0 / (RAND_MAX * 1.0) - 0.5;
                double rr3 = rand() * 1.0 / (RAND_MAX * 1.0) - 0.5;
                double ufluc = 0.02 * rr1, vfluc = 0.02 * rr2, wfluc = 0.02 * rr3;
                double wmean = 1.5 * U0 * R0 * x[i] * (1.0 - x[i] / p->Lx);
                ufluc = ufluc + 0.05 * std::sin(0.5 * M_PI * x[i]) * std::cos(2 * M_PI * y[j]);
                vfluc = vfluc + 0.05 * std::sin(0.5 * M_PI * x[i]) * std::sin(2 * M_PI * y[j]);
                size_t g = GIDX(i, j, k);
                u[g] = ufluc; v[g] = vfluc; w[g] = wmean + wfluc;
                r[g] = R0;
                e[g] = P0 / (p->gam - 1.0) + 0.5 * r[g] * (std::pow(u[g], 2) + std::pow(v[g], 2) + std::pow(w[g], 2));
            }
    return CUDNS_OK;
}

// natural cubic spline (0-based).  The reference calls 1-based Numerical-Recipes routines on 0-based
// arrays (sponge.cu:156-163,262-313); callers that want its exact tables pass the
// profile arrays shifted by one entry.
static void spline0(const double *x, const double *y, int n, std::vector<double> &y2) {
    std::vector<double> u(n);
    y2.assign(n, 0.0);
    for (int i = 1; i <= n - 2; i++) {
        double sig = (x[i] - x[i - 1]) / (x[i + 1] - x[i - 1]);
        double pp = sig * y2[i - 1] + 2.0;
        y2[i] = (sig - 1.0) / pp;
        u[i] = (y[i + 1] - y[i]) / (x[i + 1] - x[i]) - (y[i] - y[i - 1]) / (x[i] - x[i - 1]);
        u[i] = (6.0 * u[i] / (x[i + 1] - x[i - 1]) - sig * u[i - 1]) / pp;
    }
    y2[n - 1] = 0.0;
    for (int k = n - 2; k >= 0; k--) y2[k] = y2[k] * y2[k + 1] + u[k];
}
static double splint0(const double *xa, const double *ya, const std::vector<double> &y2a, int n, double x) {
    int klo = 0, khi = n - 1;
    while (khi - klo > 1) { int k = (khi + klo) >> 1; if (xa[k] > x) khi = k; else klo = k; }
    double h = xa[khi] - xa[klo];
    double a = (xa[khi] - x) / h, b = (x - xa[klo]) / h;
    return a * ya[klo] + b * ya[khi] + ((a * a * a - a) * y2a[klo] + (b * b * b - b) * y2a[khi]) * (h * h) / 6.0;
}

// calculateSponge host half, sponge.cu:115-129 (strengths), :160-169 (reference state), :185-195 (IC),
// copySpongeToDevice :58-62 (conservative references).
int cudns_build_sponge(const cudns_params *p, const double *x, const double *z,
                       const double *xIn, const double *rIn, const double *uIn, const double *wIn, int n,
                       double *sigma_x, double *sigma_z, double *ref5,
                       double *r, double *u, double *v, double *w, double *e) {
    int rc = check_params(p); if (rc) return rc;
    if (n < 3 || !xIn || !rIn || !uIn || !wIn || !sigma_x || !sigma_z || !ref5) { set_error("bad sponge inputs"); return CUDNS_EINVAL; }
    const int mx = p->mx, my = p->my, mz = p->mz;
    const double Rgas = (1.f / (p->gam * p->Ma * p->Ma));
    for (int i = 0; i < mx; i++) {
        sigma_x[i] = 0.0;
        if ((p->spTopLen > 0.0) && (x[i] >= p->Lx - p->spTopLen))
            sigma_x[i] = p->spTopStr * std::pow((x[i] - (p->Lx - p->spTopLen)) / p->spTopLen, p->spTopExp);
    }
    for (int k = 0; k < mz; k++) {
        sigma_z[k] = 0.0;
        double fz = z[k];
        if ((p->spInlLen > 0.0) && (fz <= p->spInlLen)) sigma_z[k] = p->spInlStr * std::pow((p->spInlLen - fz) / p->spInlLen, p->spInlExp);
        if ((p->spOutLen > 0.0) && (fz >= (p->Lz - p->spOutLen))) sigma_z[k] = p->spOutStr * std::pow((fz - (p->Lz - p->spOutLen)) / p->spOutLen, p->spOutExp);
    }
    std::vector<double> r2, u2, w2;
    spline0(xIn, rIn, n, r2); spline0(xIn, uIn, n, u2); spline0(xIn, wIn, n, w2);
    const size_t nq = (size_t)mx * mz;
    for (int k = 0; k < mz; k++)
        for (int i = 0; i < mx; i++) {
            double scale = std::pow(1 + z[k] / p->Re, 0.5);
            double rr = splint0(xIn, rIn, r2, n, x[i] / scale);
            double uu = splint0(xIn, uIn, u2, n, x[i] / scale);
            uu /= (scale * p->Re);
            double ww = splint0(xIn, wIn, w2, n, x[i] / scale);
            double ee = Rgas / (p->gam - 1.0) + rr * 0.5 * (uu * uu + ww * ww);
            size_t q = (size_t)i + (size_t)k * mx;
            ref5[q] = rr; ref5[nq + q] = uu * rr; ref5[2 * nq + q] = 0.0; ref5[3 * nq + q] = ww * rr; ref5[4 * nq + q] = ee;
            if (r && u && v && w && e)
                for (int j = 0; j < my; j++) {
                    size_t g = GIDX(i, j, k);
                    r[g] = rr; u[g] = uu; v[g] = 0.0; w[g] = ww; e[g] = ee;
                }
        }
    return CUDNS_OK;
}

// saveFileMPI / readFileMPI file format, comm.cpp:205-279: "fields/%c.%07d.bin", raw float64
int cudns_write_field(const char *dir, char name, int timestep, const double *var, size_t count) {
    char path[1024];
    std::snprintf(path, sizeof(path), "%s/fields/%c.%07d.bin", dir ? dir : ".", name, timestep);
    FILE *f = std::fopen(path, "wb");
    if (!f) { set_error(std::string("cannot open ") + path); return CUDNS_EINVAL; }
    size_t n = std::fwrite(var, sizeof(double), count, f);
    std::fclose(f);
    if (n != count) { set_error(std::string("short write ") + path); return CUDNS_EINVAL; }
    return CUDNS_OK;
}
int cudns_read_field(const char *dir, char name, int timestep, double *var, size_t count) {
    char path[1024];
    std::snprintf(path, sizeof(path), "%s/fields/%c.%07d.bin", dir ? dir : ".", name, timestep);
    FILE *f = std::fopen(path, "rb");
    if (!f) { set_error(std::string("cannot open ") + path); return CUDNS_EINVAL; }
    size_t n = std::fread(var, sizeof(double), count, f);
    std::fclose(f);
    if (n != count) { set_error(std::string("short read ") + path); return CUDNS_EINVAL; }
    return CUDNS_OK;
}


// Compressible self-similar boundary layer (python-utils/selfSimilarSol.py:1-98): the inflow / sponge reference profiles of the
// boundary-layer case without Python.  Same equations, boundary conditions, eta range [0,10] on 600 points and post-processing
// as the reference script (which solves the boundary-value problem with scipy's collocation solver at its default 1e-3
// tolerance); here: shooting on (F''(0), h(0)) with classical RK4 (6000 steps) and Newton, converged to 1e-13.
//   F0' = F1, F1' = F2/C, F2' = -F0 F2/C, h' = G1 Pr/C, G1' = -(F0 G1 Pr + Ec F2^2)/C, y' = sqrt(2) h,   C = h^(expMu-1)
//   F0(0) = F1(0) = G1(0) = y(0) = 0 (adiabatic wall), F1(10) = h(10) = 1.
// Outputs (n entries each, the reference writes n = 1000): x = wall distance / delta_99, r = density, u = wall-normal velocity,
// w = streamwise velocity, e = internal energy T Rgas/(gam-1); beyond the 600 computed points the free stream continues with
// steps of 0.15.  Like the reference script the viscosity exponent is fixed at 3/2 whatever the caller's viscexp is.
int cudns_blasius_profiles(double gam, double Ma, double Pr, int n, double *x, double *r, double *u, double *w, double *e) {
    if (!x || !r || !u || !w || !e || n < 600) { set_error("cudns_blasius_profiles: NULL output or n < 600"); return CUDNS_EINVAL; }
    const double expMu = 1.5, Ec = (gam - 1.0) * Ma * Ma, Rgas = 1.0 / (gam * Ma * Ma);
    const int NP = 600, SUB = 10, NS = (NP - 1) * SUB;
    const double h = 10.0 / NS;
    auto rhs = [&](const double *f, double *d) {
        const double C = std::pow(f[3], expMu - 1.0);
        d[0] = f[1]; d[1] = f[2] / C; d[2] = -f[0] * f[2] / C; d[3] = f[4] * Pr / C;
        d[4] = -f[0] * f[4] * Pr / C - Ec * f[2] * f[2] / C; d[5] = std::sqrt(2.0) * f[3];
    };
    std::vector<double> sol((size_t)NP * 6);
    auto shoot = [&](double a, double b, double *res) {
        double f[6] = {0.0, 0.0, a, b, 0.0, 0.0};
        for (int m = 0; m < 6; m++) sol[m] = f[m];
        for (int st = 0; st < NS; st++) {
            double k1[6], k2[6], k3[6], k4[6], t[6];
            rhs(f, k1);
            for (int m = 0; m < 6; m++) t[m] = f[m] + 0.5 * h * k1[m];
            rhs(t, k2);
            for (int m = 0; m < 6; m++) t[m] = f[m] + 0.5 * h * k2[m];
            rhs(t, k3);
            for (int m = 0; m < 6; m++) t[m] = f[m] + h * k3[m];
            rhs(t, k4);
            for (int m = 0; m < 6; m++) f[m] += h / 6.0 * (k1[m] + 2.0 * k2[m] + 2.0 * k3[m] + k4[m]);
            if ((st + 1) % SUB == 0) for (int m = 0; m < 6; m++) sol[(size_t)((st + 1) / SUB) * 6 + m] = f[m];
        }
        res[0] = f[1] - 1.0; res[1] = f[3] - 1.0;
    };
    double a = 0.47, b = 1.0 + 0.5 * std::sqrt(Pr) * Ec, res[2];
    bool ok = false;
    for (int it = 0; it < 50; it++) {
        shoot(a, b, res);
        if (std::fabs(res[0]) < 1e-13 && std::fabs(res[1]) < 1e-13) { ok = true; break; }
        const double da = 1e-7, db = 1e-7;
        double ra[2], rb[2];
        shoot(a + da, b, ra); shoot(a, b + db, rb);
        const double J00 = (ra[0] - res[0]) / da, J10 = (ra[1] - res[1]) / da, J01 = (rb[0] - res[0]) / db, J11 = (rb[1] - res[1]) / db;
        const double det = J00 * J11 - J01 * J10;
        if (!(std::fabs(det) > 1e-300)) break;
        a -= (J11 * res[0] - J01 * res[1]) / det;
        b -= (-J10 * res[0] + J00 * res[1]) / det;
    }
    if (!ok) { shoot(a, b, res); ok = std::fabs(res[0]) < 1e-10 && std::fabs(res[1]) < 1e-10; }
    if (!ok) { set_error("cudns_blasius_profiles: shooting did not converge"); return CUDNS_EINVAL; }
    shoot(a, b, res);
    int idx = -1;
    for (int i = 0; i < NP; i++) if (sol[(size_t)i * 6 + 1] > 0.99) { idx = i; break; }
    if (idx < 0) { set_error("cudns_blasius_profiles: no point with U > 0.99"); return CUDNS_EINVAL; }
    const double delta = sol[(size_t)idx * 6 + 5];
    for (int i = 0; i < NP; i++) {
        const double F0 = sol[(size_t)i * 6], U = sol[(size_t)i * 6 + 1], T = sol[(size_t)i * 6 + 3], yb = sol[(size_t)i * 6 + 5];
        const double rho = 1.0 / T;
        x[i] = yb / delta; r[i] = rho; w[i] = U;
        u[i] = (U * yb / std::sqrt(4.0) - F0 / (rho * std::sqrt(2.0))) / delta;
        e[i] = T * Rgas / (gam - 1.0);
    }
    for (int i = NP; i < n; i++) { x[i] = x[i - 1] + 0.15; r[i] = r[NP - 1]; u[i] = u[NP - 1]; w[i] = w[NP - 1]; e[i] = e[NP - 1]; }
    return CUDNS_OK;
}

// XDMF 2.0 sidecar for the fields/ directory (what python-utils/writexmf.py + makexmf.py produce): a 3DRectMesh with the VXVYVZ
// coordinates inline, one temporal collection, one uniform grid per saved time step whose attributes point at <name>.<%07d>.bin.
// names: one character per field ("ruvwe"); time of step t = t * dt.
int cudns_write_xdmf(const char *path, int single_precision, const double *x, int nx, const double *y, int ny, const double *z, int nz,
                     const int *timesteps, int nt, double dt, const char *names) {
    if (!path || !x || !y || !z || (!timesteps && nt > 0) || !names) { set_error("NULL argument"); return CUDNS_EINVAL; }
    FILE *f = std::fopen(path, "wt");
    if (!f) { set_error(std::string("cannot open ") + path); return CUDNS_EINVAL; }
    const int prec = single_precision ? 4 : 8;
    auto axis = [&](const char *indent, const double *v, int n) {
        std::fprintf(f, "%s<DataItem Format=\"XML\" DataType=\"Float\" Precision=\"%d\" Dimensions=\"%5d\">\n", indent, prec, n);
        for (int i = 0; i < n; i++) std::fprintf(f, "%15.6E", v[i]);
        std::fprintf(f, "\n%s</DataItem>\n", indent);
    };
    std::fprintf(f, "<?xml version=\"1.0\" ?>\n<!DOCTYPE Xdmf SYSTEM \"Xdmf.dtd\" []>\n");
    std::fprintf(f, "<Xdmf xmlns:xi=\"http://www.w3.org/2001/XInclude\" Version=\"2.0\">\n<Domain>\n");
    std::fprintf(f, "    <Topology name=\"TOPO\" TopologyType=\"3DRectMesh\" Dimensions=\"%5d%5d%5d\"/>\n", nz, ny, nx);
    std::fprintf(f, "    <Geometry name=\"GEO\" GeometryType=\"VXVYVZ\">\n");
    axis("        ", x, nx); axis("        ", y, ny); axis("        ", z, nz);
    std::fprintf(f, "    </Geometry>\n");
    std::fprintf(f, "    <Grid Name=\"TimeSeries\" GridType=\"Collection\" CollectionType=\"Temporal\">\n        <Time TimeType=\"List\">\n");
    {
        std::vector<double> tv(nt);
        for (int i = 0; i < nt; i++) tv[i] = timesteps[i] * dt;
        axis("            ", tv.data(), nt);
    }
    std::fprintf(f, "        </Time>\n");
    for (int i = 0; i < nt; i++) {
        std::fprintf(f, "        <Grid Name=\"T%07d\" GridType=\"Uniform\">\n", timesteps[i]);
        std::fprintf(f, "            <Topology Reference=\"/Xdmf/Domain/Topology[1]\"/>\n            <Geometry Reference=\"/Xdmf/Domain/Geometry[1]\"/>\n");
        for (const char *c = names; *c; c++) {
            std::fprintf(f, "            <Attribute Name=\"%c\" Center=\"Node\">\n", *c);
            std::fprintf(f, "                <DataItem Format=\"Binary\" DataType=\"Float\" Precision=\"%d\" Endian=\"Native\" Dimensions=\"%5d%5d%5d\">\n", prec, nz, ny, nx);
            std::fprintf(f, "                    %c.%07d.bin\n                </DataItem>\n            </Attribute>\n", *c, timesteps[i]);
        }
        std::fprintf(f, "        </Grid>\n");
    }
    std::fprintf(f, "    </Grid>\n</Domain>\n</Xdmf>\n");
    if (std::fclose(f) != 0) { set_error(std::string("write error ") + path); return CUDNS_EINVAL; }
    return CUDNS_OK;
}

// Variables::printFile (post.cpp:61-86) for mean.txt, fluc.txt and bulk.txt: header with the friction Reynolds number and velocity,
// the legend (the reference's own numbering, "13)" twice), then one %le row per wall-normal index: x and the 13 quantities
static int write_stats_file(const std::string &path, int n, const double *x, const double *q, int stride, double ret, double ut) {
    FILE *fp = std::fopen(path.c_str(), "w+");
    if (!fp) { set_error("cannot open " + path); return CUDNS_EINVAL; }
    std::fprintf(fp, "Reynolds number based on utau %lf with utau %lf\n", ret, ut);
    static const char *legend[] = {"1)  y", "2)  rho", "3)  uFavre", "4)  vFavre", "5)  wFavre", "6)  u", "7)  v", "8)  w", "9)  eTotal", "10) hFavre",
                                   "11) h", "12) Temperature", "13) Pressure", "13) visc"};
    for (const char *l : legend) std::fprintf(fp, "%s\n", l);
    for (int i = 0; i < 147; i++) std::fputc('-', fp);
    std::fputc('\n', fp);
    for (int i = 0; i < n; i++) {
        std::fprintf(fp, "%le", x[i]);
        for (int m = 0; m < 13; m++) std::fprintf(fp, "\t%le", q[(size_t)m * stride + i]);
        std::fputc('\n', fp);
    }
    if (std::fclose(fp) != 0) { set_error("write error " + path); return CUDNS_EINVAL; }
    return CUDNS_OK;
}
int cudns_stats_write(const char *outdir, int mx, const double *x, const double *mean, const double *fluc, const double *bulk, double retau, double utau) {
    if (!x || !mean || !fluc || !bulk || mx < 1) { set_error("cudns_stats_write: bad argument"); return CUDNS_EINVAL; }
    const std::string d = outdir && *outdir ? std::string(outdir) + "/" : std::string();
    int rc = write_stats_file(d + "mean.txt", mx, x, mean, mx, retau, utau); if (rc) return rc;
    rc = write_stats_file(d + "fluc.txt", mx, x, fluc, mx, retau, utau); if (rc) return rc;
    return write_stats_file(d + "bulk.txt", 1, x, bulk, 1, retau, utau);
}

}  // extern "C"
