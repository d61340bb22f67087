// libcudns device code: theta (dilatation) pass, fused RHS + Runge-Kutta stage kernel, ghost handling,
// reductions.  Written from scratch for sm_100a; the numerics restate the reference
// (simone-silvestri/CudaNavierStokes) -- citations are paths relative to that repository.
//
// Data layout: every state field is a padded ghost-cell array [pz][py][px] (Layout in cudns_internal.h).
// The RHS kernel marches along z: a ring of 2s+1 planes of derived quantities for the z stencils and one
// extended (x/y halo) plane for the in-plane stencils live in shared memory; one thread owns one (i,j)
// column.  The split-form convective terms are evaluated in telescoped pair form
//     -1/(4 dx) * sum_l a_l [ P(i,i+l) - P(i,i-l) ],  P = (rho_i+rho_j)(U_i+U_j)(phi_i+phi_j)
// (cuda_derivs.h:30-155 evaluates the same sum as a difference of two interface fluxes).
#include "cudns_internal.h"
#include <cstdio>
#include <cstdlib>
#include <string>

namespace cudns {

constexpr int TX = 32;
constexpr int TY = 8;
constexpr int NTHREADS = TX * TY;
constexpr int CX = TX + 2 * GX;   // extended-plane row pitch

__device__ __forceinline__ double visc_of(const KConst &c, double t) {
    // mu = T^viscexp / Re   (cuda_main.cu:237-238); common exponents avoid the generic pow
    switch (c.viscmode) {
        case 1: return t * c.invRe;
        case 2: return sqrt(t) * c.invRe;
        case 3: { double s = sqrt(t); return s * sqrt(s) * c.invRe; }
        case 4: return t * sqrt(t) * c.invRe;
        default: return pow(t, c.viscexp) * c.invRe;
    }
}

// calcState, cuda_main.cu:218-242
__device__ __forceinline__ void eos(const KConst &c, double r, double u, double v, double w, double e,
                                    double &h, double &p, double &t, double &m) {
    double rinv = 1.0 / r;
    double en = e * rinv - 0.5 * (u * u + v * v + w * w);
    t = c.cvInv * en;
    p = r * c.Rgas * t;
    h = (e + p) * rinv;
    m = visc_of(c, t);
}

// wall blowing/suction, perturbation.h:25-53.  Returns true and the ghost value of u when (j,kglob) lies in the strip.
__device__ __forceinline__ bool perturb_value(const KConst &c, int j, int kglob, double &val) {
    int kSt = c.kC - c.LP / 2, kEn = c.kC + c.LP / 2;
    if (kglob < kSt || kglob > kEn) return false;
    int alpha, beta, kappa;
    if (kglob < c.kC) { kappa = 1; alpha = kglob - kSt; beta = c.kC - kSt; }
    else              { kappa = -1; alpha = kEn - kglob; beta = kEn - c.kC; }
    double ksi = alpha * 1.0 / beta;
    double g = (15.1875 * ksi * ksi * ksi * ksi * ksi) - (35.4375 * ksi * ksi * ksi * ksi) + (20.25 * ksi * ksi * ksi);
    double y_glob = (double)j / c.d1[1];
    double tg = *c.time_on_gpu;
    val = c.amp1 * kappa * g * sin(c.omega1 * tg) + c.amp2 * kappa * g * sin(c.omega2 * tg) * cos(y_glob / c.lambdaP);
    return true;
}

// ---------------------------------------------------------------------------------------------
// theta = du/dx + dv/dy + dw/dz at viscous order (derVelX/Y/Z + calcDil, calc_stress.cu:20-96).
// One thread per point, neighbours straight from global memory (the L1/L2 absorb the re-reads; this
// pass moves 32 B/pt against ~200 B/pt of the stage kernel).  Computed for k in [-v, mz+v) so that the
// z neighbours need no second halo exchange; x/y periodic images are stored alongside.
// ---------------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(256) theta_kernel(KConst c, const double *__restrict__ q, double *__restrict__ theta) {
    const Layout &L = c.L;
    int i = blockIdx.x * 32 + (threadIdx.x & 31);
    int j = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (i >= L.mx || j >= L.my) return;
    const double *U = q + 1 * L.vol, *Vv = q + 2 * L.vol, *W = q + 3 * L.vol;
    const int kglob_lo = -c.kstart, kglob_hi = c.mz_tot - c.kstart;   // local indices of global planes 0 and mz_tot
    int kb = (int)blockIdx.z * 8 - V;
    for (int kk = 0; kk < 8; kk++) {
        int k = kb + kk;
        if (k >= L.mz + V) break;
        if (c.boundaryLayer && (k < kglob_lo || k >= kglob_hi)) continue;   // outside the global domain: ghosts are extrapolated later
        size_t g0 = L.idx(i, j, k);
        double dudx = 0.0, dvdy = 0.0, dwdz = 0.0;
        // --- x
        if (c.periodicX) {
#pragma unroll
            for (int l = 1; l <= V; l++) dudx += c.aV[l] * (U[g0 + l] - U[g0 - l]);
        } else {
            double pv = 0.0; bool pert = false;
            if (c.boundaryLayer && c.perturbed && i < V) pert = perturb_value(c, j, k + c.kstart, pv);
#pragma unroll
            for (int l = 1; l <= V; l++) {
                double up, um;
                int ip = i + l, im = i - l;
                if (ip < L.mx) up = U[g0 + l];
                else if (c.boundaryLayer) up = 2.0 * U[L.idx(L.mx - 1, j, k)] - U[L.idx(2 * (L.mx - 1) - ip, j, k)];   // topBCxExt boundary.h:150
                else up = -U[L.idx(2 * L.mx - 1 - ip, j, k)];                                                          // wallBCxVel boundary.h:111
                if (im >= 0) um = U[g0 - l];
                else { um = -U[L.idx(-1 - im, j, k)]; if (pert) um = pv; }                                             // botBCxExt/wallBCxVel (+PerturbUvel)
                dudx += c.aV[l] * (up - um);
            }
        }
        dudx *= c.d1[0];
        if (c.nonUniformX) dudx *= c.xp[i];
        // --- y (always periodic: ghosts are real)
#pragma unroll
        for (int l = 1; l <= V; l++) dvdy += c.aV[l] * (Vv[g0 + (size_t)l * L.px] - Vv[g0 - (size_t)l * L.px]) * c.d1[1];
        // --- z
#pragma unroll
        for (int l = 1; l <= V; l++) {
            double wp, wm;
            if (c.boundaryLayer && k + l >= kglob_hi) wp = 2.0 * W[L.idx(i, j, kglob_hi - 1)] - W[L.idx(i, j, 2 * (kglob_hi - 1) - (k + l))];   // topBCzExt
            else wp = W[g0 + (size_t)l * L.plane];
            if (c.boundaryLayer && k - l < kglob_lo) wm = 2.0 * W[L.idx(i, j, kglob_lo)] - W[L.idx(i, j, 2 * kglob_lo - (k - l))];               // botBCzExt
            else wm = W[g0 - (size_t)l * L.plane];
            dwdz += c.aV[l] * (wp - wm) * c.d1[2];
        }
        double th = dudx + dvdy + dwdz;
        theta[g0] = th;
        // periodic images in x / y (cross-shaped ghosts only)
        if (c.periodicX) {
            if (i < V) theta[g0 + L.mx] = th;
            if (i >= L.mx - V) theta[g0 - L.mx] = th;
        }
        if (j < V) theta[g0 + (size_t)L.my * L.px] = th;
        if (j >= L.my - V) theta[g0 - (size_t)L.my * L.px] = th;
    }
}

void launch_theta_march(const KConst &kc, const double *q, double *theta, cudaStream_t st);   // theta.cu
void launch_theta(const KConst &kc, const double *q, double *theta, cudaStream_t st) {
    static const bool legacy = [] { const char *e = getenv("CUDNS_THETA"); return e && std::string(e) == "legacy"; }();
    if (!legacy) { launch_theta_march(kc, q, theta, st); return; }
    dim3 grid((kc.L.mx + 31) / 32, (kc.L.my + 7) / 8, (kc.L.mz + 2 * kc.v + 7) / 8);
    switch (kc.v) {
        case 1: theta_kernel<1><<<grid, 256, 0, st>>>(kc, q, theta); break;
        case 2: theta_kernel<2><<<grid, 256, 0, st>>>(kc, q, theta); break;
        case 3: theta_kernel<3><<<grid, 256, 0, st>>>(kc, q, theta); break;
        default: theta_kernel<4><<<grid, 256, 0, st>>>(kc, q, theta); break;
    }
}

// ---------------------------------------------------------------------------------------------
// Fused RHS + Runge-Kutta stage kernel
// ---------------------------------------------------------------------------------------------
template <int S> struct StageSmem {
    static constexpr int R = 2 * S + 1;
    static constexpr int CY = TY + 2 * S;
    static constexpr int TSZ = TX * TY;          // doubles per quantity per ring plane
    static constexpr int SLOT = NQ * TSZ;        // doubles per ring plane
    static constexpr int CSZ = CY * CX;          // doubles per quantity of the extended plane
    static constexpr size_t bytes = (size_t)(R * SLOT + NQ * CSZ) * sizeof(double);
};

struct Center { double r, vel[3], h, p, t, mu, th; };

// neighbour accessors: value of quantity q at offset l along DIR
template <int DIR, int S> struct Nb {
    const double *curc;    // &cur[0][cy][cx]
    const double *ringc;   // ring + tid
    const int *zo;         // zo[l+S] = ring offset of plane k+l
    __device__ __forceinline__ double operator()(int q, int l) const {
        if (DIR == 0) return curc[q * StageSmem<S>::CSZ + l];
        if (DIR == 1) return curc[q * StageSmem<S>::CSZ + l * CX];
        return ringc[zo[l + S] + q * StageSmem<S>::TSZ];
    }
};

template <int DIR, int S, int V>
__device__ __forceinline__ void vel_derivs(const KConst &c, const Nb<DIR, S> &nb, const Center &C, double xpi, int i,
                                           double (&g)[3][3], double (&lap)[3][3]) {
#pragma unroll
    for (int m = 0; m < 3; m++) {
        double d1 = 0.0, d2 = c.bV[0] * C.vel[m];
#pragma unroll
        for (int l = 1; l <= V; l++) {
            double fp = nb(QU + m, l), fm = nb(QU + m, -l);
            d1 += c.aV[l] * (fp - fm);
            d2 += c.bV[l] * (fp + fm);
        }
        g[m][DIR] = d1 * c.d1[DIR];
        lap[m][DIR] = d2 * c.d2[DIR];
        if (DIR == 0 && c.nonUniformX) {
            g[m][0] *= xpi;
            double t2 = 0.0;   // derDevSharedV2x non-uniform branch, cuda_derivs.h:210-214
#pragma unroll
            for (int it = 0; it < 2 * V + 1; it++) t2 += c.cVSx[it * c.L.mx + i] * nb(QU + m, it - V);
            lap[m][0] = t2;
        }
    }
}

// One direction of cuda_rhs.cu (deviceRHSX :52-127, deviceRHSY :169-259, deviceRHSZ :303-393)
template <int DIR, int S, int V>
__device__ __forceinline__ void dir_rhs(const KConst &c, const Nb<DIR, S> &nb, const Center &C, double xpi, int i,
                                        const double (&g)[3][3], const double (&lap)[3][3], double (&rhs)[5]) {
    const bool nonuni = (DIR == 0) && c.nonUniformX;
    double tmp[3];
#pragma unroll
    for (int m = 0; m < 3; m++)
        tmp[m] = (m == DIR) ? (2.0 * g[m][DIR] - (2.0 / 3.0) * C.th) : (g[m][DIR] + g[DIR][m]);
    // viscous dissipation; quirk Q1 (cuda_rhs.cu:175): the y kernel multiplies by dv/dz where dw/dy is meant
    double g3 = (DIR == 1 && c.quirk_q1) ? g[1][2] : g[2][DIR];
    double e = C.mu * (tmp[0] * g[0][DIR] + tmp[1] * g[1][DIR] + tmp[2] * g3);
    double dmu = 0.0, dT = 0.0, d2T = c.bV[0] * C.t, dth = 0.0;
#pragma unroll
    for (int l = 1; l <= V; l++) {
        double mp = nb(QM, l), mm = nb(QM, -l);
        double tp = nb(QT, l), tm = nb(QT, -l);
        double hp = nb(QD, l), hm = nb(QD, -l);
        dmu += c.aV[l] * (mp - mm);
        dT += c.aV[l] * (tp - tm);
        d2T += c.bV[l] * (tp + tm);
        dth += c.aV[l] * (hp - hm);
    }
    dmu *= c.d1[DIR]; dT *= c.d1[DIR]; dth *= c.d1[DIR]; d2T *= c.d2[DIR];
    double dp = 0.0;
#pragma unroll
    for (int l = 1; l <= S; l++) dp += c.aF[l] * (nb(QP, l) - nb(QP, -l));
    dp *= c.d1[DIR];
    if (nonuni) {
        dmu *= xpi; dT *= xpi; dth *= xpi; dp *= xpi;
        double t2 = 0.0;
#pragma unroll
        for (int it = 0; it < 2 * V + 1; it++) t2 += c.cVSx[it * c.L.mx + i] * nb(QT, it - V);
        d2T = t2;
    }
    double m3[3];
#pragma unroll
    for (int m = 0; m < 3; m++) m3[m] = tmp[m] * dmu + lap[m][DIR] * C.mu;
    e += C.vel[0] * m3[0] + C.vel[1] * m3[1] + C.vel[2] * m3[2];
    double mth = C.mu * dth / 3.0;
    m3[DIR] += mth - dp;
    e += mth * C.vel[DIR];
    e += d2T * (C.mu * c.lamfac) + dT * (dmu * c.lamfac);     // lambda = mu/(Pr Ec), cuda_main.cu:239
    // split-form convective terms
    double aM = 0.0, a0 = 0.0, a1 = 0.0, a2 = 0.0, a4 = 0.0;
#pragma unroll
    for (int l = 1; l <= S; l++) {
        double Ap = c.aF[l] * ((C.r + nb(QR, l)) * (C.vel[DIR] + nb(QU + DIR, l)));
        double Am = c.aF[l] * ((C.r + nb(QR, -l)) * (C.vel[DIR] + nb(QU + DIR, -l)));
        aM += Ap - Am;
        a0 += Ap * nb(QU, l) - Am * nb(QU, -l);
        a1 += Ap * nb(QV, l) - Am * nb(QV, -l);
        a2 += Ap * nb(QW, l) - Am * nb(QW, -l);
        a4 += Ap * nb(QH, l) - Am * nb(QH, -l);
    }
    double fac = nonuni ? c.d1[DIR] * xpi : c.d1[DIR];
    rhs[0] += -0.5 * fac * aM;
    rhs[1] += m3[0] - 0.25 * fac * (C.vel[0] * aM + a0);
    rhs[2] += m3[1] - 0.25 * fac * (C.vel[1] * aM + a1);
    rhs[3] += m3[2] - 0.25 * fac * (C.vel[2] * aM + a2);
    rhs[4] += e - 0.25 * fac * (C.h * aM + a4);
}

template <int S, int V>
__global__ void __launch_bounds__(NTHREADS, 1) rhs_stage_kernel(KConst c, StagePtrs P, StageCoef sc, int zchunk) {
    using SM = StageSmem<S>;
    constexpr int R = SM::R, CY = SM::CY, TSZ = SM::TSZ, SLOT = SM::SLOT, CSZ = SM::CSZ;
    extern __shared__ __align__(16) double smem[];
    double *ring = smem;                 // [R][NQ][TY][TX]
    double *cur = smem + R * SLOT;       // [NQ][CY][CX]

    const Layout &L = c.L;
    const int tid = threadIdx.x;
    const int tx = tid & (TX - 1), ty = tid / TX;
    const int i0 = blockIdx.x * TX, j0 = blockIdx.y * TY;
    const int kbeg = blockIdx.z * zchunk;
    const int kend = min(kbeg + zchunk, L.mz);
    const int i = i0 + tx, j = j0 + ty;
    const bool active = (i < L.mx) && (j < L.my);
    const int ic = min(i, L.mx - 1), jc = min(j, L.my - 1);
    const size_t vol = L.vol;
    const int kglob_lo = -c.kstart, kglob_hi = c.mz_tot - c.kstart;
    const bool xlo_tile = !c.periodicX && (i0 == 0);
    const bool xhi_tile = !c.periodicX && (i0 + TX >= L.mx);
    const int nxt = min(TX, L.mx - i0);          // interior columns of this tile

    auto slot_of = [](int kk) { return ((kk + 16 * R) % R) * SLOT; };

    // load plane kk (interior columns of the tile) -> EOS -> ring
    auto ring_insert = [&](int kk) {
        double *dst = ring + slot_of(kk) + tid;
        if (c.boundaryLayer && (kk < kglob_lo || kk >= kglob_hi)) {
            if (kk >= kglob_hi) {
                // topBCzExt (boundary.h:154-156): f[mz-1+g] = 2 f[mz-1] - f[mz-1-g], every staged quantity independently
                int gq = kk - (kglob_hi - 1);
                const double *a = ring + slot_of(kglob_hi - 1) + tid, *b = ring + slot_of(kglob_hi - 1 - gq) + tid;
#pragma unroll
                for (int q = 0; q < NQ; q++) dst[q * TSZ] = 2.0 * a[q * TSZ] - b[q * TSZ];
            }
            return;   // bottom ghosts are generated once plane kglob_lo+S is in (see below)
        }
        size_t g = L.idx(ic, min(j, L.my + S - 1), kk);    // rows past my are the periodic y images (partial tiles)
        double r = P.qin[g], u = P.qin[vol + g], v = P.qin[2 * vol + g], w = P.qin[3 * vol + g], e = P.qin[4 * vol + g];
        double th = P.theta[g];
        double h, p, t, m;
        eos(c, r, u, v, w, e, h, p, t, m);
        dst[QR * TSZ] = r; dst[QU * TSZ] = u; dst[QV * TSZ] = v; dst[QW * TSZ] = w;
        dst[QH * TSZ] = h; dst[QP * TSZ] = p; dst[QT * TSZ] = t; dst[QM * TSZ] = m; dst[QD * TSZ] = th;
    };
    // botBCzExt (boundary.h:158-160): f[-g] = 2 f[0] - f[g]
    auto ring_bottom_ghosts = [&]() {
        const double *a = ring + slot_of(kglob_lo) + tid;
#pragma unroll
        for (int gq = 1; gq <= S; gq++) {
            const double *b = ring + slot_of(kglob_lo + gq) + tid;
            double *dst = ring + slot_of(kglob_lo - gq) + tid;
#pragma unroll
            for (int q = 0; q < NQ; q++) dst[q * TSZ] = 2.0 * a[q * TSZ] - b[q * TSZ];
        }
    };

    // ---- prologue: planes kbeg-S .. kbeg+S-1
    for (int kk = kbeg - S; kk < kbeg + S; kk++) ring_insert(kk);

    for (int k = kbeg; k < kend; k++) {
        ring_insert(k + S);
        if (c.boundaryLayer && k == kglob_lo) ring_bottom_ghosts();
        // ---- extended plane k: halo cells from global memory
        {
            constexpr int NXH = 2 * S * TY, NYH = 2 * S * TX;
            for (int cidx = tid; cidx < NXH + NYH; cidx += NTHREADS) {
                int cx, cy;
                if (cidx < NXH) { int hx = cidx % (2 * S), hy = cidx / (2 * S); cx = hx < S ? GX - S + hx : GX + nxt + (hx - S); cy = S + hy; }
                else { int d = cidx - NXH; int hx = d % TX, hy = d / TX; cx = GX + hx; cy = hy < S ? hy : TY + hy; }
                int gi = i0 + cx - GX, gj = j0 + cy - S;
                bool inx = c.periodicX || (gi >= 0 && gi < L.mx);
                if (!inx) continue;
                if (gj >= L.my + S) gj = L.my + S - 1;     // partial tiles in y: stay inside the allocation
                if (gi >= L.mx + GX) gi = L.mx + GX - 1;
                size_t g = L.idx(gi, gj, k);
                double r = P.qin[g], u = P.qin[vol + g], v = P.qin[2 * vol + g], w = P.qin[3 * vol + g], e = P.qin[4 * vol + g];
                double th = P.theta[g];
                double h, p, t, m;
                eos(c, r, u, v, w, e, h, p, t, m);
                double *d = cur + cy * CX + cx;
                d[QR * CSZ] = r; d[QU * CSZ] = u; d[QV * CSZ] = v; d[QW * CSZ] = w;
                d[QH * CSZ] = h; d[QP * CSZ] = p; d[QT * CSZ] = t; d[QM * CSZ] = m; d[QD * CSZ] = th;
            }
            // interior of plane k comes from the ring (already EOS'd s steps ago)
            const double *src = ring + slot_of(k) + tid;
            double *d = cur + (ty + S) * CX + (tx + GX);
            if (tx < nxt) {
#pragma unroll
                for (int q = 0; q < NQ; q++) d[q * CSZ] = src[q * TSZ];
            }
        }
        __syncthreads();
        // ---- x boundary rules on the extended plane (boundary_condition_x.h BCxNumber1-3)
        if (xlo_tile || xhi_tile) {
            for (int cidx = tid; cidx < 2 * S * TY; cidx += NTHREADS) {
                int gq = cidx % S + 1, side = (cidx / S) & 1, row = cidx / (2 * S);
                if (side == 0 && !xlo_tile) continue;
                if (side == 1 && !xhi_tile) continue;
                if (j0 + row >= L.my) continue;
                double *rowp = cur + (row + S) * CX;
                int cg, cm;             // ghost column, mirror column
                double u, v, w, p, t, th;
                if (side == 0) {
                    cg = GX - gq; cm = GX + gq - 1;   // cell mirror: f[-g] <- f[g-1]
                    u = -rowp[QU * CSZ + cm]; v = -rowp[QV * CSZ + cm]; w = -rowp[QW * CSZ + cm];     // wallBCxVel / botBCxExt(.,0)
                    p = rowp[QP * CSZ + cm];                                                         // wallBCxMir / botBCxMir
                    th = rowp[QD * CSZ + cm];                                                        // BCxNumber2
                    if (c.boundaryLayer) {
                        t = rowp[QT * CSZ + cm];                                                     // botBCxMir (adiabatic)
                        double pv;
                        if (c.perturbed && perturb_value(c, j0 + row, k + c.kstart, pv)) u = pv;      // PerturbUvel
                    } else {
                        t = 2.0 * c.TwallBot - rowp[QT * CSZ + cm];                                  // wallBCxExt
                    }
                } else {
                    int last = GX + nxt - 1;
                    cg = last + gq;
                    if (c.boundaryLayer) {
                        cm = last - gq;         // node extrapolation topBCxExt: f[mx-1+g] = 2 f[mx-1] - f[mx-1-g]
                        u = 2.0 * rowp[QU * CSZ + last] - rowp[QU * CSZ + cm];
                        v = 2.0 * rowp[QV * CSZ + last] - rowp[QV * CSZ + cm];
                        w = 2.0 * rowp[QW * CSZ + last] - rowp[QW * CSZ + cm];
                        p = 2.0 * rowp[QP * CSZ + last] - rowp[QP * CSZ + cm];
                        t = 2.0 * rowp[QT * CSZ + last] - rowp[QT * CSZ + cm];
                        th = 2.0 * rowp[QD * CSZ + last] - rowp[QD * CSZ + cm];
                    } else {
                        cm = last - gq + 1;
                        u = -rowp[QU * CSZ + cm]; v = -rowp[QV * CSZ + cm]; w = -rowp[QW * CSZ + cm];
                        p = rowp[QP * CSZ + cm];
                        th = rowp[QD * CSZ + cm];
                        t = 2.0 * c.TwallTop - rowp[QT * CSZ + cm];
                    }
                }
                rowp[QU * CSZ + cg] = u; rowp[QV * CSZ + cg] = v; rowp[QW * CSZ + cg] = w;
                rowp[QP * CSZ + cg] = p; rowp[QT * CSZ + cg] = t; rowp[QD * CSZ + cg] = th;
                rowp[QM * CSZ + cg] = visc_of(c, t);                                                     // mlBoundPT boundary.h:135
                rowp[QH * CSZ + cg] = t * c.Rgas * c.gam / (c.gam - 1.0) + 0.5 * (u * u + v * v + w * w);  // rhBoundPT boundary.h:121
                rowp[QR * CSZ + cg] = p / (c.Rgas * t);
            }
            __syncthreads();
        }

        // ---- right-hand side at (i,j,k)
        const double *curc = cur + (ty + S) * CX + (tx + GX);
        int zo[R];
#pragma unroll
        for (int m = 0; m < R; m++) zo[m] = slot_of(k + m - S);
        Nb<0, S> nx{curc, ring + tid, zo};
        Nb<1, S> ny{curc, ring + tid, zo};
        Nb<2, S> nz{curc, ring + tid, zo};
        Center C;
        C.r = curc[QR * CSZ]; C.vel[0] = curc[QU * CSZ]; C.vel[1] = curc[QV * CSZ]; C.vel[2] = curc[QW * CSZ];
        C.h = curc[QH * CSZ]; C.p = curc[QP * CSZ]; C.t = curc[QT * CSZ]; C.mu = curc[QM * CSZ]; C.th = curc[QD * CSZ];
        const double xpi = c.nonUniformX ? c.xp[ic] : 1.0;
        double g[3][3], lap[3][3];
        vel_derivs<0, S, V>(c, nx, C, xpi, ic, g, lap);
        vel_derivs<1, S, V>(c, ny, C, xpi, ic, g, lap);
        vel_derivs<2, S, V>(c, nz, C, xpi, ic, g, lap);
        double rhs[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
        dir_rhs<0, S, V>(c, nx, C, xpi, ic, g, lap, rhs);
        dir_rhs<1, S, V>(c, ny, C, xpi, ic, g, lap, rhs);
        dir_rhs<2, S, V>(c, nz, C, xpi, ic, g, lap, rhs);
        if (c.forcing) {                       // cuda_rhs.cu:392-393
            double f = *c.dpdz;
            rhs[3] += f; rhs[4] += f * C.vel[2];
        }
        const size_t gp = L.idx(ic, jc, k);
        const double e_c = P.qin[4 * vol + gp];
        if (c.boundaryLayer && c.spongeX) {    // addSponge, sponge.cu:31-41
            double sg = c.spongeX[ic] + c.spongeZ[k];
            size_t nq = (size_t)L.mx * L.mz, qi = (size_t)ic + (size_t)k * L.mx;
            rhs[0] += sg * (c.sref[qi] - C.r);
            rhs[1] += sg * (c.sref[nq + qi] - C.r * C.vel[0]);
            rhs[2] += sg * (c.sref[2 * nq + qi] - C.r * C.vel[1]);
            rhs[3] += sg * (c.sref[3 * nq + qi] - C.r * C.vel[2]);
            rhs[4] += sg * (c.sref[4 * nq + qi] - e_c);
        }
        if (active) {
            const size_t N = (size_t)L.mx * L.my * L.mz;
            const size_t n = (size_t)i + (size_t)j * L.mx + (size_t)k * L.mx * L.my;
            if (P.rhs_out) {
#pragma unroll
                for (int m = 0; m < 5; m++) P.rhs_out[m * N + n] = rhs[m];
            } else {
                // Runge-Kutta register update (sumLowStorageRK3 cuda_main.cu:244, eulerSum*/rk3final* :188-216)
                const double dt = *c.dt;
                double qb[5];
                if (P.qbase == P.qin) { qb[0] = C.r; qb[1] = C.r * C.vel[0]; qb[2] = C.r * C.vel[1]; qb[3] = C.r * C.vel[2]; qb[4] = e_c; }
                else {
                    double rb = P.qbase[gp];
                    qb[0] = rb; qb[1] = rb * P.qbase[vol + gp]; qb[2] = rb * P.qbase[2 * vol + gp]; qb[3] = rb * P.qbase[3 * vol + gp];
                    qb[4] = P.qbase[4 * vol + gp];
                }
                double qn[5];
#pragma unroll
                for (int m = 0; m < 5; m++) {
                    double inc = sc.cN * rhs[m];
                    if (P.RA) inc += sc.cA * P.RA[m * N + n];
                    if (P.RB) inc += sc.cB * P.RB[m * N + n];
                    qn[m] = qb[m] + dt * inc;
                    if (P.RW) P.RW[m * N + n] = (sc.wOld != 0.0) ? sc.wOld * P.RW[m * N + n] + sc.wNew * rhs[m] : sc.wNew * rhs[m];
                }
                double out[5] = {qn[0], qn[1] / qn[0], qn[2] / qn[0], qn[3] / qn[0], qn[4]};    // deviceDiv cuda_math.cu:36
#pragma unroll
                for (int m = 0; m < 5; m++) {
                    double *f = P.qout + m * vol;
                    f[gp] = out[m];
                    // periodic images (cross-shaped ghosts): perBCx / perBCy, boundary.h:38-46
                    if (c.periodicX) {
                        if (i < S) f[gp + L.mx] = out[m];
                        if (i >= L.mx - S) f[gp - L.mx] = out[m];
                    }
                    if (j < S) f[gp + (size_t)L.my * L.px] = out[m];
                    if (j >= L.my - S) f[gp - (size_t)L.my * L.px] = out[m];
                }
            }
        }
        __syncthreads();
    }
}

int rhs_stage_smem_bytes(int s) {
    switch (s) { case 1: return (int)StageSmem<1>::bytes; case 2: return (int)StageSmem<2>::bytes;
                 case 3: return (int)StageSmem<3>::bytes; default: return (int)StageSmem<4>::bytes; }
}
bool rhs_stage_supported(int s, int v) { return s >= 1 && s <= 4 && v >= 1 && v <= s; }

template <int S, int V>
static void launch_rhs_stage_t(const KConst &kc, const StagePtrs &p, const StageCoef &c, cudaStream_t st) {
    static bool attr_set = false;
    size_t smem = StageSmem<S>::bytes;
    if (!attr_set) { cudaFuncSetAttribute(rhs_stage_kernel<S, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr_set = true; }
    int gx = (kc.L.mx + TX - 1) / TX, gy = (kc.L.my + TY - 1) / TY;
    // z chunks: enough CTAs to fill 148 SMs several times over, but long enough to amortise the 2s-plane prologue
    int cols = gx * gy;
    int nzc = 1;
    while (cols * nzc < 148 * 4 && kc.L.mz / (nzc * 2) >= 16 * S) nzc *= 2;
    int zchunk = (kc.L.mz + nzc - 1) / nzc;
    nzc = (kc.L.mz + zchunk - 1) / zchunk;
    dim3 grid(gx, gy, nzc);
    rhs_stage_kernel<S, V><<<grid, NTHREADS, smem, st>>>(kc, p, c, zchunk);
}

void launch_rhs_stage_smem(const KConst &kc, const StagePtrs &p, const StageCoef &c, cudaStream_t st) {
    switch (kc.s * 10 + kc.v) {
        case 11: launch_rhs_stage_t<1, 1>(kc, p, c, st); break;
        case 21: launch_rhs_stage_t<2, 1>(kc, p, c, st); break;
        case 22: launch_rhs_stage_t<2, 2>(kc, p, c, st); break;
        case 31: launch_rhs_stage_t<3, 1>(kc, p, c, st); break;
        case 32: launch_rhs_stage_t<3, 2>(kc, p, c, st); break;
        case 33: launch_rhs_stage_t<3, 3>(kc, p, c, st); break;
        case 41: launch_rhs_stage_t<4, 1>(kc, p, c, st); break;
        case 42: launch_rhs_stage_t<4, 2>(kc, p, c, st); break;
        case 43: launch_rhs_stage_t<4, 3>(kc, p, c, st); break;
        case 44: launch_rhs_stage_t<4, 4>(kc, p, c, st); break;
        default: break;
    }
}

// ---------------------------------------------------------------------------------------------
// ghost handling, staging copies
// ---------------------------------------------------------------------------------------------
// periodic x/y images for nfields padded fields, all local planes incl. z ghosts (used by set_state)
__global__ void fill_xy_kernel(KConst c, double *q, int nfields) {
    const Layout &L = c.L;
    int kz = blockIdx.z;                    // padded plane index
    int f = blockIdx.y;
    double *p = q + (size_t)f * L.vol + (size_t)kz * L.plane;
    int s = c.s;
    // x images for interior rows
    if (c.periodicX) {
        for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < L.my * s; t += gridDim.x * blockDim.x) {
            int j = t / s, gq = t % s;
            double *row = p + (size_t)(j + L.gy) * L.px + GX;
            row[-1 - gq] = row[L.mx - 1 - gq];
            row[L.mx + gq] = row[gq];
        }
    }
    // y images for interior columns
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < L.mx * s; t += gridDim.x * blockDim.x) {
        int i = t % L.mx, gq = t / L.mx;
        double *col = p + (size_t)L.gy * L.px + GX + i;
        col[-(ptrdiff_t)(1 + gq) * L.px] = col[(size_t)(L.my - 1 - gq) * L.px];
        col[(size_t)(L.my + gq) * L.px] = col[(size_t)gq * L.px];
    }
}
void launch_fill_xy(const KConst &kc, double *q5, int nfields, cudaStream_t st) {
    dim3 grid(8, nfields, kc.L.pz);
    fill_xy_kernel<<<grid, 256, 0, st>>>(kc, q5, nfields);
}

// periodic z wrap on one device: copy gz full padded planes bottom<->top (perBCz boundary.h:48-51)
__global__ void zwrap_kernel(KConst c, double *q, int nfields) {
    const Layout &L = c.L;
    size_t n = (size_t)L.gz * L.plane;      // doubles per block
    int f = blockIdx.y;
    double *p = q + (size_t)f * L.vol;
    double2 *lo_ghost = (double2 *)p, *hi_ghost = (double2 *)(p + (size_t)(L.gz + L.mz) * L.plane);
    const double2 *lo_int = (const double2 *)(p + (size_t)L.gz * L.plane), *hi_int = (const double2 *)(p + (size_t)L.mz * L.plane);
    size_t n2 = n / 2;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n2; t += (size_t)gridDim.x * blockDim.x) {
        lo_ghost[t] = hi_int[t];
        hi_ghost[t] = lo_int[t];
    }
}
void launch_zwrap(const KConst &kc, double *q5, int nfields, cudaStream_t st) {
    dim3 grid(148 * 2, nfields);
    zwrap_kernel<<<grid, 256, 0, st>>>(kc, q5, nfields);
}

// pack the first/last gz interior planes of 5 fields into contiguous send blocks; unpack ghost blocks
__global__ void pack_z_kernel(KConst c, const double *q, double *send_lo, double *send_hi) {
    const Layout &L = c.L;
    size_t n2 = (size_t)L.gz * L.plane / 2;
    int f = blockIdx.y;
    const double *p = q + (size_t)f * L.vol;
    const double2 *lo_int = (const double2 *)(p + (size_t)L.gz * L.plane), *hi_int = (const double2 *)(p + (size_t)L.mz * L.plane);
    double2 *slo = (double2 *)send_lo + (size_t)f * n2, *shi = (double2 *)send_hi + (size_t)f * n2;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n2; t += (size_t)gridDim.x * blockDim.x) {
        slo[t] = lo_int[t]; shi[t] = hi_int[t];
    }
}
__global__ void unpack_z_kernel(KConst c, double *q, const double *recv_lo, const double *recv_hi) {
    const Layout &L = c.L;
    size_t n2 = (size_t)L.gz * L.plane / 2;
    int f = blockIdx.y;
    double *p = q + (size_t)f * L.vol;
    double2 *lo_ghost = (double2 *)p, *hi_ghost = (double2 *)(p + (size_t)(L.gz + L.mz) * L.plane);
    const double2 *rlo = (const double2 *)recv_lo + (size_t)f * n2, *rhi = (const double2 *)recv_hi + (size_t)f * n2;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n2; t += (size_t)gridDim.x * blockDim.x) {
        lo_ghost[t] = rlo[t]; hi_ghost[t] = rhi[t];
    }
}
void launch_pack_z(const KConst &kc, const double *q5, double *send_lo, double *send_hi, cudaStream_t st) {
    pack_z_kernel<<<dim3(148, 5), 256, 0, st>>>(kc, q5, send_lo, send_hi);
}
void launch_unpack_z(const KConst &kc, double *q5, const double *recv_lo, const double *recv_hi, cudaStream_t st) {
    unpack_z_kernel<<<dim3(148, 5), 256, 0, st>>>(kc, q5, recv_lo, recv_hi);
}

// unpadded [mz][my][mx] <-> padded interior (initDevice / getResults, cuda_utils.cu:372-413)
struct Ptr5 { const double *p[5]; };
struct Ptr5w { double *p[5]; };
__global__ void pad_kernel(KConst c, Ptr5 src, double *q) {
    const Layout &L = c.L;
    size_t N = (size_t)L.mx * L.my * L.mz;
    for (size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (size_t)gridDim.x * blockDim.x) {
        int i = (int)(n % L.mx); size_t r = n / L.mx; int j = (int)(r % L.my), k = (int)(r / L.my);
        size_t g = L.idx(i, j, k);
#pragma unroll
        for (int f = 0; f < 5; f++) q[f * L.vol + g] = src.p[f][n];
    }
}
__global__ void unpad_kernel(KConst c, const double *q, Ptr5w dst) {
    const Layout &L = c.L;
    size_t N = (size_t)L.mx * L.my * L.mz;
    for (size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (size_t)gridDim.x * blockDim.x) {
        int i = (int)(n % L.mx); size_t r = n / L.mx; int j = (int)(r % L.my), k = (int)(r / L.my);
        size_t g = L.idx(i, j, k);
#pragma unroll
        for (int f = 0; f < 5; f++) dst.p[f][n] = q[f * L.vol + g];
    }
}
void launch_pad(const KConst &kc, const double *src5[5], double *q5, cudaStream_t st) {
    Ptr5 s; for (int f = 0; f < 5; f++) s.p[f] = src5[f];
    pad_kernel<<<148 * 8, 256, 0, st>>>(kc, s, q5);
}
void launch_unpad(const KConst &kc, const double *q5, double *dst5[5], cudaStream_t st) {
    Ptr5w d; for (int f = 0; f < 5; f++) d.p[f] = dst5[f];
    unpad_kernel<<<148 * 8, 256, 0, st>>>(kc, q5, d);
}

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_max(double v) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// positive doubles order like their bit patterns
__device__ __forceinline__ void atomic_max_pos(double *addr, double v) {
    atomicMax((unsigned long long *)addr, (unsigned long long)__double_as_longlong(v));
}

// deviceCalcDt (calc_stress.cu:140-160) as two maxima: min_p CFL/max(conv_p,visc_p) = CFL/max(max_p conv_p, max_p visc_p)
__global__ void __launch_bounds__(256) dt_reduce_kernel(KConst c, const double *__restrict__ q, double *out2) {
    const Layout &L = c.L;
    const int nrows = L.my * L.mz;
    double mc = 0.0, mv = 0.0;
    for (int row = blockIdx.x; row < nrows; row += gridDim.x) {            // one (j,k) row of the interior per iteration
        const int j = row % L.my, k = row / L.my;
        const double *p = q + L.idx(0, j, k);
        for (int i = threadIdx.x; i < L.mx; i += blockDim.x) {
            const double r = p[i], u = p[L.vol + i], v = p[2 * L.vol + i], w = p[3 * L.vol + i], e = p[4 * L.vol + i];
            double ien = e / r - 0.5 * (u * u + v * v + w * w);
            double sos = sqrt(c.gam * (c.gam - 1) * ien);
            double dx = c.dxv[i], d2x = dx * dx;
            double conv = fmax((fabs(u) + sos) / dx, fmax((fabs(v) + sos) * c.d1[1], (fabs(w) + sos) * c.d1[2]));
            double mu = visc_of(c, c.cvInv * ien);
            double visc = fmax(mu / d2x, fmax(mu * c.d2[1], mu * c.d2[2]));
            mc = fmax(mc, conv); mv = fmax(mv, visc);
        }
    }
    mc = warp_max(mc); mv = warp_max(mv);
    if ((threadIdx.x & 31) == 0) { atomic_max_pos(out2, mc); atomic_max_pos(out2 + 1, mv); }
}
void launch_dt_reduce(const KConst &kc, const double *q, double *out2, cudaStream_t st) {
    cudaMemsetAsync(out2, 0, 2 * sizeof(double), st);
    dt_reduce_kernel<<<148 * 8, 256, 0, st>>>(kc, q, out2);
}

// calcBulk / calcPressureGrad integrals (calc_stress.cu:98-120,162-201; cuda_math.cu:143-201):
// out[0] = sum (u.u) w_avg, out[1] = sum rho w_int, out[2] = sum rho*w w_int, out[3] = sum rhoE w_int
// with w_int = dxv[i]/d_dy/d_dz and w_avg = w_int/Lx/Ly/Lz.  Block partials are combined in a fixed order
// by the last block (deterministic for a given launch shape).
__global__ void __launch_bounds__(256) bulk_reduce_kernel(KConst c, const double *__restrict__ q, double *out4, double *partial, unsigned int *counter) {
    const Layout &L = c.L;
    const int nrows = L.my * L.mz;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int row = blockIdx.x; row < nrows; row += gridDim.x) {
        const int j = row % L.my, k = row / L.my;
        const double *p = q + L.idx(0, j, k);
        for (int i = threadIdx.x; i < L.mx; i += blockDim.x) {
            const double r = p[i], u = p[L.vol + i], v = p[2 * L.vol + i], w = p[3 * L.vol + i], e = p[4 * L.vol + i];
            double wi = c.dxv[i] / c.d1[1] / c.d1[2];
            s0 += (u * u + v * v + w * w) * wi / c.Lx / c.Ly / c.Lz;
            s1 += r * wi; s2 += r * w * wi; s3 += e * wi;
        }
    }
    __shared__ double sh[4][8];
    s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2); s3 = warp_sum(s3);
    int wid = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { sh[0][wid] = s0; sh[1][wid] = s1; sh[2][wid] = s2; sh[3][wid] = s3; }
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) {
        for (int m = 0; m < 4; m++) { double a = 0; for (int q8 = 0; q8 < 8; q8++) a += sh[m][q8]; partial[m * gridDim.x + blockIdx.x] = a; }
        __threadfence();
        unsigned int t = atomicAdd(counter, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x < 4) {
        double a = 0; for (unsigned int b = 0; b < gridDim.x; b++) a += partial[threadIdx.x * gridDim.x + b];
        out4[threadIdx.x] = a;
        if (threadIdx.x == 0) *counter = 0;
    }
}
static double *g_bulk_partial[16] = {nullptr};
static unsigned int *g_bulk_counter[16] = {nullptr};
void launch_bulk_reduce(const KConst &kc, const double *q, double *out4, cudaStream_t st) {
    int dev = 0; cudaGetDevice(&dev);
    const int nb = 148 * 4;
    if (!g_bulk_partial[dev]) {
        cudaMalloc(&g_bulk_partial[dev], 4 * nb * sizeof(double));
        cudaMalloc(&g_bulk_counter[dev], sizeof(unsigned int));
        cudaMemset(g_bulk_counter[dev], 0, sizeof(unsigned int));
    }
    bulk_reduce_kernel<<<nb, 256, 0, st>>>(kc, q, out4, g_bulk_partial[dev], g_bulk_counter[dev]);
}

// tiny scalar kernels (deviceSumOne, deviceAdvanceTime, deviceCalcPress, ... cuda_math.cu:17-52, calc_stress.cu:12-18)
//  op 0: a = CFL / max(b[0], c[0])        (b: conv max of the new state, c: stale viscous max)    -- dt
//  op 1: a[0] += b[0]                                                                             -- time accumulation
//  op 2: a = 0.99*a - 0.5*(b[2]/b[1] - 1)  (b = bulk sums)                                        -- deviceCalcPress
//  op 3: a = b[2]/b[1]                                                                            -- par1 with forcing
__global__ void scalar_kernel(int op, double *a, const double *b, const double *c, double cfl) {
    if (op == 0) *a = cfl / fmax(b[0], c[0]);
    else if (op == 1) *a += *b;
    else if (op == 2) *a = 0.99 * (*a) - 0.5 * (b[2] / b[1] - 1);
    else if (op == 3) *a = b[2] / b[1];      // deviceDivOne: bulk velocity (calc_stress.cu:184)
}
void launch_scalar_ops(int op, double *a, const double *b, const double *c, cudaStream_t st) {
    scalar_kernel<<<1, 1, 0, st>>>(op, a, b, c, 0.0);
}
void launch_dt_combine(double *dt, const double *conv, const double *visc, double cfl, cudaStream_t st) {
    scalar_kernel<<<1, 1, 0, st>>>(0, dt, conv, visc, cfl);
}


// ---------------------------------------------------------------------------------------------
// stage hand-shake between z-slab neighbours (replaces the host-side MPI_Sendrecv + cudaDeviceSynchronize of
// updateHaloFive, comm.cpp:114-134).  The stage kernel has already stored its boundary planes into the neighbours'
// ghost planes; the signal makes them visible (system-scope release) and tells the neighbour which stage they belong to.
// ---------------------------------------------------------------------------------------------
__global__ void halo_signal_kernel(unsigned long long *lo_slot, unsigned long long *hi_slot, unsigned long long epoch) {
    __threadfence_system();
    if (lo_slot) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(lo_slot), "l"(epoch) : "memory");
    if (hi_slot) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(hi_slot), "l"(epoch) : "memory");
}
__global__ void halo_wait_kernel(const unsigned long long *slots, int need_lo, int need_hi, unsigned long long epoch) {
    // slots[0]: written by the lower neighbour, slots[1]: by the upper one
    for (int n = 0; n < 2; n++) {
        if (!(n == 0 ? need_lo : need_hi)) continue;
        unsigned long long v;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(slots + n) : "memory");
            if (v < epoch) __nanosleep(200);
        } while (v < epoch);
    }
    __threadfence_system();
}
void launch_halo_signal(unsigned long long *peer_lo_slot, unsigned long long *peer_hi_slot, unsigned long long epoch, cudaStream_t st) {
    halo_signal_kernel<<<1, 1, 0, st>>>(peer_lo_slot, peer_hi_slot, epoch);
}
void launch_halo_wait(const unsigned long long *my_slots, int need_lo, int need_hi, unsigned long long epoch, cudaStream_t st) {
    halo_wait_kernel<<<1, 1, 0, st>>>(my_slots, need_lo, need_hi, epoch);
}

}  // namespace cudns
