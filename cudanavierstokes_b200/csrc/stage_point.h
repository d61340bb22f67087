// libcudns: per-point arithmetic shared by the fourth- and fifth-generation stage kernels (stage_fast.cu, stage_duo.inc): the
// quantities a point carries, the equation of state of a stored point, the running sums and the neighbour-pair step of the
// split-form / expanded-viscous right-hand side.  Algebra: see stage_lean.inc.  Reference: cuda_rhs.cu:9-396,
// cuda_derivs.h:30-155, cuda_main.cu:218-242.
#pragma once
#include "cudns_internal.h"

namespace cudns {
namespace fast {

enum { FR = 0, FU, FV, FW, FH, FT, FD, NF };   // ring / shared-plane quantities

// H and T of a stored point: calcState (cuda_main.cu:218-242) restricted to what the kernels stage.  ONE definition for the update
// of the stage kernel and for derive_aux_kernel (same operation order -> same bits as eos_q / eos7 of the older kernels)
__device__ __forceinline__ void eos_ht(const KConst &c, real r, real rinv, real u, real v, real w, real e, real &H, real &T) {
    const real en = fma(e, rinv, RC(-0.5) * fma(u, u, fma(v, v, w * w)));
    const real t = c.cvInv * en;
    const real p = r * c.Rgas * t;
    H = (e + p) * rinv; T = t;
}


// running sums of one point
struct Acc {
    real r[5];               // convective sums; the momentum entries also collect -dp/dx_d
    real lapu[3], lapT;      // sum_d D2_d u_m + (1/3) d theta / d x_m ; sum_d D2_d T
    real g[3][3];            // g[d][m] = d u_m / d x_d
    real dT[3];
};

// both neighbours of direction D at offset l: split-form convective sums in telescoped pair form, pressure gradient from rho*T,
// viscous-order first and second differences (dir_sums of stage_lean.inc with PRT)
template <int D, int V>
__device__ __forceinline__ void pair_step(const KConst &c, const int l, const real (&C)[NF], const real (&Pn)[NF], const real (&Mn)[NF], Acc &A,
                                          real &aM) {
    const real cC = c.cf[D][l][0];
    const real cu = cC * C[FU + D];
    const real Ap = (C[FR] + Pn[FR]) * fma(cC, Pn[FU + D], cu);
    const real Am = (C[FR] + Mn[FR]) * fma(cC, Mn[FU + D], cu);
    aM += Ap - Am;
    A.r[1] = fma(Ap, Pn[FU], A.r[1]); A.r[1] = fma(-Am, Mn[FU], A.r[1]);
    A.r[2] = fma(Ap, Pn[FV], A.r[2]); A.r[2] = fma(-Am, Mn[FV], A.r[2]);
    A.r[3] = fma(Ap, Pn[FW], A.r[3]); A.r[3] = fma(-Am, Mn[FW], A.r[3]);
    A.r[4] = fma(Ap, Pn[FH], A.r[4]); A.r[4] = fma(-Am, Mn[FH], A.r[4]);
    A.r[1 + D] = fma(c.cfp[D][l], fma(Pn[FR], Pn[FT], -(Mn[FR] * Mn[FT])), A.r[1 + D]);
    if (l <= V) {
        const real k1 = c.cf[D][l][2], k2 = c.cf[D][l][3];
#pragma unroll
        for (int m = 0; m < 3; m++) {
            A.g[D][m] = fma(k1, Pn[FU + m] - Mn[FU + m], A.g[D][m]);
            A.lapu[m] = fma(k2, Pn[FU + m] + Mn[FU + m], A.lapu[m]);
        }
        A.dT[D] = fma(k1, Pn[FT] - Mn[FT], A.dT[D]);
        A.lapT = fma(k2, Pn[FT] + Mn[FT], A.lapT);
        A.lapu[D] = fma(c.c1t[D][l], Pn[FD] - Mn[FD], A.lapu[D]);
    }
}
// one neighbour of direction D at offset +l (PLUS) or -l: the same sums, one side at a time (one more FP64 instruction per pair,
// half the registers for neighbour values)
template <int D, int V, bool PLUS>
__device__ __forceinline__ void side_step(const KConst &c, const int l, const real (&C)[NF], const real (&Nq)[NF], Acc &A, real &aM) {
    const real cC = c.cf[D][l][0];
    const real cu = cC * C[FU + D];
    const real Af = (C[FR] + Nq[FR]) * fma(cC, Nq[FU + D], cu);
    const real pn = Nq[FR] * Nq[FT];
    const real sA = PLUS ? Af : -Af;
    aM += sA;
    A.r[1] = fma(sA, Nq[FU], A.r[1]); A.r[2] = fma(sA, Nq[FV], A.r[2]); A.r[3] = fma(sA, Nq[FW], A.r[3]); A.r[4] = fma(sA, Nq[FH], A.r[4]);
    A.r[1 + D] = fma(PLUS ? c.cfp[D][l] : -c.cfp[D][l], pn, A.r[1 + D]);
    if (l <= V) {
        const real k1 = PLUS ? c.cf[D][l][2] : -c.cf[D][l][2], k2 = c.cf[D][l][3];
#pragma unroll
        for (int m = 0; m < 3; m++) {
            A.g[D][m] = fma(k1, Nq[FU + m], A.g[D][m]);
            A.lapu[m] = fma(k2, Nq[FU + m], A.lapu[m]);
        }
        A.dT[D] = fma(k1, Nq[FT], A.dT[D]);
        A.lapT = fma(k2, Nq[FT], A.lapT);
        A.lapu[D] = fma(PLUS ? c.c1t[D][l] : -c.c1t[D][l], Nq[FD], A.lapu[D]);
    }
}
// the direction is complete: central values times its mass-flux sum
__device__ __forceinline__ void close_dir(const real (&C)[NF], Acc &A, const real aM) {
    A.r[0] = fma(RC(2.0), aM, A.r[0]);
    A.r[1] = fma(C[FU], aM, A.r[1]); A.r[2] = fma(C[FV], aM, A.r[2]); A.r[3] = fma(C[FW], aM, A.r[3]); A.r[4] = fma(C[FH], aM, A.r[4]);
}


// stress, dissipation, heat flux assembled once per point (cuda_rhs.cu:52-127,169-259,303-393); g[d][m] = d u_m / d x_d; fz = body
// force dpdz (cuda_rhs.cu:392-393), 0 when there is none
__device__ __forceinline__ void assemble_rhs(const KConst &c, const real (&C)[NF], const Acc &A, const real fz, real (&rhs)[5]) {
    const real g00 = A.g[0][0], g10 = A.g[0][1], g20 = A.g[0][2], dT0 = A.dT[0];
    const real g01 = A.g[1][0], g11 = A.g[1][1], g21 = A.g[1][2], dT1 = A.dT[1];
    const real g02 = A.g[2][0], g12 = A.g[2][1], g22 = A.g[2][2], dT2 = A.dT[2];
    const real mu = C[FT] * c.invRe;
    const real dm0 = dT0 * c.invRe, dm1 = dT1 * c.invRe, dm2 = dT2 * c.invRe;
    const real th23 = RC(2.0 / 3.0) * C[FD];
    const real s01 = g01 + g10, s02 = g02 + g20, s12 = g12 + g21;
    const real d00 = RC(2.0) * g00 - th23, d11 = RC(2.0) * g11 - th23, d22 = RC(2.0) * g22 - th23;
    // F_m = mu (lap u_m + (1/3) d_m theta) + sum_d (g_md + g_dm) dmu_d - (2/3) theta dmu_m
    const real F0 = fma(mu, A.lapu[0], fma(d00, dm0, fma(s01, dm1, s02 * dm2)));
    const real F1 = fma(mu, A.lapu[1], fma(s01, dm0, fma(d11, dm1, s12 * dm2)));
    const real F2 = fma(mu, A.lapu[2], fma(s02, dm0, fma(s12, dm1, d22 * dm2)));
    const real work = fma(C[FU], F0, fma(C[FV], F1, C[FW] * F2));
    // dissipation; quirk Q1 (cuda_rhs.cu:175): the y kernel multiplies (dv/dz + dw/dy) by dv/dz where dw/dy is meant
    const real g3y = c.quirk_q1 ? g12 : g21;
    real diss = d00 * g00;
    diss = fma(s01, g10, diss); diss = fma(s02, g20, diss);
    diss = fma(s01, g01, diss); diss = fma(d11, g11, diss); diss = fma(s12, g3y, diss);
    diss = fma(s02, g02, diss); diss = fma(s12, g12, diss); diss = fma(d22, g22, diss);
    rhs[0] = A.r[0];
    rhs[1] = A.r[1] + F0;
    rhs[2] = A.r[2] + F1;
    rhs[3] = A.r[3] + F2;
    // lambda = mu/(Pr Ec) (cuda_main.cu:239): lambda*lap(T) + grad(lambda).grad(T)
    const real heat = fma(mu, A.lapT, fma(dm0, dT0, fma(dm1, dT1, dm2 * dT2)));
    rhs[4] = A.r[4] + fma(mu, diss, fma(c.lamfac, heat, work));
    if (c.forcing) { rhs[3] += fz; rhs[4] = fma(fz, C[FW], rhs[4]); }
}

}  // namespace fast
}  // namespace cudns
