// libcudns: per-point arithmetic shared by the fourth- and fifth-generation stage kernels (stage_fast.cu, stage_duo.inc): the
// quantities a point carries, the equation of state of a stored point, the running sums and the neighbour-pair step of the
// split-form / expanded-viscous right-hand side.  Algebra: see stage_lean.inc.  Reference: cuda_rhs.cu:9-396,
// cuda_derivs.h:30-155, cuda_main.cu:218-242.
#pragma once
#include "cudns_internal.h"

namespace cudns {
namespace fast {

enum { FR = 0, FU, FV, FW, FH, FT, FD, NF };   // ring / shared-plane quantities

#ifdef CUDNS_F32
// Two x-adjacent points as ONE packed value.  Blackwell executes FP32 pairs in one instruction (FFMA2 / FADD2 / FMUL2: sm_100
// __ffma2_rn, __fadd2_rn, __fmul2_rn; a scalar factor is broadcast from a uniform register for free, a negated operand is an
// instruction modifier), so the single-precision stage kernel -- which is bound by instruction issue, not by the FP32 pipe
// (profiles/r02_duo_f32_512_ncu_full.txt: 77 % of the issue slots busy) -- evaluates both points of a thread with half the
// arithmetic instructions.  The per-point templates below are instantiated with T = P2 there and with T = real elsewhere.
using ::fma;                // (the overloads below would otherwise hide the scalar one inside this namespace)
struct P2 { float2 v; };
__device__ __forceinline__ P2 p2(float x, float y) { P2 r; r.v = make_float2(x, y); return r; }
__device__ __forceinline__ P2 p2(float2 a) { P2 r; r.v = a; return r; }
__device__ __forceinline__ P2 bc(float s) { return p2(s, s); }
__device__ __forceinline__ P2 operator+(P2 a, P2 b) { return p2(__fadd2_rn(a.v, b.v)); }
__device__ __forceinline__ P2 operator-(P2 a) { return p2(-a.v.x, -a.v.y); }
__device__ __forceinline__ P2 operator-(P2 a, P2 b) { return p2(__ffma2_rn(b.v, make_float2(-1.f, -1.f), a.v)); }
__device__ __forceinline__ P2 operator*(P2 a, P2 b) { return p2(__fmul2_rn(a.v, b.v)); }
__device__ __forceinline__ P2 operator*(float s, P2 a) { return p2(__fmul2_rn(make_float2(s, s), a.v)); }
__device__ __forceinline__ P2 operator*(P2 a, float s) { return s * a; }
__device__ __forceinline__ P2 operator+(P2 a, float s) { return p2(__fadd2_rn(a.v, make_float2(s, s))); }
__device__ __forceinline__ P2 fma(P2 a, P2 b, P2 c) { return p2(__ffma2_rn(a.v, b.v, c.v)); }
__device__ __forceinline__ P2 fma(float s, P2 b, P2 c) { return p2(__ffma2_rn(make_float2(s, s), b.v, c.v)); }
__device__ __forceinline__ P2 fma(P2 a, float s, P2 c) { return fma(s, a, c); }
template <typename T> __device__ __forceinline__ T zero_of();
template <> __device__ __forceinline__ P2 zero_of<P2>() { return p2(0.f, 0.f); }
template <> __device__ __forceinline__ float zero_of<float>() { return 0.f; }
#else
template <typename T> __device__ __forceinline__ T zero_of() { return T(0); }
#endif

// H and T of a stored point: calcState (cuda_main.cu:218-242) restricted to what the kernels stage.  ONE definition for the update
// of the stage kernel and for derive_aux_kernel (same operation order -> same bits as eos_q / eos7 of the older kernels)
template <typename T_>
__device__ __forceinline__ void eos_ht(const KConst &c, T_ r, T_ rinv, T_ u, T_ v, T_ w, T_ e, T_ &H, T_ &T) {
    const T_ en = fma(e, rinv, RC(-0.5) * fma(u, u, fma(v, v, w * w)));
    const T_ t = c.cvInv * en;
    const T_ p = (c.Rgas * r) * t;
    H = (e + p) * rinv; T = t;
}


// running sums of one point (T = real) or of a thread's two points (T = P2)
template <typename T> struct AccT {
    T r[5];                    // convective sums; the momentum entries also collect -dp/dx_d
    T lapu[3], lapT;           // sum_d D2_d u_m + (1/3) d theta / d x_m ; sum_d D2_d T
    T g[3][3];                 // g[d][m] = d u_m / d x_d
    T dT[3];
};
using Acc = AccT<real>;

// both neighbours of direction D at offset l: split-form convective sums in telescoped pair form, pressure gradient from rho*T,
// viscous-order first and second differences (dir_sums of stage_lean.inc with PRT)
template <int D, int V, typename T>
__device__ __forceinline__ void pair_step(const KConst &c, const int l, const T (&C)[NF], const T (&Pn)[NF], const T (&Mn)[NF], AccT<T> &A, T &aM) {
    const real cC = c.cf[D][l][0];
    const T cu = cC * C[FU + D];
    const T Ap = (C[FR] + Pn[FR]) * fma(cC, Pn[FU + D], cu);
    const T Am = (C[FR] + Mn[FR]) * fma(cC, Mn[FU + D], cu);
    aM = aM + (Ap - Am);
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
// one neighbour of direction D at offset +l (PLUS) or -l: the same sums, one side at a time (one more arithmetic instruction per
// pair, half the registers for neighbour values)
template <int D, int V, bool PLUS, typename T>
__device__ __forceinline__ void side_step(const KConst &c, const int l, const T (&C)[NF], const T (&Nq)[NF], AccT<T> &A, T &aM) {
    const real cC = c.cf[D][l][0];
    const T cu = cC * C[FU + D];
    const T Af = (C[FR] + Nq[FR]) * fma(cC, Nq[FU + D], cu);
    const T pn = Nq[FR] * Nq[FT];
    const T sA = PLUS ? Af : -Af;
    aM = aM + sA;
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
template <typename T>
__device__ __forceinline__ void close_dir(const T (&C)[NF], AccT<T> &A, const T aM) {
    A.r[0] = fma(RC(2.0), aM, A.r[0]);
    A.r[1] = fma(C[FU], aM, A.r[1]); A.r[2] = fma(C[FV], aM, A.r[2]); A.r[3] = fma(C[FW], aM, A.r[3]); A.r[4] = fma(C[FH], aM, A.r[4]);
}


// stress, dissipation, heat flux assembled once per point (cuda_rhs.cu:52-127,169-259,303-393); g[d][m] = d u_m / d x_d; fz = body
// force dpdz (cuda_rhs.cu:392-393), 0 when there is none
template <typename T>
__device__ __forceinline__ void assemble_rhs(const KConst &c, const T (&C)[NF], const AccT<T> &A, const real fz, T (&rhs)[5]) {
    const T g00 = A.g[0][0], g10 = A.g[0][1], g20 = A.g[0][2], dT0 = A.dT[0];
    const T g01 = A.g[1][0], g11 = A.g[1][1], g21 = A.g[1][2], dT1 = A.dT[1];
    const T g02 = A.g[2][0], g12 = A.g[2][1], g22 = A.g[2][2], dT2 = A.dT[2];
    const T mu = c.invRe * C[FT];
    const T dm0 = c.invRe * dT0, dm1 = c.invRe * dT1, dm2 = c.invRe * dT2;
    const T th23 = RC(2.0 / 3.0) * C[FD];
    const T s01 = g01 + g10, s02 = g02 + g20, s12 = g12 + g21;
    const T d00 = RC(2.0) * g00 - th23, d11 = RC(2.0) * g11 - th23, d22 = RC(2.0) * g22 - th23;
    // F_m = mu (lap u_m + (1/3) d_m theta) + sum_d (g_md + g_dm) dmu_d - (2/3) theta dmu_m
    const T F0 = fma(mu, A.lapu[0], fma(d00, dm0, fma(s01, dm1, s02 * dm2)));
    const T F1 = fma(mu, A.lapu[1], fma(s01, dm0, fma(d11, dm1, s12 * dm2)));
    const T F2 = fma(mu, A.lapu[2], fma(s02, dm0, fma(s12, dm1, d22 * dm2)));
    const T work = fma(C[FU], F0, fma(C[FV], F1, C[FW] * F2));
    // dissipation; quirk Q1 (cuda_rhs.cu:175): the y kernel multiplies (dv/dz + dw/dy) by dv/dz where dw/dy is meant
    const T g3y = c.quirk_q1 ? g12 : g21;
    T diss = d00 * g00;
    diss = fma(s01, g10, diss); diss = fma(s02, g20, diss);
    diss = fma(s01, g01, diss); diss = fma(d11, g11, diss); diss = fma(s12, g3y, diss);
    diss = fma(s02, g02, diss); diss = fma(s12, g12, diss); diss = fma(d22, g22, diss);
    rhs[0] = A.r[0];
    rhs[1] = A.r[1] + F0;
    rhs[2] = A.r[2] + F1;
    rhs[3] = A.r[3] + F2;
    // lambda = mu/(Pr Ec) (cuda_main.cu:239): lambda*lap(T) + grad(lambda).grad(T)
    const T heat = fma(mu, A.lapT, fma(dm0, dT0, fma(dm1, dT1, dm2 * dT2)));
    rhs[4] = A.r[4] + fma(mu, diss, fma(c.lamfac, heat, work));
    if (c.forcing) { rhs[3] = rhs[3] + fz; rhs[4] = fma(fz, C[FW], rhs[4]); }
}

}  // namespace fast
}  // namespace cudns
