// libcudns: the fused right-hand-side + Runge-Kutta stage kernel for sm_100a.
//
// One launch evaluates the full RHS of (rho, rho u, rho v, rho w, rho E) -- split-form convective fluxes, expanded viscous
// stress, heat flux, pressure gradient, body force, sponge (cuda_rhs.cu:9-396, calc_stress.cu:20-96, sponge.cu:31-41 of the
// reference) -- and applies the Runge-Kutta register update (cuda_main.cu:126-216,244-247) to every point, so that each
// conserved field crosses HBM once in and once out per stage.
//
// Blackwell mapping
//   * one CTA per SM, tile = 32 x TY columns (i,j), marching along z; one thread owns one column;
//   * TMA (cp.async.bulk.tensor, mbarrier complete_tx) stages the raw fields of the next plane -- the tile with its x/y
//     stencil halos for the in-plane terms, and the tile interior S planes ahead for the z stencil -- while the current
//     plane is being computed;
//   * the z stencil needs the 2S+1 most recent planes of 8-9 derived quantities PER COLUMN.  That data is private to the
//     owning thread, so it lives in TENSOR MEMORY: thread <-> TMEM lane, a ring of 2S+1 slots x NQ doubles in the columns
//     (tcgen05.st / tcgen05.ld 32x32b).  The 256 KB of TMEM hold what would otherwise be ~150 KB of shared memory and,
//     more importantly, a third of the shared-memory load traffic of the kernel; shared memory keeps only the current plane
//     (double-buffered, one __syncthreads per plane);
//   * FP64 FMA issue (not HBM) is the nominal limiter at 8th order: the split-form sums are evaluated in telescoped pair
//     form with pre-scaled coefficients (constant bank operands), 1/rho and the viscosity law are evaluated once per
//     point per plane, and the stress/heat terms are assembled once per point from 48 directional stencil sums.
#include "cudns_internal.h"
#include <cstdio>
#include <cstdlib>

namespace cudns {

namespace {

constexpr int TX = 32;
constexpr int CX = TX + 2 * GX;      // row pitch of the halo'd plane (doubles)

// derived quantities carried per point (mu only when the viscosity law is not linear in T)
enum { ZR = 0, ZU, ZV, ZW, ZH, ZP, ZT, ZD, ZM };

template <int S, int TY, int NQ> struct Cfg {
    static constexpr int NT = TX * TY;
    static constexpr int R = 2 * S + 1;
    static constexpr int CY = TY + 2 * S;
    static constexpr int CSZ = CX * CY;                 // doubles per quantity of the halo'd plane
    static constexpr int COLS_SLOT = 2 * NQ;            // TMEM columns (32-bit) per ring slot
    static constexpr int COLS_THREAD = R * COLS_SLOT;
    static constexpr int WPQ = (TY + 3) / 4;            // warps sharing one TMEM lane quadrant
    static_assert(WPQ * COLS_THREAD <= 512, "z ring does not fit tensor memory");
    static constexpr int NEED = WPQ * COLS_THREAD;
    static constexpr int NCOLS = NEED <= 32 ? 32 : NEED <= 64 ? 64 : NEED <= 128 ? 128 : NEED <= 256 ? 256 : 512;   // power of two >= 32
    static constexpr int NH = 2 * S * TY + 2 * S * TX;  // halo cells of one plane (no corners)
    static constexpr int NP = (NQ + 1) / 2;             // quantity pairs of the shared plane (double2 cells)
    static constexpr size_t CUR_D = (size_t)2 * 2 * NP * CSZ;
    static constexpr size_t BOX_D = (size_t)6 * CSZ;
    static constexpr size_t INT_D = (size_t)6 * NT;
    static constexpr size_t PRO_D = (size_t)2 * S * 6 * NT;
    static constexpr size_t MAIN_D = CUR_D + BOX_D + INT_D;
    static constexpr size_t DATA_D = MAIN_D > PRO_D ? MAIN_D : PRO_D;
    static constexpr size_t bytes = DATA_D * sizeof(double) + 64;
    static_assert(bytes <= 227 * 1024, "shared memory budget");
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t cnt) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(cnt) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t mbar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"((uint64_t)map), "r"(mbar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t mbar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"((uint64_t)map), "r"(mbar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// ---- tensor memory as a per-thread ring buffer ------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t holder, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(holder), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// The loads, their wait and the 2x32-bit -> double packing share one asm statement: the compiler cannot schedule a use
// before the wait, and ptxas allocates the 32-bit destinations as the halves of the 64-bit results (no moves).
template <int NQ> __device__ __forceinline__ void tmem_ld_slot(uint32_t ta, double (&q)[NQ]) {
    if constexpr (NQ == 8) {
        asm volatile("{\n\t.reg .b32 a<16>;\n\t"
                     "tcgen05.ld.sync.aligned.32x32b.x16.b32 {a0,a1,a2,a3,a4,a5,a6,a7,a8,a9,a10,a11,a12,a13,a14,a15}, [%8];\n\t"
                     "tcgen05.wait::ld.sync.aligned;\n\t"
                     "mov.b64 %0, {a0,a1};\n\tmov.b64 %1, {a2,a3};\n\tmov.b64 %2, {a4,a5};\n\tmov.b64 %3, {a6,a7};\n\t"
                     "mov.b64 %4, {a8,a9};\n\tmov.b64 %5, {a10,a11};\n\tmov.b64 %6, {a12,a13};\n\tmov.b64 %7, {a14,a15};\n\t}"
                     : "=d"(q[0]), "=d"(q[1]), "=d"(q[2]), "=d"(q[3]), "=d"(q[4]), "=d"(q[5]), "=d"(q[6]), "=d"(q[7])
                     : "r"(ta) : "memory");
    } else {
        asm volatile("{\n\t.reg .b32 a<18>;\n\t"
                     "tcgen05.ld.sync.aligned.32x32b.x16.b32 {a0,a1,a2,a3,a4,a5,a6,a7,a8,a9,a10,a11,a12,a13,a14,a15}, [%9];\n\t"
                     "tcgen05.ld.sync.aligned.32x32b.x2.b32 {a16,a17}, [%10];\n\t"
                     "tcgen05.wait::ld.sync.aligned;\n\t"
                     "mov.b64 %0, {a0,a1};\n\tmov.b64 %1, {a2,a3};\n\tmov.b64 %2, {a4,a5};\n\tmov.b64 %3, {a6,a7};\n\t"
                     "mov.b64 %4, {a8,a9};\n\tmov.b64 %5, {a10,a11};\n\tmov.b64 %6, {a12,a13};\n\tmov.b64 %7, {a14,a15};\n\t"
                     "mov.b64 %8, {a16,a17};\n\t}"
                     : "=d"(q[0]), "=d"(q[1]), "=d"(q[2]), "=d"(q[3]), "=d"(q[4]), "=d"(q[5]), "=d"(q[6]), "=d"(q[7]), "=d"(q[NQ - 1])
                     : "r"(ta), "r"(ta + 16) : "memory");
    }
}
template <int NQ> __device__ __forceinline__ void tmem_ld_pair(uint32_t ta, uint32_t tb, double (&p)[NQ], double (&m)[NQ]) {
    if constexpr (NQ == 8) {
        asm volatile("{\n\t.reg .b32 a<16>;\n\t.reg .b32 b<16>;\n\t"
                     "tcgen05.ld.sync.aligned.32x32b.x16.b32 {a0,a1,a2,a3,a4,a5,a6,a7,a8,a9,a10,a11,a12,a13,a14,a15}, [%16];\n\t"
                     "tcgen05.ld.sync.aligned.32x32b.x16.b32 {b0,b1,b2,b3,b4,b5,b6,b7,b8,b9,b10,b11,b12,b13,b14,b15}, [%17];\n\t"
                     "tcgen05.wait::ld.sync.aligned;\n\t"
                     "mov.b64 %0, {a0,a1};\n\tmov.b64 %1, {a2,a3};\n\tmov.b64 %2, {a4,a5};\n\tmov.b64 %3, {a6,a7};\n\t"
                     "mov.b64 %4, {a8,a9};\n\tmov.b64 %5, {a10,a11};\n\tmov.b64 %6, {a12,a13};\n\tmov.b64 %7, {a14,a15};\n\t"
                     "mov.b64 %8, {b0,b1};\n\tmov.b64 %9, {b2,b3};\n\tmov.b64 %10, {b4,b5};\n\tmov.b64 %11, {b6,b7};\n\t"
                     "mov.b64 %12, {b8,b9};\n\tmov.b64 %13, {b10,b11};\n\tmov.b64 %14, {b12,b13};\n\tmov.b64 %15, {b14,b15};\n\t}"
                     : "=d"(p[0]), "=d"(p[1]), "=d"(p[2]), "=d"(p[3]), "=d"(p[4]), "=d"(p[5]), "=d"(p[6]), "=d"(p[7]),
                       "=d"(m[0]), "=d"(m[1]), "=d"(m[2]), "=d"(m[3]), "=d"(m[4]), "=d"(m[5]), "=d"(m[6]), "=d"(m[7])
                     : "r"(ta), "r"(tb) : "memory");
    } else {
        asm volatile("{\n\t.reg .b32 a<18>;\n\t.reg .b32 b<18>;\n\t"
                     "tcgen05.ld.sync.aligned.32x32b.x16.b32 {a0,a1,a2,a3,a4,a5,a6,a7,a8,a9,a10,a11,a12,a13,a14,a15}, [%18];\n\t"
                     "tcgen05.ld.sync.aligned.32x32b.x16.b32 {b0,b1,b2,b3,b4,b5,b6,b7,b8,b9,b10,b11,b12,b13,b14,b15}, [%19];\n\t"
                     "tcgen05.ld.sync.aligned.32x32b.x2.b32 {a16,a17}, [%20];\n\t"
                     "tcgen05.ld.sync.aligned.32x32b.x2.b32 {b16,b17}, [%21];\n\t"
                     "tcgen05.wait::ld.sync.aligned;\n\t"
                     "mov.b64 %0, {a0,a1};\n\tmov.b64 %1, {a2,a3};\n\tmov.b64 %2, {a4,a5};\n\tmov.b64 %3, {a6,a7};\n\t"
                     "mov.b64 %4, {a8,a9};\n\tmov.b64 %5, {a10,a11};\n\tmov.b64 %6, {a12,a13};\n\tmov.b64 %7, {a14,a15};\n\t"
                     "mov.b64 %8, {a16,a17};\n\t"
                     "mov.b64 %9, {b0,b1};\n\tmov.b64 %10, {b2,b3};\n\tmov.b64 %11, {b4,b5};\n\tmov.b64 %12, {b6,b7};\n\t"
                     "mov.b64 %13, {b8,b9};\n\tmov.b64 %14, {b10,b11};\n\tmov.b64 %15, {b12,b13};\n\tmov.b64 %16, {b14,b15};\n\t"
                     "mov.b64 %17, {b16,b17};\n\t}"
                     : "=d"(p[0]), "=d"(p[1]), "=d"(p[2]), "=d"(p[3]), "=d"(p[4]), "=d"(p[5]), "=d"(p[6]), "=d"(p[7]), "=d"(p[NQ - 1]),
                       "=d"(m[0]), "=d"(m[1]), "=d"(m[2]), "=d"(m[3]), "=d"(m[4]), "=d"(m[5]), "=d"(m[6]), "=d"(m[7]), "=d"(m[NQ - 1])
                     : "r"(ta), "r"(tb), "r"(ta + 16), "r"(tb + 16) : "memory");
    }
}
template <int NQ> __device__ __forceinline__ void tmem_st_slot(uint32_t ta, const double (&q)[NQ]) {
    asm volatile("{\n\t.reg .b32 a<16>;\n\t"
                 "mov.b64 {a0,a1}, %1;\n\tmov.b64 {a2,a3}, %2;\n\tmov.b64 {a4,a5}, %3;\n\tmov.b64 {a6,a7}, %4;\n\t"
                 "mov.b64 {a8,a9}, %5;\n\tmov.b64 {a10,a11}, %6;\n\tmov.b64 {a12,a13}, %7;\n\tmov.b64 {a14,a15}, %8;\n\t"
                 "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {a0,a1,a2,a3,a4,a5,a6,a7,a8,a9,a10,a11,a12,a13,a14,a15};\n\t}"
                 ::"r"(ta), "d"(q[0]), "d"(q[1]), "d"(q[2]), "d"(q[3]), "d"(q[4]), "d"(q[5]), "d"(q[6]), "d"(q[7]) : "memory");
    if constexpr (NQ == 9)
        asm volatile("{\n\t.reg .b32 a<2>;\n\tmov.b64 {a0,a1}, %1;\n\ttcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {a0,a1};\n\t}"
                     ::"r"(ta + 16), "d"(q[NQ - 1]) : "memory");
    tmem_wait_st();
}

__device__ __forceinline__ double visc_law(const KConst &c, double t) {
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
template <int NQ> __device__ __forceinline__ void eos_q(const KConst &c, double r, double u, double v, double w, double e, double th, double (&q)[NQ]) {
    double rinv = 1.0 / r;
    double en = e * rinv - 0.5 * (u * u + v * v + w * w);
    double t = c.cvInv * en;
    double p = r * c.Rgas * t;
    q[ZR] = r; q[ZU] = u; q[ZV] = v; q[ZW] = w;
    q[ZH] = (e + p) * rinv; q[ZP] = p; q[ZT] = t; q[ZD] = th;
    if constexpr (NQ == 9) q[ZM] = visc_law(c, t);
}

// wall blowing/suction, perturbation.h:25-53
__device__ __forceinline__ bool perturb_u(const KConst &c, int j, int kglob, double &val) {
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

// stencil sums of one point, filled direction by direction
struct Sums {
    double g[3][3];       // g[m][d] = d u_m / d x_d   (viscous order)
    double lapu[3];       // sum_d D2_d u_m
    double dT[3], dth[3], dp[3], dmu[3];
    double lapT;
    double rhs[5];        // the convective part accumulates here directly
};

// Accumulate everything direction D contributes.  fetch(l, P, M) delivers the NQ quantities at offsets +l / -l.
// Coefficients are pre-scaled by the grid spacing: cC = -a_l/(4 dx), cP = a_l/dx, c1 = a_l/dx (viscous order), c2 = b_l/dx^2.
// For the stretched wall-normal grid (D == 0, nonuni) the first-derivative sums and the convective sums still lack the
// metric factor xp[i]; the caller applies it right after this call (x is processed first, so rhs holds x terms only).
template <int D, int S, int V, int NQ, class Fetch>
__device__ __forceinline__ void dir_sums(const KConst &c, const double (&C)[NQ], Fetch &&fetch, Sums &A, bool nonuni, int i) {
    const double Uc = C[ZU + D];
    double aM = 0.0;
    const double *tab = c.cVSx + i;        // non-uniform x only: cVSx[it*mx + i]  (cuda_utils.cu:107-122)
    const int mx = c.L.mx;
    {
        double k20 = c.c2[D][0];
        if (D == 0 && nonuni) k20 = tab[(size_t)V * mx];
        A.lapu[0] = fma(k20, C[ZU], A.lapu[0]); A.lapu[1] = fma(k20, C[ZV], A.lapu[1]); A.lapu[2] = fma(k20, C[ZW], A.lapu[2]);
        A.lapT = fma(k20, C[ZT], A.lapT);
    }
    A.g[0][D] = 0.0; A.g[1][D] = 0.0; A.g[2][D] = 0.0; A.dT[D] = 0.0; A.dth[D] = 0.0; A.dp[D] = 0.0; A.dmu[D] = 0.0;
#pragma unroll
    for (int l = 1; l <= S; l++) {
        double Pn[NQ], Mn[NQ];
        fetch(l, Pn, Mn);
        // split-form convective sums: A = -a_l/(4dx) (rho_c + rho_n)(U_c + U_n)
        const double cu = c.cC[D][l] * Uc;
        const double Ap = (C[ZR] + Pn[ZR]) * fma(c.cC[D][l], Pn[ZU + D], cu);
        const double Am = (C[ZR] + Mn[ZR]) * fma(c.cC[D][l], Mn[ZU + D], cu);
        aM += Ap - Am;
        A.rhs[1] = fma(Ap, Pn[ZU], A.rhs[1]); A.rhs[1] = fma(-Am, Mn[ZU], A.rhs[1]);
        A.rhs[2] = fma(Ap, Pn[ZV], A.rhs[2]); A.rhs[2] = fma(-Am, Mn[ZV], A.rhs[2]);
        A.rhs[3] = fma(Ap, Pn[ZW], A.rhs[3]); A.rhs[3] = fma(-Am, Mn[ZW], A.rhs[3]);
        A.rhs[4] = fma(Ap, Pn[ZH], A.rhs[4]); A.rhs[4] = fma(-Am, Mn[ZH], A.rhs[4]);
        A.dp[D] = fma(c.cP[D][l], Pn[ZP] - Mn[ZP], A.dp[D]);
        if (l <= V) {
            double k2 = c.c2[D][l], k2m = k2;
            if (D == 0 && nonuni) { k2 = tab[(size_t)(V + l) * mx]; k2m = tab[(size_t)(V - l) * mx]; }
#pragma unroll
            for (int m = 0; m < 3; m++) {
                A.g[m][D] = fma(c.c1[D][l], Pn[ZU + m] - Mn[ZU + m], A.g[m][D]);
                if (D == 0 && nonuni) { A.lapu[m] = fma(k2, Pn[ZU + m], A.lapu[m]); A.lapu[m] = fma(k2m, Mn[ZU + m], A.lapu[m]); }
                else A.lapu[m] = fma(k2, Pn[ZU + m] + Mn[ZU + m], A.lapu[m]);
            }
            A.dT[D] = fma(c.c1[D][l], Pn[ZT] - Mn[ZT], A.dT[D]);
            if (D == 0 && nonuni) { A.lapT = fma(k2, Pn[ZT], A.lapT); A.lapT = fma(k2m, Mn[ZT], A.lapT); }
            else A.lapT = fma(k2, Pn[ZT] + Mn[ZT], A.lapT);
            A.dth[D] = fma(c.c1[D][l], Pn[ZD] - Mn[ZD], A.dth[D]);
            if constexpr (NQ == 9) A.dmu[D] = fma(c.c1[D][l], Pn[ZM] - Mn[ZM], A.dmu[D]);
        }
    }
    A.rhs[0] = fma(2.0, aM, A.rhs[0]);
    A.rhs[1] = fma(C[ZU], aM, A.rhs[1]);
    A.rhs[2] = fma(C[ZV], aM, A.rhs[2]);
    A.rhs[3] = fma(C[ZW], aM, A.rhs[3]);
    A.rhs[4] = fma(C[ZH], aM, A.rhs[4]);
}

template <int S, int V, int TY, int NQ>
__global__ void __launch_bounds__(TX *TY, 1)
stage_kernel(const __grid_constant__ KConst c, const __grid_constant__ StagePtrs P, const __grid_constant__ StageCoef sc, int zchunk,
             const __grid_constant__ StageMaps tm) {
    using G = Cfg<S, TY, NQ>;
    constexpr int NT = G::NT, R = G::R, CSZ = G::CSZ, NP = G::NP;
    extern __shared__ __align__(1024) double smem[];
    double2 *cur0 = (double2 *)smem;                   // [2][NP][CY][CX] pairs (rho,u) (v,w) (H,p) (T,theta) (mu,-)
    double *boxraw = smem + G::CUR_D;                  // [6][CY][CX]   raw r,u,v,w,e,theta of the current plane with halos
    double *intraw = boxraw + G::BOX_D;                // [6][TY][TX]   raw fields of the plane S ahead (tile interior)
    uint64_t *mbar_p = (uint64_t *)(smem + G::DATA_D);
    uint32_t *tmem_holder = (uint32_t *)(mbar_p + 1);

    const Layout &L = c.L;
    const int tid = threadIdx.x;
    const int tx = tid & 31, ty = tid >> 5;            // warp == tile row
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
    const int nxt = min(TX, L.mx - i0);                // interior columns of this tile
    const bool nonuni = c.nonUniformX != 0;
    const bool bl = c.boundaryLayer != 0;
    const size_t N = (size_t)L.mx * L.my * L.mz;
    const size_t nxy = (size_t)L.mx * L.my;
    const size_t gp0 = L.idx(ic, jc, 0), n0 = (size_t)ic + (size_t)jc * L.mx;
    // periodic images this thread also writes (cross-shaped ghosts): perBCx / perBCy, boundary.h:38-46
    const bool img_xlo = c.periodicX && i < S, img_xhi = c.periodicX && i >= L.mx - S;
    const bool img_ylo = j < S, img_yhi = j >= L.my - S;
    const bool img_any = img_xlo || img_xhi || img_ylo || img_yhi;
    const bool do_update = active && !P.rhs_out;
    const double dt = P.rhs_out ? 0.0 : *c.dt;

    const uint32_t mbar = smem_u32(mbar_p);
    if (ty == 0) tmem_alloc(smem_u32(tmem_holder), G::NCOLS);
    if (tid == 0) mbar_init(mbar, 1);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem0 = *tmem_holder;
    const uint32_t tbase = tmem0 + ((uint32_t)(32 * (ty & 3)) << 16) + (uint32_t)((ty >> 2) * G::COLS_THREAD);
    auto tslot = [&](int kk) -> uint32_t { return tbase + (uint32_t)(((kk + 16 * R) % R) * G::COLS_SLOT); };
    // the same for kk = k + l with |l| <= S, given s0 = k mod R (no division in the plane loop)
    auto tslot_rel = [&](int s0, int l) -> uint32_t { int t = s0 + l; t = t >= R ? t - R : t; t = t < 0 ? t + R : t; return tbase + (uint32_t)(t * G::COLS_SLOT); };

    // raw fields of plane kk at this thread's column (src = [6][NT]) -> EOS -> ring slot
    auto ring_insert = [&](int kk, const double *src, uint32_t slot) {
        double q[NQ];
        if (bl && (kk < kglob_lo || kk >= kglob_hi)) {
            if (kk >= kglob_hi) {
                // topBCzExt (boundary.h:154-156): f[mz-1+g] = 2 f[mz-1] - f[mz-1-g], every staged quantity independently
                const int gq = kk - (kglob_hi - 1);
                double a[NQ], b[NQ];
                tmem_ld_pair<NQ>(tslot(kglob_hi - 1), tslot(kglob_hi - 1 - gq), a, b);
#pragma unroll
                for (int n = 0; n < NQ; n++) q[n] = 2.0 * a[n] - b[n];
                tmem_st_slot<NQ>(slot, q);
            }
            return;      // bottom ghosts are generated once plane kglob_lo+S is in (ring_bottom_ghosts)
        }
        eos_q<NQ>(c, src[tid], src[NT + tid], src[2 * NT + tid], src[3 * NT + tid], src[4 * NT + tid], src[5 * NT + tid], q);
        tmem_st_slot<NQ>(slot, q);
    };
    // botBCzExt (boundary.h:158-160): f[-g] = 2 f[0] - f[g]
    auto ring_bottom_ghosts = [&]() {
#pragma unroll
        for (int gq = 1; gq <= S; gq++) {
            double a[NQ], b[NQ], q[NQ];
            tmem_ld_pair<NQ>(tslot(kglob_lo), tslot(kglob_lo + gq), a, b);
#pragma unroll
            for (int n = 0; n < NQ; n++) q[n] = 2.0 * a[n] - b[n];
            tmem_st_slot<NQ>(tslot(kglob_lo - gq), q);
        }
    };
    auto issue_plane_loads = [&](int k) {          // thread 0: tile interior of plane k+S, halo'd tile of plane k
        mbar_expect_tx(mbar, (uint32_t)((G::INT_D + G::BOX_D) * sizeof(double)));
        tma_load_4d(smem_u32(intraw), &tm.qint, mbar, i0 + GX, j0 + L.gy, k + S + L.gz, 0);
        tma_load_3d(smem_u32(intraw + 5 * NT), &tm.thint, mbar, i0 + GX, j0 + L.gy, k + S + L.gz);
        tma_load_4d(smem_u32(boxraw), &tm.qbox, mbar, i0, j0 + L.gy - S, k + L.gz, 0);
        tma_load_3d(smem_u32(boxraw + 5 * CSZ), &tm.thbox, mbar, i0, j0 + L.gy - S, k + L.gz);
    };
    // quantity q of cell (cy,cx) in a pair-interleaved plane
    auto cell = [&](double2 *pl, int q, int cy, int cx) -> double & { return ((double *)(pl + (q >> 1) * CSZ + cy * CX + cx))[q & 1]; };
    auto store_cell = [&](double2 *pl, int cy, int cx, const double (&q)[NQ]) {
        double2 *d = pl + cy * CX + cx;
#pragma unroll
        for (int n = 0; n < NP; n++) d[n * CSZ] = make_double2(q[2 * n], (2 * n + 1 < NQ) ? q[2 * n + 1] : 0.0);
    };

    // ---- prologue: planes kbeg-S .. kbeg+S-1 into the ring (one batch of TMA loads into the whole buffer)
    uint32_t phase = 0;
    {
        double *pro = smem;
        if (tid == 0) {
            mbar_expect_tx(mbar, (uint32_t)(G::PRO_D * sizeof(double)));
            for (int n = 0; n < 2 * S; n++) {
                tma_load_4d(smem_u32(pro + (size_t)n * 6 * NT), &tm.qint, mbar, i0 + GX, j0 + L.gy, kbeg - S + n + L.gz, 0);
                tma_load_3d(smem_u32(pro + (size_t)n * 6 * NT + 5 * NT), &tm.thint, mbar, i0 + GX, j0 + L.gy, kbeg - S + n + L.gz);
            }
        }
        mbar_wait(mbar, phase); phase ^= 1;
        for (int n = 0; n < 2 * S; n++) ring_insert(kbeg - S + n, pro + (size_t)n * 6 * NT, tslot(kbeg - S + n));
        __syncthreads();
        if (tid == 0) issue_plane_loads(kbeg);
    }

    int s0 = (kbeg + 16 * R) % R;
    for (int k = kbeg; k < kend; k++, s0 = (s0 + 1 == R) ? 0 : s0 + 1) {
        double2 *cur = cur0 + (size_t)((k - kbeg) & 1) * NP * CSZ;
        const size_t gp = gp0 + (size_t)k * L.plane;          // = L.idx(ic,jc,k)
        const size_t n = n0 + (size_t)k * nxy;
        // ---- Runge-Kutta operands of this point: the loads are issued first so that their latency hides behind the whole
        //      plane; they are only consumed in the update at the bottom of the loop body
        double ra[5], rb[5], qv[5], rw[5];
#pragma unroll
        for (int m = 0; m < 5; m++) { ra[m] = 0.0; rb[m] = 0.0; qv[m] = 0.0; rw[m] = 0.0; }
        if (do_update) {
            if (P.RA) {
#pragma unroll
                for (int m = 0; m < 5; m++) ra[m] = P.RA[m * N + n];
            }
            if (P.RB) {
#pragma unroll
                for (int m = 0; m < 5; m++) rb[m] = P.RB[m * N + n];
            }
            if (P.qbase != P.qin) {
#pragma unroll
                for (int m = 0; m < 5; m++) qv[m] = P.qbase[m * vol + gp];
            }
            if (P.RW && sc.wOld != 0.0) {
#pragma unroll
                for (int m = 0; m < 5; m++) rw[m] = P.RW[m * N + n];
            }
        }
        mbar_wait(mbar, phase); phase ^= 1;
        ring_insert(k + S, intraw, tslot_rel(s0, S));
        if (bl && k == kglob_lo) ring_bottom_ghosts();
        // ---- own point of plane k: out of the ring into registers and into the shared plane
        double C[NQ];
        tmem_ld_slot<NQ>(tslot_rel(s0, 0), C);
        const double e_c = boxraw[4 * CSZ + (ty + S) * CX + (tx + GX)];
        store_cell(cur, ty + S, tx + GX, C);
        // ---- halo cells of plane k: raw -> EOS -> shared plane
        for (int cidx = tid; cidx < G::NH; cidx += NT) {
            int cx, cy;
            if (cidx < 2 * S * TY) { int hx = cidx % (2 * S), hy = cidx / (2 * S); cx = hx < S ? GX - S + hx : GX + nxt + (hx - S); cy = S + hy; }
            else { int dd = cidx - 2 * S * TY; int hx = dd % TX, hy = dd / TX; cx = GX + hx; cy = hy < S ? hy : TY + hy; }
            const int gi = i0 + cx - GX;
            if (!c.periodicX && (gi < 0 || gi >= L.mx)) continue;     // wall / extrapolation ghosts are built below
            const double *s = boxraw + cy * CX + cx;
            double q[NQ];
            eos_q<NQ>(c, s[0], s[CSZ], s[2 * CSZ], s[3 * CSZ], s[4 * CSZ], s[5 * CSZ], q);
            store_cell(cur, cy, cx, q);
        }
        __syncthreads();
        if (tid == 0 && k + 1 < kend) issue_plane_loads(k + 1);
        // ---- x boundary rules on the shared plane (boundary_condition_x.h BCxNumber1-3)
        if (xlo_tile || xhi_tile) {
            for (int cidx = tid; cidx < 2 * S * TY; cidx += NT) {
                int gq = cidx % S + 1, side = (cidx / S) & 1, row = cidx / (2 * S);
                if (side == 0 && !xlo_tile) continue;
                if (side == 1 && !xhi_tile) continue;
                if (j0 + row >= L.my) continue;
                const int cy = row + S;
                int cg, cm;             // ghost column, mirror column
                double u, v, w, p, t, th;
                if (side == 0) {
                    cg = GX - gq; cm = GX + gq - 1;   // cell mirror: f[-g] <- f[g-1]
                    u = -cell(cur, ZU, cy, cm); v = -cell(cur, ZV, cy, cm); w = -cell(cur, ZW, cy, cm);   // wallBCxVel / botBCxExt(.,0)
                    p = cell(cur, ZP, cy, cm);                                                            // wallBCxMir / botBCxMir
                    th = cell(cur, ZD, cy, cm);                                                           // BCxNumber2
                    if (bl) {
                        t = cell(cur, ZT, cy, cm);                                                        // botBCxMir (adiabatic)
                        double pv;
                        if (c.perturbed && perturb_u(c, j0 + row, k + c.kstart, pv)) u = pv;               // PerturbUvel
                    } else {
                        t = 2.0 * c.TwallBot - cell(cur, ZT, cy, cm);                                     // wallBCxExt
                    }
                } else {
                    int last = GX + nxt - 1;
                    cg = last + gq;
                    if (bl) {
                        cm = last - gq;         // node extrapolation topBCxExt: f[mx-1+g] = 2 f[mx-1] - f[mx-1-g]
                        u = 2.0 * cell(cur, ZU, cy, last) - cell(cur, ZU, cy, cm);
                        v = 2.0 * cell(cur, ZV, cy, last) - cell(cur, ZV, cy, cm);
                        w = 2.0 * cell(cur, ZW, cy, last) - cell(cur, ZW, cy, cm);
                        p = 2.0 * cell(cur, ZP, cy, last) - cell(cur, ZP, cy, cm);
                        t = 2.0 * cell(cur, ZT, cy, last) - cell(cur, ZT, cy, cm);
                        th = 2.0 * cell(cur, ZD, cy, last) - cell(cur, ZD, cy, cm);
                    } else {
                        cm = last - gq + 1;
                        u = -cell(cur, ZU, cy, cm); v = -cell(cur, ZV, cy, cm); w = -cell(cur, ZW, cy, cm);
                        p = cell(cur, ZP, cy, cm);
                        th = cell(cur, ZD, cy, cm);
                        t = 2.0 * c.TwallTop - cell(cur, ZT, cy, cm);
                    }
                }
                double q[NQ];
                q[ZU] = u; q[ZV] = v; q[ZW] = w; q[ZP] = p; q[ZT] = t; q[ZD] = th;
                if constexpr (NQ == 9) q[NQ - 1] = visc_law(c, t);                                      // mlBoundPT boundary.h:135
                q[ZH] = t * c.Rgas * c.gam / (c.gam - 1.0) + 0.5 * (u * u + v * v + w * w);             // rhBoundPT boundary.h:121
                q[ZR] = p / (c.Rgas * t);
                store_cell(cur, cy, cg, q);
            }
            __syncthreads();
        }

        // ---- the 48 directional stencil sums of point (i,j,k)
        const double2 *curc = cur + (ty + S) * CX + (tx + GX);
        Sums A;
        A.lapu[0] = A.lapu[1] = A.lapu[2] = 0.0; A.lapT = 0.0;
#pragma unroll
        for (int m = 0; m < 5; m++) A.rhs[m] = 0.0;
        dir_sums<0, S, V, NQ>(c, C, [&](int l, double (&Pn)[NQ], double (&Mn)[NQ]) {
#pragma unroll
            for (int np = 0; np < NP; np++) {
                if (np < 3 || l <= V) {
                    const double2 a = curc[np * CSZ + l], b = curc[np * CSZ - l];
                    Pn[2 * np] = a.x; Mn[2 * np] = b.x;
                    if (2 * np + 1 < NQ) { Pn[2 * np + 1] = a.y; Mn[2 * np + 1] = b.y; }
                } else {
                    Pn[2 * np] = 0.0; Mn[2 * np] = 0.0;
                    if (2 * np + 1 < NQ) { Pn[2 * np + 1] = 0.0; Mn[2 * np + 1] = 0.0; }
                }
            }
        }, A, nonuni, ic);
        if (nonuni) {      // metric of the stretched wall-normal grid (cuda_derivs.h:46-48,166-168,200-202)
            const double xpi = c.xp[ic];
            A.g[0][0] *= xpi; A.g[1][0] *= xpi; A.g[2][0] *= xpi; A.dT[0] *= xpi; A.dth[0] *= xpi; A.dp[0] *= xpi; A.dmu[0] *= xpi;
#pragma unroll
            for (int m = 0; m < 5; m++) A.rhs[m] *= xpi;
        }
        dir_sums<1, S, V, NQ>(c, C, [&](int l, double (&Pn)[NQ], double (&Mn)[NQ]) {
#pragma unroll
            for (int np = 0; np < NP; np++) {
                if (np < 3 || l <= V) {
                    const double2 a = curc[np * CSZ + l * CX], b = curc[np * CSZ - l * CX];
                    Pn[2 * np] = a.x; Mn[2 * np] = b.x;
                    if (2 * np + 1 < NQ) { Pn[2 * np + 1] = a.y; Mn[2 * np + 1] = b.y; }
                } else {
                    Pn[2 * np] = 0.0; Mn[2 * np] = 0.0;
                    if (2 * np + 1 < NQ) { Pn[2 * np + 1] = 0.0; Mn[2 * np + 1] = 0.0; }
                }
            }
        }, A, false, ic);
        dir_sums<2, S, V, NQ>(c, C, [&](int l, double (&Pn)[NQ], double (&Mn)[NQ]) {
            tmem_ld_pair<NQ>(tslot_rel(s0, l), tslot_rel(s0, -l), Pn, Mn);
        }, A, false, ic);

        // ---- stress, dissipation, heat flux, pressure gradient: assembled once per point (cuda_rhs.cu:52-127,169-259,303-393)
        const double mu = (NQ == 9) ? C[NQ - 1] : C[ZT] * c.invRe;
        double dmu[3];
#pragma unroll
        for (int d = 0; d < 3; d++) dmu[d] = (NQ == 9) ? A.dmu[d] : A.dT[d] * c.invRe;
        const double th = C[ZD];
        const double vel[3] = {C[ZU], C[ZV], C[ZW]};
        double diss = 0.0, work = 0.0;
        double F[3];
#pragma unroll
        for (int m = 0; m < 3; m++) {
            double f = mu * A.lapu[m] + (mu * (1.0 / 3.0)) * A.dth[m] - ((2.0 / 3.0) * th) * dmu[m];
#pragma unroll
            for (int d = 0; d < 3; d++) f = fma(A.g[m][d] + A.g[d][m], dmu[d], f);
            F[m] = f;
            work = fma(vel[m], f, work);
        }
#pragma unroll
        for (int d = 0; d < 3; d++) {
            double tmp[3];
#pragma unroll
            for (int m = 0; m < 3; m++) tmp[m] = (m == d) ? (2.0 * A.g[m][d] - (2.0 / 3.0) * th) : (A.g[m][d] + A.g[d][m]);
            // quirk Q1 (cuda_rhs.cu:175): the y kernel multiplies by dv/dz where dw/dy is meant
            const double g3 = (d == 1 && c.quirk_q1) ? A.g[1][2] : A.g[2][d];
            diss += tmp[0] * A.g[0][d] + tmp[1] * A.g[1][d] + tmp[2] * g3;
        }
        double rhs[5];
        rhs[0] = A.rhs[0];
        rhs[1] = A.rhs[1] + F[0] - A.dp[0];
        rhs[2] = A.rhs[2] + F[1] - A.dp[1];
        rhs[3] = A.rhs[3] + F[2] - A.dp[2];
        // lambda = mu/(Pr Ec) (cuda_main.cu:239): lambda*lap(T) + grad(lambda).grad(T)
        rhs[4] = A.rhs[4] + mu * diss + work + c.lamfac * (mu * A.lapT + dmu[0] * A.dT[0] + dmu[1] * A.dT[1] + dmu[2] * A.dT[2]);
        if (c.forcing) {                       // cuda_rhs.cu:392-393
            double f = *c.dpdz;
            rhs[3] += f; rhs[4] += f * C[ZW];
        }
        if (bl && c.spongeX) {                 // addSponge, sponge.cu:31-41
            double sg = c.spongeX[ic] + c.spongeZ[k];
            size_t nq = (size_t)L.mx * L.mz, qi = (size_t)ic + (size_t)k * L.mx;
            rhs[0] += sg * (c.sref[qi] - C[ZR]);
            rhs[1] += sg * (c.sref[nq + qi] - C[ZR] * C[ZU]);
            rhs[2] += sg * (c.sref[2 * nq + qi] - C[ZR] * C[ZV]);
            rhs[3] += sg * (c.sref[3 * nq + qi] - C[ZR] * C[ZW]);
            rhs[4] += sg * (c.sref[4 * nq + qi] - e_c);
        }
        if (active) {
            if (P.rhs_out) {
#pragma unroll
                for (int m = 0; m < 5; m++) P.rhs_out[m * N + n] = rhs[m];
            } else {
                // Runge-Kutta register update (sumLowStorageRK3 cuda_main.cu:244, eulerSum*/rk3final* :188-216):
                //   Q_out = Q_base + dt (cN K + cA RA + cB RB),   RW = wOld RW + wNew K
                double qb[5];
                if (P.qbase == P.qin) { qb[0] = C[ZR]; qb[1] = C[ZR] * C[ZU]; qb[2] = C[ZR] * C[ZV]; qb[3] = C[ZR] * C[ZW]; qb[4] = e_c; }
                else { qb[0] = qv[0]; qb[1] = qv[0] * qv[1]; qb[2] = qv[0] * qv[2]; qb[3] = qv[0] * qv[3]; qb[4] = qv[4]; }
                double qn[5];
#pragma unroll
                for (int m = 0; m < 5; m++) {
                    const double k_all = fma(sc.cN, rhs[m], fma(sc.cA, ra[m], sc.cB * rb[m]));
                    qn[m] = fma(dt, k_all, qb[m]);
                    if (P.RW) P.RW[m * N + n] = fma(sc.wNew, rhs[m], sc.wOld * rw[m]);
                }
                const double rn = 1.0 / qn[0];                                                   // deviceDiv cuda_math.cu:36
                const double out[5] = {qn[0], qn[1] * rn, qn[2] * rn, qn[3] * rn, qn[4]};
                double *f = P.qout + gp;
#pragma unroll
                for (int m = 0; m < 5; m++) f[m * vol] = out[m];
                if (img_any) {
#pragma unroll
                    for (int m = 0; m < 5; m++) {
                        double *fm = f + m * vol;
                        if (img_xlo) fm[L.mx] = out[m];
                        if (img_xhi) fm[-(ptrdiff_t)L.mx] = out[m];
                        if (img_ylo) fm[(size_t)L.my * L.px] = out[m];
                        if (img_yhi) fm[-(ptrdiff_t)((size_t)L.my * L.px)] = out[m];
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (ty == 0) tmem_dealloc(tmem0, G::NCOLS);
}

template <int S, int V, int TY, int NQ>
void launch_t(const KConst &kc, const StagePtrs &p, const StageCoef &c, const StageMaps &maps, cudaStream_t st) {
    using G = Cfg<S, TY, NQ>;
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(stage_kernel<S, V, TY, NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::bytes); attr_set = true; }
    const int gx = (kc.L.mx + TX - 1) / TX, gy = (kc.L.my + TY - 1) / TY;
    // z chunks: every chunk pays a 2S-plane prologue, so keep them >= 32 planes; more chunks smooth the tail over 148 SMs
    const int cols = gx * gy;
    int nzc = 1;
    while (cols * nzc < 148 * 16 && kc.L.mz / (nzc * 2) >= 32) nzc *= 2;
    int zchunk = (kc.L.mz + nzc - 1) / nzc;
    nzc = (kc.L.mz + zchunk - 1) / zchunk;
    dim3 grid(gx, gy, nzc);
    stage_kernel<S, V, TY, NQ><<<grid, TX * TY, G::bytes, st>>>(kc, p, c, zchunk, maps);
}

template <int TY, int NQ>
void launch_sv(const KConst &kc, const StagePtrs &p, const StageCoef &c, const StageMaps &maps, cudaStream_t st) {
    switch (kc.s * 10 + kc.v) {
        case 11: launch_t<1, 1, TY, NQ>(kc, p, c, maps, st); break;
        case 21: launch_t<2, 1, TY, NQ>(kc, p, c, maps, st); break;
        case 22: launch_t<2, 2, TY, NQ>(kc, p, c, maps, st); break;
        case 31: launch_t<3, 1, TY, NQ>(kc, p, c, maps, st); break;
        case 32: launch_t<3, 2, TY, NQ>(kc, p, c, maps, st); break;
        case 33: launch_t<3, 3, TY, NQ>(kc, p, c, maps, st); break;
        case 41: launch_t<4, 1, TY, NQ>(kc, p, c, maps, st); break;
        case 42: launch_t<4, 2, TY, NQ>(kc, p, c, maps, st); break;
        case 43: launch_t<4, 3, TY, NQ>(kc, p, c, maps, st); break;
        case 44: launch_t<4, 4, TY, NQ>(kc, p, c, maps, st); break;
        default: break;
    }
}

}  // namespace

int stage_tile_y() { return STAGE_TY; }
int stage_smem_bytes(int s, bool linear_visc) {
    switch (s) {
        case 1: return (int)(linear_visc ? Cfg<1, STAGE_TY, 8>::bytes : Cfg<1, STAGE_TY, 9>::bytes);
        case 2: return (int)(linear_visc ? Cfg<2, STAGE_TY, 8>::bytes : Cfg<2, STAGE_TY, 9>::bytes);
        case 3: return (int)(linear_visc ? Cfg<3, STAGE_TY, 8>::bytes : Cfg<3, STAGE_TY, 9>::bytes);
        default: return (int)(linear_visc ? Cfg<4, STAGE_TY, 8>::bytes : Cfg<4, STAGE_TY, 9>::bytes);
    }
}

void launch_rhs_stage(const KConst &kc, const StagePtrs &p, const StageCoef &c, const StageMaps &maps, cudaStream_t st) {
    if (kc.viscmode == 1) launch_sv<STAGE_TY, 8>(kc, p, c, maps, st);
    else launch_sv<STAGE_TY, 9>(kc, p, c, maps, st);
}

}  // namespace cudns
