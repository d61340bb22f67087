// libcudns: the fused right-hand-side + Runge-Kutta stage kernel for sm_100a, fourth generation ("fast").
//
// Same contract as lean::stage_kernel<S, V, 16, 8, false> (stage_lean.inc: periodic x, uniform grid, linear viscosity law, at most
// the RA operand tile -- every low-storage RK3 stage of the Taylor-Green / periodic-box cases) and the same Blackwell building
// blocks (tile of 32 x TY columns marching along z, one thread per column, TMA-staged planes, the z stencil's 2S+1 planes of
// (rho,u,v,w,H,T,theta) per column as a thread-private ring in tensor memory).  What changed, and why (ncu source-level sampling of
// the wide lean kernel and of two intermediate designs, profiles/r01_stage_fast_*: the x/y/z stencil sums run at ~75 % of the FP64
// peak but take only half of the plane time; the other half were latency chains -- raw -> equation of state (reciprocal) -> shared
// plane for the own and the halo cells, two CTA-wide barriers per plane, register spills reloaded from L2):
//   * THE STATE CARRIES ITS DERIVED FIELDS.  A state buffer has 8 padded fields: rho,u,v,w,rho*E and H,T (written by the update of
//     the stage that produced the state, ghosts included) and theta (dilatation pass).  TMA lands the plane with its x/y halos in
//     exactly the layout the stencil reads ([7][rows][cols], LDS.64 per quantity): there is no staging phase, no halo-cell equation
//     of state, no shared "current plane" and NO CTA-wide barrier in the plane loop -- warps only wait for TMA completion
//     (mbarrier), and one elected thread waits for "all warps have taken plane k out of the single-buffered ring tile" before it
//     issues the loads of plane k+1 (double-buffered), so the warps of a CTA drift up to a plane apart;
//   * TY = 8: two CTAs of 8 warps per SM (<= 112 KB shared memory, 256 tensor-memory columns each);
//   * 7 quantities, not 8: p is rebuilt as Rgas*rho*T in all three directions;
//   * the stencil loops are fully unrolled: every coefficient is a constant-bank operand, the ring-slot addresses of the plane
//     rotate through uniform registers (no modulo arithmetic, no LDC in the loops);
//   * dt and the body force are read once per CTA, not once per plane.
// H and T are functions of the stored (rho,u,v,w,rho*E): the update evaluates the same expression (eos_ht) the staging of the older
// kernels evaluated on load, so the results are bit-identical to theirs; derive_aux_kernel rebuilds H,T of a buffer that was not
// written by this kernel (cudns_set_state, the lean kernels).
// Reference: cuda_rhs.cu:9-396, calc_stress.cu:20-96, cuda_main.cu:126-216,218-247 (see stage_lean.inc for the algebra).
#define LEAN_TY8 CUDNS_LEAN_TY_LINEAR
#define LEAN_TY9 8
#include "stage_lean.inc"
#include "stage_point.h"

#ifndef FAST_L2_HINTS
#define FAST_L2_HINTS 0       // 1: L2 eviction hints on the TMA loads + streaming stores (measured: no gain, 143 vs 149 B/pt read)
#endif
#ifndef FAST_LAST_ISSUES
#define FAST_LAST_ISSUES 0    // 1: the last warp to empty the ring tile sends the next plane bundle (instead of thread 0 after its z sums)
#endif
#ifndef FAST_SPLIT_BAR
#define FAST_SPLIT_BAR 0      // 1: the ring tile (needed at the top of a plane) completes on a barrier of its own and is loaded first;
                              //    the halo'd plane (needed after the z sums) is waited for later
#endif
#ifndef FAST_ONESIDED
#define FAST_ONESIDED 0       // 1: one neighbour at a time (fewer registers, one more FP64 instruction per neighbour pair)
#endif

namespace cudns {
namespace fast {

using namespace lean;


template <int S, int TY_> struct FCfg {
    static constexpr int TY = TY_;                              // tile rows = warps per CTA: 16 (one CTA per SM) or 8 (two)
    static constexpr int NT = TX * TY;
    static constexpr int R = 2 * S + 1;
    static constexpr int CY = TY + 2 * S;
    static constexpr int CSZ = CX * CY;
    static constexpr int COLS_SLOT = 2 * NF;                    // 14 tensor-memory columns per ring slot
    static constexpr int COLS_THREAD = R * COLS_SLOT + 2;       // +2: the 16-column load of the last slot stays inside
    static constexpr int WPQ = TY / 4;                          // warps sharing a lane quadrant
    static constexpr int NEED = WPQ * COLS_THREAD;
    static constexpr int NCOLS = NEED <= 32 ? 32 : NEED <= 64 ? 64 : NEED <= 128 ? 128 : NEED <= 256 ? 256 : 512;
    static_assert(NEED <= 512 && (TY == 16 || NEED <= 256), "z ring does not fit tensor memory");
    // one plane bundle = halo'd tile of (rho,u,v,w,H,T,theta), rho*E and the Runge-Kutta operand of the tile interior (double-
    // buffered), and the tile interior of the plane S ahead for the ring (single-buffered)
    static constexpr size_t BOX_D = (size_t)NF * CSZ;
    static constexpr size_t E_D = (size_t)NT;
    static constexpr size_t OP_D = (size_t)5 * NT;
    static constexpr size_t BUF_D = BOX_D + E_D + OP_D;
    static constexpr size_t INT_D = (size_t)NF * NT;
    static constexpr size_t MAIN_D = 2 * BUF_D + INT_D;
    static constexpr int PRO_PLANES = S;                        // prologue batch (planes of NF interior tiles), two batches
    static constexpr size_t PRO_D = (size_t)PRO_PLANES * NF * NT;
    static constexpr size_t DATA_D = MAIN_D > PRO_D ? MAIN_D : PRO_D;
    static constexpr size_t bytes = DATA_D * sizeof(double) + 64;
    static_assert(bytes <= (TY == 16 ? 227 : 112) * 1024, "shared memory budget");
    static_assert((CSZ * 8) % 128 == 0 && (NT * 8) % 128 == 0, "TMA destinations must stay 128-byte aligned");
};

// TMA loads with an L2 eviction hint.  The stage streams ~250 B per point through L2 but re-reads only the tile interior it
// loaded S planes ahead for the ring (as part of the halo'd plane): that window survives in L2 only if everything else -- the
// halo'd planes themselves, rho*E, the Runge-Kutta operand, all stores -- is marked evict-first
constexpr uint64_t L2_EVICT_FIRST = 0x12F0000000000000ull, L2_EVICT_LAST = 0x14F0000000000000ull;
__device__ __forceinline__ void tma_load_4d_hint(uint32_t dst, const CUtensorMap *map, uint32_t mbar, int c0, int c1, int c2, int c3, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
                 ::"r"(dst), "l"((uint64_t)map), "r"(mbar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_load_3d_hint(uint32_t dst, const CUtensorMap *map, uint32_t mbar, int c0, int c1, int c2, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
                 ::"r"(dst), "l"((uint64_t)map), "r"(mbar), "r"(c0), "r"(c1), "r"(c2), "l"(pol) : "memory");
}

// tensor-memory load of ring slots, split into issue and wait so that independent work can sit in between; the wait takes the
// destination registers as read-write operands, which keeps every use behind it
struct Slot { uint32_t a[16]; };
__device__ __forceinline__ void slot_issue(uint32_t ta, Slot &s) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(s.a[0]), "=r"(s.a[1]), "=r"(s.a[2]), "=r"(s.a[3]), "=r"(s.a[4]), "=r"(s.a[5]), "=r"(s.a[6]), "=r"(s.a[7]),
                   "=r"(s.a[8]), "=r"(s.a[9]), "=r"(s.a[10]), "=r"(s.a[11]), "=r"(s.a[12]), "=r"(s.a[13]), "=r"(s.a[14]), "=r"(s.a[15])
                 : "r"(ta) : "memory");
}
__device__ __forceinline__ void slot_unpack(const Slot &s, double (&q)[NF]) {
#pragma unroll
    for (int n = 0; n < NF; n++) q[n] = __hiloint2double((int)s.a[2 * n + 1], (int)s.a[2 * n]);
}
__device__ __forceinline__ void slot_wait(Slot &s, double (&q)[NF]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(s.a[0]), "+r"(s.a[1]), "+r"(s.a[2]), "+r"(s.a[3]), "+r"(s.a[4]), "+r"(s.a[5]), "+r"(s.a[6]), "+r"(s.a[7]),
                   "+r"(s.a[8]), "+r"(s.a[9]), "+r"(s.a[10]), "+r"(s.a[11]), "+r"(s.a[12]), "+r"(s.a[13])
                 :: "memory");
    slot_unpack(s, q);
}
__device__ __forceinline__ void slot_wait2(Slot &s, Slot &t, double (&q)[NF], double (&w)[NF]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(s.a[0]), "+r"(s.a[1]), "+r"(s.a[2]), "+r"(s.a[3]), "+r"(s.a[4]), "+r"(s.a[5]), "+r"(s.a[6]), "+r"(s.a[7]),
                   "+r"(s.a[8]), "+r"(s.a[9]), "+r"(s.a[10]), "+r"(s.a[11]), "+r"(s.a[12]), "+r"(s.a[13]),
                   "+r"(t.a[0]), "+r"(t.a[1]), "+r"(t.a[2]), "+r"(t.a[3]), "+r"(t.a[4]), "+r"(t.a[5]), "+r"(t.a[6]), "+r"(t.a[7]),
                   "+r"(t.a[8]), "+r"(t.a[9]), "+r"(t.a[10]), "+r"(t.a[11]), "+r"(t.a[12]), "+r"(t.a[13])
                 :: "memory");
    slot_unpack(s, q); slot_unpack(t, w);
}
// ring slot store without the wait (tmem_st7d of stage_lean.inc waits at once): the slot is read back only by the last z step of
// the plane, tmem_wait_st() goes in front of that load
__device__ __forceinline__ void tmem_st7d_nowait(uint32_t ta, double a, double b, double c, double d, double e, double f, double g) {
    asm volatile("{\n\t.reg .b32 x<14>;\n\tmov.b64 {x0,x1}, %1;\n\tmov.b64 {x2,x3}, %2;\n\tmov.b64 {x4,x5}, %3;\n\tmov.b64 {x6,x7}, %4;\n\t"
                 "mov.b64 {x8,x9}, %5;\n\tmov.b64 {x10,x11}, %6;\n\tmov.b64 {x12,x13}, %7;\n\t"
                 "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {x0,x1,x2,x3,x4,x5,x6,x7};\n\t"
                 "tcgen05.st.sync.aligned.32x32b.x4.b32 [%8], {x8,x9,x10,x11};\n\t"
                 "tcgen05.st.sync.aligned.32x32b.x2.b32 [%9], {x12,x13};\n\t}"
                 ::"r"(ta), "d"(a), "d"(b), "d"(c), "d"(d), "d"(e), "d"(f), "d"(g), "r"(ta + 8), "r"(ta + 12) : "memory");
}
// one double parked in / fetched from two tensor-memory columns
__device__ __forceinline__ void tmem_st1d(uint32_t ta, double a) {
    asm volatile("{\n\t.reg .b32 x<2>;\n\tmov.b64 {x0,x1}, %1;\n\ttcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {x0,x1};\n\t}" ::"r"(ta), "d"(a) : "memory");
}
__device__ __forceinline__ double tmem_ld1d(uint32_t ta) {
    double a;
    asm volatile("{\n\t.reg .b32 x<2>;\n\ttcgen05.ld.sync.aligned.32x32b.x2.b32 {x0,x1}, [%1];\n\ttcgen05.wait::ld.sync.aligned;\n\tmov.b64 %0, {x0,x1};\n\t}"
                 : "=d"(a) : "r"(ta) : "memory");
    return a;
}

// ring_off[S-1][n] = (n mod (2S+1)) * 14: tensor-memory column offset of ring slot n, n < 2(2S+1) (no modulo in the plane loop)
__constant__ uint32_t ring_off[4][18] = {
    {0, 14, 28, 0, 14, 28},
    {0, 14, 28, 42, 56, 0, 14, 28, 42, 56},
    {0, 14, 28, 42, 56, 70, 84, 0, 14, 28, 42, 56, 70, 84},
    {0, 14, 28, 42, 56, 70, 84, 98, 112, 0, 14, 28, 42, 56, 70, 84, 98, 112}};
// thread index that the compiler cannot hoist out of the plane loop (everything derived from it is rebuilt where it is used
// instead of living in -- or being spilled from -- a register across the stencil sums)
__device__ __forceinline__ int tid_now() { int t; asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t)); return t; }
__device__ __forceinline__ int ctaid_z_now() { int t; asm volatile("mov.u32 %0, %%ctaid.z;" : "=r"(t)); return t; }
__device__ __forceinline__ uint32_t lds_u32_now(uint32_t a) { uint32_t v; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }

__device__ __forceinline__ void st_out(double *p, double v) {
#if FAST_L2_HINTS
    __stcs(p, v);
#else
    *p = v;
#endif
}

// MODE 0: write the right-hand side only (test path); 1: Runge-Kutta update without an RA operand; 2: with RA
template <int S, int V, int TY, int MODE>
__global__ void __launch_bounds__(TX * TY, TY == 16 ? 1 : 2)
stage_kernel(const __grid_constant__ KConst c, const __grid_constant__ StagePtrs P, const __grid_constant__ StageCoef sc, int zchunk,
             const __grid_constant__ FastMaps tm) {
    using G = FCfg<S, TY>;
    constexpr int NT = G::NT, R = G::R, CSZ = G::CSZ;
    extern __shared__ __align__(1024) double smem[];
    double *intb = smem + 2 * G::BUF_D;                // [7][TY][TX]  (rho,u,v,w,H,T,theta) of the plane S ahead, tile interior
    uint64_t *mbar_p = (uint64_t *)(smem + G::DATA_D);
    uint32_t *tmem_holder = (uint32_t *)(mbar_p + 4);
    double *dt_s = (double *)(mbar_p + 5);              // dt, read once per CTA
#if FAST_LAST_ISSUES
    unsigned int *arrivals = (unsigned int *)(mbar_p + 6);   // warps that have emptied the ring tile, counted over the chunk
#endif

    const Layout &L = c.L;
    const int tid = threadIdx.x;
    const int ty = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp == tile row (warp-uniform for the compiler)
    const int i0 = blockIdx.x * TX, j0 = blockIdx.y * TY;
    const int kbeg = blockIdx.z * zchunk;
    constexpr bool do_update = MODE != 0, useA = MODE == 2;
    auto kend_now = [&]() { return min((ctaid_z_now() + 1) * zchunk, c.L.mz); };

    // mb_full[b]: the TMA loads of a plane bundle into buffer b have landed; mb_free: every warp has taken its ring values out of intb
    const uint32_t mb_full0 = smem_u32(mbar_p), mb_free = mb_full0 + 16;
#if FAST_SPLIT_BAR
    const uint32_t mb_int = mb_full0 + 24;
#endif
    if (ty == 0) tmem_alloc(smem_u32(tmem_holder), G::NCOLS);
    if (tid == 0) {
        mbar_init(mb_full0, 1); mbar_init(mb_full0 + 8, 1); mbar_init(mb_free, TY);
#if FAST_SPLIT_BAR
        mbar_init(mb_int, 1);
#endif
        *dt_s = *c.dt;
#if FAST_LAST_ISSUES
        *arrivals = 0u;
#endif
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem0 = *tmem_holder;
    auto tbase_of = [&](int wy) { return lds_u32_now(smem_u32(tmem_holder)) + ((uint32_t)(32 * (wy & 3)) << 16) + (uint32_t)((wy >> 2) * G::COLS_THREAD); };
    const uint32_t tbase = tbase_of(ty);
    auto tslot = [&](int kk) -> uint32_t { return tbase + (uint32_t)(((kk + 16 * R) % R) * G::COLS_SLOT); };

    const uint32_t s_base = smem_u32(smem), s_int = smem_u32(intb);
    // plane bundle k: halo'd tile of plane k, rho*E and RA of its interior -> buffer b; interior of plane k+S -> intb
    auto issue_bundle = [&](int k, int b) {
        const uint32_t mb = mb_full0 + 8 * b;
        const uint32_t s_box = s_base + (uint32_t)(b * G::BUF_D * 8), s_e = s_box + (uint32_t)(G::BOX_D * 8), s_op = s_e + (uint32_t)(G::E_D * 8);
#if FAST_SPLIT_BAR
        mbar_expect_tx(mb_int, (uint32_t)(G::INT_D * sizeof(double)));
        tma_load_4d(s_int, &tm.q4int, mb_int, i0 + GX, j0 + L.gy, k + S + L.gz, 0);
        tma_load_4d(s_int + 4 * NT * 8, &tm.a3int, mb_int, i0 + GX, j0 + L.gy, k + S + L.gz, 0);
        mbar_expect_tx(mb, (uint32_t)((G::BOX_D + G::E_D + (useA ? G::OP_D : 0)) * sizeof(double)));
        tma_load_4d(s_box, &tm.q4box, mb, i0, j0 + L.gy - S, k + L.gz, 0);
        tma_load_4d(s_box + 4 * CSZ * 8, &tm.a3box, mb, i0, j0 + L.gy - S, k + L.gz, 0);
        tma_load_3d(s_e, &tm.eint, mb, i0 + GX, j0 + L.gy, k + L.gz);
        if (useA) tma_load_4d(s_op, &tm.opa, mb, i0, j0, k, 0);
#elif FAST_L2_HINTS
        mbar_expect_tx(mb, (uint32_t)((G::BOX_D + G::E_D + (useA ? G::OP_D : 0) + G::INT_D) * sizeof(double)));
        tma_load_4d_hint(s_box, &tm.q4box, mb, i0, j0 + L.gy - S, k + L.gz, 0, L2_EVICT_FIRST);
        tma_load_4d_hint(s_box + 4 * CSZ * 8, &tm.a3box, mb, i0, j0 + L.gy - S, k + L.gz, 0, L2_EVICT_FIRST);
        tma_load_4d_hint(s_int, &tm.q4int, mb, i0 + GX, j0 + L.gy, k + S + L.gz, 0, L2_EVICT_LAST);
        tma_load_4d_hint(s_int + 4 * NT * 8, &tm.a3int, mb, i0 + GX, j0 + L.gy, k + S + L.gz, 0, L2_EVICT_LAST);
        tma_load_3d_hint(s_e, &tm.eint, mb, i0 + GX, j0 + L.gy, k + L.gz, L2_EVICT_FIRST);
        if (useA) tma_load_4d_hint(s_op, &tm.opa, mb, i0, j0, k, 0, L2_EVICT_FIRST);
#else
        mbar_expect_tx(mb, (uint32_t)((G::BOX_D + G::E_D + (useA ? G::OP_D : 0) + G::INT_D) * sizeof(double)));
        tma_load_4d(s_box, &tm.q4box, mb, i0, j0 + L.gy - S, k + L.gz, 0);
        tma_load_4d(s_box + 4 * CSZ * 8, &tm.a3box, mb, i0, j0 + L.gy - S, k + L.gz, 0);
        tma_load_4d(s_int, &tm.q4int, mb, i0 + GX, j0 + L.gy, k + S + L.gz, 0);
        tma_load_4d(s_int + 4 * NT * 8, &tm.a3int, mb, i0 + GX, j0 + L.gy, k + S + L.gz, 0);
        tma_load_3d(s_e, &tm.eint, mb, i0 + GX, j0 + L.gy, k + L.gz);
        if (useA) tma_load_4d(s_op, &tm.opa, mb, i0, j0, k, 0);
#endif
    };

    // ---- prologue: planes kbeg-S .. kbeg+S-1 into the ring, two batches of S planes through the (still unused) buffers
    for (int bt = 0; bt < 2; bt++) {
        if (tid == 0) {
            mbar_expect_tx(mb_full0, (uint32_t)(G::PRO_D * sizeof(double)));
            for (int n = 0; n < S; n++) {
                const int kk = kbeg - S + bt * S + n;
                tma_load_4d(smem_u32(smem + (size_t)n * NF * NT), &tm.q4int, mb_full0, i0 + GX, j0 + L.gy, kk + L.gz, 0);
                tma_load_4d(smem_u32(smem + (size_t)n * NF * NT + 4 * NT), &tm.a3int, mb_full0, i0 + GX, j0 + L.gy, kk + L.gz, 0);
            }
        }
        mbar_wait(mb_full0, (uint32_t)bt);
        for (int n = 0; n < S; n++) {
            const double *src = smem + (size_t)n * NF * NT + tid;
            tmem_st7d(tslot(kbeg - S + bt * S + n), src[0], src[NT], src[2 * NT], src[3 * NT], src[4 * NT], src[5 * NT], src[6 * NT]);
        }
        __syncthreads();
    }
    if (tid == 0) issue_bundle(kbeg, 0);

    // ring slot of plane k-S+n: column offset ring_off[s0 + n] with s0 = (k-S) mod R
    // (recomputed from k every plane: as a loop-carried counter it was spilled, and the reload -- an L2 round trip -- stalled the
    // loop branch for 3.7 % of the kernel time, profiles/r01_stage_fast8_512_ncu_full.txt)
    const size_t N = (size_t)L.mx * L.my * L.mz;
    for (int k = kbeg; k < kend_now(); k++) {
        const int s0 = (int)((unsigned)(k - S + 16 * R) % (unsigned)R);
        // everything that depends on the thread index is rebuilt here and again before the update instead of being kept in (or
        // spilled from) registers across the stencil sums
        const int tn = tid_now();
        const int tx = tn & 31, ty = tn >> 5;
        const int kk = k - ctaid_z_now() * zchunk;
        const int b = kk & 1;
        const double *box = smem + (size_t)b * G::BUF_D;
        const uint32_t tbase = tbase_of(__shfl_sync(0xffffffffu, ty, 0));
        uint32_t zs[R];
#pragma unroll
        for (int n = 0; n < R; n++) zs[n] = tbase + ring_off[S - 1][s0 + n];
        const int own = (ty + S) * CX + (tx + GX);
        // the first buffer's barrier has already completed two phases in the prologue: the parities line up (phase 2 -> parity 0)
#if FAST_SPLIT_BAR
        mbar_wait(mb_int, (uint32_t)kk & 1u);
#else
        mbar_wait(mb_full0 + 8 * b, (uint32_t)(kk >> 1) & 1u);
#endif
        // ---- own point of plane k out of the ring; plane k+S into the ring (no arithmetic: the state carries H and T)
        Slot sl;
        slot_issue(zs[S], sl);
        {
            const double *src = intb + ty * TX + tx;
            const double r0 = src[0], r1 = src[NT], r2 = src[2 * NT], r3 = src[3 * NT], r4 = src[4 * NT], r5 = src[5 * NT], r6 = src[6 * NT];
            __syncwarp();
#if FAST_LAST_ISSUES
            tmem_st7d_nowait(zs[2 * S], r0, r1, r2, r3, r4, r5, r6);
            // the warp that arrives last knows every warp is done with intb and with plane k-1: it sends the loads of plane k+1
            // right away, a whole plane ahead (the wait completes at once; it is there for the acquire)
            if (tx == 0) {
                mbar_arrive(mb_free);
                if (atomicAdd(arrivals, 1u) == (unsigned)((kk + 1) * TY - 1) && k + 1 < kend_now()) {
                    mbar_wait(mb_free, (uint32_t)kk & 1u);
                    issue_bundle(k + 1, b ^ 1);
                }
            }
#else
            if (tx == 0) mbar_arrive(mb_free);                    // intb is consumed as soon as the values sit in registers
            tmem_st7d_nowait(zs[2 * S], r0, r1, r2, r3, r4, r5, r6);
#endif
        }
        double C[NF];
        slot_wait(sl, C);

        Acc A;
        // centre weights of the three second derivatives in one go
        A.lapu[0] = c.c20sum * C[FU]; A.lapu[1] = c.c20sum * C[FV]; A.lapu[2] = c.c20sum * C[FW]; A.lapT = c.c20sum * C[FT];
#pragma unroll
        for (int m = 0; m < 5; m++) A.r[m] = 0.0;
#pragma unroll
        for (int d = 0; d < 3; d++) { A.g[d][0] = A.g[d][1] = A.g[d][2] = 0.0; A.dT[d] = 0.0; }
        // ---- z direction: neighbours from the ring in tensor memory
        {
            double aM = 0.0;
#pragma unroll
            for (int l = 1; l <= S; l++) {
#if FAST_ONESIDED
                if (l == S) tmem_wait_st();
                { Slot sp; slot_issue(zs[S + l], sp); double Nq[NF]; slot_wait(sp, Nq); side_step<2, V, true>(c, l, C, Nq, A, aM); }
                { Slot sm; slot_issue(zs[S - l], sm); double Nq[NF]; slot_wait(sm, Nq); side_step<2, V, false>(c, l, C, Nq, A, aM); }
#else
                Slot sp, sm;
                if (l == S) tmem_wait_st();                       // the slot of plane k+S was stored at the top of this plane
                slot_issue(zs[S + l], sp); slot_issue(zs[S - l], sm);
                double Pn[NF], Mn[NF];
                slot_wait2(sp, sm, Pn, Mn);
                pair_step<2, V>(c, l, C, Pn, Mn, A, aM);
#endif
            }
            close_dir(C, A, aM);
        }
        // ---- the elected thread: once every warp has emptied intb (which also means it is done with plane k-1 and its buffer),
        // the loads of plane k+1 go out; they have the x / y directions, the assembly and the update of this plane to land
#if !FAST_LAST_ISSUES
        if (tn == 0 && k + 1 < kend_now()) {
            mbar_wait(mb_free, (uint32_t)kk & 1u);
            issue_bundle(k + 1, b ^ 1);
        }
#endif

        // ---- x and y directions: neighbours straight from the TMA-landed plane, [7][CY][CX]
#if FAST_SPLIT_BAR
        mbar_wait(mb_full0 + 8 * b, (uint32_t)(kk >> 1) & 1u);
#endif
        {
            const double *pc = box + own;
#pragma unroll
            for (int d = 0; d < 2; d++) {
                const int dstr = d ? CX : 1;                    // x: neighbouring cells, y: neighbouring rows
                double aM = 0.0;
#pragma unroll
                for (int l = 1; l <= S; l++) {
#if FAST_ONESIDED
#pragma unroll
                    for (int sgn = 0; sgn < 2; sgn++) {
                        double Nq[NF];
                        const int off = sgn ? -l * dstr : l * dstr;
#pragma unroll
                        for (int f = 0; f < NF; f++) Nq[f] = (f < FD || l <= V) ? pc[f * CSZ + off] : 0.0;
                        if (d == 0) { if (sgn) side_step<0, V, false>(c, l, C, Nq, A, aM); else side_step<0, V, true>(c, l, C, Nq, A, aM); }
                        else { if (sgn) side_step<1, V, false>(c, l, C, Nq, A, aM); else side_step<1, V, true>(c, l, C, Nq, A, aM); }
                    }
#else
                    double Pn[NF], Mn[NF];
#pragma unroll
                    for (int f = 0; f < NF; f++) {
                        if (f < FD || l <= V) { Pn[f] = pc[f * CSZ + l * dstr]; Mn[f] = pc[f * CSZ - l * dstr]; }
                        else { Pn[f] = 0.0; Mn[f] = 0.0; }
                    }
                    if (d == 0) pair_step<0, V>(c, l, C, Pn, Mn, A, aM); else pair_step<1, V>(c, l, C, Pn, Mn, A, aM);
#endif
                }
                close_dir(C, A, aM);
            }
        }

        // ---- stress, dissipation, heat flux: assembled once per point (cuda_rhs.cu:52-127,169-259,303-393); g[d][m] = d u_m / d x_d
        const double g00 = A.g[0][0], g10 = A.g[0][1], g20 = A.g[0][2], dT0 = A.dT[0];
        const double g01 = A.g[1][0], g11 = A.g[1][1], g21 = A.g[1][2], dT1 = A.dT[1];
        const double g02 = A.g[2][0], g12 = A.g[2][1], g22 = A.g[2][2], dT2 = A.dT[2];
        const double mu = C[FT] * c.invRe;
        const double dm0 = dT0 * c.invRe, dm1 = dT1 * c.invRe, dm2 = dT2 * c.invRe;
        const double th23 = (2.0 / 3.0) * C[FD];
        const double s01 = g01 + g10, s02 = g02 + g20, s12 = g12 + g21;
        const double d00 = 2.0 * g00 - th23, d11 = 2.0 * g11 - th23, d22 = 2.0 * g22 - th23;
        // F_m = mu (lap u_m + (1/3) d_m theta) + sum_d (g_md + g_dm) dmu_d - (2/3) theta dmu_m
        const double F0 = fma(mu, A.lapu[0], fma(d00, dm0, fma(s01, dm1, s02 * dm2)));
        const double F1 = fma(mu, A.lapu[1], fma(s01, dm0, fma(d11, dm1, s12 * dm2)));
        const double F2 = fma(mu, A.lapu[2], fma(s02, dm0, fma(s12, dm1, d22 * dm2)));
        const double work = fma(C[FU], F0, fma(C[FV], F1, C[FW] * F2));
        // dissipation; quirk Q1 (cuda_rhs.cu:175): the y kernel multiplies (dv/dz + dw/dy) by dv/dz where dw/dy is meant
        const double g3y = c.quirk_q1 ? g12 : g21;
        double diss = d00 * g00;
        diss = fma(s01, g10, diss); diss = fma(s02, g20, diss);
        diss = fma(s01, g01, diss); diss = fma(d11, g11, diss); diss = fma(s12, g3y, diss);
        diss = fma(s02, g02, diss); diss = fma(s12, g12, diss); diss = fma(d22, g22, diss);
        double rhs[5];
        rhs[0] = A.r[0];
        rhs[1] = A.r[1] + F0;
        rhs[2] = A.r[2] + F1;
        rhs[3] = A.r[3] + F2;
        // lambda = mu/(Pr Ec) (cuda_main.cu:239): lambda*lap(T) + grad(lambda).grad(T)
        const double heat = fma(mu, A.lapT, fma(dm0, dT0, fma(dm1, dT1, dm2 * dT2)));
        rhs[4] = A.r[4] + fma(mu, diss, fma(c.lamfac, heat, work));
        if (c.forcing) { const double fz = *c.dpdz; rhs[3] += fz; rhs[4] = fma(fz, C[FW], rhs[4]); }      // cuda_rhs.cu:392-393
        // this thread's point: flags (1 active, 2/4 periodic x images low/high, 8/16 periodic y images: perBCx / perBCy,
        // boundary.h:38-46) and element offsets inside one padded field / one unpadded register array
        const int tl = tid_now();
        const int tu = tl & 31, tw = tl >> 5;
        const int i = i0 + tu, j = j0 + tw;
        const unsigned flags = ((i < L.mx && j < L.my) ? 1u : 0u) | ((i < S) ? 2u : 0u) | ((i >= L.mx - S) ? 4u : 0u) |
                               ((j < S) ? 8u : 0u) | ((j >= L.my - S) ? 16u : 0u);
        // element offsets: a warp-uniform 64-bit part (tile origin, plane) and a 32-bit per-thread part
        const size_t gq0 = L.idx(i0, j0, k), nq0 = (size_t)i0 + (size_t)j0 * L.mx + (size_t)k * L.mx * L.my;
        const uint32_t gqt = (uint32_t)(tw * L.px + tu), nqt = (uint32_t)(tw * L.mx + tu);
        if (!do_update) {
            if (flags & 1u) {
#pragma unroll
                for (int m = 0; m < 5; m++) (P.rhs_out + m * N + nq0)[nqt] = rhs[m];
            }
        } else {
            // Runge-Kutta register update (sumLowStorageRK3 cuda_main.cu:244): Q_out = Q + dt (cN K + cA RA), RW = wNew K
            if (P.RW && (flags & 1u)) {
#pragma unroll
                for (int m = 0; m < 5; m++) st_out((P.RW + m * N + nq0) + nqt, sc.wNew * rhs[m]);
            }
            double kq[5];
#pragma unroll
            for (int m = 0; m < 5; m++) kq[m] = sc.cN * rhs[m];
            const double *eb = box + G::BOX_D;
            if (useA) {
                const double *opA = eb + G::E_D + tl;
#pragma unroll
                for (int m = 0; m < 5; m++) kq[m] = fma(sc.cA, opA[m * NT], kq[m]);
            }
            const double dt = *dt_s;
            double qn[5];
            qn[0] = fma(dt, kq[0], C[FR]);
            qn[1] = fma(dt, kq[1], C[FR] * C[FU]);
            qn[2] = fma(dt, kq[2], C[FR] * C[FV]);
            qn[3] = fma(dt, kq[3], C[FR] * C[FW]);
            qn[4] = fma(dt, kq[4], eb[tl]);
            const double rn = 1.0 / qn[0];                                                   // deviceDiv cuda_math.cu:36
            double out[7] = {qn[0], qn[1] * rn, qn[2] * rn, qn[3] * rn, qn[4], 0.0, 0.0};
            eos_ht(c, out[0], rn, out[1], out[2], out[3], out[4], out[5], out[6]);          // H and T travel with the state
            if (flags & 1u) {
                auto store_point = [&](double *fu) {                // fu: warp-uniform (tile origin of the plane)
#pragma unroll
                    for (int m = 0; m < 7; m++) st_out((fu + m * L.vol) + gqt, out[m]);
                    if (flags & 30u) {
#pragma unroll
                        for (int m = 0; m < 7; m++) {
                            double *fm = (fu + m * L.vol) + gqt;
                            if (flags & 2u) st_out(fm + L.mx, out[m]);
                            if (flags & 4u) st_out(fm - (ptrdiff_t)L.mx, out[m]);
                            if (flags & 8u) st_out(fm + (size_t)L.my * L.px, out[m]);
                            if (flags & 16u) st_out(fm - (ptrdiff_t)((size_t)L.my * L.px), out[m]);
                        }
                    }
                };
                store_point(P.qout + gq0);
                // z ghosts (replaces updateHaloFive comm.cpp:114-134 / perBCz boundary.h:48-51): the first / last gz planes also go
                // into the ghost planes of the lower / upper slab neighbour -- peer memory over NVLink, or this buffer itself
                if (k < L.gz && P.qout_lo) store_point(P.qout_lo + gq0 + (size_t)L.mz * L.plane);
                if (k >= L.mz - L.gz && P.qout_hi) store_point(P.qout_hi + gq0 - (size_t)L.mz * L.plane);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if ((tid_now() >> 5) == 0) tmem_dealloc(lds_u32_now(smem_u32(tmem_holder)), G::NCOLS);
}

template <int S, int V, int TY>
static void launch_t(const KConst &kc, const StagePtrs &p, const StageCoef &c, const FastMaps &maps, cudaStream_t st) {
    using G = FCfg<S, TY>;
    opt_in_smem<stage_kernel<S, V, TY, 0>>((int)G::bytes);
    opt_in_smem<stage_kernel<S, V, TY, 1>>((int)G::bytes);
    opt_in_smem<stage_kernel<S, V, TY, 2>>((int)G::bytes);
    const int gx = (kc.L.mx + TX - 1) / TX, gy = (kc.L.my + TY - 1) / TY;
    // z chunks: every chunk pays a 2S-plane prologue, so keep them >= 32 planes; more chunks smooth the tail over the SMs
    const int cols = gx * gy, resident = 148 * (TY == 16 ? 1 : 2);
    int nzc = pick_zchunks(cols, kc.L.mz, S, resident);
    int zchunk = (kc.L.mz + nzc - 1) / nzc;
    nzc = (kc.L.mz + zchunk - 1) / zchunk;
    dim3 grid(gx, gy, nzc);
    if (p.rhs_out) stage_kernel<S, V, TY, 0><<<grid, TX * TY, G::bytes, st>>>(kc, p, c, zchunk, maps);
    else if (!p.RA) stage_kernel<S, V, TY, 1><<<grid, TX * TY, G::bytes, st>>>(kc, p, c, zchunk, maps);
    else stage_kernel<S, V, TY, 2><<<grid, TX * TY, G::bytes, st>>>(kc, p, c, zchunk, maps);
}
template <int TY>
static void launch_ty(const KConst &kc, const StagePtrs &p, const StageCoef &c, const FastMaps &maps, cudaStream_t st) {
    switch (kc.s * 10 + kc.v) {
        case 11: launch_t<1, 1, TY>(kc, p, c, maps, st); break;
        case 21: launch_t<2, 1, TY>(kc, p, c, maps, st); break;
        case 22: launch_t<2, 2, TY>(kc, p, c, maps, st); break;
        case 31: launch_t<3, 1, TY>(kc, p, c, maps, st); break;
        case 32: launch_t<3, 2, TY>(kc, p, c, maps, st); break;
        case 33: launch_t<3, 3, TY>(kc, p, c, maps, st); break;
        case 41: launch_t<4, 1, TY>(kc, p, c, maps, st); break;
        case 42: launch_t<4, 2, TY>(kc, p, c, maps, st); break;
        case 43: launch_t<4, 3, TY>(kc, p, c, maps, st); break;
        default: launch_t<4, 4, TY>(kc, p, c, maps, st); break;
    }
}

}  // namespace fast

// same preconditions as the wide lean variant (lean_wide_ok, one operand tile); p.qin / p.qout are 8-field buffers (see the header);
// ty = 16: one CTA of 16 warps per SM, ty = 8: two CTAs of 8 warps (the TMA descriptors must have been built for that tile)
void launch_rhs_stage_fast(const KConst &kc, const StagePtrs &p, const StageCoef &c, const FastMaps &maps, int ty, cudaStream_t st) {
    if (ty == 16) fast::launch_ty<16>(kc, p, c, maps, st); else fast::launch_ty<8>(kc, p, c, maps, st);
}
int fast_smem_bytes(int s, int ty) {
    using namespace fast;
    if (ty == 16) switch (s) { case 1: return (int)FCfg<1, 16>::bytes; case 2: return (int)FCfg<2, 16>::bytes; case 3: return (int)FCfg<3, 16>::bytes; default: return (int)FCfg<4, 16>::bytes; }
    switch (s) { case 1: return (int)FCfg<1, 8>::bytes; case 2: return (int)FCfg<2, 8>::bytes; case 3: return (int)FCfg<3, 8>::bytes; default: return (int)FCfg<4, 8>::bytes; }
}

}  // namespace cudns
