// libcudns: dilatation pass  theta = du/dx + dv/dy + dw/dz  at viscous order
// (derVelX/Y/Z + calcDil of the reference, calc_stress.cu:20-96, collapsed into one streaming kernel).
//
// The stage kernel needs theta at stencil neighbours, i.e. theta of the whole field has to exist before the
// right-hand side of a stage can be evaluated; this pass is the price: it reads u,v,w once and
// writes theta once (32 B per point of HBM traffic, against 168 B of the stage kernel).
//
// Mapping: one CTA = a 32 x 16 tile of (i,j) columns marching along z, one thread per column.
//   * the planes of u (with x halos), v (with y halos) and w travel global -> shared memory with cp.async (no register
//     staging) through a ring of NST = 4 stages, three planes ahead of the one being computed: ~45 KB per CTA in flight,
//     which is what 6.5 TB/s at ~1 us latency needs (the first version staged one plane ahead through registers and
//     stopped at 56 % of the copy bandwidth); one __syncthreads per plane;
//   * du/dx and dv/dy read their neighbours from the shared plane; dw/dz keeps the 2V+1 most recent w values of the column
//     in REGISTERS (a shifting window: 2V moves per plane, noise next to the memory time of this kernel);
//   * wall / extrapolation ghosts in x are built in shared memory (BCxderVel, boundary_condition_x.h:38-40,67,94),
//     the z extrapolation of the boundary layer (BCzderVel, boundary_condition_z.h:34-40) is applied on the few
//     planes next to the global z boundaries by a slow path that reads global memory directly.
#include "cudns_internal.h"
#include <cstdlib>

namespace cudns {
namespace {

constexpr int TXT = 32, TYT = 16, NTT = TXT * TYT;
constexpr int UX = TXT + 2 * GX;
constexpr int NST = 4;                                   // cp.async ring depth (planes)

__device__ __forceinline__ void cp_async8(real *smem_dst, const real *gsrc) {        // one element
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "n"(sizeof(real)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// wall blowing/suction, perturbation.h:25-53
__device__ __forceinline__ bool perturb_theta(const KConst &c, int j, int kglob, real &val) {
    int kSt = c.kC - c.LP / 2, kEn = c.kC + c.LP / 2;
    if (kglob < kSt || kglob > kEn) return false;
    int alpha, beta, kappa;
    if (kglob < c.kC) { kappa = 1; alpha = kglob - kSt; beta = c.kC - kSt; }
    else              { kappa = -1; alpha = kEn - kglob; beta = kEn - c.kC; }
    real ksi = alpha * 1.0 / beta;
    real g = (15.1875 * ksi * ksi * ksi * ksi * ksi) - (35.4375 * ksi * ksi * ksi * ksi) + (20.25 * ksi * ksi * ksi);
    real y_glob = (real)j / c.d1[1];
    real tg = *c.time_on_gpu;
    val = c.amp1 * kappa * g * sin(c.omega1 * tg) + c.amp2 * kappa * g * sin(c.omega2 * tg) * cos(y_glob / c.lambdaP);
    return true;
}


// botBCzExt / topBCzExt (boundary.h:154-160): dw/dz next to the global z boundaries of the boundary-layer case, where w
// past the boundary is the node extrapolation f[-g] = 2 f[0] - f[g], f[mz-1+g] = 2 f[mz-1] - f[mz-1-g] (slow path)
template <int V>
__device__ __forceinline__ real dwdz_edge(const KConst &c, const real *__restrict__ W, int ic, int jc, int k, int kglob_lo, int kglob_hi) {
    const Layout &L = c.L;
    const size_t g0 = L.idx(ic, jc, k);
    real dwdz = 0.0;
#pragma unroll
    for (int l = 1; l <= V; l++) {
        real wp, wm;
        if (k + l >= kglob_hi) wp = RC(2.0) * W[L.idx(ic, jc, kglob_hi - 1)] - W[L.idx(ic, jc, 2 * (kglob_hi - 1) - (k + l))];
        else wp = W[g0 + (size_t)l * L.plane];
        if (k - l < kglob_lo) wm = RC(2.0) * W[L.idx(ic, jc, kglob_lo)] - W[L.idx(ic, jc, 2 * kglob_lo - (k - l))];
        else wm = W[g0 - (size_t)l * L.plane];
        dwdz = fma(c.c1[2][l], wp - wm, dwdz);
    }
    return dwdz;
}

template <int V, bool GEN>
__global__ void __launch_bounds__(NTT, 2)
theta_march_kernel(const __grid_constant__ KConst c, const real *__restrict__ q, real *__restrict__ theta, int zchunk) {
    constexpr int VY = TYT + 2 * V;
    constexpr int SU = TYT * UX, SV = VY * TXT, SW = NTT;  // doubles per stage
    extern __shared__ __align__(16) real sth[];
    real *su = sth;                                     // [NST][TYT][UX]
    real *sv = su + NST * SU;                           // [NST][VY][TXT]
    real *sw = sv + NST * SV;                           // [NST][TYT][TXT]

    const Layout &L = c.L;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i0 = blockIdx.x * TXT, j0 = blockIdx.y * TYT;
    const int i = i0 + tx, j = j0 + ty;
    const int nxt = min(TXT, L.mx - i0), nyt = min(TYT, L.my - j0);
    const bool inx = tx < nxt, iny = ty < nyt, active = inx && iny;
    const int ic = min(i, L.mx - 1), jc = min(j, L.my - 1);
    const int kfirst = -V + (int)blockIdx.z * zchunk;
    const int klast = min(kfirst + zchunk, L.mz + V);           // exclusive
    const int kglob_lo = -c.kstart, kglob_hi = c.mz_tot - c.kstart;
    const bool bl = GEN && c.boundaryLayer != 0;
    const bool perx = !GEN || c.periodicX != 0;
    const bool xlo = !perx && i0 == 0, xhi = !perx && (i0 + TXT >= L.mx);
    const size_t plane = L.plane;

    // this thread's column, the x-halo cell and the y-halo row it also stages: running pointers, one plane per iteration
    const size_t g00 = L.idx(ic, jc, kfirst);
    const bool hx_on = tx < 2 * V && iny;
    const int hxc = tx < V ? GX - V + tx : GX + nxt + (tx - V);           // column in su
    const int hgi = i0 + hxc - GX;
    const bool hx_load = hx_on && (perx || (hgi >= 0 && hgi < L.mx));
    const bool hy_on = ty < 2 * V && inx;
    const int hyr = ty < V ? ty : V + nyt + (ty - V);                     // row in sv
    const real *pu = q + L.vol + g00, *pv = q + 2 * L.vol + g00, *pw = q + 3 * L.vol + g00 + (size_t)V * plane;
    const real *puh = q + L.vol + L.idx(min(max(hgi, -GX), L.mx + GX - 1), jc, kfirst);
    const real *pvh = q + 2 * L.vol + L.idx(ic, j0 + hyr - V, kfirst);
    real *pt = theta + g00;
    const real xpi = (GEN && c.nonUniformX) ? c.xp[ic] : RC(1.0);
    // shared-memory slots (doubles, buffer 0)
    const int o_u = ty * UX + GX + tx, o_uh = ty * UX + hxc, o_v = (V + ty) * TXT + tx, o_vh = hyr * TXT + tx;
    // image flags: 1 x-low, 2 x-high, 4 y-low, 8 y-high (perBCx / perBCy, boundary.h:38-46)
    const unsigned img = ((perx && i < V) ? 1u : 0u) | ((perx && i >= L.mx - V) ? 2u : 0u) | ((j < V) ? 4u : 0u) | ((j >= L.my - V) ? 8u : 0u);

    real wr[2 * V + 1];                                 // wr[V + l] = w of plane k+l
#pragma unroll
    for (int m = 0; m < 2 * V; m++) wr[m + 1] = pw[(ptrdiff_t)(m - 2 * V) * (ptrdiff_t)plane];
    // plane kk (u, v at kk; w at kk+V) -> ring stage st; every thread commits exactly one group per call
    const int o_w = ty * TXT + tx;
    auto stage_plane = [&](int st) {
        if (active) { cp_async8(su + st * SU + o_u, pu); cp_async8(sv + st * SV + o_v, pv); }   // (slots past a ragged tile edge belong to halo cells)
        cp_async8(sw + st * SW + o_w, pw);
        if (hx_load) cp_async8(su + st * SU + o_uh, puh);
        if (hy_on) cp_async8(sv + st * SV + o_vh, pvh);
        pu += plane; pv += plane; pw += plane; puh += plane; pvh += plane;
    };
#pragma unroll
    for (int n = 0; n < NST - 1; n++) {
        if (kfirst + n < klast) stage_plane(n);
        cp_async_commit();
    }

    int st = 0;
    for (int k = kfirst; k < klast; k++, st = (st + 1 == NST) ? 0 : st + 1) {
        cp_async_wait<NST - 2>();                         // this thread's copies of plane k have landed ...
        __syncthreads();                                  // ... and everybody else's; plane k-1 is consumed
        if (k + NST - 1 < klast) stage_plane(st == 0 ? NST - 1 : st - 1);
        cp_async_commit();
        const int boff_u = st * SU, boff_v = st * SV;
#pragma unroll
        for (int m = 0; m < 2 * V; m++) wr[m] = wr[m + 1];
        wr[2 * V] = sw[st * SW + o_w];
        const bool outside = bl && (k < kglob_lo || k >= kglob_hi);    // ghost theta is extrapolated by the stage kernel
        if constexpr (GEN) {
            if (xlo || xhi) {
                // BCxderVel: wall (anti-mirror about the face, + blowing/suction) / node extrapolation at the free stream
                if (hx_on && !hx_load && !outside) {
                    const real *row = su + boff_u + ty * UX;
                    real val;
                    if (tx < V) {
                        const int gq = V - tx;                            // ghost -gq
                        val = -row[GX + gq - 1];
                        real pv2;
                        if (bl && c.perturbed && perturb_theta(c, j, k + c.kstart, pv2)) val = pv2;
                    } else {
                        const int gq = tx - V + 1, last = GX + nxt - 1;
                        val = bl ? RC(2.0) * row[last] - row[last - gq] : -row[last - gq + 1];
                    }
                    su[boff_u + o_uh] = val;
                }
                __syncthreads();
            }
        }
        if (active && !outside) {
            real dudx = 0.0, dvdy = 0.0, dwdz = 0.0;
            const real *ur = su + boff_u + o_u;
            const real *vr = sv + boff_v + o_v;
#pragma unroll
            for (int l = 1; l <= V; l++) {
                dudx = fma(c.c1[0][l], ur[l] - ur[-l], dudx);
                dvdy = fma(c.c1[1][l], vr[l * TXT] - vr[-l * TXT], dvdy);
            }
            if (GEN && bl && (k - kglob_lo < V || kglob_hi - 1 - k < V)) {
                dwdz = dwdz_edge<V>(c, q + 3 * L.vol, ic, jc, k, kglob_lo, kglob_hi);
            } else {
#pragma unroll
                for (int l = 1; l <= V; l++) dwdz = fma(c.c1[2][l], wr[V + l] - wr[V - l], dwdz);
            }
            const real th = (GEN ? fma(dudx, xpi, dvdy) : dudx + dvdy) + dwdz;
            *pt = th;
            if (img) {
                if (img & 1u) pt[L.mx] = th;
                if (img & 2u) pt[-(ptrdiff_t)L.mx] = th;
                if (img & 4u) pt[(size_t)L.my * L.px] = th;
                if (img & 8u) pt[-(ptrdiff_t)((size_t)L.my * L.px)] = th;
            }
        }
        pt += plane;
    }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// TMA variant for the periodic / uniform set-ups (no wall, metric or extrapolation code): the same pass with
//   * a 64 x 16 tile and FOUR points per thread (two x-adjacent cells in rows ty and ty + 8): every shared-memory access is a 128-bit
//     load, theta leaves as 128-bit stores; the y halo costs 1.5x (was 2x), the x halo 1.125x (was 1.25x) of the tile;
//   * the planes of u (x halos), v (y halos) and w land by TMA (cp.async.bulk.tensor, three boxes per plane completing on one
//     mbarrier) in a ring of TH_NS stages, two CTAs per SM: ~180 KB in flight per SM and no per-thread copy instructions (the
//     cp.async version issued five 8-byte copies per thread and plane and stopped at 66 % of the copy bandwidth);
//   * no CTA-wide barrier: a stage is refilled by the last warp that finishes with it (atomic warp count).
// ---------------------------------------------------------------------------------------------------------------------------------
constexpr int TH_TX = 64, TH_TY = 16, TH_NT = 256, TH_NS = 3, TH_UX = TH_TX + 2 * GX;

__device__ __forceinline__ uint32_t th_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void th_mbar_init(uint32_t a, uint32_t cnt) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(cnt) : "memory"); }
__device__ __forceinline__ void th_mbar_expect_tx(uint32_t a, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory"); }
__device__ __forceinline__ void th_mbar_wait(uint32_t a, uint32_t parity) {
    asm volatile("{\n\t.reg .pred P1;\n\tTH_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra TH_DONE;\n\tbra TH_WAIT;\n\tTH_DONE:\n\t}" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void th_tma_3d(uint32_t dst, const CUtensorMap *map, uint32_t mbar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"((uint64_t)map), "r"(mbar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
#ifdef CUDNS_F32
__device__ __forceinline__ void th_stg2(real *p, real a, real b) { asm volatile("st.global.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory"); }
__device__ __forceinline__ real2 make_real2(real a, real b) { return make_float2(a, b); }
#else
__device__ __forceinline__ void th_stg2(real *p, real a, real b) { asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory"); }
__device__ __forceinline__ real2 make_real2(real a, real b) { return make_double2(a, b); }
#endif

template <int V> struct ThCfg {
    static constexpr int VY = TH_TY + 2 * V;
    static constexpr int SU = TH_TY * TH_UX, SV = VY * TH_TX, SW = TH_TY * TH_TX;     // doubles per stage
    static constexpr int STAGE = SU + SV + SW;
    static constexpr size_t bytes = (size_t)TH_NS * STAGE * sizeof(real) + 64;
    static_assert((SU * sizeof(real)) % 128 == 0 && (SV * sizeof(real)) % 128 == 0 && (SW * sizeof(real)) % 128 == 0, "TMA destinations must stay 128-byte aligned");
};

// GEN adds what the wall-bounded set-ups need (BCxderVel boundary_condition_x.h:38-40,67,94; BCzderVel boundary_condition_z.h:34-40):
// the ghost columns of u at a wall / the free stream are rebuilt in the landed plane by the warp that owns the row (they depend on
// that row only), du/dx carries the metric xp, and the planes next to the global z ends of the boundary layer take the slow
// extrapolating dw/dz; planes outside the global z range are skipped (the stage kernel extrapolates theta there itself)
template <int V, bool GEN>
__global__ void __launch_bounds__(TH_NT, 2)
theta_tma_kernel(const __grid_constant__ KConst c, real *__restrict__ theta, const real *__restrict__ wfield, int zchunk,
                 const __grid_constant__ ThetaMaps tm) {
    using G = ThCfg<V>;
    extern __shared__ __align__(1024) real sth[];
    uint64_t *mbar_p = (uint64_t *)(sth + (size_t)TH_NS * G::STAGE);
    int *cnt = (int *)(mbar_p + TH_NS);

    const Layout &L = c.L;
    const int tid = threadIdx.x, lane = tid & 31, ty = tid >> 5;                 // rows ty and ty + 8
    const int i0 = blockIdx.x * TH_TX, j0 = blockIdx.y * TH_TY;
    const int i = i0 + 2 * lane;
    const int kfirst = -V + (int)blockIdx.z * zchunk;
    const int klast = min(kfirst + zchunk, L.mz + V);                             // exclusive
    const uint32_t mb0 = th_smem_u32(mbar_p), s0 = th_smem_u32(sth);
    const bool perx = !GEN || c.periodicX != 0, bl = GEN && c.boundaryLayer != 0;
    const bool xlo = !perx && i0 == 0, xhi = !perx && (i0 + TH_TX >= L.mx);
    const int nxt = min(TH_TX, L.mx - i0);
    const int kglob_lo = -c.kstart, kglob_hi = c.mz_tot - c.kstart;
    real xp0 = RC(1.0), xp1 = RC(1.0);
    if constexpr (GEN) { if (c.nonUniformX) { xp0 = c.xp[min(i, L.mx - 1)]; xp1 = c.xp[min(i + 1, L.mx - 1)]; } }
    if (tid == 0) {
        for (int n = 0; n < TH_NS; n++) { th_mbar_init(mb0 + 8 * n, 1); cnt[n] = 0; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // plane kk (u, v at kk; w at kk + V) -> ring stage st
    auto issue = [&](int kk, int st) {
        const uint32_t mb = mb0 + 8 * st, su = s0 + (uint32_t)(st * G::STAGE * sizeof(real)), sv = su + G::SU * (uint32_t)sizeof(real), sw = sv + G::SV * (uint32_t)sizeof(real);
        th_mbar_expect_tx(mb, (uint32_t)(G::STAGE * sizeof(real)));
        th_tma_3d(su, &tm.u, mb, i0, j0 + L.gy, kk + L.gz);
        th_tma_3d(sv, &tm.v, mb, i0 + GX, j0 + L.gy - V, kk + L.gz);
        th_tma_3d(sw, &tm.w, mb, i0 + GX, j0 + L.gy, kk + V + L.gz);
    };
    if (tid == 0) {
        for (int n = 0; n < TH_NS; n++) if (kfirst + n < klast) issue(kfirst + n, n);
    }
    // w of the 2V planes below the first one: this thread's four columns straight from global memory (16-byte aligned: i is even)
    real2 wr[2][2 * V + 1];                                  // wr[r][V + l] = w of plane k+l, rows ty (r = 0) and ty + 8
    const bool act[2] = {i < L.mx && j0 + ty < L.my, i < L.mx && j0 + ty + 8 < L.my};
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const real *pw = wfield + L.idx(act[r] ? i : 0, act[r] ? j0 + ty + 8 * r : 0, kfirst - V);
#pragma unroll
        for (int m = 0; m < 2 * V; m++) wr[r][m + 1] = *reinterpret_cast<const real2 *>(pw + (size_t)m * L.plane);
    }
    int st = 0; uint32_t par = 0;
    for (int k = kfirst; k < klast; k++) {
        th_mbar_wait(mb0 + 8 * st, par);
        real *su = sth + (size_t)st * G::STAGE;
        const real *sv = su + G::SU, *sw = sv + G::SV;
        const bool outside = bl && (k < kglob_lo || k >= kglob_hi);
        if constexpr (GEN) {
            if ((xlo || xhi) && !outside && lane < 2 * V) {
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    real *row = su + (ty + 8 * r) * TH_UX;
                    if (lane < V) {
                        if (xlo) {
                            const int gq = V - lane;                          // ghost -gq
                            real val = -row[GX + gq - 1], pv2;
                            if (bl && c.perturbed && perturb_theta(c, j0 + ty + 8 * r, k + c.kstart, pv2)) val = pv2;
                            row[GX - gq] = val;
                        }
                    } else if (xhi) {
                        const int gq = lane - V + 1, last = GX + nxt - 1;
                        row[last + gq] = bl ? RC(2.0) * row[last] - row[last - gq] : -row[last - gq + 1];
                    }
                }
            }
            if (xlo || xhi) __syncwarp();
        }
        real2 th[2];
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const int row = ty + 8 * r;
#pragma unroll
            for (int m = 0; m < 2 * V; m++) wr[r][m] = wr[r][m + 1];
            wr[r][2 * V] = *reinterpret_cast<const real2 *>(sw + row * TH_TX + 2 * lane);
            // u: cells i-4 .. i+5 of the row as five aligned pairs; pt0 uses offsets -l..l around cell 0, pt1 around cell 1
            const real2 *ur = reinterpret_cast<const real2 *>(su + row * TH_UX + GX + 2 * lane);
            real uc[10];
#pragma unroll
            for (int m = 0; m < 5; m++) { const real2 t = ur[m - 2]; uc[2 * m] = t.x; uc[2 * m + 1] = t.y; }     // uc[4] = cell i, uc[5] = cell i+1
            const real2 *vr = reinterpret_cast<const real2 *>(sv + (V + row) * TH_TX + 2 * lane);
            real dudx0 = 0.0, dudx1 = 0.0, dvdy0 = 0.0, dvdy1 = 0.0, dwdz0 = 0.0, dwdz1 = 0.0;
#pragma unroll
            for (int l = 1; l <= V; l++) {
                dudx0 = fma(c.c1[0][l], uc[4 + l] - uc[4 - l], dudx0);
                dudx1 = fma(c.c1[0][l], uc[5 + l] - uc[5 - l], dudx1);
                const real2 vp = vr[l * (TH_TX / 2)], vm = vr[-l * (TH_TX / 2)];
                dvdy0 = fma(c.c1[1][l], vp.x - vm.x, dvdy0);
                dvdy1 = fma(c.c1[1][l], vp.y - vm.y, dvdy1);
                dwdz0 = fma(c.c1[2][l], wr[r][V + l].x - wr[r][V - l].x, dwdz0);
                dwdz1 = fma(c.c1[2][l], wr[r][V + l].y - wr[r][V - l].y, dwdz1);
            }
            if constexpr (GEN) {
                if (bl && act[r] && !outside && (k - kglob_lo < V || kglob_hi - 1 - k < V)) {
                    dwdz0 = dwdz_edge<V>(c, wfield, i, j0 + row, k, kglob_lo, kglob_hi);
                    dwdz1 = dwdz_edge<V>(c, wfield, i + 1, j0 + row, k, kglob_lo, kglob_hi);
                }
                th[r] = make_real2(fma(dudx0, xp0, dvdy0) + dwdz0, fma(dudx1, xp1, dvdy1) + dwdz1);
            } else {
                th[r] = make_real2((dudx0 + dvdy0) + dwdz0, (dudx1 + dvdy1) + dwdz1);
            }
        }
        // this warp is done with the stage: the last of the eight refills it with plane k + TH_NS
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            if (atomicAdd(cnt + st, 1) == TH_NT / 32 - 1) {
                cnt[st] = 0; __threadfence_block();
                if (k + TH_NS < klast) issue(k + TH_NS, st);
            }
        }
        // theta and its periodic images (perBCx / perBCy, boundary.h:38-46): rows as 128-bit stores, x images point by point
#pragma unroll
        for (int r = 0; r < 2; r++) {
            if (!act[r] || outside) continue;
            const int j = j0 + ty + 8 * r;
            real *pt = theta + L.idx(i, j, k);
            th_stg2(pt, th[r].x, th[r].y);
            if (j < V) th_stg2(pt + (size_t)L.my * L.px, th[r].x, th[r].y);
            if (j >= L.my - V) th_stg2(pt - (size_t)L.my * L.px, th[r].x, th[r].y);
            if (perx) {
                if (i < V) pt[L.mx] = th[r].x;
                if (i + 1 < V) pt[L.mx + 1] = th[r].y;
                if (i >= L.mx - V) pt[-(ptrdiff_t)L.mx] = th[r].x;
                if (i + 1 >= L.mx - V) pt[1 - (ptrdiff_t)L.mx] = th[r].y;
            }
        }
        if (++st == TH_NS) { st = 0; par ^= 1u; }
    }
}

}  // namespace

int theta_tma_smem_bytes(int v) {
    switch (v) { case 1: return (int)ThCfg<1>::bytes; case 2: return (int)ThCfg<2>::bytes; case 3: return (int)ThCfg<3>::bytes; default: return (int)ThCfg<4>::bytes; }
}
// even mx; q = the padded state whose fields 1..3 the descriptors of `maps` describe
void launch_theta_tma(const KConst &kc, const real *q, real *theta, const ThetaMaps &maps, cudaStream_t st) {
    const int gx = (kc.L.mx + TH_TX - 1) / TH_TX, gy = (kc.L.my + TH_TY - 1) / TH_TY;
    const int nk = kc.L.mz + 2 * kc.v;
    // z chunks: every chunk re-reads 2V planes of w; whole waves of 148 SMs x 2 CTAs (512^3: 8 chunks = 6.9 waves, 0.70 ms; the 4
    // chunks = 3.5 waves of the first version: 0.74 ms, profiles/r02_zchunk_sweep.log)
    // (grids with fewer tiles than SMs -- the wall-bounded cases -- need short chunks to fill the machine at all)
    int nzc = pick_zchunks(gx * gy, nk, kc.v, 148 * 2, gx * gy >= 148 ? 64 : 8);
    if (const char *e = getenv("CUDNS_THETA_ZCHUNKS")) { const int n = atoi(e); if (n >= 1 && nk / n >= 8) nzc = n; }      // experiments
    int zchunk = (nk + nzc - 1) / nzc;
    nzc = (nk + zchunk - 1) / zchunk;
    dim3 grid(gx, gy, nzc);
    const real *w = q + 3 * kc.L.vol;
    const bool gen = !(kc.periodicX && !kc.nonUniformX && !kc.boundaryLayer);
#define CUDNS_THETA_TMA_CASE(VV)                                                                    \
    {                                                                                               \
        if (gen) {                                                                                  \
            opt_in_smem<theta_tma_kernel<VV, true>>((int)ThCfg<VV>::bytes);                         \
            theta_tma_kernel<VV, true><<<grid, TH_NT, ThCfg<VV>::bytes, st>>>(kc, theta, w, zchunk, maps);   \
        } else {                                                                                    \
            opt_in_smem<theta_tma_kernel<VV, false>>((int)ThCfg<VV>::bytes);                        \
            theta_tma_kernel<VV, false><<<grid, TH_NT, ThCfg<VV>::bytes, st>>>(kc, theta, w, zchunk, maps);  \
        }                                                                                           \
    }
    switch (kc.v) {
        case 1: CUDNS_THETA_TMA_CASE(1) break;
        case 2: CUDNS_THETA_TMA_CASE(2) break;
        case 3: CUDNS_THETA_TMA_CASE(3) break;
        default: CUDNS_THETA_TMA_CASE(4) break;
    }
#undef CUDNS_THETA_TMA_CASE
}

void launch_theta_march(const KConst &kc, const real *q, real *theta, cudaStream_t st) {
    const int gx = (kc.L.mx + TXT - 1) / TXT, gy = (kc.L.my + TYT - 1) / TYT;
    const int nk = kc.L.mz + 2 * kc.v;
    // z chunks: every chunk pays a 2V-plane prologue of w (registers only: cheap); aim at several waves of 148 SMs x 3 CTAs, chunks of
    // at least 6 planes (profiles/r02_theta_march_chunks.log: channel 160x192x192 0.081 -> 0.070 ms with 32 instead of 4 chunks)
    int nzc = (148 * 3 * 8 + gx * gy - 1) / (gx * gy);
    nzc = nzc < nk / 6 ? nzc : nk / 6;
    nzc = nzc < 1 ? 1 : nzc;
    if (const char *e = getenv("CUDNS_THETA_ZCHUNKS")) { const int n = atoi(e); if (n >= 1 && nk / n >= 2 * kc.v + 1) nzc = n; }   // experiments
    int zchunk = (nk + nzc - 1) / nzc;
    nzc = (nk + zchunk - 1) / zchunk;
    dim3 grid(gx, gy, nzc);
    const bool gen = !(kc.periodicX && !kc.nonUniformX && !kc.boundaryLayer);
#define CUDNS_THETA_CASE(VV)                                                                        \
    {                                                                                               \
        const size_t sm = (size_t)NST * (TYT * UX + (TYT + 2 * VV) * TXT + NTT) * sizeof(real);   \
        opt_in_smem<theta_march_kernel<VV, true>>((int)sm);                                         \
        opt_in_smem<theta_march_kernel<VV, false>>((int)sm);                                        \
        if (gen) theta_march_kernel<VV, true><<<grid, NTT, sm, st>>>(kc, q, theta, zchunk);         \
        else theta_march_kernel<VV, false><<<grid, NTT, sm, st>>>(kc, q, theta, zchunk);            \
    }
    switch (kc.v) {
        case 1: CUDNS_THETA_CASE(1) break;
        case 2: CUDNS_THETA_CASE(2) break;
        case 3: CUDNS_THETA_CASE(3) break;
        default: CUDNS_THETA_CASE(4) break;
    }
#undef CUDNS_THETA_CASE
}

}  // namespace cudns
