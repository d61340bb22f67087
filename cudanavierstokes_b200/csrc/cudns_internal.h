// Internal declarations of libcudns (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>

// Host-side helpers shared by both precisions (defined once, host_setup.cpp)
struct cudns_params;
namespace cudns_shared {
void set_error(const std::string &msg);
const double *coeff_first(int s);
const double *coeff_second(int s);
int check_params(const cudns_params *p);
}  // namespace cudns_shared

// ---- precision (`myprec` of the reference, src/globals.h:5-6).  The device side -- kernels, coefficient tables, the solver object --
// is compiled twice: as is (real = double) and with -DCUDNS_F32 (real = float; the translation units that carry it are listed in the
// Makefile).  The single-precision copy lives in namespace cudns32 and exports its C entry points as cudns32_* (cudns_abi.h); the
// public symbols of include/cudns.h dispatch on the precision recorded in the handle (abi_dispatch.cpp).  Host arrays that cross
// the C ABI are double in both builds, like the reference's (copyField casts, cuda_utils.cu:317-355).
#ifdef CUDNS_F32
#define cudns cudns32
#endif
#include "cudns_abi.h"
#include "../../include/cudns.h"

namespace cudns {
using namespace cudns_shared;
#ifdef CUDNS_F32
typedef float real;
typedef float2 real2;
typedef float4 vec16;                                   // 16 bytes of state
#define CUDNS_TMA_REAL CU_TENSOR_MAP_DATA_TYPE_FLOAT32
#else
typedef double real;
typedef double2 real2;
typedef double2 vec16;
#define CUDNS_TMA_REAL CU_TENSOR_MAP_DATA_TYPE_FLOAT64
#endif
constexpr int VEC16 = 16 / (int)sizeof(real);           // state elements per 16 bytes
#define RC(x) ((real)(x))                               // literal of the working precision

constexpr int GX = 4;            // x ghost width in memory (>= s, even: keeps interior rows 16B aligned)
constexpr int MAXS = 4;

// Padded ghost-cell layout of one field: [pz][py][px], x fastest.
struct Layout {
    int mx, my, mz;              // local interior extents (mz = this rank's slab)
    int gy, gz;                  // ghost widths: gy = s, gz = s + v
    int px, py, pz;              // padded extents
    size_t plane, vol;           // px*py, plane*pz
    __host__ __device__ inline size_t idx(int i, int j, int k) const {
        return (size_t)(k + gz) * plane + (size_t)(j + gy) * px + (size_t)(i + GX);
    }
};

// Everything a kernel needs besides field pointers; passed by value (lives in the constant bank).
struct KConst {
    Layout L;
    int s, v;
    int kstart;                  // global index of local plane 0 (perturbation strip, sponge)
    int mz_tot;
    real d1[3], d2[3];         // 1/Delta, 1/Delta^2 per direction (d_dx.. d_d2z, cuda_utils.cu:61-63,88-90)
    real aF[MAXS + 1];         // advective first-derivative weights a_l, l=1..s ( = -coeffF[s-l], globals.h:69-82)
    real aV[MAXS + 1];         // viscous   first-derivative weights
    real bV[MAXS + 1];         // viscous second-derivative weights b_0..b_v ( = coeffVS[v-l] )
    // the same weights pre-scaled by the grid spacing of direction d: c1 = a_l/dx_d (viscous order; dilatation pass)
    real c1[3][MAXS + 1];
    // stage kernels: cf[d][l] = { -a_l/(4 dx_d), -a_l/dx_d (advective order), a_l/dx_d, b_l/dx_d^2 (viscous order) },
    // c1t = a_l/(3 dx_d) (viscous order), c20sum = sum_d b_0/dx_d^2
    real cf[3][MAXS + 1][4], c1t[3][MAXS + 1], c20sum;
    real cfzp[MAXS + 1];       // -a_l Rgas / dz: z pressure gradient from rho*T (ring without p, wide variant)
    real cfp[3][MAXS + 1];     // -a_l Rgas / dx_d: pressure gradient from rho*T in every direction (fast variant)
    real gam, Rgas, cvInv, invRe, lamfac, viscexp;
    int viscmode;                // 0 generic pow, 1 n=1, 2 n=0.5, 3 n=0.75, 4 n=1.5
    int periodicX, boundaryLayer, nonUniformX, perturbed, forcing, quirk_q1;
    real TwallTop, TwallBot;
    // perturbation.h
    int kC, LP; real amp1, amp2, omega1, omega2, lambdaP;
    real Lx, Ly, Lz, CFL;
    const real *xp;            // [mx]  1/x'(xi)           (device)
    const real *cVSx;          // [(2v+1)*mx] non-uniform second-derivative table (device)
    const real *dxv;           // [mx] cell widths          (device)
    const real *spongeX, *spongeZ, *sref;   // [mx], [mz], [5][mx*mz]
    const double *dt, *dpdz, *time_on_gpu;  // device scalars: double in both builds, like every reduced scalar and statistic (the
                                            // cross-rank all-reduce callback moves doubles)
};

// One Runge-Kutta stage as a generic register update (see DESIGN.md "RK algebra"):
//   K      = RHS(Q_in)
//   Q_out  = Q_base + dt*( cN*K + cA*RA + cB*RB )        (conservative; stored primitive)
//   RW     = wOld*RW + wNew*K                             (skipped when RW == nullptr)
struct StageCoef {
    real cN, cA, cB, wOld, wNew;
};

struct StagePtrs {
    const real *qin;           // 5 padded fields, stride L.vol
    const real *qbase;         // 5 padded fields (may equal qin)
    real *qout;                // 5 padded fields
    const real *theta;         // padded
    const real *RA, *RB;       // 5 unpadded fields each (stride N) or nullptr
    real *RW;                  // 5 unpadded fields or nullptr
    real *qout_lo, *qout_hi;   // lean kernel: the SAME output buffer on the lower / upper z neighbour (peer memory over NVLink,
                                 // or qout itself for the periodic wrap on one device); the stage kernel stores its first / last
                                 // gz planes straight into the neighbour's ghost planes.  nullptr: no such neighbour / not connected
    real *rhs_out;             // test path: write K only (5 unpadded) and skip the update
};

// TMA descriptors of one state buffer (+ theta) for the stage kernel: the tile with its x/y stencil halos
// ("box", (32+2*GX) x (TY+2s) x 1 plane x 5 fields) and the tile interior ("int", 32 x TY x 1 x 5)
struct StageMaps {
    CUtensorMap qbox, qint, thbox, thint;
};
// ---- lean stage kernel (stage_lean.inc): tile rows per CTA for the linear-viscosity (8 quantities) and the general
// (9 quantities) variants, and its TMA descriptors
#ifndef CUDNS_LEAN_TY_LINEAR
#define CUDNS_LEAN_TY_LINEAR 12
#endif
#ifndef CUDNS_LEAN_TY_GENERAL_LOW
#define CUDNS_LEAN_TY_GENERAL_LOW 12
#endif
// 9 quantities: 8 rows where the rings of 12 would not fit tensor memory (s = 4) or the register cap of two resident CTAs (float)
constexpr int lean_ty_general(int s) { return (sizeof(real) == 8 && s <= 3) ? CUDNS_LEAN_TY_GENERAL_LOW : 8; }
struct LeanMaps {
    CUtensorMap qbox, qint, thbox, thint;   // input state / theta: halo'd tile and tile interior
    CUtensorMap qbint;                      // base state, tile interior (Kutta RK3 / RK4)
    CUtensorMap opa, opb;                   // RA ; RB or the old RW (unpadded register arrays, tile interior)
};
// wide: the 16-warp variant (tile 32 x 16; FAST + linear viscosity + at most the RA operand tile: see lean_wide_ok)
void launch_rhs_stage_lean(const KConst &kc, const StagePtrs &p, const StageCoef &c, const LeanMaps &maps, bool wide, cudaStream_t st);
bool lean_wide_ok(const KConst &kc);
// fourth-generation kernel (stage_fast.cu): same preconditions and TMA descriptors as the wide lean variant
// (stage_fast.cu).  Its state buffers carry 8 padded fields: rho,u,v,w,rho*E, H,T (written with the state) and theta.
constexpr int FAST_NFB = 8;
struct FastMaps {
    CUtensorMap q4box, a3box;               // halo'd tile of (rho,u,v,w) and of (H,T,theta)
    CUtensorMap q4int, a3int;               // tile interior of the same (plane S ahead, for the z ring)
    CUtensorMap eint;                       // rho*E, tile interior
    CUtensorMap opa;                        // RA (unpadded register array, tile interior)
};
void launch_rhs_stage_fast(const KConst &kc, const StagePtrs &p, const StageCoef &c, const FastMaps &maps, int ty, cudaStream_t st);
// fifth-generation kernel (stage_duo.inc): tile 64 x 8, two x-adjacent points per thread; same 8-field state buffers; needs an even mx
constexpr int DUO_TX = 64, DUO_TY = 8;
// rows of its halo'd plane: 8 + 2s, rounded up until one field of the box is a multiple of 128 bytes (TMA destinations; single
// precision with s = 1, 3 only)
constexpr int duo_box_rows(int s) { int cy = DUO_TY + 2 * s; while (((DUO_TX + 2 * GX) * cy * (int)sizeof(real)) % 128) cy++; return cy; }
// rows of the lean kernels' halo'd plane (40 cells wide): ty + 2s, rounded up the same way
constexpr int lean_box_rows(int ty, int s) { int cy = ty + 2 * s; while ((40 * cy * (int)sizeof(real)) % 128) cy++; return cy; }
struct DuoMaps {
    CUtensorMap q4box, a3box;               // halo'd tile (72 x (8+2s)) of (rho,u,v,w) and of (H,T,theta)
    CUtensorMap q4int, a3int;               // tile interior (64 x 8) of the same (plane S ahead, for the z ring)
};
void launch_rhs_stage_duo(const KConst &kc, const StagePtrs &p, const StageCoef &c, const DuoMaps &maps, cudaStream_t st);
int duo_smem_bytes(int s);
void launch_derive_aux(const KConst &kc, real *q8, cudaStream_t st);
int fast_smem_bytes(int s, int ty);
int lean_smem_wide_bytes(int s);
#define CUDNS_LEAN_TY_WIDE 16
int lean_smem_bytes(int s, bool linear_visc);

void launch_theta(const KConst &kc, const real *q, real *theta, cudaStream_t st);
// TMA variant of the dilatation pass (theta.cu) for the periodic / uniform set-ups with an even mx: boxes of u (72 x 16, x halos),
// v (64 x (16 + 2v), y halos) and w (64 x 16) of one padded state buffer
constexpr int THETA_TX = 64, THETA_TY = 16;
struct ThetaMaps { CUtensorMap u, v, w; };
void launch_theta_tma(const KConst &kc, const real *q, real *theta, const ThetaMaps &maps, cudaStream_t st);
int theta_tma_smem_bytes(int v);
void launch_fill_xy(const KConst &kc, real *q5, int nfields, cudaStream_t st);
void launch_zwrap(const KConst &kc, real *q5, int nfields, cudaStream_t st);
void launch_pack_z(const KConst &kc, const real *q5, real *send_lo, real *send_hi, cudaStream_t st);
void launch_unpack_z(const KConst &kc, real *q5, const real *recv_lo, const real *recv_hi, cudaStream_t st);
void launch_pad(const KConst &kc, const double *src5[5], real *q5, cudaStream_t st);      // the caller's arrays are double in both builds
void launch_unpad(const KConst &kc, const real *q5, double *dst5[5], cudaStream_t st);
// reductions: out[0] = max_p conv limiter, out[1] = max_p visc limiter (uses fresh mu), out[2..] sums
void launch_dt_reduce(const KConst &kc, const real *q, double *out2, cudaStream_t st);
// scratch: bulk_scratch_doubles() doubles owned by the calling solver (block partials, then the completion counter; zeroed once)
void launch_bulk_reduce(const KConst &kc, const real *q, double *out4, double *scratch, cudaStream_t st);
int bulk_scratch_doubles();
// mean square vorticity of a periodic box (same scratch)
void launch_enstrophy_reduce(const KConst &kc, const real *q, double *out, double *scratch, cudaStream_t st);
void launch_scalar_ops(int op, double *a, const double *b, const double *c, cudaStream_t st);
// wall-normal profiles / friction Reynolds number (calcAvgChan, printRes): see kernels.cu
void launch_profile_partial(const KConst &kc, const real *q, const double *mean, double *partial, int pass, cudaStream_t st);
void launch_profile_combine(const KConst &kc, const double *partial, double *out, double scale, cudaStream_t st);
void launch_profile_favre(const KConst &kc, double *mean, cudaStream_t st);
int profile_partial_doubles(const KConst &kc);
void launch_retau(const KConst &kc, const real *q, double *partial, double *out, double scale, cudaStream_t st);
// post-processing statistics (postproc/post.cpp): see kernels.cu.  acc[13][mx] += scale * (sums | squared deviations from mean)
int post_partial_doubles(const KConst &kc);
void launch_post_accumulate(const KConst &kc, const real *q, const double *mean, double *partial, double *acc, double scale, int pass, cudaStream_t st);
void launch_post_ret(const KConst &kc, const real *q, double host_dx, double *partial, double *ret2, double scale, cudaStream_t st);
void launch_post_finish_mean(const KConst &kc, double *mean, double *bulk, double *ret2, double inv_files, cudaStream_t st);
// cross-GPU stage hand-shake over peer memory: store `epoch` into the two neighbours' mailbox slots / spin until both own slots reach it
void launch_halo_signal(unsigned long long *peer_lo_slot, unsigned long long *peer_hi_slot, unsigned long long epoch, cudaStream_t st);
// the wait gives up after timeout_ns (device global timer) and stores `epoch` into *err_word (0 = never timed out): a neighbour that
// died or runs a different number of stages must not hang the device for good
void launch_halo_wait(const unsigned long long *my_slots, int need_lo, int need_hi, unsigned long long epoch, unsigned long long timeout_ns,
                      unsigned long long *err_word, cudaStream_t st);

// opt a kernel in to more than 48 KB of dynamic shared memory.  The attribute belongs to the (function, device) pair and several
// devices may be driven from one process (cudns_peer_info.local_ptr), so it is remembered per device, under a lock
template <auto Kernel>
inline void opt_in_smem(int bytes) {
    static std::mutex m;
    static unsigned long long done = 0;                  // bit d: set on device d
    int dev = 0; cudaGetDevice(&dev);
    std::lock_guard<std::mutex> g(m);
    if (dev < 64 && ((done >> dev) & 1ull)) return;
    cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (dev < 64) done |= 1ull << dev;
}

bool rhs_stage_supported(int s, int v);

// z chunks of a marching stage kernel: `cols` tiles, each split into n chunks of mz/n planes.  Every chunk re-primes its ring (2s
// planes of loads, worth ~0.8 s planes of work) and the grid runs in waves of `resident` CTAs (148 SMs x CTAs per SM); with w =
// CTAs / resident the run takes max(ceil(w), w + 1) chunk times -- whole waves when every CTA takes the same time, about one CTA
// time of tail otherwise (wall tiles, unequal clocks).  Fitted to a sweep on one B200 (profiles/r02_zchunk_sweep.log: channel
// 160x192x192 0.59 -> 0.53 ms per stage, boundary layer and Taylor-Green 512^3 unchanged within 1.5 %).  Any n is allowed, not only
// powers of two; chunks stay >= 16 planes (the bandwidth-bound dilatation pass: >= 64, its tail wave costs less).
// whole_waves: weigh the count of whole waves and the w + 1 rule equally -- the wall-bounded set-ups of the lean kernel, whose CTAs
// take nearly equal times (profiles/r02_channel_stage_chunks.log: channel 160x192x192, 80 tiles: 9 or 11 chunks = 4.9 / 5.9 waves
// 0.411 ms, 12 chunks = 6.5 waves 0.431 ms)
inline int pick_zchunks(int cols, int mz, int s, int resident, int min_chunk = 16, bool whole_waves = false) {
    if (const char *e = getenv("CUDNS_ZCHUNKS")) { const int n = atoi(e); if (n >= 1 && mz / n >= 2 * s + 1) return n; }   // experiments
    int best = 1; double best_cost = 1e300;
    for (int n = 1; n <= 128; n++) {
        const int zc = (mz + n - 1) / n;
        if (n > 1 && zc < min_chunk) break;
        const int nn = (mz + zc - 1) / zc;                         // chunks this plane count really gives
        const long ctas = (long)cols * nn;
        const double w = (double)ctas / resident, waves = (double)((ctas + resident - 1) / resident);
        const double tail = waves > w + 1.0 ? waves : w + 1.0;
        const double cost = (whole_waves ? 0.5 * waves + 0.5 * (w + 1.0) : tail) * (zc + 0.8 * s);
        if (cost < best_cost * 0.995) { best_cost = cost; best = nn; }
    }
    return best;
}


}  // namespace cudns
