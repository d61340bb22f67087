// libcudns C ABI: solver life cycle, time loop, halo plumbing.  See include/cudns.h for the mapping of
// every entry point onto the reference's functions.
#include "cudns_internal.h"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <unistd.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <condition_variable>
#include <mutex>
#include <thread>

namespace cudns {
void launch_dt_combine(double *dt, const double *conv, const double *visc, double cfl, cudaStream_t st);
}
using namespace cudns;
static constexpr bool kF32 = sizeof(real) == 4;     // this translation unit is the single-precision copy (see cudns_internal.h)

// inside cudns_create, once the solver object exists: a failing CUDA call must not leak it
#define CKC(call)                                                                                    \
    do {                                                                                             \
        cudaError_t err__ = (call);                                                                  \
        if (err__ != cudaSuccess) {                                                                  \
            set_error(std::string(#call) + ": " + cudaGetErrorString(err__));                        \
            cudns_destroy(S);                                                                        \
            return CUDNS_ECUDA;                                                                      \
        }                                                                                            \
    } while (0)
#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t err__ = (call);                                                                  \
        if (err__ != cudaSuccess) {                                                                  \
            set_error(std::string(#call) + ": " + cudaGetErrorString(err__));                        \
            return CUDNS_ECUDA;                                                                      \
        }                                                                                            \
    } while (0)

// device scalar slots
enum { SC_DT = 0, SC_DPDZ, SC_TGPU, SC_TIME, SC_RED0, SC_RED1, SC_STALE0, SC_STALE1, SC_BULK0, SC_BULK1, SC_BULK2, SC_BULK3,
       SC_HALOERR /* u64: stage number of a hand-shake that timed out, 0 = none */, SC_ENS /* mean square vorticity */, SC_N = 16 };

// asynchronous fields/ writer: one snapshot in flight
struct IoState {
    cudaStream_t cs = nullptr;          // copy stream
    double *d_stage = nullptr;          // compact copy of the slab on the device [5][mz][my][mx]
    double *h_stage = nullptr;          // the same in pinned host memory
    cudaEvent_t snap_done = nullptr, d2h_done = nullptr;
    std::thread worker;
    std::mutex m;
    std::condition_variable cv;
    bool started = false, stop = false, busy = false, have_job = false;
    std::string dir; int timestep = 0;
    int err = 0; std::string errmsg;
    uint64_t files_written = 0;
};

// per-kernel device times of the step loop under its real (sustained, possibly power-capped) conditions: CUDA events around the
// dilatation pass, the stage kernel and the hand-shake of every stage of a cudns_advance call, read back after its final sync
struct StageTimer {
    std::vector<cudaEvent_t> ev;        // 4 per stage
    size_t used = 0;
    double theta_ms = 0, stage_ms = 0, halo_ms = 0;
    uint64_t n = 0;
    bool on = false;
};

struct cudns_solver {
    int precision;               // FIRST member of both precisions' solver objects: 0 double, 1 float (abi_dispatch.cpp reads it)
    IoState *io;
    StageTimer *tm;
    cudns_params P;
    KConst kc;
    Layout L;
    size_t N;                    // local interior points
    cudaStream_t st;
    real *block;               // ONE allocation: nstate padded 5-field buffers + the neighbour mailbox (a single IPC handle covers it)
    size_t block_doubles;        // state part of the block, in doubles; the mailbox (8 x u64) follows
    real *state[3];            // padded 5-field buffers inside block
    int nstate, cur;
    // peer-memory halo transport (cudns_halo_connect): the neighbours' blocks mapped into this process
    real *peer_lo, *peer_hi;   // nullptr: not connected / no such neighbour
    void *ipc_lo, *ipc_hi;       // what cudaIpcOpenMemHandle returned (to close), nullptr for same-process peers
    bool connected;
    unsigned long long epoch;    // stage counter of the hand-shake
    real *theta;
    real *R1, *R2;
    real *d_xp, *d_cVSx, *d_dxv, *d_spx, *d_spz, *d_sref;
    double *d_scal;
    double *d_hist; int hist_cap;
    double *d_bulk;              // bulk_reduce_kernel scratch (block partials + completion counter), this solver's own
    unsigned long long halo_timeout_ns;
    double *d_prof;              // profile diagnostics scratch: partial[64][5][mx], mean[5][mx], var[5][mx], 1 scalar (lazy)
    double *d_post;              // post-processing statistics: partial[64][13][mx], mean[13][mx], fluc[13][mx], bulk[13], Re_tau, u_tau (lazy)
    int post_phase, post_files, post_added;   // 0 idle, 1 collecting means, 2 means final / collecting fluctuations, 3 fluctuations final
    real *send_lo, *send_hi, *recv_lo, *recv_hi; size_t halo_doubles;
    bool have_state, fixed_dt, have_sponge;
    cudns_allreduce_fn allreduce; void *allreduce_user;
    cudns_exchange_fn exchange; void *exchange_user;
    uint64_t launches, stages;
    size_t bytes;
    StageMaps lmaps[3];          // TMA descriptors of state[b] (+ theta), tile of the lean kernel
    CUtensorMap rmap[2];         // R1 / R2 (unpadded register arrays), tile interior of the lean kernel
    StageMaps wmaps[3];          // tile of the wide (16-warp) lean kernel
    CUtensorMap wrmap[2];
    bool wide;                   // the wide variant applies to this configuration (CUDNS_WIDE=0 disables it)
    bool fast;                   // ... and is served by the fourth-generation kernel (stage_fast.cu; CUDNS_WIDE=1: the lean wide variant)
    int fast_ty;                 // its tile rows: 8 (two CTAs per SM, default) or 16 (CUDNS_FAST_TY=16)
    int nfb;                     // padded fields per state buffer: 5, or FAST_NFB = 8 (+ H, T, theta) when the fast kernel applies
    FastMaps fmaps[3];           // its TMA descriptors per state buffer (opa is patched per launch)
    CUtensorMap frmap[2];        // R1 / R2 with its tile
    bool aux_valid[3];           // H, T of state[b] (ghosts included) match its (rho,u,v,w,rho*E)
    bool duo;                    // the fifth-generation kernel (stage_duo.inc) serves EVERY stage of this configuration (CUDNS_DUO=0: off)
    DuoMaps dmaps[3];            // its TMA descriptors per state buffer
    bool theta_tma;              // the dilatation pass runs its TMA variant (even mx; CUDNS_THETA_TMA=0: the cp.async variant)
    ThetaMaps tmaps[3];          // u, v, w boxes of every state buffer
};

// cuTensorMapEncodeTiled through the runtime's driver entry point (libcudns does not link libcuda)
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn get_encode() {
    static encode_tiled_fn fn = nullptr;
    if (!fn) {
        void *p = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (encode_tiled_fn)p;
    }
    return fn;
}
// padded field(s) [nf][pz][py][px] of doubles -> descriptor with box bx x by x 1 (x nf)
static int make_map(CUtensorMap *m, const Layout &L, real *base, int nf, int bx, int by) {
    encode_tiled_fn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return CUDNS_ECUDA; }
    cuuint64_t dims[4] = {(cuuint64_t)L.px, (cuuint64_t)L.py, (cuuint64_t)L.pz, (cuuint64_t)nf};
    cuuint64_t strides[3] = {(cuuint64_t)L.px * sizeof(real), (cuuint64_t)L.plane * sizeof(real), (cuuint64_t)L.vol * sizeof(real)};
    cuuint32_t box[4] = {(cuuint32_t)bx, (cuuint32_t)by, 1, (cuuint32_t)nf};
    cuuint32_t es[4] = {1, 1, 1, 1};
    const int rank = nf > 1 ? 4 : 3;
    CUresult r = enc(m, CUDNS_TMA_REAL, rank, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r)); return CUDNS_ECUDA; }
    return CUDNS_OK;
}

// unpadded register array [5][mz][my][mx] -> descriptor with box 32 x ty x 1 x 5
static int make_rmap(CUtensorMap *m, const Layout &L, real *base, int ty) {
    encode_tiled_fn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return CUDNS_ECUDA; }
    cuuint64_t dims[4] = {(cuuint64_t)L.mx, (cuuint64_t)L.my, (cuuint64_t)L.mz, 5};
    cuuint64_t strides[3] = {(cuuint64_t)L.mx * sizeof(real), (cuuint64_t)L.mx * L.my * sizeof(real), (cuuint64_t)L.mx * L.my * L.mz * sizeof(real)};
    cuuint32_t box[4] = {32, (cuuint32_t)ty, 1, 5};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CUDNS_TMA_REAL, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (register array) failed with CUresult " + std::to_string((int)r)); return CUDNS_ECUDA; }
    return CUDNS_OK;
}

// host table (double) -> device table of the working precision (the cast of setGPUParameters, cuda_utils.cu:60-139)
static cudaError_t upload(real *dst, const double *src, size_t n) {
    std::vector<real> tmp(src, src + n);
    return cudaMemcpy(dst, tmp.data(), n * sizeof(real), cudaMemcpyHostToDevice);
}

template <typename T> static int dmalloc(cudns_solver *S, T **p, size_t n) {
    cudaError_t e = cudaMalloc((void **)p, n * sizeof(T));
    if (e != cudaSuccess) { set_error(std::string("cudaMalloc: ") + cudaGetErrorString(e)); return CUDNS_ENOMEM; }
    S->bytes += n * sizeof(T);
    return CUDNS_OK;
}

extern "C" {

int cudns_create(const cudns_params *p, const double *x, const double *xp, const double *xpp, cudns_handle *out) {
    int rc = check_params(p); if (rc) return rc;
    if (!out || !x || !xp || !xpp) { set_error("NULL argument"); return CUDNS_EINVAL; }
    if (!rhs_stage_supported(p->stencilSize, p->stencilVisc)) { set_error("unsupported stencil"); return CUDNS_EUNSUPPORTED; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: libcudns has no CPU fallback"); return CUDNS_ECUDA; }
    CK(cudaSetDevice(p->device));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, p->device));
    if (prop.major < 10) { set_error("libcudns is built for sm_100a (B200) only"); return CUDNS_EUNSUPPORTED; }
    cudns_solver *S = new cudns_solver();
    std::memset(S, 0, sizeof(*S));
    S->P = *p;
    S->precision = kF32 ? 1 : 0;
    const int s = p->stencilSize, v = p->stencilVisc, mx = p->mx, my = p->my, mzl = p->mz / p->nranks;
    if (!p->periodicX) {
        int rem = mx % 32;
        if (rem != 0 && rem < s + 1) { set_error("non-periodic x: mx % 32 must be 0 or > stencilSize"); delete S; return CUDNS_EINVAL; }
    }
    Layout &L = S->L;
    L.mx = mx; L.my = my; L.mz = mzl; L.gy = s; L.gz = s + v;
    L.px = ((mx + 2 * GX + 15) / 16) * 16; L.py = my + 2 * s; L.pz = mzl + 2 * L.gz;
    L.plane = (size_t)L.px * L.py; L.vol = L.plane * L.pz;
    S->N = (size_t)mx * my * mzl;
    CKC(cudaStreamCreateWithFlags(&S->st, cudaStreamNonBlocking));
    S->nstate = (p->lowStorage && !p->rk4) ? 2 : 3;
    {   // the fourth-generation stage kernel (stage_fast.cu) applies to the periodic / uniform / linear-viscosity set-ups; its state
        // buffers carry three more padded fields (H, T, theta).  CUDNS_WIDE=0 / 1 select the older kernels (A/B timing)
        const char *we = getenv("CUDNS_WIDE"), *fe = getenv("CUDNS_FAST_TY");
        S->fast_ty = (fe && std::string(fe) == "16") ? 16 : 8;
        S->fast = p->viscexp == 1.0 && p->periodicX && !p->nonUniformX && !p->boundaryLayer && !(we && (std::string(we) == "0" || std::string(we) == "1")) &&
                  (size_t)fast_smem_bytes(s, S->fast_ty) <= prop.sharedMemPerBlockOptin;
        // fifth generation: two x-adjacent points per thread, every Runge-Kutta stage shape; needs an even mx (16-byte rows)
        // It serves Kutta RK3 and RK4 by default (the fourth generation cannot: those stages fell back to the lean kernel); for
        // low-storage RK3 the fourth generation is still ~5 % faster (2 x 8 warps per SM hide more latency than 8 warps with two
        // points each): CUDNS_DUO=1 forces the fifth generation there too, CUDNS_DUO=0 switches it off
        const char *de = getenv("CUDNS_DUO");
        const bool ls3 = p->lowStorage && !p->rk4;
        const bool want = kF32 || (de ? std::string(de) != "0" : !ls3);
        S->duo = S->fast && (!fe || kF32) && mx % 2 == 0 && want && (size_t)duo_smem_bytes(s) <= prop.sharedMemPerBlockOptin;
        // `myprec float` (globals.h:5-6): the fifth generation where it applies, the lean kernel for everything else (walls,
        // stretched x, boundary layer, non-linear viscosity, odd mx); there is no single-precision fourth generation
        if (kF32 && !S->duo) S->fast = false;
        if (kF32 && !S->duo && mx % 4) { set_error("precision = 1 (float) outside the periodic / uniform / linear-viscosity set-ups needs mx % 4 == 0 (16-byte rows)"); cudns_destroy(S); return CUDNS_EINVAL; }
        S->nfb = S->fast ? FAST_NFB : 5;
    }
    S->block_doubles = (size_t)S->nstate * S->nfb * L.vol;
    S->block_doubles = (S->block_doubles + 31) / 32 * 32;                       // keep the mailbox 256-byte aligned
    if ((rc = dmalloc(S, &S->block, S->block_doubles + 32))) { cudns_destroy(S); return rc; }
    CKC(cudaMemsetAsync(S->block, 0, (S->block_doubles + 32) * sizeof(real), S->st));
    for (int b = 0; b < S->nstate; b++) S->state[b] = S->block + (size_t)b * S->nfb * L.vol;
    if ((rc = dmalloc(S, &S->theta, L.vol))) { cudns_destroy(S); return rc; }
    CKC(cudaMemsetAsync(S->theta, 0, L.vol * sizeof(real), S->st));
    if ((rc = dmalloc(S, &S->R1, 5 * S->N))) { cudns_destroy(S); return rc; }
    CKC(cudaMemsetAsync(S->R1, 0, 5 * S->N * sizeof(real), S->st));
    if (!(p->lowStorage && !p->rk4)) { if ((rc = dmalloc(S, &S->R2, 5 * S->N))) { cudns_destroy(S); return rc; } }
    if ((rc = dmalloc(S, &S->d_xp, mx)) || (rc = dmalloc(S, &S->d_dxv, mx)) || (rc = dmalloc(S, &S->d_cVSx, (size_t)mx * (2 * v + 1))) ||
        (rc = dmalloc(S, &S->d_scal, SC_N)) || (rc = dmalloc(S, &S->d_spx, mx)) || (rc = dmalloc(S, &S->d_spz, mzl)) ||
        (rc = dmalloc(S, &S->d_sref, 5 * (size_t)mx * mzl)) || (rc = dmalloc(S, &S->d_bulk, bulk_scratch_doubles()))) { cudns_destroy(S); return rc; }
    CKC(cudaMemsetAsync(S->d_bulk, 0, bulk_scratch_doubles() * sizeof(double), S->st));
    {   // a neighbour that never signals (crashed rank, different stage count) must not hang the device: see halo_wait_kernel
        const char *te = getenv("CUDNS_HALO_TIMEOUT_MS");
        const double ms = te ? atof(te) : 30000.0;
        S->halo_timeout_ns = (unsigned long long)((ms > 0 ? ms : 30000.0) * 1e6);
    }
    S->halo_doubles = 5 * (size_t)L.gz * L.plane;
    if (p->nranks > 1) {
        if ((rc = dmalloc(S, &S->send_lo, S->halo_doubles)) || (rc = dmalloc(S, &S->send_hi, S->halo_doubles)) ||
            (rc = dmalloc(S, &S->recv_lo, S->halo_doubles)) || (rc = dmalloc(S, &S->recv_hi, S->halo_doubles))) { cudns_destroy(S); return rc; }
    }
    // ---- setGPUParameters, cuda_utils.cu:60-139
    const double *cF = coeff_first(s), *cVF = coeff_first(v), *cVS = coeff_second(v);
    double dx = p->nonUniformX ? p->Lx * (1.0) / mx : x[1] - x[0];   // host global dx after initGrid (init.cpp:36,63)
    double Ly_d = p->Ly * (0.5 + 1.0) / my - p->Ly * (0.5) / my;     // y[1]-y[0] exactly as initGrid builds y (init.cpp:67-68)
    double Lz_d = p->Lz * (0.5 + 1.0) / p->mz - p->Lz * (0.5) / p->mz;
    double h_dx = 1.0 / dx, h_dy = 1.0 / Ly_d, h_dz = 1.0 / Lz_d;
    KConst &kc = S->kc;
    std::memset(&kc, 0, sizeof(kc));
    kc.L = L; kc.s = s; kc.v = v; kc.kstart = p->rank * mzl; kc.mz_tot = p->mz;
    kc.d1[0] = h_dx; kc.d1[1] = h_dy; kc.d1[2] = h_dz;
    kc.d2[0] = h_dx * h_dx; kc.d2[1] = h_dy * h_dy; kc.d2[2] = h_dz * h_dz;
    for (int l = 1; l <= s; l++) kc.aF[l] = -cF[s - l];
    for (int l = 1; l <= v; l++) { kc.aV[l] = -cVF[v - l]; kc.bV[l] = cVS[v - l]; }
    kc.bV[0] = cVS[v];
    for (int d = 0; d < 3; d++) {
        // cC = -a_l/(4 dx_d) (split-form flux sums), cP = a_l/dx_d (pressure gradient, advective order), c1 = a_l/dx_d and
        // c2 = b_l/dx_d^2 (viscous order)
        double cC[MAXS + 1] = {0}, cP[MAXS + 1] = {0}, c2[MAXS + 1] = {0};
        for (int l = 1; l <= s; l++) { cC[l] = -0.25 * kc.aF[l] * kc.d1[d]; cP[l] = kc.aF[l] * kc.d1[d]; }
        for (int l = 0; l <= v; l++) { kc.c1[d][l] = kc.aV[l] * kc.d1[d]; c2[l] = kc.bV[l] * kc.d2[d]; }
        for (int l = 0; l <= MAXS; l++) {
            kc.cf[d][l][0] = cC[l]; kc.cf[d][l][1] = -cP[l]; kc.cf[d][l][2] = kc.c1[d][l]; kc.cf[d][l][3] = c2[l];
            kc.c1t[d][l] = kc.c1[d][l] / 3.0;
        }
        if (d == 2) for (int l = 0; l <= MAXS; l++) kc.cfzp[l] = -cP[l] * (1.f / (p->gam * p->Ma * p->Ma));
        for (int l = 0; l <= MAXS; l++) kc.cfp[d][l] = -cP[l] * (1.f / (p->gam * p->Ma * p->Ma));
        kc.c20sum += c2[0];
    }
    kc.gam = p->gam;
    kc.Rgas = (1.f / (p->gam * p->Ma * p->Ma));                 // globals.h:48
    double Ec = ((p->gam - 1.f) * p->Ma * p->Ma);               // globals.h:47
    kc.cvInv = (p->gam - 1.0) / kc.Rgas;
    kc.invRe = 1.0 / p->Re; kc.lamfac = 1.0 / p->Pr / Ec; kc.viscexp = p->viscexp;
    kc.viscmode = p->viscexp == 1.0 ? 1 : p->viscexp == 0.5 ? 2 : p->viscexp == 0.75 ? 3 : p->viscexp == 1.5 ? 4 : 0;
    kc.periodicX = p->periodicX; kc.boundaryLayer = p->boundaryLayer; kc.nonUniformX = p->nonUniformX;
    kc.perturbed = p->perturbed; kc.forcing = p->forcing; kc.quirk_q1 = p->quirk_q1;
    kc.TwallTop = p->TwallTop; kc.TwallBot = p->TwallBot;
    kc.kC = p->kC; kc.LP = p->LP; kc.amp1 = p->amp1; kc.amp2 = p->amp2; kc.omega1 = p->omega1; kc.omega2 = p->omega2;
    kc.lambdaP = p->Ly / (2.0 * M_PI);
    kc.Lx = p->Lx; kc.Ly = p->Ly; kc.Lz = p->Lz; kc.CFL = p->CFL;
    {
        std::vector<double> dxv(mx), tab((size_t)mx * (2 * v + 1));
        dxv[0] = (x[1] + x[0]) / 2.0;
        for (int i = 1; i < mx - 1; i++) dxv[i] = (x[i + 1] - x[i - 1]) / 2.0;
        dxv[mx - 1] = p->Lx - (x[mx - 1] + x[mx - 2]) / 2.0;
        const double h_d2x = h_dx * h_dx;
        for (int it = 0; it < v; it++)
            for (int i = 0; i < mx; i++)
                tab[i + (size_t)it * mx] = (cVS[it] * (xp[i] * xp[i]) * h_d2x - cVF[it] * xpp[i] * (xp[i] * xp[i] * xp[i]) * h_dx);
        for (int i = 0; i < mx; i++) tab[i + (size_t)v * mx] = cVS[v] * (xp[i] * xp[i]) * h_d2x;
        for (int it = v + 1; it < 2 * v + 1; it++)
            for (int i = 0; i < mx; i++)
                tab[i + (size_t)it * mx] = (cVS[2 * v - it] * (xp[i] * xp[i]) * h_d2x + cVF[2 * v - it] * xpp[i] * (xp[i] * xp[i] * xp[i]) * h_dx);
        CKC(upload(S->d_dxv, dxv.data(), mx));
        CKC(upload(S->d_cVSx, tab.data(), tab.size()));
        CKC(upload(S->d_xp, xp, mx));
    }
    kc.xp = S->d_xp; kc.cVSx = S->d_cVSx; kc.dxv = S->d_dxv;
    kc.spongeX = nullptr; kc.spongeZ = nullptr; kc.sref = nullptr;
    kc.dt = S->d_scal + SC_DT; kc.dpdz = S->d_scal + SC_DPDZ; kc.time_on_gpu = S->d_scal + SC_TGPU;
    double sc0[SC_N]; std::memset(sc0, 0, sizeof(sc0));
    sc0[SC_DPDZ] = p->forcing ? 0.00372 : 0.0;                 // cuda_utils.cu:68-70
    CKC(cudaMemcpy(S->d_scal, sc0, sizeof(sc0), cudaMemcpyHostToDevice));
    // opt in to the large dynamic shared memory the stage kernel needs; fail loudly if the device cannot give it
    const bool lean_maps = !(kF32 && S->duo);        // single precision: a solver is served by one kernel generation
    if (lean_maps && (size_t)lean_smem_bytes(s, kc.viscmode == 1) > prop.sharedMemPerBlockOptin) { set_error("device shared memory too small for the stage kernel"); cudns_destroy(S); return CUDNS_EUNSUPPORTED; }
    {   // TMA descriptors: halo'd tile and tile interior of every state buffer and of theta
        const int CXb = 32 + 2 * GX;
        const int lty = kc.viscmode == 1 ? CUDNS_LEAN_TY_LINEAR : lean_ty_general(s), LYb = lean_box_rows(lty, s);
        for (int b = 0; lean_maps && b < S->nstate; b++) {
            if ((rc = make_map(&S->lmaps[b].qbox, L, S->state[b], 5, CXb, LYb)) || (rc = make_map(&S->lmaps[b].qint, L, S->state[b], 5, 32, lty)) ||
                (rc = make_map(&S->lmaps[b].thbox, L, S->theta, 1, CXb, LYb)) || (rc = make_map(&S->lmaps[b].thint, L, S->theta, 1, 32, lty))) {
                cudns_destroy(S); return rc;
            }
        }
        if (lean_maps && ((rc = make_rmap(&S->rmap[0], L, S->R1, lty)) || (S->R2 && (rc = make_rmap(&S->rmap[1], L, S->R2, lty))))) { cudns_destroy(S); return rc; }
        const char *we = getenv("CUDNS_WIDE");
        S->wide = !kF32 && lean_wide_ok(kc) && !(we && std::string(we) == "0") && (size_t)lean_smem_wide_bytes(s) <= prop.sharedMemPerBlockOptin;
        if (S->fast && !kF32) {
            const int fty = S->fast_ty, FYb = fty + 2 * s;
            for (int b = 0; b < S->nstate; b++) {
                real *q = S->state[b];
                if ((rc = make_map(&S->fmaps[b].q4box, L, q, 4, CXb, FYb)) || (rc = make_map(&S->fmaps[b].a3box, L, q + 5 * L.vol, 3, CXb, FYb)) ||
                    (rc = make_map(&S->fmaps[b].q4int, L, q, 4, 32, fty)) || (rc = make_map(&S->fmaps[b].a3int, L, q + 5 * L.vol, 3, 32, fty)) ||
                    (rc = make_map(&S->fmaps[b].eint, L, q + 4 * L.vol, 1, 32, fty))) {
                    cudns_destroy(S); return rc;
                }
            }
            if ((rc = make_rmap(&S->frmap[0], L, S->R1, fty)) || (S->R2 && (rc = make_rmap(&S->frmap[1], L, S->R2, fty)))) { cudns_destroy(S); return rc; }
        }
        {
            const char *te = getenv("CUDNS_THETA_TMA");
            S->theta_tma = mx % 2 == 0 && !(te && std::string(te) == "0") &&
                           (size_t)theta_tma_smem_bytes(v) <= prop.sharedMemPerBlockOptin;
            for (int b = 0; S->theta_tma && b < S->nstate; b++) {
                real *q = S->state[b];
                if ((rc = make_map(&S->tmaps[b].u, L, q + L.vol, 1, THETA_TX + 2 * GX, THETA_TY)) ||
                    (rc = make_map(&S->tmaps[b].v, L, q + 2 * L.vol, 1, THETA_TX, THETA_TY + 2 * v)) ||
                    (rc = make_map(&S->tmaps[b].w, L, q + 3 * L.vol, 1, THETA_TX, THETA_TY))) {
                    cudns_destroy(S); return rc;
                }
            }
        }
        if (S->duo) {
            const int DXb = DUO_TX + 2 * GX, DYb = duo_box_rows(s);
            for (int b = 0; b < S->nstate; b++) {
                real *q = S->state[b];
                if ((rc = make_map(&S->dmaps[b].q4box, L, q, 4, DXb, DYb)) || (rc = make_map(&S->dmaps[b].a3box, L, q + 5 * L.vol, 3, DXb, DYb)) ||
                    (rc = make_map(&S->dmaps[b].q4int, L, q, 4, DUO_TX, DUO_TY)) || (rc = make_map(&S->dmaps[b].a3int, L, q + 5 * L.vol, 3, DUO_TX, DUO_TY))) {
                    cudns_destroy(S); return rc;
                }
            }
        }
        if (S->wide) {
            const int wty = CUDNS_LEAN_TY_WIDE, WYb = lean_box_rows(wty, s);
            for (int b = 0; b < S->nstate; b++) {
                if ((rc = make_map(&S->wmaps[b].qbox, L, S->state[b], 5, CXb, WYb)) || (rc = make_map(&S->wmaps[b].qint, L, S->state[b], 5, 32, wty)) ||
                    (rc = make_map(&S->wmaps[b].thbox, L, S->theta, 1, CXb, WYb)) || (rc = make_map(&S->wmaps[b].thint, L, S->theta, 1, 32, wty))) {
                    cudns_destroy(S); return rc;
                }
            }
            if ((rc = make_rmap(&S->wrmap[0], L, S->R1, wty)) || (S->R2 && (rc = make_rmap(&S->wrmap[1], L, S->R2, wty)))) { cudns_destroy(S); return rc; }
        }
    }
    CKC(cudaStreamSynchronize(S->st));
    *out = S;
    return CUDNS_OK;
}

static void io_shutdown(cudns_solver *S);

int cudns_destroy(cudns_handle S) {
    if (!S) return CUDNS_OK;
    cudaSetDevice(S->P.device);
    io_shutdown(S);
    if (S->st) cudaStreamSynchronize(S->st);
    if (S->ipc_lo) cudaIpcCloseMemHandle(S->ipc_lo);
    if (S->ipc_hi && S->ipc_hi != S->ipc_lo) cudaIpcCloseMemHandle(S->ipc_hi);
    cudaFree(S->block);
    cudaFree(S->theta); cudaFree(S->R1); cudaFree(S->R2);
    cudaFree(S->d_xp); cudaFree(S->d_cVSx); cudaFree(S->d_dxv); cudaFree(S->d_spx); cudaFree(S->d_spz); cudaFree(S->d_sref);
    cudaFree(S->d_scal); cudaFree(S->d_hist); cudaFree(S->d_prof); cudaFree(S->d_post); cudaFree(S->d_bulk);
    cudaFree(S->send_lo); cudaFree(S->send_hi); cudaFree(S->recv_lo); cudaFree(S->recv_hi);
    if (S->tm) { for (auto &e : S->tm->ev) cudaEventDestroy(e); delete S->tm; }
    if (S->st) cudaStreamDestroy(S->st);
    delete S;
    return CUDNS_OK;
}

int cudns_memory_report(cudns_handle S, size_t *solver_bytes, size_t *free_bytes, size_t *total_bytes) {
    if (!S) { set_error("NULL handle"); return CUDNS_EINVAL; }
    CK(cudaSetDevice(S->P.device));
    size_t f = 0, t = 0; CK(cudaMemGetInfo(&f, &t));
    if (solver_bytes) *solver_bytes = S->bytes;
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return CUDNS_OK;
}

int cudns_set_allreduce(cudns_handle S, cudns_allreduce_fn fn, void *user) { if (!S) return CUDNS_EINVAL; S->allreduce = fn; S->allreduce_user = user; return CUDNS_OK; }
int cudns_set_exchange(cudns_handle S, cudns_exchange_fn fn, void *user) { if (!S) return CUDNS_EINVAL; S->exchange = fn; S->exchange_user = user; return CUDNS_OK; }
int cudns_get_stream(cudns_handle S, void **stream) { if (!S || !stream) return CUDNS_EINVAL; *stream = (void *)S->st; return CUDNS_OK; }
int cudns_halo_buffers(cudns_handle S, void **send_lo, void **send_hi, void **recv_lo, void **recv_hi, size_t *bytes_each) {
    if (!S) return CUDNS_EINVAL;
    if (send_lo) *send_lo = S->send_lo; if (send_hi) *send_hi = S->send_hi;
    if (recv_lo) *recv_lo = S->recv_lo; if (recv_hi) *recv_hi = S->recv_hi;
    if (bytes_each) *bytes_each = S->halo_doubles * sizeof(real);
    return CUDNS_OK;
}
int cudns_halo_local_info(cudns_handle S, cudns_peer_info *mine) {
    if (!S || !mine) { set_error("NULL argument"); return CUDNS_EINVAL; }
    CK(cudaSetDevice(S->P.device));
    std::memset(mine, 0, sizeof(*mine));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, S->block));
    static_assert(sizeof(h) == CUDNS_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
    std::memcpy(mine->mem_handle, &h, sizeof(h));
    mine->device = S->P.device;
    mine->pid = (int)getpid();
    mine->local_ptr = (uint64_t)(uintptr_t)S->block;
    mine->block_bytes = (uint64_t)(S->block_doubles + 32) * sizeof(real);
    return CUDNS_OK;
}

// map one neighbour's block: same process -> its pointer (peer access enabled), other process -> CUDA IPC
static int open_peer(cudns_solver *S, const cudns_peer_info *pi, real **ptr, void **ipc) {
    *ptr = nullptr; *ipc = nullptr;
    if (pi->block_bytes != (uint64_t)(S->block_doubles + 32) * sizeof(real)) { set_error("neighbour block size differs (unequal slabs?)"); return CUDNS_EINVAL; }
    if (pi->pid == (int)getpid()) {
        if (pi->device != S->P.device) {
            int can = 0; CK(cudaDeviceCanAccessPeer(&can, S->P.device, pi->device));
            if (!can) { set_error("no peer access between the neighbour devices"); return CUDNS_EUNSUPPORTED; }
            cudaError_t e = cudaDeviceEnablePeerAccess(pi->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { set_error(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)); return CUDNS_ECUDA; }
            cudaGetLastError();
        }
        *ptr = (real *)(uintptr_t)pi->local_ptr;
        return CUDNS_OK;
    }
    cudaIpcMemHandle_t h; std::memcpy(&h, pi->mem_handle, sizeof(h));
    void *p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { set_error(std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e)); return CUDNS_ECUDA; }
    *ptr = (real *)p; *ipc = p;
    return CUDNS_OK;
}

int cudns_halo_connect(cudns_handle S, const cudns_peer_info *lower, const cudns_peer_info *upper) {
    if (!S || !lower || !upper) { set_error("NULL argument"); return CUDNS_EINVAL; }
    if (S->P.nranks < 2) { set_error("cudns_halo_connect needs nranks > 1"); return CUDNS_EINVAL; }
    CK(cudaSetDevice(S->P.device));
    int rc;
    if ((rc = open_peer(S, lower, &S->peer_lo, &S->ipc_lo))) return rc;
    const bool same = lower->pid == upper->pid && lower->device == upper->device && lower->local_ptr == upper->local_ptr &&
                      std::memcmp(lower->mem_handle, upper->mem_handle, CUDNS_IPC_HANDLE_BYTES) == 0;
    if (same) { S->peer_hi = S->peer_lo; S->ipc_hi = S->ipc_lo; }
    else if ((rc = open_peer(S, upper, &S->peer_hi, &S->ipc_hi))) return rc;
    S->connected = true;
    return CUDNS_OK;
}
int cudns_get_counters(cudns_handle S, uint64_t *kernel_launches, uint64_t *rk_stages) {
    if (!S) return CUDNS_EINVAL;
    if (kernel_launches) *kernel_launches = S->launches;
    if (rk_stages) *rk_stages = S->stages;
    return CUDNS_OK;
}

}  // extern "C"

// z ghosts of a padded 5-field buffer: periodic wrap on one device, exchange with the slab neighbours otherwise
static int fill_z_ghosts(cudns_solver *S, real *q) {
    if (S->P.nranks == 1) {
        if (!S->P.boundaryLayer) { launch_zwrap(S->kc, q, 5, S->st); S->launches++; }
        return CUDNS_OK;
    }
    if (!S->exchange) { set_error("nranks > 1 needs a halo transport: call cudns_set_exchange (or cudns_halo_connect)"); return CUDNS_ESTATE; }
    launch_pack_z(S->kc, q, S->send_lo, S->send_hi, S->st);
    S->exchange(S->exchange_user, (void *)S->st);
    launch_unpack_z(S->kc, q, S->recv_lo, S->recv_hi, S->st);
    S->launches += 2;
    return CUDNS_OK;
}

static void reduce_across(cudns_solver *S, double *dptr, int n, int op) {
    if (S->P.nranks > 1 && S->allreduce) S->allreduce(S->allreduce_user, dptr, n, op);
}
// every entry point that reduces scalars across the slabs (dt MAX, bulk / forcing / profile SUMs) refuses to run without the
// collective: slabs advancing with rank-local dt or dpdz would produce a wrong answer silently
static int need_allreduce(cudns_solver *S) {
    if (S->P.nranks > 1 && !S->allreduce) { set_error("nranks > 1 needs the scalar all-reduce: call cudns_set_allreduce first"); return CUDNS_ESTATE; }
    return CUDNS_OK;
}
// did a halo hand-shake of this solver give up waiting for a neighbour?  (call after the stream has been synchronised)
static int check_halo_error(cudns_solver *S) {
    if (S->P.nranks == 1) return CUDNS_OK;
    unsigned long long e = 0;
    if (cudaMemcpy(&e, S->d_scal + SC_HALOERR, sizeof(e), cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("reading the hand-shake status failed"); return CUDNS_ECUDA; }
    if (e) { set_error("halo hand-shake timed out at stage " + std::to_string(e) + ": a slab neighbour did not signal (different stage count, or it died)"); return CUDNS_ESTATE; }
    return CUDNS_OK;
}

extern "C" {

int cudns_set_state_device(cudns_handle S, const double *d_r, const double *d_u, const double *d_v, const double *d_w, const double *d_e) {
    if (!S || !d_r || !d_u || !d_v || !d_w || !d_e) { set_error("NULL argument"); return CUDNS_EINVAL; }
    { int rc0 = need_allreduce(S); if (rc0) return rc0; }
    CK(cudaSetDevice(S->P.device));
    const double *src[5] = {d_r, d_u, d_v, d_w, d_e};
    S->cur = 0;
    for (int b = 0; b < 3; b++) S->aux_valid[b] = false;
    launch_pad(S->kc, src, S->state[0], S->st);
    launch_fill_xy(S->kc, S->state[0], 5, S->st);
    S->launches += 2;
    int rc = fill_z_ghosts(S, S->state[0]); if (rc) return rc;
    // time_on_GPU = 0 (initDevice cuda_utils.cu:385) ; the "first calcState" (:332) only matters through the mu that
    // the first dt refresh sees: keep its viscous limiter
    CK(cudaMemsetAsync(S->d_scal + SC_TGPU, 0, sizeof(double), S->st));
    launch_dt_reduce(S->kc, S->state[0], S->d_scal + SC_RED0, S->st); S->launches++;
    CK(cudaMemcpyAsync(S->d_scal + SC_STALE0, S->d_scal + SC_RED0, 2 * sizeof(double), cudaMemcpyDeviceToDevice, S->st));
    reduce_across(S, S->d_scal + SC_STALE0, 2, 2);
    CK(cudaStreamSynchronize(S->st));
    S->have_state = true;
    return CUDNS_OK;
}

int cudns_get_state_device(cudns_handle S, double *d_r, double *d_u, double *d_v, double *d_w, double *d_e) {
    if (!S || !d_r || !d_u || !d_v || !d_w || !d_e) { set_error("NULL argument"); return CUDNS_EINVAL; }
    if (!S->have_state) { set_error("no state set"); return CUDNS_ESTATE; }
    CK(cudaSetDevice(S->P.device));
    double *dst[5] = {d_r, d_u, d_v, d_w, d_e};
    launch_unpad(S->kc, S->state[S->cur], dst, S->st); S->launches++;
    CK(cudaStreamSynchronize(S->st));
    return CUDNS_OK;
}

// copyField(0), cuda_utils.cu:317-333
int cudns_set_state(cudns_handle S, const double *r, const double *u, const double *v, const double *w, const double *e) {
    if (!S || !r || !u || !v || !w || !e) { set_error("NULL argument"); return CUDNS_EINVAL; }
    CK(cudaSetDevice(S->P.device));
    // stage through the register array R1 (5*N doubles), exactly the staging role of d_fr.. in the reference (single precision: R1 is
    // half that size, the double staging block is allocated for the call)
    const double *src[5] = {r, u, v, w, e};
    double *stage = nullptr;
    if (sizeof(real) == sizeof(double)) stage = reinterpret_cast<double *>(S->R1);
    else CK(cudaMalloc((void **)&stage, 5 * S->N * sizeof(double)));
    for (int f = 0; f < 5; f++) {
        cudaError_t e2 = cudaMemcpyAsync(stage + f * S->N, src[f], S->N * sizeof(double), cudaMemcpyHostToDevice, S->st);
        if (e2 != cudaSuccess) { if ((void *)stage != (void *)S->R1) cudaFree(stage); set_error(std::string("cudns_set_state: ") + cudaGetErrorString(e2)); return CUDNS_ECUDA; }
    }
    int rc = cudns_set_state_device(S, stage, stage + S->N, stage + 2 * S->N, stage + 3 * S->N, stage + 4 * S->N);
    if ((void *)stage != (void *)S->R1) cudaFree(stage);
    return rc;
}

// copyField(1), cuda_utils.cu:334-355.  Between steps the register array holds no live data for the
// low-storage scheme's first stage (alpha_0 = 0) nor for Kutta/RK4, so it doubles as the staging buffer.
int cudns_get_state(cudns_handle S, double *r, double *u, double *v, double *w, double *e) {
    if (!S || !r || !u || !v || !w || !e) { set_error("NULL argument"); return CUDNS_EINVAL; }
    CK(cudaSetDevice(S->P.device));
    double *stage = nullptr;
    if (sizeof(real) == sizeof(double)) stage = reinterpret_cast<double *>(S->R1);
    else CK(cudaMalloc((void **)&stage, 5 * S->N * sizeof(double)));
    int rc = cudns_get_state_device(S, stage, stage + S->N, stage + 2 * S->N, stage + 3 * S->N, stage + 4 * S->N);
    double *dst[5] = {r, u, v, w, e};
    cudaError_t e2 = cudaSuccess;
    for (int f = 0; f < 5 && !rc && e2 == cudaSuccess; f++) e2 = cudaMemcpyAsync(dst[f], stage + f * S->N, S->N * sizeof(double), cudaMemcpyDeviceToHost, S->st);
    if (!rc && e2 == cudaSuccess) e2 = cudaStreamSynchronize(S->st);
    if ((void *)stage != (void *)S->R1) cudaFree(stage);
    if (rc) return rc;
    if (e2 != cudaSuccess) { set_error(std::string("cudns_get_state: ") + cudaGetErrorString(e2)); return CUDNS_ECUDA; }
    return CUDNS_OK;
}

int cudns_set_sponge(cudns_handle S, const double *sigma_x, const double *sigma_z, const double *ref5) {
    if (!S || !sigma_x || !sigma_z || !ref5) { set_error("NULL argument"); return CUDNS_EINVAL; }
    CK(cudaSetDevice(S->P.device));
    CK(upload(S->d_spx, sigma_x, S->L.mx));
    CK(upload(S->d_spz, sigma_z, S->L.mz));
    CK(upload(S->d_sref, ref5, 5 * (size_t)S->L.mx * S->L.mz));
    S->kc.spongeX = S->d_spx; S->kc.spongeZ = S->d_spz; S->kc.sref = S->d_sref;
    S->have_sponge = true;
    return CUDNS_OK;
}

int cudns_set_dt(cudns_handle S, double dt, int fixed) {
    if (!S) return CUDNS_EINVAL;
    CK(cudaSetDevice(S->P.device));
    CK(cudaMemcpy(S->d_scal + SC_DT, &dt, sizeof(double), cudaMemcpyHostToDevice));
    S->fixed_dt = fixed != 0;
    return CUDNS_OK;
}

int cudns_get_scalars(cudns_handle S, double *dt, double *dpdz, double *time) {
    if (!S) return CUDNS_EINVAL;
    CK(cudaSetDevice(S->P.device));
    double h[4];
    CK(cudaStreamSynchronize(S->st));
    CK(cudaMemcpy(h, S->d_scal, 4 * sizeof(double), cudaMemcpyDeviceToHost));
    if (dt) *dt = h[SC_DT]; if (dpdz) *dpdz = h[SC_DPDZ]; if (time) *time = h[SC_TIME];
    return CUDNS_OK;
}

}  // extern "C"

// the stage kernel of the selected generation; p.qin = state[in], p.qbase = state[base]
static void launch_stage_any(cudns_solver *S, const StagePtrs &p, const StageCoef &c, int in, int base) {
    if (S->duo && !(p.RB && p.RW && c.wOld != 0.0)) {        // (RB and an accumulated RW share one stash slot: never both in our schemes)
        // 8-field buffers: theta of the input state is its field 7 (run_stage put it there), H and T are rebuilt first when the
        // buffer was not written by a kernel that stores them
        if (!S->aux_valid[in]) { launch_derive_aux(S->kc, S->state[in], S->st); S->launches++; S->aux_valid[in] = true; }
        launch_rhs_stage_duo(S->kc, p, c, S->dmaps[in], S->st);
        return;
    }
    {
        // the wide variant stages one operand tile only (RA): every low-storage RK3 stage and the test path qualify
        const bool wide = S->wide && !p.RB && !(p.RW && c.wOld != 0.0) && p.qbase == p.qin;
        if (wide && S->fast) {
            // 8-field buffers: theta of the input state is its field 7 (run_stage put it there), H and T are rebuilt first when the
            // buffer was not written by this kernel
            if (!S->aux_valid[in]) { launch_derive_aux(S->kc, S->state[in], S->st); S->launches++; S->aux_valid[in] = true; }
            FastMaps m = S->fmaps[in];
            m.opa = (p.RA == S->R2) ? S->frmap[1] : S->frmap[0];
            launch_rhs_stage_fast(S->kc, p, c, m, S->fast_ty, S->st);
            return;
        }
        const StageMaps *sm = wide ? S->wmaps : S->lmaps;
        const CUtensorMap *rm = wide ? S->wrmap : S->rmap;
        LeanMaps m;
        m.qbox = sm[in].qbox; m.qint = sm[in].qint; m.thbox = sm[in].thbox; m.thint = sm[in].thint;
        m.qbint = sm[base].qint;
        auto rmap_of = [&](const real *r) -> const CUtensorMap & { return (r == S->R2) ? rm[1] : rm[0]; };
        m.opa = rmap_of(p.RA);
        m.opb = rmap_of(p.RB ? p.RB : p.RW);
        launch_rhs_stage_lean(S->kc, p, c, m, wide, S->st);
    }
}

// will launch_stage_any serve this stage with the fast kernel?
static bool stage_is_fast(const cudns_solver *S, const StagePtrs &p, const StageCoef &c) {
    return (S->duo && !(p.RB && p.RW && c.wOld != 0.0)) || (S->fast && S->wide && !p.RB && !(p.RW && c.wOld != 0.0) && p.qbase == p.qin);
}

// does this solver's stage kernel write the z ghost planes itself (lean kernel; on one device always, across devices once
// the peer blocks are mapped)?
static bool inkernel_ghosts(const cudns_solver *S) { return S->P.nranks == 1 || S->connected; }

// neighbour-side addresses of output buffer `out` for the stage kernel (see StagePtrs::qout_lo / qout_hi)
static void ghost_targets(cudns_solver *S, int out, StagePtrs &p) {
    p.qout_lo = nullptr; p.qout_hi = nullptr;
    if (!inkernel_ghosts(S)) return;
    const size_t off = (size_t)out * S->nfb * S->L.vol;
    const bool bl = S->P.boundaryLayer != 0;                  // z is not periodic: the global bottom / top have no neighbour
    if (S->P.nranks == 1) { if (!bl) { p.qout_lo = S->state[out]; p.qout_hi = S->state[out]; } return; }
    if (!(bl && S->P.rank == 0)) p.qout_lo = S->peer_lo + off;
    if (!(bl && S->P.rank == S->P.nranks - 1)) p.qout_hi = S->peer_hi + off;
}

// after a stage kernel that wrote the neighbours' ghost planes: tell them, and wait for theirs
static void handshake(cudns_solver *S) {
    if (S->P.nranks == 1) return;
    const bool bl = S->P.boundaryLayer != 0;
    const bool has_lo = !(bl && S->P.rank == 0), has_hi = !(bl && S->P.rank == S->P.nranks - 1);
    S->epoch++;
    unsigned long long *mine = (unsigned long long *)(S->block + S->block_doubles);
    // my planes go into the lower neighbour's UPPER ghosts: its slot 1 ("written by the upper one"), and vice versa
    unsigned long long *lo_slot = has_lo ? (unsigned long long *)(S->peer_lo + S->block_doubles) + 1 : nullptr;
    unsigned long long *hi_slot = has_hi ? (unsigned long long *)(S->peer_hi + S->block_doubles) + 0 : nullptr;
    launch_halo_signal(lo_slot, hi_slot, S->epoch, S->st);
    launch_halo_wait(mine, has_lo, has_hi, S->epoch, S->halo_timeout_ns, (unsigned long long *)(S->d_scal + SC_HALOERR), S->st);
    S->launches += 2;
}

// one RHS evaluation + register update: K = RHS(state[in]); see StageCoef
static int run_stage(cudns_solver *S, int in, int base, int out, const real *RA, const real *RB, real *RW,
                     const StageCoef &c, real *rhs_out) {
    StagePtrs p;
    p.qin = S->state[in]; p.qbase = S->state[base]; p.qout = S->state[out]; p.theta = S->theta;
    p.RA = RA; p.RB = RB; p.RW = RW; p.rhs_out = rhs_out;
    const bool fast = stage_is_fast(S, p, c);
    if (fast) p.theta = S->state[in] + 7 * S->L.vol;            // 8-field buffers: theta travels with the state it belongs to
    StageTimer *tm = (S->tm && S->tm->on && !rhs_out && S->tm->used + 4 <= S->tm->ev.size()) ? S->tm : nullptr;
    cudaEvent_t *tev = tm ? &tm->ev[tm->used] : nullptr;
    if (tm) { tm->used += 4; cudaEventRecord(tev[0], S->st); }
    if (S->theta_tma) launch_theta_tma(S->kc, S->state[in], const_cast<real *>(p.theta), S->tmaps[in], S->st);
    else launch_theta(S->kc, S->state[in], const_cast<real *>(p.theta), S->st);
    if (tm) cudaEventRecord(tev[1], S->st);
    ghost_targets(S, out, p);
    if (rhs_out) { p.qout_lo = nullptr; p.qout_hi = nullptr; }
    launch_stage_any(S, p, c, in, base);
    if (tm) cudaEventRecord(tev[2], S->st);
    S->launches += 2;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error(std::string("stage launch: ") + cudaGetErrorString(e)); return CUDNS_ECUDA; }
    if (!rhs_out) {
        // H, T of the new state: written by the fast kernel together with the state (ghost planes included when it stores them
        // itself); stale after any other kernel or after a separate ghost exchange
        S->aux_valid[out] = fast && inkernel_ghosts(S);
        if (inkernel_ghosts(S)) handshake(S);
        else { int rc = fill_z_ghosts(S, S->state[out]); if (rc) return rc; }
        if (tm) cudaEventRecord(tev[3], S->st);
        S->stages++;
    }
    return CUDNS_OK;
}

// after the stream has been synchronised: fold the recorded events into the running sums
static void stage_timer_collect(cudns_solver *S) {
    StageTimer *tm = S->tm;
    if (!tm || !tm->on) return;
    for (size_t i = 0; i + 4 <= tm->used; i += 4) {
        float a = 0, b = 0, c = 0;
        if (cudaEventElapsedTime(&a, tm->ev[i], tm->ev[i + 1]) != cudaSuccess || cudaEventElapsedTime(&b, tm->ev[i + 1], tm->ev[i + 2]) != cudaSuccess ||
            cudaEventElapsedTime(&c, tm->ev[i + 2], tm->ev[i + 3]) != cudaSuccess) { cudaGetLastError(); continue; }
        tm->theta_ms += a; tm->stage_ms += b; tm->halo_ms += c; tm->n++;
    }
    tm->used = 0;
}

// viscous limiter of the state the LAST calcState saw (the reference refreshes dt with the mu field left by the
// final stage's calcState, i.e. of the state before that stage's update: cuda_main.cu:171,249-251, calc_stress.cu:131)
static void record_stale_visc(cudns_solver *S, int in) {
    launch_dt_reduce(S->kc, S->state[in], S->d_scal + SC_RED0, S->st); S->launches++;
    cudaMemcpyAsync(S->d_scal + SC_STALE1, S->d_scal + SC_RED1, sizeof(double), cudaMemcpyDeviceToDevice, S->st);
    reduce_across(S, S->d_scal + SC_STALE1, 1, 2);
}

extern "C" {

int cudns_calc_rhs(cudns_handle S, double *rhs_r, double *rhs_u, double *rhs_v, double *rhs_w, double *rhs_e) {
    if (!S) { set_error("NULL handle"); return CUDNS_EINVAL; }
    if (!S->have_state) { set_error("no state set"); return CUDNS_ESTATE; }
    CK(cudaSetDevice(S->P.device));
    real *tmp = nullptr;
    CK(cudaMalloc((void **)&tmp, 5 * S->N * sizeof(real)));
    StageCoef c = {1, 0, 0, 0, 0};
    int rc = run_stage(S, S->cur, S->cur, S->cur, nullptr, nullptr, nullptr, c, tmp);
    if (rc) { cudaFree(tmp); return rc; }
    double *dst[5] = {rhs_r, rhs_u, rhs_v, rhs_w, rhs_e};
    std::vector<real> hbuf(sizeof(real) == sizeof(double) ? 0 : S->N);          // single precision: widened on the host
    for (int f = 0; f < 5; f++) {
        if (!dst[f]) continue;
        if constexpr (!kF32) {
            CK(cudaMemcpyAsync(dst[f], tmp + f * S->N, S->N * sizeof(real), cudaMemcpyDeviceToHost, S->st));
        } else {
            CK(cudaMemcpyAsync(hbuf.data(), tmp + f * S->N, S->N * sizeof(real), cudaMemcpyDeviceToHost, S->st));
            CK(cudaStreamSynchronize(S->st));
            for (size_t n = 0; n < S->N; n++) dst[f][n] = (double)hbuf[n];
        }
    }
    CK(cudaStreamSynchronize(S->st));
    cudaFree(tmp);
    return CUDNS_OK;
}

int cudns_calc_dt(cudns_handle S, double *dt) {
    if (!S || !dt) { set_error("NULL argument"); return CUDNS_EINVAL; }
    if (!S->have_state) { set_error("no state set"); return CUDNS_ESTATE; }
    { int rc0 = need_allreduce(S); if (rc0) return rc0; }
    CK(cudaSetDevice(S->P.device));
    launch_dt_reduce(S->kc, S->state[S->cur], S->d_scal + SC_RED0, S->st); S->launches++;
    reduce_across(S, S->d_scal + SC_RED0, 2, 2);
    double h[2];
    CK(cudaStreamSynchronize(S->st));
    CK(cudaMemcpy(h, S->d_scal + SC_RED0, 2 * sizeof(double), cudaMemcpyDeviceToHost));
    *dt = S->P.CFL / fmax(h[0], h[1]);
    return CUDNS_OK;
}

int cudns_calc_bulk(cudns_handle S, double *par1, double *par2) {
    if (!S) { set_error("NULL handle"); return CUDNS_EINVAL; }
    if (!S->have_state) { set_error("no state set"); return CUDNS_ESTATE; }
    { int rc0 = need_allreduce(S); if (rc0) return rc0; }
    CK(cudaSetDevice(S->P.device));
    launch_bulk_reduce(S->kc, S->state[S->cur], S->d_scal + SC_BULK0, S->d_bulk, S->st); S->launches++;
    reduce_across(S, S->d_scal + SC_BULK0, 4, 1);
    double h[4];
    CK(cudaStreamSynchronize(S->st));
    CK(cudaMemcpy(h, S->d_scal + SC_BULK0, 4 * sizeof(double), cudaMemcpyDeviceToHost));
    if (S->P.forcing) { if (par1) *par1 = h[2] / h[1]; if (par2) *par2 = h[3]; }
    else { if (par1) *par1 = h[0]; }
    return CUDNS_OK;
}

// mean square vorticity of a periodic box (see enstrophy_reduce_kernel): the dissipation history of the unforced runs
static int enstrophy_supported(cudns_solver *S) {
    if (!S->P.periodicX || S->P.boundaryLayer) { set_error("enstrophy: periodic boxes only (periodicX = 1, boundaryLayer = 0)"); return CUDNS_EUNSUPPORTED; }
    return CUDNS_OK;
}
int cudns_calc_enstrophy(cudns_handle S, double *ens) {
    if (!S || !ens) { set_error("NULL argument"); return CUDNS_EINVAL; }
    if (!S->have_state) { set_error("no state set"); return CUDNS_ESTATE; }
    { int rc0 = need_allreduce(S); if (rc0) return rc0; }
    { int rc0 = enstrophy_supported(S); if (rc0) return rc0; }
    CK(cudaSetDevice(S->P.device));
    launch_enstrophy_reduce(S->kc, S->state[S->cur], S->d_scal + SC_ENS, S->d_bulk, S->st); S->launches++;
    reduce_across(S, S->d_scal + SC_ENS, 1, 1);
    CK(cudaStreamSynchronize(S->st));
    CK(cudaMemcpy(ens, S->d_scal + SC_ENS, sizeof(double), cudaMemcpyDeviceToHost));
    return CUDNS_OK;
}

// runSimulationLowStorage / runSimulation, cuda_main.cu:44-186
int cudns_advance(cudns_handle S, int nsteps, double *time, double *par1, double *par2) {
    if (!S) { set_error("NULL handle"); return CUDNS_EINVAL; }
    if (!S->have_state) { set_error("cudns_advance before cudns_set_state"); return CUDNS_ESTATE; }
    if (nsteps < 0) { set_error("nsteps < 0"); return CUDNS_EINVAL; }
    if (S->P.boundaryLayer && !S->have_sponge) { set_error("boundaryLayer needs cudns_set_sponge first"); return CUDNS_ESTATE; }
    { int rc0 = need_allreduce(S); if (rc0) return rc0; }
    CK(cudaSetDevice(S->P.device));
    if (nsteps == 0) return CUDNS_OK;
    if (S->hist_cap < nsteps) {                       // grows in steps of 4096 entries: no cudaFree/cudaMalloc between calls of similar length
        const int cap = (nsteps + 4095) / 4096 * 4096;
        cudaFree(S->d_hist); S->d_hist = nullptr;
        CK(cudaMalloc((void **)&S->d_hist, 3 * (size_t)cap * sizeof(double)));
        S->hist_cap = cap;
    }
    for (int a = 0; a < 3; a++) CK(cudaMemsetAsync(S->d_hist + (size_t)a * S->hist_cap, 0xff, (size_t)nsteps * sizeof(double), S->st));   // NaN = "not written"
    double *h_time = S->d_hist, *h_p1 = S->d_hist + S->hist_cap, *h_p2 = S->d_hist + 2 * S->hist_cap;
    double *sc = S->d_scal;
    const cudns_params &P = S->P;
    const bool ls = P.lowStorage && !P.rk4;
    const bool ens_hist = P.par2_enstrophy && !P.forcing;
    if (ens_hist) { int rc0 = enstrophy_supported(S); if (rc0) return rc0; }
    for (int istep = 0; istep < nsteps; istep++) {
        // ---- calcTimeStepPressGrad, cuda_main.cu:249-265
        if (istep % P.checkCFLcondition == 0) {
            if (!S->fixed_dt) {
                launch_dt_reduce(S->kc, S->state[S->cur], sc + SC_RED0, S->st); S->launches++;
                reduce_across(S, sc + SC_RED0, 1, 2);
                launch_dt_combine(sc + SC_DT, sc + SC_RED0, sc + SC_STALE1, P.CFL, S->st); S->launches++;
            }
            if (P.forcing) {
                launch_bulk_reduce(S->kc, S->state[S->cur], sc + SC_BULK0, S->d_bulk, S->st);
                reduce_across(S, sc + SC_BULK0, 4, 1);
                launch_scalar_ops(2, sc + SC_DPDZ, sc + SC_BULK0, nullptr, S->st);
                S->launches += 2;
            }
        }
        launch_scalar_ops(1, sc + SC_TIME, sc + SC_DT, nullptr, S->st);      // deviceSumOne, cuda_main.cu:117-118
        launch_scalar_ops(1, sc + SC_TGPU, sc + SC_DT, nullptr, S->st);      // deviceAdvanceTime, calc_stress.cu:12
        S->launches += 2;
        CK(cudaMemcpyAsync(h_time + istep, sc + SC_TIME, sizeof(double), cudaMemcpyDeviceToDevice, S->st));
        if (istep % P.checkBulk == 0) {                                       // calcBulk, calc_stress.cu:162-201
            launch_bulk_reduce(S->kc, S->state[S->cur], sc + SC_BULK0, S->d_bulk, S->st); S->launches++;
            reduce_across(S, sc + SC_BULK0, 4, 1);
            if (P.forcing) {
                launch_scalar_ops(3, h_p1 + istep, sc + SC_BULK0, nullptr, S->st); S->launches++;
                CK(cudaMemcpyAsync(h_p2 + istep, sc + SC_BULK3, sizeof(double), cudaMemcpyDeviceToDevice, S->st));
            } else {
                CK(cudaMemcpyAsync(h_p1 + istep, sc + SC_BULK0, sizeof(double), cudaMemcpyDeviceToDevice, S->st));
                if (ens_hist) {                                                 // extension: par2 = <w.w> of the unforced periodic box
                    launch_enstrophy_reduce(S->kc, S->state[S->cur], sc + SC_ENS, S->d_bulk, S->st); S->launches++;
                    reduce_across(S, sc + SC_ENS, 1, 1);
                    CK(cudaMemcpyAsync(h_p2 + istep, sc + SC_ENS, sizeof(double), cudaMemcpyDeviceToDevice, S->st));
                }
            }
        }
        const bool need_stale = !S->fixed_dt && (((istep + 1) % P.checkCFLcondition == 0) || istep == nsteps - 1);
        int rc = 0;
        if (ls) {
            // Wray low-storage RK3, cuda_main.cu:10-11,126-184: q += dt (alpha_s k_{s-1} + beta_s k_s); one register set
            const int a = S->cur, b = 1 - S->cur;
            StageCoef c0 = {8. / 15., 0, 0, 0, 1}, c1 = {5. / 12., -17. / 60., 0, 0, 1}, c2 = {3. / 4., -5. / 12., 0, 0, 1};
            rc = run_stage(S, a, a, b, nullptr, nullptr, S->R1, c0, nullptr); if (rc) return rc;
            rc = run_stage(S, b, b, a, S->R1, nullptr, S->R1, c1, nullptr); if (rc) return rc;
            if (need_stale) record_stale_visc(S, a);
            rc = run_stage(S, a, a, b, S->R1, nullptr, nullptr, c2, nullptr); if (rc) return rc;
            S->cur = b;
        } else if (!P.rk4) {
            // Kutta RK3, cuda_main.cu:57-105: q1 = q0 + dt/2 k1; q2 = q0 + dt(2k2 - k1); q = q0 + dt(k1 + 4k2 + k3)/6
            const int q0 = S->cur, qa = (S->cur + 1) % 3, qb = (S->cur + 2) % 3;
            StageCoef c0 = {0.5, 0, 0, 0, 1}, c1 = {2.0, -1.0, 0, 0, 1}, c2 = {1. / 6., 1. / 6., 4. / 6., 0, 0};
            rc = run_stage(S, q0, q0, qa, nullptr, nullptr, S->R1, c0, nullptr); if (rc) return rc;
            rc = run_stage(S, qa, q0, qb, S->R1, nullptr, S->R2, c1, nullptr); if (rc) return rc;
            if (need_stale) record_stale_visc(S, qb);
            rc = run_stage(S, qb, q0, qa, S->R1, S->R2, nullptr, c2, nullptr); if (rc) return rc;
            S->cur = qa;
        } else {
            // classical RK4 (extension): R1 accumulates k1 + 2k2 + 2k3
            const int q0 = S->cur, qa = (S->cur + 1) % 3, qb = (S->cur + 2) % 3;
            StageCoef c0 = {0.5, 0, 0, 0, 1}, c1 = {0.5, 0, 0, 1, 2}, c2 = {1.0, 0, 0, 1, 2}, c3 = {1. / 6., 1. / 6., 0, 0, 0};
            rc = run_stage(S, q0, q0, qa, nullptr, nullptr, S->R1, c0, nullptr); if (rc) return rc;
            rc = run_stage(S, qa, q0, qb, nullptr, nullptr, S->R1, c1, nullptr); if (rc) return rc;
            rc = run_stage(S, qb, q0, qa, nullptr, nullptr, S->R1, c2, nullptr); if (rc) return rc;
            if (need_stale) record_stale_visc(S, qa);
            rc = run_stage(S, qa, q0, qb, S->R1, nullptr, nullptr, c3, nullptr); if (rc) return rc;
            S->cur = qb;
        }
    }
    if (time) CK(cudaMemcpyAsync(time, h_time, nsteps * sizeof(double), cudaMemcpyDeviceToHost, S->st));
    CK(cudaStreamSynchronize(S->st));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error(std::string("advance: ") + cudaGetErrorString(e)); return CUDNS_ECUDA; }
    { int rc0 = check_halo_error(S); if (rc0) return rc0; }
    stage_timer_collect(S);
    if (par1 || par2) {
        std::vector<double> b1(nsteps), b2(nsteps);
        CK(cudaMemcpy(b1.data(), h_p1, nsteps * sizeof(double), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(b2.data(), h_p2, nsteps * sizeof(double), cudaMemcpyDeviceToHost));
        for (int i = 0; i < nsteps; i++) {
            if (par1 && i % P.checkBulk == 0) par1[i] = b1[i];
            if (par2 && i % P.checkBulk == 0 && (P.forcing || ens_hist)) par2[i] = b2[i];
        }
    }
    return CUDNS_OK;
}

// switch the per-stage event timing of cudns_advance on (resetting the sums) or off; read the sums of the calls made since
int cudns_set_stage_timing(cudns_handle S, int on) {
    if (!S) { set_error("NULL handle"); return CUDNS_EINVAL; }
    CK(cudaSetDevice(S->P.device));
    if (!S->tm) S->tm = new StageTimer();
    StageTimer *tm = S->tm;
    if (on && tm->ev.empty()) {
        tm->ev.resize(4 * 1024);
        for (auto &e : tm->ev) CK(cudaEventCreate(&e));
    }
    tm->on = on != 0; tm->used = 0;
    if (on) { tm->theta_ms = tm->stage_ms = tm->halo_ms = 0; tm->n = 0; }
    return CUDNS_OK;
}
int cudns_get_stage_timing(cudns_handle S, double *theta_ms, double *stage_ms, double *halo_ms, uint64_t *nstages) {
    if (!S) { set_error("NULL handle"); return CUDNS_EINVAL; }
    const StageTimer *tm = S->tm;
    if (theta_ms) *theta_ms = tm ? tm->theta_ms : 0; if (stage_ms) *stage_ms = tm ? tm->stage_ms : 0;
    if (halo_ms) *halo_ms = tm ? tm->halo_ms : 0; if (nstages) *nstages = tm ? tm->n : 0;
    return CUDNS_OK;
}

int cudns_profile_stage(cudns_handle S, int reps, float *ms_theta, float *ms_rhs, float *ms_halo) {
    if (!S) { set_error("NULL handle"); return CUDNS_EINVAL; }
    if (!S->have_state) { set_error("no state set"); return CUDNS_ESTATE; }
    CK(cudaSetDevice(S->P.device));
    if (reps < 1) reps = 1;
    cudaEvent_t e0, e1, e2, e3;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2)); CK(cudaEventCreate(&e3));
    float t_th = 0, t_rhs = 0, t_h = 0;
    const int a = S->cur, b = (S->cur + 1) % S->nstate;
    // the most frequent stage shape of the solver's scheme, with coefficients that leave state and registers unchanged (reps do not
    // drift): low-storage RK3 stages 2-3 (RA read, RW written); Kutta RK3 stage 2 (base != input, RA read, RW written); RK4
    // stages 2-3 (base != input, RW accumulated)
    const bool ls = S->P.lowStorage && !S->P.rk4;
    StageCoef c = {0.0, 0, 0, 0, 1};
    if (S->P.rk4) { c.wOld = 1.0; c.wNew = 0.0; }
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0, S->st));
        StagePtrs p; p.qin = S->state[a]; p.qbase = ls ? S->state[a] : S->state[(a + 2) % 3]; p.qout = S->state[b]; p.theta = S->theta;
        p.RA = S->P.rk4 ? nullptr : S->R1; p.RB = nullptr; p.RW = (ls || S->P.rk4) ? S->R1 : S->R2; p.rhs_out = nullptr;
        const bool fast = stage_is_fast(S, p, c);
        if (fast) {
            p.theta = S->state[a] + 7 * S->L.vol;
            if (!S->aux_valid[a]) { launch_derive_aux(S->kc, S->state[a], S->st); S->aux_valid[a] = true; }
        }
        if (S->theta_tma) launch_theta_tma(S->kc, S->state[a], const_cast<real *>(p.theta), S->tmaps[a], S->st);
        else launch_theta(S->kc, S->state[a], const_cast<real *>(p.theta), S->st);
        CK(cudaEventRecord(e1, S->st));
        ghost_targets(S, b, p);
        launch_stage_any(S, p, c, a, ls ? a : (a + 2) % 3);
        S->aux_valid[b] = false;                          // a scratch copy of the state, never advanced from
        CK(cudaEventRecord(e2, S->st));
        if (inkernel_ghosts(S)) handshake(S);
        else { int rc = fill_z_ghosts(S, S->state[b]); if (rc) return rc; }
        CK(cudaEventRecord(e3, S->st));
        CK(cudaEventSynchronize(e3));
        float x; cudaEventElapsedTime(&x, e0, e1); t_th += x; cudaEventElapsedTime(&x, e1, e2); t_rhs += x; cudaEventElapsedTime(&x, e2, e3); t_h += x;
        S->launches += 2;
    }
    if (ms_theta) *ms_theta = t_th / reps; if (ms_rhs) *ms_rhs = t_rhs / reps; if (ms_halo) *ms_halo = t_h / reps;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2); cudaEventDestroy(e3);
    return CUDNS_OK;
}

}  // extern "C"


// ---- asynchronous output and restart (writeField / initField, init.cpp:13-30; saveFileMPI / readFileMPI, comm.cpp:205-279) --------
// fields/<c>.<%07d>.bin holds the GLOBAL field [mz_tot][my][mx], raw float64, no header; with z slabs a rank's share is one
// contiguous byte range, so every rank pwrite()s its slab at its offset (what the reference does with an MPI-IO subarray view).
// The step loop is never stalled: the snapshot is a device-side copy on the solver's stream (unpad kernel, ~2 ms at 512^3), the
// D2H copy runs on a second stream into pinned memory, the file system is driven by a writer thread.
static void io_worker(cudns_solver *S) {
    IoState *io = S->io;
    cudaSetDevice(S->P.device);
    for (;;) {
        std::string dir; int ts;
        {
            std::unique_lock<std::mutex> lk(io->m);
            io->cv.wait(lk, [&] { return io->have_job || io->stop; });
            if (!io->have_job && io->stop) return;
            dir = io->dir; ts = io->timestep; io->have_job = false;
        }
        int err = 0; std::string msg;
        cudaError_t ce = cudaEventSynchronize(io->d2h_done);
        if (ce != cudaSuccess) { err = CUDNS_ECUDA; msg = std::string("fields writer: ") + cudaGetErrorString(ce); }
        const size_t N = S->N;
        const off_t off = (off_t)S->P.rank * (off_t)N * (off_t)sizeof(double);
        const char names[5] = {'r', 'u', 'v', 'w', 'e'};
        for (int f = 0; f < 5 && !err; f++) {
            char path[1200];
            std::snprintf(path, sizeof(path), "%s/fields/%c.%07d.bin", dir.c_str(), names[f], ts);
            int fd = ::open(path, O_CREAT | O_WRONLY, 0644);
            if (fd < 0) { err = CUDNS_EINVAL; msg = std::string("cannot open ") + path; break; }
            const char *src = (const char *)(io->h_stage + (size_t)f * N);
            size_t left = N * sizeof(double); off_t o = off;
            while (left > 0) {
                ssize_t w = ::pwrite(fd, src, left > ((size_t)1 << 30) ? ((size_t)1 << 30) : left, o);
                if (w <= 0) { err = CUDNS_EINVAL; msg = std::string("short write ") + path; break; }
                src += w; o += w; left -= (size_t)w;
            }
            ::close(fd);
            if (!err) io->files_written++;
        }
        {
            std::lock_guard<std::mutex> lk(io->m);
            if (err && !io->err) { io->err = err; io->errmsg = msg; }
            io->busy = false;
        }
        io->cv.notify_all();
    }
}

static int io_start(cudns_solver *S) {
    if (S->io && S->io->started) return CUDNS_OK;
    if (!S->io) S->io = new IoState();
    IoState *io = S->io;
    const size_t bytes = 5 * S->N * sizeof(double);
    CK(cudaStreamCreateWithFlags(&io->cs, cudaStreamNonBlocking));
    CK(cudaMalloc((void **)&io->d_stage, bytes));
    CK(cudaMallocHost((void **)&io->h_stage, bytes));
    CK(cudaEventCreateWithFlags(&io->snap_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&io->d2h_done, cudaEventDisableTiming));
    S->bytes += bytes;
    io->worker = std::thread(io_worker, S);
    io->started = true;
    return CUDNS_OK;
}

static void io_shutdown(cudns_solver *S) {
    IoState *io = S->io;
    if (!io) return;
    if (io->started) {
        { std::lock_guard<std::mutex> lk(io->m); io->stop = true; }
        io->cv.notify_all();
        if (io->worker.joinable()) io->worker.join();
    }
    if (io->cs) cudaStreamDestroy(io->cs);
    cudaFree(io->d_stage); cudaFreeHost(io->h_stage);
    if (io->snap_done) cudaEventDestroy(io->snap_done);
    if (io->d2h_done) cudaEventDestroy(io->d2h_done);
    delete io; S->io = nullptr;
}

extern "C" {

int cudns_write_fields_async(cudns_handle S, const char *dir, int timestep) {
    if (!S) { set_error("NULL handle"); return CUDNS_EINVAL; }
    if (!S->have_state) { set_error("no state set"); return CUDNS_ESTATE; }
    CK(cudaSetDevice(S->P.device));
    int rc = io_start(S); if (rc) return rc;
    IoState *io = S->io;
    {   // one snapshot in flight: the staging buffers are free once the previous one is on disk
        std::unique_lock<std::mutex> lk(io->m);
        io->cv.wait(lk, [&] { return !io->busy; });
        if (io->err) { set_error(io->errmsg); return io->err; }
    }
    std::string d = dir ? dir : ".";
    ::mkdir(d.c_str(), 0755);
    ::mkdir((d + "/fields").c_str(), 0755);
    const size_t N = S->N;
    double *dst[5] = {io->d_stage, io->d_stage + N, io->d_stage + 2 * N, io->d_stage + 3 * N, io->d_stage + 4 * N};
    launch_unpad(S->kc, S->state[S->cur], dst, S->st); S->launches++;
    CK(cudaEventRecord(io->snap_done, S->st));
    CK(cudaStreamWaitEvent(io->cs, io->snap_done, 0));
    CK(cudaMemcpyAsync(io->h_stage, io->d_stage, 5 * N * sizeof(double), cudaMemcpyDeviceToHost, io->cs));
    CK(cudaEventRecord(io->d2h_done, io->cs));
    {
        std::lock_guard<std::mutex> lk(io->m);
        io->dir = d; io->timestep = timestep; io->have_job = true; io->busy = true;
    }
    io->cv.notify_all();
    return CUDNS_OK;
}

int cudns_io_wait(cudns_handle S, uint64_t *files_written) {
    if (!S) { set_error("NULL handle"); return CUDNS_EINVAL; }
    IoState *io = S->io;
    if (files_written) *files_written = 0;
    if (!io || !io->started) return CUDNS_OK;
    std::unique_lock<std::mutex> lk(io->m);
    io->cv.wait(lk, [&] { return !io->busy; });
    if (files_written) *files_written = io->files_written;
    if (io->err) { set_error(io->errmsg); int e = io->err; io->err = 0; return e; }
    return CUDNS_OK;
}

// restart: this rank's slab of fields/{r,u,v,w,e}.<timestep>.bin -> cudns_set_state (readFileMPI reads the whole file on rank 0
// and broadcasts it; here every rank seeks to its own byte range)
int cudns_read_fields(cudns_handle S, const char *dir, int timestep) {
    if (!S) { set_error("NULL handle"); return CUDNS_EINVAL; }
    const size_t N = S->N;
    std::vector<double> buf(5 * N);
    const char names[5] = {'r', 'u', 'v', 'w', 'e'};
    for (int f = 0; f < 5; f++) {
        char path[1200];
        std::snprintf(path, sizeof(path), "%s/fields/%c.%07d.bin", dir ? dir : ".", names[f], timestep);
        FILE *fp = std::fopen(path, "rb");
        if (!fp) { set_error(std::string("cannot open ") + path); return CUDNS_EINVAL; }
        if (fseeko(fp, (off_t)S->P.rank * (off_t)N * (off_t)sizeof(double), SEEK_SET) != 0) { std::fclose(fp); set_error(std::string("cannot seek ") + path); return CUDNS_EINVAL; }
        size_t n = std::fread(buf.data() + (size_t)f * N, sizeof(double), N, fp);
        std::fclose(fp);
        if (n != N) { set_error(std::string("short read ") + path); return CUDNS_EINVAL; }
    }
    return cudns_set_state(S, buf.data(), buf.data() + N, buf.data() + 2 * N, buf.data() + 3 * N, buf.data() + 4 * N);
}

}  // extern "C"


// ---- on-device diagnostics of the channel / boundary-layer workflows -----------------------------
static int prof_scratch(cudns_solver *S) {
    if (S->d_prof) return CUDNS_OK;
    return dmalloc(S, &S->d_prof, (size_t)profile_partial_doubles(S->kc) + 10 * (size_t)S->L.mx + 8);
}

extern "C" {

// calcAvgChan, init.cpp:150-208
int cudns_calc_profiles(cudns_handle S, double *prof) {
    if (!S || !prof) { set_error("NULL argument"); return CUDNS_EINVAL; }
    if (!S->have_state) { set_error("no state set"); return CUDNS_ESTATE; }
    CK(cudaSetDevice(S->P.device));
    int rc = need_allreduce(S); if (rc) return rc;
    rc = prof_scratch(S); if (rc) return rc;
    const int mx = S->L.mx;
    double *partial = S->d_prof, *mean = partial + profile_partial_doubles(S->kc), *var = mean + 5 * mx;
    const double scale = 1.0 / ((double)S->P.my * (double)S->P.mz);        // global row count: slabs add up to the whole plane
    const real *q = S->state[S->cur];
    launch_profile_partial(S->kc, q, nullptr, partial, 0, S->st);
    launch_profile_combine(S->kc, partial, mean, scale, S->st);
    reduce_across(S, mean, 5 * mx, 1);
    launch_profile_favre(S->kc, mean, S->st);
    launch_profile_partial(S->kc, q, mean, partial, 1, S->st);
    launch_profile_combine(S->kc, partial, var, scale, S->st);
    reduce_across(S, var, 5 * mx, 1);
    S->launches += 5;
    CK(cudaMemcpyAsync(prof, mean, 10 * (size_t)mx * sizeof(double), cudaMemcpyDeviceToHost, S->st));
    CK(cudaStreamSynchronize(S->st));
    return CUDNS_OK;
}

// printRes, init.cpp:210-256 (the wall at i = 0; meaningful for the non-periodic-x set-ups)
int cudns_calc_retau(cudns_handle S, double *retau) {
    if (!S || !retau) { set_error("NULL argument"); return CUDNS_EINVAL; }
    if (!S->have_state) { set_error("no state set"); return CUDNS_ESTATE; }
    CK(cudaSetDevice(S->P.device));
    int rc = need_allreduce(S); if (rc) return rc;
    rc = prof_scratch(S); if (rc) return rc;
    double *partial = S->d_prof, *out = partial + profile_partial_doubles(S->kc) + 10 * (size_t)S->L.mx;
    launch_retau(S->kc, S->state[S->cur], partial, out, 1.0 / ((double)S->P.my * (double)S->P.mz), S->st);
    reduce_across(S, out, 1, 1);
    S->launches += 2;
    CK(cudaMemcpyAsync(retau, out, sizeof(double), cudaMemcpyDeviceToHost, S->st));
    CK(cudaStreamSynchronize(S->st));
    return CUDNS_OK;
}

}  // extern "C"


// ---- post-processing statistics: postproc/post.cpp as device reductions -----------------------------
namespace {
struct PostPtrs { double *partial, *mean, *fluc, *bulk, *ret; };
PostPtrs post_ptrs(cudns_solver *S) {
    PostPtrs p;
    p.partial = S->d_post; p.mean = p.partial + post_partial_doubles(S->kc); p.fluc = p.mean + 13 * (size_t)S->L.mx;
    p.bulk = p.fluc + 13 * (size_t)S->L.mx; p.ret = p.bulk + 13;
    return p;
}
size_t post_doubles(cudns_solver *S) { return (size_t)post_partial_doubles(S->kc) + 26 * (size_t)S->L.mx + 16; }
}  // namespace

extern "C" {

// main() of post.cpp up to the first loop (post.cpp:155-171): zero the accumulators; nsnapshots = endfile - initfile + 1
int cudns_stats_begin(cudns_handle S, int nsnapshots) {
    if (!S) { set_error("NULL handle"); return CUDNS_EINVAL; }
    if (nsnapshots < 1) { set_error("nsnapshots < 1"); return CUDNS_EINVAL; }
    CK(cudaSetDevice(S->P.device));
    { int rc0 = need_allreduce(S); if (rc0) return rc0; }
    if (!S->d_post) { int rc = dmalloc(S, &S->d_post, post_doubles(S)); if (rc) return rc; }
    CK(cudaMemsetAsync(S->d_post, 0, post_doubles(S) * sizeof(double), S->st));
    S->post_phase = 1; S->post_files = nsnapshots; S->post_added = 0;
    return CUDNS_OK;
}

// calcState + addMean(mean) + addMean(bulk) + calcRet of the current state (post.cpp:173-179)
int cudns_stats_add_mean(cudns_handle S) {
    if (!S) { set_error("NULL handle"); return CUDNS_EINVAL; }
    if (S->post_phase != 1) { set_error("cudns_stats_add_mean: call cudns_stats_begin first"); return CUDNS_ESTATE; }
    if (!S->have_state) { set_error("no state set"); return CUDNS_ESTATE; }
    if (S->post_added >= S->post_files) { set_error("cudns_stats_add_mean: more snapshots than cudns_stats_begin announced"); return CUDNS_ESTATE; }
    CK(cudaSetDevice(S->P.device));
    const PostPtrs p = post_ptrs(S);
    const double rows = (double)S->P.my * (double)S->P.mz;
    launch_post_accumulate(S->kc, S->state[S->cur], nullptr, p.partial, p.mean, 1.0 / (rows * S->post_files), 0, S->st);
    if (!S->P.periodicX && S->P.stencilSize <= 3)        // walls on both sides; the reference's index arithmetic leaves its array for s = 4
        launch_post_ret(S->kc, S->state[S->cur], 1.0 / S->kc.d1[0], p.partial, p.ret, 1.0 / rows, S->st);
    S->launches += 4; S->post_added++;
    return CUDNS_OK;
}

// after the last snapshot: sums over the slabs, file averages, Favre division (post.cpp:180-187)
int cudns_stats_finish_mean(cudns_handle S) {
    if (!S) { set_error("NULL handle"); return CUDNS_EINVAL; }
    if (S->post_phase != 1 || S->post_added != S->post_files) { set_error("cudns_stats_finish_mean: not every announced snapshot has been added"); return CUDNS_ESTATE; }
    CK(cudaSetDevice(S->P.device));
    const PostPtrs p = post_ptrs(S);
    reduce_across(S, p.mean, 13 * S->L.mx, 1);
    reduce_across(S, p.ret, 2, 1);
    launch_post_finish_mean(S->kc, p.mean, p.bulk, p.ret, 1.0 / S->post_files, S->st); S->launches++;
    S->post_phase = 2; S->post_added = 0;
    return CUDNS_OK;
}

// calcState + addFluc of the current state (post.cpp:188-192)
int cudns_stats_add_fluc(cudns_handle S) {
    if (!S) { set_error("NULL handle"); return CUDNS_EINVAL; }
    if (S->post_phase != 2) { set_error("cudns_stats_add_fluc: call cudns_stats_finish_mean first"); return CUDNS_ESTATE; }
    if (!S->have_state) { set_error("no state set"); return CUDNS_ESTATE; }
    if (S->post_added >= S->post_files) { set_error("cudns_stats_add_fluc: more snapshots than cudns_stats_begin announced"); return CUDNS_ESTATE; }
    CK(cudaSetDevice(S->P.device));
    const PostPtrs p = post_ptrs(S);
    launch_post_accumulate(S->kc, S->state[S->cur], p.mean, p.partial, p.fluc, 1.0 / ((double)S->P.my * (double)S->P.mz * S->post_files), 1, S->st);
    S->launches += 2; S->post_added++;
    return CUDNS_OK;
}

int cudns_stats_get(cudns_handle S, double *mean, double *fluc, double *bulk, double *retau, double *utau) {
    if (!S) { set_error("NULL handle"); return CUDNS_EINVAL; }
    if (S->post_phase < 2) { set_error("cudns_stats_get: the means are not final (cudns_stats_finish_mean)"); return CUDNS_ESTATE; }
    CK(cudaSetDevice(S->P.device));
    const PostPtrs p = post_ptrs(S);
    if (S->post_phase == 2 && S->post_added == S->post_files) { reduce_across(S, p.fluc, 13 * S->L.mx, 1); S->post_phase = 3; }
    if (fluc && S->post_phase != 3) { set_error("cudns_stats_get: fluctuations requested before every snapshot went through cudns_stats_add_fluc"); return CUDNS_ESTATE; }
    const size_t nb = 13 * (size_t)S->L.mx * sizeof(double);
    if (mean) CK(cudaMemcpyAsync(mean, p.mean, nb, cudaMemcpyDeviceToHost, S->st));
    if (fluc) CK(cudaMemcpyAsync(fluc, p.fluc, nb, cudaMemcpyDeviceToHost, S->st));
    if (bulk) CK(cudaMemcpyAsync(bulk, p.bulk, 13 * sizeof(double), cudaMemcpyDeviceToHost, S->st));
    double r2[2] = {0, 0};
    CK(cudaMemcpyAsync(r2, p.ret, 2 * sizeof(double), cudaMemcpyDeviceToHost, S->st));
    CK(cudaStreamSynchronize(S->st));
    if (retau) *retau = r2[0]; if (utau) *utau = r2[1];
    return CUDNS_OK;
}

// the whole of post.cpp's main(): two passes over fields/<c>.<first..last>.bin of `dir`, then mean.txt, fluc.txt, bulk.txt into outdir
int cudns_postprocess(cudns_handle S, const char *dir, int first, int last, const double *x, const char *outdir) {
    if (!S || !x) { set_error("NULL argument"); return CUDNS_EINVAL; }
    if (last < first) { set_error("cudns_postprocess: last < first"); return CUDNS_EINVAL; }
    int rc = cudns_stats_begin(S, last - first + 1); if (rc) return rc;
    for (int f = first; f <= last; f++) { rc = cudns_read_fields(S, dir, f); if (rc) return rc; rc = cudns_stats_add_mean(S); if (rc) return rc; }
    rc = cudns_stats_finish_mean(S); if (rc) return rc;
    for (int f = first; f <= last; f++) { rc = cudns_read_fields(S, dir, f); if (rc) return rc; rc = cudns_stats_add_fluc(S); if (rc) return rc; }
    const int mx = S->L.mx;
    std::vector<double> mean(13 * (size_t)mx), fluc(13 * (size_t)mx), bulk(13);
    double ret = 0, ut = 0;
    rc = cudns_stats_get(S, mean.data(), fluc.data(), bulk.data(), &ret, &ut); if (rc) return rc;
    if (S->P.rank != 0) return CUDNS_OK;
    return cudns_stats_write(outdir, mx, x, mean.data(), fluc.data(), bulk.data(), ret, ut);
}

}  // extern "C"
