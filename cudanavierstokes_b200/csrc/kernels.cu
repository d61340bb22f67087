// libcudns device code besides the stage kernels (stage_fast.cu, stage_lean.inc) and the dilatation pass (theta.cu): ghost
// handling, staging copies, reductions, device scalars, the cross-GPU hand-shake.  Written from scratch for sm_100a; the
// numerics restate the reference (simone-silvestri/CudaNavierStokes) -- citations are paths relative to that repository.
//
// Data layout: every state field is a padded ghost-cell array [pz][py][px] (Layout in cudns_internal.h).
#include "cudns_internal.h"
#include <cstdio>
#include <cstdlib>
#include <string>

namespace cudns {

typedef double acc_t;             // sums run in double whatever the working precision is (a float sum over 1e8 points loses five digits)

__device__ __forceinline__ real visc_of(const KConst &c, real t) {
    // mu = T^viscexp / Re   (cuda_main.cu:237-238); common exponents avoid the generic pow
    switch (c.viscmode) {
        case 1: return t * c.invRe;
        case 2: return sqrt(t) * c.invRe;
        case 3: { real s = sqrt(t); return s * sqrt(s) * c.invRe; }
        case 4: return t * sqrt(t) * c.invRe;
        default: return pow(t, c.viscexp) * c.invRe;
    }
}

void launch_theta_march(const KConst &kc, const real *q, real *theta, cudaStream_t st);   // theta.cu
void launch_theta(const KConst &kc, const real *q, real *theta, cudaStream_t st) { launch_theta_march(kc, q, theta, st); }

bool rhs_stage_supported(int s, int v) { return s >= 1 && s <= 4 && v >= 1 && v <= s; }

// ---------------------------------------------------------------------------------------------
// wall-normal profiles and friction Reynolds number on the device (calcAvgChan init.cpp:150-208, printRes :210-256: the
// reference copies the five fields to the host and loops there).  Deterministic: block partials, combined in a fixed order.
// ---------------------------------------------------------------------------------------------
constexpr int PROF_NB = 64;          // row blocks
// PASS 0: sums of rho, rho u, rho v, rho w, rho E over this slab's (j,k) rows, per i.  PASS 1: sums of squared deviations of
// rho, u, v, w, rho E from mean[5][mx].  partial[PROF_NB][5][mx]
template <int PASS>
__global__ void __launch_bounds__(256) profile_partial_kernel(KConst c, const real *__restrict__ q, const double *__restrict__ mean,
                                                              double *__restrict__ partial) {
    __shared__ acc_t red[8][5][32];
    const Layout &L = c.L;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + tx, ic = min(i, L.mx - 1);
    const long rows = (long)L.my * L.mz;
    double m[5] = {0, 0, 0, 0, 0}; acc_t a[5] = {0, 0, 0, 0, 0};
    if (PASS == 1) for (int n = 0; n < 5; n++) m[n] = mean[n * L.mx + ic];
    for (long r = (long)blockIdx.y * 8 + ty; r < rows; r += (long)PROF_NB * 8) {
        const int j = (int)(r % L.my), k = (int)(r / L.my);
        const size_t g = L.idx(ic, j, k);
        const real rr = q[g], u = q[L.vol + g], v = q[2 * L.vol + g], w = q[3 * L.vol + g], e = q[4 * L.vol + g];
        if (PASS == 0) { a[0] += rr; a[1] += rr * u; a[2] += rr * v; a[3] += rr * w; a[4] += e; }
        else { a[0] += (rr - m[0]) * (rr - m[0]); a[1] += (u - m[1]) * (u - m[1]); a[2] += (v - m[2]) * (v - m[2]);
               a[3] += (w - m[3]) * (w - m[3]); a[4] += (e - m[4]) * (e - m[4]); }
    }
    for (int n = 0; n < 5; n++) red[ty][n][tx] = a[n];
    __syncthreads();
    if (ty < 5 && i < L.mx) {
        acc_t sum = 0.0;
        for (int t = 0; t < 8; t++) sum += red[t][ty][tx];
        partial[((size_t)blockIdx.y * 5 + ty) * L.mx + i] = sum;
    }
}
// out[5][mx] = scale * sum over the PROF_NB blocks (fixed order)
__global__ void profile_combine_kernel(int mx, const double *__restrict__ partial, double *__restrict__ out, double scale) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= 5 * mx) return;
    acc_t sum = 0.0;
    for (int b = 0; b < PROF_NB; b++) sum += partial[(size_t)b * 5 * mx + n];
    out[n] = sum * scale;
}
// Favre means: <rho u_m> / <rho>  (after the cross-rank sum)
__global__ void profile_favre_kernel(int mx, double *mean) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= mx) return;
    const double rm = mean[i];
    mean[mx + i] /= rm; mean[2 * mx + i] /= rm; mean[3 * mx + i] /= rm;
}
void launch_profile_partial(const KConst &kc, const real *q, const double *mean, double *partial, int pass, cudaStream_t st) {
    dim3 grid((kc.L.mx + 31) / 32, PROF_NB);
    if (pass == 0) profile_partial_kernel<0><<<grid, 256, 0, st>>>(kc, q, mean, partial);
    else profile_partial_kernel<1><<<grid, 256, 0, st>>>(kc, q, mean, partial);
}
void launch_profile_combine(const KConst &kc, const double *partial, double *out, double scale, cudaStream_t st) {
    profile_combine_kernel<<<(5 * kc.L.mx + 127) / 128, 128, 0, st>>>(kc.L.mx, partial, out, scale);
}
void launch_profile_favre(const KConst &kc, double *mean, cudaStream_t st) { profile_favre_kernel<<<(kc.L.mx + 127) / 128, 128, 0, st>>>(kc.L.mx, mean); }
int profile_partial_doubles(const KConst &kc) { return PROF_NB * 5 * kc.L.mx; }

// friction Reynolds number of the wall at i = 0: per (j,k) the one-sided advective-order stencil on the anti-mirrored w
// (ub[g] = w[s-g-1] for the ghosts), u_tau = sqrt(mu_w |dw/dx| / rho), Re_tau += u_tau rho / mu_w.  partial[PROF_NB]
__global__ void __launch_bounds__(256) retau_partial_kernel(KConst c, const real *__restrict__ q, double *__restrict__ partial) {
    __shared__ acc_t red[256];
    const Layout &L = c.L;
    const long rows = (long)L.my * L.mz;
    const real muw = pow(1.0, c.viscexp) * c.invRe;
    acc_t a = 0.0;
    for (long r = (long)blockIdx.x * 256 + threadIdx.x; r < rows; r += (long)PROF_NB * 256) {
        const int j = (int)(r % L.my), k = (int)(r / L.my);
        const size_t g = L.idx(0, j, k);
        const real *w = q + 3 * L.vol + g;
        real dudx = 0.0;
        for (int it = 0; it < c.s; it++) dudx += -c.aF[c.s - it] * (w[c.s - it - 1] - w[c.s - it]) * c.d1[0];    // coeffF[it] = -a_{s-it}
        dudx *= c.xp[0];
        const real rr = q[g];
        a += sqrt(muw * fabs(dudx) / rr) * rr / muw;
    }
    red[threadIdx.x] = a;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) { if ((int)threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st]; __syncthreads(); }
    if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}
__global__ void retau_combine_kernel(const double *partial, double *out, double scale) {
    acc_t sum = 0.0;
    for (int b = 0; b < PROF_NB; b++) sum += partial[b];
    *out = sum * scale;
}
void launch_retau(const KConst &kc, const real *q, double *partial, double *out, double scale, cudaStream_t st) {
    retau_partial_kernel<<<PROF_NB, 256, 0, st>>>(kc, q, partial);
    retau_combine_kernel<<<1, 1, 0, st>>>(partial, out, scale);
}

// ---------------------------------------------------------------------------------------------
// post-processing statistics on the device (postproc/post.cpp:126-326: the reference reads every saved field back to the host and
// loops there, twice).  13 quantities per point in the column order of Variables::printFile (post.cpp:61-86):
// rho, rho u, rho v, rho w, u, v, w, rho E, rho h, h, T, p, mu (the Favre rows hold rho-weighted sums until the division).
// PASS 0: acc[13][mx] += scale * sum over this slab's (j,k) rows of q        (addMean, post.cpp:225-257)
// PASS 1: acc[13][mx] += scale * sum of (q' - mean)^2, q' = u,v,w,h in the Favre rows   (addFluc, post.cpp:201-223)
// Deterministic: block partials [POST_NB][13][mx], combined in a fixed order.
// ---------------------------------------------------------------------------------------------
constexpr int POST_NB = 64, POST_NQ = 13;
__device__ __forceinline__ void post_point(const KConst &c, const real *__restrict__ q, size_t g, real (&o)[POST_NQ]) {
    const Layout &L = c.L;
    const real r = q[g], u = q[L.vol + g], v = q[2 * L.vol + g], w = q[3 * L.vol + g], e = q[4 * L.vol + g];
    const real invrho = 1.0 / r;                                              // calcState, post.cpp:259-278
    const real en = e * invrho - 0.5 * (u * u + v * v + w * w);
    const real t = c.cvInv * en, p = r * c.Rgas * t, h = (e + p) * invrho, m = visc_of(c, t);
    o[0] = r; o[1] = r * u; o[2] = r * v; o[3] = r * w; o[4] = u; o[5] = v; o[6] = w; o[7] = e; o[8] = r * h; o[9] = h; o[10] = t; o[11] = p; o[12] = m;
}
template <int PASS>
__global__ void __launch_bounds__(256) post_partial_kernel(KConst c, const real *__restrict__ q, const double *__restrict__ mean,
                                                           double *__restrict__ partial) {
    __shared__ acc_t red[8][POST_NQ][32];
    const Layout &L = c.L;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + tx, ic = min(i, L.mx - 1);
    const long rows = (long)L.my * L.mz;
    double m[POST_NQ]; acc_t a[POST_NQ];
#pragma unroll
    for (int n = 0; n < POST_NQ; n++) { a[n] = 0.0; m[n] = PASS == 1 ? mean[n * L.mx + ic] : 0.0; }
    for (long r = (long)blockIdx.y * 8 + ty; r < rows; r += (long)POST_NB * 8) {
        const int j = (int)(r % L.my), k = (int)(r / L.my);
        real o[POST_NQ];
        post_point(c, q, L.idx(ic, j, k), o);
        if (PASS == 1) { o[1] = o[4]; o[2] = o[5]; o[3] = o[6]; o[8] = o[9]; }
#pragma unroll
        for (int n = 0; n < POST_NQ; n++) a[n] += PASS == 0 ? o[n] : (o[n] - m[n]) * (o[n] - m[n]);
    }
#pragma unroll
    for (int n = 0; n < POST_NQ; n++) red[ty][n][tx] = a[n];
    __syncthreads();
    for (int n = ty; n < POST_NQ; n += 8) {
        if (i < L.mx) {
            acc_t sum = 0.0;
            for (int t = 0; t < 8; t++) sum += red[t][n][tx];
            partial[((size_t)blockIdx.y * POST_NQ + n) * L.mx + i] = sum;
        }
    }
}
__global__ void post_combine_kernel(int mx, const double *__restrict__ partial, double *__restrict__ acc, double scale) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= POST_NQ * mx) return;
    acc_t sum = 0.0;
    for (int b = 0; b < POST_NB; b++) sum += partial[(size_t)b * POST_NQ * mx + n];
    acc[n] += sum * scale;
}
// calcRet (post.cpp:280-326) of one snapshot: per (j,k) the mean of |dw/dx| at the first two points next to either wall, advective
// coefficients on the anti-mirrored w, wall density from the mean wall pressure (T_wall = 1); ret2[0] += Re_tau, ret2[1] += u_tau
// (this slab's share of the mean over all rows).  partial[2][POST_NB]
__global__ void __launch_bounds__(256) post_ret_partial_kernel(KConst c, const real *__restrict__ q, double host_dx, double *__restrict__ partial) {
    __shared__ acc_t red[2][256];
    const Layout &L = c.L;
    const long rows = (long)L.my * L.mz;
    const int s = c.s, mx = L.mx;
    const real muw = c.invRe;
    acc_t aR = 0.0, aU = 0.0;
    for (long r = (long)blockIdx.x * 256 + threadIdx.x; r < rows; r += (long)POST_NB * 256) {
        const int j = (int)(r % L.my), k = (int)(r / L.my);
        const size_t g = L.idx(0, j, k);
        const real *w = q + 3 * L.vol + g;
        real o0[POST_NQ], o1[POST_NQ];
        post_point(c, q, g, o0); post_point(c, q, g + mx - 1, o1);
        const real rw = 0.5 * (o0[11] + o1[11]) / c.Rgas;
        // ub[n], n = 0 .. mx+2s+1: ub[n] = w[n-s-1] inside, -w[s-n] below the lower wall, -w[2mx+s-n] above the upper one
        auto ub = [&](int n) -> real { return n < 0 ? 0.0 : n <= s ? -w[s - n] : n <= mx + s ? w[n - s - 1] : -w[2 * mx + s - n]; };
        auto dudx = [&](int jj) -> real {
            real d = 0.0;
            for (int it = 0; it < s; it++) d += -c.aF[s - it] * (ub(jj + it - s) - ub(jj - it + s)) / host_dx;       // coeffF[it] = -a_{s-it}
            return d;
        };
        const real avg = (fabs(dudx(3)) + fabs(dudx(4)) + fabs(dudx(mx + s)) + fabs(dudx(mx + s + 1))) * 0.25 * c.xp[0];
        const real ut = sqrt(muw * avg / rw);
        aU += ut; aR += ut * rw / muw;
    }
    red[0][threadIdx.x] = aR; red[1][threadIdx.x] = aU;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if ((int)threadIdx.x < st) { red[0][threadIdx.x] += red[0][threadIdx.x + st]; red[1][threadIdx.x] += red[1][threadIdx.x + st]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { partial[blockIdx.x] = red[0][0]; partial[POST_NB + blockIdx.x] = red[1][0]; }
}
__global__ void post_ret_combine_kernel(const double *partial, double *ret2, double scale) {
    for (int n = 0; n < 2; n++) {
        acc_t sum = 0.0;
        for (int b = 0; b < POST_NB; b++) sum += partial[n * POST_NB + b];
        ret2[n] += sum * scale;
    }
}
// after the last snapshot (and the cross-rank sum): volume averages from the un-divided profiles (addMean with N = 1: weight
// dxv[i]/Lx), averages of Re_tau, u_tau over the snapshots, Favre division of rows 1-3 and 8 (post.cpp:180-187)
__global__ void post_finish_mean_kernel(KConst c, double *mean, double *bulk, double *ret2, double inv_files) {
    const int mx = c.L.mx;
    const int n = threadIdx.x;
    if (n < POST_NQ) {
        acc_t sum = 0.0;
        for (int i = 0; i < mx; i++) sum += mean[n * mx + i] * c.dxv[i] / c.Lx;
        bulk[n] = sum;
    }
    if (n == POST_NQ) { ret2[0] *= inv_files; ret2[1] *= inv_files; }
    __syncthreads();
    if (n == 0) { bulk[1] /= bulk[0]; bulk[2] /= bulk[0]; bulk[3] /= bulk[0]; bulk[8] /= bulk[0]; }
    for (int i = n; i < mx; i += blockDim.x) {
        const double r = mean[i];
        mean[mx + i] /= r; mean[2 * mx + i] /= r; mean[3 * mx + i] /= r; mean[8 * mx + i] /= r;
    }
}
int post_partial_doubles(const KConst &kc) { return POST_NB * POST_NQ * kc.L.mx; }
void launch_post_accumulate(const KConst &kc, const real *q, const double *mean, double *partial, double *acc, double scale, int pass, cudaStream_t st) {
    dim3 grid((kc.L.mx + 31) / 32, POST_NB);
    if (pass == 0) post_partial_kernel<0><<<grid, 256, 0, st>>>(kc, q, mean, partial);
    else post_partial_kernel<1><<<grid, 256, 0, st>>>(kc, q, mean, partial);
    post_combine_kernel<<<(POST_NQ * kc.L.mx + 127) / 128, 128, 0, st>>>(kc.L.mx, partial, acc, scale);
}
void launch_post_ret(const KConst &kc, const real *q, double host_dx, double *partial, double *ret2, double scale, cudaStream_t st) {
    post_ret_partial_kernel<<<POST_NB, 256, 0, st>>>(kc, q, host_dx, partial);
    post_ret_combine_kernel<<<1, 1, 0, st>>>(partial, ret2, scale);
}
void launch_post_finish_mean(const KConst &kc, double *mean, double *bulk, double *ret2, double inv_files, cudaStream_t st) {
    post_finish_mean_kernel<<<1, 128, 0, st>>>(kc, mean, bulk, ret2, inv_files);
}

// ---------------------------------------------------------------------------------------------
// ghost handling, staging copies
// ---------------------------------------------------------------------------------------------
// periodic x/y images for nfields padded fields, all local planes incl. z ghosts (used by set_state)
__global__ void fill_xy_kernel(KConst c, real *q, int nfields) {
    const Layout &L = c.L;
    int kz = blockIdx.z;                    // padded plane index
    int f = blockIdx.y;
    real *p = q + (size_t)f * L.vol + (size_t)kz * L.plane;
    int s = c.s;
    // x images for interior rows
    if (c.periodicX) {
        for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < L.my * s; t += gridDim.x * blockDim.x) {
            int j = t / s, gq = t % s;
            real *row = p + (size_t)(j + L.gy) * L.px + GX;
            row[-1 - gq] = row[L.mx - 1 - gq];
            row[L.mx + gq] = row[gq];
        }
    }
    // y images for interior columns
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < L.mx * s; t += gridDim.x * blockDim.x) {
        int i = t % L.mx, gq = t / L.mx;
        real *col = p + (size_t)L.gy * L.px + GX + i;
        col[-(ptrdiff_t)(1 + gq) * L.px] = col[(size_t)(L.my - 1 - gq) * L.px];
        col[(size_t)(L.my + gq) * L.px] = col[(size_t)gq * L.px];
    }
}
void launch_fill_xy(const KConst &kc, real *q5, int nfields, cudaStream_t st) {
    dim3 grid(8, nfields, kc.L.pz);
    fill_xy_kernel<<<grid, 256, 0, st>>>(kc, q5, nfields);
}

// periodic z wrap on one device: copy gz full padded planes bottom<->top (perBCz boundary.h:48-51)
__global__ void zwrap_kernel(KConst c, real *q, int nfields) {
    const Layout &L = c.L;
    size_t n = (size_t)L.gz * L.plane;      // doubles per block
    int f = blockIdx.y;
    real *p = q + (size_t)f * L.vol;
    vec16 *lo_ghost = (vec16 *)p, *hi_ghost = (vec16 *)(p + (size_t)(L.gz + L.mz) * L.plane);
    const vec16 *lo_int = (const vec16 *)(p + (size_t)L.gz * L.plane), *hi_int = (const vec16 *)(p + (size_t)L.mz * L.plane);
    size_t n2 = n / VEC16;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n2; t += (size_t)gridDim.x * blockDim.x) {
        lo_ghost[t] = hi_int[t];
        hi_ghost[t] = lo_int[t];
    }
}
void launch_zwrap(const KConst &kc, real *q5, int nfields, cudaStream_t st) {
    dim3 grid(148 * 2, nfields);
    zwrap_kernel<<<grid, 256, 0, st>>>(kc, q5, nfields);
}

// pack the first/last gz interior planes of 5 fields into contiguous send blocks; unpack ghost blocks
__global__ void pack_z_kernel(KConst c, const real *q, real *send_lo, real *send_hi) {
    const Layout &L = c.L;
    size_t n2 = (size_t)L.gz * L.plane / VEC16;
    int f = blockIdx.y;
    const real *p = q + (size_t)f * L.vol;
    const vec16 *lo_int = (const vec16 *)(p + (size_t)L.gz * L.plane), *hi_int = (const vec16 *)(p + (size_t)L.mz * L.plane);
    vec16 *slo = (vec16 *)send_lo + (size_t)f * n2, *shi = (vec16 *)send_hi + (size_t)f * n2;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n2; t += (size_t)gridDim.x * blockDim.x) {
        slo[t] = lo_int[t]; shi[t] = hi_int[t];
    }
}
__global__ void unpack_z_kernel(KConst c, real *q, const real *recv_lo, const real *recv_hi) {
    const Layout &L = c.L;
    size_t n2 = (size_t)L.gz * L.plane / VEC16;
    int f = blockIdx.y;
    real *p = q + (size_t)f * L.vol;
    vec16 *lo_ghost = (vec16 *)p, *hi_ghost = (vec16 *)(p + (size_t)(L.gz + L.mz) * L.plane);
    const vec16 *rlo = (const vec16 *)recv_lo + (size_t)f * n2, *rhi = (const vec16 *)recv_hi + (size_t)f * n2;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n2; t += (size_t)gridDim.x * blockDim.x) {
        lo_ghost[t] = rlo[t]; hi_ghost[t] = rhi[t];
    }
}
void launch_pack_z(const KConst &kc, const real *q5, real *send_lo, real *send_hi, cudaStream_t st) {
    pack_z_kernel<<<dim3(148, 5), 256, 0, st>>>(kc, q5, send_lo, send_hi);
}
void launch_unpack_z(const KConst &kc, real *q5, const real *recv_lo, const real *recv_hi, cudaStream_t st) {
    unpack_z_kernel<<<dim3(148, 5), 256, 0, st>>>(kc, q5, recv_lo, recv_hi);
}

// unpadded [mz][my][mx] <-> padded interior (initDevice / getResults, cuda_utils.cu:372-413)
struct Ptr5 { const double *p[5]; };
struct Ptr5w { double *p[5]; };
__global__ void pad_kernel(KConst c, Ptr5 src, real *q) {
    const Layout &L = c.L;
    size_t N = (size_t)L.mx * L.my * L.mz;
    for (size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (size_t)gridDim.x * blockDim.x) {
        int i = (int)(n % L.mx); size_t r = n / L.mx; int j = (int)(r % L.my), k = (int)(r / L.my);
        size_t g = L.idx(i, j, k);
#pragma unroll
        for (int f = 0; f < 5; f++) q[f * L.vol + g] = (real)src.p[f][n];
    }
}
__global__ void unpad_kernel(KConst c, const real *q, Ptr5w dst) {
    const Layout &L = c.L;
    size_t N = (size_t)L.mx * L.my * L.mz;
    for (size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (size_t)gridDim.x * blockDim.x) {
        int i = (int)(n % L.mx); size_t r = n / L.mx; int j = (int)(r % L.my), k = (int)(r / L.my);
        size_t g = L.idx(i, j, k);
#pragma unroll
        for (int f = 0; f < 5; f++) dst.p[f][n] = q[f * L.vol + g];
    }
}
void launch_pad(const KConst &kc, const double *src5[5], real *q5, cudaStream_t st) {
    Ptr5 s; for (int f = 0; f < 5; f++) s.p[f] = src5[f];
    pad_kernel<<<148 * 8, 256, 0, st>>>(kc, s, q5);
}
void launch_unpad(const KConst &kc, const real *q5, double *dst5[5], cudaStream_t st) {
    Ptr5w d; for (int f = 0; f < 5; f++) d.p[f] = dst5[f];
    unpad_kernel<<<148 * 8, 256, 0, st>>>(kc, q5, d);
}

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T warp_max(T v) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <typename T> __device__ __forceinline__ T warp_sum(T v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// positive doubles order like their bit patterns
__device__ __forceinline__ void atomic_max_pos(double *addr, double v) {
    atomicMax((unsigned long long *)addr, (unsigned long long)__double_as_longlong(v));
}

// deviceCalcDt (calc_stress.cu:140-160) as two maxima: min_p CFL/max(conv_p,visc_p) = CFL/max(max_p conv_p, max_p visc_p)
__global__ void __launch_bounds__(256) dt_reduce_kernel(KConst c, const real *__restrict__ q, double *out2) {
    const Layout &L = c.L;
    const int nrows = L.my * L.mz;
    real mc = 0.0, mv = 0.0;
    for (int row = blockIdx.x; row < nrows; row += gridDim.x) {            // one (j,k) row of the interior per iteration
        const int j = row % L.my, k = row / L.my;
        const real *p = q + L.idx(0, j, k);
        for (int i = threadIdx.x; i < L.mx; i += blockDim.x) {
            const real r = p[i], u = p[L.vol + i], v = p[2 * L.vol + i], w = p[3 * L.vol + i], e = p[4 * L.vol + i];
            real ien = e / r - 0.5 * (u * u + v * v + w * w);
            real sos = sqrt(c.gam * (c.gam - 1) * ien);
            real dx = c.dxv[i], d2x = dx * dx;
            real conv = fmax((fabs(u) + sos) / dx, fmax((fabs(v) + sos) * c.d1[1], (fabs(w) + sos) * c.d1[2]));
            real mu = visc_of(c, c.cvInv * ien);
            real visc = fmax(mu / d2x, fmax(mu * c.d2[1], mu * c.d2[2]));
            mc = fmax(mc, conv); mv = fmax(mv, visc);
        }
    }
    mc = warp_max(mc); mv = warp_max(mv);
    if ((threadIdx.x & 31) == 0) { atomic_max_pos(out2, (double)mc); atomic_max_pos(out2 + 1, (double)mv); }
}
void launch_dt_reduce(const KConst &kc, const real *q, double *out2, cudaStream_t st) {
    cudaMemsetAsync(out2, 0, 2 * sizeof(double), st);
    dt_reduce_kernel<<<148 * 8, 256, 0, st>>>(kc, q, out2);
}

// calcBulk / calcPressureGrad integrals (calc_stress.cu:98-120,162-201; cuda_math.cu:143-201):
// out[0] = sum (u.u) w_avg, out[1] = sum rho w_int, out[2] = sum rho*w w_int, out[3] = sum rhoE w_int
// with w_int = dxv[i]/d_dy/d_dz and w_avg = w_int/Lx/Ly/Lz.  Block partials are combined in a fixed order
// by the last block (deterministic for a given launch shape).
__global__ void __launch_bounds__(256) bulk_reduce_kernel(KConst c, const real *__restrict__ q, double *out4, double *partial, unsigned int *counter) {
    const Layout &L = c.L;
    const int nrows = L.my * L.mz;
    acc_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int row = blockIdx.x; row < nrows; row += gridDim.x) {
        const int j = row % L.my, k = row / L.my;
        const real *p = q + L.idx(0, j, k);
        for (int i = threadIdx.x; i < L.mx; i += blockDim.x) {
            const real r = p[i], u = p[L.vol + i], v = p[2 * L.vol + i], w = p[3 * L.vol + i], e = p[4 * L.vol + i];
            real wi = c.dxv[i] / c.d1[1] / c.d1[2];
            s0 += (u * u + v * v + w * w) * wi / c.Lx / c.Ly / c.Lz;
            s1 += r * wi; s2 += r * w * wi; s3 += e * wi;
        }
    }
    __shared__ acc_t sh[4][8];
    s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2); s3 = warp_sum(s3);
    int wid = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { sh[0][wid] = s0; sh[1][wid] = s1; sh[2][wid] = s2; sh[3][wid] = s3; }
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) {
        for (int m = 0; m < 4; m++) { acc_t a = 0; for (int q8 = 0; q8 < 8; q8++) a += sh[m][q8]; partial[m * gridDim.x + blockIdx.x] = a; }
        __threadfence();
        unsigned int t = atomicAdd(counter, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x < 4) {
        acc_t a = 0; for (unsigned int b = 0; b < gridDim.x; b++) a += partial[threadIdx.x * gridDim.x + b];
        out4[threadIdx.x] = a;
        if (threadIdx.x == 0) *counter = 0;
    }
}
constexpr int BULK_NB = 148 * 4;
int bulk_scratch_doubles() { return 4 * BULK_NB + 1; }
void launch_bulk_reduce(const KConst &kc, const real *q, double *out4, double *scratch, cudaStream_t st) {
    bulk_reduce_kernel<<<BULK_NB, 256, 0, st>>>(kc, q, out4, scratch, (unsigned int *)(scratch + 4 * BULK_NB));
}

// Mean square vorticity <w.w> of a periodic box: volume average (weights of calcBulk's <u.u>) of |curl u|^2 with the viscous-order
// first differences of derVelX/Y/Z (calc_stress.cu:20-86) read straight from the ghost-cell state.  Not a reference quantity -- the
// reference never writes par2 without forcing (calc_stress.cu:191-197), so its Taylor-Green runs have no dissipation history; this
// is the one libcudns offers (epsilon = <w.w>/Re in the incompressible limit).  Same deterministic two-level sum as the bulk kernel;
// it runs every checkBulk steps, not per stage (3 of 8 state fields read once: ~0.1 % of the step loop at checkBulk = 10).
__global__ void __launch_bounds__(256) enstrophy_reduce_kernel(KConst c, const real *__restrict__ q, double *out, double *partial, unsigned int *counter) {
    const Layout &L = c.L;
    const int nrows = L.my * L.mz;
    const ptrdiff_t sy = L.px, sz = (ptrdiff_t)L.plane;
    acc_t s0 = 0;
    for (int row = blockIdx.x; row < nrows; row += gridDim.x) {
        const int j = row % L.my, k = row / L.my;
        const real *pu = q + L.vol + L.idx(0, j, k), *pv = pu + L.vol, *pw = pv + L.vol;
        for (int i = threadIdx.x; i < L.mx; i += blockDim.x) {
            real uy = 0, uz = 0, vx = 0, vz = 0, wx = 0, wy = 0;
            for (int l = 1; l <= c.v; l++) {
                const real cx = c.c1[0][l], cy = c.c1[1][l], cz = c.c1[2][l];
                vx = fma(cx, pv[i + l] - pv[i - l], vx); wx = fma(cx, pw[i + l] - pw[i - l], wx);
                uy = fma(cy, pu[i + l * sy] - pu[i - l * sy], uy); wy = fma(cy, pw[i + l * sy] - pw[i - l * sy], wy);
                uz = fma(cz, pu[i + l * sz] - pu[i - l * sz], uz); vz = fma(cz, pv[i + l * sz] - pv[i - l * sz], vz);
            }
            if (c.nonUniformX) { vx *= c.xp[i]; wx *= c.xp[i]; }
            const real ox = wy - vz, oy = uz - wx, oz = vx - uy;
            s0 += (ox * ox + oy * oy + oz * oz) * c.dxv[i] / c.d1[1] / c.d1[2] / c.Lx / c.Ly / c.Lz;
        }
    }
    __shared__ acc_t sh[8];
    s0 = warp_sum(s0);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s0;
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) {
        acc_t a = 0; for (int q8 = 0; q8 < 8; q8++) a += sh[q8];
        partial[blockIdx.x] = a;
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        acc_t a = 0; for (unsigned int b = 0; b < gridDim.x; b++) a += partial[b];
        *out = a; *counter = 0;
    }
}
// scratch: the solver's bulk scratch (bulk_scratch_doubles(): the first BULK_NB partials and the counter are used)
void launch_enstrophy_reduce(const KConst &kc, const real *q, double *out, double *scratch, cudaStream_t st) {
    enstrophy_reduce_kernel<<<BULK_NB, 256, 0, st>>>(kc, q, out, scratch, (unsigned int *)(scratch + 4 * BULK_NB));
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
__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__global__ void halo_wait_kernel(const unsigned long long *slots, int need_lo, int need_hi, unsigned long long epoch,
                                 unsigned long long timeout_ns, unsigned long long *err_word) {
    // slots[0]: written by the lower neighbour, slots[1]: by the upper one
    const unsigned long long t0 = global_ns();
    for (int n = 0; n < 2; n++) {
        if (!(n == 0 ? need_lo : need_hi)) continue;
        unsigned long long v;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(slots + n) : "memory");
            if (v < epoch) {
                __nanosleep(200);
                if (global_ns() - t0 > timeout_ns) { if (*err_word == 0) *err_word = epoch; return; }    // reported by cudns_advance
            }
        } while (v < epoch);
    }
    __threadfence_system();
}
void launch_halo_signal(unsigned long long *peer_lo_slot, unsigned long long *peer_hi_slot, unsigned long long epoch, cudaStream_t st) {
    halo_signal_kernel<<<1, 1, 0, st>>>(peer_lo_slot, peer_hi_slot, epoch);
}
void launch_halo_wait(const unsigned long long *my_slots, int need_lo, int need_hi, unsigned long long epoch, unsigned long long timeout_ns,
                      unsigned long long *err_word, cudaStream_t st) {
    halo_wait_kernel<<<1, 1, 0, st>>>(my_slots, need_lo, need_hi, epoch, timeout_ns, err_word);
}

}  // namespace cudns
