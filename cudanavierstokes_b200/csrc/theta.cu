// libcudns: dilatation pass  theta = du/dx + dv/dy + dw/dz  at viscous order
// (derVelX/Y/Z + calcDil of the reference, calc_stress.cu:20-96, collapsed into one streaming kernel).
//
// The stage kernel needs theta at stencil neighbours, i.e. theta of the whole field has to exist before the
// right-hand side of a stage can be evaluated; this pass is the price: it reads u,v,w once and
// writes theta once (32 B per point of HBM traffic, against 168 B of the stage kernel).
//
// Mapping: one CTA = a 32 x 16 tile of (i,j) columns marching along z, one thread per column.
//   * du/dx and dv/dy read their neighbours from a double-buffered shared plane (u with x halos, v with y halos,
//     one __syncthreads per plane); the loads of plane k+1 are issued before plane k is computed;
//   * dw/dz keeps the 2V+1 most recent w values of the column in REGISTERS (a shifting window: 2V moves per plane,
//     noise next to the memory time of this kernel);
//   * wall / extrapolation ghosts in x are built in shared memory (BCxderVel, boundary_condition_x.h:38-40,67,94),
//     the z extrapolation of the boundary layer (BCzderVel, boundary_condition_z.h:34-40) is applied on the few
//     planes next to the global z boundaries by a slow path that reads global memory directly.
#include "cudns_internal.h"

namespace cudns {
namespace {

constexpr int TXT = 32, TYT = 16, NTT = TXT * TYT;
constexpr int UX = TXT + 2 * GX;

// wall blowing/suction, perturbation.h:25-53
__device__ __forceinline__ bool perturb_theta(const KConst &c, int j, int kglob, double &val) {
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


// botBCzExt / topBCzExt (boundary.h:154-160): dw/dz next to the global z boundaries of the boundary-layer case, where w
// past the boundary is the node extrapolation f[-g] = 2 f[0] - f[g], f[mz-1+g] = 2 f[mz-1] - f[mz-1-g] (slow path)
template <int V>
__device__ __forceinline__ double dwdz_edge(const KConst &c, const double *__restrict__ W, int ic, int jc, int k, int kglob_lo, int kglob_hi) {
    const Layout &L = c.L;
    const size_t g0 = L.idx(ic, jc, k);
    double dwdz = 0.0;
#pragma unroll
    for (int l = 1; l <= V; l++) {
        double wp, wm;
        if (k + l >= kglob_hi) wp = 2.0 * W[L.idx(ic, jc, kglob_hi - 1)] - W[L.idx(ic, jc, 2 * (kglob_hi - 1) - (k + l))];
        else wp = W[g0 + (size_t)l * L.plane];
        if (k - l < kglob_lo) wm = 2.0 * W[L.idx(ic, jc, kglob_lo)] - W[L.idx(ic, jc, 2 * kglob_lo - (k - l))];
        else wm = W[g0 - (size_t)l * L.plane];
        dwdz = fma(c.c1[2][l], wp - wm, dwdz);
    }
    return dwdz;
}

template <int V>
__global__ void __launch_bounds__(NTT, 2)
theta_march_kernel(const __grid_constant__ KConst c, const double *__restrict__ q, double *__restrict__ theta, int zchunk) {
    constexpr int R = 2 * V + 1;
    constexpr int VY = TYT + 2 * V;
    __shared__ double su[2][TYT][UX];
    __shared__ double sv[2][VY][TXT];

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
    const bool bl = c.boundaryLayer != 0;
    const bool perx = c.periodicX != 0;
    const bool xlo = !perx && i0 == 0, xhi = !perx && (i0 + TXT >= L.mx);
    const double *__restrict__ U = q + L.vol, *__restrict__ Vv = q + 2 * L.vol, *__restrict__ W = q + 3 * L.vol;
    const size_t plane = L.plane;

    // per-thread addresses (plane 0): own column, the x-halo cell and the y-halo row this thread also stages
    const size_t g00 = L.idx(ic, jc, 0);
    const bool hx_on = tx < 2 * V && iny;
    const int hxc = tx < V ? GX - V + tx : GX + nxt + (tx - V);           // column in su
    const int hgi = i0 + hxc - GX;
    const bool hx_load = hx_on && (perx || (hgi >= 0 && hgi < L.mx));
    const size_t ghx = L.idx(min(max(hgi, -GX), L.mx + GX - 1), jc, 0);
    const bool hy_on = ty < 2 * V && inx;
    const int hyr = ty < V ? ty : V + nyt + (ty - V);                     // row in sv
    const size_t ghy = L.idx(ic, j0 + hyr - V, 0);
    const double xpi = c.nonUniformX ? c.xp[ic] : 1.0;

    double wr[R];
#pragma unroll
    for (int r = 0; r < R; r++) wr[r] = 0.0;
    // prologue: w of planes kfirst-V .. kfirst+V-1 into ring positions 0 .. 2V-1 (position of plane kfirst+m-V is m)
#pragma unroll
    for (int m = 0; m < 2 * V; m++) wr[m] = W[g00 + (ptrdiff_t)(kfirst + m - V) * (ptrdiff_t)plane];
    // staged values of the next plane
    double un, vn, uhn = 0.0, vhn = 0.0, wn;
    {
        const ptrdiff_t off = (ptrdiff_t)kfirst * (ptrdiff_t)plane;
        un = U[g00 + off]; vn = Vv[g00 + off]; wn = W[g00 + off + (ptrdiff_t)V * (ptrdiff_t)plane];
        if (hx_load) uhn = U[ghx + off];
        if (hy_on) vhn = Vv[ghy + off];
    }

    int buf = 0;
    {
        for (int k = kfirst; k < klast; k++) {
            // wr[V + l] = w of plane k+l
            wr[2 * V] = wn;
            if (active) { su[buf][ty][GX + tx] = un; sv[buf][V + ty][tx] = vn; }
            if (hx_load) su[buf][ty][hxc] = uhn;
            if (hy_on) sv[buf][hyr][tx] = vhn;
            __syncthreads();
            if (k + 1 < klast) {                         // stage plane k+1 (consumed at the top of the next iteration)
                const ptrdiff_t off = (ptrdiff_t)(k + 1) * (ptrdiff_t)plane;
                un = U[g00 + off]; vn = Vv[g00 + off]; wn = W[g00 + off + (ptrdiff_t)V * (ptrdiff_t)plane];
                if (hx_load) uhn = U[ghx + off];
                if (hy_on) vhn = Vv[ghy + off];
            }
            const bool outside = bl && (k < kglob_lo || k >= kglob_hi);    // ghost theta is extrapolated by the stage kernel
            if (xlo || xhi) {
                // BCxderVel: wall (anti-mirror about the face, + blowing/suction) / node extrapolation at the free stream
                if (hx_on && !hx_load && !outside) {
                    double val;
                    if (tx < V) {
                        const int gq = V - tx;                            // ghost -gq
                        val = -su[buf][ty][GX + gq - 1];
                        double pv;
                        if (bl && c.perturbed && perturb_theta(c, j, k + c.kstart, pv)) val = pv;
                    } else {
                        const int gq = tx - V + 1, last = GX + nxt - 1;
                        val = bl ? 2.0 * su[buf][ty][last] - su[buf][ty][last - gq] : -su[buf][ty][last - gq + 1];
                    }
                    su[buf][ty][hxc] = val;
                }
                __syncthreads();
            }
            if (active && !outside) {
                double dudx = 0.0, dvdy = 0.0, dwdz = 0.0;
                const double *ur = &su[buf][ty][GX + tx];
                const double *vr = &sv[buf][V + ty][tx];
#pragma unroll
                for (int l = 1; l <= V; l++) {
                    dudx = fma(c.c1[0][l], ur[l] - ur[-l], dudx);
                    dvdy = fma(c.c1[1][l], vr[l * TXT] - vr[-l * TXT], dvdy);
                }
                if (bl && (k - kglob_lo < V || kglob_hi - 1 - k < V)) {
                    // botBCzExt / topBCzExt (boundary.h:154-160): node extrapolation of w past the global z boundaries
                    dwdz = dwdz_edge<V>(c, W, ic, jc, k, kglob_lo, kglob_hi);
                } else {
#pragma unroll
                    for (int l = 1; l <= V; l++) dwdz = fma(c.c1[2][l], wr[V + l] - wr[V - l], dwdz);
                }
                const double th = fma(dudx, xpi, dvdy) + dwdz;
                double *t = theta + g00 + (ptrdiff_t)k * (ptrdiff_t)plane;
                *t = th;
                // periodic images in x / y (cross-shaped ghosts only): perBCx / perBCy, boundary.h:38-46
                if (perx) {
                    if (i < V) t[L.mx] = th;
                    if (i >= L.mx - V) t[-(ptrdiff_t)L.mx] = th;
                }
                if (j < V) t[(size_t)L.my * L.px] = th;
                if (j >= L.my - V) t[-(ptrdiff_t)((size_t)L.my * L.px)] = th;
            }
            buf ^= 1;
#pragma unroll
            for (int m = 0; m < 2 * V; m++) wr[m] = wr[m + 1];
        }
    }
}

}  // namespace

void launch_theta_march(const KConst &kc, const double *q, double *theta, cudaStream_t st) {
    const int gx = (kc.L.mx + TXT - 1) / TXT, gy = (kc.L.my + TYT - 1) / TYT;
    const int nk = kc.L.mz + 2 * kc.v;
    // z chunks: every chunk pays a 2V-plane prologue of w; aim at a few waves of 148 SMs x 3 CTAs
    int nzc = 1;
    while (gx * gy * nzc < 148 * 3 * 4 && nk / (nzc * 2) >= 32) nzc *= 2;
    int zchunk = (nk + nzc - 1) / nzc;
    nzc = (nk + zchunk - 1) / zchunk;
    dim3 grid(gx, gy, nzc);
    switch (kc.v) {
        case 1: theta_march_kernel<1><<<grid, NTT, 0, st>>>(kc, q, theta, zchunk); break;
        case 2: theta_march_kernel<2><<<grid, NTT, 0, st>>>(kc, q, theta, zchunk); break;
        case 3: theta_march_kernel<3><<<grid, NTT, 0, st>>>(kc, q, theta, zchunk); break;
        default: theta_march_kernel<4><<<grid, NTT, 0, st>>>(kc, q, theta, zchunk); break;
    }
}

}  // namespace cudns
