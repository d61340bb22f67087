// dispatcher of the fifth-generation stage kernel over the stencil half-width (instantiations: stage_duo_s1..4.cu)
#include "cudns_internal.h"
#include "stage_point.h"
namespace cudns {
void launch_duo_s1(const KConst &, const StagePtrs &, const StageCoef &, const DuoMaps &, cudaStream_t);
void launch_duo_s2(const KConst &, const StagePtrs &, const StageCoef &, const DuoMaps &, cudaStream_t);
void launch_duo_s3(const KConst &, const StagePtrs &, const StageCoef &, const DuoMaps &, cudaStream_t);
void launch_duo_s4(const KConst &, const StagePtrs &, const StageCoef &, const DuoMaps &, cudaStream_t);
int duo_smem_s1(); int duo_smem_s2(); int duo_smem_s3(); int duo_smem_s4();

// preconditions (checked by the caller, api.cu): periodic x, uniform grid, linear viscosity law, even mx; p.qin / p.qbase / p.qout
// are 8-field buffers (rho,u,v,w,rho*E,H,T,theta); every Runge-Kutta stage shape is served
void launch_rhs_stage_duo(const KConst &kc, const StagePtrs &p, const StageCoef &c, const DuoMaps &maps, cudaStream_t st) {
    switch (kc.s) {
        case 1: launch_duo_s1(kc, p, c, maps, st); break;
        case 2: launch_duo_s2(kc, p, c, maps, st); break;
        case 3: launch_duo_s3(kc, p, c, maps, st); break;
        default: launch_duo_s4(kc, p, c, maps, st); break;
    }
}
int duo_smem_bytes(int s) {
    switch (s) { case 1: return duo_smem_s1(); case 2: return duo_smem_s2(); case 3: return duo_smem_s3(); default: return duo_smem_s4(); }
}

// H, T of every cell of a padded 8-field state buffer (ghosts included) from its (rho,u,v,w,rho*E): for buffers the stage kernel did
// not write itself (cudns_set_state, the lean kernels, a separate ghost exchange)
__global__ void __launch_bounds__(256) derive_aux_kernel(const __grid_constant__ KConst c, real *__restrict__ q8) {
    const size_t vol = c.L.vol;
    for (size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x; n < vol; n += (size_t)gridDim.x * blockDim.x) {
        const real r = q8[n], u = q8[vol + n], v = q8[2 * vol + n], w = q8[3 * vol + n], e = q8[4 * vol + n];
        real H, T;
        fast::eos_ht(c, r, RC(1.0) / r, u, v, w, e, H, T);
        q8[5 * vol + n] = H; q8[6 * vol + n] = T;
    }
}
void launch_derive_aux(const KConst &kc, real *q8, cudaStream_t st) { derive_aux_kernel<<<148 * 8, 256, 0, st>>>(kc, q8); }

#ifdef CUDNS_F32
// the fourth generation exists in double precision only (api.cu never selects it in this build)
void launch_rhs_stage_fast(const KConst &, const StagePtrs &, const StageCoef &, const FastMaps &, int, cudaStream_t) {}
int fast_smem_bytes(int, int) { return 0; }
void launch_theta_march(const KConst &, const real *, real *, cudaStream_t);
#endif
}  // namespace cudns
