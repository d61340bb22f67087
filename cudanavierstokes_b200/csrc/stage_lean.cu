// dispatcher of the lean stage kernel over the stencil half-width (instantiations: stage_lean_s1..4.cu)
#include "cudns_internal.h"
namespace cudns {
void launch_lean_s1(const KConst &, const StagePtrs &, const StageCoef &, const LeanMaps &, bool, bool, cudaStream_t);
void launch_lean_s2(const KConst &, const StagePtrs &, const StageCoef &, const LeanMaps &, bool, bool, cudaStream_t);
void launch_lean_s3(const KConst &, const StagePtrs &, const StageCoef &, const LeanMaps &, bool, bool, cudaStream_t);
void launch_lean_s4(const KConst &, const StagePtrs &, const StageCoef &, const LeanMaps &, bool, bool, cudaStream_t);
int lean_smem_s1(bool); int lean_smem_s2(bool); int lean_smem_s3(bool); int lean_smem_s4(bool);
int lean_smem_wide_s1(); int lean_smem_wide_s2(); int lean_smem_wide_s3(); int lean_smem_wide_s4();

void launch_rhs_stage_lean(const KConst &kc, const StagePtrs &p, const StageCoef &c, const LeanMaps &maps, bool wide, cudaStream_t st) {
    // FAST variant: periodic x on a uniform grid outside the boundary-layer set-up (no wall, metric, extrapolation or sponge code)
    const bool gen = !(kc.periodicX && !kc.nonUniformX && !kc.boundaryLayer);
    switch (kc.s) {
        case 1: launch_lean_s1(kc, p, c, maps, gen, wide && !gen, st); break;
        case 2: launch_lean_s2(kc, p, c, maps, gen, wide && !gen, st); break;
        case 3: launch_lean_s3(kc, p, c, maps, gen, wide && !gen, st); break;
        default: launch_lean_s4(kc, p, c, maps, gen, wide && !gen, st); break;
    }
}
bool lean_wide_ok(const KConst &kc) { return kc.viscmode == 1 && kc.periodicX && !kc.nonUniformX && !kc.boundaryLayer; }
int lean_smem_wide_bytes(int s) {
    switch (s) { case 1: return lean_smem_wide_s1(); case 2: return lean_smem_wide_s2(); case 3: return lean_smem_wide_s3(); default: return lean_smem_wide_s4(); }
}
int lean_smem_bytes(int s, bool linear_visc) {
    switch (s) { case 1: return lean_smem_s1(linear_visc); case 2: return lean_smem_s2(linear_visc);
                 case 3: return lean_smem_s3(linear_visc); default: return lean_smem_s4(linear_visc); }
}
}  // namespace cudns
