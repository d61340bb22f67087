// stage kernel instantiations for stencilSize = 2 (see stage_lean.inc)
#define LEAN_TY8 CUDNS_LEAN_TY_LINEAR
#define LEAN_TY9 lean_ty_general(2)
#include "stage_lean.inc"
namespace cudns {
void launch_lean_s2(const KConst &kc, const StagePtrs &p, const StageCoef &c, const LeanMaps &maps, bool gen, bool wide, cudaStream_t st) {
    using namespace lean;
    switch (kc.v) {
        case 1: launch_v<2, 1>(kc, p, c, maps, gen, wide, st); break;
        case 2: launch_v<2, 2>(kc, p, c, maps, gen, wide, st); break;
        default: break;
    }
}
int lean_smem_wide_s2() { return (int)lean::Cfg<2, 16, 8>::bytes; }
int lean_smem_s2(bool linear_visc) {
    return (int)(linear_visc ? lean::Cfg<2, CUDNS_LEAN_TY_LINEAR, 8>::bytes : lean::Cfg<2, lean_ty_general(2), 9>::bytes);
}
}  // namespace cudns
