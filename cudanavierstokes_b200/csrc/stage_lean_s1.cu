// stage kernel instantiations for stencilSize = 1 (see stage_lean.inc)
#define LEAN_TY8 CUDNS_LEAN_TY_LINEAR
#define LEAN_TY9 lean_ty_general(1)
#include "stage_lean.inc"
namespace cudns {
void launch_lean_s1(const KConst &kc, const StagePtrs &p, const StageCoef &c, const LeanMaps &maps, bool gen, bool wide, cudaStream_t st) {
    using namespace lean;
    switch (kc.v) {
        case 1: launch_v<1, 1>(kc, p, c, maps, gen, wide, st); break;
        default: break;
    }
}
int lean_smem_wide_s1() { return (int)lean::Cfg<1, 16, 8>::bytes; }
int lean_smem_s1(bool linear_visc) {
    return (int)(linear_visc ? lean::Cfg<1, CUDNS_LEAN_TY_LINEAR, 8>::bytes : lean::Cfg<1, lean_ty_general(1), 9>::bytes);
}
}  // namespace cudns
