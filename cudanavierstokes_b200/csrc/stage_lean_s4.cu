// stage kernel instantiations for stencilSize = 4 (see stage_lean.inc)
#define LEAN_TY8 CUDNS_LEAN_TY_LINEAR
#define LEAN_TY9 lean_ty_general(4)
#include "stage_lean.inc"
namespace cudns {
void launch_lean_s4(const KConst &kc, const StagePtrs &p, const StageCoef &c, const LeanMaps &maps, bool gen, bool wide, cudaStream_t st) {
    using namespace lean;
    switch (kc.v) {
        case 1: launch_v<4, 1>(kc, p, c, maps, gen, wide, st); break;
        case 2: launch_v<4, 2>(kc, p, c, maps, gen, wide, st); break;
        case 3: launch_v<4, 3>(kc, p, c, maps, gen, wide, st); break;
        case 4: launch_v<4, 4>(kc, p, c, maps, gen, wide, st); break;
        default: break;
    }
}
int lean_smem_wide_s4() { return (int)lean::Cfg<4, 16, 8>::bytes; }
int lean_smem_s4(bool linear_visc) {
    return (int)(linear_visc ? lean::Cfg<4, CUDNS_LEAN_TY_LINEAR, 8>::bytes : lean::Cfg<4, lean_ty_general(4), 9>::bytes);
}
}  // namespace cudns
