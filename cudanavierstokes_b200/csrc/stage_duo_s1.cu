// fifth-generation stage kernel, instantiations for stencilSize = 1 (see stage_duo.inc)
#include "stage_duo.inc"
namespace cudns {
void launch_duo_s1(const KConst &kc, const StagePtrs &p, const StageCoef &c, const DuoMaps &maps, cudaStream_t st) {
    switch (kc.v) {
        case 1: duo::launch_t<1, 1>(kc, p, c, maps, st); break;
        default: break;
    }
}
int duo_smem_s1() { return (int)duo::DCfg<1>::bytes; }
}  // namespace cudns
