// fifth-generation stage kernel, instantiations for stencilSize = 2 (see stage_duo.inc)
#include "stage_duo.inc"
namespace cudns {
void launch_duo_s2(const KConst &kc, const StagePtrs &p, const StageCoef &c, const DuoMaps &maps, cudaStream_t st) {
    switch (kc.v) {
        case 1: duo::launch_t<2, 1>(kc, p, c, maps, st); break;
        case 2: duo::launch_t<2, 2>(kc, p, c, maps, st); break;
        default: break;
    }
}
int duo_smem_s2() { return (int)duo::DCfg<2>::bytes; }
}  // namespace cudns
