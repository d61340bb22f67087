// fifth-generation stage kernel, instantiations for stencilSize = 4 (see stage_duo.inc)
#include "stage_duo.inc"
namespace cudns {
void launch_duo_s4(const KConst &kc, const StagePtrs &p, const StageCoef &c, const DuoMaps &maps, cudaStream_t st) {
    switch (kc.v) {
        case 1: duo::launch_t<4, 1>(kc, p, c, maps, st); break;
        case 2: duo::launch_t<4, 2>(kc, p, c, maps, st); break;
        case 3: duo::launch_t<4, 3>(kc, p, c, maps, st); break;
        case 4: duo::launch_t<4, 4>(kc, p, c, maps, st); break;
        default: break;
    }
}
int duo_smem_s4() { return (int)duo::DCfg<4>::bytes; }
}  // namespace cudns
