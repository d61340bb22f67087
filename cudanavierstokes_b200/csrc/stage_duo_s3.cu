// fifth-generation stage kernel, instantiations for stencilSize = 3 (see stage_duo.inc)
#include "stage_duo.inc"
namespace cudns {
void launch_duo_s3(const KConst &kc, const StagePtrs &p, const StageCoef &c, const DuoMaps &maps, cudaStream_t st) {
    switch (kc.v) {
        case 1: duo::launch_t<3, 1>(kc, p, c, maps, st); break;
        case 2: duo::launch_t<3, 2>(kc, p, c, maps, st); break;
        case 3: duo::launch_t<3, 3>(kc, p, c, maps, st); break;
        default: break;
    }
}
int duo_smem_s3() { return (int)duo::DCfg<3>::bytes; }
}  // namespace cudns
