// bas_render_tiled_kernel<8, *, *>: see render_tiled.cuh
#include "render_tiled.cuh"
namespace bas_render_detail {
BAS_INSTANTIATE_TILED(8)
}
