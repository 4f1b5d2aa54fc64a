// bas_render_tiled_kernel<6, *, *, *>: see render_tiled.cuh
#include "render_tiled.cuh"
namespace bas_render_detail {
static const TiledShape kShapes[] = {
    BAS_TILED_SHAPE(6, 1, 2),
    BAS_TILED_SHAPE(6, 2, 1),
};
const TiledShape* tiled_shapes_tw6(int* count) {
    *count = (int)(sizeof(kShapes) / sizeof(kShapes[0]));
    return kShapes;
}
}
