// bas_render_tiled_kernel<4, *, *, *>: see render_tiled.cuh
#include "render_tiled.cuh"
namespace bas_render_detail {
static const TiledShape kShapes[] = {
    BAS_TILED_SHAPE(4, 2, 2),
    BAS_TILED_SHAPE(4, 1, 2),
    BAS_TILED_SHAPE(4, 1, 3),
};
const TiledShape* tiled_shapes_tw4(int* count) {
    *count = (int)(sizeof(kShapes) / sizeof(kShapes[0]));
    return kShapes;
}
}
