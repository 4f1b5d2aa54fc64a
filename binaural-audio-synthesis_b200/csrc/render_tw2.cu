// bas_render_tiled_kernel<2, *, *, *>: see render_tiled.cuh.  Two-warp CTAs, six per SM: the same 12 warps per SM in
// finer grains, so that the last wave of a launch whose tile count is a small multiple of the resident CTAs (one
// 60 s source: 1.46 waves) leaves fewer schedulers idle.
#include "render_tiled.cuh"
namespace bas_render_detail {
static const TiledShape kShapes[] = {
    BAS_TILED_SHAPE(2, 1, 6),
};
const TiledShape* tiled_shapes_tw2(int* count) {
    *count = (int)(sizeof(kShapes) / sizeof(kShapes[0]));
    return kShapes;
}
}
