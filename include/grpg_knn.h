/* grpg_knn.h -- C-ABI of the initialisation-time nearest-neighbour statistic (last row).
 *
 * Replaces `distCUDA2` of the reference's simple-knn submodule (/root/reference/submodules/simple-knn/spatial.cu:15-27,
 * simple_knn.cu:185-219 `SimpleKNN::knn`): for every point, the MEAN OF THE SQUARED DISTANCES TO ITS THREE NEAREST
 * OTHER POINTS -- what `create_from_pcd` turns into the initial scales (lib/models/gaussian_model.py:63,
 * gaussian_model_actor.py:144).  The reference finds them exactly (Morton order, 1024-point boxes, a rejection radius
 * from the six Morton neighbours, simple_knn.cu:140-183), so the result does not depend on the search structure: it is
 * the three smallest values of  fma(dz, dz, fma(dx, dx, dy * dy))  with d = other - point (the reference build's
 * contraction of `d.x*d.x + d.y*d.y + d.z*d.z`, read from its SASS), summed smallest-first and divided by 3.0f.
 * Fewer than four points leave FLT_MAX terms in the sum, as in the reference.
 *
 * B200 design: the same Morton order (hand-written radix sort of csrc/radix_sort.cuh, no thrust / CUB), points
 * gathered into that order once, and a THREE-level box hierarchy over it (128 / 1024 / 32768 points) instead of one
 * flat level, so a point tests ~10^2 boxes instead of P / 1024 and scans 128-point leaves instead of 1024-point boxes.
 * No host synchronisation (the reference blocks twice on the bounding box, simple_knn.cu:196-199).
 */
#ifndef GRPG_KNN_H
#define GRPG_KNN_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* bytes of device scratch for P points */
size_t grpg_knn_workspace_bytes(int P);
/* points [P,3] float32 device, mean_dist2 [P] float32 device (every element written). */
int grpg_knn_mean_dist2(int P, const float* points, float* mean_dist2, void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GRPG_KNN_H */
