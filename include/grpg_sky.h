/* grpg_sky.h -- C-ABI of the sky cube-map lookup (first half).
 *
 * Replaces SkyCubeMap.forward (/root/reference/lib/models/sky_cubemap.py:77-124): per-pixel ray directions
 * (lib/utils/graphics_utils.py:186-207, get_rays_torch), bilinear lookup in a [6, res, res, 3] cube map -- the reference
 * calls nvdiffrast's `dr.texture(cube[None], rays_d[None], filter_mode='linear', boundary_mode='cube')` -- the
 * foreground mask ((1 - acc) > 1e-3, or a sky mask in training), clamp to [0, 1], CHW output; and its backward into
 * the cube-map texels.  grpg_sky_compose_rgb8 fuses the lookup in front of the post-render epilogue of grpg_image.h:
 * rgb + sky * (1 - acc), clamp, x255, uint8, CHW -> HWC in ONE kernel, the sky colour never touching memory.
 *
 * nvdiffrast is a third-party dependency that is NOT vendored in the reference tree (it is imported at
 * sky_cubemap.py:7 from the user's environment, unpinned in requirements.txt), so the lookup restates the published
 * cube-mapping it implements: major-axis face selection with the OpenGL face order and orientation
 * (+x,-x,+y,-y,+z,-z; u,v from the two minor coordinates divided by 2|major| + 0.5), texel centres at (i + 0.5) / res,
 * bilinear weights, and seamless filtering across face edges (a tap that falls off a face is taken from the texel of
 * the adjacent face that shares the edge; at a cube corner the missing tap is dropped and the other three are
 * renormalised).  PARITY UNPINNED: no nvdiffrast build exists in this image to produce golden vectors.
 */
#ifndef GRPG_SKY_H
#define GRPG_SKY_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct grpg_sky_args {
    int height, width;
    int resolution;            /* cube-map edge length in texels                                                     */
    const float* cubemap;      /* [6, res, res, 3] device                                                            */
    const float* ray_matrix;   /* [9] device, row-major M = R^T K^-1: ray direction of pixel (x, y) = M (x+jx, y+jy, 1) */
    const float* jitter;       /* [2, H, W] device sub-pixel offsets (training: torch.rand), or NULL = 0.5            */
    const uint8_t* mask;       /* [H, W] device, non-zero = look the sky up here; NULL = derive from acc or everywhere */
    const float* acc;          /* [1, H, W] device: when mask is NULL and acc is not, mask = (1 - acc) > 1e-3         */
    float fill;                /* colour of pixels outside the mask (0, or 1 with a white background)                */
    float* sky;                /* out [3, H, W] device, clamped to [0, 1]                                            */
    void* stream;
} grpg_sky_args;

int grpg_sky_forward(const grpg_sky_args* a);

/* d_cubemap [6, res, res, 3] += d(sky)/d(texel) * dL_dsky; dL_dsky [3, H, W]; `sky` = the forward's output (the clamp
 * passes the gradient where 0 <= value <= 1, like torch.clamp).  d_cubemap must be zero-initialised by the caller. */
int grpg_sky_backward(const grpg_sky_args* a, const float* dL_dsky, float* d_cubemap);

/* lookup + composite + clamp + uint8 HWC (and optionally the float image) in one kernel: see grpg_image.h for the
 * output conventions.  rgb [3,H,W], acc [1,H,W] = the rasterizer's outputs. */
int grpg_sky_compose_rgb8(const grpg_sky_args* a, const float* rgb, uint8_t* out_rgb8, float* out_rgb);

#ifdef __cplusplus
}
#endif
#endif /* GRPG_SKY_H */
