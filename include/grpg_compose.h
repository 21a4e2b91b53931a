/* grpg_compose.h -- C-ABI of the fused scene-graph compose.
 *
 * Replaces what `StreetGaussianModel` does between `parse_camera` and the rasterizer call on every render
 * (/root/reference/lib/models/street_gaussian_model.py):
 *   get_xyz       :341-367   clone, flip, quaternion_to_matrix, einsum + translate, cat
 *   get_rotation  :314-338   F.normalize, flip product, quaternion_raw_multiply, F.normalize, cat
 *   get_scaling   :296-312   exp, cat            get_opacity :438-453   sigmoid, cat
 *   get_features  :370-384   cat(dc, rest); actors: sum_f dc[:, f] * IDFT(t)[f]  (gaussian_model_actor.py:73-82)
 * -- about 40 small PyTorch kernels and ~0.4 GB of intermediate copies at 2 M Gaussians -- by ONE forward launch that
 * writes the five [P, .] rasterizer inputs, and its backward by one launch (+ a one-block finish for the pose).
 * Sub-models keep the reference's parameter layout (raw xyz, log-scales, raw quaternions, opacity logits, SH dc/rest),
 * so optimiser state, densification and checkpoints are untouched.
 *
 * Arithmetic is float32 in the reference's operation order where that is observable (exp, sigmoid, the two
 * normalisations with F.normalize's eps = 1e-12, the raw quaternion product); results agree with the PyTorch
 * formulation to float32 rounding (tests: <= 2e-6 relative on outputs, 1e-5 on gradients).
 */
#ifndef GRPG_COMPOSE_H
#define GRPG_COMPOSE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GRPG_COMPOSE_MAX_FOURIER 16

typedef struct grpg_compose_submodel {
    int n;                        /* Gaussians in this sub-model                                                  */
    int is_actor;                 /* 0 = background (no pose), 1 = actor: posed by obj_rot / obj_trans            */
    int fourier_dim;              /* rows of features_dc: 1 for the background, cfg fourier_dim for actors        */
    int reserved;
    /* parameters, device pointers, contiguous float32 */
    const float* xyz;             /* [n,3]                                                                        */
    const float* scaling;         /* [n,3]   log-scales                                                           */
    const float* rotation;        /* [n,4]   raw quaternions (r,x,y,z)                                            */
    const float* opacity;         /* [n,1]   logits                                                               */
    const float* features_dc;     /* [n,fourier_dim,3]                                                            */
    const float* features_rest;   /* [n,M-1,3]                                                                    */
    const unsigned char* flip_mask; /* [n] 0/1 or NULL (street_gaussian_model.py:284-293)                         */
    const float* obj_rot;         /* DEVICE, 4 floats: actor pose quaternion as parse_camera builds it (:270-275)  */
    const float* obj_trans;       /* DEVICE, 3 floats; both NULL for the background                               */
    float idft[GRPG_COMPOSE_MAX_FOURIER]; /* IDFT(time, fourier_dim)[0] (lib/utils/sh_utils.py:120-130)           */
    /* backward outputs (device, same shapes as the parameters); ignored by the forward */
    float* d_xyz; float* d_scaling; float* d_rotation; float* d_opacity; float* d_features_dc; float* d_features_rest;
} grpg_compose_submodel;

/* bytes of device scratch both calls need for `n_sub` sub-models (descriptor table + pose accumulators) */
size_t grpg_compose_workspace_bytes(int n_sub);

/* Outputs (device): xyz [P,3], rotation [P,4], scaling [P,3], opacity [P,1], features [P,M,3] with
 * P = sum of subs[i].n, sub-models concatenated in array order (at most 256).  `subs` is a HOST array; it is copied
 * into `workspace` on `stream` (the pose stays on the device: no host synchronisation).  Returns 0, or non-zero with grpg_last_error() set. */
int grpg_compose_forward(const grpg_compose_submodel* subs, int n_sub, int M, void* workspace, float* xyz,
                         float* rotation, float* scaling, float* opacity, float* features, void* stream);

/* Cotangents g_* (device, shapes of the forward outputs) -> subs[i].d_* and the pose gradients
 * d_obj_rots [n_sub,4], d_obj_trans [n_sub,3] (device; rows of background sub-models are zero). */
int grpg_compose_backward(const grpg_compose_submodel* subs, int n_sub, int M, void* workspace, const float* g_xyz,
                          const float* g_rotation, const float* g_scaling, const float* g_opacity,
                          const float* g_features, float* d_obj_rots, float* d_obj_trans, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GRPG_COMPOSE_H */
