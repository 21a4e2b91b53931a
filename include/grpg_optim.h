/* grpg_optim.h -- C-ABI of the fused per-iteration parameter update.
 *
 * (1) grpg_adam_step: ONE launch that applies torch.optim.Adam's update to every parameter tensor of every
 *     sub-model.  The reference builds one `torch.optim.Adam(l, lr=0.0, eps=1e-15)` per GaussianModel with seven
 *     parameter groups (/root/reference/lib/models/gaussian_model.py:286-318) and steps each of them every
 *     iteration (street_gaussian_model.py:536-541, train.py:305-307): with a background and 8 actors that is 9
 *     optimisers x 6 non-empty tensors x 7 foreach passes.  Semantics are torch's non-capturable, non-amsgrad,
 *     weight_decay = 0 path, operation for operation in float32:
 *         m <- lerp(m, g, 1 - beta1);  v <- v * beta2;  v <- v + (1 - beta2) * g * g;
 *         p <- p + step_size * (m / (sqrt(v) / bias_correction2_sqrt + eps))
 *     with step_size = -lr / (1 - beta1^t) and bias_correction2_sqrt = sqrt(1 - beta2^t) evaluated by the caller in
 *     double, exactly as torch/optim/adam.py does in Python.  Tensors without a gradient are simply not listed
 *     (torch skips them; their moments do not decay).
 * (2) grpg_densify_stats: ONE launch for `set_max_radii2D` + `add_densification_stats`
 *     (street_gaussian_model.py:555-578; called every iteration before densify_until_iter, train.py:277-281):
 *     for every visible Gaussian (radii > 0) of every sub-model
 *         max_radii2D = max(max_radii2D, radii);  xyz_gradient_accum[:,0] += |grad[:2]|;
 *         xyz_gradient_accum[:,1] += |grad[2:]|;  denom += 1
 *     -- the reference does this with boolean-mask indexing per sub-model (4 masked read-modify-writes, each a
 *     nonzero + host synchronisation).
 */
#ifndef GRPG_OPTIM_H
#define GRPG_OPTIM_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct grpg_adam_tensor {
    float* param;            /* device, updated in place                                                           */
    const float* grad;       /* device                                                                             */
    float* exp_avg;          /* device, updated in place                                                           */
    float* exp_avg_sq;       /* device, updated in place                                                           */
    long long numel;
    float one_minus_beta1;   /* (float)(1 - beta1)                                                                 */
    float beta2;             /* (float)beta2                                                                       */
    float one_minus_beta2;   /* (float)(1 - beta2)                                                                 */
    float step_size;         /* (float)(-lr / (1 - beta1^t))                                                       */
    float bias_correction2_sqrt; /* (float)sqrt(1 - beta2^t)                                                       */
    float eps;
} grpg_adam_tensor;

/* bytes of device scratch for a table of n tensors (n <= 1024 per call) */
size_t grpg_adam_workspace_bytes(int n);
/* `tensors` is a HOST array of n entries, copied into `workspace` on `stream`.  Returns 0 or sets grpg_last_error(). */
int grpg_adam_step(const grpg_adam_tensor* tensors, int n, void* workspace, void* stream);

typedef struct grpg_stats_submodel {
    int n;                        /* Gaussians of this sub-model; rows [offset, offset+n) of the composed arrays   */
    int reserved;
    float* max_radii2D;           /* [n]    device, in place                                                       */
    float* xyz_gradient_accum;    /* [n,2]  device, in place                                                       */
    float* denom;                 /* [n,1]  device, in place                                                       */
} grpg_stats_submodel;

size_t grpg_stats_workspace_bytes(int n_sub);
/* radii [P] int32, viewspace_grad [P,3] float32 (device) with P = sum of subs[i].n, sub-models in array order. */
int grpg_densify_stats(const grpg_stats_submodel* subs, int n_sub, const int* radii, const float* viewspace_grad,
                       void* workspace, void* stream);


/* The two halves on their own and with the caller's mask, for callers that keep the reference's two-call protocol
 * (`set_max_radii2D(radii, visibility_filter)` then `add_densification_stats(viewspace_point_tensor,
 * visibility_filter)`, street_gaussian_model.py:555-578) or pass a filter other than radii > 0:
 * `what` = GRPG_STATS_MAX_RADII and/or GRPG_STATS_GRADIENTS; `visibility_filter` [P] bytes (0 / non-zero, a torch
 * bool tensor) or NULL for radii > 0.  `radii` may be NULL when only GRPG_STATS_GRADIENTS is asked for with a filter,
 * `viewspace_grad` when only GRPG_STATS_MAX_RADII is. */
#define GRPG_STATS_MAX_RADII 1
#define GRPG_STATS_GRADIENTS 2
int grpg_densify_stats_ex(const grpg_stats_submodel* subs, int n_sub, const int* radii, const float* viewspace_grad,
                          const unsigned char* visibility_filter, int what, void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GRPG_OPTIM_H */
