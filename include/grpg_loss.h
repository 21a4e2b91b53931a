/* grpg_loss.h -- C-ABI of the fused L1 + SSIM image loss (forward value AND gradient in one pass).
 *
 * the per-iteration image loss of the reference's trainer,
 *   train.py:116-118            loss = (1 - lambda) * lambda_l1 * L1 + lambda * (1 - SSIM)
 *   lib/utils/loss_utils.py:21-37   l1_loss(network_output, gt, mask)
 *   lib/utils/loss_utils.py:81-124  gaussian / create_window / ssim / _ssim  (11x11 window, sigma 1.5,
 *                                   zero padding 5, C1 = 0.01^2, C2 = 0.03^2, mask zeroes both images)
 * which the reference evaluates as 5 grouped 11x11 convolutions + ~20 elementwise kernels forward and their
 * autograd transposes backward.  Here one kernel reads the two images once and writes the per-plane sums and
 * the gradient map; nothing else touches HBM.
 *
 * Plain C ABI: device pointers, sizes, a cudaStream_t; errors as in grpg_b200.h (non-zero + grpg_last_error()).
 */
#ifndef GRPG_LOSS_H
#define GRPG_LOSS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct grpg_l1_ssim_args {
    int planes;                /* B*C image planes of H x W floats, plane-major (the reference's [C,H,W] / [B,C,H,W]) */
    int planes_per_mask;       /* C: plane p uses mask image p / planes_per_mask                                      */
    int height, width;
    const float* img1;         /* [planes,H,W] rendered image (the differentiated argument)                           */
    const float* img2;         /* [planes,H,W] ground truth                                                            */
    const uint8_t* mask;       /* [planes/planes_per_mask,H,W] 0/1 bytes, or NULL = everything valid                   */
    /* gradient map written to `grad`:  coef_l1 * sign(img1-img2) * mask  +  coef_ssim * d(sum of ssim_map)/d img1.
     * The caller folds the loss weights and the two mean denominators into the coefficients
     * (L1: 1 / (#masked pixels * C), loss_utils.py:35;  SSIM: 1 / (planes*H*W), :121). */
    float coef_l1, coef_ssim;
    double* sums;              /* [planes][2] device, overwritten: (sum |img1-img2| over valid pixels, sum of ssim_map) */
    float* grad;               /* [planes,H,W] device, fully written; NULL = value only                                */
    void* stream;              /* cudaStream_t */
} grpg_l1_ssim_args;

int grpg_l1_ssim(const grpg_l1_ssim_args* a);

#ifdef __cplusplus
}
#endif
#endif /* GRPG_LOSS_H */
