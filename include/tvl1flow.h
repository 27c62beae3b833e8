/* tvl1flow.h -- drop-in C interface of the TV-L1 optical flow library of the reference
 * (reference lib/tvl1flow/tvl1flow_lib.c, which its program lib/tvl1flow/main.c:22 #includes as
 * source): the two entry points with the reference's own names, argument lists and meaning, so that
 * code written against that file links against libnlkalman_b200.so instead.  Behind them are the
 * sm_100a kernels of bwd_nlkalman_b200/csrc/nlk_tvl1.cuh through nlk_tvl1_level_host /
 * nlk_tvl1_flow_host (include/nlkalman_b200.h) on a context the library keeps for the caller.
 *
 * All pointers are HOST pointers to nx * ny floats; calls are synchronous.  As with the filter's
 * entry points (nlkalman.h) there is no CPU fallback: without a CUDA device, or for a request the
 * reference itself would abort on (a Gaussian window larger than the image, lib/tvl1flow/mask.c:232),
 * a message goes to stderr and the process exits with status 1.
 */
#ifndef TVL1FLOW_H_B200
#define TVL1FLOW_H_B200

#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

/* the flow at one scale: u1, u2 hold the initial flow on entry and the result on return;
 * verbose prints the iterations of every warping step on stderr
 * (replaces reference lib/tvl1flow/tvl1flow_lib.c:93-280) */
void Dual_TVL1_optic_flow(float *I0, float *I1, float *u1, float *u2, const int nx, const int ny,
                          const float tau, const float lambda, const float theta, const int warps,
                          const float epsilon, const bool verbose);

/* the whole estimator: normalisation, pre-smoothing, `nscales` scales by `zfactor`, the flow solved
 * from the coarsest scale down to `fscale` and upsampled from there; u1, u2 are output only
 * (replaces reference lib/tvl1flow/tvl1flow_lib.c:345-477) */
void Dual_TVL1_optic_flow_multiscale(float *I0, float *I1, float *u1, float *u2, const int nxx, const int nyy,
                                     const float tau, const float lambda, const float theta, const int nscales,
                                     const int fscale, const float zfactor, const int warps, const float epsilon,
                                     const bool verbose);

#ifdef __cplusplus
}
#endif
#endif
