/* nlkalman.h -- drop-in C interface of the NL-Kalman per-frame step.
 *
 * This header declares exactly the entry points, the parameter struct and the
 * enum of the reference's src/nlkalman.h (cited per declaration), so that the
 * reference's own drivers (src/main-flt.c:340-386, src/main-smo.c:198-212,
 * src/main-seq.c:462-583) compile and link against libnlkalman_b200.so
 * unchanged.  The implementation behind it is CUDA for sm_100a; all pointers are
 * HOST pointers to float32 images in interleaved HWC layout x[(i + j*w)*ch + l]
 * (reference lib/iio/iio.h:36-37).  Calls are synchronous: results are in host
 * memory on return.  There is no CPU fallback: without a CUDA device the
 * functions print a message to stderr and exit(1), the reference's own failure
 * convention (reference src/nlkalman.c:165-177).
 *
 * The reference selects its variant with four compile-time switches
 * (src/nlkalman.h:2,5,8,11).  This library implements the configuration the
 * reference ships and builds: DECOUPLE_FILTER2, WEIGHTED_AGGREGATION and
 * K_SIMILAR_PATCHES defined, LAMBDA_DISTANCE not defined.  The macros are kept
 * because K_SIMILAR_PATCHES decides the struct layout, i.e. the ABI.
 */
#ifndef NLKALMAN_H_B200
#define NLKALMAN_H_B200

#define DECOUPLE_FILTER2      /* reference src/nlkalman.h:2  */
/* #define LAMBDA_DISTANCE */ /* reference src/nlkalman.h:5 (off) */
#define WEIGHTED_AGGREGATION  /* reference src/nlkalman.h:8  */
#define K_SIMILAR_PATCHES     /* reference src/nlkalman.h:11 */

#ifdef __cplusplus
extern "C" {
#endif

/* orthonormal RGB <-> opponent transform, in place; no-op unless ch == 3
 * (replaces reference src/nlkalman.h:14-15, src/nlkalman.c:92-130) */
void rgb2opp(float *im, int w, int h, int ch);
void opp2rgb(float *im, int w, int h, int ch);

/* imw(x,y,:) = bicubic(im; x + of(x,y,0), y + of(x,y,1)); samples outside the
 * image and pixels with msk != 0 give NaN; msk may be NULL
 * (replaces reference src/nlkalman.h:18, src/nlkalman.c:71-88) */
void warp_bicubic(float *imw, float *im, float *of, float *msk, int w, int h, int ch);

/* parameter structure (reference src/nlkalman.h:22-37); passed BY VALUE */
struct nlkalman_params {
    int patch_sz;       /* patch size */
    int search_sz_x;    /* search window radius, spatial filtering */
    int search_sz_t;    /* search window radius, temporal filtering */
    int npatches_x;     /* number of similar patches, spatial filtering */
    int npatches_t;     /* number of similar patches, temporal filtering */
    int npatches_tagg;  /* size of the jointly filtered group */
    float dista_lambda; /* parsed but unused, as in the reference (src/nlkalman.c:642) */
    float beta_x;       /* noise multiplier, spatial filtering */
    float beta_t;       /* noise multiplier, Kalman filtering / smoothing */
};

/* reference src/nlkalman.h:40 */
enum FILTER_MODE { FLT1, FLT2, SMO1 };

/* fills every field that is < 0 with its sigma-dependent default
 * (replaces reference src/nlkalman.h:42, src/nlkalman.c:426-487) */
void nlkalman_default_params(struct nlkalman_params *p, float sigma, enum FILTER_MODE mode);

/* NL-Kalman filtering of one frame (replaces reference src/nlkalman.h:46,
 * src/nlkalman.c:518-951).  deno1: output; nisy1: noisy frame; deno0: previous
 * output warped to this frame, NaN = invalid, may be NULL; bsic1: basic estimate
 * of this frame, may be NULL; frame is unused, as in the reference. */
void nlkalman_filter_frame(float *deno1, float *nisy1, float *deno0, float *bsic1,
                           int w, int h, int ch, float sigma,
                           const struct nlkalman_params prms, int frame);

/* NL-Kalman (RTS) smoothing of one frame (replaces reference src/nlkalman.h:51,
 * src/nlkalman.c:1409-1865).  smoo1: output; filt1: filtered frame t; smoo0:
 * smoothed frame t+1 warped to t, NaN = invalid, may be NULL; bsic1 may be NULL. */
void nlkalman_smooth_frame(float *smoo1, float *filt1, float *smoo0, float *bsic1,
                           int w, int h, int ch, float sigma,
                           const struct nlkalman_params prms, int frame);

#ifdef __cplusplus
}
#endif
#endif /* NLKALMAN_H_B200 */
