/* TEST INFRASTRUCTURE ONLY (oracle/).  Plain-C restatement of the reference's
 * per-frame NL-Kalman filter / RTS smoother path (reference src/nlkalman.c).
 * It exists to check the CUDA path and is validated itself against the
 * unmodified reference build (oracle/_ref/libnlkalman_ref.so, OMP_NUM_THREADS=1).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it;
 * nothing under bwd_nlkalman_b200/ includes, links or calls it.
 */
#ifndef ORACLE_NLK_PORT_H
#define ORACLE_NLK_PORT_H

#ifdef __cplusplus
extern "C" {
#endif

/* same field order and types as struct nlkalman_params (reference src/nlkalman.h:22-37) */
typedef struct {
    int patch_sz, search_sz_x, search_sz_t;
    int npatches_x, npatches_t, npatches_tagg;
    float dista_lambda, beta_x, beta_t;
} port_params;

enum { PORT_FLT1 = 0, PORT_FLT2 = 1, PORT_SMO1 = 2 }; /* reference src/nlkalman.h:40 */
enum { PORT_PASS_FILTER = 0, PORT_PASS_SMOOTH = 1 };

/* Optional per-stage dump, all arrays caller-allocated (NULL = not wanted).
 * G = number of grid patches = gw*gh, gw = (w-psz)/step+1, gh = (h-psz)/step+1,
 * kmax = max(npatches_x, npatches_t). */
typedef struct {
    int kmax;
    int *nk;            /* [G]        candidates kept after the sort (0 = no search) */
    int *np0;           /* [G]        kept candidates with a valid previous patch */
    int *knn_xy;        /* [G*kmax*2] (qx,qy) of kept candidates, sorted order */
    float *knn_d;       /* [G*kmax]   their distances */
    unsigned char *prev_p; /* [G]     validity of the previous-frame patch at p */
    unsigned char *active; /* [G]     1 = processed (not skipped by the mask) */
    float *vp;          /* [G]        posterior variance sum of processed groups */
} port_dump;

void port_rgb2opp(float *im, int w, int h, int ch);
void port_opp2rgb(float *im, int w, int h, int ch);
void port_warp_bicubic(float *imw, const float *im, const float *of, const float *msk,
                       int w, int h, int ch);
void port_default_params(port_params *p, float sigma, int mode);
void port_window(float *w2, int psz);
void port_dct2(float *tiles, int psz, int n, int inverse);

/* mode = PORT_PASS_FILTER: nlkalman_filter_frame (reference src/nlkalman.c:518-951)
 * mode = PORT_PASS_SMOOTH: nlkalman_smooth_frame (reference src/nlkalman.c:1409-1865)
 * out: w*h*ch, in1: noisy (filter) / filtered (smoother) frame, prev0: warped previous
 * estimate with NaN = invalid (may be NULL), bsic1: basic estimate (may be NULL). */
void port_pass(int mode, float *out, const float *in1, const float *prev0, const float *bsic1,
               int w, int h, int ch, float sigma, port_params prms, port_dump *dump);

void port_filter_frame(float *deno1, const float *nisy1, const float *deno0, const float *bsic1,
                       int w, int h, int ch, float sigma, port_params prms);
void port_smooth_frame(float *smoo1, const float *filt1, const float *smoo0, const float *bsic1,
                       int w, int h, int ch, float sigma, port_params prms);

#ifdef __cplusplus
}
#endif
#endif
