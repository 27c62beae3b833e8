/* nlkalman-flt -- frame-by-frame NL-Kalman filtering on a B200, host driver.
 *
 * Command line, file conventions, mode rules, messages and exit codes follow the
 * reference driver (reference src/main-flt.c:21-400) so that the pipeline scripts run
 * unchanged; the numerics run on the GPU through the C ABI of libnlkalman_b200.so.
 * All frames of one invocation stay in HBM between the stages (colour transform, warp,
 * first filtering, second filtering); only the inputs go up and the outputs come back.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "nlk_image_io.h"
#include "nlk_opts.h"
#include "nlkalman_b200.h"

/* The CUDA runtime, the context and the kernel module take a few hundred milliseconds to come up: more
 * than everything else a per-frame invocation does.  A helper thread brings them up (a throw-away 8x8
 * context) while the main thread parses and decodes the input files. */
static void *gpu_warmup(void *arg)
{
    nlk_ctx *t = nlk_ctx_create(8, 8, 1, *(int *)arg);
    if (t) nlk_ctx_destroy(t);
    return NULL;
}

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void auto_params(struct nlkalman_params *p)
{
    /* -1 means automatic value (reference src/main-flt.c:39-52) */
    p->patch_sz = p->search_sz_x = p->search_sz_t = -1;
    p->npatches_x = p->npatches_t = p->npatches_tagg = -1;
    p->dista_lambda = p->beta_x = p->beta_t = -1.f;
}

static void print_params(const char *title, const struct nlkalman_params *p)
{
    printf("%s\n", title);
    printf("\tpatch      %d\n", p->patch_sz);
    printf("\tsearch_x   %d\n", p->search_sz_x);
    printf("\tsearch_t   %d\n", p->search_sz_t);
    printf("\tnp_x       %d\n", p->npatches_x);
    printf("\tnp_t       %d\n", p->npatches_t);
    printf("\tnp_tagg    %d\n", p->npatches_tagg);
    printf("\tlambda     %g\n", p->dista_lambda);
    printf("\tbeta_x     %g\n", p->beta_x);
    printf("\tbeta_t     %g\n", p->beta_t);
    printf("\n");
}

static int gpu_fail(const char *what)
{
    fprintf(stderr, "nlkalman-flt: %s: %s\n", what, nlk_last_error());
    return 1;
}

int main(int argc, const char *argv[])
{
    const char *noisy_path = NULL, *bflow_path = NULL, *boccl_path = NULL;
    const char *flt10_path = NULL, *flt20_path = NULL, *flt11_path = NULL, *flt21_path = NULL;
    float sigma = 0.f;
    int verbose = 0;
    struct nlkalman_params f1, f2;
    auto_params(&f1);
    auto_params(&f2);

    const struct nlk_opt options[] = {
        {NLK_OPT_GROUP, 0, "Data i/o options", NULL, NULL},
        {NLK_OPT_STRING, 'i', "nisy", &noisy_path, "input noisy frames path"},
        {NLK_OPT_STRING, 'o', "bflo", &bflow_path, "input bwd flow path"},
        {NLK_OPT_STRING, 'k', "bocc", &boccl_path, "input bwd occlusion masks path"},
        {NLK_OPT_STRING, 0, "flt10", &flt10_path, "input previous first filtering path"},
        {NLK_OPT_STRING, 0, "flt20", &flt20_path, "input previous second filtering path"},
        {NLK_OPT_STRING, 0, "flt11", &flt11_path, "input/output first filtering path"},
        {NLK_OPT_STRING, 0, "flt21", &flt21_path, "output second filtering path"},
        {NLK_OPT_FLOAT, 's', "sigma", &sigma, "noise standard dev"},
        {NLK_OPT_GROUP, 0, "First filtering options", NULL, NULL},
        {NLK_OPT_INT, 0, "f1_p", &f1.patch_sz, "patch size"},
        {NLK_OPT_INT, 0, "f1_sx", &f1.search_sz_x, "search radius (spatial filtering)"},
        {NLK_OPT_INT, 0, "f1_st", &f1.search_sz_t, "search radius (temporal filtering)"},
        {NLK_OPT_INT, 0, "f1_nx", &f1.npatches_x, "number of similar patches spatial"},
        {NLK_OPT_INT, 0, "f1_nt", &f1.npatches_t, "number of similar patches kalman"},
        {NLK_OPT_INT, 0, "f1_nt_agg", &f1.npatches_tagg, "number of similar patches kalman spatial average"},
        {NLK_OPT_FLOAT, 0, "f1_bx", &f1.beta_x, "noise multiplier in spatial filtering"},
        {NLK_OPT_FLOAT, 0, "f1_bt", &f1.beta_t, "noise multiplier in kalman filtering"},
        {NLK_OPT_FLOAT, 0, "f1_l", &f1.dista_lambda, "noisy patch weight in patch distance"},
        {NLK_OPT_GROUP, 0, "Second filtering options", NULL, NULL},
        {NLK_OPT_INT, 0, "f2_p", &f2.patch_sz, "patch size"},
        {NLK_OPT_INT, 0, "f2_sx", &f2.search_sz_x, "search radius (spatial filtering)"},
        {NLK_OPT_INT, 0, "f2_st", &f2.search_sz_t, "search radius (temporal filtering)"},
        {NLK_OPT_INT, 0, "f2_nx", &f2.npatches_x, "number of similar patches spatial"},
        {NLK_OPT_INT, 0, "f2_nt", &f2.npatches_t, "number of similar patches kalman"},
        {NLK_OPT_INT, 0, "f2_nt_agg", &f2.npatches_tagg, "number of similar patches kalman spatial average"},
        {NLK_OPT_FLOAT, 0, "f2_bx", &f2.beta_x, "noise multiplier in spatial filtering"},
        {NLK_OPT_FLOAT, 0, "f2_bt", &f2.beta_t, "noise multiplier in kalman filtering"},
        {NLK_OPT_FLOAT, 0, "f2_l", &f2.dista_lambda, "noisy patch weight in patch distance"},
        {NLK_OPT_GROUP, 0, "Program options", NULL, NULL},
        {NLK_OPT_INT, 'v', "verbose", &verbose, "verbose output"},
        {NLK_OPT_END, 0, NULL, NULL, NULL},
    };
    nlk_opts_parse(options, "nlkalman-flt [options] [[--] args]",
                   "\nPatch-based Kalman filter for video denoising.", argc, argv);

    /* mode (reference src/main-flt.c:129-149) */
    const int apply_filt1 = f1.patch_sz != 0;
    const int apply_filt2 = f2.patch_sz != 0 && flt21_path;
    if (!apply_filt1 && !apply_filt2) return fprintf(stderr, "Error: nothing to do, exiting\n"), 1;
    if (!apply_filt1 && !flt11_path)
        return fprintf(stderr, "Error: f1_p == 0 and no input path given, exiting\n"), 1;
    if (!flt11_path && !apply_filt2)
        return fprintf(stderr, "Error: no output path given for any computed output - exiting\n"), 1;
    if (!flt11_path && !flt21_path)
        return fprintf(stderr, "Error: s1_p == 0 and no output paths given for filt1 and filt2\n"), 1;
    if (f2.patch_sz == 0 && flt21_path)
        fprintf(stderr, "Warning: f2_p == 0 - no output files will be stored in %s\n", flt21_path);

    nlkalman_default_params(&f1, sigma, FLT1);
    nlkalman_default_params(&f2, sigma, FLT2);

    if (verbose) {
        printf("data input:\n");
        printf("\tnoise         %05.2f\n", sigma);
        printf("\tnoisy frames  %s\n", noisy_path);
        printf("\tbwd flows     %s\n", bflow_path);
        printf("\tbwd occlus.   %s\n", boccl_path);
        printf("\tprev filt 1   %s\n", flt10_path);
        printf("\tprev filt 2   %s\n", flt20_path);
        if (!apply_filt1) printf("\tfiltering 1   %s\n", flt11_path);
        printf("\n");
        printf("data output:\n");
        if (apply_filt1) printf("\tfiltering 1   %s\n", flt11_path);
        printf("\tfiltering 2   %s\n", flt21_path);
        printf("\n");
        if (apply_filt1) print_params("first filtering parameters:", &f1);
        if (apply_filt2) print_params("second filtering parameters:", &f2);
    }

    /* bring the GPU up beside the file decoding */
    const int dev_pick = nlk_pick_device();
    int dev = dev_pick;
    const int timing = getenv("NLK_CLI_TIMING") != NULL;
    const double t_start = now_s();
    pthread_t warm;
    const char *warm_env = getenv("NLK_CLI_WARMUP");      /* 0: bring the GPU up after the files, in this thread */
    const int warm_on = !(warm_env && atoi(warm_env) == 0) && pthread_create(&warm, NULL, gpu_warmup, &dev) == 0;

    /* load data (reference src/main-flt.c:215-332: same checks, same messages) */
    int w, h, c, w1, h1, c1;
    float *nisy = noisy_path ? nlk_read_image(noisy_path, &w, &h, &c) : NULL;
    if (!nisy) return fprintf(stderr, "Error while openning bwd optical flow\n"), 1;
    float *bflo = NULL, *bocc = NULL, *flt10 = NULL, *flt20 = NULL, *flt11 = NULL;
    if (bflow_path) {
        bflo = nlk_read_image(bflow_path, &w1, &h1, &c1);
        if (!bflo) return fprintf(stderr, "Error while openning bwd optical flow\n"), 1;
        if (w * h != w1 * h1 || c1 != 2) return fprintf(stderr, "Frame and optical flow size missmatch\n"), 1;
    }
    if (bflow_path && boccl_path) {
        bocc = nlk_read_image(boccl_path, &w1, &h1, &c1);
        if (!bocc) return fprintf(stderr, "Error while openning occlusion mask\n"), 1;
        if (w * h != w1 * h1 || c1 != 1) return fprintf(stderr, "Frame and occlusion mask size missmatch\n"), 1;
    }
    if (flt10_path) {
        flt10 = nlk_read_image(flt10_path, &w1, &h1, &c1);
        if (!flt10) fprintf(stderr, "Error while openning previous filter 1 output\n");
        if (flt10 && w * h * c != w1 * h1 * c1)
            return fprintf(stderr, "Frame and previous filter 1 output size missmatch\n"), 1;
    }
    if (flt20_path) {
        flt20 = nlk_read_image(flt20_path, &w1, &h1, &c1);
        if (!flt20) fprintf(stderr, "Error while openning previous filter 2 output\n");
        if (flt20 && w * h * c != w1 * h1 * c1)
            return fprintf(stderr, "Frame and previous filter 2 output size missmatch\n"), 1;
    }
    if (!apply_filt1) {
        flt11 = nlk_read_image(flt11_path, &w1, &h1, &c1);
        if (!flt11) return fprintf(stderr, "Error while openning filter 1 output\n"), 1;
        if (w * h * c != w1 * h1 * c1) return fprintf(stderr, "Frame and filter 1 output size missmatch\n"), 1;
    }

    /* run on the GPU (reference src/main-flt.c:335-388) */
    const double t_read = now_s();
    if (warm_on) pthread_join(warm, NULL);
    const double t_warm = now_s();
    nlk_ctx *ctx = nlk_ctx_create(w, h, c, dev);
    if (!ctx) return gpu_fail("no usable CUDA device (there is no CPU fallback)");
    const size_t ib = (size_t)w * h * c * sizeof(float), npix = (size_t)w * h;
    float *d_nisy = nlk_dev_alloc(ctx, ib), *d_warp = nlk_dev_alloc(ctx, ib), *d_tmp = nlk_dev_alloc(ctx, ib);
    float *d_flt11 = nlk_dev_alloc(ctx, ib), *d_flt21 = nlk_dev_alloc(ctx, ib);
    float *d_of = bflo ? nlk_dev_alloc(ctx, npix * 2 * sizeof(float)) : NULL;
    float *d_occ = bocc ? nlk_dev_alloc(ctx, npix * sizeof(float)) : NULL;
    if (!d_nisy || !d_warp || !d_tmp || !d_flt11 || !d_flt21 || (bflo && !d_of) || (bocc && !d_occ))
        return gpu_fail("device memory");
    int rc = nlk_upload(ctx, d_nisy, nisy, ib);
    if (!rc) rc = nlk_rgb2opp_dev(ctx, d_nisy, d_nisy);
    if (!rc && bflo) rc = nlk_upload(ctx, d_of, bflo, npix * 2 * sizeof(float));
    if (!rc && bocc) rc = nlk_upload(ctx, d_occ, bocc, npix * sizeof(float));
    if (rc) return gpu_fail("upload");

    /* previous frame of one stage -> opponent space -> warped by the backward flow */
    const float *d_prev;
#define NLK_PREVIOUS(host_img)                                                              \
    do {                                                                                    \
        d_prev = NULL;                                                                      \
        if (host_img) {                                                                     \
            rc = nlk_upload(ctx, d_tmp, host_img, ib);                                      \
            if (!rc) rc = nlk_rgb2opp_dev(ctx, d_tmp, d_tmp);                               \
            d_prev = d_tmp;                                                                 \
            if (!rc && bflo) { rc = nlk_warp_dev(ctx, d_warp, d_tmp, d_of, d_occ); d_prev = d_warp; } \
            if (rc) return gpu_fail("previous frame");                                      \
        }                                                                                   \
    } while (0)

    float *out = (float *)malloc(ib);
    if (!out) return fprintf(stderr, "out of memory\n"), 1;
    if (apply_filt1) {
        NLK_PREVIOUS(flt10);
        if (nlk_pass_dev(ctx, 0, d_flt11, d_nisy, d_prev, NULL, sigma, f1)) return gpu_fail("first filtering");
    } else {
        if (nlk_upload(ctx, d_flt11, flt11, ib) || nlk_rgb2opp_dev(ctx, d_flt11, d_flt11)) return gpu_fail("upload");
    }
    if (apply_filt2) {
        NLK_PREVIOUS(flt20);
        if (nlk_pass_dev(ctx, 0, d_flt21, d_nisy, d_prev, d_flt11, sigma, f2)) return gpu_fail("second filtering");
        if (flt11_path) { /* (sic) reference src/main-flt.c:376 */
            if (nlk_opp2rgb_dev(ctx, d_flt21, d_flt21) || nlk_download(ctx, out, d_flt21, ib) || nlk_ctx_sync(ctx))
                return gpu_fail("second filtering output");
            if (nlk_write_image(flt21_path, out, w, h, c)) return fprintf(stderr, "%s\n", nlk_io_error()), 1;
        }
    }
    if (apply_filt1 && flt11_path) {
        if (nlk_opp2rgb_dev(ctx, d_flt11, d_flt11) || nlk_download(ctx, out, d_flt11, ib) || nlk_ctx_sync(ctx))
            return gpu_fail("first filtering output");
        if (nlk_write_image(flt11_path, out, w, h, c)) return fprintf(stderr, "%s\n", nlk_io_error()), 1;
    }

    if (timing)
        fprintf(stderr, "nlkalman-flt timing: read+decode %.3f s (GPU bring-up beside it, +%.3f s waited), "
                        "GPU + output files %.3f s, total %.3f s\n",
                t_read - t_start, t_warm - t_read, now_s() - t_warm, now_s() - t_start);
    /* the outputs are on disk: leave without tearing the CUDA context down (tens of milliseconds
     * that nothing depends on) */
    fflush(NULL);
    _exit(EXIT_SUCCESS);
}
