/* tvl1flow -- Dual TV-L1 optical flow on the GPU, with the command line of the reference's program
 * (reference lib/tvl1flow/main.c:73-196, the flow estimator of scripts/nlkalman-seq.sh:60-65, :124-129):
 *
 *     tvl1flow I0 I1 [out nproc tau lambda theta nscales fscale zfactor nwarps epsilon verbose]
 *
 * Positional arguments, every one after I1 optional, an out-of-range value replaced by its default
 * (main.c:108-148; the script relies on it: it passes 0 for tau, theta and nscales); nscales capped so
 * that no scale is much smaller than 16 x 16 (:159-163).  nproc is accepted and ignored (it sets the
 * OpenMP team of the reference).  Colour inputs are read as luminance like iio_read_image_float does.
 * The flow is written by extension (.flo, .tif, .pfm) with two channels.
 */
#include <stdio.h>
#include <stdlib.h>

#include "nlk_image_io.h"
#include "nlk_opts.h"
#include "nlkalman_b200.h"

/* The numeric arguments after `out`, in command-line order, with the value that replaces a missing or
 * out-of-range one (reference lib/tvl1flow/main.c:26-35 for the values, :108-148 for the ranges). */
enum { A_NPROC, A_TAU, A_LAMBDA, A_THETA, A_NSCALES, A_FSCALE, A_ZFACTOR, A_NWARPS, A_EPSILON, A_VERBOSE, A_COUNT };
static const struct {
    const char *name;
    int integer;        /* parsed with atoi (else atof) */
    double dflt;
    double above;       /* valid: value > above ...                    */
    double upto;        /* ... and value <= upto (0: no upper bound)   */
    int below_only;     /* upper bound is exclusive (zfactor < 1)      */
    int checked;        /* 0: any value is taken as it is              */
} ARGS[A_COUNT] = {
    {"nproc",   1, 0,    0, 0,    0, 0},    /* OpenMP team of the reference: accepted, unused */
    {"tau",     0, 0.25, 0, 0.25, 0, 1},
    {"lambda",  0, 0.15, 0, 0,    0, 1},
    {"theta",   0, 0.3,  0, 0,    0, 1},
    {"nscales", 1, 100,  0, 0,    0, 1},
    {"fscale",  1, 0,    0, 0,    0, 0},
    {"zfactor", 0, 0.5,  0, 1,    1, 1},
    {"nwarps",  1, 5,    0, 0,    0, 1},
    {"epsilon", 0, 0.01, 0, 0,    0, 1},
    {"verbose", 1, 0,    0, 0,    0, 0},
};

int main(int argc, char *argv[])
{
    if (argc < 3) {
        fprintf(stderr, "Usage: %s I0 I1 [out", *argv);
        for (int k = 0; k < A_COUNT; ++k) fprintf(stderr, " %s", ARGS[k].name);
        fprintf(stderr, "]\n");
        return EXIT_FAILURE;
    }
    const char *image1_name = argv[1], *image2_name = argv[2];
    const char *outfile = argc > 3 ? argv[3] : "flow.flo";
    double v[A_COUNT];
    int replaced[A_COUNT];
    for (int k = 0; k < A_COUNT; ++k) {
        const char *a = argc > 4 + k ? argv[4 + k] : NULL;
        /* float arguments go through float like in the reference (atof assigned to a float) */
        v[k] = !a ? ARGS[k].dflt : (ARGS[k].integer ? (double)atoi(a) : (double)(float)atof(a));
        const int ok = !ARGS[k].checked ||
                       (v[k] > ARGS[k].above && (ARGS[k].upto == 0 || (ARGS[k].below_only ? v[k] < ARGS[k].upto : v[k] <= ARGS[k].upto)));
        replaced[k] = !ok;
        if (!ok) v[k] = ARGS[k].dflt;
    }
    const int verbose = (int)v[A_VERBOSE];
    if (verbose)
        for (int k = 0; k < A_COUNT; ++k)
            if (replaced[k]) fprintf(stderr, "warning: %s changed to %g\n", ARGS[k].name, v[k]);
    const float tau = (float)v[A_TAU], lambda = (float)v[A_LAMBDA], theta = (float)v[A_THETA], zfactor = (float)v[A_ZFACTOR],
                epsilon = (float)v[A_EPSILON];
    int nscales = (int)v[A_NSCALES], fscale = (int)v[A_FSCALE];
    const int nwarps = (int)v[A_NWARPS];
    if (fscale < 0) fscale = 0;     /* (the reference would index below its pyramid) */

    int nx, ny, nx2, ny2;
    float *I0 = nlk_read_image_gray(image1_name, &nx, &ny);
    if (!I0) return fprintf(stderr, "ERROR: could not read image from file \"%s\": %s\n", image1_name, nlk_io_error()), EXIT_FAILURE;
    float *I1 = nlk_read_image_gray(image2_name, &nx2, &ny2);
    if (!I1) return fprintf(stderr, "ERROR: could not read image from file \"%s\": %s\n", image2_name, nlk_io_error()), EXIT_FAILURE;
    if (nx != nx2 || ny != ny2) {
        fprintf(stderr, "ERROR: input images size mismatch %dx%d != %dx%d\n", nx, ny, nx2, ny2);
        return EXIT_FAILURE;
    }
    nscales = nlk_tvl1_scales(nx, ny, zfactor, nscales);
    if (nscales < fscale) fscale = nscales;
    if (verbose)
        fprintf(stderr, "tau=%f lambda=%f theta=%f nscales=%d zfactor=%f nwarps=%d epsilon=%g\n", tau, lambda, theta,
                nscales, zfactor, nwarps, epsilon);

    nlk_ctx *ctx = nlk_ctx_create(nx, ny, 1, nlk_pick_device());
    if (!ctx) return fprintf(stderr, "tvl1flow: %s\n", nlk_last_error()), 2;
    const size_t n = (size_t)nx * ny;
    float *planes = malloc(2 * n * sizeof(float)), *flow = malloc(2 * n * sizeof(float));
    int *its = calloc((size_t)nscales * nwarps, sizeof(int));
    if (!planes || !flow || !its) return fprintf(stderr, "tvl1flow: out of memory\n"), 2;
    if (nlk_tvl1_flow_host(ctx, I0, I1, planes, nx, ny, tau, lambda, theta, nscales, fscale, zfactor, nwarps, epsilon, its))
        return fprintf(stderr, "tvl1flow: %s\n", nlk_last_error()), 2;
    if (verbose)
        for (int s = nscales - 1; s >= fscale; --s) {
            fprintf(stderr, "Scale %d: iterations", s);
            for (int k = 0; k < nwarps; ++k) fprintf(stderr, " %d", its[s * nwarps + k]);
            fprintf(stderr, "\n");
        }
    /* u, v planes -> interleaved (iio_write_image_float_split, main.c:177) */
    for (size_t k = 0; k < n; ++k) { flow[2 * k] = planes[k]; flow[2 * k + 1] = planes[n + k]; }
    if (nlk_write_image(outfile, flow, nx, ny, 2))
        return fprintf(stderr, "tvl1flow: cannot write %s: %s\n", outfile, nlk_io_error()), EXIT_FAILURE;
    nlk_ctx_destroy(ctx);
    free(I0); free(I1); free(planes); free(flow); free(its);
    return EXIT_SUCCESS;
}
