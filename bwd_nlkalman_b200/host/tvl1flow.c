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

#define PAR_DEFAULT_OUTFLOW "flow.flo"
#define PAR_DEFAULT_TAU     0.25
#define PAR_DEFAULT_LAMBDA  0.15
#define PAR_DEFAULT_THETA   0.3
#define PAR_DEFAULT_NSCALES 100
#define PAR_DEFAULT_FSCALE  0
#define PAR_DEFAULT_ZFACTOR 0.5
#define PAR_DEFAULT_NWARPS  5
#define PAR_DEFAULT_EPSILON 0.01

int main(int argc, char *argv[])
{
    if (argc < 3) {
        fprintf(stderr, "Usage: %s I0 I1 [out nproc tau lambda theta nscales fscale zfactor nwarps epsilon verbose]\n",
                *argv);
        return EXIT_FAILURE;
    }
    int i = 1;
    const char *image1_name = argv[i++], *image2_name = argv[i++];
    const char *outfile = (argc > i) ? argv[i] : PAR_DEFAULT_OUTFLOW; i++;
    i++;                                                          /* nproc */
    float tau     = (argc > i) ? atof(argv[i]) : PAR_DEFAULT_TAU;     i++;
    float lambda  = (argc > i) ? atof(argv[i]) : PAR_DEFAULT_LAMBDA;  i++;
    float theta   = (argc > i) ? atof(argv[i]) : PAR_DEFAULT_THETA;   i++;
    int   nscales = (argc > i) ? atoi(argv[i]) : PAR_DEFAULT_NSCALES; i++;
    int   fscale  = (argc > i) ? atoi(argv[i]) : PAR_DEFAULT_FSCALE;  i++;
    float zfactor = (argc > i) ? atof(argv[i]) : PAR_DEFAULT_ZFACTOR; i++;
    int   nwarps  = (argc > i) ? atoi(argv[i]) : PAR_DEFAULT_NWARPS;  i++;
    float epsilon = (argc > i) ? atof(argv[i]) : PAR_DEFAULT_EPSILON; i++;
    int   verbose = (argc > i) ? atoi(argv[i]) : 0;                   i++;

    if (tau <= 0 || tau > 0.25) { tau = PAR_DEFAULT_TAU; if (verbose) fprintf(stderr, "warning: tau changed to %g\n", tau); }
    if (lambda <= 0) { lambda = PAR_DEFAULT_LAMBDA; if (verbose) fprintf(stderr, "warning: lambda changed to %g\n", lambda); }
    if (theta <= 0) { theta = PAR_DEFAULT_THETA; if (verbose) fprintf(stderr, "warning: theta changed to %g\n", theta); }
    if (nscales <= 0) { nscales = PAR_DEFAULT_NSCALES; if (verbose) fprintf(stderr, "warning: nscales changed to %d\n", nscales); }
    if (zfactor <= 0 || zfactor >= 1) { zfactor = PAR_DEFAULT_ZFACTOR; if (verbose) fprintf(stderr, "warning: zfactor changed to %g\n", zfactor); }
    if (nwarps <= 0) { nwarps = PAR_DEFAULT_NWARPS; if (verbose) fprintf(stderr, "warning: nwarps changed to %d\n", nwarps); }
    if (epsilon <= 0) { epsilon = PAR_DEFAULT_EPSILON; if (verbose) fprintf(stderr, "warning: epsilon changed to %f\n", epsilon); }
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
