/* nlkalman-occ -- occlusion mask from the divergence of an optical flow, on the GPU.
 *
 *     nlkalman-occ FLOW.flo TH OUT.png
 *
 * Stands in for the plambda call of the pipeline script (reference scripts/nlkalman-seq.sh:70-72):
 *     plambda FLOW "x(0,0)[0] x(-1,0)[0] - x(0,0)[1] x(0,-1)[1] - + fabs TH > 255 *" -o OUT
 * so that the script needs no image-processing tool besides the flow estimator.
 */
#include <stdio.h>
#include <stdlib.h>

#include "nlk_image_io.h"
#include "nlk_opts.h"
#include "nlkalman_b200.h"

int main(int argc, const char *argv[])
{
    if (argc != 4) return fprintf(stderr, "usage: nlkalman-occ FLOW TH OUT\n"), 1;
    int w, h, c;
    float *of = nlk_read_image(argv[1], &w, &h, &c);
    if (!of) return fprintf(stderr, "nlkalman-occ: cannot read %s: %s\n", argv[1], nlk_io_error()), 1;
    if (c != 2) return fprintf(stderr, "nlkalman-occ: %s has %d channels, a flow has 2\n", argv[1], c), 1;
    const float th = (float)atof(argv[2]);
    nlk_ctx *ctx = nlk_ctx_create(w, h, 1, nlk_pick_device());
    if (!ctx) return fprintf(stderr, "nlkalman-occ: %s\n", nlk_last_error()), 2;
    float *occ = malloc((size_t)w * h * sizeof(float));
    if (!occ || nlk_occlusion_host(ctx, occ, of, th))
        return fprintf(stderr, "nlkalman-occ: %s\n", nlk_last_error()), 2;
    if (nlk_write_image(argv[3], occ, w, h, 1))
        return fprintf(stderr, "nlkalman-occ: cannot write %s: %s\n", argv[3], nlk_io_error()), 1;
    nlk_ctx_destroy(ctx);
    free(occ);
    free(of);
    return 0;
}
