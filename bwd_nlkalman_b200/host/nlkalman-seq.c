/* nlkalman-seq -- whole-sequence NL-Kalman filtering and smoothing with the recursion
 * state resident in HBM, host driver.
 *
 * Command line and recursion follow the reference's in-memory sequence driver
 * (reference src/main-seq.c:81-597, not built by the reference's CMake): all paths are
 * printf patterns of the frame number; per frame, first filtering guided by the warped
 * previous first filtering, second filtering (from the second frame on) guided by the
 * warped previous output with the first filtering as basic estimate, then either the
 * one-frame-lag smoother inside the forward loop or the full backward smoother
 * (--s1_full 1).  The recursion is that of the pipeline script and of main-flt.c -- main-seq.c's
 * DECOUPLE_FILTER2 build (src/nlkalman.h:2, src/main-seq.c:451-467): the first filtering of a
 * frame is guided by the warped previous FIRST filtering.  Differences in mechanism, not in
 * results: frames are read one at a time instead of all up front, and every filtered frame
 * stays on the GPU (opponent colour space) until the smoother has consumed it -- nothing but
 * the inputs goes up and nothing but the requested outputs comes back.
 *
 * File I/O runs beside the GPU: a reader thread decodes the next frame's files into pinned
 * staging sets while the current frame is filtered, a writer thread encodes each output once
 * its download has landed (nlk_marker_*); the driving thread only queues uploads, kernels and
 * downloads and never waits for a file.
 *
 * --first_f2 1 applies the second filtering to the first frame too, which is what the
 * per-frame pipeline script does (reference scripts/nlkalman-seq.sh:39-41).
 *
 * --tvl1 1 makes the program the WHOLE pipeline of that script: the backward flow of a frame
 * (TV-L1 between the noisy frame and the previous output, scripts/nlkalman-seq.sh:60-65), its
 * occlusion mask (:68-72), and for the smoother the forward flow and mask (:124-137) are computed
 * on the GPU from the resident frames instead of being read from files -- no tvl1flow / plambda
 * processes, no flow or mask files (the -o / -k / --fflow / --foccl patterns then name optional
 * OUTPUT files).  --of_prms "FSCALE1 DW1 TH1 FSCALE2 DW2 TH2" are the script's OPM values.
 */
#include <pthread.h>
#include <semaphore.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "nlk_image_io.h"
#include "nlk_opts.h"
#include "nlkalman_b200.h"

static void auto_params(struct nlkalman_params *p, int patch)
{
    p->patch_sz = patch;
    p->search_sz_x = p->search_sz_t = -1;
    p->npatches_x = p->npatches_t = p->npatches_tagg = -1;
    p->dista_lambda = p->beta_x = p->beta_t = -1.f;
}

static void print_params(const char *title, const struct nlkalman_params *p, int smoother)
{
    printf("%s\n", title);
    printf("\tpatch      %d\n", p->patch_sz);
    if (!smoother) printf("\tsearch_x   %d\n", p->search_sz_x);
    printf("\tsearch_t   %d\n", p->search_sz_t);
    if (!smoother) printf("\tnp_x       %d\n", p->npatches_x);
    printf("\tnp_t       %d\n", p->npatches_t);
    printf("\tnp_tagg    %d\n", p->npatches_tagg);
    printf("\tlambda     %g\n", p->dista_lambda);
    if (!smoother) printf("\tbeta_x     %g\n", p->beta_x);
    printf("\tbeta_t     %g\n", p->beta_t);
    printf("\n");
}

static nlk_ctx *ctx;
static int w, h, c;
static size_t ib, npix;

static int gpu_fail(const char *what)
{
    fprintf(stderr, "nlkalman-seq: %s: %s\n", what, nlk_last_error());
    return 1;
}

/* reads pattern % f; checks the size against the sequence; NULL if the pattern is NULL */
static float *read_frame(const char *pattern, int f, int want_c, int *err)
{
    if (!pattern) return NULL;
    char name[1024];
    snprintf(name, sizeof name, pattern, f);
    int w1, h1, c1;
    float *x = nlk_read_image(name, &w1, &h1, &c1);
    if (!x) { fprintf(stderr, "Error: %s\n", nlk_io_error()); *err = 1; return NULL; }
    if (w1 != w || h1 != h || c1 != want_c) {
        fprintf(stderr, "Error: %s is %dx%dx%d, expected %dx%dx%d\n", name, w1, h1, c1, w, h, want_c);
        free(x);
        *err = 1;
        return NULL;
    }
    return x;
}

/* ---- reader thread: the files of the coming requests, decoded into pinned staging sets ------- */
#define N_IN 3
#define N_OUT 4
enum { REQ_FRAME = 0 /* noisy frame f + its backward flow and mask */, REQ_FLOW = 1 /* forward flow and mask of f */ };
typedef struct { int kind, f, with_flow; } in_req;
typedef struct {
    float *img, *flo, *occ;     /* pinned */
    int has_img, has_flo, has_occ, err;
    void *consumed;             /* marker: the uploads out of this set have finished */
} in_set;
static struct {
    in_req *req;
    int nreq;
    in_set set[N_IN];
    sem_t free_sets, ready_sets;
    const char *nisy, *bflo, *bocc, *fflo, *focc;
} rd;

static void *reader_main(void *arg)
{
    (void)arg;
    for (int i = 0; i < rd.nreq; ++i) {
        sem_wait(&rd.free_sets);
        in_set *s = &rd.set[i % N_IN];
        if (s->consumed) { nlk_marker_wait(s->consumed); s->consumed = NULL; }
        const in_req q = rd.req[i];
        s->has_img = s->has_flo = s->has_occ = 0;
        int err = 0;
        if (q.kind == REQ_FRAME) {
            float *x = read_frame(rd.nisy, q.f, c, &err);
            if (x) { memcpy(s->img, x, ib); free(x); s->has_img = 1; } else err = 1;
        }
        if (!err && q.with_flow) {
            float *of = read_frame(q.kind == REQ_FRAME ? rd.bflo : rd.fflo, q.f, 2, &err);
            float *oc = of ? read_frame(q.kind == REQ_FRAME ? rd.bocc : rd.focc, q.f, 1, &err) : NULL;
            if (of) { memcpy(s->flo, of, npix * 2 * sizeof(float)); s->has_flo = 1; }
            if (oc) { memcpy(s->occ, oc, npix * sizeof(float)); s->has_occ = 1; }
            free(of);
            free(oc);
        }
        s->err = err;
        sem_post(&rd.ready_sets);
        if (err) break;
    }
    return NULL;
}

static int rd_next;
/* the next request's set (blocks until the reader has it); release_set after its uploads are queued */
static in_set *acquire_set(void)
{
    sem_wait(&rd.ready_sets);
    in_set *s = &rd.set[rd_next++ % N_IN];
    return s->err ? NULL : s;
}
static void release_set(in_set *s)
{
    s->consumed = nlk_marker_record(ctx);
    sem_post(&rd.free_sets);
}

/* flow / occlusion of a staging set -> device; *d_flow_out = NULL when there is no flow */
static int upload_flow(const in_set *s, float *d_of, float *d_occ, const float **d_flow_out, const float **d_occ_out)
{
    *d_flow_out = *d_occ_out = NULL;
    if (s->has_flo) {
        if (nlk_upload(ctx, d_of, s->flo, npix * 2 * sizeof(float))) return gpu_fail("upload");
        *d_flow_out = d_of;
        if (s->has_occ) {
            if (nlk_upload(ctx, d_occ, s->occ, npix * sizeof(float))) return gpu_fail("upload");
            *d_occ_out = d_occ;
        }
    }
    return 0;
}

/* ---- writer thread: outputs encoded and written once their download has landed ---------------- */
typedef struct { void *marker; char name[1024]; int stop, ch; } out_job;
static struct {
    float *buf[N_OUT];          /* pinned */
    out_job job[N_OUT];
    sem_t free_bufs, ready_jobs;
    int head, failed;
} wr;

static void *writer_main(void *arg)
{
    (void)arg;
    for (int i = 0;; ++i) {
        sem_wait(&wr.ready_jobs);
        out_job *j = &wr.job[i % N_OUT];
        if (j->stop) break;
        if (nlk_marker_wait(j->marker)) { fprintf(stderr, "nlkalman-seq: output: %s\n", nlk_last_error()); wr.failed = 1; }
        else if (nlk_write_image(j->name, wr.buf[i % N_OUT], w, h, j->ch)) { fprintf(stderr, "Error: %s\n", nlk_io_error()); wr.failed = 1; }
        sem_post(&wr.free_bufs);
    }
    return NULL;
}

/* device (opponent space) -> RGB -> pinned buffer -> (writer thread) file pattern % f */
static int write_frame(const char *pattern, int f, const float *d_opp, float *d_scratch)
{
    sem_wait(&wr.free_bufs);
    const int k = wr.head++ % N_OUT;
    out_job *j = &wr.job[k];
    snprintf(j->name, sizeof j->name, pattern, f);
    j->stop = 0;
    j->ch = c;
    if (nlk_opp2rgb_dev(ctx, d_scratch, d_opp) || nlk_download(ctx, wr.buf[k], d_scratch, ib)) return gpu_fail("output");
    j->marker = nlk_marker_record(ctx);
    if (!j->marker) return gpu_fail("output");
    sem_post(&wr.ready_jobs);
    return wr.failed;
}

/* a computed flow (ch = 2) or mask (ch = 1) -> pinned buffer -> (writer thread) file pattern % f */
static int write_plane(const char *pattern, int f, const float *d_src, int ch)
{
    if (!pattern) return 0;
    sem_wait(&wr.free_bufs);
    const int k = wr.head++ % N_OUT;
    out_job *j = &wr.job[k];
    snprintf(j->name, sizeof j->name, pattern, f);
    j->stop = 0;
    j->ch = ch;
    if (nlk_download(ctx, wr.buf[k], d_src, npix * ch * sizeof(float))) return gpu_fail("output");
    j->marker = nlk_marker_record(ctx);
    if (!j->marker) return gpu_fail("output");
    sem_post(&wr.ready_jobs);
    return wr.failed;
}

int main(int argc, const char *argv[])
{
    const char *nisy_path = NULL, *bflo_path = NULL, *bocc_path = NULL, *fflo_path = NULL, *focc_path = NULL;
    const char *flt1_path = NULL, *flt2_path = NULL, *smo1_path = NULL;
    int fframe = 0, lframe = -1, verbose = 0, full = 1, first_f2 = 0, tvl1 = 0;
    const char *of_prms = "1 0.25 0.75 1 0.25 0.75";   /* the script's OPM default (scripts/nlkalman-seq.sh:17) */
    float sigma = 0.f;
    struct nlkalman_params f1, f2, s1;
    auto_params(&f1, -1);
    auto_params(&f2, -1);
    auto_params(&s1, 0);   /* smoothing is off unless --s1_p is given (reference src/main-seq.c:131) */

    const struct nlk_opt options[] = {
        {NLK_OPT_GROUP, 0, "Data i/o options (all paths in printf format)", NULL, NULL},
        {NLK_OPT_STRING, 'i', "nisy", &nisy_path, "input noisy frames path"},
        {NLK_OPT_STRING, 'o', "bflow", &bflo_path, "input bwd flow path"},
        {NLK_OPT_STRING, 'k', "boccl", &bocc_path, "input bwd occlusion masks path"},
        {NLK_OPT_STRING, 0, "fflow", &fflo_path, "input fwd flow path"},
        {NLK_OPT_STRING, 0, "foccl", &focc_path, "input fwd occlusion masks path"},
        {NLK_OPT_STRING, 0, "filt1", &flt1_path, "output first filtering path"},
        {NLK_OPT_STRING, 0, "filt2", &flt2_path, "output second filtering path"},
        {NLK_OPT_STRING, 0, "smoo1", &smo1_path, "output smoothing path"},
        {NLK_OPT_INT, 'f', "first", &fframe, "first frame"},
        {NLK_OPT_INT, 'l', "last", &lframe, "last frame"},
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
        {NLK_OPT_INT, 0, "first_f2", &first_f2, "1: second filtering on the first frame too (as nlkalman-seq.sh)"},
        {NLK_OPT_GROUP, 0, "Smoothing options", NULL, NULL},
        {NLK_OPT_INT, 0, "s1_p", &s1.patch_sz, "patch size"},
        {NLK_OPT_INT, 0, "s1_st", &s1.search_sz_t, "search region radius"},
        {NLK_OPT_INT, 0, "s1_nt", &s1.npatches_t, "number of similar patches kalman"},
        {NLK_OPT_INT, 0, "s1_nt_agg", &s1.npatches_tagg, "number of similar patches kalman spatial average"},
        {NLK_OPT_FLOAT, 0, "s1_bt", &s1.beta_t, "noise multiplier in kalman filtering"},
        {NLK_OPT_FLOAT, 0, "s1_l", &s1.dista_lambda, "noisy patch weight in patch distance"},
        {NLK_OPT_INT, 0, "s1_full", &full, "0: next frame smoothing, 1: full video smoothing (default)"},
        {NLK_OPT_GROUP, 0, "Optical flow options", NULL, NULL},
        {NLK_OPT_INT, 0, "tvl1", &tvl1, "1: flows and occlusion masks computed on the GPU (the flow / mask paths become outputs)"},
        {NLK_OPT_STRING, 0, "of_prms", &of_prms, "\"FSCALE1 DW1 TH1 FSCALE2 DW2 TH2\" (filtering, smoothing), as nlkalman-seq.sh"},
        {NLK_OPT_GROUP, 0, "Program options", NULL, NULL},
        {NLK_OPT_INT, 'v', "verbose", &verbose, "verbose output"},
        {NLK_OPT_END, 0, NULL, NULL, NULL},
    };
    nlk_opts_parse(options, "nlkalman-seq [options] [[--] args]",
                   "\nA video denoiser based on non-local Kalman filtering.", argc, argv);

    /* modes (reference src/main-seq.c:223-246) */
    const int second_filt = f2.patch_sz && (flt2_path || smo1_path);
    const int lag_smoother = !full && s1.patch_sz && smo1_path;
    const int full_smoother = full && s1.patch_sz && smo1_path;
    if (f1.patch_sz == 0) return fprintf(stderr, "Error: f1_p == 0, exiting\n"), 1;
    if (!flt1_path && !(flt2_path && f2.patch_sz) && !lag_smoother && !full_smoother)
        return fprintf(stderr, "Error: no output path given for any computed output - exiting\n"), 1;
    if (!flt1_path && !flt2_path && s1.patch_sz == 0)
        return fprintf(stderr, "Error: s1_p == 0 and no output paths given for filt1 and filt2\n"), 1;
    if (f2.patch_sz == 0 && flt2_path)
        fprintf(stderr, "Warning: f2_p == 0 - no output files will be stored in %s\n", flt2_path);
    if (s1.patch_sz == 0 && smo1_path)
        fprintf(stderr, "Warning: s1_p == 0 - no output files will be stored in %s\n", smo1_path);
    if (!nisy_path || lframe < fframe) return fprintf(stderr, "Error: no input frames (-i, -f, -l)\n"), 1;

    /* flow parameters: "NPROC 0 DW 0 0 FSCALE" on tvl1flow's command line (scripts/nlkalman-seq.sh:51, :111) */
    struct nlk_tvl1_params of1, of2;
    float th1 = 0.75f, th2 = 0.75f;
    nlk_tvl1_default_params(&of1);
    nlk_tvl1_default_params(&of2);
    if (tvl1) {
        int fs1, fs2;
        float dw1, dw2;
        if (sscanf(of_prms, "%d %f %f %d %f %f", &fs1, &dw1, &th1, &fs2, &dw2, &th2) != 6)
            return fprintf(stderr, "Error: --of_prms wants six values \"FSCALE1 DW1 TH1 FSCALE2 DW2 TH2\"\n"), 1;
        of1.fscale = fs1; of1.lambda = dw1;
        of2.fscale = fs2; of2.lambda = dw2;
    }
    /* with --tvl1 the flow / mask patterns are outputs; nothing is read from them */
    const char *bflo_out = tvl1 ? bflo_path : NULL, *bocc_out = tvl1 ? bocc_path : NULL;
    const char *fflo_out = tvl1 ? fflo_path : NULL, *focc_out = tvl1 ? focc_path : NULL;
    if (tvl1) bflo_path = bocc_path = fflo_path = focc_path = NULL;

    nlkalman_default_params(&f1, sigma, FLT1);
    nlkalman_default_params(&f2, sigma, FLT2);
    nlkalman_default_params(&s1, sigma, SMO1);

    if (verbose) {
        printf("data input:\n");
        printf("\tnoise         %05.2f\n", sigma);
        printf("\tfirst frame   %d\n", fframe);
        printf("\tlast frame    %d\n", lframe);
        printf("\tnoisy frames  %s\n", nisy_path);
        if (tvl1) printf("\tflows         TV-L1 on the GPU, \"%s\"\n", of_prms);
        printf("\tbwd flows     %s\n", tvl1 ? bflo_out : bflo_path);
        printf("\tfwd flows     %s\n", tvl1 ? fflo_out : fflo_path);
        printf("\tbwd occlus.   %s\n", tvl1 ? bocc_out : bocc_path);
        printf("\tfwd occlus.   %s\n", tvl1 ? focc_out : focc_path);
        printf("\n");
        printf("data output:\n");
        printf("\tfiltering 1   %s\n", flt1_path);
        printf("\tfiltering 2   %s\n", flt2_path);
        printf("\tsmoothing 1   %s\n", smo1_path);
        printf("\n");
        print_params("first filtering parameters:", &f1, 0);
        if (second_filt) print_params("second filtering parameters:", &f2, 0);
        if (lag_smoother || full_smoother)
            print_params(full_smoother ? "full smoother params:" : "single frame smoother params:", &s1, 1);
    }

    /* the first frame fixes the geometry */
    {
        char name[1024];
        snprintf(name, sizeof name, nisy_path, fframe);
        float *x = nlk_read_image(name, &w, &h, &c);
        if (!x) return fprintf(stderr, "Error: %s\n", nlk_io_error()), 1;
        free(x);
    }
    ib = (size_t)w * h * c * sizeof(float);
    npix = (size_t)w * h;
    const int dev_pick = nlk_pick_device();
    int dev = dev_pick;
    ctx = nlk_ctx_create(w, h, c, dev);
    if (!ctx) return gpu_fail("no usable CUDA device (there is no CPU fallback)");

    /* HBM-resident state: the output of every frame when a smoother needs it later,
     * else two slots used alternately; plus the previous first filtering */
    const int nframes = lframe - fframe + 1;
    const int keep_all = full_smoother;
    const int nslots = keep_all ? nframes : 2;
    float **d_deno = (float **)calloc((size_t)nslots, sizeof(float *));
    for (int i = 0; i < nslots; ++i) if (!(d_deno[i] = nlk_dev_alloc(ctx, ib))) return gpu_fail("device memory");
    float *d_nisy = nlk_dev_alloc(ctx, ib), *d_warp = nlk_dev_alloc(ctx, ib), *d_tmp = nlk_dev_alloc(ctx, ib);
    float *d_rgb[2] = {nlk_dev_alloc(ctx, ib), nlk_dev_alloc(ctx, ib)};   /* RGB scratch of the two outputs of a frame */
    float *d_bsic[2] = {nlk_dev_alloc(ctx, ib), nlk_dev_alloc(ctx, ib)};
    float *d_of = nlk_dev_alloc(ctx, npix * 2 * sizeof(float)), *d_occ = nlk_dev_alloc(ctx, npix * sizeof(float));
    float *d_of2 = nlk_dev_alloc(ctx, npix * 2 * sizeof(float)), *d_occ2 = nlk_dev_alloc(ctx, npix * sizeof(float));
    /* RGB copies of the two frames a flow is estimated between */
    float *d_from = tvl1 ? nlk_dev_alloc(ctx, ib) : NULL, *d_to = tvl1 ? nlk_dev_alloc(ctx, ib) : NULL;
    if (tvl1 && (!d_from || !d_to)) return gpu_fail("device memory");
    if (!d_nisy || !d_warp || !d_tmp || !d_rgb[0] || !d_rgb[1] || !d_bsic[0] || !d_bsic[1] || !d_of || !d_occ || !d_of2 || !d_occ2)
        return gpu_fail("device memory");
#define SLOT(f) d_deno[keep_all ? (f) - fframe : ((f) - fframe) & 1]

    /* the files, in the order the loops below consume them */
    rd.nisy = nisy_path; rd.bflo = bflo_path; rd.bocc = bocc_path; rd.fflo = fflo_path; rd.focc = focc_path;
    rd.req = (in_req *)calloc((size_t)3 * nframes + 1, sizeof(in_req));
    for (int f = fframe; f <= lframe; ++f) {
        rd.req[rd.nreq++] = (in_req){REQ_FRAME, f, f > fframe && bflo_path != NULL};
        if (lag_smoother && f > fframe) rd.req[rd.nreq++] = (in_req){REQ_FLOW, f - 1, fflo_path != NULL};
    }
    if (full_smoother)
        for (int f = lframe - 1; f >= fframe; --f) rd.req[rd.nreq++] = (in_req){REQ_FLOW, f, fflo_path != NULL};
    for (int i = 0; i < N_IN; ++i) {
        rd.set[i].img = (float *)nlk_host_alloc(ib);
        rd.set[i].flo = (float *)nlk_host_alloc(npix * 2 * sizeof(float));
        rd.set[i].occ = (float *)nlk_host_alloc(npix * sizeof(float));
        if (!rd.set[i].img || !rd.set[i].flo || !rd.set[i].occ) return gpu_fail("pinned memory");
    }
    const size_t ob = ib > npix * 2 * sizeof(float) ? ib : npix * 2 * sizeof(float);   /* a frame or a flow */
    for (int i = 0; i < N_OUT; ++i) if (!(wr.buf[i] = (float *)nlk_host_alloc(ob))) return gpu_fail("pinned memory");
    sem_init(&rd.free_sets, 0, N_IN);
    sem_init(&rd.ready_sets, 0, 0);
    sem_init(&wr.free_bufs, 0, N_OUT);
    sem_init(&wr.ready_jobs, 0, 0);
    pthread_t reader, writer;
    if (pthread_create(&reader, NULL, reader_main, NULL) || pthread_create(&writer, NULL, writer_main, NULL))
        return fprintf(stderr, "Error: cannot start the I/O threads\n"), 1;

    /* ---- forward: filtering (reference src/main-seq.c:444-550) ---------------------------- */
    for (int f = fframe; f <= lframe; ++f) {
        if (verbose) printf("processing frame %d\n", f);
        in_set *in = acquire_set();
        if (!in) return 1;
        if (nlk_upload(ctx, d_nisy, in->img, ib)) return gpu_fail("upload");
        float *bsic1 = d_bsic[(f - fframe) & 1], *bsic0 = d_bsic[(f - fframe + 1) & 1];
        const float *d_flow = NULL, *d_mask = NULL;
        if (upload_flow(in, d_of, d_occ, &d_flow, &d_mask)) return 1;
        release_set(in);
        if (tvl1 && f > fframe) {
            /* backward flow: from the noisy frame to the previous output, both as the script's files hold
             * them (RGB) (scripts/nlkalman-seq.sh:60-72) */
            if (nlk_opp2rgb_dev(ctx, d_to, SLOT(f - 1)) || nlk_flow_mask_dev(ctx, d_of, d_occ, d_nisy, d_to, of1, th1))
                return gpu_fail("optical flow");
            d_flow = d_of; d_mask = d_occ;
            if (write_plane(bflo_out, f, d_of, 2) || write_plane(bocc_out, f, d_occ, 1)) return 1;
        }
        if (nlk_rgb2opp_dev(ctx, d_nisy, d_nisy)) return gpu_fail("colour transform");

        /* first filtering, guided by the previous first filtering */
        const float *prev = NULL;
        if (f > fframe) {
            prev = bsic0;
            if (d_flow) { if (nlk_warp_dev(ctx, d_warp, bsic0, d_flow, d_mask)) return gpu_fail("warp"); prev = d_warp; }
        }
        if (nlk_pass_dev(ctx, 0, bsic1, d_nisy, prev, NULL, sigma, f1)) return gpu_fail("first filtering");
        if (flt1_path && write_frame(flt1_path, f, bsic1, d_rgb[0])) return 1;

        /* second filtering, guided by the previous output */
        float *deno1 = SLOT(f);
        if (second_filt && (f > fframe || first_f2)) {
            prev = NULL;
            if (f > fframe) {
                prev = SLOT(f - 1);
                if (d_flow) { if (nlk_warp_dev(ctx, d_warp, prev, d_flow, d_mask)) return gpu_fail("warp"); prev = d_warp; }
            }
            if (nlk_pass_dev(ctx, 0, deno1, d_nisy, prev, bsic1, sigma, f2)) return gpu_fail("second filtering");
        } else {
            /* the output of this frame is its first filtering */
            if (nlk_copy_dev(ctx, deno1, bsic1, ib)) return gpu_fail("copy");
        }
        if (second_filt && flt2_path && write_frame(flt2_path, f, deno1, d_rgb[1])) return 1;

        /* one-frame-lag smoother: frame f-1 smoothed against frame f */
        if (lag_smoother && f > fframe) {
            in_set *ff = acquire_set();
            if (!ff) return 1;
            const float *d_ff = NULL, *d_fm = NULL;
            if (upload_flow(ff, d_of2, d_occ2, &d_ff, &d_fm)) return 1;
            release_set(ff);
            if (tvl1) {   /* forward flow of frame f-1: from its filtered output to the next frame's (:124-137) */
                if (nlk_opp2rgb_dev(ctx, d_from, SLOT(f - 1)) || nlk_opp2rgb_dev(ctx, d_to, deno1) ||
                    nlk_flow_mask_dev(ctx, d_of2, d_occ2, d_from, d_to, of2, th2)) return gpu_fail("optical flow");
                d_ff = d_of2; d_fm = d_occ2;
                if (write_plane(fflo_out, f - 1, d_of2, 2) || write_plane(focc_out, f - 1, d_occ2, 1)) return 1;
            }
            const float *smoo0 = deno1;
            if (d_ff) { if (nlk_warp_dev(ctx, d_warp, deno1, d_ff, d_fm)) return gpu_fail("warp"); smoo0 = d_warp; }
            float *filt1 = SLOT(f - 1);
            if (nlk_pass_dev(ctx, 1, d_tmp, filt1, smoo0, NULL, sigma, s1)) return gpu_fail("smoothing");
            if (nlk_copy_dev(ctx, filt1, d_tmp, ib)) return gpu_fail("copy");
            if (write_frame(smo1_path, f - 1, filt1, d_rgb[0])) return 1;
        }
    }

    /* ---- backward: full smoother (reference src/main-seq.c:553-584) ------------------------ */
    if (full_smoother) {
        for (int f = lframe - 1; f >= fframe; --f) {
            if (verbose) printf("processing frame %d\n", f);
            in_set *ff = acquire_set();
            if (!ff) return 1;
            const float *d_ff = NULL, *d_fm = NULL;
            if (upload_flow(ff, d_of, d_occ, &d_ff, &d_fm)) return 1;
            release_set(ff);
            if (tvl1) {   /* forward flow: from the filtered frame to the smoothed next one (:124-137) */
                if (nlk_opp2rgb_dev(ctx, d_from, SLOT(f)) || nlk_opp2rgb_dev(ctx, d_to, SLOT(f + 1)) ||
                    nlk_flow_mask_dev(ctx, d_of, d_occ, d_from, d_to, of2, th2)) return gpu_fail("optical flow");
                d_ff = d_of; d_fm = d_occ;
                if (write_plane(fflo_out, f, d_of, 2) || write_plane(focc_out, f, d_occ, 1)) return 1;
            }
            const float *smoo0 = SLOT(f + 1);
            if (d_ff) { if (nlk_warp_dev(ctx, d_warp, smoo0, d_ff, d_fm)) return gpu_fail("warp"); smoo0 = d_warp; }
            float *filt1 = SLOT(f);
            if (nlk_pass_dev(ctx, 1, d_nisy, filt1, smoo0, NULL, sigma, s1)) return gpu_fail("smoothing");
            if (nlk_copy_dev(ctx, filt1, d_nisy, ib)) return gpu_fail("copy");
            if (write_frame(smo1_path, f, filt1, d_rgb[(f - fframe) & 1])) return 1;
        }
    }

    /* drain the writer, then leave without tearing the CUDA context down (the files are on disk) */
    sem_wait(&wr.free_bufs);
    wr.job[wr.head++ % N_OUT].stop = 1;
    sem_post(&wr.ready_jobs);
    pthread_join(writer, NULL);
    if (nlk_ctx_sync(ctx)) return gpu_fail("sync");
    fflush(NULL);
    _exit(wr.failed ? 1 : EXIT_SUCCESS);
}
