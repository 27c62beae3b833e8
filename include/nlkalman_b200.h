/* nlkalman_b200.h -- additive C ABI of libnlkalman_b200.so next to the six drop-in
 * entry points of nlkalman.h.  It exposes what the per-frame process boundary of the
 * reference hides: a context whose buffers (and the previous filtered frames, i.e. the
 * recursion state that the reference passes between processes as TIFF files,
 * scripts/nlkalman-seq.sh:80-102) stay resident in HBM across frames.
 *
 * Conventions: plain C types only; images are float32 interleaved HWC; "h_" pointers
 * are host memory (pinned memory makes the copies asynchronous-capable and faster),
 * "d_" pointers are device memory on the context's GPU.  Functions return 0 on
 * success and a negative code on failure; nlk_last_error() gives the message.  There
 * is no CPU fallback anywhere.
 */
#ifndef NLKALMAN_B200_H
#define NLKALMAN_B200_H

#include "nlkalman.h"
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nlk_ctx nlk_ctx;

#define NLK_OK 0
#define NLK_ERR_CUDA (-1)    /* CUDA runtime error (message has the details) */
#define NLK_ERR_PARAM (-2)   /* unsupported parameter (patch size, channels, radius ...) */
#define NLK_ERR_STATE (-3)   /* call not valid in the current sequence state */

/* number of visible CUDA devices, or a negative code */
int nlk_device_count(void);
const char *nlk_last_error(void);

/* context for frames of w x h x ch on CUDA device `device` (own stream, lazily sized
 * scratch).  Returns NULL on failure. */
nlk_ctx *nlk_ctx_create(int w, int h, int ch, int device);
void nlk_ctx_destroy(nlk_ctx *ctx);
/* block until everything queued on the context's stream has finished */
int nlk_ctx_sync(nlk_ctx *ctx);
/* CUDA kernels launched by this context so far */
long long nlk_ctx_launch_count(const nlk_ctx *ctx);
/* the context's cudaStream_t, for callers that queue their own work around it */
void *nlk_ctx_stream(nlk_ctx *ctx);

/* Optional per-kernel timing: when enabled every kernel the context launches is
 * bracketed by CUDA events on the context's stream.  nlk_ctx_profile_collect waits for
 * the stream, then fills ms_sum / count, both [NLK_KERNEL_COUNT][NLK_PASS_KINDS], with
 * the summed durations and the launch-group counts since the last collection. */
enum { NLK_K_COLOUR = 0, NLK_K_WARP, NLK_K_VALID, NLK_K_SEARCH, NLK_K_RESOLVE, NLK_K_GROUP,
       NLK_K_NORMALIZE, NLK_K_MEMSET, NLK_K_PEER_WAIT, NLK_K_PEER_PUSH, NLK_KERNEL_COUNT };
enum { NLK_PASS_FLT1_T = 0, /* filter, previous frame given, no basic estimate */
       NLK_PASS_FLT1_X,     /* filter, no previous frame, no basic estimate  */
       NLK_PASS_FLT2_T,     /* filter with basic estimate, previous frame given */
       NLK_PASS_FLT2_X,     /* filter with basic estimate, no previous frame */
       NLK_PASS_SMO,        /* smoother */
       NLK_PASS_OTHER,      /* outside a pass (colour transform, warp) */
       NLK_PASS_KINDS };
int nlk_ctx_profile(nlk_ctx *ctx, int enable);
int nlk_ctx_profile_collect(nlk_ctx *ctx, double *ms_sum, int *count);
/* while profiling: per pass kind, the processed patches (those the reference's mask test
 * src/nlkalman.c:597-600 does not skip) and the grid patches, summed over the passes since the
 * last call; both [NLK_PASS_KINDS].  alpha = active / grid. */
int nlk_ctx_profile_alpha(nlk_ctx *ctx, double *active_sum, double *grid_sum);

/* sustained fp32 FMA throughput of the device in TFLOP/s (2 flops per FMA), measured
 * with a register-resident FMA kernel for about `ms` milliseconds: the denominator of
 * the FP32 roofline that bounds search_knn and group_filter */
int nlk_fp32_peak(nlk_ctx *ctx, float ms, double *tflops);

/* pinned host memory (cudaMallocHost / cudaFreeHost) for the h_ arguments below */
void *nlk_host_alloc(size_t bytes);
void nlk_host_free(void *p);

/* device memory on the context's GPU and copies queued on the context's stream (the
 * copies are asynchronous when the host side is pinned; nlk_ctx_sync waits for them) */
void *nlk_dev_alloc(nlk_ctx *ctx, size_t bytes);
void nlk_dev_free(nlk_ctx *ctx, void *d_ptr);
int nlk_upload(nlk_ctx *ctx, void *d_dst, const void *h_src, size_t bytes);
int nlk_download(nlk_ctx *ctx, void *h_dst, const void *d_src, size_t bytes);
int nlk_copy_dev(nlk_ctx *ctx, void *d_dst, const void *d_src, size_t bytes);
/* A point of the context's stream (everything queued so far) that ANOTHER host thread may wait for:
 * nlk_marker_record is called by the thread that drives the context, nlk_marker_wait (blocking, frees
 * the marker) by any thread -- how a file writer learns that a download has landed while the driving
 * thread goes on queueing the next frame. */
void *nlk_marker_record(nlk_ctx *ctx);
int nlk_marker_wait(void *marker);

/* ---- single operations on device buffers (asynchronous on the context's stream) ---- */
int nlk_rgb2opp_dev(nlk_ctx *ctx, float *d_dst, const float *d_src);   /* src may equal dst */
int nlk_opp2rgb_dev(nlk_ctx *ctx, float *d_dst, const float *d_src);
int nlk_warp_dev(nlk_ctx *ctx, float *d_imw, const float *d_im, const float *d_of,
                 const float *d_msk /* may be NULL */);
/* one filter (smooth = 0) or smoother (smooth = 1) pass, semantics of
 * nlkalman_filter_frame / nlkalman_smooth_frame; d_prev0 and d_bsic1 may be NULL */
int nlk_pass_dev(nlk_ctx *ctx, int smooth, float *d_out, const float *d_in1,
                 const float *d_prev0, const float *d_bsic1, float sigma,
                 struct nlkalman_params prms);

/* occlusion mask from the divergence of a flow field, the plambda expression of the pipeline
 * script between tvl1flow and the filter (reference scripts/nlkalman-seq.sh:70-72, :95-97):
 * occ = 255 where |(u(x,y) - u(x-1,y)) + (v(x,y) - v(x,y-1))| > th, else 0; neighbours outside
 * the image are replaced by the nearest sample.  Lets flow stay on the device between the
 * flow estimator and the filter. */
int nlk_occlusion_dev(nlk_ctx *ctx, float *d_occ, const float *d_of, float th);
int nlk_occlusion_host(nlk_ctx *ctx, float *h_occ, const float *h_of, float th);

/* ---- row ranges and the strip-sharded pass (one GPU per horizontal strip) -----------
 * Within a frame the reference's loop over patches (src/nlkalman.c:590-595 `for py .. for
 * px`) shards as strips of grid-patch rows.  A rank searches and filters the reference
 * patches of its grid rows [gy0, gy1); what crosses strips is
 *   - the neighbour bitmaps of every grid row (the processed-mask chain of :597-600 with
 *     :930-931 is sequential over the whole frame): d_nbr, [gh*gw][nbw] words, rows
 *     [gy0, gy1) written by nlk_strip_search, the rest to be all-gathered by the caller;
 *   - the aggregation (:913-928) of groups near a strip border lands in the neighbour's
 *     pixel rows: d_accw, [h][w][ch+1] floats, rows [ey0, ey1) zeroed by
 *     nlk_strip_search and accumulated into by nlk_strip_filter; the caller adds rows
 *     [ey0, oy0) and [oy1, ey1) into the neighbours' buffers before normalising;
 *   - inputs must be valid on rows [ey0, ey1) (halo = search radius + patch size).
 * Everything is asynchronous on the context's stream; the caller queues its exchange
 * (NCCL through torch.distributed in bwd_nlkalman_b200/strips.py) on the same stream. */
struct nlk_strip_plan {
    int gw, gh, nbw;   /* grid of reference patches, bitmap words per patch */
    int gy0, gy1;      /* grid rows of this rank */
    int oy0, oy1;      /* pixel rows this rank owns: normalised and output here */
    int ey0, ey1;      /* pixel rows this rank reads and accumulates into (own + halo) */
    int chunk_g, chunk_y; /* grid rows / pixel rows per rank (all but the last): rank k's rows start at k * chunk */
};
/* pure host arithmetic: equal chunks of ceil(gh / nranks) grid rows, the last strip takes the
 * rest; fails if a strip would be thinner than the halo */
int nlk_strip_plan(int w, int h, int smooth, struct nlkalman_params prms, int nranks, int rank,
                   struct nlk_strip_plan *out);
/* colour transform (inverse = 0: rgb2opp, 1: opp2rgb) and warp restricted to pixel rows
 * [row0, row1); pointers are to the full frames */
int nlk_colour_rows_dev(nlk_ctx *ctx, float *d_dst, const float *d_src, int inverse, int row0, int row1);
int nlk_warp_rows_dev(nlk_ctx *ctx, float *d_imw, const float *d_im, const float *d_of,
                      const float *d_msk, int row0, int row1);
/* The strip calls, the row-range calls and the nlk_peer_* calls queue on lane 0 (the context's stream)
 * or, after nlk_strip_lane(ctx, 1, ...), on lane 1 (second stream, own pass scratch): a caller can run
 * the two filterings of a frame as two pipelines, the second filtering of frame t beside the first of
 * frame t+1, as the single-GPU recursion does.  reserve_sm: SMs the persistent group_filter launch
 * leaves free for the other lane's one-block mask_resolve.  Cross-lane order is the caller's:
 * nlk_lane_record(ctx, i) records event i (0..7) on the current lane, nlk_lane_wait(ctx, i) makes the
 * current lane wait for it.  nlk_ctx_sync waits for both lanes. */
int nlk_strip_lane(nlk_ctx *ctx, int lane, int reserve_sm);
int nlk_lane_record(nlk_ctx *ctx, int idx);
int nlk_lane_wait(nlk_ctx *ctx, int idx);
/* stage 1: zero accumulator rows, patch validity, block matching + k-NN for [gy0, gy1) */
int nlk_strip_search(nlk_ctx *ctx, int smooth, const float *d_in1, const float *d_prev0,
                     const float *d_bsic1, float sigma, struct nlkalman_params prms,
                     int gy0, int gy1, unsigned int *d_nbr, float *d_accw);
/* stage 2 (after the bitmaps of all rows are in d_nbr): processed-mask replay over the
 * whole grid, then the groups of this rank's rows */
int nlk_strip_filter(nlk_ctx *ctx);
/* stage 3 (after the border rows of the neighbours were added): pixel rows [row0, row1)
 * of the output */
int nlk_strip_normalize(nlk_ctx *ctx, float *d_out, int row0, int row1);

/* ---- peer-memory exchanges between the strips' GPUs (NVLink / NVSwitch, no collective library) ----
 * What crosses strips in the reference's patch loop (src/nlkalman.c:586-595 sharded by rows: the
 * processed-mask bitmaps of :597-600 / :930-931, the aggregation of :913-928 into rows beyond the
 * strip border, the output rows the next pass reads) is exchanged by the kernels themselves.  Every
 * rank keeps its exchange buffers in one "slab" (cudaMalloc, the same layout on all ranks, a
 * header of nlk_peer_header_bytes() first); the peers' slabs are mapped with CUDA IPC (or are plain
 * pointers when several strips live in one process) and handed to nlk_peer_bind.  All calls are
 * asynchronous on the context's stream; ranges are byte offsets into the slab.
 *   nlk_peer_push      copy [off, off+bytes) of the own slab to the same offset of the peers in
 *                      peer_mask, then store `value` into flag `slot` of those peers (slot < 0: no
 *                      flag).  side = 1: on a side stream with the copy engines (whole strips),
 *                      joined again by the next nlk_strip_normalize.
 *   nlk_peer_push_add  red.global.add the floats of the range into peer `peer` (accumulator halo
 *                      rows into their owner), then flag it.
 *   nlk_peer_signal    the flag alone.
 *   nlk_warp_rows_peer_dev  the bicubic warp with the rows it does not hold pulled from their owners.
 *   nlk_peer_wait      the stream waits until flag `slot` from every rank in src_mask is >= value
 *                      (sequence numbers, compared modulo 2^32).  A wait gives up after
 *                      NLK_PEER_TIMEOUT_MS (default 4000) and raises the error word read by
 *                      nlk_peer_error instead of hanging the GPU.
 * Flags: slots 0 .. 63, one word per (slot, source rank) in the receiver's slab. */
size_t nlk_peer_header_bytes(void);
int nlk_peer_slab_alloc(nlk_ctx *ctx, size_t bytes, void **d_slab);           /* zero-filled */
int nlk_peer_ipc_export(nlk_ctx *ctx, void *d_slab, unsigned char *handle64); /* cudaIpcMemHandle_t, 64 bytes */
int nlk_peer_ipc_import(nlk_ctx *ctx, const unsigned char *handle64, void **d_ptr);
int nlk_peer_bind(nlk_ctx *ctx, int rank, int nranks, void *const *slabs, size_t slab_bytes);
int nlk_peer_push(nlk_ctx *ctx, size_t off, size_t bytes, unsigned int peer_mask, int slot,
                  unsigned int value, int side);
int nlk_peer_push_add(nlk_ctx *ctx, size_t off, size_t bytes, int peer, int slot, unsigned int value);
int nlk_peer_signal(nlk_ctx *ctx, int slot, unsigned int value, unsigned int peer_mask);
int nlk_peer_wait(nlk_ctx *ctx, int slot, unsigned int value, unsigned int src_mask);
/* warp_bicubic of pixel rows [row0, row1) of the frame that sits at byte offset frame_off of every slab:
 * tap rows inside [local_lo, local_hi) are read from the own slab, any other row straight from its
 * owner's slab over NVLink (rank k owns rows [k * chunk_y, (k+1) * chunk_y), the last rank the rest).  The
 * caller has waited for the owners' "rows final" flags. */
int nlk_warp_rows_peer_dev(nlk_ctx *ctx, float *d_imw, size_t frame_off, const float *d_of, const float *d_msk,
                           int row0, int row1, int local_lo, int local_hi, int chunk_y);
int nlk_peer_error(nlk_ctx *ctx, unsigned int *code);   /* synchronises; 0 = no wait timed out */

/* ---- resident sequence recursion (what scripts/nlkalman-seq.sh does per frame) ------
 * The context keeps the previous frame's first and second filtering outputs in
 * opponent colour space.  One step = rgb2opp(noisy); if there is a previous frame:
 * warp(prev flt1), filter 1, warp(prev flt2), filter 2 with filter 1 as basic estimate
 * (reference src/main-flt.c:340-380); else the two spatial filterings of the first
 * frame (scripts/nlkalman-seq.sh:39-41).  f2.patch_sz == 0 skips the second filtering
 * (reference src/main-flt.c:130).  Outputs are RGB; any of them may be NULL. */
int nlk_seq_reset(nlk_ctx *ctx);
int nlk_seq_filter_dev(nlk_ctx *ctx, const float *d_noisy, const float *d_bflo,
                       const float *d_bocc, float sigma, struct nlkalman_params f1,
                       struct nlkalman_params f2, float *d_flt1_out, float *d_flt2_out);
/* same, with host buffers: copies in, runs, copies out, returns when done */
int nlk_seq_filter_host(nlk_ctx *ctx, const float *h_noisy, const float *h_bflo,
                        const float *h_bocc, float sigma, struct nlkalman_params f1,
                        struct nlkalman_params f2, float *h_flt1_out, float *h_flt2_out);

/* pipelined form of nlk_seq_filter_host for streaming a sequence: queues the frame and
 * returns.  Uploads, kernels and downloads run on their own streams with three staging
 * sets, so a frame uploads and another downloads while the two filterings work on two more
 * (at most three frames in flight: the call blocks until frame n-3 is complete).  The host
 * buffers of a frame (pinned memory, or the copies serialise) must stay untouched -- and its
 * outputs are complete only -- once nlk_seq_drain or three later submits have returned.
 * nlk_seq_filter_host = submit + drain. */
int nlk_seq_submit_host(nlk_ctx *ctx, const float *h_noisy, const float *h_bflo,
                        const float *h_bocc, float sigma, struct nlkalman_params f1,
                        struct nlkalman_params f2, float *h_flt1_out, float *h_flt2_out);
int nlk_seq_drain(nlk_ctx *ctx);
/* How nlk_seq_submit_host / nlk_seq_filter_host read their h_bocc argument:
 *   NLK_MASK_FLOAT      w*h floats, 0 = valid (the reference's in-memory form, src/main-flt.c:236-262);
 *   NLK_MASK_U8         w*h bytes, the samples of the 8-bit file the mask is read from
 *                       (scripts/nlkalman-seq.sh:70-73): a quarter of the upload;
 *   NLK_MASK_FROM_FLOW  h_bocc is ignored: the mask is built on the device from the divergence of the
 *                       frame's flow with threshold th, the script's plambda expression
 *                       (nlk_occlusion_dev) -- no mask upload at all. */
enum { NLK_MASK_FLOAT = 0, NLK_MASK_U8 = 1, NLK_MASK_FROM_FLOW = 2 };
int nlk_seq_set_mask_mode(nlk_ctx *ctx, int mode, float th);
/* The pipelined recursion runs the two filterings of a frame on two streams: the second
 * filtering of frame t (it needs flt1(t) and flt2(t-1)) overlaps the first filtering of frame
 * t+1 (it needs flt1(t) only), so that one pass's single-SM processed-mask replay and its
 * other thin phases run beside the other pass's kernels.  nlk_seq_submit_dev is that form for
 * device buffers: same arguments as nlk_seq_filter_dev, but d_flt2_out (and the context's
 * stream as seen by the caller) is complete only after nlk_seq_join (queues the wait on the
 * context's stream), nlk_seq_drain or nlk_ctx_sync (block the host); d_bflo / d_bocc of a frame
 * must stay untouched until then or until two later frames have been submitted.  Every other
 * entry point joins first, so the forms can be mixed. */
int nlk_seq_submit_dev(nlk_ctx *ctx, const float *d_noisy, const float *d_bflo,
                       const float *d_bocc, float sigma, struct nlkalman_params f1,
                       struct nlkalman_params f2, float *d_flt1_out, float *d_flt2_out);
int nlk_seq_join(nlk_ctx *ctx);

/* backward smoothing recursion (scripts/nlkalman-seq.sh:122-149): start from the last
 * filtered frame, then for each earlier frame t: warp(smoothed t+1 by the forward
 * flow), smooth (reference src/main-smo.c:198-213). */
int nlk_seq_smooth_start_dev(nlk_ctx *ctx, const float *d_last_rgb);
int nlk_seq_smooth_dev(nlk_ctx *ctx, const float *d_flt_rgb, const float *d_fflo,
                       const float *d_focc, float sigma, struct nlkalman_params s1,
                       float *d_smo_out);
int nlk_seq_smooth_start_host(nlk_ctx *ctx, const float *h_last_rgb);
int nlk_seq_smooth_host(nlk_ctx *ctx, const float *h_flt_rgb, const float *h_fflo,
                        const float *h_focc, float sigma, struct nlkalman_params s1,
                        float *h_smo_out);

/* ---- Dual TV-L1 optical flow, one scale ---------------------------------------
 * What the reference's Dual_TVL1_optic_flow does (lib/tvl1flow/tvl1flow_lib.c:93-280; called per
 * pyramid level by Dual_TVL1_optic_flow_multiscale, :345-477, the flow estimator in front of the
 * filter in scripts/nlkalman-seq.sh:60-65): `warps` times { bicubic warp of I1 and its centred
 * gradient by the current flow, then the thresholding / Chambolle dual iterations until the mean
 * squared update falls to epsilon^2 or 300 iterations }.  I0, I1: nx x ny single-channel images
 * (the level's, already normalised and smoothed); u1, u2: the flow, initial value in, result out.
 * iterations (host, [warps], may be NULL): iterations run by each warping step.  The arithmetic is
 * written with the reference's roundings (no fused multiply-add), so a level's flow is the
 * reference's bit for bit unless the float sum behind the stopping test lands on the other side of
 * epsilon^2 (its order differs; the reference's own OpenMP reduction is not ordered either). */
int nlk_tvl1_level_dev(nlk_ctx *ctx, const float *d_I0, const float *d_I1, float *d_u1, float *d_u2,
                       int nx, int ny, float tau, float lambda, float theta, int warps, float epsilon,
                       int *iterations);
int nlk_tvl1_level_host(nlk_ctx *ctx, const float *h_I0, const float *h_I1, float *h_u1, float *h_u2,
                        int nx, int ny, float tau, float lambda, float theta, int warps, float epsilon,
                        int *iterations);

/* ---- Dual TV-L1 optical flow, the whole estimator ----------------------------
 * What the reference's Dual_TVL1_optic_flow_multiscale does (lib/tvl1flow/tvl1flow_lib.c:345-477),
 * the library call behind the `tvl1flow` program of scripts/nlkalman-seq.sh:60-65, :124-129:
 * normalise both images to [0, 255] over their joint range (:305-337), Gaussian pre-smoothing
 * sigma 0.8 (:385-386, lib/tvl1flow/mask.c:216-330), `nscales` levels each zoomed out by `zfactor`
 * (Gaussian 0.6 sqrt(1/zfactor^2 - 1) + bicubic resampling, lib/tvl1flow/zoom.c:44-83), then from
 * the coarsest level down to `fscale` the level solver above, the flow zoomed in and multiplied by
 * 1 / zfactor between levels (zoom.c:91-113); levels finer than `fscale` only upsample.
 * I0, I1: nx x ny single-channel images; the flow u1, u2 (x and y displacement, nx x ny each) is
 * output only.  iterations (host, [nscales * warps], may be NULL): [s * warps + k] = iterations of
 * warping k at scale s (0 where a scale was skipped).  nlk_tvl1_flow_host returns the flow as two
 * planes, u1 then u2: the buffer the reference's driver writes (lib/tvl1flow/main.c:177).
 * nlk_tvl1_scales: the reference driver's cap on nscales (main.c:159-161), no level much below 16 px. */
int nlk_tvl1_scales(int nx, int ny, float zfactor, int nscales);
int nlk_tvl1_flow_dev(nlk_ctx *ctx, const float *d_I0, const float *d_I1, float *d_u1, float *d_u2,
                      int nx, int ny, float tau, float lambda, float theta, int nscales, int fscale,
                      float zfactor, int warps, float epsilon, int *iterations);
int nlk_tvl1_flow_host(nlk_ctx *ctx, const float *h_I0, const float *h_I1, float *h_flow, int nx, int ny,
                       float tau, float lambda, float theta, int nscales, int fscale, float zfactor,
                       int warps, float epsilon, int *iterations);

/* ---- the flow + mask step of the pipeline script, on frames resident in HBM --------------------
 * What scripts/nlkalman-seq.sh:60-72 (:124-137 for the smoother) does through files between two
 * filter invocations -- `tvl1flow FROM TO flow.flo NPROC 0 DW 0 0 FSCALE`, then the plambda mask
 * expression with threshold TH -- for two frames of the context's size and channel count, RGB,
 * interleaved, already on the device: luminance as the reference's program reads a colour file
 * (lib/iio/iio.c:3984-4003 on float samples), the estimator above, the flow interleaved [h][w][2]
 * as nlk_warp_dev takes it, the occlusion mask of nlk_occlusion_dev (d_occ may be NULL).
 * nlk_tvl1_default_params: the defaults of the reference's program (lib/tvl1flow/main.c:26-35), which
 * also replace out-of-range values there (:108-148); nscales is capped for the frame size inside. */
struct nlk_tvl1_params {
    float tau, lambda, theta;
    int nscales, fscale;
    float zfactor;
    int warps;
    float epsilon;
};
void nlk_tvl1_default_params(struct nlk_tvl1_params *p);
int nlk_flow_mask_dev(nlk_ctx *ctx, float *d_of, float *d_occ, const float *d_from_rgb, const float *d_to_rgb,
                      struct nlk_tvl1_params prms, float th);

/* ---- stage dumps for the parity tests (host arrays, any may be NULL) ------------------
 * Runs one pass on host images like nlkalman_filter_frame / nlkalman_smooth_frame and
 * also returns, per grid patch g = gy*gw + gx (gw = (w-psz)/step+1, step = psz/2):
 *   nk[g], np0[g]         candidates kept / with a valid previous patch
 *   knn_xy[g][kmax][2]    their (x, y) in sorted order;  knn_d[g][kmax] distances
 *   prev_p[g]             validity of the previous-frame patch at p
 *   active[g]             1 = processed (not skipped by the processed mask)
 *   vp[g]                 posterior variance sum of processed groups */
int nlk_pass_host_debug(nlk_ctx *ctx, int smooth, float *h_out, const float *h_in1,
                        const float *h_prev0, const float *h_bsic1, float sigma,
                        struct nlkalman_params prms, int kmax, int *nk, int *np0,
                        int *knn_xy, float *knn_d, unsigned char *prev_p,
                        unsigned char *active, float *vp);

/* batched orthonormal 2-D DCT-II (inverse = 0) / DCT-III (inverse = 1) of n tiles of
 * psz x psz on the device, host in/out, in place: the unit the reference's
 * dct_threads_forward / dct_threads_inverse compute (src/nlkalman.c:248, :307) */
int nlk_dct_host(nlk_ctx *ctx, float *h_tiles, int psz, int n, int inverse);

#ifdef __cplusplus
}
#endif
#endif /* NLKALMAN_B200_H */
