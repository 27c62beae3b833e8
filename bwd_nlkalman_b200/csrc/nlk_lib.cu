// libnlkalman_b200.so: the C ABI (include/nlkalman.h, include/nlkalman_b200.h) over the
// sm_100a kernels.  Single translation unit: the kernel headers are included here.
#include "../../include/nlkalman_b200.h"
#include "../../include/tvl1flow.h"

#include "nlk_common.cuh"
#include "nlk_prep.cuh"
#include "nlk_search.cuh"
#include "nlk_resolve.cuh"
#include "nlk_group.cuh"
#include "nlk_group_warp.cuh"
#include "nlk_peer.cuh"
#include "nlk_tvl1.cuh"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

using namespace nlk;

// ---- errors -----------------------------------------------------------------------------------

static thread_local std::string g_err;

static int set_err(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU_TRY(expr)                                                                              \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return set_err(NLK_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_),  \
                           __FILE__, __LINE__);                                                   \
    } while (0)

extern "C" const char *nlk_last_error(void) { return g_err.c_str(); }

// ---- constant tables --------------------------------------------------------------------------

static int upload_tables_dev()
{
    static float dct[MAX_PSZ + 1][MAX_PSZ * MAX_PSZ];
    static float win[MAX_PSZ + 1][MAX_PSZ * MAX_PSZ];
    static float inv[MAX_K + 1];
    const double pi = 3.14159265358979323846264338327950288;
    memset(dct, 0, sizeof dct);
    memset(win, 0, sizeof win);
    for (int n = 1; n <= MAX_PSZ; ++n) {
        for (int k = 0; k < n; ++k)
            for (int j = 0; j < n; ++j)
                dct[n][k * n + j] = (float)(sqrt((k ? 2.0 : 1.0) / n) * cos(pi * (j + 0.5) * k / n));
        // Gaussian window, reference src/nlkalman.c:367-368, :401-407, :413-416
        float w1[MAX_PSZ];
        const float N = (float)n;
        const float N2 = (float)((N - 1.) / 2.);
        for (int i = 0; i < n; ++i) {
            const float s = .4f;
            const float x = ((float)i - N2) / N2 / s;
            w1[i] = (float)exp(-.5 * x * x);
        }
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) win[n][i * n + j] = w1[i] * w1[j];
    }
    inv[0] = 0.f;
    for (int n = 1; n <= MAX_K; ++n) inv[n] = (float)(1. / (float)n);
    CU_TRY(cudaMemcpyToSymbol(c_dct, dct, sizeof dct));
    static float2 d8u[64], d8v[16];
    for (int i = 0; i < 64; ++i) d8u[i] = make_float2(dct[8][i], dct[8][i]);
    for (int j = 0; j < 4; ++j)
        for (int y = 0; y < 4; ++y) d8v[j * 4 + y] = make_float2(dct[8][(2 * j) * 8 + y], dct[8][(2 * j + 1) * 8 + y]);
    CU_TRY(cudaMemcpyToSymbol(c_dct8u, d8u, sizeof d8u));
    CU_TRY(cudaMemcpyToSymbol(c_dct8v, d8v, sizeof d8v));
    CU_TRY(cudaMemcpyToSymbol(c_win, win, sizeof win));
    CU_TRY(cudaMemcpyToSymbol(c_inv, inv, sizeof inv));
    return NLK_OK;
}

// ---- context ----------------------------------------------------------------------------------

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return NLK_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        CU_TRY(cudaMalloc(&p, bytes));
        cap = bytes;
        return NLK_OK;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// Per-pass scratch and the stream it is used on.  Two lanes: everything runs on lane 0 (the
// context's stream) except, in the pipelined recursion, the second filtering of a frame, which
// runs on lane 1 while lane 0 already works on the first filtering of the next frame (the two
// recursions only meet through flt1(t) -> flt2(t)).
struct Lane {
    cudaStream_t st = nullptr;
    DevBuf accw, valid, valid_tmp, cand, hdr, nbr, active, actflag, counters, xlist, rpack, q_warp;
    void release()
    {
        DevBuf *all[] = {&accw, &valid, &valid_tmp, &cand, &hdr, &nbr, &active, &actflag, &counters, &xlist,
                         &rpack, &q_warp};
        for (DevBuf *b : all) b->release();
    }
};

// an instantiated CUDA graph of one TV-L1 level (tvl1_level_graph) and what it was built for
struct Tvl1GraphEntry {
    const float *I0, *I1;
    float *u1, *u2, *base;
    int nx, ny, warps;
    float tau, lambda, theta, epsilon;
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    unsigned long long stamp;
};

enum { TVL1_LOOP_KERNEL = 1,   // one cooperative launch per warping step, grid barriers between half iterations
       TVL1_LOOP_GRAPH = 2,    // one CUDA graph per level, WHILE nodes around { k_tvl1_u, k_tvl1_p }
       TVL1_LOOP_HOST = 3 };   // the host queues batches of iterations and reads the error back

struct nlk_ctx {
    int w = 0, h = 0, ch = 0, device = 0, num_sms = 148;
    Lane lane[2];
    Lane *L = &lane[0];          // lane in use by the code being queued
    int reserve_sm = 0;          // pipelined recursion: group_filter leaves one SM to the other lane's mask_resolve
    bool b_pending = false;      // lane 1 has work that lane 0 has not waited for
    cudaEvent_t ev_a[2] = {nullptr, nullptr}, ev_b[2] = {nullptr, nullptr}, ev_join = nullptr;
    long long launches = 0;
    DevBuf dbg_dist, dbg_vp, tv_scratch, tv_pyr, tv_frames;
    float *tv_herr = nullptr;       // pinned word the TV-L1 level solver reads its stopping error back into
    std::vector<Tvl1GraphEntry> tv_graphs;         // instantiated level graphs (tvl1_level_graph)
    cudaStream_t tv_st2 = nullptr;  // captures the loop bodies
    int tv_loop = 0;                // how a warping step's iterations run: 0 not decided, TVL1_LOOP_* below
    int tv_coop_blocks = 0;         // co-resident blocks of k_tvl1_iterate on this device
    unsigned long long tv_stamp = 0;
    // host-call staging
    DevBuf s_in1, s_prev0, s_bsic, s_out, s_of, s_msk;
    // sequence state (opponent colour space)
    DevBuf q_noisy[2], q_flt1[2], q_flt2[2], q_smo[2], q_tmp;
    int q_cur = 0, q_have_prev = 0, q_have_flt2 = 0;
    long long q_frames = 0;
    int q_smo_cur = 0, q_have_smo = 0;
    // pipelined host-buffer recursion (nlk_seq_submit_host): copies on their own streams,
    // three staging sets so that one frame uploads and one downloads while the lanes work on two more
    cudaStream_t st_h2d = nullptr, st_d2h = nullptr;
    static constexpr int PIPE_SETS = 3;   // frames in flight between the host buffers and the lanes
    DevBuf p_in[PIPE_SETS], p_of[PIPE_SETS], p_msk[PIPE_SETS], p_o1[PIPE_SETS], p_o2[PIPE_SETS];
    cudaEvent_t ev_up[PIPE_SETS] = {}, ev_o1[PIPE_SETS] = {}, ev_o2[PIPE_SETS] = {}, ev_done[PIPE_SETS] = {};
    long long p_frames = 0;
    // strip-sharded pass in flight (nlk_strip_search .. nlk_strip_normalize)
    PassParams strip_Ps[2];
    bool strip_opens[2] = {false, false};
    int strip_lane = 0;                        // lane the strip / peer / row-range calls queue on (nlk_strip_lane)
    int strip_reserve = 0;                     // SMs group_filter leaves to the other lane's mask_resolve
    cudaEvent_t ev_lane[8] = {};               // cross-lane dependencies of the caller's schedule
    // peer-memory exchanges between the strips' GPUs (nlk_peer_*)
    PeerTable peer;
    bool peer_on = false;
    size_t slab_bytes = 0;
    std::vector<void *> ipc_opened;
    cudaStream_t st_sides[2] = {nullptr, nullptr};   // whole-strip pushes (copy engines) beside the next pass, per lane
    cudaEvent_t ev_side_forks[2] = {nullptr, nullptr}, ev_side_dones[2] = {nullptr, nullptr};
    bool side_pendings[2] = {false, false};
    unsigned int cnt_rot = 0;
    // optional per-kernel timing with CUDA events on the stream of the lane in use
    bool prof = false;
    int prof_kind = NLK_PASS_OTHER;
    struct ProfRec { int kid, kind; cudaEvent_t a, b; };
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;
    // processed / grid patches per pass while profiling (alpha of SURVEY.md section 8(d))
    struct AlphaRec { int kind, G, idx; };
    std::vector<AlphaRec> alpha_recs;
    int *h_alpha = nullptr;
    static constexpr int ALPHA_CAP = 8192;
    // how nlk_seq_submit_host reads its mask argument (nlk_seq_set_mask_mode)
    int mask_mode = NLK_MASK_FLOAT;
    float mask_th = 0.f;
    DevBuf p_msk8[PIPE_SETS];
    size_t img_bytes() const { return (size_t)w * h * ch * sizeof(float); }
};

static int ctx_use(nlk_ctx *c)
{
    if (!c) return set_err(NLK_ERR_PARAM, "null context");
    CU_TRY(cudaSetDevice(c->device));
    return NLK_OK;
}

struct ProfScope {
    nlk_ctx *c;
    cudaEvent_t a = nullptr, b = nullptr;
    int kid;
    static cudaEvent_t get(nlk_ctx *c)
    {
        if (!c->prof_pool.empty()) {
            cudaEvent_t e = c->prof_pool.back();
            c->prof_pool.pop_back();
            return e;
        }
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        return e;
    }
    ProfScope(nlk_ctx *c_, int kid_) : c(c_), kid(kid_)
    {
        if (!c->prof) return;
        a = get(c);
        b = get(c);
        cudaEventRecord(a, c->L->st);
    }
    ~ProfScope()
    {
        if (!a) return;
        cudaEventRecord(b, c->L->st);
        c->prof_recs.push_back({kid, c->prof_kind, a, b});
    }
};

extern "C" int nlk_ctx_profile(nlk_ctx *c, int enable)
{
    if (!c) return set_err(NLK_ERR_PARAM, "null context");
    c->prof = enable != 0;
    return NLK_OK;
}

extern "C" int nlk_ctx_profile_collect(nlk_ctx *c, double *ms_sum, int *count)
{
    if (int r = nlk_ctx_sync(c)) return r;
    for (int i = 0; i < NLK_KERNEL_COUNT * NLK_PASS_KINDS; ++i) { ms_sum[i] = 0; count[i] = 0; }
    for (auto &r : c->prof_recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            ms_sum[r.kid * NLK_PASS_KINDS + r.kind] += ms;
            count[r.kid * NLK_PASS_KINDS + r.kind] += 1;
        }
        c->prof_pool.push_back(r.a);
        c->prof_pool.push_back(r.b);
    }
    c->prof_recs.clear();
    return NLK_OK;
}

extern "C" int nlk_ctx_profile_alpha(nlk_ctx *c, double *active_sum, double *grid_sum)
{
    if (int r = nlk_ctx_sync(c)) return r;
    for (int i = 0; i < NLK_PASS_KINDS; ++i) { active_sum[i] = 0; grid_sum[i] = 0; }
    for (auto &a : c->alpha_recs) {
        active_sum[a.kind] += c->h_alpha[a.idx];
        grid_sum[a.kind] += a.G;
    }
    c->alpha_recs.clear();
    return NLK_OK;
}

extern "C" int nlk_seq_set_mask_mode(nlk_ctx *c, int mode, float th)
{
    if (!c) return set_err(NLK_ERR_PARAM, "null context");
    if (mode != NLK_MASK_FLOAT && mode != NLK_MASK_U8 && mode != NLK_MASK_FROM_FLOW)
        return set_err(NLK_ERR_PARAM, "mask mode %d", mode);
    c->mask_mode = mode;
    c->mask_th = th;
    return NLK_OK;
}

extern "C" int nlk_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return set_err(NLK_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return n;
}

extern "C" nlk_ctx *nlk_ctx_create(int w, int h, int ch, int device)
{
    if (w <= 0 || h <= 0 || ch <= 0 || ch > MAX_CH || w > 32767 || h > 32767) {
        set_err(NLK_ERR_PARAM, "unsupported frame %dx%dx%d (1..%d channels, sides < 32768)", w, h, ch, MAX_CH);
        return nullptr;
    }
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        set_err(NLK_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
        return nullptr;
    }
    nlk_ctx *c = new nlk_ctx();
    c->w = w; c->h = h; c->ch = ch; c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->num_sms = prop.multiProcessorCount;
    bool ok = upload_tables_dev() == NLK_OK;
    for (int i = 0; i < 2 && ok; ++i) {
        ok = cudaStreamCreateWithFlags(&c->lane[i].st, cudaStreamNonBlocking) == cudaSuccess &&
             c->lane[i].counters.ensure(64) == NLK_OK &&
             cudaEventCreateWithFlags(&c->ev_a[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&c->ev_b[i], cudaEventDisableTiming) == cudaSuccess;
    }
    ok = ok && cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        if (g_err.empty()) set_err(NLK_ERR_CUDA, "context creation failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete c;
        return nullptr;
    }
    return c;
}

static void tvl1_graphs_release(nlk_ctx *c)
{
    for (Tvl1GraphEntry &E : c->tv_graphs) {
        if (E.exec) cudaGraphExecDestroy(E.exec);
        if (E.graph) cudaGraphDestroy(E.graph);
    }
    c->tv_graphs.clear();
    if (c->tv_st2) { cudaStreamDestroy(c->tv_st2); c->tv_st2 = nullptr; }
}

extern "C" void nlk_ctx_destroy(nlk_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    for (int i = 0; i < 2; ++i) if (c->lane[i].st) cudaStreamSynchronize(c->lane[i].st);
    if (c->st_d2h) cudaStreamSynchronize(c->st_d2h);
    if (c->st_h2d) cudaStreamSynchronize(c->st_h2d);
    DevBuf *all[] = {&c->tv_scratch, &c->tv_pyr, &c->tv_frames, &c->dbg_dist, &c->dbg_vp, &c->s_in1, &c->s_prev0, &c->s_bsic, &c->s_out, &c->s_of, &c->s_msk,
                     &c->q_noisy[0], &c->q_noisy[1], &c->q_flt1[0], &c->q_flt1[1], &c->q_flt2[0], &c->q_flt2[1],
                     &c->q_smo[0], &c->q_smo[1], &c->q_tmp};
    for (DevBuf *b : all) b->release();
    for (int i = 0; i < 2; ++i) {
        c->lane[i].release();
        if (c->ev_a[i]) cudaEventDestroy(c->ev_a[i]);
        if (c->ev_b[i]) cudaEventDestroy(c->ev_b[i]);
    }
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    for (int i = 0; i < nlk_ctx::PIPE_SETS; ++i) {
        c->p_in[i].release(); c->p_of[i].release(); c->p_msk[i].release(); c->p_o1[i].release(); c->p_o2[i].release();
        cudaEvent_t *ev[] = {&c->ev_up[i], &c->ev_o1[i], &c->ev_o2[i], &c->ev_done[i]};
        for (cudaEvent_t *e : ev) if (*e) cudaEventDestroy(*e);
    }
    if (c->st_h2d) cudaStreamDestroy(c->st_h2d);
    if (c->st_d2h) cudaStreamDestroy(c->st_d2h);
    for (int i = 0; i < 2; ++i) {
        if (c->st_sides[i]) { cudaStreamSynchronize(c->st_sides[i]); cudaStreamDestroy(c->st_sides[i]); }
        if (c->ev_side_forks[i]) cudaEventDestroy(c->ev_side_forks[i]);
        if (c->ev_side_dones[i]) cudaEventDestroy(c->ev_side_dones[i]);
    }
    for (cudaEvent_t e : c->ev_lane) if (e) cudaEventDestroy(e);
    for (void *q : c->ipc_opened) cudaIpcCloseMemHandle(q);
    if (c->h_alpha) cudaFreeHost(c->h_alpha);
    if (c->tv_herr) cudaFreeHost(c->tv_herr);
    tvl1_graphs_release(c);
    for (int i = 0; i < nlk_ctx::PIPE_SETS; ++i) c->p_msk8[i].release();
    for (auto &r : c->prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (cudaEvent_t e : c->prof_pool) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) if (c->lane[i].st) cudaStreamDestroy(c->lane[i].st);
    delete c;
}

// lane 0 (the context's stream) waits for whatever the pipelined recursion left on lane 1
static int lanes_join(nlk_ctx *c)
{
    c->L = &c->lane[0];
    if (!c->b_pending) return NLK_OK;
    CU_TRY(cudaEventRecord(c->ev_join, c->lane[1].st));
    CU_TRY(cudaStreamWaitEvent(c->lane[0].st, c->ev_join, 0));
    c->b_pending = false;
    return NLK_OK;
}

static int enter(nlk_ctx *c)   // public entry points outside the pipelined recursion
{
    if (int r = ctx_use(c)) return r;
    return lanes_join(c);
}

// entry points of the strip-sharded pass: queued on the lane chosen with nlk_strip_lane.  Lane 1 is
// used by callers that run the two filterings of a frame as two pipelines (as the single-GPU
// recursion does); their cross-lane order is the caller's (nlk_lane_record / nlk_lane_wait).
static int enter_strip(nlk_ctx *c)
{
    if (!c) return set_err(NLK_ERR_PARAM, "null context");
    if (c->strip_lane == 0) return enter(c);
    if (int r = ctx_use(c)) return r;
    c->L = &c->lane[1];
    return NLK_OK;
}

extern "C" int nlk_strip_lane(nlk_ctx *c, int lane, int reserve_sm)
{
    if (!c || lane < 0 || lane > 1 || reserve_sm < 0 || reserve_sm > 8) return set_err(NLK_ERR_PARAM, "bad lane request");
    c->strip_lane = lane;
    c->strip_reserve = reserve_sm;
    return NLK_OK;
}

extern "C" int nlk_lane_record(nlk_ctx *c, int idx)
{
    if (int r = enter_strip(c)) return r;
    if (idx < 0 || idx >= 8) return set_err(NLK_ERR_PARAM, "lane event %d", idx);
    if (!c->ev_lane[idx]) CU_TRY(cudaEventCreateWithFlags(&c->ev_lane[idx], cudaEventDisableTiming));
    CU_TRY(cudaEventRecord(c->ev_lane[idx], c->L->st));
    return NLK_OK;
}

extern "C" int nlk_lane_wait(nlk_ctx *c, int idx)
{
    if (int r = enter_strip(c)) return r;
    if (idx < 0 || idx >= 8) return set_err(NLK_ERR_PARAM, "lane event %d", idx);
    if (!c->ev_lane[idx]) return NLK_OK;       // never recorded: nothing to wait for
    CU_TRY(cudaStreamWaitEvent(c->L->st, c->ev_lane[idx], 0));
    return NLK_OK;
}

extern "C" int nlk_ctx_sync(nlk_ctx *c)
{
    if (int r = ctx_use(c)) return r;
    if (int r = lanes_join(c)) return r;
    CU_TRY(cudaStreamSynchronize(c->lane[0].st));
    CU_TRY(cudaStreamSynchronize(c->lane[1].st));
    if (c->st_d2h) CU_TRY(cudaStreamSynchronize(c->st_d2h));
    for (int i = 0; i < 2; ++i) if (c->st_sides[i]) CU_TRY(cudaStreamSynchronize(c->st_sides[i]));
    return NLK_OK;
}

extern "C" long long nlk_ctx_launch_count(const nlk_ctx *c) { return c ? c->launches : 0; }
extern "C" void *nlk_ctx_stream(nlk_ctx *c) { return c ? (void *)c->lane[0].st : nullptr; }

extern "C" void *nlk_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) {
        set_err(NLK_ERR_CUDA, "cudaMallocHost(%zu) failed", bytes);
        return nullptr;
    }
    return p;
}
extern "C" void nlk_host_free(void *p) { if (p) cudaFreeHost(p); }

extern "C" void *nlk_dev_alloc(nlk_ctx *c, size_t bytes)
{
    if (ctx_use(c)) return nullptr;
    void *p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) {
        set_err(NLK_ERR_CUDA, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return p;
}

extern "C" void nlk_dev_free(nlk_ctx *c, void *d_ptr)
{
    if (!d_ptr || ctx_use(c)) return;
    cudaStreamSynchronize(c->L->st);
    cudaFree(d_ptr);
}

// a point of the context's stream another host thread can wait for (file writers, staging reuse)
extern "C" void *nlk_marker_record(nlk_ctx *c)
{
    if (enter(c)) return nullptr;
    cudaEvent_t e = nullptr;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming | cudaEventBlockingSync) != cudaSuccess ||
        cudaEventRecord(e, c->L->st) != cudaSuccess) {
        set_err(NLK_ERR_CUDA, "marker: %s", cudaGetErrorString(cudaGetLastError()));
        if (e) cudaEventDestroy(e);
        return nullptr;
    }
    return e;
}

extern "C" int nlk_marker_wait(void *marker)
{
    if (!marker) return set_err(NLK_ERR_PARAM, "null marker");
    cudaEvent_t e = static_cast<cudaEvent_t>(marker);
    const cudaError_t r = cudaEventSynchronize(e);
    cudaEventDestroy(e);
    if (r != cudaSuccess) return set_err(NLK_ERR_CUDA, "marker wait: %s", cudaGetErrorString(r));
    return NLK_OK;
}

extern "C" int nlk_upload(nlk_ctx *c, void *d_dst, const void *h_src, size_t bytes)
{
    if (int r = enter(c)) return r;
    CU_TRY(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, c->L->st));
    return NLK_OK;
}

extern "C" int nlk_download(nlk_ctx *c, void *h_dst, const void *d_src, size_t bytes)
{
    if (int r = enter(c)) return r;
    CU_TRY(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, c->L->st));
    return NLK_OK;
}

extern "C" int nlk_copy_dev(nlk_ctx *c, void *d_dst, const void *d_src, size_t bytes)
{
    if (int r = enter(c)) return r;
    CU_TRY(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, c->L->st));
    return NLK_OK;
}

// ---- one pass ---------------------------------------------------------------------------------

static int check_launch(nlk_ctx *c, int n, const char *what)
{
    if (n < 0) return set_err(NLK_ERR_PARAM, "%s: configuration not supported by the kernels", what);
    c->launches += n;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_err(NLK_ERR_CUDA, "%s launch failed: %s", what, cudaGetErrorString(e));
    return NLK_OK;
}

// fills P for one pass and sizes the scratch it points to.  The neighbour bitmaps and the
// accumulator are the context's own unless the caller supplies them (strip-sharded pass: they
// are exchanged between GPUs).
static int pass_setup(nlk_ctx *c, PassParams &P, int smooth, float *d_out, const float *d_in1,
                      const float *d_prev0, const float *d_bsic1, float sigma, const nlkalman_params &pr,
                      bool debug, uint32_t *d_nbr_ext, float *d_accw_ext)
{
    const int w = c->w, h = c->h, ch = c->ch;
    const int psz = pr.patch_sz;
    if (psz < 2 || psz > MAX_PSZ)
        return set_err(NLK_ERR_PARAM, "patch size %d not supported (2..%d)", psz, MAX_PSZ);
    if (pr.search_sz_x < 0 || pr.search_sz_t < 0 || pr.npatches_x < 0 || pr.npatches_t < 0 ||
        pr.npatches_tagg < 0)
        return set_err(NLK_ERR_PARAM, "negative parameter: call nlkalman_default_params first");
    // The smoother's single-patch branch (k <= 1 with a valid previous patch, reference
    // src/nlkalman.c:1699-1730) aggregates at uninitialised coordinates in the reference: there is
    // no defined result to reproduce, so the configuration is refused.
    if (smooth && d_prev0 && pr.npatches_t <= 1)
        return set_err(NLK_ERR_PARAM, "smoother with npatches_t = %d <= 1 is undefined in the reference "
                                      "(src/nlkalman.c:1699-1730) and not supported", pr.npatches_t);

    memset(&P, 0, sizeof P);
    P.w = w; P.h = h; P.ch = ch; P.psz = psz; P.step = psz / 2;
    const bool fits = (w >= psz && h >= psz);
    P.gw = fits ? (w - psz) / P.step + 1 : 0;
    P.gh = fits ? (h - psz) / P.step + 1 : 0;
    P.G = P.gw * P.gh;
    P.gy0 = 0; P.gy1 = P.gh;
    P.smooth = smooth;
    P.r_x = pr.search_sz_x; P.r_t = pr.search_sz_t;
    P.k_x = pr.npatches_x; P.k_t = pr.npatches_t; P.tagg = pr.npatches_tagg;
    P.sigma2 = sigma * sigma; P.beta_x = pr.beta_x; P.beta_t = pr.beta_t;
    P.has_prev = d_prev0 != nullptr; P.has_bsic = d_bsic1 != nullptr;
    P.src = d_bsic1 ? d_bsic1 : d_in1;
    P.in1 = d_in1; P.prev0 = d_prev0;
    P.vw = w - psz + 1; P.vh = h - psz + 1;
    const int rmax = smooth ? P.r_t : (P.r_t > P.r_x ? P.r_t : P.r_x);
    const int ncand = (2 * rmax + 1) * (2 * rmax + 1);
    if (ncand > SEARCH_MAX_NPAD)
        return set_err(NLK_ERR_PARAM, "search radius %d not supported (window > %d candidates)", rmax, SEARCH_MAX_NPAD);
    int kmax = P.k_x > P.k_t ? P.k_x : P.k_t;
    if (kmax > ncand) kmax = ncand;
    if (kmax < 1) kmax = 1;
    if (kmax > MAX_K) return set_err(NLK_ERR_PARAM, "more than %d patches per group", MAX_K);
    P.kstride = kmax;
    P.R = rmax / P.step;
    P.nbw = ((2 * P.R + 1) * (2 * P.R + 1) + 31) / 32;

    const size_t npix = (size_t)w * h;
    const size_t G = (size_t)(P.G > 0 ? P.G : 1);
    if (!d_accw_ext) if (int r = c->L->accw.ensure(npix * (ch + 1) * 4)) return r;
    if (int r = c->L->cand.ensure(G * kmax * 4)) return r;
    if (int r = c->L->hdr.ensure(G * sizeof(GroupHdr))) return r;
    if (!d_nbr_ext) if (int r = c->L->nbr.ensure(G * P.nbw * 4)) return r;
    if (int r = c->L->active.ensure(G * 4)) return r;
    if (int r = c->L->actflag.ensure(G)) return r;
    if (d_prev0 && P.G > 0) {
        if (int r = c->L->valid.ensure((size_t)P.vw * P.vh)) return r;
        if (int r = c->L->valid_tmp.ensure((size_t)P.vw * h)) return r;
        P.valid = c->L->valid.as<uint8_t>();
    }
    P.accw = d_accw_ext ? d_accw_ext : c->L->accw.as<float>();
    P.cand = c->L->cand.as<uint32_t>();
    P.hdr = c->L->hdr.as<GroupHdr>();
    P.nbr = d_nbr_ext ? d_nbr_ext : c->L->nbr.as<uint32_t>();
    P.active = c->L->active.as<int>();
    P.actflag = c->L->actflag.as<uint8_t>();
    P.nactive = c->L->counters.as<int>();
    P.any_nbr = c->L->counters.as<int>() + 1;
    P.work = c->L->counters.as<int>() + 2;
    P.xcount = c->L->counters.as<int>() + 3;
    // worklist of the one-block-per-patch search launch: at most every grid patch
    if (int r = c->L->xlist.ensure((G + 16) * 4)) return r;
    P.xlist = c->L->xlist.as<int>();
    P.out = d_out;
    if (debug) {
        if (int r = c->dbg_dist.ensure(G * kmax * 4)) return r;
        if (int r = c->dbg_vp.ensure(G * 4)) return r;
        P.dbg_dist = c->dbg_dist.as<float>();
        P.dbg_vp = c->dbg_vp.as<float>();
        CU_TRY(cudaMemsetAsync(P.dbg_dist, 0, G * kmax * 4, c->L->st));
        CU_TRY(cudaMemsetAsync(P.dbg_vp, 0, G * 4, c->L->st));
        CU_TRY(cudaMemsetAsync(P.cand, 0xff, G * kmax * 4, c->L->st));
    }
    return NLK_OK;
}

static int pass_kind(const PassParams &P)
{
    return P.smooth ? NLK_PASS_SMO : (P.has_bsic ? (P.has_prev ? NLK_PASS_FLT2_T : NLK_PASS_FLT2_X)
                                                  : (P.has_prev ? NLK_PASS_FLT1_T : NLK_PASS_FLT1_X));
}

struct KindScope {
    nlk_ctx *c; int saved;
    KindScope(nlk_ctx *c_, int k) : c(c_), saved(c_->prof_kind) { c->prof_kind = k; }
    ~KindScope() { c->prof_kind = saved; }
};

// pixel rows a strip of grid rows [gy0, gy1) reads (source, previous frame) and accumulates into
static void strip_rows(const PassParams &P, int *ey0, int *ey1)
{
    const int rmax = P.smooth ? P.r_t : (P.r_t > P.r_x ? P.r_t : P.r_x);
    int a = P.gy0 * P.step - rmax, b = (P.gy1 - 1) * P.step + rmax + P.psz;
    if (a < 0) a = 0;
    if (b > P.h) b = P.h;
    if (P.gy1 <= P.gy0) a = b = 0;
    *ey0 = a; *ey1 = b;
}

// zero the accumulator rows, validity map and block matching for grid rows [P.gy0, P.gy1)
static int pass_search(nlk_ctx *c, PassParams &P)
{
    int ey0, ey1;
    strip_rows(P, &ey0, &ey1);
    {
        ProfScope ps(c, NLK_K_MEMSET);
        const size_t rowb = (size_t)P.w * (P.ch + 1) * 4;
        if (ey1 > ey0) CU_TRY(cudaMemsetAsync(reinterpret_cast<char *>(P.accw) + ey0 * rowb, 0, (ey1 - ey0) * rowb, c->L->st));
        CU_TRY(cudaMemsetAsync(c->L->counters.p, 0, 64, c->L->st));
    }
    if (P.G <= 0 || P.gy1 <= P.gy0) return NLK_OK;
    if (P.prev0) {
        ProfScope ps(c, NLK_K_VALID);
        if (int r = check_launch(c, launch_valid_map(c->L->valid.as<uint8_t>(), c->L->valid_tmp.as<uint8_t>(), P.prev0,
                                                     P.w, P.h, P.ch, P.psz, ey0, ey1 - P.psz + 1, c->L->st), "valid_map")) return r;
    }
    ProfScope ps(c, NLK_K_SEARCH);
    return check_launch(c, launch_search(P, c->L->st), "search_knn");
}

// processed-mask replay over the WHOLE grid (needs every row's bitmaps), then the groups of
// rows [P.gy0, P.gy1)
static int pass_filter(nlk_ctx *c, const PassParams &P, bool strip)
{
    if (P.G <= 0) return NLK_OK;
    {
        ProfScope ps(c, NLK_K_RESOLVE);
        if (strip) {
            // the flag search_knn raises only covers this rank's rows: decide statically
            k_set_flag<<<1, 1, 0, c->L->st>>>(P.any_nbr, (P.tagg > 1 && P.R >= 1) ? 1 : 0);
            c->launches += 1;
        }
        if (int r = c->L->rpack.ensure(resolve_pack_bytes(P.gw, P.gh, P.R))) return r;
        if (int r = check_launch(c, launch_resolve(P, c->L->rpack.as<unsigned int>(), c->L->st), "mask_resolve")) return r;
        if (strip) {
            k_active_range<<<1, 32, 0, c->L->st>>>(P);
            if (int r = check_launch(c, 1, "active_range")) return r;
        }
    }
    if (P.gy1 <= P.gy0) return NLK_OK;
    ProfScope ps(c, NLK_K_GROUP);
    return check_launch(c, launch_group_filter(P, c->num_sms - c->reserve_sm, c->L->st), "group_filter");
}

static int pass_normalize(nlk_ctx *c, const PassParams &P, int row0, int row1)
{
    ProfScope ps(c, NLK_K_NORMALIZE);
    return check_launch(c, launch_normalize(P, row0, row1, c->L->st), "normalize");
}

static int run_pass(nlk_ctx *c, int smooth, float *d_out, const float *d_in1, const float *d_prev0,
                    const float *d_bsic1, float sigma, const nlkalman_params &pr, bool debug)
{
    PassParams P;
    if (int r = pass_setup(c, P, smooth, d_out, d_in1, d_prev0, d_bsic1, sigma, pr, debug, nullptr, nullptr)) return r;
    KindScope ks(c, pass_kind(P));
    if (int r = pass_search(c, P)) return r;
    if (int r = pass_filter(c, P, false)) return r;
    if (c->prof && P.G > 0) {
        if (!c->h_alpha) CU_TRY(cudaMallocHost(&c->h_alpha, nlk_ctx::ALPHA_CAP * sizeof(int)));
        const int idx = (int)c->alpha_recs.size();
        if (idx < nlk_ctx::ALPHA_CAP) {
            CU_TRY(cudaMemcpyAsync(c->h_alpha + idx, P.nactive, sizeof(int), cudaMemcpyDeviceToHost, c->L->st));
            c->alpha_recs.push_back({pass_kind(P), P.G, idx});
        }
    }
    return pass_normalize(c, P, 0, P.h);
}

extern "C" int nlk_pass_dev(nlk_ctx *c, int smooth, float *d_out, const float *d_in1,
                            const float *d_prev0, const float *d_bsic1, float sigma,
                            struct nlkalman_params prms)
{
    if (int r = enter(c)) return r;
    return run_pass(c, smooth, d_out, d_in1, d_prev0, d_bsic1, sigma, prms, false);
}

// queued on the lane in use (c->L); the extern forms below first make lane 0 current
static int colour_dev(nlk_ctx *c, float *d_dst, const float *d_src, int inverse);
static int warp_dev(nlk_ctx *c, float *d_imw, const float *d_im, const float *d_of, const float *d_msk);

extern "C" int nlk_rgb2opp_dev(nlk_ctx *c, float *d_dst, const float *d_src)
{
    if (int r = enter(c)) return r;
    return colour_dev(c, d_dst, d_src, 0);
}

extern "C" int nlk_opp2rgb_dev(nlk_ctx *c, float *d_dst, const float *d_src)
{
    if (int r = enter(c)) return r;
    return colour_dev(c, d_dst, d_src, 1);
}

extern "C" int nlk_warp_dev(nlk_ctx *c, float *d_imw, const float *d_im, const float *d_of,
                            const float *d_msk)
{
    if (int r = enter(c)) return r;
    return warp_dev(c, d_imw, d_im, d_of, d_msk);
}

static int colour_dev(nlk_ctx *c, float *d_dst, const float *d_src, int inverse)
{
    if (c->ch != 3) { // reference src/nlkalman.c:94, :114: no-op unless 3 channels
        if (d_dst != d_src) CU_TRY(cudaMemcpyAsync(d_dst, d_src, c->img_bytes(), cudaMemcpyDeviceToDevice, c->L->st));
        return NLK_OK;
    }
    ProfScope ps(c, NLK_K_COLOUR);
    return check_launch(c, launch_rgb2opp_copy(d_dst, d_src, (long)c->w * c->h, inverse, c->L->st),
                        inverse ? "opp2rgb" : "rgb2opp");
}

static int warp_dev(nlk_ctx *c, float *d_imw, const float *d_im, const float *d_of, const float *d_msk)
{
    ProfScope ps(c, NLK_K_WARP);
    return check_launch(c, launch_warp(d_imw, d_im, d_of, d_msk, c->w, c->h, c->ch, 0, c->h, c->L->st), "warp_bicubic");
}


static int stage_in(nlk_ctx *c, DevBuf &b, const float *h, size_t bytes, const float **d);

// ---- occlusion mask from the flow -----------------------------------------------

extern "C" int nlk_occlusion_dev(nlk_ctx *c, float *d_occ, const float *d_of, float th)
{
    if (int r = enter(c)) return r;
    if (!d_occ || !d_of) return set_err(NLK_ERR_PARAM, "null flow or mask");
    ProfScope ps(c, NLK_K_WARP);
    return check_launch(c, launch_occlusion(d_occ, d_of, c->w, c->h, th, c->L->st), "occlusion");
}

extern "C" int nlk_occlusion_host(nlk_ctx *c, float *h_occ, const float *h_of, float th)
{
    if (int r = enter(c)) return r;
    const size_t npix = (size_t)c->w * c->h;
    const float *d_of;
    if (int r = stage_in(c, c->s_of, h_of, npix * 2 * 4, &d_of)) return r;
    if (!d_of) return set_err(NLK_ERR_PARAM, "no flow");
    if (int r = c->s_msk.ensure(npix * 4)) return r;
    if (int r = nlk_occlusion_dev(c, c->s_msk.as<float>(), d_of, th)) return r;
    CU_TRY(cudaMemcpyAsync(h_occ, c->s_msk.p, npix * 4, cudaMemcpyDeviceToHost, c->L->st));
    CU_TRY(cudaStreamSynchronize(c->L->st));
    return NLK_OK;
}

// ---- row ranges and the strip-sharded pass --------------------------------------

static int rows_ok(nlk_ctx *c, int row0, int row1)
{
    if (row0 < 0 || row1 > c->h || row1 < row0) return set_err(NLK_ERR_PARAM, "row range [%d,%d) outside the frame", row0, row1);
    return NLK_OK;
}

extern "C" int nlk_colour_rows_dev(nlk_ctx *c, float *d_dst, const float *d_src, int inverse, int row0, int row1)
{
    if (int r = enter_strip(c)) return r;
    if (int r = rows_ok(c, row0, row1)) return r;
    const size_t off = (size_t)row0 * c->w * c->ch;
    const long npix = (long)(row1 - row0) * c->w;
    if (npix == 0) return NLK_OK;
    if (c->ch != 3) {
        if (d_dst != d_src) CU_TRY(cudaMemcpyAsync(d_dst + off, d_src + off, (size_t)npix * c->ch * 4, cudaMemcpyDeviceToDevice, c->L->st));
        return NLK_OK;
    }
    ProfScope ps(c, NLK_K_COLOUR);
    return check_launch(c, launch_rgb2opp_copy(d_dst + off, d_src + off, npix, inverse, c->L->st), "colour_rows");
}

extern "C" int nlk_warp_rows_dev(nlk_ctx *c, float *d_imw, const float *d_im, const float *d_of,
                                 const float *d_msk, int row0, int row1)
{
    if (int r = enter_strip(c)) return r;
    if (int r = rows_ok(c, row0, row1)) return r;
    ProfScope ps(c, NLK_K_WARP);
    return check_launch(c, launch_warp(d_imw, d_im, d_of, d_msk, c->w, c->h, c->ch, row0, row1, c->L->st), "warp_rows");
}

extern "C" int nlk_strip_plan(int w, int h, int smooth, struct nlkalman_params pr, int nranks, int rank,
                              struct nlk_strip_plan *out)
{
    if (!out || nranks < 1 || rank < 0 || rank >= nranks) return set_err(NLK_ERR_PARAM, "bad strip request");
    const int psz = pr.patch_sz, step = psz / 2;
    if (psz < 2 || w < psz || h < psz) return set_err(NLK_ERR_PARAM, "frame smaller than a patch");
    const int rmax = smooth ? pr.search_sz_t : (pr.search_sz_t > pr.search_sz_x ? pr.search_sz_t : pr.search_sz_x);
    const int gw = (w - psz) / step + 1, gh = (h - psz) / step + 1;
    const int R = rmax / step;
    memset(out, 0, sizeof *out);
    out->gw = gw; out->gh = gh; out->nbw = ((2 * R + 1) * (2 * R + 1) + 31) / 32;
    // equal chunks of q = ceil(gh / nranks) grid rows (the last strip takes what is left): a rank's
    // rows of the bitmaps and of an output frame then sit at rank * chunk, which lets the
    // exchanges be single in-place all-gathers
    const int q = (gh + nranks - 1) / nranks;
    const int gy0 = rank * q, gy1 = (gy0 + q < gh) ? gy0 + q : gh;
    if ((nranks - 1) * q >= gh)
        return set_err(NLK_ERR_PARAM, "%d grid rows do not split into %d strips", gh, nranks);
    out->gy0 = gy0; out->gy1 = gy1;
    out->chunk_g = q; out->chunk_y = q * step;
    // every rank must own the rows its neighbours spill into: a strip is at least r + psz rows tall
    const int last_rows = h - (nranks - 1) * q * step;
    if (nranks > 1 && (q * step < rmax + psz || last_rows < rmax + psz))
        return set_err(NLK_ERR_PARAM, "%d strips of %d grid rows are thinner than the halo (%d rows)", nranks, q, rmax + psz);
    out->oy0 = gy0 * step;
    out->oy1 = rank == nranks - 1 ? h : gy1 * step;
    int ey0 = gy0 * step - rmax, ey1 = (gy1 - 1) * step + rmax + psz;
    out->ey0 = ey0 < 0 ? 0 : ey0;
    out->ey1 = ey1 > h ? h : ey1;
    return NLK_OK;
}

extern "C" int nlk_strip_search(nlk_ctx *c, int smooth, const float *d_in1, const float *d_prev0,
                                const float *d_bsic1, float sigma, struct nlkalman_params prms,
                                int gy0, int gy1, unsigned int *d_nbr, float *d_accw)
{
    if (int r = enter_strip(c)) return r;
    if (!d_nbr || !d_accw) return set_err(NLK_ERR_PARAM, "the strip pass needs the caller's bitmap and accumulator buffers");
    PassParams &P = c->strip_Ps[c->strip_lane];
    c->strip_opens[c->strip_lane] = false;
    if (int r = pass_setup(c, P, smooth, nullptr, d_in1, d_prev0, d_bsic1, sigma, prms, false, d_nbr, d_accw)) return r;
    if (gy0 < 0 || gy1 > P.gh || gy1 < gy0) return set_err(NLK_ERR_PARAM, "grid rows [%d,%d) outside 0..%d", gy0, gy1, P.gh);
    P.gy0 = gy0; P.gy1 = gy1;
    KindScope ks(c, pass_kind(P));
    if (int r = pass_search(c, P)) return r;
    c->strip_opens[c->strip_lane] = true;
    return NLK_OK;
}

extern "C" int nlk_strip_filter(nlk_ctx *c)
{
    if (int r = enter_strip(c)) return r;
    if (!c->strip_opens[c->strip_lane]) return set_err(NLK_ERR_STATE, "nlk_strip_search must come first");
    PassParams &P = c->strip_Ps[c->strip_lane];
    KindScope ks(c, pass_kind(P));
    c->reserve_sm = c->strip_reserve;
    const int r = pass_filter(c, P, true);
    c->reserve_sm = 0;
    return r;
}

extern "C" int nlk_strip_normalize(nlk_ctx *c, float *d_out, int row0, int row1)
{
    if (int r = enter_strip(c)) return r;
    const int ln = c->strip_lane;
    if (!c->strip_opens[ln]) return set_err(NLK_ERR_STATE, "nlk_strip_search must come first");
    if (int r = rows_ok(c, row0, row1)) return r;
    c->strip_Ps[ln].out = d_out;
    if (c->side_pendings[ln]) {      // the previous frame buffer's whole-strip push reads what this may overwrite
        CU_TRY(cudaStreamWaitEvent(c->L->st, c->ev_side_dones[ln], 0));
        c->side_pendings[ln] = false;
    }
    KindScope ks(c, pass_kind(c->strip_Ps[ln]));
    return pass_normalize(c, c->strip_Ps[ln], row0, row1);
}

// ---- peer-memory exchanges between strips (nlk_peer.cuh) ---------------------------------------

extern "C" size_t nlk_peer_header_bytes(void) { return PEER_HDR_BYTES; }

extern "C" int nlk_peer_slab_alloc(nlk_ctx *c, size_t bytes, void **d_slab)
{
    if (int r = enter(c)) return r;
    if (!d_slab || bytes < PEER_HDR_BYTES) return set_err(NLK_ERR_PARAM, "slab smaller than its header (%zu bytes)", PEER_HDR_BYTES);
    void *p = nullptr;
    CU_TRY(cudaMalloc(&p, bytes));       // plain cudaMalloc: exportable with cudaIpcGetMemHandle
    CU_TRY(cudaMemsetAsync(p, 0, bytes, c->L->st));
    CU_TRY(cudaStreamSynchronize(c->L->st));
    *d_slab = p;
    return NLK_OK;
}

extern "C" int nlk_peer_ipc_export(nlk_ctx *c, void *d_slab, unsigned char *handle64)
{
    if (int r = enter(c)) return r;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    cudaIpcMemHandle_t h;
    CU_TRY(cudaIpcGetMemHandle(&h, d_slab));
    memcpy(handle64, &h, 64);
    return NLK_OK;
}

extern "C" int nlk_peer_ipc_import(nlk_ctx *c, const unsigned char *handle64, void **d_ptr)
{
    if (int r = enter(c)) return r;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void *p = nullptr;
    CU_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->ipc_opened.push_back(p);
    *d_ptr = p;
    return NLK_OK;
}

extern "C" int nlk_peer_bind(nlk_ctx *c, int rank, int nranks, void *const *slabs, size_t slab_bytes)
{
    if (int r = enter(c)) return r;
    if (nranks < 1 || nranks > PEER_MAX || rank < 0 || rank >= nranks || !slabs)
        return set_err(NLK_ERR_PARAM, "bad peer table (1..%d ranks)", PEER_MAX);
    memset(&c->peer, 0, sizeof c->peer);
    for (int i = 0; i < nranks; ++i) {
        if (!slabs[i]) return set_err(NLK_ERR_PARAM, "null slab of rank %d", i);
        c->peer.slab[i] = static_cast<char *>(slabs[i]);
    }
    c->peer.rank = rank; c->peer.nranks = nranks;
    c->slab_bytes = slab_bytes;
    for (int i = 0; i < 2; ++i) {
        if (c->st_sides[i]) continue;
        CU_TRY(cudaStreamCreateWithFlags(&c->st_sides[i], cudaStreamNonBlocking));
        CU_TRY(cudaEventCreateWithFlags(&c->ev_side_forks[i], cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&c->ev_side_dones[i], cudaEventDisableTiming));
    }
    c->peer_on = true;
    return NLK_OK;
}

static int peer_ok(nlk_ctx *c, size_t off, size_t bytes, unsigned int mask)
{
    if (!c->peer_on) return set_err(NLK_ERR_STATE, "nlk_peer_bind must come first");
    if (off + bytes > c->slab_bytes) return set_err(NLK_ERR_PARAM, "range [%zu, %zu) outside the slab", off, off + bytes);
    if (mask >> c->peer.nranks) return set_err(NLK_ERR_PARAM, "peer mask 0x%x names ranks beyond %d", mask, c->peer.nranks);
    return NLK_OK;
}

extern "C" int nlk_peer_signal(nlk_ctx *c, int slot, unsigned int value, unsigned int peer_mask)
{
    if (int r = enter_strip(c)) return r;
    if (int r = peer_ok(c, 0, 0, peer_mask)) return r;
    if (slot < 0 || slot >= PEER_SLOTS) return set_err(NLK_ERR_PARAM, "flag slot %d", slot);
    if (!peer_mask) return NLK_OK;
    ProfScope ps(c, NLK_K_PEER_PUSH);
    k_peer_signal<<<1, 32, 0, c->L->st>>>(c->peer, slot, value, peer_mask);
    return check_launch(c, 1, "peer_signal");
}

extern "C" int nlk_peer_wait(nlk_ctx *c, int slot, unsigned int value, unsigned int src_mask)
{
    if (int r = enter_strip(c)) return r;
    if (int r = peer_ok(c, 0, 0, src_mask)) return r;
    if (slot < 0 || slot >= PEER_SLOTS) return set_err(NLK_ERR_PARAM, "flag slot %d", slot);
    if (!src_mask) return NLK_OK;
    static const unsigned long long tmo = getenv("NLK_PEER_TIMEOUT_MS") ? strtoull(getenv("NLK_PEER_TIMEOUT_MS"), nullptr, 10) * 1000000ull
                                                                         : 4000000000ull;
    ProfScope ps(c, NLK_K_PEER_WAIT);
    k_peer_wait<<<1, 32, 0, c->L->st>>>(c->peer, slot, value, src_mask, tmo);
    return check_launch(c, 1, "peer_wait");
}

// side = 0: a copy kernel on the context's stream (small, latency-critical ranges).
// side = 1: copy engines on a side stream that forks from the context's stream here and runs
// beside what is queued next (whole strips of an output frame); the context's stream joins
// it again before the next nlk_strip_normalize.
extern "C" int nlk_peer_push(nlk_ctx *c, size_t off, size_t bytes, unsigned int peer_mask, int slot,
                             unsigned int value, int side)
{
    if (int r = enter_strip(c)) return r;
    if (int r = peer_ok(c, off, bytes, peer_mask)) return r;
    if (slot >= PEER_SLOTS) return set_err(NLK_ERR_PARAM, "flag slot %d", slot);
    peer_mask &= ~(1u << c->peer.rank);
    if (!peer_mask) return NLK_OK;
    if (side) {
        const int ln = c->strip_lane;
        cudaStream_t st_side = c->st_sides[ln];
        CU_TRY(cudaEventRecord(c->ev_side_forks[ln], c->L->st));
        CU_TRY(cudaStreamWaitEvent(st_side, c->ev_side_forks[ln], 0));
        const char *src = c->peer.slab[c->peer.rank] + off;
        // nearest ranks first: they read the rows soonest
        for (int d = 1; d < c->peer.nranks; ++d)
            for (int sgn = -1; sgn <= 1; sgn += 2) {
                const int p = c->peer.rank + sgn * d;
                if (p < 0 || p >= c->peer.nranks || !((peer_mask >> p) & 1u)) continue;
                if (bytes) CU_TRY(cudaMemcpyAsync(c->peer.slab[p] + off, src, bytes, cudaMemcpyDeviceToDevice, st_side));
            }
        if (slot >= 0) {
            k_peer_signal<<<1, 32, 0, st_side>>>(c->peer, slot, value, peer_mask);
            if (int r = check_launch(c, 1, "peer_signal")) return r;
        }
        CU_TRY(cudaEventRecord(c->ev_side_dones[ln], st_side));
        c->side_pendings[ln] = true;
        return NLK_OK;
    }
    ProfScope ps(c, NLK_K_PEER_PUSH);
    const int cnt = (int)(c->cnt_rot++ & 3) + 4 * c->strip_lane;
    const bool v16 = (off % 16 == 0) && (bytes % 16 == 0);
    const size_t n = v16 ? bytes / 16 : bytes / 4;
    if (bytes % 4) return set_err(NLK_ERR_PARAM, "push of %zu bytes: not a multiple of 4", bytes);
    int nb = (int)((n + 255) / 256);
    if (nb > 2 * c->num_sms) nb = 2 * c->num_sms;
    if (nb < 1) nb = 1;
    if (v16) k_peer_push<uint4><<<nb, 256, 0, c->L->st>>>(c->peer, off, n, peer_mask, slot, value, cnt);
    else k_peer_push<unsigned int><<<nb, 256, 0, c->L->st>>>(c->peer, off, n, peer_mask, slot, value, cnt);
    return check_launch(c, 1, "peer_push");
}

extern "C" int nlk_peer_push_add(nlk_ctx *c, size_t off, size_t bytes, int peer, int slot, unsigned int value)
{
    if (int r = enter_strip(c)) return r;
    if (peer < 0 || peer >= PEER_MAX) return set_err(NLK_ERR_PARAM, "peer %d", peer);
    if (int r = peer_ok(c, off, bytes, 1u << peer)) return r;
    if (slot >= PEER_SLOTS || bytes % 4) return set_err(NLK_ERR_PARAM, "bad push_add request");
    if (peer == c->peer.rank) return set_err(NLK_ERR_PARAM, "push_add to oneself");
    ProfScope ps(c, NLK_K_PEER_PUSH);
    const int cnt = (int)(c->cnt_rot++ & 3) + 4 * c->strip_lane;
    const size_t n = bytes / 4;
    const bool v4 = (off % 16 == 0) && (n % 4 == 0);
    int nb = (int)(((v4 ? n / 4 : n) + 255) / 256);
    if (nb > 2 * c->num_sms) nb = 2 * c->num_sms;
    if (nb < 1) nb = 1;
    if (v4) k_peer_push_add<4><<<nb, 256, 0, c->L->st>>>(c->peer, off, n, peer, slot, value, cnt);
    else k_peer_push_add<1><<<nb, 256, 0, c->L->st>>>(c->peer, off, n, peer, slot, value, cnt);
    return check_launch(c, 1, "peer_push_add");
}

// warp of rows [row0, row1) of the frame at frame_off of the slabs; rows [lo, hi) are read locally, any
// other row from its owner's slab (see k_warp_peer)
extern "C" int nlk_warp_rows_peer_dev(nlk_ctx *c, float *d_imw, size_t frame_off, const float *d_of, const float *d_msk,
                                      int row0, int row1, int lo, int hi, int chunk_y)
{
    if (int r = enter_strip(c)) return r;
    if (int r = rows_ok(c, row0, row1)) return r;
    if (int r = peer_ok(c, frame_off, c->img_bytes(), 0)) return r;
    if (chunk_y < 1 || lo > hi) return set_err(NLK_ERR_PARAM, "bad row ownership (chunk %d, local [%d, %d))", chunk_y, lo, hi);
    if (row1 <= row0) return NLK_OK;
    ProfScope ps(c, NLK_K_WARP);
    const dim3 nt(32, 8), nb((c->w + 31) / 32, (row1 - row0 + 7) / 8);
    if (c->ch == 3) k_warp_peer<3><<<nb, nt, 0, c->L->st>>>(d_imw, c->peer, c->peer.slab[c->peer.rank], frame_off, d_of, d_msk, c->w, c->h, c->ch, row0, row1, lo, hi, chunk_y);
    else if (c->ch == 1) k_warp_peer<1><<<nb, nt, 0, c->L->st>>>(d_imw, c->peer, c->peer.slab[c->peer.rank], frame_off, d_of, d_msk, c->w, c->h, c->ch, row0, row1, lo, hi, chunk_y);
    else k_warp_peer<0><<<nb, nt, 0, c->L->st>>>(d_imw, c->peer, c->peer.slab[c->peer.rank], frame_off, d_of, d_msk, c->w, c->h, c->ch, row0, row1, lo, hi, chunk_y);
    return check_launch(c, 1, "warp_rows_peer");
}

// nonzero once a wait has timed out (the flag never came): (0x10000 | slot << 8 | source rank)
extern "C" int nlk_peer_error(nlk_ctx *c, unsigned int *code)
{
    if (int r = nlk_ctx_sync(c)) return r;
    if (!c->peer_on) return set_err(NLK_ERR_STATE, "nlk_peer_bind must come first");
    CU_TRY(cudaMemcpy(code, c->peer.slab[c->peer.rank] + PEER_ERR_OFF, 4, cudaMemcpyDeviceToHost));
    return NLK_OK;
}

// ---- Dual TV-L1 optical flow at one scale (nlk_tvl1.cuh) ---------------------------

namespace {
// where a level's work planes sit in tv_scratch
struct Tvl1Level {
    int nx, ny, warps;
    size_t size;
    float *I1x, *I1y, *I1wx, *I1wy, *grad, *rho_c, *p11, *p12, *p21, *p22, *err;
    int *cnt, *nloop;
    unsigned *bars;                                           // one grid-barrier counter per warping step
    static constexpr int ES = TVL1_MAX_ITERATIONS + 4;        // error slots per warping step
    static size_t bytes(size_t size, int warps) { return (10 * size + (size_t)warps * ES + 64) * 4 + (size_t)warps * 12 + 64; }
    Tvl1Level(float *base, int nx_, int ny_, int warps_) : nx(nx_), ny(ny_), warps(warps_), size((size_t)nx_ * ny_)
    {
        I1x = base; I1y = I1x + size; I1wx = I1y + size; I1wy = I1wx + size; grad = I1wy + size; rho_c = grad + size;
        p11 = rho_c + size; p12 = p11 + size; p21 = p12 + size; p22 = p21 + size;
        err = p22 + size;
        cnt = reinterpret_cast<int *>(err + (size_t)warps * ES);
        nloop = cnt + warps;
        bars = reinterpret_cast<unsigned *>(nloop + warps);
    }
    size_t zero_bytes() const { return ((size_t)warps * ES) * 4 + (size_t)warps * 12; }    // err, cnt, nloop, bars
    dim3 grid() const { return dim3((nx + 31) / 32, (ny + 7) / 8); }
};
}

// One level as ONE graph launch: the loop of every warping step is a WHILE node whose body is
// { k_tvl1_u, k_tvl1_p } (nlk_tvl1.cuh; k_tvl1_p also sets the loop condition), so the stopping rule needs neither a host round
// trip nor launches past the stopping iteration.  Built by stream capture (the conditional nodes are
// added to the capturing graph by hand), instantiated once per (buffers, size, parameters) and kept.
static cudaError_t tvl1_build_graph(nlk_ctx *c, Tvl1GraphEntry &E)
{
    cudaStream_t st = c->L->st;
    if (!c->tv_st2) { cudaError_t e = cudaStreamCreateWithFlags(&c->tv_st2, cudaStreamNonBlocking); if (e != cudaSuccess) return e; }
    const Tvl1Level L(E.base, E.nx, E.ny, E.warps);
    const float l_t = E.lambda * E.theta, taut = E.tau / E.theta, eps2 = E.epsilon * E.epsilon;
    const dim3 nt(32, 8), nb = L.grid();
    cudaError_t err = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    if (err != cudaSuccess) return err;
    auto CK = [&](cudaError_t x) { if (x != cudaSuccess && err == cudaSuccess) err = x; };
    cudaGraphConditionalHandle loops[64];
    cudaGraph_t bodies[64];
    CK(cudaMemsetAsync(L.p11, 0, 4 * L.size * 4, st));                       // p = 0 (:130-134)
    CK(cudaMemsetAsync(L.err, 0, L.zero_bytes(), st));
    k_tvl1_loop_init<<<1, 64, 0, st>>>(L.nloop, L.warps);
    k_tvl1_centered_gradient<<<nb, nt, 0, st>>>(E.I1, L.I1x, L.I1y, L.nx, L.ny);
    for (int wi = 0; wi < L.warps && err == cudaSuccess; ++wi) {
        k_tvl1_warp<<<nb, nt, 0, st>>>(E.I0, E.I1, L.I1x, L.I1y, E.u1, E.u2, L.I1wx, L.I1wy, L.grad, L.rho_c, L.nx, L.ny);
        // the WHILE node goes after everything captured so far
        cudaStreamCaptureStatus status;
        cudaGraph_t g = nullptr;
        const cudaGraphNode_t *deps = nullptr;
        size_t ndeps = 0;
        CK(cudaStreamGetCaptureInfo(st, &status, nullptr, &g, &deps, &ndeps));
        if (err != cudaSuccess) break;
        cudaGraphConditionalHandle loop;
        CK(cudaGraphConditionalHandleCreate(&loop, g, 1, cudaGraphCondAssignDefault));     // iteration 1 always runs
        cudaGraphNodeParams np = {};       // (reserved fields must be zero)
        np.type = cudaGraphNodeTypeConditional;
        np.conditional.handle = loop;
        np.conditional.type = cudaGraphCondTypeWhile;
        np.conditional.size = 1;
        cudaGraphNode_t node;
        CK(cudaGraphAddNode(&node, g, deps, ndeps, &np));
        if (err != cudaSuccess) break;
        CK(cudaStreamUpdateCaptureDependencies(st, &node, 1, cudaStreamSetCaptureDependencies));
        loops[wi] = loop;
        bodies[wi] = np.conditional.phGraph_out[0];
    }
    cudaGraph_t graph = nullptr;
    CK(cudaStreamEndCapture(st, &graph));          // (also leaves capture mode after a failure)
    CK(cudaGetLastError());
    // the loop bodies, captured into the graphs the WHILE nodes own
    for (int wi = 0; wi < L.warps && err == cudaSuccess; ++wi) {
        float *e = L.err + (size_t)wi * Tvl1Level::ES;
        CK(cudaStreamBeginCaptureToGraph(c->tv_st2, bodies[wi], nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
        if (err != cudaSuccess) break;
        k_tvl1_u<<<nb, nt, 0, c->tv_st2>>>(L.rho_c, L.I1wx, L.I1wy, L.grad, L.p11, L.p12, L.p21, L.p22, E.u1, E.u2, e, 0,
                                           L.nloop + wi, L.nx, L.ny, l_t, E.theta, eps2);
        k_tvl1_p<<<nb, nt, 0, c->tv_st2>>>(E.u1, E.u2, L.p11, L.p12, L.p21, L.p22, e, 0, L.nx, L.ny, taut, eps2,
                                           (unsigned long long)loops[wi], L.nloop + wi, L.cnt + wi);
        cudaGraph_t same = nullptr;                 // (the body graph again)
        CK(cudaStreamEndCapture(c->tv_st2, &same));
        CK(cudaGetLastError());
    }
    if (err == cudaSuccess) err = cudaGraphInstantiate(&E.exec, graph, 0);
    if (err != cudaSuccess) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        return err;
    }
    E.graph = graph;
    return cudaSuccess;
}

// NLK_TVL1_LOOP = kernel (default) | graph | host
static void tvl1_pick_loop(nlk_ctx *c)
{
    if (c->tv_loop) return;
    const char *env = getenv("NLK_TVL1_LOOP");
    c->tv_loop = TVL1_LOOP_KERNEL;
    if (env && !strcmp(env, "graph")) c->tv_loop = TVL1_LOOP_GRAPH;
    if (env && !strcmp(env, "host")) c->tv_loop = TVL1_LOOP_HOST;
    if (c->tv_loop == TVL1_LOOP_KERNEL) {
        int coop = 0, per_sm = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device);
        if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tvl1_iterate, TVL1_IT_THREADS, 0) == cudaSuccess &&
            per_sm > 0)
            c->tv_coop_blocks = per_sm * c->num_sms;
        else {
            cudaGetLastError();
            c->tv_loop = TVL1_LOOP_GRAPH;
        }
    }
}

// the cached graph of this level, built on first use; nullptr: another loop form is selected, or graphs
// are not available on this driver (said once on stderr) and the caller queues the kernels itself
static cudaGraphExec_t tvl1_level_graph(nlk_ctx *c, const float *I0, const float *I1, float *u1, float *u2, float *base,
                                        int nx, int ny, float tau, float lambda, float theta, int warps, float epsilon)
{
    if (c->tv_loop != TVL1_LOOP_GRAPH || warps < 1 || warps > 64) return nullptr;
    for (Tvl1GraphEntry &E : c->tv_graphs)
        if (E.I0 == I0 && E.I1 == I1 && E.u1 == u1 && E.u2 == u2 && E.base == base && E.nx == nx && E.ny == ny &&
            E.warps == warps && E.tau == tau && E.lambda == lambda && E.theta == theta && E.epsilon == epsilon) {
            E.stamp = ++c->tv_stamp;
            return E.exec;
        }
    Tvl1GraphEntry E{I0, I1, u1, u2, base, nx, ny, warps, tau, lambda, theta, epsilon, nullptr, nullptr, ++c->tv_stamp};
    const cudaError_t e = tvl1_build_graph(c, E);
    if (e != cudaSuccess) {
        fprintf(stderr, "[nlkalman_b200] TV-L1: CUDA graph loop not available (%s); queueing the iterations from the host\n",
                cudaGetErrorString(e));
        c->tv_loop = TVL1_LOOP_HOST;
        return nullptr;
    }
    if (c->tv_graphs.size() >= 48) {      // evict the least recently used
        size_t lru = 0;
        for (size_t i = 1; i < c->tv_graphs.size(); ++i) if (c->tv_graphs[i].stamp < c->tv_graphs[lru].stamp) lru = i;
        cudaGraphExecDestroy(c->tv_graphs[lru].exec);
        cudaGraphDestroy(c->tv_graphs[lru].graph);
        c->tv_graphs[lru] = E;
    } else {
        c->tv_graphs.push_back(E);
    }
    return E.exec;
}

extern "C" int nlk_tvl1_level_dev(nlk_ctx *c, const float *d_I0, const float *d_I1, float *d_u1, float *d_u2,
                                  int nx, int ny, float tau, float lambda, float theta, int warps, float epsilon,
                                  int *iterations)
{
    if (int r = enter(c)) return r;
    if (!d_I0 || !d_I1 || !d_u1 || !d_u2 || nx < 2 || ny < 2 || warps < 0 || warps > 64)
        return set_err(NLK_ERR_PARAM, "bad TV-L1 request (%dx%d, %d warpings)", nx, ny, warps);
    const size_t size = (size_t)nx * ny;
    if (int r = c->tv_scratch.ensure(Tvl1Level::bytes(size, warps))) return r;
    const Tvl1Level L(c->tv_scratch.as<float>(), nx, ny, warps);
    cudaStream_t st = c->L->st;
    int rc = NLK_OK;
    if (!c->tv_herr) {
        CU_TRY(cudaMallocHost(&c->tv_herr, 64));
        memset(c->tv_herr, 0, 64);
    }
    int *host_flag = reinterpret_cast<int *>(c->tv_herr) + 8;     // raised by a grid barrier that gave up
    if (*(volatile int *)host_flag) return set_err(NLK_ERR_CUDA, "TV-L1: a grid barrier of k_tvl1_iterate timed out");
    tvl1_pick_loop(c);
    if (c->tv_loop == TVL1_LOOP_KERNEL && warps > 0) {
        float l_t = lambda * theta, taut = tau / theta, eps2 = epsilon * epsilon, th = theta;
        int nxv = nx, nyv = ny;
        const dim3 nt(32, 8), nb = L.grid();
        CU_TRY(cudaMemsetAsync(L.p11, 0, 4 * size * 4, st));                       // p = 0 (:130-134)
        CU_TRY(cudaMemsetAsync(L.err, 0, L.zero_bytes(), st));
        k_tvl1_centered_gradient<<<nb, nt, 0, st>>>(d_I1, L.I1x, L.I1y, nx, ny);
        int blocks = (int)((size + TVL1_IT_THREADS - 1) / TVL1_IT_THREADS);
        if (blocks > c->tv_coop_blocks) blocks = c->tv_coop_blocks;
        for (int wi = 0; wi < warps; ++wi) {
            k_tvl1_warp<<<nb, nt, 0, st>>>(d_I0, d_I1, L.I1x, L.I1y, d_u1, d_u2, L.I1wx, L.I1wy, L.grad, L.rho_c, nx, ny);
            float *e = L.err + (size_t)wi * Tvl1Level::ES;
            int *cnt = L.cnt + wi;
            unsigned *bar = L.bars + wi;
            float *rho_c = L.rho_c, *I1wx = L.I1wx, *I1wy = L.I1wy, *grad = L.grad, *p11 = L.p11, *p12 = L.p12, *p21 = L.p21,
                  *p22 = L.p22;
            void *args[] = {&rho_c, &I1wx, &I1wy, &grad, &p11, &p12, &p21, &p22, &d_u1, &d_u2, &e, &cnt, &bar, &host_flag,
                            &nxv, &nyv, &l_t, &th, &taut, &eps2};
            CU_TRY(cudaLaunchCooperativeKernel((const void *)k_tvl1_iterate, dim3(blocks), dim3(TVL1_IT_THREADS), args, 0, st));
        }
        if (int r = check_launch(c, 1 + 2 * warps, "tvl1 level")) return r;
    } else if (cudaGraphExec_t exec = tvl1_level_graph(c, d_I0, d_I1, d_u1, d_u2, c->tv_scratch.as<float>(), nx, ny, tau, lambda,
                                                theta, warps, epsilon)) {
        CU_TRY(cudaGraphLaunch(exec, st));
        c->launches += 1;
    } else {
        const float l_t = lambda * theta, taut = tau / theta, eps2 = epsilon * epsilon;
        const dim3 nt(32, 8), nb = L.grid();
        CU_TRY(cudaMemsetAsync(L.p11, 0, 4 * size * 4, st));                       // p = 0 (:130-134)
        CU_TRY(cudaMemsetAsync(L.err, 0, L.zero_bytes(), st));
        k_tvl1_centered_gradient<<<nb, nt, 0, st>>>(d_I1, L.I1x, L.I1y, nx, ny);
        if (int r = check_launch(c, 1, "tvl1 gradient")) return r;
        float *h_err = c->tv_herr;
        for (int wi = 0; wi < warps && rc == NLK_OK; ++wi) {
            float *e = L.err + (size_t)wi * Tvl1Level::ES;
            k_tvl1_warp<<<nb, nt, 0, st>>>(d_I0, d_I1, L.I1x, L.I1y, d_u1, d_u2, L.I1wx, L.I1wy, L.grad, L.rho_c, nx, ny);
            rc = check_launch(c, 1, "tvl1 warp");
            // batches of iterations; the kernels themselves stop at the reference's stopping iteration
            // (:164), the host only learns between batches that nothing is left to queue.  Batches grow
            // 4, 8, 16, 32, 32 ...: most warping steps after the first of a scale stop within a few iterations,
            // and every queued iteration past the stop is two (empty) launches.
            int batch = 4;
            for (int n0 = 1; n0 <= TVL1_MAX_ITERATIONS && rc == NLK_OK; n0 += batch, batch = batch < 32 ? 2 * batch : 32) {
                const int n1 = n0 + batch - 1 < TVL1_MAX_ITERATIONS ? n0 + batch - 1 : TVL1_MAX_ITERATIONS;
                for (int n = n0; n <= n1; ++n) {
                    k_tvl1_u<<<nb, nt, 0, st>>>(L.rho_c, L.I1wx, L.I1wy, L.grad, L.p11, L.p12, L.p21, L.p22, d_u1, d_u2, e, n,
                                                nullptr, nx, ny, l_t, theta, eps2);
                    k_tvl1_p<<<nb, nt, 0, st>>>(d_u1, d_u2, L.p11, L.p12, L.p21, L.p22, e, n, nx, ny, taut, eps2, 0ull, nullptr, nullptr);
                }
                rc = check_launch(c, 2 * (n1 - n0 + 1), "tvl1 iteration");
                if (rc != NLK_OK || n1 == TVL1_MAX_ITERATIONS) break;
                if (cudaMemcpyAsync(h_err, e + n1, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                    cudaStreamSynchronize(st) != cudaSuccess) {
                    rc = set_err(NLK_ERR_CUDA, "tvl1: %s", cudaGetErrorString(cudaGetLastError()));
                    break;
                }
                if (!(*h_err / (float)size > eps2)) break;     // converged inside this batch
            }
            k_tvl1_count<<<1, 1, 0, st>>>(e, L.cnt + wi, (float)size, eps2);
            if (rc == NLK_OK) rc = check_launch(c, 1, "tvl1 count");
        }
    }
    if (rc != NLK_OK) return rc;
    if (iterations && warps > 0) {
        CU_TRY(cudaMemcpyAsync(iterations, L.cnt, (size_t)warps * 4, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        if (*(volatile int *)host_flag) return set_err(NLK_ERR_CUDA, "TV-L1: a grid barrier of k_tvl1_iterate timed out");
    }
    return NLK_OK;
}

extern "C" int nlk_tvl1_level_host(nlk_ctx *c, const float *h_I0, const float *h_I1, float *h_u1, float *h_u2,
                                   int nx, int ny, float tau, float lambda, float theta, int warps, float epsilon,
                                   int *iterations)
{
    if (int r = enter(c)) return r;
    if (!h_I0 || !h_I1 || !h_u1 || !h_u2 || nx < 2 || ny < 2) return set_err(NLK_ERR_PARAM, "bad TV-L1 request");
    const size_t pb = (size_t)nx * ny * 4;
    if (int r = c->q_tmp.ensure(4 * pb)) return r;
    float *d = c->q_tmp.as<float>();
    const size_t n = (size_t)nx * ny;
    CU_TRY(cudaMemcpyAsync(d, h_I0, pb, cudaMemcpyHostToDevice, c->L->st));
    CU_TRY(cudaMemcpyAsync(d + n, h_I1, pb, cudaMemcpyHostToDevice, c->L->st));
    CU_TRY(cudaMemcpyAsync(d + 2 * n, h_u1, pb, cudaMemcpyHostToDevice, c->L->st));
    CU_TRY(cudaMemcpyAsync(d + 3 * n, h_u2, pb, cudaMemcpyHostToDevice, c->L->st));
    if (int r = nlk_tvl1_level_dev(c, d, d + n, d + 2 * n, d + 3 * n, nx, ny, tau, lambda, theta, warps, epsilon, iterations)) return r;
    CU_TRY(cudaMemcpyAsync(h_u1, d + 2 * n, pb, cudaMemcpyDeviceToHost, c->L->st));
    CU_TRY(cudaMemcpyAsync(h_u2, d + 3 * n, pb, cudaMemcpyDeviceToHost, c->L->st));
    CU_TRY(cudaStreamSynchronize(c->L->st));
    return NLK_OK;
}

// ---- the whole estimator: pyramid + levels (Dual_TVL1_optic_flow_multiscale, tvl1flow_lib.c:345-477) ----

extern "C" int nlk_tvl1_scales(int nx, int ny, float zfactor, int nscales)
{
    return tvl1_scales_cap(nx, ny, zfactor, nscales);
}

namespace {
// the steps of Tvl1Pyramid::run as kernel launches on the context's stream
struct Tvl1Cuda {
    nlk_ctx *c;
    cudaStream_t st;
    float tau, lambda, theta, epsilon;
    int warps, launches = 0;
    static dim3 grid(int w, int h) { return dim3((w + 31) / 32, (h + 7) / 8); }
    void upload(double *dB, const std::vector<double> &B)
    {   // (pageable source: staged before the call returns)
        cudaMemcpyAsync(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice, st);
    }
    void zero(float *p, size_t n) { cudaMemsetAsync(p, 0, n * 4, st); }
    void normalize(const float *I0, const float *I1, float *O0, float *O1, size_t n, float *part)
    {
        const int nparts = (int)std::min<size_t>(TVL1_MINMAX_BLOCKS, (n + 255) / 256);
        k_tvl1_minmax<<<nparts, 256, 0, st>>>(I0, I1, n, part);
        k_tvl1_normalize<<<nparts, 256, 0, st>>>(I0, I1, O0, O1, n, part, nparts);
        launches += 2;
    }
    void gauss(const float *in, float *tmp, float *out, int w, int h, const double *B, int taps)
    {
        k_tvl1_gauss<true><<<grid(w, h), dim3(32, 8), 0, st>>>(in, tmp, w, h, B, taps);
        k_tvl1_gauss<false><<<grid(w, h), dim3(32, 8), 0, st>>>(tmp, out, w, h, B, taps);
        launches += 2;
    }
    void zoom(const float *in, float *out, int w, int h, int ww, int hh, float fx, float fy, float scale, int scaled)
    {
        k_tvl1_zoom<<<grid(ww, hh), dim3(32, 8), 0, st>>>(in, out, w, h, ww, hh, fx, fy, scale, scaled);
        launches += 1;
    }
    int level(const float *I0, const float *I1, float *u1, float *u2, int w, int h, int *iterations)
    {
        if (int r = check_launch(c, launches, "tvl1 pyramid")) return r;
        launches = 0;
        return nlk_tvl1_level_dev(c, I0, I1, u1, u2, w, h, tau, lambda, theta, warps, epsilon, iterations);
    }
};
}

extern "C" int nlk_tvl1_flow_dev(nlk_ctx *c, const float *d_I0, const float *d_I1, float *d_u1, float *d_u2,
                                 int nxx, int nyy, float tau, float lambda, float theta, int nscales, int fscale,
                                 float zfactor, int warps, float epsilon, int *iterations)
{
    if (int r = enter(c)) return r;
    if (!d_I0 || !d_I1 || !d_u1 || !d_u2) return set_err(NLK_ERR_PARAM, "bad TV-L1 request (null image)");
    Tvl1Pyramid P;
    if (!P.plan(nxx, nyy, nscales, fscale, zfactor, warps))
        return set_err(NLK_ERR_PARAM, "%s (%dx%d, %d scales from %d, zoom %g, %d warpings)", P.error, nxx, nyy, nscales,
                       fscale, (double)zfactor, warps);
    if (int r = c->tv_pyr.ensure(P.floats * 4)) return r;
    // the level solver's planes, sized once for the finest scale it will see (it runs coarse to fine)
    if (fscale < nscales)
        if (int r = c->tv_scratch.ensure(Tvl1Level::bytes(P.size(fscale), warps))) return r;
    Tvl1Cuda ex{c, c->L->st, tau, lambda, theta, epsilon, warps};
    if (int r = P.run(ex, c->tv_pyr.as<float>(), d_I0, d_I1, d_u1, d_u2, iterations)) return r;
    return check_launch(c, ex.launches, "tvl1 pyramid");
}

extern "C" int nlk_tvl1_flow_host(nlk_ctx *c, const float *h_I0, const float *h_I1, float *h_flow, int nx, int ny,
                                  float tau, float lambda, float theta, int nscales, int fscale, float zfactor,
                                  int warps, float epsilon, int *iterations)
{
    if (int r = enter(c)) return r;
    if (!h_I0 || !h_I1 || !h_flow || nx < 2 || ny < 2) return set_err(NLK_ERR_PARAM, "bad TV-L1 request");
    const size_t n = (size_t)nx * ny, pb = n * 4;
    if (int r = c->q_tmp.ensure(4 * pb)) return r;
    float *d = c->q_tmp.as<float>();
    CU_TRY(cudaMemcpyAsync(d, h_I0, pb, cudaMemcpyHostToDevice, c->L->st));
    CU_TRY(cudaMemcpyAsync(d + n, h_I1, pb, cudaMemcpyHostToDevice, c->L->st));
    if (int r = nlk_tvl1_flow_dev(c, d, d + n, d + 2 * n, d + 3 * n, nx, ny, tau, lambda, theta, nscales, fscale, zfactor,
                                  warps, epsilon, iterations)) return r;
    // two planes, u then v: what the reference's driver hands to iio_write_image_float_split (main.c:177)
    CU_TRY(cudaMemcpyAsync(h_flow, d + 2 * n, 2 * pb, cudaMemcpyDeviceToHost, c->L->st));
    CU_TRY(cudaStreamSynchronize(c->L->st));
    return NLK_OK;
}

// ---- flow + mask of two resident frames (scripts/nlkalman-seq.sh:60-72) -----------------------------

extern "C" void nlk_tvl1_default_params(struct nlk_tvl1_params *p)
{
    if (!p) return;
    // lib/tvl1flow/main.c:26-35
    p->tau = 0.25f; p->lambda = 0.15f; p->theta = 0.3f; p->nscales = 100; p->fscale = 0; p->zfactor = 0.5f;
    p->warps = 5; p->epsilon = 0.01f;
}

extern "C" int nlk_flow_mask_dev(nlk_ctx *c, float *d_of, float *d_occ, const float *d_from_rgb, const float *d_to_rgb,
                                 struct nlk_tvl1_params p, float th)
{
    if (int r = enter(c)) return r;
    if (!d_of || !d_from_rgb || !d_to_rgb) return set_err(NLK_ERR_PARAM, "flow: null frame or flow buffer");
    if (c->ch == 2) return set_err(NLK_ERR_PARAM, "flow: a 2-channel frame has no luminance");
    // out-of-range values fall back to the defaults like in the reference's program (main.c:108-148)
    struct nlk_tvl1_params d;
    nlk_tvl1_default_params(&d);
    if (!(p.tau > 0.f && p.tau <= 0.25f)) p.tau = d.tau;
    if (!(p.lambda > 0.f)) p.lambda = d.lambda;
    if (!(p.theta > 0.f)) p.theta = d.theta;
    if (p.nscales <= 0) p.nscales = d.nscales;
    if (!(p.zfactor > 0.f && p.zfactor < 1.f)) p.zfactor = d.zfactor;
    if (p.warps <= 0) p.warps = d.warps;
    if (!(p.epsilon > 0.f)) p.epsilon = d.epsilon;
    if (p.fscale < 0) p.fscale = 0;
    p.nscales = tvl1_scales_cap(c->w, c->h, p.zfactor, p.nscales);
    if (p.nscales < p.fscale) p.fscale = p.nscales;
    const size_t npix = (size_t)c->w * c->h;
    if (int r = c->tv_frames.ensure(4 * npix * 4)) return r;
    float *l0 = c->tv_frames.as<float>(), *l1 = l0 + npix, *u1 = l1 + npix, *u2 = u1 + npix;
    cudaStream_t st = c->L->st;
    const int nb = (int)std::min<size_t>((npix + 255) / 256, 148 * 8);
    k_tvl1_luma<<<nb, 256, 0, st>>>(d_from_rgb, l0, npix, c->ch);
    k_tvl1_luma<<<nb, 256, 0, st>>>(d_to_rgb, l1, npix, c->ch);
    if (int r = check_launch(c, 2, "luminance")) return r;
    if (int r = nlk_tvl1_flow_dev(c, l0, l1, u1, u2, c->w, c->h, p.tau, p.lambda, p.theta, p.nscales, p.fscale, p.zfactor,
                                  p.warps, p.epsilon, nullptr)) return r;
    k_tvl1_interleave<<<nb, 256, 0, st>>>(u1, u2, d_of, npix);
    if (int r = check_launch(c, 1, "flow interleave")) return r;
    if (d_occ) return check_launch(c, launch_occlusion(d_occ, d_of, c->w, c->h, th, st), "occlusion");
    return NLK_OK;
}

// ---- resident sequence recursion --------------------------------------------------------------

extern "C" int nlk_seq_reset(nlk_ctx *c)
{
    if (!c) return set_err(NLK_ERR_PARAM, "null context");
    if (int r = lanes_join(c)) return r;
    c->q_cur = 0; c->q_have_prev = 0; c->q_have_flt2 = 0; c->q_smo_cur = 0; c->q_have_smo = 0;
    c->q_frames = 0;
    return NLK_OK;
}

// One frame of the forward recursion.  hook(which) is called right after output `which`
// (1: first filtering, 2: second) has been queued, with c->L the lane it was queued on.
//
// overlap = false: everything on lane 0, in order.
// overlap = true (pipelined recursion): the first filtering runs on lane 0, the second on
// lane 1, which lags: lane 0 goes on with the first filtering of the NEXT frame while lane 1
// still filters this one (flt1(t+1) needs flt1(t) only; flt2(t) needs flt1(t) and flt2(t-1)).
// The one-SM mask_resolve of one lane then runs beside the other lane's kernels.  Buffers that
// both lanes touch are double-buffered by frame parity (noisy, flt1, flt2); lane 0 waits for
// lane 1's frame t-2 before it overwrites them.  The flow and mask of a frame must stay valid
// until lane 1 is done with the frame (nlk_seq_drain, or two later frames queued).
template <class Hook>
static int seq_filter_core(nlk_ctx *c, const float *d_noisy, const float *d_bflo, const float *d_bocc,
                           float sigma, const nlkalman_params &f1, const nlkalman_params &f2,
                           float *d_flt1_out, float *d_flt2_out, bool overlap, Hook hook)
{
    const size_t ib = c->img_bytes();
    for (int i = 0; i < 2; ++i) {
        if (int r = c->q_noisy[i].ensure(ib)) return r;
        if (int r = c->q_flt1[i].ensure(ib)) return r;
        if (int r = c->q_flt2[i].ensure(ib)) return r;
        if (int r = c->lane[i].q_warp.ensure(ib)) return r;
    }
    if (f1.patch_sz == 0) return set_err(NLK_ERR_PARAM, "the resident recursion needs the first filtering (f1_p != 0)");
    const int cur = c->q_cur, prv = cur ^ 1;
    Lane *A = &c->lane[0], *B = overlap ? &c->lane[1] : &c->lane[0];
    if (!overlap) { if (int r = lanes_join(c)) return r; }
    c->L = A;
    c->reserve_sm = overlap ? 1 : 0;
    struct Restore { nlk_ctx *c; ~Restore() { c->L = &c->lane[0]; c->reserve_sm = 0; } } restore{c};
    float *noisy = c->q_noisy[cur].as<float>();
    float *flt1 = c->q_flt1[cur].as<float>(), *flt2 = c->q_flt2[cur].as<float>();
    // the buffers of this parity were last read by the second filtering two frames ago
    if (overlap && c->q_frames >= 2) CU_TRY(cudaStreamWaitEvent(A->st, c->ev_b[cur], 0));
    if (int r = colour_dev(c, noisy, d_noisy, 0)) return r;

    // first filtering (reference src/main-flt.c:345-357)
    const float *prev1 = nullptr;
    if (c->q_have_prev) {
        prev1 = c->q_flt1[prv].as<float>();
        if (d_bflo) {
            float *warp = A->q_warp.as<float>();
            if (int r = warp_dev(c, warp, prev1, d_bflo, d_bocc)) return r;
            prev1 = warp;
        }
    }
    if (int r = run_pass(c, 0, flt1, noisy, prev1, nullptr, sigma, f1, false)) return r;
    if (d_flt1_out) {
        if (int r = colour_dev(c, d_flt1_out, flt1, 1)) return r;
        if (int r = hook(1)) return r;
    }

    // second filtering (reference src/main-flt.c:361-374)
    const int do2 = f2.patch_sz != 0;
    if (do2) {
        if (overlap) {
            CU_TRY(cudaEventRecord(c->ev_a[cur], A->st));
            CU_TRY(cudaStreamWaitEvent(B->st, c->ev_a[cur], 0));
            c->L = B;
            c->b_pending = true;
        }
        const float *prev2 = nullptr;
        if (c->q_have_prev && c->q_have_flt2) {
            prev2 = c->q_flt2[prv].as<float>();
            if (d_bflo) {
                float *warp = B->q_warp.as<float>();
                if (int r = warp_dev(c, warp, prev2, d_bflo, d_bocc)) return r;
                prev2 = warp;
            }
        }
        if (int r = run_pass(c, 0, flt2, noisy, prev2, flt1, sigma, f2, false)) return r;
        if (d_flt2_out) {
            if (int r = colour_dev(c, d_flt2_out, flt2, 1)) return r;
            if (int r = hook(2)) return r;
        }
        if (overlap) CU_TRY(cudaEventRecord(c->ev_b[cur], B->st));
    } else if (overlap) {
        CU_TRY(cudaEventRecord(c->ev_b[cur], A->st));
    }
    c->q_have_prev = 1;
    c->q_have_flt2 = do2;
    c->q_cur = prv;
    c->q_frames += 1;
    return NLK_OK;
}

extern "C" int nlk_seq_filter_dev(nlk_ctx *c, const float *d_noisy, const float *d_bflo,
                                  const float *d_bocc, float sigma, struct nlkalman_params f1,
                                  struct nlkalman_params f2, float *d_flt1_out, float *d_flt2_out)
{
    if (int r = ctx_use(c)) return r;
    return seq_filter_core(c, d_noisy, d_bflo, d_bocc, sigma, f1, f2, d_flt1_out, d_flt2_out, false,
                           [](int) { return NLK_OK; });
}

extern "C" int nlk_seq_submit_dev(nlk_ctx *c, const float *d_noisy, const float *d_bflo,
                                  const float *d_bocc, float sigma, struct nlkalman_params f1,
                                  struct nlkalman_params f2, float *d_flt1_out, float *d_flt2_out)
{
    if (int r = ctx_use(c)) return r;
    return seq_filter_core(c, d_noisy, d_bflo, d_bocc, sigma, f1, f2, d_flt1_out, d_flt2_out, true,
                           [](int) { return NLK_OK; });
}

static int stage_in(nlk_ctx *c, DevBuf &b, const float *h, size_t bytes, const float **d)
{
    *d = nullptr;
    if (!h) return NLK_OK;
    if (int r = b.ensure(bytes)) return r;
    CU_TRY(cudaMemcpyAsync(b.p, h, bytes, cudaMemcpyHostToDevice, c->L->st));
    *d = b.as<float>();
    return NLK_OK;
}

static int pipe_init(nlk_ctx *c)
{
    if (c->st_h2d) return NLK_OK;
    CU_TRY(cudaStreamCreateWithFlags(&c->st_h2d, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&c->st_d2h, cudaStreamNonBlocking));
    for (int i = 0; i < nlk_ctx::PIPE_SETS; ++i) {
        CU_TRY(cudaEventCreateWithFlags(&c->ev_up[i], cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&c->ev_o1[i], cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&c->ev_o2[i], cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
    }
    return NLK_OK;
}

extern "C" int nlk_seq_submit_host(nlk_ctx *c, const float *h_noisy, const float *h_bflo,
                                   const float *h_bocc, float sigma, struct nlkalman_params f1,
                                   struct nlkalman_params f2, float *h_flt1_out, float *h_flt2_out)
{
    if (int r = ctx_use(c)) return r;
    if (!h_noisy) return set_err(NLK_ERR_PARAM, "no noisy frame");
    if (int r = pipe_init(c)) return r;
    const size_t ib = c->img_bytes(), npix = (size_t)c->w * c->h;
    const int s = (int)(c->p_frames % nlk_ctx::PIPE_SETS);
    // at most three frames in flight (one uploading or downloading beside the two the lanes
    // work on): staging set s was last used by frame n-3
    if (c->p_frames >= nlk_ctx::PIPE_SETS) CU_TRY(cudaEventSynchronize(c->ev_done[s]));
    if (int r = c->p_in[s].ensure(ib)) return r;
    CU_TRY(cudaMemcpyAsync(c->p_in[s].p, h_noisy, ib, cudaMemcpyHostToDevice, c->st_h2d));
    const float *d_of = nullptr, *d_msk = nullptr;
    if (h_bflo) {
        if (int r = c->p_of[s].ensure(npix * 2 * 4)) return r;
        CU_TRY(cudaMemcpyAsync(c->p_of[s].p, h_bflo, npix * 2 * 4, cudaMemcpyHostToDevice, c->st_h2d));
        d_of = c->p_of[s].as<float>();
    }
    // the mask: float samples (the reference's in-memory form, src/main-flt.c:236-262), the bytes of the
    // 8-bit file it is read from (scripts/nlkalman-seq.sh:70-73 writes a PNG), or none at all -- built
    // here from the divergence of the flow that was just uploaded (the script's plambda expression)
    int mask_job = 0;
    if (c->mask_mode == NLK_MASK_FROM_FLOW) {
        if (h_bflo) { if (int r = c->p_msk[s].ensure(npix * 4)) return r; d_msk = c->p_msk[s].as<float>(); mask_job = 2; }
    } else if (h_bocc) {
        if (int r = c->p_msk[s].ensure(npix * 4)) return r;
        d_msk = c->p_msk[s].as<float>();
        if (c->mask_mode == NLK_MASK_U8) {
            if (int r = c->p_msk8[s].ensure(npix)) return r;
            CU_TRY(cudaMemcpyAsync(c->p_msk8[s].p, h_bocc, npix, cudaMemcpyHostToDevice, c->st_h2d));
            mask_job = 1;
        } else {
            CU_TRY(cudaMemcpyAsync(c->p_msk[s].p, h_bocc, npix * 4, cudaMemcpyHostToDevice, c->st_h2d));
        }
    }
    CU_TRY(cudaEventRecord(c->ev_up[s], c->st_h2d));
    CU_TRY(cudaStreamWaitEvent(c->lane[0].st, c->ev_up[s], 0));
    if (mask_job) {
        c->L = &c->lane[0];
        ProfScope ps(c, NLK_K_WARP);
        const int n = mask_job == 1 ? launch_mask_u8(c->p_msk[s].as<float>(), c->p_msk8[s].as<uint8_t>(), (long)npix, c->lane[0].st)
                                    : launch_occlusion(c->p_msk[s].as<float>(), d_of, c->w, c->h, c->mask_th, c->lane[0].st);
        if (int r = check_launch(c, n, "mask")) return r;
    }
    float *d_o1 = nullptr, *d_o2 = nullptr;
    if (h_flt1_out) { if (int r = c->p_o1[s].ensure(ib)) return r; d_o1 = c->p_o1[s].as<float>(); }
    if (h_flt2_out && f2.patch_sz != 0) { if (int r = c->p_o2[s].ensure(ib)) return r; d_o2 = c->p_o2[s].as<float>(); }
    // each output goes back on the download stream as soon as it exists: the first
    // filtering's copy overlaps the second filtering
    auto hook = [&](int which) -> int {
        cudaEvent_t ev = which == 1 ? c->ev_o1[s] : c->ev_o2[s];
        CU_TRY(cudaEventRecord(ev, c->L->st));      // the lane the output was produced on
        CU_TRY(cudaStreamWaitEvent(c->st_d2h, ev, 0));
        CU_TRY(cudaMemcpyAsync(which == 1 ? h_flt1_out : h_flt2_out, which == 1 ? d_o1 : d_o2, ib,
                               cudaMemcpyDeviceToHost, c->st_d2h));
        return NLK_OK;
    };
    if (int r = seq_filter_core(c, c->p_in[s].as<float>(), d_of, d_msk, sigma, f1, f2, d_o1, d_o2, true, hook)) return r;
    // frame complete = its compute on both lanes (the staging inputs are free again) and its
    // downloads: lane 1 finishes a frame last (it waited for lane 0's part of it)
    Lane *last = f2.patch_sz != 0 ? &c->lane[1] : &c->lane[0];
    CU_TRY(cudaEventRecord(c->ev_o2[s], last->st));
    CU_TRY(cudaStreamWaitEvent(c->st_d2h, c->ev_o2[s], 0));
    CU_TRY(cudaEventRecord(c->ev_done[s], c->st_d2h));
    c->p_frames += 1;
    return NLK_OK;
}

extern "C" int nlk_seq_drain(nlk_ctx *c) { return nlk_ctx_sync(c); }

// queue a wait for the pipelined recursion on the context's stream (device-side join)
extern "C" int nlk_seq_join(nlk_ctx *c)
{
    if (int r = ctx_use(c)) return r;
    return lanes_join(c);
}

extern "C" int nlk_seq_filter_host(nlk_ctx *c, const float *h_noisy, const float *h_bflo,
                                   const float *h_bocc, float sigma, struct nlkalman_params f1,
                                   struct nlkalman_params f2, float *h_flt1_out, float *h_flt2_out)
{
    if (int r = nlk_seq_submit_host(c, h_noisy, h_bflo, h_bocc, sigma, f1, f2, h_flt1_out, h_flt2_out)) return r;
    return nlk_seq_drain(c);
}

extern "C" int nlk_seq_smooth_start_dev(nlk_ctx *c, const float *d_last_rgb)
{
    if (int r = enter(c)) return r;
    const size_t ib = c->img_bytes();
    for (int i = 0; i < 2; ++i) if (int r = c->q_smo[i].ensure(ib)) return r;
    c->q_smo_cur = 0;
    if (int r = nlk_rgb2opp_dev(c, c->q_smo[0].as<float>(), d_last_rgb)) return r;
    c->q_have_smo = 1;
    return NLK_OK;
}

extern "C" int nlk_seq_smooth_dev(nlk_ctx *c, const float *d_flt_rgb, const float *d_fflo,
                                  const float *d_focc, float sigma, struct nlkalman_params s1,
                                  float *d_smo_out)
{
    if (int r = enter(c)) return r;
    if (!c->q_have_smo) return set_err(NLK_ERR_STATE, "nlk_seq_smooth_start_* must come first");
    const size_t ib = c->img_bytes();
    if (int r = c->q_tmp.ensure(ib)) return r;
    if (int r = c->L->q_warp.ensure(ib)) return r;
    const int nxt = c->q_smo_cur, cur = nxt ^ 1; // q_smo[nxt] holds the smoothed frame t+1
    float *flt = c->q_tmp.as<float>(), *warp = c->L->q_warp.as<float>();
    if (int r = nlk_rgb2opp_dev(c, flt, d_flt_rgb)) return r;
    const float *smo0 = c->q_smo[nxt].as<float>();
    if (d_fflo) { // reference src/main-smo.c:202-206
        if (int r = nlk_warp_dev(c, warp, smo0, d_fflo, d_focc)) return r;
        smo0 = warp;
    }
    float *smo1 = c->q_smo[cur].as<float>();
    if (int r = run_pass(c, 1, smo1, flt, smo0, nullptr, sigma, s1, false)) return r;
    if (d_smo_out) if (int r = nlk_opp2rgb_dev(c, d_smo_out, smo1)) return r;
    c->q_smo_cur = cur;
    return NLK_OK;
}

extern "C" int nlk_seq_smooth_start_host(nlk_ctx *c, const float *h_last_rgb)
{
    if (int r = enter(c)) return r;
    const float *d;
    if (int r = stage_in(c, c->s_in1, h_last_rgb, c->img_bytes(), &d)) return r;
    if (!d) return set_err(NLK_ERR_PARAM, "no frame");
    return nlk_seq_smooth_start_dev(c, d);
}

extern "C" int nlk_seq_smooth_host(nlk_ctx *c, const float *h_flt_rgb, const float *h_fflo,
                                   const float *h_focc, float sigma, struct nlkalman_params s1,
                                   float *h_smo_out)
{
    if (int r = enter(c)) return r;
    const size_t ib = c->img_bytes(), npix = (size_t)c->w * c->h;
    const float *d_flt, *d_of, *d_msk;
    if (int r = stage_in(c, c->s_in1, h_flt_rgb, ib, &d_flt)) return r;
    if (int r = stage_in(c, c->s_of, h_fflo, npix * 2 * 4, &d_of)) return r;
    if (int r = stage_in(c, c->s_msk, h_focc, npix * 4, &d_msk)) return r;
    if (!d_flt) return set_err(NLK_ERR_PARAM, "no filtered frame");
    float *d_o = nullptr;
    if (h_smo_out) { if (int r = c->s_out.ensure(ib)) return r; d_o = c->s_out.as<float>(); }
    if (int r = nlk_seq_smooth_dev(c, d_flt, d_of, d_msk, sigma, s1, d_o)) return r;
    if (d_o) CU_TRY(cudaMemcpyAsync(h_smo_out, d_o, ib, cudaMemcpyDeviceToHost, c->L->st));
    CU_TRY(cudaStreamSynchronize(c->L->st));
    return NLK_OK;
}

// ---- host-image single pass (legacy entry points and the test dump) -----------------------

static int pass_host(nlk_ctx *c, int smooth, float *h_out, const float *h_in1, const float *h_prev0,
                     const float *h_bsic1, float sigma, const nlkalman_params &pr, bool debug)
{
    const size_t ib = c->img_bytes();
    const float *d_in1, *d_prev0, *d_bsic;
    if (int r = stage_in(c, c->s_in1, h_in1, ib, &d_in1)) return r;
    if (int r = stage_in(c, c->s_prev0, h_prev0, ib, &d_prev0)) return r;
    if (int r = stage_in(c, c->s_bsic, h_bsic1, ib, &d_bsic)) return r;
    if (!d_in1) return set_err(NLK_ERR_PARAM, "no input frame");
    if (int r = c->s_out.ensure(ib)) return r;
    if (int r = run_pass(c, smooth, c->s_out.as<float>(), d_in1, d_prev0, d_bsic, sigma, pr, debug)) return r;
    CU_TRY(cudaMemcpyAsync(h_out, c->s_out.p, ib, cudaMemcpyDeviceToHost, c->L->st));
    CU_TRY(cudaStreamSynchronize(c->L->st));
    return NLK_OK;
}

extern "C" int nlk_pass_host_debug(nlk_ctx *c, int smooth, float *h_out, const float *h_in1,
                                   const float *h_prev0, const float *h_bsic1, float sigma,
                                   struct nlkalman_params prms, int kmax, int *nk, int *np0,
                                   int *knn_xy, float *knn_d, unsigned char *prev_p,
                                   unsigned char *active, float *vp)
{
    if (int r = enter(c)) return r;
    if (int r = pass_host(c, smooth, h_out, h_in1, h_prev0, h_bsic1, sigma, prms, true)) return r;
    const int psz = prms.patch_sz, step = psz / 2;
    if (c->w < psz || c->h < psz) return NLK_OK;
    const int gw = (c->w - psz) / step + 1, gh = (c->h - psz) / step + 1, G = gw * gh;
    const int rmax = smooth ? prms.search_sz_t : (prms.search_sz_t > prms.search_sz_x ? prms.search_sz_t : prms.search_sz_x);
    int ks = prms.npatches_x > prms.npatches_t ? prms.npatches_x : prms.npatches_t;
    if (ks > (2 * rmax + 1) * (2 * rmax + 1)) ks = (2 * rmax + 1) * (2 * rmax + 1);
    if (ks < 1) ks = 1;
    std::vector<GroupHdr> hdr(G);
    std::vector<uint32_t> cand((size_t)G * ks);
    std::vector<float> dist((size_t)G * ks), vps(G);
    std::vector<int> act(G);
    int counters[5];
    CU_TRY(cudaMemcpy(hdr.data(), c->L->hdr.p, (size_t)G * sizeof(GroupHdr), cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(cand.data(), c->L->cand.p, (size_t)G * ks * 4, cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(dist.data(), c->dbg_dist.p, (size_t)G * ks * 4, cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(vps.data(), c->dbg_vp.p, (size_t)G * 4, cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(act.data(), c->L->active.p, (size_t)G * 4, cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(counters, c->L->counters.p, sizeof counters, cudaMemcpyDeviceToHost));
    if (counters[4]) return set_err(NLK_ERR_CUDA, "mask_resolve: a ring slot never arrived (internal error)");
    if (active) memset(active, 0, G);
    for (int i = 0; i < counters[0] && i < G; ++i)
        if (active && act[i] >= 0 && act[i] < G) active[act[i]] = 1;
    for (int g = 0; g < G; ++g) {
        if (nk) nk[g] = hdr[g].nk;
        if (np0) np0[g] = hdr[g].np0;
        if (prev_p) prev_p[g] = (hdr[g].flags & HDR_PREV_P) ? 1 : 0;
        if (vp) vp[g] = vps[g];
        for (int i = 0; i < hdr[g].nk && i < kmax; ++i) {
            const uint32_t cd = cand[(size_t)g * ks + i];
            if (knn_xy) {
                knn_xy[((size_t)g * kmax + i) * 2 + 0] = cand_x(cd);
                knn_xy[((size_t)g * kmax + i) * 2 + 1] = cand_y(cd);
            }
            if (knn_d) knn_d[(size_t)g * kmax + i] = dist[(size_t)g * ks + i];
        }
    }
    return NLK_OK;
}

// ---- batched DCT unit (tests) -----------------------------------------------------------------

template <int PSZ_T>
__global__ void k_dct_tiles(float *tiles, int n, int psz, int inverse)
{
    extern __shared__ float sm[];
    const int pp = psz * psz, TS = pp + 1;
    const int t0 = blockIdx.x * blockDim.x;
    const int cnt = min((int)blockDim.x, n - t0);
    for (int i = threadIdx.x; i < cnt * pp; i += blockDim.x) sm[(i / pp) * TS + i % pp] = tiles[(size_t)t0 * pp + i];
    __syncthreads();
    if ((int)threadIdx.x < cnt) {
        if (inverse) dct2d_tile<PSZ_T, true>(sm + threadIdx.x * TS, psz);
        else dct2d_tile<PSZ_T, false>(sm + threadIdx.x * TS, psz);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cnt * pp; i += blockDim.x) tiles[(size_t)t0 * pp + i] = sm[(i / pp) * TS + i % pp];
}

extern "C" int nlk_dct_host(nlk_ctx *c, float *h_tiles, int psz, int n, int inverse)
{
    if (int r = enter(c)) return r;
    if (psz < 1 || psz > MAX_PSZ || n < 0) return set_err(NLK_ERR_PARAM, "bad dct request");
    if (n == 0) return NLK_OK;
    const size_t bytes = (size_t)n * psz * psz * 4;
    if (int r = c->q_tmp.ensure(bytes)) return r;
    CU_TRY(cudaMemcpyAsync(c->q_tmp.p, h_tiles, bytes, cudaMemcpyHostToDevice, c->L->st));
    const int nt = psz > 12 ? 32 : 64, nb = (n + nt - 1) / nt;
    const size_t smem = (size_t)nt * (psz * psz + 1) * 4;
    if (psz == 8) k_dct_tiles<8><<<nb, nt, smem, c->L->st>>>(c->q_tmp.as<float>(), n, psz, inverse);
    else if (psz == 12) k_dct_tiles<12><<<nb, nt, smem, c->L->st>>>(c->q_tmp.as<float>(), n, psz, inverse);
    else k_dct_tiles<0><<<nb, nt, smem, c->L->st>>>(c->q_tmp.as<float>(), n, psz, inverse);
    if (int r = check_launch(c, 1, "dct_tiles")) return r;
    CU_TRY(cudaMemcpyAsync(h_tiles, c->q_tmp.p, bytes, cudaMemcpyDeviceToHost, c->L->st));
    CU_TRY(cudaStreamSynchronize(c->L->st));
    return NLK_OK;
}

// ---- the six drop-in entry points (include/nlkalman.h) ------------------------------------
// Host pointers in, host pointers out, synchronous, void: the only failure path is a
// message on stderr and exit(1), as in the reference (src/nlkalman.c:165-177).

static std::mutex g_legacy_mu;
static nlk_ctx *g_legacy = nullptr;

static void legacy_die(const char *what)
{
    fprintf(stderr, "nlkalman-b200: %s: %s\n", what, g_err.c_str());
    exit(1);
}

static nlk_ctx *legacy_ctx(int w, int h, int ch, nlk_ctx **slot = &g_legacy)
{
    if (*slot && ((*slot)->w != w || (*slot)->h != h || (*slot)->ch != ch)) {
        nlk_ctx_destroy(*slot);
        *slot = nullptr;
    }
    if (!*slot) {
        int dev = 0;
        if (const char *e = getenv("NLK_DEVICE")) dev = atoi(e);
        *slot = nlk_ctx_create(w, h, ch, dev);
        if (!*slot) legacy_die("no usable CUDA device (there is no CPU fallback)");
    }
    if (enter(*slot)) legacy_die("cudaSetDevice");
    return *slot;
}

static void legacy_colour(float *im, int w, int h, int ch, int inverse)
{
    if (ch != 3) return; // reference src/nlkalman.c:94, :114
    std::lock_guard<std::mutex> lk(g_legacy_mu);
    nlk_ctx *c = legacy_ctx(w, h, ch);
    const float *d;
    if (stage_in(c, c->s_in1, im, c->img_bytes(), &d)) legacy_die("colour transform");
    int r = inverse ? nlk_opp2rgb_dev(c, c->s_in1.as<float>(), d) : nlk_rgb2opp_dev(c, c->s_in1.as<float>(), d);
    if (r || cudaMemcpyAsync(im, c->s_in1.p, c->img_bytes(), cudaMemcpyDeviceToHost, c->L->st) != cudaSuccess ||
        cudaStreamSynchronize(c->L->st) != cudaSuccess) {
        if (!r) set_err(NLK_ERR_CUDA, "%s", cudaGetErrorString(cudaGetLastError()));
        legacy_die("colour transform");
    }
}

extern "C" void rgb2opp(float *im, int w, int h, int ch) { legacy_colour(im, w, h, ch, 0); }
extern "C" void opp2rgb(float *im, int w, int h, int ch) { legacy_colour(im, w, h, ch, 1); }

extern "C" void warp_bicubic(float *imw, float *im, float *of, float *msk, int w, int h, int ch)
{
    std::lock_guard<std::mutex> lk(g_legacy_mu);
    nlk_ctx *c = legacy_ctx(w, h, ch);
    const size_t ib = c->img_bytes(), npix = (size_t)w * h;
    const float *d_im, *d_of, *d_msk;
    if (stage_in(c, c->s_in1, im, ib, &d_im) || stage_in(c, c->s_of, of, npix * 2 * 4, &d_of) ||
        stage_in(c, c->s_msk, msk, npix * 4, &d_msk) || c->s_out.ensure(ib))
        legacy_die("warp_bicubic");
    if (nlk_warp_dev(c, c->s_out.as<float>(), d_im, d_of, d_msk)) legacy_die("warp_bicubic");
    if (cudaMemcpyAsync(imw, c->s_out.p, ib, cudaMemcpyDeviceToHost, c->L->st) != cudaSuccess ||
        cudaStreamSynchronize(c->L->st) != cudaSuccess) {
        set_err(NLK_ERR_CUDA, "%s", cudaGetErrorString(cudaGetLastError()));
        legacy_die("warp_bicubic");
    }
}

extern "C" void nlkalman_default_params(struct nlkalman_params *p, float sigma, enum FILTER_MODE mode)
{
    // host arithmetic, the reference's expressions (src/nlkalman.c:456-486): double
    // literals, truncation to int, float comparison in the max()
    if (p->patch_sz < 0) p->patch_sz = 8;
    if (p->search_sz_x < 0) p->search_sz_x = 10;
    if (p->search_sz_t < 0) p->search_sz_t = 5;
    if (p->dista_lambda < 0) p->dista_lambda = 1.0f;
    switch (mode) {
    case FLT1:
        if (p->npatches_x < 0) p->npatches_x = (int)(0.5 * sigma + 40.);
        if (p->beta_x < 0) p->beta_x = (float)(-0.04 * sigma + 3.91);
        if (p->npatches_t < 0) p->npatches_t = 30;
        if (p->npatches_tagg < 0) p->npatches_tagg = 20;
        if (p->beta_t < 0) p->beta_t = (float)(-0.005 * sigma + 2.05);
        break;
    case FLT2:
        if (p->npatches_x < 0) p->npatches_x = (int)(0.5 * sigma + 10.);
        if (p->beta_x < 0) p->beta_x = (float)(0.004 * sigma + 0.21);
        if (p->npatches_t < 0) p->npatches_t = (int)(5.f > sigma ? 5.f : sigma);
        if (p->npatches_tagg < 0) p->npatches_tagg = 1;
        if (p->beta_t < 0) p->beta_t = (float)(0.014 * sigma + 1.38);
        break;
    case SMO1:
        if (p->npatches_x < 0) p->npatches_x = 0;
        if (p->beta_x < 0) p->beta_x = 0;
        if (p->npatches_t < 0) {
            const float v = 3 * sigma - 15;
            p->npatches_t = (int)(5.f > v ? 5.f : v);
        }
        if (p->npatches_tagg < 0) p->npatches_tagg = p->npatches_t;
        if (p->beta_t < 0) {
            const double v = -0.14 * sigma + 8.0;
            p->beta_t = (float)(1.0 > v ? 1.0 : v);
        }
        break;
    }
}

extern "C" void nlkalman_filter_frame(float *deno1, float *nisy1, float *deno0, float *bsic1, int w,
                                      int h, int ch, float sigma, const struct nlkalman_params prms,
                                      int frame)
{
    (void)frame;
    std::lock_guard<std::mutex> lk(g_legacy_mu);
    nlk_ctx *c = legacy_ctx(w, h, ch);
    if (pass_host(c, 0, deno1, nisy1, deno0, bsic1, sigma, prms, false)) legacy_die("nlkalman_filter_frame");
}

extern "C" void nlkalman_smooth_frame(float *smoo1, float *filt1, float *smoo0, float *bsic1, int w,
                                      int h, int ch, float sigma, const struct nlkalman_params prms,
                                      int frame)
{
    (void)frame;
    std::lock_guard<std::mutex> lk(g_legacy_mu);
    nlk_ctx *c = legacy_ctx(w, h, ch);
    if (pass_host(c, 1, smoo1, filt1, smoo0, bsic1, sigma, prms, false)) legacy_die("nlkalman_smooth_frame");
}

// ---- fp32 FMA peak (roofline denominator for the compute-bound kernels) ----------------------

__global__ void __launch_bounds__(256) k_fma_peak(float *out, int iters, float a, float b)
{
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = (float)(threadIdx.x + i) * 1e-3f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    if (s == 123.456f) out[0] = s; // never true in practice: keeps the chain live
}

extern "C" int nlk_fp32_peak(nlk_ctx *c, float ms, double *tflops)
{
    if (int r = enter(c)) return r;
    if (int r = c->q_tmp.ensure(256)) return r;
    const int nb = c->num_sms * 8, nt = 256, iters = 4096;
    cudaEvent_t a, b;
    CU_TRY(cudaEventCreate(&a));
    CU_TRY(cudaEventCreate(&b));
    // warm up, then repeat until about `ms` of device time has been measured
    for (int i = 0; i < 3; ++i) k_fma_peak<<<nb, nt, 0, c->L->st>>>(c->q_tmp.as<float>(), iters, 0.999f, 1e-3f);
    double best = 0, spent = 0;
    while (spent < ms) {
        CU_TRY(cudaEventRecord(a, c->L->st));
        for (int i = 0; i < 8; ++i) k_fma_peak<<<nb, nt, 0, c->L->st>>>(c->q_tmp.as<float>(), iters, 0.999f, 1e-3f);
        CU_TRY(cudaEventRecord(b, c->L->st));
        CU_TRY(cudaEventSynchronize(b));
        float t = 0.f;
        CU_TRY(cudaEventElapsedTime(&t, a, b));
        const double fl = 8.0 * nb * nt * (double)iters * 16 * 2;
        const double tf = fl / (t * 1e-3) / 1e12;
        if (tf > best) best = tf;
        spent += t;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *tflops = best;
    return NLK_OK;
}

// ---- the two entry points of the reference's TV-L1 library (include/tvl1flow.h) ----------------------
// (a context of their own: a caller alternating between the flow of single-channel images and the filter
// of colour frames keeps both)
static nlk_ctx *g_legacy_flow = nullptr;

extern "C" void Dual_TVL1_optic_flow(float *I0, float *I1, float *u1, float *u2, const int nx, const int ny,
                                     const float tau, const float lambda, const float theta, const int warps,
                                     const float epsilon, const bool verbose)
{
    std::lock_guard<std::mutex> lk(g_legacy_mu);
    nlk_ctx *c = legacy_ctx(nx, ny, 1, &g_legacy_flow);
    std::vector<int> its(warps > 0 ? warps : 1, 0);
    if (nlk_tvl1_level_host(c, I0, I1, u1, u2, nx, ny, tau, lambda, theta, warps, epsilon, its.data()))
        legacy_die("Dual_TVL1_optic_flow");
    if (verbose)     // (reference tvl1flow_lib.c:251-254, without the error value, which stays on the device)
        for (int k = 0; k < warps; ++k) fprintf(stderr, "Warping: %d, Iterations: %d\n", k, its[k]);
}

extern "C" void Dual_TVL1_optic_flow_multiscale(float *I0, float *I1, float *u1, float *u2, const int nxx, const int nyy,
                                                const float tau, const float lambda, const float theta, const int nscales,
                                                const int fscale, const float zfactor, const int warps,
                                                const float epsilon, const bool verbose)
{
    std::lock_guard<std::mutex> lk(g_legacy_mu);
    nlk_ctx *c = legacy_ctx(nxx, nyy, 1, &g_legacy_flow);
    const size_t n = (size_t)nxx * nyy;
    std::vector<float> flow(2 * n);
    std::vector<int> its((size_t)(nscales > 0 ? nscales : 1) * (warps > 0 ? warps : 1), 0);
    if (nlk_tvl1_flow_host(c, I0, I1, flow.data(), nxx, nyy, tau, lambda, theta, nscales, fscale, zfactor, warps, epsilon,
                           its.data()))
        legacy_die("Dual_TVL1_optic_flow_multiscale");
    memcpy(u1, flow.data(), n * 4);
    memcpy(u2, flow.data() + n, n * 4);
    if (verbose) {   // (reference tvl1flow_lib.c:413-414, :251-254)
        Tvl1Pyramid P;
        P.plan(nxx, nyy, nscales, fscale, zfactor, warps);
        for (int s = nscales - 1; s >= fscale; --s) {
            fprintf(stderr, "Scale %d: %dx%d\n", s, P.nx[s], P.ny[s]);
            for (int k = 0; k < warps; ++k) fprintf(stderr, "Warping: %d, Iterations: %d\n", k, its[(size_t)s * warps + k]);
        }
    }
}
