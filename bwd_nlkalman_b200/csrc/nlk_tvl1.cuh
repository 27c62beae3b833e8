// Dual TV-L1 optical flow.  One scale: what the reference's
// Dual_TVL1_optic_flow computes (lib/tvl1flow/tvl1flow_lib.c:93-280, Zach-Pock-Bischof with
// Chambolle's dual update), the heavy per-frame step in front of the filter in
// scripts/nlkalman-seq.sh:60-65.  All of it is per-pixel / 5-point-stencil work on ~14 float planes:
// HBM / L2 bound at the fine scales, launch-latency bound at the coarse ones (where most iterations
// run).  An iteration is two per-pixel steps,
//
//   tvl1_u_pixel:  thresholding step v = TH(u) (:172-206), divergence of the dual variable
//                  (mask.c:43-94), u = v + theta div p and the squared update for the stopping test (:213-227)
//   tvl1_p_pixel:  forward gradient of the new u (mask.c:101-144), dual update p (:235-248)
//
// and the three bicubic warps of a warping step (I1, dI1/dx, dI1/dy at the same positions) plus rho_c
// and |grad|^2 (:140-157) are one kernel.  The stopping rule `error > eps^2 && n < 300` (:164) is evaluated
// ON THE DEVICE in every form the loop of a warping step takes (nlk_lib.cu: nlk_tvl1_level_dev):
//   k_tvl1_iterate          all iterations in one cooperative launch, grid barriers between the steps (default)
//   k_tvl1_u + k_tvl1_p     the body of a CUDA graph WHILE node (k_tvl1_p sets the loop condition), or
//                           queued in batches by the host, each kernel returning at once when the previous
//                           iteration met the stopping rule
// so the result is always that of the exact stopping iteration.
//
// Around the level solver, the pyramid of Dual_TVL1_optic_flow_multiscale (:345-477): joint
// normalisation, separable Gaussian, zoom out / zoom in (kernels at the end of this file; the sequence of
// steps is in nlk_tvl1_pyramid.h).
//
// All arithmetic is written with explicit roundings (__fmul_rn, __dadd_rn ...): the reference is built
// without FMA contraction (x86-64 baseline, no -ffast-math: lib/tvl1flow/CMakeLists.txt:18), so a * b + c
// must round twice here as well for the flow to come out bit for bit.  The per-pixel device functions
// are also compiled for the HOST by tests/models/tvl1_host_model.cpp (NLK_HOST_MODEL: the intrinsics
// become plain operators, a kernel becomes a loop over its grid) and checked there against the
// reference's library without a GPU.
#pragma once
#ifndef NLK_HOST_MODEL
#include "nlk_common.cuh"
#endif
#include "nlk_tvl1_pyramid.h"

namespace nlk {

constexpr int TVL1_MAX_ITERATIONS = 300;      // reference tvl1flow_lib.c:21
constexpr float TVL1_GRAD_IS_ZERO = 1e-10f;   // :23

// centred differences with the reference's one-sided borders (mask.c:152-215)
__global__ void k_tvl1_centered_gradient(const float *__restrict__ in, float *__restrict__ dx, float *__restrict__ dy,
                                         int nx, int ny)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
    if (j >= nx || i >= ny) return;
    const int k = i * nx + j;
    const float xl = in[j > 0 ? k - 1 : k], xr = in[j < nx - 1 ? k + 1 : k];
    const float yu = in[i > 0 ? k - nx : k], yd = in[i < ny - 1 ? k + nx : k];
    dx[k] = 0.5f * (xr - xl);
    dy[k] = 0.5f * (yd - yu);
}

__device__ __forceinline__ int tvl1_neumann(int x, int n, bool &out)
{
    if (x < 0) { out = true; return 0; }
    if (x >= n) { out = true; return n - 1; }
    return x;
}

// cubic_interpolation_cell, in double like the reference (bicubic_interpolation.c:102-110)
__device__ __forceinline__ double tvl1_cubic(double v0, double v1, double v2, double v3, double x)
{
    const double c3 = __dsub_rn(__dadd_rn(__dmul_rn(3.0, __dsub_rn(v1, v2)), v3), v0);
    const double c2 = __dadd_rn(__dsub_rn(__dadd_rn(__dsub_rn(__dmul_rn(2.0, v0), __dmul_rn(5.0, v1)), __dmul_rn(4.0, v2)), v3),
                                __dmul_rn(x, c3));
    const double c1 = __dadd_rn(__dsub_rn(v2, v0), __dmul_rn(x, c2));
    return __dadd_rn(v1, __dmul_rn(__dmul_rn(0.5, x), c1));
}

// bicubic_interpolation_at with border_out = true (bicubic_interpolation.c:138-233), Neumann
// boundary (BOUNDARY_CONDITION 0), including its use of the x step for the row above (`my`, :157)
struct Tvl1Taps { int xs[4], ys[4]; double fx, fy; bool out; };
__device__ __forceinline__ Tvl1Taps tvl1_taps(float uu, float vv, int nx, int ny)
{
    Tvl1Taps t;
    const int sx = uu < 0 ? -1 : 1, sy = vv < 0 ? -1 : 1;
    bool out = false;
    const int x = tvl1_neumann((int)uu, nx, out), y = tvl1_neumann((int)vv, ny, out);
    t.xs[0] = tvl1_neumann((int)uu - sx, nx, out);
    t.ys[0] = tvl1_neumann((int)vv - sx, ny, out);       // (sic) the reference steps by sx here
    t.xs[1] = x;
    t.ys[1] = y;
    t.xs[2] = tvl1_neumann((int)uu + sx, nx, out);
    t.ys[2] = tvl1_neumann((int)vv + sy, ny, out);
    t.xs[3] = tvl1_neumann((int)uu + 2 * sx, nx, out);
    t.ys[3] = tvl1_neumann((int)vv + 2 * sy, ny, out);
    t.fx = (double)(uu - (float)x);
    t.fy = (double)(vv - (float)y);
    t.out = out;
    return t;
}
template <bool BORDER_OUT = true>
__device__ __forceinline__ float tvl1_bicubic(const float *__restrict__ img, const Tvl1Taps &t, int nx)
{
    if (BORDER_OUT && t.out) return 0.f;
    double col[4];
#pragma unroll
    for (int a = 0; a < 4; ++a)   // pol[a][b] = input[xs[a] + nx * ys[b]], interpolated along y first
        col[a] = tvl1_cubic(img[t.xs[a] + nx * t.ys[0]], img[t.xs[a] + nx * t.ys[1]], img[t.xs[a] + nx * t.ys[2]],
                            img[t.xs[a] + nx * t.ys[3]], t.fy);
    return (float)tvl1_cubic(col[0], col[1], col[2], col[3], t.fx);
}

// one warping step: I1, I1x, I1y at (j + u1, i + u2); |grad|^2 and the constant part of rho (:140-157)
__global__ void k_tvl1_warp(const float *__restrict__ I0, const float *__restrict__ I1, const float *__restrict__ I1x,
                            const float *__restrict__ I1y, const float *__restrict__ u1, const float *__restrict__ u2,
                            float *__restrict__ I1wx, float *__restrict__ I1wy, float *__restrict__ grad,
                            float *__restrict__ rho_c, int nx, int ny)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
    if (j >= nx || i >= ny) return;
    const int p = i * nx + j;
    const float a = u1[p], b = u2[p];
    const Tvl1Taps t = tvl1_taps((float)(j + a), (float)(i + b), nx, ny);
    const float w = tvl1_bicubic(I1, t, nx), wx = tvl1_bicubic(I1x, t, nx), wy = tvl1_bicubic(I1y, t, nx);
    I1wx[p] = wx;
    I1wy[p] = wy;
    grad[p] = __fadd_rn(__fmul_rn(wx, wx), __fmul_rn(wy, wy));
    rho_c[p] = __fsub_rn(__fsub_rn(__fsub_rn(w, __fmul_rn(wx, a)), __fmul_rn(wy, b)), I0[p]);
}

// divergence with the reference's border rules (mask.c:43-94): backward differences, the field taken
// as zero beyond the first row / column and its last row / column dropped
__device__ __forceinline__ float tvl1_div(const float *__restrict__ v1, const float *__restrict__ v2, int i, int j,
                                          int nx, int ny)
{
    const int p = i * nx + j;
    // (the first and last columns sum left to right in the reference, mask.c:81-82: keep its association)
    if ((j == 0 || j == nx - 1) && i > 0 && i < ny - 1)
        return __fsub_rn(__fadd_rn(j == 0 ? v1[p] : -v1[p - 1], v2[p]), v2[p - nx]);
    const float a = j == 0 ? v1[p] : (j == nx - 1 ? -v1[p - 1] : __fsub_rn(v1[p], v1[p - 1]));
    const float b = i == 0 ? v2[p] : (i == ny - 1 ? -v2[p - nx] : __fsub_rn(v2[p], v2[p - nx]));
    return __fadd_rn(a, b);
}

// err[n]: sum over the pixels of the squared update of iteration n (n = 1 ..); err[0] unused.
// An iteration runs iff n == 1 or err[n-1] / size > eps2 (the reference's while condition, :164).
__device__ __forceinline__ bool tvl1_runs(const float *err, int n, float size, float eps2)
{
    return n == 1 || __fdiv_rn(err[n - 1], size) > eps2;
}

// thresholding step, divergence and flow update of one pixel; returns its squared update
__device__ __forceinline__ float tvl1_u_pixel(const float *__restrict__ rho_c, const float *__restrict__ I1wx,
                                              const float *__restrict__ I1wy, const float *__restrict__ grad,
                                              const float *__restrict__ p11, const float *__restrict__ p12,
                                              const float *__restrict__ p21, const float *__restrict__ p22,
                                              float *__restrict__ u1, float *__restrict__ u2, int i, int j, int nx, int ny,
                                              float l_t, float theta)
{
    const int p = i * nx + j;
    const float wx = I1wx[p], wy = I1wy[p], g = grad[p], a = u1[p], b = u2[p];
    const float rho = __fadd_rn(rho_c[p], __fadd_rn(__fmul_rn(wx, a), __fmul_rn(wy, b)));
    const float lg = __fmul_rn(l_t, g);
    float d1, d2;
    if (rho < -lg) { d1 = __fmul_rn(l_t, wx); d2 = __fmul_rn(l_t, wy); }
    else if (rho > lg) { d1 = __fmul_rn(-l_t, wx); d2 = __fmul_rn(-l_t, wy); }
    else if (g < TVL1_GRAD_IS_ZERO) { d1 = d2 = 0.f; }
    else { const float fi = __fdiv_rn(-rho, g); d1 = __fmul_rn(fi, wx); d2 = __fmul_rn(fi, wy); }
    const float na = __fadd_rn(__fadd_rn(a, d1), __fmul_rn(theta, tvl1_div(p11, p12, i, j, nx, ny)));
    const float nb = __fadd_rn(__fadd_rn(b, d2), __fmul_rn(theta, tvl1_div(p21, p22, i, j, nx, ny)));
    u1[p] = na;
    u2[p] = nb;
    const float da = __fsub_rn(na, a), db = __fsub_rn(nb, b);
    return __fadd_rn(__fmul_rn(da, da), __fmul_rn(db, db));
}

#ifndef NLK_HOST_MODEL
__global__ void __launch_bounds__(256) k_tvl1_u(const float *__restrict__ rho_c, const float *__restrict__ I1wx,
                                                const float *__restrict__ I1wy, const float *__restrict__ grad,
                                                const float *__restrict__ p11, const float *__restrict__ p12,
                                                const float *__restrict__ p21, const float *__restrict__ p22,
                                                float *__restrict__ u1, float *__restrict__ u2, float *err, int n,
                                                const int *__restrict__ n_loop, int nx, int ny, float l_t, float theta,
                                                float eps2)
{
    // stream path: iteration n, skipped when the previous one met the stopping rule; graph path (n_loop):
    // the loop counter in device memory, the loop condition has already decided that it runs
    if (n_loop) n = *n_loop;
    else if (!tvl1_runs(err, n, (float)(nx * ny), eps2)) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
    float e = 0.f;
    if (j < nx && i < ny) e = tvl1_u_pixel(rho_c, I1wx, I1wy, grad, p11, p12, p21, p22, u1, u2, i, j, nx, ny, l_t, theta);
    // block sum, one atomic per block
    __shared__ float s_red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    const int t = threadIdx.y * blockDim.x + threadIdx.x;
    if ((t & 31) == 0) s_red[t >> 5] = e;
    __syncthreads();
    if (t < 8) {
        e = s_red[t];
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) e += __shfl_xor_sync(0xffu, e, o);
        if (t == 0) atomicAdd(err + n, e);
    }
}
#endif

// forward gradient of u and the dual update of one pixel (mask.c:101-144, tvl1flow_lib.c:235-248)
__device__ __forceinline__ void tvl1_p_pixel(const float *u1, const float *u2, float *p11, float *p12, float *p21, float *p22,
                                             int i, int j, int nx, int ny, float taut)
{
    const int p = i * nx + j;
    // forward differences, zero across the last column / row
    const float a = u1[p], b = u2[p];
    const float u1x = j < nx - 1 ? __fsub_rn(u1[p + 1], a) : 0.f, u1y = i < ny - 1 ? __fsub_rn(u1[p + nx], a) : 0.f;
    const float u2x = j < nx - 1 ? __fsub_rn(u2[p + 1], b) : 0.f, u2y = i < ny - 1 ? __fsub_rn(u2[p + nx], b) : 0.f;
    // (the reference's hypot and `1.0 +` are double: :239-242)
    const double g1 = hypot((double)u1x, (double)u1y), g2 = hypot((double)u2x, (double)u2y);
    const float ng1 = (float)__dadd_rn(1.0, (double)__fmul_rn(taut, (float)g1));
    const float ng2 = (float)__dadd_rn(1.0, (double)__fmul_rn(taut, (float)g2));
    p11[p] = __fdiv_rn(__fadd_rn(p11[p], __fmul_rn(taut, u1x)), ng1);
    p12[p] = __fdiv_rn(__fadd_rn(p12[p], __fmul_rn(taut, u1y)), ng1);
    p21[p] = __fdiv_rn(__fadd_rn(p21[p], __fmul_rn(taut, u2x)), ng2);
    p22[p] = __fdiv_rn(__fadd_rn(p22[p], __fmul_rn(taut, u2y)), ng2);
}

__global__ void __launch_bounds__(256) k_tvl1_p(const float *__restrict__ u1, const float *__restrict__ u2,
                                                float *__restrict__ p11, float *__restrict__ p12,
                                                float *__restrict__ p21, float *__restrict__ p22, const float *err, int n,
                                                int nx, int ny, float taut, float eps2, unsigned long long loop,
                                                int *n_loop, int *count)
{
    // stream path (n_loop == nullptr): iteration n, skipped when the previous one met the stopping rule.
    // Graph path: the loop of a warping step is a CUDA graph WHILE node whose body is { k_tvl1_u, k_tvl1_p }
    // (nlk_lib.cu: tvl1_build_graph); the error of this iteration is complete (k_tvl1_u has finished), so
    // one thread closes it here: records the count, advances the counter (nobody reads it before the
    // next k_tvl1_u) and sets the loop condition to the reference's `error > eps^2 && n < MAX_ITERATIONS`
    // (:164) -- no host round trip, no launch past the stopping iteration.
    if (n_loop == nullptr) {
        if (!tvl1_runs(err, n, (float)(nx * ny), eps2)) return;
    }
#ifndef NLK_HOST_MODEL
    else if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && threadIdx.y == 0) {
        const int m = *n_loop;
        *count = m;
        *n_loop = m + 1;
        cudaGraphSetConditional((cudaGraphConditionalHandle)loop,
                                (m < TVL1_MAX_ITERATIONS && __fdiv_rn(err[m], (float)(nx * ny)) > eps2) ? 1u : 0u);
    }
#endif
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
    if (j >= nx || i >= ny) return;
    tvl1_p_pixel(u1, u2, p11, p12, p21, p22, i, j, nx, ny, taut);
}

#ifndef NLK_HOST_MODEL
// ---- the iterations of a warping step in ONE launch ----------------------------------------------------
// Most iterations of a pyramid run on scales of a few ten thousand pixels, where a kernel per half
// iteration is all launch latency.  k_tvl1_iterate keeps a co-resident grid (cooperative launch) for the
// whole warping step: u-phase, grid barrier, p-phase and stopping test, grid barrier, next iteration.
// Same per-pixel functions, same in-place updates, so the same flow bit for bit; the error of an iteration
// is complete after the first barrier and every thread evaluates the reference's loop condition (:164) on it.
constexpr int TVL1_IT_THREADS = 512;
constexpr long long TVL1_BARRIER_LIMIT = 400000000ll;      // cycles (~0.2 s): a barrier that long is a bug, not a wait

__device__ __forceinline__ unsigned tvl1_ld_acquire(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// all blocks of the (co-resident) grid; `target` = barriers so far * gridDim.x.  False: gave up waiting.
// Release / acquire at GPU scope through thread 0 (the block barriers on either side make it cumulative
// for the block's other threads; the acquiring load also drops the SM's stale L1 lines), no full fences.
__device__ __forceinline__ bool tvl1_grid_barrier(unsigned *bar, unsigned target, int *s_fail)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(bar) : "memory");
        const long long t0 = clock64();
        int fail = 0;
        while (tvl1_ld_acquire(bar) < target)
            if (clock64() - t0 > TVL1_BARRIER_LIMIT) { fail = 1; break; }
        *s_fail = fail;
    }
    __syncthreads();
    return *s_fail == 0;
}

__global__ void __launch_bounds__(TVL1_IT_THREADS) k_tvl1_iterate(const float *__restrict__ rho_c, const float *__restrict__ I1wx,
                                                                  const float *__restrict__ I1wy, const float *__restrict__ grad,
                                                                  float *p11, float *p12, float *p21, float *p22, float *u1,
                                                                  float *u2, float *err, int *count, unsigned *bar,
                                                                  int *host_flag, int nx, int ny, float l_t, float theta,
                                                                  float taut, float eps2)
{
    __shared__ float s_red[TVL1_IT_THREADS / 32];
    __shared__ int s_fail;
    const int size = nx * ny;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned arrived = 0;
    int n = 1;
    for (;; ++n) {
        float e = 0.f;
        for (int p = tid; p < size; p += nthr) {
            const int i = p / nx, j = p - i * nx;
            e += tvl1_u_pixel(rho_c, I1wx, I1wy, grad, p11, p12, p21, p22, u1, u2, i, j, nx, ny, l_t, theta);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
        if (lane == 0) s_red[warp] = e;
        __syncthreads();
        if (warp == 0) {
            e = lane < TVL1_IT_THREADS / 32 ? s_red[lane] : 0.f;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
            if (lane == 0) atomicAdd(err + n, e);
        }
        arrived += gridDim.x;
        if (!tvl1_grid_barrier(bar, arrived, &s_fail)) break;
        for (int p = tid; p < size; p += nthr) {
            const int i = p / nx, j = p - i * nx;
            tvl1_p_pixel(u1, u2, p11, p12, p21, p22, i, j, nx, ny, taut);
        }
        const bool more = n < TVL1_MAX_ITERATIONS && __fdiv_rn(__ldcg(err + n), (float)size) > eps2;
        if (!more) break;           // (uniform over the grid: everybody reads the same completed sum)
        arrived += gridDim.x;
        if (!tvl1_grid_barrier(bar, arrived, &s_fail)) break;
    }
    if (tid == 0) *count = n;
    if (threadIdx.x == 0 && s_fail) *host_flag = 1;      // (pinned host memory: the host sees it without a copy)
}
#endif

#ifndef NLK_HOST_MODEL
__global__ void k_tvl1_loop_init(int *n_loop, int warps)
{
    if ((int)threadIdx.x < warps) n_loop[threadIdx.x] = 1;
}
#endif

// iterations run by a warping step = the last n that passed the test
__global__ void k_tvl1_count(const float *err, int *count, float size, float eps2)
{
    int n = 1;
    while (n < TVL1_MAX_ITERATIONS && __fdiv_rn(err[n], size) > eps2) ++n;
    *count = n;
}

// ---- the pyramid around the level solver (Dual_TVL1_optic_flow_multiscale, tvl1flow_lib.c:345-477) ----

// image_normalization (tvl1flow_lib.c:305-337): 255 (I - min) / (max - min) in double, or a copy
__device__ __forceinline__ float tvl1_norm_pixel(float a, float mn, float den)
{
    return den > 0.f ? (float)__ddiv_rn(__dmul_rn(255.0, (double)__fsub_rn(a, mn)), (double)den) : a;
}

#ifndef NLK_HOST_MODEL
// joint extrema of the two images, stage 1 (getminmax, tvl1flow_lib.c:283-298): part[2b] = min, part[2b+1] = max
__global__ void __launch_bounds__(256) k_tvl1_minmax(const float *__restrict__ I0, const float *__restrict__ I1, size_t size,
                                                     float *__restrict__ part)
{
    float mn = I0[0], mx = mn;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < size; i += (size_t)gridDim.x * blockDim.x) {
        const float a = I0[i], b = I1[i];
        if (a < mn) mn = a;
        if (a > mx) mx = a;
        if (b < mn) mn = b;
        if (b > mx) mx = b;
    }
    __shared__ float s_mn[8], s_mx[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) { s_mn[threadIdx.x >> 5] = mn; s_mx[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) { mn = fminf(mn, s_mn[k]); mx = fmaxf(mx, s_mx[k]); }
        part[2 * blockIdx.x] = mn;
        part[2 * blockIdx.x + 1] = mx;
    }
}

// stage 2 + image_normalization (tvl1flow_lib.c:305-337): 255 (I - min) / (max - min) in double, or a copy
__global__ void __launch_bounds__(256) k_tvl1_normalize(const float *__restrict__ I0, const float *__restrict__ I1,
                                                        float *__restrict__ O0, float *__restrict__ O1, size_t size,
                                                        const float *__restrict__ part, int nparts)
{
    __shared__ float s_mn[8], s_mx[8];
    float mn = part[0], mx = part[1];
    for (int k = threadIdx.x; k < nparts; k += blockDim.x) { mn = fminf(mn, part[2 * k]); mx = fmaxf(mx, part[2 * k + 1]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) { s_mn[threadIdx.x >> 5] = mn; s_mx[threadIdx.x >> 5] = mx; }
    __syncthreads();
    mn = s_mn[0]; mx = s_mx[0];
    for (int k = 1; k < 8; ++k) { mn = fminf(mn, s_mn[k]); mx = fmaxf(mx, s_mx[k]); }
    const float den = __fsub_rn(mx, mn);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < size; i += (size_t)gridDim.x * blockDim.x) {
        const float a = I0[i], b = I1[i];
        O0[i] = tvl1_norm_pixel(a, mn, den);
        O1[i] = tvl1_norm_pixel(b, mn, den);
    }
}
#endif

// One direction of the separable Gaussian (mask.c:216-330): double accumulation in the reference's order,
// its boundary rule (the left / top side mirrors about the border sample, the right / bottom side
// repeats it: R[-m] = I[m], R[n + m] = I[n - 1 - m]), float result.  B: the normalised half kernel.
template <bool ROWS>
__global__ void __launch_bounds__(256) k_tvl1_gauss(const float *__restrict__ in, float *__restrict__ out, int nx, int ny,
                                                    const double *__restrict__ B, int taps)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
    if (j >= nx || i >= ny) return;
    const int n = ROWS ? nx : ny, q = ROWS ? j : i;
    const float *line = ROWS ? in + (size_t)i * nx : in + j;
    const int step = ROWS ? 1 : nx;
    double sum = __dmul_rn(B[0], (double)line[(size_t)q * step]);
    for (int k = 1; k < taps; ++k) {
        int a = q - k, b = q + k;
        if (a < 0) a = -a;
        if (b >= n) b = 2 * n - 1 - b;
        sum = __dadd_rn(sum, __dmul_rn(B[k], __dadd_rn((double)line[(size_t)a * step], (double)line[(size_t)b * step])));
    }
    out[(size_t)i * nx + j] = (float)sum;
}

// zoom_out's resampling and zoom_in (zoom.c:70-79, :104-112): bicubic at (j1 / fx, i1 / fy), samples
// beyond the border clamped (border_out = false); `scale` is the 1 / zfactor of the flow (1 for images)
__global__ void k_tvl1_zoom(const float *__restrict__ in, float *__restrict__ out, int nx, int ny, int nxx, int nyy,
                            float fx, float fy, float scale, int scaled)
{
    const int j1 = blockIdx.x * blockDim.x + threadIdx.x, i1 = blockIdx.y * blockDim.y + threadIdx.y;
    if (j1 >= nxx || i1 >= nyy) return;
    const Tvl1Taps t = tvl1_taps(__fdiv_rn((float)j1, fx), __fdiv_rn((float)i1, fy), nx, ny);
    const float g = tvl1_bicubic<false>(in, t, nx);
    out[(size_t)i1 * nxx + j1] = scaled ? __fmul_rn(g, scale) : g;
}

// ---- between the frames of the filter and the estimator (the pipeline script's file interfaces, kept on
// the device): luminance the way the reference's program reads a colour file (iio_read_image_float,
// lib/iio/iio.c:1048-1055: .299 R + .587 G + .114 B in double, rounded to float), and the flow's two
// planes interleaved as the warp and the occlusion mask take them (.flo layout)
__device__ __forceinline__ float tvl1_luma(float r, float g, float b)
{
    return (float)__dadd_rn(__dadd_rn(__dmul_rn(.299, (double)r), __dmul_rn(.587, (double)g)), __dmul_rn(.114, (double)b));
}

#ifndef NLK_HOST_MODEL
__global__ void __launch_bounds__(256) k_tvl1_luma(const float *__restrict__ img, float *__restrict__ lum, size_t npix, int ch)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x)
        lum[i] = ch == 1 ? img[i] : tvl1_luma(img[i * ch], img[i * ch + 1], img[i * ch + 2]);
}

__global__ void __launch_bounds__(256) k_tvl1_interleave(const float *__restrict__ u1, const float *__restrict__ u2,
                                                         float *__restrict__ of, size_t npix)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x)
        reinterpret_cast<float2 *>(of)[i] = make_float2(u1[i], u2[i]);
}
#endif

} // namespace nlk
