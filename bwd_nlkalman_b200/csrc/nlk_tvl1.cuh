// Dual TV-L1 optical flow at one scale (first slice): what the reference's
// Dual_TVL1_optic_flow computes (lib/tvl1flow/tvl1flow_lib.c:93-280, Zach-Pock-Bischof with
// Chambolle's dual update), the heavy per-frame step in front of the filter in
// scripts/nlkalman-seq.sh:60-65.  All of it is per-pixel / 5-point-stencil work on ~14 float planes:
// HBM / L2 bound, so the steps of an iteration are fused into two kernels,
//
//   k_tvl1_u:  thresholding step v = TH(u) (:172-206), divergence of the dual variable
//              (mask.c:43-94), u = v + theta div p and the squared update for the stopping test (:213-227)
//   k_tvl1_p:  forward gradient of the new u (mask.c:101-144), dual update p (:235-248)
//
// and the three bicubic warps of a warping step (I1, dI1/dx, dI1/dy at the same positions) plus rho_c
// and |grad|^2 (:140-157) into one.  The stopping rule `error > eps^2 && n < 300` (:164) is evaluated
// ON THE DEVICE: an iteration's kernels look at the error of the previous one and return at once when
// it is below the threshold, so a batch of iterations can be queued without a host round trip and
// the result is that of the exact stopping iteration.
#pragma once
#include "nlk_common.cuh"

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
    return v1 + 0.5 * x * (v2 - v0 + x * (2.0 * v0 - 5.0 * v1 + 4.0 * v2 - v3 + x * (3.0 * (v1 - v2) + v3 - v0)));
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
__device__ __forceinline__ float tvl1_bicubic(const float *__restrict__ img, const Tvl1Taps &t, int nx)
{
    if (t.out) return 0.f;
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
    const float a = j == 0 ? v1[p] : (j == nx - 1 ? -v1[p - 1] : v1[p] - v1[p - 1]);
    const float b = i == 0 ? v2[p] : (i == ny - 1 ? -v2[p - nx] : v2[p] - v2[p - nx]);
    return a + b;
}

// err[n]: sum over the pixels of the squared update of iteration n (n = 1 ..); err[0] unused.
// An iteration runs iff n == 1 or err[n-1] / size > eps2 (the reference's while condition, :164).
__device__ __forceinline__ bool tvl1_runs(const float *err, int n, float size, float eps2)
{
    return n == 1 || __fdiv_rn(err[n - 1], size) > eps2;
}

__global__ void __launch_bounds__(256) k_tvl1_u(const float *__restrict__ rho_c, const float *__restrict__ I1wx,
                                                const float *__restrict__ I1wy, const float *__restrict__ grad,
                                                const float *__restrict__ p11, const float *__restrict__ p12,
                                                const float *__restrict__ p21, const float *__restrict__ p22,
                                                float *__restrict__ u1, float *__restrict__ u2, float *err, int n,
                                                int nx, int ny, float l_t, float theta, float eps2)
{
    if (!tvl1_runs(err, n, (float)(nx * ny), eps2)) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
    float e = 0.f;
    if (j < nx && i < ny) {
        const int p = i * nx + j;
        const float wx = I1wx[p], wy = I1wy[p], g = grad[p], a = u1[p], b = u2[p];
        const float rho = rho_c[p] + (wx * a + wy * b);
        float d1, d2;
        if (rho < -l_t * g) { d1 = l_t * wx; d2 = l_t * wy; }
        else if (rho > l_t * g) { d1 = -l_t * wx; d2 = -l_t * wy; }
        else if (g < TVL1_GRAD_IS_ZERO) { d1 = d2 = 0.f; }
        else { const float fi = -rho / g; d1 = fi * wx; d2 = fi * wy; }
        const float na = (a + d1) + theta * tvl1_div(p11, p12, i, j, nx, ny);
        const float nb = (b + d2) + theta * tvl1_div(p21, p22, i, j, nx, ny);
        u1[p] = na;
        u2[p] = nb;
        e = (na - a) * (na - a) + (nb - b) * (nb - b);
    }
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

__global__ void __launch_bounds__(256) k_tvl1_p(const float *__restrict__ u1, const float *__restrict__ u2,
                                                float *__restrict__ p11, float *__restrict__ p12,
                                                float *__restrict__ p21, float *__restrict__ p22, const float *err, int n,
                                                int nx, int ny, float taut, float eps2)
{
    if (!tvl1_runs(err, n, (float)(nx * ny), eps2)) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
    if (j >= nx || i >= ny) return;
    const int p = i * nx + j;
    // forward differences, zero across the last column / row (mask.c:101-144)
    const float a = u1[p], b = u2[p];
    const float u1x = j < nx - 1 ? u1[p + 1] - a : 0.f, u1y = i < ny - 1 ? u1[p + nx] - a : 0.f;
    const float u2x = j < nx - 1 ? u2[p + 1] - b : 0.f, u2y = i < ny - 1 ? u2[p + nx] - b : 0.f;
    // (the reference's hypot and `1.0 +` are double: :239-242)
    const double g1 = hypot((double)u1x, (double)u1y), g2 = hypot((double)u2x, (double)u2y);
    const float ng1 = (float)(1.0 + (double)(taut * (float)g1)), ng2 = (float)(1.0 + (double)(taut * (float)g2));
    p11[p] = (p11[p] + taut * u1x) / ng1;
    p12[p] = (p12[p] + taut * u1y) / ng1;
    p21[p] = (p21[p] + taut * u2x) / ng2;
    p22[p] = (p22[p] + taut * u2y) / ng2;
}

// iterations run by a warping step = the last n that passed the test
__global__ void k_tvl1_count(const float *err, int *count, float size, float eps2)
{
    int n = 1;
    while (n < TVL1_MAX_ITERATIONS && __fdiv_rn(err[n], size) > eps2) ++n;
    *count = n;
}

} // namespace nlk
