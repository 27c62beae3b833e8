// HBM-bound per-pixel kernels: opponent colour transform, bicubic warp with NaN
// occlusion, patch-validity map, final normalisation.
#pragma once
#include "nlk_common.cuh"
#include <math.h>

namespace nlk {

// ---- colour transform (reference src/nlkalman.c:92-130) ---------------------------------
// The three coefficients are computed on the host exactly as the reference does and
// passed in, and the sums are evaluated left to right without FMA contraction.
__global__ void k_rgb2opp(float *__restrict__ dst, const float *__restrict__ src, long npix,
                          float a, float b, float c)
{
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < npix;
         i += (long)gridDim.x * blockDim.x) {
        const float r = src[3 * i], g = src[3 * i + 1], bl = src[3 * i + 2];
        const float Y = __fmul_rn(a, __fadd_rn(__fadd_rn(r, g), bl));
        const float U = __fmul_rn(b, __fsub_rn(r, bl));
        const float V = __fmul_rn(c, __fadd_rn(__fsub_rn(__fmul_rn(0.25f, r), __fmul_rn(0.5f, g)),
                                               __fmul_rn(0.25f, bl)));
        dst[3 * i] = Y; dst[3 * i + 1] = U; dst[3 * i + 2] = V;
    }
}

__global__ void k_opp2rgb(float *__restrict__ dst, const float *__restrict__ src, long npix,
                          float a, float b, float c)
{
    const float hc = __fmul_rn(0.5f, c);
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < npix;
         i += (long)gridDim.x * blockDim.x) {
        const float Y = src[3 * i], U = src[3 * i + 1], V = src[3 * i + 2];
        const float aY = __fmul_rn(a, Y), bU = __fmul_rn(b, U);
        const float R = __fadd_rn(__fadd_rn(aY, bU), __fmul_rn(hc, V));
        const float G = __fsub_rn(aY, __fmul_rn(c, V));
        const float B = __fadd_rn(__fsub_rn(aY, bU), __fmul_rn(hc, V));
        dst[3 * i] = R; dst[3 * i + 1] = G; dst[3 * i + 2] = B;
    }
}

inline int launch_rgb2opp_copy(float *dst, const float *src, long npix, int inverse, cudaStream_t st)
{
    const int nt = 256;
    const int nb = (int)((npix + nt - 1) / nt < 148 * 16 ? (npix + nt - 1) / nt : 148 * 16);
    if (!inverse) {
        const float a = 1.f / sqrtf(3.f), b = 1.f / sqrtf(2.f), c = 2.f * a * sqrtf(2.f);
        k_rgb2opp<<<nb, nt, 0, st>>>(dst, src, npix, a, b, c);
    } else {
        const float a = 1.f / sqrtf(3.f), b = 1.f / sqrtf(2.f), c = a / b;
        k_opp2rgb<<<nb, nt, 0, st>>>(dst, src, npix, a, b, c);
    }
    return 1;
}

inline int launch_rgb2opp(float *im, long npix, int inverse, cudaStream_t st)
{
    return launch_rgb2opp_copy(im, im, npix, inverse, st);
}

// ---- bicubic warp (reference src/nlkalman.c:29-88) -----------------------------------------
// Keys cubic in the reference's Horner form; its literals are double, so the polynomial
// is evaluated in double and rounded once, as the reference's compiled code does.
__device__ __forceinline__ float cubic1(float v0, float v1, float v2, float v3, float x)
{
    const double xd = x;
    const double d0 = v0, d1 = v1, d2 = v2, d3 = v3;
    return (float)(d1 + 0.5 * xd * ((double)(v2 - v0)
                 + xd * (2.0 * d0 - 5.0 * d1 + 4.0 * d2 - d3
                 + xd * (3.0 * (double)(v1 - v2) + d3 - d0))));
}

template <int CH>
__global__ void k_warp(float *__restrict__ imw, const float *__restrict__ im,
                       const float *__restrict__ of, const float *__restrict__ msk,
                       int w, int h, int ch_rt, int row0, int row1)
{
    const int ch = CH ? CH : ch_rt;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = row0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= row1) return;
    const long pix = (long)y * w + x;
    float *o = imw + pix * ch;
    const float nanv = __int_as_float(0x7fc00000);
    if (msk != nullptr && msk[pix] != 0.f) {
        for (int c = 0; c < ch; ++c) o[c] = nanv;
        return;
    }
    float xw = __fadd_rn((float)x, of[pix * 2 + 0]);
    float yw = __fadd_rn((float)y, of[pix * 2 + 1]);
    xw = __fsub_rn(xw, 1.f);
    yw = __fsub_rn(yw, 1.f);
    const int ix = (int)floorf(xw), iy = (int)floorf(yw);
    const float fx = __fsub_rn(xw, (float)ix), fy = __fsub_rn(yw, (float)iy);
    // a non-finite flow gives NaN as well (any tap outside the image does)
    const bool inside = (ix >= 0) && (ix + 3 < w) && (iy >= 0) && (iy + 3 < h) &&
                        (xw == xw) && (yw == yw);
    if (!inside) {
        for (int c = 0; c < ch; ++c) o[c] = nanv;
        return;
    }
    for (int c = 0; c < ch; ++c) {
        float col[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float *p = im + ((long)iy * w + ix + i) * ch + c;
            const long rs = (long)w * ch;
            col[i] = cubic1(p[0], p[rs], p[2 * rs], p[3 * rs], fy);
        }
        o[c] = cubic1(col[0], col[1], col[2], col[3], fx);
    }
}

// rows [row0, row1) of the warped image (the whole frame: 0, h)
inline int launch_warp(float *imw, const float *im, const float *of, const float *msk,
                       int w, int h, int ch, int row0, int row1, cudaStream_t st)
{
    if (row1 <= row0) return 0;
    dim3 nt(32, 8), nb((w + 31) / 32, (row1 - row0 + 7) / 8);
    if (ch == 3) k_warp<3><<<nb, nt, 0, st>>>(imw, im, of, msk, w, h, ch, row0, row1);
    else if (ch == 1) k_warp<1><<<nb, nt, 0, st>>>(imw, im, of, msk, w, h, ch, row0, row1);
    else k_warp<0><<<nb, nt, 0, st>>>(imw, im, of, msk, w, h, ch, row0, row1);
    return 1;
}

// ---- occlusion mask from the divergence of the flow ----------------------------------------
// What the pipeline script computes with plambda between the flow and the filter (reference
// scripts/nlkalman-seq.sh:70-72, :95-97):
//     "x(0,0)[0] x(-1,0)[0] - x(0,0)[1] x(0,-1)[1] - + fabs TH > 255 *"
// i.e. 255 where |(u(x,y) - u(x-1,y)) + (v(x,y) - v(x,y-1))| > TH, else 0, samples outside the
// image replaced by the nearest one (plambda's default boundary, reference
// lib/imscript-lite/src/getpixel.c:18-29), float arithmetic, left-to-right.
__global__ void k_occlusion(float *__restrict__ occ, const float *__restrict__ of, int w, int h, float th)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const long pix = (long)y * w + x;
    const float u = of[pix * 2], v = of[pix * 2 + 1];
    const float ul = of[((long)y * w + max(x - 1, 0)) * 2];
    const float vu = of[((long)max(y - 1, 0) * w + x) * 2 + 1];
    const float d = fabsf(__fadd_rn(__fsub_rn(u, ul), __fsub_rn(v, vu)));
    occ[pix] = d > th ? 255.f : 0.f;
}

inline int launch_occlusion(float *occ, const float *of, int w, int h, float th, cudaStream_t st)
{
    dim3 nt(32, 8), nb((w + 31) / 32, (h + 7) / 8);
    k_occlusion<<<nb, nt, 0, st>>>(occ, of, w, h, th);
    return 1;
}

// 8-bit mask samples (as stored in the PNG the script writes) to the float form warp_bicubic takes
__global__ void k_mask_u8(float *__restrict__ occ, const uint8_t *__restrict__ m8, long n)
{
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        occ[i] = (float)m8[i];
}

inline int launch_mask_u8(float *occ, const uint8_t *m8, long n, cudaStream_t st)
{
    const int nt = 256;
    const int nb = (int)((n + nt - 1) / nt < 148 * 8 ? (n + nt - 1) / nt : 148 * 8);
    k_mask_u8<<<nb, nt, 0, st>>>(occ, m8, n);
    return 1;
}

// ---- patch validity of the warped previous frame --------------------------------------------
// valid(q) <=> no NaN in channel 0 of the psz x psz patch at q (reference
// src/nlkalman.c:605-609, :725-730).  Separable: row pass then column pass.
__global__ void k_valid_rows(uint8_t *__restrict__ tmp, const float *__restrict__ prev0,
                             int w, int h, int ch, int psz, int vw, int row0)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = row0 + blockIdx.y;
    if (x >= vw) return;
    const float *p = prev0 + ((long)y * w + x) * ch;
    int bad = 0;
    for (int i = 0; i < psz; ++i) {
        const float v = p[(long)i * ch];
        bad |= (v != v);
    }
    tmp[(long)y * vw + x] = (uint8_t)bad;
}

__global__ void k_valid_cols(uint8_t *__restrict__ valid, const uint8_t *__restrict__ tmp,
                             int vw, int vh, int psz, int row0)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = row0 + blockIdx.y;
    if (x >= vw || y >= vh) return;
    int bad = 0;
    for (int i = 0; i < psz; ++i) bad |= tmp[(long)(y + i) * vw + x];
    valid[(long)y * vw + x] = (uint8_t)(!bad);
}

// validity of the patches whose top row is in [q0, q1) (the whole frame: 0, h - psz + 1)
inline int launch_valid_map(uint8_t *valid, uint8_t *tmp, const float *prev0, int w, int h, int ch,
                            int psz, int q0, int q1, cudaStream_t st)
{
    const int vw = w - psz + 1, vh = h - psz + 1;
    if (q0 < 0) q0 = 0;
    if (q1 > vh) q1 = vh;
    if (vw <= 0 || q1 <= q0) return 0;
    const int nt = 256;
    k_valid_rows<<<dim3((vw + nt - 1) / nt, q1 - q0 + psz - 1), nt, 0, st>>>(tmp, prev0, w, h, ch, psz, vw, q0);
    k_valid_cols<<<dim3((vw + nt - 1) / nt, q1 - q0), nt, 0, st>>>(valid, tmp, vw, vh, psz, q0);
    return 2;
}

// ---- normalisation (reference src/nlkalman.c:939-942, :1854-1856) --------------------------
template <int CH>
__global__ void k_normalize(float *__restrict__ out, const float *__restrict__ accw,
                            const float *__restrict__ in1, long npix, int ch_rt)
{
    const int ch = CH ? CH : ch_rt;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < npix;
         i += (long)gridDim.x * blockDim.x) {
        const float *a = accw + i * (ch + 1);
        const float wsum = a[ch];
        if (wsum > 1e-6f) {
            for (int c = 0; c < ch; ++c) out[i * ch + c] = __fdiv_rn(a[c], wsum);
        } else {
            for (int c = 0; c < ch; ++c) out[i * ch + c] = in1[i * ch + c];
        }
    }
}

// pixel rows [row0, row1) (the whole frame: 0, h)
inline int launch_normalize(const PassParams &P, int row0, int row1, cudaStream_t st)
{
    if (row1 <= row0) return 0;
    const long npix = (long)P.w * (row1 - row0), off = (long)P.w * row0;
    const int nt = 256;
    const int nb = (int)((npix + nt - 1) / nt < 148 * 16 ? (npix + nt - 1) / nt : 148 * 16);
    float *out = P.out + off * P.ch;
    const float *acc = P.accw + off * (P.ch + 1), *in1 = P.in1 + off * P.ch;
    if (P.ch == 3) k_normalize<3><<<nb, nt, 0, st>>>(out, acc, in1, npix, P.ch);
    else if (P.ch == 1) k_normalize<1><<<nb, nt, 0, st>>>(out, acc, in1, npix, P.ch);
    else k_normalize<0><<<nb, nt, 0, st>>>(out, acc, in1, npix, P.ch);
    return 1;
}

} // namespace nlk
