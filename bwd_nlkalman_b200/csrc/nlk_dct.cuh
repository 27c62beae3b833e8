// Orthonormal DCT-II / DCT-III of psz x psz tiles on CUDA cores (replaces the
// reference's per-thread FFTW plans, src/nlkalman.c:138-360, whose net effect for a
// depth-1 batch is the textbook orthonormal 2-D DCT, see :281-299 and :335-353).
//
// 1-D transforms use the even/odd split of the DCT matrix T (N even):
//   X[2m]   = sum_j T[2m][j]   (x[j] + x[N-1-j]),   X[2m+1] = sum_j T[2m+1][j] (x[j] - x[N-1-j])
// with T read from constant memory at compile-time offsets, so every product is one
// FFMA with a constant-bank operand.  A tile is owned by one thread: for 8x8 the whole
// tile lives in registers; other sizes run the row and column passes through the
// thread's own shared-memory tile.
#pragma once
#include "nlk_common.cuh"

namespace nlk {

template <int N>
__device__ __forceinline__ void dct1d_fwd(float (&x)[N])
{
    static_assert(N % 2 == 0, "even sizes only");
    constexpr int H = N / 2;
    float s[H], d[H];
#pragma unroll
    for (int j = 0; j < H; ++j) { s[j] = x[j] + x[N - 1 - j]; d[j] = x[j] - x[N - 1 - j]; }
#pragma unroll
    for (int m = 0; m < H; ++m) {
        float e = 0.f, o = 0.f;
#pragma unroll
        for (int j = 0; j < H; ++j) {
            e = fmaf(c_dct[N][(2 * m) * N + j], s[j], e);
            o = fmaf(c_dct[N][(2 * m + 1) * N + j], d[j], o);
        }
        x[2 * m] = e;
        x[2 * m + 1] = o;
    }
}

template <int N>
__device__ __forceinline__ void dct1d_inv(float (&X)[N])
{
    static_assert(N % 2 == 0, "even sizes only");
    constexpr int H = N / 2;
    float e[H], o[H];
#pragma unroll
    for (int j = 0; j < H; ++j) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int m = 0; m < H; ++m) {
            a = fmaf(c_dct[N][(2 * m) * N + j], X[2 * m], a);
            b = fmaf(c_dct[N][(2 * m + 1) * N + j], X[2 * m + 1], b);
        }
        e[j] = a;
        o[j] = b;
    }
#pragma unroll
    for (int j = 0; j < H; ++j) { X[j] = e[j] + o[j]; X[N - 1 - j] = e[j] - o[j]; }
}

// 8 points: the even half splits once more (T[2m][j] is symmetric in j <-> 3-j for
// m even, antisymmetric for m odd): 36 operations instead of 40
template <>
__device__ __forceinline__ void dct1d_fwd<8>(float (&x)[8])
{
    const float (&T)[MAX_PSZ * MAX_PSZ] = c_dct[8];
    const float s0 = x[0] + x[7], s1 = x[1] + x[6], s2 = x[2] + x[5], s3 = x[3] + x[4];
    const float d0 = x[0] - x[7], d1 = x[1] - x[6], d2 = x[2] - x[5], d3 = x[3] - x[4];
    const float ss0 = s0 + s3, ss1 = s1 + s2, sd0 = s0 - s3, sd1 = s1 - s2;
    x[0] = fmaf(T[0 * 8 + 1], ss1, T[0 * 8 + 0] * ss0);
    x[4] = fmaf(T[4 * 8 + 1], ss1, T[4 * 8 + 0] * ss0);
    x[2] = fmaf(T[2 * 8 + 1], sd1, T[2 * 8 + 0] * sd0);
    x[6] = fmaf(T[6 * 8 + 1], sd1, T[6 * 8 + 0] * sd0);
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const int k = 2 * m + 1;
        x[k] = fmaf(T[k * 8 + 3], d3, fmaf(T[k * 8 + 2], d2, fmaf(T[k * 8 + 1], d1, T[k * 8 + 0] * d0)));
    }
}

template <>
__device__ __forceinline__ void dct1d_inv<8>(float (&X)[8])
{
    const float (&T)[MAX_PSZ * MAX_PSZ] = c_dct[8];
    const float p0 = fmaf(T[4 * 8 + 0], X[4], T[0 * 8 + 0] * X[0]);
    const float p1 = fmaf(T[4 * 8 + 1], X[4], T[0 * 8 + 1] * X[0]);
    const float q0 = fmaf(T[6 * 8 + 0], X[6], T[2 * 8 + 0] * X[2]);
    const float q1 = fmaf(T[6 * 8 + 1], X[6], T[2 * 8 + 1] * X[2]);
    const float e0 = p0 + q0, e3 = p0 - q0, e1 = p1 + q1, e2 = p1 - q1;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
        o[j] = fmaf(T[7 * 8 + j], X[7], fmaf(T[5 * 8 + j], X[5], fmaf(T[3 * 8 + j], X[3], T[1 * 8 + j] * X[1])));
    X[0] = e0 + o[0]; X[7] = e0 - o[0];
    X[1] = e1 + o[1]; X[6] = e1 - o[1];
    X[2] = e2 + o[2]; X[5] = e2 - o[2];
    X[3] = e3 + o[3]; X[4] = e3 - o[3];
}

// whole 8x8 tile in registers; source element (y, x) at src[y*row_stride + x*col_stride]
template <bool INVERSE>
__device__ __forceinline__ void dct2d_8x8_strided(const float *src, int row_stride,
                                                  int col_stride, float *tile)
{
    float t[64];
#pragma unroll
    for (int y = 0; y < 8; ++y)
#pragma unroll
        for (int x = 0; x < 8; ++x) t[y * 8 + x] = src[y * row_stride + x * col_stride];
#pragma unroll
    for (int y = 0; y < 8; ++y) {
        float r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = t[y * 8 + i];
        if (INVERSE) dct1d_inv<8>(r); else dct1d_fwd<8>(r);
#pragma unroll
        for (int i = 0; i < 8; ++i) t[y * 8 + i] = r[i];
    }
#pragma unroll
    for (int x = 0; x < 8; ++x) {
        float r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = t[i * 8 + x];
        if (INVERSE) dct1d_inv<8>(r); else dct1d_fwd<8>(r);
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i * 8 + x] = r[i];
    }
#pragma unroll
    for (int i = 0; i < 64; ++i) tile[i] = t[i];
}

// row pass then column pass through the thread's own shared-memory tile
template <int N, bool INVERSE>
__device__ __forceinline__ void dct2d_passes_smem(float *__restrict__ tile)
{
    for (int y = 0; y < N; ++y) {
        float r[N];
#pragma unroll
        for (int i = 0; i < N; ++i) r[i] = tile[y * N + i];
        if (INVERSE) dct1d_inv<N>(r); else dct1d_fwd<N>(r);
#pragma unroll
        for (int i = 0; i < N; ++i) tile[y * N + i] = r[i];
    }
    for (int x = 0; x < N; ++x) {
        float r[N];
#pragma unroll
        for (int i = 0; i < N; ++i) r[i] = tile[i * N + x];
        if (INVERSE) dct1d_inv<N>(r); else dct1d_fwd<N>(r);
#pragma unroll
        for (int i = 0; i < N; ++i) tile[i * N + x] = r[i];
    }
}

// any side up to MAX_PSZ (also odd ones): plain matrix form, run-time size
template <bool INVERSE>
__device__ inline void dct2d_generic_smem(float *__restrict__ tile, int n)
{
    float r[MAX_PSZ];
    const float *T = c_dct[n];
    for (int pass = 0; pass < 2; ++pass) {
        const int line_stride = pass == 0 ? n : 1, elem_stride = pass == 0 ? 1 : n;
        for (int l = 0; l < n; ++l) {
            float *p = tile + l * line_stride;
            for (int i = 0; i < n; ++i) r[i] = p[i * elem_stride];
            for (int k = 0; k < n; ++k) {
                float acc = 0.f;
                for (int j = 0; j < n; ++j)
                    acc = fmaf(INVERSE ? T[j * n + k] : T[k * n + j], r[j], acc);
                p[k * elem_stride] = acc;
            }
        }
    }
}

template <bool INVERSE>
__device__ __forceinline__ void dct2d_8x8_smem(float *__restrict__ tile)
{
    dct2d_8x8_strided<INVERSE>(tile, 8, 1, tile);
}

template <int PSZ_T, bool INVERSE>
__device__ __forceinline__ void dct2d_tile(float *__restrict__ tile, int psz_rt)
{
    if constexpr (PSZ_T == 8) dct2d_8x8_smem<INVERSE>(tile);
    else if constexpr (PSZ_T != 0) dct2d_passes_smem<PSZ_T, INVERSE>(tile);
    else dct2d_generic_smem<INVERSE>(tile, psz_rt);
}

// forward transform of the patch whose element (y, x) is src[y*row_stride + x*col_stride]
// (a window staged in shared memory) into the thread's own tile
template <int PSZ_T>
__device__ __forceinline__ void dct2d_from_window(const float *__restrict__ src, int row_stride,
                                                  int col_stride, float *__restrict__ tile, int psz_rt)
{
    if constexpr (PSZ_T == 8) {
        dct2d_8x8_strided<false>(src, row_stride, col_stride, tile);
    } else {
        const int psz = PSZ_T ? PSZ_T : psz_rt;
        for (int y = 0; y < psz; ++y)
            for (int x = 0; x < psz; ++x) tile[y * psz + x] = src[y * row_stride + x * col_stride];
        dct2d_tile<PSZ_T, false>(tile, psz_rt);
    }
}

} // namespace nlk
