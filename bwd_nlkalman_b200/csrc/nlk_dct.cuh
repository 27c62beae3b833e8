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

// ---- packed fp32 (Blackwell FFMA2 / FADD2 / FMUL2, PTX *.f32x2) -------------------------------
// One instruction works on a 64-bit register pair: two IEEE fp32 operations for one issue
// slot (tools/micro/ffma2_rate.cu: same flop rate as FFMA, half the instructions).  The 8x8
// transform below keeps a tile as 32 pairs and needs no register shuffling between its two
// passes:
//   in   P[y][x] = (t[y][x], t[7-y][x])                 y < 4   (a row and its mirror row)
//   rows: the 36-operation 8-point transform on the four row pairs (both halves alike)
//   columns: with (lo, hi) = (X[y][k], X[7-y][k]) the even coefficients of column k come from
//        s_y = lo + hi and the odd ones from d_y = lo - hi, so the pair (s_y, d_y) times the
//        constant pair (T[2j][y], T[2j+1][y]) accumulates (coef[2j][k], coef[2j+1][k]):
//   out  P[j][k] = (coef[2j][k], coef[2j+1][k])         j < 4
// and the inverse runs the same steps backwards.  336 issue slots per 2-D transform
// against 576 scalar ones.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(f32x2 r, float &a, float &b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

#define NLK_TU(i) (*reinterpret_cast<const f32x2 *>(&c_dct8u[i]))
#define NLK_TV(i) (*reinterpret_cast<const f32x2 *>(&c_dct8v[i]))

__device__ __forceinline__ void dct8_fwd_x2(f32x2 (&x)[8])
{
    const f32x2 s0 = add2(x[0], x[7]), s1 = add2(x[1], x[6]), s2 = add2(x[2], x[5]), s3 = add2(x[3], x[4]);
    const f32x2 d0 = sub2(x[0], x[7]), d1 = sub2(x[1], x[6]), d2 = sub2(x[2], x[5]), d3 = sub2(x[3], x[4]);
    const f32x2 ss0 = add2(s0, s3), ss1 = add2(s1, s2), sd0 = sub2(s0, s3), sd1 = sub2(s1, s2);
    x[0] = fma2(NLK_TU(0 * 8 + 1), ss1, mul2(NLK_TU(0 * 8 + 0), ss0));
    x[4] = fma2(NLK_TU(4 * 8 + 1), ss1, mul2(NLK_TU(4 * 8 + 0), ss0));
    x[2] = fma2(NLK_TU(2 * 8 + 1), sd1, mul2(NLK_TU(2 * 8 + 0), sd0));
    x[6] = fma2(NLK_TU(6 * 8 + 1), sd1, mul2(NLK_TU(6 * 8 + 0), sd0));
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const int k = 2 * m + 1;
        x[k] = fma2(NLK_TU(k * 8 + 3), d3,
                    fma2(NLK_TU(k * 8 + 2), d2, fma2(NLK_TU(k * 8 + 1), d1, mul2(NLK_TU(k * 8 + 0), d0))));
    }
}

__device__ __forceinline__ void dct8_inv_x2(f32x2 (&X)[8])
{
    const f32x2 p0 = fma2(NLK_TU(4 * 8 + 0), X[4], mul2(NLK_TU(0 * 8 + 0), X[0]));
    const f32x2 p1 = fma2(NLK_TU(4 * 8 + 1), X[4], mul2(NLK_TU(0 * 8 + 1), X[0]));
    const f32x2 q0 = fma2(NLK_TU(6 * 8 + 0), X[6], mul2(NLK_TU(2 * 8 + 0), X[2]));
    const f32x2 q1 = fma2(NLK_TU(6 * 8 + 1), X[6], mul2(NLK_TU(2 * 8 + 1), X[2]));
    const f32x2 e0 = add2(p0, q0), e3 = sub2(p0, q0), e1 = add2(p1, q1), e2 = sub2(p1, q1);
    f32x2 o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
        o[j] = fma2(NLK_TU(7 * 8 + j), X[7],
                    fma2(NLK_TU(5 * 8 + j), X[5], fma2(NLK_TU(3 * 8 + j), X[3], mul2(NLK_TU(1 * 8 + j), X[1]))));
    X[0] = add2(e0, o[0]); X[7] = sub2(e0, o[0]);
    X[1] = add2(e1, o[1]); X[6] = sub2(e1, o[1]);
    X[2] = add2(e2, o[2]); X[5] = sub2(e2, o[2]);
    X[3] = add2(e3, o[3]); X[4] = sub2(e3, o[3]);
}

// P[y][x] = (t[y][x], t[7-y][x])  ->  P[j][k] = (coef[2j][k], coef[2j+1][k])
__device__ __forceinline__ void dct8x8_fwd_x2(f32x2 (&P)[4][8])
{
#pragma unroll
    for (int y = 0; y < 4; ++y) dct8_fwd_x2(P[y]);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        f32x2 sd[4];
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            float lo, hi;
            upk2(P[y][k], lo, hi);
            sd[y] = pk2(lo + hi, lo - hi);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            P[j][k] = fma2(NLK_TV(j * 4 + 3), sd[3],
                           fma2(NLK_TV(j * 4 + 2), sd[2], fma2(NLK_TV(j * 4 + 1), sd[1], mul2(NLK_TV(j * 4 + 0), sd[0]))));
    }
}

// P[j][k] = (coef[2j][k], coef[2j+1][k])  ->  P[y][x] = (t[y][x], t[7-y][x])
__device__ __forceinline__ void dct8x8_inv_x2(f32x2 (&P)[4][8])
{
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        f32x2 eo[4];
#pragma unroll
        for (int y = 0; y < 4; ++y)
            eo[y] = fma2(NLK_TV(3 * 4 + y), P[3][k],
                         fma2(NLK_TV(2 * 4 + y), P[2][k], fma2(NLK_TV(1 * 4 + y), P[1][k], mul2(NLK_TV(0 * 4 + y), P[0][k]))));
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            float e, o;
            upk2(eo[y], e, o);
            P[y][k] = pk2(e + o, e - o);
        }
    }
#pragma unroll
    for (int y = 0; y < 4; ++y) dct8_inv_x2(P[y]);
}

// forward transform, per-coefficient gain, inverse transform in one go:
//   t <- T^-1( a * T(t) + b ),  a and b as records {a[2j][k], a[2j+1][k], b[2j][k], b[2j+1][k]} at g4[j*8+k]
// Rows forward; then column by column: forward, gain, inverse (nothing but the column's
// four pairs is live in between); rows inverse.  SCALE_ONLY: b = 0.
template <bool SCALE_ONLY>
__device__ __forceinline__ void dct8x8_shrink_x2(f32x2 (&P)[4][8], const float4 *__restrict__ g4)
{
#pragma unroll
    for (int y = 0; y < 4; ++y) dct8_fwd_x2(P[y]);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        f32x2 sd[4], cf[4];
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            float lo, hi;
            upk2(P[y][k], lo, hi);
            sd[y] = pk2(lo + hi, lo - hi);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 gn = g4[j * 8 + k];
            const f32x2 c = fma2(NLK_TV(j * 4 + 3), sd[3],
                                 fma2(NLK_TV(j * 4 + 2), sd[2], fma2(NLK_TV(j * 4 + 1), sd[1], mul2(NLK_TV(j * 4 + 0), sd[0]))));
            cf[j] = SCALE_ONLY ? mul2(pk2(gn.x, gn.y), c) : fma2(pk2(gn.x, gn.y), c, pk2(gn.z, gn.w));
        }
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            const f32x2 eo = fma2(NLK_TV(3 * 4 + y), cf[3],
                                  fma2(NLK_TV(2 * 4 + y), cf[2], fma2(NLK_TV(1 * 4 + y), cf[1], mul2(NLK_TV(0 * 4 + y), cf[0]))));
            float e, o;
            upk2(eo, e, o);
            P[y][k] = pk2(e + o, e - o);
        }
    }
#pragma unroll
    for (int y = 0; y < 4; ++y) dct8_inv_x2(P[y]);
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
