// CPU model of the TV-L1 kernels (test infrastructure, not product code): the per-pixel device
// functions of bwd_nlkalman_b200/csrc/nlk_tvl1.cuh compiled for the HOST -- the rounding intrinsics
// become plain operators (build with -ffp-contract=off so they round like the intrinsics), a kernel
// becomes a loop over its grid -- and driven through the same Tvl1Pyramid::run sequence as the library.
// tests/test_tvl1_model.py checks it bit for bit against the reference's own library
// (oracle/_ref/libtvl1_ref.so), without a GPU: what is left for the GPU test is the launch plumbing.
//
//   g++ -O2 -ffp-contract=off -fPIC -shared -o libtvl1_model.so tvl1_host_model.cpp
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

#define NLK_HOST_MODEL 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(x)
struct dim3 { unsigned x = 1, y = 1, z = 1; };
static thread_local dim3 blockIdx, blockDim, threadIdx, gridDim;
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }

#include "../../bwd_nlkalman_b200/csrc/nlk_tvl1.cuh"

using namespace nlk;

// run a kernel written for 32 x 8 thread blocks over a w x h image
template <class F>
static void launch2d(int w, int h, F body)
{
    blockDim.x = 32; blockDim.y = 8;
    gridDim.x = (w + 31) / 32; gridDim.y = (h + 7) / 8;
    for (unsigned by = 0; by < gridDim.y; ++by)
        for (unsigned bx = 0; bx < gridDim.x; ++bx)
            for (unsigned ty = 0; ty < 8; ++ty)
                for (unsigned tx = 0; tx < 32; ++tx) {
                    blockIdx.x = bx; blockIdx.y = by; threadIdx.x = tx; threadIdx.y = ty;
                    body();
                }
}

// the level solver as nlk_tvl1_level_dev queues it (nlk_lib.cu): gradient; per warping step the warp
// kernel, then iterations n = 1 .. while tvl1_runs(err, n)
static void model_level(const float *I0, const float *I1, float *u1, float *u2, int nx, int ny, float tau, float lambda,
                        float theta, int warps, float epsilon, int *iterations)
{
    const size_t size = (size_t)nx * ny;
    std::vector<float> buf(10 * size, 0.f), err(TVL1_MAX_ITERATIONS + 4);
    int loop_state = 0;      // (any non-null loop counter: the host model decides about the next iteration itself)
    float *I1x = buf.data(), *I1y = I1x + size, *I1wx = I1y + size, *I1wy = I1wx + size, *grad = I1wy + size,
          *rho_c = grad + size, *p11 = rho_c + size, *p12 = p11 + size, *p21 = p12 + size, *p22 = p21 + size;
    const float l_t = lambda * theta, taut = tau / theta, eps2 = epsilon * epsilon;
    launch2d(nx, ny, [&] { k_tvl1_centered_gradient(I1, I1x, I1y, nx, ny); });
    for (int wi = 0; wi < warps; ++wi) {
        std::fill(err.begin(), err.end(), 0.f);
        launch2d(nx, ny, [&] { k_tvl1_warp(I0, I1, I1x, I1y, u1, u2, I1wx, I1wy, grad, rho_c, nx, ny); });
        int n = 1;
        for (; n <= TVL1_MAX_ITERATIONS; ++n) {
            if (!tvl1_runs(err.data(), n, (float)(nx * ny), eps2)) break;
            float e = 0.f;     // (the reference sums in float too, in another order)
            for (int i = 0; i < ny; ++i)
                for (int j = 0; j < nx; ++j)
                    e += tvl1_u_pixel(rho_c, I1wx, I1wy, grad, p11, p12, p21, p22, u1, u2, i, j, nx, ny, l_t, theta);
            err[n] = e;
            launch2d(nx, ny, [&] { k_tvl1_p(u1, u2, p11, p12, p21, p22, err.data(), 0, nx, ny, taut, eps2, 0ull, &loop_state, &loop_state); });
        }
        if (iterations) iterations[wi] = n - 1;
    }
}

struct HostEx {
    float tau, lambda, theta, epsilon;
    int warps;
    void upload(double *dB, const std::vector<double> &B) { memcpy(dB, B.data(), B.size() * 8); }
    void zero(float *p, size_t n) { memset(p, 0, n * 4); }
    void normalize(const float *I0, const float *I1, float *O0, float *O1, size_t n, float *)
    {
        float mn = I0[0], mx = mn;
        for (size_t i = 0; i < n; ++i) {
            mn = std::min(mn, std::min(I0[i], I1[i]));
            mx = std::max(mx, std::max(I0[i], I1[i]));
        }
        const float den = mx - mn;
        for (size_t i = 0; i < n; ++i) { O0[i] = tvl1_norm_pixel(I0[i], mn, den); O1[i] = tvl1_norm_pixel(I1[i], mn, den); }
    }
    void gauss(const float *in, float *tmp, float *out, int w, int h, const double *B, int taps)
    {
        launch2d(w, h, [&] { k_tvl1_gauss<true>(in, tmp, w, h, B, taps); });
        launch2d(w, h, [&] { k_tvl1_gauss<false>(tmp, out, w, h, B, taps); });
    }
    void zoom(const float *in, float *out, int w, int h, int ww, int hh, float fx, float fy, float scale, int scaled)
    {
        launch2d(ww, hh, [&] { k_tvl1_zoom(in, out, w, h, ww, hh, fx, fy, scale, scaled); });
    }
    int level(const float *I0, const float *I1, float *u1, float *u2, int w, int h, int *iterations)
    {
        model_level(I0, I1, u1, u2, w, h, tau, lambda, theta, warps, epsilon, iterations);
        return 0;
    }
};

extern "C" {

int model_tvl1_scales(int nx, int ny, float zfactor, int nscales) { return tvl1_scales_cap(nx, ny, zfactor, nscales); }

void model_tvl1_level(const float *I0, const float *I1, float *u1, float *u2, int nx, int ny, float tau, float lambda,
                      float theta, int warps, float epsilon, int *iterations)
{
    model_level(I0, I1, u1, u2, nx, ny, tau, lambda, theta, warps, epsilon, iterations);
}

// gaussian() in place (mask.c:216)
int model_tvl1_gaussian(float *I, int nx, int ny, double sigma)
{
    std::vector<double> B;
    const int taps = tvl1_gauss_kernel(sigma, B);
    if (taps < 0 || taps > nx || taps > ny) return 1;
    std::vector<float> tmp((size_t)nx * ny);
    HostEx ex{};
    ex.gauss(I, tmp.data(), I, nx, ny, B.data(), taps);
    return 0;
}

// zoom_in (zoom.c:91) / the resampling half of zoom_out
void model_tvl1_zoom(const float *in, float *out, int nx, int ny, int nxx, int nyy, float fx, float fy, float scale,
                     int scaled)
{
    HostEx ex{};
    ex.zoom(in, out, nx, ny, nxx, nyy, fx, fy, scale, scaled);
}

// flow: two planes, u then v (like nlk_tvl1_flow_host)
int model_tvl1_flow(const float *I0, const float *I1, float *flow, int nx, int ny, float tau, float lambda, float theta,
                    int nscales, int fscale, float zfactor, int warps, float epsilon, int *iterations)
{
    Tvl1Pyramid P;
    if (!P.plan(nx, ny, nscales, fscale, zfactor, warps)) return 1;
    std::vector<float> base(P.floats);
    HostEx ex{tau, lambda, theta, epsilon, warps};
    return P.run(ex, base.data(), I0, I1, flow, flow + (size_t)nx * ny, iterations);
}

}
