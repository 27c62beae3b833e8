// The pyramid of the TV-L1 flow estimator, host side (Dual_TVL1_optic_flow_multiscale, reference
// lib/tvl1flow/tvl1flow_lib.c:345-477): the sizes of the scales, the two Gaussian kernels, the layout of
// the pyramid in one allocation, and the order of the steps.  Plain C++ without CUDA types: the steps go
// through an executor, so that the library (kernel launches on a stream, nlk_lib.cu) and the CPU model
// of the kernels in tests/models/ (the same device functions compiled for the host) run the very same
// sequence.
#pragma once
#include <math.h>
#include <stddef.h>
#include <vector>

namespace nlk {

constexpr int TVL1_MAX_SCALES = 64;
constexpr int TVL1_GAUSS_MAX = 512;           // taps on one side of the Gaussian
constexpr int TVL1_MINMAX_BLOCKS = 592;       // partial extrema (4 blocks per SM)

// zoom_size (zoom.c:24-37)
inline void tvl1_zoom_size(int nx, int ny, int *nxx, int *nyy, float factor)
{
    *nxx = (int)((float)nx * factor + 0.5);
    *nyy = (int)((float)ny * factor + 0.5);
}

// half kernel of gaussian() (mask.c:226-247), computed on the host with the same libm calls
inline int tvl1_gauss_kernel(double sigma, std::vector<double> &B)
{
    const double den = 2 * sigma * sigma;
    const int size = (int)(5 * sigma) + 1;        // DEFAULT_GAUSSIAN_WINDOW_SIZE, mask.c:24
    if (size < 1 || size > TVL1_GAUSS_MAX) return -1;
    B.resize(size);
    for (int i = 0; i < size; i++) B[i] = 1 / (sigma * sqrt(2.0 * 3.1415926)) * exp(-i * i / den);
    double norm = 0;
    for (int i = 0; i < size; i++) norm += B[i];
    norm *= 2;
    norm -= B[0];
    for (int i = 0; i < size; i++) B[i] /= norm;
    return size;
}

// the cap of the reference's driver (lib/tvl1flow/main.c:159-161): no level much smaller than 16 x 16
inline int tvl1_scales_cap(int nx, int ny, float zfactor, int nscales)
{
    if (nx < 1 || ny < 1 || !(zfactor > 0.f && zfactor < 1.f)) return nscales;
    const float N = 1 + log(hypot((double)nx, (double)ny) / 16.0) / log((double)(1 / zfactor));   // (C's log: double)
    if (N < nscales) nscales = N;
    return nscales;
}

struct Tvl1Pyramid {
    int nscales = 0, fscale = 0, warps = 0;
    int nx[TVL1_MAX_SCALES], ny[TVL1_MAX_SCALES];
    size_t off[TVL1_MAX_SCALES];      // floats: I0, I1 (and u1, u2 below the finest scale) of scale s
    size_t off_tmp, off_part, off_B, floats;
    std::vector<double> Bpre, Bzoom;  // pre-smoothing (sigma 0.8) and zoom-out kernels
    float zfactor = 0.5f, up = 2.f;   // up: (float)1.0 / zfactor, the flow's factor between scales (:434-435)
    const char *error = nullptr;

    size_t size(int s) const { return (size_t)nx[s] * ny[s]; }

    bool plan(int nxx, int nyy, int nscales_, int fscale_, float zfactor_, int warps_)
    {
        if (nxx < 2 || nyy < 2 || nscales_ < 1 || nscales_ > TVL1_MAX_SCALES || fscale_ < 0 || fscale_ > nscales_ ||
            !(zfactor_ > 0.f && zfactor_ < 1.f) || warps_ < 0 || warps_ > 64) {
            error = "bad TV-L1 request";
            return false;
        }
        nscales = nscales_; fscale = fscale_; zfactor = zfactor_; warps = warps_;
        up = (float)1.0 / zfactor;
        nx[0] = nxx; ny[0] = nyy;
        for (int s = 1; s < nscales; ++s) {
            tvl1_zoom_size(nx[s - 1], ny[s - 1], &nx[s], &ny[s], zfactor);
            if (nx[s] < 2 || ny[s] < 2) { error = "TV-L1: too many scales for this image"; return false; }
        }
        const float zsigma = 0.6 * sqrt(1.0 / (zfactor * zfactor) - 1.0);      // ZOOM_SIGMA_ZERO, zoom.c:59
        const int tpre = tvl1_gauss_kernel(0.8, Bpre);                           // PRESMOOTHING_SIGMA, tvl1flow_lib.c:25
        const int tzoom = nscales > 1 ? tvl1_gauss_kernel((double)zsigma, Bzoom) : 1;
        if (tpre < 0 || tzoom < 0 || tpre > nx[0] || tpre > ny[0] ||
            (nscales > 1 && (tzoom > nx[nscales - 2] || tzoom > ny[nscales - 2]))) {
            error = "TV-L1: Gaussian window larger than the image (the reference aborts: mask.c:232)";
            return false;
        }
        size_t fl = 0;
        for (int s = 0; s < nscales; ++s) {
            off[s] = fl;
            fl += (size_t)(s ? 4 : 2) * size(s);
        }
        off_tmp = fl;   fl += 2 * size(0);                     // two work planes
        off_part = fl;  fl += 2 * TVL1_MINMAX_BLOCKS;
        fl = (fl + 1) & ~(size_t)1;
        off_B = fl;     fl += 2 * 2 * TVL1_GAUSS_MAX;          // two kernels of doubles
        floats = fl;
        return true;
    }

    // Ex: zero(p, n); normalize(I0, I1, O0, O1, n, part); gauss(in, tmp, out, w, h, B, taps);
    //     zoom(in, out, w, h, ww, hh, fx, fy, scale, scaled); upload(dB, B);
    //     level(I0, I1, u1, u2, w, h, iterations) -> 0 or an error code
    template <class Ex>
    int run(Ex &ex, float *base, const float *I0, const float *I1, float *u1, float *u2, int *iterations) const
    {
        auto I0s = [&](int s) { return base + off[s]; };
        auto I1s = [&](int s) { return base + off[s] + size(s); };
        auto u1s = [&](int s) { return s ? base + off[s] + 2 * size(s) : u1; };
        auto u2s = [&](int s) { return s ? base + off[s] + 3 * size(s) : u2; };
        float *tmpA = base + off_tmp, *tmpB = tmpA + size(0), *part = base + off_part;
        double *dBpre = reinterpret_cast<double *>(base + off_B), *dBzoom = dBpre + TVL1_GAUSS_MAX;
        ex.upload(dBpre, Bpre);
        if (nscales > 1) ex.upload(dBzoom, Bzoom);
        // normalise to [0, 255] over both images, pre-smooth (:379-386)
        ex.normalize(I0, I1, I0s(0), I1s(0), size(0), part);
        ex.gauss(I0s(0), tmpA, I0s(0), nx[0], ny[0], dBpre, (int)Bpre.size());
        ex.gauss(I1s(0), tmpA, I1s(0), nx[0], ny[0], dBpre, (int)Bpre.size());
        // the scales (:388-404): smooth, resample at (j / zfactor, i / zfactor)
        for (int s = 1; s < nscales; ++s)
            for (int im = 0; im < 2; ++im) {
                ex.gauss(im ? I1s(s - 1) : I0s(s - 1), tmpA, tmpB, nx[s - 1], ny[s - 1], dBzoom, (int)Bzoom.size());
                ex.zoom(tmpB, im ? I1s(s) : I0s(s), nx[s - 1], ny[s - 1], nx[s], ny[s], zfactor, zfactor, 1.f, 0);
            }
        // zero flow at the coarsest scale (:406-408)
        ex.zero(u1s(nscales - 1), size(nscales - 1));
        ex.zero(u2s(nscales - 1), size(nscales - 1));
        if (iterations) for (int k = 0; k < nscales * warps; ++k) iterations[k] = 0;
        for (int s = nscales - 1; s >= 0; --s) {
            if (s >= fscale)     // (:411-420; the scales finer than fscale only upsample: :443-459)
                if (int r = ex.level(I0s(s), I1s(s), u1s(s), u2s(s), nx[s], ny[s],
                                     iterations ? iterations + (size_t)s * warps : nullptr)) return r;
            if (!s) break;
            // zoom_in, times 1 / zfactor (:427-437)
            const float fx = (float)nx[s - 1] / nx[s], fy = (float)ny[s - 1] / ny[s];
            ex.zoom(u1s(s), u1s(s - 1), nx[s], ny[s], nx[s - 1], ny[s - 1], fx, fy, up, 1);
            ex.zoom(u2s(s), u2s(s - 1), nx[s], ny[s], nx[s - 1], ny[s - 1], fx, fy, up, 1);
        }
        return 0;
    }
};

} // namespace nlk
