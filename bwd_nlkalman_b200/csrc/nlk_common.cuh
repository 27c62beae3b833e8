// Shared definitions for the sm_100a kernels of the NL-Kalman per-frame step.
//
// One pass (filter or smoother, reference src/nlkalman.c:518-951 / :1409-1865) is
//   search_knn  -> mask_resolve -> group_filter -> normalize
// over a regular grid of reference patches p = (gx*step, gy*step), step = psz/2.
// Images are fp32, interleaved HWC, as at the reference boundary.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace nlk {

constexpr int MAX_PSZ = 16;   // patch side supported by the kernels
constexpr int MAX_CH = 4;     // channels supported by the kernels
constexpr int MAX_K = 4096;   // bound on candidates kept per group (table c_inv)

// candidate record: qx in bits 0..14, qy in bits 16..30, bit 31 = the previous-frame
// patch at q is valid (and so is the one at p)
__host__ __device__ inline uint32_t cand_pack(int qx, int qy, int prev)
{
    return (uint32_t)qx | ((uint32_t)qy << 16) | ((uint32_t)(prev != 0) << 31);
}
__host__ __device__ inline int cand_x(uint32_t c) { return (int)(c & 0x7fffu); }
__host__ __device__ inline int cand_y(uint32_t c) { return (int)((c >> 16) & 0x7fffu); }
__host__ __device__ inline int cand_prev(uint32_t c) { return (int)(c >> 31); }

// per grid patch header written by search_knn
struct GroupHdr {
    int nk;     // candidates kept (0: no search, k <= 1)
    int np0;    // kept candidates with a valid previous patch
    int flags;  // bit 0: prev_p, bit 1: marks (sets the processed mask for its members)
    int pxy;    // the patch position, packed like a candidate record (x | y << 16)
};
constexpr int HDR_PREV_P = 1;
constexpr int HDR_MARKS = 2;

struct PassParams {
    int w, h, ch, psz, step;
    int gw, gh, G;          // grid of reference patches
    int gy0, gy1;           // grid rows searched and filtered by this pass (a strip; whole frame: 0, gh)
    int smooth;             // 0: filter pass, 1: smoother pass
    int r_x, r_t;           // search radii (spatial / temporal)
    int k_x, k_t, tagg;     // patches kept (spatial / temporal), group size
    float sigma2, beta_x, beta_t;
    int has_prev, has_bsic;
    const float *src;       // search / statistics source: bsic1 if given else in1
    const float *in1;       // noisy frame (filter) or filtered frame (smoother)
    const float *prev0;     // warped previous estimate (NaN = invalid) or nullptr
    const uint8_t *valid;   // [vh][vw] patch validity of prev0, or nullptr
    int vw, vh;             // vw = w - psz + 1, vh = h - psz + 1
    // search output
    uint32_t *cand;         // [G][kstride]
    int kstride;
    GroupHdr *hdr;          // [G]
    uint32_t *nbr;          // [G][nbw] grid-aligned group members, bit (dy+R)*(2R+1)+(dx+R)
    int nbw, R;
    float *dbg_dist;        // [G][kstride] distances of the kept candidates, or nullptr
    int *any_nbr;           // set to 1 if some group marks a grid patch other than its own
    // grid patches of the other search radius (no valid previous patch in a temporal pass), queued by
    // the main search launch for k_search_patch_list
    int *xlist, *xcount;
    // resolve output
    uint8_t *actflag;       // [G] 1 = processed (written by mask_resolve)
    int *active;            // [G] indices of processed patches, raster order
    int *nactive;
    int *work;              // ticket counter of group_filter (zeroed at the start of a pass)
    // aggregation
    float *accw;            // [h*w][ch+1]: weighted sums, then the weight
    float *out;             // [h*w][ch]
    float *dbg_vp;          // [G] posterior variance of processed groups, or nullptr
};

// The library is built as ONE translation unit (nlk_lib.cu includes every kernel
// header), so the constant tables are plain definitions here.
// orthonormal DCT-II matrices T[k*n+j] = c(k) sqrt(2/n) cos(pi (j+1/2) k / n), one per
// supported side n (filled by the host at library initialisation)
__constant__ float c_dct[MAX_PSZ + 1][MAX_PSZ * MAX_PSZ];
// the 8-point matrix again for the packed-fp32 transform (nlk_dct.cuh): c_dct8u[k*8+j] =
// (T[k][j], T[k][j]) and c_dct8v[j*4+y] = (T[2j][y], T[2j+1][y])
__constant__ float2 c_dct8u[64];
__constant__ float2 c_dct8v[16];
// Gaussian aggregation windows (reference src/nlkalman.c:401-416), one per side
__constant__ float c_win[MAX_PSZ + 1][MAX_PSZ * MAX_PSZ];
// c_inv[n] = (float)(1. / (float)n), the reference's Welford factors (src/nlkalman.c:755-756)
__constant__ float c_inv[MAX_K + 1];

} // namespace nlk
