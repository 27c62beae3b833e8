// search_knn: block matching + k-NN selection for every grid patch (reference
// src/nlkalman.c:630-707 and :1521-1597), independent of the processed mask.
//
// One warp per reference patch.  The clamped search window of the source image is
// staged in shared memory; each lane computes whole candidate distances as the
// reference's sequential fp32 sum over (hy, hx, c) of separately rounded squares
// (no FMA), so distances -- and therefore the k-NN order -- carry the reference's bits.
// Selection is a warp bitonic sort of 64-bit keys (distance bits << 32 | scan index):
// unique keys make it the stable ascending order the reference gets from glibc's
// merge sort (src/nlkalman.c:706).
#pragma once
#include "nlk_common.cuh"

namespace nlk {

constexpr int SEARCH_MAX_NPAD = 4096;

template <int PSZ_T, int CH_T>
__device__ __forceinline__ float patch_dist(const float *__restrict__ cq, const float *__restrict__ cp,
                                            int wrow, int psz_rt, int ch_rt)
{
    const int psz = PSZ_T ? PSZ_T : psz_rt;
    const int rowlen = (PSZ_T && CH_T) ? PSZ_T * CH_T : psz * (CH_T ? CH_T : ch_rt);
    float ww = 0.f;
    if (PSZ_T && CH_T) {
#pragma unroll
        for (int hy = 0; hy < PSZ_T; ++hy) {
#pragma unroll
            for (int j = 0; j < PSZ_T * CH_T; ++j) {
                const float e = __fsub_rn(cq[hy * wrow + j], cp[hy * wrow + j]);
                ww = __fadd_rn(ww, __fmul_rn(e, e));
            }
        }
    } else {
        for (int hy = 0; hy < psz; ++hy)
            for (int j = 0; j < rowlen; ++j) {
                const float e = __fsub_rn(cq[hy * wrow + j], cp[hy * wrow + j]);
                ww = __fadd_rn(ww, __fmul_rn(e, e));
            }
    }
    return ww;
}

template <int PSZ_T, int CH_T>
__global__ void __launch_bounds__(256)
k_search(const PassParams P, int warps_per_cta, int wrow, int win_floats, int npad_max,
         int warp_smem_bytes)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x * warps_per_cta + warp;
    if (g >= P.G) return;

    const int psz = PSZ_T ? PSZ_T : P.psz;
    const int ch = CH_T ? CH_T : P.ch;
    unsigned char *base = smem_raw + (size_t)warp * warp_smem_bytes;
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(base);
    float *win = reinterpret_cast<float *>(base + (size_t)npad_max * 8);
    unsigned int *bm = reinterpret_cast<unsigned int *>(win + win_floats); // [2][nbw]

    const int px = (g % P.gw) * P.step, py = (g / P.gw) * P.step;
    const int prev_p = (P.valid != nullptr) ? (int)P.valid[(long)py * P.vw + px] : 0;
    int k = prev_p ? P.k_t : P.k_x;
    const int nbw = P.nbw;
    uint32_t *nbr_out = P.nbr + (long)g * nbw;

    if (k <= 1) {
        // no search (reference :631 / :1522).  Filter: nothing is aggregated for this
        // patch.  Smoother: the patch at p alone (prev_p) or a plain copy (!prev_p).
        if (lane == 0) {
            GroupHdr hd;
            hd.nk = 0;
            hd.np0 = (P.smooth && prev_p) ? 1 : 0;
            hd.flags = (prev_p ? HDR_PREV_P : 0) | ((P.smooth && prev_p) ? HDR_MARKS : 0);
            hd.pad = 0;
            P.hdr[g] = hd;
        }
        for (int i = lane; i < nbw; i += 32) nbr_out[i] = 0u;
        return;
    }

    const int r = P.smooth ? P.r_t : (prev_p ? P.r_t : P.r_x);
    const int x0 = max(px - r, 0), x1 = min(px + r, P.w - psz);
    const int y0 = max(py - r, 0), y1 = min(py + r, P.h - psz);
    const int nx = x1 - x0 + 1, ny = y1 - y0 + 1, n = nx * ny;

    // stage the window: rows y0 .. y1+psz-1, columns x0 .. x1+psz-1 (all channels)
    const int wlen = (nx + psz - 1) * ch, wh = ny + psz - 1;
    for (int row = 0; row < wh; ++row) {
        const float *srow = P.src + ((long)(y0 + row) * P.w + x0) * ch;
        for (int j = lane; j < wlen; j += 32) win[row * wrow + j] = srow[j];
    }
    int npad = 32;
    while (npad < n) npad <<= 1;
    __syncwarp();

    const float *cp = win + (py - y0) * wrow + (px - x0) * ch;
    const float npix = (float)psz * (float)psz * (float)ch;
    for (int ci = lane; ci < npad; ci += 32) {
        unsigned long long key = ~0ull;
        if (ci < n) {
            const int cy = ci / nx, cx = ci - cy * nx;
            const float *cq = win + cy * wrow + cx * ch;
            const float ww = patch_dist<PSZ_T, CH_T>(cq, cp, wrow, psz, ch);
            const float d = fmaxf(__fdiv_rn(ww, npix), 0.f);
            key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned int)ci;
        }
        keys[ci] = key;
    }
    __syncwarp();

    // bitonic sort, ascending
    for (int size = 2; size <= npad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = lane; t < (npad >> 1); t += 32) {
                const int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
                const int hi = lo + stride;
                const bool asc = ((lo & size) == 0);
                const unsigned long long a = keys[lo], b = keys[hi];
                if ((a > b) == asc) { keys[lo] = b; keys[hi] = a; }
            }
            __syncwarp();
        }
    }

    // keep the first k; group membership for the processed-mask resolution
    k = min(k, n);
    for (int i = lane; i < 2 * nbw; i += 32) bm[i] = 0u;
    __syncwarp();
    const int R = P.R, side = 2 * R + 1;
    int np0 = 0;
    uint32_t *cand_out = P.cand + (long)g * P.kstride;
    for (int b0 = 0; b0 < k; b0 += 32) {
        const int i = b0 + lane;
        int prev = 0, qx = 0, qy = 0;
        if (i < k) {
            const unsigned long long key = keys[i];
            const int ci = (int)(key & 0xffffffffu);
            const int cy = ci / nx;
            qy = y0 + cy;
            qx = x0 + (ci - cy * nx);
            prev = prev_p && P.valid[(long)qy * P.vw + qx];
            cand_out[i] = cand_pack(qx, qy, prev);
            if (P.dbg_dist) P.dbg_dist[(long)g * P.kstride + i] = __uint_as_float((unsigned int)(key >> 32));
        }
        const unsigned int bal = __ballot_sync(0xffffffffu, prev);
        if (i < k) {
            const int dxs = qx - px, dys = qy - py;
            if (dxs % P.step == 0 && dys % P.step == 0) {
                const int bit = (dys / P.step + R) * side + (dxs / P.step + R);
                // (A) first tagg candidates with a valid previous patch
                const int rank = np0 + __popc(bal & ((1u << lane) - 1u));
                if (prev && rank < P.tagg) atomicOr(&bm[bit >> 5], 1u << (bit & 31));
                // (B) no valid previous patch in the group: the first tagg candidates
                if (!P.smooth && i < P.tagg) atomicOr(&bm[nbw + (bit >> 5)], 1u << (bit & 31));
            }
        }
        np0 += __popc(bal);
    }
    __syncwarp();
    const int use_b = (np0 == 0);
    int marks;
    if (P.smooth) marks = np0 > 0;                 // reference :1844
    else marks = !(P.has_prev && np0 == 0);        // reference :931
    const int selfbit = R * side + R;
    unsigned int others = 0u;
    for (int i = lane; i < nbw; i += 32) {
        unsigned int v = (P.smooth && use_b) ? 0u : bm[use_b * nbw + i];
        if (!marks) v = 0u;
        nbr_out[i] = v;
        if (i == (selfbit >> 5)) v &= ~(1u << (selfbit & 31));
        others |= v;
    }
    others = __reduce_or_sync(0xffffffffu, others);
    if (lane == 0) {
        GroupHdr hd;
        hd.nk = k;
        hd.np0 = np0;
        hd.flags = (prev_p ? HDR_PREV_P : 0) | (marks ? HDR_MARKS : 0);
        hd.pad = 0;
        P.hdr[g] = hd;
        if (others) *P.any_nbr = 1;
    }
}

inline int search_max_radius(const PassParams &P) { return P.smooth ? P.r_t : max(P.r_t, P.r_x); }

inline int launch_search(const PassParams &P, cudaStream_t st)
{
    const int r = search_max_radius(P);
    const int side = 2 * r + 1;
    int npad = 32;
    while (npad < side * side) npad <<= 1;
    const int wrow = (2 * r + P.psz) * P.ch + 1;       // +1: odd stride spreads rows over banks
    const int win_floats = (2 * r + P.psz) * wrow;
    int warp_bytes = npad * 8 + win_floats * 4 + 2 * P.nbw * 4;
    warp_bytes = (warp_bytes + 15) & ~15;
    int warps = 200 * 1024 / warp_bytes;
    if (warps > 8) warps = 8;
    if (warps < 1) return -1;
    // several CTAs per SM hide the staging latency: keep each CTA below ~48 KB when possible
    while (warps > 2 && warps * warp_bytes > 56 * 1024) warps >>= 1;
    const int smem = warps * warp_bytes;
    const int nb = (P.G + warps - 1) / warps;
    const int nt = warps * 32;
#define NLK_LAUNCH_SEARCH(PS, CHN)                                                                \
    do {                                                                                          \
        cudaFuncSetAttribute(k_search<PS, CHN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
        k_search<PS, CHN><<<nb, nt, smem, st>>>(P, warps, wrow, win_floats, npad, warp_bytes);    \
    } while (0)
    if (P.psz == 8 && P.ch == 3) NLK_LAUNCH_SEARCH(8, 3);
    else if (P.psz == 8 && P.ch == 1) NLK_LAUNCH_SEARCH(8, 1);
    else if (P.psz == 12 && P.ch == 3) NLK_LAUNCH_SEARCH(12, 3);
    else if (P.psz == 12 && P.ch == 1) NLK_LAUNCH_SEARCH(12, 1);
    else NLK_LAUNCH_SEARCH(0, 0);
#undef NLK_LAUNCH_SEARCH
    return 1;
}

} // namespace nlk
