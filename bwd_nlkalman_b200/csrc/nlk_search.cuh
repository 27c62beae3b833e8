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
#include "nlk_dct.cuh"   // the packed-fp32 helpers

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

// Bitonic sort of 128 keys held four per lane (element r*32 + lane in register r):
// strides below 32 exchange through shuffles, the two larger ones between registers.
__device__ __forceinline__ void cmpswap(unsigned long long &a, unsigned long long &b, bool asc)
{
    const bool sw = (a > b) == asc;
    const unsigned long long lo = sw ? b : a, hi = sw ? a : b;
    a = lo;
    b = hi;
}

__device__ __forceinline__ void warp_sort128(unsigned long long (&key)[4], int lane)
{
#pragma unroll
    for (int size = 2; size <= 128; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= 32) {
                const int rs = stride >> 5;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    if ((r & rs) == 0) {
                        const bool asc = (((r << 5) & size) == 0);
                        cmpswap(key[r], key[r + rs], asc);
                    }
                }
            } else {
                const bool lower = (lane & stride) == 0;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const unsigned long long other = __shfl_xor_sync(0xffffffffu, key[r], stride);
                    const bool asc = ((((r << 5) | lane) & size) == 0);
                    const bool take_min = (lower == asc);
                    const bool less = key[r] < other;
                    key[r] = (less == take_min) ? key[r] : other;
                }
            }
        }
    }
}

// The 32 smallest of 128 keys (four per lane, element r*32 + lane in register r), ascending in
// key[0]: sort the four 32-key rows (0 and 2 ascending, 1 and 3 descending), then twice
// "element-wise minimum of an ascending and a descending row" (a bitonic row holding the 32
// smallest of the 64) followed by its 5-stage merge.  75 compare-exchange stages on rows
// instead of the 100 (+ 12 between registers) of the full sort; keys are unique, so the
// result is the same list the full sort starts with.
__device__ __forceinline__ void warp_row_stage(unsigned long long &k, int stride, bool asc_row, int lane)
{
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, k, stride);
    const bool lower = (lane & stride) == 0;
    const bool take_min = (lower == asc_row);
    const bool less = k < other;
    k = (less == take_min) ? k : other;
}
__device__ __forceinline__ void warp_top32_of_128(unsigned long long (&key)[4], int lane)
{
    // full bitonic sort of each row; direction of the final merge: rows 0, 2 up, rows 1, 3 down
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const bool up = (size == 32) ? ((r & 1) == 0) : ((lane & size) == 0);
                warp_row_stage(key[r], stride, up, lane);
            }
        }
    }
    key[0] = key[0] < key[1] ? key[0] : key[1];
    key[2] = key[2] < key[3] ? key[2] : key[3];
#pragma unroll
    for (int stride = 16; stride > 0; stride >>= 1) {
        warp_row_stage(key[0], stride, true, lane);
        warp_row_stage(key[2], stride, false, lane);
    }
    key[0] = key[0] < key[2] ? key[0] : key[2];
#pragma unroll
    for (int stride = 16; stride > 0; stride >>= 1) warp_row_stage(key[0], stride, true, lane);
}

// warp-level: sort npad keys (ascending), keep the first k, write the candidate
// records, the header and the grid-neighbour bitmap of patch g.  bm: 2*nbw words of
// shared scratch private to the warp.
template <bool TOP32 = false, bool PRESORTED = false>
__device__ __forceinline__ void sort_and_emit(const PassParams &P, int g, int px, int py, int prev_p, int k,
                                              unsigned long long *keys, int n, int npad, int nx, int x0,
                                              int y0, int lane, unsigned int *bm)
{
    const int nbw = P.nbw;
    uint32_t *nbr_out = P.nbr + (long)g * nbw;
    if (PRESORTED) {
        // the caller sorted `keys` (ascending) already
    } else if (npad <= 128) {
        // the usual temporal window (121 candidates): sort in registers
        unsigned long long key[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) key[r] = (r * 32 + lane < npad) ? keys[r * 32 + lane] : ~0ull;
        if (TOP32) {                          // the launch keeps at most 32 candidates per patch
            warp_top32_of_128(key, lane);     // only the first k are read below
            keys[lane] = key[0];
        } else {
            warp_sort128(key, lane);
#pragma unroll
            for (int r = 0; r < 4; ++r) if (r * 32 + lane < npad) keys[r * 32 + lane] = key[r];
        }
        __syncwarp();
    } else
    // bitonic sort in shared memory, ascending
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
        hd.pxy = (int)cand_pack(px, py, 0);
        P.hdr[g] = hd;
        if (others) *P.any_nbr = 1;
    }
}

template <int PSZ_T, int CH_T>
__global__ void __launch_bounds__(256)
k_search(const PassParams P, int warps_per_cta, int wrow, int win_floats, int npad_max,
         int warp_smem_bytes, int only_r)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = P.gy0 * P.gw + blockIdx.x * warps_per_cta + warp;
    if (g >= P.gy1 * P.gw) return;

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
    const int r = P.smooth ? P.r_t : (prev_p ? P.r_t : P.r_x);
    if (only_r >= 0 && r != only_r) return; // another launch handles this radius

    if (k <= 1) {
        // no search (reference :631 / :1522).  Filter: nothing is aggregated for this
        // patch.  Smoother: a plain copy (no valid previous patch; with one, k_t <= 1 is
        // rejected by pass_setup: the reference's branch :1699-1730 reads uninitialised memory).
        if (lane == 0) {
            GroupHdr hd;
            hd.nk = 0;
            hd.np0 = 0;
            hd.flags = prev_p ? HDR_PREV_P : 0;
            hd.pxy = (int)cand_pack(px, py, 0);
            P.hdr[g] = hd;
        }
        for (int i = lane; i < nbw; i += 32) nbr_out[i] = 0u;
        return;
    }

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

    sort_and_emit(P, g, px, py, prev_p, k, keys, n, npad, nx, x0, y0, lane, bm);
}

// ---- fast path: one thread per candidate ROW ------------------------------------------------
// A block takes a run of consecutive grid patches of one grid row and stages their common
// super-window once.  Thread (patch, candidate row) walks the window row by row and keeps
// the 2*RAD+1 distances of its candidate row in registers: every staged value is loaded
// once and used by up to PSZ candidates, and each distance still receives its terms in
// the reference's (hy, hx, c) order with separately rounded multiply and add.
template <int PSZ, int CH, int RAD, bool TOP32>
__device__ __forceinline__ void search_rows_block(const PassParams &P, int gy, int run, int np_cta, int wrow,
                                                  int wh_max, int npad, bool feed_worklist)
{
    constexpr int NX = 2 * RAD + 1, NR = 2 * RAD + 1, RL = PSZ * CH;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem_raw);    // [np_cta][npad]
    float *win = reinterpret_cast<float *>(keys + (size_t)np_cta * npad);           // [wh_max][wrow]
    unsigned int *bm = reinterpret_cast<unsigned int *>(win + (size_t)wh_max * wrow); // [8][2*nbw]
    int *s_prev = reinterpret_cast<int *>(bm + 8 * 2 * P.nbw);                       // [np_cta]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gx0 = run * np_cta;
    const int np = min(np_cta, P.gw - gx0);
    const int py = gy * P.step;

    // which patches of the run this launch handles: those whose search radius is RAD
    for (int s = tid; s < np; s += blockDim.x) {
        const int px = (gx0 + s) * P.step;
        const int prev_p = (P.valid != nullptr) ? (int)P.valid[(long)py * P.vw + px] : 0;
        const int r = P.smooth ? P.r_t : (prev_p ? P.r_t : P.r_x);
        s_prev[s] = (r == RAD) ? prev_p : -1;
        if (r != RAD && feed_worklist) {
            // a patch of the other radius (no valid previous patch: occlusions, the warp's border rows and
            // columns): queued for k_search_patch_list, one block per patch
            P.xlist[atomicAdd(P.xcount, 1)] = gy * P.gw + gx0 + s;
        }
    }
    __syncthreads();
    int any = 0;
    for (int s = 0; s < np; ++s) any |= (s_prev[s] >= 0) && ((s_prev[s] ? P.k_t : P.k_x) > 1);
    // patches without search (k <= 1) only need their header
    for (int s = tid; s < np; s += blockDim.x) {
        const int prev_p = s_prev[s];
        if (prev_p < 0 || (prev_p ? P.k_t : P.k_x) > 1) continue;
        const int g = gy * P.gw + gx0 + s;
        GroupHdr hd;
        hd.nk = 0;
        hd.np0 = 0;
        hd.flags = prev_p ? HDR_PREV_P : 0;
        hd.pxy = (int)cand_pack((gx0 + s) * P.step, py, 0);
        P.hdr[g] = hd;
        for (int i = 0; i < P.nbw; ++i) P.nbr[(long)g * P.nbw + i] = 0u;
    }
    if (!any) return;

    // super-window: rows y0 .. y1+PSZ-1, columns x0w .. x1w+PSZ-1
    const int y0 = max(py - RAD, 0), y1 = min(py + RAD, P.h - PSZ);
    const int ny = y1 - y0 + 1, wh = ny + PSZ - 1;
    const int x0w = max(gx0 * P.step - RAD, 0);
    const int x1w = min((gx0 + np - 1) * P.step + RAD, P.w - PSZ);
    const int wlen = (x1w - x0w + PSZ) * CH;
    for (int row = warp; row < wh; row += 8) {
        const float *srow = P.src + ((long)(y0 + row) * P.w + x0w) * CH;
        float *drow = win + row * wrow;
        for (int j = lane; j < wlen; j += 32) drow[j] = srow[j];
    }
    __syncthreads();

    // distances: thread = (patch slot, candidate row)
    {
        const int slot = tid / NR, ry = tid - slot * NR;
        const int prev_p = slot < np ? s_prev[slot] : -1;
        if (prev_p >= 0 && (prev_p ? P.k_t : P.k_x) > 1 && ry < ny) {
            const int px = (gx0 + slot) * P.step;
            const int x0 = max(px - RAD, 0), x1 = min(px + RAD, P.w - PSZ);
            const int nx = x1 - x0 + 1;
            const float *refp = win + (py - y0) * wrow + (px - x0w) * CH;
            const float *canp = win + ry * wrow + (x0 - x0w) * CH;
            float acc[NX];
#pragma unroll
            for (int j = 0; j < NX; ++j) acc[j] = 0.f;
            // Per window row: the RL = PSZ*CH reference values as RL/2 register pairs, the candidate
            // row walked pair by pair.  Candidate j meets reference term t at row element j*CH + t,
            // so two consecutive terms are one packed subtract and one packed multiply
            // (sub.rn.f32x2 / mul.rn.f32x2: each half rounded like the scalar instruction) against
            // the even-aligned pair (v[2m], v[2m+1]) when j*CH is even and the odd-aligned pair
            // (v[2m+1], v[2m+2]) when it is odd; the two products are then added one after the other
            // with scalar adds -- the reference's (hy, hx, c) order and roundings, two issue slots per
            // term instead of three.
            static_assert(RL % 2 == 0, "an even number of terms per patch row");
            constexpr int NV = (NX + PSZ - 1) * CH;
#pragma unroll 1
            for (int hy = 0; hy < PSZ; ++hy) {
                f32x2 ref2[RL / 2];
#pragma unroll
                for (int i = 0; i < RL / 2; ++i) ref2[i] = pk2(refp[hy * wrow + 2 * i], refp[hy * wrow + 2 * i + 1]);
                const float *cr = canp + hy * wrow;
#pragma unroll
                for (int m = 0; 2 * m < NV; ++m) {
                    const float va = cr[2 * m];
                    const float vb = (2 * m + 1 < NV) ? cr[2 * m + 1] : 0.f;
                    const float vc = (2 * m + 2 < NV) ? cr[2 * m + 2] : 0.f;
                    const f32x2 pe = pk2(va, vb), po = pk2(vb, vc);
#pragma unroll
                    for (int j = 0; j < NX; ++j) {
                        const bool odd = ((j * CH) & 1) != 0;
                        const int t = (odd ? 2 * m + 1 : 2 * m) - j * CH;     // first term of the pair
                        if (t >= 0 && t + 1 < RL) {
                            const f32x2 e = sub2(odd ? po : pe, ref2[t / 2]);
                            float lo, hi;
                            upk2(mul2(e, e), lo, hi);
                            acc[j] = __fadd_rn(__fadd_rn(acc[j], lo), hi);
                        }
                    }
                }
            }
            const float npix = (float)PSZ * (float)PSZ * (float)CH;
            unsigned long long *kp = keys + (size_t)slot * npad + ry * nx;
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                if (j < nx) {
                    const float d = fmaxf(__fdiv_rn(acc[j], npix), 0.f);
                    kp[j] = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned int)(ry * nx + j);
                }
            }
        }
    }
    __syncthreads();

    // selection: one warp per patch
    for (int slot = warp; slot < np; slot += 8) {
        const int prev_p = s_prev[slot];
        if (prev_p < 0) continue;
        const int k = prev_p ? P.k_t : P.k_x;
        if (k <= 1) continue;
        const int px = (gx0 + slot) * P.step;
        const int x0 = max(px - RAD, 0), x1 = min(px + RAD, P.w - PSZ);
        const int nx = x1 - x0 + 1, n = nx * ny;
        unsigned long long *kp = keys + (size_t)slot * npad;
        int np2 = 32;
        while (np2 < n) np2 <<= 1;
        for (int i = n + lane; i < np2; i += 32) kp[i] = ~0ull;
        __syncwarp();
        sort_and_emit<TOP32>(P, gy * P.gw + gx0 + slot, px, py, prev_p, k, kp, n, np2, nx, x0, y0, lane,
                             bm + warp * 2 * P.nbw);
    }
}

// whole grid rows [gy0, gy1): block = (grid row, run)
template <int PSZ, int CH, int RAD, bool TOP32>
__global__ void __launch_bounds__(256)
k_search_rows(const PassParams P, int np_cta, int runs_per_row, int wrow, int wh_max, int npad, int feed)
{
    const int gyl = blockIdx.x / runs_per_row, run = blockIdx.x - gyl * runs_per_row;
    search_rows_block<PSZ, CH, RAD, TOP32>(P, P.gy0 + gyl, run, np_cta, wrow, wh_max, npad, feed != 0);
}

// The patches queued by the launch of the other radius (few and scattered: occluded patches, the
// warp's border rows and columns): ONE BLOCK PER PATCH, thread = candidate, so that a handful of
// patches costs a handful of microseconds instead of the latency of whole runs.  Each distance is
// still the reference's sequential (hy, hx, c) sum with separately rounded operations; the keys are
// sorted by the whole block (bitonic network in shared memory), then one warp emits.
template <int PSZ_T, int CH_T>
__global__ void __launch_bounds__(256)
k_search_patch_list(const PassParams P, int r, int npad, int wrow)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem_raw);       // [npad]
    float *win = reinterpret_cast<float *>(keys + npad);                                 // [2r + psz][wrow]
    const int psz = PSZ_T ? PSZ_T : P.psz, ch = CH_T ? CH_T : P.ch;
    unsigned int *bm = reinterpret_cast<unsigned int *>(win + (size_t)(2 * r + psz) * wrow);   // [2][nbw]
    const int tid = threadIdx.x, lane = tid & 31;
    const int nq = *P.xcount;
    for (int wi = blockIdx.x; wi < nq; wi += gridDim.x) {
        const int g = P.xlist[wi];
        const int gy = g / P.gw, gx = g - gy * P.gw;
        const int px = gx * P.step, py = gy * P.step;
        const int prev_p = (P.valid != nullptr) ? (int)P.valid[(long)py * P.vw + px] : 0;
        const int k = prev_p ? P.k_t : P.k_x;
        if (k <= 1) {          // no search (reference :631): header only
            if (tid == 0) {
                GroupHdr hd;
                hd.nk = 0; hd.np0 = 0; hd.flags = prev_p ? HDR_PREV_P : 0; hd.pxy = (int)cand_pack(px, py, 0);
                P.hdr[g] = hd;
            }
            for (int i = tid; i < P.nbw; i += blockDim.x) P.nbr[(long)g * P.nbw + i] = 0u;
            continue;
        }
        const int x0 = max(px - r, 0), x1 = min(px + r, P.w - psz);
        const int y0 = max(py - r, 0), y1 = min(py + r, P.h - psz);
        const int nx = x1 - x0 + 1, ny = y1 - y0 + 1, n = nx * ny;
        const int wlen = (nx + psz - 1) * ch, wh = ny + psz - 1;
        for (int row = tid >> 5; row < wh; row += 8) {
            const float *srow = P.src + ((long)(y0 + row) * P.w + x0) * ch;
            for (int j = lane; j < wlen; j += 32) win[row * wrow + j] = srow[j];
        }
        __syncthreads();
        const float *cp = win + (py - y0) * wrow + (px - x0) * ch;
        const float npix = (float)psz * (float)psz * (float)ch;
        int np2 = 32;
        while (np2 < n) np2 <<= 1;
        for (int ci = tid; ci < np2; ci += blockDim.x) {
            unsigned long long key = ~0ull;
            if (ci < n) {
                const int cy = ci / nx, cx = ci - cy * nx;
                const float ww = patch_dist<PSZ_T, CH_T>(win + cy * wrow + cx * ch, cp, wrow, psz, ch);
                const float d = fmaxf(__fdiv_rn(ww, npix), 0.f);
                key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned int)ci;
            }
            keys[ci] = key;
        }
        __syncthreads();
        for (int size = 2; size <= np2; size <<= 1)
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int t = tid; t < (np2 >> 1); t += blockDim.x) {
                    const int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1)), hi = lo + stride;
                    const bool asc = (lo & size) == 0;
                    const unsigned long long a = keys[lo], b = keys[hi];
                    if ((a > b) == asc) { keys[lo] = b; keys[hi] = a; }
                }
                __syncthreads();
            }
        if (tid < 32) sort_and_emit<false, true>(P, g, px, py, prev_p, k, keys, n, np2, nx, x0, y0, lane, bm);
        __syncthreads();   // shared memory is reused by the next patch
    }
}

template <int RAD>
inline void search_rows_geom(const PassParams &P, int PSZ, int CH, int *np_cta, int *npad, int *wrow, int *wh_max,
                             size_t *smem)
{
    constexpr int NR = 2 * RAD + 1;
    *np_cta = 256 / NR;
    *npad = 32;
    while (*npad < NR * NR) *npad <<= 1;
    *wrow = (((*np_cta - 1) * P.step + 2 * RAD + PSZ) * CH) | 1; // odd stride: candidate rows fall in different banks
    *wh_max = 2 * RAD + PSZ;
    *smem = (size_t)*np_cta * *npad * 8 + (size_t)*wh_max * *wrow * 4 + 8 * 2 * P.nbw * 4 + *np_cta * 4;
}

// mode 0: all runs of the strip; 1: all runs, queueing the patches of the other radius for k_search_patch_list
template <int PSZ, int CH, int RAD>
inline int launch_search_rows(const PassParams &P, int mode, cudaStream_t st)
{
    int np_cta, npad, wrow, wh_max;
    size_t smem;
    search_rows_geom<RAD>(P, PSZ, CH, &np_cta, &npad, &wrow, &wh_max, &smem);
    if (smem > 220 * 1024) return -1;
    const int runs = (P.gw + np_cta - 1) / np_cta;
    {
        // what the patches of this launch keep: those with a valid previous patch k_t, the others k_x
        // (reference :631-707); with one radius for both, or in the smoother, a launch holds both kinds
        int kmax = P.k_t > P.k_x ? P.k_t : P.k_x;
        if (!P.smooth && P.r_t != P.r_x && P.has_prev) kmax = (RAD == P.r_t) ? P.k_t : P.k_x;
        if (!P.has_prev && !P.smooth) kmax = P.k_x;
        if (kmax <= 32 && npad <= 128) {
            cudaFuncSetAttribute(k_search_rows<PSZ, CH, RAD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k_search_rows<PSZ, CH, RAD, true><<<runs * (P.gy1 - P.gy0), 256, smem, st>>>(P, np_cta, runs, wrow, wh_max, npad, mode);
        } else {
            cudaFuncSetAttribute(k_search_rows<PSZ, CH, RAD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k_search_rows<PSZ, CH, RAD, false><<<runs * (P.gy1 - P.gy0), 256, smem, st>>>(P, np_cta, runs, wrow, wh_max, npad, mode);
        }
    }
    return 1;
}

// runs per grid row and patches per run of the fast kernel for radius r, 0 if it has none
inline int search_rows_runs(const PassParams &P, int r, int *np_cta)
{
#define NLK_ROWS_G(PS, CHN, RD) if (P.psz == PS && P.ch == CHN && r == RD) { int a, b, c; size_t d; search_rows_geom<RD>(P, PS, CHN, np_cta, &a, &b, &c, &d); return d > 220 * 1024 ? 0 : (P.gw + *np_cta - 1) / *np_cta; }
    NLK_ROWS_G(8, 3, 5) NLK_ROWS_G(8, 3, 10) NLK_ROWS_G(8, 1, 5) NLK_ROWS_G(8, 1, 10) NLK_ROWS_G(12, 3, 10) NLK_ROWS_G(12, 3, 15)
#undef NLK_ROWS_G
    return 0;
}

// the generic kernel, restricted to the patches whose search radius is r (-1: all)
inline int launch_search_generic(const PassParams &P, int r, int only_r, cudaStream_t st)
{
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
    const int nb = ((P.gy1 - P.gy0) * P.gw + warps - 1) / warps;
    const int nt = warps * 32;
#define NLK_LAUNCH_SEARCH(PS, CHN)                                                                \
    do {                                                                                          \
        cudaFuncSetAttribute(k_search<PS, CHN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
        k_search<PS, CHN><<<nb, nt, smem, st>>>(P, warps, wrow, win_floats, npad, warp_bytes, only_r); \
    } while (0)
    if (P.psz == 8 && P.ch == 3) NLK_LAUNCH_SEARCH(8, 3);
    else if (P.psz == 8 && P.ch == 1) NLK_LAUNCH_SEARCH(8, 1);
    else if (P.psz == 12 && P.ch == 3) NLK_LAUNCH_SEARCH(12, 3);
    else if (P.psz == 12 && P.ch == 1) NLK_LAUNCH_SEARCH(12, 1);
    else NLK_LAUNCH_SEARCH(0, 0);
#undef NLK_LAUNCH_SEARCH
    return 1;
}

inline int launch_search_radius(const PassParams &P, int r, int mode, cudaStream_t st)
{
#define NLK_ROWS(PS, CHN, RD) if (P.psz == PS && P.ch == CHN && r == RD) return launch_search_rows<PS, CHN, RD>(P, mode, st)
    NLK_ROWS(8, 3, 5);
    NLK_ROWS(8, 3, 10);
    NLK_ROWS(8, 1, 5);
    NLK_ROWS(8, 1, 10);
    NLK_ROWS(12, 3, 10);
    NLK_ROWS(12, 3, 15);
#undef NLK_ROWS
    return launch_search_generic(P, r, r, st);
}

// the queued patches of radius r (see k_search_patch_list)
inline int launch_search_patch_list(const PassParams &P, int r, cudaStream_t st)
{
    const int side = 2 * r + 1;
    int npad = 32;
    while (npad < side * side) npad <<= 1;
    const int wrow = ((2 * r + P.psz) * P.ch) | 1;        // odd stride: candidate rows fall in different banks
    const size_t smem = (size_t)npad * 8 + (size_t)(2 * r + P.psz) * wrow * 4 + 2 * P.nbw * 4;
    if (smem > 200 * 1024) return -1;
    const int nb = 148 * 8;      // ~14 KB of shared memory a block: a thousand queued patches in one wave
#define NLK_LAUNCH_PL(PS, CHN)                                                                        \
    do {                                                                                              \
        cudaFuncSetAttribute(k_search_patch_list<PS, CHN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        k_search_patch_list<PS, CHN><<<nb, 256, smem, st>>>(P, r, npad, wrow);                        \
    } while (0)
    if (P.psz == 8 && P.ch == 3) NLK_LAUNCH_PL(8, 3);
    else if (P.psz == 8 && P.ch == 1) NLK_LAUNCH_PL(8, 1);
    else if (P.psz == 12 && P.ch == 3) NLK_LAUNCH_PL(12, 3);
    else NLK_LAUNCH_PL(0, 0);
#undef NLK_LAUNCH_PL
    return 1;
}

inline int launch_search(PassParams &P, cudaStream_t st)
{
    // a patch searches with radius r_t (it has a valid previous patch, or the pass is the
    // smoother) or r_x (reference :637, :1527)
    if (P.smooth || P.r_x == P.r_t) return launch_search_radius(P, P.r_t, 0, st);
    if (!P.has_prev) return launch_search_radius(P, P.r_x, 0, st);   // no previous frame: all spatial
    // temporal pass: the r_t launch covers the frame and queues the patches without a valid previous
    // patch (occlusions, warp borders) for the one-block-per-patch launch of radius r_x
    int np1 = 0;
    const bool listed = search_rows_runs(P, P.r_t, &np1) > 0 && P.xlist != nullptr;
    int n = launch_search_radius(P, P.r_t, listed ? 1 : 0, st);
    if (n < 0) return n;
    const int m = listed ? launch_search_patch_list(P, P.r_x, st) : launch_search_radius(P, P.r_x, 0, st);
    return m < 0 ? m : n + m;
}

} // namespace nlk
