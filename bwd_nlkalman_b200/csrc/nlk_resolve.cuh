// mask_resolve: the reference visits grid patches in raster order and skips a patch
// when an earlier group already aggregated a patch at exactly that grid position
// (reference src/nlkalman.c:597-600 with :930-931, and :1490-1493 with :1844).  Group
// membership does not depend on the mask, so search_knn computes every group first
// and this kernel replays the greedy raster-order rule on the grid:
//
//   for p in raster order:  if (!mask[p]) { active += p;  mask[q] = 1 for q in nbr[p] }
//
// A group reaches at most R = floor(r/step) grid cells in each direction, so all cells
// with equal t = j + (R+1) i are independent: a skewed wavefront, run by ONE thread
// block (thread = grid row) with the mask as a bit set in shared memory.
#pragma once
#include "nlk_common.cuh"

namespace nlk {

template <bool SMEM_MASK>
__global__ void __launch_bounds__(1024) k_resolve(const PassParams P)
{
    extern __shared__ unsigned int sbits[];
    __shared__ int s_count;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int G = P.G, gw = P.gw, gh = P.gh;

    if (*P.any_nbr == 0) {
        // no group marks another grid patch: every patch is processed
        for (int g = tid; g < G; g += nthr) P.active[g] = g;
        if (tid == 0) *P.nactive = G;
        return;
    }
    if (SMEM_MASK) {
        for (int i = tid; i < (G + 31) / 32; i += nthr) sbits[i] = 0u;
    } else {
        for (int i = tid; i < G; i += nthr) P.gmask[i] = 0;
    }
    if (tid == 0) s_count = 0;
    __syncthreads();

    const int R = P.R, side = 2 * R + 1, skew = R + 1, nbw = P.nbw;
    const int nsteps = gw + skew * (gh - 1);
    // word 0 of the next cell of each of this thread's rows is fetched one step ahead
    // (consecutive cells of a row share cache lines, so this is an L1 hit most steps)
    constexpr int MAX_ROWS = 4;
    unsigned int nxt[MAX_ROWS];
#pragma unroll
    for (int m = 0; m < MAX_ROWS; ++m) {
        const int i = tid + m * nthr;
        nxt[m] = (i == 0 && i < gh) ? P.nbr[0] : 0u;
    }
    for (int t = 0; t < nsteps; ++t) {
#pragma unroll
        for (int m = 0; m < MAX_ROWS; ++m) {
            const int i = tid + m * nthr;
            if (i >= gh) break;
            const int j = t - skew * i;
            if (j < -1 || j >= gw) continue;
            unsigned int bits0 = nxt[m];
            if (j + 1 < gw) nxt[m] = P.nbr[(long)(i * gw + j + 1) * nbw];
            if (j < 0) continue;
            const int g = i * gw + j;
            bool done;
            if (SMEM_MASK) done = (sbits[g >> 5] >> (g & 31)) & 1u;
            else done = P.gmask[g] != 0;
            if (done) continue;
            P.active[atomicAdd(&s_count, 1)] = g;
            for (int wd = 0; wd < nbw; ++wd) {
                unsigned int bits = wd == 0 ? bits0 : P.nbr[(long)g * nbw + wd];
                while (bits) {
                    const int b = __ffs(bits) - 1;
                    bits &= bits - 1;
                    const int bit = wd * 32 + b;
                    const int dy = bit / side - R, dx = bit % side - R;
                    const int g2 = (i + dy) * gw + (j + dx);
                    if (SMEM_MASK) atomicOr(&sbits[g2 >> 5], 1u << (g2 & 31));
                    else P.gmask[g2] = 1;
                }
            }
        }
        __syncthreads();
    }
    if (tid == 0) *P.nactive = s_count;
}

inline int launch_resolve(const PassParams &P, cudaStream_t st)
{
    const size_t bytes = (size_t)((P.G + 31) / 32) * 4;
    if (P.gh > 4 * 1024) return -1; // MAX_ROWS rows per thread
    int nt = P.gh < 1024 ? ((P.gh + 31) / 32) * 32 : 1024;
    if (nt < 256) nt = 256; // the all-active fast path is a plain strided fill
    if (bytes <= 200 * 1024) {
        cudaFuncSetAttribute(k_resolve<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        k_resolve<true><<<1, nt, bytes, st>>>(P);
    } else {
        k_resolve<false><<<1, nt, 0, st>>>(P);
    }
    return 1;
}

} // namespace nlk
