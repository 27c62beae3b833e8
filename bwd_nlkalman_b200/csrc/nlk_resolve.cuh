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
// block (thread = grid row).
#pragma once
#include "nlk_common.cuh"
#include <type_traits>

namespace nlk {

// No atomics and almost no work per step: thread i owns grid row i and carries, in a
// 64-bit register, the processed-mask bits of its row for the columns j .. j+63 ahead of
// its position.  An active cell publishes its neighbour bitmap (one shared-memory slot
// per row, double-buffered by step parity).  With skew R+1, the cell that row i-dy handled
// in the previous step is (i-dy, j + dy(R+1) - 1); its members in row i are the 2R+1
// columns starting at j + (dy-1)(R+1), i.e. at or ahead of row i's position, so row i
// ORs that (2R+1)-bit field into its window at offset (dy-1)(R+1) -- R reads per step.
__device__ __forceinline__ unsigned int nbr_field(const unsigned int *wds, int nbw, int start, int width)
{
    const int w0 = start >> 5, sh = start & 31;
    unsigned long long v = wds[w0];
    if (w0 + 1 < nbw) v |= (unsigned long long)wds[w0 + 1] << 32;
    return (unsigned int)(v >> sh) & ((1u << width) - 1u);
}

// FAST: one bitmap word per cell and a 32-bit row window (R <= 4), the usual case
template <int MAX_ROWS, bool FAST>
__global__ void __launch_bounds__(1024) k_resolve(const PassParams P, int rw)
{
    typedef typename std::conditional<FAST, unsigned int, unsigned long long>::type win_t;
    extern __shared__ unsigned int s_dyn[];
    const int gw = P.gw, gh = P.gh, G = P.G, nbw = FAST ? 1 : P.nbw;
    unsigned int *s_pub = s_dyn;                                      // [2][gh][nbw]
    unsigned int *s_act = s_pub + (size_t)2 * gh * nbw;               // [gh][rw] active bits
    int *s_base = reinterpret_cast<int *>(s_act + (size_t)gh * rw);   // [gh+1] row offsets
    __shared__ int s_part[1024];
    const int tid = threadIdx.x, nthr = blockDim.x;

    if (*P.any_nbr == 0) {
        // no group marks another grid patch: every patch is processed
        for (int g = tid; g < G; g += nthr) P.active[g] = g;
        if (tid == 0) *P.nactive = G;
        return;
    }
    const int R = P.R, side = 2 * R + 1, skew = R + 1;
    const int nsteps = gw + skew * (gh - 1);
    for (int x = tid; x < 2 * gh * nbw; x += nthr) s_pub[x] = 0u;
    __syncthreads();

    // Nothing is written to global memory inside the step loop (a block barrier after a
    // global store waits for the store's round trip), and word 0 of each row's cells is
    // fetched D steps ahead into a register pipeline (loop unrolled by D: static slots;
    // only the first touch of a 128-byte line goes to L2, the rest are L1 hits).
    constexpr int D = 4;
    unsigned int q[MAX_ROWS][D], accw[MAX_ROWS];
    win_t win[MAX_ROWS];
    const unsigned int fmask = (1u << side) - 1u;
#pragma unroll
    for (int m = 0; m < MAX_ROWS; ++m) {
        const int i = tid + m * nthr;
        accw[m] = 0u;
        win[m] = 0;
#pragma unroll
        for (int u = 0; u < D; ++u) {
            const int j = u - skew * i;
            q[m][u] = (i < gh && j >= 0 && j < gw) ? P.nbr[(long)(i * gw + j) * nbw] : 0u;
        }
    }
    for (int t0 = 0; t0 < nsteps; t0 += D) {
#pragma unroll
        for (int u = 0; u < D; ++u) {
            const int t = t0 + u;
            if (t >= nsteps) break;
            const unsigned int *pub_rd = s_pub + (size_t)((t + 1) & 1) * gh * nbw; // written at step t-1
            unsigned int *pub_wr = s_pub + (size_t)(t & 1) * gh * nbw;
#pragma unroll
            for (int m = 0; m < MAX_ROWS; ++m) {
                const int i = tid + m * nthr;
                if (i >= gh) break;
                const int j = t - skew * i;
                const unsigned int bits0 = q[m][u];
                const int jn = j + D;
                if (jn >= 0 && jn < gw) q[m][u] = P.nbr[(long)(i * gw + jn) * nbw];
                if (j >= gw) {
                    // a row past its end must not leave a stale bitmap for the rows below
                    if (j < gw + 2) for (int wd = 0; wd < nbw; ++wd) pub_wr[(size_t)i * nbw + wd] = 0u;
                    continue;
                }
                if (j < 1 - R * skew) continue; // nothing above has reached this row's columns yet
                // marks from the rows above (their previous step); also collected while this
                // row has not started (j < 0): they concern columns it will visit
                win_t wnd = win[m];
                for (int dy = 1; dy <= R && dy <= i; ++dy) {
                    const unsigned int f = FAST ? ((pub_rd[i - dy] >> ((dy + R) * side)) & fmask)
                                                : nbr_field(pub_rd + (size_t)(i - dy) * nbw, nbw, (dy + R) * side, side);
                    wnd |= (win_t)f << ((dy - 1) * skew);
                }
                if (j >= 0) {
                    const bool done = wnd & 1;
                    unsigned int w0 = 0u, w1 = 0u, w2 = 0u, w3 = 0u;
                    if (!done) {
                        // own row: columns j+1 .. j+R (dy = 0, dx = 1..R)
                        w0 = bits0;
                        unsigned int f;
                        if (FAST) {
                            f = (w0 >> (R * side)) & fmask;
                        } else {
                            const long gb = (long)(i * gw + j) * nbw;
                            if (nbw > 1) w1 = P.nbr[gb + 1];
                            if (nbw > 2) w2 = P.nbr[gb + 2];
                            if (nbw > 3) w3 = P.nbr[gb + 3];
                            const int start = R * side, w_0 = start >> 5, sh = start & 31;
                            const unsigned int lo = w_0 == 0 ? w0 : (w_0 == 1 ? w1 : (w_0 == 2 ? w2 : w3));
                            const unsigned int hi = w_0 == 0 ? w1 : (w_0 == 1 ? w2 : (w_0 == 2 ? w3 : 0u));
                            f = (unsigned int)((((unsigned long long)hi << 32) | lo) >> sh) & fmask;
                        }
                        wnd |= (win_t)(f >> R); // bit 0 = own column
                    }
                    unsigned int *pw = pub_wr + (size_t)i * nbw;
                    pw[0] = w0;
                    if (nbw > 1) pw[1] = w1;
                    if (nbw > 2) pw[2] = w2;
                    if (nbw > 3) pw[3] = w3;
                    accw[m] |= (done ? 0u : 1u) << (j & 31);
                    if ((j & 31) == 31 || j == gw - 1) {
                        s_act[(size_t)i * rw + (j >> 5)] = accw[m];
                        accw[m] = 0u;
                    }
                }
                win[m] = wnd >> 1;
            }
            __syncthreads();
        }
    }

    // active list in raster order: per-row counts, block scan, then one warp per row
    for (int i = tid; i < gh; i += nthr) {
        int c = 0;
        for (int wd = 0; wd < rw; ++wd) c += __popc(s_act[(size_t)i * rw + wd]);
        s_base[i + 1] = c;
    }
    if (tid == 0) s_base[0] = 0;
    __syncthreads();
    int carry = 0;
    for (int c0 = 0; c0 < gh; c0 += nthr) {
        const int i = c0 + tid;
        s_part[tid] = i < gh ? s_base[i + 1] : 0;
        __syncthreads();
        for (int off = 1; off < nthr; off <<= 1) {
            const int a = tid >= off ? s_part[tid - off] : 0;
            __syncthreads();
            s_part[tid] += a;
            __syncthreads();
        }
        if (i < gh) s_base[i + 1] = carry + s_part[tid];
        carry += s_part[nthr - 1];
        __syncthreads();
    }
    const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
    for (int i = warp; i < gh; i += nwarps) {
        int pos = s_base[i];
        for (int wd = 0; wd < rw; ++wd) {
            const unsigned int word = s_act[(size_t)i * rw + wd];
            if ((word >> lane) & 1u) P.active[pos + __popc(word & ((1u << lane) - 1u))] = i * gw + wd * 32 + lane;
            pos += __popc(word);
        }
    }
    if (tid == 0) *P.nactive = s_base[gh];
}


// ---- blocked wavefront (one bitmap word per cell, R <= 2: every BASELINE configuration) -------
// The step count of the kernel above, gw + (R+1)(gh-1), is a chain of block barriers.  Here a
// step handles a BLOCK of C = 4 consecutive columns per row, sequentially in registers, and
// row i runs R columns plus one block behind row i-1:
//     row i, step s  ->  columns [j0, j0 + C),   j0 = C (s - i) - R i
// so that row i-1 has always finished column j + R before row i reaches column j.  Steps drop
// to gh - 1 + ceil((gw + R (gh-1)) / C).  After its block a row publishes, for each dy, the
// (C + 2R)-bit field of columns j0-R .. j0+C-1+R that its active cells mark in row i+dy (one
// shared-memory word, double-buffered by step parity).  The reader ORs the field of row i-dy
// into its window at offset (C+R)(dy-1): always at or ahead of its own position.
//
// A step is one warp-serial chain, so its instruction count is what matters: k_resolve_pack
// (whole GPU, a few microseconds) first rewrites the bitmaps into the per-row block layout
// the chain consumes -- one aligned 16-byte load per step, fields pre-extracted:
//     bits 0..R-1: own row, columns j+1..j+R;  bits 8..: row i+1;  bits 16..: row i+2
constexpr int RB_C = 4;

__host__ __device__ inline int resolve_blocks_per_row(int gw) { return (gw + RB_C - 1) / RB_C + 1; }

template <int R>
__global__ void k_resolve_pack(const PassParams P, unsigned int *__restrict__ pk, int nb)
{
    constexpr int C = RB_C, side = 2 * R + 1;
    const long n = (long)P.gh * nb * C;
    if (*P.any_nbr == 0) return;   // every patch is processed: nothing to replay
    for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x) {
        const int c = (int)(t % C);
        const long ib = t / C;
        const int b = (int)(ib % nb), i = (int)(ib / nb);
        const int j = C * b + c - (R * i) % C;
        unsigned int v = 0u;
        if (j >= 0 && j < P.gw) {
            const unsigned int w0 = P.nbr[(long)i * P.gw + j];
            v = (w0 >> (R * side + R + 1)) & ((1u << R) - 1u);
#pragma unroll
            for (int dy = 1; dy <= R; ++dy) v |= ((w0 >> ((dy + R) * side)) & ((1u << side) - 1u)) << (8 * dy);
        }
        pk[t] = v;
    }
}

template <int R>
__global__ void __launch_bounds__(1024) k_resolve_blk(const PassParams P, int rw, const uint4 *__restrict__ pk, int nb)
{
    constexpr int C = RB_C;
    extern __shared__ unsigned int s_dyn[];
    const int gw = P.gw, gh = P.gh, G = P.G;
    unsigned int *s_pub = s_dyn;                              // [2][gh]
    unsigned int *s_act = s_pub + (size_t)2 * gh;             // [gh][rw] active bits
    int *s_base = reinterpret_cast<int *>(s_act + (size_t)gh * rw); // [gh+1] row offsets
    __shared__ int s_part[1024];
    const int tid = threadIdx.x, nthr = blockDim.x;

    if (*P.any_nbr == 0) {
        // no group marks another grid patch: every patch is processed
        for (int g = tid; g < G; g += nthr) P.active[g] = g;
        if (tid == 0) *P.nactive = G;
        return;
    }
    constexpr int FW = C + 2 * R;
    static_assert(FW <= 8 && (C + R) * (R - 1) + FW <= 32 && R + C <= 8, "window / field layout");
    constexpr unsigned int ownmask = (1u << R) - 1u, fwmask = (1u << FW) - 1u, cmask = (1u << C) - 1u;
    const int nsteps = gh - 1 + (gw + R * (gh - 1) + C - 1) / C;
    for (int x = tid; x < 2 * gh; x += nthr) s_pub[x] = 0u;
    for (int x = tid; x < gh * rw; x += nthr) s_act[x] = 0u;
    __syncthreads();

    const int i = tid;
    const bool live = i < gh;
    const int o = (R * i) % C, sb = i + (R * i) / C;    // block b of the row is handled at step sb + b
    const uint4 *prow = pk + (size_t)(live ? i : 0) * nb;
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    uint4 q[2];             // blocks of steps s and s+1 (fetched two steps ahead)
#pragma unroll
    for (int u = 0; u < 2; ++u) q[u] = (live && u - sb >= 0 && u - sb < nb) ? prow[u - sb] : zero4;
    unsigned int wnd = 0u;
    for (int s0 = 0; s0 < nsteps; s0 += 2) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int s = s0 + u;
            const unsigned int *pub_rd = s_pub + (size_t)((s + 1) & 1) * gh; // written at step s-1
            unsigned int *pub_wr = s_pub + (size_t)(s & 1) * gh;
            const int b = s - sb, j0 = C * b - o;
            // in the band: the block overlaps the row, or marks for its first columns arrive
            if (live && j0 > -32 && j0 < gw) {
#pragma unroll
                for (int dy = 1; dy <= R; ++dy)
                    if (i >= dy) wnd |= ((pub_rd[i - dy] >> (8 * dy)) & fwmask) << ((C + R) * (dy - 1));
                const unsigned int x[C] = {q[u].x, q[u].y, q[u].z, q[u].w};
                // the only serial part: a column is active iff its window bit is clear, and then
                // marks the next R columns of its own row (cells outside the row are packed as 0)
#pragma unroll
                for (int c = 0; c < C; ++c) wnd |= ((wnd >> c) & 1u) ? 0u : (x[c] & ownmask) << (c + 1);
                // bits of the window below C are final: column j0+c was active iff bit c is clear
                unsigned int vm = cmask;
                if (j0 < 0) vm &= cmask << min(-j0, C);
                if (j0 + C > gw) vm &= cmask >> (j0 + C - gw);
                const unsigned int act = ~wnd & vm;
                unsigned int pw = 0u;
#pragma unroll
                for (int c = 0; c < C; ++c) pw |= (((act >> c) & 1u) ? x[c] : 0u) << c;
                pub_wr[i] = pw;   // bytes 1.. hold the fields for rows i+1.. (byte 0: not read)
                if (act) {
                    // record: row i is the only writer of its words
                    const int jb = max(j0, 0);
                    const unsigned int bits = j0 < 0 ? act >> (-j0) : act;
                    unsigned int *wp = s_act + (size_t)i * rw + (jb >> 5);
                    atomicOr(wp, bits << (jb & 31));
                    if ((jb & 31) + C > 32 && (jb >> 5) + 1 < rw) atomicOr(wp + 1, bits >> (32 - (jb & 31)));
                }
                wnd >>= C;
                q[u] = (b + 2 < nb) ? prow[b + 2] : zero4;   // block of step s+2 (b + 2 >= 0 here or zero anyway)
            } else if (live && j0 >= gw && j0 < gw + 2 * C) {
                pub_wr[i] = 0u;   // a finished row leaves no stale marks (both parities)
            } else if (live && j0 <= -32 && j0 + 2 * C > -32) {
                q[u] = (b + 2 >= 0 && b + 2 < nb) ? prow[b + 2] : zero4;   // about to enter the band
            }
            __syncthreads();
        }
    }

    // active list in raster order: per-row counts, block scan, then one warp per row
    for (int r = tid; r < gh; r += nthr) {
        int c = 0;
        for (int wd = 0; wd < rw; ++wd) c += __popc(s_act[(size_t)r * rw + wd]);
        s_base[r + 1] = c;
    }
    if (tid == 0) s_base[0] = 0;
    __syncthreads();
    {
        s_part[tid] = tid < gh ? s_base[tid + 1] : 0;
        __syncthreads();
        for (int off = 1; off < nthr; off <<= 1) {
            const int a = tid >= off ? s_part[tid - off] : 0;
            __syncthreads();
            s_part[tid] += a;
            __syncthreads();
        }
        if (tid < gh) s_base[tid + 1] = s_part[tid];
        __syncthreads();
    }
    const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
    for (int r = warp; r < gh; r += nwarps) {
        int pos = s_base[r];
        for (int wd = 0; wd < rw; ++wd) {
            const unsigned int word = s_act[(size_t)r * rw + wd];
            if ((word >> lane) & 1u) P.active[pos + __popc(word & ((1u << lane) - 1u))] = r * gw + wd * 32 + lane;
            pos += __popc(word);
        }
    }
    if (tid == 0) *P.nactive = s_base[gh];
}

// strip-sharded pass: restrict group_filter to the processed patches of grid rows [gy0, gy1).
// `active` is in raster order, so they form one contiguous range: its bounds go into the ticket
// counter (first entry) and the entry count (one past the last).
__global__ void k_active_range(const PassParams P)
{
    if (threadIdx.x != 0) return;
    const int n = *P.nactive;
    const int keys[2] = {P.gy0 * P.gw, P.gy1 * P.gw};
    int res[2];
    for (int q = 0; q < 2; ++q) {
        int lo = 0, hi = n;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (P.active[mid] < keys[q]) lo = mid + 1; else hi = mid;
        }
        res[q] = lo;
    }
    *P.work = res[0];
    *P.nactive = res[1];
}

__global__ void k_set_flag(int *p, int v) { *p = v; }

// pk: scratch of gh * resolve_blocks_per_row(gw) * 16 bytes for the blocked kernel (or nullptr)
inline int launch_resolve(const PassParams &P, unsigned int *pk, cudaStream_t st)
{
    if (P.gh > 4 * 1024) return -1;                       // MAX_ROWS rows per thread
    if ((P.R - 1) * (P.R + 1) + 2 * P.R + 1 > 64 || P.nbw > 4) return -1; // 64-bit row window
    const int rw = (P.gw + 31) / 32;
    const size_t bytes = ((size_t)2 * P.gh * P.nbw + (size_t)P.gh * rw + P.gh + 1) * 4;
    if (bytes > 200 * 1024) return -1;
    int nt = P.gh < 1024 ? ((P.gh + 31) / 32) * 32 : 1024;
    if (nt < 256) nt = 256; // the all-active fast path is a plain strided fill
    const bool fast = P.nbw == 1 && P.R <= 4;
    if (P.nbw == 1 && P.R <= 2 && P.gh <= 1024 && pk != nullptr) {
        // R = 0: no group reaches another grid cell, any_nbr stays 0 and the kernel only fills the list
        const size_t bb = ((size_t)2 * P.gh + (size_t)P.gh * rw + P.gh + 1) * 4;
        int nthr = ((P.gh + 31) / 32) * 32;
        if (nthr < 256) nthr = 256;
        const int nb = resolve_blocks_per_row(P.gw);
        const long n = (long)P.gh * nb * RB_C;
        const int pb = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
        if (P.R == 2) {
            k_resolve_pack<2><<<pb, 256, 0, st>>>(P, pk, nb);
            cudaFuncSetAttribute(k_resolve_blk<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bb);
            k_resolve_blk<2><<<1, nthr, bb, st>>>(P, rw, reinterpret_cast<const uint4 *>(pk), nb);
        } else {
            k_resolve_pack<1><<<pb, 256, 0, st>>>(P, pk, nb);
            cudaFuncSetAttribute(k_resolve_blk<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bb);
            k_resolve_blk<1><<<1, nthr, bb, st>>>(P, rw, reinterpret_cast<const uint4 *>(pk), nb);
        }
        return 2;
    }
#define NLK_LAUNCH_RESOLVE(MR, F)                                                                  \
    do {                                                                                           \
        cudaFuncSetAttribute(k_resolve<MR, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes); \
        k_resolve<MR, F><<<1, nt, bytes, st>>>(P, rw);                                             \
    } while (0)
    if (P.gh <= nt) { if (fast) NLK_LAUNCH_RESOLVE(1, true); else NLK_LAUNCH_RESOLVE(1, false); }
    else { if (fast) NLK_LAUNCH_RESOLVE(4, true); else NLK_LAUNCH_RESOLVE(4, false); }
#undef NLK_LAUNCH_RESOLVE
    return 1;
}

} // namespace nlk
