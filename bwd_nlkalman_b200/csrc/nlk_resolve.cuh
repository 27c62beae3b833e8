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
// The step count of the kernel above, gw + (R+1)(gh-1), is a chain of block barriers, and a
// step is one warp-serial instruction sequence.  This version
//   * handles a BLOCK of C = 4 consecutive columns per row and step, in registers; row i runs
//     R columns plus one block behind row i-1:
//         row i, step s  ->  block b = s - sb(i),  columns [C b - o_i, C b - o_i + C),
//         o_i = (R i) mod C,  sb(i) = i + floor(R i / C)          (C (s - i) - R i = C b - o_i)
//     so row i-1 has always finished column j + R before row i reaches column j; steps drop to
//     about gh + (gw + R gh) / C;
//   * publishes, after a block, the (C + 2R)-bit fields that the row's active cells mark in
//     rows i+1, i+2 (columns C b - o_i - R ..); the reader ORs the field of row i-dy into its
//     window at offset (C+R)(dy-1): always at or ahead of its own position;
//     (one shared-memory word per row, double-buffered by step parity, one block barrier per step);
//   * runs straight-line code: k_resolve_pack (whole GPU, a few microseconds) rewrites the
//     bitmaps into the per-row block layout with every field pre-shifted to where the step
//     ORs it, cells outside the row flagged "already processed", and a pad block on either
//     side of a row so that steps before / after the row's band need no per-row test, stored
//     step-major (pk[s][row]) so that a warp's 32 rows read 512 contiguous bytes per step;
//   * records the active cells as one nibble per block in a register, stored every 8 steps.
// Packed cell c of a block (32 bits):  bits 1+c.. : own row, columns j+1..j+R
//                                      bits 8+c.. : row i+1 (columns j-R..j+R),  bits 16+c.. : row i+2
//                                      word 0 only, bits 28..31: cells of the block outside the row
constexpr int RB_C = 4, RB_Q = 4;

__host__ __device__ inline int resolve_blocks_per_row(int gw) { return (gw + RB_C - 1) / RB_C + 1; }

// number of steps of the blocked kernel (every row must also reach its last record word: a few
// steps beyond the wavefront's end), rounded up to the unroll factor
__host__ __device__ inline int resolve_steps(int gw, int gh, int R)
{
    const int n = gh - 1 + (gw + R * (gh - 1) + RB_C - 1) / RB_C + 12;
    return (n + RB_Q - 1) / RB_Q * RB_Q + 96;     // + pad steps: the replay loop runs whole bodies and fetches ahead
}

// pk[s][i] (uint4): the block row i handles at step s, or a pad block (all four cells flagged
// as outside the row) before / after the row's band and for the lanes past the last row.  Step
// major: the 32 rows of a warp read 512 contiguous bytes per step.
template <int R>
__global__ void k_resolve_pack(const PassParams P, unsigned int *__restrict__ pk, int nb, int nthr_blk, int nsteps)
{
    constexpr int C = RB_C, side = 2 * R + 1;
    const long tid0 = blockIdx.x * (long)blockDim.x + threadIdx.x, nthr = (long)gridDim.x * blockDim.x;
    if (*P.any_nbr == 0) {
        // no group marks another grid patch: every patch is processed, nothing to replay
        for (long g = tid0; g < P.G; g += nthr) P.active[g] = (int)g;
        if (tid0 == 0) *P.nactive = P.G;
        return;
    }
    const long n = (long)nsteps * nthr_blk * C;
    for (long t = tid0; t < n; t += nthr) {
        const int c = (int)(t % C);
        const long si = t / C;
        const int i = (int)(si % nthr_blk), s = (int)(si / nthr_blk);
        const int b = s - (i + (R * i) / C);            // block of row i at step s
        const int o = (R * i) % C;
        const int j = C * b + c - o;
        unsigned int v = 0u;
        const bool row_ok = i < P.gh && b >= 0 && b < nb;
        if (row_ok && j >= 0 && j < P.gw) {
            const unsigned int w0 = P.nbr[(long)i * P.gw + j];
            v = ((w0 >> (R * side + R + 1)) & ((1u << R) - 1u)) << (1 + c);
#pragma unroll
            for (int dy = 1; dy <= R; ++dy) v |= ((w0 >> ((dy + R) * side)) & ((1u << side) - 1u)) << (8 * dy + c);
        }
        if (c == 0) {
            // flags of the four cells of the block that lie outside the row
            for (int cc = 0; cc < C; ++cc) {
                const int jj = C * b + cc - o;
                if (!(row_ok && jj >= 0 && jj < P.gw)) v |= 1u << (28 + cc);
            }
        }
        pk[t] = v;
    }
}

// The replay itself: ONE block, lane = grid row, and no block barrier in the step loop.  The 32 rows of a
// warp advance in lockstep and hand the word they publish (the fields their active cells mark in
// the rows below) to the next rows with warp shuffles; between warps it goes through a
// shared-memory ring written by the last R lanes of a warp, one slot per step of the producer's
// band, so a slot doubles as its own "ready" flag (initialised to a sentinel; nothing is ever
// overwritten, no back-pressure).  A warp only runs the steps of its own band and trails the warp
// above by one step plus the ring's visibility latency; a step is a shuffle, the 3-operation
// recurrence per column and a store: ~100 cycles instead of the ~800 of a barrier-separated step.
constexpr unsigned int RS_SENT = 0xffffffffu;   // published words only use bits 0..23
constexpr int RS_Q = 8;                         // steps per unrolled loop body = prefetch distance of the packed blocks
constexpr int RS_PAD = 2 * RS_Q;                // pad steps behind the last one (straight-line loop, no bounds tests)

// A warp publishes its progress (ring slots written so far) every RS_G steps; the warp below checks it
// once per RS_G steps -- a warp-uniform spin on one shared-memory word, bounded: a count that never
// comes (a bug, not a state of the algorithm) raises *timeout instead of hanging the GPU -- and then
// reads the slots with plain loads.
constexpr int RS_G = 4;

__device__ __forceinline__ void prog_wait(const volatile int *prog, int need, int *timeout)
{
    int spins = 0;
    while (*prog < need) {
        if (++spins > (1 << 22)) { *timeout = 1; break; }
    }
}

template <int R>
__global__ void __launch_bounds__(1024) k_resolve_sys(const PassParams P, int rw, const uint4 *__restrict__ pk, int nb,
                                                      unsigned int *__restrict__ rec, int *__restrict__ row_off,
                                                      int ring_len)
{
    constexpr int C = RB_C, Q = RS_Q;
    static_assert(Q % RS_G == 0, "progress granularity divides the loop body");
    extern __shared__ __align__(8) unsigned int s_dyn[];
    const int gh = P.gh;
    const int tid = threadIdx.x, nthr = blockDim.x, warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
    // ring w, slot j: what lanes 31 / 30 of warp w published at step first(w) + j.  Ring `nwarps` is the
    // all-zero input of warp 0.
    uint2 *s_ring = reinterpret_cast<uint2 *>(s_dyn);                                   // [nwarps + 1][ring_len]
    unsigned int *s_act = s_dyn + (size_t)(nwarps + 1) * ring_len * 2;                  // [nthr][rw] active bits, bit 4b+c of row i
    int *s_base = reinterpret_cast<int *>(s_act + (size_t)nthr * rw);                  // [gh+1] row offsets
    __shared__ int s_part[1024];
    __shared__ int s_prog[33];                                                          // slots published per warp

    if (*P.any_nbr == 0) return;   // k_resolve_pack filled the list
    constexpr int FW = C + 2 * R;
    static_assert(FW <= 8 && (C + R) * (R - 1) + FW <= 32 && 1 + C - 1 + R <= 8, "window / field layout");
    constexpr unsigned int fwmask = (1u << FW) - 1u, ownmask = ((1u << (R + C)) - 1u) << 1;
    for (int x = tid; x < (nwarps + 1) * ring_len; x += nthr) s_ring[x] = make_uint2(0u, 0u);
    for (int x = tid; x < nthr * rw; x += nthr) s_act[x] = 0u;
    if (tid <= nwarps) s_prog[tid] = tid == nwarps ? 0x7fffffff : 0;
    __syncthreads();

    const int i = tid;                                  // lanes past the last row only see pad blocks
    const unsigned int m1 = i >= 1 ? fwmask : 0u, m2 = i >= 2 ? fwmask : 0u;
    const int sb = i + (R * i) / C;                     // block b of the row is handled at step sb + b
    // steps in which some row of this warp has work: marks for its first columns arrive at most
    // three blocks before the row starts, and the last record word closes 8 blocks after its end;
    // rounded up to whole loop bodies (the extra steps see pad blocks)
    const int i0 = warp * 32, i1 = i0 + 31;
    const int w_first = max(i0 + (R * i0) / C - 3, 0);
    const int w_len = (i1 + (R * i1) / C + nb + 9 - w_first + 1 + Q - 1) / Q * Q;
    const int p0 = i0 - 32;
    const int pw_first = warp > 0 ? max(p0 + (R * p0) / C - 3, 0) : w_first - 1;   // ring slot of step s-1: s-1-pw_first
    const int in_off = w_first - 1 - pw_first;
    const int src_w = warp > 0 ? warp - 1 : nwarps;
    const volatile uint2 *ring_in = s_ring + (size_t)src_w * ring_len + in_off;
    const volatile int *prog_in = s_prog + src_w;
    volatile unsigned int *ring_out = reinterpret_cast<unsigned int *>(s_ring + (size_t)warp * ring_len) + (31 - lane);
    const uint4 *pcol = pk + (size_t)w_first * nthr + i;     // pk[s][i]
    uint4 q[Q];             // blocks of steps s .. s+Q-1 (fetched Q steps ahead: an L2 round trip)
#pragma unroll
    for (int u = 0; u < Q; ++u) { q[u] = *pcol; pcol += nthr; }
    unsigned int wnd = 0u, acc = 0u, mypub = 0u;
    unsigned int *act_row = s_act + (size_t)i * rw;
    const bool pub_lane = lane >= 32 - R;
    const bool is0 = lane == 0, is1 = lane == 1;
    int b = w_first - sb;
    for (int j0 = 0; j0 < w_len; j0 += Q) {
#pragma unroll
        for (int u = 0; u < Q; ++u) {
            if (u % RS_G == 0) {
                // the warp above has published the slots of the next RS_G steps (or is done: its
                // remaining slots are zero, the initial value)
                prog_wait(prog_in, in_off + j0 + u + RS_G, P.work + 2);   // (counters[4]: replay timeout flag)
                __threadfence_block();
            }
            // what the rows above published at step s-1: lanes 0 (and 1) take the warp above's
            const uint2 pr = make_uint2(ring_in[j0 + u].x, R > 1 ? ring_in[j0 + u].y : 0u);
            unsigned int up1 = __shfl_up_sync(0xffffffffu, mypub, 1);
            unsigned int up2 = R > 1 ? __shfl_up_sync(0xffffffffu, mypub, 2) : 0u;
            up1 = is0 ? pr.x : up1;
            if (R > 1) { up2 = is1 ? pr.x : up2; up2 = is0 ? pr.y : up2; }
            const uint4 x = q[u];
            wnd |= (up1 >> 8) & m1;
            if (R > 1) wnd |= ((up2 >> 16) & m2) << (C + R);
            wnd |= x.x >> 28;                            // cells outside the row count as processed
            // the only serial part: a column is active iff its window bit is clear, and then
            // marks the next R columns of its own row
            wnd |= x.x & ownmask & ~(unsigned int)((int)(wnd << 31) >> 31);
            wnd |= x.y & ownmask & ~(unsigned int)((int)(wnd << 30) >> 31);
            wnd |= x.z & ownmask & ~(unsigned int)((int)(wnd << 29) >> 31);
            wnd |= x.w & ownmask & ~(unsigned int)((int)(wnd << 28) >> 31);
            // bits of the window below C are final: column c of the block was active iff bit c is clear
            mypub = ((x.x & ~(unsigned int)((int)(wnd << 31) >> 31)) | (x.y & ~(unsigned int)((int)(wnd << 30) >> 31)) |
                     (x.z & ~(unsigned int)((int)(wnd << 29) >> 31)) | (x.w & ~(unsigned int)((int)(wnd << 28) >> 31))) &
                    0x00ffffffu;
            if (pub_lane) ring_out[(j0 + u) * 2] = mypub;
            if (u % RS_G == RS_G - 1) {
                __threadfence_block();
                __syncwarp();
                if (lane == 31) *reinterpret_cast<volatile int *>(s_prog + warp) = j0 + u + 1;
            }
            // record: nibble b of the row (pads and cells outside the row are never active); the word of
            // eight blocks fills from the top and is complete when b & 7 == 7
            acc = __funnelshift_r(acc, ~wnd & 0xfu, 4);
            if ((b & 7) == 7 && (unsigned int)(b >> 3) < (unsigned int)rw) act_row[b >> 3] = acc;
            wnd >>= C;
            b += 1;
            q[u] = *pcol;                                // step s + Q (pad rows behind the last step)
            pcol += nthr;
        }
    }
    // the warp below runs a few dozen steps longer than this one: the rest of the ring stays zero
    __threadfence_block();
    __syncwarp();
    if (lane == 31) *reinterpret_cast<volatile int *>(s_prog + warp) = 0x7fffffff;
    __syncthreads();

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
    // hand the record and the row offsets to k_resolve_list (whole GPU): a single block would
    // spend tens of microseconds writing the list
    for (int x = tid; x < gh * rw; x += nthr) rec[x] = s_act[x];
    for (int r = tid; r <= gh; r += nthr) row_off[r] = s_base[r];
    if (tid == 0) *P.nactive = s_base[gh];
}

// the active list in raster order from the record of k_resolve_sys: one warp per grid row
template <int R>
__global__ void k_resolve_list(const PassParams P, int rw, const unsigned int *__restrict__ rec,
                               const int *__restrict__ row_off)
{
    if (*P.any_nbr == 0) return;   // k_resolve_pack filled the list
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= P.gh) return;
    int pos = row_off[r];
    const int o = (R * r) % RB_C;        // record bit 4b+c is column 4b+c-o
    for (int wd = 0; wd < rw; ++wd) {
        const unsigned int word = rec[(size_t)r * rw + wd];
        if ((word >> lane) & 1u) P.active[pos + __popc(word & ((1u << lane) - 1u))] = r * P.gw + wd * 32 + lane - o;
        pos += __popc(word);
    }
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

// pk: scratch of resolve_pack_bytes(gw, gh, R) bytes for the blocked kernel (or nullptr)
inline size_t resolve_pack_bytes(int gw, int gh, int R)
{
    if (gh < 1 || gh > 1024 || R > 2) return 16;
    const int nthr = ((gh + 31) / 32) * 32 < 64 ? 64 : ((gh + 31) / 32) * 32;
    const int rw8 = (resolve_blocks_per_row(gw) + 7) / 8;
    return (size_t)resolve_steps(gw, gh, R > 0 ? R : 1) * nthr * 16 + ((size_t)gh * rw8 + gh + 1) * 4;
}

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
        // R = 0: no group reaches another grid cell, any_nbr stays 0 and the pack kernel fills the list
        int nthr = ((P.gh + 31) / 32) * 32;
        if (nthr < 64) nthr = 64;
        const int nb = resolve_blocks_per_row(P.gw);
        const int rw8 = (nb + 7) / 8;          // record words per row (8 blocks each)
        const int ring_len = nb + 2 * (32 + 8 * P.R) + 48;   // a warp's band and the steps the warp below runs beyond it
        const size_t bb = ((size_t)(nthr / 32 + 1) * ring_len * 2 + (size_t)nthr * rw8 + P.gh + 1) * 4;
        if (bb <= 200 * 1024) {
            const int nsteps = resolve_steps(P.gw, P.gh, P.R);
            const long n = (long)nsteps * nthr * RB_C;
            const int pb = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
            unsigned int *rec = pk + (size_t)nsteps * nthr * 4;
            int *row_off = reinterpret_cast<int *>(rec + (size_t)P.gh * rw8);
            const uint4 *pk4 = reinterpret_cast<const uint4 *>(pk);
            const int lb = (P.gh + 7) / 8;
            if (P.R == 2) {
                k_resolve_pack<2><<<pb, 256, 0, st>>>(P, pk, nb, nthr, nsteps);
                cudaFuncSetAttribute(k_resolve_sys<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bb);
                k_resolve_sys<2><<<1, nthr, bb, st>>>(P, rw8, pk4, nb, rec, row_off, ring_len);
                k_resolve_list<2><<<lb, 256, 0, st>>>(P, rw8, rec, row_off);
            } else {
                k_resolve_pack<1><<<pb, 256, 0, st>>>(P, pk, nb, nthr, nsteps);
                cudaFuncSetAttribute(k_resolve_sys<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bb);
                k_resolve_sys<1><<<1, nthr, bb, st>>>(P, rw8, pk4, nb, rec, row_off, ring_len);
                k_resolve_list<1><<<lb, 256, 0, st>>>(P, rw8, rec, row_off);
            }
            return 3;
        }
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
