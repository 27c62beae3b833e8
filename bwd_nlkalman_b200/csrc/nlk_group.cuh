// group_filter: per processed reference patch, the DCT-domain statistics, the
// Kalman / Wiener / RTS gain, the update of the patch group, the inverse transform and
// the weighted aggregation (reference src/nlkalman.c:713-932 and :1600-1845).
//
// One thread block per group (persistent over the active list).  Candidates are taken
// in chunks: their patches are gathered coalesced from HBM/L2 into shared-memory tiles
// (one tile per candidate x source x channel), one thread transforms one tile, and one
// thread per DCT coefficient runs the reference's Welford recurrences over the
// candidates in sorted order.  After the gains are known the group members are
// gathered again, transformed, shrunk, inverse-transformed and added to the
// accumulator image with vector reductions (red.global.add.v4.f32 for 3 channels).
#pragma once
#include "nlk_common.cuh"
#include "nlk_dct.cuh"

namespace nlk {

constexpr int GF_THREADS = 256;

__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_v2(float *addr, float a, float b)
{
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" :: "l"(addr), "f"(a), "f"(b) : "memory");
}

// accw[pix][0..ch] += (v[0..ch-1], wsum)
template <int CH_T>
__device__ __forceinline__ void accumulate_pixel(float *p, const float *v, float wsum, int ch)
{
    if constexpr (CH_T == 3) {
        red_add_v4(p, v[0], v[1], v[2], wsum);
    } else if constexpr (CH_T == 1) {
        red_add_v2(p, v[0], wsum);
    } else {
        for (int c = 0; c < ch; ++c) atomicAdd(p + c, v[c]);
        atomicAdd(p + ch, wsum);
    }
}

__device__ __forceinline__ float block_sum(float v, float *s_red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < GF_THREADS / 32; ++i) t += s_red[i];
    return t;
}

// stage the window rows y0..y0+wh-1, columns x0..x0+wlen/ch-1 of an HWC image
__device__ __forceinline__ void stage_window(float *__restrict__ win, int wrow, const float *__restrict__ img,
                                             int w, int ch, int x0, int y0, int wlen, int wh)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int row = warp; row < wh; row += GF_THREADS / 32) {
        const float *src = img + ((long)(y0 + row) * w + x0) * ch;
        float *dst = win + row * wrow;
        for (int j = lane; j < wlen; j += 32) dst[j] = src[j];
    }
}

// ---- separable transforms with one thread per tile LINE ----------------------------------------
// A chunk holds at most ~100 tiles for 256 threads, and a 12x12 transform by one thread is a chain
// of ~3,000 dependent-latency instructions: (tile, row) then (tile, column) work items keep every
// thread busy and cut the serial chain per phase by the patch side.
template <int PSZ_T, bool INVERSE>
__device__ __forceinline__ void line_dct(float *r, int psz)
{
    if constexpr (PSZ_T != 0) {
        float (&a)[PSZ_T] = *reinterpret_cast<float (*)[PSZ_T]>(r);
        if (INVERSE) dct1d_inv<PSZ_T>(a); else dct1d_fwd<PSZ_T>(a);
    } else {
        float o[MAX_PSZ];
        const float *T = c_dct[psz];
        for (int k = 0; k < psz; ++k) {
            float acc = 0.f;
            for (int j = 0; j < psz; ++j) acc = fmaf(INVERSE ? T[j * psz + k] : T[k * psz + j], r[j], acc);
            o[k] = acc;
        }
        for (int k = 0; k < psz; ++k) r[k] = o[k];
    }
}

// forward 2-D transform of ntiles patches out of the staged windows into tiles[t * TS]; src_of(t) is
// the patch's first sample (nullptr: no such tile), samples at src[y * wrow + x * ch].  Ends with
// a block barrier.
// sub_off != 0: the transform of src - (src + sub_off), two windows at a fixed distance
template <int PSZ_T, class SrcOf>
__device__ __forceinline__ void block_fwd_tiles(float *tiles, int TS, int ntiles, int psz, int wrow, int ch, SrcOf src_of,
                                                long sub_off = 0)
{
    constexpr int NR = PSZ_T ? PSZ_T : MAX_PSZ;
    for (int it = threadIdx.x; it < ntiles * psz; it += GF_THREADS) {
        const int t = it / psz, y = it - t * psz;
        const float *src = src_of(t);
        if (!src) continue;
        float r[NR];
#pragma unroll
        for (int i = 0; i < NR; ++i)
            if (i < psz) r[i] = sub_off ? src[y * wrow + i * ch] - src[sub_off + y * wrow + i * ch] : src[y * wrow + i * ch];
        line_dct<PSZ_T, false>(r, psz);
        float *d = tiles + t * TS + y * psz;
#pragma unroll
        for (int i = 0; i < NR; ++i) if (i < psz) d[i] = r[i];
    }
    __syncthreads();
    for (int it = threadIdx.x; it < ntiles * psz; it += GF_THREADS) {
        const int t = it / psz, x = it - t * psz;
        if (!src_of(t)) continue;
        float *d = tiles + t * TS + x;
        float r[NR];
#pragma unroll
        for (int i = 0; i < NR; ++i) if (i < psz) r[i] = d[i * psz];
        line_dct<PSZ_T, false>(r, psz);
#pragma unroll
        for (int i = 0; i < NR; ++i) if (i < psz) d[i * psz] = r[i];
    }
    __syncthreads();
}

// inverse 2-D transform, in place, of the ntiles tiles tile_of(t) (an index into `tiles`)
template <int PSZ_T, class TileOf>
__device__ __forceinline__ void block_inv_tiles(float *tiles, int TS, int ntiles, int psz, TileOf tile_of)
{
    constexpr int NR = PSZ_T ? PSZ_T : MAX_PSZ;
    for (int it = threadIdx.x; it < ntiles * psz; it += GF_THREADS) {
        const int t = it / psz, x = it - t * psz;
        float *d = tiles + tile_of(t) * TS + x;
        float r[NR];
#pragma unroll
        for (int i = 0; i < NR; ++i) if (i < psz) r[i] = d[i * psz];
        line_dct<PSZ_T, true>(r, psz);
#pragma unroll
        for (int i = 0; i < NR; ++i) if (i < psz) d[i * psz] = r[i];
    }
    __syncthreads();
    for (int it = threadIdx.x; it < ntiles * psz; it += GF_THREADS) {
        const int t = it / psz, y = it - t * psz;
        float *d = tiles + tile_of(t) * TS + y * psz;
        float r[NR];
#pragma unroll
        for (int i = 0; i < NR; ++i) if (i < psz) r[i] = d[i];
        line_dct<PSZ_T, true>(r, psz);
#pragma unroll
        for (int i = 0; i < NR; ++i) if (i < psz) d[i] = r[i];
    }
    __syncthreads();
}

template <int PSZ_T, int CH_T>
__global__ void __launch_bounds__(GF_THREADS, 2)
k_group_filter(const PassParams P, int CT, int kcap, int wrow, int win_floats)
{
    constexpr int JPT = (PSZ_T && CH_T) ? (PSZ_T * PSZ_T * CH_T + GF_THREADS - 1) / GF_THREADS
                                        : (MAX_PSZ * MAX_PSZ * MAX_CH + GF_THREADS - 1) / GF_THREADS;
    const int psz = PSZ_T ? PSZ_T : P.psz;
    const int ch = CH_T ? CH_T : P.ch;
    const int pp = psz * psz, cpp = ch * pp;
    const int TS = pp + 1;
    const int tid = threadIdx.x;

    extern __shared__ __align__(16) float smem[];
    float *tiles = smem;                                  // [CT][TS]
    float *winS = tiles + (size_t)CT * TS;                // source window (src, later in1)
    float *winP = winS + win_floats;                      // previous-frame window
    float *s_a = winP + (P.has_prev ? win_floats : 0);    // [cpp] gain
    float *s_m = s_a + cpp;                               // [cpp] group mean (M0 or M1)
    uint32_t *s_cand = reinterpret_cast<uint32_t *>(s_m + cpp); // [kcap] sorted candidates
    int *s_grp = reinterpret_cast<int *>(s_cand + kcap);  // [tagg] sorted index of each group member
    float *s_red = reinterpret_cast<float *>(s_grp + max(P.tagg, 1)); // [8]
    float *W = s_red + 8;                                 // [pp] aggregation window
    __shared__ int s_nagg;

    const int nactive = *P.nactive;
    const float sigma2 = P.sigma2;
    for (int e = tid; e < pp; e += GF_THREADS) W[e] = c_win[psz][e];

    // *P.work = first entry of the active list to handle (0 unless the pass is strip-sharded)
    for (int ai = *P.work + blockIdx.x; ai < nactive; ai += gridDim.x) {
        const int g = P.active[ai];
        const GroupHdr hd = P.hdr[g];
        const int gy = g / P.gw, gx = g - gy * P.gw;
        const int px = gx * P.step, py = gy * P.step;
        const int prev_p = hd.flags & HDR_PREV_P;
        const int k = hd.nk;
        const int np0 = hd.np0;
        if (!P.smooth && k == 0) continue; // filter, k <= 1: nothing aggregated (:815-849,:857)

        __syncthreads(); // previous group done with shared memory
        if (P.smooth && np0 == 0) {
            // reference :1795-1804: the filtered patch at p, weight 1/1e-6, mask untouched
            const float wgt = __fdiv_rn(1.f, 1e-6f);
            for (int e = tid; e < pp; e += GF_THREADS) {
                const int hy = e / psz, hx = e - hy * psz;
                const long pix = (long)(py + hy) * P.w + px + hx;
                const float wW = __fmul_rn(wgt, W[e]);
                float v[MAX_CH];
                for (int c = 0; c < ch; ++c) v[c] = __fmul_rn(wW, P.in1[pix * ch + c]);
                accumulate_pixel<CH_T>(P.accw + pix * (ch + 1), v, wW, ch);
            }
            if (P.dbg_vp && tid == 0) P.dbg_vp[g] = 0.f;
            continue;
        }

        // window geometry of this group (the search window, reference :637-639)
        const int r = P.smooth ? P.r_t : (prev_p ? P.r_t : P.r_x);
        const int x0 = max(px - r, 0), x1 = min(px + r, P.w - psz);
        const int y0 = max(py - r, 0), y1 = min(py + r, P.h - psz);
        const int wlen = (x1 - x0 + psz) * ch, wh = y1 - y0 + psz;
        stage_window(winS, wrow, P.src, P.w, ch, x0, y0, wlen, wh);
        if (prev_p) stage_window(winP, wrow, P.prev0, P.w, ch, x0, y0, wlen, wh);

        for (int i = tid; i < k; i += GF_THREADS) s_cand[i] = P.cand[(long)g * P.kstride + i];
        __syncthreads();

        // group members: the first tagg candidates with a valid previous patch, or, when
        // there is none (filter only), the first tagg candidates (:779-793, :857, :1669, :1737)
        if (tid < 32) {
            int cnt = 0;
            for (int b0 = 0; b0 < k && cnt < P.tagg; b0 += 32) {
                const int i = b0 + tid;
                const uint32_t cd = i < k ? s_cand[i] : 0u;
                const int take = (i < k) && (np0 > 0 ? cand_prev(cd) : 1);
                const unsigned int bal = __ballot_sync(0xffffffffu, take);
                const int rank = cnt + __popc(bal & ((1u << tid) - 1u));
                if (take && rank < P.tagg) s_grp[rank] = i;
                cnt += __popc(bal);
            }
            if (tid == 0) s_nagg = min(cnt, P.tagg);
        }

        // tiles of candidate slot i: (i*nsrc + s)*ch + c, s = 0 source, 1 previous frame
        const int nsrc = prev_p ? 2 : 1;
        const int tpc = nsrc * ch;            // tiles per candidate
        const int cc = CT / tpc;              // candidates per chunk
        const bool resident = (k <= cc) && !P.has_bsic;

        // ---- pass 1: statistics over the k candidates ------------------------------------
        float M1[JPT], V1[JPT], Mp[JPT], V0[JPT], V01[JPT], Mg[JPT];
#pragma unroll
        for (int u = 0; u < JPT; ++u) M1[u] = V1[u] = Mp[u] = V0[u] = V01[u] = Mg[u] = 0.f;
        int n1 = 0, n0 = 0;
        for (int c0 = 0; c0 < k; c0 += cc) {
            const int cnt = min(cc, k - c0);
            if (c0) __syncthreads();
            // transform straight out of the staged windows, one thread per tile line
            block_fwd_tiles<PSZ_T>(tiles, TS, cnt * tpc, psz, wrow, ch, [&](int t) -> const float * {
                const int slot = t / tpc, rr = t - slot * tpc;
                const int s = rr >= ch, c = rr - s * ch;
                const uint32_t cd = s_cand[c0 + slot];
                if (s == 1 && !cand_prev(cd)) return nullptr;
                return (s ? winP : winS) + (cand_y(cd) - y0) * wrow + (cand_x(cd) - x0) * ch + c;
            });
            // one thread per coefficient, candidates in sorted order
            {
                const int cstride = tpc * TS;
#pragma unroll
                for (int u = 0; u < JPT; ++u) {
                    const int j = tid + u * GF_THREADS;
                    if (j >= cpp) continue;
                    const int c = j / pp, e = j - c * pp;
                    const float *tp = tiles + c * TS + e;      // source tile of slot 0
                    const float *tq = tp + ch * TS;            // previous-frame tile of slot 0
                    int m1 = n1, m0 = n0;
                    float aM1 = M1[u], aV1 = V1[u], aMp = Mp[u], aV0 = V0[u], aV01 = V01[u], aMg = Mg[u];
                    for (int i = 0; i < cnt; ++i, tp += cstride, tq += cstride) {
                        const float p = *tp;
                        m1 += 1;
                        const float delta = p - aM1;
                        aM1 = fmaf(delta, c_inv[m1], aM1);            // :765
                        aV1 = fmaf(delta, p - aM1, aV1);              // :766
                        if (cand_prev(s_cand[c0 + i])) {              // (implies prev_p)
                            m0 += 1;
                            const float inp0 = c_inv[m0];
                            const float q = *tq;
                            const float d0 = q - aMp;                 // :770-775 / :1654-1659
                            aMp = fmaf(d0, inp0, aMp);
                            aV0 = fmaf(d0, q - aMp, aV0);
                            const float t = q - p;
                            aV01 = fmaf(t, t, aV01);                  // :777-778
                            if (m0 <= P.tagg) aMg = fmaf(q - aMg, inp0, aMg); // :783
                        }
                    }
                    M1[u] = aM1; V1[u] = aV1; Mp[u] = aMp; V0[u] = aV0; V01[u] = aV01; Mg[u] = aMg;
                }
                // the counters advance identically in every thread
                for (int i = 0; i < cnt; ++i) { n1 += 1; n0 += cand_prev(s_cand[c0 + i]); }
            }
        }
        __syncthreads();
        const int nagg = s_nagg;
        // with a basic estimate the group holds the noisy patches themselves (:785, :853):
        // the source window is no longer needed, restage it from the noisy frame
        if (P.has_bsic) stage_window(winS, wrow, P.in1, P.w, ch, x0, y0, wlen, wh);

        // ---- gains (:858-904, :1763-1777) -------------------------------------------------
        float vsum = 0.f;
        {
            const float inp1 = c_inv[max(n1, 1)];
            const float inp0 = c_inv[n0];
            const float s2 = P.has_bsic ? 0.f : sigma2;
#pragma unroll
            for (int u = 0; u < JPT; ++u) {
                const int j = tid + u * GF_THREADS;
                if (j < cpp) {
                    float v1 = V1[u], v0 = V0[u], v01 = V01[u];
                    v1 *= inp1;                                 // :805
                    if (n0) { v0 *= inp0; v01 *= inp0; }        // :806-810
                    float a, m;
                    if (P.smooth) {
                        a = __fdiv_rn(v1, v1 + P.beta_t * v01);              // :1768
                        vsum += (1.f - a * a) * v1 + a * a * fmaxf(v0 - P.beta_t * v01, 0.f);
                        m = 0.f;
                    } else if (n0 > 0) {
                        const float v = v0 + fmaxf(0.f, v01 - s2);           // :867
                        a = __fdiv_rn(v, v + P.beta_t * sigma2);             // :870
                        vsum += (1.f - a * a) * v + a * a * sigma2;          // :875
                        m = Mg[u];
                    } else {
                        const float v = fmaxf(0.f, v1 - s2);                 // :890
                        a = __fdiv_rn(v, v + P.beta_x * sigma2);             // :893
                        vsum += a * v;                                       // :898
                        m = M1[u];
                    }
                    s_a[j] = a;
                    s_m[j] = m;
                }
            }
        }
        const float vp = (float)nagg * block_sum(vsum, s_red); // (has the barriers the restage needs)
        const float wgt = __fdiv_rn(1.f, fmaxf(vp, 1e-6f)); // :911
        if (P.dbg_vp && tid == 0) P.dbg_vp[g] = vp;

        // ---- pass 2: update, inverse transform and aggregation of the group ------------
        // resident: every candidate's transform is still in `tiles`, members are updated in
        // place; otherwise members are transformed again, a chunk at a time
        // smoother, not resident: by linearity T^-1((1-a) Y1 + a Y0) = x1 + T^-1(a T(x0 - x1)) -- one forward
        // transform per member and channel (of the difference of the two windows) instead of two
        const bool diff = P.smooth && !resident;
        for (int m0 = 0; m0 < nagg; m0 += cc) {
            const int cnt = min(cc, nagg - m0);
            if (!resident) {
                __syncthreads();
                // (tile index ml * tpc + c: the tiles of a second source stay unused)
                block_fwd_tiles<PSZ_T>(tiles, TS, cnt * tpc, psz, wrow, ch, [&](int t) -> const float * {
                    const int ml = t / tpc, rr = t - ml * tpc;
                    if (rr >= ch) return nullptr;
                    const uint32_t cd = s_cand[s_grp[m0 + ml]];
                    return (diff ? winP : winS) + (cand_y(cd) - y0) * wrow + (cand_x(cd) - x0) * ch + rr;
                }, diff ? (long)(winS - winP) : 0L);
            }
            // one thread per coefficient over the members of the chunk
#pragma unroll
            for (int u = 0; u < JPT; ++u) {
                const int j = tid + u * GF_THREADS;
                if (j >= cpp) continue;
                const int c = j / pp, e = j - c * pp;
                const float a = s_a[j], oma = 1.f - a, mj = s_m[j];
                float *yb = tiles + c * TS + e;
                for (int ml = 0; ml < cnt; ++ml) {
                    const int slot = resident ? s_grp[m0 + ml] : ml;
                    float *y = yb + slot * tpc * TS;
                    if (diff) *y = a * (*y);                                 // :1775, difference form
                    else if (P.smooth) *y = oma * (*y) + a * y[ch * TS];     // :1775
                    else *y = a * (*y) + oma * mj;                           // :878 / :901
                }
            }
            __syncthreads();
            block_inv_tiles<PSZ_T>(tiles, TS, cnt * ch, psz, [&](int t) {
                const int ml = t / ch, c = t - ml * ch;
                return (resident ? s_grp[m0 + ml] : ml) * tpc + c;
            });
            for (int it = tid; it < cnt * pp; it += GF_THREADS) {
                const int ml = it / pp, e = it - ml * pp;
                const int hy = e / psz, hx = e - hy * psz;
                const int gi = s_grp[m0 + ml];
                const int slot = resident ? gi : ml;
                const uint32_t cd = s_cand[gi];
                const long pix = (long)(cand_y(cd) + hy) * P.w + cand_x(cd) + hx;
                const float wW = __fmul_rn(wgt, W[e]);                  // :923
                const float *x1 = winS + (cand_y(cd) - y0 + hy) * wrow + (cand_x(cd) - x0 + hx) * ch;
                float v[MAX_CH];
                for (int c = 0; c < ch; ++c) {
                    float px_val = tiles[(slot * tpc + c) * TS + e];
                    if (diff) px_val += x1[c];
                    v[c] = __fmul_rn(wW, px_val);                       // :926
                }
                accumulate_pixel<CH_T>(P.accw + pix * (ch + 1), v, wW, ch);
            }
        }
    }
}

inline int launch_group_team8(const PassParams &P, int num_sms, cudaStream_t st);

inline int launch_group_filter(const PassParams &P, int num_sms, cudaStream_t st)
{
    // 8x8 patches with 1 or 3 channels: the team kernel of nlk_group_warp.cuh
    if (const int n = launch_group_team8(P, num_sms, st)) return n;
    const int TS = P.psz * P.psz + 1;
    const int cpp = P.ch * P.psz * P.psz;
    const int kcap = P.kstride > 1 ? P.kstride : 1;
    const int tpc = (P.has_prev ? 2 : 1) * P.ch;
    // tiles: every candidate of a group at once when that fits (one thread per tile and
    // about 64 KB), else chunks
    int CT = kcap * tpc;
    if (CT > GF_THREADS) CT = (GF_THREADS / tpc) * tpc;
    // tile area: 64 KB by default (two blocks per SM beside the windows at 12x12); NLK_GF_TILE_KB overrides it
    static const int tile_kb = getenv("NLK_GF_TILE_KB") ? atoi(getenv("NLK_GF_TILE_KB")) : 64;
    const int by_smem = ((tile_kb > 4 ? tile_kb : 4) * 1024 / (TS * 4)) / tpc * tpc;
    if (CT > by_smem) CT = by_smem;
    if (CT < tpc) CT = tpc;
    const int r = P.smooth ? P.r_t : (P.r_t > P.r_x ? P.r_t : P.r_x);
    const int wrow = (2 * r + P.psz) * P.ch + 1;
    const int win_floats = (2 * r + P.psz) * wrow;
    const size_t smem = ((size_t)CT * TS + (size_t)win_floats * (P.has_prev ? 2 : 1) + 2 * cpp + kcap +
                         (P.tagg > 1 ? P.tagg : 1) + 8 + P.psz * P.psz) * 4;
    static const int blocks_per_sm = getenv("NLK_GF_BLOCKS") ? atoi(getenv("NLK_GF_BLOCKS")) : 2;
    const int nb = num_sms * (blocks_per_sm > 0 ? blocks_per_sm : 2);
#define NLK_LAUNCH_GF(PS, CHN)                                                                        \
    do {                                                                                              \
        cudaFuncSetAttribute(k_group_filter<PS, CHN>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                             (int)smem);                                                              \
        k_group_filter<PS, CHN><<<nb, GF_THREADS, smem, st>>>(P, CT, kcap, wrow, win_floats);        \
    } while (0)
    if (smem > 220 * 1024) return -1;
    if (P.psz == 8 && P.ch == 3) NLK_LAUNCH_GF(8, 3);
    else if (P.psz == 8 && P.ch == 1) NLK_LAUNCH_GF(8, 1);
    else if (P.psz == 12 && P.ch == 3) NLK_LAUNCH_GF(12, 3);
    else if (P.psz == 12 && P.ch == 1) NLK_LAUNCH_GF(12, 1);
    else NLK_LAUNCH_GF(0, 0);
#undef NLK_LAUNCH_GF
    return 1;
}

} // namespace nlk
