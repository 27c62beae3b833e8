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

// gather the patches of `cnt` candidates (list `cl`, entries first..first+cnt) from up
// to two images into tiles [(slot*2 + s)*ch + c][row*psz + hx]
template <int PSZ_T, int CH_T>
__device__ __forceinline__ void gather_tiles(float *__restrict__ tiles, int TS,
                                             const uint32_t *__restrict__ cl, int first, int cnt,
                                             const float *__restrict__ img0,
                                             const float *__restrict__ img1, bool img1_needs_prev,
                                             int w, int psz_rt, int ch_rt)
{
    const int psz = PSZ_T ? PSZ_T : psz_rt;
    const int ch = CH_T ? CH_T : ch_rt;
    const int rowlen = psz * ch;
    const int per_src = psz * rowlen;
    const int nsrc = img1 ? 2 : 1;
    const int total = cnt * nsrc * per_src;
    for (int it = threadIdx.x; it < total; it += GF_THREADS) {
        const int j = it % rowlen;
        int rest = it / rowlen;
        const int row = rest % psz;
        rest /= psz;
        const int s = rest % nsrc;
        const int slot = rest / nsrc;
        const uint32_t cd = cl[first + slot];
        if (s == 1 && img1_needs_prev && !cand_prev(cd)) continue;
        const float *img = s ? img1 : img0;
        const float v = img[((long)(cand_y(cd) + row) * w + cand_x(cd)) * ch + j];
        const int hx = j / ch, c = j - hx * ch;
        tiles[((slot * 2 + s) * ch + c) * TS + row * psz + hx] = v;
    }
}

template <int PSZ_T, int CH_T>
__global__ void __launch_bounds__(GF_THREADS, 2)
k_group_filter(const PassParams P, int cc, int kcap)
{
    constexpr int JPT = (PSZ_T && CH_T) ? (PSZ_T * PSZ_T * CH_T + GF_THREADS - 1) / GF_THREADS
                                        : (MAX_PSZ * MAX_PSZ * MAX_CH + GF_THREADS - 1) / GF_THREADS;
    const int psz = PSZ_T ? PSZ_T : P.psz;
    const int ch = CH_T ? CH_T : P.ch;
    const int pp = psz * psz, cpp = ch * pp;
    const int TS = pp + 1;
    const int tid = threadIdx.x;

    extern __shared__ __align__(16) float smem[];
    float *tiles = smem;                                  // [cc*2*ch][TS]
    float *s_a = tiles + (size_t)cc * 2 * ch * TS;        // [cpp] gain
    float *s_m = s_a + cpp;                               // [cpp] group mean (M0 or M1)
    uint32_t *s_cand = reinterpret_cast<uint32_t *>(s_m + cpp); // [kcap] sorted candidates
    uint32_t *s_grp = s_cand + kcap;                      // [tagg] group members (cand records)
    float *s_red = reinterpret_cast<float *>(s_grp + max(P.tagg, 1)); // [8]
    float *W = s_red + 8;                                 // [pp] aggregation window
    __shared__ int s_nagg;

    const int nactive = *P.nactive;
    const float sigma2 = P.sigma2;
    for (int e = tid; e < pp; e += GF_THREADS) W[e] = c_win[psz][e];

    for (int ai = blockIdx.x; ai < nactive; ai += gridDim.x) {
        const int g = P.active[ai];
        const GroupHdr hd = P.hdr[g];
        const int px = (g % P.gw) * P.step, py = (g / P.gw) * P.step;
        const int prev_p = hd.flags & HDR_PREV_P;
        int k = hd.nk;
        const int np0 = hd.np0;
        // smoother without search but with a valid previous patch: single-patch estimate
        // (reference :1699-1730; the group is the patch at p, see oracle/nlk_port.c)
        const bool point = P.smooth && k == 0 && prev_p;

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

        if (point) {
            if (tid == 0) s_cand[0] = cand_pack(px, py, 1);
            k = 1;
        } else {
            for (int i = tid; i < k; i += GF_THREADS) s_cand[i] = P.cand[(long)g * P.kstride + i];
        }
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
                if (take && rank < P.tagg) s_grp[rank] = cd;
                cnt += __popc(bal);
            }
            if (tid == 0) s_nagg = min(cnt, P.tagg);
        }

        // ---- pass 1: statistics over the k candidates ------------------------------------
        float M1[JPT], V1[JPT], Mp[JPT], V0[JPT], V01[JPT], Mg[JPT];
#pragma unroll
        for (int u = 0; u < JPT; ++u) M1[u] = V1[u] = Mp[u] = V0[u] = V01[u] = Mg[u] = 0.f;
        int n1 = 0, n0 = 0;
        for (int c0 = 0; c0 < k; c0 += cc) {
            const int cnt = min(cc, k - c0);
            __syncthreads();
            gather_tiles<PSZ_T, CH_T>(tiles, TS, s_cand, c0, cnt, P.src,
                                      prev_p ? P.prev0 : nullptr, true, P.w, psz, ch);
            __syncthreads();
            // one thread per tile
            for (int t = tid; t < cnt * 2 * ch; t += GF_THREADS) {
                const int slot = t / (2 * ch), s = (t / ch) & 1;
                if (s == 1 && !(prev_p && cand_prev(s_cand[c0 + slot]))) continue;
                dct2d_tile<PSZ_T, false>(tiles + t * TS, psz);
            }
            __syncthreads();
            // one thread per coefficient, candidates in sorted order
            for (int i = 0; i < cnt; ++i) {
                const int prev = prev_p && cand_prev(s_cand[c0 + i]);
                n1 += 1;
                n0 += prev;
                const float inp1 = c_inv[n1];
                const float inp0 = c_inv[n0];
#pragma unroll
                for (int u = 0; u < JPT; ++u) {
                    const int j = tid + u * GF_THREADS;
                    if (j < cpp) {
                        const int c = j / pp, e = j - c * pp;
                        const float p = tiles[((i * 2) * ch + c) * TS + e];
                        if (point) {
                            const float q = tiles[((i * 2 + 1) * ch + c) * TS + e];
                            V1[u] = p * p;
                            V0[u] = q * q;
                            V01[u] = (q - p) * (q - p);
                        } else {
                            const float delta = p - M1[u];
                            M1[u] += delta * inp1;              // :765
                            V1[u] += delta * (p - M1[u]);       // :766
                            if (prev) {
                                const float q = tiles[((i * 2 + 1) * ch + c) * TS + e];
                                const float d0 = q - Mp[u];     // :770-775 / :1654-1659
                                Mp[u] += d0 * inp0;
                                V0[u] += d0 * (q - Mp[u]);
                                const float t = q - p;
                                V01[u] += t * t;                // :777-778
                                if (n0 <= P.tagg) Mg[u] += (q - Mg[u]) * inp0; // :783
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
        const int nagg = s_nagg;

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
                    if (!point) {
                        v1 *= inp1;                             // :805
                        if (n0) { v0 *= inp0; v01 *= inp0; }    // :806-810
                    }
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
        const float vp = (float)nagg * block_sum(vsum, s_red);
        const float wgt = __fdiv_rn(1.f, fmaxf(vp, 1e-6f)); // :911
        if (P.dbg_vp && tid == 0) P.dbg_vp[g] = vp;

        // ---- pass 2: update, inverse transform and aggregation of the group ------------
        for (int n0g = 0; n0g < nagg; n0g += cc) {
            const int cnt = min(cc, nagg - n0g);
            __syncthreads();
            gather_tiles<PSZ_T, CH_T>(tiles, TS, s_grp, n0g, cnt, P.in1,
                                      P.smooth ? P.prev0 : nullptr, false, P.w, psz, ch);
            __syncthreads();
            const int nsrc = P.smooth ? 2 : 1;
            for (int t = tid; t < cnt * nsrc * ch; t += GF_THREADS) {
                const int slot = t / (nsrc * ch), r = t - slot * nsrc * ch;
                dct2d_tile<PSZ_T, false>(tiles + (slot * 2 * ch + r) * TS, psz);
            }
            __syncthreads();
            for (int it = tid; it < cnt * cpp; it += GF_THREADS) {
                const int slot = it / cpp, j = it - slot * cpp;
                const int c = j / pp, e = j - c * pp;
                float *y = tiles + ((slot * 2) * ch + c) * TS + e;
                const float a = s_a[j];
                if (P.smooth) *y = (1.f - a) * (*y) + a * y[ch * TS];   // :1775
                else *y = a * (*y) + (1.f - a) * s_m[j];                // :878 / :901
            }
            __syncthreads();
            for (int t = tid; t < cnt * ch; t += GF_THREADS) {
                const int slot = t / ch, c = t - slot * ch;
                dct2d_tile<PSZ_T, true>(tiles + ((slot * 2) * ch + c) * TS, psz);
            }
            __syncthreads();
            for (int it = tid; it < cnt * pp; it += GF_THREADS) {
                const int slot = it / pp, e = it - slot * pp;
                const int hy = e / psz, hx = e - hy * psz;
                const uint32_t cd = s_grp[n0g + slot];
                const long pix = (long)(cand_y(cd) + hy) * P.w + cand_x(cd) + hx;
                const float wW = __fmul_rn(wgt, W[e]);                  // :923
                float v[MAX_CH];
                for (int c = 0; c < ch; ++c)
                    v[c] = __fmul_rn(wW, tiles[((slot * 2) * ch + c) * TS + e]); // :926
                accumulate_pixel<CH_T>(P.accw + pix * (ch + 1), v, wW, ch);
            }
        }
    }
}

inline int group_filter_cc(const PassParams &P)
{
    const int TS = P.psz * P.psz + 1;
    int cc = GF_THREADS / (2 * P.ch);               // one thread per tile in the DCT phase
    const int by_smem = 64 * 1024 / (2 * P.ch * TS * 4);
    if (cc > by_smem) cc = by_smem;
    if (cc < 1) cc = 1;
    return cc;
}

inline int launch_group_filter(const PassParams &P, int num_sms, cudaStream_t st)
{
    const int cc = group_filter_cc(P);
    const int TS = P.psz * P.psz + 1;
    const int cpp = P.ch * P.psz * P.psz;
    const int kcap = P.kstride > 1 ? P.kstride : 1;
    const size_t smem = ((size_t)cc * 2 * P.ch * TS + 2 * cpp + kcap + (P.tagg > 1 ? P.tagg : 1) + 8 + P.psz * P.psz) * 4;
    const int nb = num_sms * 2;
#define NLK_LAUNCH_GF(PS, CHN)                                                                        \
    do {                                                                                              \
        cudaFuncSetAttribute(k_group_filter<PS, CHN>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                             (int)smem);                                                              \
        k_group_filter<PS, CHN><<<nb, GF_THREADS, smem, st>>>(P, cc, kcap);                          \
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
