// group_filter, 8x8 patches: one TEAM of two warps per group, several teams per block.
//
// Same arithmetic as the block-per-group kernel of nlk_group.cuh (reference
// src/nlkalman.c:713-932 and :1600-1845), laid out for throughput:
//   * a team synchronises with its own named barrier (bar.sync id, 64), so the eight
//     teams of a block drift apart and the FMA-heavy transform phases of some overlap the
//     shared-memory / L2 phases of the others -- no block-wide barrier in the group loop;
//   * groups are handed out dynamically (one atomic ticket per group); the chain ticket ->
//     active[] -> hdr[] of a team's next group runs under its current one;
//   * the search windows are staged with cp.async (all rows in flight at once);
//   * transforms: lane = one 8x8 tile, whole tile in registers as 32 packed-fp32 pairs
//     (nlk_dct.cuh).  The group loop is a sequence of ROUNDS: statistics rounds transform
//     (candidate x source x channel) tiles out of the staged windows into the exchange
//     buffer; update rounds transform (member x channel) tiles, shrink them, transform back
//     and leave weighted pixels -- or, in the launches whose groups are small (template UPD),
//     lane = row / column of a tile;
//   * statistics: lane = one coefficient position e (all channels), Welford recurrences
//     over the candidates in sorted order, accumulators in registers during a round and
//     partly parked in shared memory between rounds;
//   * aggregation: lane = one pixel of the patch, one red.global.add.v4.f32 per member.
// The kernel is bound by its instruction footprint and its 128 registers before anything
// else: what a launch does not run is compiled out (UPD, BSIC), loops
// over candidates and members stay rolled.
// The smoother uses the linearity of the transform:
//   T^-1((1-a) Y1 + a Y0) = x1 + T^-1(a * T(x0 - x1)),  one tile per lane instead of two.
#pragma once
#include "nlk_common.cuh"
#include "nlk_dct.cuh"
#include "nlk_group.cuh"
#include <cuda.h>      // CUtensorMap (the encoder is fetched through the runtime, no libcuda link)
#include <stdio.h>
#include <stdlib.h>

namespace nlk {

constexpr int GW_TEAM = 64;          // threads per team
constexpr int GW_MAX_TEAMS = 8;      // teams per block (named barriers 1..8)
constexpr int GW_TS = 65;            // tile stride in the exchange buffer (odd: conflict-free)

struct GroupWarpGeom {
    int teams;        // teams per block
    int team_floats;  // shared-memory floats per team
    int win_floats;   // floats of the window area (one spatial window or two temporal ones)
    int wrow_t, wrow_x;   // row strides of the staged windows (temporal / spatial radius)
    int wh_t, wh_x;       // TMA staging: rows of the window boxes
    int wpoff_t;          // TMA staging: offset of the previous-frame window (128-byte aligned)
    int mbar_off;         // TMA staging: float offset of the team's mbarrier in its area
    int kcap;         // capacity of the candidate list
    int noisy_floats; // area for the members' noisy patches (second filtering), 0 = none
};

__device__ __forceinline__ void team_sync(int bar_id)
{
    asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "n"(GW_TEAM) : "memory");
}

__device__ __forceinline__ void cp_async4(float *smem_dst, const float *gsrc)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// ---- TMA staging of the search windows ------------------------------------------------------
// One cp.async.bulk.tensor per window instead of ~30 four-byte cp.async per lane: the image is
// a 2-D tensor (w*ch floats by h rows), the window a box at (px - r, py - r); what lies outside
// the image is filled with zeros and never read (candidates are clamped to the image).  The
// copy signals the team's mbarrier; every iteration of the group loop completes exactly one
// phase of it (groups that stage nothing arrive without bytes), so the parity to wait for is
// the iteration's.
struct TeamMaps { CUtensorMap src_t, prev_t, src_x; };

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
// (the barrier is named by its shared-memory address, re-read from the per-team scalars where
// it is used: one register less across the transform phases)
__device__ __forceinline__ void mbar_init(unsigned bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, int bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, int parity)
{
    asm volatile("{\n\t.reg .pred P1;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
                 "@P1 bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int x, int y, unsigned bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(bar), "r"(x), "r"(y)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// per-team scalars of the current group, kept in shared memory and re-read where they are
// used: the transform phases hold a whole tile in registers and everything that stays live
// across them costs a spill
enum { GN_AI = 0, GN_G, GN_NK, GN_NP0, GN_FLAGS, GN_PXY, GN_COUNT = 6 };
enum { GP_WOFF0 = 0, GP_WROW, GP_WPOFF, GP_K, GP_NP0, GP_NAGG, GP_NR1, GP_FLAGS, GP_G, GP_N0, GP_N0B, GP_MBAR, GP_COUNT = 12 };
constexpr int GPF_PREV = 1;

__device__ __forceinline__ int lds_par(const int *p)
{
    int v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
}

// The ticket chain of a team's NEXT group, one dependent global access per stage:
//   stage 0: ticket = atomicAdd(work)   stage 1: g = active[ticket]   stage 2: hdr[g]
// Lane 0 issues a stage at the start of a consumer phase (few live registers there) and stores
// the result in the next iteration's slot at the end of the phase, so each round trip runs
// under the team's own work instead of in front of it.
__device__ __forceinline__ int4 chain_issue(int stage, const int *slot, const PassParams &P)
{
    int4 v = make_int4(0, 0, 0, 0);
    if (stage == 0) {
        const int t = atomicAdd(P.work, 1);
        v.x = t < *P.nactive ? t : -1;           // -1: the list is exhausted
    } else if (stage == 1) {
        const int t = slot[GN_AI];
        if (t >= 0) v.x = P.active[t];
    } else {
        if (slot[GN_AI] >= 0) v = *reinterpret_cast<const int4 *>(P.hdr + slot[GN_G]);
    }
    return v;
}
__device__ __forceinline__ void chain_store(int stage, int *slot, const int4 v)
{
    if (stage == 0) slot[GN_AI] = v.x;
    else if (stage == 1) slot[GN_G] = v.x;
    else { slot[GN_NK] = v.x; slot[GN_NP0] = v.y; slot[GN_FLAGS] = v.z; slot[GN_PXY] = v.w; }
}

// UPD selects the code of the update rounds, so that a launch carries only the one it runs
// (the kernel is instruction-cache bound: 16 warps per SM spread over its phases):
//   1: lane = (member, channel) tile, any group size
//   2: lane = row / column of a tile -- for launches whose groups have at most 8 / CH members
//      (second filtering: one member)
// BSIC: the pass has a basic estimate (second filtering).
// TMA: the windows come by cp.async.bulk.tensor (needs the image row pitch to be a multiple of 16 bytes).
template <int CH, bool SMOOTH, int UPD, bool BSIC, bool TMA>
__global__ void __launch_bounds__(GW_TEAM * GW_MAX_TEAMS, 1)
k_group_team8(const PassParams P, const GroupWarpGeom Gm, const __grid_constant__ TeamMaps M)
{
    constexpr int PSZ = 8, TS = GW_TS;
    // gain table: per channel 32 records {a[2j][k], a[2j+1][k], ((1-a)*m)[2j][k], ((1-a)*m)[2j+1][k]}
    // at j*8+k -- the coefficient pairing of the packed transform (nlk_dct.cuh)
    constexpr int GS = 33;             // 33 records per channel: the channels of a warp read different banks
    constexpr int MC = GW_TEAM / CH;   // members per update round
    constexpr int CC1 = GW_TEAM / CH, CC2 = GW_TEAM / (2 * CH);   // candidates per statistics round (1 / 2 sources)
    extern __shared__ __align__(128) float smem_team[];   // 128-byte aligned: TMA destinations
    float *const smem = smem_team;
    const int team = threadIdx.x / GW_TEAM;
    const int l64 = threadIdx.x % GW_TEAM;       // lane within the team
    const int bar = 1 + team;

    float *const tiles = smem + (size_t)team * Gm.team_floats;             // [64][TS]
    float *const win = tiles + 64 * TS;                                    // window area
    float4 *const s_am = reinterpret_cast<float4 *>(win + Gm.win_floats);  // [CH][GS] gain a and (1-a)*mean (M0 or M1)
    uint32_t *const s_cand = reinterpret_cast<uint32_t *>(s_am + CH * GS); // [kcap]
    int *const s_grp = reinterpret_cast<int *>(s_cand + Gm.kcap);          // [kcap]
    float *const s_red = reinterpret_cast<float *>(s_grp + Gm.kcap);       // [2]
    int *const s_nxt = reinterpret_cast<int *>(s_red + 2);                 // [2][GN_COUNT] ticket and header of a group, by parity
    int *const s_par = s_nxt + 2 * GN_COUNT;                               // [GP_COUNT]
    float *const s_noisy = reinterpret_cast<float *>(s_par + GP_COUNT);    // [noisy_floats]

    if (TMA) {
        if (l64 == 0) {
            const unsigned mb = smem_u32(tiles + Gm.mbar_off);
            s_par[GP_MBAR] = (int)mb;
            mbar_init(mb, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        // (the barrier that opens the first iteration orders the initialisation before its use)
    }

    // Groups are handed out by an atomic ticket; the chain ticket -> active[] -> hdr[] of the next
    // group runs under the current one (chain_issue / chain_store), its result is read from the
    // slot of the iteration's parity after the barrier that opens the iteration.
    if (l64 == 0) {
        for (int st = 0; st < 3; ++st) chain_store(st, s_nxt, chain_issue(st, s_nxt, P));
    }

    for (int it = 0;; ++it) {
        team_sync(bar);
        const int *nx = s_nxt + (it & 1) * GN_COUNT;
        if (nx[GN_AI] < 0) break;
        int round = 0;
        do {
        {
            const int lane = threadIdx.x & 31, wg = l64 >> 5;
            const int g = nx[GN_G];
            GroupHdr hd;
            hd.nk = nx[GN_NK]; hd.np0 = nx[GN_NP0]; hd.flags = nx[GN_FLAGS];
            const int px = cand_x((uint32_t)nx[GN_PXY]), py = cand_y((uint32_t)nx[GN_PXY]);
            const int prev_p = hd.flags & HDR_PREV_P;
            int k = hd.nk;
            const int np0 = hd.np0;

            if (!SMOOTH && k == 0) {                         // filter, k <= 1 (:815-849, :857)
                if (TMA && l64 == 0) mbar_arrive((unsigned)lds_par(s_par + GP_MBAR));
                break;
            }

            if (SMOOTH && np0 == 0) {
                // reference :1795-1804: the filtered patch at p, weight 1/1e-6, mask untouched
                const float wgt = __fdiv_rn(1.f, 1e-6f);
                const long pix = (long)(py + (l64 >> 3)) * P.w + px + (l64 & 7);
                const float wW = __fmul_rn(wgt, c_win[PSZ][l64]);
                float v[CH];
#pragma unroll
                for (int c = 0; c < CH; ++c) v[c] = __fmul_rn(wW, P.in1[pix * CH + c]);
                accumulate_pixel<CH>(P.accw + pix * (CH + 1), v, wW, CH);
                if (P.dbg_vp && l64 == 0) P.dbg_vp[g] = 0.f;
                if (TMA && l64 == 0) mbar_arrive((unsigned)lds_par(s_par + GP_MBAR));
                break;
            }

            // ---- stage the search window(s) of this group (reference :637-639) -----------------
            const int r = SMOOTH ? P.r_t : (prev_p ? P.r_t : P.r_x);
            const int wrow = (r == P.r_t) ? Gm.wrow_t : Gm.wrow_x;
            int woff0, wpoff;      // window origin in image coordinates, offset of the previous-frame window
            if constexpr (TMA) {
                const bool boxt = (r == P.r_t);
                const int rb = boxt ? P.r_t : P.r_x;
                // the box starts at the window's first column rounded down to a multiple of four floats:
                // TMA wants the innermost coordinate 16-byte aligned (tools/micro/tma_probe.cu: an
                // unaligned one is an illegal instruction); rows and columns outside the image are
                // filled with zeros
                const int oxf = ((px - rb) * CH) & ~3, oy = py - rb;
                woff0 = oy * wrow + oxf;
                wpoff = Gm.wpoff_t;
                if (l64 == 0) {
                    const unsigned mb = (unsigned)lds_par(s_par + GP_MBAR);
                    fence_proxy_async();   // the window was read and written through the generic proxy
                    const int bytes = (boxt ? Gm.wh_t : Gm.wh_x) * wrow * 4;
                    mbar_expect_tx(mb, prev_p ? 2 * bytes : bytes);
                    tma_load_2d(win, boxt ? &M.src_t : &M.src_x, oxf, oy, mb);
                    if (prev_p) tma_load_2d(win + wpoff, &M.prev_t, oxf, oy, mb);
                }
            } else {
                const int x0 = max(px - r, 0), x1 = min(px + r, P.w - PSZ);
                const int y0 = max(py - r, 0), y1 = min(py + r, P.h - PSZ);
                const int wlen = (x1 - x0 + PSZ) * CH, wh = y1 - y0 + PSZ;
                woff0 = y0 * wrow + x0 * CH;
                wpoff = wh * wrow;
                // warp wg takes rows wg, wg+2, ...; every element is its own 4-byte cp.async
                float *winS = win, *winP = win + wh * wrow;
                const long g0 = ((long)(y0 + wg) * P.w + x0) * CH;
                const long gstep = 2L * P.w * CH;
#pragma unroll 1
                for (int j = lane; j < wlen; j += 32) {
                    const float *gs = P.src + g0 + j;
                    float *ds = winS + wg * wrow + j;
#pragma unroll 2
                    for (int row = wg; row < wh; row += 2, gs += gstep, ds += 2 * wrow) cp_async4(ds, gs);
                    if (prev_p) {
                        const float *gp = P.prev0 + g0 + j;
                        float *dp = winP + wg * wrow + j;
#pragma unroll 2
                        for (int row = wg; row < wh; row += 2, gp += gstep, dp += 2 * wrow) cp_async4(dp, gp);
                    }
                }
            }
            for (int i = l64; i < k; i += GW_TEAM) s_cand[i] = P.cand[(long)g * P.kstride + i];
            if (l64 == 0) {
                s_par[GP_WOFF0] = woff0;
                s_par[GP_WROW] = wrow;
                s_par[GP_WPOFF] = wpoff;
                s_par[GP_K] = k;
                s_par[GP_NP0] = np0;
                s_par[GP_NR1] = prev_p ? (k + CC2 - 1) / CC2 : (k + CC1 - 1) / CC1;
                s_par[GP_FLAGS] = prev_p ? GPF_PREV : 0;
                s_par[GP_G] = g;
            }
            if (TMA) mbar_wait((unsigned)lds_par(s_par + GP_MBAR), it & 1);
            else cp_async_wait_all();
            team_sync(bar);

            // group members: the first tagg candidates with a valid previous patch, or, when
            // there is none (filter only), the first tagg candidates (:779-793, :857, :1669, :1737).
            // Both warps compute the same list.
            int cnt = 0;
            for (int b0 = 0; b0 < k && cnt < P.tagg; b0 += 32) {
                const int i = b0 + lane;
                const uint32_t cd = i < k ? s_cand[i] : 0u;
                const int take = (i < k) && (np0 > 0 ? cand_prev(cd) : 1);
                const unsigned int bal = __ballot_sync(0xffffffffu, take);
                const int rank = cnt + __popc(bal & ((1u << lane) - 1u));
                if (take && rank < P.tagg) s_grp[rank] = i;
                cnt += __popc(bal);
            }
            if (l64 == 0) s_par[GP_NAGG] = min(cnt, P.tagg);   // read after the next barrier
        }

        // statistics, in registers across the statistics rounds.  The temporal filter only uses
        // the candidates with a valid previous patch (V0, V01, M0: reference :867-878); M1 / V1
        // over all candidates are needed by the spatial branch (np0 == 0, :890-901, where no
        // candidate has a previous patch) and by the smoother (:1768) -- so the filter keeps
        // either pair in the same registers:
        //   filter:   sA = M1 or M0V, sB = V1 or V0, sC = V01, sD = M0 (group mean)
        //   smoother: sA = M1, sB = V1, sC = V01, sD = M0V, sE = V0
        // Between the statistics rounds only sA, sB (and sE) stay in registers: sC, sD wait in
        // the two gain-table slots this lane will write after the last round, and the count
        // of previous-frame candidates in the per-team scalars (by round parity) -- the
        // transform phases in between need the registers.
        float sA[CH], sB[CH], sE[SMOOTH ? CH : 1];
#pragma unroll
        for (int u = 0; u < CH; ++u) sA[u] = sB[u] = 0.f;
#pragma unroll
        for (int u = 0; u < (SMOOTH ? CH : 1); ++u) sE[u] = 0.f;

        for (;; ++round) {
            if (round) team_sync(bar);   // the previous round's consumers are done with `tiles`
            const int nr1 = lds_par(s_par + GP_NR1);
            const bool stat = round < nr1;
            const int flags = lds_par(s_par + GP_FLAGS);
            const int np0 = lds_par(s_par + GP_NP0);
            const bool need1 = SMOOTH || np0 == 0;
            int first, cnt;
            if (stat) {
                const int k = lds_par(s_par + GP_K);
                const int cc = (flags & GPF_PREV) ? CC2 : CC1;
                first = round * cc;
                cnt = min(cc, k - first);
            } else {
                const int nagg = lds_par(s_par + GP_NAGG);
                first = (round - nr1) * MC;
                if (first >= nagg) break;
                cnt = min(MC, nagg - first);
            }

            // second filtering: the group holds the NOISY patches of its members (:784-785, :853).
            // They are fetched into their own area while the statistics run (issued in round 1,
            // awaited after the gains); without room, or with a single statistics round, they
            // replace the source patches in the window after the gains instead.
            const bool noisy = !SMOOTH && BSIC && nr1 >= 2 &&
                               (round == 0 ? false : lds_par(s_par + GP_NAGG) * (PSZ * PSZ * CH) <= Gm.noisy_floats);
            if (noisy && round == 1) {
                const int nagg = lds_par(s_par + GP_NAGG);
                for (int i = l64; i < nagg * PSZ * PSZ * CH; i += GW_TEAM) {
                    const int ml = i / (PSZ * PSZ * CH), rem = i - ml * (PSZ * PSZ * CH);
                    const int hy = rem / (PSZ * CH), j = rem - hy * (PSZ * CH);
                    const uint32_t cd = s_cand[s_grp[ml]];
                    cp_async4(s_noisy + i, P.in1 + ((long)(cand_y(cd) + hy) * P.w + cand_x(cd)) * CH + j);
                }
            }

            // ---- producer ------------------------------------------------------------------------
            {
                const int woff0 = lds_par(s_par + GP_WOFF0), wrow = lds_par(s_par + GP_WROW);
                const float *winS = win - woff0;                       // indexed by image coordinates
                const float *winP = winS + lds_par(s_par + GP_WPOFF);
                if (UPD == 2 && !stat) {
                    // few members (second filtering: one): a whole tile per lane would leave the team
                    // idle behind a handful of lanes, so lane = one row / column of a tile instead
                    const int tt = l64 >> 3, y = l64 & 7;
                    const bool act = tt < cnt * CH;
                    int off = 0, c = 0, ml = 0;
                    if (act) {
                        ml = tt / CH;
                        c = tt - ml * CH;
                        const uint32_t cd = s_cand[s_grp[first + ml]];
                        off = cand_y(cd) * wrow + cand_x(cd) * CH + c;
                    }
                    float *tb = tiles + tt * TS;
                    const float *wS = noisy ? s_noisy + (first + ml) * (PSZ * PSZ * CH) + y * (PSZ * CH) + c
                                            : winS + off + y * wrow;
                    const float *wP = winP + off + y * wrow;
                    if (act) {      // rows, forward
                        float r[8];
#pragma unroll
                        for (int x = 0; x < 8; ++x) r[x] = SMOOTH ? wP[x * CH] - wS[x * CH] : wS[x * CH];
                        dct1d_fwd<8>(r);
#pragma unroll
                        for (int x = 0; x < 8; ++x) tb[y * 8 + x] = r[x];
                    }
                    team_sync(bar);
                    if (act) {      // columns (y is the column index here): forward, shrink, inverse
                        float r[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) r[i] = tb[i * 8 + y];
                        dct1d_fwd<8>(r);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float *am = reinterpret_cast<const float *>(s_am + c * GS + (i >> 1) * 8 + y) + (i & 1);
                            if (SMOOTH) r[i] *= am[0];                                       // :1775
                            else r[i] = fmaf(am[0], r[i], am[2]);                            // :878 / :901
                        }
                        dct1d_inv<8>(r);
#pragma unroll
                        for (int i = 0; i < 8; ++i) tb[i * 8 + y] = r[i];
                    }
                    team_sync(bar);
                    if (act) {      // rows, inverse
                        float r[8];
#pragma unroll
                        for (int x = 0; x < 8; ++x) r[x] = tb[y * 8 + x];
                        dct1d_inv<8>(r);
#pragma unroll
                        for (int x = 0; x < 8; ++x) tb[y * 8 + x] = SMOOTH ? r[x] + wS[x * CH] : r[x];
                    }
                } else {
                    bool act;
                    int s = 0, c;
                    uint32_t cd;
                    if (stat) {
                        int slot, rr;
                        if (flags & GPF_PREV) { slot = l64 / (2 * CH); rr = l64 - slot * (2 * CH); }
                        else { slot = l64 / CH; rr = l64 - slot * CH; }
                        s = rr >= CH; c = rr - s * CH;
                        act = slot < cnt;
                        cd = s_cand[first + (act ? slot : 0)];
                        // no previous patch: no previous-frame tile; and its source tile only
                        // matters where M1 / V1 do
                        if (!cand_prev(cd) && (s == 1 || !need1)) act = false;
                    } else {
                        const int ml = l64 / CH;
                        c = l64 - ml * CH;
                        act = l64 < cnt * CH;
                        cd = s_cand[s_grp[first + (act ? ml : 0)]];
                    }
                    if (act) {
                        const int off = cand_y(cd) * wrow + cand_x(cd) * CH + c;
                        const float *wS = winS + off, *wP = winP + off;
                        int mrow = wrow;     // the member's patch for the update: window or noisy area
                        if (noisy && !stat) { wS = s_noisy + (first + l64 / CH) * (PSZ * PSZ * CH) + c; mrow = PSZ * CH; }
                        // the tile as 32 register pairs (row y, row 7-y), packed-fp32 transform
                        float *dst = tiles + l64 * TS;
                        if (stat) {
                            const float *src = s ? wP : wS;
                            f32x2 Pq[4][8];
#pragma unroll
                            for (int y = 0; y < 4; ++y)
#pragma unroll
                                for (int x = 0; x < 8; ++x)
                                    Pq[y][x] = pk2(src[y * wrow + x * CH], src[(7 - y) * wrow + x * CH]);
                            dct8x8_fwd_x2(Pq);
#pragma unroll
                            for (int j = 0; j < 4; ++j)
#pragma unroll
                                for (int kk = 0; kk < 8; ++kk) {
                                    float lo, hi;
                                    upk2(Pq[j][kk], lo, hi);
                                    dst[(2 * j) * 8 + kk] = lo;
                                    dst[(2 * j + 1) * 8 + kk] = hi;
                                }
                        } else if (UPD != 2) {
                            f32x2 Pq[4][8];
#pragma unroll
                            for (int y = 0; y < 4; ++y)
#pragma unroll
                                for (int x = 0; x < 8; ++x) {
                                    // smoother: x1 + T^-1(a * T(x0 - x1))                        (:1775)
                                    if (SMOOTH) Pq[y][x] = pk2(wP[y * wrow + x * CH] - wS[y * wrow + x * CH],
                                                               wP[(7 - y) * wrow + x * CH] - wS[(7 - y) * wrow + x * CH]);
                                    else Pq[y][x] = pk2(wS[y * mrow + x * CH], wS[(7 - y) * mrow + x * CH]);
                                }
                            dct8x8_shrink_x2<SMOOTH>(Pq, s_am + c * GS);                             // :878 / :901
#pragma unroll
                            for (int y = 0; y < 4; ++y)
#pragma unroll
                                for (int x = 0; x < 8; ++x) {
                                    float lo, hi;
                                    upk2(Pq[y][x], lo, hi);
                                    if (SMOOTH) { lo += wS[y * wrow + x * CH]; hi += wS[(7 - y) * wrow + x * CH]; }
                                    dst[y * 8 + x] = lo;
                                    dst[(7 - y) * 8 + x] = hi;
                                }
                        }
                    }
                }
            }
            team_sync(bar);

            // lane 0: one stage of the next group's ticket chain per consumer phase
            int *const nxn = s_nxt + ((it + 1) & 1) * GN_COUNT;
            int4 pf = make_int4(0, 0, 0, 0);
            if (l64 == 0 && round < 3) pf = chain_issue(round, nxn, P);

            if (!stat) {
                // ---- aggregation: lane = pixel of the patch, one member per iteration ----------
                const float vp = (float)lds_par(s_par + GP_NAGG) * (s_red[0] + s_red[1]);
                const float wgt = __fdiv_rn(1.f, fmaxf(vp, 1e-6f));         // :911
                const float wW = __fmul_rn(wgt, c_win[PSZ][l64]);           // :923
#pragma unroll 1
                for (int ml = 0; ml < cnt; ++ml) {
                    const uint32_t cd = s_cand[s_grp[first + ml]];
                    const long pix = (long)(cand_y(cd) + (l64 >> 3)) * P.w + cand_x(cd) + (l64 & 7);
                    float v[CH];
#pragma unroll
                    for (int c = 0; c < CH; ++c) v[c] = __fmul_rn(wW, tiles[(ml * CH + c) * TS + l64]); // :926
                    accumulate_pixel<CH>(P.accw + pix * (CH + 1), v, wW, CH);
                }
                if (l64 == 0 && round < 3) chain_store(round, nxn, pf);
                continue;
            }

            // ---- statistics: lane = coefficient position, candidates in sorted order ----------
            float *const park = reinterpret_cast<float *>(s_am + (l64 >> 4) * 8 + (l64 & 7)) + ((l64 >> 3) & 1);
            float sC[CH], sD[CH];
            int n0 = 0;
            if (round) {
#pragma unroll
                for (int u = 0; u < CH; ++u) { sC[u] = park[u * (4 * GS)]; sD[u] = park[u * (4 * GS) + 2]; }
                n0 = lds_par(s_par + GP_N0 + (round & 1));
            } else {
#pragma unroll
                for (int u = 0; u < CH; ++u) sC[u] = sD[u] = 0.f;
            }
            {
                const int cstride = ((flags & GPF_PREV) ? 2 * CH : CH) * TS;
                const float *tp = tiles + l64;             // source tile of slot 0, channel 0
#pragma unroll 1
                for (int i = 0; i < cnt; ++i, tp += cstride) {
                    const int hasq = cand_prev(s_cand[first + i]);   // (implies prev_p)
                    if (need1) {
                        const float in1v = c_inv[first + i + 1];
#pragma unroll
                        for (int u = 0; u < CH; ++u) {
                            const float p = tp[u * TS];
                            const float delta = p - sA[u];
                            sA[u] = fmaf(delta, in1v, sA[u]);             // :765
                            sB[u] = fmaf(delta, p - sA[u], sB[u]);        // :766
                        }
                    }
                    if (hasq) {
                        n0 += 1;
                        const float in0v = c_inv[n0];
                        const bool ing = n0 <= P.tagg;
#pragma unroll
                        for (int u = 0; u < CH; ++u) {
                            const float p = tp[u * TS];
                            const float q = tp[(CH + u) * TS];
                            float &Mp = SMOOTH ? sD[u] : sA[u];
                            float &V0 = SMOOTH ? sE[SMOOTH ? u : 0] : sB[u];
                            const float d0 = q - Mp;                  // :770-775 / :1654-1659
                            Mp = fmaf(d0, in0v, Mp);
                            V0 = fmaf(d0, q - Mp, V0);
                            const float t = q - p;
                            sC[u] = fmaf(t, t, sC[u]);                // :777-778
                            if (!SMOOTH && ing) sD[u] = fmaf(q - sD[u], in0v, sD[u]); // :783
                        }
                    }
                }
            }
            if (l64 == 0 && round < 3) chain_store(round, nxn, pf);
            if (round != nr1 - 1) {
#pragma unroll
                for (int u = 0; u < CH; ++u) { park[u * (4 * GS)] = sC[u]; park[u * (4 * GS) + 2] = sD[u]; }
                if (l64 == 0) s_par[GP_N0 + ((round + 1) & 1)] = n0;
                continue;
            }

            // ---- after the last statistics round: gains (:858-904, :1763-1777) ----------------
            if (noisy) {
                cp_async_wait_all();
            } else if (!SMOOTH && BSIC) {
                // the group holds the NOISY patches (:784-785, :853): the source window is no
                // longer needed, put the members' noisy patches where their source patches were
                const int woff0 = lds_par(s_par + GP_WOFF0), wrow = lds_par(s_par + GP_WROW);
                const int nagg = lds_par(s_par + GP_NAGG);
                for (int i = l64; i < nagg * PSZ * PSZ * CH; i += GW_TEAM) {
                    const int ml = i / (PSZ * PSZ * CH), rem = i - ml * (PSZ * PSZ * CH);
                    const int hy = rem / (PSZ * CH), j = rem - hy * (PSZ * CH);
                    const uint32_t cd = s_cand[s_grp[ml]];
                    const int qx = cand_x(cd), qy = cand_y(cd);
                    win[(qy + hy) * wrow + qx * CH + j - woff0] = P.in1[((long)(qy + hy) * P.w + qx) * CH + j];
                }
            }
            float vsum = 0.f;
            {
                const int n1 = lds_par(s_par + GP_K);
                const float inp1 = c_inv[max(n1, 1)];
                const float inp0 = c_inv[n0];
                const float sigma2 = P.sigma2;
                const float s2 = BSIC ? 0.f : sigma2;
#pragma unroll
                for (int u = 0; u < CH; ++u) {
                    float a, m;
                    if (SMOOTH) {
                        float v1 = sB[u], v0 = sE[SMOOTH ? u : 0], v01 = sC[u];
                        v1 *= inp1;                                 // :805
                        if (n0) { v0 *= inp0; v01 *= inp0; }        // :806-810
                        a = __fdividef(v1, v1 + P.beta_t * v01);              // :1768
                        vsum += (1.f - a * a) * v1 + a * a * fmaxf(v0 - P.beta_t * v01, 0.f);
                        m = 0.f;
                    } else if (n0 > 0) {
                        const float v0 = sB[u] * inp0, v01 = sC[u] * inp0;   // :806-810
                        const float v = v0 + fmaxf(0.f, v01 - s2);           // :867
                        a = __fdividef(v, v + P.beta_t * sigma2);             // :870
                        vsum += (1.f - a * a) * v + a * a * sigma2;          // :875
                        m = sD[u];
                    } else {
                        const float v1 = sB[u] * inp1;                       // :805
                        const float v = fmaxf(0.f, v1 - s2);                 // :890
                        a = __fdividef(v, v + P.beta_x * sigma2);             // :893
                        vsum += a * v;                                       // :898
                        m = sA[u];
                    }
                    park[u * (4 * GS)] = a;
                    park[u * (4 * GS) + 2] = (1.f - a) * m;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) vsum += __shfl_xor_sync(0xffffffffu, vsum, o);
            if ((l64 & 31) == 0) s_red[l64 >> 5] = vsum;
            if (P.dbg_vp) {
                team_sync(bar);
                if (l64 == 0) P.dbg_vp[lds_par(s_par + GP_G)] = (float)lds_par(s_par + GP_NAGG) * (s_red[0] + s_red[1]);
            }
            // the barrier at the top of the next round makes the gains (and restaged members) visible
        }
        } while (0);
        // stages of the ticket chain the group had no round for (early exits, two-round groups)
        if (l64 == 0) {
            int *const nxn = s_nxt + ((it + 1) & 1) * GN_COUNT;
            for (int st = min(round, 3); st < 3; ++st) chain_store(st, nxn, chain_issue(st, nxn, P));
        }
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library links cudart only)
typedef CUresult (*nlk_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                        const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);
inline nlk_encode_tiled_fn tensor_map_encoder()
{
    static nlk_encode_tiled_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<nlk_encode_tiled_fn>(f);
    }
    return fn;
}

// the image as a (w*ch) x h fp32 tensor, boxes of bw x bh elements, zero fill outside
inline bool window_map(CUtensorMap *m, const float *img, int w, int h, int ch, int bw, int bh)
{
    nlk_encode_tiled_fn enc = tensor_map_encoder();
    if (!enc || !img) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)w * ch, (cuuint64_t)h};
    const cuuint64_t gstr[1] = {(cuuint64_t)w * ch * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh};
    const cuuint32_t est[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(img), gdim, gstr, box, est,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// returns the number of launches, 0 if this kernel does not cover the configuration
inline int launch_group_team8(const PassParams &P, int num_sms, cudaStream_t st)
{
    if (P.psz != 8 || (P.ch != 3 && P.ch != 1)) return 0;
    // smoother with a basic estimate: statistics on bsic1, members from filt1 (reference
    // src/nlkalman.c:1669) -- the team kernel takes both from one window; the block-per-group
    // kernel restages the members.  No reference driver passes one (src/main-smo.c:209).
    if (P.smooth && P.has_bsic) return 0;
    const int ch = P.ch;
    // TMA staging is opt-in (NLK_TMA=1) until it is validated on the GPU; NLK_NO_TMA wins
    static const bool no_tma = getenv("NLK_NO_TMA") != nullptr || getenv("NLK_TMA") == nullptr;
    static const bool dbg = getenv("NLK_DEBUG") != nullptr;
    GroupWarpGeom Gm;
    TeamMaps M;
    memset(&M, 0, sizeof M);
    // TMA staging (3-channel launches): row pitch and base addresses multiples of 16 bytes, boxes
    // of at most 256 elements a side, their width a multiple of 4 floats
    // (+3: the box origin is rounded down to a multiple of four floats)
    const int bw_t = ((2 * P.r_t + 8) * ch + 3 + 3) & ~3, bw_x = ((2 * P.r_x + 8) * ch + 3 + 3) & ~3;
    const int bh_t = 2 * P.r_t + 8, bh_x = 2 * P.r_x + 8;
    bool tma = ch == 3 && !no_tma && ((size_t)P.w * ch * sizeof(float)) % 16 == 0 &&
               ((uintptr_t)P.src % 16) == 0 && ((uintptr_t)P.prev0 % 16) == 0 &&
               bw_t <= 256 && bw_x <= 256 && bh_t <= 256 && bh_x <= 256;
    if (tma) {
        tma = window_map(&M.src_t, P.src, P.w, P.h, ch, bw_t, bh_t);
        if (tma && P.has_prev) tma = window_map(&M.prev_t, P.prev0, P.w, P.h, ch, bw_t, bh_t);
        if (tma && !P.smooth) tma = window_map(&M.src_x, P.src, P.w, P.h, ch, bw_x, bh_x);
    }
    Gm.wh_t = bh_t;
    Gm.wh_x = bh_x;
    int win_t, win_x;
    if (tma) {
        Gm.wrow_t = bw_t;
        Gm.wrow_x = bw_x;
        Gm.wpoff_t = (bh_t * bw_t + 31) & ~31;                      // the second window starts 128-byte aligned
        win_t = P.has_prev ? Gm.wpoff_t + bh_t * bw_t : bh_t * bw_t;
        win_x = P.smooth ? 0 : bh_x * bw_x;
    } else {
        Gm.wrow_t = ((2 * P.r_t + 8) * ch) | 1;
        Gm.wrow_x = ((2 * P.r_x + 8) * ch) | 1;
        Gm.wpoff_t = 0;
        win_t = (2 * P.r_t + 8) * Gm.wrow_t * (P.has_prev ? 2 : 1);
        win_x = P.smooth ? 0 : (2 * P.r_x + 8) * Gm.wrow_x;
    }
    Gm.win_floats = ((win_t > win_x ? win_t : win_x) + 3) & ~3;   // the float4 gain table behind it stays aligned
    Gm.kcap = P.kstride > 1 ? P.kstride : 1;
    int fl = 64 * GW_TS + Gm.win_floats + ch * 4 * 33 + 2 * Gm.kcap + 2 + 2 * GN_COUNT + GP_COUNT;
    fl = (fl + 3) & ~3;
    const int budget = 227 * 1024;
    // second filtering: room for the noisy patches of up to three members per team, if that
    // does not cost a team
    Gm.noisy_floats = 0;
    if (P.has_bsic && !P.smooth) {
        const int want = (P.tagg < 3 ? P.tagg : 3) * 64 * ch;
        const int t0 = budget / (fl * 4), t1 = budget / ((fl + want + 34) * 4);
        if ((t1 >= GW_MAX_TEAMS || t1 == t0) && t1 >= 2) { Gm.noisy_floats = want; fl += want; }
    }
    // the team's mbarrier (8 bytes) at the end; team areas are multiples of 128 bytes (TMA destinations)
    fl = (fl + 1) & ~1;
    Gm.mbar_off = fl;
    fl = (fl + 2 + 31) & ~31;
    Gm.team_floats = fl;
    int teams = budget / (fl * 4);
    if (teams > GW_MAX_TEAMS) teams = GW_MAX_TEAMS;
    if (teams < 2) return 0;
    Gm.teams = teams;
    if (dbg)
        fprintf(stderr, "[nlk] group_team8: ch %d smooth %d bsic %d tma %d: %d floats/team (window %d, noisy %d) -> %d teams\n",
                ch, P.smooth, P.has_bsic, (int)tma, fl, Gm.win_floats, Gm.noisy_floats, teams);
    const size_t smem = (size_t)teams * fl * 4;
#define NLK_LAUNCH_TEAM(CHN, SM, UP, BS, TM)                                                          \
    do {                                                                                              \
        cudaFuncSetAttribute(k_group_team8<CHN, SM, UP, BS, TM>,                                      \
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                 \
        k_group_team8<CHN, SM, UP, BS, TM><<<num_sms, teams * GW_TEAM, smem, st>>>(P, Gm, M);         \
    } while (0)
#define NLK_LAUNCH_TEAM_CH(CHN, TM)                                                                   \
    do {                                                                                              \
        const bool few = P.tagg * CHN <= 8;                                                           \
        if (P.smooth) { if (few) NLK_LAUNCH_TEAM(CHN, true, 2, false, TM); else NLK_LAUNCH_TEAM(CHN, true, 1, false, TM); } \
        else if (P.has_bsic) { if (few) NLK_LAUNCH_TEAM(CHN, false, 2, true, TM); else NLK_LAUNCH_TEAM(CHN, false, 1, true, TM); } \
        else { if (few) NLK_LAUNCH_TEAM(CHN, false, 2, false, TM); else NLK_LAUNCH_TEAM(CHN, false, 1, false, TM); } \
    } while (0)
    if (ch == 3) { if (tma) NLK_LAUNCH_TEAM_CH(3, true); else NLK_LAUNCH_TEAM_CH(3, false); }
    else NLK_LAUNCH_TEAM_CH(1, false);
#undef NLK_LAUNCH_TEAM_CH
#undef NLK_LAUNCH_TEAM
    return 1;
}

} // namespace nlk
