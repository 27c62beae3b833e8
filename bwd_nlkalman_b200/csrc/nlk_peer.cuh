// Peer-memory exchanges of the strip-sharded pass: one GPU per strip, every
// rank's exchange buffers in one "slab" at the same offsets, the peers' slabs mapped into this
// process (CUDA IPC) and reached over NVLink / NVSwitch with plain stores and reductions:
//
//   push       copy a byte range of the own slab to the same offset of the peers' slabs
//              (neighbour bitmaps after the search, halo rows of an output frame);
//   push_add   red.global.add the accumulator rows that this strip's groups wrote beyond the
//              strip border into the OWNER's accumulator (the overlap-add of the aggregation,
//              reference src/nlkalman.c:913-928, without a staging copy or a separate add);
//   signal / wait   flags[slot][source rank] in the receiver's slab, monotonically increasing
//              sequence numbers: the last block of a push stores the flag (release, system
//              scope) after its data, a one-warp kernel on the consumer's stream spins on it
//              (acquire).  No host thread ever blocks; there is no collective library call in
//              the data path.
//
// A wait gives up after a timeout and raises the slab's error word instead of hanging the GPU.
#pragma once
#include "nlk_common.cuh"
#include "nlk_prep.cuh"   // cubic1

namespace nlk {

constexpr int PEER_MAX = 16;      // ranks
constexpr int PEER_SLOTS = 64;    // flag slots per rank
// slab header: flags[PEER_SLOTS][PEER_MAX], then the error word and the block counters of the pushes
constexpr size_t PEER_FLAGS_BYTES = (size_t)PEER_SLOTS * PEER_MAX * 4;
constexpr size_t PEER_ERR_OFF = PEER_FLAGS_BYTES;            // unsigned int
constexpr size_t PEER_CNT_OFF = PEER_FLAGS_BYTES + 64;       // unsigned int[16]
constexpr size_t PEER_HDR_BYTES = PEER_FLAGS_BYTES + 256;

struct PeerTable {
    char *slab[PEER_MAX];   // base of every rank's slab as mapped here (own included)
    int rank, nranks;
};

__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void peer_signal_all(const PeerTable &T, int slot, unsigned int value, unsigned int mask)
{
    for (int p = 0; p < T.nranks; ++p)
        if ((mask >> p) & 1u)
            st_release_sys(reinterpret_cast<unsigned int *>(T.slab[p]) + slot * PEER_MAX + T.rank, value);
}

// the block that finishes last publishes the flag: every block fences its stores (system scope)
// before it counts itself done
__device__ __forceinline__ void peer_last_block_signal(const PeerTable &T, int slot, unsigned int value,
                                                       unsigned int mask, int cnt_idx)
{
    if (slot < 0) return;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int *cnt = reinterpret_cast<unsigned int *>(T.slab[T.rank] + PEER_CNT_OFF) + cnt_idx;
        if (atomicAdd(cnt, 1u) == gridDim.x - 1) {
            *cnt = 0u;
            __threadfence_system();
            peer_signal_all(T, slot, value, mask);
        }
    }
}

__global__ void k_peer_signal(const PeerTable T, int slot, unsigned int value, unsigned int mask)
{
    if (threadIdx.x == 0) {
        __threadfence_system();
        peer_signal_all(T, slot, value, mask);
    }
}

// lane p waits for rank p's flag
__global__ void k_peer_wait(const PeerTable T, int slot, unsigned int value, unsigned int mask,
                            unsigned long long timeout_ns)
{
    const int p = threadIdx.x;
    if (p >= T.nranks || !((mask >> p) & 1u)) return;
    const unsigned int *f = reinterpret_cast<const unsigned int *>(T.slab[T.rank]) + slot * PEER_MAX + p;
    const unsigned long long t0 = global_ns();
    int spins = 0;
    while ((int)(ld_acquire_sys(f) - value) < 0) {
        if ((++spins & 1023) == 0 && global_ns() - t0 > timeout_ns) {
            atomicExch(reinterpret_cast<unsigned int *>(T.slab[T.rank] + PEER_ERR_OFF), 0x10000u | (slot << 8) | p);
            return;
        }
    }
}

// copy [off, off + n * sizeof(V)) of the own slab to the peers in `mask`, then signal them
template <class V>
__global__ void __launch_bounds__(256) k_peer_push(const PeerTable T, size_t off, size_t n, unsigned int mask,
                                                   int slot, unsigned int value, int cnt_idx)
{
    const V *src = reinterpret_cast<const V *>(T.slab[T.rank] + off);
    const size_t i0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (int p = 0; p < T.nranks; ++p) {
        if (!((mask >> p) & 1u) || p == T.rank) continue;
        V *dst = reinterpret_cast<V *>(T.slab[p] + off);
        for (size_t i = i0; i < n; i += stride) dst[i] = src[i];
    }
    peer_last_block_signal(T, slot, value, mask, cnt_idx);
}

// accumulate n floats at `off` of the own slab into the same offset of peer `p`, then signal it
template <int VEC>
__global__ void __launch_bounds__(256) k_peer_push_add(const PeerTable T, size_t off, size_t n, int p,
                                                       int slot, unsigned int value, int cnt_idx)
{
    const float *src = reinterpret_cast<const float *>(T.slab[T.rank] + off);
    float *dst = reinterpret_cast<float *>(T.slab[p] + off);
    const size_t i0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    if (VEC == 4) {
        for (size_t i = i0; i < n / 4; i += stride) {
            const float4 v = reinterpret_cast<const float4 *>(src)[i];
            // rows nobody aggregated into stay untouched (most of a halo row is zeros only at the far end)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                         :: "l"(dst + 4 * i), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        }
    } else {
        for (size_t i = i0; i < n; i += stride) atomicAdd(dst + i, src[i]);
    }
    peer_last_block_signal(T, slot, value, 1u << p, cnt_idx);
}

// ---- warp with remote rows ----------------------------------------------------------------------
// The bicubic warp of a strip (reference src/nlkalman.c:66-88) reads the previous output at
// flow-displaced positions: almost always inside the rows the rank holds (its own and the halo rows its
// neighbours pushed), occasionally anywhere.  Instead of every rank pushing its whole strip to everybody
// after every pass, a tap row outside [lo, hi) is loaded straight from its owner's slab over NVLink
// (rows [k * chunk_y, (k+1) * chunk_y) belong to rank k, the rest to the last rank); the frame sits at the
// same offset in every slab.  Same arithmetic as k_warp.
template <int CH>
__global__ void k_warp_peer(float *__restrict__ imw, const PeerTable T, const char *__restrict__ local_base, size_t frame_off,
                            const float *__restrict__ of, const float *__restrict__ msk, int w, int h, int ch_rt,
                            int row0, int row1, int lo, int hi, int chunk_y)
{
    const int ch = CH ? CH : ch_rt;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = row0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= row1) return;
    const long pix = (long)y * w + x;
    float *o = imw + pix * ch;
    const float nanv = __int_as_float(0x7fc00000);
    if (msk != nullptr && msk[pix] != 0.f) {
        for (int c = 0; c < ch; ++c) o[c] = nanv;
        return;
    }
    float xw = __fadd_rn((float)x, of[pix * 2 + 0]);
    float yw = __fadd_rn((float)y, of[pix * 2 + 1]);
    xw = __fsub_rn(xw, 1.f);
    yw = __fsub_rn(yw, 1.f);
    const int ix = (int)floorf(xw), iy = (int)floorf(yw);
    const float fx = __fsub_rn(xw, (float)ix), fy = __fsub_rn(yw, (float)iy);
    const bool inside = (ix >= 0) && (ix + 3 < w) && (iy >= 0) && (iy + 3 < h) && (xw == xw) && (yw == yw);
    if (!inside) {
        for (int c = 0; c < ch; ++c) o[c] = nanv;
        return;
    }
    // (kernel parameters are only ever indexed with compile-time constants: a run-time index into the
    // table would make the compiler keep a copy of it in local memory)
    const float *rowp[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int yy = iy + k;
        const char *base = local_base;
        if (yy < lo || yy >= hi) {
            int owner = yy / chunk_y;
            if (owner > T.nranks - 1) owner = T.nranks - 1;
#pragma unroll
            for (int q = 0; q < PEER_MAX; ++q) if (q == owner) base = T.slab[q];
        }
        rowp[k] = reinterpret_cast<const float *>(base + frame_off) + ((long)yy * w + ix) * ch;
    }
    for (int c = 0; c < ch; ++c) {
        float col[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
            col[i] = cubic1(rowp[0][i * ch + c], rowp[1][i * ch + c], rowp[2][i * ch + c], rowp[3][i * ch + c], fy);
        o[c] = cubic1(col[0], col[1], col[2], col[3], fx);
    }
}

} // namespace nlk
