// Which part of a barrier-separated wavefront step is expensive?  (single block, B200)
#include <cstdio>
#include <cuda_runtime.h>
// MODE bits: 1 select chain on register data, 2 coalesced LDG.128 prefetch (4 steps ahead), 4 uncoalesced instead,
//            8 record nibble + STS every 8 steps, 16 band skipping (only ~1/3 of the warps work in a step)
template <int MODE>
__global__ void __launch_bounds__(1024) k_steps(const uint4 *__restrict__ g, unsigned *out, int nsteps, int stride)
{
    __shared__ unsigned s_pub[2][1024];
    __shared__ unsigned s_rec[1024 * 4];
    const int i = threadIdx.x, nthr = blockDim.x, warp = i >> 5;
    s_pub[0][i] = s_pub[1][i] = 0;
    __syncthreads();
    unsigned wnd = i, acc = 0;
    uint4 q[4];
    const uint4 *p = (MODE & 4) ? g + (size_t)i * stride : g + i;
    const size_t inc = (MODE & 4) ? 1 : nthr;
    for (int u = 0; u < 4; ++u) q[u] = (MODE & 6) ? p[u * inc] : make_uint4(i, i + 1, i + 2, i + 3);
    const int w0 = warp * 48, w1 = w0 + nsteps / 3;
    for (int s = 0; s < nsteps; s += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int ss = s + u;
            if (!(MODE & 16) || (ss >= w0 && ss <= w1)) {
                const unsigned v = s_pub[(ss + 1) & 1][i ? i - 1 : 0];
                wnd |= (v >> 8) & 0xff;
                uint4 x = q[u];
                if (MODE & 1) {
                    wnd |= x.x & 0x3e & ~(unsigned)((int)(wnd << 31) >> 31);
                    wnd |= x.y & 0x3e & ~(unsigned)((int)(wnd << 30) >> 31);
                    wnd |= x.z & 0x3e & ~(unsigned)((int)(wnd << 29) >> 31);
                    wnd |= x.w & 0x3e & ~(unsigned)((int)(wnd << 28) >> 31);
                }
                s_pub[ss & 1][i] = (x.x & ~(unsigned)((int)(wnd << 31) >> 31)) | (x.y & ~(unsigned)((int)(wnd << 30) >> 31)) |
                                   (x.z & ~(unsigned)((int)(wnd << 29) >> 31)) | (x.w & ~(unsigned)((int)(wnd << 28) >> 31));
                if (MODE & 8) {
                    acc |= (~wnd & 0xf) << (4 * (ss & 7));
                    if ((ss & 7) == 7) { s_rec[i * 4 + ((ss >> 3) & 3)] = acc; acc = 0; }
                }
                wnd >>= 4;
                if ((MODE & 6) && ss + 4 < nsteps) q[u] = p[(size_t)(ss + 4) * inc];
            }
            __syncthreads();
        }
    }
    out[i] = wnd + q[0].x + q[1].y + q[2].z + q[3].w + acc + s_rec[i];
}

int main()
{
    const int nsteps = 2048, stride = nsteps + 8;
    uint4 *g; unsigned *out;
    cudaMalloc(&g, (size_t)1024 * stride * 16); cudaMemset(g, 1, (size_t)1024 * stride * 16);
    cudaMalloc(&out, 4096);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int nt : {288, 544}) {
        printf("threads %d  ns/step:", nt);
#define RUN(M) do { float ms = 0; for (int rep = 0; rep < 3; ++rep) { cudaEventRecord(a); k_steps<M><<<1, nt>>>(g, out, nsteps, stride); \
        cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b); } printf("  [%d] %.0f", M, ms * 1e6 / nsteps); } while (0)
        RUN(0); RUN(1); RUN(3); RUN(5); RUN(9); RUN(11); RUN(27); RUN(17); RUN(19);
        printf("\n");
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
