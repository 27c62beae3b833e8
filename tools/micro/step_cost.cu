// Micro-benchmark: what does one barrier-separated step of a single-block wavefront cost?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o step_cost step_cost.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024) k_steps(const uint4 *__restrict__ g, unsigned *out, int nsteps, int stride)
{
    __shared__ unsigned s_pub[2][1024];
    const int i = threadIdx.x;
    s_pub[0][i] = s_pub[1][i] = 0;
    __syncthreads();
    unsigned wnd = i;
    uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0;
    const uint4 *row = g + (size_t)i * stride;
    if (MODE >= 2) { q0 = row[0]; q1 = row[1]; }
    for (int s = 0; s < nsteps; s += 2) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int ss = s + u;
            if (MODE >= 1) {
                const unsigned v = i ? s_pub[(ss + 1) & 1][i - 1] : 0u;
                wnd = (wnd | (v & 0xff)) >> 1;
                uint4 &q = u ? q1 : q0;
                if (MODE >= 2) {
                    wnd |= ((wnd & 1) ? 0u : (q.x & 3)) << 1;
                    wnd |= ((wnd & 2) ? 0u : (q.y & 3)) << 2;
                    wnd |= ((wnd & 4) ? 0u : (q.z & 3)) << 3;
                    wnd |= ((wnd & 8) ? 0u : (q.w & 3)) << 4;
                    if (ss + 2 < stride) q = row[ss + 2];
                }
                if (MODE >= 3 && (wnd & 0x10)) atomicOr(&s_pub[0][512 + (i & 511)], wnd);
                s_pub[ss & 1][i] = wnd;
            }
            __syncthreads();
        }
    }
    out[i] = wnd + q0.x + q1.y;
}

int main()
{
    const int nsteps = 2048, stride = nsteps + 8;
    uint4 *g; unsigned *out;
    cudaMalloc(&g, (size_t)1024 * stride * 16); cudaMemset(g, 1, (size_t)1024 * stride * 16);
    cudaMalloc(&out, 4096);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int nt : {32, 128, 288, 544, 1024}) {
        float ms[4];
        for (int mode = 0; mode < 4; ++mode) {
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(a);
                switch (mode) {
                case 0: k_steps<0><<<1, nt>>>(g, out, nsteps, stride); break;
                case 1: k_steps<1><<<1, nt>>>(g, out, nsteps, stride); break;
                case 2: k_steps<2><<<1, nt>>>(g, out, nsteps, stride); break;
                default: k_steps<3><<<1, nt>>>(g, out, nsteps, stride); break;
                }
                cudaEventRecord(b); cudaEventSynchronize(b);
                cudaEventElapsedTime(&ms[mode], a, b);
            }
        }
        printf("threads %4d: ns/step  barrier only %.0f | +smem chain %.0f | +global prefetch & select chain %.0f | +atomics %.0f\n",
               nt, ms[0] * 1e6 / nsteps, ms[1] * 1e6 / nsteps, ms[2] * 1e6 / nsteps, ms[3] * 1e6 / nsteps);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
