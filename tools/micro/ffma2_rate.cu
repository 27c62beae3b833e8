// Micro-benchmark: issue cost of Blackwell's packed fp32 instructions (FFMA2 / FADD2 / FMUL2,
// PTX *.f32x2) against scalar FFMA, alone and interleaved with ALU / shared-memory work.
// Question it answers: does one FFMA2 take one issue slot and two FMA-pipe cycles (so an
// issue-bound kernel gains by pairing), and does the constant (uniform-register) operand
// form run at the same rate?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 r, float &a, float &b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

__constant__ float2 c_k[8];

// MODE 0: 16 independent scalar FFMA chains (register operands)
// MODE 1: 8 independent FFMA2 chains (register operands)           -- same flops as mode 0
// MODE 2: 8 FFMA2 chains, multiplier from constant memory
// MODE 3: 16 scalar FFMA chains, multiplier from constant memory
// MODE 4: mode 0 + one LOP3 per FFMA (issue-slot competition)
// MODE 5: mode 1 + one LOP3 per *pair* of flops-equivalent (same ALU work as mode 4)
// MODE 6: mode 0 + one LDS per 4 FFMA
// MODE 7: mode 1 + one LDS per 2 FFMA2 (same LDS work as mode 6)
template <int MODE>
__global__ void __launch_bounds__(512) k_rate(float *out, int iters, float seed)
{
    __shared__ float sm[1024];
    sm[threadIdx.x] = seed; sm[threadIdx.x + 512] = seed;
    __syncthreads();
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + i + threadIdx.x;
    u64 p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = pk(a[2 * i], a[2 * i + 1]);
    const float m = seed * 0.5f;
    const u64 mm = pk(m, m);
    unsigned x = threadIdx.x, y = 0x9e3779b9u;
    float acc = 0.f;
    int li = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep) {
            if (MODE == 0 || MODE == 4 || MODE == 6) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    a[i] = fma1(a[i], m, a[(i + 1) & 15]);
                    if (MODE == 4) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(y), "r"(it)); }
                    if (MODE == 6 && (i & 3) == 0) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(&sm[(li + 33 * i) & 1023]))); acc += v; }
                }
            } else if (MODE == 3) {
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fma1(a[i], c_k[i & 7].x, a[(i + 1) & 15]);
            } else if (MODE == 1 || MODE == 5 || MODE == 7) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    p[i] = fma2(p[i], mm, p[(i + 1) & 7]);
                    if (MODE == 5) {
                        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(y), "r"(it));
                        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(y), "r"(it));
                    }
                    if (MODE == 7 && (i & 1) == 0) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(&sm[(li + 33 * i) & 1023]))); acc += v; }
                }
            } else if (MODE == 2) {
#pragma unroll
                for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], *reinterpret_cast<const u64 *>(&c_k[i]), p[(i + 1) & 7]);
            }
        }
    }
    float s = acc + (float)x;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) { float u, v; upk(p[i], u, v); s += u + v; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char *name, float *out, int sms)
{
    const int iters = 4096;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a);
        k_rate<MODE><<<sms * 2, 512>>>(out, iters, 1.0f);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    // fp32 FMA lane-operations: 128 per (it, rep) block of 16 scalar / 8 packed
    const double fmas = (double)sms * 2 * 512 * iters * 8 * 16;
    printf("%-44s %8.3f ms  %7.2f TFLOP/s\n", name, best, 2 * fmas / best * 1e-9);
}

int main()
{
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    float2 h[8]; for (int i = 0; i < 8; ++i) h[i] = make_float2(0.5f, 0.5f);
    cudaMemcpyToSymbol(c_k, h, sizeof h);
    float *out; cudaMalloc(&out, (size_t)pr.multiProcessorCount * 2 * 512 * 4);
    printf("%s, %d SMs\n", pr.name, pr.multiProcessorCount);
    run<0>("FFMA  reg (16 chains)", out, pr.multiProcessorCount);
    run<1>("FFMA2 reg (8 chains)", out, pr.multiProcessorCount);
    run<3>("FFMA  const multiplier", out, pr.multiProcessorCount);
    run<2>("FFMA2 const multiplier", out, pr.multiProcessorCount);
    run<4>("FFMA  + 1 LOP3 each", out, pr.multiProcessorCount);
    run<5>("FFMA2 + 2 LOP3 each (same ALU work)", out, pr.multiProcessorCount);
    run<6>("FFMA  + 1 LDS per 4", out, pr.multiProcessorCount);
    run<7>("FFMA2 + 1 LDS per 2 (same LDS work)", out, pr.multiProcessorCount);
    return 0;
}
