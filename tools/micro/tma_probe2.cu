// the CUDA programming guide's TMA example (libcu++ API), as a known-good reference for tma_probe.cu
#include <cstdio>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
constexpr int BW = 56, BH = 18;
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, float *out, int x, int y)
{
    __shared__ alignas(128) float smem_buffer[BH][BW];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) {
        init(&bar, blockDim.x);
        cde::fence_proxy_async_shared_cta();
    }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
    } else {
        token = bar.arrive();
    }
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = (&smem_buffer[0][0])[i];
}
typedef CUresult (*enc_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                           const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char **argv)
{
    const int w = 128, h = 96, ch = 3;
    const int bw = argc > 1 ? atoi(argv[1]) : BW;   // box width given to the encoder (the kernel buffer stays 56 wide)
    float *img, *out;
    cudaMalloc(&img, (size_t)w * h * ch * 4);
    cudaMalloc(&out, BW * BH * 4);
    float *himg = new float[w * h * ch];
    for (int i = 0; i < w * h * ch; ++i) himg[i] = (float)i;
    cudaMemcpy(img, himg, (size_t)w * h * ch * 4, cudaMemcpyHostToDevice);
    void *f = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
    CUtensorMap m;
    cuuint64_t gdim[2] = {(cuuint64_t)w * ch, (cuuint64_t)h}, gstr[1] = {(cuuint64_t)w * ch * 4};
    cuuint32_t box[2] = {(cuuint32_t)bw, BH}, est[2] = {1, 1};
    CUresult r = ((enc_fn)f)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, img, gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d (box %d x %d)\n", (int)r, bw, BH);
    kernel<<<1, 128>>>(m, out, 32, 7);
    cudaError_t e = cudaDeviceSynchronize();
    float hout[BW * BH];
    cudaMemcpy(hout, out, sizeof hout, cudaMemcpyDeviceToHost);
    printf("%s  got %.0f want %.0f\n", cudaGetErrorString(e), hout[6 * bw + 20], (float)((7 + 6) * w * ch + 32 + 20));
    return e != cudaSuccess;
}
