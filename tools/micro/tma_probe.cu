// Micro-test of the TMA window staging used by group_filter: one 2-D box load per team by one
// lane, mbarrier completion, tensor maps passed as a __grid_constant__ struct.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>

struct Maps { CUtensorMap a, b, c; };

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(128) k_probe(const __grid_constant__ Maps M, float *out, int x, int y, int sel, const CUtensorMap *gmap)
{
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned mb = smem_u32(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mb), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 64) {
        if (MODE & 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const CUtensorMap *mp = (MODE & 8) ? gmap : ((MODE & 2) ? (sel ? &M.a : &M.c) : &M.a);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mb), "r"(56 * 18 * 4) : "memory");
        if (MODE & 4)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         :: "r"(smem_u32(sm)), "l"(reinterpret_cast<unsigned long long>(mp)), "r"(mb), "r"(x), "r"(y) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         :: "r"(smem_u32(sm)), "l"(reinterpret_cast<unsigned long long>(mp)), "r"(mb), "r"(x), "r"(y) : "memory");
    }
    asm volatile("{\n\t.reg .pred P1;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 :: "r"(mb), "r"(0) : "memory");
    for (int i = threadIdx.x; i < 56 * 18; i += blockDim.x) out[i] = sm[i];
}

typedef CUresult (*enc_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                           const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv)
{
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    const int ax = argc > 3 ? atoi(argv[2]) : 30, ay = argc > 3 ? atoi(argv[3]) : 7;
    const int w = 128, h = 96, ch = 3;
    float *img, *out;
    cudaMalloc(&img, (size_t)w * h * ch * 4);
    cudaMalloc(&out, 56 * 18 * 4);
    float *himg = new float[w * h * ch];
    for (int i = 0; i < w * h * ch; ++i) himg[i] = (float)i;
    cudaMemcpy(img, himg, (size_t)w * h * ch * 4, cudaMemcpyHostToDevice);
    void *f = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
    enc_fn enc = (enc_fn)f;
    Maps M; memset(&M, 0, sizeof M);
    cuuint64_t gdim[2] = {(cuuint64_t)w * ch, (cuuint64_t)h}, gstr[1] = {(cuuint64_t)w * ch * 4};
    cuuint32_t box[2] = {56, 18}, est[2] = {1, 1};
    CUresult r = enc(&M.a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, img, gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    M.c = M.a;
    printf("encode %d\n", (int)r);
    float hout[56 * 18];
    CUtensorMap *gmap; cudaMalloc(&gmap, sizeof(CUtensorMap)); cudaMemcpy(gmap, &M.a, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
    for (int mode = 0; mode < 10; ++mode) {
        if (only >= 0 && mode != only) continue;
        for (int xy = 0; xy < 2; ++xy) {
            const int x = xy ? -ax : ax, y = xy ? -ay : ay;
            cudaMemset(out, 0, sizeof hout);
            switch (mode) {
            case 0: k_probe<0><<<1, 128, 56 * 18 * 4>>>(M, out, x, y, 1, gmap); break;
            case 1: k_probe<1><<<1, 128, 56 * 18 * 4>>>(M, out, x, y, 1, gmap); break;
            case 2: k_probe<2><<<1, 128, 56 * 18 * 4>>>(M, out, x, y, 1, gmap); break;
            case 3: k_probe<3><<<1, 128, 56 * 18 * 4>>>(M, out, x, y, 0, gmap); break;
            case 4: k_probe<4><<<1, 128, 56 * 18 * 4>>>(M, out, x, y, 1, gmap); break;
            case 5: k_probe<5><<<1, 128, 56 * 18 * 4>>>(M, out, x, y, 1, gmap); break;
            case 6: k_probe<6><<<1, 128, 56 * 18 * 4>>>(M, out, x, y, 1, gmap); break;
            case 7: k_probe<7><<<1, 128, 56 * 18 * 4>>>(M, out, x, y, 0, gmap); break;
            case 8: k_probe<8><<<1, 128, 56 * 18 * 4>>>(M, out, x, y, 0, gmap); break;
            default: k_probe<12><<<1, 128, 56 * 18 * 4>>>(M, out, x, y, 0, gmap); break;
            }
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(hout, out, sizeof hout, cudaMemcpyDeviceToHost);
            // element (row 6, col 20) of the box = image (y + 6, x + 20)
            const int iy = y + 6, ix = x + 20;
            const float want = (iy >= 0 && ix >= 0) ? (float)(iy * w * ch + ix) : 0.f;
            printf("mode %d at (%d,%d): %s  got %.0f want %.0f\n", mode, x, y, cudaGetErrorString(e), hout[6 * 56 + 20], want);
            if (e != cudaSuccess) return 1;
        }
    }
    return 0;
}
