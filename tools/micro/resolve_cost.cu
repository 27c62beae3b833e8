// Times the mask_resolve kernels of the library on synthetic bitmaps (C2 and 4K grids).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../bwd_nlkalman_b200/csrc -o resolve_cost resolve_cost.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#define NLK_RESOLVE_TIMING 1
#include "nlk_resolve.cuh"
using namespace nlk;
int main()
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int shapes[3][3] = {{479, 269, 2}, {959, 539, 2}, {959, 539, 1}};
    for (auto &sh : shapes) {
        PassParams P; memset(&P, 0, sizeof P);
        P.gw = sh[0]; P.gh = sh[1]; P.G = P.gw * P.gh; P.R = sh[2]; P.nbw = 1; P.gy0 = 0; P.gy1 = P.gh;
        const int side = 2 * P.R + 1;
        std::vector<unsigned> h(P.G);
        srand(1);
        for (auto &w : h) { unsigned v = 0; for (int bit = 0; bit < side * side; ++bit) if (rand() % 100 < 12) v |= 1u << bit; w = v | (1u << (P.R * side + P.R)); }
        unsigned *nbr, *pk; int *active, *cnt;
        cudaMalloc(&nbr, P.G * 4); cudaMemcpy(nbr, h.data(), P.G * 4, cudaMemcpyHostToDevice);
        cudaMalloc(&pk, resolve_pack_bytes(P.gw, P.gh, P.R));
        cudaMalloc(&active, P.G * 4); cudaMalloc(&cnt, 64);
        int one[4] = {0, 1, 0, 0}; cudaMemcpy(cnt, one, 16, cudaMemcpyHostToDevice);
        P.nbr = nbr; P.active = active; P.nactive = cnt; P.any_nbr = cnt + 1; P.work = cnt + 2;
        long long *dbg; cudaMalloc(&dbg, 64); cudaMemset(dbg, 0, 64); P.dbg_dist = (float *)dbg;
        float ms = 0; int n = 0;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(a);
            launch_resolve(P, pk, 0);
            cudaEventRecord(b); cudaEventSynchronize(b);
            cudaEventElapsedTime(&ms, a, b);
        }
        cudaMemcpy(&n, cnt, 4, cudaMemcpyDeviceToHost);
        long long hd[2]; cudaMemcpy(hd, dbg, 16, cudaMemcpyDeviceToHost);
        printf("[loop %.1f us, tail %.1f us @1.965GHz] ", hd[0] / 1965.0, hd[1] / 1965.0);
        const int steps = P.gh - 1 + (P.gw + P.R * (P.gh - 1) + 3) / 4;
        printf("grid %dx%d R=%d: blocked %.1f us (%d steps, %.0f ns/step), active %d of %d", P.gw, P.gh, P.R, ms * 1e3, steps,
               ms * 1e6 / steps, n, P.G);
        std::vector<int> act1(P.G); cudaMemcpy(act1.data(), active, P.G * 4, cudaMemcpyDeviceToHost);
        // the generic kernel for comparison
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(a);
            launch_resolve(P, nullptr, 0);
            cudaEventRecord(b); cudaEventSynchronize(b);
            cudaEventElapsedTime(&ms, a, b);
        }
        int n2 = 0; cudaMemcpy(&n2, cnt, 4, cudaMemcpyDeviceToHost);
        std::vector<int> act2(P.G); cudaMemcpy(act2.data(), active, P.G * 4, cudaMemcpyDeviceToHost);
        int same = n == n2; for (int k = 0; k < n && same; ++k) same = act1[k] == act2[k];
        printf(" | per-column kernel %.1f us, active %d  lists %s  %s\n", ms * 1e3, n2, same ? "IDENTICAL" : "DIFFER", cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
