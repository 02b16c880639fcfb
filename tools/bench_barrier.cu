// Micro-benchmark: cost of one grid-wide barrier (phyx_b200/csrc/barrier.cuh vs cooperative groups) on an otherwise idle grid.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -Iphyx_b200/csrc -o /tmp/bench_barrier tools/bench_barrier.cu && /tmp/bench_barrier
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
#include "barrier.cuh"
namespace cg = cooperative_groups;

__global__ void k_ours(unsigned long long* ring, int n)
{
    unsigned epoch = 0;
    for (int i = 0; i < n; ++i) phyx::grid_barrier(ring, epoch, false, false);
}
__global__ void k_cg(int n)
{
    cg::grid_group g = cg::this_grid();
    for (int i = 0; i < n; ++i) g.sync();
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned long long* ring;
    cudaMalloc(&ring, 64);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int n = 2000;
    for (int perSM = 1; perSM <= 5; ++perSM)
    {
        int grid = sms * perSM;
        float ms[2];
        for (int which = 0; which < 2; ++which)
        {
            for (int rep = 0; rep < 2; ++rep)
            {
                cudaMemset(ring, 0, 64);
                int nn = n;
                void* args0[] = { &ring, &nn };
                void* args1[] = { &nn };
                cudaEventRecord(a);
                if (which == 0) cudaLaunchCooperativeKernel((void*)k_ours, dim3(grid), dim3(256), args0, 0, 0);
                else cudaLaunchCooperativeKernel((void*)k_cg, dim3(grid), dim3(256), args1, 0, 0);
                cudaEventRecord(b);
                cudaEventSynchronize(b);
                cudaEventElapsedTime(&ms[which], a, b);
            }
        }
        printf("grid %4d (%d/SM): ours %.2f us/barrier, cooperative_groups %.2f us/barrier  (%s)\n", grid, perSM, ms[0] * 1e3 / n, ms[1] * 1e3 / n, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
