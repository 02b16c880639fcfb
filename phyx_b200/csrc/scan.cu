// phyx_b200 — exclusive prefix sum over int32 (reduce / scan-of-partials / downsweep), used by the
// radix sort offsets, the sweep's load-balanced emission and the colour-major joint layout.
// Deterministic (no atomics, fixed association order).
#include "common.cuh"

namespace phyx
{

constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ int warp_inclusive(int v)
{
    int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// exclusive scan of one value per thread across the block; returns exclusive prefix, total in *total
__device__ __forceinline__ int block_exclusive(int v, int* total)
{
    __shared__ int warpSums[kScanThreads / 32];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = warp_inclusive(v);
    if (lane == 31) warpSums[warp] = inc;
    __syncthreads();
    if (warp == 0)
    {
        int w = lane < kScanThreads / 32 ? warpSums[lane] : 0;
        int winc = warp_inclusive(w);
        if (lane < kScanThreads / 32) warpSums[lane] = winc - w;
        if (lane == kScanThreads / 32 - 1) *total = winc;
    }
    __syncthreads();
    int res = inc - v + warpSums[warp];
    return res;
}

__global__ void k_scan_reduce(const int* __restrict__ in, int n, int* __restrict__ partial)
{
    __shared__ int total;
    int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    int s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
        if (base + k < n) s += in[base + k];
    block_exclusive(s, &total);
    if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

__global__ void k_scan_down(const int* __restrict__ in, int n, const int* __restrict__ partialScanned, int* __restrict__ out,
    int* __restrict__ totalOut)
{
    __shared__ int total;
    int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    int v[kScanItems];
    int s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
    {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    int ex = block_exclusive(s, &total) + (partialScanned ? partialScanned[blockIdx.x] : 0);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
    {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
    if (totalOut && blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanThreads - 1) *totalOut = ex;
}

static int scan_rec(phyx_b200_ctx* c, const int* in, int* out, int n, int* totalDevice, int* scratch, size_t scratchInts)
{
    int blocks = (n + kScanTile - 1) / kScanTile;
    if (blocks <= 1)
    {
        k_scan_down<<<1, kScanThreads, 0, c->stream>>>(in, n, nullptr, out, totalDevice);
        c->launches++;
        PHYX_CUDA(cudaGetLastError());
        return PHYX_B200_OK;
    }
    if (size_t(blocks) * 2 > scratchInts)
    {
        set_error("scan scratch too small");
        return PHYX_B200_ERR_STATE;
    }
    int* partial = scratch;
    int* partialScanned = scratch + blocks;
    k_scan_reduce<<<blocks, kScanThreads, 0, c->stream>>>(in, n, partial);
    c->launches++;
    PHYX_TRY(scan_rec(c, partial, partialScanned, blocks, nullptr, scratch + 2 * blocks, scratchInts - 2 * size_t(blocks)));
    k_scan_down<<<blocks, kScanThreads, 0, c->stream>>>(in, n, partialScanned, out, totalDevice);
    c->launches++;
    PHYX_CUDA(cudaGetLastError());
    return PHYX_B200_OK;
}

// in and out may alias.  If totalDevice is non-null it receives the grand total.
int exclusive_scan_i32(phyx_b200_ctx* c, const int* in, int* out, int n, int* totalDevice)
{
    if (n <= 0)
    {
        if (totalDevice) PHYX_CUDA(cudaMemsetAsync(totalDevice, 0, sizeof(int), c->stream));
        return PHYX_B200_OK;
    }
    size_t need = 0;
    for (int m = n; m > kScanTile;)
    {
        m = (m + kScanTile - 1) / kScanTile;
        need += 2 * size_t(m);
    }
    need += 16;
    PHYX_TRY(c->scanTmp.reserve(need * sizeof(int)));
    return scan_rec(c, in, out, n, totalDevice, c->scanTmp.as<int>(), need);
}

} // namespace phyx
