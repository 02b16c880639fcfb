// phyx_b200 — exclusive prefix sum over int32 (single pass, decoupled look-back), used by the radix sort offsets, the
// sweep's load-balanced emission, the swap-with-last compactions and the schedule layouts.
// Deterministic: integer sums, so the result does not depend on the order tiles finish in.
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

// Single-pass scan with decoupled look-back: one launch instead of reduce / scan-of-partials / downsweep (a step makes
// about fourteen scans, which used to be 40 launches).  Tiles are taken in ticket order, so a tile only ever waits for
// tiles whose CTAs are already running.  status[t] = flag << 32 | value: flag 1 = the tile's own total is there,
// 2 = the inclusive prefix up to and including the tile is.  Integer sums: the result does not depend on who adds what.
// The status words carry the scan's EPOCH (a counter in device memory, advanced by every scan) in their upper 30 bits, so
// words left by earlier scans read as "not there yet" and nothing has to be cleared between scans; the CTA that finishes
// last puts the ticket counter back to zero and advances the epoch.  (A memset per scan was a quarter of a small world's
// step: fourteen scans, each a few microseconds.  The epoch lives on the device so that a scan's launch parameters do not
// change from step to step: the deferred step is replayed as a CUDA graph.)
__device__ __forceinline__ void scan_leave(unsigned* ticket, unsigned epoch)
{
    // ticket[0] = next tile, ticket[1] = CTAs done, ticket[2] = epoch
    if (threadIdx.x == 0 && atomicAdd(&ticket[1], 1u) == gridDim.x - 1)
    {
        ticket[0] = 0u;
        ticket[1] = 0u;
        const unsigned next = (epoch + 1u) & 0x3fffffffu;
        ticket[2] = next ? next : 1u;   // (0 = "never written"; a word 2^30 scans old could be mistaken for a fresh one: the
                                        // status area of one scan is rewritten entirely by the next scan of the same size or larger,
                                        // and a step makes the same scans in the same order)
    }
}

__global__ void __launch_bounds__(kScanThreads) k_scan_single(const int* __restrict__ in, Count nc, int* __restrict__ out, unsigned long long* __restrict__ status,
    unsigned* __restrict__ ticket, int* __restrict__ totalOut)
{
    __shared__ int total;
    __shared__ unsigned s_tile;
    __shared__ int s_prefix;
    const int n = count_of(nc);
    const unsigned epoch = __ldcg(&ticket[2]);   // (only the last CTA of a scan changes it, after every CTA has read it... see scan_leave)
    const unsigned long long tag = static_cast<unsigned long long>(epoch) << 34;
    if (threadIdx.x == 0) s_tile = atomicAdd(&ticket[0], 1u);
    __syncthreads();
    const int tile = int(s_tile);
    // (a length read from the device may be shorter than the grid was sized for: tiles past the end have no successors that matter)
    if (tile * kScanTile >= n)
    {
        if (tile == 0 && threadIdx.x == 0 && totalOut) *totalOut = 0;
        scan_leave(ticket, epoch);
        return;
    }
    const int base = tile * kScanTile + threadIdx.x * kScanItems;
    int v[kScanItems];
    int s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
    {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    const int ex = block_exclusive(s, &total);
    const unsigned mine = unsigned(total);
    if (threadIdx.x == 0)
    {
        volatile unsigned long long* st = status;
        st[tile] = tag | ((tile == 0 ? 2ull : 1ull) << 32) | mine;
    }
    if (threadIdx.x < 32)
    {
        // look back, 32 predecessors at a time
        volatile unsigned long long* st = status;
        unsigned prefix = 0;
        int t = tile - 1;
        while (t >= 0)
        {
            const int idx = t - int(threadIdx.x);
            unsigned long long w = 0;
            if (idx >= 0)
            {
                do { w = st[idx]; } while ((w >> 34) != epoch || ((w >> 32) & 3ull) == 0ull);
            }
            const unsigned flag = idx >= 0 ? unsigned(w >> 32) & 3u : 0u;
            const unsigned full = __ballot_sync(0xffffffffu, flag == 2u);   // lanes that saw a complete prefix
            // add everything up to and including the nearest complete prefix
            const int stop = full ? __ffs(full) - 1 : 31;
            unsigned part = (idx >= 0 && int(threadIdx.x) <= stop) ? unsigned(w) : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            prefix += part;
            if (full) break;
            t -= 32;
        }
        if (threadIdx.x == 0)
        {
            if (tile > 0) st[tile] = tag | (2ull << 32) | unsigned(prefix + mine);
            s_prefix = int(prefix);
        }
    }
    __syncthreads();
    int run = ex + s_prefix;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
    {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
    if (totalOut && base <= n - 1 && n - 1 < base + kScanItems) *totalOut = run;   // the thread that holds the last element
    scan_leave(ticket, epoch);
}

// in and out may alias.  If totalDevice is non-null it receives the grand total.
int exclusive_scan_i32(phyx_b200_ctx* c, const int* in, int* out, int n, int* totalDevice)
{
    Count nc;
    nc.v = n;
    nc.p = nullptr;
    nc.mul = 1;
    return exclusive_scan_count(c, in, out, nc, totalDevice);
}

// the same with the length as a Count (deferred step: n.v is the bound, the device word the length)
int exclusive_scan_count(phyx_b200_ctx* c, const int* in, int* out, Count nc, int* totalDevice)
{
    const int n = nc.v;
    if (n <= 0)
    {
        if (totalDevice) PHYX_CUDA(cudaMemsetAsync(totalDevice, 0, sizeof(int), c->stream));
        return PHYX_B200_OK;
    }
    const int tiles = (n + kScanTile - 1) / kScanTile;
    const size_t bytes = (size_t(tiles) + 4) * sizeof(unsigned long long);
    PHYX_TRY(c->scanTmp.reserve(bytes));
    if (c->scanTmp.ptr != c->scanTmpCleared)
    {
        // a new buffer: clean words, epoch 1 (0 is what cleared words carry)
        PHYX_CUDA(cudaMemsetAsync(c->scanTmp.ptr, 0, c->scanTmp.cap, c->stream));
        const unsigned one = 1u;
        PHYX_CUDA(cudaMemcpyAsync(c->scanTmp.as<unsigned>() + 2, &one, sizeof(one), cudaMemcpyHostToDevice, c->stream));
        PHYX_CUDA(cudaStreamSynchronize(c->stream));   // (`one` is on the stack; once per buffer)
        c->scanTmpCleared = c->scanTmp.ptr;
    }
    unsigned long long* status = c->scanTmp.as<unsigned long long>() + 2;
    unsigned* ticket = c->scanTmp.as<unsigned>();
    k_scan_single<<<tiles, kScanThreads, 0, c->stream>>>(in, nc, out, status, ticket, totalDevice);
    c->launches++;
    PHYX_CUDA(cudaGetLastError());
    return PHYX_B200_OK;
}

} // namespace phyx
