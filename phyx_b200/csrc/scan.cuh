// phyx_b200 — exclusive prefix sum over int32 (single pass, decoupled look-back), used by the radix sort offsets, the
// sweep's load-balanced emission, the swap-with-last compactions and the schedule layouts.  Header: the kernel is a
// template over where its values come from (scan.cu instantiates the plain array form, collide.cu the flag forms).
// Deterministic: integer sums, so the result does not depend on the order tiles finish in.
#pragma once

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

// Loader: where the values come from.  PlainLoad reads an int array (16 bytes at a time when it can); the flag loaders of
// collide.cu compute a 0/1 per element on the fly (alive manifold, new contact point, alive joint), which saves the kernel
// that used to write the flags and the array they went through.
struct PlainLoad
{
    const int* in;
    __device__ __forceinline__ bool vector_ok(const int* out) const { return ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0; }
    __device__ __forceinline__ void load4(int at, int (&v)[4]) const
    {
        const int4 q = *reinterpret_cast<const int4*>(in + at);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    }
    __device__ __forceinline__ int load(int at) const { return in[at]; }
};

template <typename Loader>
__global__ void __launch_bounds__(kScanThreads) k_scan_single(Loader loader, Count nc, int* __restrict__ out, unsigned long long* __restrict__ status,
    unsigned* __restrict__ ticket, int* __restrict__ totalOut)
{
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
    // Warp-striped tile: warp w owns the 512 ints [w * 512, (w + 1) * 512) of the tile as 4 rows of 128; lane l holds the 4
    // consecutive ints at 4 * l of every row (one 16-byte load per row when the arrays are 16-byte aligned: every load and
    // store of the warp is a full 512-byte line set; the blocked layout it replaces read 64-byte pieces per thread).
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warpBase = tile * kScanTile + warp * (32 * kScanItems);
    const bool full = tile * kScanTile + kScanTile <= n;
    const bool vec = full && loader.vector_ok(out);           // 16-byte loads
    const bool vecOut = full && (reinterpret_cast<uintptr_t>(out) & 15u) == 0;   // 16-byte stores
    int v[4][4];
    int rowSum[4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
    {
        const int at = warpBase + r * 128 + 4 * lane;
        if (vec)
            loader.load4(at, v[r]);
        else
        {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[r][e] = (at + e < n) ? loader.load(at + e) : 0;
        }
        rowSum[r] = v[r][0] + v[r][1] + v[r][2] + v[r][3];
    }
    // exclusive prefix of each lane's 4-int group inside the warp's 512 ints
    int groupEx[4];
    int warpTotal = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r)
    {
        const int inc = warp_inclusive(rowSum[r]);
        groupEx[r] = warpTotal + inc - rowSum[r];
        warpTotal += __shfl_sync(0xffffffffu, inc, 31);
    }
    // across the 8 warps of the block
    __shared__ int s_warp[kScanThreads / 32];
    if (lane == 0) s_warp[warp] = warpTotal;
    __syncthreads();
    int warpEx = 0, blockTotal = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w)
    {
        const int t = s_warp[w];
        if (w < warp) warpEx += t;
        blockTotal += t;
    }
    const unsigned mine = unsigned(blockTotal);
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
    const int blockEx = s_prefix + warpEx;
#pragma unroll
    for (int r = 0; r < 4; ++r)
    {
        const int at = warpBase + r * 128 + 4 * lane;
        int run = blockEx + groupEx[r];
        int o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
        {
            o[e] = run;
            run += v[r][e];
        }
        if (vecOut)
            *reinterpret_cast<int4*>(out + at) = make_int4(o[0], o[1], o[2], o[3]);
        else
        {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (at + e < n) out[at + e] = o[e];
        }
        // the thread that holds the last element
        if (totalOut && at <= n - 1 && n - 1 < at + 4)
        {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (at + e == n - 1) *totalOut = o[e] + v[r][e];
        }
    }
    scan_leave(ticket, epoch);
}

// host side: the scratch (ticket words + status words) of the context, a launch of the loader's instantiation
inline int scan_scratch(phyx_b200_ctx* c, int tiles, unsigned long long** status, unsigned** ticket)
{
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
    *status = c->scanTmp.as<unsigned long long>() + 2;
    *ticket = c->scanTmp.as<unsigned>();
    return PHYX_B200_OK;
}

// exclusive scan of loader(0 .. n) into out; if totalDevice is non-null it receives the grand total
template <typename Loader>
int exclusive_scan_with(phyx_b200_ctx* c, const Loader& loader, int* out, Count nc, int* totalDevice)
{
    const int n = nc.v;
    if (n <= 0)
    {
        if (totalDevice) PHYX_CUDA(cudaMemsetAsync(totalDevice, 0, sizeof(int), c->stream));
        return PHYX_B200_OK;
    }
    const int tiles = (n + kScanTile - 1) / kScanTile;
    unsigned long long* status = nullptr;
    unsigned* ticket = nullptr;
    PHYX_TRY(scan_scratch(c, tiles, &status, &ticket));
    k_scan_single<Loader><<<tiles, kScanThreads, 0, c->stream>>>(loader, nc, out, status, ticket, totalDevice);
    c->launches++;
    PHYX_CUDA(cudaGetLastError());
    return PHYX_B200_OK;
}

} // namespace phyx
