// phyx_b200 — single-axis sweep & prune broadphase on the device.
//
// Reference stages replaced here:
//   Collider::UpdateBroadphase   src/Collider.cpp:251-284   key build, radixSort3, entry gather
//   radixFloat / radixSort3      src/base/RadixSort.h:19-95 order-preserving float key, stable LSD
//                                                            sort, digits of 11/11/10 bits
//   Collider::UpdatePairs* sweep src/Collider.cpp:296-366   for i: for j>i while minx[j]<=maxx[i]:
//                                                            y-interval test -> candidate pair
//
// The sort is a stable LSD radix sort with the reference's own three digits, so the permutation
// (including the order of ties: equal keys keep body-index order) is identical to radixSort3's.
// The sweep emits pairs in exactly the reference's order (i ascending, j ascending) by
// count -> exclusive scan -> emit over fixed-size work items, so long scans (the ground body's
// covers every entry) are split over many warps instead of serialising one thread.
#include "common.cuh"
#include "pairset.cuh"

namespace phyx
{

constexpr int kBlock = 256;

// ---- keys -------------------------------------------------------------------------------------

__device__ __forceinline__ unsigned radix_float(float v)   // RadixSort.h:19-26
{
    int f = __float_as_int(v);
    unsigned mask = unsigned(f >> 31) | 0x80000000u;
    return unsigned(f) ^ mask;
}

__device__ __forceinline__ float radix_unfloat(unsigned u)   // inverse of radix_float
{
    unsigned mask = (u & 0x80000000u) ? 0x80000000u : 0xffffffffu;
    return __int_as_float(int(u ^ mask));
}

__global__ void k_make_keys(int n, const float4* __restrict__ aabb, uint2* __restrict__ kv)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    kv[i] = make_uint2(radix_float(aabb[i].x), unsigned(i));
}

// ---- radix sort pass ----------------------------------------------------------------------------
// tile = 256 threads x 16 keys; warp w owns 512 consecutive keys, read as 16 coalesced chunks of 32.

constexpr int kSortItems = 16;
constexpr int kSortWarps = kBlock / 32;
constexpr int kSortTile = kBlock * kSortItems;
constexpr int kMaxDigits = 2048;

__global__ void __launch_bounds__(kBlock) k_radix_hist(const uint2* __restrict__ src, Count nc, int shift, int digits, int numBlocks,
    int* __restrict__ table)
{
    __shared__ int h[kMaxDigits];
    const int n = count_of(nc);
    for (int d = threadIdx.x; d < digits; d += kBlock) h[d] = 0;
    __syncthreads();
    int base = blockIdx.x * kSortTile;
    unsigned mask = unsigned(digits - 1);
#pragma unroll
    for (int k = 0; k < kSortItems; ++k)
    {
        int i = base + k * kBlock + threadIdx.x;
        if (i < n) atomicAdd(&h[(src[i].x >> shift) & mask], 1);
    }
    __syncthreads();
    for (int d = threadIdx.x; d < digits; d += kBlock) table[size_t(d) * numBlocks + blockIdx.x] = h[d];
}

__global__ void __launch_bounds__(kBlock) k_radix_scatter(const uint2* __restrict__ src, uint2* __restrict__ dst, Count nc, int shift,
    int digits, int numBlocks, const int* __restrict__ tableScanned)
{
    const int n = count_of(nc);
    if (blockIdx.x * kSortTile >= n) return;   // (a count read from the device may be shorter than the grid)
    __shared__ unsigned short cnt[kSortWarps][kMaxDigits];   // per-warp digit counts, then prefixes
    __shared__ int gOff[kMaxDigits];                         // global offset of (digit, this block)

    // (only the bins this pass uses: the layout sorts of the solve schedule have 128 and <= 1024 of them)
    for (int k = threadIdx.x; k < kSortWarps * digits; k += kBlock) cnt[k / digits][k % digits] = 0;
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned mask = unsigned(digits - 1);
    const int base = blockIdx.x * kSortTile + warp * (32 * kSortItems);

    uint2 kv[kSortItems];
    unsigned packed[kSortItems];   // digit | rank-within-warp << 11
#pragma unroll
    for (int k = 0; k < kSortItems; ++k)
    {
        int i = base + k * 32 + lane;
        bool valid = i < n;
        kv[k] = valid ? src[i] : make_uint2(0, 0);
        unsigned d = (kv[k].x >> shift) & mask;
        unsigned peers = __match_any_sync(0xffffffffu, valid ? d : (0x10000u | unsigned(lane)));
        int leader = __ffs(peers) - 1;
        unsigned before = __popc(peers & ((1u << lane) - 1u));
        unsigned old = 0;
        if (lane == leader && valid)
        {
            old = cnt[warp][d];
            cnt[warp][d] = (unsigned short)(old + __popc(peers));
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        packed[k] = d | ((old + before) << 11);
        __syncwarp();   // the next chunk's leaders read counters other lanes just wrote
    }
    __syncthreads();
    for (int d = threadIdx.x; d < digits; d += kBlock)
    {
        gOff[d] = tableScanned[size_t(d) * numBlocks + blockIdx.x];
        unsigned run = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w)
        {
            unsigned t = cnt[w][d];
            cnt[w][d] = (unsigned short)run;
            run += t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSortItems; ++k)
    {
        int i = base + k * 32 + lane;
        if (i < n)
        {
            unsigned d = packed[k] & 2047u;
            int pos = gOff[d] + int(cnt[warp][d]) + int(packed[k] >> 11);
            dst[pos] = kv[k];
        }
    }
}

// ---- entries ----------------------------------------------------------------------------------

__global__ void k_gather_entries(int n, const uint2* __restrict__ sorted, const float4* __restrict__ aabb, float2* __restrict__ entryX,
    float2* __restrict__ entryY, unsigned* __restrict__ entryIndex, int* __restrict__ rowOf)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned b = sorted[i].y;
    rowOf[b] = i;   // inverse permutation: the solver stores body rows in this order (solve.cu)
    float4 bb = aabb[b];
    entryX[i] = make_float2(bb.x, bb.z);
    entryY[i] = make_float2((bb.y + bb.w) * 0.5f, (bb.w - bb.y) * 0.5f);   // Collider.cpp:276-279
    entryIndex[i] = b;
}

// ---- sweep ------------------------------------------------------------------------------------

constexpr int kChunk = 1024;   // sweep tests per work item (one warp, 32 rounds)
constexpr int kTile = 256;     // sorted bodies per tile of the shared-memory count pass
constexpr int kTileCap = 4096; // entries of {centery, extenty} a tile may hold (32 KB)
constexpr int kCells = 256;    // height cells a tile's entries are bucketed into

// end_i = first j > i with minx[j] > maxx[i] (the x-break, Collider.cpp:306); minx is sorted.
__global__ void k_sweep_end(int n, const float2* __restrict__ entryX, int* __restrict__ end, int* __restrict__ itemsOf)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float maxx = entryX[i].y;
    int lo = i + 1, hi = n;
    while (lo < hi)
    {
        int mid = (lo + hi) >> 1;
        if (entryX[mid].x > maxx) hi = mid; else lo = mid + 1;
    }
    end[i] = lo;
    itemsOf[i] = (lo - i - 1 + kChunk - 1) / kChunk;
}

__global__ void k_sweep_items(Count nc, const int* __restrict__ itemStart, const int* __restrict__ end, int2* __restrict__ items)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count_of(nc)) return;
    int cnt = (end[i] - i - 1 + kChunk - 1) / kChunk;
    int s = itemStart[i];
    for (int k = 0; k < cnt; ++k) items[s + k] = make_int2(i, k);
}

// One warp per work item.  FILTER: a hit only counts / is emitted if its (index_i, index_j) key is
// not in the manifold cache (Collider.cpp:358-362: contains() before push); the unfiltered hit
// total is still accumulated for the statistics.
template <bool EMIT, bool FILTER>
__global__ void __launch_bounds__(kBlock) k_sweep(const int* __restrict__ numItemsPtr, const int2* __restrict__ items,
    const int* __restrict__ end, const float2* __restrict__ entryY, const unsigned* __restrict__ entryIndex, int* __restrict__ itemCount,
    const int* __restrict__ itemOffset, int2* __restrict__ pairs, unsigned long long* __restrict__ totals,
    const unsigned long long* __restrict__ table, size_t tableMask, const int* __restrict__ totalOutPtr, const unsigned char* __restrict__ tileLong)
{
    const int lane = threadIdx.x & 31;
    const int warpsPerGrid = (gridDim.x * blockDim.x) >> 5;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int numItems = *numItemsPtr;
    unsigned long long localTests = 0, localHits = 0;
    // Warp w owns the items w, w + W, w + 2W, ... (W warps in the grid; neighbouring items, e.g. the thousand
    // chunks of the ground body's scan, go to different warps).  Almost all of them need no work here (the count
    // pass leaves most tiles to the tiled kernel, the emit pass only visits items with output), so the 32 lanes
    // look at 32 of the warp's items at once and the warp then walks the few that remain.
    for (long long first = warp; first < numItems; first += 32ll * warpsPerGrid)
    {
        const long long mine = first + static_cast<long long>(lane) * warpsPerGrid;
        bool need = false;
        if (mine < numItems)
        {
            if (EMIT)
            {
                // the count pass already knows which items produce output (with the cache filter almost every
                // item of a settled scene is empty)
                const int next = (mine + 1 < numItems) ? itemOffset[mine + 1] : *totalOutPtr;
                need = next != itemOffset[mine];
            }
            else
                need = !(tileLong && !tileLong[items[mine].x / kTile]);   // else: counted by the tiled kernel
        }
        unsigned todo = __ballot_sync(0xffffffffu, need);
        while (todo)
        {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1u;
            const int it = int(first + static_cast<long long>(src) * warpsPerGrid);
            int2 item = items[it];
            int i = item.x;
            int j0 = i + 1 + item.y * kChunk;
            int j1 = min(j0 + kChunk, end[i]);
            float2 yi = entryY[i];
            unsigned bi = (EMIT || FILTER) ? entryIndex[i] : 0u;
            int out = EMIT ? itemOffset[it] : 0;
            int count = 0;
            // four rounds of 32 tests per trip, their loads issued together: a round is a chain of dependent accesses
            // (entry, body index, cache probe), and the ground body's scan is a thousand items of 32 rounds each
            constexpr int U = 4;
            for (int jb = j0; jb < j1; jb += 32 * U)
            {
                float2 yj[U];
                bool hit[U];
                unsigned bj[U];
#pragma unroll
                for (int u = 0; u < U; ++u)
                {
                    const int j = jb + u * 32 + lane;
                    yj[u] = j < j1 ? entryY[j] : make_float2(0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
                {
                    const int j = jb + u * 32 + lane;
                    hit[u] = j < j1 && fabsf(yj[u].x - yi.x) <= yi.y + yj[u].y;   // Collider.cpp:309
                    bj[u] = 0;
                }
                if (FILTER)
                {
                    unsigned long long key[U], firstWord[U];
                    size_t slot[U];
#pragma unroll
                    for (int u = 0; u < U; ++u)
                    {
                        if (!EMIT) localHits += __popc(__ballot_sync(0xffffffffu, hit[u]));
                        if (hit[u]) bj[u] = entryIndex[jb + u * 32 + lane];
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u)
                    {
                        key[u] = pair_key(bi, bj[u]);
                        slot[u] = pair_slot(key[u], tableMask);
                        firstWord[u] = hit[u] ? table[slot[u]] : kEmptyPair;
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u)
                    {
                        if (!hit[u]) continue;
                        if (firstWord[u] == key[u])
                            hit[u] = false;
                        else if (firstWord[u] != kEmptyPair)
                            hit[u] = !pair_contains_from(table, tableMask, key[u], (slot[u] + 1) & tableMask);
                    }
                }
                else if (EMIT)
                {
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (hit[u]) bj[u] = entryIndex[jb + u * 32 + lane];
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
                {
                    const unsigned m = __ballot_sync(0xffffffffu, hit[u]);
                    if (EMIT)
                    {
                        if (hit[u]) pairs[out + __popc(m & ((1u << lane) - 1u))] = make_int2(int(bi), int(bj[u]));
                        out += __popc(m);
                    }
                    else
                        count += __popc(m);
                }
            }
            if (!EMIT && lane == 0)
            {
                itemCount[it] = count;
                localTests += (unsigned long long)(j1 - j0);
            }
        }
    }
    if (!EMIT && lane == 0)
    {
        if (localTests) atomicAdd(&totals[0], localTests);
        if (FILTER && localHits) atomicAdd(&totals[1], localHits);
    }
}

// Count pass, shared-memory tiled: a block takes kTile consecutive sorted bodies; the union of their
// scan ranges (their own successors up to the furthest x-break) is loaded into shared memory once and
// every warp scans its bodies from there, chunk by chunk, writing the same per-item counts the
// item kernel would.  Neighbouring bodies scan almost the same entries (on an R-row pyramid each
// scan is ~R/2 long), so this replaces ~R/2 global reads per entry by one.  Tiles whose union does
// not fit (the ground body's scan covers every entry) are flagged and left to the item kernel.

// Up to 32 queued hits, one per lane: {rel (12 bits) | body - i0 (8) | chunk (3)}; a pair that is not in the manifold cache
// counts for its (body, chunk).  Returns the new queue length.
__device__ __forceinline__ int drain_queue(const unsigned* queue, int qn, int lane, int i0, int first, const unsigned* __restrict__ entryIndex,
    const unsigned long long* __restrict__ table, size_t tableMask, unsigned* counts, int countWords)
{
    const int take = min(qn, 32);
    if (lane < take)
    {
        const unsigned e = queue[qn - 1 - lane];
        const int rel = int(e & 0xfffu), local = int((e >> 12) & 0xffu), chunk = int(e >> 20);
        const unsigned bi = entryIndex[i0 + local], bj = entryIndex[first + rel];
        if (!pair_contains(table, tableMask, pair_key(bi, bj))) atomicAdd(&counts[local * countWords + (chunk >> 1)], 1u << ((chunk & 1) * 16));
    }
    __syncwarp();
    return qn - take;
}

template <bool FILTER>
__global__ void __launch_bounds__(kBlock) k_sweep_count_tiled(Count nc, const int* __restrict__ end, const int* __restrict__ itemStart,
    const float2* __restrict__ entryY, const unsigned* __restrict__ entryIndex, int* __restrict__ itemCount, unsigned char* __restrict__ tileLong,
    unsigned long long* __restrict__ totals, const unsigned long long* __restrict__ table, size_t tableMask)
{
    const int n = count_of(nc);
    if (n == 0) return;   // (a stopped deferred step)
    __shared__ float2 tileY[kTileCap];
    __shared__ unsigned short byCell[kTileCap];     // tile entries (relative index) bucketed by y cell
    __shared__ int cellStart[kCells + 1], cellFill[kCells];
    __shared__ int sMaxEnd;
    __shared__ unsigned sYmin, sYmax, sEymax;        // order-preserving float bits
    // FILTER: hits wait in a per-warp queue and are looked up in the cache 32 at a time (one per lane); what survives is
    // counted per (body, chunk) in shared memory, two 16-bit counters to a word
    constexpr int kQueue = 56, kDrainAbove = kQueue - 32, kCountWords = (kTileCap / kChunk + 2) / 2;
    __shared__ unsigned sQueue[FILTER ? kBlock / 32 : 1][FILTER ? kQueue : 1];
    __shared__ unsigned sCount[FILTER ? kTile : 1][FILTER ? kCountWords : 1];
    const int i0 = blockIdx.x * kTile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0)
    {
        sMaxEnd = 0;
        sYmin = 0xffffffffu;
        sYmax = 0u;
        sEymax = 0u;
    }
    for (int k = threadIdx.x; k < kCells; k += kBlock) cellFill[k] = 0;
    if (FILTER)
        for (int k = threadIdx.x; k < kTile * kCountWords; k += kBlock) (&sCount[0][0])[k] = 0u;
    __syncthreads();
    {
        const int i = i0 + threadIdx.x;
        int e = (i < n) ? end[i] : 0;
        for (int o = 16; o > 0; o >>= 1) e = max(e, __shfl_xor_sync(0xffffffffu, e, o));
        if (lane == 0) atomicMax(&sMaxEnd, e);
    }
    __syncthreads();
    const int first = i0 + 1, span = sMaxEnd - first;
    if (span > kTileCap)
    {
        if (threadIdx.x == 0) tileLong[blockIdx.x] = 1;
        return;
    }
    if (threadIdx.x == 0) tileLong[blockIdx.x] = 0;
    // load the tile, find its y range and largest half extent
    {
        unsigned lo = 0xffffffffu, hi = 0u, ex = 0u;
        for (int k = threadIdx.x; k < span; k += kBlock)
        {
            const float2 y = entryY[first + k];
            tileY[k] = y;
            lo = min(lo, radix_float(y.x));
            hi = max(hi, radix_float(y.x));
            ex = max(ex, radix_float(y.y));
        }
        for (int o = 16; o > 0; o >>= 1)
        {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
            ex = max(ex, __shfl_xor_sync(0xffffffffu, ex, o));
        }
        if (lane == 0)
        {
            atomicMin(&sYmin, lo);
            atomicMax(&sYmax, hi);
            atomicMax(&sEymax, ex);
        }
    }
    __syncthreads();
    // Bucket the tile's entries by height.  A body can only touch entries whose centre lies within
    // (its own half extent + the tile's largest half extent) of its own, i.e. in a few neighbouring
    // cells, so the scan of a body visits a handful of candidates instead of its whole x range
    // (~R/2 entries on an R-row pyramid).  The counts are those of the full scan.
    const float ymin = radix_unfloat(sYmin), ymax = radix_unfloat(sYmax), eymax = radix_unfloat(sEymax);
    const float cellH = fmaxf((ymax - ymin) / float(kCells), 1e-6f);
    const float invCellH = 1.0f / cellH;
    for (int k = threadIdx.x; k < span; k += kBlock) atomicAdd(&cellFill[min(kCells - 1, max(0, int((tileY[k].x - ymin) * invCellH)))], 1);
    __syncthreads();
    if (warp == 0)
    {
        int run = 0;
        for (int c0 = 0; c0 < kCells; c0 += 32)
        {
            const int v = cellFill[c0 + lane];
            int inc = v;
            for (int o = 1; o < 32; o <<= 1)
            {
                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            cellStart[c0 + lane] = run + inc - v;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) cellStart[kCells] = run;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < kCells; k += kBlock) cellFill[k] = cellStart[k];
    __syncthreads();
    for (int k = threadIdx.x; k < span; k += kBlock)
    {
        const int cell = min(kCells - 1, max(0, int((tileY[k].x - ymin) * invCellH)));
        byCell[atomicAdd(&cellFill[cell], 1)] = (unsigned short)k;
    }
    __syncthreads();

    unsigned long long localTests = 0, localHits = 0;
    int qn = 0;   // entries in this warp's queue (warp-uniform)
    for (int i = i0 + warp; i < min(i0 + kTile, n); i += kBlock / 32)
    {
        const int e = end[i];
        const int len = e - i - 1;
        if (len <= 0) continue;
        const float2 yi = (i == i0) ? entryY[i] : tileY[i - first];
        const int item = itemStart[i];
        const int numItems = (len + kChunk - 1) / kChunk;
        // candidates: the cells within reach of this body's y interval
        const float reach = yi.y + eymax;
        // clamp in float BEFORE the conversion: a tall body (a wall, a thick slab) makes the quotient exceed the int range,
        // and int(inf-ish) +/- 1 would be signed overflow
        const float fLo = fminf(fmaxf((yi.x - reach - ymin) * invCellH, 0.0f), float(kCells - 1));
        const float fHi = fminf(fmaxf((yi.x + reach - ymin) * invCellH, 0.0f), float(kCells - 1));
        const int cLo = max(0, int(fLo) - 1);
        const int cHi = min(kCells - 1, int(fHi) + 1);
        const int kBegin = cellStart[cLo], kEnd = cellStart[cHi + 1];
        // per work item (chunk of kChunk tests) counts; a tile-able body has at most kTileCap / kChunk items
        int counts[kTileCap / kChunk + 1];
#pragma unroll
        for (int c = 0; c <= kTileCap / kChunk; ++c) counts[c] = 0;
        for (int kb = kBegin; kb < kEnd; kb += 32 * 32)
        {
            // y test for up to 32 rounds of candidates (bit t = this lane's test in round t) ...
            unsigned hits = 0;
            for (int t = 0; t < 32 && kb + t * 32 < kEnd; ++t)
            {
                const int k = kb + t * 32 + lane;
                if (k < kEnd)
                {
                    const int rel = byCell[k];
                    const int j = first + rel;
                    const float2 yj = tileY[rel];
                    if (j > i && j < e && fabsf(yj.x - yi.x) <= yi.y + yj.y) hits |= 1u << t;   // Collider.cpp:306,309
                }
            }
            if (FILTER)
            {
                // ... then the hits go to the warp's queue (the cache lookup is two dependent global loads: done one body at
                // a time it was the whole kernel, 2-4 probes in flight per warp)
                localHits += __popc(hits);
                while (__any_sync(0xffffffffu, hits != 0u))
                {
                    unsigned entry = 0u;
                    const bool have = hits != 0u;
                    if (have)
                    {
                        const int t = __ffs(hits) - 1;
                        hits &= hits - 1;
                        const int rel = byCell[kb + t * 32 + lane];
                        const int chunk = (first + rel - i - 1) / kChunk;
                        entry = unsigned(rel) | (unsigned(i - i0) << 12) | (unsigned(chunk) << 20);
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, have);
                    if (have) sQueue[warp][qn + __popc(m & ((1u << lane) - 1u))] = entry;
                    qn += __popc(m);
                    __syncwarp();
                    while (qn > kDrainAbove) qn = drain_queue(sQueue[warp], qn, lane, i0, first, entryIndex, table, tableMask, &sCount[0][0], kCountWords);
                }
            }
            else
            {
                while (hits)
                {
                    const int t = __ffs(hits) - 1;
                    hits &= hits - 1;
                    const int j = first + byCell[kb + t * 32 + lane];
                    const int chunk = (j - i - 1) / kChunk;
#pragma unroll
                    for (int c = 0; c <= kTileCap / kChunk; ++c)
                        if (c == chunk) ++counts[c];
                }
            }
        }
        if (!FILTER)
        {
#pragma unroll
            for (int c = 0; c <= kTileCap / kChunk; ++c)
            {
                if (c < numItems)
                {
                    int v = counts[c];
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (lane == 0) itemCount[item + c] = v;
                }
            }
        }
        localTests += (unsigned long long)len;   // what the reference's scan would have tested
    }
    if (FILTER)
    {
        while (qn > 0) qn = drain_queue(sQueue[warp], qn, lane, i0, first, entryIndex, table, tableMask, &sCount[0][0], kCountWords);
        __syncwarp();
        // the counts of this warp's bodies (only this warp touched them): lane = body, one item per chunk
        const int i = i0 + warp + lane * (kBlock / 32);
        if (i < min(i0 + kTile, n))
        {
            const int len = end[i] - i - 1;
            if (len > 0)
            {
                const int item = itemStart[i], numItems = (len + kChunk - 1) / kChunk;
                for (int c = 0; c < numItems; ++c) itemCount[item + c] = int((sCount[i - i0][c >> 1] >> ((c & 1) * 16)) & 0xffffu);
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) localHits += __shfl_xor_sync(0xffffffffu, localHits, o);
    if (lane == 0)
    {
        if (localTests) atomicAdd(&totals[0], localTests);
        if (FILTER && localHits) atomicAdd(&totals[1], localHits);
    }
}

// ---- deferred step: the sweep's counts stay on the device (common.cuh StepCtl) ---------------------
// counters: [0] work items, [1] emitted pairs, then two 64-bit totals (tests, unfiltered hits) at byte 16

__global__ void k_gate_items(StepCtl* ctl, int* __restrict__ counters, int capItems)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int items = counters[0];
    if (!ctl->stop)
    {
        ctl->items = items;
        if (items > capItems) ctl_stop(ctl, kStagePairs, items, 1);
    }
    if (ctl->stop) counters[0] = 0;
}

__global__ void k_gate_pairs(StepCtl* ctl, int* __restrict__ counters, int capPairs)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (!ctl->stop)
    {
        const int pairs = counters[1];
        const unsigned long long* totals = reinterpret_cast<const unsigned long long*>(counters + 4);
        ctl->newPairs = pairs;
        ctl->tests = totals[0];
        ctl->hits = totals[1];
        if (pairs > capPairs)
            ctl_stop(ctl, kStagePairs, pairs, 2);
        else
        {
            ctl->appendFirst = ctl->manifolds;   // Collider.cpp:358-362: new manifolds are appended in sweep order
            ctl->manifolds += pairs;
        }
    }
    if (ctl->stop) counters[0] = counters[1] = 0;
}

// ---- host side ----------------------------------------------------------------------------------

int radix_pass(phyx_b200_ctx* c, const uint2* src, uint2* dst, int n, int shift, int digits)
{
    Count nc;
    nc.v = n;
    nc.p = nullptr;
    nc.mul = 1;
    return radix_pass_count(c, src, dst, nc, shift, digits);
}

int radix_pass_count(phyx_b200_ctx* c, const uint2* src, uint2* dst, Count nc, int shift, int digits)
{
    const int n = nc.v;
    if (n <= 0) return PHYX_B200_OK;
    int blocks = (n + kSortTile - 1) / kSortTile;
    size_t tableInts = size_t(digits) * blocks;
    PHYX_TRY(c->hist.reserve(tableInts * sizeof(int)));
    k_radix_hist<<<blocks, kBlock, 0, c->stream>>>(src, nc, shift, digits, blocks, c->hist.as<int>());
    c->launches++;
    PHYX_TRY(exclusive_scan_i32(c, c->hist.as<int>(), c->hist.as<int>(), int(tableInts), nullptr));
    k_radix_scatter<<<blocks, kBlock, 0, c->stream>>>(src, dst, nc, shift, digits, blocks, c->hist.as<int>());
    c->launches++;
    PHYX_CUDA(cudaGetLastError());
    return PHYX_B200_OK;
}

int broadphase_update(phyx_b200_ctx* c)
{
    int n = c->bodyCount;
    size_t n1 = size_t(n > 0 ? n : 1);
    PHYX_TRY(c->sortA.reserve(n1 * sizeof(uint2)));
    PHYX_TRY(c->sortB.reserve(n1 * sizeof(uint2)));
    PHYX_TRY(c->entry.reserve(n1 * 2 * sizeof(float2)));
    PHYX_TRY(c->entryIndex.reserve(n1 * sizeof(unsigned)));
    c->broadphaseValid = true;
    if (n == 0) return PHYX_B200_OK;
    int grid = (n + kBlock - 1) / kBlock;
    PHYX_CUDA(record_event(c, c->evBp[0]));
    k_make_keys<<<grid, kBlock, 0, c->stream>>>(n, c->aabb.as<float4>(), c->sortA.as<uint2>());
    c->launches++;
    // radixSort3: digits 0-10, 11-21, 22-31; A -> B -> A -> B (RadixSort.h:74-88)
    PHYX_TRY(radix_pass(c, c->sortA.as<uint2>(), c->sortB.as<uint2>(), n, 0, 2048));
    PHYX_TRY(radix_pass(c, c->sortB.as<uint2>(), c->sortA.as<uint2>(), n, 11, 2048));
    PHYX_TRY(radix_pass(c, c->sortA.as<uint2>(), c->sortB.as<uint2>(), n, 22, 1024));
    PHYX_CUDA(record_event(c, c->evBp[1]));
    c->sortTimed = true;
    float2* entryX = c->entry.as<float2>();
    float2* entryY = entryX + n1;
    PHYX_TRY(c->rowOf.reserve(n1 * sizeof(int)));
    k_gather_entries<<<grid, kBlock, 0, c->stream>>>(n, c->sortB.as<uint2>(), c->aabb.as<float4>(), entryX, entryY, c->entryIndex.as<unsigned>(),
        c->rowOf.as<int>());
    c->launches++;
    c->rowOrderValid = true;
    c->rowOrderBodies = n;
    PHYX_CUDA(cudaGetLastError());
    return PHYX_B200_OK;
}

// Runs count + scan (+ emit into c->pairs).  Leaves the pair list on the device; c->lastPairs /
// c->lastTests hold the totals.
// filter = false: every overlapping pair (lastPairs of them) ends up in c->pairs.
// filter = true : only pairs whose key is not in c->pairTable are emitted (lastNewPairs of them);
//                 lastPairs still counts every overlapping pair.
int broadphase_sweep(phyx_b200_ctx* c, phyx_b200_broadphase_stats* stats, bool filter)
{
    if (!c->broadphaseValid)
    {
        set_error("sweep_pairs: call update_broadphase first (bodies moved since the last sort)");
        return PHYX_B200_ERR_STATE;
    }
    int n = c->bodyCount;
    c->lastPairs = c->lastTests = c->lastNewPairs = 0;
    if (stats) stats->pairs = stats->tests = 0;
    if (n < 2) return PHYX_B200_OK;
    if (filter && c->pairTableSlots == 0) PHYX_TRY(collide_rebuild_pair_table(c));
    size_t n1 = size_t(n);
    float2* entryX = c->entry.as<float2>();
    float2* entryY = entryX + n1;
    PHYX_TRY(c->sweepEnd.reserve(n1 * sizeof(int)));
    PHYX_TRY(c->itemStart.reserve((n1 + 1) * sizeof(int)));
    PHYX_TRY(c->counters.reserve(64));
    int* d_numItems = c->counters.as<int>();
    int* d_numPairs = d_numItems + 1;
    unsigned long long* d_totals = reinterpret_cast<unsigned long long*>(c->counters.as<char>() + 16);   // tests, unfiltered hits
    PHYX_CUDA(cudaMemsetAsync(c->counters.ptr, 0, 64, c->stream));

    int grid = (n + kBlock - 1) / kBlock;
    k_sweep_end<<<grid, kBlock, 0, c->stream>>>(n, entryX, c->sweepEnd.as<int>(), c->itemStart.as<int>());
    c->launches++;
    PHYX_TRY(exclusive_scan_i32(c, c->itemStart.as<int>(), c->itemStart.as<int>(), n, d_numItems));
    const bool deferred = c->def.active;   // (only the cache-filtered sweep of a whole step is ever deferred)
    int numItems = 0;
    if (deferred)
    {
        numItems = c->def.capItems;
        k_gate_items<<<1, 32, 0, c->stream>>>(c->ctl(), d_numItems, numItems);
        c->launches++;
    }
    else
        PHYX_TRY(fetch_small(c, d_numItems, sizeof(int), &numItems));
    if (numItems == 0) return PHYX_B200_OK;

    PHYX_TRY(c->items.reserve(size_t(numItems) * sizeof(int2)));
    PHYX_TRY(c->itemCount.reserve(size_t(numItems) * sizeof(int)));
    const Count nLive = c->count(n, &StepCtl::bodies);
    k_sweep_items<<<grid, kBlock, 0, c->stream>>>(nLive, c->itemStart.as<int>(), c->sweepEnd.as<int>(), c->items.as<int2>());
    c->launches++;

    const unsigned long long* table = c->pairTable.as<unsigned long long>();
    const size_t mask = c->pairTableSlots ? c->pairTableSlots - 1 : 0;
    int warpsPerBlock = kBlock / 32;
    int sweepGrid = min((numItems + warpsPerBlock - 1) / warpsPerBlock, c->numSMs * 8);
    // count pass: tiles first (shared-memory reuse), then the item kernel for the tiles that did not fit
    const int tiles = (n + kTile - 1) / kTile;
    PHYX_TRY(c->tileLong.reserve(size_t(tiles)));
    unsigned char* tileLong = c->tileLong.as<unsigned char>();
    if (filter)
    {
        k_sweep_count_tiled<true><<<tiles, kBlock, 0, c->stream>>>(nLive, c->sweepEnd.as<int>(), c->itemStart.as<int>(), entryY, c->entryIndex.as<unsigned>(),
            c->itemCount.as<int>(), tileLong, d_totals, table, mask);
        k_sweep<false, true><<<sweepGrid, kBlock, 0, c->stream>>>(d_numItems, c->items.as<int2>(), c->sweepEnd.as<int>(), entryY,
            c->entryIndex.as<unsigned>(), c->itemCount.as<int>(), nullptr, nullptr, d_totals, table, mask, nullptr, tileLong);
    }
    else
    {
        k_sweep_count_tiled<false><<<tiles, kBlock, 0, c->stream>>>(nLive, c->sweepEnd.as<int>(), c->itemStart.as<int>(), entryY, c->entryIndex.as<unsigned>(),
            c->itemCount.as<int>(), tileLong, d_totals, nullptr, 0);
        k_sweep<false, false><<<sweepGrid, kBlock, 0, c->stream>>>(d_numItems, c->items.as<int2>(), c->sweepEnd.as<int>(), entryY,
            c->entryIndex.as<unsigned>(), c->itemCount.as<int>(), nullptr, nullptr, d_totals, nullptr, 0, nullptr, tileLong);
    }
    c->launches += 2;
    struct { int items, pairs; long long pad; unsigned long long tests, hits; } host;
    if (deferred)
    {
        Count ni;
        ni.v = numItems;
        ni.p = d_numItems;
        ni.mul = 1;
        PHYX_TRY(exclusive_scan_count(c, c->itemCount.as<int>(), c->itemCount.as<int>(), ni, d_numPairs));
        k_gate_pairs<<<1, 32, 0, c->stream>>>(c->ctl(), d_numItems, c->def.capNewPairs);
        c->launches++;
        host.items = numItems;
        host.pairs = c->def.capNewPairs;   // bound; the true totals come home with the step (StepCtl)
        host.tests = host.hits = 0;
    }
    else
    {
        PHYX_TRY(exclusive_scan_i32(c, c->itemCount.as<int>(), c->itemCount.as<int>(), numItems, d_numPairs));
        PHYX_TRY(fetch_small(c, c->counters.ptr, sizeof(host), &host));
    }
    c->lastTests = (long long)host.tests;
    c->lastPairs = filter ? (long long)host.hits : host.pairs;
    c->lastNewPairs = filter ? host.pairs : 0;
    if (host.pairs > 0)
    {
        PHYX_TRY(c->pairs.reserve(size_t(host.pairs) * sizeof(int2)));
        if (filter)
            k_sweep<true, true><<<sweepGrid, kBlock, 0, c->stream>>>(d_numItems, c->items.as<int2>(), c->sweepEnd.as<int>(), entryY,
                c->entryIndex.as<unsigned>(), nullptr, c->itemCount.as<int>(), c->pairs.as<int2>(), nullptr, table, mask, d_numPairs, nullptr);
        else
            k_sweep<true, false><<<sweepGrid, kBlock, 0, c->stream>>>(d_numItems, c->items.as<int2>(), c->sweepEnd.as<int>(), entryY,
                c->entryIndex.as<unsigned>(), nullptr, c->itemCount.as<int>(), c->pairs.as<int2>(), nullptr, nullptr, 0, d_numPairs, nullptr);
        c->launches++;
    }
    PHYX_CUDA(cudaGetLastError());
    if (stats)
    {
        stats->pairs = c->lastPairs;
        stats->tests = c->lastTests;
    }
    return PHYX_B200_OK;
}

} // namespace phyx
