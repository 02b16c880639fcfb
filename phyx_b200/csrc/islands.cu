// phyx_b200 — islands of the contact graph on the device.
//
// Reference stage replaced here:
//   Solver::GatherIslands   src/Solver.cpp:285-454   union-find over the dynamic bodies joined by contact joints (static
//                                                     bodies do not merge islands, :304,316-317), islands numbered in order
//                                                     of their first body, joint counts, coalescing of consecutive islands
//                                                     into groups of at least kIslandMinSize = 256 joints (:383-415)
//
// The reference uses the islands to run SolveJointIsland once per group on its worker threads (Solver.cpp:73-92).  Here
// every schedule already relaxes all islands at once (a colour spans the whole world), so the islands serve two other
// purposes: Solver::islandCount / islandMaxSize, which the demo's HUD reads (src/main.cpp:358-360), and the partition of
// ONE world's solve over several devices by island: islands never exchange impulses, so a rank that relaxes
// whole islands needs no exchange inside the solve.
//
// Union-find: lock-free hooking of the larger root under the smaller with atomicCAS, retried until the two ends of a
// manifold have one root (every edge is processed exactly once, concurrently), then full path compression.  The root of
// an island is therefore its smallest body index, and ranking the roots in index order gives exactly the reference's
// island numbering (island_index is assigned at the first body of each island, :343-355).
#include "common.cuh"

#include <algorithm>

namespace phyx
{

constexpr int kBlock = 256;
constexpr int kIslandMinSize = 256;   // Solver.cpp:11

__device__ __forceinline__ int island_find(int* parent, int x)
{
    int p = parent[x];
    while (p != x)
    {
        const int g = parent[p];
        if (g != p) parent[x] = g;   // path halving (a benign race: parents only ever move towards the root)
        x = p;
        p = g;
    }
    return x;
}

// keep = true: start from the (compressed) forest of the previous step instead of singletons: unions are only ever added,
// so islands whose contacts have broken stay merged until the next full build (the island partition of the solve only
// needs whole islands, not minimal ones; phyx_b200_build_islands always builds exactly).  result[4] = 1 if a body changed
// between static and dynamic since the forest was built (the caller then rebuilds from singletons).
__global__ void __launch_bounds__(kBlock) k_island_init(int nb, const float4* __restrict__ params, int* __restrict__ parent, int* __restrict__ counts, bool keep,
    int* __restrict__ result)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const float4 p = params[b];
    const bool isStatic = p.x == 0.0f && p.y == 0.0f;   // Solver.cpp:304
    if (keep)
    {
        if (isStatic != (parent[b] < 0)) result[4] = 1;
    }
    else
        parent[b] = isStatic ? -1 : b;
    counts[b] = 0;
}

// Solver.cpp:310-324, for every manifold that has a joint (a live contact point)
__global__ void __launch_bounds__(kBlock) k_island_hook(int M, const int2* __restrict__ manBody, const int* __restrict__ manCount, int* __restrict__ parent)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M || manCount[m] == 0) return;
    const int2 b = manBody[m];
    if (parent[b.x] < 0 || parent[b.y] < 0) return;
    int ra = island_find(parent, b.x), rb = island_find(parent, b.y);
    while (ra != rb)
    {
        const int hi = max(ra, rb), lo = min(ra, rb);
        const int seen = atomicCAS(&parent[hi], hi, lo);
        if (seen == hi) break;
        ra = island_find(parent, hi);
        rb = island_find(parent, lo);
    }
}

__global__ void __launch_bounds__(kBlock) k_island_compress(int nb, int* __restrict__ parent, int* __restrict__ isRoot)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    int r = -1;
    if (parent[b] >= 0)
    {
        r = b;
        while (parent[r] != r) r = parent[r];
    }
    isRoot[b] = r == b;
    // written after every thread of the grid has read its chain?  No: chains only shorten, and every value on a chain
    // leads to the same root, so writing early is harmless
    if (r >= 0) parent[b] = r;
}

// island number of every body (-1: static) and joints per island (Solver.cpp:361-379)
__global__ void __launch_bounds__(kBlock) k_island_number(int nb, const int* __restrict__ parent, const int* __restrict__ rootRank, int* __restrict__ islandOf)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const int r = parent[b];
    islandOf[b] = r < 0 ? -1 : rootRank[r];
}

__global__ void __launch_bounds__(kBlock) k_island_count(int M, const int2* __restrict__ manBody, const int* __restrict__ manCount, const int* __restrict__ islandOf,
    int* __restrict__ counts)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    int island = -1, n = 0;
    if (m < M)
    {
        n = manCount[m];
        if (n > 0)
        {
            const int2 b = manBody[m];
            const int i1 = islandOf[b.x], i2 = islandOf[b.y];
            island = i1 < 0 ? i2 : i1;   // -1 if both are static
        }
    }
    // neighbouring manifolds mostly belong to one island (a pile is thousands of manifolds): one atomic per island and warp
    const unsigned peers = __match_any_sync(0xffffffffu, island);
    int sum = 0;
    for (unsigned rest = peers; rest; rest &= rest - 1) sum += __shfl_sync(peers, n, __ffs(rest) - 1);
    if (island >= 0 && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&counts[island], sum);
}

// Solver.cpp:383-415: consecutive islands are merged until a group holds at least kIslandMinSize joints.  The running sum
// with a reset is sequential in nature; ONE WARP walks the islands 32 at a time: an inclusive scan of the chunk's counts,
// then for every group that closes inside the chunk one ballot finds where.  result: [0] islands before coalescing,
// [1] groups = Solver::islandCount, [2] islandMaxSize, [3] joints in islands
__global__ void k_island_coalesce(const int* __restrict__ islandsPtr, const int* __restrict__ counts, int* __restrict__ group, int* __restrict__ groupSize,
    int* __restrict__ result)
{
    if (blockIdx.x != 0 || threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    const int islands = *islandsPtr;
    int running = 0, groups = 0, largest = 0, total = 0;
    for (int base = 0; base < islands; base += 32)
    {
        const int i = base + lane;
        const int cnt = i < islands ? counts[i] : 0;
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        int offset = 0, start = 0, closesBefore = 0, closed = 0;
        // fast path: nothing carried in and every island of the chunk is a group of its own (separate piles)
        if (running == 0 && __all_sync(0xffffffffu, i >= islands || cnt >= kIslandMinSize))
        {
            const int valid = min(32, islands - base);
            if (i < islands)
            {
                group[i] = groups + lane;
                groupSize[groups + lane] = cnt;
            }
            int mx = cnt;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            largest = max(largest, mx);
            total += __shfl_sync(0xffffffffu, incl, 31);
            groups += valid;
            continue;
        }
        for (;;)
        {
            const int val = running + incl - offset;   // the running count including island i, for lanes from `start` on
            const bool closes = lane >= start && i < islands && (val >= kIslandMinSize || (val > 0 && i == islands - 1));
            const unsigned m = __ballot_sync(0xffffffffu, closes);
            if (!m) break;
            const int j = __ffs(m) - 1;
            const int size = __shfl_sync(0xffffffffu, val, j);
            if (lane == 0) groupSize[groups + closed] = size;
            largest = max(largest, size);
            total += size;
            if (lane > j) ++closesBefore;
            ++closed;
            running = 0;
            offset = __shfl_sync(0xffffffffu, incl, j);
            start = j + 1;
        }
        if (i < islands) group[i] = groups + closesBefore;
        groups += closed;
        running += __shfl_sync(0xffffffffu, incl, 31) - offset;
    }
    if (lane == 0)
    {
        result[0] = islands;
        result[1] = groups;
        result[2] = largest;
        result[3] = total;
    }
}

// owner rank of every island group: contiguous runs of groups with about equal joint counts: a group goes to
// the rank whose share of the joints its midpoint falls into.  before[g] = joints of the groups in front of g.
__global__ void __launch_bounds__(kBlock) k_island_owner_cuts(int ranks, const int* __restrict__ result, const int* __restrict__ groupSize,
    const int* __restrict__ before, int* __restrict__ groupOwner)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= result[1]) return;
    const long long total = result[3];
    const long long mid = 2ll * before[g] + groupSize[g];
    const int owner = total > 0 ? int((mid * ranks) / (2 * total)) : 0;
    groupOwner[g] = min(max(owner, 0), ranks - 1);
}

__global__ void __launch_bounds__(kBlock) k_island_body_owner(int nb, const int* __restrict__ islandOf, const int* __restrict__ group, const int* __restrict__ groupOwner,
    const int* __restrict__ result, int* __restrict__ bodyGroup, unsigned char* __restrict__ bodyOwner)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const int i = islandOf[b];
    const int g = i < 0 ? -1 : group[i];
    bodyGroup[b] = g;
    if (bodyOwner) bodyOwner[b] = (g >= 0 && g < result[1]) ? (unsigned char)groupOwner[g] : 0;
}

// Islands of the resident manifolds.  Leaves islandOf / bodyGroup (per body) and, with ranks > 1, bodyOwner on the device.
int islands_build(phyx_b200_ctx* c, int ranks, int* islandCount, int* islandMaxSize, int* islandsBeforeCoalescing, bool exact)
{
    const int nb = c->bodyCount, M = c->manifoldCount;
    const size_t nb1 = size_t(nb > 0 ? nb : 1);
    // parent | isRoot | rootRank | counts | islandOf | group | groupSize | groupOwner | bodyGroup  (ints), result[8]
    PHYX_TRY(c->islandTmp.reserve((nb1 * 9 + 16) * sizeof(int)));
    PHYX_TRY(c->bodyOwner.reserve(nb1));
    int* parent = c->islandTmp.as<int>();
    int* isRoot = parent + nb1;
    int* rootRank = isRoot + nb1;
    int* counts = rootRank + nb1;
    int* islandOf = counts + nb1;
    int* group = islandOf + nb1;
    int* groupSize = group + nb1;
    int* groupOwner = groupSize + nb1;
    int* bodyGroup = groupOwner + nb1;
    int* result = bodyGroup + nb1;   // [0..3] counts, [4] statics changed, [8] islands before coalescing
    int host[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    // the forest of the previous step can be reused (unions only added) a few steps in a row
    bool keep = !exact && c->islandForestBodies == nb && c->islandForestAge < 8;
    for (int attempt = 0; attempt < 2 && nb > 0; ++attempt)
    {
        const int gridB = (nb + kBlock - 1) / kBlock, gridM = (M + kBlock - 1) / kBlock;
        PHYX_CUDA(cudaMemsetAsync(result, 0, 8 * sizeof(int), c->stream));
        k_island_init<<<gridB, kBlock, 0, c->stream>>>(nb, c->params.as<float4>(), parent, counts, keep, result);
        if (M > 0) k_island_hook<<<gridM, kBlock, 0, c->stream>>>(M, c->manBody.as<int2>(), c->manCount.as<int>(), parent);
        k_island_compress<<<gridB, kBlock, 0, c->stream>>>(nb, parent, isRoot);
        c->launches += 3;
        PHYX_TRY(exclusive_scan_i32(c, isRoot, rootRank, nb, result + 8));
        k_island_number<<<gridB, kBlock, 0, c->stream>>>(nb, parent, rootRank, islandOf);
        if (M > 0) k_island_count<<<gridM, kBlock, 0, c->stream>>>(M, c->manBody.as<int2>(), c->manCount.as<int>(), islandOf, counts);
        PHYX_CUDA(cudaMemsetAsync(groupSize, 0, size_t(nb) * sizeof(int), c->stream));
        k_island_coalesce<<<1, 32, 0, c->stream>>>(result + 8, counts, group, groupSize, result);
        if (ranks > 1)
        {
            // (at most one group per island, so nb bounds the grid; entries beyond the group count are never read)
            PHYX_TRY(exclusive_scan_i32(c, groupSize, counts, nb, nullptr));   // `counts` is dead: reuse it for the prefix
            k_island_owner_cuts<<<gridB, kBlock, 0, c->stream>>>(ranks, result, groupSize, counts, groupOwner);
        }
        k_island_body_owner<<<gridB, kBlock, 0, c->stream>>>(nb, islandOf, group, groupOwner, result, bodyGroup, ranks > 1 ? c->bodyOwner.as<unsigned char>() : nullptr);
        c->launches += 5;
        PHYX_TRY(fetch_small(c, result, sizeof(host), host));
        if (!(keep && host[4])) break;
        keep = false;   // a body changed between static and dynamic: the old forest is void
    }
    c->islandForestBodies = nb;
    c->islandForestAge = keep ? c->islandForestAge + 1 : 0;
    c->islandsValid = true;
    c->islandBodies = nb;
    c->islandCount = host[1];
    c->islandMaxSize = host[2];
    if (islandCount) *islandCount = host[1];
    if (islandMaxSize) *islandMaxSize = host[2];
    if (islandsBeforeCoalescing) *islandsBeforeCoalescing = host[0];
    return PHYX_B200_OK;
}

int islands_download(phyx_b200_ctx* c, int* islandOfBody, int* groupOfBody)
{
    const int nb = c->bodyCount;
    if (!c->islandsValid || c->islandBodies != nb)
    {
        set_error("download_islands: call build_islands first");
        return PHYX_B200_ERR_STATE;
    }
    if (nb == 0) return PHYX_B200_OK;
    const size_t nb1 = size_t(nb);
    const int* base = c->islandTmp.as<int>();
    if (islandOfBody) PHYX_CUDA(cudaMemcpyAsync(islandOfBody, base + 4 * nb1, nb1 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (groupOfBody) PHYX_CUDA(cudaMemcpyAsync(groupOfBody, base + 8 * nb1, nb1 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PHYX_CUDA(cudaStreamSynchronize(c->stream));
    return PHYX_B200_OK;
}

// ---- one world's solve split by island over `ranks` devices: exchange of the results ---------------------------------
// Every rank holds the whole world (the other stages run redundantly and stay bit-identical) and relaxes the manifolds of
// the islands it owns.  Afterwards each rank knows the new velocity rows of its own bodies and the cached impulses of its
// own joints; everything else is unchanged.  pack: [vel | disp] rows of the bodies and [normal, friction] impulses of the
// joints as 32-bit words, ZERO for what this rank does not own, so that an integer SUM all-reduce over the ranks (exact:
// one non-zero term per word) reassembles the full state on every rank; unpack writes it back.
__global__ void __launch_bounds__(kBlock) k_island_pack(int nb, int nj, int rank, const unsigned char* __restrict__ bodyOwner, const float4* __restrict__ vel,
    const float4* __restrict__ disp, const phyx_contact_joint* __restrict__ joints, int4* __restrict__ outRows, int2* __restrict__ outJoints)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nb)
    {
        const bool mine = bodyOwner[i] == rank;
        const float4 v = vel[i], d = disp[i];
        outRows[2 * i] = mine ? make_int4(__float_as_int(v.x), __float_as_int(v.y), __float_as_int(v.z), __float_as_int(v.w)) : make_int4(0, 0, 0, 0);
        outRows[2 * i + 1] = mine ? make_int4(__float_as_int(d.x), __float_as_int(d.y), __float_as_int(d.z), __float_as_int(d.w)) : make_int4(0, 0, 0, 0);
    }
    if (i < nj)
    {
        const phyx_contact_joint j = joints[i];
        // a joint belongs to the island of its dynamic body (both ends of a dynamic pair are in the same island)
        const int owner = max(bodyOwner[j.body1Index], bodyOwner[j.body2Index]);   // static bodies count as rank 0's
        const bool mine = owner == rank;
        outJoints[i] = mine ? make_int2(__float_as_int(j.normalLimiter_accumulatedImpulse), __float_as_int(j.frictionLimiter_accumulatedImpulse)) : make_int2(0, 0);
    }
}

__global__ void __launch_bounds__(kBlock) k_island_unpack(int nb, int nj, const int4* __restrict__ rows, const int2* __restrict__ imp, float4* __restrict__ vel,
    float4* __restrict__ disp, phyx_contact_joint* __restrict__ joints)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nb)
    {
        const int4 v = rows[2 * i], d = rows[2 * i + 1];
        vel[i] = make_float4(__int_as_float(v.x), __int_as_float(v.y), __int_as_float(v.z), __int_as_float(v.w));
        disp[i] = make_float4(__int_as_float(d.x), __int_as_float(d.y), __int_as_float(d.z), __int_as_float(d.w));
    }
    if (i < nj)
    {
        joints[i].normalLimiter_accumulatedImpulse = __int_as_float(imp[i].x);
        joints[i].frictionLimiter_accumulatedImpulse = __int_as_float(imp[i].y);
    }
}

size_t islands_exchange_words(const phyx_b200_ctx* c) { return size_t(c->bodyCount) * 8 + size_t(c->jointCount) * 2; }

int islands_pack(phyx_b200_ctx* c, int32_t* deviceBuffer)
{
    const int nb = c->bodyCount, nj = c->jointCount;
    if (c->islandRanks < 2 || !c->islandsValid || c->islandBodies != nb)
    {
        set_error("island_pack: no island partition is active (island_partition with ranks > 1, then a solve)");
        return PHYX_B200_ERR_STATE;
    }
    const int n = std::max(nb, nj);
    if (n == 0) return PHYX_B200_OK;
    int4* rows = reinterpret_cast<int4*>(deviceBuffer);
    int2* imp = reinterpret_cast<int2*>(deviceBuffer + size_t(nb) * 8);
    k_island_pack<<<(n + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(nb, nj, c->islandRank, c->bodyOwner.as<unsigned char>(), c->vel.as<float4>(), c->disp.as<float4>(),
        c->joints.as<phyx_contact_joint>(), rows, imp);
    c->launches++;
    PHYX_CUDA(cudaGetLastError());
    return PHYX_B200_OK;
}

int islands_unpack(phyx_b200_ctx* c, const int32_t* deviceBuffer)
{
    const int nb = c->bodyCount, nj = c->jointCount;
    const int n = std::max(nb, nj);
    if (n == 0) return PHYX_B200_OK;
    const int4* rows = reinterpret_cast<const int4*>(deviceBuffer);
    const int2* imp = reinterpret_cast<const int2*>(deviceBuffer + size_t(nb) * 8);
    k_island_unpack<<<(n + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(nb, nj, rows, imp, c->vel.as<float4>(), c->disp.as<float4>(), c->joints.as<phyx_contact_joint>());
    c->launches++;
    PHYX_CUDA(cudaGetLastError());
    return PHYX_B200_OK;
}

} // namespace phyx
