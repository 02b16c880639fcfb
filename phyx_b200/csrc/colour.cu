// phyx_b200 — graph colouring of the joints on the device (PHYX_B200_SCHEDULE_COLOUR).
//
// Reference stage replaced here:
//   Solver::PrepareIndices   src/Solver.cpp:217-273   serial greedy grouping into SIMD-8 sets of
//                                                      joints with pairwise distinct bodies
//
// The same idea at device width: a colour is a set of joints no two of which share a DYNAMIC body
// (static bodies never change velocity, see solve.cu), so a whole colour can be relaxed in
// parallel.  The colouring is the Jones-Plassmann parallel form of first-fit greedy: every joint
// has a fixed pseudo-random priority; in each round a joint whose priority is the smallest among
// the still uncoloured joints on BOTH of its bodies takes the smallest colour not yet used on
// either body.  The result is exactly what sequential first-fit produces when it visits the joints
// in priority order, so it is deterministic (no dependence on thread timing) and checkable on the
// host; the number of rounds is the longest priority-monotone chain, O(log J) in practice.
//
// One persistent cooperative kernel runs all rounds (two grid barriers per round); the joints are
// then laid out colour-major by one stable counting-sort pass (the broadphase's radix pass on the
// colour as a 6-bit digit), each colour padded to a multiple of 32 slots.
#include "common.cuh"
#include "barrier.cuh"

#include <stdlib.h>

namespace phyx
{

constexpr int kBlock = 256;
constexpr unsigned long long kNoClaim = ~0ull;

__device__ __forceinline__ unsigned mix32(unsigned x)   // murmur3 finaliser: the joint's priority
{
    x ^= x >> 16;
    x *= 0x85ebca6bu;
    x ^= x >> 13;
    x *= 0xc2b2ae35u;
    x ^= x >> 16;
    return x;
}

// jb[j] = {body1 or -1 if static, body2 or -1 if static}; joints with no dynamic body get colour 0
// Also validates the joint's indices (result[2] = 1 + first bad joint seen): nothing on the host
// has to walk the joint array.
// keepColours: the joint cache carried last step's colours along (incremental build); only joints
// with colour -1 (new this step) still need one.
__global__ void __launch_bounds__(kBlock) k_colour_init(int nj, int nb, int ncp, const phyx_contact_joint* __restrict__ joints,
    const float4* __restrict__ params, int2* __restrict__ jb, int* __restrict__ colour, int* __restrict__ result, bool keepColours)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nj) return;
    int b1 = joints[j].body1Index, b2 = joints[j].body2Index, cp = joints[j].contactPointIndex;
    if (b1 < 0 || b1 >= nb || b2 < 0 || b2 >= nb || cp < 0 || cp >= ncp)
    {
        atomicMax(&result[2], j + 1);
        jb[j] = make_int2(-1, -1);
        colour[j] = 0;
        return;
    }
    float4 p1 = params[b1], p2 = params[b2];
    if (p1.x == 0.0f && p1.y == 0.0f) b1 = -1;
    if (p2.x == 0.0f && p2.y == 0.0f) b2 = -1;
    jb[j] = make_int2(b1, b2);
    if (b1 < 0 && b2 < 0)
        colour[j] = 0;
    else if (!keepColours)
        colour[j] = -1;
}

// static flags the colouring is built with; mismatch = 1 if a body changed class since the last full build
__global__ void __launch_bounds__(kBlock) k_colour_body_flags(Count nb, const float4* __restrict__ params, unsigned char* __restrict__ bodyStatic, bool compare,
    int* __restrict__ result)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= count_of(nb)) return;
    float4 p = params[b];
    unsigned char st = (p.x == 0.0f && p.y == 0.0f) ? 1 : 0;
    if (compare)
    {
        if (bodyStatic[b] != st) result[3] = 1;
    }
    else
        bodyStatic[b] = st;
}

struct ColourParams
{
    int nj;
    const int* njPtr;            // deferred step: the unit count lives on the device (nj is the bound)
    const int2* jb;
    int* colour;
    unsigned long long* claim;   // per body: smallest (priority, joint) among its uncoloured joints
    unsigned long long* used;    // per body: colours taken
    int* list[2];                // worklists of still uncoloured joints (ping-pong); round 0 walks all joints
    int* listCount;              // ring of 4 counters: listCount[r & 3] = length of the list round r reads
    unsigned long long* barrier;
    int* result;                 // [0] rounds, [1] overflow (a joint needed a colour >= 64)
    const int* hint;             // optional: the colour a unit takes if it is free on both bodies (else first fit)
};

__device__ __forceinline__ unsigned long long colour_key(int j) { return (static_cast<unsigned long long>(mix32(unsigned(j))) << 32) | unsigned(j); }

__global__ void __launch_bounds__(kBlock) k_colour_rounds(ColourParams P)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    unsigned epoch = 0;
    int rounds = 0;
    bool overflow = false;
    for (;;)
    {
        const int n = rounds == 0 ? (P.njPtr ? min(__ldcg(P.njPtr), P.nj) : P.nj) : __ldcg(&P.listCount[rounds & 3]);
        const int* cur = P.list[rounds & 1];
        int* next = P.list[(rounds + 1) & 1];
        int* nextCount = &P.listCount[(rounds + 1) & 3];
        if (tid == 0) *nextCount = 0;   // last read three rounds ago
        // claim: every uncoloured joint bids on its dynamic bodies
        for (int i = tid; i < n; i += nthreads)
        {
            const int j = rounds == 0 ? i : cur[i];
            if (rounds == 0 && __ldcg(&P.colour[j]) >= 0) continue;
            const int2 b = P.jb[j];
            const unsigned long long key = colour_key(j);
            if (b.x >= 0) atomicMin(&P.claim[b.x], key);
            if (b.y >= 0) atomicMin(&P.claim[b.y], key);
        }
        grid_barrier(P.barrier, epoch, false, false);
        // commit: winners of both bids take the first colour free on both bodies; the rest queue up
        bool left = false;
        for (int i0 = tid & ~31; i0 < n; i0 += nthreads)   // warp-uniform trip count
        {
            const int i = i0 + lane;
            bool lose = false;
            int j = -1;
            if (i < n)
            {
                j = rounds == 0 ? i : cur[i];
                if (!(rounds == 0 && __ldcg(&P.colour[j]) >= 0))
                {
                    const int2 b = P.jb[j];
                    const unsigned long long key = colour_key(j);
                    const bool win1 = b.x < 0 || __ldcg(&P.claim[b.x]) == key;
                    const bool win2 = b.y < 0 || __ldcg(&P.claim[b.y]) == key;
                    if (win1 && win2)
                    {
                        unsigned long long m = (b.x >= 0 ? __ldcg(&P.used[b.x]) : 0ull) | (b.y >= 0 ? __ldcg(&P.used[b.y]) : 0ull);
                        int c = __ffsll(~m) - 1;
                        if (P.hint)
                        {
                            const int h = P.hint[j];
                            if (h >= 0 && !((m >> h) & 1ull)) c = h;
                        }
                        if (c < 0)
                        {
                            c = kMaxColours - 1;   // keep going so the kernel terminates; the host falls back
                            overflow = true;
                        }
                        const unsigned long long bit = 1ull << c;
                        if (b.x >= 0)
                        {
                            __stcg(&P.used[b.x], __ldcg(&P.used[b.x]) | bit);
                            __stcg(&P.claim[b.x], kNoClaim);
                        }
                        if (b.y >= 0)
                        {
                            __stcg(&P.used[b.y], __ldcg(&P.used[b.y]) | bit);
                            __stcg(&P.claim[b.y], kNoClaim);
                        }
                        __stcg(&P.colour[j], c);
                    }
                    else
                        lose = true;
                }
            }
            // losers go to the next round's worklist (its order does not influence the result)
            const unsigned m = __ballot_sync(0xffffffffu, lose);
            if (m)
            {
                int base = 0;
                if (lane == 0) base = atomicAdd(nextCount, __popc(m));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (lose) next[base + __popc(m & ((1u << lane) - 1u))] = j;
                left = true;
            }
        }
        ++rounds;
        BarrierResult r = grid_barrier(P.barrier, epoch, left, overflow);
        overflow = r.productive;
        if (!r.wake) break;
    }
    if (tid == 0)
    {
        P.result[0] = rounds;
        P.result[1] = overflow ? 1 : 0;
    }
}

__global__ void __launch_bounds__(kBlock) k_colour_keys(int nj, const int* __restrict__ colour, uint2* __restrict__ keys, int* __restrict__ counts)
{
    __shared__ int h[kMaxColours];
    if (threadIdx.x < kMaxColours) h[threadIdx.x] = 0;
    __syncthreads();
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < nj)
    {
        int c = colour[j];
        keys[j] = make_uint2(unsigned(c), unsigned(j));
        atomicAdd(&h[c], 1);
    }
    __syncthreads();
    if (threadIdx.x < kMaxColours && h[threadIdx.x]) atomicAdd(&counts[threadIdx.x], h[threadIdx.x]);
}

// counts[64] -> level table (colour c = level c, start aligned to 32), firstPos[c] (rank base in the
// sorted run), header {numLevels, numSlots, widestLevel}
__global__ void k_colour_levels(const int* __restrict__ counts, Level* __restrict__ levels, int* __restrict__ firstPos, int* __restrict__ header)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int numLevels = 0;
    for (int c = 0; c < kMaxColours; ++c)
        if (counts[c] > 0) numLevels = c + 1;
    int cursor = 0, run = 0, widest = 0;
    for (int c = 0; c < numLevels; ++c)
    {
        levels[c].start = cursor;
        levels[c].grouped_end = cursor;
        levels[c].end = cursor + counts[c];
        firstPos[c] = run;
        run += counts[c];
        widest = max(widest, counts[c]);
        cursor = (levels[c].end + 31) & ~31;
    }
    header[0] = numLevels;
    header[1] = cursor;
    header[2] = widest;
}

__global__ void __launch_bounds__(kBlock) k_colour_place(int nj, const uint2* __restrict__ sorted, const Level* __restrict__ levels,
    const int* __restrict__ firstPos, int* __restrict__ slotJoint)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nj) return;
    uint2 e = sorted[p];
    slotJoint[levels[e.x].start + (p - firstPos[e.x])] = int(e.y);
}

// Build the colour schedule for the resident joints.  Returns PHYX_B200_ERR_CAPACITY if more than
// 64 colours are needed (a dynamic body with dozens of joints); the caller then uses the host builder.
static int colour_joints_build(phyx_b200_ctx* c, bool* staticsChanged);
static int colour_units_build(phyx_b200_ctx* c, bool incremental, bool* staticsChanged);

// Two colourings.
//  * Units (resident pipeline): the unit is a MANIFOLD, i.e. the (up to two) joints of one body pair.
//    Colouring manifolds needs about half the colours of colouring joints (a body has half as many
//    manifolds as joints), and the two joints of a manifold go to the two halves of one FUSED level: the
//    same thread relaxes joint A then joint B with no barrier in between (B's bodies are touched by no
//    other joint of that level, so this equals the colour-major sequential order).  The colouring is
//    incremental: manifolds keep their colour from step to step (the manifold cache carries it through
//    PackManifolds, which also hands the colours of removed manifolds back to their bodies), the per-body
//    colour masks persist, and only manifolds without a colour go through the rounds.  A full rebuild
//    happens on first use, when a body changes between static and dynamic, or when the colour count has
//    drifted upwards.
//  * Joints (host-array API, no manifold information): every joint is a unit; always a full build.
int colour_schedule_build(phyx_b200_ctx* c)
{
    bool changed = false;
    if (!c->jointUnitsValid || c->manifoldCount == 0) return colour_joints_build(c, &changed);
    bool incremental = c->colourStateValid && c->colourStateBodies == c->bodyCount && c->manColour.ptr && c->bodyUsed.ptr;
    int st = colour_units_build(c, incremental, &changed);
    // colours in use over ALL coloured manifolds (not only the ones this rank lays out: an island partition must take the same
    // rebuild decisions as the one-device run, or the incremental colourings drift apart)
    const int colours = c->part.ranks > 1 ? c->partColours : c->coloursInUse;
    // every colour is a level = a grid barrier and a latency chain per pass (~8-10 us x 22 passes on the bench scene),
    // a full rebuild costs about a millisecond once: rebuild as soon as the incremental colouring has drifted two
    // colours above the last full build (kColourDrift; a deferred step takes the same decision on the device and stops)
    if (c->def.active) return st;
    if (st == PHYX_B200_OK && incremental && (changed || colours > c->coloursAtFullBuild + kColourDrift))
        st = colour_units_build(c, false, &changed);
    return st;
}

static int colour_joints_build(phyx_b200_ctx* c, bool* staticsChanged)
{
    const bool incremental = false;
    c->strip.valid = false;
    c->hostLevelsStale = false;

    const int nj = c->jointCount, nb = c->bodyCount;
    c->hostSlots.clear();
    c->hostSlotPos.clear();
    c->hostLevels.clear();
    c->slotPosValid = false;
    c->slotCount = c->levelCount = 0;
    c->strictLevelCount = c->numMultiStatics = 0;
    if (nj == 0) return PHYX_B200_OK;

    const size_t nb1 = size_t(nb > 0 ? nb : 1);
    // scratch: jb[nj] int2 | colour[nj] int | claim[nb] u64 | used[nb] u64 | counts[64] | firstPos[64] | header[4] | result[4] | barrier[4] u64
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~size_t(255); return o; };
    const size_t oJb = take(size_t(nj) * sizeof(int2)), oColour = take(size_t(nj) * sizeof(int)), oClaim = take(nb1 * 8),
                 oList0 = take(size_t(nj) * sizeof(int)), oList1 = take(size_t(nj) * sizeof(int)),
                 oCounts = take(kMaxColours * sizeof(int)), oFirst = take(kMaxColours * sizeof(int)), oHeader = take(16), oResult = take(16),
                 oBarrier = take(32), oListCount = take(16);
    PHYX_TRY(c->colourTmp.reserve(off));
    char* base = c->colourTmp.as<char>();
    int2* jb = reinterpret_cast<int2*>(base + oJb);
    PHYX_TRY(c->bodyUsed.reserve(nb1 * 8));
    PHYX_TRY(c->bodyStatic.reserve(nb1));
    int* colour = reinterpret_cast<int*>(base + oColour);
    unsigned long long* claim = reinterpret_cast<unsigned long long*>(base + oClaim);
    unsigned long long* used = c->bodyUsed.as<unsigned long long>();
    int* counts = reinterpret_cast<int*>(base + oCounts);
    int* firstPos = reinterpret_cast<int*>(base + oFirst);
    int* header = reinterpret_cast<int*>(base + oHeader);
    int* result = reinterpret_cast<int*>(base + oResult);
    unsigned long long* barrier = reinterpret_cast<unsigned long long*>(base + oBarrier);

    PHYX_CUDA(cudaMemsetAsync(claim, 0xff, nb1 * 8, c->stream));
    if (!incremental) PHYX_CUDA(cudaMemsetAsync(used, 0, nb1 * 8, c->stream));
    PHYX_CUDA(cudaMemsetAsync(base + oCounts, 0, off - oCounts, c->stream));   // counts .. barrier

    const int grid = (nj + kBlock - 1) / kBlock;
    if (nb > 0)
    {
        k_colour_body_flags<<<(nb + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(c->count(nb, &StepCtl::bodies), c->params.as<float4>(),
            c->bodyStatic.as<unsigned char>(), incremental, result);
        c->launches++;
    }
    k_colour_init<<<grid, kBlock, 0, c->stream>>>(nj, nb, c->contactPointCount, c->joints.as<phyx_contact_joint>(), c->params.as<float4>(), jb, colour,
        result, incremental);
    c->launches++;

    if (c->colourBlocksPerSM == 0)
    {
        int per = 0;
        PHYX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_colour_rounds, kBlock, 0));
        c->colourBlocksPerSM = per > 0 ? per : 1;
    }
    ColourParams P = { nj, nullptr, jb, colour, claim, used, { reinterpret_cast<int*>(base + oList0), reinterpret_cast<int*>(base + oList1) },
        reinterpret_cast<int*>(base + oListCount), barrier, result, nullptr };
    int cgrid = std::max(1, std::min(grid, c->numSMs * c->colourBlocksPerSM));
    void* args[] = { &P };
    PHYX_CUDA(cudaLaunchCooperativeKernel((void*)k_colour_rounds, dim3(cgrid), dim3(kBlock), args, 0, c->stream));
    c->launches++;

    // colour-major layout: stable counting sort of {colour, joint} on a 6-bit digit
    PHYX_TRY(c->colourKeys.reserve(size_t(nj) * sizeof(uint2)));
    PHYX_TRY(c->colourSorted.reserve(size_t(nj) * sizeof(uint2)));
    k_colour_keys<<<grid, kBlock, 0, c->stream>>>(nj, colour, c->colourKeys.as<uint2>(), counts);
    c->launches++;
    PHYX_TRY(radix_pass(c, c->colourKeys.as<uint2>(), c->colourSorted.as<uint2>(), nj, 0, kMaxColours));
    PHYX_TRY(c->levels.reserve(kMaxColours * sizeof(Level)));
    k_colour_levels<<<1, 32, 0, c->stream>>>(counts, c->levels.as<Level>(), firstPos, header);
    c->launches++;
    const size_t maxSlots = size_t(nj) + 32 * kMaxColours;
    PHYX_TRY(c->slotJoint.reserve(maxSlots * sizeof(int)));
    PHYX_CUDA(cudaMemsetAsync(c->slotJoint.ptr, 0xff, maxSlots * sizeof(int), c->stream));
    k_colour_place<<<grid, kBlock, 0, c->stream>>>(nj, c->colourSorted.as<uint2>(), c->levels.as<Level>(), firstPos, c->slotJoint.as<int>());
    c->launches++;
    PHYX_CUDA(cudaGetLastError());

    struct { int header[4]; int result[4]; } host;
    PHYX_CUDA(cudaMemcpyAsync(host.header, header, 16, cudaMemcpyDeviceToHost, c->stream));
    PHYX_CUDA(cudaMemcpyAsync(host.result, result, 16, cudaMemcpyDeviceToHost, c->stream));
    Level lv[kMaxColours];
    PHYX_CUDA(cudaMemcpyAsync(lv, c->levels.ptr, sizeof(lv), cudaMemcpyDeviceToHost, c->stream));
    PHYX_CUDA(cudaStreamSynchronize(c->stream));
    if (host.result[2])
    {
        set_error("solve: joint %d references a body outside [0,%d) or a contact point outside [0,%d)", host.result[2] - 1, nb, c->contactPointCount);
        return PHYX_B200_ERR_ARGUMENT;
    }
    if (host.result[1])
    {
        set_error("colouring needs more than %d colours", kMaxColours);
        return PHYX_B200_ERR_CAPACITY;
    }
    c->levelCount = host.header[0];
    c->slotCount = host.header[1];
    c->colourRounds = host.result[0];
    *staticsChanged = host.result[3] != 0;
    c->colourStateValid = false;   // per-joint colours are not carried from step to step
    c->hostLevels.assign(lv, lv + c->levelCount);
    c->hostSlotsStale = true;
    return PHYX_B200_OK;
}

// ---- manifold units -----------------------------------------------------------------------------------

constexpr int kSkipUnit = kMaxColours;   // working colour of a manifold without contact points

// Experimental (PHYX_COLOUR_HINTS=1): colour hint of a manifold from geometry: layer parity of the lower body and
// the side the upper one sits on.  On a regular brick pile these four combinations never meet at a body.
__device__ __forceinline__ int unit_hint(float4 p1, float2 s1, float4 p2, float2 s2)
{
    const bool static1 = p1.x == 0.0f && p1.y == 0.0f, static2 = p2.x == 0.0f && p2.y == 0.0f;
    const bool swap = static2 ? !static1 : (!static1 && (p2.w < p1.w || (p2.w == p1.w && p2.z < p1.z)));
    const float4 lo = swap ? p2 : p1, hi = swap ? p1 : p2;
    const float2 slo = swap ? s2 : s1;
    if (static1 || static2)
    {
        const float h = fmaxf(2.0f * (swap ? s1 : s2).y, 1e-3f);
        const int layer = int(floorf(hi.w / h)) & 1;
        return 2 * (1 - layer);
    }
    const float h = fmaxf(2.0f * slo.y, 1e-3f);
    const int layer = int(floorf(lo.w / h)) & 1;
    return 2 * layer + (hi.z > lo.z ? 1 : 0);
}

__global__ void __launch_bounds__(kBlock) k_unit_init(Count Mc, const int2* __restrict__ manBody, const int* __restrict__ manCount,
    const float4* __restrict__ params, const float2* __restrict__ size, int* __restrict__ manColour, int2* __restrict__ jb, int* __restrict__ work,
    int* __restrict__ hint, unsigned long long* __restrict__ bodyUsed, bool keepColours)
{
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= count_of(Mc)) return;
    int2 b = manBody[m];
    const float4 p1 = params[b.x], p2 = params[b.y];
    if (hint) hint[m] = unit_hint(p1, size[b.x], p2, size[b.y]);
    if (p1.x == 0.0f && p1.y == 0.0f) b.x = -1;
    if (p2.x == 0.0f && p2.y == 0.0f) b.y = -1;
    jb[m] = b;
    const int had = keepColours ? manColour[m] : -1;
    if (manCount[m] == 0)
    {
        // lost its last contact: give the colour back to the bodies
        if (had >= 0)
        {
            const unsigned long long keep = ~(1ull << had);
            if (b.x >= 0) atomicAnd(&bodyUsed[b.x], keep);
            if (b.y >= 0) atomicAnd(&bodyUsed[b.y], keep);
        }
        manColour[m] = -1;
        work[m] = kSkipUnit;
    }
    else if (b.x < 0 && b.y < 0)
        work[m] = 0;   // no dynamic body: conflicts with nothing
    else
        work[m] = had;
}

// persist the colours, emit {colour, manifold} sort keys and the per-colour histogram (bin 64 = skipped)
__global__ void __launch_bounds__(kBlock) k_unit_keys(int M, const int* __restrict__ work, int* __restrict__ manColour, uint2* __restrict__ keys,
    int* __restrict__ counts, const int2* __restrict__ manBody, const unsigned char* __restrict__ bodyOwner, int rank)
{
    __shared__ int h[kMaxColours + 1];
    if (threadIdx.x <= kMaxColours) h[threadIdx.x] = 0;
    __syncthreads();
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m < M)
    {
        int c = work[m];
        manColour[m] = (c == kSkipUnit) ? -1 : c;
        if (c < kMaxColours) atomicMax(&counts[kMaxColours + 1], c + 1);   // colours in use, whoever owns the manifold
        // island partition (islands.cu): manifolds of another rank's islands keep their colour but are not laid out
        if (bodyOwner && max(bodyOwner[manBody[m].x], bodyOwner[manBody[m].y]) != rank) c = kSkipUnit;
        keys[m] = make_uint2(unsigned(c), unsigned(m));
        atomicAdd(&h[c], 1);
    }
    __syncthreads();
    if (threadIdx.x <= kMaxColours && h[threadIdx.x]) atomicAdd(&counts[threadIdx.x], h[threadIdx.x]);
}

// counts[64] -> paired level table: level k = {start, -1, start + 2*count}, start a multiple of 64; slots
// 2r and 2r+1 of a level are the first and second joint of its r-th manifold (solve.cu, paired levels)
__global__ void k_unit_levels(const int* __restrict__ counts, Level* __restrict__ levels, int* __restrict__ firstPos, int* __restrict__ header)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int numLevels = 0;
    for (int c = 0; c < kMaxColours; ++c)
        if (counts[c] > 0) numLevels = c + 1;
    int cursor = 0, run = 0, widest = 0;
    for (int c = 0; c < numLevels; ++c)
    {
        levels[c].start = cursor;
        levels[c].grouped_end = -1;
        levels[c].end = cursor + 2 * counts[c];
        firstPos[c] = run;
        run += counts[c];
        widest = max(widest, 2 * counts[c]);
        cursor = (levels[c].end + 63) & ~63;
    }
    header[0] = numLevels;
    header[1] = cursor;
    header[2] = widest;
}

__global__ void __launch_bounds__(kBlock) k_unit_place(int M, const uint2* __restrict__ sorted, const Level* __restrict__ levels,
    const int* __restrict__ firstPos, const int* __restrict__ manCount, const float4* __restrict__ contactPoints, int* __restrict__ slotJoint, unsigned skipBin)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= M) return;
    const uint2 e = sorted[p];
    if (e.x >= skipBin) return;   // manifold without contact points
    const int m = int(e.y);
    const int slot = levels[e.x].start + 2 * (p - firstPos[e.x]);
    // the joints of a manifold are the solverIndex of its contact points (World.cpp:100-103,140)
    slotJoint[slot] = __float_as_int(contactPoints[size_t(2 * m) * 2 + 1].w);
    slotJoint[slot + 1] = manCount[m] > 1 ? __float_as_int(contactPoints[size_t(2 * m + 1) * 2 + 1].w) : -1;
}


// ---- partitioned layout (one world over several devices, partition.cu) -----------------------------------
// The solver rows are in sorted-x order, so a contiguous row range is a vertical strip of the world.  Rank q
// owns the rows [cuts[q], cuts[q+1]); the cuts balance the number of manifolds per rank (a manifold counts
// for its lower dynamic row).  A manifold whose dynamic bodies all belong to rank q is INTERIOR to q
// (class q); one whose two bodies belong to different ranks is CUT (class `ranks`).  Slots are laid out
// class-major, colour-minor: each rank's interior manifolds are one contiguous slot range holding its colour
// levels, the cut manifolds follow with theirs.  Read as ONE sequential order this is a valid Gauss-Seidel
// sweep (interior classes are mutually independent), which is what the oracle checks the devices against.

__device__ __forceinline__ int part_rank_of(const int* __restrict__ cuts, int ranks, int row)
{
    int k = 0;
    for (int q = 1; q < ranks; ++q) k += (cuts[q] <= row) ? 1 : 0;
    return k;
}

// hist[row] = manifolds (with a colour) whose lower dynamic row it is
__global__ void __launch_bounds__(kBlock) k_part_hist(int M, const int2* __restrict__ jb, const int* __restrict__ work, const int* __restrict__ rowOf,
    int* __restrict__ hist)
{
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    if (work[m] >= kMaxColours) return;
    const int2 b = jb[m];
    int r1 = b.x < 0 ? -1 : (rowOf ? rowOf[b.x] : b.x), r2 = b.y < 0 ? -1 : (rowOf ? rowOf[b.y] : b.y);
    int home = r1 < 0 ? r2 : (r2 < 0 ? r1 : min(r1, r2));
    if (home >= 0) atomicAdd(&hist[home], 1);
}

// cuts[q] = first row with at least q/ranks of the manifolds before it; plan[0..ranks] = cuts
__global__ void k_part_cuts(int nb, int ranks, const int* __restrict__ prefix, const int* __restrict__ total, int* __restrict__ cuts)
{
    const int q = threadIdx.x;
    if (q > ranks) return;
    if (q == 0) { cuts[0] = 0; return; }
    if (q == ranks) { cuts[q] = nb; return; }
    const long long target = (static_cast<long long>(*total) * q + ranks - 1) / ranks;
    int lo = 0, hi = nb;   // smallest row r in [0, nb] with prefix[r] >= target (prefix[nb] := total)
    while (lo < hi)
    {
        const int mid = (lo + hi) >> 1;
        if (prefix[mid] >= target) hi = mid; else lo = mid + 1;
    }
    cuts[q] = lo;
}

// persist the colours, classify, emit {class * 64 + colour, manifold} sort keys and the histogram over bins,
// flag the rows cut manifolds touch (the boundary rows exchanged between the ranks)
__global__ void __launch_bounds__(kBlock) k_part_keys(int M, const int2* __restrict__ jb, const int* __restrict__ work, const int* __restrict__ rowOf,
    const int* __restrict__ cuts, int ranks, int* __restrict__ manColour, uint2* __restrict__ keys, int* __restrict__ counts, int* __restrict__ rowFlag)
{
    __shared__ int h[(kMaxRanks + 1) * kMaxColours + 1];
    const int bins1 = (ranks + 1) * kMaxColours + 1;
    for (int b = threadIdx.x; b < bins1; b += blockDim.x) h[b] = 0;
    __syncthreads();
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = m < M ? work[m] : kMaxColours;
    if (m < M) manColour[m] = (c >= kMaxColours) ? -1 : c;
    int bin = (ranks + 1) * kMaxColours;   // skipped
    if (m < M && c < kMaxColours)
    {
        const int2 b = jb[m];
        const int r1 = b.x < 0 ? -1 : (rowOf ? rowOf[b.x] : b.x), r2 = b.y < 0 ? -1 : (rowOf ? rowOf[b.y] : b.y);
        int cls = 0;
        if (r1 >= 0 && r2 >= 0)
        {
            const int k1 = part_rank_of(cuts, ranks, r1), k2 = part_rank_of(cuts, ranks, r2);
            if (k1 == k2)
                cls = k1;
            else
            {
                cls = ranks;
                rowFlag[r1] = 1;
                rowFlag[r2] = 1;
            }
        }
        else if (r1 >= 0 || r2 >= 0)
            cls = part_rank_of(cuts, ranks, r1 >= 0 ? r1 : r2);
        bin = cls * kMaxColours + c;
    }
    if (m < M)
    {
        keys[m] = make_uint2(unsigned(bin), unsigned(m));
        atomicAdd(&h[bin], 1);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < bins1; b += blockDim.x)
        if (h[b]) atomicAdd(&counts[b], h[b]);
}

// counts[bins] -> paired level table over ALL bins (empty bins are empty levels), firstPos, header {bins, numSlots, widest}
__global__ void k_part_levels(int bins, const int* __restrict__ counts, Level* __restrict__ levels, int* __restrict__ firstPos, int* __restrict__ header)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int cursor = 0, run = 0, widest = 0;
    for (int b = 0; b < bins; ++b)
    {
        levels[b].start = cursor;
        levels[b].grouped_end = -1;
        levels[b].end = cursor + 2 * counts[b];
        firstPos[b] = run;
        run += counts[b];
        widest = max(widest, 2 * counts[b]);
        if (counts[b] > 0) cursor = (levels[b].end + 63) & ~63;
    }
    header[0] = bins;
    header[1] = cursor;
    header[2] = widest;
}

__global__ void __launch_bounds__(kBlock) k_part_blist(int nb, const int* __restrict__ rowFlag, const int* __restrict__ rowPrefix, int* __restrict__ bRows)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nb) return;
    if (rowFlag[r]) bRows[rowPrefix[r]] = r;
}

// plan words: [0..ranks] cuts, [16..16+ranks] first boundary row of each rank in bRows, [16+ranks] = total
__global__ void k_part_bstart(int nb, int ranks, const int* __restrict__ rowPrefix, const int* __restrict__ total, int* __restrict__ plan)
{
    const int q = threadIdx.x;
    if (q > ranks) return;
    const int row = plan[q];
    plan[16 + q] = row >= nb ? *total : rowPrefix[row];
}

// Layout of the coloured manifolds for a partitioned solve; `work` holds the colours.  Fills the context's
// schedule (slots, level table over all bins) and the partition plan.
static int part_layout(phyx_b200_ctx* c, const int2* jb, const int* work, int* result, bool incremental, bool* staticsChanged, int* coloursOut)
{
    Partition& pt = c->part;
    const int M = c->manifoldCount, nb = c->bodyCount, R = pt.ranks;
    const int bins = (R + 1) * kMaxColours;
    const int grid = (M + kBlock - 1) / kBlock;
    const int* rowOf = (c->rowOrderValid && c->rowOrderBodies == nb) ? c->rowOf.as<int>() : nullptr;
    pt.planValid = false;

    PHYX_TRY(pt.rowFlag.reserve(size_t(nb + 1) * sizeof(int)));
    PHYX_TRY(pt.rowPrefix.reserve(size_t(nb + 1) * sizeof(int)));
    PHYX_TRY(pt.bRows.reserve(size_t(nb + 1) * sizeof(int)));
    PHYX_TRY(pt.planWords.reserve((64 + size_t(bins) + 1 + bins) * sizeof(int)));
    int* plan = pt.planWords.as<int>();        // [0..15] cuts, [16..31] bStart, [32..35] header, [36] scan total, [64..] counts[bins + 1], then firstPos[bins]
    int* header = plan + 32;
    int* total = plan + 36;
    int* counts = plan + 64;
    int* firstPos = counts + bins + 1;
    PHYX_CUDA(cudaMemsetAsync(plan, 0, (64 + size_t(bins) + 1 + bins) * sizeof(int), c->stream));

    // cuts that balance the manifold count
    PHYX_CUDA(cudaMemsetAsync(pt.rowFlag.ptr, 0, size_t(nb + 1) * sizeof(int), c->stream));
    k_part_hist<<<grid, kBlock, 0, c->stream>>>(M, jb, work, rowOf, pt.rowFlag.as<int>());
    PHYX_TRY(exclusive_scan_i32(c, pt.rowFlag.as<int>(), pt.rowPrefix.as<int>(), nb, total));
    k_part_cuts<<<1, 32, 0, c->stream>>>(nb, R, pt.rowPrefix.as<int>(), total, plan);
    c->launches += 2;

    // classes, sort keys, boundary flags
    PHYX_CUDA(cudaMemsetAsync(pt.rowFlag.ptr, 0, size_t(nb + 1) * sizeof(int), c->stream));
    PHYX_TRY(c->colourKeys.reserve(size_t(M) * sizeof(uint2)));
    PHYX_TRY(c->colourSorted.reserve(size_t(M) * sizeof(uint2)));
    k_part_keys<<<grid, kBlock, 0, c->stream>>>(M, jb, work, rowOf, plan, R, c->manColour.as<int>(), c->colourKeys.as<uint2>(), counts, pt.rowFlag.as<int>());
    c->launches++;
    int digits = 1;
    while (digits < bins + 1) digits <<= 1;
    PHYX_TRY(radix_pass(c, c->colourKeys.as<uint2>(), c->colourSorted.as<uint2>(), M, 0, digits));
    PHYX_TRY(pt.binLevels.reserve(size_t(bins) * sizeof(Level)));
    k_part_levels<<<1, 32, 0, c->stream>>>(bins, counts, pt.binLevels.as<Level>(), firstPos, header);
    c->launches++;
    const size_t maxSlots = 2 * size_t(M) + 64 * size_t(bins);
    PHYX_TRY(c->slotJoint.reserve(maxSlots * sizeof(int)));
    PHYX_CUDA(cudaMemsetAsync(c->slotJoint.ptr, 0xff, maxSlots * sizeof(int), c->stream));
    k_unit_place<<<grid, kBlock, 0, c->stream>>>(M, c->colourSorted.as<uint2>(), pt.binLevels.as<Level>(), firstPos, c->manCount.as<int>(),
        c->contactPoints.as<float4>(), c->slotJoint.as<int>(), unsigned(bins));
    c->launches++;

    // boundary rows, in row order: each rank's own ones are one contiguous run
    PHYX_TRY(exclusive_scan_i32(c, pt.rowFlag.as<int>(), pt.rowPrefix.as<int>(), nb, total));
    k_part_blist<<<(nb + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(nb, pt.rowFlag.as<int>(), pt.rowPrefix.as<int>(), pt.bRows.as<int>());
    k_part_bstart<<<1, 32, 0, c->stream>>>(nb, R, pt.rowPrefix.as<int>(), total, plan);
    c->launches += 2;
    PHYX_CUDA(cudaGetLastError());

    std::vector<int> host(64 + size_t(bins) + 1);
    int res[4];
    PHYX_CUDA(cudaMemcpyAsync(host.data(), plan, host.size() * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PHYX_CUDA(cudaMemcpyAsync(res, result, 16, cudaMemcpyDeviceToHost, c->stream));
    PHYX_CUDA(cudaStreamSynchronize(c->stream));
    if (res[1])
    {
        set_error("colouring needs more than %d colours", kMaxColours);
        return PHYX_B200_ERR_CAPACITY;
    }
    c->colourRounds = res[0];
    *staticsChanged = res[3] != 0;
    c->colourStateValid = true;
    c->colourStateBodies = nb;

    // host copies: global level list (non-empty bins, in slot order), this rank's level table, slot ranges per class
    const int* hc = host.data() + 64;
    std::vector<Level> mine, cutLevels;
    int cursor = 0, colours = 0;
    pt.widestInterior = pt.widestCut = 0;
    for (int q = 0; q <= R; ++q)
    {
        pt.classSlotStart[q] = cursor;
        for (int k = 0; k < kMaxColours; ++k)
        {
            const int n = hc[q * kMaxColours + k];
            if (n == 0) continue;
            const Level L = { cursor, -1, cursor + 2 * n };
            c->hostLevels.push_back(L);
            if (q == pt.rank)
            {
                mine.push_back(L);
                pt.widestInterior = std::max(pt.widestInterior, 2 * n);
            }
            if (q == R)
            {
                cutLevels.push_back(L);
                pt.widestCut = std::max(pt.widestCut, 2 * n);
            }
            colours = std::max(colours, k + 1);
            cursor = (cursor + 2 * n + 63) & ~63;
        }
    }
    pt.classSlotStart[R + 1] = cursor;
    c->slotCount = cursor;
    c->levelCount = int(c->hostLevels.size());
    if (!incremental) c->coloursAtFullBuild = colours;
    pt.numInterior = int(mine.size());
    pt.numCut = int(cutLevels.size());
    mine.insert(mine.end(), cutLevels.begin(), cutLevels.end());
    PHYX_TRY(pt.partLevels.reserve((mine.size() + 1) * sizeof(Level)));
    if (!mine.empty()) PHYX_CUDA(cudaMemcpyAsync(pt.partLevels.ptr, mine.data(), mine.size() * sizeof(Level), cudaMemcpyHostToDevice, c->stream));
    // the whole order as one level table: what a single device executes when it runs this schedule alone
    PHYX_TRY(c->levels.reserve((c->hostLevels.size() + 1) * sizeof(Level)));
    if (!c->hostLevels.empty())
        PHYX_CUDA(cudaMemcpyAsync(c->levels.ptr, c->hostLevels.data(), c->hostLevels.size() * sizeof(Level), cudaMemcpyHostToDevice, c->stream));
    PHYX_CUDA(cudaStreamSynchronize(c->stream));   // `mine` is pageable stack memory
    for (int q = 0; q <= R; ++q)
    {
        pt.cuts[q] = host[q];
        pt.bStart[q] = host[16 + q];
    }
    pt.planValid = true;
    c->hostSlotsStale = true;
    *coloursOut = colours;
    return PHYX_B200_OK;
}

static int colour_units_build(phyx_b200_ctx* c, bool incremental, bool* staticsChanged)
{
    const int M = c->manifoldCount, nb = c->bodyCount, nj = c->jointCount;
    c->hostSlots.clear();
    c->hostSlotPos.clear();
    c->hostLevels.clear();
    c->slotPosValid = false;
    c->slotCount = c->levelCount = 0;
    c->strictLevelCount = c->numMultiStatics = 0;
    if (nj == 0 || M == 0) return PHYX_B200_OK;

    const size_t nb1 = size_t(nb > 0 ? nb : 1);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~size_t(255); return o; };
    const size_t oJb = take(size_t(M) * sizeof(int2)), oWork = take(size_t(M) * sizeof(int)), oHint = take(size_t(M) * sizeof(int)), oClaim = take(nb1 * 8),
                 oList0 = take(size_t(M) * sizeof(int)), oList1 = take(size_t(M) * sizeof(int)),
                 oCounts = take((kMaxColours + 2) * sizeof(int)), oFirst = take(kMaxColours * sizeof(int)), oHeader = take(16), oResult = take(16),
                 oBarrier = take(32), oListCount = take(16);
    PHYX_TRY(c->colourTmp.reserve(off));
    char* base = c->colourTmp.as<char>();
    int2* jb = reinterpret_cast<int2*>(base + oJb);
    int* work = reinterpret_cast<int*>(base + oWork);
    PHYX_TRY(c->manColour.reserve_keep(size_t(M) * sizeof(int), incremental ? size_t(M) * sizeof(int) : 0, c->stream));
    PHYX_TRY(c->bodyUsed.reserve_keep(nb1 * 8, incremental ? nb1 * 8 : 0, c->stream));
    PHYX_TRY(c->bodyStatic.reserve_keep(nb1, incremental ? nb1 : 0, c->stream));
    unsigned long long* claim = reinterpret_cast<unsigned long long*>(base + oClaim);
    unsigned long long* used = c->bodyUsed.as<unsigned long long>();
    int* counts = reinterpret_cast<int*>(base + oCounts);
    int* firstPos = reinterpret_cast<int*>(base + oFirst);
    int* header = reinterpret_cast<int*>(base + oHeader);
    int* result = reinterpret_cast<int*>(base + oResult);
    unsigned long long* barrier = reinterpret_cast<unsigned long long*>(base + oBarrier);

    PHYX_CUDA(cudaMemsetAsync(claim, 0xff, nb1 * 8, c->stream));
    if (!incremental) PHYX_CUDA(cudaMemsetAsync(used, 0, nb1 * 8, c->stream));
    PHYX_CUDA(cudaMemsetAsync(base + oCounts, 0, off - oCounts, c->stream));   // counts .. list counters

    const int grid = (M + kBlock - 1) / kBlock;
    if (nb > 0)
    {
        k_colour_body_flags<<<(nb + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(c->count(nb, &StepCtl::bodies), c->params.as<float4>(),
            c->bodyStatic.as<unsigned char>(), incremental, result);
        c->launches++;
    }
    int* hint = nullptr;   // (geometric colour hints: a measured round-1 experiment, not used)
    (void)oHint;
    k_unit_init<<<grid, kBlock, 0, c->stream>>>(c->count(M, &StepCtl::manifolds), c->manBody.as<int2>(), c->manCount.as<int>(), c->params.as<float4>(), c->size.as<float2>(),
        c->manColour.as<int>(), jb, work, hint, used, incremental);
    c->launches++;

    if (c->colourBlocksPerSM == 0)
    {
        int per = 0;
        PHYX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_colour_rounds, kBlock, 0));
        c->colourBlocksPerSM = per > 0 ? per : 1;
    }
    ColourParams P = { M, c->def.active ? &c->ctl()->manifolds : nullptr, jb, work, claim, used, { reinterpret_cast<int*>(base + oList0), reinterpret_cast<int*>(base + oList1) },
        reinterpret_cast<int*>(base + oListCount), barrier, result, hint };
    int cgrid = std::max(1, std::min(grid, c->numSMs * c->colourBlocksPerSM));
    void* args[] = { &P };
    PHYX_CUDA(cudaLaunchCooperativeKernel((void*)k_colour_rounds, dim3(cgrid), dim3(kBlock), args, 0, c->stream));
    c->launches++;

    if (c->part.ranks > 1)
    {
        c->strip.valid = false;
        return part_layout(c, jb, work, result, incremental, staticsChanged, &c->partColours);
    }

    // strip layout (strips.cu): class-major over S row ranges, run by the strip-local kernel; rejected layouts
    // (a manifold across non-adjacent strips, a strip too large for shared memory) fall through to colour-major
    c->strip.valid = false;
    c->hostLevelsStale = false;
    // (deferred step: the strip count of the previous layout; the verdict on the layout is taken on the device, strips.cu)
    int S = c->def.active ? c->strip.strips : c->forceKernelForm == 1 || c->forceKernelForm == 2 ? 0 : strip_choose(c, M, nb);
    if (S > 0)
    {
        int res[4] = { 0, 0, 0, 0 };
        bool usable = false;
        for (;;)
        {
            PHYX_TRY(strip_layout(c, S, jb, work, result, res, &usable));   // waits for the layout header and the colouring's result words
            if (c->def.active)
            {
                // bounds until the step's counts come home (deferred_finish, api.cu)
                c->def.colourResult = result;
                c->slotCount = 2 * M;
                c->levelCount = std::max(1, c->strip.colours);   // (the previous layout's; the solve only asks whether there are any)
                c->colourStateValid = true;
                c->colourStateBodies = nb;
                c->hostSlotsStale = true;
                c->hostLevelsStale = true;
                return PHYX_B200_OK;
            }
            // strips narrower than the bodies' reach (a manifold across non-adjacent strips, a row in two cut sets): try
            // half as many, and remember what worked for the next steps of this world (one strip always works, if it fits)
            // too large for shared memory only (the side buffers grew, e.g. larger bins after a full recolouring): once
            // more with the rows that the side buffers just measured leave room for, instead of a step on the fallback forms
            if (!usable && c->strip.rejected == 4 && c->strip.rowLimitForce == 0 && c->strip.maxStripRows > 0)
            {
                const long long side = (long long)(c->strip.maxCutRows + 64) * 24 + (long long)(c->strip.maxBin + 64) * 2;
                const long long rows = ((227ll * 1024 - 4096 - side) / 16 - 32) & ~63ll;
                if (rows >= (nb + S - 1) / S + 64 && rows < c->strip.maxStripRows)
                {
                    c->strip.rowLimitForce = int(rows);
                    continue;
                }
            }
            if (usable || c->strip.want > 0 || S == 1 || !(c->strip.rejected & 3)) break;
            S = std::max(1, S / 2);
            c->strip.autoLimit = S;
            c->strip.limitAge = 0;
        }
        c->strip.rowLimitForce = 0;
        if (res[1])
        {
            c->strip.valid = false;   // the layout was built on colours that do not exist
            c->hostLevelsStale = false;
            set_error("colouring needs more than %d colours", kMaxColours);
            return PHYX_B200_ERR_CAPACITY;
        }
        if (usable)
        {
            strip_limit_recover(c);
            c->slotCount = 2 * c->strip.manifolds;
            c->levelCount = c->strip.colours;
            c->coloursInUse = c->strip.colours;
            c->colourRounds = res[0];
            *staticsChanged = res[3] != 0;
            c->colourStateValid = true;
            c->colourStateBodies = nb;
            if (!incremental) c->coloursAtFullBuild = c->coloursInUse;
            c->hostSlotsStale = true;
            c->hostLevelsStale = true;
            return PHYX_B200_OK;
        }
        if (c->forceKernelForm == 3)
        {
            set_error("strip layout rejected (reason mask %d: 1 manifold across non-adjacent strips, 2 row in two cut sets, 4 shared memory, 16 no manifolds)",
                c->strip.rejected);
            return PHYX_B200_ERR_STATE;
        }
    }

    // colour-major layout of the manifolds: stable counting sort on a 7-bit digit (bin 64 = skipped)
    PHYX_TRY(c->colourKeys.reserve(size_t(M) * sizeof(uint2)));
    PHYX_TRY(c->colourSorted.reserve(size_t(M) * sizeof(uint2)));
    const bool split = c->islandRanks > 1 && c->islandsValid && c->islandBodies == nb;
    k_unit_keys<<<grid, kBlock, 0, c->stream>>>(M, work, c->manColour.as<int>(), c->colourKeys.as<uint2>(), counts, c->manBody.as<int2>(),
        split ? c->bodyOwner.as<unsigned char>() : nullptr, c->islandRank);
    c->launches++;
    PHYX_TRY(radix_pass(c, c->colourKeys.as<uint2>(), c->colourSorted.as<uint2>(), M, 0, 2 * kMaxColours));
    // per-colour table for the placement (a colour nobody uses is an empty entry); the solve gets a compact copy
    // without empty entries below: an empty level would still cost a grid barrier per pass
    PHYX_TRY(c->part.binLevels.reserve(kMaxColours * sizeof(Level)));
    PHYX_TRY(c->levels.reserve(kMaxColours * sizeof(Level)));
    Level* colourTable = c->part.binLevels.as<Level>();
    k_unit_levels<<<1, 32, 0, c->stream>>>(counts, colourTable, firstPos, header);
    c->launches++;
    const size_t maxSlots = 2 * size_t(M) + 64 * kMaxColours;
    PHYX_TRY(c->slotJoint.reserve(maxSlots * sizeof(int)));
    PHYX_CUDA(cudaMemsetAsync(c->slotJoint.ptr, 0xff, maxSlots * sizeof(int), c->stream));
    k_unit_place<<<grid, kBlock, 0, c->stream>>>(M, c->colourSorted.as<uint2>(), colourTable, firstPos, c->manCount.as<int>(),
        c->contactPoints.as<float4>(), c->slotJoint.as<int>(), unsigned(kMaxColours));
    c->launches++;
    PHYX_CUDA(cudaGetLastError());

    struct { int header[4]; int result[4]; int counts[kMaxColours + 2]; } host;
    PHYX_CUDA(cudaMemcpyAsync(host.header, header, 16, cudaMemcpyDeviceToHost, c->stream));
    PHYX_CUDA(cudaMemcpyAsync(host.result, result, 16, cudaMemcpyDeviceToHost, c->stream));
    PHYX_CUDA(cudaMemcpyAsync(host.counts, counts, sizeof(host.counts), cudaMemcpyDeviceToHost, c->stream));
    PHYX_CUDA(cudaStreamSynchronize(c->stream));
    if (host.result[1])
    {
        set_error("colouring needs more than %d colours", kMaxColours);
        return PHYX_B200_ERR_CAPACITY;
    }
    c->slotCount = host.header[1];
    c->colourRounds = host.result[0];
    *staticsChanged = host.result[3] != 0;
    c->colourStateValid = true;
    c->colourStateBodies = nb;
    int cursor = 0;
    for (int k = 0; k < host.header[0]; ++k)
    {
        if (host.counts[k] > 0) c->hostLevels.push_back({ cursor, -1, cursor + 2 * host.counts[k] });
        cursor = (cursor + 2 * host.counts[k] + 63) & ~63;   // as k_unit_levels: an empty colour takes no slots
    }
    c->levelCount = int(c->hostLevels.size());
    c->coloursInUse = host.counts[kMaxColours + 1];
    if (!incremental) c->coloursAtFullBuild = c->coloursInUse;
    if (c->levelCount > 0)
        PHYX_CUDA(cudaMemcpyAsync(c->levels.ptr, c->hostLevels.data(), size_t(c->levelCount) * sizeof(Level), cudaMemcpyHostToDevice, c->stream));
    c->hostSlotsStale = true;
    return PHYX_B200_OK;
}

} // namespace phyx
