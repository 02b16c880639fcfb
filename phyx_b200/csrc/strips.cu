// phyx_b200 — strip-local sequential-impulse solve: the iteration loops of Solver::SolveJointIsland
// (reference src/Solver.cpp:130-215) with NO grid-wide barrier on the critical path.
//
// Reference stages replaced here (all src/Solver.cpp), same arithmetic as solve.cu (solve_math.cuh):
//   PreStepJoints            :697-758   warm start pass
//   SolveJointsImpulses      :760-914   impulse passes
//   SolveJointsDisplacement  :916-1018  displacement passes
//   SolveJointIsland loops   :159-211   iteration control incl. the productive early-out (:189, :210)
//
// Why: the grid-barrier forms (solve.cu) run one colour of the WHOLE world between two grid barriers; a pass over
// one colour costs ~9 us whatever it holds (4-5 us barrier + skew, one DRAM latency chain), 7-8 colours x 22 passes
// per solve (profiles/README.md, round 1).  Here the solver rows, which are stored in the broadphase's sorted-x
// order, are cut into S contiguous ranges ("strips", S <= number of SMs), one per CTA of ONE persistent kernel:
//   * the strip's body rows {v, w, lastIteration} live in SHARED MEMORY for the whole solve (a contiguous row range:
//     one 1-D bulk TMA copy, cp.async.bulk + mbarrier, SASS UBLKCP, per phase);
//   * a manifold whose dynamic bodies lie in one strip is INTERIOR to it: all colours of a strip's interior manifolds
//     are relaxed by its CTA under __syncthreads;
//   * a manifold whose bodies lie in adjacent strips k, k+1 is CUT (class S + k).  Cut sets of different boundaries
//     touch disjoint rows (checked when the layout is built), so ALL cut sets are relaxed concurrently, set k by
//     CTA k, on a shared-memory copy of the rows the set touches: the strip's own right-boundary rows and the
//     neighbour's left-boundary rows, which travel through global memory (L2) under point-to-point flags:
//         CTA k+1: interior -> store L(k+1) -> release flagA[k+1] .............. acquire flagB[k] -> load L(k+1)
//         CTA k  : interior -> acquire flagA[k+1] -> load L(k+1) -> cut set k -> store L(k+1) -> release flagB[k]
//     A CTA waits for its two neighbours only; skew does not accumulate across the grid;
//   * the productive early-out needs an OR over all CTAs: every CTA adds its flag to a per-iteration counter and
//     reads the counter of iteration it-2 before starting iteration it (complete by then without waiting, in
//     practice).  Running one more iteration after a non-productive one changes nothing (every joint fails the
//     lastIteration test, Solver.cpp:790-798), so results and the reported iteration counts equal the reference
//     loop's.
// Read as one order [strip 0 colours .. strip S-1 colours, cut set 0 .. cut set S-2] this is a valid sequential
// Gauss-Seidel sweep over the slot order, which is what the oracle checks bit for bit; as in the partitioned solve
// each class keeps its own lastIteration word per static body (tests: partition.sequential_equivalent).
//
// Inside a CTA a colour ("bin") is relaxed in two steps: every thread applies the skip test (Solver.cpp:790-798) to
// its candidates, which needs the 8-byte index word (prefetched during the previous bin) and two shared-memory rows,
// and pushes the ones that pass onto a shared-memory worklist; then the worklist is dealt round-robin over all
// threads, and only those fetch their 96-byte record from HBM.  Work is
// balanced however the active manifolds cluster, and no byte is fetched for a skipped manifold.
//
// Fallback: if a manifold spans non-adjacent strips, a row belongs to two cut sets, or a strip does not fit in shared
// memory, the layout is rejected and the colour-major layout + grid-barrier kernels of solve.cu run instead.
#include "common.cuh"
#include "scan.cuh"
#include "solve_math.cuh"
#include "tma.cuh"

#include <algorithm>

namespace phyx
{

constexpr int kBlock = 256;
constexpr int kStripU = 8;         // candidates per thread whose index words are prefetched one bin ahead
constexpr int kStripBins = kMaxColours;   // bins (colours) per class
constexpr int kStripMax = 1023;           // strips per layout (the class digit of the layout sort has 2048 bins)
constexpr size_t kStripSmemLimit = 227 * 1024 - 4096;   // dynamic shared memory of the strip kernel: the opt-in maximum minus its static tables
constexpr int kStripRowLimit = 10240;     // rows per strip (160 KB of the SM's 228 KB: the record fetches of step 2 want the rest as L1)

// layout header (device ints)
enum
{
    H_REJECT = 0,   // bit 0: a manifold spans non-adjacent strips, bit 1: a row is in two cut sets
    H_MAXROWS,      // rows of the largest strip
    H_MAXCUT,       // rows of the largest cut set (own right-boundary rows + the neighbour's left-boundary rows)
    H_ROW_LO,       // first / one past the last row the laid-out manifolds touch (an island partition owns a part of the world)
    H_MANIFOLDS,    // manifolds with a colour (slots / 2)
    H_MAXBIN,       // manifolds of the largest (class, colour) bin
    H_COLOURS,      // colours in use
    H_CUTM,         // cut manifolds
    H_TOTAL_R,
    H_TOTAL_L,
    H_SCAN_TOTAL,
    H_ROW_HI,
    H_WORDS = 16
};

enum
{
    kRejectFar = 1,
    kRejectBothSides = 2,
    kRejectSmem = 4,
    kRejectEmpty = 16
};

// ---- layout ----------------------------------------------------------------------------------------------------

// hist[row] = predicted work of the manifolds whose lower dynamic row it is.  A manifold costs one skip test per pass
// and, in the iterations in which it is relaxed, a record fetch + two relaxations (about 4x a test, measured).  How many
// iterations that will be is predicted from the previous step: activity[b] = last iteration in which body b received a
// productive impulse (FinishBodies keeps it), and a manifold is relaxed while either body is at most one iteration stale
// (Solver.cpp:790-798).  Without a previous step every manifold counts the same.
// (cover: +1 at the lower row and -1 at the upper row of every manifold between two dynamic bodies; its exclusive prefix
// sum at row r counts the manifolds a cut in front of row r would split.)
// owner / rank: with an island partition (islands.cu) only the manifolds of this rank's islands are laid out.
__device__ __forceinline__ bool manifold_is_mine(const unsigned char* __restrict__ bodyOwner, int rank, int2 bodies)
{
    return !bodyOwner || max(bodyOwner[bodies.x], bodyOwner[bodies.y]) == rank;
}

__global__ void __launch_bounds__(kBlock) k_strip_hist(Count Mc, const int2* __restrict__ jb, const int* __restrict__ work, const int* __restrict__ rowOf,
    const int* __restrict__ activity, const int2* __restrict__ manBody, const unsigned char* __restrict__ bodyOwner, int rank, int* __restrict__ hist,
    int* __restrict__ cover, int prevS, const int* __restrict__ prevCuts, const float* __restrict__ factor, int* __restrict__ header)
{
    __shared__ int s_lo, s_hi;
    if (threadIdx.x == 0)
    {
        s_lo = 0x7fffffff;
        s_hi = 0;
    }
    __syncthreads();
    const int M = count_of(Mc);
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    int r1 = -1, r2 = -1, home = -1;
    int2 b = make_int2(-1, -1);
    if (m < M && work[m] < kMaxColours && manifold_is_mine(bodyOwner, rank, manBody[m]))
    {
        b = jb[m];
        r1 = b.x < 0 ? -1 : (rowOf ? rowOf[b.x] : b.x);
        r2 = b.y < 0 ? -1 : (rowOf ? rowOf[b.y] : b.y);
        home = r1 < 0 ? r2 : (r2 < 0 ? r1 : min(r1, r2));
    }
    // the span of rows the laid-out manifolds touch: one atomic per block
    {
        int lo = home >= 0 ? home : 0x7fffffff, hi = home >= 0 ? max(r1, r2) + 1 : 0;
        for (int o = 16; o > 0; o >>= 1)
        {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if ((threadIdx.x & 31) == 0)
        {
            if (lo != 0x7fffffff) atomicMin(&s_lo, lo);
            if (hi) atomicMax(&s_hi, hi);
        }
        __syncthreads();
        if (threadIdx.x == 0)
        {
            if (s_lo != 0x7fffffff) atomicMin(&header[H_ROW_LO], s_lo);
            if (s_hi) atomicMax(&header[H_ROW_HI], s_hi);
        }
    }
    if (home < 0) return;
    if (r1 >= 0 && r2 >= 0 && r1 != r2)
    {
        atomicAdd(&cover[min(r1, r2)], 1);
        atomicAdd(&cover[max(r1, r2)], -1);
    }
    int weight = 8;
    if (activity)
    {
        const int last = max(b.x >= 0 ? activity[b.x] : -1, b.y >= 0 ? activity[b.y] : -1);
        weight = 5 + min(max(last + 2, 1), 24);
    }
    weight *= 8;   // fixed point for the feedback factor
    if (factor)
    {
        int lo = 0, hi = prevS - 1;   // the strip this row lived in last step
        while (lo < hi)
        {
            const int mid = (lo + hi + 1) >> 1;
            if (prevCuts[mid] <= home) lo = mid; else hi = mid - 1;
        }
        weight = max(1, int(float(weight) * factor[lo]));
    }
    atomicAdd(&hist[home], weight);
}

// Balance feedback.  The predicted weights are only a model; the kernel measures what each strip really cost (SM clocks
// its CTA spent working) and the next layout scales the weights of the manifolds that lived in strip k by factor[k], a
// damped running product of (cost of strip k / mean cost).  Strips move from step to step, so factors are looked up by
// row through the previous cuts and re-sampled onto the new cuts afterwards.
// (out of place: the factors of the last layout stay untouched until k_strip_commit, so a deferred step that stops after
// the layout can be laid out again by the stage path from the same state)
__global__ void k_strip_feedback(int S, const long long* __restrict__ cost, const float* __restrict__ factor, float* __restrict__ factorOut, bool useCost)
{
    // (sum in a fixed order: thread t adds its strips, then the 256 partial sums are added in index order by thread 0)
    __shared__ float s_part[256];
    __shared__ float s_sum;
    float part = 0.f;
    for (int k = threadIdx.x; k < S; k += blockDim.x) part += float(cost[k]);
    s_part[threadIdx.x] = part;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        float sum = 0.f;
        for (int t = 0; t < int(blockDim.x); ++t) sum += s_part[t];
        s_sum = sum;
    }
    __syncthreads();
    const float mean = useCost ? s_sum / float(S) : 0.f;   // (measured feedback switched off: the factors stay what they are)
    for (int k = threadIdx.x; k < S; k += blockDim.x)
    {
        float f = factor[k];
        if (mean > 0.f)
        {
            const float ratio = fminf(fmaxf(float(cost[k]) / mean, 0.5f), 2.0f);
            f = fminf(fmaxf(f * sqrtf(ratio), 0.25f), 4.0f);
        }
        factorOut[k] = f;
    }
}

// the new layout's balance state becomes the current one (not for a deferred step that has stopped)
__global__ void k_strip_commit(const StepCtl* ctl, int S, const float* __restrict__ factorNext, const int* __restrict__ cutsNext, float* __restrict__ factor,
    int* __restrict__ prevCuts)
{
    if (ctl && ctl->stop) return;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k <= S; k += gridDim.x * blockDim.x)
    {
        prevCuts[k] = cutsNext[k];
        if (k < S) factor[k] = factorNext[k];
    }
}

__global__ void k_strip_resample(int S, const int* __restrict__ cuts, const int* __restrict__ prevCuts, const float* __restrict__ factor, float* __restrict__ factorOut,
    int* __restrict__ prevCutsOut)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > S) return;
    prevCutsOut[k] = cuts[k];
    if (k == S) return;
    const int mid = (cuts[k] + cuts[k + 1]) >> 1;
    int lo = 0, hi = S - 1;   // strip of `mid` under the previous cuts
    while (lo < hi)
    {
        const int m = (lo + hi + 1) >> 1;
        if (prevCuts[m] <= mid) lo = m; else hi = m - 1;
    }
    factorOut[k] = factor[lo];
}

// the work of a row as the cut scan sees it: its manifolds' weights, plus a constant for every dynamic row (shared memory per
// strip is what limits its width); added on the fly by the scan's loader (scan.cuh)
struct RowWork
{
    const int* hist;
    const unsigned* order;
    const unsigned char* bodyStatic;
    const unsigned char* bodyOwner;
    int rank;
    __device__ __forceinline__ bool vector_ok(const int*) const { return false; }
    __device__ __forceinline__ void load4(int, int (&)[4]) const {}
    __device__ __forceinline__ int load(int r) const
    {
        const unsigned body = order ? order[r] : unsigned(r);
        return hist[r] + ((!bodyStatic[body] && (!bodyOwner || bodyOwner[body] == rank)) ? 24 : 0);
    }
};

// cuts[q] = first row with at least q/S of the manifolds before it
__global__ void __launch_bounds__(kBlock) k_strip_cuts(int nb, int S, const int* __restrict__ prefix, const int* __restrict__ total, const int* __restrict__ header,
    int* __restrict__ cuts)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q > S) return;
    // the strips cover the rows the laid-out manifolds touch; rows outside (other ranks' islands) belong to no strip
    const int rowLo = min(header[H_ROW_LO], nb), rowHi = max(min(header[H_ROW_HI], nb), rowLo);
    if (q == 0) { cuts[0] = rowLo; return; }
    if (q == S) { cuts[q] = rowHi; return; }
    const long long target = (static_cast<long long>(*total) * q + S - 1) / S;
    int lo = 0, hi = nb;   // smallest row r in [0, nb] with prefix[r] >= target (prefix[nb] := total)
    while (lo < hi)
    {
        const int mid = (lo + hi) >> 1;
        if (prefix[mid] >= target) hi = mid; else lo = mid + 1;
    }
    cuts[q] = min(max(lo, rowLo), rowHi);
}

// Move every cut to the nearest row in front of which NO manifold would be split, if there is one within `reach` rows
// (between separate piles there is: then whole islands lie inside a strip, the cut sets are empty, and the order in which
// an island's manifolds are relaxed, colour-major, no longer depends on where the cuts are: the island-parallel solve over
// several devices relies on that for bit-identical results).  covered[r] > 0: a cut in front of row r splits manifolds.
__global__ void __launch_bounds__(kBlock) k_strip_snap(int nb, int S, int reach, const int* __restrict__ covered, int* __restrict__ cuts)
{
    // one block per cut: the window [cut - reach, cut + reach] is searched by all threads, nearest clean row wins (the lower
    // one on a tie)
    __shared__ int s_best;
    const int q = blockIdx.x + 1;
    if (q >= S) return;
    if (threadIdx.x == 0) s_best = 0x7fffffff;
    __syncthreads();
    const int r0 = cuts[q], first = cuts[0], last = cuts[S];
    int best = 0x7fffffff;
    for (int i = threadIdx.x; i <= 2 * reach; i += blockDim.x)
    {
        const int r = r0 - reach + i;
        if (r < first || r > last || r >= nb || r < 0) continue;
        if (covered[r] != 0) continue;
        const int d = r < r0 ? r0 - r : r - r0;
        best = min(best, 2 * d + (r > r0 ? 1 : 0));
    }
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0 && best != 0x7fffffff) atomicMin(&s_best, best);
    __syncthreads();
    if (threadIdx.x == 0 && s_best != 0x7fffffff)
    {
        const int d = s_best >> 1;
        cuts[q] = (s_best & 1) ? r0 + d : r0 - d;
    }
}

// cuts in order, and no strip wider than `limit` rows (its rows must fit in shared memory, with room left for the L1);
// possible whenever bodies <= S * limit
__global__ void k_strip_monotonic(int S, int limit, int* __restrict__ cuts)
{
    // the two sweeps are serial; run them on a shared-memory copy (in global memory every step of the chain is an L2 round trip)
    __shared__ int s_cuts[kStripMax + 2];
    if (blockIdx.x != 0) return;
    for (int q = threadIdx.x; q <= S; q += blockDim.x) s_cuts[q] = cuts[q];
    __syncthreads();
    if (threadIdx.x == 0)
    {
        for (int q = 1; q < S; ++q) s_cuts[q] = min(max(s_cuts[q], s_cuts[q - 1]), s_cuts[q - 1] + limit);
        for (int q = S - 1; q >= 1; --q) s_cuts[q] = max(s_cuts[q], s_cuts[q + 1] - limit);
    }
    __syncthreads();
    for (int q = threadIdx.x; q <= S; q += blockDim.x) cuts[q] = s_cuts[q];
}

// largest k in [0, S) with cuts[k] <= row: the strip that holds the row (empty strips are never returned)
__device__ __forceinline__ int strip_of(const int* cuts, int S, int row)
{
    int lo = 0, hi = S - 1;
    while (lo < hi)
    {
        const int mid = (lo + hi + 1) >> 1;
        if (cuts[mid] <= row) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// classify, persist the colours, emit {class << 7 | colour, manifold} sort keys, flag the rows cut manifolds touch
// (flags[row]: bit 0 = right-boundary row of its strip, bit 1 = left-boundary row)
__global__ void __launch_bounds__(kBlock) k_strip_keys(Count Mc, int S, const int2* __restrict__ jb, const int* __restrict__ work, const int* __restrict__ rowOf,
    const int* __restrict__ cutsG, const int2* __restrict__ manBody, const unsigned char* __restrict__ bodyOwner, int rank, int* __restrict__ manColour,
    uint2* __restrict__ keys, int* __restrict__ flags, int* __restrict__ header)
{
    extern __shared__ int s_cuts[];
    for (int q = threadIdx.x; q <= S; q += blockDim.x) s_cuts[q] = cutsG[q];
    __syncthreads();
    const int M = count_of(Mc);
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    bool coloured = false, cut = false;
    int colour = 0;
    if (m < M)
    {
        const int c = work[m];
        manColour[m] = (c >= kMaxColours) ? -1 : c;
        unsigned key = unsigned(2 * S) << 7;   // skipped (no contact points, or another rank's island): behind every class
        if (c < kMaxColours) colour = c + 1;   // colours in use: counted whoever owns the manifold
        if (c < kMaxColours && manifold_is_mine(bodyOwner, rank, manBody[m]))
        {
            coloured = true;
            const int2 b = jb[m];
            const int r1 = b.x < 0 ? -1 : (rowOf ? rowOf[b.x] : b.x), r2 = b.y < 0 ? -1 : (rowOf ? rowOf[b.y] : b.y);
            int cls = 0;
            if (r1 >= 0 && r2 >= 0)
            {
                const int k1 = strip_of(s_cuts, S, r1), k2 = strip_of(s_cuts, S, r2);
                if (k1 == k2)
                    cls = k1;
                else
                {
                    const int lo = min(k1, k2), hi = max(k1, k2);
                    if (hi - lo > 1) atomicOr(&header[H_REJECT], kRejectFar);
                    cls = S + lo;
                    cut = true;
                    atomicOr(&flags[k1 < k2 ? r1 : r2], 1);
                    atomicOr(&flags[k1 < k2 ? r2 : r1], 2);
                }
            }
            else if (r1 >= 0 || r2 >= 0)
                cls = strip_of(s_cuts, S, r1 >= 0 ? r1 : r2);
            key = (unsigned(cls) << 7) | unsigned(c);
        }
        keys[m] = make_uint2(key, unsigned(m));
    }
    const int nCol = __syncthreads_count(coloured), nCut = __syncthreads_count(cut);
    // colours in use: block maximum first
    __shared__ int s_max;
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
    if (colour) atomicMax(&s_max, colour);
    __syncthreads();
    if (threadIdx.x == 0)
    {
        if (nCol) atomicAdd(&header[H_MANIFOLDS], nCol);
        if (nCut) atomicAdd(&header[H_CUTM], nCut);
        if (s_max) atomicMax(&header[H_COLOURS], s_max);
    }
}

// header and bin table of a new layout
__global__ void __launch_bounds__(kBlock) k_strip_init(int* __restrict__ header, int2* __restrict__ binRange, int bins)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 64) header[i] = i == H_ROW_LO ? 0x7f7f7f7f : 0;   // (larger than any row)
    if (i < bins) binRange[i] = make_int2(0, 0);
}

// one bit of the boundary flags as the value the boundary-list scans rank (scan.cuh loader)
struct FlagBit
{
    const int* flags;
    int shift;
    __device__ __forceinline__ bool vector_ok(const int*) const { return false; }
    __device__ __forceinline__ void load4(int, int (&)[4]) const {}
    __device__ __forceinline__ int load(int r) const { return (flags[r] >> shift) & 1; }
};

// boundary row lists in row order
__global__ void __launch_bounds__(kBlock) k_strip_lists(int nb, const int* __restrict__ flags, const int* __restrict__ prefixR, const int* __restrict__ prefixL,
    int* __restrict__ bR, int* __restrict__ bL, int* __restrict__ header)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nb) return;
    const int f = flags[r];
    if (f & 1) bR[prefixR[r]] = r;
    if (f & 2) bL[prefixL[r]] = r;
    if (f == 3) atomicOr(&header[H_REJECT], kRejectBothSides);
}

// per strip: first boundary row of each kind, sizes for the kernel's shared memory
// bStart: [0 .. S] right lists, [S+1 .. 2S+1] left lists
__global__ void __launch_bounds__(kBlock) k_strip_starts(int nb, int S, const int* __restrict__ cuts, const int* __restrict__ prefixR, const int* __restrict__ prefixL,
    int* __restrict__ bStart, int* __restrict__ header)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q > S) return;
    auto at = [&](const int* prefix, int total, int k) { const int row = cuts[k]; return row >= nb ? total : prefix[row]; };
    const int totalR = header[H_TOTAL_R], totalL = header[H_TOTAL_L];
    bStart[q] = at(prefixR, totalR, q);
    bStart[S + 1 + q] = at(prefixL, totalL, q);
    if (q < S)
    {
        atomicMax(&header[H_MAXROWS], cuts[q + 1] - cuts[q]);
        const int nR = at(prefixR, totalR, q + 1) - at(prefixR, totalR, q);
        const int nLn = q + 2 <= S ? at(prefixL, totalL, q + 2) - at(prefixL, totalL, q + 1) : 0;
        atomicMax(&header[H_MAXCUT], nR + nLn);
        const int nL = at(prefixL, totalL, q + 1) - at(prefixL, totalL, q);
        atomicMax(&header[H_MAXCUT], nL);   // the own-left list is staged with the same capacity
    }
}

// sorted position p = manifold slot p (joint slots 2p, 2p+1): joints, index words, bin table
__global__ void __launch_bounds__(kBlock) k_strip_place(Count Mc, int S, const uint2* __restrict__ sorted, const int2* __restrict__ manBody,
    const int* __restrict__ manCount, const float4* __restrict__ contactPoints, const int* __restrict__ rowOf, const unsigned char* __restrict__ bodyStatic,
    const int* __restrict__ cuts, const int* __restrict__ prefixR, const int* __restrict__ prefixL, const int* __restrict__ bStart,
    int* __restrict__ slotJoint, int2* __restrict__ pairIdx, unsigned* __restrict__ pairTest, int2* __restrict__ binRange)
{
    const int M = count_of(Mc);
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= M) return;
    const uint2 e = sorted[p];
    const int cls = int(e.x >> 7), colour = int(e.x & 127u);
    if (cls >= 2 * S) return;   // manifold without contact points
    const int m = int(e.y);
    const int bin = cls * kStripBins + colour;
    if (p == 0 || sorted[p - 1].x != e.x) binRange[bin].x = p;
    if (p == M - 1 || sorted[p + 1].x != e.x) binRange[bin].y = p + 1;
    // the joints of a manifold are the solverIndex of its contact points (World.cpp:100-103,140)
    slotJoint[2 * p] = __float_as_int(contactPoints[size_t(2 * m) * 2 + 1].w);
    const bool hasB = manCount[m] > 1;
    slotJoint[2 * p + 1] = hasB ? __float_as_int(contactPoints[size_t(2 * m + 1) * 2 + 1].w) : -1;
    const int2 b = manBody[m];
    const int r1 = rowOf ? rowOf[b.x] : b.x, r2 = rowOf ? rowOf[b.y] : b.y;
    const bool st1 = bodyStatic[b.x] != 0, st2 = bodyStatic[b.y] != 0;
    int x, y;
    if (cls < S)
    {
        const int row0 = cuts[cls];
        x = st1 ? (r1 | kStaticBit) : r1 - row0;
        y = st2 ? (r2 | kStaticBit) : r2 - row0;
    }
    else
    {
        // cut set k: buffer = [right-boundary rows of strip k][left-boundary rows of strip k+1]
        const int k = cls - S;
        const int nR = bStart[k + 1] - bStart[k];
        const int split = cuts[k + 1];
        x = r1 < split ? prefixR[r1] - bStart[k] : nR + prefixL[r1] - bStart[S + 1 + k + 1];
        y = r2 < split ? prefixR[r2] - bStart[k] : nR + prefixL[r2] - bStart[S + 1 + k + 1];
    }
    pairIdx[p] = make_int2(x, y | (hasB ? kPairHasB : 0));
    // the skip test's view: two 16-bit local rows, 0xffff = static body (its lastIteration lives in a global word)
    pairTest[p] = (st1 ? 0xffffu : unsigned(x)) | ((st2 ? 0xffffu : unsigned(y)) << 16);
}

// Cut sets get their own colouring.  The global colours (7-8 on a pile) leave a cut set of ~1000 manifolds in as many
// bins, and every bin costs the strip kernel two block barriers and a latency chain whatever it holds; within a cut set
// a body meets two or three manifolds, so first fit needs 3-4 colours.  One CTA per cut set: Jones-Plassmann rounds in
// shared memory with the slot position as priority (= sequential first fit in slot order: deterministic), then a stable
// counting sort of the set's slots by the new colour.  Sets too large for the shared-memory tables keep their bins.
constexpr int kRecolourThreads = 256;
constexpr int kRecolourItems = 16;                                  // manifolds per thread
constexpr int kRecolourCap = kRecolourThreads * kRecolourItems;     // manifolds per cut set
constexpr int kRecolourRows = 4096;                                 // rows per cut set

__global__ void __launch_bounds__(kRecolourThreads) k_strip_recolour_cuts(int S, int* __restrict__ slotJoint, int2* __restrict__ pairIdx,
    unsigned* __restrict__ pairTest, int2* __restrict__ binRange)
{
    __shared__ int s_claim[kRecolourRows];
    __shared__ unsigned s_used[kRecolourRows];
    __shared__ signed char s_col[kRecolourCap];
    __shared__ int s_lo, s_hi, s_rows;
    __shared__ int s_count[32], s_base[32], s_warp[kRecolourThreads / 32];
    const int cls = S + blockIdx.x;
    int2* bins = binRange + size_t(cls) * kStripBins;
    if (threadIdx.x == 0)
    {
        int lo = 0x7fffffff, hi = 0;
        for (int c = 0; c < kStripBins; ++c)
            if (bins[c].y > bins[c].x)
            {
                lo = min(lo, bins[c].x);
                hi = max(hi, bins[c].y);
            }
        s_lo = lo;
        s_hi = hi;
        s_rows = 0;
    }
    if (threadIdx.x < 32) s_count[threadIdx.x] = 0;
    __syncthreads();
    const int lo = s_lo, n = s_hi - s_lo;
    if (n <= 0 || n > kRecolourCap) return;
    unsigned t[kRecolourItems];
    int maxRow = 0;
#pragma unroll
    for (int k = 0; k < kRecolourItems; ++k)
    {
        const int i = int(threadIdx.x) * kRecolourItems + k;   // a thread owns consecutive slots: the final ranking is a plain scan
        t[k] = i < n ? pairTest[lo + i] : 0u;
        if (i < n)
        {
            maxRow = max(maxRow, int(max(t[k] & 0xffffu, t[k] >> 16)));
            s_col[i] = -1;
        }
    }
    atomicMax(&s_rows, maxRow + 1);
    __syncthreads();
    const int rows = s_rows;
    if (rows > kRecolourRows) return;   // (cut manifolds have two dynamic bodies: no 0xffff entries)
    for (int r = threadIdx.x; r < rows; r += kRecolourThreads) s_used[r] = 0u;
    for (;;)
    {
        for (int r = threadIdx.x; r < rows; r += kRecolourThreads) s_claim[r] = 0x7fffffff;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kRecolourItems; ++k)
        {
            const int i = int(threadIdx.x) * kRecolourItems + k;
            if (i < n && s_col[i] < 0)
            {
                atomicMin(&s_claim[t[k] & 0xffffu], i);
                atomicMin(&s_claim[t[k] >> 16], i);
            }
        }
        __syncthreads();
        bool left = false;
#pragma unroll
        for (int k = 0; k < kRecolourItems; ++k)
        {
            const int i = int(threadIdx.x) * kRecolourItems + k;
            if (i < n && s_col[i] < 0)
            {
                const int ra = t[k] & 0xffffu, rb = t[k] >> 16;
                if (s_claim[ra] == i && s_claim[rb] == i)
                {
                    const unsigned m = s_used[ra] | s_used[rb];
                    const int c = m == 0xffffffffu ? 31 : __ffs(~m) - 1;   // (32 manifolds on one body of a cut set: give up on exactness of the colouring? no: see below)
                    s_col[i] = (signed char)c;
                    s_used[ra] |= 1u << c;
                    s_used[rb] |= 1u << c;
                    if (m == 0xffffffffu) s_lo = -1;   // more than 32 colours: keep the global bins
                }
                else
                    left = true;
            }
        }
        if (!__syncthreads_or(left ? 1 : 0)) break;
    }
    if (s_lo < 0) return;
    // stable counting sort by colour: per colour an exclusive scan of "has this colour" in slot order
#pragma unroll
    for (int k = 0; k < kRecolourItems; ++k)
    {
        const int i = int(threadIdx.x) * kRecolourItems + k;
        if (i < n) atomicAdd(&s_count[s_col[i]], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        int run = 0;
        for (int c = 0; c < 32; ++c)
        {
            s_base[c] = run;
            run += s_count[c];
        }
    }
    __syncthreads();
    int dest[kRecolourItems];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = 0; c < 32; ++c)
    {
        if (s_count[c] == 0) continue;   // uniform
        int mine = 0;
#pragma unroll
        for (int k = 0; k < kRecolourItems; ++k)
        {
            const int i = int(threadIdx.x) * kRecolourItems + k;
            if (i < n && s_col[i] == c) ++mine;
        }
        int inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        int before = inc - mine;
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        __syncthreads();
        int at = s_base[c] + before;
#pragma unroll
        for (int k = 0; k < kRecolourItems; ++k)
        {
            const int i = int(threadIdx.x) * kRecolourItems + k;
            if (i < n && s_col[i] == c) dest[k] = at++;
        }
    }
    // permute the set's slots
    int ja[kRecolourItems], jb2[kRecolourItems];
    int2 idx[kRecolourItems];
#pragma unroll
    for (int k = 0; k < kRecolourItems; ++k)
    {
        const int i = int(threadIdx.x) * kRecolourItems + k;
        if (i < n)
        {
            ja[k] = slotJoint[2 * (lo + i)];
            jb2[k] = slotJoint[2 * (lo + i) + 1];
            idx[k] = pairIdx[lo + i];
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kRecolourItems; ++k)
    {
        const int i = int(threadIdx.x) * kRecolourItems + k;
        if (i < n)
        {
            const int d = lo + dest[k];
            slotJoint[2 * d] = ja[k];
            slotJoint[2 * d + 1] = jb2[k];
            pairIdx[d] = idx[k];
            pairTest[d] = t[k];
        }
    }
    if (threadIdx.x < kStripBins)
    {
        const int c = threadIdx.x;
        bins[c] = (c < 32 && s_count[c] > 0) ? make_int2(lo + s_base[c], lo + s_base[c] + s_count[c]) : make_int2(0, 0);
    }
}

__global__ void __launch_bounds__(kBlock) k_strip_maxbin(int bins, const int2* __restrict__ binRange, int* __restrict__ header)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= bins) return;
    const int2 r = binRange[b];
    if (r.y > r.x) atomicMax(&header[H_MAXBIN], r.y - r.x);
}

// Deferred step (common.cuh StepCtl): the decisions strip_layout's caller takes on the host after reading the header are
// taken here, and the step stops when any of them says "not this layout": rejected layout, shared-memory shape above the
// predicted one, colour overflow, a body that changed between static and dynamic, colour drift (colour.cu).
__host__ __device__ inline int strip_count_for(int manifolds, int bodies, int numSMs, int autoLimit)
{
    int S = manifolds / 1024;
    S = S < numSMs ? S : numSMs;
    S = S > 1 ? S : 1;
    if (autoLimit > 0 && S > autoLimit) S = autoLimit;   // what the last rejected layouts of this world allowed
    // a strip's rows must fit in shared memory: large worlds get more strips than SMs (several strips per CTA)
    const int needed = (bodies + kStripRowLimit * 3 / 4 - 1) / (kStripRowLimit * 3 / 4);
    if (needed > S) S = needed < kStripMax ? needed : kStripMax;
    return S;
}

__global__ void k_strip_verdict(StepCtl* ctl, const int* __restrict__ header, const int* __restrict__ colourResult, int rowCap, int cutCap, int workCap,
    int coloursAtFullBuild, int S, int numSMs, int autoLimit, int bodies)
{
    if (threadIdx.x != 0 || blockIdx.x != 0 || ctl->stop) return;
    int reason = 0;
    if (strip_count_for(ctl->manifolds, bodies, numSMs, autoLimit) != S) reason = 15;   // the stage path would choose another strip count
    else if (header[H_REJECT] || header[H_MANIFOLDS] == 0) reason = 10;
    else if (((header[H_MAXROWS] + 8) & ~7) > rowCap || ((header[H_MAXCUT] + 8) & ~7) > cutCap || ((header[H_MAXBIN] + 8) & ~7) > workCap) reason = 11;
    else if (colourResult[1]) reason = 12;
    else if (colourResult[3]) reason = 13;
    else if (header[H_COLOURS] > coloursAtFullBuild + kColourDrift) reason = 14;
    if (reason)
        ctl_stop(ctl, kStageSolve, header[H_MAXROWS], reason);
    else
        ctl->slots = 2 * header[H_MANIFOLDS];
}

// Rows a strip may have.  kStripRowLimit bounds what fits at all; within it, a layout whose shared memory (rows + cut-set
// buffers + worklist) stays inside the SM's 196 KB carve-out keeps 32 KB of L1 for the record fetches, one that spills into
// the 228 KB carve-out keeps none (0.95 against 1.15 ms on the 1 M pyramid, whose balanced strips sit right on that edge).
// So when the previous layout of this shape says the side buffers leave room for strips a seventh wider than average, the
// limit is what keeps the total inside 196 KB.  In steps of 256 rows: the limit is a launch parameter of the deferred step.
int strip_row_limit(const phyx_b200_ctx* c, int S)
{
    const StripPlan& sp = c->strip;
    if (S <= 1 || sp.feedbackStrips != S || sp.feedbackBodies != c->bodyCount || sp.maxStripRows <= 0)
        return sp.rowLimitForce > 0 ? std::min(sp.rowLimitForce, kStripRowLimit) : kStripRowLimit;
    const long long side = (long long)(sp.maxCutRows + sp.maxCutRows / 8 + 72) * 24 + (long long)(sp.maxBin + sp.maxBin / 8 + 72) * 2 + 5 * 1024;
    long long rows = ((196 * 1024 - side) / 16 - 16) & ~255ll;
    const int avg = (c->bodyCount + S - 1) / S;
    if (rows < avg + avg / 7 || rows < 2048)
    {
        // these strips do not fit that carve-out anyway: what the whole shared memory leaves for rows beside the side
        // buffers (kStripRowLimit assumes ~60 KB of them; a layout with more must keep its strips narrower or it is rejected)
        rows = ((long long)kStripSmemLimit + 5 * 1024 - side) / 16 - 32;
        rows &= ~63ll;
        if (rows < avg + 64) return kStripRowLimit;
    }
    if (sp.rowLimitForce > 0) rows = std::min<long long>(rows, sp.rowLimitForce);
    return int(std::min<long long>(rows, kStripRowLimit));
}

// strips for a world of this size: enough manifolds per strip to keep a CTA busy, at most one strip per SM
int strip_choose(const phyx_b200_ctx* c, int manifolds, int bodies)
{
    if (c->strip.want < 0) return 0;
    if (c->strip.want > 0) return std::min(c->strip.want, kStripMax);
    return strip_count_for(manifolds, bodies, c->numSMs, c->strip.autoLimit);
}

static size_t strip_smem_bytes(int rowCap, int cutCap, int workCap)
{
    return size_t(rowCap) * 16 + size_t(cutCap) * 16 + size_t(workCap) * 2 + size_t(cutCap) * (2 + 2 + 4);
}

// Class-major layout of the coloured manifolds over S strips; `work` holds the colours.  On success with *usable the
// context's schedule (slotJoint, pairIdx, bin table) is the strip layout; otherwise the caller lays out colour-major.
int strip_layout(phyx_b200_ctx* c, int S, const int2* jb, const int* work, const int* colourResult, int* colourResultHost, bool* usable)
{
    StripPlan& sp = c->strip;
    const int M = c->manifoldCount, nb = c->bodyCount;
    *usable = false;
    sp.valid = false;
    sp.rejected = 0;
    const int grid = (M + kBlock - 1) / kBlock, gridB = (nb + kBlock - 1) / kBlock, gridS = (S + 1 + kBlock - 1) / kBlock;
    const bool sortedRows = c->rowOrderValid && c->rowOrderBodies == nb;
    const int* rowOf = sortedRows ? c->rowOf.as<int>() : nullptr;
    const unsigned* order = sortedRows ? c->entryIndex.as<unsigned>() : nullptr;
    const int bins = 2 * S * kStripBins;
    const int rowLimit = strip_row_limit(c, S);
    sp.rowLimit = rowLimit;

    PHYX_TRY(sp.header.reserve(64 * sizeof(int)));
    PHYX_TRY(sp.flags.reserve(size_t(nb + 1) * 3 * sizeof(int)));
    PHYX_TRY(sp.prefixR.reserve(size_t(nb + 1) * sizeof(int)));
    PHYX_TRY(sp.prefixL.reserve(size_t(nb + 1) * sizeof(int)));
    PHYX_TRY(sp.bR.reserve(size_t(nb + 1) * sizeof(int)));
    PHYX_TRY(sp.bL.reserve(size_t(nb + 1) * sizeof(int)));
    PHYX_TRY(sp.cuts.reserve(size_t(S + 2) * sizeof(int)));
    PHYX_TRY(sp.bStart.reserve(size_t(2 * S + 4) * sizeof(int)));
    PHYX_TRY(sp.binRange.reserve(size_t(bins) * sizeof(int2)));
    int* header = sp.header.as<int>();
    // the three per-row arrays that start from zero share one buffer and one memset: boundary flags | work histogram | cover
    int* flags = sp.flags.as<int>();
    int* hist = flags + (nb + 1);
    int* cover = flags + 2 * (nb + 1);
    PHYX_CUDA(cudaMemsetAsync(flags, 0, size_t(nb + 1) * 3 * sizeof(int), c->stream));
    k_strip_init<<<(bins + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(header, sp.binRange.as<int2>(), bins);
    c->launches++;

    // cuts that balance the manifold count (a manifold counts for its lower dynamic row)
    const int* activity = (c->activityValid && c->activityBodies == nb) ? c->bodyActivity.as<int>() : nullptr;
    const bool split = c->islandRanks > 1 && c->islandsValid && c->islandBodies == nb;
    const unsigned char* bodyOwner = split ? c->bodyOwner.as<unsigned char>() : nullptr;
    // balance feedback from the previous solve of this world (same strip count, same bodies)
    PHYX_TRY(sp.cost.reserve(size_t(2 * S + 2) * sizeof(long long)));
    PHYX_TRY(sp.factor.reserve(size_t(3 * S + 3) * sizeof(float)));
    PHYX_TRY(sp.prevCuts.reserve(size_t(2 * S + 4) * sizeof(int)));
    const bool feedback = sp.feedbackStrips == S && sp.feedbackBodies == nb && activity != nullptr;
    float* factor = sp.factor.as<float>();
    float* factorNext = factor + S + 1;
    float* factorFb = factor + 2 * (S + 1);   // this step's factors: last layout's x the measured cost of its strips
    int* prevCuts = sp.prevCuts.as<int>();
    int* prevCutsNext = prevCuts + S + 2;
    if (feedback)
    {
        k_strip_feedback<<<1, 256, 0, c->stream>>>(S, sp.cost.as<long long>(), factor, factorFb, sp.measuredFeedback);
        c->launches++;
        // A step that rebuilds its colouring lays out twice (colour.cu), and the second layout must not apply the same
        // measured cost again on top of the first one's committed factors: squared factors squeeze strips until the layout
        // is rejected and the step falls back to the grid-barrier forms (seen as one form-1 solve, and the first-time
        // allocation of its streams, in the step of every rebuild).  The cost is consumed here; the strip kernel writes
        // the next.  (Not in a deferred step: if it stops, the stage path lays out again from the same state.)
        if (!c->def.active) PHYX_CUDA(cudaMemsetAsync(sp.cost.ptr, 0, size_t(S) * sizeof(long long), c->stream));
    }
    const Count Mc = c->count(M, &StepCtl::manifolds);
    k_strip_hist<<<grid, kBlock, 0, c->stream>>>(Mc, jb, work, rowOf, activity, c->manBody.as<int2>(), bodyOwner, c->islandRank, hist, cover, S,
        feedback ? prevCuts : nullptr, feedback ? factorFb : nullptr, header);
    PHYX_TRY(exclusive_scan_with(c, RowWork{ hist, order, c->bodyStatic.as<unsigned char>(), bodyOwner, c->islandRank }, sp.prefixR.as<int>(),
        Count(nb), header + H_SCAN_TOTAL));
    k_strip_cuts<<<gridS, kBlock, 0, c->stream>>>(nb, S, sp.prefixR.as<int>(), header + H_SCAN_TOTAL, header, sp.cuts.as<int>());
    c->launches += 2;
    if (S > 1)
    {
        PHYX_TRY(exclusive_scan_i32(c, cover, cover, nb, nullptr));
        // width limit first, then clean cuts, then the width limit again for the strips a snap made too wide
        k_strip_monotonic<<<1, 128, 0, c->stream>>>(S, rowLimit, sp.cuts.as<int>());
        k_strip_snap<<<S - 1, kBlock, 0, c->stream>>>(nb, S, std::max(1, nb / S / 2), cover, sp.cuts.as<int>());
        // (the snap may have pulled a cut up to half a strip away: a strip that became wider than the limit gets its cut
        // pulled back, which costs that cut its cleanness; leaving it rejected the whole layout for the step, and the
        // step then ran on the grid-barrier forms, with the first-time allocation of their streams: seen at 1 M bodies
        // as one form-1 solve and 2-140 ms of cudaMalloc in the step of a colour rebuild)
        k_strip_monotonic<<<1, 128, 0, c->stream>>>(S, rowLimit, sp.cuts.as<int>());
        c->launches += 3;
    }

    // carry the balance factors over to the new cuts
    if (feedback)
    {
        k_strip_resample<<<gridS, kBlock, 0, c->stream>>>(S, sp.cuts.as<int>(), prevCuts, factorFb, factorNext, prevCutsNext);
        c->launches++;
        if (!c->def.active)
        {
            k_strip_commit<<<gridS, kBlock, 0, c->stream>>>(nullptr, S, factorNext, prevCutsNext, factor, prevCuts);
            c->launches++;
        }
    }
    else
    {
        if (c->def.active)
        {
            set_error("deferred step without balance feedback (internal: the step was not eligible)");
            return PHYX_B200_ERR_STATE;
        }
        std::vector<float> ones(size_t(S), 1.0f);
        PHYX_CUDA(cudaMemcpyAsync(factor, ones.data(), size_t(S) * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        PHYX_CUDA(cudaMemcpyAsync(prevCuts, sp.cuts.ptr, size_t(S + 1) * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
        PHYX_CUDA(cudaMemsetAsync(sp.cost.ptr, 0, size_t(S) * sizeof(long long), c->stream));
        PHYX_CUDA(cudaStreamSynchronize(c->stream));   // `ones` is pageable
    }
    sp.feedbackStrips = S;
    sp.feedbackBodies = nb;

    // classes, sort keys, boundary flags; two stable passes: colour (7 bits), then class
    PHYX_TRY(c->colourKeys.reserve(size_t(M) * sizeof(uint2)));
    PHYX_TRY(c->colourSorted.reserve(size_t(M) * sizeof(uint2)));
    k_strip_keys<<<grid, kBlock, size_t(S + 1) * sizeof(int), c->stream>>>(Mc, S, jb, work, rowOf, sp.cuts.as<int>(), c->manBody.as<int2>(), bodyOwner, c->islandRank,
        c->manColour.as<int>(), c->colourKeys.as<uint2>(), flags, header);
    c->launches++;
    int digits = 1;
    while (digits < 2 * S + 1) digits <<= 1;
    PHYX_TRY(radix_pass_count(c, c->colourKeys.as<uint2>(), c->colourSorted.as<uint2>(), Mc, 0, 128));
    PHYX_TRY(radix_pass_count(c, c->colourSorted.as<uint2>(), c->colourKeys.as<uint2>(), Mc, 7, digits));
    const uint2* sorted = c->colourKeys.as<uint2>();

    // boundary row lists
    PHYX_TRY(exclusive_scan_with(c, FlagBit{ flags, 0 }, sp.prefixR.as<int>(), Count(nb), header + H_TOTAL_R));
    PHYX_TRY(exclusive_scan_with(c, FlagBit{ flags, 1 }, sp.prefixL.as<int>(), Count(nb), header + H_TOTAL_L));
    k_strip_lists<<<gridB, kBlock, 0, c->stream>>>(nb, flags, sp.prefixR.as<int>(), sp.prefixL.as<int>(),
        sp.bR.as<int>(), sp.bL.as<int>(), header);
    k_strip_starts<<<gridS, kBlock, 0, c->stream>>>(nb, S, sp.cuts.as<int>(), sp.prefixR.as<int>(), sp.prefixL.as<int>(), sp.bStart.as<int>(), header);
    c->launches += 2;

    // slots
    const size_t maxSlots = 2 * size_t(M) + 64;
    PHYX_TRY(c->slotJoint.reserve(maxSlots * sizeof(int)));
    PHYX_TRY(c->pairIdx.reserve((size_t(M) + 64) * sizeof(int2)));
    PHYX_TRY(sp.pairTest.reserve((size_t(M) + 64) * sizeof(unsigned)));
    k_strip_place<<<grid, kBlock, 0, c->stream>>>(Mc, S, sorted, c->manBody.as<int2>(), c->manCount.as<int>(), c->contactPoints.as<float4>(), rowOf,
        c->bodyStatic.as<unsigned char>(), sp.cuts.as<int>(), sp.prefixR.as<int>(), sp.prefixL.as<int>(), sp.bStart.as<int>(), c->slotJoint.as<int>(),
        c->pairIdx.as<int2>(), sp.pairTest.as<unsigned>(), sp.binRange.as<int2>());
    if (S > 1)
    {
        k_strip_recolour_cuts<<<S - 1, kRecolourThreads, 0, c->stream>>>(S, c->slotJoint.as<int>(), c->pairIdx.as<int2>(), sp.pairTest.as<unsigned>(), sp.binRange.as<int2>());
        c->launches++;
    }
    k_strip_maxbin<<<(bins + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(bins, sp.binRange.as<int2>(), header);
    c->launches += 2;
    PHYX_CUDA(cudaGetLastError());

    if (c->def.active)
    {
        // the verdict is taken on the device; the header comes home with the step's counts (deferred_finish, api.cu)
        k_strip_verdict<<<1, 32, 0, c->stream>>>(c->ctl(), header, colourResult, c->def.rowCap, c->def.cutCap, c->def.workCap, c->coloursAtFullBuild,
            S, c->numSMs, sp.autoLimit, nb);
        k_strip_commit<<<gridS, kBlock, 0, c->stream>>>(c->ctl(), S, factorNext, prevCutsNext, factor, prevCuts);
        c->launches += 2;
        PHYX_CUDA(cudaGetLastError());
        sp.strips = S;
        sp.valid = true;
        *usable = true;
        return PHYX_B200_OK;
    }
    int host[16];
    PHYX_TRY(mailbox_stage(c, header, sizeof(host), 0));
    if (colourResult) PHYX_TRY(mailbox_stage(c, colourResult, 16, 64));
    PHYX_TRY(mailbox_wait(c));
    memcpy(host, mailbox_at(c, 0), sizeof(host));
    if (colourResultHost) memcpy(colourResultHost, mailbox_at(c, 64), 16);

    sp.strips = S;
    *usable = strip_apply_header(c, host);
    return PHYX_B200_OK;
}

// A rejected layout halves the strip count and the world remembers the limit (autoLimit).  Rejections are often transient
// (a pile that is still collapsing, a strip squeezed narrow by the balance feedback), and a limit that stays for good
// leaves SMs idle: after 32 usable layouts under a limit the limit is doubled (dropped once it reaches the SM count); if
// the wider layout is rejected again, the halving comes back at the price of one extra layout build.
void strip_limit_recover(phyx_b200_ctx* c)
{
    StripPlan& sp = c->strip;
    if (sp.autoLimit <= 0)
    {
        sp.limitAge = 0;
        return;
    }
    if (++sp.limitAge < 32) return;
    sp.limitAge = 0;
    sp.autoLimit = sp.autoLimit * 2 >= c->numSMs ? 0 : sp.autoLimit * 2;
}

// the layout header (16 words, read back from the device) -> the plan's host fields; true if the layout is usable
bool strip_apply_header(phyx_b200_ctx* c, const int* host)
{
    StripPlan& sp = c->strip;
    sp.maxStripRows = host[H_MAXROWS];
    sp.maxCutRows = host[H_MAXCUT];
    sp.manifolds = host[H_MANIFOLDS];
    sp.maxBin = host[H_MAXBIN];
    sp.colours = host[H_COLOURS];
    sp.cutManifolds = host[H_CUTM];
    int rejected = host[H_REJECT];
    if (sp.manifolds == 0) rejected |= kRejectEmpty;
    if (strip_smem_bytes(sp.maxStripRows + 9, sp.maxCutRows + 9, sp.maxBin + 9) > kStripSmemLimit) rejected |= kRejectSmem;
    if (sp.maxStripRows > 65000 || sp.maxCutRows > 65000 || sp.maxBin > 65000) rejected |= kRejectSmem;
    sp.rejected = rejected;
    if (rejected) sp.lastReject = (rejected & 0xff) | (sp.strips << 8) | ((sp.maxStripRows / 64) << 20);
    sp.valid = rejected == 0;
    return sp.valid;
}

// Deferred step: the shared-memory shape the strip kernel will be launched with, predicted from the previous layouts
// (strips move a little from step to step; k_strip_verdict stops the step if the new layout does not fit).  Headroom is
// free only up to the next shared-memory carve-out of the SM: past it the kernel loses L1 (measured: +20 % kernel time on
// the 1 M pyramid for 10 KB too much), so the bound grows into the slack of the bucket the exact shape falls into and
// crosses into the next bucket only for a minimal margin.  False if the previous layout leaves no usable prediction.
bool strip_predict_caps(phyx_b200_ctx* c, int* rowCap, int* cutCap, int* workCap)
{
    const StripPlan& sp = c->strip;
    if (!sp.valid || sp.strips <= 0) return false;
    Deferred& d = c->def;
    // base shape: the largest of the last few layouts (islands hop between strips: the maxima oscillate), slowly forgotten
    d.baseRows = std::max(float(sp.maxStripRows), d.baseRows * 0.985f);
    d.baseCut = std::max(float(sp.maxCutRows), d.baseCut * 0.97f);
    d.baseBin = std::max(float(sp.maxBin), d.baseBin * 0.97f);
    auto round8 = [](float v) { return (int(v) + 8) & ~7; };
    // (rows: when the layout holds the strips inside the 196 KB carve-out, the limit itself is the bound: it never changes)
    const int limitRows = strip_row_limit(c, sp.strips);
    const bool fixedRows = limitRows < kStripRowLimit && round8(float(limitRows)) > sp.maxStripRows;
    const int r0 = fixedRows ? round8(float(limitRows)) : std::min(round8(d.baseRows), (kStripRowLimit + 8) & ~7), k0 = round8(d.baseCut), w0 = round8(d.baseBin);
    if (r0 <= sp.maxStripRows || k0 <= sp.maxCutRows || w0 <= sp.maxBin) return false;
    auto fits = [](int r, int k, int w) { return strip_smem_bytes(r + 1, k + 1, w + 1) <= kStripSmemLimit && r < 65000 && k < 65000 && w < 65000; };
    if (!fits(r0, k0, w0)) return false;
    int r = r0, k = k0, w = w0;
    if (!d.tight)
    {
        // carve-outs of an sm_100 SM (KB); the kernel's static tables and the system's 1 KB come on top of the dynamic part
        static const int buckets[] = { 8, 16, 32, 64, 100, 132, 164, 196, 228 };
        const size_t overhead = 5 * 1024, exact = strip_smem_bytes(r0, k0, w0) + overhead;
        size_t limit = 228 * 1024;
        for (int b : buckets)
            if (size_t(b) * 1024 >= exact) { limit = size_t(b) * 1024; break; }
        // wanted: rows + 25 % + 64, cut rows and bin + 50 % + 128; at least rows + 1 % + 16, the others + 3 % + 32
        // (room inside the bucket is free: take a lot of it, the shape then also stays valid for many steps)
        const int rWant = fixedRows ? r0 : r0 + r0 / 4 + 64, kWant = k0 + k0 / 2 + 128, wWant = w0 + w0 / 2 + 128;
        const int rMin = fixedRows ? r0 : r0 + r0 / 100 + 16, kMin = k0 + k0 / 32 + 32, wMin = w0 + w0 / 32 + 32;
        auto total = [&](int rr, int kk, int ww) { return strip_smem_bytes(rr, kk, ww) + overhead; };
        // the shape of the previous step still does (roomy enough, same carve-out): keep it, launch parameters that do not
        // change let the step be replayed as a graph
        if (*rowCap >= rMin && *cutCap >= kMin && *workCap >= wMin && total(*rowCap, *cutCap, *workCap) <= limit && fits(*rowCap, *cutCap, *workCap))
            return true;
        // binary search on the share t of (want - min) that still stays inside the bucket
        r = rMin; k = kMin; w = wMin;
        int lo = 0, hi = 64;
        while (lo < hi)
        {
            const int mid = (lo + hi + 1) / 2;
            const int rr = rMin + (rWant - rMin) * mid / 64, kk = kMin + (kWant - kMin) * mid / 64, ww = wMin + (wWant - wMin) * mid / 64;
            if (total(rr, kk, ww) <= limit) lo = mid; else hi = mid - 1;
        }
        r = rMin + (rWant - rMin) * lo / 64;
        k = kMin + (kWant - kMin) * lo / 64;
        w = wMin + (wWant - wMin) * lo / 64;
        r = std::min((r + 7) & ~7, (kStripRowLimit + 8) & ~7);
        k = (k + 7) & ~7;
        w = (w + 7) & ~7;
        while (!fits(r, k, w) && (r > r0 || k > k0 || w > w0))
        {
            r = std::max(r0, r - 8);
            k = std::max(k0, k - 8);
            w = std::max(w0, w - 8);
        }
    }
    *rowCap = r;
    *cutCap = k;
    *workCap = w;
    return true;
}

// host copy of the schedule for phyx_b200_get_schedule: one level per non-empty (class, colour) bin, in slot order;
// classStart[cls] = first slot of class cls (2S + 1 entries)
int strip_host_levels(phyx_b200_ctx* c, std::vector<int>* classStart)
{
    StripPlan& sp = c->strip;
    c->hostLevels.clear();
    if (classStart) classStart->clear();
    if (!sp.valid) return PHYX_B200_OK;
    const int bins = 2 * sp.strips * kStripBins;
    std::vector<int2> r;
    r.resize(size_t(bins));
    PHYX_CUDA(cudaMemcpyAsync(r.data(), sp.binRange.ptr, size_t(bins) * sizeof(int2), cudaMemcpyDeviceToHost, c->stream));
    PHYX_CUDA(cudaStreamSynchronize(c->stream));
    int cursor = 0;
    for (int b = 0; b < bins; ++b)
    {
        if (classStart && b % kStripBins == 0) classStart->push_back(cursor);
        if (r[b].y > r[b].x)
        {
            c->hostLevels.push_back({ 2 * r[b].x, -1, 2 * r[b].y });
            cursor = 2 * r[b].y;
        }
    }
    if (classStart) classStart->push_back(cursor);
    c->hostLevelsStale = false;
    return PHYX_B200_OK;
}

void strip_release(phyx_b200_ctx* c)
{
    StripPlan& sp = c->strip;
    DevBuf* bufs[] = { &sp.cuts, &sp.binRange, &sp.flags, &sp.prefixR, &sp.prefixL, &sp.bR, &sp.bL, &sp.bStart, &sp.header, &sp.sync, &sp.trace, &sp.pairTest, &sp.cost, &sp.factor, &sp.prevCuts };
    for (DevBuf* b : bufs) b->release();
    sp.valid = false;
}

// ---- the kernel ------------------------------------------------------------------------------------------------

struct StripParams
{
    float4* rows[2];                 // solver rows: impulse (velocity) and displacement arrays
    const float4* pairQ;             // PairRecord per manifold slot
    const int2* pairIdx;             // {row1, row2 | hasB}: rows local to the class's shared-memory buffer, or global | static bit
    const unsigned* pairTest;        // row1 | row2 << 16 (local; 0xffff = static body): all the skip test needs
    float2* accNF;                   // per joint slot
    float* accD;
    const int2* binRange;            // [2S][kStripBins] manifold slot ranges
    const int* cuts;                 // [S + 1]
    const int* bR;                   // right-boundary rows (global row numbers), row order
    const int* bL;                   // left-boundary rows
    const int* bStart;               // [0..S] starts in bR, [S+1 .. 2S+1] starts in bL
    unsigned long long* flagA;       // [S] "my interior pass `seq` is done and my left-boundary rows are in global memory"
    unsigned long long* flagB;       // [S] "cut set k of pass `seq` is done and strip k+1's left-boundary rows are in global memory"
    unsigned long long* done;        // [2][doneStride] per iteration: arrivals | productive CTAs << 32
    int doneStride;
    int S, contactIters, penetrationIters;
    int rowCap, cutCap, workCap;
    int* result;                     // [0] impulse iterations run, [1] displacement iterations run, [2] wake passes
    unsigned long long* activeTotal; // [2]
    long long* cost;                 // [S] SM clocks this strip's CTA spent working (not waiting for neighbours) in this solve: next step's balance
    unsigned long long* trace;       // developer aid (phyx_b200_strip_trace): [S][tracePasses][8] globaltimer stamps, or null
    int tracePasses;
    const StepCtl* ctl;              // deferred step: a stopped step (ctl->stop) must not be solved; null otherwise
};

__device__ __forceinline__ void flag_release(unsigned long long* flag, unsigned long long value)
{
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(flag), "l"(value) : "memory");
}

__device__ __forceinline__ unsigned long long flag_peek(const unsigned long long* flag)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long flag_acquire(const unsigned long long* flag)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
    return v;
}

// thread 0 spins until *flag >= want, then the CTA proceeds
__device__ __forceinline__ void cta_wait_flag(const unsigned long long* flag, unsigned long long want)
{
    if (threadIdx.x == 0)
    {
        while (flag_peek(flag) < want) {}
        flag_acquire(flag);
    }
    __syncthreads();
}

__device__ __forceinline__ void trace_mark(const StripParams& P, int k, int passIndex, int what)
{
    if (P.trace && threadIdx.x == 0 && passIndex < P.tracePasses)
    {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        P.trace[(size_t(k) * P.tracePasses + passIndex) * 8 + what] = t;
    }
}

struct StripCta
{
    float4* s_rows;     // the strip's rows
    float4* s_cut;      // rows of the cut set
    unsigned short* s_work;    // worklist: positions in the bin
    unsigned short* s_listL;   // own left-boundary rows (local)
    unsigned short* s_listR;   // own right-boundary rows (local)
    int* s_listN;              // the right neighbour's left-boundary rows (global row numbers)
    int2* s_bins;       // [nInt interior bins][nCut cut bins]
    int* s_count;       // [2] worklist lengths (alternating)
    int k, row0, nRows, nL, nR, nLn, nInt, nCut;
    int parity;         // which worklist counter the next bin pass uses
    unsigned active[2];
    long long clk[3];   // developer aid: SM clocks spent in step 1 / step 2 / (spare) of the bins of the current pass (thread 0)
    long long busy;     // SM clocks spent working, all passes (thread 0)
};

// warm start (PreStepJoints, Solver.cpp:736-750) of one bin: every manifold, no test; two manifolds per thread in flight
// (the manifolds of a bin touch disjoint rows)
template <int T>
__device__ __forceinline__ void prestep_bin(const StripParams& P, float4* rowsS, const float4* __restrict__ rowsG, int2 bin)
{
    constexpr int K = 2;
    const int n = bin.y - bin.x;
    for (int i0 = threadIdx.x; i0 < n; i0 += K * T)
    {
        int2 idx[K];
        float4 a0[K], b0[K], c2[K], a1[K], b1[K], accs[K];
#pragma unroll
        for (int u = 0; u < K; ++u)
        {
            const int i = i0 + u * T;
            idx[u] = make_int2(-1, -1);
            if (i < n)
            {
                const int p = bin.x + i;
                idx[u] = __ldg(&P.pairIdx[p]);
                const float4* rec = P.pairQ + size_t(p) * kPairRecordWords;
                a0[u] = __ldcs(rec);
                b0[u] = __ldcs(rec + 1);
                c2[u] = __ldcs(rec + 2);
                a1[u] = __ldcs(rec + 4);
                b1[u] = __ldcs(rec + 5);
                accs[u] = __ldcs(reinterpret_cast<const float4*>(&P.accNF[2 * p]));
            }
        }
#pragma unroll
        for (int u = 0; u < K; ++u)
        {
            if (i0 + u * T >= n) continue;
            const bool st1 = idx[u].x & kStaticBit, st2 = idx[u].y & kStaticBit, haveB = idx[u].y < 0;
            const int r1 = idx[u].x & kBodyMask, r2 = idx[u].y & kBodyMask;
            float4 v1 = st1 ? __ldcg(&rowsG[r1]) : rowsS[r1];
            float4 v2 = st2 ? __ldcg(&rowsG[r2]) : rowsS[r2];
#pragma unroll
            for (int h = 0; h < 2; ++h)
            {
                if (h == 1 && !haveB) break;
                const float4 c0 = h ? b0[u] : a0[u], c1 = h ? b1[u] : a1[u];
                const float nx = c0.x, ny = c0.y;
                const float accN = h ? accs[u].z : accs[u].x, accF = h ? accs[u].w : accs[u].y;
                v1.x += (nx * c2[u].x) * accN;
                v1.y += (ny * c2[u].x) * accN;
                v1.z += (c0.z * c2[u].y) * accN;
                v2.x += ((-nx) * c2[u].z) * accN;
                v2.y += ((-ny) * c2[u].z) * accN;
                v2.z += (c0.w * c2[u].w) * accN;
                const float tx = -ny, ty = nx;
                v1.x += (tx * c2[u].x) * accF;
                v1.y += (ty * c2[u].x) * accF;
                v1.z += (c1.x * c2[u].y) * accF;
                v2.x += ((-tx) * c2[u].z) * accF;
                v2.y += ((-ty) * c2[u].z) * accF;
                v2.z += (c1.y * c2[u].w) * accF;
            }
            if (!st1) rowsS[r1] = v1;
            if (!st2) rowsS[r2] = v2;
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// the record and accumulators of manifold slot p are about to be read: start them towards L2
template <int PHASE>
__device__ __forceinline__ void prefetch_manifold(const StripParams& P, int p)
{
    const float4* rec = P.pairQ + size_t(p) * kPairRecordWords;
    prefetch_l2(rec);
    prefetch_l2(rec + (PHASE == 0 ? 5 : 3));   // a 96-byte record may straddle two 128-byte lines
    if (PHASE == 0)
        prefetch_l2(&P.accNF[2 * p]);
    else
        prefetch_l2(&P.accD[2 * p]);
}

template <int T>
__device__ __forceinline__ void prefetch_idx(const StripParams& P, int2 bin, unsigned (&pre)[kStripU], int2& preIdx)
{
    const int n = bin.y - bin.x;
    // a small bin is relaxed without a worklist, by the thread that tests the candidate: it will want the index words too
    if (n <= T && int(threadIdx.x) < n) preIdx = __ldg(&P.pairIdx[bin.x + int(threadIdx.x)]);
#pragma unroll
    for (int u = 0; u < kStripU; ++u)
    {
        const int i = int(threadIdx.x) + u * T;
        pre[u] = i < n ? __ldg(&P.pairTest[bin.x + i]) : 0xffffffffu;
    }
}

// relax the (up to two) joints of manifold slot p on the shared-memory rows; returns "productive"
template <int PHASE>
__device__ __forceinline__ bool relax_manifold(const StripParams& P, StripCta& s, float4* rowsS, const float4* __restrict__ rowsG, int p, int2 idx, int it)
{
    const float4* rec = P.pairQ + size_t(p) * kPairRecordWords;
    const float4 a0 = __ldcs(rec), b0 = __ldcs(rec + 1), c2 = __ldcs(rec + 2), nd = __ldcs(rec + 3);
    float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = a1, accv;
    if (PHASE == 0)
    {
        a1 = __ldcs(rec + 4);
        b1 = __ldcs(rec + 5);
        accv = __ldcs(reinterpret_cast<const float4*>(&P.accNF[2 * p]));
    }
    else
    {
        const float2 a = __ldcs(reinterpret_cast<const float2*>(&P.accD[2 * p]));
        accv = make_float4(a.x, a.y, 0.f, 0.f);
    }
    const bool st1 = idx.x & kStaticBit, st2 = idx.y & kStaticBit, haveB = idx.y < 0;
    const int r1 = idx.x & kBodyMask, r2 = idx.y & kBodyMask;
    float4 w1 = st1 ? __ldcg(&rowsG[r1]) : rowsS[r1];
    float4 w2 = st2 ? __ldcg(&rowsG[r2]) : rowsS[r2];
    const int last1 = __float_as_int(w1.w), last2 = __float_as_int(w2.w);   // used for dynamic bodies only
    const float4 a3 = make_float4(0.f, 0.f, nd.x, nd.y), b3 = make_float4(0.f, 0.f, nd.z, nd.w);
    float2 accA, accB;
    if (PHASE == 0)
    {
        accA = make_float2(accv.x, accv.y);
        accB = make_float2(accv.z, accv.w);
    }
    else
    {
        accA = make_float2(accv.x, 0.f);
        accB = make_float2(accv.y, 0.f);
    }
    s.active[PHASE] += haveB ? 2u : 1u;
    const bool productiveA = relax<PHASE>(a0, a1, c2, a3, accA, w1, w2, false);
    bool productiveB = false;
    if (haveB) productiveB = relax<PHASE>(b0, b1, c2, b3, accB, w1, w2, false);
    if (PHASE == 0)
        __stcs(reinterpret_cast<float4*>(&P.accNF[2 * p]), make_float4(accA.x, accA.y, accB.x, accB.y));
    else
        __stcs(reinterpret_cast<float2*>(&P.accD[2 * p]), make_float2(accA.x, accB.x));
    // lastIteration = it where productive (Solver.cpp:903-910)
    const bool productive = productiveA || productiveB;
    if (!st1)
    {
        w1.w = __int_as_float(productive ? it : last1);
        rowsS[r1] = w1;
    }
    if (!st2)
    {
        w2.w = __int_as_float(productive ? it : last2);
        rowsS[r2] = w2;
    }
    return productive;
}

// One bin (one colour of one class) of one iteration.  PHASE 0: SolveJointsImpulses (Solver.cpp:781-910), PHASE 1:
// SolveJointsDisplacement (:937-1014).  `pre` holds the test words of this thread's first kStripU candidates; on return
// it holds those of `next`.  dummyRow: a shared-memory row nobody owns, whose lastIteration is "never" (the test words of
// static bodies and of slots beyond the bin point there).  Returns "this thread saw a productive joint".
//
// Static bodies.  The reference keeps lastIteration on static bodies too and the skip test reads it (Solver.cpp:790-798,
// 903-910): a productive ground contact of one pile keeps the ground contacts of every other pile awake.  That coupling is
// an artefact of sharing one record, not physics, and it would make an island's result depend on which other islands
// happen to be relaxed by the same CTA (or device).  The strip kernel therefore tracks a static body's lastIteration PER
// DYNAMIC PARTNER; since a productive joint marks both of its bodies, the partner's own lastIteration already says
// everything such a private copy would, and the static side of the test drops out.  The oracle reproduces this exactly
// on the equivalent problem in which every (dynamic body, static body) pair has its own copy of the static body
// (tests/conftest.py); the grid-barrier forms of solve.cu keep the reference's shared record.
template <int PHASE, int T>
__device__ __forceinline__ bool solve_bin(const StripParams& P, StripCta& s, float4* rowsS, int dummyRow, const float4* __restrict__ rowsG, int2 bin, int2 next, int it,
    unsigned (&pre)[kStripU], int2& preIdx)
{
    const int n = bin.y - bin.x;
    const int lane = threadIdx.x & 31;
    const unsigned below = (1u << lane) - 1u;
    bool anyProductive = false;
    const long long c0 = P.trace ? clock64() : 0;
    if (n <= T)
    {
        // small bin: every thread owns at most one candidate; test it and relax it on the spot (no worklist, one barrier)
        const unsigned t = pre[0];
        const int2 idx = preIdx;
        prefetch_idx<T>(P, next, pre, preIdx);
        const int ra = min(int(t & 0xffffu), dummyRow), rb = min(int(t >> 16), dummyRow);
        const bool active = (__float_as_int(rowsS[ra].w) > it - 2) || (__float_as_int(rowsS[rb].w) > it - 2);
        if (active) anyProductive = relax_manifold<PHASE>(P, s, rowsS, rowsG, bin.x + int(threadIdx.x), idx, it);
        __syncthreads();
        if (P.trace) s.clk[1] += clock64() - c0;
        return anyProductive;
    }
    int* count = &s.s_count[s.parity];
    // ---- step 1: the skip test (Solver.cpp:790-792, for both joints of a manifold at once, see solve.cu paired levels).
    // kStripU candidates per thread from prefetched test words; slots beyond the bin are skipped as a block (n is uniform)
    {
        bool active[kStripU];
        unsigned m[kStripU];
        int warpTotal = 0;
#pragma unroll
        for (int u = 0; u < kStripU; ++u)
        {
            active[u] = false;
            m[u] = 0u;
            if (u * T < n)
            {
                const unsigned t = pre[u];
                const int ra = min(int(t & 0xffffu), dummyRow), rb = min(int(t >> 16), dummyRow);
                const int la = __float_as_int(rowsS[ra].w), lb = __float_as_int(rowsS[rb].w);
                active[u] = (la > it - 2) || (lb > it - 2);
                // whoever relaxes it after the barrier finds its record on the way to L2 already
                if (active[u] && it > 0) prefetch_manifold<PHASE>(P, bin.x + int(threadIdx.x) + u * T);   // (iteration 0 relaxes everything: nothing to get ahead of)
                m[u] = __ballot_sync(0xffffffffu, active[u]);
                warpTotal += __popc(m[u]);
            }
        }
        // the test words are used up: fetch those of the next bin now, a whole bin pass ahead of their use
        prefetch_idx<T>(P, next, pre, preIdx);
        if (warpTotal)
        {
            int base = 0;
            if (lane == 0) base = atomicAdd(count, warpTotal);
            base = __shfl_sync(0xffffffffu, base, 0);
            // candidates of this warp in (u, lane) order
#pragma unroll
            for (int u = 0; u < kStripU; ++u)
            {
                if (active[u]) s.s_work[base + __popc(m[u] & below)] = (unsigned short)(int(threadIdx.x) + u * T);
                base += __popc(m[u]);
            }
        }
    }
    // candidates beyond the prefetched ones (bins larger than kStripU * T)
    for (int i0 = kStripU * T; i0 < n; i0 += T)
    {
        const int i = i0 + int(threadIdx.x);
        bool active = false;
        if (i < n)
        {
            const unsigned t = __ldg(&P.pairTest[bin.x + i]);
            const int ra = min(int(t & 0xffffu), dummyRow), rb = min(int(t >> 16), dummyRow);
            active = (__float_as_int(rowsS[ra].w) > it - 2) || (__float_as_int(rowsS[rb].w) > it - 2);
        }
        const unsigned m = __ballot_sync(0xffffffffu, active);
        if (m)
        {
            int base = 0;
            if (lane == 0) base = atomicAdd(count, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (active) s.s_work[base + __popc(m & below)] = (unsigned short)i;
        }
    }
    __syncthreads();
    const long long c1 = P.trace ? clock64() : 0;
    const int total = *count;
    if (threadIdx.x == 0) s.s_count[s.parity ^ 1] = 0;  // the next bin pass counts there
    s.parity ^= 1;

    // ---- step 2: relax the worklist, one entry per thread and round.  Consecutive entries (neighbouring records) go to
    // consecutive lanes: dealing them across the warps instead was measured 3x slower (DRAM locality of the record fetch)
    for (int w = threadIdx.x; w < total; w += T) {
        const int p = bin.x + s.s_work[w];
        anyProductive |= relax_manifold<PHASE>(P, s, rowsS, rowsG, p, __ldg(&P.pairIdx[p]), it);
    }
    // rows of this bin are written: the next bin may read them
    __syncthreads();
    if (P.trace)
    {
        s.clk[0] += c1 - c0;
        s.clk[1] += clock64() - c1;
    }
    return anyProductive;
}

// One pass over the strip's classes.  MODE -1: warm start, 0: impulse iteration `it`, 1: displacement iteration `it`.
template <int MODE, int T>
__device__ __forceinline__ bool run_pass(const StripParams& P, StripCta& s, int it, unsigned long long seq, int passIndex, unsigned (&pre)[kStripU], int2& preIdx)
{
    constexpr int PHASE = MODE == 1 ? 1 : 0;
    float4* rowsG = P.rows[PHASE];
    const int k = s.k, nBins = s.nInt + s.nCut;
    bool any = false;
    trace_mark(P, k, passIndex, 0);
    s.clk[0] = s.clk[1] = s.clk[2] = 0;
    const long long w0 = clock64();
    // interior
    for (int b = 0; b < s.nInt; ++b)
    {
        if (MODE < 0)
            prestep_bin<T>(P, s.s_rows, rowsG, s.s_bins[b]);
        else
            any |= solve_bin<PHASE, T>(P, s, s.s_rows, P.rowCap - 1, rowsG, s.s_bins[b], s.s_bins[b + 1 < nBins ? b + 1 : 0], it, pre, preIdx);
    }
    trace_mark(P, k, passIndex, 1);
    s.busy += clock64() - w0;
    if (P.trace && threadIdx.x == 0 && passIndex < P.tracePasses)
    {
        P.trace[(size_t(k) * P.tracePasses + passIndex) * 8 + 5] = (unsigned long long)s.clk[0];
        P.trace[(size_t(k) * P.tracePasses + passIndex) * 8 + 6] = (unsigned long long)s.clk[1];
    }
    if (P.S == 1) return any;
    // my left-boundary rows -> global memory, for cut set k-1 (which exists exactly if I have left-boundary rows)
    if (k > 0 && s.nL > 0)
    {
        for (int i = threadIdx.x; i < s.nL; i += T) __stcg(&rowsG[s.row0 + s.s_listL[i]], s.s_rows[s.s_listL[i]]);
        __syncthreads();
        if (threadIdx.x == 0) flag_release(&P.flagA[k], seq);
    }
    if (k + 1 < P.S && s.nCut > 0)
    {
        cta_wait_flag(&P.flagA[k + 1], seq);
        trace_mark(P, k, passIndex, 2);
        const long long w1 = clock64();
        for (int i = threadIdx.x; i < s.nR; i += T) s.s_cut[i] = s.s_rows[s.s_listR[i]];
        for (int i = threadIdx.x; i < s.nLn; i += T) s.s_cut[s.nR + i] = __ldcg(&rowsG[s.s_listN[i]]);
        __syncthreads();
        for (int b = s.nInt; b < nBins; ++b)
        {
            if (MODE < 0)
                prestep_bin<T>(P, s.s_cut, rowsG, s.s_bins[b]);
            else
                any |= solve_bin<PHASE, T>(P, s, s.s_cut, P.cutCap - 1, rowsG, s.s_bins[b], s.s_bins[b + 1 < nBins ? b + 1 : 0], it, pre, preIdx);
        }
        for (int i = threadIdx.x; i < s.nR; i += T) s.s_rows[s.s_listR[i]] = s.s_cut[i];
        for (int i = threadIdx.x; i < s.nLn; i += T) __stcg(&rowsG[s.s_listN[i]], s.s_cut[s.nR + i]);
        s.busy += clock64() - w1;
        __syncthreads();
        if (threadIdx.x == 0) flag_release(&P.flagB[k], seq);   // (strip k+1 waits for it exactly if it has left-boundary rows)
        trace_mark(P, k, passIndex, 3);
    }
    if (k > 0 && s.nL > 0)
    {
        cta_wait_flag(&P.flagB[k - 1], seq);
        for (int i = threadIdx.x; i < s.nL; i += T) s.s_rows[s.s_listL[i]] = __ldcg(&rowsG[s.row0 + s.s_listL[i]]);
        __syncthreads();
    }
    trace_mark(P, k, passIndex, 4);
    return any;
}

// per-strip tables into shared memory: sizes, the bin lists of its two classes (empty bins dropped), its boundary row lists
template <int T>
__device__ __forceinline__ void setup_strip(const StripParams& P, StripCta& s, int k, int2* s_tmp, int* s_n)
{
    const int S = P.S;
    s.k = k;
    s.row0 = P.cuts[k];
    s.nRows = P.cuts[k + 1] - s.row0;
    const int* bRs = P.bStart;
    const int* bLs = P.bStart + S + 1;
    s.nR = bRs[k + 1] - bRs[k];
    s.nL = bLs[k + 1] - bLs[k];
    s.nLn = k + 1 < S ? bLs[k + 2] - bLs[k + 1] : 0;
    __syncthreads();   // (whoever still reads the previous strip's tables is done)
    if (threadIdx.x < 2 * kStripBins)
    {
        const int cls = threadIdx.x < kStripBins ? k : S + k;
        const int c = threadIdx.x & (kStripBins - 1);
        s_tmp[threadIdx.x] = P.binRange[cls * kStripBins + c];   // the table has 2S classes; class 2S-1 is always empty
    }
    for (int i = threadIdx.x; i < s.nL; i += T) s.s_listL[i] = (unsigned short)(P.bL[bLs[k] + i] - s.row0);
    for (int i = threadIdx.x; i < s.nR; i += T) s.s_listR[i] = (unsigned short)(P.bR[bRs[k] + i] - s.row0);
    for (int i = threadIdx.x; i < s.nLn; i += T) s.s_listN[i] = P.bL[bLs[k + 1] + i];
    __syncthreads();
    if (threadIdx.x == 0)
    {
        int n = 0;
        for (int c = 0; c < kStripBins; ++c)
            if (s_tmp[c].y > s_tmp[c].x) s.s_bins[n++] = s_tmp[c];
        s_n[0] = n;
        if (k + 1 < S)
            for (int c = 0; c < kStripBins; ++c)
                if (s_tmp[kStripBins + c].y > s_tmp[kStripBins + c].x) s.s_bins[n++] = s_tmp[kStripBins + c];
        s_n[1] = n - s_n[0];
    }
    __syncthreads();
    s.nInt = s_n[0];
    s.nCut = s_n[1];
}

// One pass when there are MORE STRIPS THAN CTAs (worlds whose strips would not fit in shared memory otherwise: above ~1.5 M
// bodies per device).  CTA c owns the strips c, c + G, c + 2G, ...; a strip's rows are staged into shared memory for each
// visit and live in global memory in between.  Phase I: every owned strip's interior bins (its whole row range is written
// back, which also publishes the left-boundary rows; flagA).  Phase C: every owned cut set (flagB).  A strip's interior
// visit of pass p waits for the cut set on its left of pass p-1.  No wait can cycle: phase I of a pass waits only for
// phase C of the previous pass, phase C only for phase I of the same pass.
template <int MODE, int T>
__device__ __forceinline__ bool run_pass_streaming(const StripParams& P, StripCta& s, int it, unsigned long long seq, int passIndex, unsigned (&pre)[kStripU], int2& preIdx,
    int2* s_tmp, int* s_n)
{
    constexpr int PHASE = MODE == 1 ? 1 : 0;
    float4* rowsG = P.rows[PHASE];
    const int2 none = make_int2(0, 0);
    bool any = false;
    for (int k = blockIdx.x; k < P.S; k += gridDim.x)
    {
        setup_strip<T>(P, s, k, s_tmp, s_n);
        // the cut set on my left of the previous pass has written my left-boundary rows (not across a phase change: the
        // displacement rows are untouched until the first displacement pass)
        if (k > 0 && s.nL > 0 && seq > 1 && !(MODE == 1 && it == 0)) cta_wait_flag(&P.flagB[k - 1], seq - 1);
        const long long w0 = clock64();
        for (int i = threadIdx.x; i < s.nRows; i += T) s.s_rows[i] = __ldcg(&rowsG[s.row0 + i]);
        if (MODE >= 0 && s.nInt > 0) prefetch_idx<T>(P, s.s_bins[0], pre, preIdx);
        __syncthreads();
        for (int b = 0; b < s.nInt; ++b)
        {
            if (MODE < 0)
                prestep_bin<T>(P, s.s_rows, rowsG, s.s_bins[b]);
            else
                any |= solve_bin<PHASE, T>(P, s, s.s_rows, P.rowCap - 1, rowsG, s.s_bins[b], b + 1 < s.nInt ? s.s_bins[b + 1] : none, it, pre, preIdx);
        }
        for (int i = threadIdx.x; i < s.nRows; i += T) __stcg(&rowsG[s.row0 + i], s.s_rows[i]);
        if (threadIdx.x == 0) P.cost[k] += clock64() - w0;
        __syncthreads();
        if (threadIdx.x == 0 && k > 0 && s.nL > 0) flag_release(&P.flagA[k], seq);
    }
    for (int k = blockIdx.x; k + 1 < P.S; k += gridDim.x)
    {
        setup_strip<T>(P, s, k, s_tmp, s_n);
        if (s.nCut == 0) continue;
        cta_wait_flag(&P.flagA[k + 1], seq);
        const long long w1 = clock64();
        const int nBins = s.nInt + s.nCut;
        for (int i = threadIdx.x; i < s.nR; i += T) s.s_cut[i] = __ldcg(&rowsG[s.row0 + s.s_listR[i]]);
        for (int i = threadIdx.x; i < s.nLn; i += T) s.s_cut[s.nR + i] = __ldcg(&rowsG[s.s_listN[i]]);
        if (MODE >= 0) prefetch_idx<T>(P, s.s_bins[s.nInt], pre, preIdx);
        __syncthreads();
        for (int b = s.nInt; b < nBins; ++b)
        {
            if (MODE < 0)
                prestep_bin<T>(P, s.s_cut, rowsG, s.s_bins[b]);
            else
                any |= solve_bin<PHASE, T>(P, s, s.s_cut, P.cutCap - 1, rowsG, s.s_bins[b], b + 1 < nBins ? s.s_bins[b + 1] : none, it, pre, preIdx);
        }
        for (int i = threadIdx.x; i < s.nR; i += T) __stcg(&rowsG[s.row0 + s.s_listR[i]], s.s_cut[i]);
        for (int i = threadIdx.x; i < s.nLn; i += T) __stcg(&rowsG[s.s_listN[i]], s.s_cut[s.nR + i]);
        if (threadIdx.x == 0) P.cost[k] += clock64() - w1;
        __syncthreads();
        if (threadIdx.x == 0) flag_release(&P.flagB[k], seq);
    }
    (void)passIndex;
    return any;
}

// all iterations of one phase; returns the number of passes executed (>= the reference's count: see the header)
template <int PHASE, int T, bool STREAM>
__device__ __forceinline__ int run_phase(const StripParams& P, StripCta& s, int iters, unsigned long long& seq, int& passIndex, unsigned (&pre)[kStripU], int2& preIdx,
    int2* s_tmp, int* s_n)
{
    __shared__ unsigned long long s_word;
    unsigned long long* done = P.done + size_t(PHASE) * P.doneStride;
    int it = 0;
    for (; it < iters; ++it)
    {
        if (it >= 2)
        {
            // early-out (Solver.cpp:189 / :210), two iterations late: by now every CTA has reported iteration it-2
            if (threadIdx.x == 0)
            {
                unsigned long long v;
                while (((v = flag_peek(&done[it - 2])) & 0xffffffffull) != gridDim.x) {}
                s_word = v;
            }
            __syncthreads();
            const bool productive = (s_word >> 32) != 0;
            __syncthreads();
            if (!productive) break;
        }
        ++seq;
        ++passIndex;
        const bool mine = STREAM ? run_pass_streaming<PHASE, T>(P, s, it, seq, passIndex, pre, preIdx, s_tmp, s_n)
                                 : run_pass<PHASE, T>(P, s, it, seq, passIndex, pre, preIdx);
        const int any = __syncthreads_or(mine ? 1 : 0);
        if (threadIdx.x == 0) atomicAdd(&done[it], 1ull | (any ? (1ull << 32) : 0ull));
    }
    return it;
}

template <int T>
__global__ void __launch_bounds__(T, 1) k_solve_strips(StripParams P)
{
    extern __shared__ __align__(128) unsigned char stripSmem[];
    __shared__ int2 s_bins[2 * kStripBins];
    __shared__ int2 s_tmp[2 * kStripBins];
    __shared__ int s_count[2];
    __shared__ int s_n[2];
    __shared__ unsigned long long s_mbar;

    if (P.ctl && P.ctl->stop) return;   // (every CTA reads the same word: the whole grid leaves)
    StripCta s;
    s.s_rows = reinterpret_cast<float4*>(stripSmem);
    s.s_cut = s.s_rows + P.rowCap;
    s.s_listN = reinterpret_cast<int*>(s.s_cut + P.cutCap);
    s.s_listL = reinterpret_cast<unsigned short*>(s.s_listN + P.cutCap);
    s.s_listR = s.s_listL + P.cutCap;
    s.s_work = s.s_listR + P.cutCap;
    s.s_bins = s_bins;
    s.s_count = s_count;
    const int S = P.S;
    const bool streaming = S > int(gridDim.x);   // more strips than CTAs: rows are staged per visit (run_pass_streaming)
    if (streaming && threadIdx.x == 0)
        for (int k = blockIdx.x; k < S; k += gridDim.x) P.cost[k] = 0;   // accumulated per visit (only this CTA touches them)
    s.parity = 0;
    if (threadIdx.x == 0)   // the rows that test words of static bodies and of empty slots point at: lastIteration = never
    {
        s.s_rows[P.rowCap - 1] = make_float4(0.f, 0.f, 0.f, __int_as_float(-(1 << 30)));
        s.s_cut[P.cutCap - 1] = make_float4(0.f, 0.f, 0.f, __int_as_float(-(1 << 30)));
        mbar_init(&s_mbar, 1);
        mbar_fence_init();
        s_count[0] = s_count[1] = 0;
    }
    s.active[0] = s.active[1] = 0u;
    s.busy = 0;
    __syncthreads();

    unsigned long long seq = 0;
    int passIndex = 0;
    unsigned pre[kStripU];
#pragma unroll
    for (int u = 0; u < kStripU; ++u) pre[u] = 0xffffffffu;
    int2 preIdx = make_int2(-1, -1);
    int ranI = 0, ranD = 0;

    if (!streaming)
    {
        const int k = blockIdx.x;
        setup_strip<T>(P, s, k, s_tmp, s_n);
        // the strip's rows: one bulk copy; they stay in shared memory for the whole phase
        if (threadIdx.x == 0)
        {
            const unsigned bytes = unsigned(s.nRows) * 16u;
            mbar_expect_tx(&s_mbar, bytes);
            if (bytes) bulk_g2s(s.s_rows, P.rows[0] + s.row0, bytes, &s_mbar);
        }
        mbar_wait(&s_mbar, 0);

        // warm start
        ++seq;
        run_pass<-1, T>(P, s, 0, seq, passIndex, pre, preIdx);
        if (s.nInt + s.nCut > 0) prefetch_idx<T>(P, s_bins[0], pre, preIdx);
        ranI = run_phase<0, T, false>(P, s, P.contactIters, seq, passIndex, pre, preIdx, s_tmp, s_n);

        // velocity rows back, displacement rows in
        __syncthreads();
        if (threadIdx.x == 0 && s.nRows > 0)
        {
            bulk_s2g_fence();
            bulk_s2g(P.rows[0] + s.row0, s.s_rows, unsigned(s.nRows) * 16u);
            bulk_commit_wait_all();
        }
        if (P.penetrationIters > 0)
        {
            __syncthreads();
            if (threadIdx.x == 0)
            {
                const unsigned bytes = unsigned(s.nRows) * 16u;
                mbar_expect_tx(&s_mbar, bytes);
                if (bytes) bulk_g2s(s.s_rows, P.rows[1] + s.row0, bytes, &s_mbar);
            }
            mbar_wait(&s_mbar, 1);
            ranD = run_phase<1, T, false>(P, s, P.penetrationIters, seq, passIndex, pre, preIdx, s_tmp, s_n);
            __syncthreads();
            if (threadIdx.x == 0 && s.nRows > 0)
            {
                bulk_s2g_fence();
                bulk_s2g(P.rows[1] + s.row0, s.s_rows, unsigned(s.nRows) * 16u);
                bulk_commit_wait_all();
            }
        }
        if (threadIdx.x == 0) P.cost[k] = s.busy;
    }
    else
    {
        ++seq;
        run_pass_streaming<-1, T>(P, s, 0, seq, passIndex, pre, preIdx, s_tmp, s_n);
        ranI = run_phase<0, T, true>(P, s, P.contactIters, seq, passIndex, pre, preIdx, s_tmp, s_n);
        if (P.penetrationIters > 0) ranD = run_phase<1, T, true>(P, s, P.penetrationIters, seq, passIndex, pre, preIdx, s_tmp, s_n);
    }

    for (int phase = 0; phase < 2; ++phase)
    {
        unsigned v = s.active[phase];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&P.activeTotal[phase], static_cast<unsigned long long>(v));
    }
    // iteration counts as the reference loop reports them: up to and including the first non-productive iteration
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        const int executed[2] = { ranI, ranD };
        for (int phase = 0; phase < 2; ++phase)
        {
            const unsigned long long* done = P.done + size_t(phase) * P.doneStride;
            int ran = executed[phase];
            for (int it = 0; it < executed[phase]; ++it)
            {
                unsigned long long v;
                while (((v = flag_peek(&done[it])) & 0xffffffffull) != gridDim.x) {}
                if ((v >> 32) == 0)
                {
                    ran = it + 1;
                    break;
                }
            }
            P.result[phase] = ran;
        }
    }
}

// launch the strip kernel on the context's stream (rows are prepared, records refreshed)
int strip_solve_launch(phyx_b200_ctx* c, const phyx_b200_solve_config* cfg, float4* rowsVel, float4* rowsDisp)
{
    StripPlan& sp = c->strip;
    const int S = sp.strips, I = cfg->contactIterationsCount, D = cfg->penetrationIterationsCount;
    const int stride = std::max(I, D) + 2;
    const size_t syncWords = size_t(2) * S + 2 * size_t(stride) + 8;
    PHYX_TRY(sp.sync.reserve(syncWords * 8));
    PHYX_CUDA(cudaMemsetAsync(sp.sync.ptr, 0, syncWords * 8, c->stream));

    StripParams P;
    memset(&P, 0, sizeof(P));
    P.rows[0] = rowsVel;
    P.rows[1] = rowsDisp;
    P.pairQ = c->pairQ.as<float4>();
    P.pairIdx = c->pairIdx.as<int2>();
    P.pairTest = sp.pairTest.as<unsigned>();
    P.accNF = c->accNF.as<float2>();
    P.accD = c->accD.as<float>();
    P.binRange = sp.binRange.as<int2>();
    P.cuts = sp.cuts.as<int>();
    P.bR = sp.bR.as<int>();
    P.bL = sp.bL.as<int>();
    P.bStart = sp.bStart.as<int>();
    unsigned long long* sync = sp.sync.as<unsigned long long>();
    P.flagA = sync;
    P.flagB = sync + S;
    P.done = sync + 2 * S;
    P.doneStride = stride;
    P.S = S;
    P.contactIters = I;
    P.penetrationIters = D;
    // (deferred step: the shape predicted from the previous layout; k_strip_verdict has checked that this layout fits)
    P.rowCap = c->def.active ? c->def.rowCap : (sp.maxStripRows + 8) & ~7;
    P.cutCap = c->def.active ? c->def.cutCap : (sp.maxCutRows + 8) & ~7;
    P.workCap = c->def.active ? c->def.workCap : (sp.maxBin + 8) & ~7;
    P.ctl = c->def.active ? c->ctl() : nullptr;
    P.result = reinterpret_cast<int*>(c->solveFlags.as<char>() + 32);
    P.activeTotal = reinterpret_cast<unsigned long long*>(c->solveFlags.as<char>() + 48);
    P.cost = sp.cost.as<long long>();
    P.trace = nullptr;
    if (sp.tracePasses > 0)
    {
        PHYX_TRY(sp.trace.reserve(size_t(S) * sp.tracePasses * 8 * 8));
        PHYX_CUDA(cudaMemsetAsync(sp.trace.ptr, 0, size_t(S) * sp.tracePasses * 8 * 8, c->stream));
        P.trace = sp.trace.as<unsigned long long>();
        P.tracePasses = sp.tracePasses;
    }
    const size_t smem = strip_smem_bytes(P.rowCap, P.cutCap, P.workCap);
    if (!sp.attributeSet)   // per device
    {
        PHYX_CUDA(cudaFuncSetAttribute(k_solve_strips<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kStripSmemLimit)));
        sp.attributeSet = true;
    }
    // every CTA waits for its neighbours: all S must be resident at once, which the cooperative launch guarantees (or refuses)
    void* args[] = { &P };
    // one CTA per strip; a world with more strips than SMs is run by numSMs CTAs, each visiting several strips per pass
    PHYX_CUDA(cudaLaunchCooperativeKernel((void*)k_solve_strips<512>, dim3(std::min(S, c->numSMs)), dim3(512), args, smem, c->stream));
    c->launches++;
    return PHYX_B200_OK;
}

} // namespace phyx
