// phyx_b200 — grid-wide barrier for persistent cooperative kernels (solve.cu, colour.cu).
#pragma once

namespace phyx
{

// ---- grid barrier -------------------------------------------------------------------------------------
// One barrier per level.  Besides synchronising, it ORs two flags over the whole grid at no extra
// latency by packing them into the arrival count: [19:0] arrivals, [39:20] CTAs asking for a wake
// pass, [59:40] CTAs that saw a productive joint in this iteration.  A ring of four words avoids
// sense reversal: word e+1 is cleared by CTA 0 before it arrives at barrier e.
struct BarrierResult
{
    bool wake, productive;
};

// Split form: arrive as soon as the CTA's last store of the level is issued, do independent work (prefetches for
// the next level), then wait.  Thread 0's release fence waits for ITS outstanding memory operations, so loads
// issued before the arrival would delay the whole CTA's arrival by their latency; issued between arrive and
// wait they overlap the barrier instead.
__device__ __forceinline__ void grid_arrive(unsigned long long* ring, unsigned epoch, bool wake, bool productive, unsigned long long& ticket)
{
    const int w = __syncthreads_or(wake ? 1 : 0);
    const int pr = __syncthreads_or(productive ? 1 : 0);
    ticket = 0ull;
    if (threadIdx.x == 0)
    {
        unsigned long long* word = ring + (epoch & 3u);
        if (blockIdx.x == 0) ring[(epoch + 1u) & 3u] = 0ull;
        const unsigned long long add = 1ull | (w ? (1ull << 20) : 0ull) | (pr ? (1ull << 40) : 0ull);
        // release: the CTA's writes (ordered before this point by the __syncthreads above) become visible before
        // the arrival (a release atomic: MEMBAR + ATOM, without the L1 invalidation an acq_rel fence adds)
        unsigned long long before;
        asm volatile("atom.release.gpu.global.add.u64 %0, [%1], %2;" : "=l"(before) : "l"(word), "l"(add) : "memory");
        ticket = before + add;
    }
}

__device__ __forceinline__ BarrierResult grid_wait(unsigned long long* ring, unsigned& epoch, unsigned long long ticket)
{
    __shared__ unsigned long long s_value;
    if (threadIdx.x == 0)
    {
        unsigned long long* word = ring + (epoch & 3u);
        unsigned long long v = ticket;
        // poll with relaxed loads and acquire once at the end: an acquire load per poll costs an L1
        // invalidation of the whole SM (SASS CCTL.IVALL, ~75 polls per barrier in the ncu capture of r1g)
        while ((v & 0xfffffull) != gridDim.x)
            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(word) : "memory");
        // one acquire load instead of a fence: a fence (MEMBAR) also waits for this thread's own outstanding loads,
        // i.e. for the next level's prefetch it issued between arrive and wait
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(word) : "memory");
        s_value = v;
    }
    __syncthreads();
    const unsigned long long v = s_value;
    epoch++;
    BarrierResult r;
    r.wake = ((v >> 20) & 0xfffffull) != 0;
    r.productive = ((v >> 40) & 0xfffffull) != 0;
    return r;
}

__device__ __forceinline__ BarrierResult grid_barrier(unsigned long long* ring, unsigned& epoch, bool wake, bool productive)
{
    unsigned long long ticket;
    grid_arrive(ring, epoch, wake, productive, ticket);
    return grid_wait(ring, epoch, ticket);
}

} // namespace phyx
