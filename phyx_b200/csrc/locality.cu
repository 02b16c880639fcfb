// phyx_b200 — memory order of the solver's body rows.
//
// The iterations are bound by the gather / scatter of 16-byte body rows: a warp pays
// one L1 wavefront per distinct 128-byte line its 32 rows fall into.  Contacts connect bodies that are
// neighbours in space, mostly vertically (things rest on things), so rows are laid out strip by strip:
// bodies are binned into vertical strips about two boxes wide and ordered bottom-to-top inside a strip.
// A joint's two rows are then often in the same line, and joints that are neighbours in slot order
// (sweep order) touch neighbouring strips.
//
// The order is a pure performance hint: any permutation gives the same results (the rows are a
// packed copy, reference PrepareBodies / FinishBodies, src/Solver.cpp:456-494), so it is refreshed
// only every few steps.  Built with the broadphase's stable radix passes (key = strip << 11 | height).
#include "common.cuh"

namespace phyx
{

constexpr int kBlock = 256;

// scene statistics over dynamic bodies: [0] min of aabb.min.x (as ordered uint), [1] min of aabb.min.y,
// [2] count, [3] sum of x extents (fixed point), [4] sum of y extents (fixed point)
__device__ __forceinline__ unsigned ordered_bits(float v)
{
    int f = __float_as_int(v);
    return unsigned(f) ^ (unsigned(f >> 31) | 0x80000000u);
}
__device__ __forceinline__ float from_ordered(unsigned u)
{
    unsigned m = (u & 0x80000000u) ? 0x80000000u : 0xffffffffu;
    return __int_as_float(int(u ^ m));
}

__global__ void __launch_bounds__(kBlock) k_locality_stats(int n, const float4* __restrict__ aabb, const float4* __restrict__ params,
    unsigned long long* __restrict__ stats)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned minx = 0xffffffffu, miny = 0xffffffffu;
    unsigned long long cnt = 0, sx = 0, sy = 0;
    if (b < n)
    {
        float4 p = params[b];
        if (!(p.x == 0.0f && p.y == 0.0f))
        {
            float4 a = aabb[b];
            minx = ordered_bits(a.x);
            miny = ordered_bits(a.y);
            cnt = 1;
            sx = (unsigned long long)(fminf(fmaxf(a.z - a.x, 0.f), 1e6f) * 64.f);
            sy = (unsigned long long)(fminf(fmaxf(a.w - a.y, 0.f), 1e6f) * 64.f);
        }
    }
    for (int o = 16; o > 0; o >>= 1)
    {
        minx = min(minx, __shfl_xor_sync(0xffffffffu, minx, o));
        miny = min(miny, __shfl_xor_sync(0xffffffffu, miny, o));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
    }
    if ((threadIdx.x & 31) == 0 && cnt)
    {
        atomicMin(&stats[0], (unsigned long long)minx);
        atomicMin(&stats[1], (unsigned long long)miny);
        atomicAdd(&stats[2], cnt);
        atomicAdd(&stats[3], sx);
        atomicAdd(&stats[4], sy);
    }
}

__global__ void __launch_bounds__(kBlock) k_locality_keys(int n, const float4* __restrict__ aabb, const float4* __restrict__ params,
    const unsigned long long* __restrict__ stats, uint2* __restrict__ kv)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    float4 p = params[b];
    unsigned key = 0xffffffffu;   // static bodies last
    const unsigned long long cnt = stats[2];
    if (!(p.x == 0.0f && p.y == 0.0f) && cnt)
    {
        const float x0 = from_ordered(unsigned(stats[0])), y0 = from_ordered(unsigned(stats[1]));
        const float meanW = fmaxf(float(stats[3]) / 64.f / float(cnt), 1e-3f), meanH = fmaxf(float(stats[4]) / 64.f / float(cnt), 1e-3f);
        float4 a = aabb[b];
        const float cx = 0.5f * (a.x + a.z), cy = 0.5f * (a.y + a.w);
        const float strip = floorf((cx - x0) / (2.0f * meanW));
        const float height = floorf((cy - y0) / (0.5f * meanH));
        const unsigned s = unsigned(fminf(fmaxf(strip, 0.f), 2097150.f));   // 21 bits
        const unsigned h = unsigned(fminf(fmaxf(height, 0.f), 2047.f));     // 11 bits
        key = (s << 11) | h;
    }
    kv[b] = make_uint2(key, unsigned(b));
}

__global__ void __launch_bounds__(kBlock) k_locality_finish(int n, const uint2* __restrict__ sorted, unsigned* __restrict__ order, int* __restrict__ rowOf)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned b = sorted[i].y;
    order[i] = b;
    rowOf[b] = i;
}

// Refresh the strip order (every kRefreshEvery solves, or when the body set changed).
int locality_order_update(phyx_b200_ctx* c)
{
    constexpr int kRefreshEvery = 8;
    const int n = c->bodyCount;
    const bool fresh = c->locValid && c->locBodies == n && (c->locAge++ % kRefreshEvery) != 0;
    if (fresh || n == 0) return PHYX_B200_OK;
    const size_t n1 = size_t(n);
    PHYX_TRY(c->locKeysA.reserve(n1 * sizeof(uint2)));
    PHYX_TRY(c->locKeysB.reserve(n1 * sizeof(uint2)));
    PHYX_TRY(c->locOrder.reserve(n1 * sizeof(unsigned)));
    PHYX_TRY(c->locRowOf.reserve(n1 * sizeof(int)));
    PHYX_TRY(c->locStats.reserve(64));
    unsigned long long* stats = c->locStats.as<unsigned long long>();
    PHYX_CUDA(cudaMemsetAsync(stats, 0xff, 16, c->stream));
    PHYX_CUDA(cudaMemsetAsync(stats + 2, 0, 48, c->stream));
    const int grid = (n + kBlock - 1) / kBlock;
    k_locality_stats<<<grid, kBlock, 0, c->stream>>>(n, c->aabb.as<float4>(), c->params.as<float4>(), stats);
    k_locality_keys<<<grid, kBlock, 0, c->stream>>>(n, c->aabb.as<float4>(), c->params.as<float4>(), stats, c->locKeysA.as<uint2>());
    c->launches += 2;
    PHYX_TRY(radix_pass(c, c->locKeysA.as<uint2>(), c->locKeysB.as<uint2>(), n, 0, 2048));
    PHYX_TRY(radix_pass(c, c->locKeysB.as<uint2>(), c->locKeysA.as<uint2>(), n, 11, 2048));
    PHYX_TRY(radix_pass(c, c->locKeysA.as<uint2>(), c->locKeysB.as<uint2>(), n, 22, 1024));
    k_locality_finish<<<grid, kBlock, 0, c->stream>>>(n, c->locKeysB.as<uint2>(), c->locOrder.as<unsigned>(), c->locRowOf.as<int>());
    c->launches++;
    PHYX_CUDA(cudaGetLastError());
    c->locValid = true;
    c->locBodies = n;
    if (c->locAge == 0) c->locAge = 1;
    return PHYX_B200_OK;
}

} // namespace phyx
