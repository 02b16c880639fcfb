// phyx_b200 — arithmetic of one contact joint and the static-body lastIteration words, shared by the
// iteration kernels (solve.cu: grid-barrier forms; strips.cu: strip-local form).
#pragma once
#include "common.cuh"

namespace phyx
{

constexpr float kProductiveImpulse = 1e-4f;    // Solver.cpp:8
constexpr float kFrictionCoefficient = 0.3f;   // Solver.cpp:9
constexpr int kPairHasB = int(0x80000000u);    // pairIdx.y: the manifold has a second joint
constexpr int kPairRecordWords = 6;            // float4 per manifold record (PairRecord below)

__device__ __forceinline__ float vmax(float l, float r) { return l > r ? l : r; }   // SIMD max: l>r?l:r

// ---- static bodies -----------------------------------------------------------------------------------
// A static body (invMass = invInertia = 0) never changes velocity, so joints that share one do not
// conflict and may sit in the same level; thousands of ground contacts would otherwise serialise.
// What they DO share is the body's lastIteration, which the reference updates joint by joint
// (Solver.cpp:903-910) and reads in the skip test (:790-798).  To reproduce the sequential
// semantics exactly, each static body has one 64-bit word per phase:
//     [63:48] latest iteration in which a joint on it was productive, +1 (0 = never)
//     [47:32] the productive iteration before that, +1
//     [31:0]  smallest sequential position among the productive joints of the latest iteration
// A joint at position p in iteration it therefore sees lastIteration = it exactly when an EARLIER
// joint (position < p) on that body was productive in this iteration, else the value carried over
// from previous iterations.  Schedules keep the joints of one static body in non-decreasing level
// order, so every earlier joint is in the same or an earlier level; if a joint of the same level
// becomes productive on a body whose carried-over value is stale ("cold"), the level is re-scanned
// for joints that this wakes up (k_solve, wake passes) until nothing changes.
__device__ __forceinline__ int static_visible_last(const unsigned long long* p, int it, unsigned pos)
{
    unsigned long long w = __ldcg(p);
    unsigned latest = unsigned(w >> 48), prev = unsigned(w >> 32) & 0xffffu, minPos = unsigned(w);
    if (latest == unsigned(it + 1)) return (minPos < pos) ? it : int(prev) - 1;
    return int(latest) - 1;
}

// Record "productive at (it, pos)".  Returns true if this changed the word while the body was cold
// (its carried-over lastIteration <= it-2), i.e. if it can wake up later joints of the same level.
__device__ __forceinline__ bool static_mark(unsigned long long* p, int it, unsigned pos, int* hotCounter)
{
    unsigned long long old = __ldcg(p);
    for (;;)
    {
        unsigned latest = unsigned(old >> 48), prev = unsigned(old >> 32) & 0xffffu, minPos = unsigned(old);
        unsigned long long nw;
        int carried;
        if (latest == unsigned(it + 1))
        {
            if (pos >= minPos) return false;
            nw = (old & 0xffffffff00000000ull) | pos;
            carried = int(prev) - 1;
        }
        else
        {
            nw = (static_cast<unsigned long long>(it + 1) << 48) | (static_cast<unsigned long long>(latest) << 32) | pos;
            carried = int(latest) - 1;
        }
        unsigned long long seen = atomicCAS(p, old, nw);
        if (seen == old)
        {
            // first productive joint on this body in this iteration: it will be "hot" in the next one
            if (hotCounter && latest != unsigned(it + 1)) atomicAdd(hotCounter, 1);
            return carried <= it - 2;
        }
        old = seen;
    }
}

__device__ __forceinline__ float flipsign_bits(float x, float y)   // SIMD_AVX2.h:272-275
{
    return __int_as_float(__float_as_int(x) ^ (__float_as_int(y) & 0x80000000));
}

// ---- the arithmetic of one joint --------------------------------------------------------------------
// PHASE 0: SolveJointsImpulses (Solver.cpp:833-901): normal impulse clamped to >= -accumulated, then
// friction clamped to the Coulomb cone of the updated normal impulse.  PHASE 1:
// SolveJointsDisplacement (:971-1003).  v1 / v2 are the two body rows; acc = {accN, accF} resp. {accD, -}.
// `wide` selects the SIMD (xor) form of flipsign over the scalar one (SIMD_AVX2.h:272 / SIMD_Scalar.h:265).
template <int PHASE>
__device__ __forceinline__ bool relax(const float4 c0, const float4 c1, const float4 c2, const float4 c3, float2& acc, float4& v1, float4& v2, bool wide)
{
    const float nx = c0.x, ny = c0.y, aN1 = c0.z, aN2 = c0.w;
    const float im1 = c2.x, ii1 = c2.y, im2 = c2.z, ii2 = c2.w;
    const float cinvN = c3.z;
    // normal limiter: projectors (n, -n), compMass = projector * invMass
    const float n2x = -nx, n2y = -ny;
    const float cm1x = nx * im1, cm1y = ny * im1, cm1a = aN1 * ii1;
    const float cm2x = n2x * im2, cm2y = n2y * im2, cm2a = aN2 * ii2;

    if (PHASE == 0)
    {
        const float aF1 = c1.x, aF2 = c1.y, cinvF = c1.z, dstVel = c1.w;

        float dV = dstVel;
        dV -= nx * v1.x;
        dV -= ny * v1.y;
        dV -= aN1 * v1.z;
        dV -= n2x * v2.x;
        dV -= n2y * v2.y;
        dV -= aN2 * v2.z;
        float dN = dV * cinvN;
        dN = vmax(dN, -acc.x);
        v1.x += cm1x * dN;
        v1.y += cm1y * dN;
        v1.z += cm1a * dN;
        v2.x += cm2x * dN;
        v2.y += cm2y * dN;
        v2.z += cm2a * dN;
        acc.x += dN;

        const float tx = -ny, ty = nx, t2x = -tx, t2y = -ty;
        float fV = 0.0f;
        fV -= tx * v1.x;
        fV -= ty * v1.y;
        fV -= aF1 * v1.z;
        fV -= t2x * v2.x;
        fV -= t2y * v2.y;
        fV -= aF2 * v2.z;
        float dF = fV * cinvF;
        const float force = acc.y + dF;
        const float limit = acc.x * kFrictionCoefficient;
        const float limitSigned = wide ? flipsign_bits(limit, force) : (force < 0.0f ? -limit : limit);
        const float adjusted = limitSigned - acc.y;
        dF = (fabsf(force) > limit) ? adjusted : dF;
        acc.y += dF;
        v1.x += (tx * im1) * dF;
        v1.y += (ty * im1) * dF;
        v1.z += (aF1 * ii1) * dF;
        v2.x += (t2x * im2) * dF;
        v2.y += (t2y * im2) * dF;
        v2.z += (aF2 * ii2) * dF;
        return vmax(fabsf(dN), fabsf(dF)) > kProductiveImpulse;
    }
    else
    {
        float accD = acc.x;
        float dV = c3.w;   // dstDisplacingVelocity
        dV -= nx * v1.x;
        dV -= ny * v1.y;
        dV -= aN1 * v1.z;
        dV -= n2x * v2.x;
        dV -= n2y * v2.y;
        dV -= aN2 * v2.z;
        float d = dV * cinvN;
        d = vmax(d, -accD);
        v1.x += cm1x * d;
        v1.y += cm1y * d;
        v1.z += cm1a * d;
        v2.x += cm2x * d;
        v2.y += cm2y * d;
        v2.z += cm2a * d;
        acc.x = accD + d;
        return fabsf(d) > kProductiveImpulse;
    }
}

struct PairRecord   // layout of pairQ: 6 float4 = 96 bytes per manifold; the displacement phase needs the first 64 bytes only
{
    float4 a0;   // joint a: {n.x, n.y, angN1, angN2}
    float4 b0;   // joint b: the same
    float4 m;    // {invMass1, invInertia1, invMass2, invInertia2}: shared, both joints have the same two bodies
    float4 nd;   // {compInvMassN a, dstDisplacingVelocity a, compInvMassN b, dstDisplacingVelocity b}
    float4 a1;   // joint a: {angF1, angF2, compInvMassF, dstVelocity}
    float4 b1;   // joint b: the same
};
static_assert(sizeof(PairRecord) == kPairRecordWords * sizeof(float4), "record layout");

} // namespace phyx
