// phyx_b200 — sequential-impulse contact solve on the device.
//
// Reference stages replaced here (all src/Solver.cpp):
//   PrepareBodies            :456-480   nothing to do: the resident vel/disp rows ARE SolveBody
//   PrepareJoints copy       :509-521 } k_refresh: gather joint + contact point + body params by
//   RefreshJoints            :592-695 }            slot, write the packed per-slot streams
//   PreStepJoints            :697-758   k_solve pass 0 (warm start, schedule order)
//   SolveJointsImpulses      :760-914   k_solve impulse iterations
//   SolveJointsDisplacement  :916-1018  k_solve displacement iterations
//   SolveJointIsland loops   :130-215   k_solve iteration control incl. the productive early-out
//   FinishJoints             :527-547   k_finish
//   FinishBodies             :482-494   nothing to do (rows are already the body velocities)
//
// Packed joint streams (one float4 per slot and stream, slot order = schedule order, so every
// level is a contiguous, perfectly coalesced range):
//   Q0 {n.x, n.y, angN1, angN2}            normal, angular projectors of the normal limiter
//   Q1 {angF1, angF2, compInvMassF, dstVelocity}
//   Q2 {invMass1, invInertia1, invMass2, invInertia2}
//   Q3 {body1 | static<<30, body2 | static<<30, compInvMassN, dstDisplacingVelocity}
//   ACC {accumulatedImpulse N, F} (float2, read+write)   ACCD accumulatedDisplacingImpulse (float)
// The reference stores 29 floats per joint (ContactJointPacked, Solver.h:26-45); the other
// projectors and all compMass terms are single products/negations of the values above
// (n2 = -n1, t = (-n.y, n.x), compMass = projector * invMass), recomputed in registers with the
// same single rounding, so results are bit-identical while an impulse iteration streams 80 B
// per joint instead of 132 B.
//
// Body rows are float4 {vx, vy, w, lastIteration} gathered/scattered through L2 (ld/st .cg):
// 1 M bodies x 16 B x 2 arrays fit the 126 MB L2, the joint streams are marked evict-first.
//
// Compiled with --fmad=false: every float operation is the reference's, in the reference's order.
#include "common.cuh"

#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace phyx
{

constexpr int kBlock = 256;
constexpr float kProductiveImpulse = 1e-4f;    // Solver.cpp:8
constexpr float kFrictionCoefficient = 0.3f;   // Solver.cpp:9

struct SolveParams
{
    float4* vel;
    float4* disp;
    const float4* q0;
    const float4* q1;
    const float4* q2;
    const float4* q3;
    float2* accNF;
    float* accD;
    const Level* levels;
    int numLevels;
    unsigned long long* stampImp;
    unsigned long long* stampDisp;
    int contactIters, penetrationIters;
    unsigned* flags;   // [0, I) impulse productive words, [I, I+D) displacement
    int* result;       // [0] impulse iterations run, [1] displacement iterations run, [2] static hazards
};

__device__ __forceinline__ float vmax(float l, float r) { return l > r ? l : r; }   // SIMD max: l>r?l:r

// ---- RefreshLimiter, Solver.cpp:549-590: angular projectors + compInvMass ------------------------
__device__ __forceinline__ void refresh_limiter(float n1x, float n1y, float w1x, float w1y, float w2x, float w2y, float im1, float ii1,
    float im2, float ii2, float& a1, float& a2, float& cinv)
{
    float n2x = -n1x, n2y = -n1y;
    a1 = n1x * w1y - n1y * w1x;
    a2 = n2x * w2y - n2y * w2x;
    float cm1x = n1x * im1, cm1y = n1y * im1, cm1a = a1 * ii1;
    float cm2x = n2x * im2, cm2y = n2y * im2, cm2a = a2 * ii2;
    float c1 = n1x * cm1x + n1y * cm1y + a1 * cm1a;
    float c2 = n2x * cm2x + n2y * cm2y + a2 * cm2a;
    float c = c1 + c2;
    cinv = (fabsf(c) > 0.0f) ? __fdiv_rn(1.0f, c) : 0.0f;
}

// ---- PrepareJoints copy + RefreshJoints ----------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_refresh(int numSlots, const int* __restrict__ slotJoint, const phyx_contact_joint* __restrict__ joints,
    const float4* __restrict__ contactPoints, const float4* __restrict__ params, float4* __restrict__ q0, float4* __restrict__ q1,
    float4* __restrict__ q2, float4* __restrict__ q3, float2* __restrict__ accNF, float* __restrict__ accD)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= numSlots) return;
    int j = slotJoint[s];
    if (j < 0)
    {
        q3[s] = make_float4(__int_as_float(-1), __int_as_float(-1), 0.f, 0.f);
        return;
    }
    phyx_contact_joint jt = joints[j];
    float4 c0 = contactPoints[size_t(jt.contactPointIndex) * 2];       // delta1, delta2
    float4 c1 = contactPoints[size_t(jt.contactPointIndex) * 2 + 1];   // normal, flags, solverIndex
    float4 p1 = params[jt.body1Index], p2 = params[jt.body2Index];     // invMass, invInertia, pos

    float nx = c1.x, ny = c1.y;
    float point1x = c0.x + p1.z, point1y = c0.y + p1.w;
    float point2x = c0.z + p2.z, point2y = c0.w + p2.w;
    float w1x = c0.x, w1y = c0.y;
    float w2x = point1x - p2.z, w2y = point1y - p2.w;

    float aN1, aN2, cinvN, aF1, aF2, cinvF;
    refresh_limiter(nx, ny, w1x, w1y, w2x, w2y, p1.x, p1.y, p2.x, p2.y, aN1, aN2, cinvN);

    // Solver.cpp:658-679.  bounce is 0, so dv = (-0)*(relative velocity . n) and
    // dstVelocity = max(dv - 1, 0) = 0 for every finite input: the point-velocity terms of the
    // reference drop out and the body velocities need not be gathered here.
    float depth = (point2x - point1x) * nx + (point2y - point1y) * ny;
    float dstVelocity = 0.0f;
    float dstVel = (depth < 1.0f) ? dstVelocity - 0.1f : dstVelocity;
    float dstDisp = 0.1f * vmax(0.0f, depth - 2.0f * 1.0f);

    float tx = -ny, ty = nx;
    refresh_limiter(tx, ty, w1x, w1y, w2x, w2y, p1.x, p1.y, p2.x, p2.y, aF1, aF2, cinvF);

    int b1 = jt.body1Index | ((p1.x == 0.0f && p1.y == 0.0f) ? kStaticBit : 0);
    int b2 = jt.body2Index | ((p2.x == 0.0f && p2.y == 0.0f) ? kStaticBit : 0);

    q0[s] = make_float4(nx, ny, aN1, aN2);
    q1[s] = make_float4(aF1, aF2, cinvF, dstVel);
    q2[s] = make_float4(p1.x, p1.y, p2.x, p2.y);
    q3[s] = make_float4(__int_as_float(b1), __int_as_float(b2), cinvN, dstDisp);
    accNF[s] = make_float2(jt.normalLimiter_accumulatedImpulse, jt.frictionLimiter_accumulatedImpulse);
    accD[s] = 0.0f;
}

// ---- FinishJoints ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_finish(int numSlots, const int* __restrict__ slotJoint, const float2* __restrict__ accNF,
    phyx_contact_joint* __restrict__ joints)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= numSlots) return;
    int j = slotJoint[s];
    if (j < 0) return;
    float2 a = accNF[s];
    joints[j].normalLimiter_accumulatedImpulse = a.x;
    joints[j].frictionLimiter_accumulatedImpulse = a.y;
}

// ---- static-body lastIteration stamps ----------------------------------------------------------------
// A static body can sit in many joints of one level, so its lastIteration cannot live in a row
// that those joints race on.  Per static body one 64-bit word {hi = tick of the latest productive
// level, lo = tick of the one before}; tick = iteration*numLevels + level + 1.  A reader in tick t
// sees hi if hi < t, else lo: writes of the CURRENT level are invisible, whatever the timing, so
// the result is deterministic ("visible to later levels only").
__device__ __forceinline__ int stamp_visible_last(const unsigned long long* p, unsigned tick, int numLevels)
{
    unsigned long long w = __ldcg(p);
    unsigned hi = unsigned(w >> 32), lo = unsigned(w);
    unsigned t = hi < tick ? hi : lo;
    return t == 0 ? -1 : int((t - 1) / unsigned(numLevels));
}

__device__ __forceinline__ void stamp_write(unsigned long long* p, unsigned tick)
{
    unsigned long long old = __ldcg(p);
    while (unsigned(old >> 32) != tick)
    {
        unsigned long long nw = (static_cast<unsigned long long>(tick) << 32) | (old >> 32);
        unsigned long long prev = atomicCAS(p, old, nw);
        if (prev == old) break;
        old = prev;
    }
}

__device__ __forceinline__ float flipsign_bits(float x, float y)   // SIMD_AVX2.h:272-275
{
    return __int_as_float(__float_as_int(x) ^ (__float_as_int(y) & 0x80000000));
}

// ---- PreStepJoints, Solver.cpp:736-750 ---------------------------------------------------------------
__device__ __forceinline__ void prestep_slot(const SolveParams& P, int s)
{
    float4 c3 = __ldcs(&P.q3[s]);
    int r1 = __float_as_int(c3.x), r2 = __float_as_int(c3.y);
    if (r1 < 0) return;
    int b1 = r1 & kBodyMask, b2 = r2 & kBodyMask;
    bool st1 = r1 & kStaticBit, st2 = r2 & kStaticBit;
    float4 c0 = __ldcs(&P.q0[s]), c1 = __ldcs(&P.q1[s]), c2 = __ldcs(&P.q2[s]);
    float2 acc = __ldcs(&P.accNF[s]);
    float4 v1 = __ldcg(&P.vel[b1]), v2 = __ldcg(&P.vel[b2]);
    float nx = c0.x, ny = c0.y;
    float accN = acc.x, accF = acc.y;

    v1.x += (nx * c2.x) * accN;
    v1.y += (ny * c2.x) * accN;
    v1.z += (c0.z * c2.y) * accN;
    v2.x += ((-nx) * c2.z) * accN;
    v2.y += ((-ny) * c2.z) * accN;
    v2.z += (c0.w * c2.w) * accN;

    float tx = -ny, ty = nx;
    v1.x += (tx * c2.x) * accF;
    v1.y += (ty * c2.x) * accF;
    v1.z += (c1.x * c2.y) * accF;
    v2.x += ((-tx) * c2.z) * accF;
    v2.y += ((-ty) * c2.z) * accF;
    v2.z += (c1.y * c2.w) * accF;

    if (!st1) __stcg(&P.vel[b1], v1);
    if (!st2) __stcg(&P.vel[b2], v2);
}

// ---- one level of one iteration ------------------------------------------------------------------------
// PHASE 0: SolveJointsImpulses (Solver.cpp:781-910); PHASE 1: SolveJointsDisplacement (:937-1014)
template <int PHASE>
__device__ __forceinline__ bool solve_level(const SolveParams& P, const Level L, int it, unsigned tick, int tid, int nthreads)
{
    float4* rows = PHASE == 0 ? P.vel : P.disp;
    unsigned long long* stamps = PHASE == 0 ? P.stampImp : P.stampDisp;
    const int lane = threadIdx.x & 31;
    bool anyProductive = false;

    for (int s0 = L.start + (tid & ~31); s0 < L.end; s0 += nthreads)   // warp-uniform trip count
    {
        const int s = s0 + lane;
        bool valid = s < L.end;
        float4 c3 = make_float4(__int_as_float(-1), __int_as_float(-1), 0.f, 0.f);
        if (valid) c3 = __ldcs(&P.q3[s]);
        const int r1 = __float_as_int(c3.x), r2 = __float_as_int(c3.y);
        valid = valid && r1 >= 0;
        const int b1 = r1 & kBodyMask, b2 = r2 & kBodyMask;
        const bool st1 = r1 & kStaticBit, st2 = r2 & kStaticBit;

        float4 v1 = make_float4(0.f, 0.f, 0.f, 0.f), v2 = v1;
        int last1 = -1, last2 = -1;
        bool active = false;
        if (valid)
        {
            v1 = __ldcg(&rows[b1]);
            v2 = __ldcg(&rows[b2]);
            last1 = st1 ? stamp_visible_last(&stamps[b1], tick, P.numLevels) : __float_as_int(v1.w);
            last2 = st2 ? stamp_visible_last(&stamps[b2], tick, P.numLevels) : __float_as_int(v2.w);
            active = (last1 > it - 2) || (last2 > it - 2);   // Solver.cpp:790-792
        }
        const bool wide = s < L.grouped_end;
        unsigned m = __ballot_sync(0xffffffffu, active);
        if (wide) active = valid && ((m >> (lane & ~7)) & 0xffu);   // AVX2: skip only if none of the 8 lanes (:797)

        bool productive = false;
        if (active)
        {
            const float4 c0 = __ldcs(&P.q0[s]);
            const float4 c2 = __ldcs(&P.q2[s]);
            const float nx = c0.x, ny = c0.y, aN1 = c0.z, aN2 = c0.w;
            const float im1 = c2.x, ii1 = c2.y, im2 = c2.z, ii2 = c2.w;
            const float cinvN = c3.z;
            // normal limiter: projectors (n, -n), compMass = projector * invMass
            const float n2x = -nx, n2y = -ny;
            const float cm1x = nx * im1, cm1y = ny * im1, cm1a = aN1 * ii1;
            const float cm2x = n2x * im2, cm2y = n2y * im2, cm2a = aN2 * ii2;

            if (PHASE == 0)
            {
                const float4 c1 = __ldcs(&P.q1[s]);
                float2 acc = __ldcs(&P.accNF[s]);
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
                __stcs(&P.accNF[s], acc);
                productive = vmax(fabsf(dN), fabsf(dF)) > kProductiveImpulse;
            }
            else
            {
                float accD = __ldcs(&P.accD[s]);
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
                accD += d;
                __stcs(&P.accD[s], accD);
                productive = fabsf(d) > kProductiveImpulse;
            }

            // lastIteration = it where productive (Solver.cpp:903-910); static bodies via stamps
            if (!st1)
            {
                v1.w = __int_as_float(productive ? it : last1);
                __stcg(&rows[b1], v1);
            }
            else if (productive)
            {
                if (last1 <= it - 2) atomicAdd(&P.result[2], 1);
                stamp_write(&stamps[b1], tick);
            }
            if (!st2)
            {
                v2.w = __int_as_float(productive ? it : last2);
                __stcg(&rows[b2], v2);
            }
            else if (productive)
            {
                if (last2 <= it - 2) atomicAdd(&P.result[2], 1);
                stamp_write(&stamps[b2], tick);
            }
        }
        anyProductive |= productive;
    }
    return anyProductive;
}

// Persistent cooperative kernel: the whole SolveJointIsland loop nest (Solver.cpp:159-211) in one
// launch, one grid-wide barrier per level.
__global__ void __launch_bounds__(kBlock) k_solve(SolveParams P)
{
    cg::grid_group grid = cg::this_grid();
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;

    for (int l = 0; l < P.numLevels; ++l)
    {
        const Level L = P.levels[l];
        for (int s = L.start + tid; s < L.end; s += nthreads) prestep_slot(P, s);
        grid.sync();
    }

    int ranImpulse = 0;
    for (int it = 0; it < P.contactIters; ++it)
    {
        bool any = false;
        for (int l = 0; l < P.numLevels; ++l)
        {
            any |= solve_level<0>(P, P.levels[l], it, unsigned(it) * unsigned(P.numLevels) + unsigned(l) + 1u, tid, nthreads);
            if (l + 1 < P.numLevels) grid.sync();
        }
        if (__any_sync(0xffffffffu, any) && lane == 0) atomicOr(&P.flags[it], 1u);
        grid.sync();
        ranImpulse++;
        if (__ldcg(&P.flags[it]) == 0u) break;   // Solver.cpp:189
    }

    int ranDisp = 0;
    for (int it = 0; it < P.penetrationIters; ++it)
    {
        bool any = false;
        for (int l = 0; l < P.numLevels; ++l)
        {
            any |= solve_level<1>(P, P.levels[l], it, unsigned(it) * unsigned(P.numLevels) + unsigned(l) + 1u, tid, nthreads);
            if (l + 1 < P.numLevels) grid.sync();
        }
        if (__any_sync(0xffffffffu, any) && lane == 0) atomicOr(&P.flags[P.contactIters + it], 1u);
        grid.sync();
        ranDisp++;
        if (__ldcg(&P.flags[P.contactIters + it]) == 0u) break;   // Solver.cpp:210
    }
    if (tid == 0)
    {
        P.result[0] = ranImpulse;
        P.result[1] = ranDisp;
    }
}

// ---- host orchestration -----------------------------------------------------------------------------

static float elapsed(cudaEvent_t a, cudaEvent_t b)
{
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

// joints + contact points are resident (c->joints, c->contactPoints), schedule is resident.
int solve_run(phyx_b200_ctx* c, const phyx_b200_solve_config* cfg, phyx_b200_solve_stats* stats)
{
    const int ns = c->slotCount, nl = c->levelCount, nb = c->bodyCount;
    const int I = cfg->contactIterationsCount, D = cfg->penetrationIterationsCount;
    if (I < 0 || D < 0)
    {
        set_error("solve: negative iteration count");
        return PHYX_B200_ERR_ARGUMENT;
    }
    size_t ns1 = size_t(ns > 0 ? ns : 1);
    PHYX_TRY(c->q0.reserve(ns1 * sizeof(float4)));
    PHYX_TRY(c->q1.reserve(ns1 * sizeof(float4)));
    PHYX_TRY(c->q2.reserve(ns1 * sizeof(float4)));
    PHYX_TRY(c->q3.reserve(ns1 * sizeof(float4)));
    PHYX_TRY(c->accNF.reserve(ns1 * sizeof(float2)));
    PHYX_TRY(c->accD.reserve(ns1 * sizeof(float)));
    PHYX_TRY(c->stamps.reserve(size_t(nb > 0 ? nb : 1) * 2 * sizeof(unsigned long long)));
    size_t flagWords = size_t(I + D) + 8;
    PHYX_TRY(c->solveFlags.reserve(flagWords * sizeof(unsigned)));

    cudaEvent_t e0 = c->ev[0], e1 = c->ev[1], e2 = c->ev[2], e3 = c->ev[3];
    PHYX_CUDA(cudaEventRecord(e0, c->stream));
    int ranI = I > 0 ? 1 : 0, ranD = D > 0 ? 1 : 0, hazards = 0;
    if (ns > 0 && nl > 0)
    {
        PHYX_CUDA(cudaMemsetAsync(c->stamps.ptr, 0, size_t(nb) * 2 * sizeof(unsigned long long), c->stream));
        PHYX_CUDA(cudaMemsetAsync(c->solveFlags.ptr, 0, flagWords * sizeof(unsigned), c->stream));
        int grid = (ns + kBlock - 1) / kBlock;
        k_refresh<<<grid, kBlock, 0, c->stream>>>(ns, c->slotJoint.as<int>(), c->joints.as<phyx_contact_joint>(), c->contactPoints.as<float4>(),
            c->params.as<float4>(), c->q0.as<float4>(), c->q1.as<float4>(), c->q2.as<float4>(), c->q3.as<float4>(), c->accNF.as<float2>(),
            c->accD.as<float>());
        c->launches++;
        PHYX_CUDA(cudaEventRecord(e1, c->stream));

        SolveParams P;
        P.vel = c->vel.as<float4>();
        P.disp = c->disp.as<float4>();
        P.q0 = c->q0.as<float4>();
        P.q1 = c->q1.as<float4>();
        P.q2 = c->q2.as<float4>();
        P.q3 = c->q3.as<float4>();
        P.accNF = c->accNF.as<float2>();
        P.accD = c->accD.as<float>();
        P.levels = c->levels.as<Level>();
        P.numLevels = nl;
        P.stampImp = c->stamps.as<unsigned long long>();
        P.stampDisp = P.stampImp + nb;
        P.contactIters = I;
        P.penetrationIters = D;
        P.flags = c->solveFlags.as<unsigned>();
        P.result = reinterpret_cast<int*>(P.flags + (I + D));
        if (c->solveBlocksPerSM == 0)
        {
            int per = 0;
            PHYX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_solve, kBlock, 0));
            if (per < 1)
            {
                set_error("solve kernel does not fit on an SM");
                return PHYX_B200_ERR_CUDA;
            }
            c->solveBlocksPerSM = per;
        }
        // persistent grid: every SM full, but no more CTAs than the widest level can use
        int maxLevel = 0;
        for (const Level& L : c->hostLevels) maxLevel = max(maxLevel, L.end - L.start);
        int want = (maxLevel + kBlock - 1) / kBlock;
        int sgrid = max(1, min(want, c->numSMs * c->solveBlocksPerSM));
        void* args[] = { &P };
        PHYX_CUDA(cudaLaunchCooperativeKernel((void*)k_solve, dim3(sgrid), dim3(kBlock), args, 0, c->stream));
        c->launches++;
        PHYX_CUDA(cudaEventRecord(e2, c->stream));
        k_finish<<<grid, kBlock, 0, c->stream>>>(ns, c->slotJoint.as<int>(), c->accNF.as<float2>(), c->joints.as<phyx_contact_joint>());
        c->launches++;
        PHYX_CUDA(cudaEventRecord(e3, c->stream));
        int host[3];
        PHYX_CUDA(cudaMemcpyAsync(host, P.result, sizeof(host), cudaMemcpyDeviceToHost, c->stream));
        PHYX_CUDA(cudaStreamSynchronize(c->stream));
        ranI = host[0];
        ranD = host[1];
        hazards = host[2];
        if (stats)
        {
            stats->ms_refresh = elapsed(e0, e1);
            stats->ms_iterations = elapsed(e1, e2);
            stats->ms_finish = elapsed(e2, e3);
        }
    }
    else if (stats)
        stats->ms_refresh = stats->ms_iterations = stats->ms_finish = 0.f;
    if (stats)
    {
        stats->joints = c->jointCount;
        stats->slots = ns;
        stats->levels = nl;
        stats->contactIterationsRun = ranI;
        stats->penetrationIterationsRun = ranD;
        stats->staticHazards = hazards;
    }
    return PHYX_B200_OK;
}

} // namespace phyx
