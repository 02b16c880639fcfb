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
#include "barrier.cuh"
#include "solve_math.cuh"

#include <stdlib.h>

namespace phyx
{

constexpr int kBlock = 256;

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
    const int* slotPos;                // position of the slot's unit in the sequential order (null: the slot index)
    int* processed;                    // per slot: tick of the last pass that ran it (units with a static body only)
    unsigned long long* staticImp;     // per body: static-body lastIteration word, impulse phase
    unsigned long long* staticDisp;    //           ... displacement phase
    int contactIters, penetrationIters;
    unsigned long long* barrier;       // ring of 4 grid-barrier words
    int* result;                       // [0] impulse iterations run, [1] displacement iterations run, [2] extra wake passes
    unsigned long long* activeTotal;   // [2] joint-iterations relaxed (not skipped) per phase
    // strict companion schedule of a replay (schedule.cu StaticRule); numStrictLevels == 0: single schedule
    const Level* strictLevels;
    int numStrictLevels;
    // paired levels, record form (k_solve_pairs2): one 128-byte record and one index word pair per manifold
    const float4* pairQ;
    const int2* pairIdx;
    const int* strictMap;              // strict position -> slot (or -1)
    const unsigned char* rowsMulti;    // per body row: static body with at least two units
    int numMultiStatics;
    int* hotCount;                     // [2 phases][4]: multi-unit static bodies productive in iteration it, at [it % 3]
};

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

// ---- PrepareBodies / FinishBodies, Solver.cpp:456-494 ---------------------------------------------------
// Like the reference, the solve works on its own packed copy of the body rows {v, w, lastIteration}.
// Here the copy is also the place to fix memory locality: rows are stored in the broadphase's
// sorted-x order (order[i] = body at sorted position i), so joints that are neighbours in the slot
// order (sweep order) gather neighbouring rows: the gather / scatter of body rows through L1 is what
// bounds the iterations (one 32-byte sector per lane when the rows are scattered).
__global__ void __launch_bounds__(kBlock) k_prepare_bodies(Count nc, const unsigned* __restrict__ order, const float4* __restrict__ vel,
    const float4* __restrict__ disp, float4* __restrict__ rowsVel, float4* __restrict__ rowsDisp, const unsigned char* __restrict__ multi,
    unsigned char* __restrict__ rowsMulti)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count_of(nc)) return;
    const unsigned b = order ? order[i] : unsigned(i);
    if (multi) rowsMulti[i] = multi[b];
    float4 v = vel[b], d = disp[b];
    v.w = __int_as_float(-1);
    d.w = __int_as_float(-1);
    rowsVel[i] = v;
    rowsDisp[i] = d;
}

__global__ void __launch_bounds__(kBlock) k_finish_bodies(Count nc, const unsigned* __restrict__ order, const float4* __restrict__ rowsVel,
    const float4* __restrict__ rowsDisp, float4* __restrict__ vel, float4* __restrict__ disp, int* __restrict__ activity)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count_of(nc)) return;
    const unsigned b = order ? order[i] : unsigned(i);
    if (activity) activity[b] = __float_as_int(rowsVel[i].w);   // last productive impulse iteration: next step's strip balance
    vel[b] = rowsVel[i];
    disp[b] = rowsDisp[i];
}

// ---- PrepareJoints copy + RefreshJoints ----------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_refresh(Count numSlotsC, const int* __restrict__ slotJoint, const phyx_contact_joint* __restrict__ joints,
    const float4* __restrict__ contactPoints, const float4* __restrict__ params, const int* __restrict__ rowOf, float4* __restrict__ q0,
    float4* __restrict__ q1, float4* __restrict__ q2, float4* __restrict__ q3, float2* __restrict__ accNF, float* __restrict__ accD,
    float4* __restrict__ pairQ, int2* __restrict__ pairIdx, int firstSlot, bool writeIdx)
{
    int s = firstSlot + blockIdx.x * blockDim.x + threadIdx.x;   // slots [firstSlot, numSlots)
    if (s >= count_of(numSlotsC)) return;
    int j = slotJoint[s];
    if (j < 0)
    {
        if (!pairQ)
            q3[s] = make_float4(__int_as_float(-1), __int_as_float(-1), 0.f, 0.f);
        else if (!(s & 1) && writeIdx)
            pairIdx[s >> 1] = make_int2(-1, -1);
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

    // the packed joint addresses the solver's row copy (rowOf: body -> row, see k_prepare_bodies)
    int b1 = (rowOf ? rowOf[jt.body1Index] : jt.body1Index) | ((p1.x == 0.0f && p1.y == 0.0f) ? kStaticBit : 0);
    int b2 = (rowOf ? rowOf[jt.body2Index] : jt.body2Index) | ((p2.x == 0.0f && p2.y == 0.0f) ? kStaticBit : 0);

    if (pairQ)
    {
        // record form (PairRecord below): six float4 per manifold, a = slot 2p, b = slot 2p+1
        const int h = s & 1;
        float4* rec = pairQ + size_t(s >> 1) * kPairRecordWords;
        rec[h ? 1 : 0] = make_float4(nx, ny, aN1, aN2);
        rec[h ? 5 : 4] = make_float4(aF1, aF2, cinvF, dstVel);
        reinterpret_cast<float2*>(rec + 3)[h] = make_float2(cinvN, dstDisp);
        if (!h)
        {
            rec[2] = make_float4(p1.x, p1.y, p2.x, p2.y);   // both joints of a manifold have the same two bodies
            if (writeIdx) pairIdx[s >> 1] = make_int2(b1, b2 | (slotJoint[s + 1] >= 0 ? kPairHasB : 0));   // numSlots is a multiple of 64
        }
    }
    else
    {
        q0[s] = make_float4(nx, ny, aN1, aN2);
        q1[s] = make_float4(aF1, aF2, cinvF, dstVel);
        q2[s] = make_float4(p1.x, p1.y, p2.x, p2.y);
        q3[s] = make_float4(__int_as_float(b1), __int_as_float(b2), cinvN, dstDisp);
    }
    accNF[s] = make_float2(jt.normalLimiter_accumulatedImpulse, jt.frictionLimiter_accumulatedImpulse);
    accD[s] = 0.0f;
}

// ---- FinishJoints ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_finish(Count numSlotsC, const int* __restrict__ slotJoint, const float2* __restrict__ accNF,
    phyx_contact_joint* __restrict__ joints)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= count_of(numSlotsC)) return;
    int j = slotJoint[s];
    if (j < 0) return;
    float2 a = accNF[s];
    joints[j].normalLimiter_accumulatedImpulse = a.x;
    joints[j].frictionLimiter_accumulatedImpulse = a.y;
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


// ---- one pass over one level of one iteration -----------------------------------------------------------
// The read-only joint streams of a slot (and its accumulators, which only this thread ever writes)
// do not depend on what other CTAs do, so they are loaded AHEAD of the grid barrier that opens the
// level; after the barrier only the two body rows (L2) stand between a thread and its arithmetic.
// All streams are fetched speculatively, before the skip test: the kernel is latency-bound, not
// bandwidth-bound (ncu: DRAM ~10 % busy), so spending bytes to shorten the dependent chain wins.
template <int PHASE>
struct SlotData
{
    float4 c0, c1, c2, c3;   // c1: impulse phase only
    float2 acc;              // impulse: {accN, accF}; displacement: {accD, -}
};

#ifndef PHYX_SOLVE_SPECULATIVE
#define PHYX_SOLVE_SPECULATIVE 1   // 1: fetch every stream before the skip test; 0: only Q3, the rest once the joint is known to be active
#endif

template <int PHASE>
__device__ __forceinline__ void load_rest(const SolveParams& P, int s, SlotData<PHASE>& d)
{
    d.c0 = __ldcs(&P.q0[s]);
    d.c2 = __ldcs(&P.q2[s]);
    if (PHASE == 0)
    {
        d.c1 = __ldcs(&P.q1[s]);
        d.acc = __ldcs(&P.accNF[s]);
    }
    else
        d.acc.x = __ldcs(&P.accD[s]);
}

template <int PHASE>
__device__ __forceinline__ void load_slot(const SolveParams& P, int s, bool inRange, SlotData<PHASE>& d)
{
    d.c3 = make_float4(__int_as_float(-1), __int_as_float(-1), 0.f, 0.f);
    if (inRange)
    {
        d.c3 = __ldcs(&P.q3[s]);
        if (PHYX_SOLVE_SPECULATIVE) load_rest<PHASE>(P, s, d);
    }
}

// PHASE 0: SolveJointsImpulses (Solver.cpp:781-910); PHASE 1: SolveJointsDisplacement (:937-1014).
// firstPass = false is a wake pass: only units that contain a static body and have not run yet in
// this (iteration, level) are reconsidered.  Returns productive; sets `wake` if a joint of this
// pass turned a cold static body productive.  `pre` holds the streams of this thread's first slot
// when havePre is set.
template <int PHASE, bool DUAL>
__device__ __forceinline__ bool solve_level(const SolveParams& P, const Level L, const int* __restrict__ slotMapArg, int it, int tick, bool firstPass, int tid,
    int nthreads, bool& wake, unsigned& activeCount, SlotData<PHASE>& pre, bool havePre)
{
    const int* __restrict__ slotMap = DUAL ? slotMapArg : nullptr;   // compile-time null for single-schedule kernels
    float4* rows = PHASE == 0 ? P.vel : P.disp;
    unsigned long long* statics = PHASE == 0 ? P.staticImp : P.staticDisp;
    int* hotCounter = (DUAL && P.hotCount) ? P.hotCount + PHASE * 4 + (it % 3) : nullptr;
    const int lane = threadIdx.x & 31;
    const unsigned seg = 0xffu << (lane & ~7);   // the 8-lane unit this lane belongs to
    bool anyProductive = false;

    for (int s0 = L.start + (tid & ~31); s0 < L.end; s0 += nthreads)   // warp-uniform trip count
    {
        const int k = s0 + lane;   // position in the level; the slot it names goes through slotMap if there is one
        bool valid = k < L.end;
        int s = k;
        if (slotMap && valid)
        {
            s = slotMap[k];
            valid = s >= 0;
        }
        if (!havePre) load_slot<PHASE>(P, s, valid, pre);
        havePre = false;
        const float4 c3 = pre.c3;
        const int r1 = __float_as_int(c3.x), r2 = __float_as_int(c3.y);
        valid = valid && r1 >= 0;
        const int b1 = r1 & kBodyMask, b2 = r2 & kBodyMask;
        const bool st1 = valid && (r1 & kStaticBit), st2 = valid && (r2 & kStaticBit);
        const bool wide = k < L.grouped_end;

        // units with a static body are the only ones a wake pass can affect
        const unsigned mStatic = __ballot_sync(0xffffffffu, st1 || st2);
        const bool unitHasStatic = wide ? (mStatic & seg) != 0 : (st1 || st2);
        if (!firstPass)
        {
            valid = valid && unitHasStatic;
            if (valid) valid = __ldcg(&P.processed[s]) != tick;
        }

        float4 v1 = make_float4(0.f, 0.f, 0.f, 0.f), v2 = v1;
        int last1 = -1, last2 = -1;
        unsigned pos = 0;
        bool active = false;
        if (valid)
        {
            v1 = __ldcg(&rows[b1]);
            v2 = __ldcg(&rows[b2]);
            if (st1 || st2) pos = P.slotPos ? unsigned(P.slotPos[s]) : unsigned(s);
            last1 = st1 ? static_visible_last(&statics[b1], it, pos) : __float_as_int(v1.w);
            last2 = st2 ? static_visible_last(&statics[b2], it, pos) : __float_as_int(v2.w);
            active = (last1 > it - 2) || (last2 > it - 2);   // Solver.cpp:790-792
        }
        const unsigned mActive = __ballot_sync(0xffffffffu, active);
        if (wide) active = valid && (mActive & seg) != 0;   // AVX2: skip only if none of the 8 lanes is active (:797)

        bool productive = false;
        if (active)
        {
            ++activeCount;
            if (!PHYX_SOLVE_SPECULATIVE) load_rest<PHASE>(P, s, pre);
            float2 acc = pre.acc;
            productive = relax<PHASE>(pre.c0, pre.c1, pre.c2, c3, acc, v1, v2, wide);
            if (PHASE == 0)
                __stcs(&P.accNF[s], acc);
            else
                __stcs(&P.accD[s], acc.x);

            // lastIteration = it where productive (Solver.cpp:903-910)
            if (!st1)
            {
                v1.w = __int_as_float(productive ? it : last1);
                __stcg(&rows[b1], v1);
            }
            else if (productive)
                wake |= static_mark(&statics[b1], it, pos, (DUAL && P.rowsMulti && P.rowsMulti[b1]) ? hotCounter : nullptr);
            if (!st2)
            {
                v2.w = __int_as_float(productive ? it : last2);
                __stcg(&rows[b2], v2);
            }
            else if (productive)
                wake |= static_mark(&statics[b2], it, pos, (DUAL && P.rowsMulti && P.rowsMulti[b2]) ? hotCounter : nullptr);
            if (unitHasStatic) __stcg(&P.processed[s], tick);
        }
        anyProductive |= productive;
    }
    return anyProductive;
}

// All iterations of one phase (Solver.cpp:175-190 / :196-211), one grid barrier per level plus one per
// wake pass (rare).  Returns the number of iterations run.
template <int PHASE, bool DUAL>
__device__ __forceinline__ int run_phase(const SolveParams& P, int iters, int tid, int nthreads, unsigned& epoch, int& tick, int& wakePasses,
    unsigned& activeCount)
{
    SlotData<PHASE> pre;
    bool havePre = false;
    int ran = 0;
    const bool dual = DUAL && P.numStrictLevels > 0;
    int* hot = P.hotCount ? P.hotCount + PHASE * 4 : nullptr;
    for (int it = 0; it < iters; ++it)
    {
        // Replay schedules come in two forms (schedule.cu StaticRule).  The fast one ignores static bodies
        // and is exact whenever every static body with several joints is "hot" (was productive in the
        // previous iteration: then all its joints are active whatever their order; iteration 0 is always
        // hot).  Otherwise this iteration runs on the strict form.
        bool strict = false;
        if (dual)
        {
            if (it > 0) strict = __ldcg(&hot[(it - 1) % 3]) < P.numMultiStatics;
            if (tid == 0) hot[(it + 1) % 3] = 0;
            havePre = false;   // the slot prefetched across the iteration boundary may belong to the other form
        }
        const Level* levels = strict ? P.strictLevels : P.levels;
        const int numLevels = strict ? P.numStrictLevels : P.numLevels;
        const int* slotMap = strict ? P.strictMap : nullptr;

        bool any = false, productiveAnywhere = false;
        for (int l = 0; l < numLevels; ++l)
        {
            const Level L = levels[l];
            ++tick;
            bool wake = false;
            any |= solve_level<PHASE, DUAL>(P, L, slotMap, it, tick, true, tid, nthreads, wake, activeCount, pre, havePre);
            // streams of this thread's first slot in the level that follows (next level, or level 0 of
            // the next iteration), fetched while the grid drains into the barrier
            havePre = false;
            if (l + 1 < numLevels || !dual)
            {
                const Level N = levels[l + 1 < numLevels ? l + 1 : 0];
                const int kN = N.start + tid;
                bool inRange = kN < N.end;
                int sN = kN;
                if (DUAL && slotMap && inRange)
                {
                    sN = slotMap[kN];
                    inRange = sN >= 0;
                }
                load_slot<PHASE>(P, sN, inRange, pre);
                havePre = true;
            }
            BarrierResult r = grid_barrier(P.barrier, epoch, wake, any);
            while (r.wake)
            {
                SlotData<PHASE> scratch;
                wake = false;
                any |= solve_level<PHASE, DUAL>(P, L, slotMap, it, tick, false, tid, nthreads, wake, activeCount, scratch, false);
                ++wakePasses;
                r = grid_barrier(P.barrier, epoch, wake, any);
            }
            productiveAnywhere = r.productive;
        }
        ++ran;
        if (strict) ++wakePasses;   // reported together: iterations that needed the strict form
        if (!productiveAnywhere) break;   // Solver.cpp:189 / :210
    }
    return ran;
}

// Persistent cooperative kernel: the whole SolveJointIsland loop nest (Solver.cpp:159-211) in one launch.
// THREADS x MIN_BLOCKS is the register budget: 256x4, 512x2 and 1024x1 all give 64 registers and
// 1024 threads per SM; fewer, larger CTAs make the grid barrier cheaper (148 arrivals instead of 592).
template <int THREADS, int MIN_BLOCKS, bool DUAL>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS) k_solve(SolveParams P)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    unsigned epoch = 0;
    int wakePasses = 0, tick = 0;
    unsigned active[2] = { 0u, 0u };

    for (int l = 0; l < P.numLevels; ++l)
    {
        const Level L = P.levels[l];
        for (int s = L.start + tid; s < L.end; s += nthreads) prestep_slot(P, s);
        grid_barrier(P.barrier, epoch, false, false);
    }

    const int ranImpulse = run_phase<0, DUAL>(P, P.contactIters, tid, nthreads, epoch, tick, wakePasses, active[0]);
    const int ranDisplacement = run_phase<1, DUAL>(P, P.penetrationIters, tid, nthreads, epoch, tick, wakePasses, active[1]);

    for (int phase = 0; phase < 2; ++phase)
    {
        unsigned v = active[phase];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&P.activeTotal[phase], static_cast<unsigned long long>(v));
    }
    if (tid == 0)
    {
        P.result[0] = ranImpulse;
        P.result[1] = ranDisplacement;
        P.result[2] = wakePasses;
    }
}

// =====================================================================================================
// Paired levels (manifold-unit colouring, colour.cu)
// =====================================================================================================
// A paired level holds whole manifolds: slots 2u and 2u+1 are the two joints of one body pair (2u+1 is
// empty when the manifold has one contact point).  No other unit of the level touches these two bodies,
// so ONE thread relaxes both joints back to back on the same two body rows held in registers: half the
// levels (grid barriers) per iteration, and half the row gathers / scatters per joint.  The result is
// that of the sequential sweep in slot order: joint 2u+1 follows 2u directly, it sees exactly the rows
// 2u left, and since both joints have the same bodies the skip test (Solver.cpp:790-798) gives the
// same answer for both: if 2u runs, lastIteration of its bodies can only have grown, so 2u+1 runs;
// if 2u is skipped nothing changed, so 2u+1 is skipped.
template <int PHASE>
struct PairData
{
    float4 a0, a1, a2, a3;   // streams of slot 2u
    float4 b0, b1, b2, b3;   // streams of slot 2u+1
    float4 acc;              // impulse: {accN, accF} of 2u, {accN, accF} of 2u+1; displacement: {accD 2u, accD 2u+1, -, -}
};

template <int PHASE>
__device__ __forceinline__ void load_pair(const SolveParams& P, int s, bool inRange, PairData<PHASE>& d)
{
    d.a3 = make_float4(__int_as_float(-1), __int_as_float(-1), 0.f, 0.f);
    if (inRange)
    {
        d.a3 = __ldcs(&P.q3[s]);
        d.b3 = __ldcs(&P.q3[s + 1]);
        d.a0 = __ldcs(&P.q0[s]);
        d.b0 = __ldcs(&P.q0[s + 1]);
        d.a2 = __ldcs(&P.q2[s]);
        d.b2 = __ldcs(&P.q2[s + 1]);
        if (PHASE == 0)
        {
            d.a1 = __ldcs(&P.q1[s]);
            d.b1 = __ldcs(&P.q1[s + 1]);
            d.acc = __ldcs(reinterpret_cast<const float4*>(&P.accNF[s]));   // s is even: 16-byte aligned
        }
        else
        {
            const float2 a = __ldcs(reinterpret_cast<const float2*>(&P.accD[s]));
            d.acc = make_float4(a.x, a.y, 0.f, 0.f);
        }
    }
}

__device__ __forceinline__ void prestep_pair(const SolveParams& P, int s)
{
    const float4 c3 = __ldcs(&P.q3[s]);
    const int r1 = __float_as_int(c3.x), r2 = __float_as_int(c3.y);
    if (r1 < 0) return;
    const int b1 = r1 & kBodyMask, b2 = r2 & kBodyMask;
    float4 v1 = __ldcg(&P.vel[b1]), v2 = __ldcg(&P.vel[b2]);
    const float4 accs = __ldcs(reinterpret_cast<const float4*>(&P.accNF[s]));
    const bool haveB = __float_as_int(__ldcs(&P.q3[s + 1]).x) >= 0;
#pragma unroll
    for (int h = 0; h < 2; ++h)
    {
        if (h == 1 && !haveB) break;
        const float4 c0 = __ldcs(&P.q0[s + h]), c1 = __ldcs(&P.q1[s + h]), c2 = __ldcs(&P.q2[s + h]);
        const float nx = c0.x, ny = c0.y;
        const float accN = h ? accs.z : accs.x, accF = h ? accs.w : accs.y;
        v1.x += (nx * c2.x) * accN;
        v1.y += (ny * c2.x) * accN;
        v1.z += (c0.z * c2.y) * accN;
        v2.x += ((-nx) * c2.z) * accN;
        v2.y += ((-ny) * c2.z) * accN;
        v2.z += (c0.w * c2.w) * accN;
        const float tx = -ny, ty = nx;
        v1.x += (tx * c2.x) * accF;
        v1.y += (ty * c2.x) * accF;
        v1.z += (c1.x * c2.y) * accF;
        v2.x += ((-tx) * c2.z) * accF;
        v2.y += ((-ty) * c2.z) * accF;
        v2.z += (c1.y * c2.w) * accF;
    }
    if (!(r1 & kStaticBit)) __stcg(&P.vel[b1], v1);
    if (!(r2 & kStaticBit)) __stcg(&P.vel[b2], v2);
}

template <int PHASE>
__device__ __forceinline__ bool solve_pairs(const SolveParams& P, const Level L, int it, int tick, bool firstPass, int tid, int nthreads, bool& wake,
    unsigned& activeCount, PairData<PHASE>& pre, bool havePre)
{
    float4* rows = PHASE == 0 ? P.vel : P.disp;
    unsigned long long* statics = PHASE == 0 ? P.staticImp : P.staticDisp;
    bool anyProductive = false;
    for (int s = L.start + 2 * tid; s < L.end; s += 2 * nthreads)
    {
        if (!havePre) load_pair<PHASE>(P, s, true, pre);
        havePre = false;
        const int r1 = __float_as_int(pre.a3.x), r2 = __float_as_int(pre.a3.y);
        if (r1 < 0) continue;
        const int b1 = r1 & kBodyMask, b2 = r2 & kBodyMask;
        const bool st1 = r1 & kStaticBit, st2 = r2 & kStaticBit;
        if (!firstPass && (!(st1 || st2) || __ldcg(&P.processed[s]) == tick)) continue;   // wake pass: only pairs a static body can wake

        float4 v1 = __ldcg(&rows[b1]), v2 = __ldcg(&rows[b2]);
        const unsigned pos = unsigned(s);
        const int last1 = st1 ? static_visible_last(&statics[b1], it, pos) : __float_as_int(v1.w);
        const int last2 = st2 ? static_visible_last(&statics[b2], it, pos) : __float_as_int(v2.w);
        if (!((last1 > it - 2) || (last2 > it - 2))) continue;   // Solver.cpp:790-792, for both joints (see above)

        const bool haveB = __float_as_int(pre.b3.x) >= 0;
        activeCount += haveB ? 2u : 1u;
        float2 accA = make_float2(pre.acc.x, PHASE == 0 ? pre.acc.y : 0.f);
        float2 accB = PHASE == 0 ? make_float2(pre.acc.z, pre.acc.w) : make_float2(pre.acc.y, 0.f);
        const bool productiveA = relax<PHASE>(pre.a0, pre.a1, pre.a2, pre.a3, accA, v1, v2, false);
        bool productiveB = false;
        if (haveB) productiveB = relax<PHASE>(pre.b0, pre.b1, pre.b2, pre.b3, accB, v1, v2, false);
        if (PHASE == 0)
            __stcs(reinterpret_cast<float4*>(&P.accNF[s]), make_float4(accA.x, accA.y, accB.x, accB.y));
        else
            __stcs(reinterpret_cast<float2*>(&P.accD[s]), make_float2(accA.x, accB.x));

        // lastIteration = it where productive (Solver.cpp:903-910); a static body is marked at the position
        // of the first productive joint
        const bool productive = productiveA || productiveB;
        const unsigned markPos = productiveA ? pos : pos + 1;
        if (!st1)
        {
            v1.w = __int_as_float(productive ? it : last1);
            __stcg(&rows[b1], v1);
        }
        else if (productive)
            wake |= static_mark(&statics[b1], it, markPos, nullptr);
        if (!st2)
        {
            v2.w = __int_as_float(productive ? it : last2);
            __stcg(&rows[b2], v2);
        }
        else if (productive)
            wake |= static_mark(&statics[b2], it, markPos, nullptr);
        if (st1 || st2) __stcg(&P.processed[s], tick);
        anyProductive |= productive;
    }
    return anyProductive;
}

template <int PHASE>
__device__ __forceinline__ int run_phase_pairs(const SolveParams& P, int iters, int tid, int nthreads, unsigned& epoch, int& tick, int& wakePasses,
    unsigned& activeCount)
{
    PairData<PHASE> pre;
    bool havePre = false;
    int ran = 0;
    for (int it = 0; it < iters; ++it)
    {
        bool any = false, productiveAnywhere = false;
        for (int l = 0; l < P.numLevels; ++l)
        {
            const Level L = P.levels[l];
            ++tick;
            bool wake = false;
            any |= solve_pairs<PHASE>(P, L, it, tick, true, tid, nthreads, wake, activeCount, pre, havePre);
            unsigned long long ticket;
            grid_arrive(P.barrier, epoch, wake, any, ticket);
            // streams of this thread's first pair of the level that follows, fetched while the grid drains into the barrier
            const Level N = P.levels[l + 1 < P.numLevels ? l + 1 : 0];
            const int sN = N.start + 2 * tid;
            havePre = sN < N.end;
            if (havePre) load_pair<PHASE>(P, sN, true, pre);
            BarrierResult r = grid_wait(P.barrier, epoch, ticket);
            while (r.wake)
            {
                PairData<PHASE> scratch;
                wake = false;
                any |= solve_pairs<PHASE>(P, L, it, tick, false, tid, nthreads, wake, activeCount, scratch, false);
                ++wakePasses;
                r = grid_barrier(P.barrier, epoch, wake, any);
            }
            productiveAnywhere = r.productive;
        }
        ++ran;
        if (!productiveAnywhere) break;   // Solver.cpp:189 / :210
    }
    return ran;
}

template <int THREADS, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS) k_solve_pairs(SolveParams P)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    unsigned epoch = 0;
    int wakePasses = 0, tick = 0;
    unsigned active[2] = { 0u, 0u };

    for (int l = 0; l < P.numLevels; ++l)
    {
        const Level L = P.levels[l];
        for (int s = L.start + 2 * tid; s < L.end; s += 2 * nthreads) prestep_pair(P, s);
        grid_barrier(P.barrier, epoch, false, false);
    }

    const int ranImpulse = run_phase_pairs<0>(P, P.contactIters, tid, nthreads, epoch, tick, wakePasses, active[0]);
    const int ranDisplacement = run_phase_pairs<1>(P, P.penetrationIters, tid, nthreads, epoch, tick, wakePasses, active[1]);

    for (int phase = 0; phase < 2; ++phase)
    {
        unsigned v = active[phase];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&P.activeTotal[phase], static_cast<unsigned long long>(v));
    }
    if (tid == 0)
    {
        P.result[0] = ranImpulse;
        P.result[1] = ranDisplacement;
        P.result[2] = wakePasses;
    }
}

// =====================================================================================================
// Paired levels, record form: activity prediction instead of speculative streaming
// =====================================================================================================
// Measured on the 1 M-box pyramid (profiles/k_solve_r1g_metrics.json): k_solve_pairs moves 6.2 GB per
// launch at 61 % of the measured copy bandwidth, yet only ~30 % of the joint-iterations pass the
// lastIteration test (Solver.cpp:790-798): most of the streamed bytes belong to joints that are
// skipped.  This form reads, for every manifold, only an 8-byte index pair {row1, row2 | hasB} and the
// two body rows (L2) to decide; the 128-byte record (one cache line: both joints of the manifold) is
// fetched for active manifolds only.  Activity is strongly correlated between consecutive iterations, so
// each thread keeps one bit per manifold it owns (a manifold is visited by the same thread in every
// iteration) and, while the grid drains into the level barrier, prefetches into L2 the records of the
// next level's manifolds that were active last time: the DRAM latency of the records overlaps the
// barrier instead of sitting behind the skip test.  A thread handles its manifolds of a level in
// batches of kPairU: all index words first, then all row gathers (2 * kPairU independent L2 loads in
// flight), then the tests and the relaxations; no two manifolds of a level share a dynamic body, so
// gathering the rows of a whole batch up front reads nothing stale.  The arithmetic and the order of
// relaxations are those of solve_pairs, hence the same results bit for bit.

// Records of the manifolds that pass the skip test are staged through shared memory with asynchronous copies
// (LDGSTS): a thread issues the copies for ALL its active manifolds of a batch at once and waits once, instead
// of paying one DRAM round trip per manifold in sequence (measured: the chain of the busiest thread, not
// bandwidth, bounds a level pass).  Layout [manifold of the batch][float4 of the record][thread]: conflict-free.
__device__ __forceinline__ void cp_async16(float4* smemDst, const float4* src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(smemDst))), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

constexpr int kPairU = 4;   // manifolds a thread handles together (default; the 1024-threads-per-SM shape uses 2)



__device__ __forceinline__ void prestep_pair2(const SolveParams& P, int p)
{
    const int2 idx = __ldcs(&P.pairIdx[p]);
    if (idx.x < 0) return;
    const int b1 = idx.x & kBodyMask, b2 = idx.y & kBodyMask;
    const bool haveB = idx.y < 0;
    float4 v1 = __ldcg(&P.vel[b1]), v2 = __ldcg(&P.vel[b2]);
    const float4 accs = __ldcs(reinterpret_cast<const float4*>(&P.accNF[2 * p]));
    const float4* rec = P.pairQ + size_t(p) * kPairRecordWords;
    const float4 c2 = __ldcs(rec + 2);
#pragma unroll
    for (int h = 0; h < 2; ++h)
    {
        if (h == 1 && !haveB) break;
        const float4 c0 = __ldcs(rec + (h ? 1 : 0)), c1 = __ldcs(rec + (h ? 5 : 4));
        const float nx = c0.x, ny = c0.y;
        const float accN = h ? accs.z : accs.x, accF = h ? accs.w : accs.y;
        v1.x += (nx * c2.x) * accN;
        v1.y += (ny * c2.x) * accN;
        v1.z += (c0.z * c2.y) * accN;
        v2.x += ((-nx) * c2.z) * accN;
        v2.y += ((-ny) * c2.z) * accN;
        v2.z += (c0.w * c2.w) * accN;
        const float tx = -ny, ty = nx;
        v1.x += (tx * c2.x) * accF;
        v1.y += (ty * c2.x) * accF;
        v1.z += (c1.x * c2.y) * accF;
        v2.x += ((-tx) * c2.z) * accF;
        v2.y += ((-ty) * c2.z) * accF;
        v2.z += (c1.y * c2.w) * accF;
    }
    if (!(idx.x & kStaticBit)) __stcg(&P.vel[b1], v1);
    if (!(idx.y & kStaticBit)) __stcg(&P.vel[b2], v2);
}

// Which manifold of a level a thread handles as its k-th one.  Activity is clustered (a moving region of the pile is
// a contiguous run of manifolds of a level), and a level pass takes as long as its busiest CTA: with a plain
// "thread id + k * threads" mapping a CTA owns runs of 256 consecutive manifolds and some CTAs get three or four hot
// runs while the average is one.  Here consecutive 32-manifold chunks go to consecutive CTAs, and a CTA's consecutive
// chunks to its consecutive warps, so a hot run of any length is dealt evenly over all CTAs and over the warps of
// each; a warp still reads 32 consecutive index words / records.
__device__ __forceinline__ long long pair_of(int k)
{
    const long long chunk = static_cast<long long>(k * int(blockDim.x >> 5) + int(threadIdx.x >> 5)) * gridDim.x + blockIdx.x;
    return chunk * 32 + (threadIdx.x & 31);
}
// true while some thread of the grid still has a k-th manifold in a level of numPairs
__device__ __forceinline__ bool pairs_left(int k, int numPairs)
{
    return static_cast<long long>(k) * int(blockDim.x >> 5) * gridDim.x * 32 < numPairs;
}

// One pass over one paired level.  `pre` holds the index words of this thread's first batch when havePre
// is set.  First passes record which manifolds were active in `activity` (bit = bitCursor + position in
// the thread's visiting order, first 64 only) and advance bitCursor; wake passes leave both alone.
template <int PHASE, int U>
__device__ __forceinline__ bool solve_pairs2(const SolveParams& P, const Level L, int it, int tick, bool firstPass, int tid, int nthreads, bool& wake,
    unsigned& activeCount, const int2 (&pre)[U], bool havePre, unsigned& bitCursor, unsigned long long& activity, float4* stage)
{
    float4* rows = PHASE == 0 ? P.vel : P.disp;
    unsigned long long* statics = PHASE == 0 ? P.staticImp : P.staticDisp;
    const int pairBase = L.start >> 1, numPairs = (L.end - L.start) >> 1;
    bool anyProductive = false;
    for (int k0 = 0; pairs_left(k0, numPairs); k0 += U)
    {
        int2 idx[U];
        float4 v1[U], v2[U];
        int pv[U];   // global manifold index (pairBase + position in the level)
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            const long long p = pair_of(k0 + u);
            pv[u] = pairBase + int(p < numPairs ? p : 0);
            if (havePre)
                idx[u] = pre[u];
            else
                idx[u] = p < numPairs ? __ldcs(&P.pairIdx[pv[u]]) : make_int2(-1, -1);
        }
        havePre = false;
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            bool valid = idx[u].x >= 0;
            if (!firstPass && valid)   // wake pass: only manifolds a static body can wake, and only those that have not run yet
            {
                valid = ((idx[u].x | idx[u].y) & kStaticBit) && __ldcg(&P.processed[2 * pv[u]]) != tick;
            }
            if (!valid) idx[u].x = -1;
            v1[u] = v2[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid)
            {
                v1[u] = __ldcg(&rows[idx[u].x & kBodyMask]);
                v2[u] = __ldcg(&rows[idx[u].y & kBodyMask]);
            }
        }
        // skip test (Solver.cpp:790-792) for the whole batch, for both joints of a manifold at once (see
        // solve_pairs).  Testing every manifold of the batch before relaxing any is exact: a manifold that
        // an earlier one of this batch wakes up through a static body is caught by the wake pass, like one
        // woken by another thread.  (Pulling the records of the active ones into L1 at this point with
        // LDGSTS.ca was measured: 11.3 instead of 10.4 us per level pass, the L1 does not survive.)
        int last1v[U], last2v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            const int r1 = idx[u].x, r2 = idx[u].y;
            if (r1 < 0) continue;
            const int p = pv[u];
            const unsigned pos = unsigned(2 * p);
            last1v[u] = (r1 & kStaticBit) ? static_visible_last(&statics[r1 & kBodyMask], it, pos) : __float_as_int(v1[u].w);
            last2v[u] = (r2 & kStaticBit) ? static_visible_last(&statics[r2 & kBodyMask], it, pos) : __float_as_int(v2[u].w);
            const bool active = ((last1v[u] > it - 2) || (last2v[u] > it - 2));
            if (!active)
            {
                idx[u].x = -1;
                continue;
            }
            // stage the record: [u][word][thread]
            const float4* rec = P.pairQ + size_t(p) * kPairRecordWords;
#pragma unroll
            for (int f = 0; f < (PHASE == 0 ? kPairRecordWords : 4); ++f) cp_async16(stage + (u * kPairRecordWords + f) * blockDim.x, rec + f);
        }
        // the accumulators go to registers while the copies are in flight
        float4 accv[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            accv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx[u].x < 0) continue;
            const int s = 2 * pv[u];
            if (PHASE == 0)
                accv[u] = __ldcs(reinterpret_cast<const float4*>(&P.accNF[s]));
            else
            {
                const float2 a = __ldcs(reinterpret_cast<const float2*>(&P.accD[s]));
                accv[u] = make_float4(a.x, a.y, 0.f, 0.f);
            }
        }
        cp_async_wait_all();
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            const int r1 = idx[u].x, r2 = idx[u].y;
            if (r1 < 0) continue;
            const int s = 2 * pv[u];
            const int b1 = r1 & kBodyMask, b2 = r2 & kBodyMask;
            const bool st1 = r1 & kStaticBit, st2 = r2 & kStaticBit, haveB = r2 < 0;
            const unsigned pos = unsigned(s);
            const int last1 = last1v[u], last2 = last2v[u];

            if (firstPass && bitCursor + u < 64u) activity |= 1ull << (bitCursor + u);
            const float4* mine = stage + u * kPairRecordWords * blockDim.x;
            const float4 a0 = mine[0], b0 = mine[blockDim.x], a2 = mine[2 * blockDim.x], nd = mine[3 * blockDim.x];
            float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), b1r = a1;
            if (PHASE == 0)
            {
                a1 = mine[4 * blockDim.x];
                b1r = mine[5 * blockDim.x];
            }
            const float4 a3 = make_float4(0.f, 0.f, nd.x, nd.y), b3 = make_float4(0.f, 0.f, nd.z, nd.w);
            float2 accA, accB;
            if (PHASE == 0)
            {
                accA = make_float2(accv[u].x, accv[u].y);
                accB = make_float2(accv[u].z, accv[u].w);
            }
            else
            {
                accA = make_float2(accv[u].x, 0.f);
                accB = make_float2(accv[u].y, 0.f);
            }
            activeCount += haveB ? 2u : 1u;
            float4 w1 = v1[u], w2 = v2[u];
            const bool productiveA = relax<PHASE>(a0, a1, a2, a3, accA, w1, w2, false);
            bool productiveB = false;
            if (haveB) productiveB = relax<PHASE>(b0, b1r, a2, b3, accB, w1, w2, false);
            if (PHASE == 0)
                __stcs(reinterpret_cast<float4*>(&P.accNF[s]), make_float4(accA.x, accA.y, accB.x, accB.y));
            else
                __stcs(reinterpret_cast<float2*>(&P.accD[s]), make_float2(accA.x, accB.x));

            // lastIteration = it where productive (Solver.cpp:903-910); a static body is marked at the position
            // of the first productive joint
            const bool productive = productiveA || productiveB;
            const unsigned markPos = productiveA ? pos : pos + 1;
            if (!st1)
            {
                w1.w = __int_as_float(productive ? it : last1);
                __stcg(&rows[b1], w1);
            }
            else if (productive)
                wake |= static_mark(&statics[b1], it, markPos, nullptr);
            if (!st2)
            {
                w2.w = __int_as_float(productive ? it : last2);
                __stcg(&rows[b2], w2);
            }
            else if (productive)
                wake |= static_mark(&statics[b2], it, markPos, nullptr);
            if (st1 || st2) __stcg(&P.processed[s], tick);
            anyProductive |= productive;
        }
        if (firstPass) bitCursor += U;
    }
    return anyProductive;
}

template <int PHASE, int U>
__device__ __forceinline__ int run_phase_pairs2(const SolveParams& P, const Level* levels, float4* stage, int iters, int tid, int nthreads, unsigned& epoch,
    int& tick, int& wakePasses, unsigned& activeCount)
{
    int2 pre[U];
    bool havePre = false;
    int ran = 0;
    unsigned long long predicted = ~0ull;   // iteration 0 relaxes every joint (lastIteration = -1 > -2)
    for (int it = 0; it < iters; ++it)
    {
        bool any = false, productiveAnywhere = false;
        unsigned long long activity = 0ull;
        unsigned bitCursor = 0u;
        for (int l = 0; l < P.numLevels; ++l)
        {
            const Level L = levels[l];
            ++tick;
            bool wake = false;
            any |= solve_pairs2<PHASE, U>(P, L, it, tick, true, tid, nthreads, wake, activeCount, pre, havePre, bitCursor, activity, stage);
            unsigned long long ticket;
            grid_arrive(P.barrier, epoch, wake, any, ticket);
            // while the grid drains into the barrier: index words of this thread's first batch of the level that follows
            {
                const bool wrap = l + 1 == P.numLevels;
                const Level N = levels[wrap ? 0 : l + 1];
                const int pairBaseN = N.start >> 1, numPairsN = (N.end - N.start) >> 1;
#pragma unroll
                for (int u = 0; u < U; ++u)
                {
                    const long long pN = pair_of(u);
                    pre[u] = make_int2(-1, -1);
                    if (pN < numPairsN)
                    {
                        const int p = pairBaseN + int(pN);
                        pre[u] = __ldcs(&P.pairIdx[p]);
                    }
                }
                havePre = true;
            }
            BarrierResult r = grid_wait(P.barrier, epoch, ticket);
            while (r.wake)
            {
                int2 none[U];
                unsigned dummyCursor = 0u;
                unsigned long long dummyActivity = 0ull;
                wake = false;
                any |= solve_pairs2<PHASE, U>(P, L, it, tick, false, tid, nthreads, wake, activeCount, none, false, dummyCursor, dummyActivity, stage);
                ++wakePasses;
                r = grid_barrier(P.barrier, epoch, wake, any);
            }
            productiveAnywhere = r.productive;
        }
        ++ran;
        predicted = activity;
        if (!productiveAnywhere) break;   // Solver.cpp:189 / :210
    }
    return ran;
}

template <int THREADS, int MIN_BLOCKS, int U>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS) k_solve_pairs2(SolveParams P)
{
    // the level table is read once per level pass, right behind the grid barrier: keep it in shared memory
    constexpr int kSmemLevels = 128;
    __shared__ Level s_levels[kSmemLevels];
    const bool tableFits = P.numLevels <= kSmemLevels;
    if (tableFits)
        for (int l = threadIdx.x; l < P.numLevels; l += THREADS) s_levels[l] = P.levels[l];
    __syncthreads();
    const Level* levels = tableFits ? s_levels : P.levels;
    extern __shared__ float4 s_stage[];   // U * kPairRecordWords * THREADS float4 (dynamic): staging of the active records
    float4* stage = s_stage + threadIdx.x;

    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    unsigned epoch = 0;
    int wakePasses = 0, tick = 0;
    unsigned active[2] = { 0u, 0u };

    for (int l = 0; l < P.numLevels; ++l)
    {
        const Level L = levels[l];
        const int pairBase = L.start >> 1, numPairs = (L.end - L.start) >> 1;
        for (long long p = tid; p < numPairs; p += nthreads) prestep_pair2(P, pairBase + int(p));
        grid_barrier(P.barrier, epoch, false, false);
    }

    const int ranImpulse = run_phase_pairs2<0, U>(P, levels, stage, P.contactIters, tid, nthreads, epoch, tick, wakePasses, active[0]);
    const int ranDisplacement = run_phase_pairs2<1, U>(P, levels, stage, P.penetrationIters, tid, nthreads, epoch, tick, wakePasses, active[1]);

    for (int phase = 0; phase < 2; ++phase)
    {
        unsigned v = active[phase];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&P.activeTotal[phase], static_cast<unsigned long long>(v));
    }
    if (tid == 0)
    {
        P.result[0] = ranImpulse;
        P.result[1] = ranDisplacement;
        P.result[2] = wakePasses;
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
// Iteration kernels (stats->kernelForm): 0 = k_solve (joint units: host-array API and replay schedules), 1 = k_solve_pairs
// (manifold units, streaming), 2 = k_solve_pairs2 (manifold units, record form), 3 = k_solve_strips (strips.cu; the default
// of the resident pipeline whenever the strip layout is usable).  1 and 2 give the same results bit for bit on the same
// (colour-major) schedule; phyx_b200_solve_tuning forces a form.
int solve_run(phyx_b200_ctx* c, const phyx_b200_solve_config* cfg, phyx_b200_solve_stats* stats)
{
    const int ns = c->slotCount, nl = c->levelCount, nb = c->bodyCount;
    const int I = cfg->contactIterationsCount, D = cfg->penetrationIterationsCount;
    if (I < 0 || D < 0 || I > 60000 || D > 60000)
    {
        set_error("solve: iteration counts must be in [0, 60000]");
        return PHYX_B200_ERR_ARGUMENT;
    }
    const bool strips = c->strip.valid;
    const bool paired = strips || (!c->hostLevels.empty() && c->hostLevels[0].grouped_end < 0);
    // streaming against record form: 1.65 against 1.78 ms with 38 % of the joint-iterations active, 1.64 against 1.50 ms with
    // 27 % (1 M pyramid, round 1): the choice follows the activity of the previous solve of this world
    const bool records = strips || (paired && (c->forceKernelForm ? c->forceKernelForm == 2 : c->lastActiveFraction < 0.32f));
    size_t ns1 = size_t(ns > 0 ? ns : 1);
    if (!records)
    {
        PHYX_TRY(c->q0.reserve(ns1 * sizeof(float4)));
        PHYX_TRY(c->q1.reserve(ns1 * sizeof(float4)));
        PHYX_TRY(c->q2.reserve(ns1 * sizeof(float4)));
        PHYX_TRY(c->q3.reserve(ns1 * sizeof(float4)));
    }
    else
    {
        PHYX_TRY(c->pairQ.reserve(ns1 / 2 * kPairRecordWords * sizeof(float4) + 128));
        PHYX_TRY(c->pairIdx.reserve(ns1 / 2 * sizeof(int2) + 64));
    }
    PHYX_TRY(c->accNF.reserve(ns1 * sizeof(float2) + 64));
    PHYX_TRY(c->accD.reserve(ns1 * sizeof(float) + 64));
    PHYX_TRY(c->stamps.reserve(size_t(nb > 0 ? nb : 1) * 2 * sizeof(unsigned long long)));
    PHYX_TRY(c->processed.reserve(ns1 * sizeof(int)));
    PHYX_TRY(c->solveFlags.reserve(128));

    cudaEvent_t e0 = c->ev[0], e1 = c->ev[1], e2 = c->ev[2], e3 = c->ev[3];
    PHYX_CUDA(record_event(c, e0));
    int ranI = I > 0 ? 1 : 0, ranD = D > 0 ? 1 : 0, wakePasses = 0;
    if (ns > 0 && nl > 0)
    {
        if (!strips) PHYX_CUDA(cudaMemsetAsync(c->stamps.ptr, 0, size_t(nb) * 2 * sizeof(unsigned long long), c->stream));
        PHYX_CUDA(cudaMemsetAsync(c->solveFlags.ptr, 0, 128, c->stream));
        if (!strips) PHYX_CUDA(cudaMemsetAsync(c->processed.ptr, 0, size_t(ns) * sizeof(int), c->stream));
        // memory order of the solver rows: the broadphase's sorted-x order when there is one (measured best: k_solve 2.40 ms
        // vs 3.08 ms for body order on the 1 M pyramid), else body order
        const bool sorted = c->rowOrderValid && c->rowOrderBodies == nb;
        const unsigned* order = sorted ? c->entryIndex.as<unsigned>() : nullptr;
        const int* rowOf = sorted ? c->rowOf.as<int>() : nullptr;
        PHYX_TRY(c->solveRows.reserve(size_t(nb) * 2 * sizeof(float4)));
        float4* rowsVel = c->solveRows.as<float4>();
        float4* rowsDisp = rowsVel + nb;
        const bool dual = c->strictLevelCount > 0 && c->slotPosValid;
        if (dual) PHYX_TRY(c->rowsMulti.reserve(size_t(nb)));
        // (deferred step: ns is a bound, the true slot count is StepCtl::slots; a stopped step has zero counts everywhere)
        const Count nbc = c->count(nb, &StepCtl::bodies), nsc = c->count(ns, &StepCtl::slots);
        k_prepare_bodies<<<(nb + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(nbc, order, c->vel.as<float4>(), c->disp.as<float4>(), rowsVel, rowsDisp,
            dual ? c->staticMulti.as<unsigned char>() : nullptr, dual ? c->rowsMulti.as<unsigned char>() : nullptr);
        c->launches++;
        int grid = (ns + kBlock - 1) / kBlock;
        c->lastKernelForm = strips ? 3 : records ? 2 : paired ? 1 : 0;
        // the strip layout writes its own index words (rows local to a strip's shared memory)
        k_refresh<<<grid, kBlock, 0, c->stream>>>(nsc, c->slotJoint.as<int>(), c->joints.as<phyx_contact_joint>(), c->contactPoints.as<float4>(),
            c->params.as<float4>(), rowOf, c->q0.as<float4>(), c->q1.as<float4>(), c->q2.as<float4>(), c->q3.as<float4>(), c->accNF.as<float2>(),
            c->accD.as<float>(), records ? c->pairQ.as<float4>() : nullptr, records ? c->pairIdx.as<int2>() : nullptr, 0, !strips);
        c->launches++;
        PHYX_CUDA(record_event(c, e1));

        SolveParams P;
        memset(&P, 0, sizeof(P));
        P.result = reinterpret_cast<int*>(c->solveFlags.as<char>() + 32);
        if (strips)
            PHYX_TRY(strip_solve_launch(c, cfg, rowsVel, rowsDisp));
        else
        {
            P.vel = rowsVel;
            P.disp = rowsDisp;
            P.q0 = c->q0.as<float4>();
            P.q1 = c->q1.as<float4>();
            P.q2 = c->q2.as<float4>();
            P.q3 = c->q3.as<float4>();
            P.accNF = c->accNF.as<float2>();
            P.accD = c->accD.as<float>();
            P.levels = c->levels.as<Level>();
            P.numLevels = nl;
            P.slotPos = c->slotPosValid ? c->slotPos.as<int>() : nullptr;
            P.processed = c->processed.as<int>();
            P.staticImp = c->stamps.as<unsigned long long>();
            P.staticDisp = P.staticImp + nb;
            P.contactIters = I;
            P.penetrationIters = D;
            P.barrier = c->solveFlags.as<unsigned long long>();          // 4 words
            P.activeTotal = reinterpret_cast<unsigned long long*>(c->solveFlags.as<char>() + 48);
            P.pairQ = records ? c->pairQ.as<float4>() : nullptr;
            P.pairIdx = records ? c->pairIdx.as<int2>() : nullptr;
            P.strictLevels = dual ? c->strictLevels.as<Level>() : nullptr;
            P.numStrictLevels = dual ? c->strictLevelCount : 0;
            P.strictMap = dual ? c->strictMap.as<int>() : nullptr;
            P.rowsMulti = dual ? c->rowsMulti.as<unsigned char>() : nullptr;
            P.numMultiStatics = dual ? c->numMultiStatics : 0;
            P.hotCount = dual ? reinterpret_cast<int*>(c->solveFlags.as<char>() + 64) : nullptr;
            // CTA shapes (measured, round 1): manifold units 256 threads x 2 CTAs per SM, joint units 512 x 2
            const int sblock = paired ? 256 : 512;
            void* solveKernel = records ? (void*)k_solve_pairs2<256, 2, kPairU> : paired ? (void*)k_solve_pairs<256, 2>
                                        : dual  ? (void*)k_solve<512, 2, true> : (void*)k_solve<512, 2, false>;
            // the record form stages the active records in dynamic shared memory: [U][6][threads] float4
            const size_t solveSmem = records ? size_t(kPairU) * kPairRecordWords * sblock * sizeof(float4) : 0;
            if (records) PHYX_CUDA(cudaFuncSetAttribute(solveKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(solveSmem)));
            int per = 0;
            PHYX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, solveKernel, sblock, solveSmem));
            if (per < 1)
            {
                set_error("solve kernel does not fit on an SM");
                return PHYX_B200_ERR_CUDA;
            }
            c->solveBlocksPerSM = per;
            // persistent grid: every SM full, but no more CTAs than the widest level can use
            int maxLevel = 0;
            for (const Level& L : c->hostLevels) maxLevel = max(maxLevel, L.end - L.start);
            if (paired) maxLevel = (maxLevel + 1) / 2;   // one thread per pair
            const int want = (maxLevel + sblock - 1) / sblock;
            const int sgrid = max(1, min(want, c->numSMs * c->solveBlocksPerSM));
            void* args[] = { &P };
            PHYX_CUDA(cudaLaunchCooperativeKernel(solveKernel, dim3(sgrid), dim3(sblock), args, solveSmem, c->stream));
            c->launches++;
        }
        PHYX_CUDA(record_event(c, e2));
        k_finish<<<grid, kBlock, 0, c->stream>>>(nsc, c->slotJoint.as<int>(), c->accNF.as<float2>(), c->joints.as<phyx_contact_joint>());
        PHYX_TRY(c->bodyActivity.reserve(size_t(nb) * sizeof(int)));
        k_finish_bodies<<<(nb + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(nbc, order, rowsVel, rowsDisp, c->vel.as<float4>(), c->disp.as<float4>(),
            c->bodyActivity.as<int>());
        c->activityValid = true;
        c->activityBodies = nb;
        c->launches += 2;
        PHYX_CUDA(record_event(c, e3));
        if (c->def.active) return PHYX_B200_OK;   // the results come home with the step's counts (deferred_finish, api.cu)
        int host[8];   // result[0..2], pad, activeTotal[2] as two 64-bit words
        PHYX_TRY(fetch_small(c, P.result, sizeof(host), host));
        PHYX_CUDA(cudaEventSynchronize(e3));   // (complete by now: the read-back was enqueued behind it)
        ranI = host[0];
        ranD = host[1];
        wakePasses = host[2];
        {
            long long act0 = 0;
            memcpy(&act0, &host[4], 8);
            const double nominal = double(c->jointCount) * double(ranI > 0 ? ranI : 1);
            c->lastActiveFraction = nominal > 0.0 ? float(double(act0) / nominal) : 1.0f;
        }
        if (stats)
        {
            memcpy(&stats->activeJointIterations[0], &host[4], 8);
            memcpy(&stats->activeJointIterations[1], &host[6], 8);
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
        stats->wakePasses = wakePasses;
        stats->colourRounds = c->colourRounds;
        stats->kernelForm = c->lastKernelForm;
    }
    return PHYX_B200_OK;
}


// =====================================================================================================
// One world over several devices: the partitioned solve (an island that spans devices)
// =====================================================================================================
// colour.cu (part_layout) lays the coloured manifolds out class-major: rank q's INTERIOR manifolds (all
// dynamic bodies in q's row range), then the CUT manifolds (bodies of two ranks).  Every pass of the solve
// (warm start, each impulse iteration, each displacement iteration) then runs as
//     interior levels, every rank its own, in parallel              (kernel, mode 0)
//     exchange of the boundary rows (rows that cut manifolds touch)  (fused into the two kernels, see below)
//     cut levels, every rank all of them, redundantly                (kernel, mode 1)
// which equals ONE sequential sweep over the slot order [rank 0 levels .. rank R-1 levels, cut levels]: the
// interior classes touch disjoint dynamic rows, and when the cut levels start every rank holds the same
// fresh boundary rows, so all ranks compute identical cut results and no second exchange is needed.  This is
// the ghost-body exchange at iteration boundaries that BASELINE.json's north_star describes, done without a
// collective call: the tail of the interior kernel STORES the rank's own boundary rows straight into every
// peer's receive buffer over NVLink (peer pointers from CUDA IPC), fences, and releases a per-sender
// sequence flag in the peer's memory; the head of the cut kernel acquires the flags of all peers and copies
// the received rows into its row array.  Two receive slots alternate, which is enough because a rank can
// be at most one exchange ahead of a peer that has not consumed the previous one.
// The productive early-out (Solver.cpp:189,210) needs an OR over all ranks: the interior flag travels with
// the boundary rows, the cut flag is computed identically everywhere, and the decision is left in device
// memory (stop[phase]) where the kernels of the remaining iterations see it and return at once, so the
// host never waits for it.  Static bodies (never written) keep their lastIteration words per rank: a rank
// sees the productive ground contacts of its own manifolds only (documented deviation from the one-device
// order; the oracle reproduces it by giving every rank its own copy of each static body).

struct PartDevState
{
    unsigned long long xseq;       // boundary exchanges completed (persists across solves; identical on all ranks)
    int stop[2];                   // phase ended early: no productive joint anywhere
    int ran[2];                    // iterations run per phase
    int interiorProductive;        // this rank's last interior launch saw a productive joint
    int error;                     // 1: gave up waiting for a peer
    unsigned long long active[2];  // joint-iterations relaxed by this rank (interior and cut)
    int wakePasses, pad_;
};

constexpr size_t kPartFlagBytes = 1024;      // exchange buffer: [0,64) boundary flags, [64,128) bulk flags, then the areas
constexpr unsigned long long kPartTimeoutNs = 4000000000ull;

struct PartArgs
{
    int phase, it, levelBegin, levelEnd, tickBase, mode, ring, rank, ranks, bCap;
    int bStart[kMaxRanks + 1];
    const int* bRows;
    PartDevState* st;
    char* self;
    char* peer[kMaxRanks];
};

__host__ __device__ __forceinline__ size_t part_boundary_bytes(int ranks, int bCap) { return size_t(2) * ranks * (size_t(bCap) + 1) * sizeof(float4); }

__device__ __forceinline__ float4* part_recv(char* xbuf, int ranks, int bCap, int slot, int sender)
{
    return reinterpret_cast<float4*>(xbuf + kPartFlagBytes) + (size_t(slot) * ranks + sender) * (size_t(bCap) + 1);
}

__device__ __forceinline__ unsigned long long global_timer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// spin until *flag >= want (written by a peer device with st.release.sys); false after kPartTimeoutNs
__device__ __forceinline__ bool part_wait(const unsigned long long* flag, unsigned long long want)
{
    const unsigned long long t0 = global_timer();
    bool ok = true;
    for (;;)
    {
        unsigned long long v;
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
        if (v >= want) break;
        if (global_timer() - t0 > kPartTimeoutNs)
        {
            ok = false;
            break;
        }
    }
    asm volatile("fence.acq_rel.sys;" ::: "memory");
    return ok;
}

__device__ __forceinline__ void part_signal(unsigned long long* flag, unsigned long long value)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(value) : "memory");
}

template <int THREADS, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS) k_part_solve(SolveParams P, PartArgs A)
{
    extern __shared__ float4 s_stage[];   // kPairU * kPairRecordWords * THREADS float4 (dynamic)
    float4* stage = s_stage + threadIdx.x;
    PartDevState* st = A.st;
    // stop[] is only written by the last thread standing of an earlier cut launch: the same answer for every thread
    if (A.phase >= 0 && __ldcg(&st->stop[A.phase])) return;

    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    unsigned long long* ring = P.barrier + 4 * A.ring;
    if (tid == 0)   // the next launch uses the other ring
    {
        unsigned long long* other = P.barrier + 4 * (A.ring ^ 1);
        other[0] = other[1] = other[2] = other[3] = 0ull;
    }
    unsigned epoch = 0;
    const unsigned long long xseq = __ldcg(&st->xseq);
    const int slot = int(xseq & 1ull);
    float4* rows = A.phase == 1 ? P.disp : P.vel;
    unsigned long long* selfFlags = reinterpret_cast<unsigned long long*>(A.self);
    bool gathered = false;

    if (A.mode == 1)
    {
        // head of the cut launch: take over the peers' boundary rows of exchange `xseq`
        if (threadIdx.x == 0)
        {
            bool ok = true;
            for (int q = 0; q < A.ranks; ++q)
                if (q != A.rank) ok = part_wait(selfFlags + q, xseq + 1ull) && ok;
            if (!ok) atomicExch(&st->error, 1);
        }
        __syncthreads();
        for (int q = 0; q < A.ranks; ++q)
        {
            if (q == A.rank) continue;
            const float4* src = part_recv(A.self, A.ranks, A.bCap, slot, q);
            const int first = A.bStart[q], n = A.bStart[q + 1] - first;
            for (int i = tid; i < n; i += nthreads) __stcg(&rows[A.bRows[first + i]], __ldcg(src + i));
            gathered |= __ldcg(src + A.bCap).x != 0.0f;
        }
        gathered |= __ldcg(&st->interiorProductive) != 0;
        grid_barrier(ring, epoch, false, false);
    }

    bool any = false, productive = false;
    int wakePasses = 0;
    unsigned active = 0u;
    for (int l = A.levelBegin; l < A.levelEnd; ++l)
    {
        const Level L = P.levels[l];
        const int tick = A.tickBase + (l - A.levelBegin) + 1;
        if (A.phase < 0)
        {
            const int pairBase = L.start >> 1, numPairs = (L.end - L.start) >> 1;
            for (long long p = tid; p < numPairs; p += nthreads) prestep_pair2(P, pairBase + int(p));
            grid_barrier(ring, epoch, false, false);
            continue;
        }
        int2 none[kPairU];
        unsigned cursor = 0u;
        unsigned long long activity = 0ull;
        bool wake = false;
        if (A.phase == 0)
            any |= solve_pairs2<0, kPairU>(P, L, A.it, tick, true, tid, nthreads, wake, active, none, false, cursor, activity, stage);
        else
            any |= solve_pairs2<1, kPairU>(P, L, A.it, tick, true, tid, nthreads, wake, active, none, false, cursor, activity, stage);
        BarrierResult r = grid_barrier(ring, epoch, wake, any);
        while (r.wake)
        {
            wake = false;
            if (A.phase == 0)
                any |= solve_pairs2<0, kPairU>(P, L, A.it, tick, false, tid, nthreads, wake, active, none, false, cursor, activity, stage);
            else
                any |= solve_pairs2<1, kPairU>(P, L, A.it, tick, false, tid, nthreads, wake, active, none, false, cursor, activity, stage);
            ++wakePasses;
            r = grid_barrier(ring, epoch, wake, any);
        }
        productive = r.productive;
    }

    if (A.mode == 0)
    {
        // tail of the interior launch: this rank's own boundary rows (final: the last level barrier is behind
        // us) go straight into every peer's receive slot, then the flag
        const int first = A.bStart[A.rank], n = A.bStart[A.rank + 1] - first;
        for (int q = 0; q < A.ranks; ++q)
        {
            if (q == A.rank) continue;
            float4* dst = part_recv(A.peer[q], A.ranks, A.bCap, slot, A.rank);
            for (int i = tid; i < n; i += nthreads) dst[i] = __ldcg(&rows[A.bRows[first + i]]);
            if (tid == 0) dst[A.bCap] = make_float4(productive ? 1.0f : 0.0f, 0.f, 0.f, 0.f);
        }
        __threadfence_system();
        grid_barrier(ring, epoch, false, false);
        if (tid == 0)
        {
            st->interiorProductive = productive ? 1 : 0;
            for (int q = 0; q < A.ranks; ++q)
                if (q != A.rank) part_signal(reinterpret_cast<unsigned long long*>(A.peer[q]) + A.rank, xseq + 1ull);
        }
    }
    else if (tid == 0)
    {
        if (A.phase >= 0)
        {
            st->ran[A.phase] = A.it + 1;
            if (!(gathered || productive)) st->stop[A.phase] = 1;
        }
        st->xseq = xseq + 1ull;
    }

    if (A.phase >= 0)
    {
        unsigned v = active;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&st->active[A.phase], static_cast<unsigned long long>(v));
        if (tid == 0 && wakePasses) atomicAdd(&st->wakePasses, wakePasses);
    }
}

// end-of-solve exchange: flags around plain peer copies
__global__ void k_part_bulk_signal(PartArgs A, unsigned long long seq)
{
    __threadfence_system();
    for (int q = 0; q < A.ranks; ++q)
        if (q != A.rank) part_signal(reinterpret_cast<unsigned long long*>(A.peer[q]) + 8 + A.rank, seq);
}

__global__ void k_part_bulk_wait(PartArgs A, unsigned long long seq)
{
    unsigned long long* flags = reinterpret_cast<unsigned long long*>(A.self) + 8;
    bool ok = true;
    for (int q = 0; q < A.ranks; ++q)
        if (q != A.rank) ok = part_wait(flags + q, seq) && ok;
    if (!ok) atomicExch(&A.st->error, 1);
}

static PartArgs part_args(phyx_b200_ctx* c)
{
    Partition& pt = c->part;
    PartArgs A;
    memset(&A, 0, sizeof(A));
    A.rank = pt.rank;
    A.ranks = pt.ranks;
    A.bCap = pt.boundaryCapacity;
    for (int q = 0; q <= pt.ranks; ++q) A.bStart[q] = pt.bStart[q];
    A.bRows = pt.bRows.as<int>();
    A.st = pt.state.as<PartDevState>();
    A.self = pt.xbuf;
    for (int q = 0; q < pt.ranks; ++q) A.peer[q] = pt.peer[q];
    return A;
}

static size_t part_bulk_offset(const Partition& pt) { return (kPartFlagBytes + part_boundary_bytes(pt.ranks, pt.boundaryCapacity) + 255) & ~size_t(255); }

int part_create(phyx_b200_ctx* c, int rank, int ranks, int boundaryCapacity, size_t bulkBytes, void* ipcHandleOut, void** localOut)
{
    Partition& pt = c->part;
    if (ranks < 2 || ranks > kMaxRanks || rank < 0 || rank >= ranks || boundaryCapacity < 1)
    {
        set_error("partition_create: need 2 <= ranks <= %d, 0 <= rank < ranks, boundaryCapacity >= 1", kMaxRanks);
        return PHYX_B200_ERR_ARGUMENT;
    }
    if (pt.xbuf)
    {
        set_error("partition_create: context is already partitioned");
        return PHYX_B200_ERR_STATE;
    }
    pt.rank = rank;
    pt.ranks = ranks;
    pt.boundaryCapacity = boundaryCapacity;
    pt.bulkCapacity = (bulkBytes + 255) & ~size_t(255);
    pt.xbytes = part_bulk_offset(pt) + size_t(ranks) * pt.bulkCapacity;
    PHYX_CUDA(cudaMalloc(reinterpret_cast<void**>(&pt.xbuf), pt.xbytes));
    PHYX_CUDA(cudaMemset(pt.xbuf, 0, kPartFlagBytes));
    PHYX_TRY(pt.state.reserve(sizeof(PartDevState)));
    PHYX_CUDA(cudaMemset(pt.state.ptr, 0, sizeof(PartDevState)));
    if (!pt.evReady) PHYX_CUDA(cudaEventCreateWithFlags(&pt.evReady, cudaEventDisableTiming));
    pt.bulkSeq = 0;
    pt.planValid = false;
    pt.failed = false;
    for (int q = 0; q < kMaxRanks; ++q)
    {
        pt.peer[q] = nullptr;
        pt.peerOpened[q] = false;
    }
    pt.peer[rank] = pt.xbuf;
    if (ipcHandleOut)
    {
        cudaIpcMemHandle_t h;
        PHYX_CUDA(cudaIpcGetMemHandle(&h, pt.xbuf));
        static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
        memcpy(ipcHandleOut, &h, sizeof(h));
    }
    if (localOut) *localOut = pt.xbuf;
    // a schedule built before this call has the one-device layout
    c->scheduleMode = -1;
    c->colourStateValid = false;
    return PHYX_B200_OK;
}

int part_attach(phyx_b200_ctx* c, const void* ipcHandles, void* const* localPointers, const int* peerDevices)
{
    Partition& pt = c->part;
    if (!pt.xbuf || (!ipcHandles && !localPointers))
    {
        set_error("partition_attach: call partition_create first and pass the peers' IPC handles or pointers");
        return PHYX_B200_ERR_STATE;
    }
    for (int q = 0; q < pt.ranks; ++q)
    {
        if (q == pt.rank) continue;
        if (localPointers)
        {
            pt.peer[q] = static_cast<char*>(localPointers[q]);
            if (peerDevices && peerDevices[q] != c->device)
            {
                cudaError_t e = cudaDeviceEnablePeerAccess(peerDevices[q], 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled)
                    cudaGetLastError();
                else if (e != cudaSuccess)
                {
                    set_error("cudaDeviceEnablePeerAccess(%d) from device %d failed: %s", peerDevices[q], c->device, cudaGetErrorString(e));
                    return PHYX_B200_ERR_CUDA;
                }
            }
        }
        else
        {
            cudaIpcMemHandle_t h;
            memcpy(&h, static_cast<const char*>(ipcHandles) + size_t(q) * sizeof(h), sizeof(h));
            void* p = nullptr;
            PHYX_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            pt.peer[q] = static_cast<char*>(p);
            pt.peerOpened[q] = true;
        }
        if (!pt.peer[q])
        {
            set_error("partition_attach: no buffer for rank %d", q);
            return PHYX_B200_ERR_ARGUMENT;
        }
    }
    return PHYX_B200_OK;
}

void part_destroy(phyx_b200_ctx* c)
{
    Partition& pt = c->part;
    for (int q = 0; q < kMaxRanks; ++q)
    {
        if (pt.peerOpened[q] && pt.peer[q]) cudaIpcCloseMemHandle(pt.peer[q]);
        pt.peer[q] = nullptr;
        pt.peerOpened[q] = false;
    }
    if (pt.xbuf) cudaFree(pt.xbuf);
    pt.xbuf = nullptr;
    if (pt.evReady) cudaEventDestroy(pt.evReady);
    pt.evReady = nullptr;
    delete static_cast<SolveParams*>(pt.params);
    pt.params = nullptr;
    DevBuf* bufs[] = { &pt.rowFlag, &pt.rowPrefix, &pt.bRows, &pt.planWords, &pt.binLevels, &pt.partLevels, &pt.state };
    for (DevBuf* b : bufs) b->release();
    pt.ranks = 1;
    pt.rank = 0;
    pt.planValid = false;
    c->scheduleMode = -1;
    c->colourStateValid = false;
}

constexpr int kPartThreads = 256;
constexpr size_t kPartSmem = size_t(kPairU) * kPairRecordWords * kPartThreads * sizeof(float4);

// schedule is built (part_layout): pack the rows, refresh this rank's joints, reset the per-solve state
int part_begin(phyx_b200_ctx* c, const phyx_b200_solve_config* cfg)
{
    Partition& pt = c->part;
    const int ns = c->slotCount, nb = c->bodyCount, R = pt.ranks;
    if (!pt.xbuf)
    {
        set_error("solve_partitioned: partition_create / partition_attach first");
        return PHYX_B200_ERR_STATE;
    }
    for (int q = 0; q < R; ++q)
        if (!pt.peer[q])
        {
            set_error("solve_partitioned: rank %d is not attached", q);
            return PHYX_B200_ERR_STATE;
        }
    if (cfg->contactIterationsCount < 0 || cfg->penetrationIterationsCount < 0 || cfg->contactIterationsCount > 60000 || cfg->penetrationIterationsCount > 60000)
    {
        set_error("solve: iteration counts must be in [0, 60000]");
        return PHYX_B200_ERR_ARGUMENT;
    }
    pt.launchIndex = 0;
    if (pt.failed)
    {
        set_error("solve_partitioned: an earlier solve timed out waiting for a peer; partition_destroy / partition_create / partition_attach on all ranks first");
        return PHYX_B200_ERR_STATE;
    }
    if (c->jointCount == 0) return PHYX_B200_OK;
    if (!pt.planValid)
    {
        set_error("solve_partitioned: needs the manifold-unit schedule of the resident pipeline (RefreshContactJoints before SolveJoints)");
        return PHYX_B200_ERR_STATE;
    }
    for (int q = 0; q < R; ++q)
    {
        const size_t need = size_t(pt.cuts[q + 1] - pt.cuts[q]) * 2 * sizeof(float4) + size_t(pt.classSlotStart[q + 1] - pt.classSlotStart[q]) * sizeof(float2);
        if (pt.bStart[q + 1] - pt.bStart[q] > pt.boundaryCapacity || need > pt.bulkCapacity)
        {
            set_error("solve_partitioned: rank %d has %d boundary rows / %zu bulk bytes, capacity is %d / %zu", q, pt.bStart[q + 1] - pt.bStart[q], need,
                pt.boundaryCapacity, pt.bulkCapacity);
            return PHYX_B200_ERR_CAPACITY;
        }
    }
    const size_t ns1 = size_t(ns > 0 ? ns : 1);
    PHYX_TRY(c->pairQ.reserve(ns1 / 2 * kPairRecordWords * sizeof(float4) + 128));
    PHYX_TRY(c->pairIdx.reserve(ns1 / 2 * sizeof(int2) + 64));
    PHYX_TRY(c->accNF.reserve(ns1 * sizeof(float2)));
    PHYX_TRY(c->accD.reserve(ns1 * sizeof(float)));
    PHYX_TRY(c->stamps.reserve(size_t(nb > 0 ? nb : 1) * 2 * sizeof(unsigned long long)));
    PHYX_TRY(c->processed.reserve(ns1 * sizeof(int)));
    PHYX_TRY(c->solveFlags.reserve(128));
    PHYX_TRY(c->solveRows.reserve(size_t(nb) * 2 * sizeof(float4)));
    PHYX_CUDA(cudaMemsetAsync(c->stamps.ptr, 0, size_t(nb) * 2 * sizeof(unsigned long long), c->stream));
    PHYX_CUDA(cudaMemsetAsync(c->solveFlags.ptr, 0, 128, c->stream));
    PHYX_CUDA(cudaMemsetAsync(c->processed.ptr, 0, size_t(ns) * sizeof(int), c->stream));
    PHYX_CUDA(cudaMemsetAsync(pt.state.as<char>() + 8, 0, sizeof(PartDevState) - 8, c->stream));   // everything but xseq

    // rows in the broadphase's sorted-x order when there is one (the layout was built on the same choice)
    const bool sorted = c->rowOrderValid && c->rowOrderBodies == nb;
    const unsigned* order = sorted ? c->entryIndex.as<unsigned>() : nullptr;
    const int* rowOf = sorted ? c->rowOf.as<int>() : nullptr;
    float4* rowsVel = c->solveRows.as<float4>();
    float4* rowsDisp = rowsVel + nb;
    k_prepare_bodies<<<(nb + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(nb, order, c->vel.as<float4>(), c->disp.as<float4>(), rowsVel, rowsDisp, nullptr, nullptr);
    c->launches++;
    // RefreshJoints for the slots this rank relaxes: its own class and the cut class
    const int ranges[2][2] = { { pt.classSlotStart[pt.rank], pt.classSlotStart[pt.rank + 1] }, { pt.classSlotStart[R], pt.classSlotStart[R + 1] } };
    for (const auto& rg : ranges)
    {
        const int n = rg[1] - rg[0];
        if (n <= 0) continue;
        k_refresh<<<(n + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(rg[1], c->slotJoint.as<int>(), c->joints.as<phyx_contact_joint>(), c->contactPoints.as<float4>(),
            c->params.as<float4>(), rowOf, nullptr, nullptr, nullptr, nullptr, c->accNF.as<float2>(), c->accD.as<float>(), c->pairQ.as<float4>(),
            c->pairIdx.as<int2>(), rg[0], true);
        c->launches++;
    }
    if (!pt.params) pt.params = new SolveParams;
    SolveParams& P = *static_cast<SolveParams*>(pt.params);
    memset(&P, 0, sizeof(P));
    P.vel = rowsVel;
    P.disp = rowsDisp;
    P.accNF = c->accNF.as<float2>();
    P.accD = c->accD.as<float>();
    P.levels = pt.partLevels.as<Level>();
    P.numLevels = pt.numInterior + pt.numCut;
    P.processed = c->processed.as<int>();
    P.staticImp = c->stamps.as<unsigned long long>();
    P.staticDisp = P.staticImp + nb;
    P.contactIters = cfg->contactIterationsCount;
    P.penetrationIters = cfg->penetrationIterationsCount;
    P.barrier = c->solveFlags.as<unsigned long long>();   // two rings of 4 words
    P.pairQ = c->pairQ.as<float4>();
    P.pairIdx = c->pairIdx.as<int2>();
    {
        int per = 0;
        PHYX_CUDA(cudaFuncSetAttribute(k_part_solve<kPartThreads, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kPartSmem)));
        PHYX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_part_solve<kPartThreads, 2>, kPartThreads, kPartSmem));
        if (per < 1)
        {
            set_error("partitioned solve kernel does not fit on an SM");
            return PHYX_B200_ERR_CUDA;
        }
        c->solveBlocksPerSM = per;
    }
    PHYX_CUDA(cudaGetLastError());
    return PHYX_B200_OK;
}

// one launch: phase -1 (warm start), 0 (impulse iteration `it`), 1 (displacement iteration `it`); mode 0 interior, 1 cut
int part_launch(phyx_b200_ctx* c, int phase, int it, int mode)
{
    Partition& pt = c->part;
    if (c->jointCount == 0) return PHYX_B200_OK;
    PartArgs A = part_args(c);
    A.phase = phase;
    A.it = it;
    A.mode = mode;
    A.levelBegin = mode == 0 ? 0 : pt.numInterior;
    A.levelEnd = mode == 0 ? pt.numInterior : pt.numInterior + pt.numCut;
    A.ring = pt.launchIndex & 1;
    A.tickBase = (pt.launchIndex + 1) * 128;
    pt.launchIndex++;
    const int widest = (mode == 0 ? pt.widestInterior : pt.widestCut) / 2;   // manifolds of the widest level
    const int want = (widest + kPartThreads - 1) / kPartThreads;
    const int grid = std::max(1, std::min(want, c->numSMs * c->solveBlocksPerSM));
    SolveParams& P = *static_cast<SolveParams*>(pt.params);
    void* args[] = { &P, &A };
    PHYX_CUDA(cudaLaunchCooperativeKernel((void*)k_part_solve<kPartThreads, 2>, dim3(grid), dim3(kPartThreads), args, kPartSmem, c->stream));
    c->launches++;
    return PHYX_B200_OK;
}

// end of the solve: every rank hands its own rows (both arrays) and the accumulated impulses of its own slots
// to all peers, so that the replicas are identical again for FinishJoints / FinishBodies and the next stages
int part_bulk_push(phyx_b200_ctx* c)
{
    Partition& pt = c->part;
    if (c->jointCount == 0) return PHYX_B200_OK;
    const int nb = c->bodyCount, r = pt.rank;
    const size_t rowsN = size_t(pt.cuts[r + 1] - pt.cuts[r]), slotsN = size_t(pt.classSlotStart[r + 1] - pt.classSlotStart[r]);
    const float4* rowsVel = c->solveRows.as<float4>();
    const float4* rowsDisp = rowsVel + nb;
    const size_t base = part_bulk_offset(pt) + size_t(r) * pt.bulkCapacity;
    for (int q = 0; q < pt.ranks; ++q)
    {
        if (q == r) continue;
        char* dst = pt.peer[q] + base;
        if (rowsN)
        {
            PHYX_CUDA(cudaMemcpyAsync(dst, rowsVel + pt.cuts[r], rowsN * sizeof(float4), cudaMemcpyDefault, c->stream));
            PHYX_CUDA(cudaMemcpyAsync(dst + rowsN * sizeof(float4), rowsDisp + pt.cuts[r], rowsN * sizeof(float4), cudaMemcpyDefault, c->stream));
        }
        if (slotsN)
            PHYX_CUDA(cudaMemcpyAsync(dst + 2 * rowsN * sizeof(float4), c->accNF.as<float2>() + pt.classSlotStart[r], slotsN * sizeof(float2), cudaMemcpyDefault, c->stream));
    }
    pt.bulkSeq++;
    k_part_bulk_signal<<<1, 1, 0, c->stream>>>(part_args(c), pt.bulkSeq);
    c->launches++;
    PHYX_CUDA(cudaGetLastError());
    return PHYX_B200_OK;
}

int part_bulk_pull(phyx_b200_ctx* c)
{
    Partition& pt = c->part;
    if (c->jointCount == 0) return PHYX_B200_OK;
    const int nb = c->bodyCount;
    k_part_bulk_wait<<<1, 1, 0, c->stream>>>(part_args(c), pt.bulkSeq);
    c->launches++;
    float4* rowsVel = c->solveRows.as<float4>();
    float4* rowsDisp = rowsVel + nb;
    for (int q = 0; q < pt.ranks; ++q)
    {
        if (q == pt.rank) continue;
        const size_t rowsN = size_t(pt.cuts[q + 1] - pt.cuts[q]), slotsN = size_t(pt.classSlotStart[q + 1] - pt.classSlotStart[q]);
        const char* src = pt.xbuf + part_bulk_offset(pt) + size_t(q) * pt.bulkCapacity;
        if (rowsN)
        {
            PHYX_CUDA(cudaMemcpyAsync(rowsVel + pt.cuts[q], src, rowsN * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
            PHYX_CUDA(cudaMemcpyAsync(rowsDisp + pt.cuts[q], src + rowsN * sizeof(float4), rowsN * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
        }
        if (slotsN)
            PHYX_CUDA(cudaMemcpyAsync(c->accNF.as<float2>() + pt.classSlotStart[q], src + 2 * rowsN * sizeof(float4), slotsN * sizeof(float2), cudaMemcpyDeviceToDevice, c->stream));
    }
    return PHYX_B200_OK;
}

// FinishJoints / FinishBodies on the (again identical) replicas; results of this rank
int part_end(phyx_b200_ctx* c, phyx_b200_solve_stats* stats)
{
    Partition& pt = c->part;
    const int ns = c->slotCount, nb = c->bodyCount;
    PartDevState host;
    memset(&host, 0, sizeof(host));
    if (c->jointCount > 0)
    {
        const bool sorted = c->rowOrderValid && c->rowOrderBodies == nb;
        const unsigned* order = sorted ? c->entryIndex.as<unsigned>() : nullptr;
        float4* rowsVel = c->solveRows.as<float4>();
        k_finish<<<(ns + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(ns, c->slotJoint.as<int>(), c->accNF.as<float2>(), c->joints.as<phyx_contact_joint>());
        k_finish_bodies<<<(nb + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(nb, order, rowsVel, rowsVel + nb, c->vel.as<float4>(), c->disp.as<float4>(), nullptr);
        c->launches += 2;
        PHYX_CUDA(cudaMemcpyAsync(&host, pt.state.ptr, sizeof(host), cudaMemcpyDeviceToHost, c->stream));
        PHYX_CUDA(cudaStreamSynchronize(c->stream));
        if (host.error)
        {
            // the ranks' exchange counters and flags may have diverged for good: refuse further partitioned solves on this
            // context until partition_destroy + partition_create + partition_attach have run again on ALL ranks
            pt.failed = true;
            pt.planValid = false;
            set_error("solve_partitioned: rank %d gave up waiting for a peer's boundary rows (a peer failed or did not call the solve); "
                      "re-create the partition on all ranks", pt.rank);
            return PHYX_B200_ERR_STATE;
        }
    }
    if (stats)
    {
        stats->joints = c->jointCount;
        stats->slots = ns;
        stats->levels = c->levelCount;
        stats->contactIterationsRun = c->jointCount ? host.ran[0] : 0;
        stats->penetrationIterationsRun = c->jointCount ? host.ran[1] : 0;
        stats->wakePasses = host.wakePasses;
        stats->colourRounds = c->colourRounds;
        stats->kernelForm = 2;
        stats->activeJointIterations[0] = (long long)host.active[0];
        stats->activeJointIterations[1] = (long long)host.active[1];
    }
    return PHYX_B200_OK;
}

} // namespace phyx
