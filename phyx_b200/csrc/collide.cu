// phyx_b200 — the stages between broadphase and solve, resident on the device so a step never leaves
// HBM: manifold cache, box-box narrowphase, manifold packing, contact-joint cache.
//
// Reference stages replaced here:
//   Collider::UpdatePairs*         src/Collider.cpp:286-366  new manifold for every overlapping pair that is
//                                                            not in the cache yet, appended in sweep order
//   UpdateManifold & helpers       src/Collider.cpp:8-245    SAT (least penetration) + support-set clipping,
//   (Collider::UpdateManifolds     src/Collider.cpp:368-377)  <= 2 points, merged with last step's by proximity
//   Geom support mapping           src/Geom.h:11-77
//   Collider::PackManifolds        src/Collider.cpp:379-416  drop empty & separated manifolds (swap-with-last)
//   World::RefreshContactJoints    src/World.cpp:72-149      match / create / delete persistent joints
//
// Order is part of the result: manifolds are appended in sweep order, removals fill
// holes with the LAST element, new joints are appended in (manifold, point) order.  The sequential
// swap-with-last loops are reproduced in parallel: after a pass of removals, the k-th hole (ascending)
// receives the k-th surviving element counted from the end — that is exactly what the serial loop
// leaves behind — so the arrays are bit-identical to the reference's.
//
// All float arithmetic mirrors the reference operation by operation (--fmad=false).
#include "common.cuh"
#include "pairset.cuh"
#include "scan.cuh"

namespace phyx
{

constexpr int kBlock = 256;

// ================================================================================================
// UpdatePairs: append the sweep's new pairs as manifolds
// ================================================================================================

__global__ void __launch_bounds__(kBlock) k_append_manifolds(Count count, const int2* __restrict__ pairs, int first, const int* __restrict__ firstPtr,
    int2* __restrict__ manBody, int* __restrict__ manCount, int* __restrict__ manColour, unsigned long long* __restrict__ table, size_t mask)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count_of(count)) return;
    if (firstPtr) first = __ldcg(firstPtr);   // (deferred step: the manifold count lives on the device)
    int2 p = pairs[k];
    manBody[first + k] = p;        // Manifold(index_i, index_j, size*2): pointIndex is implicit (2*m)
    manCount[first + k] = 0;
    manColour[first + k] = -1;     // coloured when it gets its first contact point (colour.cu)
    if (table) pair_insert(table, mask, pair_key(unsigned(p.x), unsigned(p.y)));
}

__global__ void __launch_bounds__(kBlock) k_table_insert(Count count, const int2* __restrict__ manBody, unsigned long long* __restrict__ table, size_t mask)
{
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= count_of(count)) return;
    int2 p = manBody[m];
    pair_insert(table, mask, pair_key(unsigned(p.x), unsigned(p.y)));
}

static int reserve_manifolds(phyx_b200_ctx* c, int count)
{
    size_t n = size_t(count > 0 ? count : 1);
    size_t have = size_t(c->manifoldCount);
    PHYX_TRY(c->manBody.reserve_keep(n * sizeof(int2), have * sizeof(int2), c->stream));
    PHYX_TRY(c->manCount.reserve_keep(n * sizeof(int), have * sizeof(int), c->stream));
    PHYX_TRY(c->manColour.reserve_keep(n * sizeof(int), have * sizeof(int), c->stream));
    PHYX_TRY(c->contactPoints.reserve_keep(n * 2 * sizeof(phyx_contact_point), have * 2 * sizeof(phyx_contact_point), c->stream));
    return PHYX_B200_OK;
}

// (re)build the pair table from the live manifolds; sized for a load factor <= 1/4 with headroom
int collide_rebuild_pair_table(phyx_b200_ctx* c) { return collide_rebuild_pair_table_for(c, c->manifoldCount); }

int collide_rebuild_pair_table_for(phyx_b200_ctx* c, int manifolds)
{
    size_t want = 1024;
    while (want < size_t(manifolds + c->bodyCount) * 4) want <<= 1;
    if (want > c->pairTableSlots)
    {
        PHYX_TRY(c->pairTable.reserve(want * sizeof(unsigned long long)));
        c->pairTableSlots = want;
    }
    PHYX_CUDA(cudaMemsetAsync(c->pairTable.ptr, 0xff, c->pairTableSlots * sizeof(unsigned long long), c->stream));
    if (c->manifoldCount > 0)
    {
        k_table_insert<<<(c->manifoldCount + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(c->count(c->manifoldCount, &StepCtl::manifolds), c->manBody.as<int2>(),
            c->pairTable.as<unsigned long long>(), c->pairTableSlots - 1);
        c->launches++;
    }
    PHYX_CUDA(cudaGetLastError());
    return PHYX_B200_OK;
}

int collide_update_pairs(phyx_b200_ctx* c, phyx_b200_broadphase_stats* stats)
{
    PHYX_TRY(broadphase_sweep(c, stats, true));
    const int fresh = int(c->lastNewPairs);
    if (fresh == 0) return PHYX_B200_OK;
    PHYX_TRY(reserve_manifolds(c, c->manifoldCount + fresh));
    // the table must stay sparse (linear probing never terminates on a full table): if the additions
    // would push it past half full, append without inserting and rebuild it at the right size
    const bool rebuild = size_t(c->manifoldCount + fresh) * 2 > c->pairTableSlots;
    // (deferred step: `fresh` and manifoldCount are bounds, the kernel reads the counts the sweep left on the device)
    k_append_manifolds<<<(fresh + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(c->count(fresh, &StepCtl::newPairs), c->pairs.as<int2>(), c->manifoldCount,
        c->def.active ? &c->ctl()->appendFirst : nullptr, c->manBody.as<int2>(),
        c->manCount.as<int>(), c->manColour.as<int>(), rebuild ? nullptr : c->pairTable.as<unsigned long long>(), c->pairTableSlots - 1);
    c->launches++;
    c->manifoldCount += fresh;
    c->contactPointCount = 2 * c->manifoldCount;
    PHYX_CUDA(cudaGetLastError());
    if (rebuild) PHYX_TRY(collide_rebuild_pair_table(c));
    return PHYX_B200_OK;
}

// ================================================================================================
// Narrowphase
// ================================================================================================

struct V2
{
    float x, y;
};
__device__ __forceinline__ V2 v2(float x, float y) { V2 r; r.x = x; r.y = y; return r; }
__device__ __forceinline__ V2 operator+(V2 a, V2 b) { return v2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ V2 operator-(V2 a, V2 b) { return v2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ V2 operator-(V2 a) { return v2(-a.x, -a.y); }
__device__ __forceinline__ V2 operator*(V2 a, float s) { return v2(a.x * s, a.y * s); }
__device__ __forceinline__ float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ float sqlen(V2 a) { return a.x * a.x + a.y * a.y; }
__device__ __forceinline__ V2 perp(V2 a) { return v2(-a.y, a.x); }

struct Box
{
    V2 pos, ax, ay, half;
};

struct Point   // ContactPoint, src/Manifold.h:12-43
{
    V2 d1, d2, n;
    int merged, fresh, solver;
};

__device__ __forceinline__ Box load_box(int b, const float4* __restrict__ params, const float4* __restrict__ rot, const float2* __restrict__ size)
{
    float4 p = params[b], r = rot[b];
    float2 s = size[b];
    Box g;
    g.pos = v2(p.z, p.w);
    g.ax = v2(r.x, r.y);
    g.ay = v2(r.z, r.w);
    g.half = v2(s.x, s.y);
    return g;
}

// ComputeSeparatingAxis, src/Collider.cpp:8-55
__device__ bool least_penetration_axis(const Box& A, const Box& B, V2& axisOut)
{
    const V2 d = A.pos - B.pos;
    float dot00 = fabsf(dot(A.ax, B.ax));
    float dot01 = fabsf(dot(A.ax, B.ay));
    float reach = A.half.x + B.half.x * dot00 + B.half.y * dot01;
    float gap = fabsf(dot(A.ax, d)) - reach;
    if (gap > 0) return false;
    float best = gap;
    V2 axis = A.ax;

    float dot10 = fabsf(dot(A.ay, B.ax));
    float dot11 = fabsf(dot(A.ay, B.ay));
    reach = A.half.y + B.half.x * dot10 + B.half.y * dot11;
    gap = fabsf(dot(A.ay, d)) - reach;
    if (gap > 0) return false;
    if (gap > best) { best = gap; axis = A.ay; }

    reach = B.half.x + A.half.x * dot00 + A.half.y * dot10;
    gap = fabsf(dot(B.ax, d)) - reach;
    if (gap > 0) return false;
    if (gap > best) { best = gap; axis = B.ax; }

    reach = B.half.y + A.half.x * dot01 + A.half.y * dot11;
    gap = fabsf(dot(B.ay, d)) - reach;
    if (gap > 0) return false;
    if (gap > best) { best = gap; axis = B.ay; }

    axisOut = axis;
    return true;
}

// Geom::GetSupportPointSet / GetClippingEdge / GetClippingVertex, src/Geom.h:11-77
__device__ int support_set(const Box& g, V2 axis, V2& p1, V2& p2)
{
    const V2 xdim = g.ax * g.half.x;
    const V2 ydim = g.ay * g.half.y;
    const float xdiff = dot(axis, g.ax);
    const float ydiff = dot(axis, g.ay);
    if (fabsf(xdiff) < 0.1f || fabsf(ydiff) < 0.1f)
    {
        V2 a = g.pos, b = g.pos;
        V2 offset = v2(0.f, 0.f);
        if (fabsf(xdiff) < fabsf(ydiff))
        {
            if (dot(axis, ydim) > 0.0f) { offset = offset + ydim; a = a + xdim; b = b - xdim; }
            else                        { offset = offset - ydim; a = a - xdim; b = b + xdim; }
        }
        else
        {
            if (dot(axis, xdim) > 0.0f) { offset = offset + xdim; a = a - ydim; b = b + ydim; }
            else                        { offset = offset - xdim; a = a + ydim; b = b - ydim; }
        }
        p1 = a + offset;
        p2 = b + offset;
        return 2;
    }
    float xs = xdiff < 0.0f ? -1.0f : 1.0f;
    float ys = ydiff < 0.0f ? -1.0f : 1.0f;
    p1 = g.pos + xdim * xs + ydim * ys;
    return 1;
}

// ProjectPointToLine, src/Vector2.h:276-281
__device__ __forceinline__ V2 project_along(V2 point, V2 linePoint, V2 n, V2 dir)
{
    float mult = __fdiv_rn(1.0f, dot(dir, n));
    return point + dir * (dot(linePoint, n) - dot(point, n)) * mult;
}

__device__ __forceinline__ bool within_segment(V2 p, V2 a, V2 b) { return (dot(p - a, b - a) >= 0.0f) && (dot(p - b, a - b) >= 0.0f); }

// AddPoint, src/Collider.cpp:57-92
__device__ void merge_point(Point* pts, int& count, V2 point1, V2 point2, V2 axis, V2 pos1, V2 pos2)
{
    Point fresh;
    fresh.d1 = point1 - pos1;   // ContactPoint ctor, src/Manifold.h:18-26
    fresh.d2 = point2 - pos2;
    fresh.n = axis;
    int closest = -1;
    float best = 3.402823466e+38f;
    for (int i = 0; i < count; ++i)
    {
        float a = sqlen(fresh.d1 - pts[i].d1), b = sqlen(fresh.d2 - pts[i].d2);
        // ContactPoint::Equals(tolerance 2): false only if BOTH deltas moved by more than 2
        if (!((sqlen(pts[i].d1 - fresh.d1) > 4.0f) && (sqlen(pts[i].d2 - fresh.d2) > 4.0f)))
        {
            float dist = a + b;
            if (dist < best)
            {
                best = dist;
                closest = i;
            }
        }
    }
    if (closest >= 0)
    {
        pts[closest].merged = 1;
        pts[closest].fresh = 0;
        pts[closest].n = fresh.n;
        pts[closest].d1 = fresh.d1;
        pts[closest].d2 = fresh.d2;
    }
    else if (count < 4)
    {
        fresh.merged = 1;
        fresh.fresh = 1;
        fresh.solver = -1;
        pts[count++] = fresh;
    }
}

// GenerateContacts, src/Collider.cpp:94-211
__device__ void generate_contacts(const Box& A, const Box& B, Point* pts, int& count, V2 axis)
{
    if (dot(axis, A.pos - B.pos) < 0.0f) axis = -axis;
    V2 s1a, s1b, s2a, s2b;
    int n1 = support_set(A, -axis, s1a, s1b);
    int n2 = support_set(B, axis, s2a, s2b);
    if (n1 == 2 && sqlen(s1a - s1b) < 4.0f) { s1a = (s1a + s1b) * 0.5f; n1 = 1; }
    if (n2 == 2 && sqlen(s2a - s2b) < 4.0f) { s2a = (s2a + s2b) * 0.5f; n2 = 1; }

    if (n1 == 1 && n2 == 1)
    {
        V2 delta = s2a - s1a;
        if (dot(delta, axis) >= 0.0f) merge_point(pts, count, s1a, s2a, axis, A.pos, B.pos);
    }
    else if (n1 == 1 && n2 == 2)
    {
        V2 n = perp(s2b - s2a);
        V2 p = project_along(s1a, s2a, n, axis);
        if (within_segment(p, s2a, s2b)) merge_point(pts, count, s1a, p, axis, A.pos, B.pos);
    }
    else if (n1 == 2 && n2 == 1)
    {
        V2 n = perp(s1b - s1a);
        V2 p = project_along(s2a, s1a, n, axis);
        if (within_segment(p, s1a, s1b)) merge_point(pts, count, p, s2a, axis, A.pos, B.pos);
    }
    else
    {
        V2 on1[4], on2[4];
        int found = 0;
        const V2 e1[2] = { s1a, s1b }, e2[2] = { s2a, s2b };
        for (int i = 0; i < 2; ++i)
        {
            V2 n = perp(s2b - s2a);
            if (dot(e1[i] - s2a, n) >= 0.0f)
            {
                V2 p = project_along(e1[i], s2a, n, axis);
                if (within_segment(p, s2a, s2b)) { on1[found] = e1[i]; on2[found] = p; found++; }
            }
        }
        for (int i = 0; i < 2; ++i)
        {
            V2 n = perp(s1b - s1a);
            if (dot(e2[i] - s1a, n) >= 0.0f)
            {
                V2 p = project_along(e2[i], s1a, n, axis);
                if (within_segment(p, s1a, s1b)) { on1[found] = p; on2[found] = e2[i]; found++; }
            }
        }
        if (found == 1) merge_point(pts, count, on1[0], on2[0], axis, A.pos, B.pos);
        if (found >= 2)
        {
            merge_point(pts, count, on1[0], on2[0], axis, A.pos, B.pos);
            merge_point(pts, count, on1[1], on2[1], axis, A.pos, B.pos);
        }
    }
}

// UpdateManifold, src/Collider.cpp:213-245: one thread per manifold; contact point = 2 x float4
__global__ void __launch_bounds__(kBlock) k_update_manifolds(Count count, const int2* __restrict__ manBody, int* __restrict__ manCount,
    float4* __restrict__ contactPoints, const float4* __restrict__ params, const float4* __restrict__ rot, const float2* __restrict__ size)
{
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= count_of(count)) return;
    const int2 bodies = manBody[m];
    const int had = manCount[m];
    Point pts[4];
    for (int i = 0; i < had; ++i)
    {
        float4 c0 = contactPoints[size_t(2 * m + i) * 2], c1 = contactPoints[size_t(2 * m + i) * 2 + 1];
        pts[i].d1 = v2(c0.x, c0.y);
        pts[i].d2 = v2(c0.z, c0.w);
        pts[i].n = v2(c1.x, c1.y);
        pts[i].merged = 0;
        pts[i].fresh = 0;
        pts[i].solver = __float_as_int(c1.w);
    }
    int n = had;
    const Box A = load_box(bodies.x, params, rot, size), B = load_box(bodies.y, params, rot, size);
    V2 axis;
    if (least_penetration_axis(A, B, axis)) generate_contacts(A, B, pts, n, axis);
    int kept = 0;
    for (int i = 0; i < n; ++i)
    {
        if (!pts[i].merged) continue;
        if (kept < 2)
        {
            int flags = 1 | (pts[i].fresh ? 0x100 : 0);   // isMerged = 1, isNewlyCreated
            contactPoints[size_t(2 * m + kept) * 2] = make_float4(pts[i].d1.x, pts[i].d1.y, pts[i].d2.x, pts[i].d2.y);
            contactPoints[size_t(2 * m + kept) * 2 + 1] = make_float4(pts[i].n.x, pts[i].n.y, __int_as_float(flags), __int_as_float(pts[i].solver));
        }
        kept++;
    }
    manCount[m] = kept < 2 ? kept : 2;
}

int collide_update_manifolds(phyx_b200_ctx* c)
{
    int M = c->manifoldCount;
    c->jointUnitsValid = false;   // new contact points have no joint until RefreshContactJoints
    if (M == 0) return PHYX_B200_OK;
    k_update_manifolds<<<(M + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(c->count(M, &StepCtl::manifolds), c->manBody.as<int2>(), c->manCount.as<int>(),
        c->contactPoints.as<float4>(), c->params.as<float4>(), c->rot.as<float4>(), c->size.as<float2>());
    c->launches++;
    PHYX_CUDA(cudaGetLastError());
    return PHYX_B200_OK;
}

// ================================================================================================
// swap-with-last compaction (shared by PackManifolds and the joint cleanup)
// ================================================================================================
// alive[i] (0/1) -> prefix (exclusive scan).  K = number alive.  Movers = alive elements at index
// >= K, numbered from the END; holes = dead elements at index < K, numbered from the FRONT.

// (the flags themselves are never stored: element i is alive iff the exclusive prefix steps up after it)
__device__ __forceinline__ bool alive_at(const int* __restrict__ prefix, int i, int n, int total) { return (i + 1 < n ? prefix[i + 1] : total) != prefix[i]; }

__global__ void __launch_bounds__(kBlock) k_list_movers(Count nc, const int* __restrict__ prefix, const int* __restrict__ totalPtr, int* __restrict__ moverIndex)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = count_of(nc);
    if (i >= n) return;
    const int K = *totalPtr;
    if (i >= K && alive_at(prefix, i, n, K)) moverIndex[K - prefix[i] - 1] = i;   // alive elements after i: K - prefix[i] - 1
}

// ================================================================================================
// PackManifolds
// ================================================================================================

// alive flag of a manifold, computed inside the scan that ranks the survivors (scan.cuh): Collider.cpp:390
struct ManifoldAlive
{
    const int2* manBody;
    const int* manCount;
    const float4* aabb;
    const int* manColour;
    unsigned long long* bodyUsed;
    __device__ __forceinline__ bool vector_ok(const int*) const { return false; }
    __device__ __forceinline__ void load4(int, int (&)[4]) const {}
    __device__ __forceinline__ int load(int m) const
    {
        int2 b = manBody[m];
        float4 a1 = aabb[b.x], a2 = aabb[b.y];
        // AABB2::Intersects, src/AABB2.h:18-23
        bool apart = (a1.x > a2.z) || (a2.x > a1.z) || (a1.y > a2.w) || (a2.y > a1.w);
        const bool keep = !(manCount[m] == 0 && apart);   // Collider.cpp:390
        // a removed manifold hands its solver colour back to its bodies (colour.cu; the masks of static bodies are never read)
        if (!keep && bodyUsed && manColour[m] >= 0)
        {
            const unsigned long long mask = ~(1ull << manColour[m]);
            atomicAnd(&bodyUsed[b.x], mask);
            atomicAnd(&bodyUsed[b.y], mask);
        }
        return keep ? 1 : 0;
    }
};

__global__ void __launch_bounds__(kBlock) k_manifold_fill(Count nc, const int* __restrict__ prefix, const int* __restrict__ totalPtr,
    const int* __restrict__ moverIndex, int2* __restrict__ manBody, int* __restrict__ manCount, int* __restrict__ manColour,
    float4* __restrict__ contactPoints)
{
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    const int K = *totalPtr, n = count_of(nc);
    if (m >= n || m >= K || alive_at(prefix, m, n, K)) return;
    const int src = moverIndex[m - prefix[m]];   // holes before m: m - prefix[m]
    const int cnt = manCount[src];
    manBody[m] = manBody[src];
    manCount[m] = cnt;
    manColour[m] = manColour[src];
    for (int k = 0; k < cnt; ++k)   // Collider.cpp:400-401: only the live points travel
    {
        contactPoints[size_t(2 * m + k) * 2] = contactPoints[size_t(2 * src + k) * 2];
        contactPoints[size_t(2 * m + k) * 2 + 1] = contactPoints[size_t(2 * src + k) * 2 + 1];
    }
}

// deferred step: small single-thread kernels move the counts along on the device
__global__ void k_ctl_packed(StepCtl* ctl, const int* __restrict__ total)
{
    if (threadIdx.x != 0 || blockIdx.x != 0 || ctl->stop) return;
    ctl->packed = *total;
    ctl->manifolds = *total;
}

// new joints are known: check the bound, fix the joint count the cleanup starts from
__global__ void k_gate_fresh(StepCtl* ctl, const int* __restrict__ total, int capFresh)
{
    if (threadIdx.x != 0 || blockIdx.x != 0 || ctl->stop) return;
    const int fresh = *total;
    ctl->fresh = fresh;
    if (fresh > capFresh)
        ctl_stop(ctl, kStageRefresh, fresh, 3);
    else
        ctl->jointsGrown = ctl->joints + fresh;
}

__global__ void k_ctl_joints(StepCtl* ctl, const int* __restrict__ total)
{
    if (threadIdx.x != 0 || blockIdx.x != 0 || ctl->stop) return;
    ctl->jointsKept = *total;
    ctl->created = ctl->fresh;
    ctl->deleted = ctl->jointsGrown - *total;
    ctl->joints = *total;
}

int collide_pack_manifolds(phyx_b200_ctx* c)
{
    const int M = c->manifoldCount;
    c->jointUnitsValid = false;   // contact points move; RefreshContactJoints re-links the joints
    if (M == 0) return PHYX_B200_OK;
    // scratch: prefix[M] | movers[M] | total  (the alive flags are computed inside the scan and never stored)
    PHYX_TRY(c->collideTmp.reserve((size_t(M) * 2 + 4) * sizeof(int)));
    int* prefix = c->collideTmp.as<int>();
    int* movers = prefix + M;
    int* total = movers + M;
    const int grid = (M + kBlock - 1) / kBlock;
    const bool track = c->colourStateValid && c->colourStateBodies == c->bodyCount && c->bodyUsed.ptr;
    const Count Mc = c->count(M, &StepCtl::manifolds);
    ManifoldAlive aliveOf = { c->manBody.as<int2>(), c->manCount.as<int>(), c->aabb.as<float4>(), c->manColour.as<int>(),
        track ? c->bodyUsed.as<unsigned long long>() : nullptr };
    PHYX_TRY(exclusive_scan_with(c, aliveOf, prefix, Mc, total));
    k_list_movers<<<grid, kBlock, 0, c->stream>>>(Mc, prefix, total, movers);
    k_manifold_fill<<<grid, kBlock, 0, c->stream>>>(Mc, prefix, total, movers, c->manBody.as<int2>(), c->manCount.as<int>(),
        c->manColour.as<int>(), c->contactPoints.as<float4>());
    c->launches += 2;
    if (c->def.active)
    {
        // deferred step: the survivor count stays on the device (M remains the host's bound); the cache is rebuilt from the
        // survivors whether or not anything was removed (in a moving world something always is)
        k_ctl_packed<<<1, 32, 0, c->stream>>>(c->ctl(), total);
        c->launches++;
        return collide_rebuild_pair_table(c);
    }
    int K = 0;
    PHYX_TRY(fetch_small(c, total, sizeof(int), &K));
    const bool removed = K != M;
    c->manifoldCount = K;
    c->contactPointCount = 2 * K;
    // removed keys must leave the cache (Collider.cpp:392): rebuild it from the survivors
    if (removed) PHYX_TRY(collide_rebuild_pair_table(c));
    return PHYX_B200_OK;
}

// ================================================================================================
// RefreshContactJoints
// ================================================================================================

// World.cpp:83-86 marks every joint dead (contactPointIndex = -1) before the contact points claim theirs.  Here a joint is
// alive iff its STAMP is this refresh's epoch: the match kernel stamps the joints it touches, nothing has to be reset (a
// pass over the 20-byte joint records just to clear one word), and the survivor scan reads 4 contiguous bytes per joint.
// The epoch lives in device memory (its launch parameter must not change from step to step: graph replay).
__global__ void k_joint_epoch(int* __restrict__ epoch)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) *epoch += 1;
}

// 1 for live contact points without a joint yet (solverIndex < 0), computed inside the scan that ranks them
struct PointIsNew
{
    const int* manCount;
    const float4* contactPoints;
    __device__ __forceinline__ bool vector_ok(const int*) const { return false; }
    __device__ __forceinline__ void load4(int, int (&)[4]) const {}
    __device__ __forceinline__ int load(int p) const
    {
        const bool live = (p & 1) < manCount[p >> 1];
        return (live && __float_as_int(contactPoints[size_t(p) * 2 + 1].w) < 0) ? 1 : 0;
    }
};

struct JointAlive
{
    const int* stamp;
    const int* epoch;
    __device__ __forceinline__ bool vector_ok(const int*) const { return false; }
    __device__ __forceinline__ void load4(int, int (&)[4]) const {}
    __device__ __forceinline__ int load(int j) const { return stamp[j] == __ldg(epoch) ? 1 : 0; }
};

// World.cpp:91-124: new points get a joint appended in (manifold, point) order, known points re-attach
__global__ void __launch_bounds__(kBlock) k_joint_match(Count numPoints, Count oldJointsC, const int2* __restrict__ manBody, const int* __restrict__ manCount,
    float4* __restrict__ contactPoints, const int* __restrict__ newRank, const int* __restrict__ newTotal, phyx_contact_joint* __restrict__ joints,
    int* __restrict__ stamp, const int* __restrict__ epochPtr)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int nP = count_of(numPoints);
    if (p >= nP) return;
    const int oldJoints = count_of(oldJointsC);
    const int epoch = __ldg(epochPtr);
    if ((p & 1) >= manCount[p >> 1]) return;
    if (alive_at(newRank, p, nP, *newTotal))
    {
        int j = oldJoints + newRank[p];
        int2 b = manBody[p >> 1];
        phyx_contact_joint jt;
        jt.contactPointIndex = p;
        jt.body1Index = b.x;
        jt.body2Index = b.y;
        jt.normalLimiter_accumulatedImpulse = 0.f;
        jt.frictionLimiter_accumulatedImpulse = 0.f;
        joints[j] = jt;
        stamp[j] = epoch;
        reinterpret_cast<int*>(contactPoints + size_t(p) * 2 + 1)[3] = j;
    }
    else
    {
        int j = __float_as_int(contactPoints[size_t(p) * 2 + 1].w);
        joints[j].contactPointIndex = p;
        stamp[j] = epoch;
    }
}

__global__ void __launch_bounds__(kBlock) k_joint_fill(Count nc, const int* __restrict__ prefix, const int* __restrict__ totalPtr,
    const int* __restrict__ moverIndex, phyx_contact_joint* __restrict__ joints)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int K = *totalPtr, n = count_of(nc);
    if (j >= n || j >= K || alive_at(prefix, j, n, K)) return;
    joints[j] = joints[moverIndex[j - prefix[j]]];   // World.cpp:131-134
}

__global__ void __launch_bounds__(kBlock) k_joint_backlink(const int* __restrict__ totalPtr, const phyx_contact_joint* __restrict__ joints,
    float4* __restrict__ contactPoints)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= *totalPtr) return;
    reinterpret_cast<int*>(contactPoints + size_t(joints[j].contactPointIndex) * 2 + 1)[3] = j;   // World.cpp:140
}

// stamps of the joints (see k_joint_epoch); a new buffer starts from zeros (epochs start at 1)
static int reserve_stamps(phyx_b200_ctx* c, int joints)
{
    const void* before = c->jointStamp.ptr;
    PHYX_TRY(c->jointStamp.reserve(size_t(joints > 0 ? joints : 1) * sizeof(int)));
    if (c->jointStamp.ptr != before) PHYX_CUDA(cudaMemsetAsync(c->jointStamp.ptr, 0, c->jointStamp.cap, c->stream));
    return PHYX_B200_OK;
}

int collide_refresh_joints(phyx_b200_ctx* c, int* matched, int* created, int* deleted)
{
    // deferred step: J0 is exact (nothing before this stage changes the joints), P, `fresh`, J1 are bounds; the true counts
    // are StepCtl::joints / manifolds x 2 / fresh / jointsGrown / jointsKept
    const bool deferred = c->def.active;
    const int J0 = c->jointCount, P = 2 * c->manifoldCount;
    int fresh = 0;
    const Count J0c = c->count(J0, &StepCtl::joints);
    int* epoch = c->mailboxSeqDev + 1;   // (a spare word of that small device block)
    k_joint_epoch<<<1, 32, 0, c->stream>>>(epoch);
    c->launches++;
    // scratch A (points): newRank[P] | total
    size_t needA = size_t(P) + 4;
    PHYX_TRY(c->collideTmp.reserve(needA * sizeof(int)));
    if (P > 0)
    {
        int* newRank = c->collideTmp.as<int>();
        int* total = newRank + P;
        const Count Pc = c->count(P, &StepCtl::manifolds, 2);
        PointIsNew isNewOf = { c->manCount.as<int>(), c->contactPoints.as<float4>() };
        PHYX_TRY(exclusive_scan_with(c, isNewOf, newRank, Pc, total));
        if (deferred)
        {
            fresh = c->def.capFresh;
            k_gate_fresh<<<1, 32, 0, c->stream>>>(c->ctl(), total, fresh);
            c->launches++;
        }
        else
            PHYX_TRY(fetch_small(c, total, sizeof(int), &fresh));
        PHYX_TRY(c->joints.reserve_keep(size_t(J0 + fresh > 0 ? J0 + fresh : 1) * sizeof(phyx_contact_joint), size_t(J0) * sizeof(phyx_contact_joint), c->stream));
        PHYX_TRY(reserve_stamps(c, J0 + fresh));
        k_joint_match<<<(P + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(Pc, J0c, c->manBody.as<int2>(), c->manCount.as<int>(),
            c->contactPoints.as<float4>(), newRank, total, c->joints.as<phyx_contact_joint>(), c->jointStamp.as<int>(), epoch);
        c->launches++;
    }
    const int J1 = J0 + fresh;
    int K = 0;
    if (J1 > 0)
    {
        // scratch B (joints): prefix[J1] | movers[J1] | total  (scratch A is dead by now)
        PHYX_TRY(c->collideTmp.reserve((size_t(J1) * 2 + 4) * sizeof(int)));
        int* prefix = c->collideTmp.as<int>();
        int* movers = prefix + J1;
        int* total = movers + J1;
        const int grid = (J1 + kBlock - 1) / kBlock;
        const Count J1c = c->count(J1, &StepCtl::jointsGrown);
        PHYX_TRY(reserve_stamps(c, J1));   // (P == 0: no match ran, every stamp is older than the epoch: all joints go)
        JointAlive aliveOf = { c->jointStamp.as<int>(), epoch };
        PHYX_TRY(exclusive_scan_with(c, aliveOf, prefix, J1c, total));
        k_list_movers<<<grid, kBlock, 0, c->stream>>>(J1c, prefix, total, movers);
        k_joint_fill<<<grid, kBlock, 0, c->stream>>>(J1c, prefix, total, movers, c->joints.as<phyx_contact_joint>());
        k_joint_backlink<<<grid, kBlock, 0, c->stream>>>(total, c->joints.as<phyx_contact_joint>(), c->contactPoints.as<float4>());
        c->launches += 3;
        if (deferred)
        {
            k_ctl_joints<<<1, 32, 0, c->stream>>>(c->ctl(), total);
            c->launches++;
            K = J1;   // bound
        }
        else
            PHYX_TRY(fetch_small(c, total, sizeof(int), &K));
    }
    PHYX_CUDA(cudaGetLastError());
    c->jointCount = K;
    c->jointUnitsValid = true;   // every live contact point now names its joint (solverIndex) and vice versa
    if (created) *created = fresh;
    if (deleted) *deleted = J1 - K;
    if (matched) *matched = K - fresh > 0 ? K - fresh : 0;
    return PHYX_B200_OK;
}

int collide_reset(phyx_b200_ctx* c)
{
    c->colourStateValid = false;
    c->jointUnitsValid = false;
    c->manifoldCount = 0;
    c->contactPointCount = 0;
    c->jointCount = 0;
    if (c->pairTableSlots) PHYX_CUDA(cudaMemsetAsync(c->pairTable.ptr, 0xff, c->pairTableSlots * sizeof(unsigned long long), c->stream));
    return PHYX_B200_OK;
}

} // namespace phyx
