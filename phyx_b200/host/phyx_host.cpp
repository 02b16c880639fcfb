// phyx_b200 host mirror — implementation (see phyx_host.h).
//
// Stage map (reference src/World.cpp:19-37):
//   1 IntegrateVelocity      -> phyx_b200_integrate_velocity                      (device)
//   2 UpdateBroadphase       -> phyx_b200_update_broadphase                       (device)
//   3 UpdatePairs            -> phyx_b200_sweep_pairs + pair-cache filter         (device + host)
//   4 UpdateManifolds        -> box-box SAT + clipping                             (host, §8f "next")
//   5 PackManifolds          -> swap-with-last compaction                          (host)
//   6 RefreshContactJoints   -> joint cache match / create / delete                (host)
//   7 SolveJoints            -> phyx_b200_solve_joints                             (device)
//   8 IntegratePosition      -> phyx_b200_integrate_position                       (device)
//
// Compile with -ffp-contract=off: the host stages perform the reference's float operations in the
// reference's order so that whole-step results are bit-comparable with its strict-FP build.
#include "phyx_host.h"

#include <chrono>
#include <cstdio>
#include <limits>
#include <thread>

namespace phyx_host
{

static double now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

struct StageTimer
{
    double* slot;
    double t0;
    explicit StageTimer(double* s) : slot(s), t0(now_ms()) {}
    ~StageTimer() { *slot += now_ms() - t0; }
};

void fail(const char* what, int status)
{
    // the reference has no error channel (no exceptions, no return codes: SURVEY §8b); a failed GPU
    // call must not silently continue on stale data, and there is no CPU path to fall back to
    std::fprintf(stderr, "phyx_b200: %s failed (status %d): %s\n", what, status, phyx_b200_last_error());
    std::abort();
}

#define PHYX_CALL(expr)                   \
    do                                    \
    {                                     \
        int st_ = (expr);                 \
        if (st_ != 0) phyx_host::fail(#expr, st_); \
    } while (0)

void Device::ensure()
{
    if (!ctx) PHYX_CALL(phyx_b200_create(deviceIndex, &ctx));
}

void Device::upload(RigidBody* bodies, int count)
{
    ensure();
    PHYX_CALL(phyx_b200_upload_bodies(ctx, reinterpret_cast<const phyx_rigid_body*>(bodies), count));
    resident = true;
    residentCount = count;
}

void Device::download(RigidBody* bodies, int count)
{
    PHYX_CALL(phyx_b200_download_bodies(ctx, reinterpret_cast<phyx_rigid_body*>(bodies), count));
}

Device::~Device()
{
    if (ctx) phyx_b200_destroy(ctx);
}

} // namespace phyx_host

using phyx_host::Device;
using phyx_host::StageTimer;

unsigned int WorkQueue::getIdealWorkerCount() { return std::max(std::thread::hardware_concurrency(), 1u); }

// ==== narrowphase (host; reference src/Collider.cpp:8-245, src/Geom.h:11-77) ===========================

namespace
{

struct Box
{
    Vector2f pos, ax, ay, half;
};

inline Box box_of(const RigidBody& b)
{
    Box r;
    r.pos = b.geom.coords.pos;
    r.ax = b.geom.coords.xVector;
    r.ay = b.geom.coords.yVector;
    r.half = b.geom.size;
    return r;
}

// Separating-axis test over the four face normals; returns the axis of least penetration
// (reference ComputeSeparatingAxis, src/Collider.cpp:8-55).
bool least_penetration_axis(const RigidBody& b1, const RigidBody& b2, Vector2f& axisOut)
{
    const Vector2f a0[2] = { b1.coords.xVector, b1.coords.yVector };
    const Vector2f a1[2] = { b2.coords.xVector, b2.coords.yVector };
    const Vector2f e0 = b1.geom.size, e1 = b2.geom.size;
    const Vector2f d = b1.coords.pos - b2.coords.pos;

    float dot00 = std::fabs(a0[0] * a1[0]);
    float dot01 = std::fabs(a0[0] * a1[1]);
    float reach = e0.x + e1.x * dot00 + e1.y * dot01;
    float gap = std::fabs(a0[0] * d) - reach;
    if (gap > 0) return false;
    float best = gap;
    Vector2f axis = a0[0];

    float dot10 = std::fabs(a0[1] * a1[0]);
    float dot11 = std::fabs(a0[1] * a1[1]);
    reach = e0.y + e1.x * dot10 + e1.y * dot11;
    gap = std::fabs(a0[1] * d) - reach;
    if (gap > 0) return false;
    if (gap > best) { best = gap; axis = a0[1]; }

    reach = e1.x + e0.x * dot00 + e0.y * dot10;
    gap = std::fabs(a1[0] * d) - reach;
    if (gap > 0) return false;
    if (gap > best) { best = gap; axis = a1[0]; }

    reach = e1.y + e0.x * dot01 + e0.y * dot11;
    gap = std::fabs(a1[1] * d) - reach;
    if (gap > 0) return false;
    if (gap > best) { best = gap; axis = a1[1]; }

    axisOut = axis;
    return true;
}

// Support set of a box along `axis`: a face (2 points) when the axis is within ~5.7 degrees of a
// face normal, else a vertex (reference Geom::GetSupportPointSet / GetClippingEdge /
// GetClippingVertex, src/Geom.h:11-77).
int support_set(const Box& g, const Vector2f& axis, Vector2f out[2])
{
    const Vector2f xdim = g.ax * g.half.x;
    const Vector2f ydim = g.ay * g.half.y;
    const float xdiff = axis * g.ax;
    const float ydiff = axis * g.ay;
    if (std::fabs(xdiff) < 0.1f || std::fabs(ydiff) < 0.1f)
    {
        Vector2f p1 = g.pos, p2 = g.pos;
        Vector2f offset = Vector2f::zero();
        if (std::fabs(xdiff) < std::fabs(ydiff))
        {
            if (axis * ydim > 0.0f) { offset += ydim; p1 += xdim; p2 -= xdim; }
            else                    { offset -= ydim; p1 -= xdim; p2 += xdim; }
        }
        else
        {
            if (axis * xdim > 0.0f) { offset += xdim; p1 -= ydim; p2 += ydim; }
            else                    { offset -= xdim; p1 += ydim; p2 -= ydim; }
        }
        p1 += offset;
        p2 += offset;
        out[0] = p1;
        out[1] = p2;
        return 2;
    }
    float xs = xdiff < 0.0f ? -1.0f : 1.0f;
    float ys = ydiff < 0.0f ? -1.0f : 1.0f;
    out[0] = g.pos + xs * xdim + ys * ydim;
    return 1;
}

// projection of `point` along `dir` onto the line through `linePoint` with normal `n`
// (reference ProjectPointToLine, src/Vector2.h:276-281)
inline Vector2f project_along(const Vector2f& point, const Vector2f& linePoint, const Vector2f& n, const Vector2f& dir)
{
    float mult = 1.0f / (dir * n);
    return point + dir * ((linePoint * n) - (point * n)) * mult;
}

inline bool within_segment(const Vector2f& p, const Vector2f& a, const Vector2f& b)
{
    return (((p - a) * (b - a)) >= 0.0f) && (((p - b) * (a - b)) >= 0.0f);
}

// merge a candidate into the manifold's working set (reference AddPoint, src/Collider.cpp:57-92)
void merge_point(ContactPoint* points, int& count, ContactPoint& fresh)
{
    ContactPoint* closest = nullptr;
    float best = std::numeric_limits<float>::max();
    for (int i = 0; i < count; ++i)
    {
        ContactPoint& old = points[i];
        if (fresh.Equals(old, 2.0f))
        {
            float dist = (fresh.delta1 - old.delta1).SquareLen() + (fresh.delta2 - old.delta2).SquareLen();
            if (dist < best)
            {
                best = dist;
                closest = &old;
            }
        }
    }
    if (closest)
    {
        closest->isMerged = 1;
        closest->isNewlyCreated = 0;
        closest->normal = fresh.normal;
        closest->delta1 = fresh.delta1;
        closest->delta2 = fresh.delta2;
    }
    else
    {
        assert(count < 4);
        fresh.isMerged = 1;
        fresh.isNewlyCreated = 1;
        points[count++] = fresh;
    }
}

// reference GenerateContacts, src/Collider.cpp:94-211
void generate_contacts(const RigidBody& b1, const RigidBody& b2, ContactPoint* points, int& count, Vector2f axis)
{
    if (axis * (b1.coords.pos - b2.coords.pos) < 0.0f) axis.Invert();

    Vector2f s1[2], s2[2];
    const float tol = 2.0f;
    int n1 = support_set(box_of(b1), -axis, s1);
    int n2 = support_set(box_of(b2), axis, s2);
    if (n1 == 2 && (s1[0] - s1[1]).SquareLen() < tol * tol) { s1[0] = (s1[0] + s1[1]) * 0.5f; n1 = 1; }
    if (n2 == 2 && (s2[0] - s2[1]).SquareLen() < tol * tol) { s2[0] = (s2[0] + s2[1]) * 0.5f; n2 = 1; }

    if (n1 == 1 && n2 == 1)
    {
        Vector2f delta = s2[0] - s1[0];
        if (delta * axis >= 0.0f)
        {
            ContactPoint c(s1[0], s2[0], axis, &b1, &b2);
            merge_point(points, count, c);
        }
    }
    else if (n1 == 1 && n2 == 2)
    {
        Vector2f n = (s2[1] - s2[0]).GetPerpendicular();
        Vector2f p = project_along(s1[0], s2[0], n, axis);
        if (within_segment(p, s2[0], s2[1]))
        {
            ContactPoint c(s1[0], p, axis, &b1, &b2);
            merge_point(points, count, c);
        }
    }
    else if (n1 == 2 && n2 == 1)
    {
        Vector2f n = (s1[1] - s1[0]).GetPerpendicular();
        Vector2f p = project_along(s2[0], s1[0], n, axis);
        if (within_segment(p, s1[0], s1[1]))
        {
            ContactPoint c(p, s2[0], axis, &b1, &b2);
            merge_point(points, count, c);
        }
    }
    else if (n1 == 2 && n2 == 2)
    {
        Vector2f onOne[4], onTwo[4];
        int found = 0;
        for (int i = 0; i < 2; ++i)
        {
            Vector2f n = (s2[1] - s2[0]).GetPerpendicular();
            if ((s1[i] - s2[0]) * n >= 0.0)
            {
                Vector2f p = project_along(s1[i], s2[0], n, axis);
                if (within_segment(p, s2[0], s2[1])) { onOne[found] = s1[i]; onTwo[found] = p; found++; }
            }
        }
        for (int i = 0; i < 2; ++i)
        {
            Vector2f n = (s1[1] - s1[0]).GetPerpendicular();
            if ((s2[i] - s1[0]) * n >= 0.0)
            {
                Vector2f p = project_along(s2[i], s1[0], n, axis);
                if (within_segment(p, s1[0], s1[1])) { onOne[found] = p; onTwo[found] = s2[i]; found++; }
            }
        }
        if (found == 1)
        {
            ContactPoint c(onOne[0], onTwo[0], axis, &b1, &b2);
            merge_point(points, count, c);
        }
        if (found >= 2)
        {
            ContactPoint c0(onOne[0], onTwo[0], axis, &b1, &b2);
            merge_point(points, count, c0);
            ContactPoint c1(onOne[1], onTwo[1], axis, &b1, &b2);
            merge_point(points, count, c1);
        }
    }
}

// reference UpdateManifold, src/Collider.cpp:213-245
void update_manifold(Manifold& m, const RigidBody* bodies, ContactPoint* points)
{
    ContactPoint work[kMaxContactPoints * 2];
    for (int i = 0; i < m.pointCount; ++i)
    {
        work[i] = points[i];
        work[i].isMerged = 0;
        work[i].isNewlyCreated = 0;
    }
    int count = m.pointCount;
    const RigidBody& b1 = bodies[m.body1Index];
    const RigidBody& b2 = bodies[m.body2Index];
    Vector2f axis;
    if (least_penetration_axis(b1, b2, axis)) generate_contacts(b1, b2, work, count, axis);
    m.pointCount = 0;
    for (int i = 0; i < count; ++i)
        if (work[i].isMerged)
        {
            assert(m.pointCount < kMaxContactPoints);
            points[m.pointCount++] = work[i];
        }
}

} // namespace

// ==== Collider =========================================================================================

Collider::Collider() {}

NOINLINE void Collider::UpdateBroadphase(RigidBody* bodies, size_t bodiesCount)
{
    StageTimer t(&device->stageMs[1]);
    if (!device->inUpdate) device->upload(bodies, int(bodiesCount));
    PHYX_CALL(phyx_b200_update_broadphase(device->ctx));
    if (mirrorBroadphase || !device->inUpdate)
    {
        broadphase.resize(int(bodiesCount));
        static_assert(sizeof(BroadphaseEntry) == sizeof(phyx_broadphase_entry), "BroadphaseEntry layout");
        PHYX_CALL(phyx_b200_download_broadphase(device->ctx, reinterpret_cast<phyx_broadphase_entry*>(broadphase.data), int(bodiesCount)));
    }
}

// reference UpdatePairs*, src/Collider.cpp:286-366: the device returns every overlapping pair in
// sweep order; pairs not yet in the cache become new manifolds, appended in that order
NOINLINE void Collider::UpdatePairs(WorkQueue&, RigidBody*, size_t bodiesCount)
{
    StageTimer t(&device->stageMs[2]);
    if (pairBuffer.size() < 4 * bodiesCount + 1024) pairBuffer.resize(4 * bodiesCount + 1024);
    int64_t count = 0;
    int st = phyx_b200_sweep_pairs(device->ctx, pairBuffer.data(), int64_t(pairBuffer.size()), &count, &device->lastBroadphase);
    if (st == PHYX_B200_ERR_CAPACITY)
    {
        pairBuffer.resize(size_t(count) + size_t(count) / 4);
        st = phyx_b200_sweep_pairs(device->ctx, pairBuffer.data(), int64_t(pairBuffer.size()), &count, &device->lastBroadphase);
    }
    if (st != 0) phyx_host::fail("phyx_b200_sweep_pairs", st);
    for (int64_t k = 0; k < count; ++k)
    {
        std::pair<unsigned, unsigned> key(unsigned(pairBuffer[k].body1Index), unsigned(pairBuffer[k].body2Index));
        if (manifoldMap.insert(key)) manifolds.push_back(Manifold(int(key.first), int(key.second), manifolds.size * kMaxContactPoints));
    }
}

NOINLINE void Collider::UpdateManifolds(WorkQueue&, RigidBody* bodies)
{
    StageTimer t(&device->stageMs[3]);
    contactPoints.resize_copy(manifolds.size * kMaxContactPoints);
    for (int i = 0; i < manifolds.size; ++i) update_manifold(manifolds.data[i], bodies, contactPoints.data + manifolds.data[i].pointIndex);
}

// reference PackManifolds, src/Collider.cpp:379-416
NOINLINE void Collider::PackManifolds(RigidBody* bodies)
{
    StageTimer t(&device->stageMs[4]);
    for (int i = 0; i < manifolds.size;)
    {
        Manifold& m = manifolds.data[i];
        if (m.pointCount == 0 && !bodies[m.body1Index].geom.aabb.Intersects(bodies[m.body2Index].geom.aabb))
        {
            manifoldMap.erase(std::make_pair(unsigned(m.body1Index), unsigned(m.body2Index)));
            const Manifold& last = manifolds.data[manifolds.size - 1];
            const int slot = m.pointIndex;
            for (int k = 0; k < last.pointCount; ++k) contactPoints.data[slot + k] = contactPoints.data[last.pointIndex + k];
            m = last;
            m.pointIndex = slot;
            manifolds.size--;
        }
        else
            ++i;
    }
    contactPoints.truncate(manifolds.size * kMaxContactPoints);
}

// ==== Solver ===========================================================================================

Solver::Solver() : islandCount(0), islandMaxSize(0) {}

NOINLINE void Solver::SolveJoints(WorkQueue&, RigidBody* bodies, int bodiesCount, ContactPoint* contactPoints, const Configuration& configuration)
{
    StageTimer t(&device->stageMs[6]);
    if (!device->inUpdate) device->upload(bodies, bodiesCount);
    phyx_b200_solve_config cfg;
    cfg.contactIterationsCount = configuration.contactIterationsCount;
    cfg.penetrationIterationsCount = configuration.penetrationIterationsCount;
    switch (configuration.solveMode)
    {
    case Configuration::Solve_Scalar: cfg.schedule = PHYX_B200_SCHEDULE_REPLAY_SCALAR; break;
    case Configuration::Solve_SSE2: cfg.schedule = PHYX_B200_SCHEDULE_REPLAY_SSE2; break;
    case Configuration::Solve_AVX2: cfg.schedule = PHYX_B200_SCHEDULE_REPLAY_AVX2; break;
    default: cfg.schedule = PHYX_B200_SCHEDULE_COLOUR; break;
    }
    cfg.flags = solveFlags;
    PHYX_CALL(phyx_b200_solve_joints(device->ctx, reinterpret_cast<phyx_contact_joint*>(contactJoints.data), contactJoints.size,
        reinterpret_cast<const phyx_contact_point*>(contactPoints), contactPointCount, &cfg, &device->lastSolve));
    islandCount = 1;                       // Island_Single bookkeeping (reference src/Solver.cpp:108-109)
    islandMaxSize = contactJoints.size;
    if (!device->inUpdate) device->download(bodies, bodiesCount);
}

// ==== World ============================================================================================

World::World() : collisionTime(0), mergeTime(0), solveTime(0), gravity(0)
{
    collider.device = &device;
    solver.device = &device;
}

World::~World() {}

RigidBody* World::AddBody(Coords2f coords, Vector2f size)
{
    RigidBody fresh(coords, size, 1e-5f);
    fresh.index = bodies.size;
    bodies.push_back(fresh);
    device.resident = false;
    return &bodies.data[bodies.size - 1];
}

void World::Update(WorkQueue& queue, float dt, const Configuration& configuration)
{
    collisionTime = mergeTime = solveTime = 0;
    // The host AoS is the source of truth at entry: the caller edits bodies between steps (statics
    // after AddBody, accelerations every frame: reference src/main.cpp:91-93,337-346).
    device.upload(bodies.data, bodies.size);
    device.inUpdate = true;

    IntegrateVelocity(queue, dt);

    collider.UpdateBroadphase(bodies.data, bodies.size);
    collider.UpdatePairs(queue, bodies.data, bodies.size);
    collider.UpdateManifolds(queue, bodies.data);
    collider.PackManifolds(bodies.data);

    RefreshContactJoints();

    solver.contactPointCount = collider.contactPoints.size;
    solver.SolveJoints(queue, bodies.data, bodies.size, collider.contactPoints.data, configuration);

    IntegratePosition(queue, dt);

    device.inUpdate = false;
    device.download(bodies.data, bodies.size);
}

NOINLINE void World::IntegrateVelocity(WorkQueue&, float dt)
{
    StageTimer t(&device.stageMs[0]);
    if (!device.inUpdate) device.upload(bodies.data, bodies.size);
    PHYX_CALL(phyx_b200_integrate_velocity(device.ctx, dt, gravity));
    if (!device.inUpdate) device.download(bodies.data, bodies.size);
}

NOINLINE void World::IntegratePosition(WorkQueue&, float dt)
{
    StageTimer t(&device.stageMs[7]);
    if (!device.inUpdate) device.upload(bodies.data, bodies.size);
    PHYX_CALL(phyx_b200_integrate_position(device.ctx, dt));
    if (!device.inUpdate) device.download(bodies.data, bodies.size);
}

// reference World::RefreshContactJoints, src/World.cpp:72-149: joints persist across steps as the
// warm-start cache; contact points carry the index of their joint
NOINLINE void World::RefreshContactJoints()
{
    StageTimer t(&device.stageMs[5]);
    AlignedArray<ContactJoint>& joints = solver.contactJoints;
    for (int j = 0; j < joints.size; ++j) joints.data[j].contactPointIndex = -1;

    for (int mi = 0; mi < collider.manifolds.size; ++mi)
    {
        const Manifold& man = collider.manifolds.data[mi];
        for (int k = 0; k < man.pointCount; ++k)
        {
            int cpIndex = man.pointIndex + k;
            ContactPoint& cp = collider.contactPoints.data[cpIndex];
            if (cp.solverIndex < 0)
            {
                cp.solverIndex = joints.size;
                joints.push_back(ContactJoint(man.body1Index, man.body2Index, cpIndex));
            }
            else
            {
                ContactJoint& joint = joints.data[cp.solverIndex];
                assert(joint.body1Index == man.body1Index && joint.body2Index == man.body2Index);
                joint.contactPointIndex = cpIndex;
            }
        }
    }

    for (int j = 0; j < joints.size;)
    {
        ContactJoint& joint = joints.data[j];
        if (joint.contactPointIndex < 0)
        {
            joint = joints.data[joints.size - 1];
            joints.size--;
        }
        else
        {
            collider.contactPoints.data[joint.contactPointIndex].solverIndex = j;
            ++j;
        }
    }
}
