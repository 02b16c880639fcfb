// phyx_b200 host mirror — implementation (see phyx_host.h).
//
// Every stage of World::Update (reference src/World.cpp:19-37) is a call into the C ABI; the whole
// step runs on the device without leaving HBM:
//   1 IntegrateVelocity      -> phyx_b200_integrate_velocity
//   2 UpdateBroadphase       -> phyx_b200_update_broadphase
//   3 UpdatePairs            -> phyx_b200_update_pairs
//   4 UpdateManifolds        -> phyx_b200_update_manifolds
//   5 PackManifolds          -> phyx_b200_pack_manifolds
//   6 RefreshContactJoints   -> phyx_b200_refresh_contact_joints
//   7 SolveJoints            -> phyx_b200_solve_resident
//   8 IntegratePosition      -> phyx_b200_integrate_position
// There is no host implementation of any stage: this file only moves the caller-visible arrays
// (`bodies` in, `bodies` + the collider / solver mirrors out) and keeps their sizes in step.
//
// Synchronisation rules
//   * `World::bodies` is the source of truth at Update entry (the caller edits it between steps:
//     statics after AddBody, accelerations every frame, reference src/main.cpp:91-93,337-346); it
//     is uploaded once per Update and downloaded once at the end.
//   * The manifold cache and the joint cache live on the device.  `collider.manifolds`,
//     `collider.contactPoints` and `solver.contactJoints` are read-only mirrors: their sizes are
//     always current, their contents are refreshed after every Update / stage call while
//     `World::mirrorCollider` is true (default; the demo reads them for rendering and the HUD).
//     Clearing them (resetWorld, reference src/main.cpp:86-89) resets the device caches.
//   * Calling the public stage functions one by one (as the reference allows) works: each call
//     uploads `bodies` if it consumes body state the caller may have edited, runs on the device and
//     refreshes the mirrors it changes.
#include "phyx_host.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
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
    if (hostStale && resident && residentCount == count) download(bodies, count);   // never push a stale host copy over newer device state
    // keep World::bodies page-locked (it is copied both ways every Update); re-pin when it was reallocated
    if (pinBodies && (bodies != pinnedPtr || size_t(count) * sizeof(RigidBody) > pinnedBytes))
    {
        if (pinnedPtr) phyx_b200_host_unregister(ctx, pinnedPtr);
        pinnedPtr = nullptr;
        pinnedBytes = 0;
        if (count > 0 && phyx_b200_host_register(ctx, bodies, size_t(count) * sizeof(RigidBody)) == 0)
        {
            pinnedPtr = bodies;
            pinnedBytes = size_t(count) * sizeof(RigidBody);
        }
    }
    // inside World::Update with the array page-locked, nobody touches it before the step's own read-back: no need to wait
    if (inUpdateUpload && pinnedPtr == bodies)
        PHYX_CALL(phyx_b200_upload_bodies_async(ctx, reinterpret_cast<const phyx_rigid_body*>(bodies), count));
    else
        PHYX_CALL(phyx_b200_upload_bodies(ctx, reinterpret_cast<const phyx_rigid_body*>(bodies), count));
    resident = true;
    residentCount = count;
}

void Device::download(RigidBody* bodies, int count)
{
    PHYX_CALL(phyx_b200_download_bodies(ctx, reinterpret_cast<phyx_rigid_body*>(bodies), count));
    hostStale = false;
}

void Device::unpin(void* buffer)
{
    if (buffer && buffer == pinnedPtr)
    {
        if (ctx) phyx_b200_host_unregister(ctx, pinnedPtr);
        pinnedPtr = nullptr;
        pinnedBytes = 0;
    }
}

Device::~Device()
{
    if (ctx && pinnedPtr) phyx_b200_host_unregister(ctx, pinnedPtr);
    if (ctx) phyx_b200_destroy(ctx);
}

// Bring the caller-visible collider / solver arrays in line with the device caches.
void Device::mirror(Collider& collider, Solver& solver, bool contents)
{
    int32_t m = 0, p = 0, j = 0;
    PHYX_CALL(phyx_b200_collider_counts(ctx, &m, &p, &j));
    collider.manifolds.resize(m);
    collider.contactPoints.resize(p);
    solver.contactJoints.resize(j);
    if (contents)
    {
        PHYX_CALL(phyx_b200_download_manifolds(ctx, reinterpret_cast<phyx_manifold*>(collider.manifolds.data), m));
        PHYX_CALL(phyx_b200_download_contact_points(ctx, reinterpret_cast<phyx_contact_point*>(collider.contactPoints.data), p));
        PHYX_CALL(phyx_b200_download_joints(ctx, reinterpret_cast<phyx_contact_joint*>(solver.contactJoints.data), j));
    }
    mirroredManifolds = m;
    mirroredJoints = j;
}

// resetWorld() empties the host arrays (reference src/main.cpp:86-89): follow it on the device
void Device::followReset(Collider& collider, Solver& solver)
{
    ensure();
    if ((mirroredManifolds > 0 || mirroredJoints > 0) && collider.manifolds.size == 0 && solver.contactJoints.size == 0)
    {
        PHYX_CALL(phyx_b200_reset_collider(ctx));
        collider.contactPoints.resize(0);
        mirroredManifolds = mirroredJoints = 0;
    }
}

} // namespace phyx_host

using phyx_host::Device;
using phyx_host::StageTimer;

unsigned int WorkQueue::getIdealWorkerCount() { return std::max(std::thread::hardware_concurrency(), 1u); }

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

NOINLINE void Collider::UpdatePairs(WorkQueue&, RigidBody*, size_t)
{
    StageTimer t(&device->stageMs[2]);
    if (!device->inUpdate) device->followReset(*this, *solver);
    PHYX_CALL(phyx_b200_update_pairs(device->ctx, &device->lastBroadphase));
    if (!device->inUpdate) device->mirror(*this, *solver, mirrorContents);
}

NOINLINE void Collider::UpdateManifolds(WorkQueue&, RigidBody*)
{
    StageTimer t(&device->stageMs[3]);
    device->ensure();
    PHYX_CALL(phyx_b200_update_manifolds(device->ctx));
    if (!device->inUpdate) device->mirror(*this, *solver, mirrorContents);
}

NOINLINE void Collider::PackManifolds(RigidBody*)
{
    StageTimer t(&device->stageMs[4]);
    device->ensure();
    PHYX_CALL(phyx_b200_pack_manifolds(device->ctx));
    if (!device->inUpdate) device->mirror(*this, *solver, mirrorContents);
}

// ==== Solver ===========================================================================================

Solver::Solver() : islandCount(0), islandMaxSize(0) {}

NOINLINE void Solver::SolveJoints(WorkQueue&, RigidBody* bodies, int bodiesCount, ContactPoint*, const Configuration& configuration)
{
    StageTimer t(&device->stageMs[6]);
    if (!device->inUpdate) device->upload(bodies, bodiesCount);
    phyx_b200_solve_config cfg = MakeConfig(configuration);
    PHYX_CALL(phyx_b200_solve_resident(device->ctx, &cfg, &device->lastSolve));
    FillIslandCounters(configuration);
    if (!device->inUpdate)
    {
        device->download(bodies, bodiesCount);
        device->mirror(*collider, *this, collider->mirrorContents);
    }
}

phyx_b200_solve_config Solver::MakeConfig(const Configuration& configuration) const
{
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
    return cfg;
}

void Solver::FillIslandCounters(const Configuration& configuration)
{
    if (configuration.islandMode == Configuration::Island_Multiple || configuration.islandMode == Configuration::Island_MultipleSloppy)
    {
        // GatherIslands (reference src/Solver.cpp:285-454) on the device: the counters the demo's HUD reads.  The solve itself
        // relaxes all islands at once whatever the island mode (a colour spans the world).
        int32_t groups = 0, largest = 0;
        PHYX_CALL(phyx_b200_build_islands(device->ctx, &groups, &largest, nullptr));
        islandCount = groups;
        islandMaxSize = largest;
    }
    else
    {
        islandCount = 1;                   // Island_Single bookkeeping (reference src/Solver.cpp:108-109)
        islandMaxSize = device->lastSolve.joints;
    }
}

// ==== World ============================================================================================

World::World() : collisionTime(0), mergeTime(0), solveTime(0), gravity(0)
{
    // World::bodies is page-locked in place: whatever makes it reallocate must undo that first (AlignedArray::releaseHook)
    bodies.releaseHook = [](void* buffer, void* user) { static_cast<phyx_host::Device*>(user)->unpin(buffer); };
    bodies.releaseUser = &device;
    collider.device = &device;
    collider.solver = &solver;
    solver.device = &device;
    solver.collider = &collider;
}

World::~World()
{
    // members are destroyed in reverse order (device before bodies): undo the page-lock while the context is still alive
    device.unpin(bodies.data);
    bodies.releaseHook = nullptr;
}

RigidBody* World::AddBody(Coords2f coords, Vector2f size)
{
    SyncBodies();
    RigidBody fresh(coords, size, 1e-5f);
    fresh.index = bodies.size;
    bodies.push_back(fresh);
    device.resident = false;
    device.hostEdited = true;
    return &bodies.data[bodies.size - 1];
}

void World::SyncBodies()
{
    if (device.hostStale && device.ctx && device.resident && device.residentCount == int(bodies.size)) device.download(bodies.data, bodies.size);
    device.hostStale = false;
}

RigidBody* World::EditBodies()
{
    SyncBodies();
    device.hostEdited = true;
    return bodies.data;
}

void World::Update(WorkQueue& queue, float dt, const Configuration& configuration)
{
    collisionTime = mergeTime = solveTime = 0;
    device.followReset(collider, solver);
    // the reference's semantics: World::bodies is the source of truth at entry.  With the opt-in lazyBodies contract
    // it only is when the caller said so.
    if (!device.lazyBodies || device.hostEdited || !device.resident || device.residentCount != int(bodies.size))
    {
        device.inUpdateUpload = device.fusedUpdate && !collider.mirrorBroadphase;   // (the fused step ends with a wait)
        device.upload(bodies.data, bodies.size);
        device.inUpdateUpload = false;
    }
    device.hostEdited = false;
    device.inUpdate = true;

    if (device.fusedUpdate && !collider.mirrorBroadphase)
    {
        // the eight stages as ONE call into the C ABI (phyx_b200_world_step): same results, no read-back per stage
        StageTimer t(&device.stepMs);
        phyx_b200_solve_config cfg = solver.MakeConfig(configuration);
        phyx_b200_step_info info;
        PHYX_CALL(phyx_b200_world_step(device.ctx, dt, gravity, &cfg, &device.lastSolve, &device.lastBroadphase, &info));
        createdJoints = info.deferred ? info.jointsCreated : createdJoints;
        deletedJoints = info.deferred ? info.jointsDeleted : deletedJoints;
        matchedJoints = info.deferred ? info.joints - info.jointsCreated : matchedJoints;
        device.lastStepDeferred = info.deferred != 0;
        solver.FillIslandCounters(configuration);
    }
    else
    {
        IntegrateVelocity(queue, dt);

        collider.UpdateBroadphase(bodies.data, bodies.size);
        collider.UpdatePairs(queue, bodies.data, bodies.size);
        collider.UpdateManifolds(queue, bodies.data);
        collider.PackManifolds(bodies.data);

        RefreshContactJoints();

        solver.SolveJoints(queue, bodies.data, bodies.size, collider.contactPoints.data, configuration);

        IntegratePosition(queue, dt);
    }

    device.inUpdate = false;
    {
        StageTimer t(&device.syncMs);
        if (device.lazyBodies)
            device.hostStale = true;
        else
            device.download(bodies.data, bodies.size);
        device.mirror(collider, solver, collider.mirrorContents);
    }
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

NOINLINE void World::RefreshContactJoints()
{
    StageTimer t(&device.stageMs[5]);
    device.ensure();
    PHYX_CALL(phyx_b200_refresh_contact_joints(device.ctx, &matchedJoints, &createdJoints, &deletedJoints));
    if (!device.inUpdate) device.mirror(collider, solver, collider.mirrorContents);
}
