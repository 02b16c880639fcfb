// C wrapper around the host mirror's World so that Python tests and bench.py can drive the same
// C++ object a reference user would (World::AddBody / Update / the public stage functions).
#include "phyx_host.h"

#include <chrono>

namespace
{
struct Handle
{
    World world;
    WorkQueue queue;
    Handle(int device, int workers) : queue(workers) { world.device.deviceIndex = device; }
};

Configuration make_config(int solveMode, int islandMode, int contactIters, int penetrationIters)
{
    Configuration c;
    c.solveMode = Configuration::SolveMode(solveMode);
    c.islandMode = Configuration::IslandMode(islandMode);
    c.contactIterationsCount = contactIters;
    c.penetrationIterationsCount = penetrationIters;
    return c;
}
}

#define API extern "C" __attribute__((visibility("default")))

API void* phyxw_create(int device) { return new Handle(device, 0); }
API void phyxw_destroy(void* h) { delete static_cast<Handle*>(h); }
API void phyxw_set_gravity(void* h, float g) { static_cast<Handle*>(h)->world.gravity = g; }
API void phyxw_set_solve_flags(void* h, int flags) { static_cast<Handle*>(h)->world.solver.solveFlags = flags; }
API void phyxw_set_mirror_contents(void* h, int on) { static_cast<Handle*>(h)->world.collider.mirrorContents = on != 0; }
API double phyxw_get_sync_ms(void* h) { return static_cast<Handle*>(h)->world.device.syncMs; }
API void phyxw_reset_world(void* h)
{
    // what the demo's resetWorld does (reference src/main.cpp:86-89)
    World& w = static_cast<Handle*>(h)->world;
    w.bodies.clear();
    w.collider.manifolds.clear();
    w.collider.manifoldMap.clear();
    w.solver.contactJoints.clear();
}

API int phyxw_add_body(void* h, float x, float y, float angle, float sx, float sy, int is_static)
{
    World& w = static_cast<Handle*>(h)->world;
    RigidBody* b = w.AddBody(Coords2f(Vector2f(x, y), angle), Vector2f(sx, sy));
    if (is_static == 2)
        b->invMass = 0.f;   // the demo's platforms (reference src/main.cpp:172-173): immovable but free to rotate
    else if (is_static)
    {
        b->invMass = 0.f;
        b->invInertia = 0.f;
    }
    return int(b->index);
}

API void phyxw_add_bodies(void* h, const float* rows6, int count)
{
    for (int i = 0; i < count; ++i)
        phyxw_add_body(h, rows6[6 * i], rows6[6 * i + 1], rows6[6 * i + 2], rows6[6 * i + 3], rows6[6 * i + 4], int(rows6[6 * i + 5]));
}

API void phyxw_step(void* h, float dt, int solveMode, int islandMode, int contactIters, int penetrationIters)
{
    Handle* s = static_cast<Handle*>(h);
    s->world.Update(s->queue, dt, make_config(solveMode, islandMode, contactIters, penetrationIters));
}

// the eight public stage functions, individually (bit i of mask = stage i), as a reference user may call them
API void phyxw_step_staged(void* h, float dt, int solveMode, int islandMode, int contactIters, int penetrationIters, int mask)
{
    Handle* s = static_cast<Handle*>(h);
    World& w = s->world;
    Configuration c = make_config(solveMode, islandMode, contactIters, penetrationIters);
    if (mask & 1) w.IntegrateVelocity(s->queue, dt);
    if (mask & 2) w.collider.UpdateBroadphase(w.bodies.data, w.bodies.size);
    if (mask & 4) w.collider.UpdatePairs(s->queue, w.bodies.data, w.bodies.size);
    if (mask & 8) w.collider.UpdateManifolds(s->queue, w.bodies.data);
    if (mask & 16) w.collider.PackManifolds(w.bodies.data);
    if (mask & 32) w.RefreshContactJoints();
    if (mask & 64) w.solver.SolveJoints(s->queue, w.bodies.data, w.bodies.size, w.collider.contactPoints.data, c);
    if (mask & 128) w.IntegratePosition(s->queue, dt);
}

API int phyxw_body_count(void* h) { return static_cast<Handle*>(h)->world.bodies.size; }
API int phyxw_joint_count(void* h) { return static_cast<Handle*>(h)->world.solver.contactJoints.size; }
API int phyxw_manifold_count(void* h) { return static_cast<Handle*>(h)->world.collider.manifolds.size; }
API int phyxw_contact_point_count(void* h) { return static_cast<Handle*>(h)->world.collider.contactPoints.size; }
API int phyxw_broadphase_count(void* h) { return static_cast<Handle*>(h)->world.collider.broadphase.size; }

API void phyxw_get_bodies(void* h, void* out)
{
    World& w = static_cast<Handle*>(h)->world;
    w.SyncBodies();
    memcpy(out, w.bodies.data, size_t(w.bodies.size) * sizeof(RigidBody));
}
API void phyxw_set_bodies(void* h, const void* in, int count)
{
    World& w = static_cast<Handle*>(h)->world;
    w.bodies.resize(count);
    memcpy(w.bodies.data, in, size_t(count) * sizeof(RigidBody));   // the whole array is replaced: nothing to sync first
    w.device.hostStale = false;
    w.device.hostEdited = true;
}
API void phyxw_set_lazy_bodies(void* h, int on) { static_cast<Handle*>(h)->world.device.lazyBodies = on != 0; }
API void phyxw_sync_bodies(void* h) { static_cast<Handle*>(h)->world.SyncBodies(); }
API void phyxw_get_joints(void* h, void* out)
{
    Solver& s = static_cast<Handle*>(h)->world.solver;
    memcpy(out, s.contactJoints.data, size_t(s.contactJoints.size) * sizeof(ContactJoint));
}
API void phyxw_get_manifolds(void* h, void* out)
{
    Collider& c = static_cast<Handle*>(h)->world.collider;
    memcpy(out, c.manifolds.data, size_t(c.manifolds.size) * sizeof(Manifold));
}
API void phyxw_get_contact_points(void* h, void* out)
{
    Collider& c = static_cast<Handle*>(h)->world.collider;
    memcpy(out, c.contactPoints.data, size_t(c.contactPoints.size) * sizeof(ContactPoint));
}
API void phyxw_get_broadphase(void* h, void* out)
{
    Collider& c = static_cast<Handle*>(h)->world.collider;
    memcpy(out, c.broadphase.data, size_t(c.broadphase.size) * sizeof(Collider::BroadphaseEntry));
}
API void phyxw_get_stage_ms(void* h, double* out8) { memcpy(out8, static_cast<Handle*>(h)->world.device.stageMs, sizeof(double) * 8); }
API void phyxw_reset_stage_ms(void* h)
{
    memset(static_cast<Handle*>(h)->world.device.stageMs, 0, sizeof(double) * 8);
    static_cast<Handle*>(h)->world.device.syncMs = 0;
    static_cast<Handle*>(h)->world.device.stepMs = 0;
}
API double phyxw_get_step_ms(void* h) { return static_cast<Handle*>(h)->world.device.stepMs; }
API void phyxw_set_fused_update(void* h, int on) { static_cast<Handle*>(h)->world.device.fusedUpdate = on != 0; }
API int phyxw_last_step_deferred(void* h) { return static_cast<Handle*>(h)->world.device.lastStepDeferred ? 1 : 0; }
API void phyxw_get_solve_stats(void* h, phyx_b200_solve_stats* out) { *out = static_cast<Handle*>(h)->world.device.lastSolve; }
API void phyxw_get_broadphase_stats(void* h, phyx_b200_broadphase_stats* out) { *out = static_cast<Handle*>(h)->world.device.lastBroadphase; }
API void phyxw_get_island_counts(void* h, int* out2)
{
    const Solver& s = static_cast<Handle*>(h)->world.solver;
    out2[0] = s.islandCount;
    out2[1] = s.islandMaxSize;
}
API void* phyxw_context(void* h)
{
    Handle* s = static_cast<Handle*>(h);
    s->world.device.ensure();
    return s->world.device.ctx;
}
