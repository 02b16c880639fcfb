// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// C-ABI harness around the UNMODIFIED reference (zeux/phyx @ 327b6c96) so that tests, the
// golden-fixture generator and bench.py's `--impl reference` / `cpu_baseline` legs can drive the
// reference's own World/Collider/Solver from Python (ctypes).
//
// The reference sources are compiled WHERE THEY LIE (/root/reference/src, passed with -I and
// -DPHYX_REF_SRC); nothing from the reference is copied into this repository.  The build recipe
// is oracle/Makefile; outputs go to oracle/_ref/ only (git-ignored, but shipped to the GPU box).
//
// Only the reference's PUBLIC surface is used here (World / Collider / Solver members); the four
// reference translation units are compiled next to this file by the Makefile.
#include <cstdio>
#include <cstring>
#include <cassert>
#include <chrono>
#include <cstdint>
#include <thread>

#include "World.h"
#include "Configuration.h"
#include "base/WorkQueue.h"
#include "base/RadixSort.h"

static_assert(sizeof(RigidBody) == 128, "RigidBody layout");
static_assert(sizeof(ContactPoint) == 32, "ContactPoint layout");
static_assert(sizeof(Manifold) == 16, "Manifold layout");
static_assert(sizeof(ContactJoint) == 20, "ContactJoint layout");
static_assert(sizeof(Solver::SolveBody) == 16, "SolveBody layout");
static_assert(sizeof(Collider::BroadphaseEntry) == 20, "BroadphaseEntry layout");

namespace
{
struct RefWorld
{
    World world;
    WorkQueue* queue;
    double stage_ms[9]; // accumulated per-stage wall time (8 stages + PrepareIndices probe)
    long long pair_tests;

    explicit RefWorld(int workers) : queue(new WorkQueue(workers)), pair_tests(0)
    {
        memset(stage_ms, 0, sizeof(stage_ms));
    }
    ~RefWorld() { delete queue; }
};

double now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

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

extern "C" {

const char* ref_build_flavour()
{
#if defined(__FAST_MATH__)
    return "fast";
#else
    return "strict";
#endif
}

int ref_hardware_concurrency() { return int(std::thread::hardware_concurrency()); }

void* ref_world_create(int workers) { return new RefWorld(workers); }
void ref_world_destroy(void* h) { delete static_cast<RefWorld*>(h); }

void ref_set_gravity(void* h, float g) { static_cast<RefWorld*>(h)->world.gravity = g; }

// World::AddBody through the reference's own Coords2f(pos, angle) constructor.
int ref_add_body(void* h, float x, float y, float angle, float sx, float sy, int is_static)
{
    World& w = static_cast<RefWorld*>(h)->world;
    RigidBody* b = w.AddBody(Coords2f(Vector2f(x, y), angle), Vector2f(sx, sy));
    if (is_static == 2)
        b->invMass = 0.f;   // the demo's platforms (src/main.cpp:172-173): immovable but free to rotate
    else if (is_static)
    {
        b->invMass = 0.f;
        b->invInertia = 0.f;
    }
    return int(b->index);
}

void ref_add_bodies(void* h, const float* rows6, int count)
{
    for (int i = 0; i < count; ++i)
        ref_add_body(h, rows6[6 * i], rows6[6 * i + 1], rows6[6 * i + 2], rows6[6 * i + 3], rows6[6 * i + 4], int(rows6[6 * i + 5]));
}

int ref_body_count(void* h) { return static_cast<RefWorld*>(h)->world.bodies.size; }
int ref_joint_count(void* h) { return static_cast<RefWorld*>(h)->world.solver.contactJoints.size; }
int ref_manifold_count(void* h) { return static_cast<RefWorld*>(h)->world.collider.manifolds.size; }
int ref_contact_point_count(void* h) { return static_cast<RefWorld*>(h)->world.collider.contactPoints.size; }

void ref_get_bodies(void* h, void* out)
{
    World& w = static_cast<RefWorld*>(h)->world;
    memcpy(out, w.bodies.data, size_t(w.bodies.size) * sizeof(RigidBody));
}
void ref_set_bodies(void* h, const void* in, int n)
{
    World& w = static_cast<RefWorld*>(h)->world;
    w.bodies.resize(n);
    memcpy(w.bodies.data, in, size_t(n) * sizeof(RigidBody));
}
void ref_get_joints(void* h, void* out)
{
    Solver& s = static_cast<RefWorld*>(h)->world.solver;
    memcpy(out, s.contactJoints.data, size_t(s.contactJoints.size) * sizeof(ContactJoint));
}
void ref_get_manifolds(void* h, void* out)
{
    Collider& c = static_cast<RefWorld*>(h)->world.collider;
    memcpy(out, c.manifolds.data, size_t(c.manifolds.size) * sizeof(Manifold));
}
void ref_get_contact_points(void* h, void* out)
{
    Collider& c = static_cast<RefWorld*>(h)->world.collider;
    memcpy(out, c.contactPoints.data, size_t(c.contactPoints.size) * sizeof(ContactPoint));
}
int ref_get_broadphase(void* h, void* out)
{
    Collider& c = static_cast<RefWorld*>(h)->world.collider;
    if (out) memcpy(out, c.broadphase.data, size_t(c.broadphase.size) * sizeof(Collider::BroadphaseEntry));
    return c.broadphase.size;
}
// joint_index as left by the last SolveJoints (the order the SIMD loops ran in).
int ref_get_joint_index(void* h, int* out)
{
    Solver& s = static_cast<RefWorld*>(h)->world.solver;
    if (out) memcpy(out, s.joint_index.data, size_t(s.joint_index.size) * sizeof(int));
    return s.joint_index.size;
}

// One World::Update, exactly as the demo calls it.
void ref_step(void* h, float dt, int solveMode, int islandMode, int contactIters, int penetrationIters)
{
    RefWorld* r = static_cast<RefWorld*>(h);
    Configuration c = make_config(solveMode, islandMode, contactIters, penetrationIters);
    r->world.Update(*r->queue, dt, c);
}

// The same eight stages World::Update runs (World.cpp:25-36), called through the reference's
// public stage functions with a steady_clock around each; times accumulate in stage_ms[0..7].
// stage_mask selects which stages run (bit i = stage i) so tests can capture mid-step state.
void ref_step_staged(void* h, float dt, int solveMode, int islandMode, int contactIters, int penetrationIters, int stage_mask)
{
    RefWorld* r = static_cast<RefWorld*>(h);
    World& w = r->world;
    Configuration c = make_config(solveMode, islandMode, contactIters, penetrationIters);
    double t0;
#define STAGE(i, call)                 \
    if (stage_mask & (1 << (i)))       \
    {                                  \
        t0 = now_ms();                 \
        call;                          \
        r->stage_ms[i] += now_ms() - t0; \
    }
    STAGE(0, w.IntegrateVelocity(*r->queue, dt));
    STAGE(1, w.collider.UpdateBroadphase(w.bodies.data, w.bodies.size));
    // bit 8 of stage_mask: take UpdatePairsParallel even with 0 workers.  With WorkQueue(0) its
    // parallelFor runs inline on the caller, so the pair order is the serial sweep order, but it
    // goes through contains()+insert() and so avoids the DenseHash tombstone corruption that
    // UpdatePairsSerial trips on scenes with manifold churn.
    if (stage_mask & 0x100)
    {
        STAGE(2, w.collider.UpdatePairsParallel(*r->queue, w.bodies.data, w.bodies.size));
    }
    else
    {
        STAGE(2, w.collider.UpdatePairs(*r->queue, w.bodies.data, w.bodies.size));
    }
    STAGE(3, w.collider.UpdateManifolds(*r->queue, w.bodies.data));
    STAGE(4, w.collider.PackManifolds(w.bodies.data));
    STAGE(5, w.RefreshContactJoints());
    STAGE(6, w.solver.SolveJoints(*r->queue, w.bodies.data, w.bodies.size, w.collider.contactPoints.data, c));
    STAGE(7, w.IntegratePosition(*r->queue, dt));
#undef STAGE
}

void ref_get_stage_ms(void* h, double* out9) { memcpy(out9, static_cast<RefWorld*>(h)->stage_ms, sizeof(double) * 9); }
void ref_reset_stage_ms(void* h) { memset(static_cast<RefWorld*>(h)->stage_ms, 0, sizeof(double) * 9); }

// Count sweep tests / overlapping pairs of the current broadphase array (Collider.cpp:296-320 loop
// shape, without touching the manifold map): out[0] = tests (inner-loop iterations that passed the
// x-break), out[1] = pairs that also pass the y test.
void ref_count_sweep(void* h, long long* out2)
{
    Collider& c = static_cast<RefWorld*>(h)->world.collider;
    long long tests = 0, pairs = 0;
    int n = c.broadphase.size;
    for (int i = 0; i < n; ++i)
    {
        const Collider::BroadphaseEntry& a = c.broadphase.data[i];
        for (int j = i + 1; j < n; ++j)
        {
            const Collider::BroadphaseEntry& b = c.broadphase.data[j];
            if (b.minx > a.maxx) break;
            tests++;
            if (fabsf(b.centery - a.centery) <= a.extenty + b.extenty) pairs++;
        }
    }
    out2[0] = tests;
    out2[1] = pairs;
}

// Stand-alone probe of the reference's serial grouper on the CURRENT joints (Solver.cpp:217-273):
// fills joint_index with identity the way SolveJoints<N> does (Solver.cpp:95-106), runs
// PrepareIndices(0, J, N), returns groupOffset and (optionally) the resulting order.  Time is
// accumulated in stage_ms[8].
int ref_prepare_indices(void* h, int groupSize, int* out_index)
{
    RefWorld* r = static_cast<RefWorld*>(h);
    Solver& s = r->world.solver;
    int jointCount = s.contactJoints.size;
    int bodiesCount = r->world.bodies.size;
    s.joint_index.resize(jointCount);
    s.jointGroup_joints.resize(jointCount);
    s.jointGroup_bodies.resize(bodiesCount);
    for (int i = 0; i < jointCount; ++i) s.joint_index[i] = i;
    for (int i = 0; i < bodiesCount; ++i) s.jointGroup_bodies[i] = 0;
    double t0 = now_ms();
    int groupOffset = s.PrepareIndices(0, jointCount, groupSize);
    r->stage_ms[8] += now_ms() - t0;
    if (out_index) memcpy(out_index, s.joint_index.data, size_t(jointCount) * sizeof(int));
    return groupOffset;
}

// Solver::GatherIslands (src/Solver.cpp:285-454) on the CURRENT joints: per body its island (island_index of its root,
// -1 for static bodies) and its coalesced group (island_indexremap of that); out3 = {islands before coalescing,
// islandCount, islandMaxSize}.
void ref_gather_islands(void* h, int groupSize, int* out_island, int* out_group, int* out3)
{
    RefWorld* r = static_cast<RefWorld*>(h);
    Solver& s = r->world.solver;
    const int n = r->world.bodies.size;
    s.GatherIslands(r->world.bodies.data, n, groupSize);
    int islands = 0;
    for (int i = 0; i < n; ++i)
    {
        const int root = s.island_remap[i];
        const int island = root < 0 ? -1 : s.island_index[root];
        if (island + 1 > islands) islands = island + 1;
        if (out_island) out_island[i] = island;
        if (out_group) out_group[i] = island < 0 ? -1 : s.island_indexremap[island];
    }
    out3[0] = islands;
    out3[1] = s.islandCount;
    out3[2] = s.islandMaxSize;
}

// ---- function-level entry points on caller-provided arrays (captured-input parity) ----------

// Solver::SolveJoints on caller arrays: bodies (RigidBody AoS, in/out), joints (ContactJoint AoS,
// in/out: cached impulses), contact points (in).  Returns the joint order used via out_index
// (may be null).  Runs with WorkQueue(workers).
void ref_solve_joints(void* bodies, int bodiesCount, void* joints, int jointCount, const void* contactPoints,
    int contactPointCount, int solveMode, int islandMode, int contactIters, int penetrationIters, int workers, int* out_index)
{
    // the SIMD gathers use aligned 32-byte loads on ContactPoint rows: keep them in an AlignedArray
    AlignedArray<ContactPoint> cps;
    cps.resize(contactPointCount);
    memcpy(cps.data, contactPoints, size_t(contactPointCount) * sizeof(ContactPoint));
    WorkQueue queue(workers);
    Solver solver;
    solver.contactJoints.resize(jointCount);
    memcpy(solver.contactJoints.data, joints, size_t(jointCount) * sizeof(ContactJoint));
    Configuration c = make_config(solveMode, islandMode, contactIters, penetrationIters);
    solver.SolveJoints(queue, static_cast<RigidBody*>(bodies), bodiesCount, cps.data, c);
    memcpy(joints, solver.contactJoints.data, size_t(jointCount) * sizeof(ContactJoint));
    if (out_index) memcpy(out_index, solver.joint_index.data, size_t(jointCount) * sizeof(int));
}

// Collider::UpdateBroadphase on caller bodies; out = BroadphaseEntry[n] (20 B each).
void ref_update_broadphase(void* bodies, int bodiesCount, void* out_entries)
{
    Collider c;
    c.UpdateBroadphase(static_cast<RigidBody*>(bodies), size_t(bodiesCount));
    memcpy(out_entries, c.broadphase.data, size_t(bodiesCount) * sizeof(Collider::BroadphaseEntry));
}

// Collider::UpdateBroadphase + UpdatePairsSerial on caller bodies with an EMPTY manifold map:
// returns every overlapping pair in the reference's emission order as int pairs.
int ref_all_pairs(void* bodies, int bodiesCount, int* out_pairs, int capacity)
{
    WorkQueue queue(0);
    Collider c;
    c.UpdateBroadphase(static_cast<RigidBody*>(bodies), size_t(bodiesCount));
    c.UpdatePairs(queue, static_cast<RigidBody*>(bodies), size_t(bodiesCount));
    int n = c.manifolds.size;
    for (int i = 0; i < n && i < capacity; ++i)
    {
        out_pairs[2 * i + 0] = c.manifolds.data[i].body1Index;
        out_pairs[2 * i + 1] = c.manifolds.data[i].body2Index;
    }
    return n;
}

// World::IntegrateVelocity / IntegratePosition on caller bodies.
void ref_integrate_velocity(void* bodies, int bodiesCount, float dt, float gravity)
{
    WorkQueue queue(0);
    World w;
    w.gravity = gravity;
    w.bodies.resize(bodiesCount);
    memcpy(w.bodies.data, bodies, size_t(bodiesCount) * sizeof(RigidBody));
    w.IntegrateVelocity(queue, dt);
    memcpy(bodies, w.bodies.data, size_t(bodiesCount) * sizeof(RigidBody));
}
void ref_integrate_position(void* bodies, int bodiesCount, float dt)
{
    WorkQueue queue(0);
    World w;
    w.bodies.resize(bodiesCount);
    memcpy(w.bodies.data, bodies, size_t(bodiesCount) * sizeof(RigidBody));
    w.IntegratePosition(queue, dt);
    memcpy(bodies, w.bodies.data, size_t(bodiesCount) * sizeof(RigidBody));
}

// radixFloat / radixSort3 on caller {value,index} pairs (RadixSort.h:19-95): sorts in place.
void ref_radix_sort3(uint32_t* value_index_pairs, int count)
{
    typedef Collider::BroadphaseSortEntry E;
    AlignedArray<E> tmp;
    tmp.resize(count);
    E* src = reinterpret_cast<E*>(value_index_pairs);
    E* res = radixSort3(src, tmp.data, size_t(count), [](const E& e) { return e.value; });
    if (res != src) memcpy(src, res, size_t(count) * sizeof(E));
}
uint32_t ref_radix_float(float v) { return radixFloat(v); }

} // extern "C"
