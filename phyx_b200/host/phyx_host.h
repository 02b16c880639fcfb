// phyx_b200 host mirror — the reference's C++ surface (World / Collider / Solver / Configuration /
// RigidBody / AlignedArray / WorkQueue: same type names, same public members, same call
// signatures) re-implemented on top of the C ABI in include/phyx_b200.h.
//
// A program written against the reference's headers (reference src/main.cpp is the model:
// world.AddBody(...), world.gravity, world.Update(queue, dt, config), world.bodies[i].coords,
// world.collider.manifolds, world.solver.contactJoints ...) compiles against these headers
// unchanged and gets every stage of the step — integration, sweep & prune broadphase, narrowphase,
// manifold and joint caches, contact solve — executed by the sm_100a kernels.  Nothing is computed
// on the host: the classes below only keep the caller-visible arrays in step with the device.
//
// Records keep the reference's exact memory layout (static_asserts below): they are handed to the
// C ABI as they are.
#pragma once

#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <unordered_set>
#include <utility>
#include <vector>

#include "phyx_b200.h"

#ifdef _MSC_VER
#define NOINLINE __declspec(noinline)
#else
#define NOINLINE __attribute__((noinline))
#endif

// ---- value types (reference src/Vector2.h, Coords2.h, AABB2.h: only what the API needs) ---------

template <typename T> struct Vector2
{
    T x, y;

    Vector2() {}
    Vector2(T x_, T y_) : x(x_), y(y_) {}

    T SquareLen() const { return x * x + y * y; }
    T Len() const { return std::sqrt(x * x + y * y); }
    void Invert() { x = -x; y = -y; }
    Vector2 GetPerpendicular() const { return Vector2(-y, x); }

    Vector2 operator-() const { return Vector2(-x, -y); }
    Vector2 operator+(const Vector2& o) const { return Vector2(x + o.x, y + o.y); }
    Vector2 operator-(const Vector2& o) const { return Vector2(x - o.x, y - o.y); }
    Vector2 operator*(T s) const { return Vector2(x * s, y * s); }
    T operator*(const Vector2& o) const { return x * o.x + y * o.y; }   // dot product, as in the reference
    Vector2& operator+=(const Vector2& o) { x += o.x; y += o.y; return *this; }
    Vector2& operator-=(const Vector2& o) { x -= o.x; y -= o.y; return *this; }
    Vector2& operator*=(T s) { x *= s; y *= s; return *this; }

    // reference Vector2::Rotate (src/Vector2.h:48-56): cos/sin in double, narrowed to T
    void Rotate(T angle)
    {
        T c = T(std::cos(double(angle))), s = T(std::sin(double(angle)));
        Vector2 px(-y, x);
        Vector2 d = (*this) * c + px * s - (*this);
        *this += d;
    }

    static Vector2 zero() { return Vector2(0, 0); }
    static Vector2 zeroVector() { return Vector2(0, 0); }
};

template <typename T> inline Vector2<T> operator*(T s, const Vector2<T>& v) { return Vector2<T>(v.x * s, v.y * s); }
template <typename T> inline T operator^(const Vector2<T>& a, const Vector2<T>& b) { return a.x * b.y - a.y * b.x; }

typedef Vector2<float> Vector2f;

template <typename T> struct Coords2
{
    Vector2<T> xVector, yVector;
    Vector2<T> pos;

    Coords2() {}
    // reference Coords2(pos, angle) (src/Coords2.h:10-17): pi truncated to 3.141592f, trig in double
    Coords2(const Vector2<T>& p, T angle)
    {
        float pi = 3.141592f;
        T quarter = angle + T(pi) / T(2.0);
        xVector = Vector2<T>(T(std::cos(double(angle))), T(std::sin(double(angle))));
        yVector = Vector2<T>(T(std::cos(double(quarter))), T(std::sin(double(quarter))));
        pos = p;
    }
    void Rotate(T angle)
    {
        xVector.Rotate(angle);
        yVector.Rotate(angle);
    }
};

typedef Coords2<float> Coords2f;

template <typename T> struct AABB2
{
    Vector2<T> boxPoint1, boxPoint2;

    AABB2() : boxPoint1(0, 0), boxPoint2(0, 0) {}
    void Set(const Vector2<T>& a, const Vector2<T>& b) { boxPoint1 = a; boxPoint2 = b; }
    bool Intersects(const AABB2& o) const
    {
        if (boxPoint1.x > o.boxPoint2.x || o.boxPoint1.x > boxPoint2.x) return false;
        if (boxPoint1.y > o.boxPoint2.y || o.boxPoint1.y > boxPoint2.y) return false;
        return true;
    }
};

typedef AABB2<float> AABB2f;

// reference src/Geom.h (box: half extents + frame + cached AABB)
struct Geom
{
    Vector2f size;
    Coords2f coords;
    AABB2f aabb;

    void RecomputeAABB()
    {
        Vector2f d(std::fabs(coords.xVector.x) * size.x + std::fabs(coords.yVector.x) * size.y,
            std::fabs(coords.xVector.y) * size.x + std::fabs(coords.yVector.y) * size.y);
        aabb.Set(coords.pos - d, coords.pos + d);
    }
};

// reference src/RigidBody.h:12-58
struct RigidBody
{
    RigidBody() {}
    RigidBody(Coords2f c, Vector2f size, float density)
    {
        coords = c;
        displacingVelocity = Vector2f(0.f, 0.f);
        displacingAngularVelocity = 0.f;
        acceleration = Vector2f(0.f, 0.f);
        angularAcceleration = 0.f;
        velocity = Vector2f(0.f, 0.f);
        angularVelocity = 0.f;
        geom.size = size;
        float mass = density * (size.x * size.y);
        float inertia = mass * (size.x * size.x + size.y * size.y);
        invMass = 1.0f / mass;
        invInertia = 1.0f / inertia;
        lastIteration = lastDisplacementIteration = 0;
        UpdateGeom();
    }
    void UpdateGeom()
    {
        geom.coords = coords;
        geom.RecomputeAABB();
    }

    unsigned int index;
    Geom geom;
    Vector2f velocity, acceleration;
    Vector2f displacingVelocity;
    float angularVelocity, angularAcceleration;
    float displacingAngularVelocity;
    float invMass, invInertia;
    Coords2f coords;
    int lastIteration;
    int lastDisplacementIteration;
};

static const int kMaxContactPoints = 2;

// reference src/Manifold.h:12-43
struct ContactPoint
{
    ContactPoint() {}
    ContactPoint(const Vector2f& p1, const Vector2f& p2, const Vector2f& n, const RigidBody* b1, const RigidBody* b2)
    {
        delta1 = p1 - b1->coords.pos;
        delta2 = p2 - b2->coords.pos;
        normal = n;
        isMerged = 0;
        isNewlyCreated = 1;
        pad_[0] = pad_[1] = 0;
        solverIndex = -1;
    }
    bool Equals(const ContactPoint& o, float tolerance) const
    {
        float t2 = tolerance * tolerance;
        return !(((o.delta1 - delta1).SquareLen() > t2) && ((o.delta2 - delta2).SquareLen() > t2));
    }
    Vector2f delta1, delta2;
    Vector2f normal;
    bool isMerged;
    bool isNewlyCreated;
    unsigned char pad_[2];
    int solverIndex;
};

// reference src/Manifold.h:45-68
struct Manifold
{
    Manifold() : body1Index(-1), body2Index(-1), pointCount(0), pointIndex(0) {}
    Manifold(int b1, int b2, int firstPoint) : body1Index(b1), body2Index(b2), pointCount(0), pointIndex(firstPoint) {}
    int body1Index, body2Index;
    int pointCount, pointIndex;
};

// reference src/Joints.h:6-23
struct ContactJoint
{
    ContactJoint() {}
    ContactJoint(int b1, int b2, int collisionIndex)
        : contactPointIndex(collisionIndex), body1Index(b1), body2Index(b2), normalLimiter_accumulatedImpulse(0.f), frictionLimiter_accumulatedImpulse(0.f)
    {
    }
    int contactPointIndex;
    int body1Index;
    int body2Index;
    float normalLimiter_accumulatedImpulse;
    float frictionLimiter_accumulatedImpulse;
};

static_assert(sizeof(RigidBody) == sizeof(phyx_rigid_body), "RigidBody layout");
static_assert(offsetof(RigidBody, velocity) == offsetof(phyx_rigid_body, velocity), "RigidBody layout");
static_assert(offsetof(RigidBody, coords) == offsetof(phyx_rigid_body, xVector), "RigidBody layout");
static_assert(sizeof(ContactPoint) == sizeof(phyx_contact_point), "ContactPoint layout");
static_assert(sizeof(Manifold) == sizeof(phyx_manifold), "Manifold layout");
static_assert(sizeof(ContactJoint) == sizeof(phyx_contact_joint), "ContactJoint layout");

// ---- containers / queue (reference src/base/AlignedArray.h, WorkQueue.h) -------------------------

// 32-byte aligned POD vector with the reference's public members (data, size, capacity) and calls.
template <typename T> struct AlignedArray
{
    T* data;
    int size;
    int capacity;
    // Called with (data, releaseUser) right before a buffer is given back to the allocator.  World::bodies is page-locked
    // in place (cudaHostRegister) for full-rate PCIe copies: CUDA requires the range to be unregistered before it is freed,
    // whichever call makes the array grow (AddBody, push_back, resize, move assignment, destruction).
    void (*releaseHook)(void* buffer, void* user) = nullptr;
    void* releaseUser = nullptr;

    AlignedArray() : data(nullptr), size(0), capacity(0) {}
    ~AlignedArray() { release(); }
    AlignedArray(const AlignedArray&) = delete;
    AlignedArray& operator=(const AlignedArray&) = delete;
    AlignedArray(AlignedArray&& o) : data(o.data), size(o.size), capacity(o.capacity) { o.data = nullptr; o.size = o.capacity = 0; }
    AlignedArray& operator=(AlignedArray&& o)
    {
        if (this != &o)
        {
            release();
            data = o.data; size = o.size; capacity = o.capacity;
            o.data = nullptr; o.size = o.capacity = 0;
        }
        return *this;
    }

    T* begin() { return data; }
    T* end() { return data + size; }
    T& operator[](int i) { assert(i >= 0 && i < size); return data[i]; }
    const T& operator[](int i) const { assert(i >= 0 && i < size); return data[i]; }

    void push_back(const T& v)
    {
        if (size == capacity)
        {
            T keep = v;   // v may live inside this array
            grow(size + 1, true);
            data[size++] = keep;
        }
        else
            data[size++] = v;
    }
    void truncate(int n) { assert(n <= size); size = n; }
    void clear() { size = 0; }
    void resize(int n) { if (n > capacity) grow(n, false); size = n; }
    void resize_copy(int n) { if (n > capacity) grow(n, true); size = n; }

  private:
    void release()
    {
        if (data && releaseHook) releaseHook(data, releaseUser);
        std::free(data);
        data = nullptr;
    }
    void grow(int need, bool keep)
    {
        int cap = capacity;
        while (cap < need) cap += cap / 2 + 1;
        size_t bytes = (size_t(cap) * sizeof(T) + 32 + 31) & ~size_t(31);   // tail pad as in the reference
        T* fresh = static_cast<T*>(std::aligned_alloc(32, bytes));
        if (data && keep) std::memcpy(fresh, data, size_t(size) * sizeof(T));
        release();
        data = fresh;
        capacity = cap;
    }
};

// The GPU grid replaces the reference's thread pool; the type stays because World::Update takes it.
class WorkQueue
{
  public:
    static unsigned int getIdealWorkerCount();
    explicit WorkQueue(unsigned int workerCount) : workers_(workerCount) {}
    unsigned int getWorkerCount() const { return workers_; }

  private:
    unsigned int workers_;
};

// reference src/Configuration.h, plus one value for the throughput schedule
struct Configuration
{
    enum SolveMode
    {
        Solve_Scalar,   // replay of the reference's scalar order         (bit-comparable with the reference)
        Solve_SSE2,     // replay of the reference's SIMD-4 order
        Solve_AVX2,     // replay of the reference's SIMD-8 order
        Solve_B200,     // device graph colouring (throughput mode)
    };
    enum IslandMode
    {
        Island_Single,
        Island_Multiple,
        Island_SingleSloppy,
        Island_MultipleSloppy
    };
    SolveMode solveMode;
    IslandMode islandMode;
    int contactIterationsCount;
    int penetrationIterationsCount;
};

// ---- device binding shared by World, Collider and Solver --------------------------------------------

struct Collider;
struct Solver;

namespace phyx_host
{
struct Device
{
    phyx_b200_ctx* ctx = nullptr;
    int deviceIndex = 0;
    bool resident = false;     // device SoA mirrors `bodies` (set by upload, cleared when the host edits)
    bool inUpdate = false;     // World::Update is driving: stage functions skip their own sync
    int residentCount = 0;
    double stageMs[8] = {};    // wall time per stage (reference scope taxonomy)
    double syncMs = 0;         // wall time of the end-of-Update download + mirrors
    double stepMs = 0;         // wall time of the fused step (phyx_b200_world_step) when World::Update takes it
    bool fusedUpdate = true;   // World::Update = one phyx_b200_world_step call (false: the eight stage calls, same results)
    bool lastStepDeferred = false;
    bool inUpdateUpload = false;   // this upload is World::Update's own (may skip the wait: phyx_b200_upload_bodies_async)
    bool pinBodies = true;     // page-lock World::bodies in place for full-rate PCIe copies
    void* pinnedPtr = nullptr;
    size_t pinnedBytes = 0;
    int mirroredManifolds = 0, mirroredJoints = 0;   // sizes last written into the host mirrors
    // Opt-in contract for World::bodies.  Off (default, the reference's semantics): the host array is
    // the source of truth at every Update entry and current at its exit: 2 x 128 B per body over PCIe per Update.
    // On: Update uploads only when the host copy was declared edited (World::EditBodies, AddBody) and never
    // downloads; World::SyncBodies brings World::bodies up to date when the caller wants to look at it.
    bool lazyBodies = false;
    bool hostEdited = true;    // the host copy has changes the device has not seen
    bool hostStale = false;    // the device is ahead of the host copy
    phyx_b200_solve_stats lastSolve = {};
    phyx_b200_broadphase_stats lastBroadphase = {};

    void ensure();                                   // create the context (aborts with a message on failure)
    void upload(RigidBody* bodies, int count);       // AoS -> HBM
    void download(RigidBody* bodies, int count);     // HBM -> AoS
    void unpin(void* buffer);                        // `buffer` is about to be freed: undo the page-lock if it is the pinned one
    void mirror(Collider& collider, Solver& solver, bool contents);   // device caches -> host arrays
    void followReset(Collider& collider, Solver& solver);             // host arrays cleared -> clear device caches
    ~Device();
};
[[noreturn]] void fail(const char* what, int status);
}

// ---- Collider (reference src/Collider.h) ----------------------------------------------------------------

struct Collider
{
    Collider();

    void UpdateBroadphase(RigidBody* bodies, size_t bodiesCount);
    void UpdatePairs(WorkQueue& queue, RigidBody* bodies, size_t bodiesCount);
    void UpdateManifolds(WorkQueue& queue, RigidBody* bodies);
    void PackManifolds(RigidBody* bodies);

    struct BroadphaseEntry
    {
        float minx, maxx;
        float centery, extenty;
        unsigned int index;
    };
    struct BroadphaseSortEntry
    {
        unsigned int value;
        unsigned int index;
    };

    // The pair cache itself lives on the device (pairset.cuh).  This member only exists because
    // callers clear it when they reset the world (reference src/main.cpp:88).
    struct PairSet
    {
        std::unordered_set<uint64_t> keys;
        static uint64_t key(unsigned a, unsigned b) { return (uint64_t(a) << 32) | b; }
        bool contains(const std::pair<unsigned, unsigned>& p) const { return keys.count(key(p.first, p.second)) != 0; }
        bool insert(const std::pair<unsigned, unsigned>& p) { return keys.insert(key(p.first, p.second)).second; }
        void erase(const std::pair<unsigned, unsigned>& p) { keys.erase(key(p.first, p.second)); }
        void clear() { keys.clear(); }
        size_t size() const { return keys.size(); }
    };

    PairSet manifoldMap;
    AlignedArray<Manifold> manifolds;
    AlignedArray<ContactPoint> contactPoints;
    AlignedArray<BroadphaseEntry> broadphase;          // filled on demand (see World::mirrorBroadphase)
    AlignedArray<BroadphaseSortEntry> broadphaseSort[2];

    phyx_host::Device* device = nullptr;
    Solver* solver = nullptr;
    bool mirrorBroadphase = false;   // also copy the sorted entries back into `broadphase` every step
    bool mirrorContents = true;      // refresh manifolds / contactPoints / contactJoints contents after every step
                                     // (sizes are always current); switch off when nothing on the host reads them
};

// ---- Solver (reference src/Solver.h) -------------------------------------------------------------------

struct Solver
{
    Solver();

    void SolveJoints(WorkQueue& queue, RigidBody* bodies, int bodiesCount, ContactPoint* contactPoints, const Configuration& configuration);

    int islandCount;
    int islandMaxSize;
    AlignedArray<ContactJoint> contactJoints;

    // (not in the reference) Configuration -> the C ABI's solve config; islandCount / islandMaxSize after a solve
    phyx_b200_solve_config MakeConfig(const Configuration& configuration) const;
    void FillIslandCounters(const Configuration& configuration);

    phyx_host::Device* device = nullptr;
    Collider* collider = nullptr;
    int solveFlags = 0;              // PHYX_B200_SOLVE_*
};

// ---- World (reference src/World.h) ---------------------------------------------------------------------

struct World
{
    World();
    ~World();

    RigidBody* AddBody(Coords2f coords, Vector2f size);

    void Update(WorkQueue& queue, float dt, const Configuration& configuration);

    NOINLINE void IntegrateVelocity(WorkQueue& queue, float dt);
    NOINLINE void IntegratePosition(WorkQueue& queue, float dt);
    NOINLINE void RefreshContactJoints();

    // --- additions (not in the reference), only meaningful with device.lazyBodies ---
    RigidBody* EditBodies();   // call BEFORE changing World::bodies between Updates: syncs the host copy, marks it edited
    void SyncBodies();         // bring World::bodies up to date with the device

    float collisionTime;
    float mergeTime;
    float solveTime;

    AlignedArray<RigidBody> bodies;
    Collider collider;
    Solver solver;

    float gravity;

    // --- additions (not in the reference) ---
    phyx_host::Device device;        // the HBM-resident state; select the GPU with device.deviceIndex before the first Update
    int matchedJoints = 0, createdJoints = 0, deletedJoints = 0;   // the reference's Matched/Created/Deleted counters (World.cpp:146-148)
};
