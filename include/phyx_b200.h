/* phyx_b200 — C ABI of the B200 (sm_100a) hot path behind World::Update().
 *
 * The reference (zeux/phyx @ 327b6c96) has no plugin/FFI layer: its "operator API" is the C++
 * member surface World / Collider / Solver.  This header is the boundary a
 * drop-in replacement binds instead: every entry point replaces one reference stage and takes
 * the reference's own POD records (plain pointers and sizes, no C++ or torch types), so the host
 * mirror in phyx_b200/host/ (same class and member names as the reference) — or a patched copy
 * of the reference itself, see INTEGRATION.md — forwards its stage calls here.
 *
 * All functions return 0 on success, a non-zero phyx_b200_status otherwise; the message is
 * available from phyx_b200_last_error().  Nothing here falls back to the CPU: if no CUDA device
 * is present phyx_b200_create fails.
 *
 * One context = one World on one device.  A context is driven by one host thread at a time
 * (the reference's Update is not re-entrant either, World.cpp:19-37).
 */
#ifndef PHYX_B200_H
#define PHYX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define PHYX_B200_API __declspec(dllexport)
#else
#define PHYX_B200_API __attribute__((visibility("default")))
#endif

/* ---- records: the reference's own layouts, sizes pinned by static_assert in the library ------ */

typedef struct { float x, y; } phyx_vec2;

/* RigidBody, reference src/RigidBody.h:12-58 (128 B) */
typedef struct {
    uint32_t index;
    phyx_vec2 geom_size;
    phyx_vec2 geom_xVector, geom_yVector, geom_pos;
    phyx_vec2 aabb_min, aabb_max;
    phyx_vec2 velocity, acceleration, displacingVelocity;
    float angularVelocity, angularAcceleration, displacingAngularVelocity;
    float invMass, invInertia;
    phyx_vec2 xVector, yVector, pos;
    int32_t lastIteration, lastDisplacementIteration;
} phyx_rigid_body;

/* ContactJoint, reference src/Joints.h:6-23 (20 B) */
typedef struct {
    int32_t contactPointIndex, body1Index, body2Index;
    float normalLimiter_accumulatedImpulse, frictionLimiter_accumulatedImpulse;
} phyx_contact_joint;

/* ContactPoint, reference src/Manifold.h:12-43 (32 B) */
typedef struct {
    phyx_vec2 delta1, delta2, normal;
    uint8_t isMerged, isNewlyCreated, pad_[2];
    int32_t solverIndex;
} phyx_contact_point;

/* Manifold, reference src/Manifold.h:45-68 (16 B) */
typedef struct { int32_t body1Index, body2Index, pointCount, pointIndex; } phyx_manifold;

/* Collider::BroadphaseEntry, reference src/Collider.h:45-50 (20 B) */
typedef struct { float minx, maxx, centery, extenty; uint32_t index; } phyx_broadphase_entry;

typedef struct { int32_t body1Index, body2Index; } phyx_pair;

typedef enum {
    PHYX_B200_OK = 0,
    PHYX_B200_ERR_CUDA = 1,        /* a CUDA runtime call failed (message has the cudaError) */
    PHYX_B200_ERR_ARGUMENT = 2,    /* null pointer, negative size, index out of range */
    PHYX_B200_ERR_NO_DEVICE = 3,   /* no CUDA device / device is not sm_100 */
    PHYX_B200_ERR_CAPACITY = 4,    /* caller buffer too small (required size is reported) */
    PHYX_B200_ERR_STATE = 5        /* call order violated (e.g. sweep before update_broadphase) */
} phyx_b200_status;

/* How the joints are ordered for the sequential-impulse sweeps. */
typedef enum {
    /* Graph colouring on the device (the reference's independent-joint grouping, Solver.cpp:217-273,
     * widened from SIMD-8 groups to whole colours); per-joint skip rule.  Throughput mode. */
    PHYX_B200_SCHEDULE_COLOUR = 0,
    /* Replay of the reference's AVX2 order: PrepareIndices(N=8) groups + scalar tail, executed as
     * dependency levels, with the AVX2 8-lane skip rule.  Dependency-equivalent to
     * Solve_AVX2 / Island_Single, hence bit-comparable with it.  Parity mode. */
    PHYX_B200_SCHEDULE_REPLAY_AVX2 = 1,
    /* As above for Solve_SSE2 (N=4) and Solve_Scalar (N=1). */
    PHYX_B200_SCHEDULE_REPLAY_SSE2 = 2,
    PHYX_B200_SCHEDULE_REPLAY_SCALAR = 3
} phyx_b200_schedule;

typedef struct {
    int32_t contactIterationsCount;       /* Configuration.h:21 */
    int32_t penetrationIterationsCount;   /* Configuration.h:22 */
    int32_t schedule;                     /* phyx_b200_schedule */
    int32_t flags;                        /* PHYX_B200_SOLVE_* */
} phyx_b200_solve_config;

#define PHYX_B200_SOLVE_STATIC_DEPS 1     /* replay: also put joints that share a STATIC body on strictly increasing levels
                                             (serialises ground contacts; only useful to cross-check the wake passes) */
#define PHYX_B200_SOLVE_KEEP_SCHEDULE 2   /* reuse the schedule built by the previous solve call if the
                                             joint (body1,body2) list is unchanged */
#define PHYX_B200_SOLVE_HOST_COLOURING 4  /* COLOUR schedule: build it with the serial host greedy instead of
                                             the device kernel (cross-check only) */

typedef struct {
    int32_t joints, slots, levels;             /* schedule shape: slots >= joints (padding), levels = colours */
    int32_t contactIterationsRun, penetrationIterationsRun;  /* with the productive early-out */
    int32_t wakePasses;                        /* extra level passes run for static-body wake-ups (DESIGN.md "static bodies") */
    int32_t colourRounds;                      /* rounds the device colouring needed (0: host-built schedule) */
    int32_t kernelForm;                        /* iteration kernel that ran: 0 joint units (k_solve), 1 manifold units streaming
                                                  (k_solve_pairs), 2 manifold units record form (k_solve_pairs2 / partitioned),
                                                  3 strip-local (k_solve_strips, the default of the resident pipeline) */
    int64_t activeJointIterations[2];          /* joint-iterations actually relaxed (not skipped by the lastIteration
                                                  test) in the impulse / displacement loops */
    float ms_schedule, ms_refresh, ms_iterations, ms_finish, ms_total;   /* CUDA-event times */
    float ms_h2d, ms_d2h;
} phyx_b200_solve_stats;

typedef struct {
    int64_t tests;     /* sweep tests: j visited before the x-break (Collider.cpp:303-307) */
    int64_t pairs;     /* pairs that also pass the y test (hash lookups in the reference) */
    float ms_sort, ms_sweep, ms_total;
} phyx_b200_broadphase_stats;

typedef struct phyx_b200_ctx phyx_b200_ctx;

/* ---- lifetime ------------------------------------------------------------------------------- */
PHYX_B200_API int phyx_b200_create(int device, phyx_b200_ctx** out);
PHYX_B200_API void phyx_b200_destroy(phyx_b200_ctx* ctx);
PHYX_B200_API const char* phyx_b200_last_error(void);
PHYX_B200_API const char* phyx_b200_version(void);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
PHYX_B200_API int64_t phyx_b200_launch_count(const phyx_b200_ctx* ctx);
/* device (re)allocations made by this process's contexts so far and the host time spent in them; a steady-state
 * step makes none (buffers grow with 2x headroom), so a difference across a timed region flags an outlier */
PHYX_B200_API void phyx_b200_alloc_stats(int64_t* count, double* hostMs);
/* CUDA stream all work of this context is issued on (cudaStream_t as void*), for event timing */
PHYX_B200_API void* phyx_b200_stream(const phyx_b200_ctx* ctx);
PHYX_B200_API int phyx_b200_synchronize(phyx_b200_ctx* ctx);

/* ---- bodies: World::bodies (AlignedArray<RigidBody>, World.h:32) <-> SoA in HBM ---------------- */
/* Replaces nothing in the reference (there is no device boundary there); it is the H2D/D2H hop of
 * the drop-in.  upload converts the AoS records to the device SoA, download converts back and
 * fills every field the reference's stages write (velocity .. coords, geom.coords, geom.aabb). */
PHYX_B200_API int phyx_b200_upload_bodies(phyx_b200_ctx* ctx, const phyx_rigid_body* bodies, int count);
/* upload without waiting for the copy: `bodies` must be page-locked (phyx_b200_host_register) and stay untouched until the
 * next call that waits for the device (world_step, download_bodies, synchronize).  What World::Update of the host mirror
 * uses: the copy engine and the launch of the step then overlap the host's bookkeeping */
PHYX_B200_API int phyx_b200_upload_bodies_async(phyx_b200_ctx* ctx, const phyx_rigid_body* bodies, int count);
PHYX_B200_API int phyx_b200_download_bodies(phyx_b200_ctx* ctx, phyx_rigid_body* bodies, int count);
PHYX_B200_API int phyx_b200_body_count(const phyx_b200_ctx* ctx);
/* page-lock / release a caller buffer in place (e.g. World::bodies.data) so the copies above run at
 * full PCIe rate; optional */
PHYX_B200_API int phyx_b200_host_register(phyx_b200_ctx* ctx, void* ptr, size_t bytes);
PHYX_B200_API int phyx_b200_host_unregister(phyx_b200_ctx* ctx, void* ptr);

/* AABB of all dynamic bodies {min.x, min.y, max.x, max.y} ({+inf, +inf, -inf, -inf}-like extremes if there are none): what a
 * caller that shards a scene by island over several devices exchanges to check that the shards stay apart */
PHYX_B200_API int phyx_b200_dynamic_extent(phyx_b200_ctx* ctx, float* minMax4);

/* ---- World::IntegrateVelocity / IntegratePosition, reference src/World.cpp:39-70 -------------- */
PHYX_B200_API int phyx_b200_integrate_velocity(phyx_b200_ctx* ctx, float dt, float gravity);
PHYX_B200_API int phyx_b200_integrate_position(phyx_b200_ctx* ctx, float dt);

/* ---- Collider::UpdateBroadphase, reference src/Collider.cpp:251-284 --------------------------- */
/* key = radixFloat(aabb.min.x), stable 3-pass (11/11/10 bit) LSD radix sort, gather entries. */
PHYX_B200_API int phyx_b200_update_broadphase(phyx_b200_ctx* ctx);
/* Collider::broadphase (Collider.h:65) for host readers */
PHYX_B200_API int phyx_b200_download_broadphase(phyx_b200_ctx* ctx, phyx_broadphase_entry* entries, int capacity);

/* ---- the sweep of Collider::UpdatePairs*, reference src/Collider.cpp:296-366 ------------------- */
/* Every (index_i, index_j) that passes the x-break and the y test, in the reference's emission
 * order (i ascending, then j ascending in sorted order).  The manifold-map filter stays with the
 * caller (Collider.cpp:311,358).  If the list is longer than `capacity` nothing is written,
 * *count receives the required size and PHYX_B200_ERR_CAPACITY is returned. */
PHYX_B200_API int phyx_b200_sweep_pairs(phyx_b200_ctx* ctx, phyx_pair* pairs, int64_t capacity, int64_t* count,
    phyx_b200_broadphase_stats* stats);

/* ---- Solver::SolveJoints, reference src/Solver.cpp:17-119 ------------------------------------- */
/* Runs PrepareBodies .. FinishBodies on the bodies resident in the context: velocities and
 * displacing velocities are updated on the device, the cached impulses are written back into
 * `joints` (Solver::contactJoints, Solver.h:108). */
PHYX_B200_API int phyx_b200_solve_joints(phyx_b200_ctx* ctx, phyx_contact_joint* joints, int jointCount,
    const phyx_contact_point* contactPoints, int contactPointCount, const phyx_b200_solve_config* config,
    phyx_b200_solve_stats* stats);

/* The schedule the last solve used: slots[k] = joint index or -1, levels[l] = {start, grouped_end,
 * end} (same meaning as oracle/phyx_oracle.h).  Pass NULL to query sizes. */
PHYX_B200_API int phyx_b200_get_schedule(phyx_b200_ctx* ctx, int32_t* slots, int32_t slotCapacity,
    int32_t* levels3, int32_t levelCapacity, int32_t* slotCount, int32_t* levelCount);

/* Tuning of Solver::SolveJoints on the resident joint cache (the reference's counterpart is the choice of SIMD width and
 * island mode in Configuration, src/Configuration.h:3-23).  kernelForm: 0 = choose (strip-local whenever its layout is
 * usable, else streaming / record form by the previous step's activity), 1 / 2 = force the streaming / record form on the
 * colour-major layout, 3 = require the strip-local form (the solve fails with PHYX_B200_ERR_STATE if the layout is
 * rejected).  strips: 0 = choose, -1 = never, n = cut the solver rows into n strips (<= number of SMs).  All forms
 * relax the same joints; 1 and 2 give identical results, 3 uses the slot order phyx_b200_get_schedule reports. */
PHYX_B200_API int phyx_b200_solve_tuning(phyx_b200_ctx* ctx, int kernelForm, int strips);
/* Balance feedback of the strip layout.  measured != 0 (default): the row cuts of a step are balanced with the SM clocks
 * each strip's CTA spent in the previous solve, so on worlds whose strips share manifolds (one wide island) the slot order,
 * and with it the last bits of the result, depend on timing: every run is a valid sweep (and is checked as such against
 * the oracle), two runs are not bit-identical.  measured = 0: cuts follow the predicted work only; runs are reproducible
 * bit for bit at a few percent of solve time.  (Reference counterpart: none; its multithreaded island modes are not
 * reproducible either, SURVEY App. B.) */
PHYX_B200_API int phyx_b200_strip_feedback(phyx_b200_ctx* ctx, int measured);
/* The strip layout of the last solve: *strips = S (0: the last solve did not use strips), cuts[S+1] = row cuts,
 * classSlotStart[2S+1] = first slot of each class (class k < S: interior of strip k, class S+k: cut set between strips
 * k and k+1), info[8] = {usable, reject mask, rows of the largest strip, rows of the largest cut set, largest bin,
 * last rejection (sticky: mask | strips << 8 | widest strip rows / 64 << 20), colours, cut manifolds}.  capacity = entries available in cuts / classSlotStart; any pointer may be NULL. */
PHYX_B200_API int phyx_b200_strip_plan(phyx_b200_ctx* ctx, int32_t* strips, int32_t* cuts, int32_t* classSlotStart, int32_t capacity, int32_t* info);

/* Developer aid (no reference counterpart; the reference instruments its solve with microprofile scopes, Solver.cpp:132):
 * time stamps of the strip-local kernel.  passes >= 0: the following solves record, per CTA and pass (warm start = 0),
 * 8 words of %globaltimer: [0] pass start, [1] interior done, [2] neighbour's rows arrived, [3] cut set done, [4] pass end.
 * out != NULL: copy the stamps of the last solve ([strips][passes][8] words; *strips = 0 if none were recorded). */
PHYX_B200_API int phyx_b200_strip_trace(phyx_b200_ctx* ctx, int passes, uint64_t* out, int64_t capacity, int32_t* strips);

/* ---- resident collider stages: a whole World::Update without leaving HBM ------------------------ */
/* The manifold cache (Collider::manifolds / contactPoints / manifoldMap, Collider.h:58-61) and the joint
 * cache (Solver::contactJoints, Solver.h:108) live in the context; every stage below works on them in
 * place and reproduces the reference's ordering rules (sweep-order append, swap-with-last removal),
 * so the arrays read back with the download calls are bit-identical to the reference's. */

/* Collider::UpdatePairs, reference src/Collider.cpp:286-366: sweep + "not in the cache yet" filter; new
 * manifolds are appended in sweep order.  stats->pairs counts every overlapping pair (cache lookups). */
PHYX_B200_API int phyx_b200_update_pairs(phyx_b200_ctx* ctx, phyx_b200_broadphase_stats* stats);
/* Collider::UpdateManifolds, reference src/Collider.cpp:8-245,368-377: box-box SAT + clipping, <= 2 points */
PHYX_B200_API int phyx_b200_update_manifolds(phyx_b200_ctx* ctx);
/* Collider::PackManifolds, reference src/Collider.cpp:379-416 */
PHYX_B200_API int phyx_b200_pack_manifolds(phyx_b200_ctx* ctx);
/* World::RefreshContactJoints, reference src/World.cpp:72-149; the three counters are the reference's
 * Matched / Created / Deleted meta counters (World.cpp:146-148); any of them may be NULL */
PHYX_B200_API int phyx_b200_refresh_contact_joints(phyx_b200_ctx* ctx, int32_t* matched, int32_t* created, int32_t* deleted);
/* Solver::SolveJoints on the resident joint cache (same as solve_joints without the host arrays) */
PHYX_B200_API int phyx_b200_solve_resident(phyx_b200_ctx* ctx, const phyx_b200_solve_config* config, phyx_b200_solve_stats* stats);
/* ---- World::Update as ONE call: reference src/World.cpp:19-37 ------------------------------------------------------- */
/* IntegrateVelocity, UpdateBroadphase, UpdatePairs, UpdateManifolds, PackManifolds, RefreshContactJoints, SolveJoints,
 * IntegratePosition on the resident state, in that order, with the results of calling the eight stage functions above one
 * after the other.  The difference is on the host side: the stage functions each return their counts, which costs a
 * device read-back per stage; this call keeps the counts of the step in flight on the device (buffers and grids are sized
 * by bounds predicted from the previous step) and reads everything back once, at the end.  A step whose counts outgrow the
 * bounds, or in which the device takes a decision the stage path takes on the host (strip layout rejected, colouring
 * rebuilt), stops on the device before the stage in question has changed anything and is finished by the stage functions
 * (info->stopStage / stopReason say why).  The first steps of a world, replay schedules, forced kernel forms and
 * partitioned worlds always take the stage path (info->deferred = 0). */
typedef struct {
    int32_t deferred;                    /* 1: the whole step ran with device-side counts and one read-back */
    int32_t stopStage, stopReason;       /* deferred attempt stopped at stage (3 UpdatePairs, 6 RefreshContactJoints, 7 SolveJoints;
                                            0 = not); reason: 1 sweep items, 2 new pairs, 3 new joints above their bound, 10 strip
                                            layout rejected, 11 strip shape above the predicted one, 12 more than 64 colours,
                                            13 a body changed between static and dynamic, 14 colouring drifted (rebuilt) */
    int32_t manifolds, contactPoints, joints;   /* sizes of Collider::manifolds / contactPoints, Solver::contactJoints after the step */
    int32_t newPairs, jointsCreated, jointsDeleted;
    int32_t graphReplay;                 /* 1: the step was one CUDA-graph launch (steady state: same bounds and buffers as the step before) */
    int32_t graphStatus;                 /* why not, when it was not: 0 first step with these launches, 2 captured and launched now,
                                            3 launches differ from the previous step's (bounds or buffers changed), -1 capture failed */
    int32_t pad_;
    int64_t pairs, tests;                /* overlapping pairs / sweep tests of this step's broadphase */
    int64_t deferredSteps, deferredStops;       /* totals of this context */
} phyx_b200_step_info;
PHYX_B200_API int phyx_b200_world_step(phyx_b200_ctx* ctx, float dt, float gravity, const phyx_b200_solve_config* config,
    phyx_b200_solve_stats* solveStats, phyx_b200_broadphase_stats* broadphaseStats, phyx_b200_step_info* info);
/* deferred = 1 (default): world_step may run steps with device-side counts; 0: it always calls the stage functions;
 * 2: as 1 with bounds that leave no headroom (test aid: every growing count exercises the stop-and-resume path);
 * 3: as 1 without CUDA-graph replay of the steady-state step */
PHYX_B200_API int phyx_b200_step_mode(phyx_b200_ctx* ctx, int deferred);

/* resetWorld() clears manifolds, manifoldMap and contactJoints (reference src/main.cpp:86-89) */
PHYX_B200_API int phyx_b200_reset_collider(phyx_b200_ctx* ctx);
/* sizes of Collider::manifolds, Collider::contactPoints, Solver::contactJoints */
PHYX_B200_API int phyx_b200_collider_counts(phyx_b200_ctx* ctx, int32_t* manifolds, int32_t* contactPoints, int32_t* joints);
/* host mirrors of the three arrays (the demo reads them for rendering and the HUD, main.cpp:357-413) */
PHYX_B200_API int phyx_b200_download_manifolds(phyx_b200_ctx* ctx, phyx_manifold* out, int capacity);
PHYX_B200_API int phyx_b200_download_contact_points(phyx_b200_ctx* ctx, phyx_contact_point* out, int capacity);
PHYX_B200_API int phyx_b200_download_joints(phyx_b200_ctx* ctx, phyx_contact_joint* out, int capacity);
/* push a complete collider state built elsewhere (contactPoints holds 2*manifoldCount records) */
PHYX_B200_API int phyx_b200_upload_collider(phyx_b200_ctx* ctx, const phyx_manifold* manifolds, int manifoldCount,
    const phyx_contact_point* contactPoints, const phyx_contact_joint* joints, int jointCount);

/* ---- islands: Solver::GatherIslands, reference src/Solver.cpp:285-454 ----------------------------------------------- */
/* Union-find over the dynamic bodies joined by the resident contact joints (static bodies do not merge islands), islands
 * numbered in order of their first body, consecutive islands coalesced into groups of >= 256 joints.  Call after
 * refresh_contact_joints.  *islandCount / *islandMaxSize are Solver::islandCount / Solver::islandMaxSize (groups and the
 * joints of the largest group, what the demo's HUD shows, src/main.cpp:358-360); *islands = islands before coalescing. */
PHYX_B200_API int phyx_b200_build_islands(phyx_b200_ctx* ctx, int32_t* islandCount, int32_t* islandMaxSize, int32_t* islands);
/* per body: its island (the reference's island_index[island_remap[body]]) and its group (island_indexremap of that), -1 for
 * static bodies; either pointer may be NULL */
PHYX_B200_API int phyx_b200_download_islands(phyx_b200_ctx* ctx, int32_t* islandOfBody, int32_t* groupOfBody, int32_t capacity);
/* One world's Solver::SolveJoints split BY ISLAND over `ranks` devices (the reference's island dispatch, src/Solver.cpp:73-92,
 * with devices in place of worker threads).  Every rank holds the whole world and runs the other stages redundantly; with
 * ranks > 1 solve_resident (colour schedule) builds the islands, gives every rank a contiguous run of island groups with
 * about equal joint counts, and relaxes only the manifolds of this rank's islands: islands exchange no impulses, so there is
 * no communication inside the solve.  Afterwards island_pack writes this rank's results (velocity / displacement rows of its
 * bodies, cached impulses of its joints; zero elsewhere) as island_exchange_words 32-bit words into a device buffer of the
 * caller; an integer SUM all-reduce over the ranks (e.g. NCCL, one non-zero term per word: exact) followed by island_unpack
 * on every rank leaves all replicas with the complete, identical state.  ranks = 1 switches the split off. */
PHYX_B200_API int phyx_b200_island_partition(phyx_b200_ctx* ctx, int rank, int ranks);
/* the rank that owns each body under the active island partition (static bodies: rank 0), as built by the last
 * build_islands / solve_resident */
PHYX_B200_API int phyx_b200_download_body_owners(phyx_b200_ctx* ctx, uint8_t* ownerOfBody, int32_t capacity);
PHYX_B200_API int phyx_b200_island_exchange_words(phyx_b200_ctx* ctx, int64_t* words);
PHYX_B200_API int phyx_b200_island_pack(phyx_b200_ctx* ctx, int32_t* deviceBuffer);
PHYX_B200_API int phyx_b200_island_unpack(phyx_b200_ctx* ctx, const int32_t* deviceBuffer);

/* ---- one world over several devices (an island that spans devices) ------------------------ */
/* The reference has no counterpart (one process, one address space); in its terms this is Solver::SolveJoints
 * (src/Solver.cpp:68-215) of ONE island executed by `ranks` devices.  Every rank holds the whole world and
 * runs the other stages redundantly (deterministic, so the replicas stay bit-identical); the solve is split by
 * solver row (sorted-x order): a rank relaxes the manifolds whose dynamic bodies lie in its row range, all
 * ranks relax the few manifolds that straddle a cut, and the rows those touch ("ghost bodies") travel between
 * the devices once per pass (warm start, every impulse iteration, every displacement iteration) over NVLink
 * peer memory.  Results equal the one-device sweep over the same slot order (phyx_b200_get_schedule), with
 * static bodies' lastIteration tracked per rank.
 *
 * partition_create allocates this rank's exchange buffer (boundaryCapacity rows per peer and pass, bulkBytes
 * per peer for the end-of-solve exchange: 32 B per body + 8 B per schedule slot of the largest rank) and
 * returns its CUDA IPC handle (64 bytes) and/or its device pointer; partition_attach receives the handles of
 * all ranks (ranks x 64 bytes, other processes) or their pointers (same process; peerDevices may name the
 * devices so that peer access gets enabled). */
PHYX_B200_API int phyx_b200_partition_create(phyx_b200_ctx* ctx, int rank, int ranks, int boundaryCapacity, size_t bulkBytes,
    void* ipcHandleOut, void** localPointerOut);
PHYX_B200_API int phyx_b200_partition_attach(phyx_b200_ctx* ctx, const void* ipcHandles, void* const* localPointers, const int* peerDevices);
PHYX_B200_API int phyx_b200_partition_destroy(phyx_b200_ctx* ctx);
/* the plan of the last partitioned schedule: row cuts [ranks+1], first boundary row of each rank in the boundary
 * list [ranks+1], first slot of each class [ranks+2] (class `ranks` = the cut manifolds); any may be NULL */
PHYX_B200_API int phyx_b200_partition_plan(phyx_b200_ctx* ctx, int32_t* cuts, int32_t* boundaryStart, int32_t* classSlotStart);
/* Solver::SolveJoints on the resident joint cache, this rank's share; collective: every rank's process calls
 * it for the same step.  A rank whose peers do not show up gives up after 4 s with PHYX_B200_ERR_STATE. */
PHYX_B200_API int phyx_b200_solve_partitioned(phyx_b200_ctx* ctx, const phyx_b200_solve_config* config, phyx_b200_solve_stats* stats);
/* the same for all ranks living in ONE process (group[k] is rank k; stats has `count` entries or is NULL) */
PHYX_B200_API int phyx_b200_solve_partitioned_group(phyx_b200_ctx* const* group, int count, const phyx_b200_solve_config* config,
    phyx_b200_solve_stats* stats);

/* ---- device-resident variants (inputs already in HBM; used for kernel-only timing) ------------ */
/* Stage joints + contact points in HBM once ... */
PHYX_B200_API int phyx_b200_stage_joints(phyx_b200_ctx* ctx, const phyx_contact_joint* joints, int jointCount,
    const phyx_contact_point* contactPoints, int contactPointCount);
/* ... then solve them in place (cached impulses stay on the device; read with fetch_joints). */
PHYX_B200_API int phyx_b200_solve_staged(phyx_b200_ctx* ctx, const phyx_b200_solve_config* config, phyx_b200_solve_stats* stats);
PHYX_B200_API int phyx_b200_fetch_joints(phyx_b200_ctx* ctx, phyx_contact_joint* joints, int jointCount);
/* snapshot / restore of the resident body state and of the staged joints' cached impulses (so every
 * timed step does identical work) */
PHYX_B200_API int phyx_b200_snapshot_bodies(phyx_b200_ctx* ctx);
PHYX_B200_API int phyx_b200_restore_bodies(phyx_b200_ctx* ctx);
/* sweep without the D2H copy: counts only */
PHYX_B200_API int phyx_b200_sweep_pairs_resident(phyx_b200_ctx* ctx, phyx_b200_broadphase_stats* stats);

#ifdef __cplusplus
}
#endif
#endif
