// phyx_b200 — shared declarations of the sm_100a hot path (context, buffers, error plumbing).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "phyx_b200.h"

namespace phyx
{

void set_error(const char* fmt, ...);

#define PHYX_CUDA(call)                                                                          \
    do                                                                                           \
    {                                                                                            \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess)                                                                   \
        {                                                                                        \
            phyx::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return PHYX_B200_ERR_CUDA;                                                           \
        }                                                                                        \
    } while (0)

#define PHYX_TRY(call)              \
    do                              \
    {                               \
        int s_ = (call);            \
        if (s_ != PHYX_B200_OK)     \
            return s_;              \
    } while (0)

// grow-only device buffer
struct DevBuf
{
    void* ptr = nullptr;
    size_t cap = 0;

    template <typename T> T* as() const { return static_cast<T*>(ptr); }
    int reserve(size_t bytes);
    int reserve_keep(size_t bytes, size_t keepBytes, cudaStream_t stream);
    void release();
};

// One level of the solve schedule: slots [start, grouped_end) are 8-wide units (AVX2 skip rule),
// slots [grouped_end, end) are 1-wide units.  start % 8 == 0.
struct Level
{
    int start, grouped_end, end;
};

constexpr int kStaticBit = 1 << 30;       // body reference flag in the packed joint: body is static
constexpr int kBodyMask = kStaticBit - 1;

constexpr int kMaxRanks = 8;              // devices one world can be partitioned over
constexpr int kMaxColours = 64;           // colours per class of the schedule
constexpr int kColourDrift = 1;           // the incremental colouring is rebuilt when it uses this many colours more than the last full build

// Partition of one world's solve over `ranks` devices (partition.cu).  Every rank holds the whole world and runs
// the collider stages redundantly (they are deterministic, so the replicas stay bit-identical); the solve is
// split by solver row (sorted-x order): rank q owns the rows [cuts[q], cuts[q+1]) and the manifolds whose
// dynamic bodies all lie in them ("interior"); manifolds whose two bodies belong to different ranks are "cut".
struct Partition
{
    int rank = 0, ranks = 1;
    // exchange buffer of this rank (one cudaMalloc, exported with CUDA IPC) and the peers' views of theirs
    char* xbuf = nullptr;
    size_t xbytes = 0;
    char* peer[kMaxRanks] = {};
    bool peerOpened[kMaxRanks] = {};   // opened with cudaIpcOpenMemHandle (else: same process, not ours to close)
    int boundaryCapacity = 0;          // rows per sender and slot of the boundary exchange
    size_t bulkCapacity = 0;           // bytes per sender of the end-of-solve exchange
    unsigned long long bulkSeq = 0;
    // plan of the current schedule (host copies; built by colour.cu with the schedule)
    int cuts[kMaxRanks + 1] = {};      // row cuts
    int bStart[kMaxRanks + 1] = {};    // boundary rows owned by rank q: bRows[bStart[q] .. bStart[q+1])
    int classSlotStart[kMaxRanks + 2] = {};   // slots of class q (q = ranks: the cut class): [classSlotStart[q], classSlotStart[q+1])
    int numInterior = 0, numCut = 0;   // levels of this rank in partLevels: interior first, then cut
    bool planValid = false;
    bool failed = false;               // a solve timed out waiting for a peer: the partition must be re-created
    DevBuf rowFlag, rowPrefix, bRows, planWords, binLevels, partLevels, state;
    int widestInterior = 0, widestCut = 0;   // slots of the widest level of each kind (grid sizing)
    void* params = nullptr;            // SolveParams of the solve in flight (solve.cu)
    cudaEvent_t evReady = nullptr;     // group driver: "this rank has sent"
    int launchIndex = 0;
};

// Strip-local solve (strips.cu): the solver rows (sorted-x order) are cut into S contiguous ranges ("strips"), one per
// CTA of a persistent kernel.  A manifold whose dynamic bodies lie in one strip is INTERIOR to it (class k); one whose
// bodies lie in adjacent strips k, k+1 is CUT (class S + k).  Slots are laid out class-major, colour-minor.
struct StripPlan
{
    bool valid = false;        // the schedule of the resident joints is a strip layout and k_solve_strips can run it
    int want = 0;              // tuning (phyx_b200_solve_tuning): 0 = choose, -1 = never, > 0 = this many strips
    int strips = 0;            // S of the current layout
    int autoLimit = 0;         // largest S worth trying when choosing (halved whenever a layout is rejected as too narrow)
    int limitAge = 0;          // usable layouts since autoLimit last changed (strip_limit_recover)
    int rowLimit = 0;          // rows a strip of the current layout may have (strips.cu strip_row_limit)
    int rowLimitForce = 0;     // > 0: a tighter limit for the retry of a layout that was rejected for its shared-memory size
    int maxStripRows = 0, maxCutRows = 0, maxBin = 0, colours = 0, cutManifolds = 0, manifolds = 0;
    bool attributeSet = false;
    int rejected = 0;          // why the last layout attempt was not usable (bit mask, see strips.cu), 0 = usable
    int lastReject = 0;        // the last rejection, sticky: mask | strips << 8 | (widest strip's rows / 64) << 20 (diagnostics)
    DevBuf cuts, binRange, flags, prefixR, prefixL, bR, bL, bStart, header, sync, trace, pairTest, cost, factor, prevCuts;
    int feedbackStrips = 0, feedbackBodies = 0;   // the balance feedback (measured cost per strip) belongs to this layout shape
    bool measuredFeedback = true;   // phyx_b200_strip_feedback: balance the cuts by the measured cost of the previous solve's strips
    int tracePasses = 0;       // developer aid (phyx_b200_strip_trace): passes of the next solves to time-stamp per CTA
};

// ---- deferred step (phyx_b200_world_step): the counts of a step in flight live on the device -----------------------
// The stage functions of the C ABI return their counts (pairs, manifolds, joints) to the host, which costs a read-back
// per stage.  A whole World::Update issued through phyx_b200_world_step keeps them in this block instead: the host
// sizes buffers and grids by UPPER BOUNDS predicted from the previous step, every kernel reads the true count from
// here, and ONE read-back at the end of the step brings all of them home.  A count that outgrows its bound stops the
// step on the device (every later kernel sees zero counts); the host then finishes it with the stage functions.
struct StepCtl
{
    int bodies, manifolds, joints, slots;      // live counts (zeroed by a stop; `saved` keeps them)
    int items, newPairs, appendFirst, packed;  // sweep work items, new pairs, first manifold slot of the append, survivors of the pack
    int fresh, jointsGrown, jointsKept, stop;  // new joints, joints before the cleanup, after it; stage that stopped the step (0 = none)
    int stopNeed, stopReason, pad0, pad1;
    int saved[4];                              // bodies, manifolds, joints, slots at the moment of the stop
    unsigned long long tests, hits;            // sweep totals
    int created, deleted, pad2, pad3;
};
static_assert(sizeof(StepCtl) == 112, "StepCtl layout");

enum StepStage { kStageNone = 0, kStagePairs = 3, kStageRefresh = 6, kStageSolve = 7 };   // numbered as in World::Update

// a count a kernel receives: the host's value (stage functions) or a device word times `mul` (deferred step; v is then the bound
// the grid was sized by)
struct Count
{
    int v;
    const int* p;
    int mul;
    __host__ __device__ Count(int value = 0) : v(value), p(nullptr), mul(1) {}   // a plain host count converts silently
    __host__ __device__ Count(int bound, const int* word, int times) : v(bound), p(word), mul(times) {}
};
#ifdef __CUDACC__
__device__ __forceinline__ int count_of(const Count& c) { return c.p ? min(__ldcg(c.p) * c.mul, c.v) : c.v; }
// stop the step in flight: later kernels see zero counts
__device__ __forceinline__ void ctl_stop(StepCtl* ctl, int stage, int need, int reason)
{
    if (ctl->stop) return;
    ctl->stop = stage;
    ctl->stopNeed = need;
    ctl->stopReason = reason;
    ctl->saved[0] = ctl->bodies;
    ctl->saved[1] = ctl->manifolds;
    ctl->saved[2] = ctl->joints;
    ctl->saved[3] = ctl->slots;
    ctl->bodies = ctl->manifolds = ctl->joints = ctl->slots = 0;
    ctl->items = ctl->newPairs = ctl->packed = ctl->fresh = ctl->jointsGrown = ctl->jointsKept = 0;
}
#endif

// host side of the deferred step
struct Deferred
{
    bool enabled = true;       // phyx_b200_step_mode
    bool tight = false;        // test mode (step_mode 2): bounds without headroom
    bool active = false;       // a deferred step is being issued: the context's counts are upper bounds
    bool ctlStale = true;      // the device block does not hold the host's counts (a stage function changed them)
    int capItems = 0, capNewPairs = 0, capFresh = 0;   // bounds of this step
    int rowCap = 0, cutCap = 0, workCap = 0;           // shared-memory shape of the strip kernel, from the previous layouts
    float baseRows = 0.f, baseCut = 0.f, baseBin = 0.f; // ... their running maxima (strips.cu strip_predict_caps)
    int lastItems = 0, lastNewPairs = 0, lastFresh = 0;
    int64_t steps = 0, stops = 0, ineligible = 0;      // statistics
    int lastStopStage = 0, lastStopReason = 0;
    const int* colourResult = nullptr;                 // device words of the colouring's result (colour.cu), read back with the step
    int ubManifolds = 0, ubJoints = 0;                 // bounds on the manifold / joint arrays (grids and scratch are sized by them)
    int ctlManifolds = -1, ctlJoints = -1;             // what the device block holds as the starting counts (-1: unknown)
    // the steady-state step as a CUDA graph
    bool useGraph = true, graphBroken = false, capturing = false;
    cudaGraphExec_t graphExec = nullptr;
    unsigned long long graphKey = 0, lastKey = 0;
    long long lastAllocCount = -1;
    int64_t graphLaunches = 0, graphReplays = 0;
    int graphStatus = 0;
};

} // namespace phyx

struct phyx_b200_ctx
{
    int device = 0;
    int numSMs = 0;
    cudaStream_t stream = nullptr;
    int64_t launches = 0;
    long long allocCount = 0;    // device (re)allocations made for this context (api.cu AllocTimer)

    // ---- bodies (SoA in HBM) --------------------------------------------------------------
    int bodyCount = 0;
    phyx::DevBuf vel;      // float4 {vx, vy, w, lastIteration}       = reference SolveBody (impulse)
    phyx::DevBuf disp;     // float4 {dvx, dvy, dw, lastIteration}    = reference SolveBody (displacement)
    phyx::DevBuf acc;      // float4 {ax, ay, alpha, 0}
    phyx::DevBuf params;   // float4 {invMass, invInertia, pos.x, pos.y}
    phyx::DevBuf rot;      // float4 {xVector.x, xVector.y, yVector.x, yVector.y}
    phyx::DevBuf aabb;     // float4 {min.x, min.y, max.x, max.y}
    phyx::DevBuf size;     // float2 half extents
    phyx::DevBuf aos;      // staging for the 128-byte AoS records
    phyx::DevBuf snap;     // snapshot of vel/disp/acc/params/rot/aabb
    phyx::DevBuf snapJoints;   // ... and of the staged joints (cached impulses)
    int snapJointCount = -1;
    bool hasSnapshot = false;

    // ---- broadphase -----------------------------------------------------------------------
    phyx::DevBuf sortA, sortB;   // uint2 {key, index}
    phyx::DevBuf hist;           // per-block digit counts / offsets
    phyx::DevBuf scanTmp;
    void* scanTmpCleared = nullptr;
    phyx::DevBuf entry;          // float4 {minx, maxx, centery, extenty}, sorted order
    phyx::DevBuf entryIndex;     // uint32 body index, sorted order
    phyx::DevBuf sweepEnd;       // int: end_i
    phyx::DevBuf itemStart;      // int: first work item of body i (exclusive scan)
    phyx::DevBuf items;          // int2 {i, chunk}
    phyx::DevBuf itemCount;      // int per item -> exclusive scan = output offset
    phyx::DevBuf tileLong;       // u8 per sweep tile: 1 = left to the item kernel
    phyx::DevBuf pairs;          // int2 output
    phyx::DevBuf counters;       // misc device scalars
    bool broadphaseValid = false;
    int64_t lastPairs = 0, lastTests = 0;

    // ---- collider state (persistent across steps) ---------------------------------------------
    int manifoldCount = 0;
    phyx::DevBuf manBody;        // int2 {body1, body2} per manifold (sweep order at creation: NOT canonical)
    phyx::DevBuf manCount;       // int pointCount per manifold; pointIndex is always 2*m
    phyx::DevBuf pairTable;      // open-addressing set of (body1<<32 | body2) keys of the live manifolds
    size_t pairTableSlots = 0;
    phyx::DevBuf collideTmp;     // flags / prefix sums / mover tables
    int64_t lastNewPairs = 0;

    // ---- solve ----------------------------------------------------------------------------
    int jointCount = 0, contactPointCount = 0;
    phyx::DevBuf joints;         // phyx_contact_joint AoS (device copy)
    phyx::DevBuf jointStamp;     // int per joint: refresh epoch in which a contact point last claimed it (collide.cu)
    phyx::DevBuf contactPoints;  // phyx_contact_point AoS (device copy)
    int slotCount = 0, levelCount = 0;
    phyx::DevBuf slotJoint;      // int: joint index of slot (or -1)
    phyx::DevBuf levels;         // Level[levelCount]
    phyx::DevBuf q0, q1, q2, q3; // float4 per slot (see solve.cu)
    phyx::DevBuf pairQ, pairIdx; // paired levels, record form: 128-byte record + int2 index words per manifold (solve.cu)
    phyx::DevBuf accNF;          // float2 per slot
    phyx::DevBuf accD;           // float per slot
    phyx::DevBuf stamps;         // 2 x u64 per body (impulse / displacement static-body words)
    phyx::DevBuf slotPos;        // int per slot: sequential position of the slot's unit (replay schedules)
    bool slotPosValid = false;
    phyx::DevBuf processed;      // int per slot: tick of the pass that ran it
    // strict companion of a replay schedule (schedule.cu StaticRule): used for iterations in which a static
    // body with several joints starts "cold"
    phyx::DevBuf strictLevels, strictMap, staticMulti, rowsMulti;
    int strictLevelCount = 0, numMultiStatics = 0;
    phyx::DevBuf solveFlags;     // productive flags + result words
    phyx::DevBuf solveRows;      // 2 x float4 per body: the solver's packed copy of the velocity / displacement rows
    phyx::DevBuf rowOf;          // int per body: its row in the sorted-x order of the last broadphase
    bool rowOrderValid = false;
    int rowOrderBodies = 0;
    phyx::DevBuf colourTmp;      // colouring scratch
    phyx::DevBuf colourKeys, colourSorted;   // uint2 {colour, joint} before / after the counting sort
    bool hostLevelsStale = false; // strip layout: the level list is rebuilt from the device's bin table on demand
    bool hostSlotsStale = false; // schedule lives on the device only; get_schedule fetches it on demand
    // persistent colouring state (incremental recolouring of the joint cache)
    phyx::DevBuf manColour;      // int per manifold, moves with it through PackManifolds; -1 = none (no contact points yet)
    phyx::DevBuf bodyUsed;       // u64 per body: colours taken by its manifolds
    phyx::DevBuf bodyStatic;     // u8 per body: static flag the colouring was built with
    bool colourStateValid = false;
    bool jointUnitsValid = false;   // joints are exactly the contact points of the resident manifolds (set by RefreshContactJoints)
    int colourStateBodies = 0, coloursAtFullBuild = 0, partColours = 0, coloursInUse = 0;
    std::vector<int> hostSlotPos;
    std::vector<int> hostSlots;  // last schedule (host copy, for get_schedule / KEEP_SCHEDULE)
    std::vector<phyx::Level> hostLevels;
    std::vector<int> hostPairKey; // (b1,b2) list the schedule was built for
    int scheduleMode = -1, scheduleFlags = 0;
    std::vector<phyx_contact_joint> hostJoints; // host copy of the resident joints (host-built schedules need it)
    bool hostJointsValid = false;

    // small results come back through page-locked, device-mapped host memory (a tiny kernel copies them and raises a sequence
    // flag, the host spins on the flag): a fraction of the latency of cudaMemcpyAsync to pageable memory + cudaStreamSynchronize
    int* mailboxHost = nullptr;
    int* mailboxDev = nullptr;
    int* mailboxSeqDev = nullptr;   // device copy of the last sequence number posted
    unsigned mailboxSeq = 0;
    cudaEvent_t ev[8] = {};
    cudaEvent_t evBp[4] = {};    // broadphase timing: sort start / end (update_broadphase), sweep start / end (update_pairs)
    bool sortTimed = false;
    int solveBlocksPerSM = 0, colourBlocksPerSM = 0, colourRounds = 0;
    int lastKernelForm = 0;
    float lastActiveFraction = 1.0f;   // share of the impulse joint-iterations the previous solve relaxed (kernel choice, solve.cu)

    phyx::StripPlan strip;
    // islands of the contact graph (islands.cu) and the split of one world's solve by island over several devices
    phyx::DevBuf islandTmp, bodyOwner;
    bool islandsValid = false;
    int islandBodies = 0, islandCount = 0, islandMaxSize = 0;
    int islandRank = 0, islandRanks = 1;
    int islandForestBodies = -1, islandForestAge = 0;   // union-find forest kept between steps (partition path only)
    phyx::DevBuf bodyActivity;   // int per body: last impulse iteration with a productive joint on it, previous solve (strip balance)
    bool activityValid = false;
    int activityBodies = 0;
    int forceKernelForm = 0;     // tuning: 0 = choose, 1 streaming, 2 record, 3 strips (error if the layout cannot be built)

    // ---- one world over several devices (partition.cu; SURVEY.md §8e: an island that spans devices) -----
    phyx::Partition part;

    // ---- deferred step -------------------------------------------------------------------------------------
    phyx::DevBuf ctlBuf;         // StepCtl
    phyx::Deferred def;
    phyx::StepCtl* ctl() const { return ctlBuf.as<phyx::StepCtl>(); }
    // a count for a kernel launch: the host value, or (deferred step) the device word with the host value as the bound
    phyx::Count count(int hostValue, int phyx::StepCtl::*field, int mul = 1) const
    {
        return def.active ? phyx::Count(hostValue, &(ctl()->*field), mul) : phyx::Count(hostValue);
    }
};

namespace phyx
{
// cudaEventRecord that also works while the deferred step is being captured into a graph (an external event node)
inline cudaError_t record_event(phyx_b200_ctx* c, cudaEvent_t ev)
{
    return cudaEventRecordWithFlags(ev, c->stream, c->def.capturing ? cudaEventRecordExternal : cudaEventRecordDefault);
}

// bodies.cu
int bodies_upload(phyx_b200_ctx* c, const phyx_rigid_body* bodies, int n, bool wait = true);
int bodies_download(phyx_b200_ctx* c, phyx_rigid_body* bodies, int n);
int bodies_integrate_velocity(phyx_b200_ctx* c, float dt, float gravity);
int bodies_integrate_position(phyx_b200_ctx* c, float dt);
int bodies_snapshot(phyx_b200_ctx* c, bool restore);
int bodies_dynamic_extent(phyx_b200_ctx* c, float* out4);

// broadphase.cu
int broadphase_update(phyx_b200_ctx* c);
int broadphase_sweep(phyx_b200_ctx* c, phyx_b200_broadphase_stats* stats, bool filter);
// one stable LSD pass over {key, value} pairs on `digits` (power of two <= 2048) bins of key >> shift
int radix_pass(phyx_b200_ctx* c, const uint2* src, uint2* dst, int n, int shift, int digits);
int radix_pass_count(phyx_b200_ctx* c, const uint2* src, uint2* dst, Count n, int shift, int digits);

// colour.cu
int colour_schedule_build(phyx_b200_ctx* c);

// islands.cu
int islands_build(phyx_b200_ctx* c, int ranks, int* islandCount, int* islandMaxSize, int* islandsBeforeCoalescing, bool exact);
int islands_download(phyx_b200_ctx* c, int* islandOfBody, int* groupOfBody);
size_t islands_exchange_words(const phyx_b200_ctx* c);
int islands_pack(phyx_b200_ctx* c, int32_t* deviceBuffer);
int islands_unpack(phyx_b200_ctx* c, const int32_t* deviceBuffer);

// strips.cu
int strip_choose(const phyx_b200_ctx* c, int manifolds, int bodies);
// colourResult (device, 4 ints, may be null) is read back together with the layout header into colourResultHost
int strip_layout(phyx_b200_ctx* c, int S, const int2* jb, const int* work, const int* colourResult, int* colourResultHost, bool* usable);
int strip_solve_launch(phyx_b200_ctx* c, const phyx_b200_solve_config* cfg, float4* rowsVel, float4* rowsDisp);
int strip_host_levels(phyx_b200_ctx* c, std::vector<int>* classStart);
bool strip_apply_header(phyx_b200_ctx* c, const int* host16);
void strip_limit_recover(phyx_b200_ctx* c);
int strip_row_limit(const phyx_b200_ctx* c, int S);
bool strip_predict_caps(phyx_b200_ctx* c, int* rowCap, int* cutCap, int* workCap);
void strip_release(phyx_b200_ctx* c);

// schedule.cu
int schedule_build(phyx_b200_ctx* c, const phyx_contact_joint* hostJoints, int nj, int mode, int flags);

// solve.cu
int solve_run(phyx_b200_ctx* c, const phyx_b200_solve_config* cfg, phyx_b200_solve_stats* stats);

// solve.cu, partitioned solve (one world over several devices)
int part_create(phyx_b200_ctx* c, int rank, int ranks, int boundaryCapacity, size_t bulkBytes, void* ipcHandleOut, void** localOut);
int part_attach(phyx_b200_ctx* c, const void* ipcHandles, void* const* localPointers, const int* peerDevices);
void part_destroy(phyx_b200_ctx* c);
int part_begin(phyx_b200_ctx* c, const phyx_b200_solve_config* cfg);
int part_launch(phyx_b200_ctx* c, int phase, int it, int mode);
int part_bulk_push(phyx_b200_ctx* c);
int part_bulk_pull(phyx_b200_ctx* c);
int part_end(phyx_b200_ctx* c, phyx_b200_solve_stats* stats);

// collide.cu
int collide_update_pairs(phyx_b200_ctx* c, phyx_b200_broadphase_stats* stats);
int collide_update_manifolds(phyx_b200_ctx* c);
int collide_pack_manifolds(phyx_b200_ctx* c);
int collide_refresh_joints(phyx_b200_ctx* c, int* matched, int* created, int* deleted);
int collide_reset(phyx_b200_ctx* c);
int collide_rebuild_pair_table(phyx_b200_ctx* c);
int collide_rebuild_pair_table_for(phyx_b200_ctx* c, int manifolds);   // sized for this many manifolds

// api.cu: small device -> host read-backs (see phyx_b200_ctx::mailboxHost).  stage() enqueues a copy of `bytes` (a multiple of 4, the
// sum of all staged pieces <= 3.5 KB) to byte offset `offset` of the mailbox; wait() raises the flag and spins until it arrives;
// at(offset) is the host view of what was staged.
int mailbox_stage(phyx_b200_ctx* c, const void* dev, size_t bytes, size_t offset);
int mailbox_wait(phyx_b200_ctx* c);
inline const void* mailbox_at(const phyx_b200_ctx* c, size_t offset) { return reinterpret_cast<const char*>(c->mailboxHost) + offset; }
// stage + wait + copy out
int fetch_small(phyx_b200_ctx* c, const void* dev, size_t bytes, void* out);

// scan.cu
int exclusive_scan_i32(phyx_b200_ctx* c, const int* in, int* out, int n, int* totalDevice /* may be null */);
int exclusive_scan_count(phyx_b200_ctx* c, const int* in, int* out, Count n, int* totalDevice /* may be null */);
} // namespace phyx
