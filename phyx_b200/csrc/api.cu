// phyx_b200 — the C ABI (include/phyx_b200.h): context lifetime, host<->device staging, stage calls.
#include "common.cuh"

#include <stdarg.h>

#include <algorithm>
#include <atomic>
#include <chrono>

static_assert(sizeof(phyx_rigid_body) == 128, "RigidBody must keep the reference layout (128 B)");
static_assert(sizeof(phyx_contact_joint) == 20, "ContactJoint must keep the reference layout (20 B)");
static_assert(sizeof(phyx_contact_point) == 32, "ContactPoint must keep the reference layout (32 B)");
static_assert(sizeof(phyx_manifold) == 16, "Manifold must keep the reference layout (16 B)");
static_assert(sizeof(phyx_broadphase_entry) == 20, "BroadphaseEntry must keep the reference layout (20 B)");
static_assert(sizeof(phyx::Level) == 12, "Level is three ints");

namespace phyx
{

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

// Device allocations made so far and the host time they took (phyx_b200_alloc_stats).  cudaMalloc / cudaFree are
// driver round trips that synchronise the device; on shared hosts a single one was seen to take 0.3 s, so buffers
// grow with generous headroom (request + 25 %, at least 2x the old size and 1 MiB) and a steady-state step makes none.
static std::atomic<long long> g_allocNs{ 0 };   // contexts may live on different host threads
static std::atomic<long long> g_allocCount{ 0 };

static size_t grow_to(size_t bytes, size_t cap)
{
    // 25 % beyond the request even on the first allocation: the persistent arrays (manifolds, contact points,
    // joints) creep upwards by a few hundred records per step
    size_t want = std::max(std::max(bytes + bytes / 4, cap * 2), size_t(1) << 20);
    return (want + 255) & ~size_t(255);
}

// allocations of the context the calling thread is working on (set at every API entry): a buffer that moved invalidates the
// context's captured step graph
static thread_local long long* t_ctxAllocs = nullptr;

struct AllocTimer
{
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    ~AllocTimer()
    {
        g_allocNs += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
        g_allocCount++;
        if (t_ctxAllocs) ++*t_ctxAllocs;
    }
};

int DevBuf::reserve(size_t bytes)
{
    if (bytes <= cap) return PHYX_B200_OK;
    AllocTimer timer;
    const size_t want = grow_to(bytes, cap);
    void* p = nullptr;
    PHYX_CUDA(cudaMalloc(&p, want));
    if (ptr) cudaFree(ptr);   // contents are scratch or re-filled by the caller: no copy
    ptr = p;
    cap = want;
    return PHYX_B200_OK;
}

// grow, preserving the first keepBytes (persistent arrays: manifolds, contact points, joints)
int DevBuf::reserve_keep(size_t bytes, size_t keepBytes, cudaStream_t stream)
{
    if (bytes <= cap) return PHYX_B200_OK;
    AllocTimer timer;
    const size_t want = grow_to(bytes, cap);
    void* p = nullptr;
    PHYX_CUDA(cudaMalloc(&p, want));
    if (ptr)
    {
        if (keepBytes) PHYX_CUDA(cudaMemcpyAsync(p, ptr, std::min(keepBytes, cap), cudaMemcpyDeviceToDevice, stream));
        PHYX_CUDA(cudaStreamSynchronize(stream));
        cudaFree(ptr);
    }
    ptr = p;
    cap = want;
    return PHYX_B200_OK;
}

void DevBuf::release()
{
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
}

constexpr size_t kMailboxBytes = 4096, kMailboxFlag = 4096 - 64;   // flag word at the start of the last cache line

__global__ void k_mailbox_copy(const int* __restrict__ src, int words, int* __restrict__ dst)
{
    for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
}

// (the sequence number of the last post is also kept in device memory: k_mailbox_post counts on from it, so that its launch
// parameters do not change from step to step and the deferred step can be replayed as a CUDA graph)
__global__ void k_mailbox_flag(volatile int* flag, int seq, int* __restrict__ seqDev)
{
    __threadfence_system();   // (kernels of one stream run in order: the staged copies are complete; make them visible first)
    *seqDev = seq;
    *flag = seq;
}

int mailbox_stage(phyx_b200_ctx* c, const void* dev, size_t bytes, size_t offset)
{
    if (bytes % 4 || offset % 4 || offset + bytes > kMailboxFlag)
    {
        set_error("mailbox: bad piece (%zu bytes at %zu)", bytes, offset);
        return PHYX_B200_ERR_ARGUMENT;
    }
    k_mailbox_copy<<<1, 64, 0, c->stream>>>(static_cast<const int*>(dev), int(bytes / 4), c->mailboxDev + offset / 4);
    return PHYX_B200_OK;
}

static int mailbox_spin(phyx_b200_ctx* c, int seq);

int mailbox_wait(phyx_b200_ctx* c)
{
    const int seq = int(++c->mailboxSeq & 0x7fffffffu);
    k_mailbox_flag<<<1, 1, 0, c->stream>>>(c->mailboxDev + kMailboxFlag / 4, seq, c->mailboxSeqDev);
    PHYX_CUDA(cudaGetLastError());
    return mailbox_spin(c, seq);
}

// several pieces and the flag in ONE launch (the deferred step's only read-back)
struct MailPieces
{
    const int* src[4];
    int words[4], offset[4];   // offsets in words
    int count;
};

__global__ void k_mailbox_post(MailPieces p, int* __restrict__ box, volatile int* flag, int* __restrict__ seqDev)
{
    for (int k = 0; k < p.count; ++k)
        for (int i = threadIdx.x; i < p.words[k]; i += blockDim.x) box[p.offset[k] + i] = p.src[k][i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
    {
        __threadfence_system();
        const int seq = (*seqDev + 1) & 0x7fffffff;
        *seqDev = seq;
        *flag = seq;
    }
}

static int mailbox_post(phyx_b200_ctx* c, const MailPieces& p)
{
    k_mailbox_post<<<1, 128, 0, c->stream>>>(p, c->mailboxDev, c->mailboxDev + kMailboxFlag / 4, c->mailboxSeqDev);
    c->launches++;
    PHYX_CUDA(cudaGetLastError());
    return PHYX_B200_OK;
}
// the host side of a post: wait for the next sequence number
static int mailbox_await_post(phyx_b200_ctx* c)
{
    const int seq = int(++c->mailboxSeq & 0x7fffffffu);
    return mailbox_spin(c, seq);
}

static int mailbox_spin(phyx_b200_ctx* c, int seq)
{
    volatile int* flag = c->mailboxHost + kMailboxFlag / 4;
    for (unsigned spins = 0;; ++spins)
    {
        if (*flag == seq) break;
        if ((spins & 0xffffu) == 0xffffu)
        {
            // a faulted kernel never raises the flag: ask the driver now and then
            const cudaError_t e = cudaStreamQuery(c->stream);
            if (e != cudaSuccess && e != cudaErrorNotReady)
            {
                set_error("device error while waiting for a result: %s", cudaGetErrorString(e));
                return PHYX_B200_ERR_CUDA;
            }
            if (e == cudaSuccess && *flag != seq)
            {
                // the stream has drained without the flag (cannot happen unless the launch itself failed)
                if (*flag == seq) break;
                set_error("mailbox flag lost");
                return PHYX_B200_ERR_CUDA;
            }
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    return PHYX_B200_OK;
}

int fetch_small(phyx_b200_ctx* c, const void* dev, size_t bytes, void* out)
{
    PHYX_TRY(mailbox_stage(c, dev, bytes, 0));
    PHYX_TRY(mailbox_wait(c));
    memcpy(out, mailbox_at(c, 0), bytes);
    return PHYX_B200_OK;
}

static int check(phyx_b200_ctx* c)
{
    if (!c)
    {
        set_error("null context");
        return PHYX_B200_ERR_ARGUMENT;
    }
    PHYX_CUDA(cudaSetDevice(c->device));
    t_ctxAllocs = &c->allocCount;
    return PHYX_B200_OK;
}

static float elapsed_ms(cudaEvent_t a, cudaEvent_t b)
{
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

// H2D of joints + contact points into the resident buffers
static int stage_joints(phyx_b200_ctx* c, const phyx_contact_joint* joints, int nj, const phyx_contact_point* cps, int ncp)
{
    if (nj < 0 || ncp < 0 || (nj > 0 && (!joints || !cps)))
    {
        set_error("solve: bad joint / contact point arrays");
        return PHYX_B200_ERR_ARGUMENT;
    }
    // index validation happens where the schedule is built (on the device for the colour schedule)
    PHYX_TRY(c->joints.reserve(size_t(std::max(nj, 1)) * sizeof(phyx_contact_joint)));
    PHYX_TRY(c->contactPoints.reserve(size_t(std::max(ncp, 1)) * sizeof(phyx_contact_point)));
    if (nj > 0) PHYX_CUDA(cudaMemcpyAsync(c->joints.ptr, joints, size_t(nj) * sizeof(phyx_contact_joint), cudaMemcpyHostToDevice, c->stream));
    if (ncp > 0)
        PHYX_CUDA(cudaMemcpyAsync(c->contactPoints.ptr, cps, size_t(ncp) * sizeof(phyx_contact_point), cudaMemcpyHostToDevice, c->stream));
    c->jointCount = nj;
    c->contactPointCount = ncp;
    // joints / contact points now come from the caller: the resident manifold cache no longer describes them
    c->manifoldCount = 0;
    c->pairTableSlots = 0;
    c->hostJointsValid = false;
    c->colourStateValid = false;
    c->jointUnitsValid = false;
    return PHYX_B200_OK;
}

} // namespace phyx

using namespace phyx;

extern "C" {

const char* phyx_b200_last_error(void) { return g_error; }
const char* phyx_b200_version(void) { return "phyx_b200 0.1 (sm_100a)"; }

int phyx_b200_create(int device, phyx_b200_ctx** out)
{
    if (!out)
    {
        set_error("create: null out pointer");
        return PHYX_B200_ERR_ARGUMENT;
    }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
    {
        set_error("no CUDA device available (%s); phyx_b200 has no CPU fallback", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return PHYX_B200_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= count)
    {
        set_error("create: device %d outside [0,%d)", device, count);
        return PHYX_B200_ERR_ARGUMENT;
    }
    PHYX_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PHYX_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
    {
        set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return PHYX_B200_ERR_NO_DEVICE;
    }
    phyx_b200_ctx* c = new phyx_b200_ctx();
    c->device = device;
    c->numSMs = prop.multiProcessorCount;
    auto init = [&]() -> int {
        PHYX_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        for (auto& ev : c->ev) PHYX_CUDA(cudaEventCreate(&ev));
        for (auto& ev : c->evBp) PHYX_CUDA(cudaEventCreate(&ev));
        void* host = nullptr;
        PHYX_CUDA(cudaHostAlloc(&host, kMailboxBytes, cudaHostAllocMapped | cudaHostAllocPortable));
        memset(host, 0, kMailboxBytes);
        c->mailboxHost = static_cast<int*>(host);
        void* dev = nullptr;
        PHYX_CUDA(cudaHostGetDevicePointer(&dev, host, 0));
        c->mailboxDev = static_cast<int*>(dev);
        PHYX_CUDA(cudaMalloc(&dev, 64));
        PHYX_CUDA(cudaMemset(dev, 0, 64));
        c->mailboxSeqDev = static_cast<int*>(dev);
        return PHYX_B200_OK;
    };
    const int st = init();
    if (st != PHYX_B200_OK)
    {
        // do not leak the half-built context
        for (auto& ev : c->ev)
            if (ev) cudaEventDestroy(ev);
        for (auto& ev : c->evBp)
            if (ev) cudaEventDestroy(ev);
        if (c->stream) cudaStreamDestroy(c->stream);
        if (c->mailboxHost) cudaFreeHost(c->mailboxHost);
        if (c->mailboxSeqDev) cudaFree(c->mailboxSeqDev);
        delete c;
        return st;
    }
    *out = c;
    return PHYX_B200_OK;
}

void phyx_b200_destroy(phyx_b200_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    part_destroy(c);
    strip_release(c);
    if (c->def.graphExec) cudaGraphExecDestroy(c->def.graphExec);
    c->def.graphExec = nullptr;
    DevBuf* bufs[] = { &c->vel, &c->disp, &c->acc, &c->params, &c->rot, &c->aabb, &c->size, &c->aos, &c->snap, &c->snapJoints, &c->sortA, &c->sortB, &c->hist,
        &c->scanTmp, &c->entry, &c->entryIndex, &c->sweepEnd, &c->itemStart, &c->items, &c->itemCount, &c->pairs, &c->counters, &c->joints,
        &c->contactPoints, &c->slotJoint, &c->levels, &c->q0, &c->q1, &c->q2, &c->q3, &c->accNF, &c->accD, &c->stamps, &c->solveFlags, &c->slotPos, &c->processed,
        &c->colourTmp, &c->colourKeys, &c->colourSorted, &c->manBody, &c->manCount, &c->pairTable, &c->collideTmp, &c->manColour, &c->bodyUsed, &c->bodyStatic, &c->solveRows, &c->rowOf, &c->tileLong, &c->strictLevels, &c->strictMap, &c->staticMulti, &c->rowsMulti, &c->pairQ, &c->pairIdx, &c->bodyActivity, &c->islandTmp, &c->bodyOwner, &c->ctlBuf, &c->jointStamp };
    for (DevBuf* b : bufs) b->release();
    for (auto& ev : c->ev)
        if (ev) cudaEventDestroy(ev);
    for (auto& ev : c->evBp)
        if (ev) cudaEventDestroy(ev);
    cudaStreamDestroy(c->stream);
    if (c->mailboxHost) cudaFreeHost(c->mailboxHost);
    if (c->mailboxSeqDev) cudaFree(c->mailboxSeqDev);
    if (t_ctxAllocs == &c->allocCount) t_ctxAllocs = nullptr;   // (this thread's allocation counter pointed into the context)
    delete c;
}

int64_t phyx_b200_launch_count(const phyx_b200_ctx* c) { return c ? c->launches : 0; }

void phyx_b200_alloc_stats(int64_t* count, double* hostMs)
{
    if (count) *count = g_allocCount.load();
    if (hostMs) *hostMs = double(g_allocNs.load()) * 1e-6;
}
void* phyx_b200_stream(const phyx_b200_ctx* c) { return c ? (void*)c->stream : nullptr; }

int phyx_b200_synchronize(phyx_b200_ctx* c)
{
    PHYX_TRY(check(c));
    PHYX_CUDA(cudaStreamSynchronize(c->stream));
    return PHYX_B200_OK;
}

int phyx_b200_body_count(const phyx_b200_ctx* c) { return c ? c->bodyCount : 0; }

// Page-lock a caller buffer in place (cudaHostRegister) so uploads / downloads of World::bodies run at
// full PCIe rate without a staging copy.  Returns OK also when the range is already registered.
int phyx_b200_host_register(phyx_b200_ctx* c, void* ptr, size_t bytes)
{
    PHYX_TRY(check(c));
    if (!ptr || bytes == 0) return PHYX_B200_OK;
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterDefault);
    if (e == cudaErrorHostMemoryAlreadyRegistered)
    {
        cudaGetLastError();
        return PHYX_B200_OK;
    }
    if (e != cudaSuccess)
    {
        cudaGetLastError();
        set_error("host_register(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        return PHYX_B200_ERR_CUDA;
    }
    return PHYX_B200_OK;
}

int phyx_b200_host_unregister(phyx_b200_ctx* c, void* ptr)
{
    PHYX_TRY(check(c));
    if (!ptr) return PHYX_B200_OK;
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) cudaGetLastError();   // not registered: nothing to undo
    return PHYX_B200_OK;
}

int phyx_b200_upload_bodies(phyx_b200_ctx* c, const phyx_rigid_body* bodies, int count)
{
    PHYX_TRY(check(c));
    if (count < 0 || (count > 0 && !bodies))
    {
        set_error("upload_bodies: bad arguments");
        return PHYX_B200_ERR_ARGUMENT;
    }
    return bodies_upload(c, bodies, count);
}

int phyx_b200_upload_bodies_async(phyx_b200_ctx* c, const phyx_rigid_body* bodies, int count)
{
    PHYX_TRY(check(c));
    if (count < 0 || (count > 0 && !bodies))
    {
        set_error("upload_bodies_async: bad arguments");
        return PHYX_B200_ERR_ARGUMENT;
    }
    return bodies_upload(c, bodies, count, false);
}

int phyx_b200_download_bodies(phyx_b200_ctx* c, phyx_rigid_body* bodies, int count)
{
    PHYX_TRY(check(c));
    if (count != c->bodyCount || (count > 0 && !bodies))
    {
        set_error("download_bodies: count %d does not match the %d resident bodies", count, c->bodyCount);
        return PHYX_B200_ERR_ARGUMENT;
    }
    return bodies_download(c, bodies, count);
}

int phyx_b200_integrate_velocity(phyx_b200_ctx* c, float dt, float gravity)
{
    PHYX_TRY(check(c));
    return bodies_integrate_velocity(c, dt, gravity);
}

int phyx_b200_integrate_position(phyx_b200_ctx* c, float dt)
{
    PHYX_TRY(check(c));
    return bodies_integrate_position(c, dt);
}

int phyx_b200_dynamic_extent(phyx_b200_ctx* c, float* minMax4)
{
    PHYX_TRY(check(c));
    if (!minMax4)
    {
        set_error("dynamic_extent: null output");
        return PHYX_B200_ERR_ARGUMENT;
    }
    return bodies_dynamic_extent(c, minMax4);
}

int phyx_b200_snapshot_bodies(phyx_b200_ctx* c)
{
    PHYX_TRY(check(c));
    return bodies_snapshot(c, false);
}

int phyx_b200_restore_bodies(phyx_b200_ctx* c)
{
    PHYX_TRY(check(c));
    return bodies_snapshot(c, true);
}

int phyx_b200_update_broadphase(phyx_b200_ctx* c)
{
    PHYX_TRY(check(c));
    return broadphase_update(c);
}

int phyx_b200_download_broadphase(phyx_b200_ctx* c, phyx_broadphase_entry* entries, int capacity)
{
    PHYX_TRY(check(c));
    int n = c->bodyCount;
    if (!c->broadphaseValid)
    {
        set_error("download_broadphase: call update_broadphase first");
        return PHYX_B200_ERR_STATE;
    }
    if (capacity < n || (n > 0 && !entries))
    {
        set_error("download_broadphase: capacity %d < %d entries", capacity, n);
        return PHYX_B200_ERR_CAPACITY;
    }
    if (n == 0) return PHYX_B200_OK;
    std::vector<float2> x(n), y(n);
    std::vector<unsigned> idx(n);
    const float2* ex = c->entry.as<float2>();
    PHYX_CUDA(cudaMemcpyAsync(x.data(), ex, size_t(n) * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
    PHYX_CUDA(cudaMemcpyAsync(y.data(), ex + n, size_t(n) * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
    PHYX_CUDA(cudaMemcpyAsync(idx.data(), c->entryIndex.ptr, size_t(n) * sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    PHYX_CUDA(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < n; ++i)
    {
        entries[i].minx = x[i].x;
        entries[i].maxx = x[i].y;
        entries[i].centery = y[i].x;
        entries[i].extenty = y[i].y;
        entries[i].index = idx[i];
    }
    return PHYX_B200_OK;
}

int phyx_b200_sweep_pairs_resident(phyx_b200_ctx* c, phyx_b200_broadphase_stats* stats)
{
    PHYX_TRY(check(c));
    return broadphase_sweep(c, stats, false);
}

int phyx_b200_sweep_pairs(phyx_b200_ctx* c, phyx_pair* pairs, int64_t capacity, int64_t* count, phyx_b200_broadphase_stats* stats)
{
    PHYX_TRY(check(c));
    if (!count || capacity < 0 || (capacity > 0 && !pairs))
    {
        set_error("sweep_pairs: bad arguments");
        return PHYX_B200_ERR_ARGUMENT;
    }
    PHYX_TRY(broadphase_sweep(c, stats, false));
    *count = c->lastPairs;
    if (c->lastPairs > capacity)
    {
        set_error("sweep_pairs: %lld pairs do not fit capacity %lld", (long long)c->lastPairs, (long long)capacity);
        return PHYX_B200_ERR_CAPACITY;
    }
    if (c->lastPairs > 0)
    {
        PHYX_CUDA(cudaMemcpyAsync(pairs, c->pairs.ptr, size_t(c->lastPairs) * sizeof(phyx_pair), cudaMemcpyDeviceToHost, c->stream));
        PHYX_CUDA(cudaStreamSynchronize(c->stream));
    }
    return PHYX_B200_OK;
}

int phyx_b200_stage_joints(phyx_b200_ctx* c, const phyx_contact_joint* joints, int jointCount, const phyx_contact_point* contactPoints,
    int contactPointCount)
{
    PHYX_TRY(check(c));
    PHYX_TRY(stage_joints(c, joints, jointCount, contactPoints, contactPointCount));
    c->hostJoints.assign(joints, joints + jointCount);
    c->hostJointsValid = true;
    PHYX_CUDA(cudaStreamSynchronize(c->stream));
    return PHYX_B200_OK;
}

int phyx_b200_solve_staged(phyx_b200_ctx* c, const phyx_b200_solve_config* cfg, phyx_b200_solve_stats* stats)
{
    PHYX_TRY(check(c));
    if (!cfg)
    {
        set_error("solve: null config");
        return PHYX_B200_ERR_ARGUMENT;
    }
    if (stats) memset(stats, 0, sizeof(*stats));
    cudaEvent_t t0 = c->ev[4], t1 = c->ev[5], t2 = c->ev[6];
    PHYX_CUDA(record_event(c, t0));
    PHYX_TRY(schedule_build(c, c->hostJointsValid ? c->hostJoints.data() : nullptr, c->jointCount, cfg->schedule, cfg->flags));
    PHYX_CUDA(record_event(c, t1));
    PHYX_TRY(solve_run(c, cfg, stats));
    PHYX_CUDA(record_event(c, t2));
    if (c->def.active) return PHYX_B200_OK;   // deferred step: nothing waits here (deferred_finish fills the statistics)
    PHYX_CUDA(cudaEventSynchronize(t2));
    if (stats)
    {
        stats->ms_schedule = elapsed_ms(t0, t1);
        stats->ms_total = elapsed_ms(t0, t2);
    }
    return PHYX_B200_OK;
}

// ---- tuning and introspection of the solve ----------------------------------------------------------------

int phyx_b200_solve_tuning(phyx_b200_ctx* c, int kernelForm, int strips)
{
    PHYX_TRY(check(c));
    if (kernelForm < 0 || kernelForm > 3 || strips < -1)
    {
        set_error("solve_tuning: kernelForm must be 0 (choose), 1, 2 or 3; strips -1 (never), 0 (choose) or a count");
        return PHYX_B200_ERR_ARGUMENT;
    }
    c->forceKernelForm = kernelForm;
    c->strip.want = strips;
    // schedules built so far may have the other layout
    c->scheduleMode = -1;
    c->colourStateValid = false;
    return PHYX_B200_OK;
}

int phyx_b200_strip_feedback(phyx_b200_ctx* c, int measured)
{
    PHYX_TRY(check(c));
    c->strip.measuredFeedback = measured != 0;
    return PHYX_B200_OK;
}

int phyx_b200_strip_plan(phyx_b200_ctx* c, int32_t* strips, int32_t* cuts, int32_t* classSlotStart, int32_t capacity, int32_t* info)
{
    PHYX_TRY(check(c));
    const StripPlan& sp = c->strip;
    if (strips) *strips = sp.valid ? sp.strips : 0;
    if (info)
    {
        // [5]: the last rejection of this context, sticky: reason mask | strips << 8 | (rows of the widest strip / 64) << 20
        const int v[8] = { sp.valid ? 1 : 0, sp.rejected, sp.maxStripRows, sp.maxCutRows, sp.maxBin, sp.lastReject, sp.colours, sp.cutManifolds };
        memcpy(info, v, sizeof(v));
    }
    if (!sp.valid || (!cuts && !classSlotStart)) return PHYX_B200_OK;
    if (capacity < 2 * sp.strips + 1)
    {
        set_error("strip_plan: need room for %d entries", 2 * sp.strips + 1);
        return PHYX_B200_ERR_CAPACITY;
    }
    if (cuts)
    {
        PHYX_CUDA(cudaMemcpyAsync(cuts, sp.cuts.ptr, size_t(sp.strips + 1) * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        PHYX_CUDA(cudaStreamSynchronize(c->stream));
    }
    if (classSlotStart)
    {
        std::vector<int> cs;
        PHYX_TRY(strip_host_levels(c, &cs));
        memcpy(classSlotStart, cs.data(), cs.size() * sizeof(int));
    }
    return PHYX_B200_OK;
}

int phyx_b200_strip_trace(phyx_b200_ctx* c, int passes, uint64_t* out, int64_t capacity, int32_t* strips)
{
    PHYX_TRY(check(c));
    StripPlan& sp = c->strip;
    if (out && sp.trace.ptr && sp.tracePasses > 0 && sp.valid)
    {
        const int64_t n = int64_t(sp.strips) * sp.tracePasses * 8;
        if (capacity < n)
        {
            set_error("strip_trace: need room for %lld words", (long long)n);
            return PHYX_B200_ERR_CAPACITY;
        }
        PHYX_CUDA(cudaMemcpyAsync(out, sp.trace.ptr, size_t(n) * 8, cudaMemcpyDeviceToHost, c->stream));
        PHYX_CUDA(cudaStreamSynchronize(c->stream));
        if (strips) *strips = sp.strips;
    }
    else if (strips)
        *strips = 0;
    if (passes >= 0) sp.tracePasses = passes;
    return PHYX_B200_OK;
}

// ---- islands ------------------------------------------------------------------------------------------------

int phyx_b200_build_islands(phyx_b200_ctx* c, int32_t* islandCount, int32_t* islandMaxSize, int32_t* islandsBeforeCoalescing)
{
    PHYX_TRY(check(c));
    return islands_build(c, c->islandRanks, islandCount, islandMaxSize, islandsBeforeCoalescing, true);
}

int phyx_b200_download_islands(phyx_b200_ctx* c, int32_t* islandOfBody, int32_t* groupOfBody, int32_t capacity)
{
    PHYX_TRY(check(c));
    if (capacity < c->bodyCount)
    {
        set_error("download_islands: capacity %d < %d bodies", capacity, c->bodyCount);
        return PHYX_B200_ERR_CAPACITY;
    }
    return islands_download(c, islandOfBody, groupOfBody);
}

int phyx_b200_island_partition(phyx_b200_ctx* c, int rank, int ranks)
{
    PHYX_TRY(check(c));
    if (ranks < 1 || ranks > 255 || rank < 0 || rank >= ranks)
    {
        set_error("island_partition: need 1 <= ranks <= 255 and 0 <= rank < ranks");
        return PHYX_B200_ERR_ARGUMENT;
    }
    if (c->part.ranks > 1)
    {
        set_error("island_partition: the context is partitioned by solver row (partition_create); the two are exclusive");
        return PHYX_B200_ERR_STATE;
    }
    c->islandRank = rank;
    c->islandRanks = ranks;
    c->islandsValid = false;
    c->scheduleMode = -1;   // schedules built so far cover other manifolds
    return PHYX_B200_OK;
}

int phyx_b200_download_body_owners(phyx_b200_ctx* c, uint8_t* ownerOfBody, int32_t capacity)
{
    PHYX_TRY(check(c));
    const int nb = c->bodyCount;
    if (c->islandRanks < 2 || !c->islandsValid || c->islandBodies != nb)
    {
        set_error("download_body_owners: no island partition is active (island_partition with ranks > 1, then build_islands or a solve)");
        return PHYX_B200_ERR_STATE;
    }
    if (capacity < nb)
    {
        set_error("download_body_owners: capacity %d < %d bodies", capacity, nb);
        return PHYX_B200_ERR_CAPACITY;
    }
    if (nb > 0)
    {
        PHYX_CUDA(cudaMemcpyAsync(ownerOfBody, c->bodyOwner.ptr, size_t(nb), cudaMemcpyDeviceToHost, c->stream));
        PHYX_CUDA(cudaStreamSynchronize(c->stream));
    }
    return PHYX_B200_OK;
}

int phyx_b200_island_exchange_words(phyx_b200_ctx* c, int64_t* words)
{
    PHYX_TRY(check(c));
    if (words) *words = int64_t(islands_exchange_words(c));
    return PHYX_B200_OK;
}

int phyx_b200_island_pack(phyx_b200_ctx* c, int32_t* deviceBuffer)
{
    PHYX_TRY(check(c));
    return islands_pack(c, deviceBuffer);
}

int phyx_b200_island_unpack(phyx_b200_ctx* c, const int32_t* deviceBuffer)
{
    PHYX_TRY(check(c));
    return islands_unpack(c, deviceBuffer);
}

// ---- one world over several devices ---------------------------------------------------------------------

int phyx_b200_partition_create(phyx_b200_ctx* c, int rank, int ranks, int boundaryCapacity, size_t bulkBytes, void* ipcHandleOut, void** localPointerOut)
{
    PHYX_TRY(check(c));
    return part_create(c, rank, ranks, boundaryCapacity, bulkBytes, ipcHandleOut, localPointerOut);
}

int phyx_b200_partition_attach(phyx_b200_ctx* c, const void* ipcHandles, void* const* localPointers, const int* peerDevices)
{
    PHYX_TRY(check(c));
    return part_attach(c, ipcHandles, localPointers, peerDevices);
}

int phyx_b200_partition_destroy(phyx_b200_ctx* c)
{
    PHYX_TRY(check(c));
    PHYX_CUDA(cudaStreamSynchronize(c->stream));
    part_destroy(c);
    return PHYX_B200_OK;
}

int phyx_b200_partition_plan(phyx_b200_ctx* c, int32_t* cuts, int32_t* boundaryStart, int32_t* classSlotStart)
{
    PHYX_TRY(check(c));
    const Partition& pt = c->part;
    if (pt.ranks < 2 || !pt.planValid)
    {
        set_error("partition_plan: no partitioned schedule has been built yet");
        return PHYX_B200_ERR_STATE;
    }
    for (int q = 0; q <= pt.ranks; ++q)
    {
        if (cuts) cuts[q] = pt.cuts[q];
        if (boundaryStart) boundaryStart[q] = pt.bStart[q];
    }
    if (classSlotStart)
        for (int q = 0; q <= pt.ranks + 1; ++q) classSlotStart[q] = pt.classSlotStart[q];
    return PHYX_B200_OK;
}

// The passes of one partitioned solve over a set of contexts.  One context per process: `count` is 1 and the
// kernels themselves wait for the peers' flags.  Several contexts in ONE process (tests, or one host process
// driving several devices): every pass is issued on all of them before the next one, and an event of each
// context orders "has sent" before the others' "takes over" (two cooperative kernels of one device cannot
// wait for each other inside the kernel).
static int solve_partitioned_passes(phyx_b200_ctx* const* group, int count, const phyx_b200_solve_config* cfg, phyx_b200_solve_stats* stats)
{
    auto all = [&](auto&& fn) -> int {
        for (int k = 0; k < count; ++k)
        {
            PHYX_TRY(check(group[k]));
            PHYX_TRY(fn(group[k]));
        }
        return PHYX_B200_OK;
    };
    auto sent = [&]() -> int {
        if (count < 2) return PHYX_B200_OK;
        for (int k = 0; k < count; ++k)
        {
            PHYX_TRY(check(group[k]));
            PHYX_CUDA(cudaEventRecord(group[k]->part.evReady, group[k]->stream));
        }
        for (int k = 0; k < count; ++k)
        {
            PHYX_TRY(check(group[k]));
            for (int j = 0; j < count; ++j)
                if (j != k) PHYX_CUDA(cudaStreamWaitEvent(group[k]->stream, group[j]->part.evReady, 0));
        }
        return PHYX_B200_OK;
    };
    auto pass = [&](int phase, int it) -> int {
        PHYX_TRY(all([&](phyx_b200_ctx* c) -> int { return part_launch(c, phase, it, 0); }));
        PHYX_TRY(sent());
        return all([&](phyx_b200_ctx* c) -> int { return part_launch(c, phase, it, 1); });
    };
    for (int k = 0; k < count; ++k)
    {
        phyx_b200_ctx* c = group[k];
        PHYX_TRY(check(c));
        if (stats) memset(&stats[k], 0, sizeof(stats[k]));
        PHYX_CUDA(cudaEventRecord(c->ev[4], c->stream));
        PHYX_TRY(schedule_build(c, nullptr, c->jointCount, PHYX_B200_SCHEDULE_COLOUR, 0));
        PHYX_CUDA(cudaEventRecord(c->ev[5], c->stream));
        PHYX_TRY(part_begin(c, cfg));
        PHYX_CUDA(cudaEventRecord(c->ev[0], c->stream));
    }
    PHYX_TRY(pass(-1, 0));
    for (int it = 0; it < cfg->contactIterationsCount; ++it) PHYX_TRY(pass(0, it));
    for (int it = 0; it < cfg->penetrationIterationsCount; ++it) PHYX_TRY(pass(1, it));
    PHYX_TRY(all([&](phyx_b200_ctx* c) -> int {
        PHYX_CUDA(cudaEventRecord(c->ev[1], c->stream));
        return part_bulk_push(c);
    }));
    PHYX_TRY(sent());
    PHYX_TRY(all([&](phyx_b200_ctx* c) -> int { return part_bulk_pull(c); }));
    for (int k = 0; k < count; ++k)
    {
        phyx_b200_ctx* c = group[k];
        PHYX_TRY(check(c));
        PHYX_CUDA(cudaEventRecord(c->ev[2], c->stream));
        PHYX_TRY(part_end(c, stats ? &stats[k] : nullptr));
        PHYX_CUDA(cudaEventRecord(c->ev[6], c->stream));
        PHYX_CUDA(cudaEventSynchronize(c->ev[6]));
        if (stats)
        {
            stats[k].ms_schedule = elapsed_ms(c->ev[4], c->ev[5]);
            stats[k].ms_refresh = elapsed_ms(c->ev[5], c->ev[0]);
            stats[k].ms_iterations = elapsed_ms(c->ev[0], c->ev[1]);
            stats[k].ms_finish = elapsed_ms(c->ev[1], c->ev[6]);   // end-of-solve exchange + FinishJoints + FinishBodies
            stats[k].ms_total = elapsed_ms(c->ev[4], c->ev[6]);
        }
    }
    return PHYX_B200_OK;
}

static int check_partitioned(phyx_b200_ctx* c, const phyx_b200_solve_config* cfg)
{
    PHYX_TRY(check(c));
    if (!cfg || cfg->schedule != PHYX_B200_SCHEDULE_COLOUR)
    {
        set_error("solve_partitioned: needs a config with schedule = PHYX_B200_SCHEDULE_COLOUR");
        return PHYX_B200_ERR_ARGUMENT;
    }
    if (c->part.ranks < 2)
    {
        set_error("solve_partitioned: the context is not partitioned (partition_create)");
        return PHYX_B200_ERR_STATE;
    }
    return PHYX_B200_OK;
}

int phyx_b200_solve_partitioned(phyx_b200_ctx* c, const phyx_b200_solve_config* cfg, phyx_b200_solve_stats* stats)
{
    PHYX_TRY(check_partitioned(c, cfg));
    return solve_partitioned_passes(&c, 1, cfg, stats);
}

int phyx_b200_solve_partitioned_group(phyx_b200_ctx* const* group, int count, const phyx_b200_solve_config* cfg, phyx_b200_solve_stats* stats)
{
    if (!group || count < 2 || count > kMaxRanks)
    {
        set_error("solve_partitioned_group: need 2..%d contexts", kMaxRanks);
        return PHYX_B200_ERR_ARGUMENT;
    }
    for (int k = 0; k < count; ++k)
    {
        PHYX_TRY(check_partitioned(group[k], cfg));
        if (group[k]->part.ranks != count || group[k]->part.rank != k)
        {
            set_error("solve_partitioned_group: context %d is rank %d of %d", k, group[k]->part.rank, group[k]->part.ranks);
            return PHYX_B200_ERR_ARGUMENT;
        }
    }
    return solve_partitioned_passes(group, count, cfg, stats);
}

// ---- resident collider stages -----------------------------------------------------------------------

int phyx_b200_update_pairs(phyx_b200_ctx* c, phyx_b200_broadphase_stats* stats)
{
    PHYX_TRY(check(c));
    c->hostJointsValid = false;
    PHYX_CUDA(cudaEventRecord(c->evBp[2], c->stream));
    PHYX_TRY(collide_update_pairs(c, stats));
    if (stats)
    {
        // CUDA-event times of the radix sort (key build + 3 passes, run by the last update_broadphase) and of this sweep
        PHYX_CUDA(cudaEventRecord(c->evBp[3], c->stream));
        PHYX_CUDA(cudaEventSynchronize(c->evBp[3]));
        stats->ms_sort = c->sortTimed ? elapsed_ms(c->evBp[0], c->evBp[1]) : 0.f;
        stats->ms_sweep = elapsed_ms(c->evBp[2], c->evBp[3]);
        stats->ms_total = stats->ms_sort + stats->ms_sweep;
    }
    return PHYX_B200_OK;
}

int phyx_b200_update_manifolds(phyx_b200_ctx* c)
{
    PHYX_TRY(check(c));
    return collide_update_manifolds(c);
}

int phyx_b200_pack_manifolds(phyx_b200_ctx* c)
{
    PHYX_TRY(check(c));
    return collide_pack_manifolds(c);
}

int phyx_b200_refresh_contact_joints(phyx_b200_ctx* c, int32_t* matched, int32_t* created, int32_t* deleted)
{
    PHYX_TRY(check(c));
    c->hostJointsValid = false;
    return collide_refresh_joints(c, matched, created, deleted);
}

int phyx_b200_solve_resident(phyx_b200_ctx* c, const phyx_b200_solve_config* cfg, phyx_b200_solve_stats* stats)
{
    PHYX_TRY(check(c));
    // island partition: which islands are this rank's is decided on the contact graph of this step
    if (c->islandRanks > 1)
    {
        if (!cfg || cfg->schedule != PHYX_B200_SCHEDULE_COLOUR)
        {
            set_error("solve: an island partition needs schedule = PHYX_B200_SCHEDULE_COLOUR");
            return PHYX_B200_ERR_ARGUMENT;
        }
        PHYX_TRY(islands_build(c, c->islandRanks, nullptr, nullptr, nullptr, false));
        c->scheduleMode = -1;
    }
    return phyx_b200_solve_staged(c, cfg, stats);
}

// ---- World::Update as one call ------------------------------------------------------------------------------------
// Reference src/World.cpp:19-37.  The eight stage functions above return their counts to the host, one read-back each;
// here the step is issued without waiting: buffers and grids are sized by bounds predicted from the previous step, the
// kernels read the true counts from a device block (common.cuh StepCtl) and ONE read-back at the end brings the counts,
// the strip layout's header and the solve's results home.  When a bound does not hold, or the device takes a decision
// the stage path takes on the host (rejected strip layout, colour rebuild), the step stops on the device before anything
// was changed by the stage in question and the host finishes it with the stage functions: results are those of the
// stage path either way.

// the counts the device block starts a step from: written from the host only when they are not what the previous deferred
// step left there (first deferred step, or stage functions ran in between) ...
__global__ void k_ctl_set(StepCtl* ctl, int manifolds, int joints)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    ctl->manifolds = manifolds;
    ctl->joints = joints;
}

// ... and the per-step part: manifolds / joints carry over from the previous step on the device, so this launch has the same
// parameters every step (the step is replayed as a CUDA graph)
__global__ void k_ctl_reset(StepCtl* ctl, int bodies)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    StepCtl z;
    memset(&z, 0, sizeof(z));
    z.bodies = bodies;
    z.manifolds = ctl->manifolds;
    z.joints = ctl->joints;
    z.jointsGrown = z.joints;
    *ctl = z;
}

// stages of World::Update by number
static int run_stages(phyx_b200_ctx* c, int first, float dt, float gravity, const phyx_b200_solve_config* cfg, phyx_b200_solve_stats* solveStats,
    phyx_b200_broadphase_stats* bpStats)
{
    if (first <= 0) PHYX_TRY(phyx_b200_integrate_velocity(c, dt, gravity));
    if (first <= 1) PHYX_TRY(phyx_b200_update_broadphase(c));
    if (first <= 2) PHYX_TRY(phyx_b200_update_pairs(c, bpStats));
    if (first <= 3) PHYX_TRY(phyx_b200_update_manifolds(c));
    if (first <= 4) PHYX_TRY(phyx_b200_pack_manifolds(c));
    if (first <= 5) PHYX_TRY(phyx_b200_refresh_contact_joints(c, nullptr, nullptr, nullptr));
    if (first <= 6) PHYX_TRY(phyx_b200_solve_resident(c, cfg, solveStats));
    return phyx_b200_integrate_position(c, dt);
}

static int bound_of(int last, int floorValue)
{
    const long long v = std::max<long long>(floorValue, (long long)last + last / 2 + 1024);
    return int(std::min<long long>((v + 1023) & ~1023ll, 1ll << 30));
}

// can this step run deferred?  Only the steady state of the default pipeline does: colour schedule on the strip layout of
// the previous step, one device, nothing forced
static bool deferred_eligible(phyx_b200_ctx* c, const phyx_b200_solve_config* cfg)
{
    const int nb = c->bodyCount;
    if (!c->def.enabled || !cfg || cfg->schedule != PHYX_B200_SCHEDULE_COLOUR || cfg->flags != 0) return false;
    if (cfg->contactIterationsCount < 0 || cfg->penetrationIterationsCount < 0 || cfg->contactIterationsCount > 60000 || cfg->penetrationIterationsCount > 60000) return false;
    if (c->part.ranks > 1 || c->islandRanks > 1) return false;
    if (c->forceKernelForm == 1 || c->forceKernelForm == 2 || c->strip.want != 0 || c->strip.tracePasses > 0) return false;
    if (nb < 2 || c->manifoldCount <= 0 || c->jointCount <= 0) return false;
    if (!c->jointUnitsValid || !c->colourStateValid || c->colourStateBodies != nb || !c->manColour.ptr || !c->bodyUsed.ptr) return false;
    if (c->scheduleMode != PHYX_B200_SCHEDULE_COLOUR || c->scheduleFlags != 0) return false;
    const StripPlan& sp = c->strip;
    if (!sp.valid || sp.strips <= 0 || sp.feedbackStrips != sp.strips || sp.feedbackBodies != nb) return false;
    if (!c->activityValid || c->activityBodies != nb) return false;
    if (c->pairTableSlots == 0) return false;
    // the strip count follows the world's size: leave the choice to the stage path when it would change by much
    const int S = strip_choose(c, c->manifoldCount, nb);
    if (S <= 0 || S * 8 > sp.strips * 9 || S * 9 < sp.strips * 8) return false;
    return strip_predict_caps(c, &c->def.rowCap, &c->def.cutCap, &c->def.workCap);
}

// keep `cap` while it is neither too tight nor wasteful for `need`: the bounds are launch parameters, and launch parameters
// that do not change let the step be replayed as a graph
static void sticky_bound(int* cap, int need, int floorValue)
{
    if (*cap < need + need / 8 + 256 || *cap > 4 * need + 4 * floorValue) *cap = bound_of(need, floorValue);
}

// bounds of this step, the device block's starting counts; returns the key of everything the issued launches depend on
static int deferred_prepare(phyx_b200_ctx* c, float dt, float gravity, const phyx_b200_solve_config* cfg, unsigned long long* keyOut)
{
    Deferred& d = c->def;
    PHYX_TRY(c->ctlBuf.reserve(sizeof(StepCtl)));
    if (d.tight)
    {
        // test mode: no headroom at all, so that any growth exercises the stop-and-resume path
        d.capItems = std::max(d.lastItems, 1);
        d.capNewPairs = std::max(d.lastNewPairs, 1);
        d.capFresh = std::max(d.lastFresh, 1);
        d.ubManifolds = c->manifoldCount + d.capNewPairs;
        d.ubJoints = c->jointCount + d.capFresh;
    }
    else
    {
        sticky_bound(&d.capItems, std::max(d.lastItems, c->bodyCount), 8192);
        sticky_bound(&d.capNewPairs, std::max(d.lastNewPairs, c->manifoldCount / 64), 4096);
        sticky_bound(&d.capFresh, std::max(d.lastFresh, c->jointCount / 64), 4096);
        // bounds on the arrays themselves (what grids and scratch are sized by)
        auto sticky_array = [](int* ub, int need) {
            if (*ub < need || *ub > need + need / 3 + 65536) *ub = int(std::min<long long>(((long long)need + need / 8 + 16383) & ~16383ll, 1ll << 30));
        };
        sticky_array(&d.ubManifolds, c->manifoldCount + d.capNewPairs);
        sticky_array(&d.ubJoints, c->jointCount + d.capFresh);
    }
    // the cache must take the new pairs without a rebuild in the middle of the step
    if (size_t(d.ubManifolds) * 2 > c->pairTableSlots) PHYX_TRY(collide_rebuild_pair_table_for(c, d.ubManifolds));
    if (d.ctlManifolds != c->manifoldCount || d.ctlJoints != c->jointCount)
    {
        k_ctl_set<<<1, 32, 0, c->stream>>>(c->ctl(), c->manifoldCount, c->jointCount);
        c->launches++;
        d.ctlManifolds = c->manifoldCount;
        d.ctlJoints = c->jointCount;
    }
    // everything the launches of deferred_enqueue depend on besides the buffers' addresses (covered by the allocation count)
    unsigned long long h = 1469598103934665603ull;
    auto mix = [&h](unsigned long long v) { h = (h ^ v) * 1099511628211ull; };
    unsigned dtBits, gBits;
    memcpy(&dtBits, &dt, 4);
    memcpy(&gBits, &gravity, 4);
    mix(unsigned(c->bodyCount)); mix(unsigned(d.capItems)); mix(unsigned(d.capNewPairs)); mix(unsigned(d.capFresh));
    mix(unsigned(d.ubManifolds)); mix(unsigned(d.ubJoints)); mix(unsigned(d.rowCap)); mix(unsigned(d.cutCap)); mix(unsigned(d.workCap));
    mix(unsigned(c->strip.strips)); mix(unsigned(c->strip.autoLimit)); mix(unsigned(strip_row_limit(c, c->strip.strips))); mix(unsigned(c->coloursAtFullBuild)); mix(c->strip.measuredFeedback ? 1u : 0u);
    mix(unsigned(cfg->contactIterationsCount)); mix(unsigned(cfg->penetrationIterationsCount)); mix(dtBits); mix(gBits);
    mix((unsigned long long)c->pairTableSlots); mix((unsigned long long)c->allocCount);
    // (belt and braces: the addresses of the buffers that grow with the world)
    const DevBuf* watched[] = { &c->manBody, &c->manCount, &c->manColour, &c->contactPoints, &c->joints, &c->collideTmp, &c->colourTmp, &c->colourKeys,
        &c->colourSorted, &c->slotJoint, &c->pairQ, &c->pairIdx, &c->accNF, &c->accD, &c->solveRows, &c->items, &c->itemCount, &c->pairs, &c->pairTable,
        &c->hist, &c->scanTmp, &c->strip.pairTest, &c->strip.sync, &c->bodyActivity, &c->jointStamp };
    for (const DevBuf* b : watched) mix((unsigned long long)(uintptr_t)b->ptr);
    *keyOut = h;
    return PHYX_B200_OK;
}

// the launches of one step, in stream order (capturable: nothing in here waits for the device)
static int deferred_enqueue(phyx_b200_ctx* c, float dt, float gravity, const phyx_b200_solve_config* cfg)
{
    Deferred& d = c->def;
    k_ctl_reset<<<1, 32, 0, c->stream>>>(c->ctl(), c->bodyCount);
    c->launches++;
    d.colourResult = nullptr;
    c->hostJointsValid = false;
    // the context's counts are BOUNDS from here on (the true counts are in the device block)
    c->manifoldCount = d.ubManifolds - d.capNewPairs;
    c->contactPointCount = 2 * c->manifoldCount;
    c->jointCount = d.ubJoints - d.capFresh;
    PHYX_TRY(bodies_integrate_velocity(c, dt, gravity));
    PHYX_TRY(broadphase_update(c));
    PHYX_CUDA(record_event(c, c->evBp[2]));
    PHYX_TRY(collide_update_pairs(c, nullptr));
    PHYX_CUDA(record_event(c, c->evBp[3]));
    PHYX_TRY(collide_update_manifolds(c));
    PHYX_TRY(collide_pack_manifolds(c));
    PHYX_TRY(collide_refresh_joints(c, nullptr, nullptr, nullptr));
    PHYX_TRY(phyx_b200_solve_staged(c, cfg, nullptr));
    PHYX_TRY(bodies_integrate_position(c, dt));
    if (!d.colourResult || !c->strip.header.ptr)
    {
        set_error("deferred step: the solve did not take the strip path (internal)");
        return PHYX_B200_ERR_STATE;
    }
    // one read-back: counts | layout header | colouring result | solve result
    MailPieces p;
    p.count = 4;
    p.src[0] = reinterpret_cast<const int*>(c->ctl());
    p.words[0] = int(sizeof(StepCtl) / 4);
    p.offset[0] = 0;
    p.src[1] = c->strip.header.as<int>();
    p.words[1] = 16;
    p.offset[1] = 128 / 4;
    p.src[2] = d.colourResult;
    p.words[2] = 4;
    p.offset[2] = 192 / 4;
    p.src[3] = reinterpret_cast<const int*>(c->solveFlags.as<char>() + 32);
    p.words[3] = 8;
    p.offset[3] = 208 / 4;
    return mailbox_post(c, p);
}

static void graph_drop(phyx_b200_ctx* c)
{
    if (c->def.graphExec) cudaGraphExecDestroy(c->def.graphExec);
    c->def.graphExec = nullptr;
    c->def.graphKey = 0;
}

// Issue one deferred step and wait for its read-back.  Steps whose launches are the same as the previous step's (same
// bounds, same buffers: the steady state) are captured into a CUDA graph once and replayed: ~110 stream operations become
// one launch.
static int deferred_issue(phyx_b200_ctx* c, float dt, float gravity, const phyx_b200_solve_config* cfg, bool* replayed)
{
    Deferred& d = c->def;
    *replayed = false;
    d.graphStatus = 0;
    unsigned long long key = 0;
    PHYX_TRY(deferred_prepare(c, dt, gravity, cfg, &key));
    d.active = true;
    if (d.useGraph && d.graphExec && d.graphKey == key)
    {
        PHYX_CUDA(cudaGraphLaunch(d.graphExec, c->stream));
        c->launches += d.graphLaunches;
        *replayed = true;
        d.graphStatus = 1;
        d.graphReplays++;
        // (the host-side bookkeeping of deferred_enqueue is the same every step: it still stands from the captured one,
        // except the bounds-as-counts, which deferred_finish replaces by the true counts anyway)
        return mailbox_await_post(c);
    }
    if (d.useGraph && !d.graphBroken && d.lastKey == key && key != 0)
    {
        // second step in a row with the same launches: capture it
        graph_drop(c);
        const int64_t launches0 = c->launches;
        cudaGraph_t graph = nullptr;
        d.capturing = true;
        cudaError_t e = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed);
        int st = PHYX_B200_OK;
        if (e == cudaSuccess)
        {
            st = deferred_enqueue(c, dt, gravity, cfg);
            e = cudaStreamEndCapture(c->stream, &graph);
        }
        d.capturing = false;
        if (st == PHYX_B200_OK && e == cudaSuccess && graph)
        {
            e = cudaGraphInstantiate(&d.graphExec, graph, 0);
            cudaGraphDestroy(graph);
            if (e == cudaSuccess && c->allocCount == d.lastAllocCount)
            {
                d.graphKey = key;
                d.graphLaunches = c->launches - launches0;
                PHYX_CUDA(cudaGraphLaunch(d.graphExec, c->stream));
                *replayed = true;
                d.graphStatus = 2;
                d.graphReplays++;
                return mailbox_await_post(c);
            }
            graph_drop(c);
        }
        else if (graph)
            cudaGraphDestroy(graph);
        // the capture did not work out (nothing of it has run): issue the step on the stream, and do not try again
        set_error("graph capture of the deferred step failed: %s (status %d)", cudaGetErrorString(e), st);
        cudaGetLastError();
        c->launches = launches0;
        d.graphStatus = -1;
        if (c->allocCount == d.lastAllocCount) d.graphBroken = true;
    }
    else if (d.lastKey != key)
        d.graphStatus = 3;
    d.lastKey = key;
    d.lastAllocCount = c->allocCount;
    PHYX_TRY(deferred_enqueue(c, dt, gravity, cfg));
    // (an allocation during the enqueue moves a buffer: the next step's launches differ, whatever the key says)
    if (c->allocCount != d.lastAllocCount) d.lastKey = 0;
    return mailbox_await_post(c);
}

int phyx_b200_world_step(phyx_b200_ctx* c, float dt, float gravity, const phyx_b200_solve_config* cfg, phyx_b200_solve_stats* solveStats,
    phyx_b200_broadphase_stats* bpStats, phyx_b200_step_info* info)
{
    PHYX_TRY(check(c));
    if (!cfg)
    {
        set_error("world_step: null config");
        return PHYX_B200_ERR_ARGUMENT;
    }
    if (info) memset(info, 0, sizeof(*info));
    if (solveStats) memset(solveStats, 0, sizeof(*solveStats));
    Deferred& d = c->def;
    int resumeFrom = 0;
    bool ranDeferred = false;
    if (deferred_eligible(c, cfg))
    {
        bool replayed = false;
        const int st = deferred_issue(c, dt, gravity, cfg, &replayed);
        d.active = false;
        d.capturing = false;
        if (info) info->graphReplay = replayed ? 1 : 0;
        if (info) info->graphStatus = d.graphStatus;
        if (st != PHYX_B200_OK)
        {
            // the context's counts may be bounds: nothing sensible can continue from here
            c->jointUnitsValid = false;
            c->colourStateValid = false;
            c->strip.valid = false;
            d.ctlManifolds = d.ctlJoints = -1;
            graph_drop(c);
            return st;
        }
        d.steps++;
        StepCtl ctl;
        int header[16], colourResult[4], result[8];
        memcpy(&ctl, mailbox_at(c, 0), sizeof(ctl));
        memcpy(header, mailbox_at(c, 128), sizeof(header));
        memcpy(colourResult, mailbox_at(c, 192), sizeof(colourResult));
        memcpy(result, mailbox_at(c, 208), sizeof(result));
        if (ctl.stop)
        {
            // the device stopped before stage ctl.stop changed anything: exact counts, then the stage path from there
            d.stops++;
            d.lastStopStage = ctl.stop;
            d.lastStopReason = ctl.stopReason;
            d.ctlManifolds = d.ctlJoints = -1;    // (the device block was zeroed by the stop)
            c->manifoldCount = ctl.saved[1];
            c->contactPointCount = 2 * c->manifoldCount;
            c->jointCount = ctl.saved[2];
            c->broadphaseValid = true;            // (IntegratePosition did not run)
            c->jointUnitsValid = ctl.stop == kStageSolve;
            c->strip.valid = false;
            c->scheduleMode = -1;
            if (ctl.stop == kStageSolve && ctl.stopReason == 12) c->colourStateValid = false;   // colour overflow: the stage path rebuilds (and falls back)
            if (ctl.stop == kStagePairs)
            {
                d.lastItems = std::max(d.lastItems, ctl.stopReason == 1 ? ctl.stopNeed : d.lastItems);
                d.lastNewPairs = std::max(d.lastNewPairs, ctl.stopReason == 2 ? ctl.stopNeed : d.lastNewPairs);
            }
            if (ctl.stop == kStageRefresh) d.lastFresh = std::max(d.lastFresh, ctl.stopNeed);
            resumeFrom = ctl.stop == kStagePairs ? 2 : ctl.stop == kStageRefresh ? 5 : 6;
            // (the stopped step's PackManifolds rebuilt the cache from zero manifolds)
            if (ctl.stop == kStagePairs) PHYX_TRY(collide_rebuild_pair_table(c));
        }
        else
        {
            ranDeferred = true;
            d.ctlManifolds = ctl.manifolds;
            d.ctlJoints = ctl.joints;
            c->manifoldCount = ctl.manifolds;
            c->contactPointCount = 2 * ctl.manifolds;
            c->jointCount = ctl.joints;
            c->lastTests = (long long)ctl.tests;
            c->lastPairs = (long long)ctl.hits;
            c->lastNewPairs = ctl.newPairs;
            d.lastItems = ctl.items;
            d.lastNewPairs = ctl.newPairs;
            d.lastFresh = ctl.fresh;
            strip_apply_header(c, header);        // (usable: the device has checked it)
            strip_limit_recover(c);
            c->slotCount = 2 * c->strip.manifolds;
            c->levelCount = c->strip.colours;
            c->coloursInUse = c->strip.colours;
            c->colourRounds = colourResult[0];
            c->lastKernelForm = 3;
            long long act0 = 0, act1 = 0;
            memcpy(&act0, &result[4], 8);
            memcpy(&act1, &result[6], 8);
            const double nominal = double(c->jointCount) * double(result[0] > 0 ? result[0] : 1);
            c->lastActiveFraction = nominal > 0.0 ? float(double(act0) / nominal) : 1.0f;
            if (solveStats)
            {
                solveStats->joints = c->jointCount;
                solveStats->slots = c->slotCount;
                solveStats->levels = c->levelCount;
                solveStats->contactIterationsRun = result[0];
                solveStats->penetrationIterationsRun = result[1];
                solveStats->wakePasses = result[2];
                solveStats->colourRounds = c->colourRounds;
                solveStats->kernelForm = 3;
                solveStats->activeJointIterations[0] = act0;
                solveStats->activeJointIterations[1] = act1;
                solveStats->ms_schedule = elapsed_ms(c->ev[4], c->ev[5]);
                solveStats->ms_refresh = elapsed_ms(c->ev[0], c->ev[1]);
                solveStats->ms_iterations = elapsed_ms(c->ev[1], c->ev[2]);
                solveStats->ms_finish = elapsed_ms(c->ev[2], c->ev[3]);
                solveStats->ms_total = elapsed_ms(c->ev[4], c->ev[6]);
            }
            if (bpStats)
            {
                bpStats->tests = c->lastTests;
                bpStats->pairs = c->lastPairs;
                bpStats->ms_sort = c->sortTimed ? elapsed_ms(c->evBp[0], c->evBp[1]) : 0.f;
                bpStats->ms_sweep = elapsed_ms(c->evBp[2], c->evBp[3]);
                bpStats->ms_total = bpStats->ms_sort + bpStats->ms_sweep;
            }
            if (info)
            {
                info->newPairs = ctl.newPairs;
                info->jointsCreated = ctl.created;
                info->jointsDeleted = ctl.deleted;
            }
        }
        if (info)
        {
            info->stopStage = ctl.stop;
            info->stopReason = ctl.stopReason;
        }
    }
    else
        d.ineligible++;
    if (!ranDeferred) PHYX_TRY(run_stages(c, resumeFrom, dt, gravity, cfg, solveStats, bpStats));
    if (info)
    {
        info->deferred = ranDeferred ? 1 : 0;
        info->manifolds = c->manifoldCount;
        info->contactPoints = c->contactPointCount;
        info->joints = c->jointCount;
        info->pairs = c->lastPairs;
        info->tests = c->lastTests;
        info->deferredSteps = d.steps;
        info->deferredStops = d.stops;
    }
    return PHYX_B200_OK;
}

int phyx_b200_step_mode(phyx_b200_ctx* c, int deferred)
{
    PHYX_TRY(check(c));
    c->def.enabled = deferred != 0;
    c->def.tight = deferred == 2;
    c->def.useGraph = deferred != 3;
    return PHYX_B200_OK;
}

int phyx_b200_reset_collider(phyx_b200_ctx* c)
{
    PHYX_TRY(check(c));
    c->hostJointsValid = false;
    return collide_reset(c);
}

int phyx_b200_collider_counts(phyx_b200_ctx* c, int32_t* manifolds, int32_t* contactPoints, int32_t* joints)
{
    PHYX_TRY(check(c));
    if (manifolds) *manifolds = c->manifoldCount;
    if (contactPoints) *contactPoints = c->contactPointCount;
    if (joints) *joints = c->jointCount;
    return PHYX_B200_OK;
}

int phyx_b200_download_manifolds(phyx_b200_ctx* c, phyx_manifold* out, int capacity)
{
    PHYX_TRY(check(c));
    const int M = c->manifoldCount;
    if (capacity < M || (M > 0 && !out))
    {
        set_error("download_manifolds: capacity %d < %d", capacity, M);
        return PHYX_B200_ERR_CAPACITY;
    }
    if (M == 0) return PHYX_B200_OK;
    std::vector<int2> body(M);
    std::vector<int> count(M);
    PHYX_CUDA(cudaMemcpyAsync(body.data(), c->manBody.ptr, size_t(M) * sizeof(int2), cudaMemcpyDeviceToHost, c->stream));
    PHYX_CUDA(cudaMemcpyAsync(count.data(), c->manCount.ptr, size_t(M) * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PHYX_CUDA(cudaStreamSynchronize(c->stream));
    for (int m = 0; m < M; ++m)
    {
        out[m].body1Index = body[m].x;
        out[m].body2Index = body[m].y;
        out[m].pointCount = count[m];
        out[m].pointIndex = 2 * m;   // invariant of the reference's bookkeeping (Collider.cpp:315,405)
    }
    return PHYX_B200_OK;
}

int phyx_b200_download_contact_points(phyx_b200_ctx* c, phyx_contact_point* out, int capacity)
{
    PHYX_TRY(check(c));
    const int n = c->contactPointCount;
    if (capacity < n || (n > 0 && !out))
    {
        set_error("download_contact_points: capacity %d < %d", capacity, n);
        return PHYX_B200_ERR_CAPACITY;
    }
    if (n > 0)
    {
        PHYX_CUDA(cudaMemcpyAsync(out, c->contactPoints.ptr, size_t(n) * sizeof(phyx_contact_point), cudaMemcpyDeviceToHost, c->stream));
        PHYX_CUDA(cudaStreamSynchronize(c->stream));
    }
    return PHYX_B200_OK;
}

int phyx_b200_download_joints(phyx_b200_ctx* c, phyx_contact_joint* out, int capacity)
{
    PHYX_TRY(check(c));
    const int n = c->jointCount;
    if (capacity < n || (n > 0 && !out))
    {
        set_error("download_joints: capacity %d < %d", capacity, n);
        return PHYX_B200_ERR_CAPACITY;
    }
    if (n > 0)
    {
        PHYX_CUDA(cudaMemcpyAsync(out, c->joints.ptr, size_t(n) * sizeof(phyx_contact_joint), cudaMemcpyDeviceToHost, c->stream));
        PHYX_CUDA(cudaStreamSynchronize(c->stream));
    }
    return PHYX_B200_OK;
}

// Push a complete collider state (e.g. one produced by the reference, or edited by the caller).
int phyx_b200_upload_collider(phyx_b200_ctx* c, const phyx_manifold* manifolds, int manifoldCount, const phyx_contact_point* contactPoints,
    const phyx_contact_joint* joints, int jointCount)
{
    PHYX_TRY(check(c));
    if (manifoldCount < 0 || jointCount < 0 || (manifoldCount > 0 && (!manifolds || !contactPoints)) || (jointCount > 0 && !joints))
    {
        set_error("upload_collider: bad arguments");
        return PHYX_B200_ERR_ARGUMENT;
    }
    const int M = manifoldCount;
    std::vector<int2> body(size_t(M > 0 ? M : 1));
    std::vector<int> count(size_t(M > 0 ? M : 1));
    for (int m = 0; m < M; ++m)
    {
        if (manifolds[m].pointIndex != 2 * m || manifolds[m].pointCount < 0 || manifolds[m].pointCount > 2)
        {
            set_error("upload_collider: manifold %d breaks the pointIndex = 2*index / pointCount <= 2 invariant", m);
            return PHYX_B200_ERR_ARGUMENT;
        }
        body[m] = make_int2(manifolds[m].body1Index, manifolds[m].body2Index);
        count[m] = manifolds[m].pointCount;
    }
    c->manifoldCount = 0;
    c->jointCount = 0;
    PHYX_TRY(c->manBody.reserve(body.size() * sizeof(int2)));
    PHYX_TRY(c->manCount.reserve(count.size() * sizeof(int)));
    PHYX_TRY(c->manColour.reserve(count.size() * sizeof(int)));
    PHYX_TRY(c->contactPoints.reserve(size_t(M > 0 ? M : 1) * 2 * sizeof(phyx_contact_point)));
    PHYX_TRY(c->joints.reserve(size_t(jointCount > 0 ? jointCount : 1) * sizeof(phyx_contact_joint)));
    if (M > 0)
    {
        PHYX_CUDA(cudaMemcpyAsync(c->manBody.ptr, body.data(), size_t(M) * sizeof(int2), cudaMemcpyHostToDevice, c->stream));
        PHYX_CUDA(cudaMemcpyAsync(c->manCount.ptr, count.data(), size_t(M) * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        PHYX_CUDA(cudaMemcpyAsync(c->contactPoints.ptr, contactPoints, size_t(M) * 2 * sizeof(phyx_contact_point), cudaMemcpyHostToDevice, c->stream));
    }
    if (jointCount > 0)
        PHYX_CUDA(cudaMemcpyAsync(c->joints.ptr, joints, size_t(jointCount) * sizeof(phyx_contact_joint), cudaMemcpyHostToDevice, c->stream));
    PHYX_CUDA(cudaStreamSynchronize(c->stream));
    c->manifoldCount = M;
    c->contactPointCount = 2 * M;
    c->jointCount = jointCount;
    c->hostJointsValid = false;
    c->colourStateValid = false;
    c->jointUnitsValid = false;
    return collide_rebuild_pair_table(c);
}

int phyx_b200_fetch_joints(phyx_b200_ctx* c, phyx_contact_joint* joints, int jointCount)
{
    PHYX_TRY(check(c));
    if (jointCount != c->jointCount || (jointCount > 0 && !joints))
    {
        set_error("fetch_joints: count %d does not match the %d resident joints", jointCount, c->jointCount);
        return PHYX_B200_ERR_ARGUMENT;
    }
    if (jointCount > 0)
    {
        PHYX_CUDA(cudaMemcpyAsync(joints, c->joints.ptr, size_t(jointCount) * sizeof(phyx_contact_joint), cudaMemcpyDeviceToHost, c->stream));
        PHYX_CUDA(cudaStreamSynchronize(c->stream));
    }
    return PHYX_B200_OK;
}

int phyx_b200_solve_joints(phyx_b200_ctx* c, phyx_contact_joint* joints, int jointCount, const phyx_contact_point* contactPoints,
    int contactPointCount, const phyx_b200_solve_config* cfg, phyx_b200_solve_stats* stats)
{
    PHYX_TRY(check(c));
    if (!cfg)
    {
        set_error("solve: null config");
        return PHYX_B200_ERR_ARGUMENT;
    }
    if (stats) memset(stats, 0, sizeof(*stats));
    cudaEvent_t t0 = c->ev[4], t1 = c->ev[5], t2 = c->ev[6], t3 = c->ev[7];
    PHYX_CUDA(cudaEventRecord(t0, c->stream));
    PHYX_TRY(stage_joints(c, joints, jointCount, contactPoints, contactPointCount));
    PHYX_CUDA(cudaEventRecord(t1, c->stream));
    PHYX_TRY(schedule_build(c, joints, jointCount, cfg->schedule, cfg->flags));
    PHYX_CUDA(cudaEventRecord(t2, c->stream));
    PHYX_TRY(solve_run(c, cfg, stats));
    PHYX_CUDA(cudaEventRecord(t3, c->stream));
    float h2d = 0.f, sched = 0.f;
    if (jointCount > 0)
        PHYX_CUDA(cudaMemcpyAsync(joints, c->joints.ptr, size_t(jointCount) * sizeof(phyx_contact_joint), cudaMemcpyDeviceToHost, c->stream));
    PHYX_CUDA(cudaEventRecord(c->ev[0], c->stream));
    PHYX_CUDA(cudaEventSynchronize(c->ev[0]));
    h2d = elapsed_ms(t0, t1);
    sched = elapsed_ms(t1, t2);
    if (stats)
    {
        stats->ms_h2d = h2d;
        stats->ms_schedule = sched;
        stats->ms_d2h = elapsed_ms(t3, c->ev[0]);
        stats->ms_total = elapsed_ms(t0, c->ev[0]);
    }
    return PHYX_B200_OK;
}

int phyx_b200_get_schedule(phyx_b200_ctx* c, int32_t* slots, int32_t slotCapacity, int32_t* levels3, int32_t levelCapacity, int32_t* slotCount,
    int32_t* levelCount)
{
    PHYX_TRY(check(c));
    if (c->hostLevelsStale) PHYX_TRY(strip_host_levels(c, nullptr));
    if (c->hostSlotsStale)
    {
        // device-built schedule: fetch the slot -> joint table on demand
        c->hostSlots.resize(size_t(c->slotCount));
        if (c->slotCount)
        {
            PHYX_CUDA(cudaMemcpyAsync(c->hostSlots.data(), c->slotJoint.ptr, size_t(c->slotCount) * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            PHYX_CUDA(cudaStreamSynchronize(c->stream));
        }
        c->hostSlotsStale = false;
    }
    int ns = int(c->hostSlots.size()), nl = int(c->hostLevels.size());
    if (slotCount) *slotCount = ns;
    if (levelCount) *levelCount = nl;
    if (!slots && !levels3) return PHYX_B200_OK;
    if (slotCapacity < ns || levelCapacity < nl)
    {
        set_error("get_schedule: need %d slots and %d levels", ns, nl);
        return PHYX_B200_ERR_CAPACITY;
    }
    if (slots && ns) memcpy(slots, c->hostSlots.data(), size_t(ns) * sizeof(int));
    if (levels3 && nl) memcpy(levels3, c->hostLevels.data(), size_t(nl) * sizeof(Level));
    return PHYX_B200_OK;
}

} // extern "C"
