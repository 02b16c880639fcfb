// phyx_b200 — body state in HBM: AoS <-> SoA conversion and the two integration kernels.
//
// Reference stages replaced here:
//   World::IntegrateVelocity   src/World.cpp:39-55
//   World::IntegratePosition   src/World.cpp:57-70  (+ Vector2::Rotate src/Vector2.h:48-56,
//                                                     Geom::RecomputeAABB src/Geom.h:79-85)
//
// Layout (DESIGN.md "HBM layout"): one float4 (or float2) array per field group, one thread per
// body, every access a full-width coalesced vector load/store.  All arithmetic is written with
// the reference's operation order and compiled with --fmad=false so results are bit-equal to the
// reference's strict-FP build.
#include "common.cuh"

namespace phyx
{

constexpr int kBlock = 256;

__device__ __forceinline__ float as_f(int v) { return __int_as_float(v); }

// One RigidBody record = 32 floats = 8 float4.
__global__ void k_bodies_unpack(const float4* __restrict__ aos, int n, float4* __restrict__ vel, float4* __restrict__ disp,
    float4* __restrict__ acc, float4* __restrict__ params, float4* __restrict__ rot, float4* __restrict__ aabb, float2* __restrict__ size)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4* r = aos + size_t(i) * 8;
    float4 f0 = r[0]; // index, size.x, size.y, gxV.x
    float4 f2 = r[2]; // gpos.y, aabbmin.x, aabbmin.y, aabbmax.x
    float4 f3 = r[3]; // aabbmax.y, vel.x, vel.y, acc.x
    float4 f4 = r[4]; // acc.y, dvel.x, dvel.y, angVel
    float4 f5 = r[5]; // angAcc, dAngVel, invMass, invInertia
    float4 f6 = r[6]; // xV.x, xV.y, yV.x, yV.y
    float4 f7 = r[7]; // pos.x, pos.y, lastIt, lastDispIt
    size[i] = make_float2(f0.y, f0.z);
    aabb[i] = make_float4(f2.y, f2.z, f2.w, f3.x);
    vel[i] = make_float4(f3.y, f3.z, f4.w, as_f(-1));
    acc[i] = make_float4(f3.w, f4.x, f5.x, 0.f);
    disp[i] = make_float4(f4.y, f4.z, f5.y, as_f(-1));
    params[i] = make_float4(f5.z, f5.w, f7.x, f7.y);
    rot[i] = f6;
}

__global__ void k_bodies_pack(float4* __restrict__ aos, int n, const float4* __restrict__ vel, const float4* __restrict__ disp,
    const float4* __restrict__ acc, const float4* __restrict__ params, const float4* __restrict__ rot, const float4* __restrict__ aabb,
    const float2* __restrict__ size)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 v = vel[i], d = disp[i], a = acc[i], p = params[i], r = rot[i], bb = aabb[i];
    float2 s = size[i];
    float4* o = aos + size_t(i) * 8;
    o[0] = make_float4(__uint_as_float(unsigned(i)), s.x, s.y, r.x);   // index, size, geom.coords.xVector.x
    o[1] = make_float4(r.y, r.z, r.w, p.z);                            // geom xV.y, yV, pos.x
    o[2] = make_float4(p.w, bb.x, bb.y, bb.z);                         // geom pos.y, aabb
    o[3] = make_float4(bb.w, v.x, v.y, a.x);
    o[4] = make_float4(a.y, d.x, d.y, v.z);
    o[5] = make_float4(a.z, d.z, p.x, p.y);
    o[6] = r;
    o[7] = make_float4(p.z, p.w, as_f(0), as_f(0));
}

// World::IntegrateVelocity, src/World.cpp:43-54
__global__ void k_integrate_velocity(int n, float dt, float gravity, float4* __restrict__ vel, float4* __restrict__ acc,
    const float4* __restrict__ params)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 v = vel[i], a = acc[i];
    float invMass = params[i].x;
    if (invMass > 0.0f)
        a.y = a.y + gravity;
    v.x = v.x + a.x * dt;
    v.y = v.y + a.y * dt;
    v.z = v.z + a.z * dt;
    vel[i] = v;
    acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// Vector2::Rotate, src/Vector2.h:48-56 (cos/sin evaluated in double, narrowed to float)
__device__ __forceinline__ void rotate_vec(float& x, float& y, float c, float s)
{
    float xx = x, xy = y;
    float yx = -xy, yy = xx;
    float dx = (xx * c + yx * s) - xx;
    float dy = (xy * c + yy * s) - xy;
    x = xx + dx;
    y = xy + dy;
}

// World::IntegratePosition, src/World.cpp:61-69
__global__ void k_integrate_position(Count nc, float dt, const float4* __restrict__ vel, float4* __restrict__ disp,
    float4* __restrict__ params, float4* __restrict__ rot, float4* __restrict__ aabb, const float2* __restrict__ size)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count_of(nc)) return;   // (a stopped deferred step must not move the bodies: the host finishes it first)
    float4 v = vel[i], d = disp[i], p = params[i], r = rot[i];
    float2 sz = size[i];
    float mx = d.x + v.x * dt;
    float my = d.y + v.y * dt;
    p.z = p.z + mx;
    p.w = p.w + my;
    float angle = -(d.z + v.z * dt);
    double sd, cd;
    sincos(double(angle), &sd, &cd);
    float c = float(cd), s = float(sd);
    rotate_vec(r.x, r.y, c, s);
    rotate_vec(r.z, r.w, c, s);
    float ex = fabsf(r.x) * sz.x + fabsf(r.z) * sz.y;
    float ey = fabsf(r.y) * sz.x + fabsf(r.w) * sz.y;
    params[i] = p;
    rot[i] = r;
    disp[i] = make_float4(0.f, 0.f, 0.f, as_f(-1));
    aabb[i] = make_float4(p.z - ex, p.w - ey, p.z + ex, p.w + ey);
}

// AABB of all dynamic bodies: {min.x, min.y, max.x, max.y} as order-preserving unsigned keys (atomicMin / atomicMax)
__device__ __forceinline__ unsigned float_key(float v)
{
    const unsigned u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__global__ void k_dynamic_extent(int n, const float4* __restrict__ aabb, const float4* __restrict__ params, unsigned* __restrict__ out4)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned k[4] = { 0xffffffffu, 0xffffffffu, 0u, 0u };
    if (i < n)
    {
        const float4 p = params[i];
        if (!(p.x == 0.0f && p.y == 0.0f))
        {
            const float4 a = aabb[i];
            k[0] = float_key(a.x);
            k[1] = float_key(a.y);
            k[2] = float_key(a.z);
            k[3] = float_key(a.w);
        }
    }
    for (int o = 16; o > 0; o >>= 1)
    {
        k[0] = min(k[0], __shfl_xor_sync(0xffffffffu, k[0], o));
        k[1] = min(k[1], __shfl_xor_sync(0xffffffffu, k[1], o));
        k[2] = max(k[2], __shfl_xor_sync(0xffffffffu, k[2], o));
        k[3] = max(k[3], __shfl_xor_sync(0xffffffffu, k[3], o));
    }
    if ((threadIdx.x & 31) == 0)
    {
        atomicMin(&out4[0], k[0]);
        atomicMin(&out4[1], k[1]);
        atomicMax(&out4[2], k[2]);
        atomicMax(&out4[3], k[3]);
    }
}

int bodies_dynamic_extent(phyx_b200_ctx* c, float* out4)
{
    const int n = c->bodyCount;
    PHYX_TRY(c->counters.reserve(64));
    unsigned* d = c->counters.as<unsigned>() + 8;
    const unsigned init[4] = { 0xffffffffu, 0xffffffffu, 0u, 0u };
    PHYX_CUDA(cudaMemcpyAsync(d, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
    if (n > 0)
    {
        k_dynamic_extent<<<(n + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(n, c->aabb.as<float4>(), c->params.as<float4>(), d);
        c->launches++;
    }
    unsigned host[4];
    PHYX_CUDA(cudaMemcpyAsync(host, d, sizeof(host), cudaMemcpyDeviceToHost, c->stream));
    PHYX_CUDA(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 4; ++k)
    {
        const unsigned u = host[k];
        const unsigned bits = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
        memcpy(&out4[k], &bits, 4);
    }
    return PHYX_B200_OK;
}

static int reserve_bodies(phyx_b200_ctx* c, int n)
{
    size_t n4 = size_t(n > 0 ? n : 1) * sizeof(float4);
    PHYX_TRY(c->vel.reserve(n4));
    PHYX_TRY(c->disp.reserve(n4));
    PHYX_TRY(c->acc.reserve(n4));
    PHYX_TRY(c->params.reserve(n4));
    PHYX_TRY(c->rot.reserve(n4));
    PHYX_TRY(c->aabb.reserve(n4));
    PHYX_TRY(c->size.reserve(n4 / 2));
    PHYX_TRY(c->aos.reserve(size_t(n > 0 ? n : 1) * sizeof(phyx_rigid_body)));
    return PHYX_B200_OK;
}

int bodies_upload(phyx_b200_ctx* c, const phyx_rigid_body* bodies, int n, bool wait)
{
    PHYX_TRY(reserve_bodies(c, n));
    c->bodyCount = n;
    c->broadphaseValid = false;
    c->rowOrderValid = false;
    c->hasSnapshot = false;
    if (n == 0) return PHYX_B200_OK;
    PHYX_CUDA(cudaMemcpyAsync(c->aos.ptr, bodies, size_t(n) * sizeof(phyx_rigid_body), cudaMemcpyHostToDevice, c->stream));
    k_bodies_unpack<<<(n + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(c->aos.as<float4>(), n, c->vel.as<float4>(), c->disp.as<float4>(),
        c->acc.as<float4>(), c->params.as<float4>(), c->rot.as<float4>(), c->aabb.as<float4>(), c->size.as<float2>());
    c->launches++;
    PHYX_CUDA(cudaGetLastError());
    // the source may be pageable memory that the caller mutates right after: finish the copy now (unless the caller has
    // promised to leave the array alone until its next synchronising call: phyx_b200_upload_bodies_async)
    if (wait) PHYX_CUDA(cudaStreamSynchronize(c->stream));
    return PHYX_B200_OK;
}

int bodies_download(phyx_b200_ctx* c, phyx_rigid_body* bodies, int n)
{
    if (n == 0) return PHYX_B200_OK;
    k_bodies_pack<<<(n + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(c->aos.as<float4>(), n, c->vel.as<float4>(), c->disp.as<float4>(),
        c->acc.as<float4>(), c->params.as<float4>(), c->rot.as<float4>(), c->aabb.as<float4>(), c->size.as<float2>());
    c->launches++;
    PHYX_CUDA(cudaGetLastError());
    PHYX_CUDA(cudaMemcpyAsync(bodies, c->aos.ptr, size_t(n) * sizeof(phyx_rigid_body), cudaMemcpyDeviceToHost, c->stream));
    PHYX_CUDA(cudaStreamSynchronize(c->stream));
    return PHYX_B200_OK;
}

int bodies_integrate_velocity(phyx_b200_ctx* c, float dt, float gravity)
{
    int n = c->bodyCount;
    if (n == 0) return PHYX_B200_OK;
    k_integrate_velocity<<<(n + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(n, dt, gravity, c->vel.as<float4>(), c->acc.as<float4>(),
        c->params.as<float4>());
    c->launches++;
    PHYX_CUDA(cudaGetLastError());
    return PHYX_B200_OK;
}

int bodies_integrate_position(phyx_b200_ctx* c, float dt)
{
    int n = c->bodyCount;
    if (n == 0) return PHYX_B200_OK;
    k_integrate_position<<<(n + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(c->count(n, &StepCtl::bodies), dt, c->vel.as<float4>(), c->disp.as<float4>(),
        c->params.as<float4>(), c->rot.as<float4>(), c->aabb.as<float4>(), c->size.as<float2>());
    c->launches++;
    c->broadphaseValid = false;
    PHYX_CUDA(cudaGetLastError());
    return PHYX_B200_OK;
}

int bodies_snapshot(phyx_b200_ctx* c, bool restore)
{
    int n = c->bodyCount;
    size_t n4 = size_t(n) * sizeof(float4);
    if (!restore)
    {
        PHYX_TRY(c->snap.reserve(6 * (n4 > 0 ? n4 : 16)));
        c->hasSnapshot = true;
    }
    else if (!c->hasSnapshot)
    {
        set_error("restore_bodies without snapshot_bodies");
        return PHYX_B200_ERR_STATE;
    }
    DevBuf* bufs[6] = { &c->vel, &c->disp, &c->acc, &c->params, &c->rot, &c->aabb };
    for (int k = 0; k < 6 && n > 0; ++k)
    {
        char* s = c->snap.as<char>() + size_t(k) * n4;
        if (restore)
            PHYX_CUDA(cudaMemcpyAsync(bufs[k]->ptr, s, n4, cudaMemcpyDeviceToDevice, c->stream));
        else
            PHYX_CUDA(cudaMemcpyAsync(s, bufs[k]->ptr, n4, cudaMemcpyDeviceToDevice, c->stream));
    }
    // staged joints (cached impulses are updated in place by solve_staged) travel with the snapshot
    size_t jb = size_t(c->jointCount) * sizeof(phyx_contact_joint);
    if (!restore)
    {
        c->snapJointCount = c->jointCount;
        if (jb)
        {
            PHYX_TRY(c->snapJoints.reserve(jb));
            PHYX_CUDA(cudaMemcpyAsync(c->snapJoints.ptr, c->joints.ptr, jb, cudaMemcpyDeviceToDevice, c->stream));
        }
    }
    else if (jb && c->snapJointCount == c->jointCount)
        PHYX_CUDA(cudaMemcpyAsync(c->joints.ptr, c->snapJoints.ptr, jb, cudaMemcpyDeviceToDevice, c->stream));
    if (restore) c->broadphaseValid = false;
    return PHYX_B200_OK;
}

} // namespace phyx
