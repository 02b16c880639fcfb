/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See phyx_oracle.h.
 *
 * Plain-C scalar restatement of the reference hot path.  Compile with
 *   gcc -std=c11 -O2 -ffp-contract=off -fno-fast-math
 * (no FMA contraction, no reassociation): every float operation below is written in the order
 * the reference performs it, so the result is bit-equal to the reference's strict-FP build
 * (oracle/_ref/libphyx_ref_strict.so), which tests/test_oracle_vs_reference.py verifies.
 */
#include "phyx_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

_Static_assert(sizeof(pxo_body) == 128, "RigidBody is 128 B");
_Static_assert(sizeof(pxo_joint) == 20, "ContactJoint is 20 B");
_Static_assert(sizeof(pxo_contact_point) == 32, "ContactPoint is 32 B");
_Static_assert(sizeof(pxo_broadphase_entry) == 20, "BroadphaseEntry is 20 B");

/* ------------------------------------------------------------------------------------------ */
/* a2: World::IntegrateVelocity, src/World.cpp:43-54                                           */
void pxo_integrate_velocity(pxo_body* bodies, int n, float dt, float gravity)
{
    for (int i = 0; i < n; ++i)
    {
        pxo_body* b = &bodies[i];
        if (b->invMass > 0.0f)
            b->acceleration.y += gravity;
        b->velocity.x += b->acceleration.x * dt;
        b->velocity.y += b->acceleration.y * dt;
        b->acceleration.x = 0.0f;
        b->acceleration.y = 0.0f;
        b->angularVelocity += b->angularAcceleration * dt;
        b->angularAcceleration = 0.0f;
    }
}

/* Vector2::Rotate, src/Vector2.h:48-56.  `cos(angle)` there is the double overload (the call is
 * unqualified in the global namespace), its result is narrowed to float by operator*(const T&). */
static void rotate_vec(pxo_vec2* v, float angle)
{
    float c = (float)cos((double)angle);
    float s = (float)sin((double)angle);
    float xx = v->x, xy = v->y;   /* x = self           */
    float yx = -xy, yy = xx;      /* y = (-x.y, x.x)    */
    float dx = (xx * c + yx * s) - xx;
    float dy = (xy * c + yy * s) - xy;
    v->x = xx + dx;
    v->y = xy + dy;
}

/* a13: World::IntegratePosition, src/World.cpp:61-69; RigidBody::UpdateGeom, src/RigidBody.h:38-42;
 *      Geom::RecomputeAABB, src/Geom.h:79-85 */
void pxo_integrate_position(pxo_body* bodies, int n, float dt)
{
    for (int i = 0; i < n; ++i)
    {
        pxo_body* b = &bodies[i];
        float mx = b->displacingVelocity.x + b->velocity.x * dt;
        float my = b->displacingVelocity.y + b->velocity.y * dt;
        b->pos.x += mx;
        b->pos.y += my;
        float angle = -(b->displacingAngularVelocity + b->angularVelocity * dt);
        rotate_vec(&b->xVector, angle);
        rotate_vec(&b->yVector, angle);
        b->displacingVelocity.x = 0.0f;
        b->displacingVelocity.y = 0.0f;
        b->displacingAngularVelocity = 0.0f;
        b->geom_xVector = b->xVector;
        b->geom_yVector = b->yVector;
        b->geom_pos = b->pos;
        float ex = fabsf(b->xVector.x) * b->size.x + fabsf(b->yVector.x) * b->size.y;
        float ey = fabsf(b->xVector.y) * b->size.x + fabsf(b->yVector.y) * b->size.y;
        b->aabb_min.x = b->pos.x - ex;
        b->aabb_min.y = b->pos.y - ey;
        b->aabb_max.x = b->pos.x + ex;
        b->aabb_max.y = b->pos.y + ey;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* a3: radixFloat, src/base/RadixSort.h:19-26 */
uint32_t pxo_radix_float(float v)
{
    int32_t f;
    memcpy(&f, &v, 4);
    uint32_t mask = (uint32_t)(f >> 31) | 0x80000000u;
    return (uint32_t)f ^ mask;
}

/* a3: radixSort3, src/base/RadixSort.h:28-95: one histogram pass over three digits
 * (11/11/10 bits), exclusive prefix sums, three stable scatter passes; the sorted run ends up in
 * the second buffer, copied back here. */
void pxo_radix_sort3(uint32_t* kv, int n)
{
    uint32_t* hist = (uint32_t*)calloc(3 * 2048, sizeof(uint32_t));
    uint32_t* tmp = (uint32_t*)malloc((size_t)(n > 0 ? n : 1) * 8);
    uint32_t *h0 = hist, *h1 = hist + 2048, *h2 = hist + 4096;
    for (int i = 0; i < n; ++i)
    {
        uint32_t k = kv[2 * i];
        h0[k & 2047]++;
        h1[(k >> 11) & 2047]++;
        h2[k >> 22]++;
    }
    uint32_t s0 = 0, s1 = 0, s2 = 0;
    for (int i = 0; i < 2048; ++i)
    {
        uint32_t c0 = h0[i], c1 = h1[i], c2 = h2[i];
        h0[i] = s0; h1[i] = s1; h2[i] = s2;
        s0 += c0; s1 += c1; s2 += c2;
    }
    for (int i = 0; i < n; ++i)
    {
        uint32_t d = h0[kv[2 * i] & 2047]++;
        tmp[2 * d] = kv[2 * i]; tmp[2 * d + 1] = kv[2 * i + 1];
    }
    for (int i = 0; i < n; ++i)
    {
        uint32_t d = h1[(tmp[2 * i] >> 11) & 2047]++;
        kv[2 * d] = tmp[2 * i]; kv[2 * d + 1] = tmp[2 * i + 1];
    }
    for (int i = 0; i < n; ++i)
    {
        uint32_t d = h2[kv[2 * i] >> 22]++;
        tmp[2 * d] = kv[2 * i]; tmp[2 * d + 1] = kv[2 * i + 1];
    }
    memcpy(kv, tmp, (size_t)n * 8);
    free(tmp);
    free(hist);
}

/* a3: Collider::UpdateBroadphase, src/Collider.cpp:251-284 */
void pxo_update_broadphase(const pxo_body* bodies, int n, pxo_broadphase_entry* out)
{
    uint32_t* kv = (uint32_t*)malloc((size_t)(n > 0 ? n : 1) * 8);
    for (int i = 0; i < n; ++i)
    {
        kv[2 * i] = pxo_radix_float(bodies[i].aabb_min.x);
        kv[2 * i + 1] = (uint32_t)i;
    }
    pxo_radix_sort3(kv, n);
    for (int i = 0; i < n; ++i)
    {
        uint32_t bi = kv[2 * i + 1];
        const pxo_body* b = &bodies[bi];
        out[i].minx = b->aabb_min.x;
        out[i].maxx = b->aabb_max.x;
        out[i].centery = (b->aabb_min.y + b->aabb_max.y) * 0.5f;
        out[i].extenty = (b->aabb_max.y - b->aabb_min.y) * 0.5f;
        out[i].index = bi;
    }
    free(kv);
}

/* a4: sweep of Collider::UpdatePairsSerial, src/Collider.cpp:298-319 (manifold map left out) */
long long pxo_sweep_pairs(const pxo_broadphase_entry* e, int n, int* pairs, long long cap, long long* tests)
{
    long long count = 0, t = 0;
    for (int i = 0; i < n; ++i)
    {
        float maxx = e[i].maxx;
        for (int j = i + 1; j < n; ++j)
        {
            if (e[j].minx > maxx)
                break;
            t++;
            if (fabsf(e[j].centery - e[i].centery) <= e[i].extenty + e[j].extenty)
            {
                if (count < cap)
                {
                    pairs[2 * count] = (int)e[i].index;
                    pairs[2 * count + 1] = (int)e[j].index;
                }
                count++;
            }
        }
    }
    if (tests) *tests = t;
    return count;
}

/* ------------------------------------------------------------------------------------------ */
/* a6: Solver::PrepareIndices, src/Solver.cpp:217-273 (joint_index starts as identity,
 *     jointGroup_bodies zeroed: Solver.cpp:95-106) */
int pxo_prepare_indices(const pxo_joint* joints, int nj, int nbodies, int group, int* joint_index)
{
    for (int i = 0; i < nj; ++i) joint_index[i] = i;
    if (group == 1)
        return nj;

    int* pool = (int*)malloc((size_t)(nj > 0 ? nj : 1) * sizeof(int));
    int* body_tag = (int*)calloc((size_t)(nbodies > 0 ? nbodies : 1), sizeof(int));
    for (int i = 0; i < nj; ++i) pool[i] = joint_index[i];

    int tag = 0, remaining = nj, offset = 0;
    while (remaining >= group)
    {
        int taken = 0;
        tag++;
        for (int i = 0; i < remaining && taken < group;)
        {
            int j = pool[i];
            int b1 = joints[j].body1Index, b2 = joints[j].body2Index;
            if (body_tag[b1] < tag && body_tag[b2] < tag)
            {
                body_tag[b1] = tag;
                body_tag[b2] = tag;
                joint_index[offset + taken] = j;
                taken++;
                pool[i] = pool[remaining - 1];
                remaining--;
            }
            else
                i++;
        }
        offset += taken;
        if (taken < group)
            break;
    }
    for (int i = 0; i < remaining; ++i)
        joint_index[offset + i] = pool[i];
    free(pool);
    free(body_tag);
    return offset & ~(group - 1);
}

/* ------------------------------------------------------------------------------------------ */
/* Solver data, as the reference packs it */
typedef struct { float vx, vy, w; int last; } solve_body;               /* Solver.h:97-103 */
typedef struct { float invMass, invInertia, px, py; } solve_params;      /* Solver.h:86-95 (used part) */

typedef struct {                                                          /* Solver.h:7-24 */
    float p1x, p1y, p2x, p2y, a1, a2;
    float cm1x, cm1y, cm2x, cm2y, cm1a, cm2a;
    float cinv;
} limiter;

typedef struct {                                                          /* Solver.h:26-45 */
    int b1, b2, cp;
    limiter n;
    float accN, dstVel, dstDisp, accD;
    limiter f;
    float accF;
} packed_joint;

static float vmax(float l, float r) { return l > r ? l : r; }            /* SIMD_*: max = l>r?l:r */

/* RefreshLimiter, src/Solver.cpp:549-590 */
static void refresh_limiter(limiter* L, float n1x, float n1y, float n2x, float n2y, float w1x, float w1y,
    float w2x, float w2y, float im1, float ii1, float im2, float ii2)
{
    L->p1x = n1x; L->p1y = n1y; L->p2x = n2x; L->p2y = n2y;
    L->a1 = n1x * w1y - n1y * w1x;
    L->a2 = n2x * w2y - n2y * w2x;
    L->cm1x = L->p1x * im1;
    L->cm1y = L->p1y * im1;
    L->cm1a = L->a1 * ii1;
    L->cm2x = L->p2x * im2;
    L->cm2y = L->p2y * im2;
    L->cm2a = L->a2 * ii2;
    float c1 = L->p1x * L->cm1x + L->p1y * L->cm1y + L->a1 * L->cm1a;
    float c2 = L->p2x * L->cm2x + L->p2y * L->cm2y + L->a2 * L->cm2a;
    float c = c1 + c2;
    L->cinv = (fabsf(c) > 0.0f) ? 1.0f / c : 0.0f;
}

/* RefreshJoints, src/Solver.cpp:592-695 */
static void refresh_joint(packed_joint* j, const solve_body* imp, const solve_params* par, const pxo_contact_point* cps)
{
    const solve_body *v1 = &imp[j->b1], *v2 = &imp[j->b2];
    const solve_params *q1 = &par[j->b1], *q2 = &par[j->b2];
    const pxo_contact_point* c = &cps[j->cp];

    float p1x = c->delta1.x + q1->px, p1y = c->delta1.y + q1->py;
    float p2x = c->delta2.x + q2->px, p2y = c->delta2.y + q2->py;
    float w1x = c->delta1.x, w1y = c->delta1.y;
    float w2x = p1x - q2->px, w2y = p1y - q2->py;
    float nx = c->normal.x, ny = c->normal.y;

    refresh_limiter(&j->n, nx, ny, -nx, -ny, w1x, w1y, w2x, w2y, q1->invMass, q1->invInertia, q2->invMass, q2->invInertia);

    float bounce = 0.0f, deltaVelocity = 1.0f, maxPenetrationVelocity = 0.1f, deltaDepth = 1.0f, errorReduction = 0.1f;

    float pv1x = (q1->py - p1y) * v1->w + v1->vx;
    float pv1y = (p1x - q1->px) * v1->w + v1->vy;
    float pv2x = (q2->py - p2y) * v2->w + v2->vx;
    float pv2y = (p2x - q2->px) * v2->w + v2->vy;
    float rvx = pv1x - pv2x, rvy = pv1y - pv2y;

    float dv = -bounce * (rvx * nx + rvy * ny);
    float depth = (p2x - p1x) * nx + (p2y - p1y) * ny;
    float dstVelocity = vmax(dv - deltaVelocity, 0.0f);

    j->dstVel = (depth < deltaDepth) ? dstVelocity - maxPenetrationVelocity : dstVelocity;
    j->dstDisp = errorReduction * vmax(0.0f, depth - 2.0f * deltaDepth);
    j->accD = 0.0f;

    float tx = -ny, ty = nx;
    refresh_limiter(&j->f, tx, ty, -tx, -ty, w1x, w1y, w2x, w2y, q1->invMass, q1->invInertia, q2->invMass, q2->invInertia);
}

/* PreStepJoints, src/Solver.cpp:697-758 */
static void prestep_joint(const packed_joint* j, solve_body* imp)
{
    solve_body *b1 = &imp[j->b1], *b2 = &imp[j->b2];
    b1->vx += j->n.cm1x * j->accN;
    b1->vy += j->n.cm1y * j->accN;
    b1->w += j->n.cm1a * j->accN;
    b2->vx += j->n.cm2x * j->accN;
    b2->vy += j->n.cm2y * j->accN;
    b2->w += j->n.cm2a * j->accN;
    b1->vx += j->f.cm1x * j->accF;
    b1->vy += j->f.cm1y * j->accF;
    b1->w += j->f.cm1a * j->accF;
    b2->vx += j->f.cm2x * j->accF;
    b2->vy += j->f.cm2y * j->accF;
    b2->w += j->f.cm2a * j->accF;
}

static float flipsign_bits(float x, float y)        /* SIMD_AVX2.h:272-275 / SIMD_SSE2: xor sign bit */
{
    uint32_t xb, yb;
    memcpy(&xb, &x, 4); memcpy(&yb, &y, 4);
    xb ^= (yb & 0x80000000u);
    memcpy(&x, &xb, 4);
    return x;
}
static float flipsign_cmp(float x, float y) { return y < 0.0f ? -x : x; }   /* SIMD_Scalar.h:265-268 */

/* SolveJointsImpulses body for one joint, src/Solver.cpp:833-901.  v1/v2 are working copies of
 * the two SolveBody rows.  Returns productive. */
static int impulse_joint(packed_joint* j, solve_body* v1, solve_body* v2, int wide)
{
    float dV = j->dstVel;
    dV -= j->n.p1x * v1->vx;
    dV -= j->n.p1y * v1->vy;
    dV -= j->n.a1 * v1->w;
    dV -= j->n.p2x * v2->vx;
    dV -= j->n.p2y * v2->vy;
    dV -= j->n.a2 * v2->w;

    float dN = dV * j->n.cinv;
    dN = vmax(dN, -j->accN);

    v1->vx += j->n.cm1x * dN;
    v1->vy += j->n.cm1y * dN;
    v1->w += j->n.cm1a * dN;
    v2->vx += j->n.cm2x * dN;
    v2->vy += j->n.cm2y * dN;
    v2->w += j->n.cm2a * dN;
    j->accN += dN;

    float fV = 0.0f;
    fV -= j->f.p1x * v1->vx;
    fV -= j->f.p1y * v1->vy;
    fV -= j->f.a1 * v1->w;
    fV -= j->f.p2x * v2->vx;
    fV -= j->f.p2y * v2->vy;
    fV -= j->f.a2 * v2->w;

    float dF = fV * j->f.cinv;
    float reaction = j->accN;
    float acc = j->accF;
    float force = acc + dF;
    float limit = reaction * 0.3f;                              /* kFrictionCoefficient */
    float forceAbs = fabsf(force);
    float limitSigned = wide ? flipsign_bits(limit, force) : flipsign_cmp(limit, force);
    float adjusted = limitSigned - acc;
    dF = (forceAbs > limit) ? adjusted : dF;
    j->accF += dF;

    v1->vx += j->f.cm1x * dF;
    v1->vy += j->f.cm1y * dF;
    v1->w += j->f.cm1a * dF;
    v2->vx += j->f.cm2x * dF;
    v2->vy += j->f.cm2y * dF;
    v2->w += j->f.cm2a * dF;

    float cumulative = vmax(fabsf(dN), fabsf(dF));
    return cumulative > 1e-4f;                                  /* kProductiveImpulse */
}

/* SolveJointsDisplacement body for one joint, src/Solver.cpp:971-1003 */
static int displacement_joint(packed_joint* j, solve_body* v1, solve_body* v2)
{
    float dV = j->dstDisp;
    dV -= j->n.p1x * v1->vx;
    dV -= j->n.p1y * v1->vy;
    dV -= j->n.a1 * v1->w;
    dV -= j->n.p2x * v2->vx;
    dV -= j->n.p2y * v2->vy;
    dV -= j->n.a2 * v2->w;

    float d = dV * j->n.cinv;
    d = vmax(d, -j->accD);

    v1->vx += j->n.cm1x * d;
    v1->vy += j->n.cm1y * d;
    v1->w += j->n.cm1a * d;
    v2->vx += j->n.cm2x * d;
    v2->vy += j->n.cm2y * d;
    v2->w += j->n.cm2a * d;
    j->accD += d;

    return fabsf(d) > 1e-4f;
}

/* ------------------------------------------------------------------------------------------ */
/* One pass (phase 0 = impulses, 1 = displacement) of iteration `it` over the whole schedule, strictly
 * sequentially: level by level, slot by slot, every write visible to the next unit at once - the
 * reference's own loop (Solver.cpp:774-911 / 929-1015).  Returns any-productive (:189,:210). */
static int run_iteration(packed_joint* P, const int* slots, const pxo_level* levels, int nlevels, solve_body* rows, int phase, int it)
{
    int any = 0;
    for (int l = 0; l < nlevels; ++l)
    {
        const pxo_level* L = &levels[l];
        int k = L->start;
        while (k < L->end)
        {
            int width = (k < L->grouped_end) ? 8 : 1;
            /* skip rule, Solver.cpp:787-798 / 946-957: lastIteration > it-2 on either body of ANY lane */
            int active = 0;
            for (int u = 0; u < width; ++u)
            {
                int s = slots[k + u];
                if (s < 0) continue;
                if (rows[P[s].b1].last > it - 2 || rows[P[s].b2].last > it - 2) active = 1;
            }
            if (active)
            {
                for (int u = 0; u < width; ++u)
                {
                    int s = slots[k + u];
                    if (s < 0) continue;
                    packed_joint* j = &P[s];
                    solve_body v1 = rows[j->b1], v2 = rows[j->b2];
                    int productive = phase == 0 ? impulse_joint(j, &v1, &v2, width == 8) : displacement_joint(j, &v1, &v2);
                    any |= productive;
                    if (productive) { v1.last = it; v2.last = it; }      /* Solver.cpp:903-904 */
                    rows[j->b1] = v1;
                    rows[j->b2] = v2;
                }
            }
            k += width;
        }
    }
    return any;
}

void pxo_solve_scheduled(pxo_body* bodies, int nb, pxo_joint* joints, int nj, const pxo_contact_point* cps,
    const int* slots, const pxo_level* levels, int nlevels, int contact_iters, int penetration_iters,
    int* iters_out)
{
    size_t nbs = (size_t)(nb > 0 ? nb : 1), njs = (size_t)(nj > 0 ? nj : 1);
    solve_params* par = (solve_params*)malloc(nbs * sizeof(solve_params));
    solve_body* imp = (solve_body*)malloc(nbs * sizeof(solve_body));
    solve_body* dis = (solve_body*)malloc(nbs * sizeof(solve_body));
    packed_joint* P = (packed_joint*)malloc(njs * sizeof(packed_joint));

    /* PrepareBodies, src/Solver.cpp:456-480 */
    for (int i = 0; i < nb; ++i)
    {
        par[i].invMass = bodies[i].invMass; par[i].invInertia = bodies[i].invInertia;
        par[i].px = bodies[i].pos.x; par[i].py = bodies[i].pos.y;
        imp[i].vx = bodies[i].velocity.x; imp[i].vy = bodies[i].velocity.y; imp[i].w = bodies[i].angularVelocity; imp[i].last = -1;
        dis[i].vx = bodies[i].displacingVelocity.x; dis[i].vy = bodies[i].displacingVelocity.y;
        dis[i].w = bodies[i].displacingAngularVelocity; dis[i].last = -1;
    }
    /* PrepareJoints copy, src/Solver.cpp:509-521 (packed storage is addressed by joint id here;
     * the processing ORDER is what the schedule fixes) */
    for (int i = 0; i < nj; ++i)
    {
        P[i].b1 = joints[i].body1Index; P[i].b2 = joints[i].body2Index; P[i].cp = joints[i].contactPointIndex;
        P[i].accN = joints[i].normalImpulse; P[i].accF = joints[i].frictionImpulse;
    }
    /* RefreshJoints (order-independent) then PreStepJoints (schedule order), Solver.cpp:147-165 */
    for (int i = 0; i < nj; ++i) refresh_joint(&P[i], imp, par, cps);
    for (int l = 0; l < nlevels; ++l)
        for (int k = levels[l].start; k < levels[l].end; ++k)
            if (slots[k] >= 0) prestep_joint(&P[slots[k]], imp);

    int ran0 = 0, ran1 = 0;
    for (int it = 0; it < contact_iters; ++it)                 /* Solver.cpp:175-190 */
    {
        ran0++;
        if (!run_iteration(P, slots, levels, nlevels, imp, 0, it)) break;
    }
    for (int it = 0; it < penetration_iters; ++it)             /* Solver.cpp:196-211 */
    {
        ran1++;
        if (!run_iteration(P, slots, levels, nlevels, dis, 1, it)) break;
    }
    /* FinishJoints / FinishBodies, src/Solver.cpp:482-494, 537-545 */
    for (int i = 0; i < nj; ++i)
    {
        joints[i].normalImpulse = P[i].accN;
        joints[i].frictionImpulse = P[i].accF;
    }
    for (int i = 0; i < nb; ++i)
    {
        bodies[i].velocity.x = imp[i].vx; bodies[i].velocity.y = imp[i].vy; bodies[i].angularVelocity = imp[i].w;
        bodies[i].displacingVelocity.x = dis[i].vx; bodies[i].displacingVelocity.y = dis[i].vy;
        bodies[i].displacingAngularVelocity = dis[i].w;
    }
    if (iters_out) { iters_out[0] = ran0; iters_out[1] = ran1; }
    free(P); free(dis); free(imp); free(par);
}

/* SolveJoints<N>, Island_Single: the reference order is "groups of N from PrepareIndices, then the
 * tail one by one"; expressed as a schedule in which every unit is its own level, the scheduled
 * solve above IS the sequential reference loop. */
void pxo_solve_joints(pxo_body* bodies, int nb, pxo_joint* joints, int nj, const pxo_contact_point* cps,
    int group, int contact_iters, int penetration_iters, int* joint_index_out, int* iters_out)
{
    int* order = (int*)malloc((size_t)(nj > 0 ? nj : 1) * sizeof(int));
    int group_offset = pxo_prepare_indices(joints, nj, nb, group, order);
    if (group == 1) group_offset = 0;

    /* slots: wide units padded to 8 lanes (N=4 uses lanes 0..3), singles after */
    int nunits = (group > 1 ? group_offset / group : 0) + (nj - group_offset);
    int nslots = (group > 1 ? (group_offset / group) * 8 : 0) + (nj - group_offset) * 8;
    int* slots = (int*)malloc((size_t)(nslots > 0 ? nslots : 1) * sizeof(int));
    pxo_level* levels = (pxo_level*)malloc((size_t)(nunits > 0 ? nunits : 1) * sizeof(pxo_level));
    int k = 0, l = 0;
    for (int g = 0; group > 1 && g < group_offset / group; ++g)
    {
        for (int u = 0; u < 8; ++u) slots[k + u] = (u < group) ? order[g * group + u] : -1;
        levels[l].start = k; levels[l].grouped_end = k + 8; levels[l].end = k + 8;
        k += 8; l++;
    }
    for (int i = group_offset; i < nj; ++i)
    {
        slots[k] = order[i];
        for (int u = 1; u < 8; ++u) slots[k + u] = -1;
        levels[l].start = k; levels[l].grouped_end = k; levels[l].end = k + 1;
        k += 8; l++;
    }
    pxo_solve_scheduled(bodies, nb, joints, nj, cps, slots, levels, nunits, contact_iters, penetration_iters, iters_out);
    if (joint_index_out) memcpy(joint_index_out, order, (size_t)nj * sizeof(int));
    free(levels); free(slots); free(order);
}

/* ------------------------------------------------------------------------------------------ */
/* One rank of the PARTITIONED solve (one world over several devices), executed
 * literally on the CPU so that the multi-rank protocol itself can be tested without a GPU: every
 * rank holds its own copy of all body rows and accumulators, relaxes the levels of its own class
 * ("interior"), receives the other ranks' boundary rows, and relaxes the levels of the cut class on
 * top of them.  The arithmetic is the functions above; what is restated here is the order of
 * operations of solve.cu's k_part_solve and its drivers.  level_class[l] = owning rank of level l,
 * `ranks` for a cut level. */
struct pxo_rank
{
    int nb, nj, nlevels, rank, ranks;
    solve_params* par;
    solve_body* rows[2];    /* impulse / displacement rows, this rank's copy */
    packed_joint* P;        /* this rank's copy of the packed joints (accumulators) */
    int* slots;
    pxo_level* levels;
    int* level_class;
    pxo_level* scratch;
};

pxo_rank* pxo_rank_create(const pxo_body* bodies, int nb, const pxo_joint* joints, int nj, const pxo_contact_point* cps,
    const int* slots, int nslots, const pxo_level* levels, const int* level_class, int nlevels, int rank, int ranks)
{
    pxo_rank* r = (pxo_rank*)calloc(1, sizeof(pxo_rank));
    size_t nbs = (size_t)(nb > 0 ? nb : 1), njs = (size_t)(nj > 0 ? nj : 1), nls = (size_t)(nlevels > 0 ? nlevels : 1);
    r->nb = nb; r->nj = nj; r->nlevels = nlevels; r->rank = rank; r->ranks = ranks;
    r->par = (solve_params*)malloc(nbs * sizeof(solve_params));
    r->rows[0] = (solve_body*)malloc(nbs * sizeof(solve_body));
    r->rows[1] = (solve_body*)malloc(nbs * sizeof(solve_body));
    r->P = (packed_joint*)malloc(njs * sizeof(packed_joint));
    r->slots = (int*)malloc((size_t)(nslots > 0 ? nslots : 1) * sizeof(int));
    r->levels = (pxo_level*)malloc(nls * sizeof(pxo_level));
    r->scratch = (pxo_level*)malloc(nls * sizeof(pxo_level));
    r->level_class = (int*)malloc(nls * sizeof(int));
    memcpy(r->slots, slots, (size_t)nslots * sizeof(int));
    memcpy(r->levels, levels, (size_t)nlevels * sizeof(pxo_level));
    memcpy(r->level_class, level_class, (size_t)nlevels * sizeof(int));
    for (int i = 0; i < nb; ++i)     /* PrepareBodies */
    {
        r->par[i].invMass = bodies[i].invMass; r->par[i].invInertia = bodies[i].invInertia;
        r->par[i].px = bodies[i].pos.x; r->par[i].py = bodies[i].pos.y;
        r->rows[0][i].vx = bodies[i].velocity.x; r->rows[0][i].vy = bodies[i].velocity.y; r->rows[0][i].w = bodies[i].angularVelocity;
        r->rows[0][i].last = -1;
        r->rows[1][i].vx = bodies[i].displacingVelocity.x; r->rows[1][i].vy = bodies[i].displacingVelocity.y;
        r->rows[1][i].w = bodies[i].displacingAngularVelocity; r->rows[1][i].last = -1;
    }
    for (int i = 0; i < nj; ++i)     /* PrepareJoints copy + RefreshJoints */
    {
        r->P[i].b1 = joints[i].body1Index; r->P[i].b2 = joints[i].body2Index; r->P[i].cp = joints[i].contactPointIndex;
        r->P[i].accN = joints[i].normalImpulse; r->P[i].accF = joints[i].frictionImpulse;
        refresh_joint(&r->P[i], r->rows[0], r->par, cps);
    }
    return r;
}

void pxo_rank_destroy(pxo_rank* r)
{
    if (!r) return;
    free(r->par); free(r->rows[0]); free(r->rows[1]); free(r->P); free(r->slots); free(r->levels); free(r->scratch); free(r->level_class);
    free(r);
}

/* One launch of k_part_solve: phase -1 = warm start, 0 = impulse iteration `it`, 1 = displacement iteration `it`;
 * cut = 0: the levels of this rank's own class, cut = 1: the levels of the cut class.  Returns any-productive. */
int pxo_rank_pass(pxo_rank* r, int phase, int it, int cut)
{
    const int want = cut ? r->ranks : r->rank;
    int n = 0;
    for (int l = 0; l < r->nlevels; ++l)
        if (r->level_class[l] == want) r->scratch[n++] = r->levels[l];
    if (phase < 0)
    {
        for (int l = 0; l < n; ++l)
            for (int k = r->scratch[l].start; k < r->scratch[l].end; ++k)
                if (r->slots[k] >= 0) prestep_joint(&r->P[r->slots[k]], r->rows[0]);
        return 0;
    }
    return run_iteration(r->P, r->slots, r->scratch, n, r->rows[phase], phase, it);
}

/* body rows {vx, vy, w, lastIteration} as 4 x 32 bit, the unit of the boundary exchange */
void pxo_rank_get_rows(const pxo_rank* r, int phase, const int* ids, int n, float* out4)
{
    const solve_body* rows = r->rows[phase == 1 ? 1 : 0];
    for (int i = 0; i < n; ++i) memcpy(out4 + 4 * (size_t)i, &rows[ids[i]], sizeof(solve_body));
}

void pxo_rank_set_rows(pxo_rank* r, int phase, const int* ids, int n, const float* in4)
{
    solve_body* rows = r->rows[phase == 1 ? 1 : 0];
    for (int i = 0; i < n; ++i) memcpy(&rows[ids[i]], in4 + 4 * (size_t)i, sizeof(solve_body));
}

/* accumulated impulses {normal, friction} of joints, the second part of the end-of-solve exchange */
void pxo_rank_get_acc(const pxo_rank* r, const int* ids, int n, float* out2)
{
    for (int i = 0; i < n; ++i) { out2[2 * (size_t)i] = r->P[ids[i]].accN; out2[2 * (size_t)i + 1] = r->P[ids[i]].accF; }
}

void pxo_rank_set_acc(pxo_rank* r, const int* ids, int n, const float* in2)
{
    for (int i = 0; i < n; ++i) { r->P[ids[i]].accN = in2[2 * (size_t)i]; r->P[ids[i]].accF = in2[2 * (size_t)i + 1]; }
}

/* FinishJoints / FinishBodies from this rank's copy (after the end-of-solve exchange it is the whole result) */
void pxo_rank_finish(const pxo_rank* r, pxo_body* bodies, pxo_joint* joints)
{
    for (int i = 0; i < r->nj; ++i) { joints[i].normalImpulse = r->P[i].accN; joints[i].frictionImpulse = r->P[i].accF; }
    for (int i = 0; i < r->nb; ++i)
    {
        bodies[i].velocity.x = r->rows[0][i].vx; bodies[i].velocity.y = r->rows[0][i].vy; bodies[i].angularVelocity = r->rows[0][i].w;
        bodies[i].displacingVelocity.x = r->rows[1][i].vx; bodies[i].displacingVelocity.y = r->rows[1][i].vy;
        bodies[i].displacingAngularVelocity = r->rows[1][i].w;
    }
}
