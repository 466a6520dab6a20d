// prl_device.cuh -- one environment per thread: the whole of `Environment::step(Move)`
// (reference src/common.rs:215-340) and of the rapier2d 0.17.2 `PhysicsPipeline::step` it calls
// (common.rs:315-329), for the scene class `from_world` can build:
// one rotation-locked capsule (the player), parent-less cuboids (fixed blocks), dynamic cuboids
// (pushable blocks) and goal rectangles.
//
// Design (B200-first, not a translation of rapier's object graph):
//  * body state lives in registers for the whole launch; a launch runs 1..number_of_steps steps;
//  * the world's static colliders / pairs / goals are read from a shared-memory DevWorld;
//  * contact manifolds live in a small per-thread slot table (<= MAXC live manifolds) kept
//    sorted by pair index through a 64-bit nibble list, so the Gauss-Seidel order is the
//    canonical ascending (collider1, collider2) order without moving manifold data;
//  * every stage is a flat loop over "my k-th item" so that the lanes of a warp run the narrow
//    phase / constraint solve of different environments in lockstep instead of serialising;
//  * arithmetic is f32 in the reference's expression order; the translation unit is compiled
//    with -fmad=false because rustc never contracts a*b+c, so results are reproducible against
//    a scalar CPU evaluation of the same expressions.
#pragma once
#include <math.h>
#include <stdint.h>

#include "prl_libm.h"
#include "prl_world.h"

// The product builds this header with nvcc for sm_100a only. tests/hostsim compiles the same
// templates with g++ (no CUDA) to single-step the device logic against the oracle on machines
// without a GPU; that build is test infrastructure and is never shipped.
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define PRL_D __host__ __device__ __forceinline__
#else
#define PRL_D inline
#endif

namespace prl {

PRL_D int first_zero_bit(uint32_t used) {  // index of the lowest clear bit, 32 if none
#if defined(__CUDA_ARCH__)
    int r = __ffs((int)~used) - 1;
    return r < 0 ? 32 : r;
#else
    for (int i = 0; i < 32; ++i) if (!((used >> i) & 1u)) return i;
    return 32;
#endif
}
PRL_D float f32_inf() {
#if defined(__CUDA_ARCH__)
    return __int_as_float(0x7f800000);
#else
    return INFINITY;
#endif
}

struct v2 { float x, y; };
struct Iso { v2 t; float c, s; };  // nalgebra Isometry2: translation + unit complex

PRL_D v2 V(float x, float y) { v2 r; r.x = x; r.y = y; return r; }
PRL_D v2 operator+(v2 a, v2 b) { return V(a.x + b.x, a.y + b.y); }
PRL_D v2 operator-(v2 a, v2 b) { return V(a.x - b.x, a.y - b.y); }
PRL_D v2 operator-(v2 a) { return V(-a.x, -a.y); }
PRL_D v2 operator*(v2 a, float k) { return V(a.x * k, a.y * k); }
PRL_D float dot(v2 a, v2 b) { return a.x * b.x + a.y * b.y; }
PRL_D float cross(v2 a, v2 b) { return a.x * b.y - a.y * b.x; }      // Vector2 x Vector2
PRL_D v2 wcross(float w, v2 v) { return V(-w * v.y, w * v.x); }       // scalar x Vector2
PRL_D float inv0(float x) { return x == 0.0f ? 0.0f : 1.0f / x; }
PRL_D v2 rot(float c, float s, v2 v) { return V(c * v.x - s * v.y, s * v.x + c * v.y); }
PRL_D v2 rot_inv(float c, float s, v2 v) { float ns = -s; return V(c * v.x - ns * v.y, ns * v.x + c * v.y); }
PRL_D v2 xform(const Iso& m, v2 p) { return rot(m.c, m.s, p) + m.t; }
PRL_D v2 xform_inv(const Iso& m, v2 p) { return rot_inv(m.c, m.s, p - m.t); }
PRL_D Iso inv_mul(const Iso& a, const Iso& b) {  // a^-1 * b
    Iso r;
    float ns = -a.s;
    r.t = rot_inv(a.c, a.s, b.t - a.t);
    r.c = a.c * b.c - ns * b.s;
    r.s = a.c * b.s + ns * b.c;
    return r;
}
PRL_D Iso inverse(const Iso& a) {
    Iso r;
    r.c = a.c;
    r.s = -a.s;
    r.t = rot(r.c, r.s, -a.t);
    return r;
}

// register arrays with a run-time index: selects over a compile-time unrolled range
template <int N> PRL_D float sel(const float (&a)[N], int i) {
    float r = a[0];
#pragma unroll
    for (int k = 1; k < N; ++k) r = (i == k) ? a[k] : r;
    return r;
}
template <int N> PRL_D void put(float (&a)[N], int i, float v) {
#pragma unroll
    for (int k = 0; k < N; ++k) if (i == k) a[k] = v;
}

constexpr float kCos1Deg = 0.99984769515f;   // parry utils::COS_1_DEGREES
constexpr float kPersistDistSq = 1.0e-6f;    // parry try_update_contacts_eps distance threshold^2
constexpr float kFriction = 0.5f;            // rapier collider default, Average combine rule
constexpr float kFltMax = 3.40282347e+38f;
constexpr float kTimeUntilSleep = 2.0f;      // rapier RigidBodyActivation defaults
constexpr float kLinThreshold = 0.4f, kAngThreshold = 0.5f;

// A contact manifold of one collider pair (parry ContactManifold restricted to <= 2 points):
// normals local to collider 1 / 2, points local to collider 1 / 2, signed distance.
struct Manifold {
    int npts;
    v2 n1, n2;
    v2 p1[2], p2[2];
    float dist[2];
};

// parry ContactManifold::try_update_contacts_eps(pos12, cos 1 deg, 1e-6): keep the cached
// manifold if the normals still agree and no point slid by more than 1 mm or changed side.
PRL_D bool manifold_try_update(Manifold& m, const Iso& pos12) {
    if (m.npts == 0) return false;
    v2 ln2 = rot(pos12.c, pos12.s, m.n2);
    if (-dot(m.n1, ln2) < kCos1Deg) return false;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (k < m.npts) {
            v2 lp2 = xform(pos12, m.p2[k]);
            v2 dpt = lp2 - m.p1[k];
            float dist = dot(dpt, m.n1);
            if (dist * m.dist[k] < 0.0f) return false;
            v2 np1 = lp2 - m.n1 * dist;
            v2 d = m.p1[k] - np1;
            if (dot(d, d) > kPersistDistSq) return false;
            m.dist[k] = dist;
            m.p1[k] = np1;
        }
    }
    return true;
}

// parry Cuboid::support_face (2-D): the face most aligned with dir
PRL_D void cuboid_support_face(v2 he, v2 dir, v2& a, v2& b) {
    if (fabsf(dir.y) < fabsf(dir.x)) {
        float xj = copysignf(he.x, dir.x);
        a = V(xj, he.y);
        b = V(xj, -he.y);
    } else {
        float yj = copysignf(he.y, dir.y);
        a = V(he.x, yj);
        b = V(-he.x, yj);
    }
}

// parry clip_segment_segment_with_normal + PolygonalFeature::face_face_contacts (2-D): clip two
// faces against each other along the tangent; two contact points when the ranges overlap.
PRL_D void face_face_contacts(Manifold& m, const Iso& pos12, v2 s1a, v2 s1b, v2 normal1, v2 s2a, v2 s2b) {
    v2 tangent = V(-normal1.y, normal1.x);
    float r1a = dot(s1a, tangent), r1b = dot(s1b, tangent);
    float r2a = dot(s2a, tangent), r2b = dot(s2b, tangent);
    if (r1b < r1a) { float t = r1a; r1a = r1b; r1b = t; v2 p = s1a; s1a = s1b; s1b = p; }
    if (r2b < r2a) { float t = r2a; r2a = r2b; r2b = t; v2 p = s2a; s2a = s2b; s2b = p; }
    if (r2a > r1b || r1a > r2b) return;
    v2 c0, c1, c2, c3;
    if (r2a > r1a) {
        float bc = (r2a - r1a) * inv0(r1b - r1a);
        c0 = s1a + (s1b - s1a) * bc;
        c1 = s2a;
    } else {
        float bc = (r1a - r2a) * inv0(r2b - r2a);
        c0 = s1a;
        c1 = s2a + (s2b - s2a) * bc;
    }
    if (r2b < r1b) {
        float bc = (r2b - r1a) * inv0(r1b - r1a);
        c2 = s1a + (s1b - s1a) * bc;
        c3 = s2b;
    } else {
        float bc = (r1b - r2a) * inv0(r2b - r2a);
        c2 = s1b;
        c3 = s2a + (s2b - s2a) * bc;
    }
    m.p1[0] = c0;
    m.p2[0] = xform_inv(pos12, c1);
    m.dist[0] = dot(c1 - c0, normal1);
    m.p1[1] = c2;
    m.p2[1] = xform_inv(pos12, c3);
    m.dist[1] = dot(c3 - c2, normal1);
    m.npts = 2;
}

// parry sat::cuboid_cuboid_find_local_separating_normal_oneway
PRL_D float cuboid_sat_oneway(v2 he1, v2 he2, const Iso& pos12, v2& best_dir) {
    float best = -kFltMax;
    best_dir = V(0.0f, 0.0f);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        float ti = i == 0 ? pos12.t.x : pos12.t.y;
        float sign = copysignf(1.0f, ti);
        v2 axis1 = i == 0 ? V(sign, 0.0f) : V(0.0f, sign);
        v2 axis2 = rot_inv(pos12.c, pos12.s, -axis1);
        v2 lp2 = V(copysignf(he2.x, axis2.x), copysignf(he2.y, axis2.y));
        v2 pt2 = xform(pos12, lp2);
        float sep = (i == 0 ? pt2.x : pt2.y) * sign - (i == 0 ? he1.x : he1.y);
        if (sep > best) { best = sep; best_dir = axis1; }
    }
    return best;
}

// parry contact_manifold_cuboid_cuboid
PRL_D void contact_cuboid_cuboid(Manifold& m, const Iso& pos12, v2 he1, v2 he2, float prediction) {
    if (manifold_try_update(m, pos12)) return;
    Iso pos21 = inverse(pos12);
    v2 d1, d2;
    float sep1 = cuboid_sat_oneway(he1, he2, pos12, d1);
    if (sep1 > prediction) { m.npts = 0; return; }
    float sep2 = cuboid_sat_oneway(he2, he1, pos21, d2);
    if (sep2 > prediction) { m.npts = 0; return; }
    v2 best = d1;
    if (sep2 > sep1) best = rot(pos12.c, pos12.s, -d2);
    m.npts = 0;
    v2 ln2 = rot(pos21.c, pos21.s, -best);
    v2 f1a, f1b, f2a, f2b;
    cuboid_support_face(he1, best, f1a, f1b);
    cuboid_support_face(he2, ln2, f2a, f2b);
    f2a = xform(pos12, f2a);
    f2b = xform(pos12, f2b);
    face_face_contacts(m, pos12, f1a, f1b, best, f2a, f2b);
    m.n1 = best;
    m.n2 = ln2;
}

// Closest points between the capsule's inner segment [a,b] and the box [-he,he], in the box
// frame: exact closest-feature enumeration (2 segment endpoints vs the box, 4 box corners vs the
// segment; fixed order, strict '<'). parry reaches the same pair with GJK up to its tolerance.
// Returns false if the segment touches the box.
PRL_D bool segment_box_closest(v2 a, v2 b, v2 he, float& u_out, v2& q_out) {
    v2 d = b - a;
    {
        float t0 = 0.0f, t1 = 1.0f;
        bool hit = true;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            float av = i == 0 ? a.x : a.y, dv = i == 0 ? d.x : d.y, hv = i == 0 ? he.x : he.y;
            if (hit) {
                if (dv == 0.0f) {
                    if (av < -hv || av > hv) hit = false;
                } else {
                    float inv = 1.0f / dv;
                    float ta = (-hv - av) * inv, tb = (hv - av) * inv;
                    if (ta > tb) { float t = ta; ta = tb; tb = t; }
                    if (ta > t0) t0 = ta;
                    if (tb < t1) t1 = tb;
                    if (t0 > t1) hit = false;
                }
            }
        }
        if (hit) return false;
    }
    float best = kFltMax, bu = 0.0f;
    v2 bq = V(0.0f, 0.0f);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        v2 p = e == 0 ? a : b;
        v2 q = V(fminf(fmaxf(p.x, -he.x), he.x), fminf(fmaxf(p.y, -he.y), he.y));
        v2 r = q - p;
        float dd = dot(r, r);
        if (dd < best) { best = dd; bu = e == 0 ? 0.0f : 1.0f; bq = q; }
    }
    float dd2 = dot(d, d);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v2 cn = V((k == 1 || k == 2) ? he.x : -he.x, (k >= 2) ? he.y : -he.y);
        float u = dot(cn - a, d) * inv0(dd2);
        u = fminf(fmaxf(u, 0.0f), 1.0f);
        v2 ps = a + d * u;
        v2 r = cn - ps;
        float dd = dot(r, r);
        if (dd < best) { best = dd; bu = u; bq = cn; }
    }
    u_out = bu;
    q_out = bq;
    return true;
}

// parry contact_manifold_pfm_pfm with pfm1 = the capsule's segment (border radius r) and
// pfm2 = the cuboid: closest points, then segment-vs-support-face clipping, then the radius shift.
PRL_D void contact_capsule_cuboid(Manifold& m, const Iso& pos12, float radius, float hh, v2 he2, float prediction) {
    if (manifold_try_update(m, pos12)) return;
    const float total_prediction = prediction + radius + 0.0f;
    v2 sa = V(0.0f, -hh), sb = V(0.0f, hh);
    v2 a2 = xform_inv(pos12, sa), b2 = xform_inv(pos12, sb);
    float u;
    v2 q, p1, p2_1, dir;
    m.npts = 0;
    if (segment_box_closest(a2, b2, he2, u, q)) {
        p1 = sa + (sb - sa) * u;
        p2_1 = xform(pos12, q);
        v2 dp = p2_1 - p1;
        float n = sqrtf(dot(dp, dp));
        if (n > total_prediction) return;
        dir = V(dp.x / n, dp.y / n);
    } else {
        // inner segment inside the box (parry: EPA): minimum face penetration of the deeper endpoint
        float best = kFltMax;
        v2 nb = V(0.0f, 1.0f), pe = a2, qf = a2;
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            v2 nf = f == 0 ? V(1.0f, 0.0f) : f == 1 ? V(-1.0f, 0.0f) : f == 2 ? V(0.0f, 1.0f) : V(0.0f, -1.0f);
            float hf = f < 2 ? he2.x : he2.y;
            float da = dot(a2, nf), db = dot(b2, nf);
            v2 e = da < db ? a2 : b2;
            float depth = hf - fminf(da, db);
            if (depth < best) { best = depth; nb = nf; pe = e; qf = e + nf * depth; }
        }
        p1 = xform(pos12, pe);
        p2_1 = xform(pos12, qf);
        dir = -rot(pos12.c, pos12.s, nb);
    }
    v2 n2 = rot_inv(pos12.c, pos12.s, -dir);
    v2 f2a, f2b;
    cuboid_support_face(he2, n2, f2a, f2b);
    f2a = xform(pos12, f2a);
    f2b = xform(pos12, f2b);
    face_face_contacts(m, pos12, sa, sb, dir, f2a, f2b);
    if (m.npts == 0) {
        m.p1[0] = p1;
        m.p2[0] = xform_inv(pos12, p2_1);
        m.dist[0] = dot(p2_1 - p1, dir);
        m.npts = 1;
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (k < m.npts) {
            m.p1[k] = m.p1[k] + dir * radius;
            m.dist[k] -= radius + 0.0f;
        }
    }
    m.n1 = dir;
    m.n2 = n2;
}

// rapier RigidBodyVelocity::integrate with local_com = origin:
// shift * Isometry::new(linvel*dt, angvel*dt) * shift^-1 * init, then renormalize_fast
PRL_D Iso integrate_pose(v2 linvel, float angvel, const Iso& init, float dt) {
    v2 com = init.t;
    float ang = angvel * dt;
    float rc = libm_cosf(ang), rs = libm_sinf(ang);
    v2 tr = linvel * dt;
    v2 t = com + tr;
    t = t + rot(rc, rs, -com);
    t = t + rot(rc, rs, init.t);
    Iso r;
    r.t = t;
    r.c = rc * init.c - rs * init.s;
    r.s = rc * init.s + rs * init.c;
    float sq = r.c * r.c + r.s * r.s;
    float k = 0.5f * (3.0f - sq);
    r.c *= k;
    r.s *= k;
    return r;
}

// ------------------------------------------------------------------------------------------
// Per-thread environment. NB = bodies (player + pushable blocks) as a compile-time bound so the
// body arrays stay in registers; S = manifold slots.
template <int NB, int S = MAXC>
struct Env {
    // rigid bodies (0 = player)
    float x[NB], y[NB], c[NB], s[NB], vx[NB], vy[NB], w[NB], timer[NB];
    uint32_t sleeping;   // bit b
    uint32_t flags;      // FLAG_WON | FLAG_FRESH | FLAG_OVERFLOW
    // manifold slot table
    uint64_t order;      // nibble k = slot of the k-th live manifold in ascending pair order
    int count;
    uint32_t used;       // bit s = slot s is live
    int spair[S];        // pair index of slot s
    int snpts[S];
    float mf[S][14];     // n1, n2, (p1, p2, dist) x 2
    // per-step scratch, indexed by slot: solver contacts and velocity constraints
    float cnx[S], cny[S];            // world normal of the manifold
    int cn[S];                       // number of solver contacts (0..2)
    float cpx[S][2], cpy[S][2], cpd[S][2];
    int cb[S];                       // body1 | body2 << 8 (body1 = 255: ground constraint)
    float cdx[S], cdy[S], cim1[S], cim2[S], ccfm[S];
    float ce[S][2][11];              // per point: gcross1, gcross2, rhs, rhs_wo_bias, impulse, r,
                                     //            t_gcross1, t_gcross2, t_rhs, t_impulse, t_r
};

template <int NB, int S>
struct Stepper {
    using E = Env<NB, S>;
    const DevWorld& W;
    E& e;
    PRL_D Stepper(const DevWorld& w, E& env) : W(w), e(env) {}

    PRL_D int slot_at(int k) const { return (int)((e.order >> (4 * k)) & 15u); }
    PRL_D Iso body_pose(int b) const { Iso p; p.t = V(sel(e.x, b), sel(e.y, b)); p.c = sel(e.c, b); p.s = sel(e.s, b); return p; }
    PRL_D Iso collider_pose(int c) const {
        int b = W.body_of[c];
        if (b >= 0) return body_pose(b);
        Iso p; p.t = V(W.cx[c], W.cy[c]); p.c = W.cc[c]; p.s = W.cs[c];
        return p;
    }
    PRL_D void wake_strong(int b) { e.sleeping &= ~(1u << b); put(e.timer, b, 0.0f); }

    PRL_D void load_manifold(int slot, Manifold& m) const {
        m.npts = e.snpts[slot];
        const float* f = e.mf[slot];
        m.n1 = V(f[0], f[1]); m.n2 = V(f[2], f[3]);
        m.p1[0] = V(f[4], f[5]); m.p2[0] = V(f[6], f[7]); m.dist[0] = f[8];
        m.p1[1] = V(f[9], f[10]); m.p2[1] = V(f[11], f[12]); m.dist[1] = f[13];
    }
    PRL_D void store_manifold(int slot, const Manifold& m) {
        e.snpts[slot] = m.npts;
        float* f = e.mf[slot];
        f[0] = m.n1.x; f[1] = m.n1.y; f[2] = m.n2.x; f[3] = m.n2.y;
        f[4] = m.p1[0].x; f[5] = m.p1[0].y; f[6] = m.p2[0].x; f[7] = m.p2[0].y; f[8] = m.dist[0];
        f[9] = m.p1[1].x; f[10] = m.p1[1].y; f[11] = m.p2[1].x; f[12] = m.p2[1].y; f[13] = m.dist[1];
    }
    // position in the sorted list of pair p, or -1
    PRL_D int find(int p) const {
        int r = -1;
        for (int k = 0; k < e.count; ++k) if (e.spair[slot_at(k)] == p) r = k;
        return r;
    }
    PRL_D void remove_at(int k) {
        int slot = slot_at(k);
        e.used &= ~(1u << slot);
        uint64_t low = k == 0 ? 0ull : (e.order & (~0ull >> (64 - 4 * k)));
        uint64_t high = k >= 15 ? 0ull : (e.order >> (4 * (k + 1))) << (4 * k);
        e.order = low | high;
        --e.count;
    }
    // allocate a slot for pair p and link it at its sorted position; -1 on overflow
    PRL_D int insert(int p) {
        int slot = first_zero_bit(e.used);
        if (slot >= S) { e.flags |= FLAG_OVERFLOW; return -1; }
        int k = 0;
        while (k < e.count && e.spair[slot_at(k)] < p) ++k;
        uint64_t low = k == 0 ? 0ull : (e.order & (~0ull >> (64 - 4 * k)));
        uint64_t high = k == 0 ? e.order : (e.order >> (4 * k));
        e.order = low | ((uint64_t)slot << (4 * k)) | (k >= 15 ? 0ull : (high << (4 * (k + 1))));
        e.used |= 1u << slot;
        e.spair[slot] = p;
        ++e.count;
        return slot;
    }

    // rapier NarrowPhase::compute_contacts tail: solver contacts of one manifold. They are a pure
    // function of the cached manifold and collider 1's pose, which cannot change between the end
    // of one pipeline step and the constraint build of the next, so they are recomputed here
    // rather than stored.
    PRL_D void solver_contacts(int slot) {
        int p = e.spair[slot];
        Iso pose1 = collider_pose(W.pair_c1[p]);
        Manifold m;
        load_manifold(slot, m);
        v2 normal = rot(pose1.c, pose1.s, m.n1);
        int ns = 0;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (k < m.npts && m.dist[k] < W.prediction) {
                v2 wp = xform(pose1, m.p1[k]);
                v2 half = normal * m.dist[k];
                e.cpx[slot][ns] = wp.x + half.x / 2.0f;
                e.cpy[slot][ns] = wp.y + half.y / 2.0f;
                e.cpd[slot][ns] = m.dist[k];
                ++ns;
            }
        }
        e.cnx[slot] = normal.x;
        e.cny[slot] = normal.y;
        e.cn[slot] = ns;
    }

    // rapier RigidBody::apply_impulse; a zero impulse neither moves nor wakes the body
    PRL_D void apply_impulse(int b, v2 j) {
        if (j.x != 0.0f || j.y != 0.0f) {
            float im = W.inv_mass[b];
            put(e.vx, b, sel(e.vx, b) + j.x * im);
            put(e.vy, b, sel(e.vy, b) + j.y * im);
            wake_strong(b);
        }
    }
    // rapier RigidBody::apply_impulse_at_point
    PRL_D void apply_impulse_at_point(int b, v2 j, v2 pt) {
        float torque = cross(pt - V(sel(e.x, b), sel(e.y, b)), j);
        apply_impulse(b, j);
        if (torque != 0.0f) {
            float iis = W.inv_inertia_sqrt[b];
            put(e.w, b, sel(e.w, b) + iis * (iis * torque));
            wake_strong(b);
        }
    }

    // common.rs:216-313: ground-contact scan over the player's contact pairs, then the move impulses
    PRL_D void player_move(uint32_t mv) {
        v2 lower = V(e.x[0], e.y[0] - W.lower_off);
        int nfc = 0;
        v2 pmin = V(0.0f, 0.0f), pmax = V(0.0f, 0.0f);
        int bmin = -1, bmax = -1;
        int nplayer = 0;  // live manifolds of player pairs: the head of the sorted list
        for (int k = 0; k < e.count; ++k) {
            int slot = slot_at(k);
            int p = e.spair[slot];
            if (p >= W.n_player_pairs) break;
            ++nplayer;
            int body = W.body_of[W.pair_c2[p]];
            for (int i = 0; i < e.cn[slot]; ++i) {
                v2 pt = V(e.cpx[slot][i], e.cpy[slot][i]);
                v2 q = pt - lower;
                q = V(q.x / W.radius, q.y / W.radius);
                if (q.y < -0.707f) {
                    if (nfc == 0 || pt.x < pmin.x) { pmin = pt; bmin = body; }    // min_by: first minimum
                    if (nfc == 0 || !(pt.x < pmax.x)) { pmax = pt; bmax = body; } // max_by: last maximum
                    ++nfc;
                }
            }
        }
        if (nfc == 0) return;
        v2 player_impulse = V(0.0f, 0.0f);
        if (mv & 1u) {
            v2 n = pmin - lower;
            float len = sqrtf(n.x * n.x + n.y * n.y);
            n = V(n.x / len, n.y / len);
            v2 imp = V(0.003f * n.y, -0.003f * n.x);
            if (NB > 1 && bmin >= 0) apply_impulse_at_point(bmin, -imp, pmin);
            player_impulse = player_impulse + imp;
        }
        if (mv & 2u) {
            v2 n = pmax - lower;
            float len = sqrtf(n.x * n.x + n.y * n.y);
            n = V(n.x / len, n.y / len);
            v2 imp = V(-0.003f * n.y, 0.003f * n.x);
            if (NB > 1 && bmax >= 0) apply_impulse_at_point(bmax, -imp, pmax);
            player_impulse = player_impulse + imp;
        }
        if (mv & 4u) {
            float cnt = (float)nfc;
            for (int k = 0; k < nplayer; ++k) {
                int slot = slot_at(k);
                int body = W.body_of[W.pair_c2[e.spair[slot]]];
                for (int i = 0; i < e.cn[slot]; ++i) {
                    v2 pt = V(e.cpx[slot][i], e.cpy[slot][i]);
                    v2 q = pt - lower;
                    q = V(q.x / W.radius, q.y / W.radius);
                    if (q.y < -0.707f) {
                        v2 n = pt - lower;
                        float len = sqrtf(n.x * n.x + n.y * n.y);
                        n = V(n.x / len, n.y / len);
                        v2 imp = V(-0.1f * n.x / cnt, -0.1f * n.y / cnt);
                        if (NB > 1 && body >= 0) apply_impulse_at_point(body, -imp, pt);
                        player_impulse = player_impulse + imp;
                    }
                }
            }
        }
        apply_impulse(0, player_impulse);
    }

    // rapier IslandManager::update_active_set_with_contacts: sleep timers, then everything
    // reachable from an awake body through a manifold with solver contacts stays awake.
    PRL_D void update_active_set() {
        uint32_t tentative = 0, reached = 0;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            if (b < W.n_bodies && !((e.sleeping >> b) & 1u)) {
                float sq_lin = e.vx[b] * e.vx[b] + e.vy[b] * e.vy[b];
                float sq_ang = e.w[b] * e.w[b];
                if (sq_lin < kLinThreshold * fabsf(kLinThreshold) && sq_ang < kAngThreshold * fabsf(kAngThreshold))
                    e.timer[b] += W.dt;
                else
                    e.timer[b] = 0.0f;
                if (e.timer[b] >= kTimeUntilSleep) tentative |= 1u << b;
                else reached |= 1u << b;
            }
        }
        if (NB > 1) {
            for (int round = 0; round < NB - 1; ++round) {
                for (int k = 0; k < e.count; ++k) {
                    int slot = slot_at(k);
                    if (e.cn[slot] == 0) continue;
                    int p = e.spair[slot];
                    int b1 = W.body_of[W.pair_c1[p]], b2 = W.body_of[W.pair_c2[p]];
                    if (b1 < 0 || b2 < 0) continue;
                    uint32_t m1 = 1u << b1, m2 = 1u << b2;
                    if ((reached & m1) || (reached & m2)) reached |= m1 | m2;
                }
            }
        }
        e.sleeping = (e.sleeping | tentative) & ~reached;
        uint32_t fell_asleep = tentative & e.sleeping;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            if ((fell_asleep >> b) & 1u) {  // RigidBody::sleep
                e.timer[b] = kTimeUntilSleep;
                e.vx[b] = 0.0f; e.vy[b] = 0.0f; e.w[b] = 0.0f;
            }
        }
    }

    // rapier VelocityConstraint::generate / VelocityGroundConstraint::generate for one manifold
    PRL_D void generate_constraint(int slot) {
        int p = e.spair[slot];
        int h1 = W.body_of[W.pair_c1[p]], h2 = W.body_of[W.pair_c2[p]];
        bool two_body = h1 >= 0 && h2 >= 0;
        v2 normal = V(e.cnx[slot], e.cny[slot]);
        v2 dir1;
        if (two_body) dir1 = -normal;
        else if (h2 < 0) { h2 = h1; h1 = -1; dir1 = normal; }  // the dynamic body becomes body 2
        else dir1 = -normal;
        v2 v1 = V(0.0f, 0.0f), com1 = V(0.0f, 0.0f);
        float w1 = 0.0f, im1 = 0.0f, ii1 = 0.0f;
        v2 v2_ = V(sel(e.vx, h2), sel(e.vy, h2)), com2 = V(sel(e.x, h2), sel(e.y, h2));
        float w2 = sel(e.w, h2), im2 = W.inv_mass[h2], ii2 = W.inv_inertia_sqrt[h2];
        float ccd = W.ccd_thickness[h2];
        if (two_body) {
            v1 = V(sel(e.vx, h1), sel(e.vy, h1));
            com1 = V(sel(e.x, h1), sel(e.y, h1));
            w1 = sel(e.w, h1);
            im1 = W.inv_mass[h1];
            ii1 = W.inv_inertia_sqrt[h1];
            ccd = W.ccd_thickness[h1] + W.ccd_thickness[h2];
        }
        v2 tan1 = V(-dir1.y, dir1.x);
        e.cb[slot] = (two_body ? h1 : 255) | (h2 << 8);
        e.cdx[slot] = dir1.x;
        e.cdy[slot] = dir1.y;
        e.cim1[slot] = im1;
        e.cim2[slot] = im2;
        bool fast = false;
        for (int k = 0; k < e.cn[slot]; ++k) {
            v2 pt = V(e.cpx[slot][k], e.cpy[slot][k]);
            float dist = e.cpd[slot][k];
            v2 dp1 = pt - com1, dp2 = pt - com2;
            v2 vel1 = v1 + wcross(w1, dp1);
            v2 vel2 = v2_ + wcross(w2, dp2);
            float* el = e.ce[slot][k];
            {
                float gcross2 = ii2 * cross(dp2, -dir1);
                float gcross1 = 0.0f, pm;
                if (two_body) {
                    gcross1 = ii1 * cross(dp1, dir1);
                    float imsum = im1 + im2;
                    pm = inv0(dot(dir1, dir1 * imsum) + gcross1 * gcross1 + gcross2 * gcross2);
                } else {
                    pm = inv0(dot(dir1, dir1 * im2) + gcross2 * gcross2);
                }
                float rhs_wo_bias = (1.0f + 0.0f * 0.0f) * dot(vel1 - vel2, dir1);
                rhs_wo_bias += fmaxf(dist, 0.0f) * W.inv_dt;
                rhs_wo_bias *= 0.0f + 1.0f;
                float pen = dist + W.allowed_linear_error;
                pen = fminf(fmaxf(pen, -kFltMax), 0.0f);
                float rhs_bias = W.erp_inv_dt * pen;
                float rhs = rhs_wo_bias + rhs_bias;
                fast = fast || (-rhs * W.dt > ccd * 0.5f);
                el[0] = gcross1; el[1] = gcross2; el[2] = rhs; el[3] = rhs_wo_bias; el[4] = 0.0f; el[5] = pm;
            }
            {
                float gcross2 = ii2 * cross(dp2, -tan1);
                float gcross1 = 0.0f, r;
                if (two_body) {
                    gcross1 = ii1 * cross(dp1, tan1);
                    float imsum = im1 + im2;
                    r = dot(tan1, tan1 * imsum) + gcross1 * gcross1 + gcross2 * gcross2;
                } else {
                    r = dot(tan1, tan1 * im2) + gcross2 * gcross2;
                }
                v2 dv = (vel1 - vel2) + V(0.0f, 0.0f);
                el[6] = gcross1; el[7] = gcross2; el[8] = dot(dv, tan1); el[9] = 0.0f; el[10] = inv0(r);
            }
        }
        e.ccfm[slot] = fast ? 1.0f : W.cfm_factor;
    }

    // rapier VelocityConstraintElement::solve_group / VelocityGroundConstraintElement::solve_group
    PRL_D void solve_constraint(int slot, bool normal_part, bool friction_part, float (&dlx)[NB], float (&dly)[NB], float (&da)[NB]) {
        int b1 = e.cb[slot] & 255, b2 = e.cb[slot] >> 8;
        v2 dir1 = V(e.cdx[slot], e.cdy[slot]);
        v2 tan1 = V(-dir1.y, dir1.x);
        float im2 = e.cim2[slot];
        v2 l2 = V(sel(dlx, b2), sel(dly, b2));
        float a2 = sel(da, b2);
        int n = e.cn[slot];
        if (NB > 1 && b1 != 255) {
            float im1 = e.cim1[slot];
            v2 l1 = V(sel(dlx, b1), sel(dly, b1));
            float a1 = sel(da, b1);
            if (normal_part) {
                float cfm = e.ccfm[slot];
                for (int k = 0; k < n; ++k) {
                    float* el = e.ce[slot][k];
                    float dvel = dot(dir1, l1) + el[0] * a1 - dot(dir1, l2) + el[1] * a2 + el[2];
                    float ni = cfm * fmaxf(el[4] - el[5] * dvel, 0.0f);
                    float dl = ni - el[4];
                    el[4] = ni;
                    l1 = l1 + (dir1 * im1) * dl;
                    a1 += el[0] * dl;
                    l2 = l2 + (dir1 * im2) * -dl;
                    a2 += el[1] * dl;
                }
            }
            if (friction_part) {
                for (int k = 0; k < n; ++k) {
                    float* el = e.ce[slot][k];
                    float limit = kFriction * el[4];
                    float dvel = dot(tan1, l1) + el[6] * a1 - dot(tan1, l2) + el[7] * a2 + el[8];
                    float ni = fminf(fmaxf(el[9] - el[10] * dvel, -limit), limit);
                    float dl = ni - el[9];
                    el[9] = ni;
                    l1 = l1 + (tan1 * im1) * dl;
                    a1 += el[6] * dl;
                    l2 = l2 + (tan1 * im2) * -dl;
                    a2 += el[7] * dl;
                }
            }
            put(dlx, b1, l1.x); put(dly, b1, l1.y); put(da, b1, a1);
        } else {
            if (normal_part) {
                float cfm = e.ccfm[slot];
                for (int k = 0; k < n; ++k) {
                    float* el = e.ce[slot][k];
                    float dvel = -dot(dir1, l2) + el[1] * a2 + el[2];
                    float ni = cfm * fmaxf(el[4] - el[5] * dvel, 0.0f);
                    float dl = ni - el[4];
                    el[4] = ni;
                    l2 = l2 + (dir1 * im2) * -dl;
                    a2 += el[1] * dl;
                }
            }
            if (friction_part) {
                for (int k = 0; k < n; ++k) {
                    float* el = e.ce[slot][k];
                    float limit = kFriction * el[4];
                    float dvel = -dot(tan1, l2) + el[7] * a2 + el[8];
                    float ni = fminf(fmaxf(el[9] - el[10] * dvel, -limit), limit);
                    float dl = ni - el[9];
                    el[9] = ni;
                    l2 = l2 + (tan1 * im2) * -dl;
                    a2 += el[7] * dl;
                }
            }
        }
        put(dlx, b2, l2.x); put(dly, b2, l2.y); put(da, b2, a2);
    }

    // rapier islands + constraint build + VelocitySolver::solve + position update for the
    // active bodies (SURVEY A.3 stages 2-7). Returns the mask of bodies that moved.
    PRL_D uint32_t solve_and_integrate() {
        update_active_set();
        uint32_t active = ~e.sleeping & ((1u << W.n_bodies) - 1u);
        // constraints: two-body manifolds first, then one-body ("ground") ones, each in pair order
        uint64_t cons = 0;
        int ncons = 0;
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            if (NB == 1 && pass == 0) continue;
            for (int k = 0; k < e.count; ++k) {
                int slot = slot_at(k);
                if (e.cn[slot] == 0) continue;
                int p = e.spair[slot];
                int h1 = W.body_of[W.pair_c1[p]], h2 = W.body_of[W.pair_c2[p]];
                bool two_body = h1 >= 0 && h2 >= 0;
                if ((pass == 0) != two_body) continue;
                bool awake1 = h1 >= 0 && ((active >> h1) & 1u);
                bool awake2 = h2 >= 0 && ((active >> h2) & 1u);
                if (!awake1 && !awake2) continue;
                cons |= (uint64_t)slot << (4 * ncons);
                ++ncons;
            }
        }
        for (int i = 0; i < ncons; ++i) generate_constraint((int)((cons >> (4 * i)) & 15u));

        float dlx[NB], dly[NB], da[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            dlx[b] = b < W.n_bodies ? W.gdv_x[b] : 0.0f;
            dly[b] = b < W.n_bodies ? W.gdv_y[b] : 0.0f;
            da[b] = 0.0f;
        }
        // 4 velocity iterations with interleaved friction, then the 4 remaining friction iterations
        for (int it = 0; it < 4; ++it) {
            for (int i = 0; i < ncons; ++i) solve_constraint((int)((cons >> (4 * i)) & 15u), true, false, dlx, dly, da);
            for (int i = 0; i < ncons; ++i) solve_constraint((int)((cons >> (4 * i)) & 15u), false, true, dlx, dly, da);
        }
        for (int it = 0; it < 4; ++it)
            for (int i = 0; i < ncons; ++i) solve_constraint((int)((cons >> (4 * i)) & 15u), false, true, dlx, dly, da);
        // positions integrate the biased velocities
        float nx[NB], ny[NB], nc[NB], ns[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            nx[b] = e.x[b]; ny[b] = e.y[b]; nc[b] = e.c[b]; ns[b] = e.s[b];
            if ((active >> b) & 1u) {
                float dangvel = W.inv_inertia_sqrt[b] * da[b];
                Iso init; init.t = V(e.x[b], e.y[b]); init.c = e.c[b]; init.s = e.s[b];
                Iso np = integrate_pose(V(e.vx[b] + dlx[b], e.vy[b] + dly[b]), e.w[b] + dangvel, init, W.dt);
                nx[b] = np.t.x; ny[b] = np.t.y; nc[b] = np.c; ns[b] = np.s;
            }
        }
        // one stabilisation sweep without the penetration bias, then the final velocities
        for (int i = 0; i < ncons; ++i) {
            int slot = (int)((cons >> (4 * i)) & 15u);
            e.ccfm[slot] = 1.0f;
            for (int k = 0; k < e.cn[slot]; ++k) e.ce[slot][k][2] = e.ce[slot][k][3];
        }
        for (int i = 0; i < ncons; ++i) solve_constraint((int)((cons >> (4 * i)) & 15u), true, false, dlx, dly, da);
        for (int i = 0; i < ncons; ++i) solve_constraint((int)((cons >> (4 * i)) & 15u), false, true, dlx, dly, da);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            if ((active >> b) & 1u) {
                float dangvel = W.inv_inertia_sqrt[b] * da[b];
                e.vx[b] = e.vx[b] + dlx[b];
                e.vy[b] = e.vy[b] + dly[b];
                e.w[b] += dangvel;
                e.x[b] = nx[b]; e.y[b] = ny[b]; e.c[b] = nc[b]; e.s[b] = ns[b];
            }
        }
        return active;
    }

    // rapier PhysicsPipeline::detect_collisions for the colliders of the bodies in `moved`
    // (`all` on the first step): loosened-AABB broad phase, then the narrow phase on the
    // overlapping pairs. A brute-force AABB test over the world's candidate pairs is
    // result-equivalent to rapier's sweep-and-prune.
    PRL_D void detect_collisions(uint32_t moved, bool all) {
        // loosened AABBs of the body colliders
        float bmnx[NB], bmny[NB], bmxx[NB], bmxy[NB];
        {
            // player capsule: segment endpoints +- radius
            Iso pp = body_pose(0);
            v2 a = xform(pp, V(0.0f, -W.half_height)), b = xform(pp, V(0.0f, W.half_height));
            bmnx[0] = (fminf(a.x, b.x) - W.radius) - W.aabb_margin;
            bmny[0] = (fminf(a.y, b.y) - W.radius) - W.aabb_margin;
            bmxx[0] = (fmaxf(a.x, b.x) + W.radius) + W.aabb_margin;
            bmxy[0] = (fmaxf(a.y, b.y) + W.radius) + W.aabb_margin;
        }
#pragma unroll
        for (int b = 1; b < NB; ++b) {
            if (b < W.n_bodies) {
                int c = W.collider_of[b];
                float ac = fabsf(e.c[b]), as = fabsf(e.s[b]);
                float wx = ac * W.hx[c] + as * W.hy[c], wy = as * W.hx[c] + ac * W.hy[c];
                bmnx[b] = (e.x[b] - wx) - W.aabb_margin;
                bmny[b] = (e.y[b] - wy) - W.aabb_margin;
                bmxx[b] = (e.x[b] + wx) + W.aabb_margin;
                bmxy[b] = (e.y[b] + wy) + W.aabb_margin;
            } else {
                bmnx[b] = bmny[b] = bmxx[b] = bmxy[b] = 0.0f;
            }
        }
        auto pair_overlaps = [&](int p, bool& modified) -> bool {
            int c1 = W.pair_c1[p], c2 = W.pair_c2[p];
            int b1 = W.body_of[c1], b2 = W.body_of[c2];
            modified = all || (b1 >= 0 && ((moved >> b1) & 1u)) || (b2 >= 0 && ((moved >> b2) & 1u));
            float amnx, amny, amxx, amxy, qmnx, qmny, qmxx, qmxy;
            if (b1 >= 0) { amnx = sel(bmnx, b1); amny = sel(bmny, b1); amxx = sel(bmxx, b1); amxy = sel(bmxy, b1); }
            else { amnx = W.amnx[c1]; amny = W.amny[c1]; amxx = W.amxx[c1]; amxy = W.amxy[c1]; }
            if (b2 >= 0) { qmnx = sel(bmnx, b2); qmny = sel(bmny, b2); qmxx = sel(bmxx, b2); qmxy = sel(bmxy, b2); }
            else { qmnx = W.amnx[c2]; qmny = W.amny[c2]; qmxx = W.amxx[c2]; qmxy = W.amxy[c2]; }
            return amnx <= qmxx && qmnx <= amxx && amny <= qmxy && qmny <= amxy;
        };
        // 1. manifolds whose pair left the broad phase: NarrowPhase::remove_pair wakes both
        //    bodies if the pair had an active contact
        for (int k = e.count - 1; k >= 0; --k) {
            int slot = slot_at(k);
            int p = e.spair[slot];
            bool modified;
            bool ov = pair_overlaps(p, modified);
            if (modified && !ov) {
                bool active_contact = false;
                const float* f = e.mf[slot];
                if (e.snpts[slot] > 0 && f[8] < W.prediction) active_contact = true;
                if (e.snpts[slot] > 1 && f[13] < W.prediction) active_contact = true;
                if (active_contact) {
                    int b1 = W.body_of[W.pair_c1[p]], b2 = W.body_of[W.pair_c2[p]];
                    if (b1 >= 0) wake_strong(b1);
                    if (b2 >= 0) wake_strong(b2);
                }
                remove_at(k);
            }
        }
        // 2. overlapping pairs with a moved collider, in ascending pair order
        int cand[S];
        int ncand = 0;
        for (int p = 0; p < W.n_pairs; ++p) {
            bool modified;
            bool ov = pair_overlaps(p, modified);
            if (modified && ov) {
                if (ncand < S) cand[ncand++] = p;
                else e.flags |= FLAG_OVERFLOW;
            }
        }
        // 3. narrow phase; player pairs (capsule-cuboid) come first in pair order
        int k = 0;
        for (; k < ncand && cand[k] < W.n_player_pairs; ++k) {
            int p = cand[k];
            int c2 = W.pair_c2[p];
            int at = find(p);
            int slot = at >= 0 ? slot_at(at) : -1;
            Manifold m;
            if (slot >= 0) load_manifold(slot, m);
            else m.npts = 0;
            Iso pos12 = inv_mul(body_pose(0), collider_pose(c2));
            contact_capsule_cuboid(m, pos12, W.radius, W.half_height, V(W.hx[c2], W.hy[c2]), W.prediction);
            commit(p, at, slot, m);
        }
        if (NB > 1) {
            for (; k < ncand; ++k) {
                int p = cand[k];
                int c1 = W.pair_c1[p], c2 = W.pair_c2[p];
                int at = find(p);
                int slot = at >= 0 ? slot_at(at) : -1;
                Manifold m;
                if (slot >= 0) load_manifold(slot, m);
                else m.npts = 0;
                Iso pos12 = inv_mul(collider_pose(c1), collider_pose(c2));
                contact_cuboid_cuboid(m, pos12, V(W.hx[c1], W.hy[c1]), V(W.hx[c2], W.hy[c2]), W.prediction);
                commit(p, at, slot, m);
            }
        }
    }
    // write a narrow-phase result back to the slot table. Only manifolds with points are kept:
    // an empty manifold carries no state the next step can observe.
    PRL_D void commit(int p, int at, int slot, const Manifold& m) {
        if (m.npts > 0) {
            if (slot < 0) slot = insert(p);
            if (slot >= 0) store_manifold(slot, m);
        } else if (at >= 0) {
            remove_at(at);
        }
    }

    // common.rs:175-195 distance_to_goals, editor pixels; +inf for a goal-less world
    PRL_D float distance_to_goals() const {
        float best = f32_inf();
        for (int i = 0; i < W.n_goals; ++i) {
            v2 d = V(e.x[0] - W.gx[i], e.y[0] - W.gy[i]);
            float dx = fmaxf(fabsf(d.x * W.gxax[i] + d.y * W.gxay[i]) - W.ghw[i], 0.0f);
            float dy = fmaxf(fabsf(d.x * W.gyax[i] + d.y * W.gyay[i]) - W.ghh[i], 0.0f);
            float dist = sqrtf(dx * dx + dy * dy) / W.scale;
            best = i == 0 ? dist : fminf(best, dist);
        }
        return best;
    }

    // Environment::step (common.rs:215-340); returns distance_to_goals after the step
    PRL_D float step(uint32_t mv) {
        for (int k = 0; k < e.count; ++k) solver_contacts(slot_at(k));
        player_move(mv);
        // first step only: every collider is new, so rapier detects collisions on the initial
        // poses before anything moves (SURVEY A.3 stage 1)
        int pass = (e.flags & FLAG_FRESH) ? 0 : 1;
        for (; pass < 2; ++pass) {
            uint32_t moved = 0;
            if (pass == 1) moved = solve_and_integrate();
            detect_collisions(moved, pass == 0);
            if (pass == 0) {
                e.flags &= ~FLAG_FRESH;
                for (int k = 0; k < e.count; ++k) solver_contacts(slot_at(k));
            }
        }
        // query_pipeline.update (common.rs:330-331) is never queried: omitted
        float d = distance_to_goals();
        if (!(e.flags & FLAG_WON) && W.n_goals > 0 && d < 1e-7f) e.flags |= FLAG_WON;
        return d;
    }
};

}  // namespace prl
