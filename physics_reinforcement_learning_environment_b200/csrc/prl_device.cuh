// prl_device.cuh -- one environment per thread: the whole of `Environment::step(Move)`
// (reference src/common.rs:215-340) and of the rapier2d 0.17.2 `PhysicsPipeline::step` it calls
// (common.rs:315-329), for the scene class `from_world` can build:
// one rotation-locked capsule (the player), parent-less cuboids (fixed blocks), dynamic cuboids
// (pushable blocks) and goal rectangles.
//
// Design (B200-first, not a translation of rapier's object graph):
//  * body state lives in registers for the whole launch; a launch runs 1..number_of_steps steps;
//  * the world's static colliders / pairs / goals are read from a shared-memory DevWorld;
//  * contact manifolds live in a small per-thread slot table (<= MAXC live manifolds): slot ==
//    pair index and bit masks for worlds with <= MAXC candidate pairs (every shipped level), else
//    arbitrary slots kept sorted by pair index through a 64-bit nibble list; either way the
//    Gauss-Seidel order is the canonical ascending (collider1, collider2) order and no manifold
//    data ever moves;
//  * the solver's constraint rows live in shared memory, one column per lane (conflict-free);
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
// rarely taken paths are kept out of line: the step loop is instruction-fetch bound, so the hot
// path should be as few cache lines as possible
#define PRL_COLD static __host__ __device__ __noinline__
#else
#define PRL_D inline
#define PRL_COLD static inline
#endif
// rarely taken branches: laid out away from the hot path (the step is instruction-fetch sensitive)
#define PRL_UNLIKELY(x) __builtin_expect(!!(x), 0)

// PRL_PHASE_TIMING (profiling builds only, never the product): per-warp cycle counts of the phases
// of a step, printed by block 0 of the rollout kernel
#if defined(PRL_PHASE_TIMING) && defined(__CUDA_ARCH__)
#define PRL_TICK(i) do { long long c_ = clock64(); tacc[i] += c_ - tlast; tlast = c_; } while (0)
#else
#define PRL_TICK(i) do { } while (0)
#endif

// PRL_HOST_TRACE (tests/hostsim only): per-step work counters for offline studies of how
// environments should be grouped into warps
#if defined(PRL_HOST_TRACE)
#define PRL_TRACE(stmt) do { stmt; } while (0)
#else
#define PRL_TRACE(stmt) do { } while (0)
#endif

// PRL_BOUNDS_CHECK (libprl_debug.so, run by tests/test_bounds_gpu.py; never the product build): every
// run-time index into the row planes, the manifold records, the world tables, the state arrays and the
// genomes is checked, and a failed check sets bit `code` of a device word that
// prl_debug_bounds_violations() reads back. compute-sanitizer is not available on the GPU pool.
#if defined(PRL_BOUNDS_CHECK) && defined(__CUDACC__)
__device__ unsigned int g_prl_bounds_violations;
#endif
#if defined(PRL_BOUNDS_CHECK) && defined(__CUDA_ARCH__)
#define PRL_CHECK(cond, code) do { if (!(cond)) atomicOr(&g_prl_bounds_violations, 1u << (code)); } while (0)
#else
#define PRL_CHECK(cond, code) do { } while (0)
#endif
// codes: 0 row entry read, 1 row entry written, 2 manifold slot, 3 candidate pair, 4 state slot in HBM,
//        5 gene, 6 regroup table, 7 body

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
// maximum of v over the lanes of the warp (all 32 lanes must call it); the host build is a single "lane"
PRL_D int warp_max(int v) {
#if defined(__CUDA_ARCH__)
    return (int)__reduce_max_sync(0xffffffffu, (unsigned)v);
#else
    return v;
#endif
}
PRL_D int popcount(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
PRL_D int first_bit_from(uint32_t mask, int from) {  // lowest set bit at or above `from`, 32 if none
    uint32_t m = from >= 32 ? 0u : (mask >> from) << from;
#if defined(__CUDA_ARCH__)
    return m ? __ffs((int)m) - 1 : 32;
#else
    for (int i = 0; i < 32; ++i) if ((m >> i) & 1u) return i;
    return 32;
#endif
}
PRL_D bool warp_all(bool p) {
#if defined(__CUDA_ARCH__)
    return __all_sync(0xffffffffu, p) != 0;
#else
    return p;
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
    // Both ends of the overlap: the end that lies inside the other segment is interpolated on it,
    // bc = (r - ra) * inv0(rb - ra). One expression (and one division) per end; which segment is
    // interpolated is a select of the operands, the arithmetic is parry's in either case.
    const float inv1 = inv0(r1b - r1a), inv2 = inv0(r2b - r2a);
    v2 c0, c1, c2, c3;
    {
        const bool on1 = r2a > r1a;  // the low end of segment 2 falls inside segment 1
        const float bc = on1 ? (r2a - r1a) * inv1 : (r1a - r2a) * inv2;
        const v2 base = on1 ? s1a : s2a, span = on1 ? s1b - s1a : s2b - s2a;
        const v2 ip = base + span * bc;
        c0 = on1 ? ip : s1a;
        c1 = on1 ? s2a : ip;
    }
    {
        const bool on1 = r2b < r1b;  // the high end of segment 2 falls inside segment 1
        const float bc = on1 ? (r2b - r1a) * inv1 : (r1b - r2a) * inv2;
        const v2 base = on1 ? s1a : s2a, span = on1 ? s1b - s1a : s2b - s2a;
        const v2 ip = base + span * bc;
        c2 = on1 ? ip : s1b;
        c3 = on1 ? s2b : ip;
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
// (the out-of-line routines take and return everything BY VALUE: a local whose address is passed
// to a call lives in memory, hot path included)
PRL_COLD Manifold contact_cuboid_cuboid_full(Iso pos12, v2 he1, v2 he2, float prediction) {
    Manifold m = {};
    Iso pos21 = inverse(pos12);
    v2 d1, d2;
    float sep1 = cuboid_sat_oneway(he1, he2, pos12, d1);
    if (sep1 > prediction) return m;
    float sep2 = cuboid_sat_oneway(he2, he1, pos21, d2);
    if (sep2 > prediction) return m;
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
    return m;
}

// Closest points between the capsule's inner segment [a,b] and the box [-he,he], in the box
// frame: exact closest-feature enumeration (2 segment endpoints vs the box, 4 box corners vs the
// segment; fixed order, strict '<'). parry reaches the same pair with GJK up to its tolerance.
// Returns false if the segment touches the box.
// (Rolled loops on purpose: this runs for a few lanes of nearly every warp-step, so what it costs is
// instruction-cache footprint, not issue slots -- profiles/r02_experiments.md.)
PRL_D bool segment_box_closest(v2 a, v2 b, v2 he, float& u_out, v2& q_out) {
    v2 d = b - a;
    {
        float t0 = 0.0f, t1 = 1.0f;
        bool hit = true;
#pragma unroll 1
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
    const float idd2 = inv0(dot(d, d));
    // candidates in fixed order with strict '<': the two segment endpoints against the box, then the
    // four box corners (-,-), (+,-), (+,+), (-,+) against the segment
#pragma unroll 1
    for (int c = 0; c < 6; ++c) {
        v2 q, r;
        float u;
        if (c < 2) {
            v2 p = c == 0 ? a : b;
            q = V(fminf(fmaxf(p.x, -he.x), he.x), fminf(fmaxf(p.y, -he.y), he.y));
            r = q - p;
            u = c == 0 ? 0.0f : 1.0f;
        } else {
            const int kc = c - 2;
            q = V((kc == 1 || kc == 2) ? he.x : -he.x, (kc >= 2) ? he.y : -he.y);
            u = dot(q - a, d) * idd2;
            u = fminf(fmaxf(u, 0.0f), 1.0f);
            v2 ps = a + d * u;
            r = q - ps;
        }
        float dd = dot(r, r);
        if (dd < best) { best = dd; bu = u; bq = q; }
    }
    u_out = bu;
    q_out = bq;
    return true;
}

// parry's EPA branch of contact_manifold_pfm_pfm: the capsule's inner segment is inside the box.
// Minimum face penetration of the deeper endpoint (unreachable at this game's speeds except by a
// level that spawns the player inside a block). Returns (p1 | p2_1) and writes dir.
struct InsideBox { v2 p1, p2_1, dir; };
PRL_COLD InsideBox capsule_inside_box(Iso pos12, v2 a2, v2 b2, v2 he2) {
    float best = kFltMax;
    v2 nb = V(0.0f, 1.0f), pe = a2, qf = a2;
#pragma unroll 1
    for (int f = 0; f < 4; ++f) {
        v2 nf = f == 0 ? V(1.0f, 0.0f) : f == 1 ? V(-1.0f, 0.0f) : f == 2 ? V(0.0f, 1.0f) : V(0.0f, -1.0f);
        float hf = f < 2 ? he2.x : he2.y;
        float da = dot(a2, nf), db = dot(b2, nf);
        v2 e = da < db ? a2 : b2;
        float depth = hf - fminf(da, db);
        if (depth < best) { best = depth; nb = nf; pe = e; qf = e + nf * depth; }
    }
    InsideBox r;
    r.p1 = xform(pos12, pe);
    r.p2_1 = xform(pos12, qf);
    r.dir = -rot(pos12.c, pos12.s, nb);
    return r;
}

// parry contact_manifold_pfm_pfm with pfm1 = the capsule's segment (border radius r) and
// pfm2 = the cuboid: closest points, then segment-vs-support-face clipping, then the radius shift.
PRL_COLD Manifold contact_capsule_cuboid_full(Iso pos12, float radius, float hh, v2 he2, float prediction) {
    Manifold m = {};
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
        if (n > total_prediction) return m;
        dir = V(dp.x / n, dp.y / n);
    } else {
        const InsideBox in = capsule_inside_box(pos12, a2, b2, he2);
        p1 = in.p1; p2_1 = in.p2_1; dir = in.dir;
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
    return m;
}

// hot path of both contact routines: the cached manifold usually survives (persistence first)
PRL_D void contact_cuboid_cuboid(Manifold& m, const Iso& pos12, v2 he1, v2 he2, float prediction) {
    if (!manifold_try_update(m, pos12)) m = contact_cuboid_cuboid_full(pos12, he1, he2, prediction);
}
PRL_D void contact_capsule_cuboid(Manifold& m, const Iso& pos12, float radius, float hh, v2 he2, float prediction) {
    if (!manifold_try_update(m, pos12)) m = contact_capsule_cuboid_full(pos12, radius, hh, he2, prediction);
}

// sin and cos of a pushable block's rotation increment; kept out of line so the pose update of
// every body shares one copy of the argument reduction
#if defined(__CUDA_ARCH__)
static __device__ __noinline__ v2 libm_sincosf(float x) { return V(libm_sinf(x), libm_cosf(x)); }  // (sin, cos)
#else
inline v2 libm_sincosf(float x) { return V(libm_sinf(x), libm_cosf(x)); }
#endif

// rapier RigidBodyVelocity::integrate with local_com = origin:
// shift * Isometry::new(linvel*dt, angvel*dt) * shift^-1 * init, then renormalize_fast.
// `locked`: rotations are locked (the player), so angvel is identically zero and the rotation
// increment is exactly (cos, sin) = (1, 0).
PRL_D Iso integrate_pose(v2 linvel, float angvel, const Iso& init, float dt, bool locked) {
    v2 com = init.t;
    float rc = 1.0f, rs = 0.0f;
    if (!locked) { v2 sc = libm_sincosf(angvel * dt); rs = sc.x; rc = sc.y; }
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
// body arrays stay in registers; S = manifold slots. Everything indexed at run time lives in
// thread-local memory as 16-byte rows, so one record moves with one LDL.128 / STL.128.
struct alignas(16) f4 { float x, y, z, w; };

PRL_D float bits_f32(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

template <int NB, int S = MAXC>
struct Env {
    // rigid bodies (0 = player)
    float x[NB], y[NB], c[NB], s[NB], vx[NB], vy[NB], w[NB], timer[NB];
    uint32_t sleeping;   // bit b
    uint32_t flags;      // FLAG_WON | FLAG_FRESH | FLAG_OVERFLOW
    // manifold slot table. Worlds with at most S candidate pairs use slot == pair index ("direct":
    // `used` is the set of pairs with a manifold; order / count / spair are dead). Larger worlds
    // keep up to S live manifolds in arbitrary slots, sorted by pair through `order`.
    uint64_t order;      // nibble k = slot of the k-th live manifold in ascending pair order
    int count;
    uint32_t used;       // bit s = slot s is live
    uint64_t spair[2];   // byte s = pair index of slot s (registers: the lists are walked every stage)
    uint32_t ns1, ns2;   // bit s = slot s has at least one / has two solver contacts
};

// The contact records of one environment, indexed at run time and therefore in thread-local
// memory (16-byte rows, so one record moves with one LDL.128 / STL.128). They are a separate
// object from Env on purpose: a struct that is indexed with a run-time subscript anywhere is
// kept in memory as a whole by the compiler, and the body state must stay in registers.
template <int S = MAXC>
struct Contacts {
    // mf[s]: the cached contact manifold of one collider pair
    //   [0] = (hdr, n1.x, n1.y, n2.x)      hdr bits = pair | npts << 8 | n_solver_contacts << 12
    //   [1] = (n2.y, p1a.x, p1a.y, p2a.x)
    //   [2] = (p2a.y, dist_a, p1b.x, p1b.y)
    //   [3] = (p2b.x, p2b.y, dist_b, -)
    f4 mf[S][4];
    // sc[s]: its solver contacts (world space), refreshed whenever the narrow phase visits the pair
    //   [0] = (normal.x, normal.y, pt0.x, pt0.y)   [1] = (dist0, pt1.x, pt1.y, dist1)
    f4 sc[S][2];
    // S > 16 ("wide": worlds with pushable blocks and more than 16 candidate pairs): the sorted manifold
    // list, the slot -> pair table, the step's constraint list and the narrow-phase candidates no longer
    // fit 64-bit registers as nibbles / bytes and live here, one byte per entry
    uint8_t order[S > 16 ? S : 1], spair[S > 16 ? S : 1], cons[S > 16 ? S : 1], cand[S > 16 ? S : 1];
};

struct alignas(8) f2 { float x, y; };

// manifold slots per environment: MAXC, or twice that ("wide" lists in thread-local bytes, Stepper::kWide)
// for worlds with pushable blocks and more candidate pairs than the slot == pair tables hold
template <int NB, class WT> struct SlotsOf { static constexpr int value = (NB > 1 && WT::kMaxPairs > MAXC) ? MAXC_WIDE : MAXC; };

template <int NB, int S, class WT = DevWorld, int CAP = 8>
struct Stepper {
    using E = Env<NB, S>;
    static constexpr int kGround = NB;  // pseudo body index of the fixed world: zero velocity, zero inverse mass
    // Velocity-constraint rows, one per solver contact, in Gauss-Seidel order (per-step scratch).
    // The 14 solver sweeps re-read every row, so rows live in shared memory, and the shared-memory
    // data pipe is what bounds the sweeps: a row is therefore split into planes that are each
    // read with the widest conflict-free access, and a lane's rows sit in the lane's own COLUMN
    // (entry = row * 32 + lane), so that every access of a warp covers consecutive addresses:
    //   plane 0  f4  (dir.x, dir.y, gcross2, t_gcross2)                  both sweeps
    //   plane 1  f4  (rhs, r, rhs_wo_bias, softness)                     normal sweeps
    //                softness = the constraint's cfm_factor as a float (NB <= 2), or the bits
    //                body1 | body2 << 8 | rigid << 16 (NB > 2)
    //   plane 2  f2  (t_rhs, t_r)                                        friction sweeps
    //   plane 3  f   accumulated normal impulse                          read + written
    //   plane 4  f   accumulated tangent impulse                         read + written
    //   plane 5  f2  (gcross1, t_gcross1)  body 1's angular terms        only with NB > 2
    // With at most one pushable block (NB <= 2) the only two-body rows are the <= 2 rows of the
    // player-block manifold, whose plane 5 stays in registers (tb_g1 / tb_tg1), and a ground
    // row's body follows from its position in the row list (player pairs sort first).
    // A column holds CAP rows. If an environment of the warp needs more, the warp solves in G
    // rounds (G = 2, 4, 8; see the kernels): the 32 / G lanes of a round re-shape the planes into
    // 32 / G columns of G * CAP rows. Either way row j of a lane is entry ent0 + j * stride with
    // stride = 32 / G and ent0 = the lane's column, so the lanes of a round always touch
    // consecutive entries. G == 1 in all but a fraction of a percent of the steps.
    static constexpr int kCap = CAP;
    static_assert(CAP >= 4 && CAP <= 10, "column height");
    static constexpr int kEntries = 32 * CAP;   // entries per plane per warp
    static constexpr int kRowBytes = 48 + (NB > 2 ? 8 : 0);
    static constexpr int kWarpBytes = kEntries * kRowBytes;
    static constexpr int kOff1 = 16 * kEntries, kOff2 = 32 * kEntries, kOff3 = 40 * kEntries, kOff4 = 44 * kEntries,
                         kOff5 = 48 * kEntries;
    static constexpr int kMaxRows = 2 * S;  // rows one environment can need
    static_assert(kMaxRows <= 8 * CAP, "eight columns must hold every row of one environment");
    const WT& W;
    E& e;
    Contacts<S>& k;
    unsigned char* pool;    // the warp's planes
    int ent0, stride;       // entry of my row 0 and distance between my rows, this step
    float cfm_reg;          // cfm_factor
    float tb_g1[2], tb_tg1[2];
    // constraint selection of the current step (begin_step -> finish_step)
    uint32_t active;
    uint64_t cons;      // nibble i = slot of my i-th constraint
    uint32_t cm_tb, cm_gr;  // slot == pair worlds instead: masks of the two-body / one-body manifolds that become constraints
    int ncons, need, ntb, nplayer;   // constraints, rows, two-body rows (the head of the row list), rows of player pairs
    float im_reg[NB];       // inverse masses (registers: every row update needs one, off the critical path)
    // collision detection owed for the poses of the last position update (detect_pending)
    uint32_t pend_moved;
    bool pend_all, first_step;
    // Ground-contact scan (common.rs:216-244) gathered while the narrow phase has the player's
    // solver contacts in registers: the first two floor contacts in the reference's iteration
    // order (pair order, then point order). gs_n = their number (more than two: count only), or -1
    // if the narrow phase did not visit the player's pairs since the last scan (player asleep,
    // state just loaded): player_move then reads the stored solver contacts instead.
    int gs_n;
    float gs_x[2], gs_y[2];
    int gs_b[2];
#if defined(PRL_PHASE_TIMING)
    long long tacc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tlast = 0;
#endif
#if defined(PRL_HOST_TRACE)
    int tr_cand = 0, tr_full = 0;  // narrow-phase pairs visited / recomputed from scratch by the last detection
#endif
    PRL_D Stepper(const WT& w, E& env, Contacts<S>& contacts, void* warp_pool)
        : W(w), e(env), k(contacts), pool(static_cast<unsigned char*>(warp_pool)), ent0(0), stride(1), cfm_reg(w.cfm_factor), active(0), cons(0), cm_tb(0), cm_gr(0), ncons(0), need(0), ntb(0),
          nplayer(0), pend_moved(0), pend_all((env.flags & FLAG_FRESH) != 0), first_step(false), gs_n(-1) {
        gs_x[0] = gs_x[1] = gs_y[0] = gs_y[1] = 0.0f;
        gs_b[0] = gs_b[1] = -1;
        tb_g1[0] = tb_g1[1] = tb_tg1[0] = tb_tg1[1] = 0.0f;
#pragma unroll
        for (int b = 0; b < NB; ++b) im_reg[b] = w.inv_mass[b];
    }
    // The sweeps fetch one row ahead without a bounds check: the entry after a lane's last row may
    // lie up to 32 entries past a plane (it is never used), so a block's planes are followed by kPoolPad bytes.
    static constexpr int kPoolPad = 256;
    PRL_D int row_entry(int j) const { return ent0 + j * stride; }
    // (reads may run one row -- 32 entries -- past a plane: the sweeps fetch one row ahead, kPoolPad)
    PRL_D f4& q0(int en) const { PRL_CHECK(en >= 0 && en < kEntries + 32, 0); return reinterpret_cast<f4*>(pool)[en]; }
    PRL_D f4& q1(int en) const { PRL_CHECK(en >= 0 && en < kEntries + 32, 0); return reinterpret_cast<f4*>(pool + kOff1)[en]; }
    PRL_D f2& q2(int en) const { PRL_CHECK(en >= 0 && en < kEntries + 32, 0); return reinterpret_cast<f2*>(pool + kOff2)[en]; }
    PRL_D float& acc_n(int en) const { PRL_CHECK(en >= 0 && en < kEntries + 32, 0); return reinterpret_cast<float*>(pool + kOff3)[en]; }
    PRL_D float& acc_t(int en) const { PRL_CHECK(en >= 0 && en < kEntries + 32, 0); return reinterpret_cast<float*>(pool + kOff4)[en]; }
    PRL_D f2& q5(int en) const { PRL_CHECK(en >= 0 && en < kEntries + 32, 0); return reinterpret_cast<f2*>(pool + kOff5)[en]; }

    // slot == pair index for worlds with at most S candidate pairs (the shipped levels): the live
    // manifolds are a bit mask, and selection / broad phase are mask algebra
    static constexpr bool kDirect = WT::kMaxPairs <= S;
    static constexpr int kSlots = S;
    static constexpr bool kWide = S > 16;  // lists in thread-local bytes instead of register nibbles
    static_assert(S <= 32, "the slot masks are 32 bits wide");
    PRL_D int slot_at(int i) const { return kWide ? (int)k.order[i] : (int)((e.order >> (4 * i)) & 15u); }
    PRL_D int pair_of(int slot) const {
        if (kDirect) return slot;
        if (kWide) return (int)k.spair[slot];
        return (int)(((slot < 8 ? e.spair[0] : e.spair[1]) >> (8 * (slot & 7))) & 0xffu);
    }
    PRL_D void set_pair(int slot, int pair) {
        if (kDirect) return;
        if (kWide) { k.spair[slot] = (uint8_t)pair; return; }
        uint64_t sh = 8 * (slot & 7);
        if (slot < 8) e.spair[0] = (e.spair[0] & ~(0xffull << sh)) | ((uint64_t)pair << sh);
        else e.spair[1] = (e.spair[1] & ~(0xffull << sh)) | ((uint64_t)pair << sh);
    }
    // the step's constraints (slots, two-body first) and the narrow-phase candidates (pair indices)
    PRL_D int cons_at(int i) const { return kWide ? (int)k.cons[i] : (int)((cons >> (4 * i)) & 15u); }
    PRL_D void cons_push(int slot) {
        if (kWide) k.cons[ncons] = (uint8_t)slot;
        else cons |= (uint64_t)slot << (4 * ncons);
        ++ncons;
    }
    // iteration over the live manifolds in ascending pair order: for (it = it_begin(); it_ok(it); it = it_next(it))
    PRL_D int it_begin() const { return kDirect ? first_bit_from(e.used, 0) : 0; }
    PRL_D bool it_ok(int it) const { return kDirect ? it < 32 : it < e.count; }
    PRL_D int it_next(int it) const { return kDirect ? first_bit_from(e.used, it + 1) : it + 1; }
    PRL_D int it_slot(int it) const { return kDirect ? it : slot_at(it); }
    PRL_D int live_count() const {
        if (!kDirect) return e.count;
        int c = 0;
        for (int it = it_begin(); it_ok(it); it = it_next(it)) ++c;
        return c;
    }
    PRL_D int ns_of(int slot) const { return (int)(((e.ns1 >> slot) & 1u) + ((e.ns2 >> slot) & 1u)); }
    PRL_D void set_ns(int slot, int ns) {
        const uint32_t bit = 1u << slot;
        e.ns1 = ns >= 1 ? (e.ns1 | bit) : (e.ns1 & ~bit);
        e.ns2 = ns >= 2 ? (e.ns2 | bit) : (e.ns2 & ~bit);
    }
    // bodies of pair p's colliders from the packed pair word: -1 for a fixed collider
    PRL_D static int info_b1(uint32_t info) { return (int)(int8_t)((info >> 16) & 0xffu); }
    PRL_D static int info_b2(uint32_t info) { return (int)(int8_t)(info >> 24); }
    PRL_D static int info_c1(uint32_t info) { return (int)(info & 0xffu); }
    PRL_D static int info_c2(uint32_t info) { return (int)((info >> 8) & 0xffu); }
    PRL_D int pair_at(int k) const { return pair_of(slot_at(k)); }
    PRL_D Iso body_pose(int b) const { PRL_CHECK(b >= 0 && b < NB, 7); Iso p; p.t = V(sel(e.x, b), sel(e.y, b)); p.c = sel(e.c, b); p.s = sel(e.s, b); return p; }
    PRL_D Iso collider_pose(int c) const {
        int b = W.body_of[c];
        if (b >= 0) return body_pose(b);
        Iso p; p.t = V(W.cx[c], W.cy[c]); p.c = W.cc[c]; p.s = W.cs[c];
        return p;
    }
    // pose of pair p's collider with body index b (b < 0: the pair's fixed collider)
    PRL_D Iso pair_collider_pose(int p, int b) const {
        if (b >= 0) return body_pose(b);
        const f4 q = *reinterpret_cast<const f4*>(W.pair_pose[p]);
        Iso r; r.t = V(q.x, q.y); r.c = q.z; r.s = q.w;
        return r;
    }
    PRL_D void wake_strong(int b) { e.sleeping &= ~(1u << b); put(e.timer, b, 0.0f); }

    PRL_D void load_manifold(int slot, Manifold& m) const {
        PRL_CHECK(slot >= 0 && slot < S, 2);
        f4 a = k.mf[slot][0], b = k.mf[slot][1], c = k.mf[slot][2], d = k.mf[slot][3];
        m.npts = (int)((f32_bits(a.x) >> 8) & 0xfu);
        m.n1 = V(a.y, a.z); m.n2 = V(a.w, b.x);
        m.p1[0] = V(b.y, b.z); m.p2[0] = V(b.w, c.x); m.dist[0] = c.y;
        m.p1[1] = V(c.z, c.w); m.p2[1] = V(d.x, d.y); m.dist[1] = d.z;
    }
    // store a manifold with points and refresh its solver contacts (rapier
    // NarrowPhase::compute_contacts tail: points with dist < prediction, at the midpoint between
    // the two surfaces, in world space)
    // scan: the pair is one of the player's -- its floor contacts feed the ground-contact scan
    template <bool SCAN = false>
    PRL_D void store_manifold(int slot, int pair, const Manifold& m, const Iso& pose1, bool scan = SCAN) {
        PRL_CHECK(slot >= 0 && slot < S, 2);
        PRL_CHECK(pair >= 0 && pair < W.n_pairs, 3);
        v2 normal = rot(pose1.c, pose1.s, m.n1);
        int ns = 0;
        float px[2] = {0.0f, 0.0f}, py[2] = {0.0f, 0.0f}, pd[2] = {0.0f, 0.0f};
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            if (i < m.npts && m.dist[i] < W.prediction) {
                v2 wp = xform(pose1, m.p1[i]);
                v2 half = normal * m.dist[i];
                float qx = wp.x + half.x / 2.0f, qy = wp.y + half.y / 2.0f;
                if (ns == 0) { px[0] = qx; py[0] = qy; pd[0] = m.dist[i]; }
                else { px[1] = qx; py[1] = qy; pd[1] = m.dist[i]; }
                ++ns;
            }
        }
        if (scan && gs_n >= 0) {  // a player pair: floor contacts among its solver contacts (common.rs:234-238)
            const int body = info_b2(W.pair_info[pair]);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                if (i < ns) {
                    float qy = (py[i] - (e.y[0] - W.lower_off)) / W.radius;
                    if (qy < -0.707f) {
                        if (gs_n == 0) { gs_x[0] = px[i]; gs_y[0] = py[i]; gs_b[0] = body; }
                        else if (gs_n == 1) { gs_x[1] = px[i]; gs_y[1] = py[i]; gs_b[1] = body; }
                        ++gs_n;
                    }
                }
            }
        }
        f4 a, b, c, d;
        a.x = bits_f32((uint32_t)pair | ((uint32_t)m.npts << 8));
        set_ns(slot, ns);
        set_pair(slot, pair);
        a.y = m.n1.x; a.z = m.n1.y; a.w = m.n2.x;
        b.x = m.n2.y; b.y = m.p1[0].x; b.z = m.p1[0].y; b.w = m.p2[0].x;
        c.x = m.p2[0].y; c.y = m.dist[0]; c.z = m.p1[1].x; c.w = m.p1[1].y;
        d.x = m.p2[1].x; d.y = m.p2[1].y; d.z = m.dist[1]; d.w = 0.0f;
        k.mf[slot][0] = a; k.mf[slot][1] = b; k.mf[slot][2] = c; k.mf[slot][3] = d;
        f4 s0, s1;
        s0.x = normal.x; s0.y = normal.y; s0.z = px[0]; s0.w = py[0];
        s1.x = pd[0]; s1.y = px[1]; s1.z = py[1]; s1.w = pd[1];
        k.sc[slot][0] = s0; k.sc[slot][1] = s1;
    }
    // after a load from HBM: solver contacts are a pure function of the manifold and collider 1's
    // pose, so they are not part of the stored state
    PRL_D void refresh_solver_contacts() {
        gs_n = -1;
#pragma unroll 1
        for (int it = it_begin(); it_ok(it); it = it_next(it)) {
            int slot = it_slot(it);
            Manifold m;
            load_manifold(slot, m);
            int p = (int)(f32_bits(k.mf[slot][0].x) & 0xffu);
            store_manifold(slot, p, m, collider_pose(W.pair_c1[p]));
        }
    }
    PRL_D void remove_at(int i) {
        int slot = slot_at(i);
        e.used &= ~(1u << slot);
        if (kWide) {
#pragma unroll 1
            for (int j = i; j + 1 < e.count; ++j) k.order[j] = k.order[j + 1];
        } else {
            uint64_t low = i == 0 ? 0ull : (e.order & (~0ull >> (64 - 4 * i)));
            uint64_t high = i >= 15 ? 0ull : (e.order >> (4 * (i + 1))) << (4 * i);
            e.order = low | high;
        }
        --e.count;
    }
    // allocate a slot and link it at sorted position i; -1 on overflow
    PRL_D int insert_at(int i) {
        int slot = first_zero_bit(e.used);
        if (slot >= S) { e.flags |= FLAG_OVERFLOW; return -1; }
        if (kWide) {
#pragma unroll 1
            for (int j = e.count; j > i; --j) k.order[j] = k.order[j - 1];
            k.order[i] = (uint8_t)slot;
        } else {
            uint64_t low = i == 0 ? 0ull : (e.order & (~0ull >> (64 - 4 * i)));
            uint64_t high = i == 0 ? e.order : (e.order >> (4 * i));
            e.order = low | ((uint64_t)slot << (4 * i)) | (i >= 15 ? 0ull : (high << (4 * (i + 1))));
        }
        e.used |= 1u << slot;
        ++e.count;
        return slot;
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

    // common.rs:216-313: ground-contact scan over the player's contact pairs, then the move
    // impulses. The scan normally comes from the narrow phase (gs_*, see store_manifold).
    PRL_D void player_move(uint32_t mv) {
        if (PRL_UNLIKELY(gs_n < 0 || gs_n > 2)) { player_move_stored(mv); return; }
        const int nfc = gs_n;
        if (nfc == 0) return;
        v2 lower = V(e.x[0], e.y[0] - W.lower_off);
        // min_by keeps the first minimum, max_by the last maximum (common.rs:249-254, 273-278)
        const bool second_min = nfc == 2 && gs_x[1] < gs_x[0];
        const bool second_max = nfc == 2 && !(gs_x[1] < gs_x[0]);
        v2 player_impulse = V(0.0f, 0.0f);
        const float cnt = (float)nfc;
        // The reference's three blocks -- left: tangent impulse at the min-x floor contact
        // (common.rs:249-271), right: at the max-x one (:273-295), up: the jump impulse shared by all
        // floor contacts (:297-310) -- as ONE loop over up to four impulse applications in that order
        // (one copy of the code: the step is bound by its instruction footprint). Per application:
        //   n = normalize(point - lower centre)
        //   left  J = ( 0.003 n.y, -0.003 n.x)    right J = (-0.003 n.y, 0.003 n.x)    up J = -0.1 n / count
        //   a dynamic contact body gets -J at the point; the player accumulates J
        uint32_t todo = (mv & 3u) | ((mv & 4u) ? (nfc == 2 ? 12u : 4u) : 0u);
#pragma unroll 1
        for (; todo != 0; todo &= todo - 1) {
            const int a = first_bit_from(todo, 0);
            const bool second = a == 0 ? second_min : (a == 1 ? second_max : a == 3);
            const v2 pt = second ? V(gs_x[1], gs_y[1]) : V(gs_x[0], gs_y[0]);
            const int body = second ? gs_b[1] : gs_b[0];
            v2 n = pt - lower;
            const float len = sqrtf(n.x * n.x + n.y * n.y);
            n = V(n.x / len, n.y / len);
            const bool side = a < 2;
            const float cx = a == 0 ? 0.003f : (a == 1 ? -0.003f : -0.1f), cy = a == 0 ? -0.003f : (a == 1 ? 0.003f : -0.1f);
            v2 imp = V(cx * (side ? n.y : n.x), cy * (side ? n.x : n.y));
            if (!side) imp = V(imp.x / cnt, imp.y / cnt);
            if (NB > 1 && body >= 0) apply_impulse_at_point(body, -imp, pt);
            player_impulse = player_impulse + imp;
        }
        apply_impulse(0, player_impulse);
    }
    // The same from the stored solver contacts (the narrow phase did not visit the player's pairs
    // since the last scan, or there are more than two floor contacts). The player's pairs are the
    // head of the sorted manifold list.
    PRL_D void player_move_stored(uint32_t mv) {
        v2 lower = V(e.x[0], e.y[0] - W.lower_off);
        int nfc = 0;
        v2 pmin = V(0.0f, 0.0f), pmax = V(0.0f, 0.0f);
        int bmin = -1, bmax = -1;
#pragma unroll 1
        for (int it = it_begin(); it_ok(it); it = it_next(it)) {
            int slot = it_slot(it);
            int p = pair_of(slot), ns = ns_of(slot);
            if (p >= W.n_player_pairs) break;
            int body = info_b2(W.pair_info[p]);
            f4 s0 = k.sc[slot][0], s1 = k.sc[slot][1];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                if (i < ns) {
                    v2 pt = i == 0 ? V(s0.z, s0.w) : V(s1.y, s1.z);
                    v2 q = pt - lower;
                    q = V(q.x / W.radius, q.y / W.radius);
                    if (q.y < -0.707f) {
                        if (nfc == 0 || pt.x < pmin.x) { pmin = pt; bmin = body; }    // min_by: first minimum
                        if (nfc == 0 || !(pt.x < pmax.x)) { pmax = pt; bmax = body; } // max_by: last maximum
                        ++nfc;
                    }
                }
            }
        }
        if (nfc == 0) return;
        v2 player_impulse = V(0.0f, 0.0f);
        // left / right: tangent impulse at the extreme floor contact (common.rs:249-295)
#pragma unroll 1
        for (int side = 0; side < 2; ++side) {
            if (mv & (1u << side)) {
                v2 pt = side == 0 ? pmin : pmax;
                int body = side == 0 ? bmin : bmax;
                v2 n = pt - lower;
                float len = sqrtf(n.x * n.x + n.y * n.y);
                n = V(n.x / len, n.y / len);
                v2 imp = side == 0 ? V(0.003f * n.y, -0.003f * n.x) : V(-0.003f * n.y, 0.003f * n.x);
                if (NB > 1 && body >= 0) apply_impulse_at_point(body, -imp, pt);
                player_impulse = player_impulse + imp;
            }
        }
        if (mv & 4u) {  // up: the jump impulse is shared by all floor contacts (common.rs:297-310)
            float cnt = (float)nfc;
#pragma unroll 1
            for (int it = it_begin(); it_ok(it); it = it_next(it)) {
                int slot = it_slot(it);
                int ns = ns_of(slot);
                if (pair_of(slot) >= W.n_player_pairs) break;
                int body = info_b2(W.pair_info[pair_of(slot)]);
                f4 s0 = k.sc[slot][0], s1 = k.sc[slot][1];
#pragma unroll 1
                for (int i = 0; i < ns; ++i) {
                    v2 pt = i == 0 ? V(s0.z, s0.w) : V(s1.y, s1.z);
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
            // a graph walk is only needed if some body is asleep or about to be
            if ((e.sleeping | tentative) != 0) {
#pragma unroll 1
                for (int round = 0; round < NB - 1; ++round) {
#pragma unroll 1
                    for (int it = it_begin(); it_ok(it); it = it_next(it)) {
                        int slot = it_slot(it);
                        if (ns_of(slot) == 0) continue;
                        int p = pair_of(slot);
                        const uint32_t info = W.pair_info[p];
                        int b1 = info_b1(info), b2 = info_b2(info);
                        if (b1 < 0 || b2 < 0) continue;
                        uint32_t m1 = 1u << b1, m2 = 1u << b2;
                        if ((reached & m1) || (reached & m2)) reached |= m1 | m2;
                    }
                }
            }
        }
        e.sleeping = (e.sleeping | tentative) & ~reached;
        uint32_t fell_asleep = tentative & e.sleeping;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            if ((fell_asleep >> b) & 1u) {  // RigidBody::sleep
                e.timer[b] = kTimeUntilSleep;
                e.vx[b] = 0.0f; e.vy[b] = 0.0f;
                if (b > 0) e.w[b] = 0.0f;  // the player's rotation is locked: (cos, sin, w) stay (1, 0, 0)
            }
        }
    }

    // rapier VelocityConstraint::generate / VelocityGroundConstraint::generate for one manifold:
    // appends one row per solver contact. A one-body ("ground") constraint is written as a
    // two-body row against the pseudo body kGround (zero velocity, zero inverse mass, zero
    // angular response), which evaluates to the same values as rapier's ground formulas.
    PRL_D int generate_constraint(int slot, int nrows, const f4& s0, const f4& s1) {
        int p = pair_of(slot), ns = ns_of(slot);
        PRL_CHECK(slot >= 0 && slot < S, 2);
        PRL_CHECK(p >= 0 && p < W.n_pairs, 3);
        const uint32_t info = W.pair_info[p];
        int h1 = info_b1(info), h2 = info_b2(info);
        bool two_body = h1 >= 0 && h2 >= 0;
        v2 normal = V(s0.x, s0.y);
        v2 dir1;
        if (two_body) dir1 = -normal;
        else if (h2 < 0) { h2 = h1; h1 = -1; dir1 = normal; }  // the dynamic body becomes body 2
        else dir1 = -normal;
        v2 v1 = V(0.0f, 0.0f), com1 = V(0.0f, 0.0f);
        float w1 = 0.0f, im1 = 0.0f, ii1 = 0.0f;
        v2 v2_ = V(sel(e.vx, h2), sel(e.vy, h2)), com2 = V(sel(e.x, h2), sel(e.y, h2));
        float w2 = sel(e.w, h2), im2 = W.inv_mass[h2], ii2 = W.inv_inertia_sqrt[h2];
        float ccd = W.ccd_thickness[h2];
        if (NB > 1 && two_body) {
            v1 = V(sel(e.vx, h1), sel(e.vy, h1));
            com1 = V(sel(e.x, h1), sel(e.y, h1));
            w1 = sel(e.w, h1);
            im1 = W.inv_mass[h1];
            ii1 = W.inv_inertia_sqrt[h1];
            ccd = W.ccd_thickness[h1] + W.ccd_thickness[h2];
        }
        v2 tan1 = V(-dir1.y, dir1.x);
        uint32_t bodies = (uint32_t)(two_body ? h1 : kGround) | ((uint32_t)h2 << 8);
        bool fast = false;
#pragma unroll 1
        for (int k = 0; k < ns; ++k) {
            v2 pt = k == 0 ? V(s0.z, s0.w) : V(s1.y, s1.z);
            float dist = k == 0 ? s1.x : s1.w;
            v2 dp1 = pt - com1, dp2 = pt - com2;
            v2 vel1 = v1 + wcross(w1, dp1);
            v2 vel2 = v2_ + wcross(w2, dp2);
            f4 r0, r1;
            float gcross2 = ii2 * cross(dp2, -dir1);
            float gcross1 = 0.0f, pm;
            if (NB > 1 && two_body) {
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
            float t_gcross2 = ii2 * cross(dp2, -tan1);
            float t_gcross1 = 0.0f, tr;
            if (NB > 1 && two_body) {
                t_gcross1 = ii1 * cross(dp1, tan1);
                float imsum = im1 + im2;
                tr = dot(tan1, tan1 * imsum) + t_gcross1 * t_gcross1 + t_gcross2 * t_gcross2;
            } else {
                tr = dot(tan1, tan1 * im2) + t_gcross2 * t_gcross2;
            }
            v2 dv = (vel1 - vel2) + V(0.0f, 0.0f);
            r0.x = dir1.x; r0.y = dir1.y; r0.z = gcross2; r0.w = t_gcross2;
            r1.x = rhs; r1.y = pm; r1.z = rhs_wo_bias; r1.w = NB <= 2 ? cfm_reg : bits_f32(bodies);
            f2 r2;
            r2.x = dot(dv, tan1); r2.y = inv0(tr);
            int j = nrows + k;
            int en = row_entry(j);
            PRL_CHECK(en >= 0 && en < kEntries, 1);
            q0(en) = r0; q1(en) = r1; q2(en) = r2;
            acc_n(en) = 0.0f; acc_t(en) = 0.0f;
            if (NB > 2) {
                f2 r5;
                r5.x = gcross1; r5.y = t_gcross1;
                q5(en) = r5;
            } else if (NB == 2 && two_body) {
                if (j == 0) { tb_g1[0] = gcross1; tb_tg1[0] = t_gcross1; }
                else { tb_g1[1] = gcross1; tb_tg1[1] = t_gcross1; }
            }
        }
        if (fast) {  // a fast contact makes the whole constraint rigid (cfm_factor 1)
#pragma unroll 1
            for (int k = 0; k < ns; ++k) q1(row_entry(nrows + k)).w = NB <= 2 ? 1.0f : bits_f32(bodies | (1u << 16));
        }
        return nrows + ns;
    }

    // Gauss-Seidel row updates: rapier VelocityConstraintElement::solve_group (two bodies) and
    // VelocityGroundConstraintElement::solve_group (one body), normal part and friction part.
    // Two-body rows come first in the row list (ntb of them), ground rows after. The body
    // increments dl* stay in registers. BODY >= 0: every row of the range acts on that body (with
    // one pushable block the ground rows of the player's pairs precede the block's, so the sweep
    // is two ranges with a compile-time body); BODY < 0: the body comes from the row's bits.
    // The sweep fetches the NEXT row from shared memory while the current row is computed (a
    // row's impulses are only written by its own update, so reading them one row early is safe).
    struct RowN { f4 a, b; float acc; };            // plane 0, plane 1, normal impulse
    struct RowT { f4 a; f2 c; float acc, accn; uint32_t bits; };  // plane 0, plane 2, tangent impulse, normal impulse
    PRL_D void fetch(int en, RowN& r) const { r.a = q0(en); r.b = q1(en); r.acc = acc_n(en); }
    template <int BODY> PRL_D void fetch(int en, RowT& r) const {
        r.a = q0(en); r.c = q2(en); r.acc = acc_t(en); r.accn = acc_n(en);
        r.bits = BODY < 0 ? f32_bits(q1(en).w) : 0u;
    }
    // softness of a row: rapier's cfm_factor, 1 for a constraint with a fast contact and for every
    // constraint once the bias has been removed (remove_bias)
    PRL_D float row_cfm(float w) const { return NB <= 2 ? w : ((f32_bits(w) >> 16) ? 1.0f : cfm_reg); }
    template <int BODY>
    PRL_D void ground_row_n(int en, const RowN& r, float (&dlx)[NB], float (&dly)[NB], float (&da)[NB]) {
        constexpr int B = BODY >= 0 ? BODY : 0;
        const v2 dir = V(r.a.x, r.a.y);
        const float g2 = r.a.z, rr = r.b.y, acc = r.acc, rhs = r.b.x;
        const int b2 = BODY >= 0 ? BODY : (int)((f32_bits(r.b.w) >> 8) & 0xffu);
        const float im2 = BODY >= 0 ? im_reg[B] : sel(im_reg, b2);
        v2 l2 = BODY >= 0 ? V(dlx[B], dly[B]) : V(sel(dlx, b2), sel(dly, b2));
        float a2 = BODY >= 0 ? da[B] : sel(da, b2);
        float dvel = -dot(dir, l2) + g2 * a2 + rhs;
        float ni = row_cfm(r.b.w) * fmaxf(acc - rr * dvel, 0.0f);
        PRL_CHECK(en >= 0 && en < kEntries, 1);
        acc_n(en) = ni;
        float dl = ni - acc;
        l2 = l2 + (dir * im2) * -dl;
        a2 += g2 * dl;
        if (BODY >= 0) { dlx[B] = l2.x; dly[B] = l2.y; da[B] = a2; }
        else { put(dlx, b2, l2.x); put(dly, b2, l2.y); put(da, b2, a2); }
    }
    template <int BODY>
    PRL_D void ground_row_t(int en, const RowT& r, float (&dlx)[NB], float (&dly)[NB], float (&da)[NB]) {
        constexpr int B = BODY >= 0 ? BODY : 0;
        const v2 dir = V(-r.a.y, r.a.x);  // tangent
        const float g2 = r.a.w, rhs = r.c.x, rr = r.c.y, acc = r.acc;
        const int b2 = BODY >= 0 ? BODY : (int)((r.bits >> 8) & 0xffu);
        const float im2 = BODY >= 0 ? im_reg[B] : sel(im_reg, b2);
        v2 l2 = BODY >= 0 ? V(dlx[B], dly[B]) : V(sel(dlx, b2), sel(dly, b2));
        float a2 = BODY >= 0 ? da[B] : sel(da, b2);
        float dvel = -dot(dir, l2) + g2 * a2 + rhs;
        float limit = kFriction * r.accn;
        float ni = fminf(fmaxf(acc - rr * dvel, -limit), limit);
        PRL_CHECK(en >= 0 && en < kEntries, 1);
        acc_t(en) = ni;
        float dl = ni - acc;
        l2 = l2 + (dir * im2) * -dl;
        a2 += g2 * dl;
        if (BODY >= 0) { dlx[B] = l2.x; dly[B] = l2.y; da[B] = a2; }
        else { put(dlx, b2, l2.x); put(dly, b2, l2.y); put(da, b2, a2); }
    }
    // one sweep over my rows [jbegin, jend): two rows per trip, each fetched while the previous one
    // is computed, in two alternating register sets
    template <bool NORMAL, int BODY>
    PRL_D void ground_rows(int jbegin, int jend, float (&dlx)[NB], float (&dly)[NB], float (&da)[NB]) {
        if (jbegin >= jend) return;
        int en = row_entry(jbegin);
        int left = jend - jbegin;
        if (NORMAL) {
            RowN ra, rb;
            fetch(en, ra);
#pragma unroll 1
            while (true) {
                const int en1 = en + stride;
                fetch(en1, rb);
                ground_row_n<BODY>(en, ra, dlx, dly, da);
                if (--left == 0) break;
                en = en1 + stride;
                fetch(en, ra);
                ground_row_n<BODY>(en1, rb, dlx, dly, da);
                if (--left == 0) break;
            }
        } else {
            RowT ra, rb;
            fetch<BODY>(en, ra);
#pragma unroll 1
            while (true) {
                const int en1 = en + stride;
                fetch<BODY>(en1, rb);
                ground_row_t<BODY>(en, ra, dlx, dly, da);
                if (--left == 0) break;
                en = en1 + stride;
                fetch<BODY>(en, ra);
                ground_row_t<BODY>(en1, rb, dlx, dly, da);
                if (--left == 0) break;
            }
        }
    }
    template <bool NORMAL>
    PRL_D void two_body_row(int j, float (&dlx)[NB], float (&dly)[NB], float (&da)[NB]) {
        const int en = row_entry(j);
        PRL_CHECK(en >= 0 && en < kEntries, 1);
        const f4 r0 = q0(en);
        v2 dir = V(r0.x, r0.y);
        uint32_t bodies = 0;
        float g1, g2, rhs, rr, acc, accn = 0.0f, cfm = 1.0f;
        if (NORMAL) {
            const f4 r1 = q1(en);
            if (NB > 2) bodies = f32_bits(r1.w);
            cfm = row_cfm(r1.w);
            g2 = r0.z; rr = r1.y; rhs = r1.x; acc = acc_n(en);
        } else {
            const f2 r2 = q2(en);
            if (NB > 2) bodies = f32_bits(q1(en).w);
            dir = V(-dir.y, dir.x); g2 = r0.w; rhs = r2.x; rr = r2.y; acc = acc_t(en); accn = acc_n(en);
        }
        int b1 = NB == 2 ? 0 : (int)(bodies & 0xffu), b2 = NB == 2 ? 1 : (int)((bodies >> 8) & 0xffu);
        float im1 = sel(im_reg, b1), im2 = sel(im_reg, b2);
        if (NB > 2) { const f2 r5 = q5(en); g1 = NORMAL ? r5.x : r5.y; }
        else g1 = NORMAL ? (j == 0 ? tb_g1[0] : tb_g1[1]) : (j == 0 ? tb_tg1[0] : tb_tg1[1]);
        v2 l1 = V(sel(dlx, b1), sel(dly, b1)), l2 = V(sel(dlx, b2), sel(dly, b2));
        float a1 = sel(da, b1), a2 = sel(da, b2);
        float dvel = dot(dir, l1) + g1 * a1 - dot(dir, l2) + g2 * a2 + rhs;
        float ni;
        if (NORMAL) {
            ni = cfm * fmaxf(acc - rr * dvel, 0.0f);
            acc_n(en) = ni;
        } else {
            float limit = kFriction * accn;
            ni = fminf(fmaxf(acc - rr * dvel, -limit), limit);
            acc_t(en) = ni;
        }
        float dl = ni - acc;
        l1 = l1 + (dir * im1) * dl;
        a1 += g1 * dl;
        l2 = l2 + (dir * im2) * -dl;
        a2 += g2 * dl;
        put(dlx, b1, l1.x); put(dly, b1, l1.y); put(da, b1, a1);
        put(dlx, b2, l2.x); put(dly, b2, l2.y); put(da, b2, a2);
    }
    // rapier's remove_bias before the stabilisation sweep: every row continues without the
    // penetration bias (rhs := rhs_wo_bias) and without softness (cfm_factor := 1)
    PRL_D void remove_bias(int nrows) {
        int en = row_entry(0);
#pragma unroll 1
        for (int j = 0; j < nrows; ++j, en += stride) {
            PRL_CHECK(en >= 0 && en < kEntries, 1);
            f4 r = q1(en);
            r.x = r.z;
            r.w = NB <= 2 ? 1.0f : bits_f32(f32_bits(r.w) | (1u << 16));
            q1(en) = r;
        }
    }

    // rapier IslandManager::update_active_set_with_contacts + NarrowPhase::select_active_contacts:
    // which manifolds become constraints this step -- two-body manifolds first, then one-body
    // ("ground") ones, each in pair order -- and how many rows they need.
    PRL_D void select_constraints() {
        update_active_set();
        active = ~e.sleeping & ((1u << W.n_bodies) - 1u);
        cons = 0;
        ncons = 0;
        need = 0;
        ntb = 0;
        nplayer = 0;
        if (kDirect) {
            // slot == pair: the selection is bit-mask algebra on the live / solver-contact masks and
            // the world's body -> pairs table; row counts are population counts (ns is 1 or 2)
            uint32_t awake_pairs = 0;
#pragma unroll
            for (int b = 0; b < NB; ++b) if ((active >> b) & 1u) awake_pairs |= W.body_pairs[b];
            const uint32_t sel_ = e.used & e.ns1 & awake_pairs;
            cm_tb = sel_ & W.two_body_pairs;
            cm_gr = sel_ & ~W.two_body_pairs;
            const uint32_t pl = cm_gr & ((1u << W.n_player_pairs) - 1u);
            ntb = popcount(cm_tb) + popcount(cm_tb & e.ns2);
            nplayer = popcount(pl) + popcount(pl & e.ns2);
            need = popcount(sel_) + popcount(sel_ & e.ns2);
        } else {
#pragma unroll 1
        for (int pass = NB == 1 ? 1 : 0; pass < 2; ++pass) {
#pragma unroll 1
            for (int it = it_begin(); it_ok(it); it = it_next(it)) {
                int slot = it_slot(it);
                int ns = ns_of(slot);
                if (ns == 0) continue;
                int p = pair_of(slot);
                const uint32_t info = W.pair_info[p];
                int h1 = info_b1(info), h2 = info_b2(info);
                bool two_body = h1 >= 0 && h2 >= 0;
                if ((pass == 0) != two_body) continue;
                bool awake1 = h1 >= 0 && ((active >> h1) & 1u);
                bool awake2 = h2 >= 0 && ((active >> h2) & 1u);
                if (!awake1 && !awake2) continue;
                cons_push(slot);
                need += ns;
                if (two_body) ntb += ns;
                else if (p < W.n_player_pairs) nplayer += ns;
            }
        }
        }
    }

    // constraint build + VelocitySolver::solve + position update for the active bodies (SURVEY
    // A.3 stages 4-7), with this lane's rows at entries first_entry + j * entry_stride of the
    // warp's planes. Returns the mask of bodies that moved.
    PRL_D uint32_t solve_and_integrate(int first_entry, int entry_stride) {
        ent0 = first_entry;
        stride = entry_stride;
        int nrows = 0;
        if (kDirect) {
            // two-body manifolds first, then one-body ("ground") ones, each in ascending pair order --
            // as ONE sequence per lane, so that a warp runs max(constraints) trips, not
            // max(two-body) + max(ground): nearly every warp has a lane whose player stands on the block
            uint32_t mt = NB > 1 ? cm_tb : 0u, mg = cm_gr;
#pragma unroll 1
            while ((mt | mg) != 0) {
                const uint32_t cur = mt != 0 ? mt : mg;
                const int slot = first_bit_from(cur, 0);
                if (mt != 0) mt &= mt - 1; else mg &= mg - 1;
                const f4 c0 = k.sc[slot][0], c1 = k.sc[slot][1];
                nrows = generate_constraint(slot, nrows, c0, c1);
            }
        } else {
            f4 n0 = {}, n1 = {};
            if (ncons > 0) { const int s0_ = cons_at(0); n0 = k.sc[s0_][0]; n1 = k.sc[s0_][1]; }
#pragma unroll 1
            for (int i = 0; i < ncons; ++i) {
                const f4 c0 = n0, c1 = n1;
                nrows = generate_constraint(cons_at(i), nrows, c0, c1);
                if (i + 1 < ncons) { const int sn = cons_at(i + 1); n0 = k.sc[sn][0]; n1 = k.sc[sn][1]; }
            }
        }
        PRL_TICK(5);
        float dlx[NB], dly[NB], da[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            dlx[b] = b < W.n_bodies ? W.gdv_x[b] : 0.0f;
            dly[b] = b < W.n_bodies ? W.gdv_y[b] : 0.0f;
            da[b] = 0.0f;
        }
        float nx[NB], ny[NB], nc[NB], ns[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) { nx[b] = e.x[b]; ny[b] = e.y[b]; nc[b] = e.c[b]; ns[b] = e.s[b]; }
        // sweep schedule (rapier VelocitySolver::solve with 4 velocity, 8 friction and 1
        // stabilisation iteration): N F N F N F N F F F F F | integrate, drop bias | N F
#pragma unroll 1
        for (int sweep = 0; sweep < 14; ++sweep) {
            if (sweep == 12) {
                // positions integrate the biased velocities
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    if ((active >> b) & 1u) {
                        float dangvel = W.inv_inertia_sqrt[b] * da[b];
                        Iso init; init.t = V(e.x[b], e.y[b]); init.c = e.c[b]; init.s = e.s[b];
                        Iso np = integrate_pose(V(e.vx[b] + dlx[b], e.vy[b] + dly[b]), e.w[b] + dangvel, init, W.dt, b == 0);
                        nx[b] = np.t.x; ny[b] = np.t.y; nc[b] = np.c; ns[b] = np.s;
                    }
                }
                remove_bias(nrows);
            }
            // ground rows: rows [ntb, jsplit) belong to the player's pairs, [jsplit, nrows) to block pairs
            const int jsplit = ntb + nplayer;
            if (sweep == 12 || (sweep < 8 && (sweep & 1) == 0)) {
                if (NB > 1) {
#pragma unroll 1
                    for (int j = 0; j < ntb; ++j) two_body_row<true>(j, dlx, dly, da);
                }
                if (NB == 1) ground_rows<true, 0>(0, nrows, dlx, dly, da);
                else if (NB == 2) {
                    ground_rows<true, 0>(ntb, jsplit, dlx, dly, da);
                    ground_rows<true, 1>(jsplit, nrows, dlx, dly, da);
                }
                else ground_rows<true, -1>(ntb, nrows, dlx, dly, da);
            } else {
                if (NB > 1) {
#pragma unroll 1
                    for (int j = 0; j < ntb; ++j) two_body_row<false>(j, dlx, dly, da);
                }
                if (NB == 1) ground_rows<false, 0>(0, nrows, dlx, dly, da);
                else if (NB == 2) {
                    ground_rows<false, 0>(ntb, jsplit, dlx, dly, da);
                    ground_rows<false, 1>(jsplit, nrows, dlx, dly, da);
                }
                else ground_rows<false, -1>(ntb, nrows, dlx, dly, da);
            }
        }
        PRL_TICK(6);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            if ((active >> b) & 1u) {
                float dangvel = W.inv_inertia_sqrt[b] * da[b];
                e.vx[b] = e.vx[b] + dlx[b];
                e.vy[b] = e.vy[b] + dly[b];
                e.x[b] = nx[b]; e.y[b] = ny[b];
                // the player's rotation is locked (inverse inertia 0): its angular increment is an
                // exact zero and (cos, sin, w) stay the constants (1, 0, 0), so they are never
                // written and cost no registers
                if (b > 0) { e.w[b] += dangvel; e.c[b] = nc[b]; e.s[b] = ns[b]; }
            }
        }
        return active;
    }

    // loosened AABB overlap of pair p and whether one of its colliders moved. A pair has at most
    // one fixed collider, whose box comes with the pair.
    PRL_D bool pair_overlaps(int p, uint32_t moved, bool all, const float (&bmnx)[NB], const float (&bmny)[NB], const float (&bmxx)[NB],
                             const float (&bmxy)[NB], bool& modified) const {
        const uint32_t info = W.pair_info[p];
        const f4 box = *reinterpret_cast<const f4*>(W.pair_box[p]);
        int b1 = info_b1(info), b2 = info_b2(info);
        modified = all || (b1 >= 0 && ((moved >> b1) & 1u)) || (b2 >= 0 && ((moved >> b2) & 1u));
        float amnx, amny, amxx, amxy, qmnx, qmny, qmxx, qmxy;
        if (b1 >= 0) { amnx = sel(bmnx, b1); amny = sel(bmny, b1); amxx = sel(bmxx, b1); amxy = sel(bmxy, b1); }
        else { amnx = box.x; amny = box.y; amxx = box.z; amxy = box.w; }
        if (b2 >= 0) { qmnx = sel(bmnx, b2); qmny = sel(bmny, b2); qmxx = sel(bmxx, b2); qmxy = sel(bmxy, b2); }
        else { qmnx = box.x; qmny = box.y; qmxx = box.z; qmxy = box.w; }
        return amnx <= qmxx && qmnx <= amxx && amny <= qmxy && qmny <= amxy;
    }

    // rapier PhysicsPipeline::detect_collisions for the colliders of the bodies in `moved`
    // (`all` on the first step): loosened-AABB broad phase, then the narrow phase on the
    // overlapping pairs. A brute-force AABB test over the world's candidate pairs is
    // result-equivalent to rapier's sweep-and-prune.
    PRL_D void detect_collisions(uint32_t moved, bool all) {
        // every live manifold of the player is revisited iff the player moved
        gs_n = (all || (moved & 1u)) ? 0 : -1;
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
        if (kDirect) {
            // slot == pair: the broad phase is a branch-free pass over the pairs that leaves bit
            // masks -- overlapping pairs, and (from the world's body -> pairs table) pairs with a
            // moved collider -- so its table look-ups overlap instead of waiting for each other
            uint32_t ov = 0;
#pragma unroll 2  // (4: -4 % on jump.json, the single-body kernels are bound by their instruction footprint)
            for (int p = 0; p < W.n_pairs; ++p) {
                bool modified;
                if (pair_overlaps(p, moved, all, bmnx, bmny, bmxx, bmxy, modified)) ov |= 1u << p;
            }
            uint32_t mod = 0;
#pragma unroll
            for (int b = 0; b < NB; ++b) if (all || ((moved >> b) & 1u)) mod |= W.body_pairs[b];
            // a pair that left the broad phase drops its manifold; NarrowPhase::remove_pair wakes
            // both bodies if the pair had an active contact
#pragma unroll 1
            for (uint32_t stale = mod & ~ov & e.used; stale != 0; stale &= stale - 1) {
                const int p = first_bit_from(stale, 0);
                if (ns_of(p) != 0) {
                    const uint32_t info = W.pair_info[p];
                    int b1 = info_b1(info), b2 = info_b2(info);
                    if (b1 >= 0) wake_strong(b1);
                    if (b2 >= 0) wake_strong(b2);
                }
                e.used &= ~(1u << p);
                e.ns1 &= ~(1u << p);
                e.ns2 &= ~(1u << p);
            }
            const uint32_t cand = mod & ov;
            PRL_TICK(0);
            // Narrow phase on "my i-th candidate", player pairs and block pairs in ONE loop: what runs
            // almost always -- manifold persistence, solver contacts, the store -- is the same code
            // for both, only the out-of-line recompute differs. (Two loops doubled the hot code and
            // ran max(player pairs) + max(block pairs) trips per warp instead of max(pairs).)
            const int npp = W.n_player_pairs;
            PRL_TRACE(tr_cand = 0; tr_full = 0);
#pragma unroll 1
            for (uint32_t m_ = cand; m_ != 0; m_ &= m_ - 1) {
                PRL_TRACE(++tr_cand);
                const int p = first_bit_from(m_, 0);
                PRL_CHECK(p < W.n_pairs && p < S, 3);
                const uint32_t info = W.pair_info[p];
                const f4 he = *reinterpret_cast<const f4*>(W.pair_he[p]);
                Manifold m;
                if (e.used & (1u << p)) load_manifold(p, m);
                else m.npts = 0;
                const Iso pose1 = pair_collider_pose(p, info_b1(info));
                const Iso pos12 = inv_mul(pose1, pair_collider_pose(p, info_b2(info)));
                if (!manifold_try_update(m, pos12)) {
                    PRL_TRACE(++tr_full);
                    if (NB == 1 || p < npp) m = contact_capsule_cuboid_full(pos12, W.radius, W.half_height, V(he.z, he.w), W.prediction);
                    else m = contact_cuboid_cuboid_full(pos12, V(he.x, he.y), V(he.z, he.w), W.prediction);
                }
                if (m.npts > 0) {
                    e.used |= 1u << p;
                    store_manifold(p, p, m, pose1, NB == 1 || p < npp);
                } else {
                    e.used &= ~(1u << p);
                    e.ns1 &= ~(1u << p);
                    e.ns2 &= ~(1u << p);
                }
            }
            PRL_TICK(2);
            return;
        }
        // Broad phase over the world's candidate pairs (ascending), merged with the sorted manifold
        // list: a pair that left the broad phase drops its manifold (NarrowPhase::remove_pair wakes
        // both bodies if it had an active contact); an overlapping pair with a moved collider
        // becomes a narrow-phase candidate.
        uint64_t cand[2] = {0ull, 0ull};  // byte i = pair index of my i-th candidate
        int ncand = 0;
        {
            int k = 0;  // position in the sorted manifold list of the first entry with pair >= p
            auto visit = [&](int p) {
                bool modified;
                bool ov = pair_overlaps(p, moved, all, bmnx, bmny, bmxx, bmxy, modified);
                bool have = k < e.count && pair_at(k) == p;
                if (modified && !ov && have) {
                    if (ns_of(slot_at(k)) != 0) {
                        const uint32_t info = W.pair_info[p];
                        int b1 = info_b1(info), b2 = info_b2(info);
                        if (b1 >= 0) wake_strong(b1);
                        if (b2 >= 0) wake_strong(b2);
                    }
                    remove_at(k);
                    have = false;
                }
                if (have) ++k;
                if (modified && ov) {
                    if (ncand < S) {
                        if (kWide) this->k.cand[ncand] = (uint8_t)p;
                        else if (ncand < 8) cand[0] |= (uint64_t)p << (8 * ncand);
                        else cand[1] |= (uint64_t)p << (8 * (ncand - 8));
                        ++ncand;
                    }
                    else e.flags |= FLAG_OVERFLOW;
                }
            };
            if (NB == 1 && W.grid_n > 0) {
                // single-body world: every pair is (player, fixed collider). Only the pairs the x grid
                // lists for the player's cells can overlap; the pairs of the live manifolds are
                // visited too (one that no longer overlaps drops its manifold), all in pair order.
                uint32_t vm[4] = {0u, 0u, 0u, 0u};
                const int c0 = grid_cell(W, bmnx[0]), c1 = grid_cell(W, bmxx[0]);
#pragma unroll 1
                for (int c = c0; c <= c1; ++c) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) vm[q] |= W.grid_mask[c][q];
                }
#pragma unroll 1
                for (int j = 0; j < e.count; ++j) {
                    const int p = pair_at(j);
#pragma unroll
                    for (int q = 0; q < 4; ++q) if ((p >> 5) == q) vm[q] |= 1u << (p & 31);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
#pragma unroll 1
                    for (uint32_t m_ = vm[q]; m_ != 0; m_ &= m_ - 1) visit(32 * q + first_bit_from(m_, 0));
                }
            } else {
#pragma unroll 1
                for (int p = 0; p < W.n_pairs; ++p) visit(p);
            }
        }
        // Narrow phase on "my i-th candidate": the lanes of a warp work on different pairs in
        // lockstep. Player pairs (capsule-cuboid) sort first, block pairs (cuboid-cuboid) after.
        int k = 0;
        int i = 0;
        auto cand_at = [&](int q) { return kWide ? (int)this->k.cand[q] : (int)(((q < 8 ? cand[0] : cand[1]) >> (8 * (q & 7))) & 0xffu); };
#pragma unroll 1
        for (; i < ncand && cand_at(i) < W.n_player_pairs; ++i) {
            int p = cand_at(i);
            while (k < e.count && pair_at(k) < p) ++k;
            bool have = k < e.count && pair_at(k) == p;
            int slot = have ? slot_at(k) : -1;
            const f4 he = *reinterpret_cast<const f4*>(W.pair_he[p]);
            Manifold m;
            if (have) load_manifold(slot, m);
            else m.npts = 0;
            Iso pose1 = body_pose(0);
            Iso pos12 = inv_mul(pose1, pair_collider_pose(p, info_b2(W.pair_info[p])));
            contact_capsule_cuboid(m, pos12, W.radius, W.half_height, V(he.z, he.w), W.prediction);
            k = commit<true>(p, k, have, slot, m, pose1);
        }
        if (NB > 1) {
#pragma unroll 1
            for (; i < ncand; ++i) {
                int p = cand_at(i);
                while (k < e.count && pair_at(k) < p) ++k;
                bool have = k < e.count && pair_at(k) == p;
                int slot = have ? slot_at(k) : -1;
                const uint32_t info = W.pair_info[p];
                const f4 he = *reinterpret_cast<const f4*>(W.pair_he[p]);
                Manifold m;
                if (have) load_manifold(slot, m);
                else m.npts = 0;
                Iso pose1 = pair_collider_pose(p, info_b1(info));
                Iso pos12 = inv_mul(pose1, pair_collider_pose(p, info_b2(info)));
                contact_cuboid_cuboid(m, pos12, V(he.x, he.y), V(he.z, he.w), W.prediction);
                k = commit(p, k, have, slot, m, pose1);
            }
        }
    }
    // write a narrow-phase result back at sorted position k; returns the position after it. Only
    // manifolds with points are kept: an empty one carries no state the next step can observe.
    template <bool SCAN = false>
    PRL_D int commit(int p, int k, bool have, int slot, const Manifold& m, const Iso& pose1) {
        if (m.npts > 0) {
            if (!have) slot = insert_at(k);
            if (slot >= 0) { store_manifold<SCAN>(slot, p, m, pose1); return k + 1; }
            return k;
        }
        if (have) remove_at(k);
        return k;
    }

    // common.rs:175-195 distance_to_goals, editor pixels; +inf for a goal-less world
    PRL_D float distance_to_goals() const {
        float best = f32_inf();
#pragma unroll 1
        for (int i = 0; i < W.n_goals; ++i) {
            v2 d = V(e.x[0] - W.gx[i], e.y[0] - W.gy[i]);
            float dx = fmaxf(fabsf(d.x * W.gxax[i] + d.y * W.gxay[i]) - W.ghw[i], 0.0f);
            float dy = fmaxf(fabsf(d.x * W.gyax[i] + d.y * W.gyay[i]) - W.ghh[i], 0.0f);
            float dist = sqrtf(dx * dx + dy * dy) / W.scale;
            best = i == 0 ? dist : fminf(best, dist);
        }
        return best;
    }

    // ---- Observation for policies evaluated on the device: what Agent::get_move(&mut self,
    // &Environment) (algorithm.rs:9) can read between two steps -- every body through
    // rigid_body_set() (common.rs:201-211), the ground-contact scan Environment::step itself makes
    // from narrow_phase / collider_set() (common.rs:223-244), distance_to_goals() (common.rs:175-195):
    //   obs[9 b + 0..8]   body b: x, y, cos, sin, vx, vy, angvel, sleep timer, sleeping (0/1)
    //   obs[9 nb + 0]     on_ground: 1 if the player has a floor contact (common.rs:244), else 0
    //   obs[9 nb + 1]     number of floor contacts
    //   obs[9 nb + 2]     1 if the player has an active contact with a pushable block, else 0
    //   obs[9 nb + 3]     distance_to_goals() in editor pixels (+inf without goals)
    // Call after detect_pending(). Before the first step the narrow phase is empty, so
    // the contact entries are 0 then.
    static constexpr int kObsExtra = 4;
    PRL_D int floor_contact_count() const {
        if (gs_n >= 0) return gs_n;
        const float ly = e.y[0] - W.lower_off;
        int nfc = 0;
#pragma unroll 1
        for (int it = it_begin(); it_ok(it); it = it_next(it)) {
            const int slot = it_slot(it);
            if (pair_of(slot) >= W.n_player_pairs) break;
            const int ns = ns_of(slot);
            const f4 s0 = k.sc[slot][0], s1 = k.sc[slot][1];
            if (ns > 0 && (s0.w - ly) / W.radius < -0.707f) ++nfc;
            if (ns > 1 && (s1.z - ly) / W.radius < -0.707f) ++nfc;
        }
        return nfc;
    }
    PRL_D bool touches_pushable_block() const {
        bool touching = false;
        if (NB > 1) {
#pragma unroll 1
            for (int it = it_begin(); it_ok(it) && !touching; it = it_next(it)) {
                const int slot = it_slot(it);
                const int p = pair_of(slot);
                if (p >= W.n_player_pairs) break;
                touching = ns_of(slot) > 0 && info_b2(W.pair_info[p]) >= 0;
            }
        }
        return touching;
    }
    PRL_D void observe(float* obs) const {
        const int nb = W.n_bodies;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            if (b < nb) {
                float* o = obs + 9 * b;
                o[0] = e.x[b]; o[1] = e.y[b]; o[2] = e.c[b]; o[3] = e.s[b]; o[4] = e.vx[b]; o[5] = e.vy[b];
                o[6] = e.w[b]; o[7] = e.timer[b]; o[8] = (float)((e.sleeping >> b) & 1u);
            }
        }
        const bool before_first_step = first_step || pend_all || (e.flags & FLAG_FRESH) != 0;
        const int nfc = before_first_step ? 0 : floor_contact_count();
        float* x = obs + 9 * nb;
        x[0] = nfc > 0 ? 1.0f : 0.0f;
        x[1] = (float)nfc;
        x[2] = (!before_first_step && touches_pushable_block()) ? 1.0f : 0.0f;
        x[3] = distance_to_goals();
    }

    // min over the goals of the SQUARED physics-space distance (the argument of distance_to_goals'
    // sqrt): the fused rollout keeps the running minimum of this and takes one sqrt and one division at
    // its end -- sqrt and the division by the scale are monotone, so min and win test commute with them
    PRL_D float goal_distance_sq() const {
        float best = f32_inf();
#pragma unroll 1
        for (int i = 0; i < W.n_goals; ++i) {
            v2 d = V(e.x[0] - W.gx[i], e.y[0] - W.gy[i]);
            float dx = fmaxf(fabsf(d.x * W.gxax[i] + d.y * W.gxay[i]) - W.ghw[i], 0.0f);
            float dy = fmaxf(fabsf(d.x * W.gyax[i] + d.y * W.gyay[i]) - W.ghh[i], 0.0f);
            best = fminf(best, dx * dx + dy * dy);  // (finite state: no NaN to propagate)
        }
        return best;
    }

    // Environment::step (common.rs:215-340) in three parts, so that the collision detection has
    // one call site and the warp can agree on the shape of its row storage between the halves (every
    // lane of the warp takes part in that vote, including lanes whose rollout has ended):
    //
    //   detect_pending();                 // broad + narrow phase owed by the previous position update
    //   need = begin_step(move);          // ground scan + move impulses + constraint selection
    //   (kernel) the warp agrees on how many columns of row storage a lane gets this step
    //   d = finish_step(entry, stride);   // solve, integrate, win latch
    //
    // rapier runs the detection at the END of PhysicsPipeline::step; running it at the start of
    // the next step (and once more after the last one) is the same sequence of operations.
    PRL_D void detect_pending() {
        if (pend_all || pend_moved != 0) {
            // first step only: every collider is new, so rapier detects collisions on the initial
            // poses before anything moves (SURVEY A.3 stage 1) -- but Environment::step's ground
            // scan ran before that and saw no contacts (common.rs:224)
            first_step = pend_all;
            detect_collisions(pend_moved, pend_all);
            pend_all = false;
            pend_moved = 0;
            e.flags &= ~FLAG_FRESH;
        }
    }
    PRL_D int begin_step(uint32_t mv) {
        if (!first_step) player_move(mv);
        first_step = false;
        PRL_TICK(3);
        select_constraints();
        PRL_TICK(4);
        return need;
    }
    PRL_D float finish_step(int first_entry, int entry_stride) {
        pend_moved = solve_and_integrate(first_entry, entry_stride);
        // query_pipeline.update (common.rs:330-331) is never queried: omitted
        float d = distance_to_goals();
        if (!(e.flags & FLAG_WON) && W.n_goals > 0 && d < 1e-7f) e.flags |= FLAG_WON;
        PRL_TICK(7);
        return d;
    }
    // finish_step for rollouts that only need the running minimum and the win latch: returns the
    // squared distance; won() latches iff sqrt(q) / scale < 1e-7, i.e. iff q <= W.win_q
    PRL_D float finish_step_sq(int first_entry, int entry_stride) {
        pend_moved = solve_and_integrate(first_entry, entry_stride);
        const float q = goal_distance_sq();
        if (!(e.flags & FLAG_WON) && W.n_goals > 0 && q <= W.win_q) e.flags |= FLAG_WON;
        PRL_TICK(7);
        return q;
    }
    PRL_D float step_q(uint32_t mv) {  // host build
        detect_pending();
        begin_step(mv);
        return finish_step_sq(0, 1);
    }
    // the whole step for one lane on its own (host build); call detect_pending() once more after
    // the last step
    PRL_D float step(uint32_t mv) {
        detect_pending();
        begin_step(mv);
        return finish_step(0, 1);  // a single lane owns the planes
    }
};

// Policy of the observation-driven rollout (prl_kernels.cuh policy_rollout_kernel; layout in include/prl.h)
constexpr int kMaxObs = 9 * MAXB + 4;
constexpr int kMaxHidden = 32;
PRL_D int policy_obs_dim(int n_bodies) { return 9 * n_bodies + 4; }
PRL_D int policy_num_params(int n_bodies, int hidden) { return hidden * (policy_obs_dim(n_bodies) + 1) + 3 * (hidden + 1); }

PRL_D uint32_t policy_move(const float* theta, int64_t n, int64_t a, int hidden, int D, const float* obs) {
    float h[kMaxHidden];
    const float* w = theta + a;
#pragma unroll 1
    for (int j = 0; j < hidden; ++j) {
        const float* row = w + (int64_t)j * (D + 1) * n;
        float acc = row[(int64_t)D * n];
#pragma unroll 4
        for (int i = 0; i < D; ++i) acc = acc + row[(int64_t)i * n] * obs[i];
        h[j] = fmaxf(acc, 0.0f);
    }
    uint32_t mv = 0;
    const float* w2 = w + (int64_t)hidden * (D + 1) * n;
#pragma unroll 1
    for (int k = 0; k < 3; ++k) {
        const float* row = w2 + (int64_t)k * (hidden + 1) * n;
        float acc = row[(int64_t)hidden * n];
#pragma unroll 4
        for (int j = 0; j < hidden; ++j) acc = acc + row[(int64_t)j * n] * h[j];
        if (acc > 0.0f) mv |= 1u << k;
    }
    return mv;
}


}  // namespace prl
