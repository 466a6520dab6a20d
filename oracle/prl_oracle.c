/*
 * prl_oracle.c -- CPU ORACLE (test infrastructure, NOT a product path). See prl_oracle.h.
 *
 * PARITY UNPINNED: no reference test/golden vector exists for this path and the real crate
 * cannot be built here; this file restates the published algorithms of
 *   rapier2d 0.17.2 / parry2d 0.13.4 / nalgebra 0.32.2 / libm 0.2.7 / glam 0.23.0
 * (named per function below as "rapier:", "parry:", "nalgebra:", "libm:", "glam:") around a
 * literal restatement of /root/reference/src/common.rs ("ref:" citations).
 *
 * Known, documented deviations from the crates (DESIGN.md "Oracle fidelity"):
 *  D1  capsule-vs-cuboid closest points: parry runs GJK (+EPA when the inner segment
 *      overlaps the box); the oracle uses an exact closest-feature enumeration
 *      (2 endpoints + 4 corners) and a min-face-penetration fallback. Same result up to
 *      GJK's relative tolerance sqrt(10*eps) on the direction; not bit-level.
 *  D2  contact-graph edge order (Gauss-Seidel order, floor-contact scan order) is canonical:
 *      ascending (collider1, collider2); rapier's order is the broad phase's event order.
 *  D3  broad phase is a brute-force loosened-AABB overlap test (observably equivalent to SAP
 *      up to pair order).
 *  D4  RigidBody::apply_impulse ignores a zero impulse (does not wake) -- SURVEY A.2 [L].
 *
 * Each deviation can be flipped at run time (orc_set_default_variant / orc_env_set_variant,
 * ORC_VARIANT_*) to MEASURE how much the results depend on it (tests/test_oracle_sensitivity.py,
 * DESIGN.md "Oracle"): GJK closest points (D1), descending / contact-graph edge order (D2),
 * waking on a zero impulse (D4). The default (0) is the canonical oracle the CUDA path matches.
 *
 * Everything is f32, evaluated in the crates' expression order, with no FMA contraction
 * (-ffp-contract=off), because rustc never contracts.
 */
#include "prl_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ constants */
/* ref: common.rs:7-9 */
static const float PLAYER_DEPTH = 20.0f;
static const float PLAYER_RADIUS = 20.0f;
#define BEVY_TO_PHYSICS_SCALE (0.25f / (2.0f * 20.0f))

/* rapier: IntegrationParameters::default() */
static const float P_DT = 1.0f / 60.0f;
static const float P_ERP = 0.8f;
static const float P_DAMPING_RATIO = 0.25f;
static const float P_ALLOWED_LINEAR_ERROR = 0.001f;
static const float P_PREDICTION = 0.002f;
#define P_MAX_PENETRATION_CORRECTION 3.40282347e+38f
#define P_VELOCITY_ITERS 4
#define P_FRICTION_ITERS 8
#define P_STABILIZATION_ITERS 1
/* rapier: RigidBodyActivation::default_*() */
static const float ACT_LINEAR_THRESHOLD = 0.4f;
static const float ACT_ANGULAR_THRESHOLD = 0.5f;
static const float ACT_TIME_UNTIL_SLEEP = 2.0f;
/* parry: utils::COS_1_DEGREES, ContactManifold::try_update_contacts thresholds */
static const float COS_1_DEGREES = 0.99984769515f;
static const float PERSIST_DIST_SQ = 1.0e-6f;
/* rapier collider defaults: friction 0.5 (Average rule -> 0.5), restitution 0 */
static const float FRICTION = 0.5f;
#define F32_EPS 1.1920929e-7f

typedef struct { float x, y; } v2;
/* nalgebra Isometry2: translation + UnitComplex (re = c, im = s) */
typedef struct { v2 t; float c, s; } iso;

enum { SHAPE_CAPSULE = 0, SHAPE_CUBOID = 1 };

typedef struct {
    iso pos;            /* rb.pos.position */
    v2 linvel;
    float angvel;
    float inv_mass;          /* mprops.effective_inv_mass (same on both axes) */
    float inv_inertia_sqrt;  /* mprops.effective_world_inv_inertia_sqrt (0: rotations locked) */
    float ccd_thickness;
    int sleeping;
    float timer;             /* activation.time_since_can_sleep */
    int collider;
    /* per-step scratch */
    v2 force;
    iso next_pos;            /* rb.pos.next_position */
    v2 dlin;                 /* mj_lambda.linear */
    float dang;              /* mj_lambda.angular (scaled by inv_inertia_sqrt) */
    int active;
} body_t;

typedef struct {
    int shape;
    int body;            /* -1: parent-less (fixed) collider */
    iso pos;
    v2 he;               /* cuboid half-extents; capsule: (radius, half_height) */
    v2 aabb_min, aabb_max; /* loosened by prediction/2 */
    int modified;
} collider_t;

typedef struct { v2 p1, p2; float dist; } mpoint_t;
typedef struct { v2 point; float dist; } scontact_t;

typedef struct {
    int c1, c2;
    int exists;          /* broad-phase pair present (contact-graph edge exists) */
    int npts;
    v2 n1, n2;           /* local_n1, local_n2 */
    mpoint_t pt[2];
    /* manifold.data (rapier: narrow_phase.rs compute_contacts) */
    v2 normal;
    int nsolver;
    scontact_t sc[2];
    int has_active;      /* pair.has_any_active_contact */
} pair_t;

typedef struct { float x, y, w, h, rot; v2 x_axis, y_axis; } goal_t;

struct orc_env {
    int nbodies, ncolliders, npairs, ngoals;
    body_t* bodies;
    collider_t* colliders;
    pair_t* pairs;
    goal_t* goals;
    int won;
    int fresh;           /* no pipeline step taken yet: narrow phase empty */
    /* from_world snapshot for reset */
    body_t* bodies0;
    collider_t* colliders0;
    /* derived integration parameters */
    float inv_dt, erp_inv_dt, cfm_factor;
    /* sensitivity studies only (orc_set_default_variant): which documented deviation to flip */
    int variant;
    int* edges;          /* pair indices in contact-graph edge order: appended when a pair is created,
                            swap-removed when it goes (rapier: InteractionGraph over petgraph) */
    int nedges;
};
static int g_default_variant = 0;
void orc_set_default_variant(int32_t flags) { g_default_variant = flags; }

/* ------------------------------------------------------------------ small maths */
static inline v2 V2(float x, float y) { v2 r = {x, y}; return r; }
static inline v2 vadd(v2 a, v2 b) { return V2(a.x + b.x, a.y + b.y); }
static inline v2 vsub(v2 a, v2 b) { return V2(a.x - b.x, a.y - b.y); }
static inline v2 vscale(v2 a, float k) { return V2(a.x * k, a.y * k); }
static inline v2 vneg(v2 a) { return V2(-a.x, -a.y); }
/* nalgebra: Vector2 dot = a.x*b.x + a.y*b.y */
static inline float vdot(v2 a, v2 b) { return a.x * b.x + a.y * b.y; }
/* rapier utils: Vector2.gcross(Vector2) */
static inline float vcross(v2 a, v2 b) { return a.x * b.y - a.y * b.x; }
/* rapier utils: Real.gcross(Vector2) = (-w*v.y, w*v.x) */
static inline v2 wcross(float w, v2 v) { return V2(-w * v.y, w * v.x); }
/* parry/rapier utils::inv */
static inline float inv0(float x) { return x == 0.0f ? 0.0f : 1.0f / x; }

/* nalgebra: UnitComplex * Vector2 = (re*x - im*y, im*x + re*y) */
static inline v2 rot_mul(float c, float s, v2 v) { return V2(c * v.x - s * v.y, s * v.x + c * v.y); }
/* nalgebra: UnitComplex::inverse_transform_vector = conj * v */
static inline v2 rot_inv_mul(float c, float s, v2 v) {
    float ns = -s;
    return V2(c * v.x - ns * v.y, ns * v.x + c * v.y);
}
/* nalgebra: Isometry * Point = rotation * p + translation */
static inline v2 iso_mul_pt(const iso* m, v2 p) { return vadd(rot_mul(m->c, m->s, p), m->t); }
/* nalgebra: Isometry::inverse_transform_point = rot^-1 * (p - t) */
static inline v2 iso_inv_mul_pt(const iso* m, v2 p) { return rot_inv_mul(m->c, m->s, vsub(p, m->t)); }
/* nalgebra: Isometry::inv_mul(self, rhs) */
static inline iso iso_inv_mul(const iso* a, const iso* b) {
    iso r;
    float ns = -a->s;
    r.t = rot_inv_mul(a->c, a->s, vsub(b->t, a->t));
    r.c = a->c * b->c - ns * b->s;
    r.s = a->c * b->s + ns * b->c;
    return r;
}
/* nalgebra: Isometry::inverse */
static inline iso iso_inverse(const iso* a) {
    iso r;
    r.c = a->c;
    r.s = -a->s;
    r.t = rot_mul(r.c, r.s, vneg(a->t));
    return r;
}

/* ------------------------------------------------------------------ libm 0.2.7 sinf/cosf
 * libm: src/math/{sinf,cosf,k_sinf,k_cosf,rem_pio2f}.rs (musl-derived). Used by nalgebra's
 * UnitComplex::new under rapier's enhanced-determinism (simba libm_force). */
static float k_sinf(double x) {
    const double S1 = -0.166666666416265235595, S2 = 0.0083333293858894631756,
                 S3 = -0.000198393348360966317347, S4 = 0.0000027183114939898219064;
    double z = x * x;
    double w = z * z;
    double r = S3 + z * S4;
    double s = z * x;
    return (float)((x + s * (S1 + z * S2)) + s * w * r);
}
static float k_cosf(double x) {
    const double C0 = -0.499999997251031003120, C1 = 0.0416666233237390631894,
                 C2 = -0.00138867637746099294692, C3 = 0.0000243904487962774090654;
    double z = x * x;
    double w = z * z;
    double r = C2 + z * C3;
    return (float)(((1.0 + z * C0) + w * C1) + (w * z) * r);
}
static int rem_pio2f_medium(float x, double* y) {
    const double TOINT = 1.5 / 2.220446049250313e-16, INV_PIO2 = 6.36619772367581382433e-01,
                 PIO2_1 = 1.57079631090164184570e+00, PIO2_1T = 1.58932547735281966916e-08;
    double x64 = (double)x;
    double f_n = x64 * INV_PIO2 + TOINT - TOINT;
    *y = x64 - f_n * PIO2_1 - f_n * PIO2_1T;
    return (int)f_n;
}
static const double PIO2_D = 1.57079632679489661923;
static float libm_sinf(float x) {
    double x64 = (double)x;
    uint32_t ix;
    memcpy(&ix, &x, 4);
    int sign = (int)(ix >> 31);
    ix &= 0x7fffffffu;
    if (ix <= 0x3f490fdau) {
        if (ix < 0x39800000u) return x;
        return k_sinf(x64);
    }
    if (ix <= 0x407b53d1u) {
        if (ix <= 0x4016cbe3u) return sign ? -k_cosf(x64 + 1.0 * PIO2_D) : k_cosf(x64 - 1.0 * PIO2_D);
        return k_sinf(sign ? -(x64 + 2.0 * PIO2_D) : -(x64 - 2.0 * PIO2_D));
    }
    if (ix <= 0x40e231d5u) {
        if (ix <= 0x40afeddfu) return sign ? k_cosf(x64 + 3.0 * PIO2_D) : -k_cosf(x64 - 3.0 * PIO2_D);
        return k_sinf(sign ? x64 + 4.0 * PIO2_D : x64 - 4.0 * PIO2_D);
    }
    if (ix >= 0x7f800000u) return x - x;
    double y;
    int n = rem_pio2f_medium(x, &y); /* |x| < 2^28*pi/2; larger arguments are out of scope */
    switch (n & 3) {
        case 0: return k_sinf(y);
        case 1: return k_cosf(y);
        case 2: return k_sinf(-y);
        default: return -k_cosf(y);
    }
}
static float libm_cosf(float x) {
    double x64 = (double)x;
    uint32_t ix;
    memcpy(&ix, &x, 4);
    int sign = (int)(ix >> 31);
    ix &= 0x7fffffffu;
    if (ix <= 0x3f490fdau) {
        if (ix < 0x39800000u) return 1.0f;
        return k_cosf(x64);
    }
    if (ix <= 0x407b53d1u) {
        if (ix > 0x4016cbe3u) return -k_cosf(sign ? x64 + 2.0 * PIO2_D : x64 - 2.0 * PIO2_D);
        return sign ? k_sinf(x64 + 1.0 * PIO2_D) : k_sinf(1.0 * PIO2_D - x64);
    }
    if (ix <= 0x40e231d5u) {
        if (ix > 0x40afeddfu) return k_cosf(sign ? x64 + 4.0 * PIO2_D : x64 - 4.0 * PIO2_D);
        return sign ? k_sinf(-x64 - 3.0 * PIO2_D) : k_sinf(x64 - 3.0 * PIO2_D);
    }
    if (ix >= 0x7f800000u) return x - x;
    double y;
    int n = rem_pio2f_medium(x, &y);
    switch (n & 3) {
        case 0: return k_cosf(y);
        case 1: return k_sinf(-y);
        case 2: return -k_cosf(y);
        default: return k_sinf(y);
    }
}

/* ------------------------------------------------------------------ AABBs (parry) */
/* parry: Cuboid::aabb(pos) -> center +- |R|*he ; Capsule::aabb -> transformed segment +- r;
 * rapier: broad_phase handle_modified_collider loosens by prediction/2. */
static void collider_update_aabb(collider_t* co) {
    v2 mins, maxs;
    if (co->shape == SHAPE_CUBOID) {
        float ac = fabsf(co->pos.c), as = fabsf(co->pos.s);
        /* |R| * he with R = [[c,-s],[s,c]] */
        v2 ws = V2(ac * co->he.x + as * co->he.y, as * co->he.x + ac * co->he.y);
        mins = vsub(co->pos.t, ws);
        maxs = vadd(co->pos.t, ws);
    } else {
        float r = co->he.x, hh = co->he.y;
        v2 a = iso_mul_pt(&co->pos, V2(0.0f, -hh));
        v2 b = iso_mul_pt(&co->pos, V2(0.0f, hh));
        mins = V2(fminf(a.x, b.x) - r, fminf(a.y, b.y) - r);
        maxs = V2(fmaxf(a.x, b.x) + r, fmaxf(a.y, b.y) + r);
    }
    float l = P_PREDICTION / 2.0f;
    co->aabb_min = V2(mins.x - l, mins.y - l);
    co->aabb_max = V2(maxs.x + l, maxs.y + l);
}
static int aabb_overlap(const collider_t* a, const collider_t* b) {
    return a->aabb_min.x <= b->aabb_max.x && b->aabb_min.x <= a->aabb_max.x &&
           a->aabb_min.y <= b->aabb_max.y && b->aabb_min.y <= a->aabb_max.y;
}

/* ------------------------------------------------------------------ narrow phase (parry) */

/* parry: ContactManifold::try_update_contacts_eps(pos12, COS_1_DEGREES, 1e-6) */
static int manifold_try_update(pair_t* m, const iso* pos12) {
    if (m->npts == 0) return 0;
    v2 local_n2 = rot_mul(pos12->c, pos12->s, m->n2);
    if (-vdot(m->n1, local_n2) < COS_1_DEGREES) return 0;
    for (int k = 0; k < m->npts; ++k) {
        mpoint_t* pt = &m->pt[k];
        v2 local_p2 = iso_mul_pt(pos12, pt->p2);
        v2 dpt = vsub(local_p2, pt->p1);
        float dist = vdot(dpt, m->n1);
        if (dist * pt->dist < 0.0f) return 0;
        v2 new_p1 = vsub(local_p2, vscale(m->n1, dist));
        v2 d = vsub(pt->p1, new_p1);
        if (vdot(d, d) > PERSIST_DIST_SQ) return 0;
        pt->dist = dist;
        pt->p1 = new_p1;
    }
    return 1;
}

/* parry: Cuboid::support_face (dim2): the face whose normal is most aligned with dir.
 * i = iamin(dir) (first index on ties) is the EDGE direction. */
static void cuboid_support_face(v2 he, v2 dir, v2* a, v2* b) {
    if (fabsf(dir.y) < fabsf(dir.x)) { /* i = 1 (edge along y), j = 0 */
        float xj = copysignf(he.x, dir.x);
        *a = V2(xj, he.y);
        *b = V2(xj, -he.y);
    } else { /* i = 0 (edge along x), j = 1 */
        float yj = copysignf(he.y, dir.y);
        *a = V2(he.x, yj);
        *b = V2(-he.x, yj);
    }
}

/* parry: query::details::clip_segment_segment_with_normal (dim2). Returns 0 when the
 * tangent ranges are disjoint. out = {a.p1, a.p2, b.p1, b.p2}. */
static int clip_segment_segment_with_normal(v2 s1a, v2 s1b, v2 s2a, v2 s2b, v2 normal, v2 out[4]) {
    v2 tangent = V2(-normal.y, normal.x);
    float r1a = vdot(s1a, tangent), r1b = vdot(s1b, tangent);
    float r2a = vdot(s2a, tangent), r2b = vdot(s2b, tangent);
    if (r1b < r1a) { float t = r1a; r1a = r1b; r1b = t; v2 p = s1a; s1a = s1b; s1b = p; }
    if (r2b < r2a) { float t = r2a; r2a = r2b; r2b = t; v2 p = s2a; s2a = s2b; s2b = p; }
    if (r2a > r1b || r1a > r2b) return 0;
    if (r2a > r1a) {
        float bc = (r2a - r1a) * inv0(r1b - r1a);
        out[0] = vadd(s1a, vscale(vsub(s1b, s1a), bc));
        out[1] = s2a;
    } else {
        float bc = (r1a - r2a) * inv0(r2b - r2a);
        out[0] = s1a;
        out[1] = vadd(s2a, vscale(vsub(s2b, s2a), bc));
    }
    if (r2b < r1b) {
        float bc = (r2b - r1a) * inv0(r1b - r1a);
        out[2] = vadd(s1a, vscale(vsub(s1b, s1a), bc));
        out[3] = s2b;
    } else {
        float bc = (r1b - r2a) * inv0(r2b - r2a);
        out[2] = s1b;
        out[3] = vadd(s2a, vscale(vsub(s2b, s2a), bc));
    }
    return 1;
}

/* parry: PolygonalFeature::face_face_contacts (dim2): always two points when clipping
 * succeeds (`if true { // dist < prediction`). Points are (p on 1 in frame 1, p on 2 in
 * frame 2, dist along normal1). */
static void face_face_contacts(pair_t* m, const iso* pos12, v2 f1a, v2 f1b, v2 normal1, v2 f2a, v2 f2b) {
    v2 c[4];
    if (!clip_segment_segment_with_normal(f1a, f1b, f2a, f2b, normal1, c)) return;
    m->pt[0].p1 = c[0];
    m->pt[0].p2 = iso_inv_mul_pt(pos12, c[1]);
    m->pt[0].dist = vdot(vsub(c[1], c[0]), normal1);
    m->pt[1].p1 = c[2];
    m->pt[1].p2 = iso_inv_mul_pt(pos12, c[3]);
    m->pt[1].dist = vdot(vsub(c[3], c[2]), normal1);
    m->npts = 2;
}

/* parry: sat::cuboid_cuboid_find_local_separating_normal_oneway */
static float cuboid_sat_oneway(v2 he1, v2 he2, const iso* pos12, v2* best_dir) {
    float best = -3.40282347e+38f;
    *best_dir = V2(0.0f, 0.0f);
    for (int i = 0; i < 2; ++i) {
        float ti = i == 0 ? pos12->t.x : pos12->t.y;
        float sign = copysignf(1.0f, ti);
        v2 axis1 = i == 0 ? V2(sign, 0.0f) : V2(0.0f, sign);
        v2 axis2 = rot_inv_mul(pos12->c, pos12->s, vneg(axis1));
        /* parry: Cuboid::local_support_point = copysign(he, dir) per component */
        v2 local_pt2 = V2(copysignf(he2.x, axis2.x), copysignf(he2.y, axis2.y));
        v2 pt2 = iso_mul_pt(pos12, local_pt2);
        float sep = (i == 0 ? pt2.x : pt2.y) * sign - (i == 0 ? he1.x : he1.y);
        if (sep > best) { best = sep; *best_dir = axis1; }
    }
    return best;
}

/* parry: contact_manifold_cuboid_cuboid */
static void contact_cuboid_cuboid(pair_t* m, const iso* pos12, v2 he1, v2 he2) {
    if (manifold_try_update(m, pos12)) return;
    iso pos21 = iso_inverse(pos12);
    v2 d1, d2;
    float sep1 = cuboid_sat_oneway(he1, he2, pos12, &d1);
    if (sep1 > P_PREDICTION) { m->npts = 0; return; }
    float sep2 = cuboid_sat_oneway(he2, he1, &pos21, &d2);
    if (sep2 > P_PREDICTION) { m->npts = 0; return; }
    /* dim2: sep3 = (-MAX, x) never wins */
    v2 best = d1;
    if (sep2 > sep1) best = rot_mul(pos12->c, pos12->s, vneg(d2));
    m->npts = 0;
    v2 local_n2 = rot_mul(pos21.c, pos21.s, vneg(best));
    v2 f1a, f1b, f2a, f2b;
    cuboid_support_face(he1, best, &f1a, &f1b);
    cuboid_support_face(he2, local_n2, &f2a, &f2b);
    f2a = iso_mul_pt(pos12, f2a);
    f2b = iso_mul_pt(pos12, f2b);
    face_face_contacts(m, pos12, f1a, f1b, best, f2a, f2b);
    m->n1 = best;
    m->n2 = local_n2;
}

/* Deviation D1: exact closest points between segment [a,b] and the box [-he,he], all in the
 * BOX frame (replaces parry's gjk::closest_points on the capsule's inner segment).
 * Returns 0 if the segment touches/overlaps the box, else 1 with u (parameter on the
 * segment) and q (point on the box). Candidates in fixed order, strict '<'. */
static int segment_box_closest(v2 a, v2 b, v2 he, float* u_out, v2* q_out) {
    v2 d = vsub(b, a);
    /* slab test for overlap */
    {
        float t0 = 0.0f, t1 = 1.0f;
        int hit = 1;
        const float av[2] = {a.x, a.y}, dv[2] = {d.x, d.y}, hv[2] = {he.x, he.y};
        for (int i = 0; i < 2 && hit; ++i) {
            if (dv[i] == 0.0f) {
                if (av[i] < -hv[i] || av[i] > hv[i]) hit = 0;
            } else {
                float inv = 1.0f / dv[i];
                float ta = (-hv[i] - av[i]) * inv, tb = (hv[i] - av[i]) * inv;
                if (ta > tb) { float t = ta; ta = tb; tb = t; }
                if (ta > t0) t0 = ta;
                if (tb < t1) t1 = tb;
                if (t0 > t1) hit = 0;
            }
        }
        if (hit) return 0;
    }
    float best = 3.40282347e+38f;
    float bu = 0.0f;
    v2 bq = V2(0.0f, 0.0f);
    /* segment endpoints against the box */
    for (int e = 0; e < 2; ++e) {
        v2 p = e == 0 ? a : b;
        v2 q = V2(fminf(fmaxf(p.x, -he.x), he.x), fminf(fmaxf(p.y, -he.y), he.y));
        v2 r = vsub(q, p);
        float dd = vdot(r, r);
        if (dd < best) { best = dd; bu = e == 0 ? 0.0f : 1.0f; bq = q; }
    }
    /* box corners against the segment */
    float dd2 = vdot(d, d);
    for (int k = 0; k < 4; ++k) {
        v2 cn = V2((k == 1 || k == 2) ? he.x : -he.x, (k >= 2) ? he.y : -he.y);
        float u = vdot(vsub(cn, a), d) * inv0(dd2);
        u = fminf(fmaxf(u, 0.0f), 1.0f);
        v2 ps = vadd(a, vscale(d, u));
        v2 r = vsub(cn, ps);
        float dd = vdot(r, r);
        if (dd < best) { best = dd; bu = u; bq = cn; }
    }
    *u_out = bu;
    *q_out = bq;
    return 1;
}

/* Variant ORC_VARIANT_GJK (sensitivity to deviation D1): the same closest points from a textbook
 * GJK on the Minkowski difference segment - box (parry: query::gjk::closest_points), in the box
 * frame: simplex of <= 3 support points, start direction `dir0` (parry passes the cached
 * normal, else the centre offset), termination when the support bound shows the distance is
 * known to parry's relative tolerance sqrt(10 * f32::EPSILON) or the simplex stops changing.
 * Returns 0 if the origin is enclosed / reached (the shapes touch), else 1 with the closest
 * points expressed as (u on the segment, q on the box). */
typedef struct { v2 w, pa, pb; } cso_pt; /* w = pa - pb, pa on the segment, pb on the box */
static cso_pt cso_support(v2 a, v2 b, v2 he, v2 dir) {
    cso_pt r;
    r.pa = vdot(a, dir) >= vdot(b, dir) ? a : b;
    v2 nd = vneg(dir);
    r.pb = V2(copysignf(he.x, nd.x), copysignf(he.y, nd.y)); /* parry: Cuboid::local_support_point */
    r.w = vsub(r.pa, r.pb);
    return r;
}
static int segment_box_closest_gjk(v2 a, v2 b, v2 he, v2 dir0, float* u_out, v2* q_out) {
    const float eps_tol = F32_EPS * 10.0f, eps_rel = sqrtf(eps_tol);
    cso_pt sx[3];
    int n = 0;
    float len0 = sqrtf(vdot(dir0, dir0));
    v2 dir = len0 > eps_tol ? V2(dir0.x / len0, dir0.y / len0) : V2(1.0f, 0.0f);
    sx[n++] = cso_support(a, b, he, vneg(dir));
    v2 proj = sx[0].w; /* closest point of the simplex to the origin */
    float bary[3] = {1.0f, 0.0f, 0.0f};
    float old_dist = 3.40282347e+38f;
    for (int iter = 0; iter < 100; ++iter) {
        float dist = sqrtf(vdot(proj, proj));
        if (dist <= eps_tol) return 0; /* origin on the simplex: shapes touch */
        if (!(dist < old_dist)) break;  /* no progress: converged to rounding */
        old_dist = dist;
        dir = V2(-proj.x / dist, -proj.y / dist);
        cso_pt s = cso_support(a, b, he, dir);
        float min_bound = -vdot(dir, s.w); /* lower bound on the distance */
        if (min_bound > 0.0f && dist - min_bound <= eps_rel * dist) break; /* parry's exit test */
        int dup = 0;
        for (int i = 0; i < n; ++i) if (sx[i].w.x == s.w.x && sx[i].w.y == s.w.y) dup = 1;
        if (dup) break;
        sx[n++] = s;
        /* project the origin on the simplex and reduce it to the supporting face */
        if (n == 2) {
            v2 ab = vsub(sx[1].w, sx[0].w);
            float t = -vdot(sx[0].w, ab) * inv0(vdot(ab, ab));
            if (t <= 0.0f) { n = 1; bary[0] = 1.0f; proj = sx[0].w; }
            else if (t >= 1.0f) { sx[0] = sx[1]; n = 1; bary[0] = 1.0f; proj = sx[0].w; }
            else { bary[0] = 1.0f - t; bary[1] = t; proj = vadd(sx[0].w, vscale(ab, t)); }
        } else { /* n == 3: triangle */
            float best = 3.40282347e+38f;
            int bi = -1, bj = -1;
            float bt = 0.0f;
            v2 bp = V2(0.0f, 0.0f);
            for (int i = 0; i < 3; ++i) {
                int j = (i + 1) % 3;
                v2 ab = vsub(sx[j].w, sx[i].w);
                float t = -vdot(sx[i].w, ab) * inv0(vdot(ab, ab));
                t = fminf(fmaxf(t, 0.0f), 1.0f);
                v2 pp = vadd(sx[i].w, vscale(ab, t));
                float dd = vdot(pp, pp);
                if (dd < best) { best = dd; bi = i; bj = j; bt = t; bp = pp; }
            }
            /* origin inside the triangle? (same sign of the three edge cross products) */
            float c0 = vcross(vsub(sx[1].w, sx[0].w), vneg(sx[0].w));
            float c1 = vcross(vsub(sx[2].w, sx[1].w), vneg(sx[1].w));
            float c2 = vcross(vsub(sx[0].w, sx[2].w), vneg(sx[2].w));
            if ((c0 >= 0.0f && c1 >= 0.0f && c2 >= 0.0f) || (c0 <= 0.0f && c1 <= 0.0f && c2 <= 0.0f)) return 0;
            cso_pt pi = sx[bi], pj = sx[bj];
            if (bt <= 0.0f) { sx[0] = pi; n = 1; bary[0] = 1.0f; }
            else if (bt >= 1.0f) { sx[0] = pj; n = 1; bary[0] = 1.0f; }
            else { sx[0] = pi; sx[1] = pj; n = 2; bary[0] = 1.0f - bt; bary[1] = bt; }
            proj = bp;
        }
    }
    v2 pa = V2(0.0f, 0.0f), pb = V2(0.0f, 0.0f);
    for (int i = 0; i < n; ++i) { pa = vadd(pa, vscale(sx[i].pa, bary[i])); pb = vadd(pb, vscale(sx[i].pb, bary[i])); }
    v2 d = vsub(b, a);
    float u = vdot(vsub(pa, a), d) * inv0(vdot(d, d));
    *u_out = fminf(fmaxf(u, 0.0f), 1.0f);
    *q_out = pb;
    return 1;
}

/* parry: contact_manifold_pfm_pfm with pfm1 = capsule segment (border radius r), pfm2 = cuboid */
static void contact_capsule_cuboid(pair_t* m, const iso* pos12, float radius, float hh, v2 he2, int variant) {
    v2 cached_n1 = m->n1;
    int had_normal = (m->n1.x != 0.0f || m->n1.y != 0.0f);
    if (manifold_try_update(m, pos12)) return;
    const float total_prediction = P_PREDICTION + radius + 0.0f;
    v2 sa = V2(0.0f, -hh), sb = V2(0.0f, hh);
    /* segment in the box frame */
    v2 a2 = iso_inv_mul_pt(pos12, sa), b2 = iso_inv_mul_pt(pos12, sb);
    float u;
    v2 q;
    v2 p1, p2_1, dir;
    m->npts = 0;
    int separated;
    if (variant & ORC_VARIANT_GJK) {
        /* start direction in the box frame: the cached normal, else towards the box centre */
        v2 d0 = had_normal ? rot_inv_mul(pos12->c, pos12->s, cached_n1) : vneg(vscale(vadd(a2, b2), 0.5f));
        separated = segment_box_closest_gjk(a2, b2, he2, d0, &u, &q);
        if (!separated) separated = segment_box_closest(a2, b2, he2, &u, &q); /* touching: exact routine decides */
    } else {
        separated = segment_box_closest(a2, b2, he2, &u, &q);
    }
    if (separated) {
        p1 = vadd(sa, vscale(vsub(sb, sa), u));
        p2_1 = iso_mul_pt(pos12, q);
        v2 dp = vsub(p2_1, p1);
        float sq = vdot(dp, dp);
        float n = sqrtf(sq);
        if (n > total_prediction) return; /* GJKResult::NoIntersection */
        dir = V2(dp.x / n, dp.y / n);
    } else {
        /* inner segment overlaps the box (parry: EPA). Min face penetration of the deepest
         * endpoint; documented approximation, unreachable at this game's speeds. */
        float best = 3.40282347e+38f;
        v2 nb = V2(0.0f, 1.0f), pe = a2, qf = a2;
        for (int f = 0; f < 4; ++f) {
            v2 nf = f == 0 ? V2(1.0f, 0.0f) : f == 1 ? V2(-1.0f, 0.0f) : f == 2 ? V2(0.0f, 1.0f) : V2(0.0f, -1.0f);
            float hf = f < 2 ? he2.x : he2.y;
            float da = vdot(a2, nf), db = vdot(b2, nf);
            v2 e = da < db ? a2 : b2;
            float depth = hf - fminf(da, db);
            if (depth < best) {
                best = depth;
                nb = nf;
                pe = e;
                qf = vadd(e, vscale(nf, depth));
            }
        }
        p1 = iso_mul_pt(pos12, pe);
        p2_1 = iso_mul_pt(pos12, qf);
        dir = vneg(rot_mul(pos12->c, pos12->s, nb));
    }
    v2 n2 = rot_inv_mul(pos12->c, pos12->s, vneg(dir));
    v2 f2a, f2b;
    cuboid_support_face(he2, n2, &f2a, &f2b);
    f2a = iso_mul_pt(pos12, f2a);
    f2b = iso_mul_pt(pos12, f2b);
    face_face_contacts(m, pos12, sa, sb, dir, f2a, f2b);
    if (m->npts == 0) {
        m->pt[0].p1 = p1;
        m->pt[0].p2 = iso_inv_mul_pt(pos12, p2_1);
        m->pt[0].dist = vdot(vsub(p2_1, p1), dir);
        m->npts = 1;
    }
    for (int k = 0; k < m->npts; ++k) {
        m->pt[k].p1 = vadd(m->pt[k].p1, vscale(dir, radius));
        m->pt[k].dist -= radius + 0.0f;
    }
    m->n1 = dir;
    m->n2 = n2;
}

/* rapier: NarrowPhase::compute_contacts, solver-contact generation for one pair */
static void pair_update_solver_contacts(pair_t* m, const collider_t* co1) {
    m->has_active = 0;
    m->nsolver = 0;
    m->normal = rot_mul(co1->pos.c, co1->pos.s, m->n1);
    for (int k = 0; k < m->npts; ++k) {
        if (m->pt[k].dist < P_PREDICTION) {
            v2 wp = iso_mul_pt(&co1->pos, m->pt[k].p1);
            v2 half = vscale(m->normal, m->pt[k].dist);
            scontact_t* sc = &m->sc[m->nsolver++];
            sc->point = vadd(wp, V2(half.x / 2.0f, half.y / 2.0f));
            sc->dist = m->pt[k].dist;
            m->has_active = 1;
        }
    }
}

static void body_wake_strong(body_t* b) {
    b->sleeping = 0;
    b->timer = 0.0f;
}

/* rapier: PhysicsPipeline::detect_collisions = BroadPhase::update + NarrowPhase::register_pairs
 * + compute_contacts, restricted to pairs with a modified collider. */
static void detect_collisions(orc_env* e) {
    for (int i = 0; i < e->ncolliders; ++i)
        if (e->colliders[i].modified) collider_update_aabb(&e->colliders[i]);
    for (int p = 0; p < e->npairs; ++p) {
        pair_t* m = &e->pairs[p];
        collider_t* co1 = &e->colliders[m->c1];
        collider_t* co2 = &e->colliders[m->c2];
        if (!co1->modified && !co2->modified) continue;
        if (!aabb_overlap(co1, co2)) {
            if (m->exists) {
                /* rapier: NarrowPhase::remove_pair wakes both bodies if the pair was active */
                if (m->has_active) {
                    if (co1->body >= 0) body_wake_strong(&e->bodies[co1->body]);
                    if (co2->body >= 0) body_wake_strong(&e->bodies[co2->body]);
                }
                m->exists = 0;
                m->npts = 0;
                m->nsolver = 0;
                m->has_active = 0;
                if (e->variant & ORC_VARIANT_GJK) m->n1 = V2(0.0f, 0.0f); /* the workspace dies with the pair */
                for (int i = 0; i < e->nedges; ++i)
                    if (e->edges[i] == p) { e->edges[i] = e->edges[--e->nedges]; break; } /* swap_remove */
            }
            continue;
        }
        if (!m->exists) e->edges[e->nedges++] = p;
        m->exists = 1;
        iso pos12 = iso_inv_mul(&co1->pos, &co2->pos);
        if (co1->shape == SHAPE_CAPSULE)
            contact_capsule_cuboid(m, &pos12, co1->he.x, co1->he.y, co2->he, e->variant);
        else
            contact_cuboid_cuboid(m, &pos12, co1->he, co2->he);
        pair_update_solver_contacts(m, co1);
    }
    for (int i = 0; i < e->ncolliders; ++i) e->colliders[i].modified = 0;
}

/* ------------------------------------------------------------------ bodies */

/* rapier: RigidBody::apply_impulse (deviation D4: zero impulse is a no-op) */
static int g_wake_on_zero_impulse = 0; /* set from the env's variant at every step (single-threaded studies only) */
static void body_apply_impulse(body_t* b, v2 j) {
    if (g_wake_on_zero_impulse && j.x == 0.0f && j.y == 0.0f) { body_wake_strong(b); return; }
    if (j.x != 0.0f || j.y != 0.0f) {
        b->linvel = vadd(b->linvel, vscale(j, b->inv_mass));
        body_wake_strong(b);
    }
}
/* rapier: RigidBody::apply_impulse_at_point */
static void body_apply_impulse_at_point(body_t* b, v2 j, v2 point) {
    float torque_impulse = vcross(vsub(point, b->pos.t), j);
    body_apply_impulse(b, j);
    if (torque_impulse != 0.0f) {
        b->angvel += b->inv_inertia_sqrt * (b->inv_inertia_sqrt * torque_impulse);
        body_wake_strong(b);
    }
}

/* rapier: RigidBodyVelocity::integrate, local_com = origin:
 *   shift * Isometry::new(linvel*dt, angvel*dt) * shift^-1 * init_pos, then renormalize_fast */
static iso integrate_pose(v2 linvel, float angvel, const iso* init) {
    v2 com = init->t; /* init_pos * origin */
    float ang = angvel * P_DT;
    float rc = libm_cosf(ang), rs = libm_sinf(ang);
    v2 tr = vscale(linvel, P_DT);
    /* (shift * iso): t = com + tr ; * shift^-1: t += R*(-com) ; * init: t += R*init.t, rot = R*init.rot */
    v2 t = vadd(com, tr);
    t = vadd(t, rot_mul(rc, rs, vneg(com)));
    t = vadd(t, rot_mul(rc, rs, init->t));
    iso r;
    r.t = t;
    r.c = rc * init->c - rs * init->s;
    r.s = rc * init->s + rs * init->c;
    float sq = r.c * r.c + r.s * r.s;
    float k = 0.5f * (3.0f - sq);
    r.c *= k;
    r.s *= k;
    return r;
}

/* ------------------------------------------------------------------ velocity constraints */
typedef struct {
    float gcross1, gcross2, rhs, rhs_wo_bias, impulse, r;
    float t_gcross1, t_gcross2, t_rhs, t_impulse, t_r;
} vc_elem_t;
typedef struct {
    int b1, b2;   /* b1 = -1: ground constraint (rapier: VelocityGroundConstraint) */
    v2 dir1, tangent1;
    float im1, im2, cfm_factor, limit;
    int n;
    vc_elem_t e[2];
} vc_t;

/* rapier: VelocityConstraint::generate / VelocityGroundConstraint::generate */
static void vc_generate(const orc_env* env, const pair_t* m, vc_t* c) {
    const collider_t* co1 = &env->colliders[m->c1];
    const collider_t* co2 = &env->colliders[m->c2];
    int h1 = co1->body, h2 = co2->body;
    int two_body = (h1 >= 0 && h2 >= 0);
    v2 force_dir1;
    v2 vel1_lin = V2(0.0f, 0.0f), com1 = V2(0.0f, 0.0f);
    float vel1_ang = 0.0f, im1 = 0.0f, ii1 = 0.0f, ccd;
    if (two_body) {
        force_dir1 = vneg(m->normal);
    } else if (h2 < 0) { /* relative_dominance < 0: flipped, dynamic body becomes body 2 */
        h2 = h1;
        h1 = -1;
        force_dir1 = m->normal;
    } else {
        force_dir1 = vneg(m->normal);
    }
    const body_t* rb2 = &env->bodies[h2];
    ccd = rb2->ccd_thickness;
    if (two_body) {
        const body_t* rb1 = &env->bodies[h1];
        vel1_lin = rb1->linvel;
        vel1_ang = rb1->angvel;
        com1 = rb1->pos.t;
        im1 = rb1->inv_mass;
        ii1 = rb1->inv_inertia_sqrt;
        ccd = rb1->ccd_thickness + rb2->ccd_thickness;
    }
    v2 tangent1 = V2(-force_dir1.y, force_dir1.x);
    c->b1 = h1;
    c->b2 = h2;
    c->dir1 = force_dir1;
    c->tangent1 = tangent1;
    c->im1 = im1;
    c->im2 = rb2->inv_mass;
    c->limit = 0.0f;
    c->n = m->nsolver;
    int is_fast = 0;
    for (int k = 0; k < m->nsolver; ++k) {
        const scontact_t* mp = &m->sc[k];
        v2 dp1 = vsub(mp->point, com1);
        v2 dp2 = vsub(mp->point, rb2->pos.t);
        v2 vel1 = vadd(vel1_lin, wcross(vel1_ang, dp1));
        v2 vel2 = vadd(rb2->linvel, wcross(rb2->angvel, dp2));
        c->limit = FRICTION;
        vc_elem_t* el = &c->e[k];
        {
            float gcross2 = rb2->inv_inertia_sqrt * vcross(dp2, vneg(force_dir1));
            float pm;
            float gcross1 = 0.0f;
            if (two_body) {
                gcross1 = ii1 * vcross(dp1, force_dir1);
                float imsum = im1 + rb2->inv_mass;
                pm = inv0(vdot(force_dir1, vscale(force_dir1, imsum)) + gcross1 * gcross1 + gcross2 * gcross2);
            } else {
                pm = inv0(vdot(force_dir1, vscale(force_dir1, rb2->inv_mass)) + gcross2 * gcross2);
            }
            /* restitution 0: is_bouncy = 0, is_resting = 1 */
            float rhs_wo_bias = (1.0f + 0.0f * 0.0f) * vdot(vsub(vel1, vel2), force_dir1);
            rhs_wo_bias += fmaxf(mp->dist, 0.0f) * env->inv_dt;
            rhs_wo_bias *= 0.0f + 1.0f;
            float pen = mp->dist + P_ALLOWED_LINEAR_ERROR;
            pen = fminf(fmaxf(pen, -P_MAX_PENETRATION_CORRECTION), 0.0f);
            float rhs_bias = env->erp_inv_dt * pen;
            float rhs = rhs_wo_bias + rhs_bias;
            is_fast = is_fast || (-rhs * P_DT > ccd * 0.5f);
            el->gcross1 = gcross1;
            el->gcross2 = gcross2;
            el->rhs = rhs;
            el->rhs_wo_bias = rhs_wo_bias;
            el->impulse = 0.0f;
            el->r = pm;
        }
        {
            el->t_impulse = 0.0f;
            float gcross2 = rb2->inv_inertia_sqrt * vcross(dp2, vneg(tangent1));
            float gcross1 = 0.0f, r;
            if (two_body) {
                gcross1 = ii1 * vcross(dp1, tangent1);
                float imsum = im1 + rb2->inv_mass;
                r = vdot(tangent1, vscale(tangent1, imsum)) + gcross1 * gcross1 + gcross2 * gcross2;
            } else {
                r = vdot(tangent1, vscale(tangent1, rb2->inv_mass)) + gcross2 * gcross2;
            }
            /* + tangent_velocity (zero vector) */
            v2 dv = vadd(vsub(vel1, vel2), V2(0.0f, 0.0f));
            el->t_rhs = vdot(dv, tangent1);
            el->t_gcross1 = gcross1;
            el->t_gcross2 = gcross2;
            el->t_r = inv0(r);
        }
    }
    c->cfm_factor = is_fast ? 1.0f : env->cfm_factor;
}

/* rapier: VelocityConstraintElement::solve_group / VelocityGroundConstraintElement::solve_group */
static void vc_solve(orc_env* env, vc_t* c, int solve_normal, int solve_friction) {
    body_t* rb2 = &env->bodies[c->b2];
    v2 l2 = rb2->dlin;
    float a2 = rb2->dang;
    if (c->b1 >= 0) {
        body_t* rb1 = &env->bodies[c->b1];
        v2 l1 = rb1->dlin;
        float a1 = rb1->dang;
        if (solve_normal) {
            for (int k = 0; k < c->n; ++k) {
                vc_elem_t* el = &c->e[k];
                float dvel = vdot(c->dir1, l1) + el->gcross1 * a1 - vdot(c->dir1, l2) + el->gcross2 * a2 + el->rhs;
                float new_impulse = c->cfm_factor * fmaxf(el->impulse - el->r * dvel, 0.0f);
                float dl = new_impulse - el->impulse;
                el->impulse = new_impulse;
                l1 = vadd(l1, vscale(vscale(c->dir1, c->im1), dl));
                a1 += el->gcross1 * dl;
                l2 = vadd(l2, vscale(vscale(c->dir1, c->im2), -dl));
                a2 += el->gcross2 * dl;
            }
        }
        if (solve_friction) {
            for (int k = 0; k < c->n; ++k) {
                vc_elem_t* el = &c->e[k];
                float limit = c->limit * el->impulse;
                float dvel = vdot(c->tangent1, l1) + el->t_gcross1 * a1 - vdot(c->tangent1, l2) + el->t_gcross2 * a2 + el->t_rhs;
                float new_impulse = fminf(fmaxf(el->t_impulse - el->t_r * dvel, -limit), limit);
                float dl = new_impulse - el->t_impulse;
                el->t_impulse = new_impulse;
                l1 = vadd(l1, vscale(vscale(c->tangent1, c->im1), dl));
                a1 += el->t_gcross1 * dl;
                l2 = vadd(l2, vscale(vscale(c->tangent1, c->im2), -dl));
                a2 += el->t_gcross2 * dl;
            }
        }
        rb1->dlin = l1;
        rb1->dang = a1;
    } else {
        if (solve_normal) {
            for (int k = 0; k < c->n; ++k) {
                vc_elem_t* el = &c->e[k];
                float dvel = -vdot(c->dir1, l2) + el->gcross2 * a2 + el->rhs;
                float new_impulse = c->cfm_factor * fmaxf(el->impulse - el->r * dvel, 0.0f);
                float dl = new_impulse - el->impulse;
                el->impulse = new_impulse;
                l2 = vadd(l2, vscale(vscale(c->dir1, c->im2), -dl));
                a2 += el->gcross2 * dl;
            }
        }
        if (solve_friction) {
            for (int k = 0; k < c->n; ++k) {
                vc_elem_t* el = &c->e[k];
                float limit = c->limit * el->impulse;
                float dvel = -vdot(c->tangent1, l2) + el->t_gcross2 * a2 + el->t_rhs;
                float new_impulse = fminf(fmaxf(el->t_impulse - el->t_r * dvel, -limit), limit);
                float dl = new_impulse - el->t_impulse;
                el->t_impulse = new_impulse;
                l2 = vadd(l2, vscale(vscale(c->tangent1, c->im2), -dl));
                a2 += el->t_gcross2 * dl;
            }
        }
    }
    rb2->dlin = l2;
    rb2->dang = a2;
}

/* ------------------------------------------------------------------ islands / sleeping */
/* rapier: IslandManager::update_active_set_with_contacts (+ update_energy) */
static void update_active_set(orc_env* e) {
    int nb = e->nbodies;
    int* stack = (int*)malloc(sizeof(int) * (size_t)(nb + 2 * e->npairs + 1));
    int* can_sleep = (int*)malloc(sizeof(int) * (size_t)nb);
    char* visited = (char*)calloc((size_t)nb, 1);
    int sp = 0, ncs = 0;
    for (int i = nb - 1; i >= 0; --i) {
        body_t* rb = &e->bodies[i];
        if (rb->sleeping) continue; /* not in the active set */
        float sq_lin = vdot(rb->linvel, rb->linvel);
        float sq_ang = rb->angvel * rb->angvel;
        if (sq_lin < ACT_LINEAR_THRESHOLD * fabsf(ACT_LINEAR_THRESHOLD) &&
            sq_ang < ACT_ANGULAR_THRESHOLD * fabsf(ACT_ANGULAR_THRESHOLD))
            rb->timer += P_DT;
        else
            rb->timer = 0.0f;
        if (rb->timer >= ACT_TIME_UNTIL_SLEEP) {
            rb->sleeping = 1;
            can_sleep[ncs++] = i;
        } else {
            stack[sp++] = i;
        }
    }
    while (sp > 0) {
        int h = stack[--sp];
        if (visited[h]) continue;
        visited[h] = 1;
        int ch = e->bodies[h].collider;
        for (int p = 0; p < e->npairs; ++p) {
            const pair_t* m = &e->pairs[p];
            if (!m->exists || m->nsolver == 0) continue;
            int other = -1;
            if (m->c1 == ch) other = m->c2;
            else if (m->c2 == ch) other = m->c1;
            if (other < 0) continue;
            int ob = e->colliders[other].body;
            if (ob >= 0) stack[sp++] = ob;
        }
        e->bodies[h].sleeping = 0; /* wake_up(false) */
    }
    for (int i = 0; i < ncs; ++i) {
        body_t* rb = &e->bodies[can_sleep[i]];
        if (rb->sleeping) { /* RigidBody::sleep */
            rb->timer = ACT_TIME_UNTIL_SLEEP;
            rb->linvel = V2(0.0f, 0.0f);
            rb->angvel = 0.0f;
        }
    }
    for (int i = 0; i < nb; ++i) e->bodies[i].active = !e->bodies[i].sleeping;
    free(stack);
    free(can_sleep);
    free(visited);
}

/* ------------------------------------------------------------------ pipeline step */
/* rapier: PhysicsPipeline::step, gravity (0,-2) (ref: common.rs:315-329) */
static void pipeline_step(orc_env* e) {
    const v2 gravity = V2(0.0f, -2.0f);
    if (e->fresh) {
        /* first step: every collider is new => detect_collisions on the initial poses */
        for (int i = 0; i < e->ncolliders; ++i) e->colliders[i].modified = 1;
        detect_collisions(e);
        e->fresh = 0;
    }
    update_active_set(e);
    /* rapier: NarrowPhase::select_active_contacts + categorize_contacts:
     * two-body constraints first, then ground constraints, each in pair order (D2) */
    vc_t* cons = (vc_t*)malloc(sizeof(vc_t) * (size_t)(e->npairs + 1));
    int* cons_pair = (int*)malloc(sizeof(int) * (size_t)(e->npairs + 1));
    int ncons = 0;
    for (int pass = 0; pass < 2; ++pass) {
        /* canonical: ascending pair index. Variants (sensitivity to deviation D2): descending,
         * or the contact graph's edge order (creation order with swap-remove) */
        const int nwalk = (e->variant & ORC_VARIANT_CREATION_ORDER) ? e->nedges : e->npairs;
        for (int w = 0; w < nwalk; ++w) {
            int p = (e->variant & ORC_VARIANT_CREATION_ORDER) ? e->edges[w] : ((e->variant & ORC_VARIANT_REVERSE_ORDER) ? e->npairs - 1 - w : w);
            const pair_t* m = &e->pairs[p];
            if (!m->exists || m->nsolver == 0) continue;
            int h1 = e->colliders[m->c1].body, h2 = e->colliders[m->c2].body;
            int two_body = (h1 >= 0 && h2 >= 0);
            if ((pass == 0) != two_body) continue;
            int awake1 = h1 >= 0 && !e->bodies[h1].sleeping;
            int awake2 = h2 >= 0 && !e->bodies[h2].sleeping;
            if (!awake1 && !awake2) continue;
            cons_pair[ncons++] = p; /* generated after the forces below */
        }
    }
    /* forces: force = gravity (*) mass * gravity_scale, mass = inv(effective_inv_mass) */
    for (int i = 0; i < e->nbodies; ++i) {
        body_t* rb = &e->bodies[i];
        if (!rb->active) continue;
        float mass = inv0(rb->inv_mass);
        rb->force = vadd(V2(0.0f, 0.0f), vscale(V2(gravity.x * mass, gravity.y * mass), 1.0f));
    }
    if (ncons > 0) {
        for (int i = 0; i < ncons; ++i) vc_generate(e, &e->pairs[cons_pair[i]], &cons[i]);
        /* rapier: VelocitySolver::solve */
        for (int i = 0; i < e->nbodies; ++i) {
            body_t* rb = &e->bodies[i];
            if (!rb->active) continue;
            rb->dang = 0.0f + rb->inv_inertia_sqrt * 0.0f * P_DT;
            rb->dlin = vadd(V2(0.0f, 0.0f), vscale(vscale(rb->force, rb->inv_mass), P_DT));
        }
        for (int it = 0; it < P_VELOCITY_ITERS; ++it) {
            int solve_friction = (P_FRICTION_ITERS + it >= P_VELOCITY_ITERS);
            for (int i = 0; i < ncons; ++i) vc_solve(e, &cons[i], 1, 0);
            if (solve_friction)
                for (int i = 0; i < ncons; ++i) vc_solve(e, &cons[i], 0, 1);
        }
        int remaining = P_FRICTION_ITERS > P_VELOCITY_ITERS ? P_FRICTION_ITERS - P_VELOCITY_ITERS : 0;
        for (int it = 0; it < remaining; ++it)
            for (int i = 0; i < ncons; ++i) vc_solve(e, &cons[i], 0, 1);
        /* integrate positions with the biased velocities */
        for (int i = 0; i < e->nbodies; ++i) {
            body_t* rb = &e->bodies[i];
            if (!rb->active) continue;
            float dangvel = rb->inv_inertia_sqrt * rb->dang;
            v2 nl = vadd(rb->linvel, rb->dlin);
            float na = rb->angvel + dangvel;
            rb->next_pos = integrate_pose(nl, na, &rb->pos);
        }
        /* remove bias */
        for (int i = 0; i < ncons; ++i) {
            cons[i].cfm_factor = 1.0f;
            for (int k = 0; k < cons[i].n; ++k) cons[i].e[k].rhs = cons[i].e[k].rhs_wo_bias;
        }
        for (int it = 0; it < P_STABILIZATION_ITERS; ++it) {
            for (int i = 0; i < ncons; ++i) vc_solve(e, &cons[i], 1, 0);
            for (int i = 0; i < ncons; ++i) vc_solve(e, &cons[i], 0, 1);
        }
        for (int i = 0; i < e->nbodies; ++i) {
            body_t* rb = &e->bodies[i];
            if (!rb->active) continue;
            float dangvel = rb->inv_inertia_sqrt * rb->dang;
            rb->linvel = vadd(rb->linvel, rb->dlin);
            rb->angvel += dangvel;
            /* advance_to_final_positions */
            rb->pos = rb->next_pos;
        }
    } else {
        /* rapier: IslandSolver::init_and_solve, no-constraint branch */
        for (int i = 0; i < e->nbodies; ++i) {
            body_t* rb = &e->bodies[i];
            if (!rb->active) continue;
            rb->linvel = vadd(rb->linvel, vscale(vscale(rb->force, rb->inv_mass), P_DT));
            rb->angvel += rb->inv_inertia_sqrt * (rb->inv_inertia_sqrt * 0.0f * P_DT);
            rb->pos = integrate_pose(rb->linvel, rb->angvel, &rb->pos);
        }
    }
    free(cons);
    free(cons_pair);
    /* colliders follow their bodies (pos_wrt_parent = identity) and are marked modified */
    for (int i = 0; i < e->nbodies; ++i) {
        body_t* rb = &e->bodies[i];
        if (!rb->active) continue;
        collider_t* co = &e->colliders[rb->collider];
        co->pos = rb->pos;
        co->modified = 1;
    }
    detect_collisions(e);
}

/* ------------------------------------------------------------------ Environment (ref) */

/* parry: MassProperties::from_cuboid (dim2), density 1 */
static void cuboid_mass(v2 he, float* inv_mass, float* inv_inertia_sqrt) {
    float volume = he.x * he.y * 4.0f;
    float ix = (he.x * he.x) / 3.0f;
    float iy = (he.y * he.y) / 3.0f;
    float mass = volume * 1.0f;
    float inertia = (ix + iy) * mass;
    *inv_mass = inv0(mass);
    *inv_inertia_sqrt = inv0(sqrtf(inertia));
}
/* parry: MassProperties::from_capsule (dim2), density 1; only the mass matters (rotations locked) */
static float capsule_inv_mass(float hh_in, float radius) {
    float half_height = (hh_in - (-hh_in)) / 2.0f; /* |b - a| / 2 with a=(0,-hh), b=(0,hh) */
    float cyl_vol = radius * half_height * 4.0f;
    float ball_vol = 3.14159265358979323846f * radius * radius;
    float cap_vol = cyl_vol + ball_vol;
    float cap_mass = cap_vol * 1.0f;
    return inv0(cap_mass);
}

orc_env* orc_env_create(const float player_position[2], const orc_object* objs, int32_t n) {
    orc_env* e = (orc_env*)calloc(1, sizeof(orc_env));
    if (!e) return NULL;
    int nblocks = 0, ndyn = 0, ngoals = 0;
    for (int i = 0; i < n; ++i) {
        if (objs[i].kind == 2) ++ngoals;
        else { ++nblocks; if (objs[i].kind == 1) ++ndyn; }
    }
    e->nbodies = 1 + ndyn;
    e->ncolliders = 1 + nblocks;
    e->ngoals = ngoals;
    e->bodies = (body_t*)calloc((size_t)e->nbodies, sizeof(body_t));
    e->bodies0 = (body_t*)calloc((size_t)e->nbodies, sizeof(body_t));
    e->colliders = (collider_t*)calloc((size_t)e->ncolliders, sizeof(collider_t));
    e->colliders0 = (collider_t*)calloc((size_t)e->ncolliders, sizeof(collider_t));
    e->goals = (goal_t*)calloc((size_t)(ngoals ? ngoals : 1), sizeof(goal_t));

    /* rapier: IntegrationParameters::{inv_dt, erp_inv_dt, cfm_factor} */
    e->inv_dt = 1.0f / P_DT;
    e->erp_inv_dt = P_ERP * e->inv_dt;
    {
        float inv_erp_minus_one = 1.0f / P_ERP - 1.0f;
        float cfm_coeff = inv_erp_minus_one * inv_erp_minus_one /
                          ((1.0f + inv_erp_minus_one) * 4.0f * P_DAMPING_RATIO * P_DAMPING_RATIO);
        e->cfm_factor = 1.0f / (1.0f + cfm_coeff);
    }

    /* ref: common.rs:70-86 Environment::new */
    {
        body_t* rb = &e->bodies[0];
        rb->pos.t = V2(player_position[0] * BEVY_TO_PHYSICS_SCALE, player_position[1] * BEVY_TO_PHYSICS_SCALE);
        rb->pos.c = 1.0f;
        rb->pos.s = 0.0f;
        float hh = 0.5f * PLAYER_DEPTH * BEVY_TO_PHYSICS_SCALE;
        float radius = PLAYER_RADIUS * BEVY_TO_PHYSICS_SCALE;
        rb->inv_mass = capsule_inv_mass(hh, radius);
        rb->inv_inertia_sqrt = 0.0f; /* lock_rotations */
        rb->ccd_thickness = radius;
        rb->collider = 0;
        collider_t* co = &e->colliders[0];
        co->shape = SHAPE_CAPSULE;
        co->body = 0;
        co->he = V2(radius, hh);
        co->pos = rb->pos;
    }
    /* ref: common.rs:107-159 add_object */
    int ci = 1, bi = 1, gi = 0;
    for (int i = 0; i < n; ++i) {
        const orc_object* o = &objs[i];
        if (o->kind == 2) {
            goal_t* g = &e->goals[gi++];
            g->x = o->position[0] * BEVY_TO_PHYSICS_SCALE;
            g->y = o->position[1] * BEVY_TO_PHYSICS_SCALE;
            g->w = fabsf(o->scale[0]) * BEVY_TO_PHYSICS_SCALE;
            g->h = fabsf(o->scale[1]) * BEVY_TO_PHYSICS_SCALE;
            g->rot = o->rotation;
            /* glam: Quat::from_rotation_z(rot) * Vec3::X / Vec3::Y (std sin_cos -> host libm) */
            float s = sinf(g->rot * 0.5f), c = cosf(g->rot * 0.5f);
            float k = c * c - ((0.0f * 0.0f + 0.0f * 0.0f) + s * s);
            float w2 = c * 2.0f;
            g->x_axis = V2(k + 0.0f + 0.0f, 0.0f + 0.0f + s * w2);
            g->y_axis = V2(0.0f + 0.0f + (0.0f - s) * w2, k + 0.0f + 0.0f);
            continue;
        }
        collider_t* co = &e->colliders[ci];
        co->shape = SHAPE_CUBOID;
        co->he = V2(0.5f * fabsf(o->scale[0]) * BEVY_TO_PHYSICS_SCALE, 0.5f * fabsf(o->scale[1]) * BEVY_TO_PHYSICS_SCALE);
        co->pos.t = V2(o->position[0] * BEVY_TO_PHYSICS_SCALE, o->position[1] * BEVY_TO_PHYSICS_SCALE);
        co->pos.c = libm_cosf(o->rotation);
        co->pos.s = libm_sinf(o->rotation);
        if (o->kind == 0) {
            co->body = -1;
        } else {
            body_t* rb = &e->bodies[bi];
            rb->pos = co->pos;
            cuboid_mass(co->he, &rb->inv_mass, &rb->inv_inertia_sqrt);
            rb->ccd_thickness = fminf(co->he.x, co->he.y);
            rb->collider = ci;
            co->body = bi++;
        }
        ++ci;
    }
    /* candidate pairs, ascending (c1, c2); fixed-fixed never collide */
    int np = 0;
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) e->pairs = (pair_t*)calloc((size_t)(np ? np : 1), sizeof(pair_t));
        int k = 0;
        for (int a = 0; a < e->ncolliders; ++a)
            for (int b = a + 1; b < e->ncolliders; ++b) {
                if (e->colliders[a].body < 0 && e->colliders[b].body < 0) continue;
                if (pass == 1) { e->pairs[k].c1 = a; e->pairs[k].c2 = b; }
                ++k;
            }
        np = k;
    }
    e->npairs = np;
    e->edges = (int*)calloc((size_t)(np ? np : 1), sizeof(int));
    e->variant = g_default_variant;
    memcpy(e->bodies0, e->bodies, sizeof(body_t) * (size_t)e->nbodies);
    memcpy(e->colliders0, e->colliders, sizeof(collider_t) * (size_t)e->ncolliders);
    orc_env_reset(e);
    return e;
}

void orc_env_reset(orc_env* e) {
    memcpy(e->bodies, e->bodies0, sizeof(body_t) * (size_t)e->nbodies);
    memcpy(e->colliders, e->colliders0, sizeof(collider_t) * (size_t)e->ncolliders);
    for (int p = 0; p < e->npairs; ++p) {
        int c1 = e->pairs[p].c1, c2 = e->pairs[p].c2;
        memset(&e->pairs[p], 0, sizeof(pair_t));
        e->pairs[p].c1 = c1;
        e->pairs[p].c2 = c2;
    }
    e->nedges = 0;
    e->won = 0;
    e->fresh = 1;
}

void orc_env_destroy(orc_env* e) {
    if (!e) return;
    free(e->bodies); free(e->bodies0); free(e->colliders); free(e->colliders0);
    free(e->pairs); free(e->goals); free(e->edges); free(e);
}

/* ref: common.rs:175-195 */
int orc_env_distance_to_goals(const orc_env* e, float* out) {
    if (e->ngoals == 0) return -1;
    v2 p = e->bodies[0].pos.t;
    float best = 0.0f;
    for (int i = 0; i < e->ngoals; ++i) {
        const goal_t* g = &e->goals[i];
        v2 d = vsub(p, V2(g->x, g->y));
        float dx = fmaxf(fabsf(d.x * g->x_axis.x + d.y * g->x_axis.y) - g->w / 2.0f, 0.0f);
        float dy = fmaxf(fabsf(d.x * g->y_axis.x + d.y * g->y_axis.y) - g->h / 2.0f, 0.0f);
        float dist = sqrtf(dx * dx + dy * dy) / BEVY_TO_PHYSICS_SCALE;
        best = i == 0 ? dist : fminf(best, dist);
    }
    *out = best;
    return 0;
}

int orc_env_won(const orc_env* e) { return e->won; }
void orc_env_set_variant(orc_env* e, int32_t flags) { e->variant = flags; }

/* ref: common.rs:215-340 */
void orc_env_step(orc_env* e, uint32_t mv) {
    body_t* player = &e->bodies[0];
    v2 lower = V2(player->pos.t.x, player->pos.t.y - PLAYER_DEPTH * BEVY_TO_PHYSICS_SCALE / 2.0f);
    /* floor contacts (point, other collider's body); at most 2 per pair */
    int cap = 2 * e->npairs + 2, nfc = 0;
    v2* fpt = (v2*)malloc(sizeof(v2) * (size_t)cap);
    int* fbody = (int*)malloc(sizeof(int) * (size_t)cap);
    g_wake_on_zero_impulse = (e->variant & ORC_VARIANT_WAKE_ON_ZERO_IMPULSE) != 0;
    if (!e->fresh) {
        /* contacts_with(player_collider): canonical ascending pair order; variants: descending, or
         * the player's adjacency list in the contact graph (petgraph: newest edge first) */
        const int nwalk = (e->variant & ORC_VARIANT_CREATION_ORDER) ? e->nedges : e->npairs;
        for (int w = 0; w < nwalk; ++w) {
            int p = (e->variant & ORC_VARIANT_CREATION_ORDER) ? e->edges[e->nedges - 1 - w] : ((e->variant & ORC_VARIANT_REVERSE_ORDER) ? e->npairs - 1 - w : w);
            const pair_t* m = &e->pairs[p];
            if (m->c1 != 0 || !m->exists) continue; /* contacts_with(player_collider) */
            if (!m->has_active) continue;
            for (int k = 0; k < m->nsolver; ++k) {
                v2 q = vsub(m->sc[k].point, lower);
                float r = PLAYER_RADIUS * BEVY_TO_PHYSICS_SCALE;
                q = V2(q.x / r, q.y / r);
                if (q.y < -0.707f) {
                    fpt[nfc] = m->sc[k].point;
                    fbody[nfc] = e->colliders[m->c2].body;
                    ++nfc;
                }
            }
        }
    }
    if (nfc > 0) {
        v2 player_impulse = V2(0.0f, 0.0f);
        if (mv & 1u) { /* left: min_by x, first minimum */
            int b = 0;
            for (int i = 1; i < nfc; ++i) if (fpt[i].x < fpt[b].x) b = i;
            v2 normal = vsub(fpt[b], lower);
            float len = sqrtf(normal.x * normal.x + normal.y * normal.y);
            normal = V2(normal.x / len, normal.y / len);
            v2 impulse = V2(0.003f * normal.y, -0.003f * normal.x);
            if (fbody[b] >= 0) body_apply_impulse_at_point(&e->bodies[fbody[b]], vneg(impulse), fpt[b]);
            player_impulse = vadd(player_impulse, impulse);
        }
        if (mv & 2u) { /* right: max_by x, last maximum */
            int b = 0;
            for (int i = 1; i < nfc; ++i) if (!(fpt[i].x < fpt[b].x)) b = i;
            v2 normal = vsub(fpt[b], lower);
            float len = sqrtf(normal.x * normal.x + normal.y * normal.y);
            normal = V2(normal.x / len, normal.y / len);
            v2 impulse = V2(-0.003f * normal.y, 0.003f * normal.x);
            if (fbody[b] >= 0) body_apply_impulse_at_point(&e->bodies[fbody[b]], vneg(impulse), fpt[b]);
            player_impulse = vadd(player_impulse, impulse);
        }
        if (mv & 4u) {
            for (int i = 0; i < nfc; ++i) {
                v2 normal = vsub(fpt[i], lower);
                float len = sqrtf(normal.x * normal.x + normal.y * normal.y);
                normal = V2(normal.x / len, normal.y / len);
                float cnt = (float)nfc;
                v2 impulse = V2(-0.1f * normal.x / cnt, -0.1f * normal.y / cnt);
                if (fbody[i] >= 0) body_apply_impulse_at_point(&e->bodies[fbody[i]], vneg(impulse), fpt[i]);
                player_impulse = vadd(player_impulse, impulse);
            }
        }
        body_apply_impulse(player, player_impulse);
    }
    free(fpt);
    free(fbody);

    pipeline_step(e);
    /* query_pipeline.update (common.rs:330-331) is never queried: omitted */

    if (!e->won) {
        float d;
        if (orc_env_distance_to_goals(e, &d) == 0 && d < 1e-7f) e->won = 1;
    }
}

/* ------------------------------------------------------------------ observation */
int32_t orc_env_num_bodies(const orc_env* e) { return e->nbodies; }
void orc_env_body_state(const orc_env* e, int32_t b, float o[9]) {
    const body_t* rb = &e->bodies[b];
    o[0] = rb->pos.t.x; o[1] = rb->pos.t.y; o[2] = rb->pos.c; o[3] = rb->pos.s;
    o[4] = rb->linvel.x; o[5] = rb->linvel.y; o[6] = rb->angvel; o[7] = rb->timer;
    o[8] = (float)rb->sleeping;
}
void orc_env_set_body_state(orc_env* e, int32_t b, const float in[9]) {
    body_t* rb = &e->bodies[b];
    rb->pos.t = V2(in[0], in[1]); rb->pos.c = in[2]; rb->pos.s = in[3];
    rb->linvel = V2(in[4], in[5]); rb->angvel = in[6]; rb->timer = in[7];
    rb->sleeping = in[8] != 0.0f;
    e->colliders[rb->collider].pos = rb->pos;
}
int32_t orc_env_num_pairs(const orc_env* e) { return e->npairs; }
void orc_env_pair(const orc_env* e, int32_t p, int32_t out2[2]) { out2[0] = e->pairs[p].c1; out2[1] = e->pairs[p].c2; }
void orc_env_manifold(const orc_env* e, int32_t p, float o[16]) {
    const pair_t* m = &e->pairs[p];
    memset(o, 0, sizeof(float) * 16);
    o[0] = (float)m->npts;
    o[1] = m->n1.x; o[2] = m->n1.y; o[3] = m->n2.x; o[4] = m->n2.y;
    for (int k = 0; k < m->npts; ++k) {
        o[5 + 5 * k] = m->pt[k].p1.x; o[6 + 5 * k] = m->pt[k].p1.y;
        o[7 + 5 * k] = m->pt[k].p2.x; o[8 + 5 * k] = m->pt[k].p2.y;
        o[9 + 5 * k] = m->pt[k].dist;
    }
    o[15] = (float)m->exists;
}

/* ------------------------------------------------------------------ GA rollout (ref: main.rs:44-70) */
float orc_rollout(orc_env* e, const uint8_t* genome, int32_t genome_len, int32_t repeat_move,
                  int32_t number_of_steps, int32_t early_exit, int32_t* step_of_win, int32_t* steps_run) {
    orc_env_reset(e);
    float score = INFINITY;
    int32_t steps = 0, win_step = -1;
    int stop = 0;
    for (int32_t g = 0; g < genome_len && !stop; ++g) {
        for (int32_t r = 0; r < repeat_move; ++r) {
            orc_env_step(e, genome[g] & 7u);
            ++steps;
            float d = 0.0f;
            orc_env_distance_to_goals(e, &d);
            score = fminf(score, d);
            if (e->won && win_step < 0) win_step = steps;
            if (e->won && early_exit) break;
        }
        if (e->won && early_exit) stop = 1;
    }
    /* the tail runs even after a win in the reference (one extra step, then break) */
    for (int32_t r = 0; r < number_of_steps % repeat_move; ++r) {
        orc_env_step(e, 0u);
        ++steps;
        float d = 0.0f;
        orc_env_distance_to_goals(e, &d);
        score = fminf(score, d);
        if (e->won && win_step < 0) win_step = steps;
        if (e->won && early_exit) break;
    }
    if (step_of_win) *step_of_win = win_step;
    if (steps_run) *steps_run = steps;
    return score;
}

typedef struct {
    const float* player; const orc_object* objs; int32_t n;
    const uint8_t* genomes; int64_t lo, hi; int32_t genome_len, repeat_move, number_of_steps, early_exit;
    float* out_score; int32_t* out_win; int64_t steps;
} batch_job_t;

static void* batch_worker(void* arg) {
    batch_job_t* j = (batch_job_t*)arg;
    orc_env* e = orc_env_create(j->player, j->objs, j->n);
    j->steps = 0;
    for (int64_t i = j->lo; i < j->hi; ++i) {
        int32_t win, run;
        float s = orc_rollout(e, j->genomes + i * j->genome_len, j->genome_len, j->repeat_move,
                              j->number_of_steps, j->early_exit, &win, &run);
        if (j->out_score) j->out_score[i] = s;
        if (j->out_win) j->out_win[i] = win;
        j->steps += run;
    }
    orc_env_destroy(e);
    return NULL;
}

int64_t orc_rollout_batch(const float player_position[2], const orc_object* objs, int32_t n,
                          const uint8_t* genomes, int64_t n_envs, int32_t genome_len,
                          int32_t repeat_move, int32_t number_of_steps, int32_t early_exit,
                          int32_t n_threads, float* out_score, int32_t* out_step_of_win) {
    if (n_threads < 1) n_threads = 1;
    if ((int64_t)n_threads > n_envs) n_threads = (int32_t)(n_envs > 0 ? n_envs : 1);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
    batch_job_t* jobs = (batch_job_t*)calloc((size_t)n_threads, sizeof(batch_job_t));
    for (int t = 0; t < n_threads; ++t) {
        batch_job_t* j = &jobs[t];
        j->player = player_position; j->objs = objs; j->n = n; j->genomes = genomes;
        j->lo = n_envs * t / n_threads; j->hi = n_envs * (t + 1) / n_threads;
        j->genome_len = genome_len; j->repeat_move = repeat_move; j->number_of_steps = number_of_steps;
        j->early_exit = early_exit; j->out_score = out_score; j->out_win = out_step_of_win;
        pthread_create(&th[t], NULL, batch_worker, j);
    }
    int64_t total = 0;
    for (int t = 0; t < n_threads; ++t) { pthread_join(th[t], NULL); total += jobs[t].steps; }
    free(th);
    free(jobs);
    return total;
}

/* ------------------------------------------------------------------ observation-driven agents
 * The batched form of Agent::get_move(&mut self, &Environment) (ref: algorithm.rs:8-11) that the
 * product evaluates inside its fused rollout (include/prl.h prl_batch_rollout_policy): the
 * observation an agent can read between two steps and a two-layer perceptron on it. Restated
 * here independently so the kernels can be checked bit for bit. */
int32_t orc_env_observation_dim(const orc_env* e) { return 9 * e->nbodies + 4; }

/* obs: per body x, y, cos, sin, vx, vy, angvel, sleep timer, sleeping (ref: common.rs:201-211);
 * on_ground, number of floor contacts, touching a pushable block (the scan of ref:
 * common.rs:223-244 over narrow_phase.contacts_with(player_collider)); distance_to_goals in px */
void orc_env_observation(const orc_env* e, float* obs) {
    for (int b = 0; b < e->nbodies; ++b) orc_env_body_state(e, b, obs + 9 * b);
    const body_t* player = &e->bodies[0];
    float lower_y = player->pos.t.y - PLAYER_DEPTH * BEVY_TO_PHYSICS_SCALE / 2.0f;
    int nfc = 0, touching = 0;
    if (!e->fresh) {
        for (int p = 0; p < e->npairs; ++p) {
            const pair_t* m = &e->pairs[p];
            if (m->c1 != 0 || !m->exists || !m->has_active) continue;
            if (e->colliders[m->c2].body >= 0) touching = 1;
            for (int k = 0; k < m->nsolver; ++k) {
                float r = PLAYER_RADIUS * BEVY_TO_PHYSICS_SCALE;
                float qy = (m->sc[k].point.y - lower_y) / r;
                if (qy < -0.707f) ++nfc;
            }
        }
    }
    float* x = obs + 9 * e->nbodies;
    x[0] = nfc > 0 ? 1.0f : 0.0f;
    x[1] = (float)nfc;
    x[2] = touching ? 1.0f : 0.0f;
    float d = INFINITY;
    orc_env_distance_to_goals(e, &d);
    x[3] = d;
}

/* h = max(W1 obs + b1, 0), out = W2 h + b2, Move bit k = out_k > 0; sums left to right in f32.
 * theta[p * stride]: parameter p of this agent (parameter-major storage, stride = agents) */
uint32_t orc_policy_move(const float* theta, int64_t stride, int32_t hidden, int32_t D, const float* obs) {
    float h[64];
    for (int j = 0; j < hidden; ++j) {
        const float* row = theta + (int64_t)j * (D + 1) * stride;
        float acc = row[(int64_t)D * stride];
        for (int i = 0; i < D; ++i) acc = acc + row[(int64_t)i * stride] * obs[i];
        h[j] = fmaxf(acc, 0.0f);
    }
    uint32_t mv = 0;
    const float* w2 = theta + (int64_t)hidden * (D + 1) * stride;
    for (int k = 0; k < 3; ++k) {
        const float* row = w2 + (int64_t)k * (hidden + 1) * stride;
        float acc = row[(int64_t)hidden * stride];
        for (int j = 0; j < hidden; ++j) acc = acc + row[(int64_t)j * stride] * h[j];
        if (acc > 0.0f) mv |= 1u << k;
    }
    return mv;
}

/* ref: main.rs:44-70 with the agent's get_move in place of the genome, repeat_move = 1 */
float orc_rollout_policy(orc_env* e, const float* theta, int64_t stride, int32_t hidden, int32_t number_of_steps,
                         int32_t early_exit, int32_t* step_of_win, int32_t* steps_run) {
    orc_env_reset(e);
    float score = INFINITY;
    int32_t steps = 0, win_step = -1;
    const int32_t D = orc_env_observation_dim(e);
    float* obs = (float*)malloc(sizeof(float) * (size_t)D);
    for (int32_t t = 0; t < number_of_steps; ++t) {
        orc_env_observation(e, obs);
        orc_env_step(e, orc_policy_move(theta, stride, hidden, D, obs));
        ++steps;
        float d = 0.0f;
        orc_env_distance_to_goals(e, &d);
        score = fminf(score, d);
        if (e->won && win_step < 0) win_step = steps;
        if (e->won && early_exit) break;
    }
    free(obs);
    if (step_of_win) *step_of_win = win_step;
    if (steps_run) *steps_run = steps;
    return score;
}

typedef struct {
    const float* player; const orc_object* objs; int32_t n;
    const float* theta; int64_t n_envs, lo, hi; int32_t hidden, number_of_steps, early_exit;
    float* out_score; int32_t* out_win; int64_t steps;
} policy_job_t;

static void* policy_worker(void* arg) {
    policy_job_t* j = (policy_job_t*)arg;
    orc_env* e = orc_env_create(j->player, j->objs, j->n);
    j->steps = 0;
    for (int64_t i = j->lo; i < j->hi; ++i) {
        int32_t win, run;
        float s = orc_rollout_policy(e, j->theta + i, j->n_envs, j->hidden, j->number_of_steps, j->early_exit, &win, &run);
        if (j->out_score) j->out_score[i] = s;
        if (j->out_win) j->out_win[i] = win;
        j->steps += run;
    }
    orc_env_destroy(e);
    return NULL;
}

int64_t orc_rollout_policy_batch(const float player_position[2], const orc_object* objs, int32_t n, const float* theta,
                                 int64_t n_envs, int32_t hidden, int32_t number_of_steps, int32_t early_exit,
                                 int32_t n_threads, float* out_score, int32_t* out_step_of_win) {
    if (n_threads < 1) n_threads = 1;
    if ((int64_t)n_threads > n_envs) n_threads = (int32_t)(n_envs > 0 ? n_envs : 1);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
    policy_job_t* jobs = (policy_job_t*)calloc((size_t)n_threads, sizeof(policy_job_t));
    for (int t = 0; t < n_threads; ++t) {
        policy_job_t* j = &jobs[t];
        j->player = player_position; j->objs = objs; j->n = n; j->theta = theta; j->n_envs = n_envs;
        j->lo = n_envs * t / n_threads; j->hi = n_envs * (t + 1) / n_threads;
        j->hidden = hidden; j->number_of_steps = number_of_steps; j->early_exit = early_exit;
        j->out_score = out_score; j->out_win = out_step_of_win;
        pthread_create(&th[t], NULL, policy_worker, j);
    }
    int64_t total = 0;
    for (int t = 0; t < n_threads; ++t) { pthread_join(th[t], NULL); total += jobs[t].steps; }
    free(th);
    free(jobs);
    return total;
}
