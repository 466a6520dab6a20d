// prl_world.cc -- `World` ingest (serde JSON schema, reference src/common.rs:19-49) and the host
// side of `Environment::from_world` (common.rs:70-172): every level object becomes a row of the
// static DevWorld tables the kernels stage in shared memory.
//
// All physics-space values are computed in f32 with the expression order of the reference and
// of the crates it calls (rapier2d 0.17.2 builders, parry2d 0.13.4 mass properties); this file
// must be compiled without FMA contraction (-ffp-contract=off).
#include <math.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cmath>

#include <charconv>

#include "prl_host.h"
#include "prl_libm.h"

// layout of the PODs of include/prl.h as the Rust (rust/src/ffi.rs) and ctypes (_lib.py) mirrors declare them
static_assert(sizeof(prl_object) == 28 && offsetof(prl_object, scale) == 12 && offsetof(prl_object, rotation) == 20 && offsetof(prl_object, kind) == 24, "prl_object layout");
static_assert(sizeof(prl_ga_config) == 40 && offsetof(prl_ga_config, number_of_agents) == 8 && offsetof(prl_ga_config, repeat_move) == 16 &&
              offsetof(prl_ga_config, mutation_rate) == 20 && offsetof(prl_ga_config, keep_best) == 24 && offsetof(prl_ga_config, seed) == 32, "prl_ga_config layout");

namespace prl {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }

static inline float inv0(float x) { return x == 0.0f ? 0.0f : 1.0f / x; }

// common.rs:7-9
static const float kPlayerDepth = 20.0f, kPlayerRadius = 20.0f;
static const float kScale = 0.25f / (2.0f * 20.0f);

int build_dev_world(prl_world* w) {
    DevWorld& d = w->dev;
    memset(&d, 0, sizeof d);
    const int n = (int)w->objects.size();
    int nblocks = 0, ndyn = 0, ngoals = 0;
    for (int i = 0; i < n; ++i) {
        int k = w->objects[i].kind;
        if (k == 2) ++ngoals;
        else if (k == 0 || k == 1) { ++nblocks; if (k == 1) ++ndyn; }
        else { set_error("object %d: kind %d is not 0 (fixed block), 1 (pushable block) or 2 (goal)", i, k); return PRL_ERR_INVALID; }
    }
    if (nblocks > PRL_MAX_BLOCKS) { set_error("world has %d blocks, limit PRL_MAX_BLOCKS=%d", nblocks, PRL_MAX_BLOCKS); return PRL_ERR_CAPACITY; }
    if (ndyn > PRL_MAX_DYNAMIC) { set_error("world has %d pushable blocks, limit PRL_MAX_DYNAMIC=%d", ndyn, PRL_MAX_DYNAMIC); return PRL_ERR_CAPACITY; }
    if (ngoals > PRL_MAX_GOALS) { set_error("world has %d goals, limit PRL_MAX_GOALS=%d", ngoals, PRL_MAX_GOALS); return PRL_ERR_CAPACITY; }
    d.n_colliders = 1 + nblocks;
    d.n_bodies = 1 + ndyn;
    d.n_goals = ngoals;

    // rapier IntegrationParameters::default() and its derived inv_dt / erp_inv_dt / cfm_factor
    const float dt = 1.0f / 60.0f, erp = 0.8f, damping_ratio = 0.25f;
    d.dt = dt;
    d.inv_dt = 1.0f / dt;
    d.erp_inv_dt = erp * d.inv_dt;
    {
        float a = 1.0f / erp - 1.0f;
        float cfm_coeff = a * a / ((1.0f + a) * 4.0f * damping_ratio * damping_ratio);
        d.cfm_factor = 1.0f / (1.0f + cfm_coeff);
    }
    d.prediction = 0.002f;
    d.allowed_linear_error = 0.001f;
    d.aabb_margin = d.prediction / 2.0f;
    d.scale = kScale;
    {
        // largest f32 q with sqrtf(q) / scale < 1e-7f (Environment::step's win threshold, common.rs:335):
        // both operations are correctly rounded and monotone in q, so bisect on the bit pattern
        uint32_t lo = 0u, hi = 0x7f800000u;  // q = 0 passes, +inf does not
        while (hi - lo > 1u) {
            uint32_t mid = lo + (hi - lo) / 2u;
            float q;
            memcpy(&q, &mid, 4);
            if (sqrtf(q) / d.scale < 1e-7f) lo = mid; else hi = mid;
        }
        memcpy(&d.win_q, &lo, 4);
    }

    // Environment::new (common.rs:70-86): rotation-locked dynamic body + capsule_y collider
    d.half_height = 0.5f * kPlayerDepth * kScale;
    d.radius = kPlayerRadius * kScale;
    d.lower_off = kPlayerDepth * kScale / 2.0f;  // common.rs:219
    d.cx[0] = w->player_position[0] * kScale;
    d.cy[0] = w->player_position[1] * kScale;
    d.cc[0] = 1.0f;
    d.cs[0] = 0.0f;
    d.hx[0] = d.radius;
    d.hy[0] = d.half_height;
    d.body_of[0] = 0;
    d.collider_of[0] = 0;
    {
        // parry MassProperties::from_capsule (2-D), density 1; rotations locked -> no angular part
        float hh = (d.half_height - (-d.half_height)) / 2.0f;
        float cyl = d.radius * hh * 4.0f;
        float ball = 3.14159265358979323846f * d.radius * d.radius;
        float mass = (cyl + ball) * 1.0f;
        d.inv_mass[0] = inv0(mass);
        d.inv_inertia_sqrt[0] = 0.0f;
        d.ccd_thickness[0] = d.radius;
    }

    // add_object (common.rs:107-159)
    w->object_body.assign(n, -1);
    int ci = 1, bi = 1, gi = 0;
    for (int i = 0; i < n; ++i) {
        const prl_object& o = w->objects[i];
        if (o.kind == 2) {
            d.gx[gi] = o.position[0] * kScale;
            d.gy[gi] = o.position[1] * kScale;
            float gw = fabsf(o.scale[0]) * kScale, gh = fabsf(o.scale[1]) * kScale;
            d.ghw[gi] = gw / 2.0f;
            d.ghh[gi] = gh / 2.0f;
            // glam Quat::from_rotation_z(rot) * Vec3::X / Vec3::Y, truncated (common.rs:183-184)
            float s = sinf(o.rotation * 0.5f), c = cosf(o.rotation * 0.5f);
            float k = c * c - ((0.0f * 0.0f + 0.0f * 0.0f) + s * s);
            float w2 = c * 2.0f;
            d.gxax[gi] = k + 0.0f + 0.0f;
            d.gxay[gi] = 0.0f + 0.0f + s * w2;
            d.gyax[gi] = 0.0f + 0.0f + (0.0f - s) * w2;
            d.gyay[gi] = k + 0.0f + 0.0f;
            ++gi;
            continue;
        }
        d.hx[ci] = 0.5f * fabsf(o.scale[0]) * kScale;
        d.hy[ci] = 0.5f * fabsf(o.scale[1]) * kScale;
        d.cx[ci] = o.position[0] * kScale;
        d.cy[ci] = o.position[1] * kScale;
        d.cc[ci] = libm_cosf(o.rotation);
        d.cs[ci] = libm_sinf(o.rotation);
        if (o.kind == 0) {
            d.body_of[ci] = -1;
            // parry Cuboid::aabb, loosened by prediction / 2 per side (rapier broad phase)
            float ac = fabsf(d.cc[ci]), as = fabsf(d.cs[ci]);
            float wx = ac * d.hx[ci] + as * d.hy[ci], wy = as * d.hx[ci] + ac * d.hy[ci];
            d.amnx[ci] = (d.cx[ci] - wx) - d.aabb_margin;
            d.amny[ci] = (d.cy[ci] - wy) - d.aabb_margin;
            d.amxx[ci] = (d.cx[ci] + wx) + d.aabb_margin;
            d.amxy[ci] = (d.cy[ci] + wy) + d.aabb_margin;
        } else {
            // parry MassProperties::from_cuboid (2-D), density 1
            float volume = d.hx[ci] * d.hy[ci] * 4.0f;
            float ix = (d.hx[ci] * d.hx[ci]) / 3.0f, iy = (d.hy[ci] * d.hy[ci]) / 3.0f;
            float mass = volume * 1.0f;
            float inertia = (ix + iy) * mass;
            d.inv_mass[bi] = inv0(mass);
            d.inv_inertia_sqrt[bi] = inv0(sqrtf(inertia));
            d.ccd_thickness[bi] = fminf(d.hx[ci], d.hy[ci]);
            d.collider_of[bi] = ci;
            d.body_of[ci] = (int8_t)bi;
            w->object_body[i] = bi;
            ++bi;
        }
        ++ci;
    }
    // gravity (0,-2) (common.rs:316): force = mass * gravity; velocity increment (force * inv_mass) * dt
    for (int b = 0; b < d.n_bodies; ++b) {
        float mass = inv0(d.inv_mass[b]);
        float fx = 0.0f + (0.0f * mass) * 1.0f, fy = 0.0f + (-2.0f * mass) * 1.0f;
        d.gdv_x[b] = 0.0f + (fx * d.inv_mass[b]) * dt;
        d.gdv_y[b] = 0.0f + (fy * d.inv_mass[b]) * dt;
    }
    // candidate collider pairs in ascending (collider1, collider2) order; two parent-less
    // colliders never collide
    int np = 0, npp = 0;
    for (int b = 0; b < MAXB; ++b) d.body_pairs[b] = 0;
    d.two_body_pairs = 0;
    for (int a = 0; a < d.n_colliders; ++a)
        for (int b = a + 1; b < d.n_colliders; ++b) {
            if (d.body_of[a] < 0 && d.body_of[b] < 0) continue;
            if (np >= PRL_MAX_PAIRS) { set_error("world needs more than PRL_MAX_PAIRS=%d candidate collider pairs", PRL_MAX_PAIRS); return PRL_ERR_CAPACITY; }
            d.pair_c1[np] = (uint8_t)a;
            d.pair_c2[np] = (uint8_t)b;
            d.pair_info[np] = (uint32_t)a | ((uint32_t)b << 8) | ((uint32_t)(d.body_of[a] & 0xff) << 16) | ((uint32_t)(d.body_of[b] & 0xff) << 24);
            if (np < 32) {
                if (d.body_of[a] >= 0 && d.body_of[b] >= 0) d.two_body_pairs |= 1u << np;
                if (d.body_of[a] >= 0) d.body_pairs[d.body_of[a]] |= 1u << np;
                if (d.body_of[b] >= 0) d.body_pairs[d.body_of[b]] |= 1u << np;
            }
            const int fixed = d.body_of[a] < 0 ? a : (d.body_of[b] < 0 ? b : -1);
            for (int q = 0; q < 4; ++q) d.pair_box[np][q] = d.pair_pose[np][q] = 0.0f;
            d.pair_he[np][0] = d.hx[a]; d.pair_he[np][1] = d.hy[a]; d.pair_he[np][2] = d.hx[b]; d.pair_he[np][3] = d.hy[b];
            if (fixed >= 0) {
                d.pair_pose[np][0] = d.cx[fixed]; d.pair_pose[np][1] = d.cy[fixed];
                d.pair_pose[np][2] = d.cc[fixed]; d.pair_pose[np][3] = d.cs[fixed];
                d.pair_box[np][0] = d.amnx[fixed]; d.pair_box[np][1] = d.amny[fixed];
                d.pair_box[np][2] = d.amxx[fixed]; d.pair_box[np][3] = d.amxy[fixed];
            }
            if (a == 0) ++npp;
            ++np;
        }
    d.n_pairs = np;
    d.n_player_pairs = npp;
    // x grid over the player's pairs with a fixed collider (used by single-body worlds)
    d.grid_n = 0;
    d.grid_x0 = 0.0f;
    d.grid_inv_w = 0.0f;
    for (int c = 0; c < DevWorld::kGrid; ++c) for (int q = 0; q < 4; ++q) d.grid_mask[c][q] = 0;
    if (d.n_bodies == 1 && np > 0) {
        float lo = 3.0e38f, hi = -3.0e38f;
        for (int p = 0; p < np; ++p) { lo = fminf(lo, d.pair_box[p][0]); hi = fmaxf(hi, d.pair_box[p][2]); }
        const int gn = DevWorld::kGrid;
        const float w = (hi - lo) / (float)gn;
        if (w > 0.0f && w < 3.0e38f) {
            d.grid_n = gn;
            d.grid_x0 = lo;
            d.grid_inv_w = 1.0f / w;
            for (int p = 0; p < np; ++p) {
                const int c0 = grid_cell(d, d.pair_box[p][0]), c1 = grid_cell(d, d.pair_box[p][2]);
                for (int c = c0; c <= c1; ++c) d.grid_mask[c][p >> 5] |= 1u << (p & 31);
            }
        }
    }
    {
        // manifold slots kept per environment in HBM: as many as the kernel variant holds live at once
        const int cap = (d.n_bodies > 1 && np > MAXC) ? MAXC_WIDE : MAXC;
        d.n_slots = np < cap ? np : cap;
    }
    return PRL_OK;
}

// ---------------------------------------------------------------------------- JSON
// serde_json::from_str::<World> (editor.rs:916-936) with #[derive(Deserialize)] on World,
// ObjectAndTransform and WorldObject (common.rs:19-49), restated:
//  * RFC 8259 grammar, whitespace = space \t \n \r, nothing but whitespace after the value,
//    nesting limit 128 (serde_json's recursion limit);
//  * a struct is a map or a sequence of its fields in declaration order; in a map unknown keys
//    are IGNORED (their value may be any JSON value: no deny_unknown_fields), a known key
//    twice is an error, a missing field is an error; keys are compared after unescaping;
//  * [f32; N] is a sequence of exactly N numbers; a number is parsed as f64 and narrowed with
//    `as f32` (integers are accepted; a literal that overflows f64 is an error);
//  * the externally tagged enum: "Goal" | {"Goal":null} | {"Block":{"fixed":b}} | {"Block":[b]};
//    a map tag has exactly one key.
namespace {
struct Cur {
    const char* p;
    const char* e;
    void ws() { while (p < e && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p; }
    bool peek(char c) { ws(); return p < e && *p == c; }
    bool eat(char c) { ws(); if (p < e && *p == c) { ++p; return true; } return false; }
    bool lit(const char* w) {
        size_t n = strlen(w);
        if ((size_t)(e - p) >= n && !strncmp(p, w, n)) { p += n; return true; }
        return false;
    }
    static void utf8(std::string* out, unsigned cp) {
        if (cp < 0x80) out->push_back((char)cp);
        else if (cp < 0x800) { out->push_back((char)(0xC0 | (cp >> 6))); out->push_back((char)(0x80 | (cp & 0x3F))); }
        else if (cp < 0x10000) { out->push_back((char)(0xE0 | (cp >> 12))); out->push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out->push_back((char)(0x80 | (cp & 0x3F))); }
        else { out->push_back((char)(0xF0 | (cp >> 18))); out->push_back((char)(0x80 | ((cp >> 12) & 0x3F))); out->push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out->push_back((char)(0x80 | (cp & 0x3F))); }
    }
    bool hex4(unsigned* v) {
        if (e - p < 4) return false;
        unsigned r = 0;
        for (int i = 0; i < 4; ++i) {
            char ch = p[i];
            unsigned d = ch >= '0' && ch <= '9' ? ch - '0' : ch >= 'a' && ch <= 'f' ? ch - 'a' + 10 : ch >= 'A' && ch <= 'F' ? ch - 'A' + 10 : 99;
            if (d == 99) return false;
            r = r * 16 + d;
        }
        p += 4;
        *v = r;
        return true;
    }
    // a JSON string, unescaped; the input must be valid UTF-8 (from_str takes a &str)
    bool str(std::string* out) {
        ws();
        if (p >= e || *p != '"') return false;
        ++p;
        out->clear();
        while (p < e) {
            unsigned char ch = (unsigned char)*p;
            if (ch == '"') { ++p; return true; }
            if (ch < 0x20) return false;  // control characters must be escaped
            if (ch == '\\') {
                ++p;
                if (p >= e) return false;
                char c2 = *p++;
                switch (c2) {
                    case '"': out->push_back('"'); break;
                    case '\\': out->push_back('\\'); break;
                    case '/': out->push_back('/'); break;
                    case 'b': out->push_back('\b'); break;
                    case 'f': out->push_back('\f'); break;
                    case 'n': out->push_back('\n'); break;
                    case 'r': out->push_back('\r'); break;
                    case 't': out->push_back('\t'); break;
                    case 'u': {
                        unsigned cp;
                        if (!hex4(&cp)) return false;
                        if (cp >= 0xDC00 && cp <= 0xDFFF) return false;  // lone trailing surrogate
                        if (cp >= 0xD800 && cp <= 0xDBFF) {
                            unsigned lo;
                            if (e - p < 2 || p[0] != '\\' || p[1] != 'u') return false;
                            p += 2;
                            if (!hex4(&lo) || lo < 0xDC00 || lo > 0xDFFF) return false;
                            cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                        }
                        utf8(out, cp);
                        break;
                    }
                    default: return false;
                }
                continue;
            }
            if (ch < 0x80) { out->push_back((char)ch); ++p; continue; }
            // multi-byte UTF-8 sequence: validate (no overlongs, no surrogates, <= U+10FFFF)
            int n = ch >= 0xF0 ? 4 : ch >= 0xE0 ? 3 : ch >= 0xC2 ? 2 : 0;
            if (n == 0 || ch > 0xF4 || e - p < n) return false;
            unsigned cp = ch & (0xFF >> (n + 1));
            for (int i = 1; i < n; ++i) {
                unsigned char cc = (unsigned char)p[i];
                if ((cc & 0xC0) != 0x80) return false;
                cp = (cp << 6) | (cc & 0x3F);
            }
            if ((n == 3 && cp < 0x800) || (n == 4 && cp < 0x10000) || cp > 0x10FFFF || (cp >= 0xD800 && cp <= 0xDFFF)) return false;
            out->append(p, (size_t)n);
            p += n;
        }
        return false;
    }
    // RFC 8259 number; serde_json: f64 parse, then `as f32`
    bool num(float* out) {
        ws();
        const char* s0 = p;
        if (p < e && *p == '-') ++p;
        if (p >= e) return false;
        if (*p == '0') ++p;
        else if (*p >= '1' && *p <= '9') { while (p < e && *p >= '0' && *p <= '9') ++p; }
        else return false;
        if (p < e && *p == '.') {
            ++p;
            if (p >= e || *p < '0' || *p > '9') return false;
            while (p < e && *p >= '0' && *p <= '9') ++p;
        }
        if (p < e && (*p == 'e' || *p == 'E')) {
            ++p;
            if (p < e && (*p == '+' || *p == '-')) ++p;
            if (p >= e || *p < '0' || *p > '9') return false;
            while (p < e && *p >= '0' && *p <= '9') ++p;
        }
        std::string tmp(s0, p);
        double d = strtod(tmp.c_str(), nullptr);
        if (std::isinf(d)) return false;  // serde_json: "number out of range"
        if (out) *out = (float)d;         // `as f32`: rounds to nearest, saturates to +-inf
        return true;
    }
    bool boolean(bool* out) {
        ws();
        if (lit("true")) { *out = true; return true; }
        if (lit("false")) { *out = false; return true; }
        return false;
    }
    // serde::de::IgnoredAny: any JSON value, validated and dropped
    bool skip(int depth) {
        ws();
        if (p >= e || depth > 128) return false;
        char ch = *p;
        if (ch == '"') { std::string t; return str(&t); }
        if (ch == 't') return lit("true");
        if (ch == 'f') return lit("false");
        if (ch == 'n') return lit("null");
        if (ch == '[') {
            ++p;
            if (eat(']')) return true;
            for (;;) {
                if (!skip(depth + 1)) return false;
                if (eat(',')) continue;
                return eat(']');
            }
        }
        if (ch == '{') {
            ++p;
            if (eat('}')) return true;
            for (;;) {
                std::string k;
                if (!str(&k) || !eat(':') || !skip(depth + 1)) return false;
                if (eat(',')) continue;
                return eat('}');
            }
        }
        return num(nullptr);
    }
    bool farray(float* out, int n) {  // [f32; n]
        if (!eat('[')) return false;
        for (int i = 0; i < n; ++i) {
            if (i && !eat(',')) return false;
            if (!num(&out[i])) return false;
        }
        return eat(']');
    }
};

// a struct as a map (unknown keys ignored, duplicates and missing fields are errors) or as a
// sequence of its fields in order. field(i) parses the value of field i.
template <class F>
bool parse_struct(Cur& c, const char* const* names, int n, F&& field) {
    if (c.eat('[')) {
        for (int i = 0; i < n; ++i) {
            if (i && !c.eat(',')) return false;
            if (!field(i)) return false;
        }
        return c.eat(']');
    }
    if (!c.eat('{')) return false;
    unsigned seen = 0;
    if (!c.eat('}')) {
        for (;;) {
            std::string k;
            if (!c.str(&k) || !c.eat(':')) return false;
            int f = -1;
            for (int i = 0; i < n; ++i) if (k == names[i]) f = i;
            if (f < 0) { if (!c.skip(1)) return false; }
            else {
                if (seen & (1u << f)) return false;  // duplicate field
                if (!field(f)) return false;
                seen |= 1u << f;
            }
            if (c.eat(',')) continue;
            if (c.eat('}')) break;
            return false;
        }
    }
    return seen == (1u << n) - 1u;  // missing field
}

// WorldObject (common.rs:46-49), externally tagged
bool parse_world_object(Cur& c, int* kind) {
    std::string tag;
    if (c.peek('"')) {  // a bare string names a unit variant
        if (!c.str(&tag) || tag != "Goal") return false;
        *kind = 2;
        return true;
    }
    if (!c.eat('{') || !c.str(&tag) || !c.eat(':')) return false;
    if (tag == "Goal") {
        c.ws();
        if (!c.lit("null")) return false;
        *kind = 2;
    } else if (tag == "Block") {
        bool fixed = false;
        static const char* const names[] = {"fixed"};
        if (!parse_struct(c, names, 1, [&](int) { return c.boolean(&fixed); })) return false;
        *kind = fixed ? 0 : 1;
    } else {
        return false;  // unknown variant
    }
    return c.eat('}');  // exactly one key
}

// ObjectAndTransform (common.rs:26-31)
bool parse_object(Cur& c, prl_object* o) {
    static const char* const names[] = {"object", "position", "scale", "rotation"};
    return parse_struct(c, names, 4, [&](int f) {
        switch (f) {
            case 0: return parse_world_object(c, &o->kind);
            case 1: return c.farray(o->position, 3);
            case 2: return c.farray(o->scale, 2);
            default: return c.num(&o->rotation);
        }
    });
}
}  // namespace

int parse_world_json(const char* s, size_t len, prl_world* out) {
    Cur c{s, s + len};
    static const char* const names[] = {"player_position", "objects"};
    bool ok = parse_struct(c, names, 2, [&](int f) {
        if (f == 0) return c.farray(out->player_position, 2);
        if (!c.eat('[')) return false;  // Vec<ObjectAndTransform>
        out->objects.clear();
        if (c.eat(']')) return true;
        for (;;) {
            prl_object o;
            memset(&o, 0, sizeof o);
            if (!parse_object(c, &o)) return false;
            out->objects.push_back(o);
            if (c.eat(',')) continue;
            return c.eat(']');
        }
    });
    c.ws();
    if (ok && c.p == c.e) return PRL_OK;
    set_error("World JSON does not match the serde schema of common.rs:19-49 near byte %ld", (long)(c.p - s));
    return PRL_ERR_PARSE;
}

// serde_json prints f32 with ryu (ryu::Buffer::format_finite -> pretty::format32): the shortest digits
// that round-trip, laid out by decimal exponent. With `length` digits d1 d2 ... and value
// digits * 10^k, kk = length + k:
//   0 <= k  and kk <= 13   ->  digits, k zeros, ".0"            (1234e7  -> 12340000000.0)
//   0 <  kk <= 13          ->  point inside the digits           (1234e-2 -> 12.34)
//   -6 < kk <= 0           ->  "0.", -kk zeros, digits           (1234e-7 -> 0.0001234)
//   otherwise              ->  d1[.d2...]e<kk-1>                 (1e30, 1.234e-7)
// Non-finite values are written as `null` (serde_json's serialize_f32).
static void put_f32(std::string& s, float v) {
    if (!std::isfinite(v)) { s += "null"; return; }
    if (std::signbit(v)) s += '-';
    if (v == 0.0f) { s += "0.0"; return; }
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, std::fabs(v), std::chars_format::scientific);  // shortest round-trip digits
    std::string t(buf, r.ptr);
    const size_t ep = t.find('e');
    std::string digits;
    for (char ch : t.substr(0, ep)) if (ch != '.') digits += ch;
    const int exp10 = atoi(t.c_str() + ep + 1);  // of the first digit
    const int length = (int)digits.size();
    const int kk = exp10 + 1, k = kk - length;
    if (0 <= k && kk <= 13) {
        s += digits;
        s.append((size_t)k, '0');
        s += ".0";
    } else if (0 < kk && kk <= 13) {
        s += digits.substr(0, (size_t)kk);
        s += '.';
        s += digits.substr((size_t)kk);
    } else if (-6 < kk && kk <= 0) {
        s += "0.";
        s.append((size_t)(-kk), '0');
        s += digits;
    } else {
        s += digits[0];
        if (length > 1) { s += '.'; s += digits.substr(1); }
        s += 'e';
        s += std::to_string(kk - 1);
    }
}

std::string world_to_json(const prl_world* w) {
    std::string s = "{\"player_position\":[";
    put_f32(s, w->player_position[0]);
    s += ",";
    put_f32(s, w->player_position[1]);
    s += "],\"objects\":[";
    for (size_t i = 0; i < w->objects.size(); ++i) {
        const prl_object& o = w->objects[i];
        if (i) s += ",";
        s += "{\"object\":";
        if (o.kind == 2) s += "\"Goal\"";
        else s += o.kind == 0 ? "{\"Block\":{\"fixed\":true}}" : "{\"Block\":{\"fixed\":false}}";
        s += ",\"position\":[";
        put_f32(s, o.position[0]); s += ","; put_f32(s, o.position[1]); s += ","; put_f32(s, o.position[2]);
        s += "],\"scale\":[";
        put_f32(s, o.scale[0]); s += ","; put_f32(s, o.scale[1]);
        s += "],\"rotation\":";
        put_f32(s, o.rotation);
        s += "}";
    }
    s += "]}";
    return s;
}

}  // namespace prl
