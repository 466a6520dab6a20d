// prl_world.h -- the per-world static tables shared by the host builder and the kernels.
//
// One `DevWorld` is the physics-space image of `Environment::from_world(&World)`
// (reference src/common.rs:70-172): the player capsule, every block's cuboid collider, the
// candidate collider pairs, the goals and the constants of
// `IntegrationParameters::default()` (common.rs:89). It is identical for every environment of a
// batch, lives once in HBM and is staged into shared memory by every thread block.
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/prl.h"

namespace prl {

constexpr int MAXCOL = PRL_MAX_BLOCKS + 1;   // colliders: player + blocks
constexpr int MAXB = PRL_MAX_DYNAMIC + 1;    // bodies: player + pushable blocks
constexpr int MIDB = 5;                      // kernels keep the bodies in registers: variants for 1, 2, <= MIDB and <= MAXB bodies
constexpr int MAXP = PRL_MAX_PAIRS;
constexpr int MAXG = PRL_MAX_GOALS;
constexpr int MAXC = PRL_MAX_CONTACTS;       // manifold slots per environment
constexpr int MAXC_WIDE = 2 * MAXC;          // ... in worlds with pushable blocks and more than MAXC candidate pairs
constexpr int MF_WORDS = 15;                 // HBM words per manifold slot (see prl_device.cuh)

// MC / MP / MG: collider, pair and goal capacity. The host always builds the full-size DevWorld;
// kernels of small worlds stage a compact copy (DevWorldSmall) so that shared memory goes to the
// solver's row pool instead of empty table slots.
template <int MC, int MP, int MG>
struct alignas(16) DevWorldT {
    static constexpr int kMaxColliders = MC, kMaxPairs = MP, kMaxGoals = MG;
    int32_t n_colliders, n_bodies, n_pairs, n_goals;
    int32_t n_player_pairs;  // pairs whose collider1 is the player capsule come first
    int32_t n_slots;         // manifold slots kept per environment in HBM = min(n_pairs, MAXC)
    uint32_t two_body_pairs; // bit p = both colliders of candidate pair p belong to bodies (first 32 pairs)
    // player capsule (common.rs:74-86) and glue constants (common.rs:216-220, 234-236)
    float radius, half_height, lower_off;
    // integration parameters (rapier IntegrationParameters::default, derived values)
    float dt, inv_dt, erp_inv_dt, cfm_factor, prediction, allowed_linear_error, aabb_margin;
    float scale;  // BEVY_TO_PHYSICS_SCALE (distance_to_goals divides by it, common.rs:192)
    // the win test of common.rs:335 on the SQUARED physics-space distance q: sqrt(q) / scale < 1e-7
    // iff q <= win_q (sqrt and the division are monotone; the host finds the largest such q bit by bit)
    float win_q;
    // colliders in insertion order; 0 = player capsule (common.rs:86, 125, 140-144)
    float cx[MC], cy[MC], cc[MC], cs[MC];  // pose (initial pose for pushable blocks)
    float hx[MC], hy[MC];                  // cuboid half extents
    float amnx[MC], amny[MC], amxx[MC], amxy[MC];  // loosened AABB of fixed colliders
    int8_t body_of[MC];                    // -1: parent-less (fixed) collider
    // bodies in handle order; 0 = player
    float inv_mass[MAXB], inv_inertia_sqrt[MAXB], ccd_thickness[MAXB];
    float gdv_x[MAXB], gdv_y[MAXB];        // 0 + (force * inv_mass) * dt with force = mass * gravity
    int32_t collider_of[MAXB];
    uint32_t body_pairs[MAXB];             // bit p = candidate pair p involves this body (first 32 pairs)
    // candidate pairs, ascending (collider1, collider2), fixed-fixed excluded
    uint8_t pair_c1[MP], pair_c2[MP];
    // the same per pair in one word, so the per-step walks over pairs need no dependent table
    // look-ups: collider1 | collider2 << 8 | body1 << 16 | body2 << 24 (body 0xff: fixed collider),
    // and the loosened AABB (min x, min y, max x, max y) of the pair's fixed collider, if any
    // pair_pose: pose (x, y, cos, sin) of that fixed collider; pair_he: half extents of collider1
    // and collider2 (hx1, hy1, hx2, hy2)
    uint32_t pair_info[MP];
    alignas(16) float pair_box[MP][4];
    alignas(16) float pair_pose[MP][4];
    alignas(16) float pair_he[MP][4];
    // x grid over the fixed colliders of the player's pairs (worlds too large for slot == pair
    // manifolds): grid_mask[c] = pairs whose fixed collider's loosened AABB reaches cell c, cells
    // c = clamp(floor((x - grid_x0) * grid_inv_w), 0, grid_n - 1), the end cells open-ended. The
    // broad phase of a single-body world then tests a handful of pairs instead of all of them.
    static constexpr int kGrid = MP > 32 ? 64 : 1;
    int32_t grid_n;
    float grid_x0, grid_inv_w;
    uint32_t grid_mask[kGrid][4];
    // goals (common.rs:148-157, 175-195)
    float gx[MG], gy[MG], ghw[MG], ghh[MG];        // centre, width/2, height/2
    float gxax[MG], gxay[MG], gyax[MG], gyay[MG];  // Quat::from_rotation_z(rot) * X / Y, truncated
};
using DevWorld = DevWorldT<MAXCOL, MAXP, MAXG>;
using DevWorldSmall = DevWorldT<16, MAXC, 4>;  // <= MAXC pairs: manifold slot == pair index (Stepper::kDirect)

// cell of x in the world's x grid; monotone in x, identical on host and device (plain f32 ops)
template <class WT>
#if defined(__CUDACC__)
__host__ __device__
#endif
inline int grid_cell(const WT& w, float x) {
    float f = floorf((x - w.grid_x0) * w.grid_inv_w);
    if (!(f >= 0.0f)) return 0;  // left of the grid (or NaN)
    return f > (float)(w.grid_n - 1) ? w.grid_n - 1 : (int)f;
}

template <class WT>
inline bool world_fits(const DevWorld& w) {
    return w.n_colliders <= WT::kMaxColliders && w.n_pairs <= WT::kMaxPairs && w.n_goals <= WT::kMaxGoals;
}

// ---- HBM state layout of a batch: structure of arrays, one 32-bit word per field per environment:
// word(f, env) = state[f * n_envs + env]. Fields:
//   F_FLAGS                          bits 0-7 sleeping mask (bit b = body b), bit 8 won, bit 9 fresh
//                                    (no pipeline step taken yet), bit 10 manifold-slot overflow,
//                                    bits 16-23 number of live manifold slots
//   F_SCORE, F_WIN_STEP, F_STEPS     running min of distance_to_goals, 1-based step of the first
//                                    won() (-1: none), steps since reset
//   F_ROLL                           position of a rollout in its move schedule (only meaningful
//                                    inside rollout_kernel, which parks environments here when it
//                                    regroups them between threads)
//   F_PLAYER .. +4                   x, y, vx, vy, sleep timer (rotations locked: cos=1, sin=0, w=0)
//   F_DYN + 8*(b-1) .. +7            x, y, cos, sin, vx, vy, angvel, sleep timer of pushable block b
//   slots                            n_slots x MF_WORDS: word 0 = pair | npts << 8, then local_n1,
//                                    local_n2, 2 x (local_p1, local_p2, dist)
enum { F_FLAGS = 0, F_SCORE = 1, F_WIN_STEP = 2, F_STEPS = 3, F_ROLL = 4, F_PLAYER = 5, F_DYN = 10 };
constexpr uint32_t FLAG_WON = 1u << 8, FLAG_FRESH = 1u << 9, FLAG_OVERFLOW = 1u << 10;

inline int state_slot_base(int n_bodies) { return F_DYN + 8 * (n_bodies - 1); }
inline int state_words(int n_bodies, int n_slots) { return state_slot_base(n_bodies) + MF_WORDS * n_slots; }

}  // namespace prl
