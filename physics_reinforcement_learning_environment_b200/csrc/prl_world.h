// prl_world.h -- the per-world static tables shared by the host builder and the kernels.
//
// One `DevWorld` is the physics-space image of `Environment::from_world(&World)`
// (reference src/common.rs:70-172): the player capsule, every block's cuboid collider, the
// candidate collider pairs, the goals and the constants of
// `IntegrationParameters::default()` (common.rs:89). It is identical for every environment of a
// batch, lives once in HBM and is staged into shared memory by every thread block.
#pragma once
#include <stdint.h>

#include "../../include/prl.h"

namespace prl {

constexpr int MAXCOL = PRL_MAX_BLOCKS + 1;   // colliders: player + blocks
constexpr int MAXB = PRL_MAX_DYNAMIC + 1;    // bodies: player + pushable blocks
constexpr int MAXP = PRL_MAX_PAIRS;
constexpr int MAXG = PRL_MAX_GOALS;
constexpr int MAXC = PRL_MAX_CONTACTS;
constexpr int MF_WORDS = 14;                 // manifold cache words per pair (see prl_device.cuh)

struct DevWorld {
    int32_t n_colliders, n_bodies, n_pairs, n_goals;
    // player capsule (common.rs:74-86) and glue constants (common.rs:216-220, 234-236)
    float radius, half_height, lower_off;
    // integration parameters (rapier IntegrationParameters::default, derived values)
    float dt, inv_dt, erp_inv_dt, cfm_factor, prediction, allowed_linear_error, aabb_margin;
    float inv_scale_div;  // BEVY_TO_PHYSICS_SCALE (distance_to_goals divides by it, common.rs:193)
    // colliders in insertion order; 0 = player capsule (common.rs:86, 125, 140-144)
    float cx[MAXCOL], cy[MAXCOL], cc[MAXCOL], cs[MAXCOL];  // pose (initial pose for pushable blocks)
    float hx[MAXCOL], hy[MAXCOL];                          // cuboid half extents
    float amnx[MAXCOL], amny[MAXCOL], amxx[MAXCOL], amxy[MAXCOL];  // loosened AABB of fixed colliders
    int8_t body_of[MAXCOL];                                // -1: parent-less (fixed) collider
    // bodies in handle order; 0 = player
    float inv_mass[MAXB], inv_inertia_sqrt[MAXB], ccd_thickness[MAXB];
    float gdv_x[MAXB], gdv_y[MAXB];                        // (force * inv_mass) * dt with force = mass * gravity
    int32_t collider_of[MAXB];
    // candidate pairs, ascending (collider1, collider2), fixed-fixed excluded
    uint8_t pair_c1[MAXP], pair_c2[MAXP];
    // goals (common.rs:148-157, 175-195)
    float gx[MAXG], gy[MAXG], ghw[MAXG], ghh[MAXG];        // centre, width/2, height/2
    float gxax[MAXG], gxay[MAXG], gyax[MAXG], gyay[MAXG];  // Quat::from_rotation_z(rot) * X / Y, truncated
};

}  // namespace prl
