/*
 * prl_oracle.h -- CPU ORACLE (test infrastructure, NOT a product path).
 *
 * Scalar-f32 restatement of the reference's `Environment::step` hot path:
 *   /root/reference/src/common.rs:70-172   (Environment::new / add_object / from_world)
 *   /root/reference/src/common.rs:175-199  (distance_to_goals / won)
 *   /root/reference/src/common.rs:215-340  (step)
 *   /root/reference/src/main.rs:44-70      (GA agent_score rollout)
 * and of the third-party arithmetic `step` calls into, which is NOT vendored in
 * /root/reference: rapier2d 0.17.2 (feature enhanced-determinism), parry2d 0.13.4,
 * nalgebra 0.32.2, simba 0.8.1 + libm 0.2.7, glam 0.23.0 (Cargo.toml:19, Cargo.lock).
 * Their published algorithms are restated from SURVEY.md Appendix A; each function in
 * prl_oracle.c names the crate routine it follows.
 *
 * PARITY UNPINNED: the reference holds no tests, golden vectors or fixtures for this path
 *  and neither cargo nor the rapier sources exist in this image, so
 * the oracle cannot be checked against the real crate. It is pinned only by analytic
 * known-answer tests (tests/test_oracle_known_answers.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library. Build: `make -C oracle` (gcc -O2 -ffp-contract=off: rustc never
 * contracts a*b+c into an FMA, so neither does the oracle).
 */
#ifndef PRL_ORACLE_H
#define PRL_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same POD as include/prl.h's prl_object (kept separate on purpose: the oracle shares no
 * code with the product). kind: 0 = Block{fixed:true}, 1 = Block{fixed:false}, 2 = Goal. */
typedef struct {
    float position[3];
    float scale[2];
    float rotation;
    int32_t kind;
} orc_object;

typedef struct orc_env orc_env;

/* Sensitivity studies (tests/test_oracle_sensitivity.py): flip one of the oracle's documented
 * deviations from the crates. 0 = the canonical oracle. The default applies to environments created
 * afterwards (orc_rollout_batch included); not thread safe, tests only. */
enum {
    ORC_VARIANT_REVERSE_ORDER = 1,        /* D2: Gauss-Seidel / floor-contact scan in descending pair order */
    ORC_VARIANT_CREATION_ORDER = 2,       /* D2: contact-graph edge order (pair creation order, swap-remove on removal) */
    ORC_VARIANT_GJK = 4,                  /* D1: capsule-cuboid closest points by GJK with parry's tolerance */
    ORC_VARIANT_WAKE_ON_ZERO_IMPULSE = 8  /* D4: apply_impulse(0, wake_up = true) wakes the body */
};
void orc_set_default_variant(int32_t flags);
void orc_env_set_variant(orc_env*, int32_t flags);

/* Environment::from_world (common.rs:162-172). Returns NULL on allocation failure. */
orc_env* orc_env_create(const float player_position[2], const orc_object* objs, int32_t n);
void orc_env_destroy(orc_env*);
/* back to the from_world state */
void orc_env_reset(orc_env*);

/* Environment::step (common.rs:215-340). move: bit0=left bit1=right bit2=up. */
void orc_env_step(orc_env*, uint32_t move);
/* Environment::distance_to_goals (common.rs:175-195); returns 0 and writes *out, or -1 if the
 * world has no goal (the reference returns None). Editor pixels. */
int orc_env_distance_to_goals(const orc_env*, float* out);
/* Environment::won (common.rs:197-199) */
int orc_env_won(const orc_env*);

/* Observation (common.rs:201-211 accessors): bodies in from_world handle order
 * (0 = player, then dynamic blocks in object order). */
int32_t orc_env_num_bodies(const orc_env*);
/* out9 = x, y, cos, sin, vx, vy, angvel, sleep_timer, sleeping(0/1) */
void orc_env_body_state(const orc_env*, int32_t body, float out9[9]);
/* overwrite a body's dynamic state (teacher forcing in tests) */
void orc_env_set_body_state(orc_env*, int32_t body, const float in9[9]);
/* debug: number of candidate pairs and the cached manifold of pair p:
 * out16 = npts, n1x,n1y,n2x,n2y, (p1x,p1y,p2x,p2y,dist) x 2, pad */
int32_t orc_env_num_pairs(const orc_env*);
void orc_env_manifold(const orc_env*, int32_t pair, float out16[16]);
void orc_env_pair(const orc_env*, int32_t pair, int32_t out2[2]);

/* GA fitness rollout, main.rs:44-70, on a fresh env. genome: genome_len bytes (bit0..2).
 * early_exit=1 reproduces the reference's break-on-win; 0 always runs number_of_steps steps
 * (throughput benches). Returns min distance; *step_of_win = 1-based step at which won() first
 * became true or -1; *steps_run = steps executed. */
float orc_rollout(orc_env*, const uint8_t* genome, int32_t genome_len, int32_t repeat_move,
                  int32_t number_of_steps, int32_t early_exit, int32_t* step_of_win,
                  int32_t* steps_run);

/* Many rollouts of one world on n_threads host threads (static partition), for the CPU
 * baseline. genomes: n_envs x genome_len. Returns total env-steps executed. */
int64_t orc_rollout_batch(const float player_position[2], const orc_object* objs, int32_t n,
                          const uint8_t* genomes, int64_t n_envs, int32_t genome_len,
                          int32_t repeat_move, int32_t number_of_steps, int32_t early_exit,
                          int32_t n_threads, float* out_score, int32_t* out_step_of_win);

/* Observation-driven agents (ref: algorithm.rs:8-11 Agent::get_move(&mut self, &Environment)):
 * observation = 9 floats per body (orc_env_body_state) + on_ground, floor-contact count, touching a
 * pushable block (the scan of common.rs:223-244) + distance_to_goals (px); policy = two-layer
 * perceptron, parameter p of an agent at theta[p * stride] (layout of include/prl.h
 * prl_batch_rollout_policy). */
int32_t orc_env_observation_dim(const orc_env*);
void orc_env_observation(const orc_env*, float* obs);
uint32_t orc_policy_move(const float* theta, int64_t stride, int32_t hidden, int32_t obs_dim, const float* obs);
float orc_rollout_policy(orc_env*, const float* theta, int64_t stride, int32_t hidden, int32_t number_of_steps,
                         int32_t early_exit, int32_t* step_of_win, int32_t* steps_run);
int64_t orc_rollout_policy_batch(const float player_position[2], const orc_object* objs, int32_t n, const float* theta,
                                 int64_t n_envs, int32_t hidden, int32_t number_of_steps, int32_t early_exit,
                                 int32_t n_threads, float* out_score, int32_t* out_step_of_win);

#ifdef __cplusplus
}
#endif
#endif
