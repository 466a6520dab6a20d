/*
 * prl.h -- C ABI of the B200 batched physics environment.
 *
 * This is the drop-in boundary for the reference's one data-parallel hot path: stepping many
 * independent `Environment`s through `Environment::step(Move)` and the genetic algorithm that
 * drives them. The reference (a Rust crate) has no FFI of its own; each entry point below names
 * the reference interface it replaces (file:line into the reference repo). A Rust `extern "C"`
 * block binding these symbols is shown in INTEGRATION.md and rust/src/ffi.rs.
 *
 * Conventions
 *  - every function returns PRL_OK (0) or a negative prl_status; prl_last_error() returns a
 *    thread-local, NUL-terminated description of the last failure on the calling thread;
 *  - handles are opaque, caller-owned and freed with the matching *_destroy;
 *  - a batch / GA handle is bound to one CUDA device and one stream; it may be used from any
 *    host thread (Algorithm::train runs off the main thread, train.rs:80) but by one thread at
 *    a time;
 *  - buffer arguments documented "host or device" are classified with
 *    cudaPointerGetAttributes; host buffers are staged through the handle's stream and the
 *    call returns after the data has arrived (device buffers: the call is asynchronous on the
 *    handle's stream);
 *  - a Move is one byte: bit0 = left, bit1 = right, bit2 = up (common.rs:351-356);
 *  - there is NO CPU fallback: every compute entry point fails with PRL_ERR_CUDA when no
 *    sm_100 device is usable.
 */
#ifndef PRL_H
#define PRL_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PRL_ABI_VERSION 2

typedef enum {
    PRL_OK = 0,
    PRL_ERR_INVALID = -1,   /* bad argument */
    PRL_ERR_PARSE = -2,     /* World JSON does not match the serde schema */
    PRL_ERR_CAPACITY = -3,  /* world exceeds the compiled limits (see prl_limits) */
    PRL_ERR_NO_GOAL = -4,   /* distance_to_goals() on a goal-less world (reference: None / unwrap panic) */
    PRL_ERR_CUDA = -5,      /* CUDA runtime failure, or no usable device */
    PRL_ERR_OVERFLOW = -6   /* an environment needed more simultaneous contact manifolds than the kernels hold (PRL_MAX_CONTACTS) */
} prl_status;

#define PRL_MOVE_LEFT 1u
#define PRL_MOVE_RIGHT 2u
#define PRL_MOVE_UP 4u

/* compiled limits */
#define PRL_MAX_DYNAMIC 7     /* pushable blocks per world (the per-environment sleeping mask has 8 bits: player + 7) */
#define PRL_MAX_BLOCKS 128    /* fixed + pushable blocks per world */
#define PRL_MAX_PAIRS 128     /* candidate collider pairs per world */
#define PRL_MAX_GOALS 16
#define PRL_MAX_CONTACTS 16   /* simultaneous contact manifolds per environment in worlds with more candidate pairs
                                 than that (smaller worlds have a slot per pair and no limit); twice as many,
                                 32, when such a world has pushable blocks */

/* One `ObjectAndTransform` (common.rs:26-31). kind: 0 = WorldObject::Block{fixed:true},
 * 1 = Block{fixed:false}, 2 = Goal (common.rs:46-49). */
typedef struct {
    float position[3];
    float scale[2];
    float rotation;
    int32_t kind;
} prl_object;

typedef struct prl_world prl_world;
typedef struct prl_batch prl_batch;
typedef struct prl_ga prl_ga;

const char* prl_last_error(void);
int32_t prl_abi_version(void);
/* Debug builds only (libprl_debug.so, -DPRL_BOUNDS_CHECK: every run-time index of the kernels is
 * checked): bit mask of the checks that failed on the current device since the library was loaded,
 * 0 = none; -1 = this library was built without the checks (the product); -2 = CUDA failure. */
int32_t prl_debug_bounds_violations(void);
/* number of CUDA devices visible to the library (0 when none / no driver) */
int32_t prl_device_count(void);

/* ------------------------------------------------------------------ World (common.rs:19-49) */
/* `World { player_position, objects }` */
int prl_world_create(const float player_position[2], const prl_object* objs, int32_t n_objects,
                     prl_world** out);
/* serde_json::from_str::<World> (editor.rs:916-936): externally tagged enum
 * {"Block":{"fixed":bool}} | "Goal" */
int prl_world_from_json(const char* utf8, size_t len, prl_world** out);
/* serde_json::to_string(&World) (editor.rs:938-962). Writes at most cap bytes (incl. NUL) and
 * always sets *needed to the size required. */
int prl_world_to_json(const prl_world*, char* buf, size_t cap, size_t* needed);
int32_t prl_world_num_objects(const prl_world*);
int prl_world_get_object(const prl_world*, int32_t index, prl_object* out);
int prl_world_player_position(const prl_world*, float out[2]);
/* scene sizes after Environment::from_world (common.rs:162-172): bodies = 1 + pushable blocks,
 * colliders = 1 + blocks */
int prl_world_counts(const prl_world*, int32_t* n_bodies, int32_t* n_colliders, int32_t* n_goals,
                     int32_t* n_pairs);
/* from_world's Vec<Option<RigidBodyHandle>> (common.rs:162-172): body index of object `index`
 * (>= 1), or -1 for None */
int32_t prl_world_object_body(const prl_world*, int32_t index);
void prl_world_destroy(prl_world*);

/* ------------------------------------------------------------------ BatchedEnvironment */
/* n_envs x Environment::from_world(world) (common.rs:162-172) on `device`. `stream` is a
 * cudaStream_t (NULL = the legacy default stream). */
int prl_batch_create(const prl_world*, int64_t n_envs, int32_t device, void* stream, prl_batch** out);
void prl_batch_destroy(prl_batch*);
int64_t prl_batch_num_envs(const prl_batch*);
int32_t prl_batch_num_bodies(const prl_batch*);
/* every environment back to its from_world state */
int prl_batch_reset(prl_batch*);
/* Environment::step (common.rs:215-340) for every environment; moves: n_envs bytes, host or device */
int prl_batch_step(prl_batch*, const uint8_t* moves);
/* the same Move for every environment, n_steps times (README.md:105-114 loop body) */
int prl_batch_step_uniform(prl_batch*, uint32_t move, int32_t n_steps);
/* Environment::distance_to_goals (common.rs:175-195), editor pixels; out: n_envs floats, host or device */
int prl_batch_distance_to_goals(prl_batch*, float* out);
/* Environment::won (common.rs:197-199); out: n_envs bytes (0/1), host or device */
int prl_batch_won(prl_batch*, uint8_t* out);
/* Observation, the batched form of rigid_body_set()/player_handle() (common.rs:201-211) as the
 * viewers use them (train.rs:146-157, game.rs:134-145). Bodies in from_world handle order
 * (0 = player). out: n x n_bodies x 9 floats, host only:
 * x, y, cos, sin, vx, vy, angvel, sleep_timer, sleeping(0/1) */
int prl_batch_get_state(prl_batch*, int64_t env0, int64_t n, float* out);
/* GA fitness rollout (main.rs:44-70) for every environment, from the from_world state:
 * genomes = n_envs x genome_len Moves (row per agent); the tail of
 * number_of_steps % repeat_move default Moves is run as the reference does. early_exit = 1
 * reproduces break-on-win; 0 always runs every step (throughput runs).
 * out_score: min distance_to_goals over the rollout; out_step_of_win: 1-based step at which
 * won() first became true, or -1. All buffers host or device; outputs may be NULL. */
int prl_batch_rollout(prl_batch*, const uint8_t* genomes, int32_t genome_len, int32_t repeat_move,
                      int32_t number_of_steps, int32_t early_exit, float* out_score,
                      int32_t* out_step_of_win);
/* Per environment since the last reset / rollout start: running minimum of distance_to_goals
 * (the GA's score, main.rs:50), 1-based step at which won() first became true (-1: never) and
 * the number of steps taken. Host or device buffers; any may be NULL. Synchronises. */
int prl_batch_rollout_info(prl_batch*, float* out_score, int32_t* out_step_of_win, int32_t* out_steps);
/* ---- Observation-driven agents: Agent::get_move(&mut self, &Environment) (algorithm.rs:8-11) in batch.
 * Observation of one environment, D = prl_batch_observation_dim = 9 * n_bodies + 4 floats: per body
 * (from_world handle order) x, y, cos, sin, vx, vy, angvel, sleep_timer, sleeping (what
 * rigid_body_set() exposes, common.rs:201-211), then the ground-contact scan Environment::step makes
 * from narrow_phase / collider_set() (common.rs:223-244): on_ground (0/1), number of floor contacts,
 * 1 if the player has an active contact with a pushable block; then distance_to_goals() in pixels
 * (+inf without goals). Contacts are those of the end of the previous step (none before the first). */
int32_t prl_batch_observation_dim(const prl_batch*);
/* out: n_envs x D floats, host or device */
int prl_batch_get_observation(prl_batch*, float* out);
/* A per-agent two-layer perceptron evaluated INSIDE the fused rollout, every step, in place of the
 * genome look-up of prl_batch_rollout: h = max(W1 obs + b1, 0) (hidden <= 32 units), out = W2 h + b2,
 * Move = (out[0] > 0, out[1] > 0, out[2] > 0) = (left, right, up). params: P x n_envs floats,
 * parameter-major (params[p * n_envs + env]), P = prl_policy_num_params = hidden * (D + 1) +
 * 3 * (hidden + 1): W1 row j at p = j * (D + 1) + i with its bias at i = D, then W2 row k at
 * hidden * (D + 1) + k * (hidden + 1) + j with its bias at j = hidden. From the from_world state;
 * scores / win steps as prl_batch_rollout (main.rs:44-70 with repeat_move = 1). Host or device buffers. */
int32_t prl_policy_num_params(const prl_batch*, int32_t hidden);
int prl_batch_rollout_policy(prl_batch*, const float* params, int32_t hidden, int32_t number_of_steps,
                             int32_t early_exit, float* out_score, int32_t* out_step_of_win);
/* Replay for the viewers (train.rs:132-159: `update_visualization` steps a fresh Environment with
 * `agent.get_move()` once per frame and copies every body's transform to its sprite): n fresh
 * from_world environments play one GeneticAgent each (main.rs:276-284: genes repeated
 * repeat_move times, then Move::default() forever; no early exit) for number_of_steps steps in ONE
 * launch, and every body is recorded after every step. Independent of the batch's own
 * environments (only its world, device and stream are used).
 * genomes: n x genome_len Moves. out_state: number_of_steps x n x n_bodies x 9 floats (layout of
 * prl_batch_get_state); out_distance: number_of_steps x n floats (distance_to_goals, +inf without
 * goals); out_won: number_of_steps x n bytes. Host or device buffers; any output may be NULL. */
int prl_batch_trajectory(prl_batch*, int64_t n, const uint8_t* genomes, int32_t genome_len, int32_t repeat_move,
                         int32_t number_of_steps, float* out_state, float* out_distance, uint8_t* out_won);
/* wait for the handle's stream; returns PRL_ERR_OVERFLOW if any environment needed more than
 * PRL_MAX_CONTACTS simultaneous contact manifolds since the last reset */
int prl_batch_sync(prl_batch*);
/* kernels launched by this handle so far */
int64_t prl_batch_launch_count(const prl_batch*);
/* bytes of HBM holding the batch's state arrays */
int64_t prl_batch_state_bytes(const prl_batch*);

/* ------------------------------------------------------------------ Genetic algorithm (main.rs:18-163) */
typedef struct {
    int32_t number_of_steps;   /* main.rs:22, default 1000 */
    int64_t number_of_agents;  /* main.rs:23, default 1000; GLOBAL population over all ranks */
    int32_t repeat_move;       /* main.rs:24, default 20 */
    float mutation_rate;       /* main.rs:25, default 0.1 */
    int32_t keep_best;         /* main.rs:26, default false */
    uint64_t seed;             /* the reference uses thread_rng (main.rs:42); here a Philox key */
} prl_ga_config;

/* Three ways to run the GA (csrc/prl_ga.cu):
 *  - prl_ga_create(rank 0 of 1): the whole population on one device;
 *  - prl_ga_create(rank r of R), one process per GPU: this handle owns the slice
 *    [r*N/R, (r+1)*N/R); the HOST exchanges the generation between prl_ga_best calls
 *    (ncclAllGather of the local scores and genomes into the global buffers, prl_ga_buffers);
 *  - prl_ga_create_multi(devices[]): one process, one slice and one stream per device; the
 *    library exchanges the generation itself over peer memory (NVLink): scores are pushed into
 *    every peer's score vector, parent genomes are read from the owning device by the breeding
 *    kernel. Needs peer access between all listed devices. A device may be listed more than once
 *    (several slices on one GPU).
 * The result -- every genome and score of every generation -- is the same for all three. */
int prl_ga_create(const prl_world*, const prl_ga_config*, int32_t device, void* stream,
                  int32_t rank, int32_t n_ranks, prl_ga** out);
int prl_ga_create_multi(const prl_world*, const prl_ga_config*, const int32_t* devices, int32_t n_devices,
                        prl_ga** out);
void prl_ga_destroy(prl_ga*);
/* slices held by this handle: n_devices for prl_ga_create_multi, else 1 */
int32_t prl_ga_num_shards(const prl_ga*);
/* main.rs:72-84: random genomes for the handle's slice(s) and their scores */
int prl_ga_init(prl_ga*);
/* Device buffers of a prl_ga_create handle, for the per-generation exchange by the host when
 * n_ranks > 1: local_*: this rank's slice; global_*: all agents. With n_ranks == 1 local_* and
 * global_* are the same memory, and the genome pointer is only valid until the next
 * prl_ga_next_generation (parents and children are two buffers that swap). Not for
 * prl_ga_create_multi handles (PRL_ERR_INVALID): use prl_ga_read_generation. */
int prl_ga_buffers(prl_ga*, float** local_scores, uint8_t** local_genomes, float** global_scores,
                   uint8_t** global_genomes, int64_t* local_count, int64_t* local_offset,
                   int32_t* genome_len);
/* main.rs:87-124: (best score, best genome) of the current global generation; first minimum
 * wins ties. best_genome: genome_len bytes, host or device; outputs may be NULL. With
 * n_ranks > 1 the host must have all-gathered the generation (prl_ga_buffers). Synchronises.
 * Returns PRL_ERR_OVERFLOW if an agent's environment ran out of contact-manifold slots while it
 * was scored (its score is not exact). */
int prl_ga_best(prl_ga*, float* best_score, uint8_t* best_genome, int64_t* best_index);
/* main.rs:126-160 for the handle's slice(s): elitism, weighted 2-parent choice, crossover,
 * mutation, then scores the children. */
int prl_ga_next_generation(prl_ga*);
/* One iteration of the loop at main.rs:86-162 for single-process handles (prl_ga_create with
 * n_ranks == 1, prl_ga_create_multi): reports the best agent of the current generation (the first
 * call initialises the population, main.rs:72-84), then breeds and scores the next one. */
int prl_ga_step_generation(prl_ga*, float* best_score, uint8_t* best_genome, int64_t* best_index);
/* the whole current generation to the host (single-process handles): scores: number_of_agents
 * floats, genomes: number_of_agents x genome_len Moves; either may be NULL */
int prl_ga_read_generation(prl_ga*, float* scores, uint8_t* genomes);
int64_t prl_ga_generation(const prl_ga*);
int64_t prl_ga_launch_count(const prl_ga*);
/* environment steps executed by the most recent scoring pass of the handle's slice(s) (early exit
 * on win makes this less than n_local * number_of_steps); -1 on CUDA failure */
int64_t prl_ga_env_steps(prl_ga*);

#ifdef __cplusplus
}
#endif
#endif /* PRL_H */
