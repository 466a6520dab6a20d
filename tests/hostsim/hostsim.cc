// hostsim.cc -- TEST INFRASTRUCTURE, never shipped: compiles the product's per-environment device
// logic (csrc/prl_device.cuh) with g++ so the `-m "not gpu"` tests can run it, one environment at
// a time, against the oracle on a machine without a GPU. The product library (libprl.so) does not
// contain or call any of this; its only execution path is the CUDA kernels.
#include <stdint.h>
#include <string.h>

#include "../../physics_reinforcement_learning_environment_b200/csrc/prl_device.cuh"
#include "../../physics_reinforcement_learning_environment_b200/csrc/prl_host.h"

using namespace prl;

static int g_row_cap = 8;
// statistics for sizing the kernels' row storage: histogram of solver rows per environment step,
// and of two-body rows / body-0 ground rows
static long g_need_hist[3][40];
extern "C" void hs_need_hist(long* out, int reset) {
    memcpy(out, g_need_hist, sizeof g_need_hist);
    if (reset) memset(g_need_hist, 0, sizeof g_need_hist);
}
// optional per-step trace (offline grouping studies, scripts/regroup_study.py): 4 ints per step --
// solver rows, narrow-phase pairs visited / recomputed by the detection that preceded the step,
// sleeping mask
static int* g_trace = nullptr;
static long g_trace_cap = 0, g_trace_len = 0;
extern "C" void hs_set_trace(int* buf, long cap_steps) { g_trace = buf; g_trace_cap = cap_steps; g_trace_len = 0; }
extern "C" long hs_trace_len() { return g_trace_len; }
template <class SIM> static void count_need(const SIM& sim) {
    if (g_trace && g_trace_len < g_trace_cap) {
        int* o = g_trace + 4 * g_trace_len++;
        o[0] = sim.need; o[1] = sim.tr_cand; o[2] = sim.tr_full; o[3] = (int)sim.e.sleeping;
    }
    ++g_need_hist[0][sim.need < 39 ? sim.need : 39];
    ++g_need_hist[1][sim.ntb < 39 ? sim.ntb : 39];
    int live = sim.live_count();
    ++g_need_hist[2][live < 39 ? live : 39];
}
// rows per column of the solver's row storage (4 or 8): the layout must not change any result
extern "C" void hs_set_row_cap(int cap) { g_row_cap = cap <= 4 ? 4 : 8; }

template <int NB, class WT, int S>
static void init_env(const WT& W, Env<NB, S>& e) {
    memset(&e, 0, sizeof e);
    for (int b = 0; b < NB; ++b) {
        int c = b < W.n_bodies ? W.collider_of[b] : 0;
        e.x[b] = W.cx[c]; e.y[b] = W.cy[c]; e.c[b] = W.cc[c]; e.s[b] = W.cs[c];
    }
    e.flags = FLAG_FRESH;
}

template <int NB, int PR, class WT>
static float rollout(const WT& W, const uint8_t* genome, int genome_len, int repeat_move, int number_of_steps,
                     int early_exit, float* traj, int* win_out, int* steps_out, int* overflow) {
    using Sim = Stepper<NB, SlotsOf<NB, WT>::value, WT, PR>;
    static typename Sim::E e;
    static Contacts<Sim::kSlots> k;
    init_env(W, e);
    memset(&k, 0, sizeof k);
    alignas(16) static unsigned char pool[Sim::kWarpBytes + Sim::kPoolPad];
    Sim sim(W, e, k, pool);
    float qmin = INFINITY;
    int win = -1, steps = 0;
    auto record = [&]() {
        if (!traj) return;
        for (int b = 0; b < W.n_bodies; ++b) {
            float* o = traj + ((size_t)(steps - 1) * W.n_bodies + b) * 9;
            o[0] = e.x[b]; o[1] = e.y[b]; o[2] = e.c[b]; o[3] = e.s[b]; o[4] = e.vx[b]; o[5] = e.vy[b];
            o[6] = e.w[b]; o[7] = e.timer[b]; o[8] = (float)((e.sleeping >> b) & 1u);
        }
    };
    bool stop = false;
    for (int g = 0; g < genome_len && !stop; ++g) {
        for (int r = 0; r < repeat_move; ++r) {
            // the rollout kernel's path: running minimum of the squared distance, one sqrt at the end
            qmin = fminf(qmin, sim.step_q(genome[g] & 7u));
            count_need(sim);
            ++steps;
            if (traj) sim.detect_pending();  // what a kernel does before it stores the state
            record();
            bool won = (e.flags & FLAG_WON) != 0;
            if (won && win < 0) win = steps;
            if (won && early_exit) { stop = true; break; }
        }
    }
    for (int r = 0; r < number_of_steps % repeat_move; ++r) {
        qmin = fminf(qmin, sim.step_q(0u));
        count_need(sim);
        ++steps;
        if (traj) sim.detect_pending();
        record();
        bool won = (e.flags & FLAG_WON) != 0;
        if (won && win < 0) win = steps;
        if (won && early_exit) break;
    }
    sim.detect_pending();
    *win_out = win;
    *steps_out = steps;
    *overflow = (e.flags & FLAG_OVERFLOW) ? 1 : 0;
    return sqrtf(qmin) / W.scale;
}

// the policy rollout of prl_kernels.cuh policy_rollout_kernel for one environment (theta: this
// agent's parameters, parameter p at theta[p * stride]); optionally records every observation
template <int NB, int PR, class WT>
static float policy_rollout(const WT& W, const float* theta, long stride, int hidden, int number_of_steps, int early_exit,
                            float* obs_out, int* win_out, int* steps_out, int* overflow) {
    using Sim = Stepper<NB, SlotsOf<NB, WT>::value, WT, PR>;
    static typename Sim::E e;
    static Contacts<Sim::kSlots> k;
    init_env(W, e);
    memset(&k, 0, sizeof k);
    alignas(16) static unsigned char pool[Sim::kWarpBytes + Sim::kPoolPad];
    Sim sim(W, e, k, pool);
    float score = INFINITY;
    int win = -1, steps = 0;
    const int D = policy_obs_dim(W.n_bodies);
    for (int t = 0; t < number_of_steps; ++t) {
        sim.detect_pending();
        float obs[kMaxObs];
        sim.observe(obs);
        if (obs_out) memcpy(obs_out + (size_t)t * D, obs, sizeof(float) * D);
        uint32_t mv = policy_move(theta, stride, 0, hidden, D, obs);
        sim.begin_step(mv);
        float d = sim.finish_step(0, 1);
        ++steps;
        score = fminf(score, d);
        bool won = (e.flags & FLAG_WON) != 0;
        if (won && win < 0) win = steps;
        if (won && early_exit) break;
    }
    sim.detect_pending();
    *win_out = win;
    *steps_out = steps;
    *overflow = (e.flags & FLAG_OVERFLOW) ? 1 : 0;
    return score;
}

// host copy of the kernels' stage_world(): compact the full-size tables into WT's capacities
template <class WT>
static void copy_world(WT& d, const DevWorld& s) {
    memset(&d, 0, sizeof d);
    d.n_colliders = s.n_colliders; d.n_bodies = s.n_bodies; d.n_pairs = s.n_pairs; d.n_goals = s.n_goals;
    d.n_player_pairs = s.n_player_pairs; d.n_slots = s.n_slots; d.two_body_pairs = s.two_body_pairs;
    d.radius = s.radius; d.half_height = s.half_height; d.lower_off = s.lower_off;
    d.dt = s.dt; d.inv_dt = s.inv_dt; d.erp_inv_dt = s.erp_inv_dt; d.cfm_factor = s.cfm_factor;
    d.prediction = s.prediction; d.allowed_linear_error = s.allowed_linear_error; d.aabb_margin = s.aabb_margin; d.scale = s.scale; d.win_q = s.win_q;
    d.grid_n = WT::kGrid > 1 ? s.grid_n : 0; d.grid_x0 = s.grid_x0; d.grid_inv_w = s.grid_inv_w;
    if (WT::kGrid > 1) for (int c = 0; c < WT::kGrid; ++c) for (int q = 0; q < 4; ++q) d.grid_mask[c][q] = s.grid_mask[c][q];
    for (int c = 0; c < s.n_colliders; ++c) {
        d.cx[c] = s.cx[c]; d.cy[c] = s.cy[c]; d.cc[c] = s.cc[c]; d.cs[c] = s.cs[c]; d.hx[c] = s.hx[c]; d.hy[c] = s.hy[c];
        d.amnx[c] = s.amnx[c]; d.amny[c] = s.amny[c]; d.amxx[c] = s.amxx[c]; d.amxy[c] = s.amxy[c]; d.body_of[c] = s.body_of[c];
    }
    for (int b = 0; b < MAXB; ++b) {
        d.inv_mass[b] = s.inv_mass[b]; d.inv_inertia_sqrt[b] = s.inv_inertia_sqrt[b]; d.ccd_thickness[b] = s.ccd_thickness[b];
        d.gdv_x[b] = s.gdv_x[b]; d.gdv_y[b] = s.gdv_y[b]; d.collider_of[b] = s.collider_of[b]; d.body_pairs[b] = s.body_pairs[b];
    }
    for (int p = 0; p < s.n_pairs; ++p) {
        d.pair_c1[p] = s.pair_c1[p]; d.pair_c2[p] = s.pair_c2[p]; d.pair_info[p] = s.pair_info[p];
        for (int q = 0; q < 4; ++q) { d.pair_box[p][q] = s.pair_box[p][q]; d.pair_pose[p][q] = s.pair_pose[p][q]; d.pair_he[p][q] = s.pair_he[p][q]; }
    }
    for (int g = 0; g < s.n_goals; ++g) {
        d.gx[g] = s.gx[g]; d.gy[g] = s.gy[g]; d.ghw[g] = s.ghw[g]; d.ghh[g] = s.ghh[g];
        d.gxax[g] = s.gxax[g]; d.gxay[g] = s.gxay[g]; d.gyax[g] = s.gyax[g]; d.gyay[g] = s.gyay[g];
    }
}

extern "C" int hs_rollout(const char* json, size_t len, const uint8_t* genome, int genome_len, int repeat_move,
                          int number_of_steps, int early_exit, float* traj, float* score, int* win, int* steps,
                          int* overflow) {
    prl_world w;
    int rc = parse_world_json(json, len, &w);
    if (rc == PRL_OK) rc = build_dev_world(&w);
    if (rc != PRL_OK) return rc;
    const DevWorld& W = w.dev;
#define HS_ARGS genome, genome_len, repeat_move, number_of_steps, early_exit, traj, win, steps, overflow
    const bool small_pool = g_row_cap < 8;
    // the same variant choice as the library (prl_api.cu pick_variant): small worlds run with the
    // compact tables and slot == pair manifolds, the rest with the general tables
    if (world_fits<DevWorldSmall>(W) && W.n_bodies <= MIDB) {
        static DevWorldSmall S;
        copy_world(S, W);
        if (W.n_bodies == 1) *score = small_pool ? rollout<1, 4>(S, HS_ARGS) : rollout<1, 8>(S, HS_ARGS);
        else if (W.n_bodies == 2) *score = small_pool ? rollout<2, 4>(S, HS_ARGS) : rollout<2, 9>(S, HS_ARGS);
        else *score = small_pool ? rollout<MIDB, 4>(S, HS_ARGS) : rollout<MIDB, 8>(S, HS_ARGS);
    } else if (W.n_bodies == 1) {
        *score = small_pool ? rollout<1, 4>(W, HS_ARGS) : rollout<1, 8>(W, HS_ARGS);
    } else if (W.n_bodies <= MIDB) {
        *score = rollout<MIDB, 8>(W, HS_ARGS);  // (32 manifold slots = up to 64 rows: 8-row columns, as the kernels)
    } else {
        *score = rollout<MAXB, 8>(W, HS_ARGS);
    }
#undef HS_ARGS
    return 0;
}

extern "C" int hs_policy_rollout(const char* json, size_t len, const float* theta, long stride, int hidden, int number_of_steps,
                                 int early_exit, float* obs_out, float* score, int* win, int* steps, int* overflow) {
    prl_world w;
    int rc = parse_world_json(json, len, &w);
    if (rc == PRL_OK) rc = build_dev_world(&w);
    if (rc != PRL_OK) return rc;
    const DevWorld& W = w.dev;
#define HS_ARGS theta, stride, hidden, number_of_steps, early_exit, obs_out, win, steps, overflow
    if (world_fits<DevWorldSmall>(W) && W.n_bodies <= MIDB) {
        static DevWorldSmall S;
        copy_world(S, W);
        if (W.n_bodies == 1) *score = policy_rollout<1, 8>(S, HS_ARGS);
        else if (W.n_bodies == 2) *score = policy_rollout<2, 9>(S, HS_ARGS);
        else *score = policy_rollout<MIDB, 8>(S, HS_ARGS);
    } else if (W.n_bodies == 1) {
        *score = policy_rollout<1, 8>(W, HS_ARGS);
    } else if (W.n_bodies <= MIDB) {
        *score = policy_rollout<MIDB, 8>(W, HS_ARGS);
    } else {
        *score = policy_rollout<MAXB, 8>(W, HS_ARGS);
    }
#undef HS_ARGS
    return 0;
}
