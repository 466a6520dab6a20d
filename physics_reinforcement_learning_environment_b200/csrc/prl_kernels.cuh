// prl_kernels.cuh -- the batch kernels: one thread per environment, the world's static tables in
// shared memory, the environment's state in registers for the whole launch.
//
//   reset_kernel     n x Environment::from_world (reference src/common.rs:162-172)
//   step_kernel      Environment::step (common.rs:215-340) for every environment, n_steps times
//                    (per-environment moves, or one move for all)
//   rollout_kernel   the GA fitness rollout (src/main.rs:44-70) for every agent, from a fresh
//                    from_world state, whole horizon in one launch
//   trajectory_kernel  the viewer's replay of an agent (src/train.rs:132-159): fresh environments,
//                    every body recorded after every step
//   observe_kernel / gather_state_kernel   distance_to_goals / won / body state of the stored batch
//
// HBM layout: structure of arrays, word(f, env) = state[f * n + env] (prl_world.h), so a warp's
// loads and stores of one field are one coalesced 128-byte transaction.
#pragma once
#include "prl_device.cuh"

#if defined(PRL_PHASE_TIMING)
#include <stdio.h>
#define PRL_TICK_K(i) do { long long c_ = clock64(); sim.tacc[i] += c_ - sim.tlast; sim.tlast = c_; } while (0)
#else
#define PRL_TICK_K(i) do { } while (0)
#endif

namespace prl {

constexpr int kBlock = 64;
constexpr int kWarpBlock = 32;  // rollouts of populations too small to fill the SMs with 64-thread blocks
// Shared memory of a block: the world tables plus, per warp, the planes of velocity-constraint
// rows (prl_device.cuh, Stepper): 32 columns of RowCap rows. Single-body worlds need at most three
// rows per environment and get 4-row columns (6 KB per warp). With pushable blocks the column must
// hold EVERY row an environment can reasonably need: a warp with one lane over the limit solves in
// rounds, i.e. takes twice as long, and the block's other 13 warps wait for it at the step barrier
// (push.json needs 8 or 9 rows in only 0.2 % of the environment steps, yet with 7-row columns 63 % of
// the steps had such a lane somewhere in the block: 9-row columns were 27 % faster than 7-row ones).
// One pushable block: 9-row columns (13.5 KB per warp, 194 KB per 448-thread block, which still
// leaves the SM 60 KB of L1 for the manifold records: +1 % over 10-row columns); more: 8-row
// columns of 56-byte rows (all that fits). A 65536-environment batch puts 7 64-thread blocks (or
// one 448-thread block) on each SM.
#ifndef PRL_ROWCAP2
#define PRL_ROWCAP2 9
#endif
// (kernels of 64- and 32-thread blocks have no block barrier, so a round only costs the warp that
// needs it: they keep 7-row columns, which lets seven 64-thread blocks share an SM)
template <int NB, int BLOCK> struct RowCap { static constexpr int value = NB == 1 ? 4 : (NB == 2 ? (BLOCK > 64 ? PRL_ROWCAP2 : 7) : 8); };
template <int NB, class WT, int BLOCK = 64> using SimOf = Stepper<NB, SlotsOf<NB, WT>::value, WT, RowCap<NB, BLOCK>::value>;
// Rollouts of large batches run ONE block of 448 threads per SM (the same 14 warps, 148 x 448 >=
// 65536) with a block barrier at the top of every step: the step is ~4000 straight-line
// instructions outside the solver's loops, and warps of one SM that are in the same phase of the
// step share their instruction fetches (measured +19 % over 7 independent 64-thread blocks, whose
// top stall was no_instruction). Small batches keep 64-thread blocks so they spread over the SMs.
constexpr int kBigBlock = 448;
// Every kRegroupEvery steps (measured: 32: -1.7 %, 64: -0.5 %, 128 and 256 equal) a 448-thread block re-deals its environments to its threads, sorted
// by how many solver rows they needed last step (finished rollouts last): environments parked in
// the HBM state arrays, a counting sort in shared memory, and each thread picks up its new
// environment. Warps then hold environments of similar cost, so fewer lanes idle while their
// neighbours solve contacts (and warps of finished rollouts only hit the barrier).
#ifndef PRL_REGROUP_EVERY
#define PRL_REGROUP_EVERY 128
#endif
constexpr int kRegroupEvery = PRL_REGROUP_EVERY;
constexpr int kRegroupKeys = 34;  // 0 = finished, 1 + min(rows, 32)
template <int BLOCK> constexpr size_t regroup_bytes() { return BLOCK >= kBigBlock ? (kRegroupKeys + 2 * BLOCK) * sizeof(int) : 0; }
template <int NB, class WT, int BLOCK>
constexpr size_t rollout_smem_bytes() {
    return ((sizeof(WT) + 15) / 16) * 16 + (size_t)(BLOCK / 32) * SimOf<NB, WT, BLOCK>::kWarpBytes + SimOf<NB, WT, BLOCK>::kPoolPad +
           regroup_bytes<BLOCK>();
}

// One step's solve for the lanes of a warp. The warp first agrees on the shape of its row
// storage: 32 columns of kCap rows unless some environment needs more, in which case the lanes
// take turns in G rounds and the 32 / G lanes of a round use 32 / G columns of G * kCap rows
// (prl_device.cuh, Stepper). `body(entry, stride)` runs finish_step and the caller's bookkeeping.
// Every lane of the warp must call this (lanes with nothing to do pass run = false, need = 0).
template <class SIM, class F>
__device__ __forceinline__ void solve_in_rounds(bool run, int need, F&& body) {
    const int mx = warp_max(need);
    const int lane = (int)(threadIdx.x & 31u);
    if (!PRL_UNLIKELY(mx > SIM::kCap)) {
        if (run) body(lane, 32);
        return;
    }
    const int lg = mx <= 2 * SIM::kCap ? 1 : (mx <= 4 * SIM::kCap ? 2 : 3);
    __syncwarp();  // the rounds overwrite other lanes' columns of the previous step
#pragma unroll 1
    for (int r = 0; r < (1 << lg); ++r) {
        if (run && (lane & ((1 << lg) - 1)) == r) body(lane >> lg, 32 >> lg);
        __syncwarp();
    }
}

// block-wide copy of the world tables into shared memory (compacting into WT's capacities)
template <class WT>
__device__ __forceinline__ void stage_world(WT* dst, const DevWorld* __restrict__ src) {
    const int t = threadIdx.x, nt = blockDim.x;
    if (t == 0) {
        dst->n_colliders = src->n_colliders; dst->n_bodies = src->n_bodies; dst->n_pairs = src->n_pairs; dst->n_goals = src->n_goals;
        dst->n_player_pairs = src->n_player_pairs; dst->n_slots = src->n_slots; dst->two_body_pairs = src->two_body_pairs;
        dst->radius = src->radius; dst->half_height = src->half_height; dst->lower_off = src->lower_off;
        dst->dt = src->dt; dst->inv_dt = src->inv_dt; dst->erp_inv_dt = src->erp_inv_dt; dst->cfm_factor = src->cfm_factor;
        dst->prediction = src->prediction; dst->allowed_linear_error = src->allowed_linear_error; dst->aabb_margin = src->aabb_margin;
        dst->scale = src->scale; dst->win_q = src->win_q;
        dst->grid_n = WT::kGrid > 1 ? src->grid_n : 0; dst->grid_x0 = src->grid_x0; dst->grid_inv_w = src->grid_inv_w;
    }
    if (WT::kGrid > 1)
        for (int c = t; c < WT::kGrid * 4; c += nt) dst->grid_mask[c >> 2][c & 3] = src->grid_mask[c >> 2][c & 3];
    for (int c = t; c < src->n_colliders; c += nt) {
        dst->cx[c] = src->cx[c]; dst->cy[c] = src->cy[c]; dst->cc[c] = src->cc[c]; dst->cs[c] = src->cs[c];
        dst->hx[c] = src->hx[c]; dst->hy[c] = src->hy[c];
        dst->amnx[c] = src->amnx[c]; dst->amny[c] = src->amny[c]; dst->amxx[c] = src->amxx[c]; dst->amxy[c] = src->amxy[c];
        dst->body_of[c] = src->body_of[c];
    }
    for (int b = t; b < MAXB; b += nt) {
        dst->inv_mass[b] = src->inv_mass[b]; dst->inv_inertia_sqrt[b] = src->inv_inertia_sqrt[b];
        dst->ccd_thickness[b] = src->ccd_thickness[b]; dst->gdv_x[b] = src->gdv_x[b]; dst->gdv_y[b] = src->gdv_y[b];
        dst->collider_of[b] = src->collider_of[b]; dst->body_pairs[b] = src->body_pairs[b];
    }
    for (int p = t; p < src->n_pairs; p += nt) {
        dst->pair_c1[p] = src->pair_c1[p]; dst->pair_c2[p] = src->pair_c2[p]; dst->pair_info[p] = src->pair_info[p];
        dst->pair_box[p][0] = src->pair_box[p][0]; dst->pair_box[p][1] = src->pair_box[p][1];
        dst->pair_box[p][2] = src->pair_box[p][2]; dst->pair_box[p][3] = src->pair_box[p][3];
#pragma unroll
        for (int q = 0; q < 4; ++q) { dst->pair_pose[p][q] = src->pair_pose[p][q]; dst->pair_he[p][q] = src->pair_he[p][q]; }
    }
    for (int g = t; g < src->n_goals; g += nt) {
        dst->gx[g] = src->gx[g]; dst->gy[g] = src->gy[g]; dst->ghw[g] = src->ghw[g]; dst->ghh[g] = src->ghh[g];
        dst->gxax[g] = src->gxax[g]; dst->gxay[g] = src->gxay[g]; dst->gyax[g] = src->gyax[g]; dst->gyay[g] = src->gyay[g];
    }
    __syncthreads();
}

template <int NB, class WT, int S>
__device__ __forceinline__ void env_init(const WT& W, Env<NB, S>& e) {
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        int c = b < W.n_bodies ? W.collider_of[b] : 0;
        e.x[b] = W.cx[c]; e.y[b] = W.cy[c];
        e.c[b] = b == 0 ? 1.0f : W.cc[c];  // the player is never rotated (common.rs:74-79)
        e.s[b] = b == 0 ? 0.0f : W.cs[c];
        e.vx[b] = 0.0f; e.vy[b] = 0.0f; e.w[b] = 0.0f; e.timer[b] = 0.0f;
    }
    e.sleeping = 0;
    e.flags = FLAG_FRESH;
    e.order = 0;
    e.count = 0;
    e.used = 0;
    e.spair[0] = e.spair[1] = 0;
    e.ns1 = e.ns2 = 0;
}

template <int NB, class WT, int S>
__device__ __forceinline__ void env_load(const WT& W, Env<NB, S>& e, Contacts<S>& ct, const uint32_t* __restrict__ st, int64_t n, int64_t i) {
    auto ld = [&](int f) { return __uint_as_float(st[(int64_t)f * n + i]); };
    uint32_t fl = st[(int64_t)F_FLAGS * n + i];
    e.sleeping = fl & 0xffu;
    e.flags = fl & (FLAG_WON | FLAG_FRESH | FLAG_OVERFLOW);
    e.count = (int)((fl >> 16) & 0xffu);
    e.x[0] = ld(F_PLAYER); e.y[0] = ld(F_PLAYER + 1); e.c[0] = 1.0f; e.s[0] = 0.0f;
    e.vx[0] = ld(F_PLAYER + 2); e.vy[0] = ld(F_PLAYER + 3); e.w[0] = 0.0f; e.timer[0] = ld(F_PLAYER + 4);
#pragma unroll
    for (int b = 1; b < NB; ++b) {
        if (b < W.n_bodies) {
            int f = F_DYN + 8 * (b - 1);
            e.x[b] = ld(f); e.y[b] = ld(f + 1); e.c[b] = ld(f + 2); e.s[b] = ld(f + 3);
            e.vx[b] = ld(f + 4); e.vy[b] = ld(f + 5); e.w[b] = ld(f + 6); e.timer[b] = ld(f + 7);
        } else {
            e.x[b] = e.y[b] = e.s[b] = e.vx[b] = e.vy[b] = e.w[b] = e.timer[b] = 0.0f;
            e.c[b] = 1.0f;
        }
    }
    // live manifolds are stored compacted in pair order: HBM slot k = k-th manifold. The 15 HBM
    // words of a slot are the first 15 words of the 16-word thread-local record. Small worlds
    // keep the manifold of pair p in record p (Stepper::kDirect), larger ones in record k.
    constexpr bool direct = SimOf<NB, WT>::kDirect;
    int base = F_DYN + 8 * (W.n_bodies - 1);
    const int count = e.count;
    e.order = 0xfedcba9876543210ull;
    if (S > 16) for (int q = 0; q < S; ++q) ct.order[q] = (uint8_t)q;  // wide lists: HBM slot k = manifold k = record k
    e.used = direct ? 0u : (count >= 32 ? ~0u : (1u << count) - 1u);
    e.spair[0] = e.spair[1] = 0;
    e.ns1 = e.ns2 = 0;
    // all 15 words of a slot are requested before the first one is used:
    // this loop is a chain of global-memory latencies, not of bytes
#pragma unroll 1
    for (int k = 0; k < count; ++k) {
        PRL_CHECK(k < W.n_slots, 4);
        int f = base + MF_WORDS * k;
        uint32_t w[MF_WORDS];
#pragma unroll
        for (int j = 0; j < MF_WORDS; ++j) w[j] = st[(int64_t)(f + j) * n + i];
        const uint32_t h = w[0];
        int slot = direct ? (int)(h & 0xffu) : k;
        PRL_CHECK(slot >= 0 && slot < S, 2);
        if (direct) e.used |= 1u << slot;
        f4* rec = ct.mf[slot];
        f4 q;
        q.x = __uint_as_float(w[0]); q.y = __uint_as_float(w[1]); q.z = __uint_as_float(w[2]); q.w = __uint_as_float(w[3]); rec[0] = q;
        q.x = __uint_as_float(w[4]); q.y = __uint_as_float(w[5]); q.z = __uint_as_float(w[6]); q.w = __uint_as_float(w[7]); rec[1] = q;
        q.x = __uint_as_float(w[8]); q.y = __uint_as_float(w[9]); q.z = __uint_as_float(w[10]); q.w = __uint_as_float(w[11]); rec[2] = q;
        q.x = __uint_as_float(w[12]); q.y = __uint_as_float(w[13]); q.z = __uint_as_float(w[14]); q.w = 0.0f; rec[3] = q;
    }
}

template <int NB, class WT, int S>
__device__ __forceinline__ void env_store(const WT& W, const Env<NB, S>& e, const Contacts<S>& ct, uint32_t* __restrict__ st, int64_t n, int64_t i) {
    auto sf = [&](int f, float v) { st[(int64_t)f * n + i] = __float_as_uint(v); };
    constexpr bool direct = SimOf<NB, WT>::kDirect;
    int count = direct ? __popc(e.used) : e.count;
    count = count < W.n_slots ? count : W.n_slots;
    st[(int64_t)F_FLAGS * n + i] = (e.sleeping & 0xffu) | e.flags | ((uint32_t)count << 16);
    sf(F_PLAYER, e.x[0]); sf(F_PLAYER + 1, e.y[0]); sf(F_PLAYER + 2, e.vx[0]); sf(F_PLAYER + 3, e.vy[0]);
    sf(F_PLAYER + 4, e.timer[0]);
#pragma unroll
    for (int b = 1; b < NB; ++b) {
        if (b < W.n_bodies) {
            int f = F_DYN + 8 * (b - 1);
            sf(f, e.x[b]); sf(f + 1, e.y[b]); sf(f + 2, e.c[b]); sf(f + 3, e.s[b]);
            sf(f + 4, e.vx[b]); sf(f + 5, e.vy[b]); sf(f + 6, e.w[b]); sf(f + 7, e.timer[b]);
        }
    }
    int base = F_DYN + 8 * (W.n_bodies - 1);
    uint32_t left = e.used;
#pragma unroll 1
    for (int k = 0; k < count; ++k) {
        int slot = direct ? __ffs((int)left) - 1 : (S > 16 ? (int)ct.order[k] : (int)((e.order >> (4 * k)) & 15u));
        PRL_CHECK(k < W.n_slots && slot >= 0 && slot < S, 4);
        left &= left - 1;
        int f = base + MF_WORDS * k;
        const f4* rec = ct.mf[slot];
        const f4 a = rec[0], b = rec[1], c = rec[2], d = rec[3];
        st[(int64_t)f * n + i] = __float_as_uint(a.x) & 0xfffu;  // pair | npts << 8 (solver-contact count is derived)
        sf(f + 1, a.y); sf(f + 2, a.z); sf(f + 3, a.w);
        sf(f + 4, b.x); sf(f + 5, b.y); sf(f + 6, b.z); sf(f + 7, b.w);
        sf(f + 8, c.x); sf(f + 9, c.y); sf(f + 10, c.z); sf(f + 11, c.w);
        sf(f + 12, d.x); sf(f + 13, d.y); sf(f + 14, d.z);
    }
}

// n x Environment::from_world (common.rs:162-172)
__global__ void __launch_bounds__(kBlock) reset_kernel(const DevWorld* __restrict__ dw, uint32_t* __restrict__ st, int64_t n) {
    __shared__ DevWorld W;
    stage_world(&W, dw);
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    auto sf = [&](int f, float v) { st[(int64_t)f * n + i] = __float_as_uint(v); };
    st[(int64_t)F_FLAGS * n + i] = FLAG_FRESH;
    sf(F_SCORE, __int_as_float(0x7f800000));
    st[(int64_t)F_WIN_STEP * n + i] = (uint32_t)-1;
    st[(int64_t)F_STEPS * n + i] = 0;
    sf(F_PLAYER, W.cx[0]); sf(F_PLAYER + 1, W.cy[0]); sf(F_PLAYER + 2, 0.0f); sf(F_PLAYER + 3, 0.0f); sf(F_PLAYER + 4, 0.0f);
    for (int b = 1; b < W.n_bodies; ++b) {
        int c = W.collider_of[b], f = F_DYN + 8 * (b - 1);
        sf(f, W.cx[c]); sf(f + 1, W.cy[c]); sf(f + 2, W.cc[c]); sf(f + 3, W.cs[c]);
        sf(f + 4, 0.0f); sf(f + 5, 0.0f); sf(f + 6, 0.0f); sf(f + 7, 0.0f);
    }
}

// Environment::step for every environment. moves != nullptr: one Move byte per environment and
// n_steps == 1; else `uniform_move` for all, n_steps times. Also maintains the running minimum of
// distance_to_goals, the step of the first win and the step counter.
template <int NB, class WT>
__global__ void __launch_bounds__(kBlock, NB > 2 ? 6 : 7) step_kernel(const DevWorld* __restrict__ dw, uint32_t* __restrict__ st, int64_t n,
                                                      const uint8_t* __restrict__ moves, uint32_t uniform_move, int n_steps,
                                                      int* __restrict__ overflow) {
    using Sim = SimOf<NB, WT>;
    __shared__ WT W;
    __shared__ __align__(16) unsigned char row_pool[(kBlock / 32) * Sim::kWarpBytes + Sim::kPoolPad];
    stage_world(&W, dw);
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // the warp agrees on its row storage with a warp vote every step, so every lane stays alive:
    // lanes past the end replay the last environment and skip the stores
    const bool live = i < n;
    if (!live) i = n - 1;
    typename Sim::E e;
    Contacts<Sim::kSlots> ct;
    env_load(W, e, ct, st, n, i);
    float score = __uint_as_float(st[(int64_t)F_SCORE * n + i]);
    int win_step = (int)st[(int64_t)F_WIN_STEP * n + i];
    int steps = (int)st[(int64_t)F_STEPS * n + i];
    uint32_t mv = moves ? (moves[i] & 7u) : (uniform_move & 7u);
    Sim sim(W, e, ct, row_pool + (threadIdx.x >> 5) * Sim::kWarpBytes);
    sim.refresh_solver_contacts();
#pragma unroll 1
    for (int t = 0;; ++t) {
        sim.detect_pending();
        if (t == n_steps) break;
        const int need = sim.begin_step(mv);
        solve_in_rounds<Sim>(true, need, [&](int entry, int stride) {
            float d = sim.finish_step(entry, stride);
            ++steps;
            score = fminf(score, d);
            if ((e.flags & FLAG_WON) && win_step < 0) win_step = steps;
        });
    }
    if (!live) return;
    if (e.flags & FLAG_OVERFLOW) atomicOr(overflow, 1);  // more live manifolds than PRL_MAX_CONTACTS: reported by prl_batch_sync
    env_store(W, e, ct, st, n, i);
    st[(int64_t)F_SCORE * n + i] = __float_as_uint(score);
    st[(int64_t)F_WIN_STEP * n + i] = (uint32_t)win_step;
    st[(int64_t)F_STEPS * n + i] = (uint32_t)steps;
}

// GA fitness rollout (main.rs:44-70) for every agent: fresh from_world state, genome g played
// repeat_move times per gene, then number_of_steps % repeat_move default moves; score = min
// distance_to_goals. early_exit reproduces the reference's break-on-win (including the tail's
// extra step after a win); without it every step is executed (throughput runs).
template <int NB, class WT, int BLOCK>
__global__ void __launch_bounds__(BLOCK, BLOCK >= kBigBlock ? 1 : (NB > 2 ? 6 : 7)) rollout_kernel(const DevWorld* __restrict__ dw, uint32_t* __restrict__ st, int64_t n,
                                                         const uint8_t* __restrict__ genomes, int genome_len, int repeat_move,
                                                         int number_of_steps, int early_exit,
                                                         float* __restrict__ out_score, int32_t* __restrict__ out_win,
                                                         int* __restrict__ overflow, int lanes, int per_block) {
    using Sim = SimOf<NB, WT, BLOCK>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WT& W = *reinterpret_cast<WT*>(smem_raw);
    unsigned char* row_pool = smem_raw + ((sizeof(WT) + 15) / 16) * 16;
    stage_world(&W, dw);
    // `lanes` environments per warp. Populations too small to fill the GPU (BLOCK == kWarpBlock) run in
    // SPARSE warps: a rollout of few environments is bound by the latency of a warp's step, and a warp
    // executes the union of its lanes' paths -- with 7 environments instead of 32 most steps of a random
    // population are pure free flight, and four such warps per SM overlap their latencies.
    // General mapping: a block owns the `per_block` environments [blockIdx * per_block, ...), warp w of it the
    // `lanes` environments [w * lanes, ...) of those (dense: per_block = BLOCK, lanes = 32).
    const int lane_ = (int)(threadIdx.x & 31u), warp_ = (int)(threadIdx.x >> 5);
    int local = warp_ * lanes + lane_;
    int64_t i = (int64_t)blockIdx.x * per_block + local;
    bool live = lane_ < lanes && local < per_block && i < n;
    if (!live) i = n - 1;  // pad lanes only take part in the warp's votes and the block's barriers
    typename Sim::E e;
    Contacts<Sim::kSlots> ct;
    env_init(W, e);
    Sim sim(W, e, ct, row_pool + (threadIdx.x >> 5) * Sim::kWarpBytes);
    // running minimum of the SQUARED physics-space distance to the goals (Stepper::finish_step_sq):
    // one sqrt and one division at the end of the rollout instead of one per step
    float score = __int_as_float(0x7f800000);
    int win_step = -1, steps = 0;
    const uint8_t* genome = genomes + i * (int64_t)genome_len;
    // one flat loop over the gene phase (genome_len x repeat_move steps) and the tail phase
    // (number_of_steps % repeat_move default moves): a win in the gene phase skips to the tail,
    // which then runs one more step before it sees won() -- the reference's two loops. A lane
    // whose rollout has ended idles in the loop until the whole block is done: the warp votes on
    // its row storage every step, and the block barrier at the top of the step keeps the SM's
    // warps in the same phase of the step (shared instruction fetches).
    const int main_steps = genome_len * repeat_move;
    const int total = main_steps + number_of_steps % repeat_move;
    int g = 0, rep = 0, t = live ? 0 : total, iter = 0;
    uint32_t gene = genome_len > 0 ? (genome[0] & 7u) : 0u;
    int* regroup = reinterpret_cast<int*>(row_pool + (BLOCK / 32) * Sim::kWarpBytes + Sim::kPoolPad);  // only if BLOCK >= kBigBlock
#if defined(PRL_PHASE_TIMING)
    sim.tlast = clock64();
#endif
#pragma unroll 1
    while (true) {
        sim.detect_pending();
        // the barrier every step is what keeps the block's warps in the same phase of the step
        // (measured on push, 448-thread blocks: every step 2.59e9 env-steps/s, every 2nd 2.39e9,
        // every 4th 2.16e9, only when regrouping 2.02e9; a second barrier before the solve: no gain)
        const bool all_done = __syncthreads_and(!(t < total));
        PRL_TICK_K(8);
        if (all_done) break;
#if defined(PRL_ICACHE_SCATTER)
        {
            // measurement builds only: PRL_ICACHE_SCATTER chunks of code, each reached by a taken branch
            // over 7 never-executed instructions, i.e. ~160 B of NON-sequential instruction footprint per
            // chunk (profiles/r02_experiments.md)
            float p0 = score;
#pragma unroll
            for (int q = 0; q < PRL_ICACHE_SCATTER; ++q) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %1, -77;\n\t@p bra SKIP%=;\n\t"
                    "add.f32 %0, %0, 0f3F800000;\n\tmul.f32 %0, %0, 0f3F800001;\n\tadd.f32 %0, %0, 0f3F800002;\n\t"
                    "mul.f32 %0, %0, 0f3F800003;\n\tadd.f32 %0, %0, 0f3F800004;\n\tmul.f32 %0, %0, 0f3F800005;\n\t"
                    "add.f32 %0, %0, 0f3F800006;\n\t"
                    "SKIP%=:\n\tadd.f32 %0, %0, 0f3F800000;\n\t}" : "+f"(p0) : "r"(number_of_steps));
            }
            if (p0 == -12345.0f) score = p0;
        }
#endif
#if defined(PRL_ICACHE_PAD)
        {
            // measurement builds only: PRL_ICACHE_PAD KB of straight-line code executed once per step, to
            // measure what a KB of per-step instruction footprint costs (profiles/r02_icache_pad.md)
            float p0 = score, p1 = score, p2 = score, p3 = score;
#pragma unroll
            for (int q = 0; q < PRL_ICACHE_PAD * 16; ++q) {
                asm volatile("add.f32 %0, %0, 0f3F800000;" : "+f"(p0));
                asm volatile("add.f32 %0, %0, 0f3F800000;" : "+f"(p1));
                asm volatile("add.f32 %0, %0, 0f3F800000;" : "+f"(p2));
                asm volatile("add.f32 %0, %0, 0f3F800000;" : "+f"(p3));
            }
            if (p0 + p1 + p2 + p3 == -12345.0f) score = p0;
        }
#endif
        // (single-body worlds in SM-wide blocks run in lock step but are not regrouped: sorting the few
        // walking players of a population into their own warps does not shorten those warps' chains)
        if (BLOCK >= kBigBlock && NB > 1 && PRL_UNLIKELY(iter > 0 && iter % kRegroupEvery == 0)) {
            // ---- regroup: park, sort, pick up (all threads of the block are here, nothing pending)
            int* hist = regroup;
            int* perm_env = regroup + kRegroupKeys;
            int* perm_live = perm_env + BLOCK;
            if (live) {
                env_store(W, e, ct, st, n, i);
                st[(int64_t)F_SCORE * n + i] = __float_as_uint(score);
                st[(int64_t)F_WIN_STEP * n + i] = (uint32_t)win_step;
                st[(int64_t)F_STEPS * n + i] = (uint32_t)steps;
                st[(int64_t)F_ROLL * n + i] = (uint32_t)t;
            }
            const int key = t < total ? 1 + (sim.need < 32 ? sim.need : 32) : 0;  // finished rollouts last
            if (threadIdx.x < kRegroupKeys) hist[threadIdx.x] = 0;
            __syncthreads();
            const int pos = atomicAdd(&hist[key], 1);
            __syncthreads();
            int off = 0;
            for (int k = key + 1; k < kRegroupKeys; ++k) off += hist[k];
            PRL_CHECK(off + pos >= 0 && off + pos < BLOCK && key >= 0 && key < kRegroupKeys, 6);
            perm_env[off + pos] = live ? (int)(i - (int64_t)blockIdx.x * per_block) : -1;
            perm_live[off + pos] = live ? 1 : 0;
            __syncthreads();
            // running rollouts sort first: the warps' real lanes take them in order, pad lanes take the rest
            const int slots = (BLOCK / 32) * lanes;
            const int mine = lane_ < lanes ? warp_ * lanes + lane_ : slots + warp_ * (32 - lanes) + (lane_ - lanes);
            PRL_CHECK(mine >= 0 && mine < BLOCK, 6);
            live = perm_live[mine] != 0;
            i = live ? (int64_t)blockIdx.x * per_block + perm_env[mine] : n - 1;
            env_load(W, e, ct, st, n, i);
            score = __uint_as_float(st[(int64_t)F_SCORE * n + i]);
            win_step = (int)st[(int64_t)F_WIN_STEP * n + i];
            steps = (int)st[(int64_t)F_STEPS * n + i];
            t = live ? (int)st[(int64_t)F_ROLL * n + i] : total;
            sim.refresh_solver_contacts();
            genome = genomes + i * (int64_t)genome_len;
            g = t < main_steps ? t / repeat_move : genome_len;
            rep = t < main_steps ? t % repeat_move : 0;
            gene = g < genome_len ? (genome[g] & 7u) : 0u;
            __syncthreads();  // hist / perm are reused by the next regroup
            PRL_TICK_K(9);
        }
        ++iter;
        const bool run = t < total;
        const bool in_genes = t < main_steps;
        const int need = run ? sim.begin_step(in_genes ? gene : 0u) : 0;
        solve_in_rounds<Sim>(run, need, [&](int entry, int stride) {
            const float q = sim.finish_step_sq(entry, stride);
            ++steps;
            ++t;
            score = fminf(score, q);
            bool won = (e.flags & FLAG_WON) != 0;
            if (won && win_step < 0) win_step = steps;
            if (in_genes && ++rep == repeat_move) {
                rep = 0;
                if (++g < genome_len) gene = genome[g] & 7u;
                PRL_CHECK(g <= genome_len && i >= 0 && i < n, 5);
            }
            if (won && early_exit) t = in_genes ? main_steps : total;
        });
    }
#if defined(PRL_PHASE_TIMING)
    if (blockIdx.x == 3 && (threadIdx.x & 31u) == 0)
        printf("warp %2d: broad %lld player-np %lld block-np %lld move %lld select %lld gen %lld sweeps %lld rest %lld barrier %lld regroup %lld\n",
               (int)(threadIdx.x >> 5), sim.tacc[0] / iter, sim.tacc[1] / iter, sim.tacc[2] / iter, sim.tacc[3] / iter, sim.tacc[4] / iter,
               sim.tacc[5] / iter, sim.tacc[6] / iter, sim.tacc[7] / iter, sim.tacc[8] / iter, sim.tacc[9] / iter);
#endif
    if (!live) return;
    if (e.flags & FLAG_OVERFLOW) atomicOr(overflow, 1);  // more live manifolds than PRL_MAX_CONTACTS: reported by prl_batch_sync
    env_store(W, e, ct, st, n, i);
    score = sqrtf(score) / W.scale;  // distance_to_goals (common.rs:192) of the closest approach, editor pixels
    st[(int64_t)F_SCORE * n + i] = __float_as_uint(score);
    st[(int64_t)F_WIN_STEP * n + i] = (uint32_t)win_step;
    st[(int64_t)F_STEPS * n + i] = (uint32_t)steps;
    if (out_score) out_score[i] = score;
    if (out_win) out_win[i] = win_step;
}

// ---- Closed-loop agents evaluated inside the fused rollout: the batched form of
// Agent::get_move(&mut self, &Environment) (algorithm.rs:8-11) for a small per-agent MLP.
// Observation = Stepper::observe (bodies, ground-contact scan of common.rs:223-244, distance to the
// goals), D = 9 n_bodies + 4 values. Policy of agent a, parameters theta[p * n + a] (parameter-major,
// so a warp reads consecutive agents' copies of one parameter in one transaction):
//   h_j   = max(b1_j + sum_i W1[j][i] obs_i, 0)      j < hidden      p = j (D + 1) + i, bias at i = D
//   out_k = b2_k + sum_j W2[k][j] h_j                 k < 3           p = hidden (D + 1) + k (hidden + 1) + j
//   Move  = (left, right, up) = (out_0 > 0, out_1 > 0, out_2 > 0)
// Sums run left to right in f32 without contraction, so the CPU oracle evaluates the same bits.
// The GA-style fitness rollout (main.rs:44-70 with repeat_move = 1) driven by the policy: every step
// mv = policy(observation), Environment::step(mv), score = min distance; early_exit: break on win.
template <int NB, class WT>
__global__ void __launch_bounds__(kBlock, NB > 2 ? 6 : 7) policy_rollout_kernel(const DevWorld* __restrict__ dw, uint32_t* __restrict__ st, int64_t n,
                                                                                 const float* __restrict__ theta, int hidden, int number_of_steps,
                                                                                 int early_exit, float* __restrict__ out_score,
                                                                                 int32_t* __restrict__ out_win, int* __restrict__ overflow) {
    using Sim = SimOf<NB, WT>;
    __shared__ WT W;
    __shared__ __align__(16) unsigned char row_pool[(kBlock / 32) * Sim::kWarpBytes + Sim::kPoolPad];
    stage_world(&W, dw);
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n;
    if (!live) i = n - 1;
    typename Sim::E e;
    Contacts<Sim::kSlots> ct;
    env_init(W, e);
    Sim sim(W, e, ct, row_pool + (threadIdx.x >> 5) * Sim::kWarpBytes);
    float score = __int_as_float(0x7f800000);
    int win_step = -1, steps = 0, t = 0;
    const int D = policy_obs_dim(W.n_bodies);
#pragma unroll 1
    while (true) {
        sim.detect_pending();
        if (warp_all(!(t < number_of_steps))) break;  // the warp votes on its row storage every step: lanes leave together
        const bool run = t < number_of_steps;
        uint32_t mv = 0;
        if (run) {
            float obs[kMaxObs];
            sim.observe(obs);
            mv = policy_move(theta, n, i, hidden, D, obs);
        }
        const int need = run ? sim.begin_step(mv) : 0;
        solve_in_rounds<Sim>(run, need, [&](int entry, int stride) {
            float d = sim.finish_step(entry, stride);
            ++steps;
            ++t;
            score = fminf(score, d);
            const bool won = (e.flags & FLAG_WON) != 0;
            if (won && win_step < 0) win_step = steps;
            if (won && early_exit) t = number_of_steps;
        });
    }
    if (!live) return;
    if (e.flags & FLAG_OVERFLOW) atomicOr(overflow, 1);
    env_store(W, e, ct, st, n, i);
    st[(int64_t)F_SCORE * n + i] = __float_as_uint(score);
    st[(int64_t)F_WIN_STEP * n + i] = (uint32_t)win_step;
    st[(int64_t)F_STEPS * n + i] = (uint32_t)steps;
    if (out_score) out_score[i] = score;
    if (out_win) out_win[i] = win_step;
}

// Observation of the stored batch (layout of Stepper::observe): n x D floats
template <int NB, class WT>
__global__ void __launch_bounds__(kBlock, NB > 2 ? 6 : 7) observation_kernel(const DevWorld* __restrict__ dw, const uint32_t* __restrict__ st, int64_t n,
                                                                             float* __restrict__ out) {
    using Sim = SimOf<NB, WT>;
    __shared__ WT W;
    __shared__ __align__(16) unsigned char row_pool[(kBlock / 32) * Sim::kWarpBytes + Sim::kPoolPad];
    stage_world(&W, dw);
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    typename Sim::E e;
    Contacts<Sim::kSlots> ct;
    env_load(W, e, ct, st, n, i);
    Sim sim(W, e, ct, row_pool + (threadIdx.x >> 5) * Sim::kWarpBytes);
    sim.refresh_solver_contacts();
    float obs[kMaxObs];
    sim.observe(obs);
    const int D = policy_obs_dim(W.n_bodies);
    for (int q = 0; q < D; ++q) out[i * D + q] = obs[q];
}

// Replay (train.rs:132-159): n fresh environments each play a GeneticAgent (main.rs:276-284) for
// number_of_steps steps; every body's state, the distance to the goals and won() are recorded
// after every step. out_state: steps x n x n_bodies x 9 (layout of gather_state_kernel).
template <int NB, class WT>
__global__ void __launch_bounds__(kBlock, NB > 2 ? 6 : 7) trajectory_kernel(const DevWorld* __restrict__ dw, int64_t n,
                                                                            const uint8_t* __restrict__ genomes, int genome_len,
                                                                            int repeat_move, int number_of_steps,
                                                                            float* __restrict__ out_state, float* __restrict__ out_dist,
                                                                            uint8_t* __restrict__ out_won, int* __restrict__ overflow) {
    using Sim = SimOf<NB, WT>;
    __shared__ WT W;
    __shared__ __align__(16) unsigned char row_pool[(kBlock / 32) * Sim::kWarpBytes + Sim::kPoolPad];
    stage_world(&W, dw);
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n;
    if (!live) i = n - 1;
    typename Sim::E e;
    Contacts<Sim::kSlots> ct;
    env_init(W, e);
    Sim sim(W, e, ct, row_pool + (threadIdx.x >> 5) * Sim::kWarpBytes);
    const uint8_t* genome = genomes + i * (int64_t)genome_len;
    const int nb = W.n_bodies;
#pragma unroll 1
    for (int t = 0; t < number_of_steps; ++t) {
        sim.detect_pending();
        const int g = t / repeat_move;
        const uint32_t mv = g < genome_len ? (genome[g] & 7u) : 0u;  // GeneticAgent::get_move
        const int need = sim.begin_step(mv);
        float d = 0.0f;
        solve_in_rounds<Sim>(true, need, [&](int entry, int stride) { d = sim.finish_step(entry, stride); });
        sim.detect_pending();  // pair removal may wake a body: the recorded state is the state a step kernel would store
        if (live) {
            const int64_t rec = (int64_t)t * n + i;
            if (out_dist) out_dist[rec] = d;
            if (out_won) out_won[rec] = (e.flags & FLAG_WON) ? 1 : 0;
            if (out_state) {
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    if (b < nb) {
                        float* o = out_state + (rec * nb + b) * 9;
                        o[0] = e.x[b]; o[1] = e.y[b]; o[2] = e.c[b]; o[3] = e.s[b]; o[4] = e.vx[b]; o[5] = e.vy[b];
                        o[6] = e.w[b]; o[7] = e.timer[b]; o[8] = (float)((e.sleeping >> b) & 1u);
                    }
                }
            }
        }
    }
    if (live && (e.flags & FLAG_OVERFLOW)) atomicOr(overflow, 1);
}

// Environment::distance_to_goals (common.rs:175-195) and Environment::won (common.rs:197-199)
__global__ void __launch_bounds__(kBlock) observe_kernel(const DevWorld* __restrict__ dw, const uint32_t* __restrict__ st, int64_t n,
                                                         float* __restrict__ out_dist, uint8_t* __restrict__ out_won) {
    __shared__ DevWorld W;
    stage_world(&W, dw);
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (out_won) out_won[i] = (st[(int64_t)F_FLAGS * n + i] & FLAG_WON) ? 1 : 0;
    if (out_dist) {
        float px = __uint_as_float(st[(int64_t)F_PLAYER * n + i]), py = __uint_as_float(st[(int64_t)(F_PLAYER + 1) * n + i]);
        float best = __int_as_float(0x7f800000);
        for (int g = 0; g < W.n_goals; ++g) {
            float dx0 = px - W.gx[g], dy0 = py - W.gy[g];
            float dx = fmaxf(fabsf(dx0 * W.gxax[g] + dy0 * W.gxay[g]) - W.ghw[g], 0.0f);
            float dy = fmaxf(fabsf(dx0 * W.gyax[g] + dy0 * W.gyay[g]) - W.ghh[g], 0.0f);
            float dist = sqrtf(dx * dx + dy * dy) / W.scale;
            best = g == 0 ? dist : fminf(best, dist);
        }
        out_dist[i] = best;
    }
}

// Observation: bodies of environments [env0, env0+cnt) as cnt x n_bodies x 9 floats
// (x, y, cos, sin, vx, vy, angvel, sleep timer, sleeping)
__global__ void __launch_bounds__(kBlock) gather_state_kernel(const uint32_t* __restrict__ st, int64_t n, int n_bodies, int64_t env0,
                                                              int64_t cnt, float* __restrict__ out) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cnt * n_bodies) return;
    int64_t i = env0 + t / n_bodies;
    int b = (int)(t % n_bodies);
    auto ld = [&](int f) { return __uint_as_float(st[(int64_t)f * n + i]); };
    float* o = out + t * 9;
    uint32_t fl = st[(int64_t)F_FLAGS * n + i];
    if (b == 0) {
        o[0] = ld(F_PLAYER); o[1] = ld(F_PLAYER + 1); o[2] = 1.0f; o[3] = 0.0f;
        o[4] = ld(F_PLAYER + 2); o[5] = ld(F_PLAYER + 3); o[6] = 0.0f; o[7] = ld(F_PLAYER + 4);
    } else {
        int f = F_DYN + 8 * (b - 1);
        for (int j = 0; j < 8; ++j) o[j] = ld(f + j);
    }
    o[8] = (float)((fl >> b) & 1u);
}

}  // namespace prl
