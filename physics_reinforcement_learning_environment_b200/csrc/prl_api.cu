// prl_api.cu -- the C ABI of include/prl.h over the batch kernels. No CPU fallback: every compute
// entry point needs a CUDA device and fails with PRL_ERR_CUDA otherwise.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <string>
#include <vector>

#include "prl_host.h"
#include "prl_batch.h"
#include "prl_kernels.cuh"

namespace prl {
const char* last_error();
}
using namespace prl;



namespace {


bool is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int ensure_stage(void** buf, size_t* cap, size_t bytes) {
    if (*cap >= bytes) return PRL_OK;
    if (*buf) cudaFree(*buf);
    *buf = nullptr;
    *cap = 0;
    CUDA_TRY(cudaMalloc(buf, bytes));
    *cap = bytes;
    return PRL_OK;
}

inline unsigned grid_for(int64_t n) { return (unsigned)((n + kBlock - 1) / kBlock); }

// Kernel variants: (bodies, world-table capacity). Small worlds (<= 15 blocks, the shipped
// levels) with at most one pushable block get the compact tables and the 48-byte solver rows;
// everything else runs the general variant.
enum Variant { V1_SMALL, V2_SMALL, VN_SMALL, V1_BIG, VN_BIG, VX_BIG };
Variant pick_variant(const DevWorld& w) {
    bool small = world_fits<DevWorldSmall>(w);
    if (w.n_bodies == 1) return small ? V1_SMALL : V1_BIG;
    if (w.n_bodies > MIDB) return VX_BIG;  // 5 - 7 pushable blocks: body state no longer fits the register file, exact but slower
    if (small) return w.n_bodies == 2 ? V2_SMALL : VN_SMALL;
    return VN_BIG;
}
#define PRL_DISPATCH(variant, KERNEL, grid, stream, ...)                                                        \
    switch (variant) {                                                                                           \
        case V1_SMALL: KERNEL<1, DevWorldSmall><<<grid, kBlock, 0, stream>>>(__VA_ARGS__); break;               \
        case V2_SMALL: KERNEL<2, DevWorldSmall><<<grid, kBlock, 0, stream>>>(__VA_ARGS__); break;               \
        case VN_SMALL: KERNEL<MIDB, DevWorldSmall><<<grid, kBlock, 0, stream>>>(__VA_ARGS__); break;            \
        case V1_BIG: KERNEL<1, DevWorld><<<grid, kBlock, 0, stream>>>(__VA_ARGS__); break;                      \
        case VN_BIG: KERNEL<MIDB, DevWorld><<<grid, kBlock, 0, stream>>>(__VA_ARGS__); break;                   \
        default: KERNEL<MAXB, DevWorld><<<grid, kBlock, 0, stream>>>(__VA_ARGS__); break;                       \
    }

}  // namespace

extern "C" {

const char* prl_last_error(void) { return prl::last_error(); }
int32_t prl_abi_version(void) { return PRL_ABI_VERSION; }
#if defined(PRL_BOUNDS_CHECK)
// negative control of the checking machinery: a check that must fail (code 30)
__global__ void bounds_selftest_kernel(int n) { PRL_CHECK((int)threadIdx.x >= n, 30); }
#endif
int32_t prl_debug_bounds_violations(void) {
#if defined(PRL_BOUNDS_CHECK)
    unsigned int v = 0;
    if (getenv("PRL_BOUNDS_SELFTEST")) bounds_selftest_kernel<<<1, 32>>>(64);
    if (cudaDeviceSynchronize() != cudaSuccess) return -2;
    if (cudaMemcpyFromSymbol(&v, g_prl_bounds_violations, sizeof v) != cudaSuccess) return -2;
    return (int32_t)v;
#else
    return -1;
#endif
}
int32_t prl_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ------------------------------------------------------------------------------- World
int prl_world_create(const float player_position[2], const prl_object* objs, int32_t n_objects, prl_world** out) {
    if (!player_position || !out || n_objects < 0 || (n_objects > 0 && !objs)) { set_error("prl_world_create: null argument"); return PRL_ERR_INVALID; }
    prl_world* w = new (std::nothrow) prl_world();
    if (!w) { set_error("out of memory"); return PRL_ERR_INVALID; }
    w->player_position[0] = player_position[0];
    w->player_position[1] = player_position[1];
    w->objects.assign(objs, objs + n_objects);
    int rc = build_dev_world(w);
    if (rc != PRL_OK) { delete w; return rc; }
    *out = w;
    return PRL_OK;
}

int prl_world_from_json(const char* utf8, size_t len, prl_world** out) {
    if (!utf8 || !out) { set_error("prl_world_from_json: null argument"); return PRL_ERR_INVALID; }
    prl_world* w = new (std::nothrow) prl_world();
    if (!w) { set_error("out of memory"); return PRL_ERR_INVALID; }
    int rc = parse_world_json(utf8, len, w);
    if (rc == PRL_OK) rc = build_dev_world(w);
    if (rc != PRL_OK) { delete w; return rc; }
    *out = w;
    return PRL_OK;
}

int prl_world_to_json(const prl_world* w, char* buf, size_t cap, size_t* needed) {
    if (!w) { set_error("prl_world_to_json: null world"); return PRL_ERR_INVALID; }
    std::string s = world_to_json(w);
    if (needed) *needed = s.size() + 1;
    if (buf && cap > 0) {
        size_t k = s.size() < cap - 1 ? s.size() : cap - 1;
        memcpy(buf, s.data(), k);
        buf[k] = 0;
    }
    return PRL_OK;
}

int32_t prl_world_num_objects(const prl_world* w) { return w ? (int32_t)w->objects.size() : 0; }
int prl_world_get_object(const prl_world* w, int32_t index, prl_object* out) {
    if (!w || !out || index < 0 || index >= (int32_t)w->objects.size()) { set_error("prl_world_get_object: bad index"); return PRL_ERR_INVALID; }
    *out = w->objects[index];
    return PRL_OK;
}
int prl_world_player_position(const prl_world* w, float out[2]) {
    if (!w || !out) { set_error("prl_world_player_position: null argument"); return PRL_ERR_INVALID; }
    out[0] = w->player_position[0];
    out[1] = w->player_position[1];
    return PRL_OK;
}
int prl_world_counts(const prl_world* w, int32_t* n_bodies, int32_t* n_colliders, int32_t* n_goals, int32_t* n_pairs) {
    if (!w) { set_error("prl_world_counts: null world"); return PRL_ERR_INVALID; }
    if (n_bodies) *n_bodies = w->dev.n_bodies;
    if (n_colliders) *n_colliders = w->dev.n_colliders;
    if (n_goals) *n_goals = w->dev.n_goals;
    if (n_pairs) *n_pairs = w->dev.n_pairs;
    return PRL_OK;
}
int32_t prl_world_object_body(const prl_world* w, int32_t index) {
    if (!w || index < 0 || index >= (int32_t)w->object_body.size()) return -1;
    return w->object_body[index];
}
void prl_world_destroy(prl_world* w) { delete w; }

// ------------------------------------------------------------------------------- Batch
int prl_batch_create(const prl_world* w, int64_t n_envs, int32_t device, void* stream, prl_batch** out) {
    if (!w || !out || n_envs <= 0) { set_error("prl_batch_create: bad argument"); return PRL_ERR_INVALID; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device is usable; this library has no CPU path");
        return PRL_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { set_error("device %d out of range (%d visible)", device, ndev); return PRL_ERR_INVALID; }
    DeviceGuard g(device);
    prl_batch* b = new (std::nothrow) prl_batch();
    if (!b) { set_error("out of memory"); return PRL_ERR_INVALID; }
    b->world_src = w;
    b->host_world = w->dev;
    b->n = n_envs;
    b->device = device;
    b->stream = (cudaStream_t)stream;
    if (cudaDeviceGetAttribute(&b->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) b->sm_count = 148;
    b->words = state_words(w->dev.n_bodies, w->dev.n_slots);
    cudaError_t e2 = cudaMalloc(&b->dev_world, sizeof(DevWorld));
    if (e2 == cudaSuccess) e2 = cudaMalloc(&b->state, (size_t)b->words * (size_t)n_envs * 4u);
    if (e2 == cudaSuccess) e2 = cudaMalloc(&b->overflow, sizeof(int));
    if (e2 != cudaSuccess) {
        set_error("cudaMalloc of %d state words x %lld environments failed: %s", b->words, (long long)n_envs, cudaGetErrorString(e2));
        prl_batch_destroy(b);
        return PRL_ERR_CUDA;
    }
    cudaError_t e3 = cudaMemcpyAsync(b->dev_world, &b->host_world, sizeof(DevWorld), cudaMemcpyHostToDevice, b->stream);
    if (e3 == cudaSuccess) e3 = cudaStreamSynchronize(b->stream);
    if (e3 != cudaSuccess) {
        set_error("uploading the world tables failed: %s", cudaGetErrorString(e3));
        prl_batch_destroy(b);
        return PRL_ERR_CUDA;
    }
    int rc = prl_batch_reset(b);
    if (rc != PRL_OK) { prl_batch_destroy(b); return rc; }
    *out = b;
    return PRL_OK;
}

void prl_batch_destroy(prl_batch* b) {
    if (!b) return;
    DeviceGuard g(b->device);
    cudaStreamSynchronize(b->stream);
    if (b->dev_world) cudaFree(b->dev_world);
    if (b->state) cudaFree(b->state);
    if (b->overflow) cudaFree(b->overflow);
    if (b->stage) cudaFree(b->stage);
    if (b->stage2) cudaFree(b->stage2);
    if (b->traj) cudaFree(b->traj);
    if (b->pinned) cudaFreeHost(b->pinned);
    if (b->pinned_free) cudaEventDestroy(b->pinned_free);
    delete b;
}

int64_t prl_batch_num_envs(const prl_batch* b) { return b ? b->n : 0; }
int32_t prl_batch_num_bodies(const prl_batch* b) { return b ? b->host_world.n_bodies : 0; }
int64_t prl_batch_launch_count(const prl_batch* b) { return b ? b->launches : 0; }
int64_t prl_batch_state_bytes(const prl_batch* b) { return b ? (int64_t)b->words * b->n * 4 : 0; }

int prl_batch_reset(prl_batch* b) {
    if (!b) { set_error("prl_batch_reset: null batch"); return PRL_ERR_INVALID; }
    DeviceGuard g(b->device);
    CUDA_TRY(cudaMemsetAsync(b->overflow, 0, sizeof(int), b->stream));
    reset_kernel<<<grid_for(b->n), kBlock, 0, b->stream>>>(b->dev_world, b->state, b->n);
    ++b->launches;
    CUDA_TRY(cudaGetLastError());
    return PRL_OK;
}

static int launch_step(prl_batch* b, const uint8_t* dev_moves, uint32_t uniform, int n_steps) {
    PRL_DISPATCH(pick_variant(b->host_world), step_kernel, grid_for(b->n), b->stream, b->dev_world, b->state, b->n, dev_moves, uniform, n_steps, b->overflow);
    ++b->launches;
    CUDA_TRY(cudaGetLastError());
    return PRL_OK;
}

int prl_batch_step(prl_batch* b, const uint8_t* moves) {
    if (!b || !moves) { set_error("prl_batch_step: null argument"); return PRL_ERR_INVALID; }
    DeviceGuard g(b->device);
    const uint8_t* dm = moves;
    if (!is_device_ptr(moves)) {
        // host moves: through a pinned buffer owned by the batch, so that the transfer is asynchronous
        // and the call does not wait for the stream (a closed loop on the host steps once per call)
        int rc = ensure_stage(&b->stage, &b->stage_bytes, (size_t)b->n);
        if (rc != PRL_OK) return rc;
        if (b->pinned_bytes < (size_t)b->n) {
            if (b->pinned) { CUDA_TRY(cudaStreamSynchronize(b->stream)); cudaFreeHost(b->pinned); b->pinned = nullptr; b->pinned_bytes = 0; }
            CUDA_TRY(cudaMallocHost(&b->pinned, (size_t)b->n));
            b->pinned_bytes = (size_t)b->n;
        }
        if (!b->pinned_free) CUDA_TRY(cudaEventCreateWithFlags(&b->pinned_free, cudaEventDisableTiming));
        else CUDA_TRY(cudaEventSynchronize(b->pinned_free));  // the previous step's transfer has left the buffer
        memcpy(b->pinned, moves, (size_t)b->n);
        CUDA_TRY(cudaMemcpyAsync(b->stage, b->pinned, (size_t)b->n, cudaMemcpyHostToDevice, b->stream));
        CUDA_TRY(cudaEventRecord(b->pinned_free, b->stream));
        dm = (const uint8_t*)b->stage;
    }
    return launch_step(b, dm, 0, 1);
}

int prl_batch_step_uniform(prl_batch* b, uint32_t move, int32_t n_steps) {
    if (!b || n_steps < 0) { set_error("prl_batch_step_uniform: bad argument"); return PRL_ERR_INVALID; }
    if (n_steps == 0) return PRL_OK;
    DeviceGuard g(b->device);
    return launch_step(b, nullptr, move, n_steps);
}

static int observe(prl_batch* b, float* out_dist, uint8_t* out_won) {
    DeviceGuard g(b->device);
    void* host_out = nullptr;
    size_t bytes = 0;
    float* dd = out_dist;
    uint8_t* dw = out_won;
    if (out_dist && !is_device_ptr(out_dist)) {
        bytes = (size_t)b->n * 4;
        int rc = ensure_stage(&b->stage2, &b->stage2_bytes, bytes);
        if (rc != PRL_OK) return rc;
        dd = (float*)b->stage2;
        host_out = out_dist;
    }
    if (out_won && !is_device_ptr(out_won)) {
        bytes = (size_t)b->n;
        int rc = ensure_stage(&b->stage2, &b->stage2_bytes, bytes);
        if (rc != PRL_OK) return rc;
        dw = (uint8_t*)b->stage2;
        host_out = out_won;
    }
    observe_kernel<<<grid_for(b->n), kBlock, 0, b->stream>>>(b->dev_world, b->state, b->n, dd, dw);
    ++b->launches;
    CUDA_TRY(cudaGetLastError());
    if (host_out) {
        CUDA_TRY(cudaMemcpyAsync(host_out, b->stage2, bytes, cudaMemcpyDeviceToHost, b->stream));
        CUDA_TRY(cudaStreamSynchronize(b->stream));
    }
    return PRL_OK;
}

int prl_batch_distance_to_goals(prl_batch* b, float* out) {
    if (!b || !out) { set_error("prl_batch_distance_to_goals: null argument"); return PRL_ERR_INVALID; }
    if (b->host_world.n_goals == 0) { set_error("the world has no goal: distance_to_goals() is None"); return PRL_ERR_NO_GOAL; }
    return observe(b, out, nullptr);
}

int prl_batch_won(prl_batch* b, uint8_t* out) {
    if (!b || !out) { set_error("prl_batch_won: null argument"); return PRL_ERR_INVALID; }
    return observe(b, nullptr, out);
}

int prl_batch_get_state(prl_batch* b, int64_t env0, int64_t cnt, float* out) {
    if (!b || !out || env0 < 0 || cnt < 0 || env0 + cnt > b->n) { set_error("prl_batch_get_state: bad range"); return PRL_ERR_INVALID; }
    if (cnt == 0) return PRL_OK;
    DeviceGuard g(b->device);
    const int nb = b->host_world.n_bodies;
    size_t bytes = (size_t)cnt * nb * 9 * 4;
    bool dev = is_device_ptr(out);
    float* d = out;
    if (!dev) {
        int rc = ensure_stage(&b->stage2, &b->stage2_bytes, bytes);
        if (rc != PRL_OK) return rc;
        d = (float*)b->stage2;
    }
    gather_state_kernel<<<grid_for(cnt * nb), kBlock, 0, b->stream>>>(b->state, b->n, nb, env0, cnt, d);
    ++b->launches;
    CUDA_TRY(cudaGetLastError());
    if (!dev) {
        CUDA_TRY(cudaMemcpyAsync(out, d, bytes, cudaMemcpyDeviceToHost, b->stream));
        CUDA_TRY(cudaStreamSynchronize(b->stream));
    }
    return PRL_OK;
}

int prl_batch_rollout_info(prl_batch* b, float* out_score, int32_t* out_step_of_win, int32_t* out_steps) {
    if (!b) { set_error("prl_batch_rollout_info: null batch"); return PRL_ERR_INVALID; }
    DeviceGuard g(b->device);
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    if (out_score) CUDA_TRY(cudaMemcpy(out_score, b->state + (size_t)F_SCORE * b->n, (size_t)b->n * 4, cudaMemcpyDefault));
    if (out_step_of_win) CUDA_TRY(cudaMemcpy(out_step_of_win, b->state + (size_t)F_WIN_STEP * b->n, (size_t)b->n * 4, cudaMemcpyDefault));
    if (out_steps) CUDA_TRY(cudaMemcpy(out_steps, b->state + (size_t)F_STEPS * b->n, (size_t)b->n * 4, cudaMemcpyDefault));
    return PRL_OK;
}

int prl_batch_rollout(prl_batch* b, const uint8_t* genomes, int32_t genome_len, int32_t repeat_move, int32_t number_of_steps,
                      int32_t early_exit, float* out_score, int32_t* out_step_of_win) {
    if (!b || !genomes || genome_len < 0 || repeat_move < 1 || number_of_steps < 0) { set_error("prl_batch_rollout: bad argument"); return PRL_ERR_INVALID; }
    if (b->host_world.n_goals == 0) { set_error("the world has no goal: the rollout's distance_to_goals().unwrap() would panic (main.rs:50)"); return PRL_ERR_NO_GOAL; }
    DeviceGuard g(b->device);
    const uint8_t* dg = genomes;
    size_t gbytes = (size_t)b->n * (size_t)genome_len;
    if (!is_device_ptr(genomes) && gbytes > 0) {
        int rc = ensure_stage(&b->stage, &b->stage_bytes, gbytes);
        if (rc != PRL_OK) return rc;
        CUDA_TRY(cudaMemcpyAsync(b->stage, genomes, gbytes, cudaMemcpyHostToDevice, b->stream));
        dg = (const uint8_t*)b->stage;
    }
    bool score_host = out_score && !is_device_ptr(out_score);
    bool win_host = out_step_of_win && !is_device_ptr(out_step_of_win);
    float* ds = score_host ? nullptr : out_score;
    int32_t* dw = win_host ? nullptr : out_step_of_win;
    {
        // one 448-thread block per SM once the batch can fill half the SMs that way, else 64-thread
        // blocks: the SM's warps then run in lock step (one barrier per step) and share their instruction
        // fetches -- the step's code is larger than the SM's 32 KB instruction cache
        // (profiles/r02_icache_probe.txt). Kernels with pushable blocks are also regrouped every 128 steps.
        const bool big = b->n >= (int64_t)b->sm_count * kBigBlock / 2;
        // a population too small to give every SM a 64-thread block (config C2: 4096 agents) runs one
        // warp per block, so that it spreads over twice as many SMs: such a rollout is bound by the
        // latency of a single warp's step, not by throughput
        const bool tiny = (b->n + kBlock - 1) / kBlock < (int64_t)b->sm_count;
        // ... in sparse warps: as many environments per warp as give every SM 16 warps (single-body
        // worlds) or 4 (with pushable blocks: their row storage is 13.5 KB of shared memory per warp).
        // Measured on one B200, 1000 steps: jump.json 4096 environments 7.5e8 env-steps/s in dense warps,
        // 9.6e8 / 9.9e8 / 1.03e9 with 4 / 8 / 16 warps per SM; push.json 4096: 2.6e8 dense, 3.0e8 with 4,
        // 1.6e8 with 16. One environment alone in a warp steps in 3.06 us on jump.json: the floor.
        static const int warps_env = [] { const char* v = getenv("PRL_TINY_WARPS_PER_SM"); int k = v ? atoi(v) : 0; return k; }();
        const int warps_per_sm = warps_env > 0 ? warps_env : (b->host_world.n_bodies == 1 ? 16 : 4);
        int64_t sparse = (b->n + (int64_t)warps_per_sm * b->sm_count - 1) / ((int64_t)warps_per_sm * b->sm_count);
        const int sparse_lanes = (int)(sparse < 1 ? 1 : (sparse > 32 ? 32 : sparse));
        cudaError_t le = cudaSuccess;
#define PRL_ROLLOUT(NB_, WT_, BLOCK_)                                                                                     \
    do {                                                                                                                  \
        const size_t smem = rollout_smem_bytes<NB_, WT_, BLOCK_>();                                                         \
        static unsigned long long configured = 0; /* function attributes are per device */                                \
        const unsigned long long dev_bit = 1ull << (b->device & 63);                                                      \
        if (!(configured & dev_bit)) {                                                                                    \
            le = cudaFuncSetAttribute(rollout_kernel<NB_, WT_, BLOCK_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (le == cudaSuccess) configured |= dev_bit;                                                                 \
        }                                                                                                                 \
        /* environments per warp and per block: dense, or sparse warps (populations too small to fill the GPU). \
           (512-thread blocks -- 16 warps of 28 environments, the whole register file -- measured -2.3 %.) */   \
        int lanes_ = 32, per_block = BLOCK_;                                                                              \
        if (BLOCK_ == kWarpBlock) { lanes_ = sparse_lanes; per_block = sparse_lanes; }                                    \
        if (le == cudaSuccess)                                                                                            \
            rollout_kernel<NB_, WT_, BLOCK_><<<(unsigned)((b->n + per_block - 1) / per_block), BLOCK_, smem, b->stream>>>( \
                b->dev_world, b->state, b->n, dg, genome_len, repeat_move, number_of_steps, early_exit, ds, dw, b->overflow, lanes_, per_block); \
    } while (0)
        switch (pick_variant(b->host_world)) {
            case V1_SMALL: if (tiny) PRL_ROLLOUT(1, DevWorldSmall, kWarpBlock); else PRL_ROLLOUT(1, DevWorldSmall, kBlock); break;
            case V2_SMALL:
                if (big) PRL_ROLLOUT(2, DevWorldSmall, kBigBlock);
                else if (tiny) PRL_ROLLOUT(2, DevWorldSmall, kWarpBlock);
                else PRL_ROLLOUT(2, DevWorldSmall, kBlock);
                break;
            case VN_SMALL: if (big) PRL_ROLLOUT(MIDB, DevWorldSmall, kBigBlock); else PRL_ROLLOUT(MIDB, DevWorldSmall, kBlock); break;
            // large single-body worlds (general tables, x grid: the largest single-body step): SM-wide blocks
            // in lock step, +6.5 % on the synthetic 64-obstacle world (5.79e9 -> 6.17e9); the compact
            // single-body kernel loses 9 % that way on jump.json and stays in 64-thread blocks
            case V1_BIG:
                if (big) PRL_ROLLOUT(1, DevWorld, kBigBlock);
                else if (tiny) PRL_ROLLOUT(1, DevWorld, kWarpBlock);
                else PRL_ROLLOUT(1, DevWorld, kBlock);
                break;
            case VN_BIG: if (big) PRL_ROLLOUT(MIDB, DevWorld, kBigBlock); else PRL_ROLLOUT(MIDB, DevWorld, kBlock); break;
            default: PRL_ROLLOUT(MAXB, DevWorld, kBlock); break;
        }
#undef PRL_ROLLOUT
        CUDA_TRY(le);
    }
    ++b->launches;
    CUDA_TRY(cudaGetLastError());
    // host outputs come straight from the state arrays the kernel also fills
    if (score_host) CUDA_TRY(cudaMemcpyAsync(out_score, b->state + (size_t)F_SCORE * b->n, (size_t)b->n * 4, cudaMemcpyDeviceToHost, b->stream));
    if (win_host) CUDA_TRY(cudaMemcpyAsync(out_step_of_win, b->state + (size_t)F_WIN_STEP * b->n, (size_t)b->n * 4, cudaMemcpyDeviceToHost, b->stream));
    if (score_host || win_host || dg != genomes) CUDA_TRY(cudaStreamSynchronize(b->stream));
    return PRL_OK;
}

int32_t prl_batch_observation_dim(const prl_batch* b) { return b ? policy_obs_dim(b->host_world.n_bodies) : 0; }
int32_t prl_policy_num_params(const prl_batch* b, int32_t hidden) {
    if (!b || hidden < 1 || hidden > kMaxHidden) return -1;
    return policy_num_params(b->host_world.n_bodies, hidden);
}

int prl_batch_get_observation(prl_batch* b, float* out) {
    if (!b || !out) { set_error("prl_batch_get_observation: null argument"); return PRL_ERR_INVALID; }
    DeviceGuard g(b->device);
    const size_t bytes = (size_t)b->n * policy_obs_dim(b->host_world.n_bodies) * 4;
    const bool dev = is_device_ptr(out);
    float* d = out;
    if (!dev) {
        int rc = ensure_stage(&b->stage2, &b->stage2_bytes, bytes);
        if (rc != PRL_OK) return rc;
        d = (float*)b->stage2;
    }
    PRL_DISPATCH(pick_variant(b->host_world), observation_kernel, grid_for(b->n), b->stream, b->dev_world, b->state, b->n, d);
    ++b->launches;
    CUDA_TRY(cudaGetLastError());
    if (!dev) {
        CUDA_TRY(cudaMemcpyAsync(out, d, bytes, cudaMemcpyDeviceToHost, b->stream));
        CUDA_TRY(cudaStreamSynchronize(b->stream));
    }
    return PRL_OK;
}

int prl_batch_rollout_policy(prl_batch* b, const float* params, int32_t hidden, int32_t number_of_steps, int32_t early_exit,
                             float* out_score, int32_t* out_step_of_win) {
    if (!b || !params || hidden < 1 || hidden > kMaxHidden || number_of_steps < 0) {
        set_error("prl_batch_rollout_policy: bad argument (1 <= hidden <= %d)", kMaxHidden);
        return PRL_ERR_INVALID;
    }
    if (b->host_world.n_goals == 0) { set_error("the world has no goal: the rollout's distance_to_goals().unwrap() would panic (main.rs:50)"); return PRL_ERR_NO_GOAL; }
    DeviceGuard g(b->device);
    const size_t pbytes = (size_t)policy_num_params(b->host_world.n_bodies, hidden) * (size_t)b->n * 4;
    const float* dp = params;
    if (!is_device_ptr(params)) {
        int rc = ensure_stage(&b->stage, &b->stage_bytes, pbytes);
        if (rc != PRL_OK) return rc;
        CUDA_TRY(cudaMemcpyAsync(b->stage, params, pbytes, cudaMemcpyHostToDevice, b->stream));
        dp = (const float*)b->stage;
    }
    const bool score_host = out_score && !is_device_ptr(out_score), win_host = out_step_of_win && !is_device_ptr(out_step_of_win);
    PRL_DISPATCH(pick_variant(b->host_world), policy_rollout_kernel, grid_for(b->n), b->stream, b->dev_world, b->state, b->n, dp, hidden,
                 number_of_steps, early_exit, score_host ? nullptr : out_score, win_host ? nullptr : out_step_of_win, b->overflow);
    ++b->launches;
    CUDA_TRY(cudaGetLastError());
    if (score_host) CUDA_TRY(cudaMemcpyAsync(out_score, b->state + (size_t)F_SCORE * b->n, (size_t)b->n * 4, cudaMemcpyDeviceToHost, b->stream));
    if (win_host) CUDA_TRY(cudaMemcpyAsync(out_step_of_win, b->state + (size_t)F_WIN_STEP * b->n, (size_t)b->n * 4, cudaMemcpyDeviceToHost, b->stream));
    if (score_host || win_host || dp != params) CUDA_TRY(cudaStreamSynchronize(b->stream));
    return PRL_OK;
}

int prl_batch_trajectory(prl_batch* b, int64_t n, const uint8_t* genomes, int32_t genome_len, int32_t repeat_move,
                         int32_t number_of_steps, float* out_state, float* out_distance, uint8_t* out_won) {
    if (!b || n <= 0 || genome_len < 0 || (genome_len > 0 && !genomes) || repeat_move < 1 || number_of_steps < 0) {
        set_error("prl_batch_trajectory: bad argument");
        return PRL_ERR_INVALID;
    }
    if (number_of_steps == 0) return PRL_OK;
    DeviceGuard g(b->device);
    const int nb = b->host_world.n_bodies;
    const size_t gbytes = (size_t)n * (size_t)genome_len, recs = (size_t)number_of_steps * (size_t)n;
    const size_t state_bytes = out_state ? recs * nb * 9 * 4 : 0, dist_bytes = out_distance ? recs * 4 : 0, won_bytes = out_won ? recs : 0;
    // one scratch allocation for whatever lives on the host: genomes | state | distance | won
    const bool g_host = gbytes > 0 && !is_device_ptr(genomes);
    const bool s_host = out_state && !is_device_ptr(out_state), d_host = out_distance && !is_device_ptr(out_distance);
    const bool w_host = out_won && !is_device_ptr(out_won);
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t off_s = up(g_host ? gbytes : 0), off_d = off_s + up(s_host ? state_bytes : 0), off_w = off_d + up(d_host ? dist_bytes : 0);
    const size_t total = off_w + up(w_host ? won_bytes : 0);
    if (total > 0) {
        int rc = ensure_stage(&b->traj, &b->traj_bytes, total);
        if (rc != PRL_OK) return rc;
    }
    unsigned char* scratch = (unsigned char*)b->traj;
    const uint8_t* dg = genomes;
    if (g_host) { CUDA_TRY(cudaMemcpyAsync(scratch, genomes, gbytes, cudaMemcpyHostToDevice, b->stream)); dg = scratch; }
    float* ds = s_host ? (float*)(scratch + off_s) : out_state;
    float* dd = d_host ? (float*)(scratch + off_d) : out_distance;
    uint8_t* dwon = w_host ? (uint8_t*)(scratch + off_w) : out_won;
    PRL_DISPATCH(pick_variant(b->host_world), trajectory_kernel, grid_for(n), b->stream, b->dev_world, n, dg, genome_len, repeat_move,
                 number_of_steps, ds, dd, dwon, b->overflow);
    ++b->launches;
    CUDA_TRY(cudaGetLastError());
    if (s_host) CUDA_TRY(cudaMemcpyAsync(out_state, ds, state_bytes, cudaMemcpyDeviceToHost, b->stream));
    if (d_host) CUDA_TRY(cudaMemcpyAsync(out_distance, dd, dist_bytes, cudaMemcpyDeviceToHost, b->stream));
    if (w_host) CUDA_TRY(cudaMemcpyAsync(out_won, dwon, won_bytes, cudaMemcpyDeviceToHost, b->stream));
    if (total > 0) CUDA_TRY(cudaStreamSynchronize(b->stream));
    return PRL_OK;
}

int prl_batch_sync(prl_batch* b) {
    if (!b) { set_error("prl_batch_sync: null batch"); return PRL_ERR_INVALID; }
    DeviceGuard g(b->device);
    int flag = 0;
    CUDA_TRY(cudaMemcpyAsync(&flag, b->overflow, sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    if (flag) {
        set_error("an environment needed more simultaneous contact manifolds than the kernels hold (%d); its trajectory is no longer exact",
                  (b->host_world.n_bodies > 1 && b->host_world.n_pairs > MAXC) ? MAXC_WIDE : MAXC);
        return PRL_ERR_OVERFLOW;
    }
    return PRL_OK;
}

}  // extern "C"
