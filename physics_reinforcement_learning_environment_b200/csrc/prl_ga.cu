// prl_ga.cu -- the genetic algorithm of the reference's src/main.rs:41-163 on the device.
//
// The reference evaluates agents serially and draws two parents per child with rand 0.8.5's
// `choose_multiple_weighted` (Efraimidis-Spirakis keys over the whole generation: O(N) per child,
// O(N^2) per generation). Here one generation is:
//   best      min (first index on ties, Iterator::min_by) and max of the global scores
//   weights   w_i = max_score + 1 - score_i (main.rs:135) in f32, then fixed point (2^-20) so that
//             their inclusive prefix sum is exact integer arithmetic: the same on any GPU count
//             and in any reduction order
//   breed     one thread per child: parent 1 by inverse-CDF binary search, parent 2 likewise with
//             rejection of parent 1 (the same distribution as weighted sampling of 2 without
//             replacement), uniform crossover per gene, per-bit mutation (main.rs:140-158),
//             elitism (main.rs:126-131)
//   score     the fused rollout kernel on the children (main.rs:44-70, early exit on win)
// Random numbers are Philox4x32-10 keyed by the seed with counter (child index, generation, draw),
// so a child's genome does not depend on how the population is sharded over GPUs. (The reference
// itself is not reproducible: thread_rng, main.rs:42.)
//
// Multi-GPU: each rank owns the slice [rank*N/R, (rank+1)*N/R) of the population. After scoring,
// the host all-gathers the local scores and genomes into the global buffers (NCCL over NVLink,
// see prl_ga_buffers); every rank then computes the same best / prefix sums and breeds only its
// own children. The physics itself needs no communication.
#include <cuda_runtime.h>
#include <stdint.h>

#include <new>

#include "prl_batch.h"

using namespace prl;

struct prl_ga {
    prl_ga_config cfg;
    int device = 0;
    cudaStream_t stream = nullptr;
    int rank = 0, n_ranks = 1;
    int64_t n_global = 0, local_off = 0, n_local = 0;
    int genome_len = 0;
    prl_batch* batch = nullptr;
    uint8_t* local_genomes = nullptr;   // n_local x genome_len
    float* local_scores = nullptr;      // n_local
    uint8_t* global_genomes = nullptr;  // n_global x genome_len (parents)
    float* global_scores = nullptr;     // n_global
    unsigned long long* cdf = nullptr;  // n_global inclusive prefix sums of the fixed-point weights
    unsigned long long* block_sums = nullptr;
    unsigned long long* best_key = nullptr;  // [0] = score bits << 32 | index (min), [1] = max score bits
    unsigned long long* steps_sum = nullptr;
    int64_t generation = 0;
    int64_t launches = 0;
    bool best_valid = false;
    float best_score = 0.0f, max_score = 0.0f;
    int64_t best_index = 0;
};

namespace {

constexpr int kGaBlock = 256;
constexpr int kScanChunk = 1024;  // elements per scan block

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
// Philox4x32-10 (Salmon et al., SC'11)
__device__ __forceinline__ void philox4x32(uint64_t seed, uint64_t index, uint32_t generation, uint32_t draw, uint32_t (&out)[4]) {
    uint32_t c[4] = {(uint32_t)index, (uint32_t)(index >> 32), generation, draw};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

// draws: 0..genome_len-1 = per-gene words, then parent draws
constexpr uint32_t kDrawParents = 0x40000000u;

// main.rs:72-82: every gene is three fair coin flips
__global__ void init_genomes_kernel(uint8_t* __restrict__ genomes, int64_t n_local, int64_t local_off, int genome_len, uint64_t seed) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_local * genome_len) return;
    int64_t a = t / genome_len;
    int g = (int)(t % genome_len);
    uint32_t r[4];
    philox4x32(seed, (uint64_t)(local_off + a), 0u, (uint32_t)g, r);
    genomes[t] = (uint8_t)((r[0] >> 31) | ((r[1] >> 31) << 1) | ((r[2] >> 31) << 2));
}

// main.rs:87-111: min_by keeps the first minimum; scores are >= 0 (or +inf), so their bit
// patterns order like the floats and (bits << 32 | index) orders like (score, index).
__global__ void best_kernel(const float* __restrict__ scores, int64_t n, unsigned long long* __restrict__ out) {
    unsigned long long kmin = ~0ull;
    unsigned int kmax = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        unsigned int bits = __float_as_uint(scores[i]);
        unsigned long long key = ((unsigned long long)bits << 32) | (unsigned long long)(uint32_t)i;
        kmin = key < kmin ? key : kmin;
        kmax = bits > kmax ? bits : kmax;
    }
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, kmin, o);
        kmin = other < kmin ? other : kmin;
        unsigned int om = __shfl_xor_sync(0xffffffffu, kmax, o);
        kmax = om > kmax ? om : kmax;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&out[0], kmin);
        atomicMax(&out[1], (unsigned long long)kmax);
    }
}

// main.rs:135: weight of an agent, as 44.20 fixed point (scores are pixel distances, far below 2^24)
__device__ __forceinline__ unsigned long long fixed_weight(float max_score, float score) {
    float w = max_score + 1.0f - score;
    if (!(w > 0.0f)) return 0ull;  // +inf scores (no step taken) and NaN get no offspring
    return (unsigned long long)llrint((double)w * 1048576.0);
}
// per-chunk sums of the weights
__global__ void weight_block_sums_kernel(const float* __restrict__ scores, int64_t n, float max_score, unsigned long long* __restrict__ block_sums) {
    __shared__ unsigned long long warp_sums[kGaBlock / 32];
    int64_t base = (int64_t)blockIdx.x * kScanChunk;
    unsigned long long s = 0;
    for (int j = threadIdx.x; j < kScanChunk; j += kGaBlock) {
        int64_t i = base + j;
        if (i < n) s += fixed_weight(max_score, scores[i]);
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < kGaBlock / 32; ++w) t += warp_sums[w];
        block_sums[blockIdx.x] = t;
    }
}
// exclusive scan of the chunk sums, in place, by one thread (<= N/1024 values)
__global__ void scan_block_sums_kernel(unsigned long long* __restrict__ block_sums, int64_t nblocks) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int64_t b = 0; b < nblocks; ++b) { unsigned long long v = block_sums[b]; block_sums[b] = run; run += v; }
    }
}
// inclusive prefix sums inside each chunk + the chunk's offset
__global__ void weight_scan_kernel(const float* __restrict__ scores, int64_t n, float max_score, const unsigned long long* __restrict__ block_offsets,
                                   unsigned long long* __restrict__ cdf) {
    __shared__ unsigned long long warp_tot[kGaBlock / 32];
    int64_t base = (int64_t)blockIdx.x * kScanChunk;
    constexpr int per = kScanChunk / kGaBlock;  // consecutive elements per thread
    unsigned long long v[per];
    unsigned long long run = 0;
#pragma unroll
    for (int j = 0; j < per; ++j) {
        int64_t i = base + (int64_t)threadIdx.x * per + j;
        run += i < n ? fixed_weight(max_score, scores[i]) : 0ull;
        v[j] = run;
    }
    unsigned long long incl = run;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    unsigned long long off = block_offsets[blockIdx.x];
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    off += incl - run;
#pragma unroll
    for (int j = 0; j < per; ++j) {
        int64_t i = base + (int64_t)threadIdx.x * per + j;
        if (i < n) cdf[i] = off + v[j];
    }
}

// weighted pick by inverse CDF: a 64-bit random word scaled to [0, total), then the first index
// whose inclusive prefix sum exceeds it
__device__ __forceinline__ int64_t pick(const unsigned long long* __restrict__ cdf, int64_t n, unsigned long long total, uint32_t hi, uint32_t lo) {
    unsigned long long target = __umul64hi(((unsigned long long)hi << 32) | lo, total);
    int64_t a = 0, b = n - 1;
    while (a < b) {
        int64_t m = (a + b) >> 1;
        if (cdf[m] > target) b = m; else a = m + 1;
    }
    return a;
}

// main.rs:126-158 for the children [local_off, local_off + n_local)
__global__ void breed_kernel(const uint8_t* __restrict__ parents, const unsigned long long* __restrict__ cdf, int64_t n_global, int genome_len,
                             uint8_t* __restrict__ children, int64_t n_local, int64_t local_off, float mutation_rate, int keep_best,
                             int64_t best_index, uint64_t seed, uint32_t generation) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_local) return;
    int64_t child = local_off + t;
    uint8_t* out = children + t * genome_len;
    if (keep_best && child == 0) {  // new_generation = vec![min_agent.clone()]
        const uint8_t* src = parents + best_index * genome_len;
        for (int g = 0; g < genome_len; ++g) out[g] = src[g];
        return;
    }
    unsigned long long total = cdf[n_global - 1];
    uint32_t r[4];
    philox4x32(seed, (uint64_t)child, generation, kDrawParents, r);
    int64_t p1 = pick(cdf, n_global, total, r[0], r[1]);
    int64_t p2 = pick(cdf, n_global, total, r[2], r[3]);
    for (uint32_t attempt = 1; p2 == p1 && n_global > 1 && attempt < 64; ++attempt) {
        philox4x32(seed, (uint64_t)child, generation, kDrawParents + attempt, r);
        p2 = pick(cdf, n_global, total, r[0], r[1]);
        if (p2 == p1) p2 = pick(cdf, n_global, total, r[2], r[3]);
    }
    if (p2 == p1 && n_global > 1) p2 = (p1 + 1) % n_global;
    const uint8_t* g1 = parents + p1 * genome_len;
    const uint8_t* g2 = parents + p2 * genome_len;
    for (int g = 0; g < genome_len; ++g) {
        philox4x32(seed, (uint64_t)child, generation, (uint32_t)g, r);
        uint32_t mv = (r[0] >> 31) ? g1[g] : g2[g];  // rng.gen::<bool>()
        // rng.gen::<f32>() < mutation_rate: 24 random bits scaled to [0,1); the re-drawn bit
        // comes from the low byte of the same word, which the uniform does not use
#pragma unroll
        for (int bit = 0; bit < 3; ++bit) {
            float u = (float)(r[1 + bit] >> 8) * (1.0f / 16777216.0f);
            if (u < mutation_rate) mv = (mv & ~(1u << bit)) | ((r[1 + bit] & 1u) << bit);
        }
        out[g] = (uint8_t)(mv & 7u);
    }
}

__global__ void sum_steps_kernel(const uint32_t* __restrict__ steps, int64_t n, unsigned long long* __restrict__ out) {
    unsigned long long s = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s += steps[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

inline unsigned blocks_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

int score_local(prl_ga* ga) {
    int rc = prl_batch_rollout(ga->batch, ga->local_genomes, ga->genome_len, ga->cfg.repeat_move, ga->cfg.number_of_steps, 1,
                               ga->local_scores, nullptr);
    ga->best_valid = false;
    return rc;
}

}  // namespace

extern "C" {

int prl_ga_create(const prl_world* w, const prl_ga_config* cfg, int32_t device, void* stream, int32_t rank, int32_t n_ranks, prl_ga** out) {
    if (!w || !cfg || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks) { set_error("prl_ga_create: bad argument"); return PRL_ERR_INVALID; }
    if (cfg->number_of_agents < 2 || cfg->repeat_move < 1 || cfg->number_of_steps < 1 || cfg->number_of_steps / cfg->repeat_move < 1) {
        set_error("prl_ga_create: need number_of_agents >= 2 (two distinct parents, main.rs:134-138), repeat_move >= 1 and number_of_steps >= repeat_move");
        return PRL_ERR_INVALID;
    }
    if (cfg->number_of_agents > 0xffffffffll) { set_error("prl_ga_create: number_of_agents exceeds 2^32-1"); return PRL_ERR_INVALID; }
    if (w->dev.n_goals == 0) { set_error("the world has no goal: agent_score's distance_to_goals().unwrap() would panic (main.rs:50)"); return PRL_ERR_NO_GOAL; }
    prl_ga* ga = new (std::nothrow) prl_ga();
    if (!ga) { set_error("out of memory"); return PRL_ERR_INVALID; }
    ga->cfg = *cfg;
    ga->device = device;
    ga->stream = (cudaStream_t)stream;
    ga->rank = rank;
    ga->n_ranks = n_ranks;
    ga->n_global = cfg->number_of_agents;
    ga->local_off = ga->n_global * rank / n_ranks;
    ga->n_local = ga->n_global * (rank + 1) / n_ranks - ga->local_off;
    ga->genome_len = cfg->number_of_steps / cfg->repeat_move;  // main.rs:75
    if (ga->n_local < 1) { set_error("prl_ga_create: rank %d of %d owns no agent", rank, n_ranks); delete ga; return PRL_ERR_INVALID; }
    int rc = prl_batch_create(w, ga->n_local, device, stream, &ga->batch);
    if (rc != PRL_OK) { delete ga; return rc; }
    DeviceGuard g(device);
    size_t L = (size_t)ga->genome_len;
    int64_t nblocks = (ga->n_global + kScanChunk - 1) / kScanChunk;
    cudaError_t e = cudaMalloc(&ga->local_genomes, (size_t)ga->n_local * L);
    if (e == cudaSuccess) e = cudaMalloc(&ga->local_scores, (size_t)ga->n_local * 4);
    if (e == cudaSuccess) e = cudaMalloc(&ga->global_genomes, (size_t)ga->n_global * L);
    if (e == cudaSuccess) e = cudaMalloc(&ga->global_scores, (size_t)ga->n_global * 4);
    if (e == cudaSuccess) e = cudaMalloc(&ga->cdf, (size_t)ga->n_global * 8);
    if (e == cudaSuccess) e = cudaMalloc(&ga->block_sums, (size_t)nblocks * 8);
    if (e == cudaSuccess) e = cudaMalloc(&ga->best_key, 16);
    if (e == cudaSuccess) e = cudaMalloc(&ga->steps_sum, 8);
    if (e != cudaSuccess) {
        set_error("prl_ga_create: cudaMalloc failed: %s", cudaGetErrorString(e));
        prl_ga_destroy(ga);
        return PRL_ERR_CUDA;
    }
    *out = ga;
    return PRL_OK;
}

void prl_ga_destroy(prl_ga* ga) {
    if (!ga) return;
    {
        DeviceGuard g(ga->device);
        cudaStreamSynchronize(ga->stream);
        cudaFree(ga->local_genomes); cudaFree(ga->local_scores); cudaFree(ga->global_genomes); cudaFree(ga->global_scores);
        cudaFree(ga->cdf); cudaFree(ga->block_sums); cudaFree(ga->best_key); cudaFree(ga->steps_sum);
    }
    prl_batch_destroy(ga->batch);
    delete ga;
}

int prl_ga_init(prl_ga* ga) {
    if (!ga) { set_error("prl_ga_init: null handle"); return PRL_ERR_INVALID; }
    DeviceGuard g(ga->device);
    int64_t total = ga->n_local * ga->genome_len;
    init_genomes_kernel<<<blocks_for(total, kGaBlock), kGaBlock, 0, ga->stream>>>(ga->local_genomes, ga->n_local, ga->local_off, ga->genome_len, ga->cfg.seed);
    ++ga->launches;
    CUDA_TRY(cudaGetLastError());
    ga->generation = 0;
    return score_local(ga);
}

int prl_ga_buffers(prl_ga* ga, float** local_scores, uint8_t** local_genomes, float** global_scores, uint8_t** global_genomes,
                   int64_t* local_count, int64_t* local_offset, int32_t* genome_len) {
    if (!ga) { set_error("prl_ga_buffers: null handle"); return PRL_ERR_INVALID; }
    if (local_scores) *local_scores = ga->local_scores;
    if (local_genomes) *local_genomes = ga->local_genomes;
    if (global_scores) *global_scores = ga->global_scores;
    if (global_genomes) *global_genomes = ga->global_genomes;
    if (local_count) *local_count = ga->n_local;
    if (local_offset) *local_offset = ga->local_off;
    if (genome_len) *genome_len = ga->genome_len;
    return PRL_OK;
}

int prl_ga_best(prl_ga* ga, float* best_score, uint8_t* best_genome, int64_t* best_index) {
    if (!ga) { set_error("prl_ga_best: null handle"); return PRL_ERR_INVALID; }
    DeviceGuard g(ga->device);
    if (!ga->best_valid) {
        if (ga->n_ranks == 1) {
            // a single rank is its own all-gather
            CUDA_TRY(cudaMemcpyAsync(ga->global_scores, ga->local_scores, (size_t)ga->n_local * 4, cudaMemcpyDeviceToDevice, ga->stream));
            CUDA_TRY(cudaMemcpyAsync(ga->global_genomes, ga->local_genomes, (size_t)ga->n_local * ga->genome_len, cudaMemcpyDeviceToDevice, ga->stream));
        }
        unsigned long long init[2] = {~0ull, 0ull};
        CUDA_TRY(cudaMemcpyAsync(ga->best_key, init, 16, cudaMemcpyHostToDevice, ga->stream));
        unsigned nb = blocks_for(ga->n_global, kGaBlock);
        if (nb > 1184) nb = 1184;  // 148 SMs x 8 resident blocks, grid-stride
        best_kernel<<<nb, kGaBlock, 0, ga->stream>>>(ga->global_scores, ga->n_global, ga->best_key);
        ++ga->launches;
        CUDA_TRY(cudaGetLastError());
        unsigned long long keys[2];
        CUDA_TRY(cudaMemcpyAsync(keys, ga->best_key, 16, cudaMemcpyDeviceToHost, ga->stream));
        CUDA_TRY(cudaStreamSynchronize(ga->stream));
        uint32_t bmin = (uint32_t)(keys[0] >> 32), bmax = (uint32_t)keys[1];
        memcpy(&ga->best_score, &bmin, 4);
        memcpy(&ga->max_score, &bmax, 4);
        ga->best_index = (int64_t)(keys[0] & 0xffffffffull);
        ga->best_valid = true;
    }
    if (best_score) *best_score = ga->best_score;
    if (best_index) *best_index = ga->best_index;
    if (best_genome) {
        CUDA_TRY(cudaMemcpyAsync(best_genome, ga->global_genomes + (size_t)ga->best_index * ga->genome_len, (size_t)ga->genome_len,
                                 cudaMemcpyDefault, ga->stream));
        CUDA_TRY(cudaStreamSynchronize(ga->stream));
    }
    return PRL_OK;
}

int prl_ga_next_generation(prl_ga* ga) {
    if (!ga) { set_error("prl_ga_next_generation: null handle"); return PRL_ERR_INVALID; }
    int rc = prl_ga_best(ga, nullptr, nullptr, nullptr);
    if (rc != PRL_OK) return rc;
    DeviceGuard g(ga->device);
    int64_t nblocks = (ga->n_global + kScanChunk - 1) / kScanChunk;
    weight_block_sums_kernel<<<(unsigned)nblocks, kGaBlock, 0, ga->stream>>>(ga->global_scores, ga->n_global, ga->max_score, ga->block_sums);
    scan_block_sums_kernel<<<1, 32, 0, ga->stream>>>(ga->block_sums, nblocks);
    weight_scan_kernel<<<(unsigned)nblocks, kGaBlock, 0, ga->stream>>>(ga->global_scores, ga->n_global, ga->max_score, ga->block_sums, ga->cdf);
    ++ga->generation;
    breed_kernel<<<blocks_for(ga->n_local, kGaBlock), kGaBlock, 0, ga->stream>>>(
        ga->global_genomes, ga->cdf, ga->n_global, ga->genome_len, ga->local_genomes, ga->n_local, ga->local_off, ga->cfg.mutation_rate,
        ga->cfg.keep_best, ga->best_index, ga->cfg.seed, (uint32_t)ga->generation);
    ga->launches += 4;
    CUDA_TRY(cudaGetLastError());
    return score_local(ga);
}

int64_t prl_ga_generation(const prl_ga* ga) { return ga ? ga->generation : 0; }
int64_t prl_ga_launch_count(const prl_ga* ga) { return ga ? ga->launches + prl_batch_launch_count(ga->batch) : 0; }

int64_t prl_ga_env_steps(prl_ga* ga) {
    if (!ga) return -1;
    DeviceGuard g(ga->device);
    if (cudaMemsetAsync(ga->steps_sum, 0, 8, ga->stream) != cudaSuccess) return -1;
    unsigned nb = blocks_for(ga->n_local, kGaBlock);
    if (nb > 1184) nb = 1184;
    sum_steps_kernel<<<nb, kGaBlock, 0, ga->stream>>>(ga->batch->state + (size_t)F_STEPS * ga->batch->n, ga->n_local, ga->steps_sum);
    ++ga->launches;
    unsigned long long s = 0;
    if (cudaMemcpyAsync(&s, ga->steps_sum, 8, cudaMemcpyDeviceToHost, ga->stream) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(ga->stream) != cudaSuccess) return -1;
    return (int64_t)s;
}

}  // extern "C"
