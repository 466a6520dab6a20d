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
// Multi-GPU. The population is cut into contiguous slices [r*N/R, (r+1)*N/R); the physics needs no
// communication, selection needs every score (the reference draws parents from the WHOLE
// generation, main.rs:134-138) and the genomes of the parents actually drawn. Three ways to run:
//   one shard        prl_ga_create(rank 0 of 1): parents and children are two buffers that swap,
//                    nothing is copied between generations;
//   host exchange    prl_ga_create(rank r of R), one process per GPU: the host all-gathers the
//                    local scores and genomes into the global buffers once per generation
//                    (ncclAllGather through torch.distributed, genetic.py / parallel.py);
//   peer shards      prl_ga_create_multi(devices[]): one process, one shard and one stream per
//                    device. After scoring, every shard pushes its 4 B/agent of scores into its
//                    peers' copies of the score vector (cudaMemcpyPeerAsync over NVLink, ordered
//                    by events, no host synchronisation), and the breeding kernel of every shard
//                    then reads the genomes of the parents it drew STRAIGHT FROM THE OWNING
//                    SHARD'S MEMORY through peer-mapped pointers: the 100 B/agent genome
//                    all-gather (100 MB per generation per GPU at one million agents) is never
//                    made -- each child pulls its two parents, 2 x genome_len bytes, over NVLink.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <new>
#include <vector>

#include "prl_batch.h"

using namespace prl;

namespace {
constexpr int kMaxShards = 16;

// where the breeding kernel finds the genome of parent p: the shard that owns agent p
struct ParentTable {
    const uint8_t* base[kMaxShards];  // genomes of agents [off[k], off[k+1])
    int64_t off[kMaxShards + 1];
    int n;
};

// One slice of the population on one device.
struct GaShard {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int64_t local_off = 0, n_local = 0;
    prl_batch* batch = nullptr;
    uint8_t* genomes[2] = {nullptr, nullptr};  // [cur] = the current (scored) generation of the slice
    int cur = 0;
    float* local_scores = nullptr;      // n_local; peer / single shards: global_scores + local_off
    bool scores_aliased = false;
    uint8_t* global_genomes = nullptr;  // host-exchange mode only: n_global x genome_len
    float* global_scores = nullptr;     // n_global
    unsigned long long* cdf = nullptr;  // n_global inclusive prefix sums of the fixed-point weights
    unsigned long long* block_sums = nullptr;
    unsigned long long* best_key = nullptr;  // [0] = score bits << 32 | index (min), [1] = max score bits
    unsigned long long* steps_sum = nullptr;
    cudaEvent_t scored = nullptr;       // peer mode: my scores have landed in every peer's score vector
    int64_t launches = 0;
};
}  // namespace

struct prl_ga {
    enum Mode { SINGLE, HOST_EXCHANGE, PEER };
    prl_ga_config cfg;
    Mode mode = SINGLE;
    int rank = 0, n_ranks = 1;          // host-exchange mode: my rank among the processes
    int64_t n_global = 0;
    int genome_len = 0;
    std::vector<GaShard> shards;
    int64_t generation = -1;            // -1: prl_ga_init not called yet
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
__device__ __forceinline__ const uint8_t* parent_genome(const ParentTable& pt, int64_t p, int genome_len) {
    int k = 0;
    while (k + 1 < pt.n && p >= pt.off[k + 1]) ++k;
    return pt.base[k] + (p - pt.off[k]) * genome_len;
}
__global__ void breed_kernel(const ParentTable pt, const unsigned long long* __restrict__ cdf, int64_t n_global, int genome_len,
                             uint8_t* __restrict__ children, int64_t n_local, int64_t local_off, float mutation_rate, int keep_best,
                             int64_t best_index, uint64_t seed, uint32_t generation) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_local) return;
    int64_t child = local_off + t;
    uint8_t* out = children + t * genome_len;
    if (keep_best && child == 0) {  // new_generation = vec![min_agent.clone()]
        const uint8_t* src = parent_genome(pt, best_index, genome_len);
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
    // the parents may live on another GPU (peer shards): 2 x genome_len bytes over NVLink
    const uint8_t* g1 = parent_genome(pt, p1, genome_len);
    const uint8_t* g2 = parent_genome(pt, p2, genome_len);
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

int validate_config(const prl_world* w, const prl_ga_config* cfg) {
    if (cfg->number_of_agents < 2 || cfg->repeat_move < 1 || cfg->number_of_steps < 1 || cfg->number_of_steps / cfg->repeat_move < 1) {
        set_error("prl_ga_create: need number_of_agents >= 2 (two distinct parents, main.rs:134-138), repeat_move >= 1 and number_of_steps >= repeat_move");
        return PRL_ERR_INVALID;
    }
    if (cfg->number_of_agents > 0xffffffffll) { set_error("prl_ga_create: number_of_agents exceeds 2^32-1"); return PRL_ERR_INVALID; }
    if (w->dev.n_goals == 0) { set_error("the world has no goal: agent_score's distance_to_goals().unwrap() would panic (main.rs:50)"); return PRL_ERR_NO_GOAL; }
    return PRL_OK;
}

// device memory of one shard; `stream` == nullptr and own_stream: the shard makes its own
int shard_alloc(prl_ga* ga, GaShard& s, const prl_world* w) {
    DeviceGuard g(s.device);
    if (s.own_stream) CUDA_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    int rc = prl_batch_create(w, s.n_local, s.device, s.stream, &s.batch);
    if (rc != PRL_OK) return rc;
    const size_t L = (size_t)ga->genome_len;
    const int64_t nblocks = (ga->n_global + kScanChunk - 1) / kScanChunk;
    cudaError_t e = cudaMalloc(&s.genomes[0], (size_t)s.n_local * L);
    // host-exchange shards breed from the gathered copy into their slice in place (the host keeps
    // pointers to both); the others alternate between two buffers
    if (e == cudaSuccess && ga->mode != prl_ga::HOST_EXCHANGE) e = cudaMalloc(&s.genomes[1], (size_t)s.n_local * L);
    if (e == cudaSuccess) e = cudaMalloc(&s.global_scores, (size_t)ga->n_global * 4);
    if (e == cudaSuccess && ga->mode == prl_ga::HOST_EXCHANGE) {
        e = cudaMalloc(&s.global_genomes, (size_t)ga->n_global * L);
        if (e == cudaSuccess) e = cudaMalloc(&s.local_scores, (size_t)s.n_local * 4);
    } else if (e == cudaSuccess) {
        s.local_scores = s.global_scores + s.local_off;  // my rollouts write their scores in place
        s.scores_aliased = true;
    }
    if (e == cudaSuccess) e = cudaMalloc(&s.cdf, (size_t)ga->n_global * 8);
    if (e == cudaSuccess) e = cudaMalloc(&s.block_sums, (size_t)nblocks * 8);
    if (e == cudaSuccess) e = cudaMalloc(&s.best_key, 16);
    if (e == cudaSuccess) e = cudaMalloc(&s.steps_sum, 8);
    if (e == cudaSuccess && ga->mode == prl_ga::PEER) e = cudaEventCreateWithFlags(&s.scored, cudaEventDisableTiming);
    if (e != cudaSuccess) { set_error("prl_ga_create: device allocation failed on device %d: %s", s.device, cudaGetErrorString(e)); return PRL_ERR_CUDA; }
    return PRL_OK;
}

void shard_free(GaShard& s) {
    {
        DeviceGuard g(s.device);
        if (s.stream || !s.own_stream) cudaStreamSynchronize(s.stream);
        cudaFree(s.genomes[0]); cudaFree(s.genomes[1]); cudaFree(s.global_genomes); cudaFree(s.global_scores);
        if (!s.scores_aliased) cudaFree(s.local_scores);
        cudaFree(s.cdf); cudaFree(s.block_sums); cudaFree(s.best_key); cudaFree(s.steps_sum);
        if (s.scored) cudaEventDestroy(s.scored);
    }
    prl_batch_destroy(s.batch);
    if (s.own_stream && s.stream) { DeviceGuard g(s.device); cudaStreamDestroy(s.stream); }
}

// main.rs:44-70 for every agent of the shard (asynchronous on the shard's stream)
int shard_score(prl_ga* ga, GaShard& s) {
    return prl_batch_rollout(s.batch, s.genomes[s.cur], ga->genome_len, ga->cfg.repeat_move, ga->cfg.number_of_steps, 1, s.local_scores, nullptr);
}

// peer shards: every shard's fresh scores into every other shard's score vector, ordered by
// events: a shard's stream continues only when all slices have arrived in its copy. No host sync.
int exchange_scores(prl_ga* ga) {
    if (ga->mode != prl_ga::PEER) return PRL_OK;
    for (GaShard& s : ga->shards) {
        DeviceGuard g(s.device);
        for (GaShard& d : ga->shards) {
            if (&d == &s) continue;
            CUDA_TRY(cudaMemcpyPeerAsync(d.global_scores + s.local_off, d.device, s.local_scores, s.device, (size_t)s.n_local * 4, s.stream));
        }
        CUDA_TRY(cudaEventRecord(s.scored, s.stream));
    }
    for (GaShard& d : ga->shards) {
        DeviceGuard g(d.device);
        for (GaShard& s : ga->shards)
            if (&d != &s) CUDA_TRY(cudaStreamWaitEvent(d.stream, s.scored, 0));
    }
    return PRL_OK;
}

int score_all(prl_ga* ga) {
    ga->best_valid = false;
    for (GaShard& s : ga->shards) {
        int rc = shard_score(ga, s);
        if (rc != PRL_OK) return rc;
    }
    return exchange_scores(ga);
}

ParentTable parent_table(const prl_ga* ga, const GaShard& me) {
    ParentTable pt;
    memset(&pt, 0, sizeof(pt));
    if (ga->mode == prl_ga::HOST_EXCHANGE) {
        pt.n = 1; pt.base[0] = me.global_genomes; pt.off[0] = 0; pt.off[1] = ga->n_global;
        return pt;
    }
    pt.n = (int)ga->shards.size();
    for (int k = 0; k < pt.n; ++k) { pt.base[k] = ga->shards[k].genomes[ga->shards[k].cur]; pt.off[k] = ga->shards[k].local_off; }
    pt.off[pt.n] = ga->n_global;
    return pt;
}

int create_common(const prl_world* w, const prl_ga_config* cfg, prl_ga::Mode mode, int rank, int n_ranks, const int32_t* devices, int n_shards,
                  void* stream, prl_ga** out) {
    prl_ga* ga = new (std::nothrow) prl_ga();
    if (!ga) { set_error("out of memory"); return PRL_ERR_INVALID; }
    ga->cfg = *cfg;
    ga->mode = mode;
    ga->rank = rank;
    ga->n_ranks = n_ranks;
    ga->n_global = cfg->number_of_agents;
    ga->genome_len = cfg->number_of_steps / cfg->repeat_move;  // main.rs:75
    ga->shards.resize(n_shards);
    const int parts = mode == prl_ga::HOST_EXCHANGE ? n_ranks : n_shards;
    for (int k = 0; k < n_shards; ++k) {
        GaShard& s = ga->shards[k];
        const int r = mode == prl_ga::HOST_EXCHANGE ? rank : k;
        s.device = devices[k];
        s.stream = mode == prl_ga::PEER ? nullptr : (cudaStream_t)stream;
        s.own_stream = mode == prl_ga::PEER;
        s.local_off = ga->n_global * r / parts;
        s.n_local = ga->n_global * (r + 1) / parts - s.local_off;
        if (s.n_local < 1) { set_error("prl_ga_create: slice %d of %d owns no agent", r, parts); prl_ga_destroy(ga); return PRL_ERR_INVALID; }
    }
    for (GaShard& s : ga->shards) {
        int rc = shard_alloc(ga, s, w);
        if (rc != PRL_OK) { prl_ga_destroy(ga); return rc; }
    }
    *out = ga;
    return PRL_OK;
}

}  // namespace

extern "C" {

int prl_ga_create(const prl_world* w, const prl_ga_config* cfg, int32_t device, void* stream, int32_t rank, int32_t n_ranks, prl_ga** out) {
    if (!w || !cfg || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks) { set_error("prl_ga_create: bad argument"); return PRL_ERR_INVALID; }
    int rc = validate_config(w, cfg);
    if (rc != PRL_OK) return rc;
    return create_common(w, cfg, n_ranks == 1 ? prl_ga::SINGLE : prl_ga::HOST_EXCHANGE, rank, n_ranks, &device, 1, stream, out);
}

int prl_ga_create_multi(const prl_world* w, const prl_ga_config* cfg, const int32_t* devices, int32_t n_devices, prl_ga** out) {
    if (!w || !cfg || !out || !devices || n_devices < 1 || n_devices > kMaxShards) {
        set_error("prl_ga_create_multi: bad argument (1..%d devices)", kMaxShards);
        return PRL_ERR_INVALID;
    }
    int rc = validate_config(w, cfg);
    if (rc != PRL_OK) return rc;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device is usable; this library has no CPU path");
        return PRL_ERR_CUDA;
    }
    for (int a = 0; a < n_devices; ++a)
        if (devices[a] < 0 || devices[a] >= ndev) { set_error("prl_ga_create_multi: device %d out of range (%d visible)", devices[a], ndev); return PRL_ERR_INVALID; }
    // every shard reads its peers' genomes and writes their score vectors directly
    for (int a = 0; a < n_devices; ++a) {
        DeviceGuard g(devices[a]);
        for (int b = 0; b < n_devices; ++b) {
            if (devices[a] == devices[b]) continue;
            int can = 0;
            CUDA_TRY(cudaDeviceCanAccessPeer(&can, devices[a], devices[b]));
            if (!can) { set_error("prl_ga_create_multi: device %d cannot map the memory of device %d (no NVLink / PCIe peer path)", devices[a], devices[b]); return PRL_ERR_CUDA; }
            cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else CUDA_TRY(e);
        }
    }
    return create_common(w, cfg, prl_ga::PEER, 0, 1, devices, n_devices, nullptr, out);
}

void prl_ga_destroy(prl_ga* ga) {
    if (!ga) return;
    // nobody may still be reading a peer's memory when it is freed
    for (GaShard& s : ga->shards) { DeviceGuard g(s.device); if (s.stream || !s.own_stream) cudaStreamSynchronize(s.stream); }
    for (GaShard& s : ga->shards) shard_free(s);
    delete ga;
}

int32_t prl_ga_num_shards(const prl_ga* ga) { return ga ? (int32_t)ga->shards.size() : 0; }

int prl_ga_init(prl_ga* ga) {
    if (!ga) { set_error("prl_ga_init: null handle"); return PRL_ERR_INVALID; }
    for (GaShard& s : ga->shards) {
        DeviceGuard g(s.device);
        s.cur = 0;
        int64_t total = s.n_local * ga->genome_len;
        init_genomes_kernel<<<blocks_for(total, kGaBlock), kGaBlock, 0, s.stream>>>(s.genomes[0], s.n_local, s.local_off, ga->genome_len, ga->cfg.seed);
        ++s.launches;
        CUDA_TRY(cudaGetLastError());
    }
    ga->generation = 0;
    return score_all(ga);
}

int prl_ga_buffers(prl_ga* ga, float** local_scores, uint8_t** local_genomes, float** global_scores, uint8_t** global_genomes,
                   int64_t* local_count, int64_t* local_offset, int32_t* genome_len) {
    if (!ga) { set_error("prl_ga_buffers: null handle"); return PRL_ERR_INVALID; }
    if (ga->mode == prl_ga::PEER) { set_error("prl_ga_buffers: a prl_ga_create_multi handle exchanges its generation itself; read it with prl_ga_read_generation"); return PRL_ERR_INVALID; }
    GaShard& s = ga->shards[0];
    if (local_scores) *local_scores = s.local_scores;
    if (local_genomes) *local_genomes = s.genomes[s.cur];
    if (global_scores) *global_scores = s.global_scores;
    if (global_genomes) *global_genomes = ga->mode == prl_ga::SINGLE ? s.genomes[s.cur] : s.global_genomes;
    if (local_count) *local_count = s.n_local;
    if (local_offset) *local_offset = s.local_off;
    if (genome_len) *genome_len = ga->genome_len;
    return PRL_OK;
}

int prl_ga_best(prl_ga* ga, float* best_score, uint8_t* best_genome, int64_t* best_index) {
    if (!ga) { set_error("prl_ga_best: null handle"); return PRL_ERR_INVALID; }
    if (ga->generation < 0) { set_error("prl_ga_best: call prl_ga_init first"); return PRL_ERR_INVALID; }
    if (!ga->best_valid) {
        // every shard reduces its own copy of the score vector (they are identical); all shards are
        // queued before the first is waited for
        unsigned long long keys[kMaxShards][2];
        int over[kMaxShards];
        const unsigned long long init[2] = {~0ull, 0ull};
        for (size_t k = 0; k < ga->shards.size(); ++k) {
            GaShard& s = ga->shards[k];
            DeviceGuard g(s.device);
            CUDA_TRY(cudaMemcpyAsync(s.best_key, init, 16, cudaMemcpyHostToDevice, s.stream));
            unsigned nb = blocks_for(ga->n_global, kGaBlock);
            if (nb > 1184) nb = 1184;  // 148 SMs x 8 resident blocks, grid-stride
            best_kernel<<<nb, kGaBlock, 0, s.stream>>>(s.global_scores, ga->n_global, s.best_key);
            ++s.launches;
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaMemcpyAsync(keys[k], s.best_key, 16, cudaMemcpyDeviceToHost, s.stream));
            CUDA_TRY(cudaMemcpyAsync(&over[k], s.batch->overflow, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
        }
        for (GaShard& s : ga->shards) { DeviceGuard g(s.device); CUDA_TRY(cudaStreamSynchronize(s.stream)); }
        for (size_t k = 0; k < ga->shards.size(); ++k) {
            if (over[k]) {
                set_error("an agent's environment needed more than PRL_MAX_CONTACTS=%d simultaneous contact manifolds: its score is not exact, the generation is not usable", PRL_MAX_CONTACTS);
                return PRL_ERR_OVERFLOW;
            }
            if (keys[k][0] != keys[0][0] || keys[k][1] != keys[0][1]) { set_error("prl_ga_best: the shards disagree on the generation (internal error)"); return PRL_ERR_CUDA; }
        }
        uint32_t bmin = (uint32_t)(keys[0][0] >> 32), bmax = (uint32_t)keys[0][1];
        memcpy(&ga->best_score, &bmin, 4);
        memcpy(&ga->max_score, &bmax, 4);
        ga->best_index = (int64_t)(keys[0][0] & 0xffffffffull);
        ga->best_valid = true;
    }
    if (best_score) *best_score = ga->best_score;
    if (best_index) *best_index = ga->best_index;
    if (best_genome) {
        GaShard& s0 = ga->shards[0];
        const ParentTable pt = parent_table(ga, s0);
        int k = 0;
        while (k + 1 < pt.n && ga->best_index >= pt.off[k + 1]) ++k;
        GaShard& owner = ga->mode == prl_ga::PEER ? ga->shards[k] : s0;
        DeviceGuard g(owner.device);
        CUDA_TRY(cudaMemcpyAsync(best_genome, pt.base[k] + (size_t)(ga->best_index - pt.off[k]) * ga->genome_len, (size_t)ga->genome_len,
                                 cudaMemcpyDefault, owner.stream));
        CUDA_TRY(cudaStreamSynchronize(owner.stream));
    }
    return PRL_OK;
}

int prl_ga_next_generation(prl_ga* ga) {
    if (!ga) { set_error("prl_ga_next_generation: null handle"); return PRL_ERR_INVALID; }
    int rc = prl_ga_best(ga, nullptr, nullptr, nullptr);
    if (rc != PRL_OK) return rc;
    const int64_t nblocks = (ga->n_global + kScanChunk - 1) / kScanChunk;
    ++ga->generation;
    // breeding reads the current buffers of every shard, so all tables are taken before any swap
    ParentTable tables[kMaxShards];
    for (size_t k = 0; k < ga->shards.size(); ++k) tables[k] = parent_table(ga, ga->shards[k]);
    for (size_t k = 0; k < ga->shards.size(); ++k) {
        GaShard& s = ga->shards[k];
        DeviceGuard g(s.device);
        weight_block_sums_kernel<<<(unsigned)nblocks, kGaBlock, 0, s.stream>>>(s.global_scores, ga->n_global, ga->max_score, s.block_sums);
        scan_block_sums_kernel<<<1, 32, 0, s.stream>>>(s.block_sums, nblocks);
        weight_scan_kernel<<<(unsigned)nblocks, kGaBlock, 0, s.stream>>>(s.global_scores, ga->n_global, ga->max_score, s.block_sums, s.cdf);
        uint8_t* children = ga->mode == prl_ga::HOST_EXCHANGE ? s.genomes[0] : s.genomes[s.cur ^ 1];
        breed_kernel<<<blocks_for(s.n_local, kGaBlock), kGaBlock, 0, s.stream>>>(
            tables[k], s.cdf, ga->n_global, ga->genome_len, children, s.n_local, s.local_off, ga->cfg.mutation_rate, ga->cfg.keep_best,
            ga->best_index, ga->cfg.seed, (uint32_t)ga->generation);
        s.launches += 4;
        CUDA_TRY(cudaGetLastError());
        if (ga->mode == prl_ga::PEER) CUDA_TRY(cudaEventRecord(s.scored, s.stream));  // here: "my breeding has read its parents"
    }
    if (ga->mode == prl_ga::PEER) {
        // a shard's rollouts overwrite its slice of every score vector, and two generations on its
        // next breeding overwrites the buffer its peers are reading parents from now: both wait
        // until every shard has bred
        for (GaShard& d : ga->shards) {
            DeviceGuard g(d.device);
            for (GaShard& s : ga->shards)
                if (&d != &s) CUDA_TRY(cudaStreamWaitEvent(d.stream, s.scored, 0));
        }
    }
    if (ga->mode != prl_ga::HOST_EXCHANGE)
        for (GaShard& s : ga->shards) s.cur ^= 1;
    return score_all(ga);
}

int prl_ga_step_generation(prl_ga* ga, float* best_score, uint8_t* best_genome, int64_t* best_index) {
    if (!ga) { set_error("prl_ga_step_generation: null handle"); return PRL_ERR_INVALID; }
    if (ga->mode == prl_ga::HOST_EXCHANGE) { set_error("prl_ga_step_generation: a rank of a multi-process GA exchanges through the host between prl_ga_best and prl_ga_next_generation"); return PRL_ERR_INVALID; }
    int rc = ga->generation < 0 ? prl_ga_init(ga) : PRL_OK;
    if (rc == PRL_OK) rc = prl_ga_best(ga, best_score, best_genome, best_index);
    if (rc == PRL_OK) rc = prl_ga_next_generation(ga);
    return rc;
}

int prl_ga_read_generation(prl_ga* ga, float* scores, uint8_t* genomes) {
    if (!ga) { set_error("prl_ga_read_generation: null handle"); return PRL_ERR_INVALID; }
    if (ga->generation < 0) { set_error("prl_ga_read_generation: call prl_ga_init first"); return PRL_ERR_INVALID; }
    if (ga->mode == prl_ga::HOST_EXCHANGE) { set_error("prl_ga_read_generation: a rank of a multi-process GA holds the gathered generation in its prl_ga_buffers"); return PRL_ERR_INVALID; }
    for (GaShard& s : ga->shards) {
        DeviceGuard g(s.device);
        if (scores) CUDA_TRY(cudaMemcpyAsync(scores + s.local_off, s.local_scores, (size_t)s.n_local * 4, cudaMemcpyDefault, s.stream));
        if (genomes) CUDA_TRY(cudaMemcpyAsync(genomes + (size_t)s.local_off * ga->genome_len, s.genomes[s.cur], (size_t)s.n_local * ga->genome_len, cudaMemcpyDefault, s.stream));
    }
    for (GaShard& s : ga->shards) { DeviceGuard g(s.device); CUDA_TRY(cudaStreamSynchronize(s.stream)); }
    return PRL_OK;
}

int64_t prl_ga_generation(const prl_ga* ga) { return ga ? (ga->generation < 0 ? 0 : ga->generation) : 0; }
int64_t prl_ga_launch_count(const prl_ga* ga) {
    if (!ga) return 0;
    int64_t n = 0;
    for (const GaShard& s : ga->shards) n += s.launches + prl_batch_launch_count(s.batch);
    return n;
}

int64_t prl_ga_env_steps(prl_ga* ga) {
    if (!ga) return -1;
    unsigned long long sums[kMaxShards];
    for (size_t k = 0; k < ga->shards.size(); ++k) {
        GaShard& s = ga->shards[k];
        DeviceGuard g(s.device);
        if (cudaMemsetAsync(s.steps_sum, 0, 8, s.stream) != cudaSuccess) return -1;
        unsigned nb = blocks_for(s.n_local, kGaBlock);
        if (nb > 1184) nb = 1184;
        sum_steps_kernel<<<nb, kGaBlock, 0, s.stream>>>(s.batch->state + (size_t)F_STEPS * s.batch->n, s.n_local, s.steps_sum);
        ++s.launches;
        if (cudaMemcpyAsync(&sums[k], s.steps_sum, 8, cudaMemcpyDeviceToHost, s.stream) != cudaSuccess) return -1;
    }
    int64_t total = 0;
    for (size_t k = 0; k < ga->shards.size(); ++k) {
        DeviceGuard g(ga->shards[k].device);
        if (cudaStreamSynchronize(ga->shards[k].stream) != cudaSuccess) return -1;
        total += (int64_t)sums[k];
    }
    return total;
}

}  // extern "C"
