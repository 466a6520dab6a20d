// prl_batch.h -- the object behind a `prl_batch*` handle (include/prl.h), shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "prl_host.h"

#define CUDA_TRY(expr)                                                                       \
    do {                                                                                     \
        cudaError_t err__ = (expr);                                                          \
        if (err__ != cudaSuccess) {                                                          \
            prl::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
            return PRL_ERR_CUDA;                                                             \
        }                                                                                    \
    } while (0)

struct prl_batch {
    const prl_world* world_src = nullptr;
    prl::DevWorld host_world;
    prl::DevWorld* dev_world = nullptr;
    uint32_t* state = nullptr;     // SoA words, [state_words][n]
    int* overflow = nullptr;       // device flag: some environment ran out of manifold slots
    int64_t n = 0;
    int words = 0;
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    // staging for host-side buffers
    void* stage = nullptr;
    size_t stage_bytes = 0;
    void* stage2 = nullptr;
    size_t stage2_bytes = 0;
    // pinned host buffer for the per-step moves of a closed loop on the host: the caller's array is
    // copied here and the transfer is asynchronous, so prl_batch_step does not wait for the stream
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    cudaEvent_t pinned_free = nullptr;  // the last transfer out of `pinned` has completed
    // scratch of prl_batch_trajectory (kept between calls)
    void* traj = nullptr;
    size_t traj_bytes = 0;
    int64_t launches = 0;
};

namespace prl {
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
}  // namespace prl
