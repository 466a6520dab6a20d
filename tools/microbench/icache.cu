// Instruction-cache capacity probe for sm_100a: a loop whose body is KB kilobytes of straight-line
// independent FADDs (4 chains), run by 448-thread blocks (14 warps per SM, one block per SM) with and
// without a block barrier per iteration. Prints cycles per instruction per warp: flat while the body
// fits the SM's instruction cache, rising beyond. profiles/r02_experiments.md quotes the result.
#include <cstdio>
#include <cuda_runtime.h>

template <int KB, bool BARRIER>
__global__ void __launch_bounds__(448, 1) body(float* out, int iters, long long* cycles) {
    float p0 = threadIdx.x, p1 = 1.0f, p2 = 2.0f, p3 = 3.0f;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if (BARRIER) __syncthreads();
#pragma unroll
        for (int q = 0; q < KB * 16; ++q) {
            asm volatile("add.f32 %0, %0, 0f3F800000;" : "+f"(p0));
            asm volatile("add.f32 %0, %0, 0f3F800000;" : "+f"(p1));
            asm volatile("add.f32 %0, %0, 0f3F800000;" : "+f"(p2));
            asm volatile("add.f32 %0, %0, 0f3F800000;" : "+f"(p3));
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (p0 + p1 + p2 + p3 == -1.0f) out[0] = p0;
}

template <int KB, bool BARRIER>
void run(float* out, long long* dcyc) {
    const int iters = 2000;
    body<KB, BARRIER><<<148, 448>>>(out, 10, dcyc);
    cudaDeviceSynchronize();
    body<KB, BARRIER><<<148, 448>>>(out, iters, dcyc);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, dcyc, sizeof h, cudaMemcpyDeviceToHost);
    double mx = 0;
    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    const double instr = (double)KB * 64;
    printf("body %3d KB  barrier %d  cycles/iteration %9.0f  cycles per instruction per warp %.3f  (14 warps: SM issues %.2f inst/cycle)\n", KB,
           (int)BARRIER, mx / iters, mx / iters / instr, 14.0 * instr / (mx / iters));
}

int main() {
    float* out; long long* dcyc;
    cudaMalloc(&out, 4); cudaMalloc(&dcyc, 148 * 8);
    run<8, true>(out, dcyc); run<16, true>(out, dcyc); run<24, true>(out, dcyc); run<32, true>(out, dcyc); run<40, true>(out, dcyc);
    run<48, true>(out, dcyc); run<64, true>(out, dcyc); run<96, true>(out, dcyc); run<128, true>(out, dcyc);
    run<8, false>(out, dcyc); run<16, false>(out, dcyc); run<24, false>(out, dcyc); run<32, false>(out, dcyc); run<40, false>(out, dcyc);
    run<48, false>(out, dcyc); run<64, false>(out, dcyc); run<96, false>(out, dcyc); run<128, false>(out, dcyc);
    return 0;
}
