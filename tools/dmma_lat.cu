// DMMA (mma.sync.m8n8k4.f64) issue/latency microbenchmark for B200: throughput against the number of independent
// accumulator chains per warp and warps per sub-partition, with and without interleaved SHFL/FSEL (the transposes of
// csrc/blocked.cuh).  Prints one JSON object.  Not part of the product path.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 2048;
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int CH, int NSH>
__global__ void k(double* out, double a, double b) {
    double c0[CH], c1[CH], s = threadIdx.x;
#pragma unroll
    for (int i = 0; i < CH; i++) { c0[i] = i; c1[i] = -i; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            dmma(c0[i], c1[i], a, b);
#pragma unroll
            for (int q = 0; q < NSH; q++) s = __shfl_xor_sync(0xffffffffu, (threadIdx.x & 4) ? s : -s, 1 + q);
        }
    }
    double r = s;
#pragma unroll
    for (int i = 0; i < CH; i++) r += c0[i] + c1[i];
    if (r == 123.456) out[0] = r;
}
template <int CH, int NSH>
static void run(const char* name, int warps) {
    double* out; cudaMalloc(&out, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148;
    k<CH, NSH><<<grid, warps * 32>>>(out, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    k<CH, NSH><<<grid, warps * 32>>>(out, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * 1.965e9;
    // cycles per DMMA per sub-partition: warps/4 warps per SMSP each issuing ITERS*CH
    const double per = cyc / ((double)ITERS * CH * (warps / 4.0));
    printf("\"%s_w%d\": %.2f, ", name, warps, per);
    cudaFree(out);
}
int main() {
    printf("{\"unit\": \"cycles per DMMA.8x8x4 per sub-partition (16 = pipe peak)\", ");
    run<1, 0>("ch1", 4); run<2, 0>("ch2", 4); run<3, 0>("ch3", 4); run<4, 0>("ch4", 4); run<8, 0>("ch8", 4);
    run<1, 0>("ch1", 8); run<2, 0>("ch2", 8); run<4, 0>("ch4", 8);
    run<1, 0>("ch1", 12); run<2, 0>("ch2", 12);
    run<4, 1>("ch4_shfl1", 4); run<4, 2>("ch4_shfl2", 4); run<4, 4>("ch4_shfl4", 4);
    run<4, 1>("ch4_shfl1", 8); run<4, 2>("ch4_shfl2", 8); run<4, 4>("ch4_shfl4", 8);
    printf("\"clock_ghz\": 1.965}\n");
    return 0;
}
