// FP64 FMA throughput vs the number of "fresh" register operands per instruction (no spills: 32 x, 32 y, 32 z, 4 u).  (not product code)
// MODE 0: x = fma(x, a, b)            1 fresh
// MODE 1: x[i] = fma(y[i], u[i/8], x[i])   2 fresh + 1 shared by 8 consecutive instructions
// MODE 2: x[i] = fma(y[i], z[i], x[i])     3 fresh
// MODE 3: x[i] = fma(y[i], z[(i+5)&31], x[i])  3 fresh, other placement
// MODE 4: x[i] = y[i] * x[i] (DMUL 2 fresh)    MODE 5: x[i] = u * x[i] (DMUL 1 fresh)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
constexpr int ITERS = 2048;
template<int MODE>
__global__ void __launch_bounds__(256, 1) k(double* out, const double* in) {
    double x[32], y[32], z[32], u[4];
#pragma unroll
    for (int i = 0; i < 32; i++) { x[i] = in[threadIdx.x + i]; y[i] = in[threadIdx.x + 64 + i]; z[i] = in[threadIdx.x + 128 + i]; }
#pragma unroll
    for (int i = 0; i < 4; i++) u[i] = in[threadIdx.x + 200 + i];
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 32; i++) {
            if (MODE == 0) x[i] = fma(x[i], u[0], u[1]);
            else if (MODE == 1) x[i] = fma(y[i], u[i >> 3], x[i]);
            else if (MODE == 2) x[i] = fma(y[i], z[i], x[i]);
            else if (MODE == 3) x[i] = fma(y[i], z[(i + 5) & 31], x[i]);
            else if (MODE == 4) x[i] = y[i] * x[i];
            else x[i] = u[i >> 3] * x[i];
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template<typename F> float time_ms(F launch) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize()); float best = 1e30f;
    for (int r = 0; r < 5; r++) { CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
    return best;
}
int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount; double *out, *in; CK(cudaMalloc(&out, 8 * 256 * sms * 8)); CK(cudaMalloc(&in, 8 * 1024));
    CK(cudaMemset(in, 0, 8 * 1024));
    printf("{");
#define RUN(NAME, MODE, W) { float ms = time_ms([&]{ k<MODE><<<sms * 8, W * 32>>>(out, in); }); \
      double ins = (double)sms * 8 * W * 32 * ITERS * 32; printf("\"%s_%dw\": %.2f, ", NAME, W, ins / (ms * 1e-3) / 1e12); }
    RUN("dfma_1fresh", 0, 8) RUN("dfma_2fresh_1shared", 1, 8) RUN("dfma_3fresh", 2, 8) RUN("dfma_3fresh_b", 3, 8) RUN("dmul_2fresh", 4, 8) RUN("dmul_1fresh", 5, 8)
    RUN("dfma_1fresh", 0, 4) RUN("dfma_2fresh_1shared", 1, 4) RUN("dfma_3fresh", 2, 4) RUN("dfma_3fresh_b", 3, 4) RUN("dmul_2fresh", 4, 4) RUN("dmul_1fresh", 5, 4)
    printf("\"unit\": \"T lane-instructions/s (peak = 148 SM x 64 lanes x 1.965 GHz = 18.6)\"}\n");
    return 0;
}
