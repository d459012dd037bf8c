// FP64 ceiling microbenchmark for B200 (sm_100a): DFMA, DMMA (mma.sync f64), mixed issue,
// and DFMA co-issued with SHFL / LDS.  Prints one JSON object.  Not part of the product path.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

constexpr int ITERS = 4096;

template<int CH>
__global__ void k_dfma(double* out, double a, double b) {
    double x[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) x[i] = fma(x[i], a, b);
    }
    double s = 0; 
#pragma unroll
    for (int i = 0; i < CH; i++) s += x[i];
    if (s == 123.456) out[0] = s;
}

// DFMA chains + one 64-bit shuffle (2 SHFL) per SH_EVERY fma
template<int CH, int NSH>
__global__ void k_dfma_shfl(double* out, double a, double b) {
    double x[CH]; double t = threadIdx.x;
#pragma unroll
    for (int i = 0; i < CH; i++) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) x[i] = fma(x[i], a, b);
#pragma unroll
        for (int i = 0; i < NSH; i++) t = __shfl_xor_sync(0xffffffffu, t, 1 + i);
    }
    double s = t; 
#pragma unroll
    for (int i = 0; i < CH; i++) s += x[i];
    if (s == 123.456) out[0] = s;
}

template<int CH, int NLD>
__global__ void k_dfma_lds(double* out, double a, double b) {
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i * 1e-9;
    __syncthreads();
    double x[CH]; double t = 0; int idx = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < CH; i++) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) x[i] = fma(x[i], a, b);
#pragma unroll
        for (int i = 0; i < NLD; i++) { t += 0; x[i % CH] += sm[(idx + i * 32 + it) & 1023]; }
    }
    double s = t; 
#pragma unroll
    for (int i = 0; i < CH; i++) s += x[i];
    if (s == 123.456) out[0] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template<int CH>
__global__ void k_dmma884(double* out, double a, double b) {
    double c0[CH], c1[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { c0[i] = i; c1[i] = -i; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) dmma884(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += c0[i] + c1[i];
    if (s == 123.456) out[0] = s;
}

template<int CH>
__global__ void k_dmma16816(double* out, double a, double b) {
    double c[CH][4]; double av[8], bv[4];
#pragma unroll
    for (int i = 0; i < 8; i++) av[i] = a + i;
#pragma unroll
    for (int i = 0; i < 4; i++) bv[i] = b + i;
#pragma unroll
    for (int i = 0; i < CH; i++) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) dmma16816(c[i], av, bv);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456) out[0] = s;
}

// interleave: CHM dmma884 + CHF dfma per iteration
template<int CHM, int CHF>
__global__ void k_mixed(double* out, double a, double b) {
    double c0[CHM], c1[CHM], x[CHF];
#pragma unroll
    for (int i = 0; i < CHM; i++) { c0[i] = i; c1[i] = -i; }
#pragma unroll
    for (int i = 0; i < CHF; i++) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < (CHM > CHF ? CHM : CHF); i++) {
            if (i < CHM) dmma884(c0[i], c1[i], a, b);
            if (i < CHF) x[i] = fma(x[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHM; i++) s += c0[i] + c1[i];
#pragma unroll
    for (int i = 0; i < CHF; i++) s += x[i];
    if (s == 123.456) out[0] = s;
}

// reciprocal + log throughput (warp-level): how expensive are 1/x and log(x) per call
__global__ void k_div(double* out, double a) {
    double x = 1.0 + threadIdx.x * 1e-3, acc = 0;
    for (int it = 0; it < ITERS; it++) { x = 1.0 / (x + a); acc += x; }
    if (acc == 123.456) out[0] = acc;
}
__global__ void k_log(double* out, double a) {
    double x = 1.0 + threadIdx.x * 1e-3, acc = 0;
    for (int it = 0; it < ITERS; it++) { x = log(x + a) + 2.0; acc += x; }
    if (acc == 123.456) out[0] = acc;
}
__global__ void k_sincos(double* out, double a) {
    double x = 1.0 + threadIdx.x * 1e-3, acc = 0;
    for (int it = 0; it < ITERS; it++) { double s, c; sincos(x * 1000.0 + a, &s, &c); x = s + 2.0; acc += c; }
    if (acc == 123.456) out[0] = acc;
}
__global__ void k_exp(double* out, double a) {
    double x = 1.0 + threadIdx.x * 1e-3, acc = 0;
    for (int it = 0; it < ITERS; it++) { x = exp(-x + a) + 0.5; acc += x; }
    if (acc == 123.456) out[0] = acc;
}

template<typename F>
float time_ms(F launch, int reps = 5) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, 64));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", p.name, sms, p.clockRate);
    const int TPB = 512; const int GRID = sms * 4;  // 2048 threads/SM
    auto tf = [&](double flops_per_thread_iter, float ms, int grid, int tpb) {
        return flops_per_thread_iter * (double)ITERS * grid * tpb / (ms * 1e-3) / 1e12; };
    { float ms = time_ms([&]{ k_dfma<8><<<GRID, TPB>>>(out, 1.0000001, 1e-9); });
      printf(", \"dfma_tflops\": %.3f", tf(2.0 * 8, ms, GRID, TPB)); }
    { float ms = time_ms([&]{ k_dfma<8><<<sms, 128>>>(out, 1.0000001, 1e-9); });
      printf(", \"dfma_tflops_1warp_per_smsp\": %.3f", tf(2.0 * 8, ms, sms, 128)); }
    { float ms = time_ms([&]{ k_dfma<2><<<sms, 128>>>(out, 1.0000001, 1e-9); });
      printf(", \"dfma_tflops_1warp_2chains\": %.3f", tf(2.0 * 2, ms, sms, 128)); }
    { float ms = time_ms([&]{ k_dfma<1><<<sms, 128>>>(out, 1.0000001, 1e-9); });
      // 1 chain, 1 warp per SMSP: cycles per dependent DFMA = latency
      double cyc = ms * 1e-3 * p.clockRate * 1e3 / ITERS;
      printf(", \"dfma_dep_latency_cyc_at_nominal_clock\": %.2f", cyc); }
    { float ms = time_ms([&]{ k_dmma884<8><<<GRID, TPB>>>(out, 1.0000001, 1e-9); });
      printf(", \"dmma_m8n8k4_tflops\": %.3f", tf(512.0 / 32 * 8, ms, GRID, TPB)); }
    { float ms = time_ms([&]{ k_dmma16816<4><<<GRID, TPB>>>(out, 1.0000001, 1e-9); });
      printf(", \"dmma_m16n8k16_tflops\": %.3f", tf(2.0 * 16 * 8 * 16 / 32 * 4, ms, GRID, TPB)); }
    { float ms = time_ms([&]{ k_mixed<4, 4><<<GRID, TPB>>>(out, 1.0000001, 1e-9); });
      printf(", \"mixed_4dmma_4dfma_tflops\": %.3f", tf(512.0 / 32 * 4 + 2.0 * 4, ms, GRID, TPB)); }
    { float ms = time_ms([&]{ k_mixed<2, 8><<<GRID, TPB>>>(out, 1.0000001, 1e-9); });
      printf(", \"mixed_2dmma_8dfma_tflops\": %.3f", tf(512.0 / 32 * 2 + 2.0 * 8, ms, GRID, TPB)); }
    { float ms = time_ms([&]{ k_dfma_shfl<8, 2><<<GRID, TPB>>>(out, 1.0000001, 1e-9); });
      printf(", \"dfma8_shfl64x2_dfma_tflops\": %.3f", tf(2.0 * 8, ms, GRID, TPB)); }
    { float ms = time_ms([&]{ k_dfma_shfl<8, 4><<<GRID, TPB>>>(out, 1.0000001, 1e-9); });
      printf(", \"dfma8_shfl64x4_dfma_tflops\": %.3f", tf(2.0 * 8, ms, GRID, TPB)); }
    { float ms = time_ms([&]{ k_dfma_shfl<8, 8><<<GRID, TPB>>>(out, 1.0000001, 1e-9); });
      printf(", \"dfma8_shfl64x8_dfma_tflops\": %.3f", tf(2.0 * 8, ms, GRID, TPB)); }
    { float ms = time_ms([&]{ k_dfma_lds<8, 2><<<GRID, TPB>>>(out, 1.0000001, 1e-9); });
      printf(", \"dfma8_lds64x2_dfma_tflops\": %.3f", tf(2.0 * 8 + 2, ms, GRID, TPB)); }
    { float ms = time_ms([&]{ k_dfma_lds<8, 4><<<GRID, TPB>>>(out, 1.0000001, 1e-9); });
      printf(", \"dfma8_lds64x4_dfma_tflops\": %.3f", tf(2.0 * 8 + 4, ms, GRID, TPB)); }
    { float ms = time_ms([&]{ k_dfma_lds<8, 8><<<GRID, TPB>>>(out, 1.0000001, 1e-9); });
      printf(", \"dfma8_lds64x8_dfma_tflops\": %.3f", tf(2.0 * 8 + 8, ms, GRID, TPB)); }
    // transcendental costs in "DFMA-equivalent warp slots": time per call relative to time per DFMA at peak
    float ms_fma = time_ms([&]{ k_dfma<8><<<GRID, TPB>>>(out, 1.0000001, 1e-9); });
    double per_fma = ms_fma / (8.0 * ITERS);
    { float ms = time_ms([&]{ k_div<<<GRID, TPB>>>(out, 1e-9); });   printf(", \"div_cost_in_dfma\": %.1f", ms / ITERS / per_fma); }
    { float ms = time_ms([&]{ k_log<<<GRID, TPB>>>(out, 1e-9); });   printf(", \"log_cost_in_dfma\": %.1f", ms / ITERS / per_fma); }
    { float ms = time_ms([&]{ k_sincos<<<GRID, TPB>>>(out, 1e-9); });printf(", \"sincos_cost_in_dfma\": %.1f", ms / ITERS / per_fma); }
    { float ms = time_ms([&]{ k_exp<<<GRID, TPB>>>(out, 1e-9); });   printf(", \"exp_cost_in_dfma\": %.1f", ms / ITERS / per_fma); }
    // sustained DFMA for ~2 s to see the power-capped clock
    { cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      int n = 0; CK(cudaEventRecord(e0));
      for (; n < 400; n++) k_dfma<8><<<GRID, TPB>>>(out, 1.0000001, 1e-9);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      printf(", \"dfma_tflops_sustained\": %.3f, \"sustained_seconds\": %.2f", tf(2.0 * 8, ms / n, GRID, TPB), ms * 1e-3); }
    printf("}\n");
    return 0;
}
