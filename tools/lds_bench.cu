// Shared-memory load throughput under the multicast patterns the celerite kernel uses (not product code).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
constexpr int ITERS = 2048;
// MODE: 0 lane-distinct, 1 group-of-4 multicast (8 distinct addrs, stride STRIDE doubles), 2 uniform
template<int MODE, int VEC, int STRIDE>
__global__ void k_lds(double* out, int salt) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 1e-9;
    __syncthreads();
    int lane = threadIdx.x & 31;
    int base = (MODE == 0) ? lane * VEC : (MODE == 1) ? (lane >> 2) * STRIDE : 0;
    base += salt;
    double acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            int idx = (base + u * 64 + (it & 7) * 2) & 4095;
            if (VEC == 1) { acc0 += sm[idx]; }
            else { double2 v = *reinterpret_cast<const double2*>(&sm[idx & ~1]); acc0 += v.x; acc1 += v.y; }
        }
    }
    double s = acc0 + acc1 + acc2 + acc3;
    if (s == 123.456) out[0] = s;
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
    int sms = p.multiProcessorCount; double* out; CK(cudaMalloc(&out, 64));
    const int TPB = 512, GRID = sms * 2; // 32 warps/SM
    printf("{");
#define RUN(NAME, MODE, VEC, STRIDE) { float ms = time_ms([&]{ k_lds<MODE, VEC, STRIDE><<<GRID, TPB, 4096 * 8>>>(out, 0); }); \
      double lds_per_sm = (double)ITERS * 8 * (TPB / 32) * 2; /* warp-level LDS per SM */ \
      double cyc = ms * 1e-3 * 1.965e9; /* upper bound on cycles (boost clock) */ \
      printf("\"%s_cyc_per_warp_lds_per_sm\": %.3f, ", NAME, cyc / lds_per_sm); }
    RUN("lds64_distinct", 0, 1, 0)
    RUN("lds64_group4_stride5", 1, 1, 5)
    RUN("lds64_group4_stride6", 1, 1, 6)
    RUN("lds64_uniform", 2, 1, 0)
    RUN("lds128_distinct", 0, 2, 0)
    RUN("lds128_group4_stride6", 1, 2, 6)
    RUN("lds128_group4_stride8", 1, 2, 8)
    RUN("lds128_uniform", 2, 2, 0)
    printf("\"note\": \"cycles at 1.965 GHz per warp-level LDS, SM-wide (includes 1-2 DADD each)\"}\n");
    return 0;
}
