// Does a shared-memory load cost less when only some lanes are active?  (not product code)
// MODE 0: all lanes, 8 distinct 16-byte addresses (quad-shared);  1: only lanes with (lane&3)==0 (predicated ld.shared);
// 2: only lanes 0..7;  3: all lanes distinct addresses.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
constexpr int ITERS = 2048;
template<int MODE, int VEC>
__global__ void k(unsigned long long* out, int salt) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 1e-9;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int sel = (MODE == 1) ? (lane & 3) : (MODE == 2) ? (lane >> 3) : 0;     // active iff sel == 0
    const int base = ((MODE == 3) ? lane * VEC : (lane >> 2) * 10) + salt;
    unsigned long long a0 = 0, a1 = 0;
    for (int it = 0; it < ITERS; it++) {
        const int idx = ((base + (it & 7) * 2) & 1022);
        const unsigned addr = (unsigned)__cvta_generic_to_shared(&sm[idx]);
        unsigned long long x0 = 0, x1 = 0, x2 = 0, x3 = 0, x4 = 0, x5 = 0, x6 = 0, x7 = 0;
        unsigned long long y0 = 0, y1 = 0, y2 = 0, y3 = 0, y4 = 0, y5 = 0, y6 = 0, y7 = 0;
        if (VEC == 2)
            asm volatile("{.reg .pred p; setp.eq.s32 p, %17, 0;\n"
                         "@p ld.shared.v2.u64 {%0,%1}, [%16];\n @p ld.shared.v2.u64 {%2,%3}, [%16+768];\n"
                         "@p ld.shared.v2.u64 {%4,%5}, [%16+1536];\n @p ld.shared.v2.u64 {%6,%7}, [%16+2304];\n"
                         "@p ld.shared.v2.u64 {%8,%9}, [%16+3072];\n @p ld.shared.v2.u64 {%10,%11}, [%16+3840];\n"
                         "@p ld.shared.v2.u64 {%12,%13}, [%16+4608];\n @p ld.shared.v2.u64 {%14,%15}, [%16+5376];}"
                         : "+l"(x0), "+l"(y0), "+l"(x1), "+l"(y1), "+l"(x2), "+l"(y2), "+l"(x3), "+l"(y3), "+l"(x4), "+l"(y4),
                           "+l"(x5), "+l"(y5), "+l"(x6), "+l"(y6), "+l"(x7), "+l"(y7) : "r"(addr), "r"(sel));
        else
            asm volatile("{.reg .pred p; setp.eq.s32 p, %9, 0;\n"
                         "@p ld.shared.u64 %0, [%8];\n @p ld.shared.u64 %1, [%8+768];\n @p ld.shared.u64 %2, [%8+1536];\n"
                         "@p ld.shared.u64 %3, [%8+2304];\n @p ld.shared.u64 %4, [%8+3072];\n @p ld.shared.u64 %5, [%8+3840];\n"
                         "@p ld.shared.u64 %6, [%8+4608];\n @p ld.shared.u64 %7, [%8+5376];}"
                         : "+l"(x0), "+l"(x1), "+l"(x2), "+l"(x3), "+l"(x4), "+l"(x5), "+l"(x6), "+l"(x7) : "r"(addr), "r"(sel));
        a0 ^= x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
        a1 ^= y0 ^ y1 ^ y2 ^ y3 ^ y4 ^ y5 ^ y6 ^ y7;
    }
    if ((a0 ^ a1) == 0x123456789ULL) out[0] = a0;
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
    int sms = p.multiProcessorCount; unsigned long long* out; CK(cudaMalloc(&out, 64));
    const int TPB = 512, GRID = sms * 2; // 32 warps/SM
    printf("{");
#define RUN(NAME, MODE, VEC) { float ms = time_ms([&]{ k<MODE, VEC><<<GRID, TPB, 4096 * 8>>>(out, 0); }); \
      double lds_per_sm = (double)ITERS * 8 * (TPB / 32) * 2; double cyc = ms * 1e-3 * 1.965e9; \
      printf("\"%s\": %.3f, ", NAME, cyc / lds_per_sm); }
    RUN("lds64_all_quadshared", 0, 1) RUN("lds64_one_lane_per_quad", 1, 1) RUN("lds64_lanes0to7", 2, 1) RUN("lds64_all_distinct", 3, 1)
    RUN("lds128_all_quadshared", 0, 2) RUN("lds128_one_lane_per_quad", 1, 2) RUN("lds128_lanes0to7", 2, 2) RUN("lds128_all_distinct", 3, 2)
    printf("\"unit\": \"SM cycles (1.965 GHz) per warp-level LDS, 32 warps/SM\"}\n");
    return 0;
}
