// Does FP64 FMA throughput depend on register-file operand traffic?  (not product code)
// MODE 0: x[i] = fma(x[i], a, b)            — two operands shared by every instruction (what a peak test does)
// MODE 1: x[i] = fma(y[i], z[i], x[i])      — three distinct register pairs per instruction, no operand shared
// MODE 2: x[r][c] = fma(q[r], w[c], x[r][c]) — outer-product update (the block phase's access pattern), c outer / r inner
// MODE 3: as 2 plus the dependent mul and the two matvec FMAs of the celerite block phase (4 FP64 per entry)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
constexpr int ITERS = 512;
template<int MODE>
__global__ void __launch_bounds__(256, 1) k(double* out, const double* in, int n) {
    double x[8][8], q[8], w[8], y[8][8];
    for (int r = 0; r < 8; r++) { q[r] = in[(threadIdx.x + r) % n]; w[r] = in[(threadIdx.x + 2 * r + 1) % n];
        for (int c = 0; c < 8; c++) { x[r][c] = in[(threadIdx.x + r * 8 + c) % n]; y[r][c] = in[(threadIdx.x * 3 + r * 8 + c) % n]; } }
    double rp[8] = {0, 0, 0, 0, 0, 0, 0, 0}, cs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int it = 0; it < ITERS; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int c = 0; c < 8; c++)
#pragma unroll
                for (int r = 0; r < 8; r++) x[r][c] = fma(x[r][c], q[0], w[0]);
        } else if (MODE == 1) {
#pragma unroll
            for (int c = 0; c < 8; c++)
#pragma unroll
                for (int r = 0; r < 8; r++) x[r][c] = fma(y[r][c], y[(r + 3) & 7][(c + 5) & 7], x[r][c]);
        } else if (MODE == 2) {
#pragma unroll
            for (int c = 0; c < 8; c++)
#pragma unroll
                for (int r = 0; r < 8; r++) x[r][c] = fma(q[r], w[c], x[r][c]);
        } else if (MODE == 4) {      // DMUL, one operand shared
#pragma unroll
            for (int c = 0; c < 8; c++)
#pragma unroll
                for (int r = 0; r < 8; r++) x[r][c] = x[r][c] * q[c];
        } else if (MODE == 5) {      // DMUL, both operands distinct
#pragma unroll
            for (int c = 0; c < 8; c++)
#pragma unroll
                for (int r = 0; r < 8; r++) x[r][c] = x[r][c] * y[(r + 3) & 7][(c + 5) & 7];
        } else if (MODE == 6) {      // DADD, both operands distinct
#pragma unroll
            for (int c = 0; c < 8; c++)
#pragma unroll
                for (int r = 0; r < 8; r++) x[r][c] = x[r][c] + y[(r + 3) & 7][(c + 5) & 7];
        } else if (MODE == 7) {      // DFMA, accumulate chains: x[r][c] = fma(y[r][c], w[c], x[r][c]) (2 fresh + 1 shared)
#pragma unroll
            for (int c = 0; c < 8; c++)
#pragma unroll
                for (int r = 0; r < 8; r++) x[r][c] = fma(y[r][c], w[c], x[r][c]);
        } else if (MODE == 12 || MODE == 13) {   // as 7 through inline PTX with a fixed operand order: 12: (y, w, x)   13: (w, y, x)
#pragma unroll
            for (int c = 0; c < 8; c++)
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    if (MODE == 12) asm("fma.rn.f64 %0, %1, %2, %0;" : "+d"(x[r][c]) : "d"(y[r][c]), "d"(w[c]));
                    else            asm("fma.rn.f64 %0, %1, %2, %0;" : "+d"(x[r][c]) : "d"(w[c]), "d"(y[r][c]));
                }
        } else if (MODE >= 8 && MODE <= 11) {   // as 7 with y taken at another position: different relative register placement
            constexpr int DR = (MODE == 8) ? 1 : (MODE == 9) ? 3 : (MODE == 10) ? 0 : 4, DC = (MODE == 8) ? 0 : (MODE == 9) ? 5 : (MODE == 10) ? 1 : 4;
#pragma unroll
            for (int c = 0; c < 8; c++)
#pragma unroll
                for (int r = 0; r < 8; r++) x[r][c] = fma(y[(r + DR) & 7][(c + DC) & 7], w[c], x[r][c]);
        } else {
#pragma unroll
            for (int c = 0; c < 8; c++)
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    const double m = q[(r + 1) & 7] * fma(q[r], w[c], x[r][c]);
                    x[r][c] = m;
                    rp[r] = fma(m, w[(c + 3) & 7], rp[r]);
                    cs[c] = fma(m, q[(r + 5) & 7], cs[c]);
                }
        }
    }
    double s = 0;
    for (int r = 0; r < 8; r++) { s += rp[r] + cs[r]; for (int c = 0; c < 8; c++) s += x[r][c]; }
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
#define RUN(NAME, MODE, TPB, FPE) { float ms = time_ms([&]{ k<MODE><<<sms * 8, TPB>>>(out, in, 1024); }); \
      double fl = (double)sms * 8 * TPB * ITERS * 64 * FPE; printf("\"%s\": %.2f, ", NAME, fl / (ms * 1e-3) / 1e12); }
    RUN("shared_operands_8w", 0, 256, 2) RUN("distinct_operands_8w", 1, 256, 2) RUN("outer_product_8w", 2, 256, 2) RUN("block_phase_8w", 3, 256, 7)
    RUN("dmul_shared_8w", 4, 256, 1) RUN("dmul_distinct_8w", 5, 256, 1) RUN("dadd_distinct_8w", 6, 256, 1) RUN("dfma_2fresh_8w", 7, 256, 2)
    RUN("dfma_2fresh_dr1_8w", 8, 256, 2) RUN("dfma_2fresh_dr3dc5_8w", 9, 256, 2) RUN("dfma_2fresh_dc1_8w", 10, 256, 2) RUN("dfma_2fresh_dr4dc4_8w", 11, 256, 2)
    RUN("dfma_ptx_y_w_x_8w", 12, 256, 2) RUN("dfma_ptx_w_y_x_8w", 13, 256, 2)
    RUN("shared_operands_4w", 0, 128, 2) RUN("distinct_operands_4w", 1, 128, 2) RUN("outer_product_4w", 2, 128, 2) RUN("block_phase_4w", 3, 128, 7)
    printf("\"unit\": \"TFLOP/s (FMA = 2, MUL = 1); block_phase issues 4 FP64 per 7 flops\"}\n");
    return 0;
}
