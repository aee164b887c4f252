// Issue-rate microbenchmark for the FP32 add/fma pipes, guards, conversions and FP64 (developer tool).
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int V>
__global__ void k(float *out, float one, float c, double dc) {
    float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    #pragma unroll 8
    for (int i = 0; i < ITERS; ++i) {
        if (V == 0) { a0 = __fadd_rn(a0, c); a1 = __fadd_rn(a1, c); a2 = __fadd_rn(a2, c); a3 = __fadd_rn(a3, c); a4 = __fadd_rn(a4, c); a5 = __fadd_rn(a5, c); a6 = __fadd_rn(a6, c); a7 = __fadd_rn(a7, c); }
        if (V == 1) { a0 = __fmaf_rn(a0, one, c); a1 = __fmaf_rn(a1, one, c); a2 = __fmaf_rn(a2, one, c); a3 = __fmaf_rn(a3, one, c); a4 = __fmaf_rn(a4, one, c); a5 = __fmaf_rn(a5, one, c); a6 = __fmaf_rn(a6, one, c); a7 = __fmaf_rn(a7, one, c); }
        if (V == 2) { asm volatile("fma.rn.f32 %0, %0, 0f3F800000, %1;" : "+f"(a0) : "f"(c)); asm volatile("fma.rn.f32 %0, %0, 0f3F800000, %1;" : "+f"(a1) : "f"(c)); asm volatile("fma.rn.f32 %0, %0, 0f3F800000, %1;" : "+f"(a2) : "f"(c)); asm volatile("fma.rn.f32 %0, %0, 0f3F800000, %1;" : "+f"(a3) : "f"(c));
                      asm volatile("fma.rn.f32 %0, %0, 0f3F800000, %1;" : "+f"(a4) : "f"(c)); asm volatile("fma.rn.f32 %0, %0, 0f3F800000, %1;" : "+f"(a5) : "f"(c)); asm volatile("fma.rn.f32 %0, %0, 0f3F800000, %1;" : "+f"(a6) : "f"(c)); asm volatile("fma.rn.f32 %0, %0, 0f3F800000, %1;" : "+f"(a7) : "f"(c)); }
        if (V == 3) { a0 = __fadd_rn(a0, c); a1 = __fmaf_rn(a1, one, c); a2 = __fadd_rn(a2, c); a3 = __fmaf_rn(a3, one, c); a4 = __fadd_rn(a4, c); a5 = __fmaf_rn(a5, one, c); a6 = __fadd_rn(a6, c); a7 = __fmaf_rn(a7, one, c); }   // half FADD, half FFMA
        if (V == 4) { a0 = __fmul_rn(a0, c); a1 = __fmul_rn(a1, c); a2 = __fmul_rn(a2, c); a3 = __fmul_rn(a3, c); a4 = __fmul_rn(a4, c); a5 = __fmul_rn(a5, c); a6 = __fmul_rn(a6, c); a7 = __fmul_rn(a7, c); }
        if (V == 5) { a0 = __fmul_rn(a0, 0.25f); a1 = __fmul_rn(a1, 0.25f); a2 = __fmul_rn(a2, 0.25f); a3 = __fmul_rn(a3, 0.25f); a4 = __fmul_rn(a4, 0.25f); a5 = __fmul_rn(a5, 0.25f); a6 = __fmul_rn(a6, 0.25f); a7 = __fmul_rn(a7, 0.25f); }
        if (V == 6) { // conversions + DP: float->double, 1 dmul, 2 dfma, ->float
            double A0 = (double)a0, A1 = (double)a1, A2 = (double)a2, A3 = (double)a3;
            double q0 = __dmul_rn(A0, dc), q1 = __dmul_rn(A1, dc), q2 = __dmul_rn(A2, dc), q3 = __dmul_rn(A3, dc);
            double e0 = __fma_rn(dc, q0, -A0), e1 = __fma_rn(dc, q1, -A1), e2 = __fma_rn(dc, q2, -A2), e3 = __fma_rn(dc, q3, -A3);
            a0 = __double2float_rn(__fma_rn(-e0, dc, q0)); a1 = __double2float_rn(__fma_rn(-e1, dc, q1)); a2 = __double2float_rn(__fma_rn(-e2, dc, q2)); a3 = __double2float_rn(__fma_rn(-e3, dc, q3)); }
        if (V == 7) { // conversions only
            a0 = __double2float_rn((double)a0 + dc); a1 = __double2float_rn((double)a1 + dc); a2 = __double2float_rn((double)a2 + dc); a3 = __double2float_rn((double)a3 + dc); }
        if (V == 8) { // shuffles
            a0 = __shfl_up_sync(0xffffffffu, a0, 1); a1 = __shfl_down_sync(0xffffffffu, a1, 1); a2 = __shfl_up_sync(0xffffffffu, a2, 1); a3 = __shfl_down_sync(0xffffffffu, a3, 1); }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
int main() {
    float *out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char *names[] = {"FADD x8", "FFMA(reg one) x8", "FFMA(imm 1.0) x8", "4 FADD + 4 FFMA(reg)", "FMUL(reg) x8", "FMUL(imm) x8", "4x [F2F, DMUL, 2 DFMA, F2F]", "4x [F2F, DADD, F2F]", "4 SHFL"};
    int nops[] = {8, 8, 8, 8, 8, 8, 4, 4, 4};
    for (int v = 0; v < 9; ++v) {
        float ms = 0;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            switch (v) {
                case 0: k<0><<<148 * 8, 256>>>(out, 1.f, 1e-3f, 1.0000001); break; case 1: k<1><<<148 * 8, 256>>>(out, 1.f, 1e-3f, 1.0000001); break;
                case 2: k<2><<<148 * 8, 256>>>(out, 1.f, 1e-3f, 1.0000001); break; case 3: k<3><<<148 * 8, 256>>>(out, 1.f, 1e-3f, 1.0000001); break;
                case 4: k<4><<<148 * 8, 256>>>(out, 1.f, 1.0000001f, 1.0000001); break; case 5: k<5><<<148 * 8, 256>>>(out, 1.f, 1e-3f, 1.0000001); break;
                case 6: k<6><<<148 * 8, 256>>>(out, 1.f, 1e-3f, 1.0000001); break; case 7: k<7><<<148 * 8, 256>>>(out, 1.f, 1e-3f, 1.0000001); break;
                case 8: k<8><<<148 * 8, 256>>>(out, 1.f, 1e-3f, 1.0000001); break;
            }
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        }
        double groups = 148.0 * 8 * 256 * (double)ITERS * nops[v];
        printf("%-32s %8.3f ms  %7.1f lane-groups/clk/SM @1.9GHz (per listed unit)\n", names[v], ms, groups / (ms * 1e-3) / 148 / 1.9e9);
    }
    return 0;
}
