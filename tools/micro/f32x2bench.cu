// Issue rate of the two-wide binary32 instructions of sm_100 (FADD2 / FMUL2 / FFMA2) against their scalar
// forms, alone and mixed with scalar work (developer tool; numbers quoted in DESIGN.md).
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
__device__ __forceinline__ void add2(float &x, float &y, float cx, float cy)
{
    asm volatile("{\n.reg .b64 a, b;\nmov.b64 a, {%0, %1};\nmov.b64 b, {%2, %3};\nadd.rn.f32x2 a, a, b;\nmov.b64 {%0, %1}, a;\n}\n" : "+f"(x), "+f"(y) : "f"(cx), "f"(cy));
}
__device__ __forceinline__ void mul2(float &x, float &y, float cx, float cy)
{
    asm volatile("{\n.reg .b64 a, b;\nmov.b64 a, {%0, %1};\nmov.b64 b, {%2, %3};\nmul.rn.f32x2 a, a, b;\nmov.b64 {%0, %1}, a;\n}\n" : "+f"(x), "+f"(y) : "f"(cx), "f"(cy));
}
__device__ __forceinline__ void fma2(float &x, float &y, float mx, float my, float cx, float cy)
{
    asm volatile("{\n.reg .b64 a, b, c;\nmov.b64 a, {%0, %1};\nmov.b64 b, {%2, %3};\nmov.b64 c, {%4, %5};\nfma.rn.f32x2 a, a, b, c;\nmov.b64 {%0, %1}, a;\n}\n"
                 : "+f"(x), "+f"(y) : "f"(mx), "f"(my), "f"(cx), "f"(cy));
}
template <int V>
__global__ void k(float *out, float one, float c, float c2)
{
    float a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = threadIdx.x + j;
    const float vc = c + (threadIdx.x & 1) * 1e-6f, vc2 = c2 + (threadIdx.x & 1) * 1e-6f;   // per-lane operands (vector registers)
#pragma unroll 4
    for (int i = 0; i < ITERS; ++i) {
        if (V == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j] = __fadd_rn(a[j], vc);                 // 16 FADD
        }
        if (V == 1) {
#pragma unroll
            for (int j = 0; j < 16; j += 2) add2(a[j], a[j + 1], vc, vc2);           // 8 FADD2
        }
        if (V == 2) {
#pragma unroll
            for (int j = 0; j < 16; j += 2) mul2(a[j], a[j + 1], vc, vc2);           // 8 FMUL2
        }
        if (V == 3) {
#pragma unroll
            for (int j = 0; j < 16; j += 2) fma2(a[j], a[j + 1], one, one, vc, vc2); // 8 FFMA2
        }
        if (V == 4) {                                                                 // 4 FADD2 + 8 FADD
#pragma unroll
            for (int j = 0; j < 8; j += 2) add2(a[j], a[j + 1], vc, vc2);
#pragma unroll
            for (int j = 8; j < 16; ++j) a[j] = __fadd_rn(a[j], vc);
        }
        if (V == 5) {                                                                 // 4 FADD2 + 8 IADD (alu pipe)
#pragma unroll
            for (int j = 0; j < 8; j += 2) add2(a[j], a[j + 1], vc, vc2);
#pragma unroll
            for (int j = 8; j < 16; ++j) a[j] = __int_as_float(__float_as_int(a[j]) * 2 - 1);
        }
        if (V == 6) {                                                                 // 8 FADD2, uniform (scalar-broadcast) operand
#pragma unroll
            for (int j = 0; j < 16; j += 2) add2(a[j], a[j + 1], c, c);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int V>
float run(float *out, int warps_per_sm)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<V><<<148 * warps_per_sm / 4, 128>>>(out, 1.f, 1e-3f, 2e-3f);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    }
    return ms;
}
int main()
{
    float *out; cudaMalloc(&out, 148 * 16 * 32 * 4 * 4);
    const char *names[] = {"16 FADD", "8 FADD2", "8 FMUL2", "8 FFMA2", "4 FADD2 + 8 FADD", "4 FADD2 + 8 int ops", "8 FADD2 (uniform operand)"};
    const int instr[] = {16, 8, 8, 8, 12, 12, 8}, flops[] = {16, 16, 16, 16, 16, 8, 16};
    for (int wps : {12, 16, 32}) {
        printf("-- %d warps per SM\n", wps);
        for (int v = 0; v < 7; ++v) {
            float ms = 0;
            switch (v) {
                case 0: ms = run<0>(out, wps); break; case 1: ms = run<1>(out, wps); break; case 2: ms = run<2>(out, wps); break;
                case 3: ms = run<3>(out, wps); break; case 4: ms = run<4>(out, wps); break; case 5: ms = run<5>(out, wps); break;
                case 6: ms = run<6>(out, wps); break;
            }
            const double warps = 148.0 * wps, clk = ms * 1e-3 * 1.965e9;
            printf("%-28s %8.3f ms   %6.2f warp-instr/clk/SM   %6.1f fp32 lane-ops/clk/SM (at 1.965 GHz)\n", names[v], ms,
                   warps * ITERS * instr[v] / clk / 148.0, warps * ITERS * flops[v] * 32 / clk / 148.0);
        }
    }
    return 0;
}
