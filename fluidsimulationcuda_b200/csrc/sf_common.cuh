// Shared device/host definitions for the stable-fluids kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sf {

// Geometry of the (slab of the) grid a kernel works on.  Fields are row-major, pitch G = N+2.
// `row_base` is the GLOBAL row number of the first row stored in a field array, so global row r
// lives at  field + (r - row_base) * G.  Single GPU: row_base = 0, own = [0, G).
struct Geom {
    int N;         // interior width (Stam's N)
    int G;         // N + 2
    int row_base;  // global row of local row 0 (= own_lo - halo)
    int own_lo;    // first owned global row
    int own_hi;    // one past the last owned global row
    int rows;      // rows stored locally (own_hi - own_lo + 2*halo)
};

enum ArithMode { MODE_STRICT = 0, MODE_PRESSURE = 1, MODE_FAST = 2 };

// One Jacobi cell update with the reference's operand order (FluidSequential.c:95-96):
//   ((left + right) + up) + down ;  x0 + alpha*sum ;  / beta.
// __fadd_rn/__fmul_rn are never contracted into FMAs by nvcc; __fdiv_rn is the IEEE division.
//   MODE_PRESSURE: alpha == 1, beta == 4 exactly: 1*sum == sum and /4 == *0.25f are exact
//                  identities in binary32 (also for subnormal results), so this is bit-identical
//                  to the STRICT formula at a third of the instructions.
//   MODE_FAST:     FMA + reciprocal multiply (opt-in, not bit-identical).
template <int MODE>
__device__ __forceinline__ float jacobi_cell(float l, float r, float up, float dn, float b, float alpha,
                                             float beta, float rbeta)
{
    float s = __fadd_rn(__fadd_rn(__fadd_rn(l, r), up), dn);
    if (MODE == MODE_PRESSURE) return __fmul_rn(__fadd_rn(b, s), 0.25f);
    if (MODE == MODE_FAST) return __fmul_rn(__fmaf_rn(alpha, s, b), rbeta);
    return __fdiv_rn(__fadd_rn(b, __fmul_rn(alpha, s)), beta);
}

__host__ __device__ __forceinline__ uint32_t hash100(uint64_t seed, uint64_t field, uint64_t cell)
{
    // splitmix64 finaliser over (seed, field, global cell id)
    uint64_t z = seed * 0x9E3779B97F4A7C15ull + field * 0xD1B54A32D192ED03ull + cell;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (uint32_t)(z % 100ull);
}

// ---- launch wrappers implemented in the .cu files (all enqueue on `st`, return cudaError_t) ----
struct JacobiLaunch {
    const float *xin, *rhs;
    float *xout;
    float alpha, beta;
    int b;          // boundary kind 0/1/2
    int sweeps;     // 1..8 fused sweeps
    int mode;       // ArithMode
    int out_lo, out_hi;  // global rows to produce, within [own_lo, own_hi)
    int chunk_rows;      // 0 = auto
    int zero_guess;      // xin is known to be all zeros: do not read it
};
cudaError_t launch_jacobi_stream(const Geom &g, const JacobiLaunch &L, int sm_count, cudaStream_t st);
cudaError_t launch_jacobi_generic(const Geom &g, const JacobiLaunch &L, cudaStream_t st);
bool jacobi_stream_supported(const Geom &g);

cudaError_t launch_set_bnd(const Geom &g, int b, float *x, cudaStream_t st);
cudaError_t launch_add_source(const Geom &g, int nfields, float *const *x, const float *const *s, float dt,
                              cudaStream_t st);
cudaError_t launch_advect(const Geom &g, int b, float *d, const float *d0, const float *u, const float *v, float dt,
                          cudaStream_t st);
// both velocity components in one pass: d_u <- advect(b=1, u0), d_v <- advect(b=2, v0) by (u0, v0)
cudaError_t launch_advect_uv(const Geom &g, float *du, float *dv, const float *u0, const float *v0, float dt,
                             cudaStream_t st);
cudaError_t launch_divergence(const Geom &g, const float *u, const float *v, float *p, float *div, int write_p,
                              cudaStream_t st);
cudaError_t launch_last_project(const Geom &g, float *u, float *v, const float *p, cudaStream_t st);
cudaError_t launch_init(const Geom &g, uint64_t seed, float *dens, float *dens_prev, float *u, float *u_prev, float *v,
                        float *v_prev, cudaStream_t st);
cudaError_t launch_max_abs(const Geom &g, const float *x, float *dev_out, cudaStream_t st);
cudaError_t launch_residual(const Geom &g, const float *x, const float *x0, float alpha, float beta, double *dev_out,
                            cudaStream_t st);

}  // namespace sf
