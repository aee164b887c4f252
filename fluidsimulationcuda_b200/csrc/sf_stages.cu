// Non-Jacobi stages of the stable-fluids step, each with set_bnd fused in (the thread that
// produces an interior cell next to a wall also writes the wall cell, and the four threads at
// the interior corners write the grid corners), so every field is read and written once per stage.
// Reference semantics: FluidSequential.c:62-82 (set_bnd, add_source), :107-141 (advect),
// :143-158 (computeDivergenceAndPressure), :161-173 (lastProject), :244-271 (initial condition).
// All arithmetic uses __f*_rn intrinsics so nvcc cannot contract mul+add into FMA: results are
// bit-identical to the reference's sequential build.
#include <cuda.h>   // CUtensorMap (types only: the encoder is looked up at run time, the library does not link libcuda)

#include <initializer_list>

#include "sf_common.cuh"

namespace sf {
namespace {

struct Wall {
    // which wall/corner cells the thread owning interior cell (row, col) must also write
    bool L, R, T, B;
};
__device__ __forceinline__ Wall wall_of(const Geom &g, int row, int col)
{
    Wall w;
    w.L = (col == 1);
    w.R = (col == g.N);
    w.T = (row == 1) && (g.own_lo == 0);
    w.B = (row == g.N) && (g.own_hi == g.G);
    return w;
}
// write interior value `o` at flat index i plus the wall/corner cells it determines (set_bnd(b))
__device__ __forceinline__ void store_with_walls(float *x, size_t i, size_t G, float o, const Wall &w, float sx, float sy)
{
    x[i] = o;
    if (w.L | w.R | w.T | w.B) {
        const float wx = __fmul_rn(sx, o), wy = __fmul_rn(sy, o);
        const float cn = __fmul_rn(0.5f, __fadd_rn(wy, wx));
        if (w.L) x[i - 1] = wx;
        if (w.R) x[i + 1] = wx;
        if (w.T) x[i - G] = wy;
        if (w.B) x[i + G] = wy;
        if (w.L && w.T) x[i - G - 1] = cn;
        if (w.R && w.T) x[i - G + 1] = cn;
        if (w.L && w.B) x[i + G - 1] = cn;
        if (w.R && w.B) x[i + G + 1] = cn;
    }
}

// interior rows this slab produces
__device__ __forceinline__ void interior_rows(const Geom &g, int &lo, int &hi)
{
    lo = max(g.own_lo, 1);
    hi = min(g.own_hi, g.N + 1);
}

// ---- set_bnd ---------------------------------------------------------------------------------
__global__ void set_bnd_kernel(float *x, Geom g, float sx, float sy)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x + 1;   // 1..N
    if (k > g.N) return;
    const size_t G = (size_t)g.G;
    const int N = g.N;
    const bool top = (g.own_lo == 0), bot = (g.own_hi == g.G);
    if (k >= max(g.own_lo, 1) && k < min(g.own_hi, N + 1)) {   // side walls of owned row k
        const size_t r = (size_t)(k - g.row_base) * G;
        x[r] = __fmul_rn(sx, x[r + 1]);
        x[r + N + 1] = __fmul_rn(sx, x[r + N]);
    }
    if (top) x[(size_t)(0 - g.row_base) * G + k] = __fmul_rn(sy, x[(size_t)(1 - g.row_base) * G + k]);
    if (bot) x[(size_t)(N + 1 - g.row_base) * G + k] = __fmul_rn(sy, x[(size_t)(N - g.row_base) * G + k]);
    // corners: computed from the interior values that define the adjacent wall cells, so no
    // ordering between threads is needed (x[0][1] = sy*x[1][1], x[1][0] = sx*x[1][1], ...)
    // (two separate tests: at N == 1 the one thread owns the left AND the right corners)
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        if (k != (side == 0 ? 1 : N)) continue;
        const int col = k, wc = (side == 0) ? 0 : N + 1;
        if (top) {
            const float a = x[(size_t)(1 - g.row_base) * G + col];
            x[(size_t)(0 - g.row_base) * G + wc] = __fmul_rn(0.5f, __fadd_rn(__fmul_rn(sy, a), __fmul_rn(sx, a)));
        }
        if (bot) {
            const float a = x[(size_t)(N - g.row_base) * G + col];
            x[(size_t)(N + 1 - g.row_base) * G + wc] = __fmul_rn(0.5f, __fadd_rn(__fmul_rn(sy, a), __fmul_rn(sx, a)));
        }
    }
}

// ---- add_source: x += dt*s on every owned cell (ring included), up to 3 fields per launch ------
struct AddSrcArgs {
    float *x[3];
    const float *s[3];
    size_t first, count;   // flat range of owned cells (multiple of G)
    float dt;
    int vec;               // 1: first/count/pointers allow float4
};
__global__ void add_source_kernel(AddSrcArgs A)
{
    // selects, not A.x[blockIdx.y]: a dynamic index would copy the parameter block to local memory
    float *x = blockIdx.y == 0 ? A.x[0] : (blockIdx.y == 1 ? A.x[1] : A.x[2]);
    const float *s = blockIdx.y == 0 ? A.s[0] : (blockIdx.y == 1 ? A.s[1] : A.s[2]);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (A.vec) {
        float4 *x4 = reinterpret_cast<float4 *>(x + A.first);
        const float4 *s4 = reinterpret_cast<const float4 *>(s + A.first);
        const size_t n4 = A.count / 4;
        for (; i < n4; i += stride) {
            float4 a = x4[i];
            const float4 b = __ldg(s4 + i);
            a.x = __fadd_rn(a.x, __fmul_rn(A.dt, b.x));
            a.y = __fadd_rn(a.y, __fmul_rn(A.dt, b.y));
            a.z = __fadd_rn(a.z, __fmul_rn(A.dt, b.z));
            a.w = __fadd_rn(a.w, __fmul_rn(A.dt, b.w));
            x4[i] = a;
        }
    } else {
        for (; i < A.count; i += stride) x[A.first + i] = __fadd_rn(x[A.first + i], __fmul_rn(A.dt, s[A.first + i]));
    }
}

// ---- advect ------------------------------------------------------------------------------------
// One thread per interior cell.  The gather is data dependent but spatially smooth (neighbouring
// cells trace back to neighbouring sources), so a warp's four gathers land in a few 128-B lines.
// NF = 1: d <- advect(b, d0 by u, v).  NF = 2: the two velocity components in one pass
// (FluidSequential.c:232,237: advect(1,u,u0,u0,v0); advect(2,v,v0,u0,v0)) sharing the back-trace.
template <int NF>
__global__ void advect_kernel(float *__restrict__ dA, float *__restrict__ dB, const float *__restrict__ srcA,
                              const float *__restrict__ srcB, const float *__restrict__ u, const float *__restrict__ v,
                              Geom g, float dt0, int bA)
{
    int lo, hi;
    interior_rows(g, lo, hi);
    const int col = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int row = blockIdx.y * blockDim.y + threadIdx.y + lo;
    if (col > g.N || row >= hi) return;
    const size_t G = (size_t)g.G;
    const size_t i = (size_t)(row - g.row_base) * G + col;
    // FluidSequential.c:114-134
    float px = __fsub_rn((float)col, __fmul_rn(dt0, u[i]));
    float py = __fsub_rn((float)row, __fmul_rn(dt0, v[i]));
    const float hiC = (float)g.N + 0.5f;
    if (px < 0.5f) px = 0.5f;
    if (px > hiC) px = hiC;
    if (py < 0.5f) py = 0.5f;
    if (py > hiC) py = hiC;
    const int c0 = (int)px, r0 = (int)py;
    const float wx1 = __fsub_rn(px, (float)c0), wx0 = __fsub_rn(1.0f, wx1);
    const float wy1 = __fsub_rn(py, (float)r0), wy0 = __fsub_rn(1.0f, wy1);
    const int rs = min(max(r0, g.row_base), g.row_base + g.rows - 2);   // see advect_cell
    const size_t j = (size_t)(rs - g.row_base) * G + c0;
    const Wall w = wall_of(g, row, col);
    {
        const float a00 = __ldg(srcA + j), a10 = __ldg(srcA + j + G), a01 = __ldg(srcA + j + 1), a11 = __ldg(srcA + j + G + 1);
        const float colA = __fadd_rn(__fmul_rn(wy0, a00), __fmul_rn(wy1, a10));
        const float colB = __fadd_rn(__fmul_rn(wy0, a01), __fmul_rn(wy1, a11));
        const float o = __fadd_rn(__fmul_rn(wx0, colA), __fmul_rn(wx1, colB));
        store_with_walls(dA, i, G, o, w, bA == 1 ? -1.0f : 1.0f, bA == 2 ? -1.0f : 1.0f);
    }
    if (NF == 2) {
        const float a00 = __ldg(srcB + j), a10 = __ldg(srcB + j + G), a01 = __ldg(srcB + j + 1), a11 = __ldg(srcB + j + G + 1);
        const float colA = __fadd_rn(__fmul_rn(wy0, a00), __fmul_rn(wy1, a10));
        const float colB = __fadd_rn(__fmul_rn(wy0, a01), __fmul_rn(wy1, a11));
        const float o = __fadd_rn(__fmul_rn(wx0, colA), __fmul_rn(wx1, colB));
        store_with_walls(dB, i, G, o, w, 1.0f, -1.0f);   // b = 2
    }
}

// ---- divergence (+ p = 0) --------------------------------------------------------------------
__global__ void divergence_kernel(const float *__restrict__ u, const float *__restrict__ v, float *__restrict__ p,
                                  float *__restrict__ div, Geom g, float scale, int write_p)
{
    int lo, hi;
    interior_rows(g, lo, hi);
    const int col = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int row = blockIdx.y * blockDim.y + threadIdx.y + lo;
    if (col > g.N || row >= hi) return;
    const size_t G = (size_t)g.G;
    const size_t i = (size_t)(row - g.row_base) * G + col;
    // FluidSequential.c:151-152: (-0.5f*h) * (((u_r - u_l) + v_d) - v_u)
    float acc = __fsub_rn(__ldg(u + i + 1), __ldg(u + i - 1));
    acc = __fadd_rn(acc, __ldg(v + i + G));
    acc = __fsub_rn(acc, __ldg(v + i - G));
    const Wall w = wall_of(g, row, col);
    store_with_walls(div, i, G, __fmul_rn(scale, acc), w, 1.0f, 1.0f);
    if (write_p) store_with_walls(p, i, G, 0.0f, w, 1.0f, 1.0f);
}

// ---- lastProject (gradient subtract) ------------------------------------------------------------
__global__ void last_project_kernel(float *__restrict__ u, float *__restrict__ v, const float *__restrict__ p, Geom g,
                                    float h)
{
    int lo, hi;
    interior_rows(g, lo, hi);
    const int col = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int row = blockIdx.y * blockDim.y + threadIdx.y + lo;
    if (col > g.N || row >= hi) return;
    const size_t G = (size_t)g.G;
    const size_t i = (size_t)(row - g.row_base) * G + col;
    // FluidSequential.c:167-168: u -= (0.5f*(p_r - p_l)) / h
    const float gx = __fmul_rn(0.5f, __fsub_rn(__ldg(p + i + 1), __ldg(p + i - 1)));
    const float gy = __fmul_rn(0.5f, __fsub_rn(__ldg(p + i + G), __ldg(p + i - G)));
    const float nu = __fsub_rn(u[i], __fdiv_rn(gx, h));
    const float nv = __fsub_rn(v[i], __fdiv_rn(gy, h));
    const Wall w = wall_of(g, row, col);
    store_with_walls(u, i, G, nu, w, -1.0f, 1.0f);   // set_bnd(1, u)
    store_with_walls(v, i, G, nv, w, 1.0f, -1.0f);   // set_bnd(2, v)
}

// ---- synthetic initial condition ---------------------------------------------------------------
__global__ void init_kernel(float *dens, float *dens_prev, float *u, float *u_prev, float *v, float *v_prev, Geom g,
                            uint64_t seed)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = blockIdx.y * blockDim.y + threadIdx.y + g.own_lo;
    if (col >= g.G || row >= g.own_hi) return;
    const size_t cell = (size_t)row * g.G + col;                       // GLOBAL cell id feeds the hash
    const size_t i = (size_t)(row - g.row_base) * g.G + col;
    const int mid = g.G / 2, half = g.G / 8;
    const bool inside = (col < mid + half) && (col >= mid - half) && (row < mid + half) && (row >= mid - half);
    if (dens_prev) dens_prev[i] = inside ? __fdiv_rn((float)hash100(seed, 0, cell), 1000.0f) : 0.0f;
    if (u_prev) u_prev[i] = __fdiv_rn((float)hash100(seed, 1, cell), 100.0f);
    if (v_prev) v_prev[i] = __fdiv_rn((float)hash100(seed, 2, cell), 100.0f);
    if (dens) dens[i] = 0.0f;
    if (u) u[i] = 0.0f;
    if (v) v[i] = 0.0f;
}
// float4 variant (G % 4 == 0): one thread writes four consecutive cells of each field.  The twelve IEEE divisions per thread
// (h / 100.0f, h / 1000.0f with h = 0..99) made the kernel compute-bound at 3.9 TB/s of stores: the 100 possible quotients of
// each divisor are formed once per block with the same __fdiv_rn and looked up (same bits).
__global__ void __launch_bounds__(256) init4_kernel(float *dens, float *dens_prev, float *u, float *u_prev, float *v,
                                                    float *v_prev, Geom g, uint64_t seed)
{
    __shared__ float q100[100], q1000[100];
    {
        const int t = threadIdx.y * blockDim.x + threadIdx.x;
        if (t < 100) { q100[t] = __fdiv_rn((float)t, 100.0f); q1000[t] = __fdiv_rn((float)t, 1000.0f); }
    }
    __syncthreads();
    const int col = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int row = blockIdx.y * blockDim.y + threadIdx.y + g.own_lo;
    if (col >= g.G || row >= g.own_hi) return;
    const size_t cell = (size_t)row * g.G + col;
    const size_t i = (size_t)(row - g.row_base) * g.G + col;
    const int mid = g.G / 2, half = g.G / 8;
    const bool rin = (row < mid + half) && (row >= mid - half);
    float d[4], a[4], b[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const bool inside = rin && (col + k < mid + half) && (col + k >= mid - half);
        d[k] = inside ? q1000[hash100(seed, 0, cell + k)] : 0.0f;
        a[k] = q100[hash100(seed, 1, cell + k)];
        b[k] = q100[hash100(seed, 2, cell + k)];
    }
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (dens_prev) *reinterpret_cast<float4 *>(dens_prev + i) = make_float4(d[0], d[1], d[2], d[3]);
    if (u_prev) *reinterpret_cast<float4 *>(u_prev + i) = make_float4(a[0], a[1], a[2], a[3]);
    if (v_prev) *reinterpret_cast<float4 *>(v_prev + i) = make_float4(b[0], b[1], b[2], b[3]);
    if (dens) *reinterpret_cast<float4 *>(dens + i) = z;
    if (u) *reinterpret_cast<float4 *>(u + i) = z;
    if (v) *reinterpret_cast<float4 *>(v + i) = z;
}

// ---- reductions (warp shuffle, then one atomic per block) ------------------------------------
__global__ void max_abs_kernel(const float *__restrict__ x, size_t first, size_t count, float *out)
{
    // integer order of |v|'s bit pattern == magnitude order, and every NaN pattern sorts above +inf: a NaN anywhere in the
    // field comes out as NaN (fmaxf would drop it, and a caller sizing an advection halo from the result would see 0)
    int m = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        m = max(m, __float_as_int(fabsf(x[first + i])));
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ int part[32];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = (threadIdx.x < (blockDim.x >> 5)) ? part[threadIdx.x] : 0;
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) atomicMax(reinterpret_cast<int *>(out), m);
    }
}

// sum over the owned interior cells of r^2, r = x0 - (beta*x - alpha*(x_l + x_r + x_u + x_d)): the residual of the system the
// lin_solve relaxes (FluidSequential.c:95-96 solved for x0).  A diagnostic, not part of the reference's arithmetic: every
// operation is binary64 (products of binary32 values are exact there), so the result does not depend on contraction or on
// the association of the neighbour sum beyond 1e-15 relative.
__global__ void residual_kernel(const float *__restrict__ x, const float *__restrict__ x0, Geom g, float alpha,
                                float beta, double *out)
{
    int lo, hi;
    interior_rows(g, lo, hi);
    double acc = 0.0;
    const size_t G = (size_t)g.G;
    const double al = (double)alpha, be = (double)beta;
    for (int row = lo + blockIdx.y; row < hi; row += gridDim.y)
        for (int col = 1 + blockIdx.x * blockDim.x + threadIdx.x; col <= g.N; col += gridDim.x * blockDim.x) {
            const size_t i = (size_t)(row - g.row_base) * G + col;
            const double nb = __dadd_rn(__dadd_rn((double)x[i - 1], (double)x[i + 1]), __dadd_rn((double)x[i - G], (double)x[i + G]));
            const double r = __dsub_rn((double)x0[i], __dsub_rn(__dmul_rn(be, (double)x[i]), __dmul_rn(al, nb)));
            acc = __fma_rn(r, r, acc);
        }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ double part[32];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        acc = (threadIdx.x < (blockDim.x >> 5)) ? part[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (threadIdx.x == 0) atomicAdd(out, acc);
    }
}

// ================================================================================================
// Row-vectorised variants for G % 4 == 0 (every BASELINE width): one thread owns four adjacent
// columns [c, c+4) of one row, c % 4 == 0, so every access is an aligned float4 and the wall
// columns 0 / N+1 sit in the same thread as interior columns 1 / N (set_bnd on columns is a
// register operation).  Horizontal neighbours c-1 / c+4 come from the adjacent lanes by shuffle;
// only the first / last lane of a warp loads them as scalars.  A warp covers 128 consecutive
// columns of a row (512 contiguous bytes per access).
struct Row4 {
    float4 v;
    float l, r;   // values at columns c-1 and c+4 (0 outside the grid)
};
__device__ __forceinline__ Row4 load_row4(const float *__restrict__ row, int c, int G, int lane)
{
    Row4 o;
    o.v = __ldg(reinterpret_cast<const float4 *>(row + c));
    o.l = __shfl_up_sync(0xffffffffu, o.v.w, 1);
    o.r = __shfl_down_sync(0xffffffffu, o.v.x, 1);
    if (lane == 0) o.l = (c > 0) ? __ldg(row + c - 1) : 0.0f;
    if (lane == 31) o.r = (c + 4 < G) ? __ldg(row + c + 4) : 0.0f;
    return o;
}
// store four interior-row values plus the wall cells they determine (set_bnd(b) fused)
__device__ __forceinline__ void store_row4_walls(float *__restrict__ x, const Geom &g, int row, int c, float4 o, float sx, float sy)
{
    const size_t G = (size_t)g.G;
    const bool ownsL = (c == 0), ownsR = (c + 4 == g.G);
    if (ownsL) o.x = __fmul_rn(sx, o.y);
    if (ownsR) o.w = __fmul_rn(sx, o.z);
    float *dst = x + (size_t)(row - g.row_base) * G + c;
    *reinterpret_cast<float4 *>(dst) = o;
    const bool top = (row == 1) && (g.own_lo == 0), bot = (row == g.N) && (g.own_hi == g.G);
    if (top | bot) {
        float4 w = make_float4(__fmul_rn(sy, o.x), __fmul_rn(sy, o.y), __fmul_rn(sy, o.z), __fmul_rn(sy, o.w));
        if (ownsL) w.x = __fmul_rn(0.5f, __fadd_rn(w.y, o.x));
        if (ownsR) w.w = __fmul_rn(0.5f, __fadd_rn(w.z, o.w));
        if (top) *reinterpret_cast<float4 *>(dst - G) = w;
        if (bot) *reinterpret_cast<float4 *>(dst + G) = w;
    }
}
#define SF_ROW4_PROLOGUE                                                      \
    int lo, hi;                                                               \
    interior_rows(g, lo, hi);                                                 \
    const int lane = threadIdx.x & 31;                                        \
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;                \
    const int row = blockIdx.y * blockDim.y + threadIdx.y + lo;               \
    if (row >= hi) return;            /* uniform per warp: a warp is one row */\
    const bool active = c < g.G;                                              \
    const int cs = active ? c : g.G - 4;   /* idle lanes mirror the last float4: loads stay legal */ \
    const size_t G = (size_t)g.G;                                             \
    const size_t rowoff = (size_t)(row - g.row_base) * G;

__global__ void __launch_bounds__(256) divergence4_kernel(const float *__restrict__ u, const float *__restrict__ v,
                                                          float *__restrict__ p, float *__restrict__ div, Geom g,
                                                          float scale, int write_p)
{
    SF_ROW4_PROLOGUE
    const Row4 U = load_row4(u + rowoff, cs, g.G, lane);
    const float4 vd = __ldg(reinterpret_cast<const float4 *>(v + rowoff + G + cs));
    const float4 vu = __ldg(reinterpret_cast<const float4 *>(v + rowoff - G + cs));
    if (!active) return;
    // FluidSequential.c:151-152: (-0.5f*h) * (((u_r - u_l) + v_d) - v_u)
    float4 o;
    o.x = __fmul_rn(scale, __fsub_rn(__fadd_rn(__fsub_rn(U.v.y, U.l), vd.x), vu.x));
    o.y = __fmul_rn(scale, __fsub_rn(__fadd_rn(__fsub_rn(U.v.z, U.v.x), vd.y), vu.y));
    o.z = __fmul_rn(scale, __fsub_rn(__fadd_rn(__fsub_rn(U.v.w, U.v.y), vd.z), vu.z));
    o.w = __fmul_rn(scale, __fsub_rn(__fadd_rn(__fsub_rn(U.r, U.v.z), vd.w), vu.w));
    store_row4_walls(div, g, row, c, o, 1.0f, 1.0f);
    if (write_p) store_row4_walls(p, g, row, c, make_float4(0.f, 0.f, 0.f, 0.f), 1.0f, 1.0f);
}

__global__ void __launch_bounds__(256) last_project4_kernel(float *__restrict__ u, float *__restrict__ v,
                                                            const float *__restrict__ p, Geom g, float h)
{
    SF_ROW4_PROLOGUE
    const Row4 P = load_row4(p + rowoff, cs, g.G, lane);
    const float4 pd = __ldg(reinterpret_cast<const float4 *>(p + rowoff + G + cs));
    const float4 pu = __ldg(reinterpret_cast<const float4 *>(p + rowoff - G + cs));
    if (!active) return;
    float4 uu = *reinterpret_cast<const float4 *>(u + rowoff + c);
    float4 vv = *reinterpret_cast<const float4 *>(v + rowoff + c);
    // FluidSequential.c:167-168: u -= (0.5f*(p_r - p_l)) / h ; v -= (0.5f*(p_d - p_u)) / h
    uu.x = __fsub_rn(uu.x, __fdiv_rn(__fmul_rn(0.5f, __fsub_rn(P.v.y, P.l)), h));
    uu.y = __fsub_rn(uu.y, __fdiv_rn(__fmul_rn(0.5f, __fsub_rn(P.v.z, P.v.x)), h));
    uu.z = __fsub_rn(uu.z, __fdiv_rn(__fmul_rn(0.5f, __fsub_rn(P.v.w, P.v.y)), h));
    uu.w = __fsub_rn(uu.w, __fdiv_rn(__fmul_rn(0.5f, __fsub_rn(P.r, P.v.z)), h));
    vv.x = __fsub_rn(vv.x, __fdiv_rn(__fmul_rn(0.5f, __fsub_rn(pd.x, pu.x)), h));
    vv.y = __fsub_rn(vv.y, __fdiv_rn(__fmul_rn(0.5f, __fsub_rn(pd.y, pu.y)), h));
    vv.z = __fsub_rn(vv.z, __fdiv_rn(__fmul_rn(0.5f, __fsub_rn(pd.z, pu.z)), h));
    vv.w = __fsub_rn(vv.w, __fdiv_rn(__fmul_rn(0.5f, __fsub_rn(pd.w, pu.w)), h));
    store_row4_walls(u, g, row, c, uu, -1.0f, 1.0f);   // set_bnd(1, u)
    store_row4_walls(v, g, row, c, vv, 1.0f, -1.0f);   // set_bnd(2, v)
}

// the back-trace of one cell (FluidSequential.c:114-124): position clamped to [0.5, N + 0.5] in both directions
__device__ __forceinline__ void trace_back(int row, int col, float uu, float vv, float dt0, float hiC, float &px, float &py)
{
    px = __fsub_rn((float)col, __fmul_rn(dt0, uu));
    py = __fsub_rn((float)row, __fmul_rn(dt0, vv));
    if (px < 0.5f) px = 0.5f;
    if (px > hiC) px = hiC;
    if (py < 0.5f) py = 0.5f;
    if (py > hiC) py = hiC;
}
// the interpolation of :125-137 from the four source cells (a10 = next row, a01 = next column)
__device__ __forceinline__ float bilinear4(float a00, float a10, float a01, float a11, float wx0, float wx1, float wy0, float wy1)
{
    return __fadd_rn(__fmul_rn(wx0, __fadd_rn(__fmul_rn(wy0, a00), __fmul_rn(wy1, a10))),
                     __fmul_rn(wx1, __fadd_rn(__fmul_rn(wy0, a01), __fmul_rn(wy1, a11))));
}
// bilinear gather at a traced position from global memory; NF source fields share the trace
template <int NF>
__device__ __forceinline__ void gather_global(const float *__restrict__ srcA, const float *__restrict__ srcB, const Geom &g,
                                              float px, float py, float &oA, float &oB)
{
    const int c0 = (int)px, r0 = (int)py;
    const float wx1 = __fsub_rn(px, (float)c0), wx0 = __fsub_rn(1.0f, wx1);
    const float wy1 = __fsub_rn(py, (float)r0), wy0 = __fsub_rn(1.0f, wy1);
    const size_t G = (size_t)g.G;
    // a slab context stores rows [row_base, row_base + rows): a back-trace that leaves them (the caller's ghost rows were
    // too few -- slab.SlabSolver verifies the reach afterwards and raises) reads the nearest stored rows, never outside the array
    const int rs = min(max(r0, g.row_base), g.row_base + g.rows - 2);
    const size_t j = (size_t)(rs - g.row_base) * G + c0;
    oA = bilinear4(__ldg(srcA + j), __ldg(srcA + j + G), __ldg(srcA + j + 1), __ldg(srcA + j + G + 1), wx0, wx1, wy0, wy1);
    if (NF == 2)
        oB = bilinear4(__ldg(srcB + j), __ldg(srcB + j + G), __ldg(srcB + j + 1), __ldg(srcB + j + G + 1), wx0, wx1, wy0, wy1);
}
// one back-trace + bilinear gather (FluidSequential.c:114-137)
template <int NF>
__device__ __forceinline__ void advect_cell(const float *__restrict__ srcA, const float *__restrict__ srcB, const Geom &g,
                                            int row, int col, float uu, float vv, float dt0, float hiC, float &oA, float &oB)
{
    float px, py;
    trace_back(row, col, uu, vv, dt0, hiC, px, py);
    gather_global<NF>(srcA, srcB, g, px, py, oA, oB);
}

// ---- advect on a peer-memory slab ---------------------------------------------------------------
// Same arithmetic; the gather's two source rows are resolved per cell: a row outside this slab's
// owned range is read straight from the neighbour GPU that owns it (peer loads over NVLink), so the
// advection needs no halo exchange and no bound on the back-trace other than "not beyond the
// neighbour's slab" (which sets an error bit instead of reading out of bounds).
struct PeerView {
    const float *loc, *up, *dn;
};
// which array holds global row r (0 = this slab, 1 = up neighbour, 2 = down neighbour) and the row's offset in it
__device__ __forceinline__ int peer_row(const Geom &g, const PeerGeom &pg, int r, size_t &off)
{
    const size_t G = (size_t)g.G;
    if (r < g.own_lo) {
        if (r < pg.up_lo) { atomicOr(pg.error, 2u); r = pg.up_lo; }          // SF_SLAB_ERR_REACH
        off = (size_t)(r - pg.up_row_base) * G;
        return 1;
    }
    if (r >= g.own_hi) {
        if (r >= pg.dn_hi) { atomicOr(pg.error, 2u); r = pg.dn_hi - 1; }
        off = (size_t)(r - pg.dn_row_base) * G;
        return 2;
    }
    off = (size_t)(r - g.row_base) * G;
    return 0;
}
__device__ __forceinline__ const float *peer_base(const PeerView &f, int which) { return which == 0 ? f.loc : (which == 1 ? f.up : f.dn); }
__device__ __forceinline__ float bilinear(const float *p0, const float *p1, float wx0, float wx1, float wy0, float wy1)
{
    const float a00 = __ldg(p0), a10 = __ldg(p1), a01 = __ldg(p0 + 1), a11 = __ldg(p1 + 1);
    return __fadd_rn(__fmul_rn(wx0, __fadd_rn(__fmul_rn(wy0, a00), __fmul_rn(wy1, a10))),
                     __fmul_rn(wx1, __fadd_rn(__fmul_rn(wy0, a01), __fmul_rn(wy1, a11))));
}
template <int NF>
__device__ __forceinline__ void gather_peer(const PeerView &sA, const PeerView &sB, const Geom &g, const PeerGeom &pg,
                                            float px, float py, float &oA, float &oB)
{
    const int c0 = (int)px, r0 = (int)py;
    const float wx1 = __fsub_rn(px, (float)c0), wx0 = __fsub_rn(1.0f, wx1);
    const float wy1 = __fsub_rn(py, (float)r0), wy0 = __fsub_rn(1.0f, wy1);
    if (r0 >= g.own_lo && r0 + 1 < g.own_hi) {
        // both source rows are this slab's own (all but the cells within reach of the slab edges)
        const size_t G = (size_t)g.G;
        const size_t j = (size_t)(r0 - g.row_base) * G + c0;
        oA = bilinear(sA.loc + j, sA.loc + j + G, wx0, wx1, wy0, wy1);
        if (NF == 2) oB = bilinear(sB.loc + j, sB.loc + j + G, wx0, wx1, wy0, wy1);
        return;
    }
    size_t off0, off1;
    const int w0 = peer_row(g, pg, r0, off0), w1 = peer_row(g, pg, r0 + 1, off1);
    oA = bilinear(peer_base(sA, w0) + off0 + c0, peer_base(sA, w1) + off1 + c0, wx0, wx1, wy0, wy1);
    if (NF == 2) oB = bilinear(peer_base(sB, w0) + off0 + c0, peer_base(sB, w1) + off1 + c0, wx0, wx1, wy0, wy1);
}
template <int NF>
__device__ __forceinline__ void advect_cell_peer(const PeerView &sA, const PeerView &sB, const Geom &g, const PeerGeom &pg,
                                                 int row, int col, float uu, float vv, float dt0, float hiC, float &oA, float &oB)
{
    float px, py;
    trace_back(row, col, uu, vv, dt0, hiC, px, py);
    gather_peer<NF>(sA, sB, g, pg, px, py, oA, oB);
}

// store one row's four lane-strided cells per field, with set_bnd fused (see advect_lanes_kernel)
template <int NF>
__device__ __forceinline__ void store_lanes_row(float *__restrict__ dA, float *__restrict__ dB, const Geom &g, int row, int seg,
                                                int lane, float (&oA)[4], float (&oB)[4], int bA)
{
    const size_t G = (size_t)g.G;
    const size_t rowoff = (size_t)(row - g.row_base) * G;
    const bool top = (row == 1) && (g.own_lo == 0), bot = (row == g.N) && (g.own_hi == g.G);
    if (seg > 0 && seg + 128 < g.G && !(top | bot)) {
        // warp-uniform fast path (all but the first / last segment of a row and the two rows next to a wall row):
        // no wall column, no idle lane, no wall row -- four coalesced stores per field
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            dA[rowoff + seg + 32 * k + lane] = oA[k];
            if (NF == 2) dB[rowoff + seg + 32 * k + lane] = oB[k];
        }
        return;
    }
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        float *o = f == 0 ? oA : oB;
        float *dst = (f == 0 ? dA : dB) + rowoff;
        const float sx = (f == 0 ? bA == 1 : false) ? -1.0f : 1.0f;     // field B of the pair is v: b = 2
        const float sy = (f == 0 ? bA == 2 : true) ? -1.0f : 1.0f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int col = seg + 32 * k + lane;
            const float nxt = __shfl_down_sync(0xffffffffu, o[k], 1), prv = __shfl_up_sync(0xffffffffu, o[k], 1);
            const bool wallL = (col == 0), wallR = (col == g.G - 1);
            float val = o[k];
            if (wallL) val = __fmul_rn(sx, nxt);                       // x[row][0]   = sx * x[row][1]
            if (wallR) val = __fmul_rn(sx, prv);                       // x[row][N+1] = sx * x[row][N]
            if (col < g.G) {
                dst[col] = val;
                if (top | bot) {
                    float w = __fmul_rn(sy, val);                      // wall row = sy * adjacent interior row
                    if (wallL) w = __fmul_rn(0.5f, __fadd_rn(__fmul_rn(sy, nxt), val));   // corner = .5 * (wall-row nbr + wall-col nbr)
                    if (wallR) w = __fmul_rn(0.5f, __fadd_rn(__fmul_rn(sy, prv), val));
                    if (top) (dst - G)[col] = w;
                    if (bot) (dst + G)[col] = w;
                }
            }
        }
    }
}

// ---- advect, lane-strided mapping (the product path for G % 4 == 0) ---------------------------------
// A warp owns 128 consecutive columns of one row and lane l handles the four columns seg + 32 k + l (k = 0..3), NOT four
// adjacent ones: every load, gather and store instruction of the warp then touches 32 CONSECUTIVE cells -- or their
// back-traced sources, which are consecutive up to the (smooth) variation of the velocity field.  With four adjacent cells
// per thread (advect4_kernel, round 1) each gather request was strided by 16 bytes and touched 16-27 sectors: ncu showed the
// L1 tag stage at 76 % and DRAM at 38 % (profiles/r02/s1_86a0b52_stage_ncu.csv); one cell per thread coalesces as well but
// keeps too few loads in flight per thread to cover two dependent DRAM round trips.  Same arithmetic per cell, same bits.
// set_bnd is fused as before: the lane that holds wall column 0 / N+1 takes the adjacent interior value from lane 1 / lane-1
// by shuffle; the warp of row 1 / row N also writes the wall row and its corners.
// PEER = true: the two source rows of a back-trace may lie in a neighbour GPU's slab (see advect_cell_peer).
template <int NF, bool PEER>
__global__ void __launch_bounds__(256) advect_lanes_kernel(float *__restrict__ dA, float *__restrict__ dB, PeerView sA, PeerView sB,
                                                           const float *__restrict__ u, const float *__restrict__ v, Geom g,
                                                           PeerGeom pg, float dt0, int bA)
{
    int lo, hi;
    interior_rows(g, lo, hi);
    const int lane = threadIdx.x;                                   // blockDim.x == 32: one warp per row segment
    const int seg = blockIdx.x * 128;
    const int row = blockIdx.y * blockDim.y + threadIdx.y + lo;
    if (row >= hi) return;                                          // uniform per warp
    const size_t G = (size_t)g.G;
    const size_t rowoff = (size_t)(row - g.row_base) * G;
    const float hiC = (float)g.N + 0.5f;
    float uu[4], vv[4], oA[4], oB[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {                                   // idle lanes (col >= G) mirror the last column: loads stay legal
        const int cl = min(seg + 32 * k + lane, g.G - 1);
        uu[k] = __ldg(u + rowoff + cl);
        vv[k] = __ldg(v + rowoff + cl);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        // columns 0 and N+1 are wall cells: their values are replaced below, but the trace still has to stay inside the
        // array, which the clamp to [0.5, N+0.5] guarantees
        const int cl = min(seg + 32 * k + lane, g.G - 1);
        oA[k] = 0.0f; oB[k] = 0.0f;
        if (PEER) advect_cell_peer<NF>(sA, sB, g, pg, row, cl, uu[k], vv[k], dt0, hiC, oA[k], oB[k]);
        else advect_cell<NF>(sA.loc, sB.loc, g, row, cl, uu[k], vv[k], dt0, hiC, oA[k], oB[k]);
    }
    store_lanes_row<NF>(dA, dB, g, row, seg, lane, oA, oB, bA);
}

// [emu-cut-begin] (tools/emu compiles this file for the host up to here: the TMA kernel has no host twin; its arithmetic is
// trace_back / bilinear4 / store_lanes_row above, which the emulated kernels share)
// ---- advect with the source tile staged in shared memory by the TMA unit ---------------------------
// The lane-strided kernel above moves the minimum DRAM traffic but is bound by the L1 tag stage: the 32 sources of one gather
// instruction spread over 8-10 rows of the source field (12-17 sectors per request, profiles/r02/final_*_ncu_advect_lanes_*).
// Here a CTA (8 warps) owns a tile of 8 * RPW rows x 128 columns (RPW = 2 or 4 rows per warp) and works in three steps:
//  1. every thread loads u, v of its 4 * RPW cells (lane-strided as above), traces them back and the CTA reduces the bounding box
//     of the traces (redux.sync + one shared-memory exchange);
//  2. if the box fits (<= AT_BOXW columns, <= 8 * maxsub rows, inside the rows this slab stores / owns) ONE thread asks the TMA
//     unit for it: 2-D tiled tensor copies (cp.async.bulk.tensor.2d, SASS UTMALDG) of AT_SUB rows x AT_BOXW columns each, only as
//     many as the box is high, completing on an mbarrier; cells beyond the array come back as zeros and are never read;
//  3. the gathers become shared-memory loads (pitch 160 words = 0 mod 32 banks: the lanes of a warp sit in consecutive columns
//     +- their jitter, so conflicts only arise where neighbouring traces share a column but not a row).
// A tile whose box does not fit (a velocity field rougher than ~+-12 cells of back-trace within a tile, or a trace that
// leaves the slab in the PEER build) falls back to the global / peer gathers of the kernel above, as a whole CTA.
// Same arithmetic per cell (trace_back, bilinear4), same bits.
#ifndef SF_AT_MINB
#define SF_AT_MINB 3      // CTAs per SM the register allocation allows
#endif
constexpr int AT_BOXW = 160;    // box width in cells
constexpr int AT_SUB = 8;       // rows per tensor copy
constexpr int AT_MAXSUB = 8;    // at most 64 source rows per field in shared memory


__device__ __forceinline__ void at_mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void at_mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void at_mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
// one box of the tensor map's size whose first cell is (column c, stored row r) -> shared memory, completing on `bar`
__device__ __forceinline__ void at_tensor_load(void *smem_dst, const CUtensorMap *tm, int c, int r, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(tm), "r"(c), "r"(r), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

// truncation of a clamped trace coordinate (0.5 <= x <= N + 0.5 < 2^23) without the conversion unit: adding 2^23 with rounding
// toward zero leaves the integer part in the low mantissa bits.  `fl` = (float)(int)x and `i` = (int)x, exactly.
__device__ __forceinline__ void trunc_pos(float x, float &fl, int &i)
{
    const float t = __fadd_rz(x, 8388608.0f);
    i = __float_as_int(t) - 0x4B000000;
    fl = __fsub_rn(t, 8388608.0f);
}

template <int NF, bool PEER, int RPW>
__global__ void __launch_bounds__(256, (RPW == 2 ? 4 : SF_AT_MINB)) advect_tile_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                          float *__restrict__ dA, float *__restrict__ dB, PeerView sA, PeerView sB,
                                                          const float *__restrict__ u, const float *__restrict__ v, Geom g,
                                                          PeerGeom pg, float dt0, int bA, int maxsub, unsigned int *__restrict__ stats)
{
    extern __shared__ unsigned char at_dyn[];
    __shared__ int s_red[8][4];
    __shared__ __align__(8) uint64_t s_bar;
    // tensor copies need a 128-byte aligned destination
    float *tile = reinterpret_cast<float *>(at_dyn + ((128u - ((unsigned)__cvta_generic_to_shared(at_dyn) & 127u)) & 127u));
    int lo, hi;
    interior_rows(g, lo, hi);
    const int lane = threadIdx.x, w = threadIdx.y;
    const int seg = blockIdx.x * 128;
    const int row0 = blockIdx.y * (8 * RPW) + lo + RPW * w;
    const size_t G = (size_t)g.G;
    const float hiC = (float)g.N + 0.5f;
    if (threadIdx.x == 0 && threadIdx.y == 0) at_mbar_init(&s_bar, 1);

    // 1. traces of the thread's 4 * RPW cells (rows past the slab's last row and idle lanes mirror the last row / column: loads stay legal
    //    and the box does not grow)
    float px[RPW][4], py[RPW][4];
    {
        float uu[RPW][4], vv[RPW][4];
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr) {
            const size_t rowoff = (size_t)(min(row0 + rr, hi - 1) - g.row_base) * G;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int cl = min(seg + 32 * k + lane, g.G - 1);
                uu[rr][k] = __ldg(u + rowoff + cl);
                vv[rr][k] = __ldg(v + rowoff + cl);
            }
        }
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                trace_back(min(row0 + rr, hi - 1), min(seg + 32 * k + lane, g.G - 1), uu[rr][k], vv[rr][k], dt0, hiC, px[rr][k], py[rr][k]);
    }
    float fcmin = 3.0e38f, fcmax = 0.0f, frmin = 3.0e38f, frmax = 0.0f;
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            // the extremes of the positions give the extremes of their integer parts (truncation is monotone)
            fcmin = fminf(fcmin, px[rr][k]); fcmax = fmaxf(fcmax, px[rr][k]);
            frmin = fminf(frmin, py[rr][k]); frmax = fmaxf(frmax, py[rr][k]);
            // a NaN velocity passes the clamps (and fminf / fmaxf): such a tile must not index shared memory -- it takes the
            // gather path, where (int)NaN = 0 reads cell (0, 0) like the other kernels
            if (!(px[rr][k] <= hiC) || !(py[rr][k] <= hiC)) frmax = 1.0e9f;
        }
    int cmin = (int)fcmin, cmax = (int)fcmax, rmin = (int)frmin, rmax = (int)frmax;
    cmin = __reduce_min_sync(0xffffffffu, cmin); cmax = __reduce_max_sync(0xffffffffu, cmax);
    rmin = __reduce_min_sync(0xffffffffu, rmin); rmax = __reduce_max_sync(0xffffffffu, rmax);
    if (lane == 0) { s_red[w][0] = cmin; s_red[w][1] = cmax; s_red[w][2] = rmin; s_red[w][3] = rmax; }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        cmin = min(cmin, s_red[i][0]); cmax = max(cmax, s_red[i][1]);
        rmin = min(rmin, s_red[i][2]); rmax = max(rmax, s_red[i][3]);
    }
    // 2. does the box [rmin, rmax + 1] x [cmin, cmax + 1] fit?  (the TMA unit wants the first cell of a box 16-byte aligned:
    //    an odd first column is an illegal instruction -- tools/micro/tma_probe.cu)
    cmin &= ~3;
    const int span_r = rmax + 2 - rmin;
    bool fit = (cmax + 2 - cmin <= AT_BOXW) && (span_r <= AT_SUB * maxsub);
    if (PEER) fit = fit && rmin >= g.own_lo && rmax + 1 < g.own_hi;                 // see gather_peer
    else fit = fit && rmin >= g.row_base && rmax + 1 < g.row_base + g.rows;         // see gather_global
    if (!fit) {
        if (threadIdx.x == 0 && threadIdx.y == 0) atomicAdd(stats + 1, 1u);      // CTAs served by the fallback
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr) {
            const int row = row0 + rr;
            if (row >= hi) break;                                           // uniform per warp
            float oA[4], oB[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                oA[k] = 0.0f; oB[k] = 0.0f;
                if (PEER) gather_peer<NF>(sA, sB, g, pg, px[rr][k], py[rr][k], oA[k], oB[k]);
                else gather_global<NF>(sA.loc, sB.loc, g, px[rr][k], py[rr][k], oA[k], oB[k]);
            }
            store_lanes_row<NF>(dA, dB, g, row, seg, lane, oA, oB, bA);
        }
        return;
    }
    const int nsub = (span_r + AT_SUB - 1) / AT_SUB;
    float *tA = tile, *tB = tile + (size_t)maxsub * AT_SUB * AT_BOXW;
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        atomicAdd(stats, 1u);          // CTAs served by the TMA box
        at_mbar_expect_tx(&s_bar, (unsigned)(nsub * AT_SUB * AT_BOXW * sizeof(float) * NF));
        for (int s = 0; s < nsub; ++s) {
            at_tensor_load(tA + s * AT_SUB * AT_BOXW, &tmA, cmin, rmin - g.row_base + s * AT_SUB, &s_bar);
            if (NF == 2) at_tensor_load(tB + s * AT_SUB * AT_BOXW, &tmB, cmin, rmin - g.row_base + s * AT_SUB, &s_bar);
        }
    }
    at_mbar_wait(&s_bar, 0);
    // 3. gathers from shared memory
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr) {
        const int row = row0 + rr;
        if (row >= hi) break;                                               // uniform per warp
        float oA[4], oB[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float x = px[rr][k], y = py[rr][k];
            int c0, r0;
            float fc, fr;
            trunc_pos(x, fc, c0); trunc_pos(y, fr, r0);
            const float wx1 = __fsub_rn(x, fc), wx0 = __fsub_rn(1.0f, wx1);
            const float wy1 = __fsub_rn(y, fr), wy0 = __fsub_rn(1.0f, wy1);
            const int j = (r0 - rmin) * AT_BOXW + (c0 - cmin);
            oA[k] = bilinear4(tA[j], tA[j + AT_BOXW], tA[j + 1], tA[j + AT_BOXW + 1], wx0, wx1, wy0, wy1);
            oB[k] = 0.0f;
            if (NF == 2) oB[k] = bilinear4(tB[j], tB[j + AT_BOXW], tB[j + 1], tB[j + AT_BOXW + 1], wx0, wx1, wy0, wy1);
        }
        store_lanes_row<NF>(dA, dB, g, row, seg, lane, oA, oB, bA);
    }
}

// [emu-cut-end]

inline bool row4_ok(const Geom &g, std::initializer_list<const void *> ptrs)
{
    if (g.G % 4 != 0) return false;
    for (const void *p : ptrs)
        if ((uintptr_t)p % 16 != 0) return false;
    return true;
}
inline dim3 row4_grid(const Geom &g, dim3 block, int rows)
{
    const int per_block = block.x * 4;
    return dim3((g.G + per_block - 1) / per_block, (rows + block.y - 1) / block.y);
}

// advect_lanes_kernel: block (32, 8) = eight rows of one 128-column segment
inline dim3 lanes_grid(const Geom &g, int rows) { return dim3((g.G + 127) / 128, (rows + 7) / 8); }
inline dim3 cell_grid(const Geom &g, dim3 block, int rows) { return dim3((g.N + block.x - 1) / block.x, (rows + block.y - 1) / block.y); }
inline int interior_row_count(const Geom &g)
{
    const int lo = g.own_lo > 1 ? g.own_lo : 1, hi = g.own_hi < g.N + 1 ? g.own_hi : g.N + 1;
    return hi > lo ? hi - lo : 0;
}

// [emu-cut-begin]
// ---- host side of advect_tile_kernel -----------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn tensor_map_encoder()
{
    static const EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        (void)cudaGetLastError();
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}
// the locally stored rows of one field as a 2-D tensor (columns fastest), box = AT_SUB rows x AT_BOXW columns, zeros out of bounds
inline bool make_field_map(CUtensorMap *tm, const float *field, const Geom &g)
{
    const EncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)g.G, (cuuint64_t)g.rows};
    const cuuint64_t strides[1] = {(cuuint64_t)g.G * sizeof(float)};
    const cuuint32_t box[2] = {AT_BOXW, AT_SUB};
    const cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(field), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
inline bool tile_ok(const Geom &g, int tile, const unsigned int *stats, std::initializer_list<const void *> srcs)
{
    if (tile <= 0 || !stats || g.G % 4 != 0 || g.G < 2 * AT_BOXW || g.G > (1 << 22) || g.rows < AT_SUB * AT_MAXSUB) return false;
    for (const void *p : srcs)
        if ((uintptr_t)p % 16 != 0) return false;
    return true;
}
inline size_t tile_smem(int nf, int maxsub) { return (size_t)nf * maxsub * AT_SUB * AT_BOXW * sizeof(float) + 128; }
inline dim3 tile_grid(const Geom &g, int rows, int rpw) { return dim3((g.G + 127) / 128, (rows + 8 * rpw - 1) / (8 * rpw)); }
// SF_OPT_ADVECT_TILE value -> rows per warp and 8-row copies of shared memory per field: 1 = the default (see below),
// 2..8 = 32-row tiles with that many copies, 12..18 = 16-row tiles with (value - 10) copies
#ifndef SF_AT_DEFAULT
#define SF_AT_DEFAULT 15
#endif
inline void tile_shape(int tile, int nf, int &rpw, int &maxsub)
{
    // measured (profiles/r02/final_*_advect_ab.log): one field is fastest with 32-row tiles (0.226 vs 0.249 ms), the u, v pair
    // with 16-row tiles, whose two boxes leave room for four CTAs per SM
    if (tile < 2) tile = (nf == 1) ? AT_MAXSUB : SF_AT_DEFAULT;
    rpw = tile >= 10 ? 2 : 4;
    maxsub = tile >= 10 ? tile - 10 : tile;
    if (maxsub < 2) maxsub = 2;
    if (maxsub > AT_MAXSUB) maxsub = AT_MAXSUB;
}

template <int NF, bool PEER, int RPW>
cudaError_t launch_advect_tile_rpw(const CUtensorMap &tmA, const CUtensorMap &tmB, const Geom &g, int maxsub, float *dA, float *dB, PeerView sA,
                                   PeerView sB, const float *u, const float *v, PeerGeom pg, float dt0, int bA, int rows,
                                   unsigned int *stats, cudaStream_t st)
{
    // per device, and cheap: set on every launch (one process may drive several devices)
    cudaError_t e = cudaFuncSetAttribute(advect_tile_kernel<NF, PEER, RPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem(NF, AT_MAXSUB));
    if (e != cudaSuccess) return e;
    advect_tile_kernel<NF, PEER, RPW><<<tile_grid(g, rows, RPW), dim3(32, 8), tile_smem(NF, maxsub), st>>>(tmA, tmB, dA, dB, sA, sB, u, v, g, pg, dt0, bA,
                                                                                                      maxsub, stats);
    return cudaGetLastError();
}
template <int NF, bool PEER>
cudaError_t launch_advect_tile(const Geom &g, int tile, float *dA, float *dB, PeerView sA, PeerView sB, const float *u, const float *v,
                               PeerGeom pg, float dt0, int bA, int rows, unsigned int *stats, cudaStream_t st, bool &done)
{
    done = false;
    CUtensorMap tmA, tmB;
    if (!make_field_map(&tmA, sA.loc, g)) return cudaSuccess;       // no encoder (old driver): the caller takes the gather kernel
    if (NF == 2) { if (!make_field_map(&tmB, sB.loc, g)) return cudaSuccess; }
    else tmB = tmA;
    int rpw, maxsub;
    tile_shape(tile, NF, rpw, maxsub);
    done = true;
    if (rpw == 2) return launch_advect_tile_rpw<NF, PEER, 2>(tmA, tmB, g, maxsub, dA, dB, sA, sB, u, v, pg, dt0, bA, rows, stats, st);
    return launch_advect_tile_rpw<NF, PEER, 4>(tmA, tmB, g, maxsub, dA, dB, sA, sB, u, v, pg, dt0, bA, rows, stats, st);
}

// [emu-cut-end]

}  // namespace

void preload_stage_kernels()
{
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, set_bnd_kernel); cudaFuncGetAttributes(&a, add_source_kernel);
    cudaFuncGetAttributes(&a, advect_kernel<1>); cudaFuncGetAttributes(&a, advect_kernel<2>);
    cudaFuncGetAttributes(&a, advect_lanes_kernel<1, false>); cudaFuncGetAttributes(&a, advect_lanes_kernel<2, false>);
    cudaFuncGetAttributes(&a, advect_lanes_kernel<1, true>); cudaFuncGetAttributes(&a, advect_lanes_kernel<2, true>);
    cudaFuncGetAttributes(&a, advect_tile_kernel<1, false, 2>); cudaFuncGetAttributes(&a, advect_tile_kernel<2, false, 2>);
    cudaFuncGetAttributes(&a, advect_tile_kernel<1, true, 2>); cudaFuncGetAttributes(&a, advect_tile_kernel<2, true, 2>);
    cudaFuncGetAttributes(&a, advect_tile_kernel<1, false, 4>); cudaFuncGetAttributes(&a, advect_tile_kernel<2, false, 4>);
    cudaFuncGetAttributes(&a, advect_tile_kernel<1, true, 4>); cudaFuncGetAttributes(&a, advect_tile_kernel<2, true, 4>);
    cudaFuncGetAttributes(&a, divergence_kernel); cudaFuncGetAttributes(&a, divergence4_kernel);
    cudaFuncGetAttributes(&a, last_project_kernel); cudaFuncGetAttributes(&a, last_project4_kernel);
    cudaFuncGetAttributes(&a, init_kernel); cudaFuncGetAttributes(&a, init4_kernel);
    cudaFuncGetAttributes(&a, max_abs_kernel); cudaFuncGetAttributes(&a, residual_kernel);
    (void)cudaGetLastError();
}

cudaError_t launch_set_bnd(const Geom &g, int b, float *x, cudaStream_t st)
{
    set_bnd_kernel<<<(g.N + 255) / 256, 256, 0, st>>>(x, g, b == 1 ? -1.0f : 1.0f, b == 2 ? -1.0f : 1.0f);
    return cudaGetLastError();
}

cudaError_t launch_add_source(const Geom &g, int nfields, float *const *x, const float *const *s, float dt, cudaStream_t st)
{
    if (nfields < 1 || nfields > 3) return cudaErrorInvalidValue;
    AddSrcArgs A;
    for (int k = 0; k < 3; ++k) { A.x[k] = x[k < nfields ? k : 0]; A.s[k] = s[k < nfields ? k : 0]; }
    A.first = (size_t)(g.own_lo - g.row_base) * g.G;
    A.count = (size_t)(g.own_hi - g.own_lo) * g.G;
    A.dt = dt;
    bool vec = (A.first % 4 == 0) && (A.count % 4 == 0);
    for (int k = 0; k < nfields; ++k) vec = vec && ((uintptr_t)x[k] % 16 == 0) && ((uintptr_t)s[k] % 16 == 0);
    A.vec = vec ? 1 : 0;
    const size_t work = vec ? A.count / 4 : A.count;
    size_t blocks = (work + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    add_source_kernel<<<dim3((unsigned)blocks, nfields), 256, 0, st>>>(A);
    return cudaGetLastError();
}

cudaError_t launch_advect(const Geom &g, int b, float *d, const float *d0, const float *u, const float *v, float dt, int tile,
                          unsigned int *tile_stats, cudaStream_t st)
{
    const int rows = interior_row_count(g);
    if (rows == 0) return cudaSuccess;
    const dim3 block(64, 4);
    const float dt0 = dt * (float)g.N;   // FluidSequential.c:111, rounded once in binary32
    if (g.G % 4 == 0) {
        const PeerView sA{d0, nullptr, nullptr};
        if (tile_ok(g, tile, tile_stats, {d0})) {
            bool done;
            const cudaError_t e = launch_advect_tile<1, false>(g, tile, d, nullptr, sA, sA, u, v, PeerGeom(), dt0, b, rows, tile_stats, st, done);
            if (done || e != cudaSuccess) return e;
        }
        advect_lanes_kernel<1, false><<<lanes_grid(g, rows), dim3(32, 8), 0, st>>>(d, nullptr, sA, sA, u, v, g, PeerGeom(), dt0, b);
        return cudaGetLastError();
    }
    advect_kernel<1><<<cell_grid(g, block, rows), block, 0, st>>>(d, nullptr, d0, nullptr, u, v, g, dt0, b);
    return cudaGetLastError();
}

cudaError_t launch_advect_uv(const Geom &g, float *du, float *dv, const float *u0, const float *v0, float dt, int tile,
                             unsigned int *tile_stats, cudaStream_t st)
{
    const int rows = interior_row_count(g);
    if (rows == 0) return cudaSuccess;
    const dim3 block(64, 4);
    const float dt0 = dt * (float)g.N;
    if (g.G % 4 == 0) {
        const PeerView sA{u0, nullptr, nullptr}, sB{v0, nullptr, nullptr};
        if (tile_ok(g, tile, tile_stats, {u0, v0})) {
            bool done;
            const cudaError_t e = launch_advect_tile<2, false>(g, tile, du, dv, sA, sB, u0, v0, PeerGeom(), dt0, 1, rows, tile_stats, st, done);
            if (done || e != cudaSuccess) return e;
        }
        advect_lanes_kernel<2, false><<<lanes_grid(g, rows), dim3(32, 8), 0, st>>>(du, dv, sA, sB, u0, v0, g, PeerGeom(), dt0, 1);
        return cudaGetLastError();
    }
    advect_kernel<2><<<cell_grid(g, block, rows), block, 0, st>>>(du, dv, u0, v0, u0, v0, g, dt0, 1);
    return cudaGetLastError();
}

cudaError_t launch_advect_peer(const Geom &g, int b, float *d, const float *d0, const float *u, const float *v, float dt,
                               PeerSrc d0p, PeerGeom pg, int tile, unsigned int *tile_stats, cudaStream_t st)
{
    const int rows = interior_row_count(g);
    if (rows == 0) return cudaSuccess;
    if (g.G % 4 != 0) return cudaErrorInvalidValue;
    const float dt0 = dt * (float)g.N;   // FluidSequential.c:111
    const PeerView sA{d0, d0p.up, d0p.dn};
    if (tile_ok(g, tile, tile_stats, {d0})) {
        bool done;
        const cudaError_t e = launch_advect_tile<1, true>(g, tile, d, nullptr, sA, sA, u, v, pg, dt0, b, rows, tile_stats, st, done);
        if (done || e != cudaSuccess) return e;
    }
    advect_lanes_kernel<1, true><<<lanes_grid(g, rows), dim3(32, 8), 0, st>>>(d, nullptr, sA, sA, u, v, g, pg, dt0, b);
    return cudaGetLastError();
}

cudaError_t launch_advect_uv_peer(const Geom &g, float *du, float *dv, const float *u0, const float *v0, float dt,
                                  PeerSrc u0p, PeerSrc v0p, PeerGeom pg, int tile, unsigned int *tile_stats, cudaStream_t st)
{
    const int rows = interior_row_count(g);
    if (rows == 0) return cudaSuccess;
    if (g.G % 4 != 0) return cudaErrorInvalidValue;
    const float dt0 = dt * (float)g.N;
    const PeerView sA{u0, u0p.up, u0p.dn}, sB{v0, v0p.up, v0p.dn};
    if (tile_ok(g, tile, tile_stats, {u0, v0})) {
        bool done;
        const cudaError_t e = launch_advect_tile<2, true>(g, tile, du, dv, sA, sB, u0, v0, pg, dt0, 1, rows, tile_stats, st, done);
        if (done || e != cudaSuccess) return e;
    }
    advect_lanes_kernel<2, true><<<lanes_grid(g, rows), dim3(32, 8), 0, st>>>(du, dv, sA, sB, u0, v0, g, pg, dt0, 1);
    return cudaGetLastError();
}

cudaError_t launch_divergence(const Geom &g, const float *u, const float *v, float *p, float *div, int write_p, cudaStream_t st)
{
    const int rows = interior_row_count(g);
    if (rows == 0) return cudaSuccess;
    const dim3 block(64, 4);
    const float h = 1.0f / (float)g.N;
    const float scale = -0.5f * h;
    if (row4_ok(g, {u, v, p, div})) {
        const dim3 b4(32, 8);
        divergence4_kernel<<<row4_grid(g, b4, rows), b4, 0, st>>>(u, v, p, div, g, scale, write_p);
        return cudaGetLastError();
    }
    divergence_kernel<<<cell_grid(g, block, rows), block, 0, st>>>(u, v, p, div, g, scale, write_p);
    return cudaGetLastError();
}

cudaError_t launch_last_project(const Geom &g, float *u, float *v, const float *p, cudaStream_t st)
{
    const int rows = interior_row_count(g);
    if (rows == 0) return cudaSuccess;
    const dim3 block(64, 4);
    const float h = 1.0f / (float)g.N;
    if (row4_ok(g, {u, v, p})) {
        const dim3 b4(32, 8);
        last_project4_kernel<<<row4_grid(g, b4, rows), b4, 0, st>>>(u, v, p, g, h);
        return cudaGetLastError();
    }
    last_project_kernel<<<cell_grid(g, block, rows), block, 0, st>>>(u, v, p, g, h);
    return cudaGetLastError();
}

cudaError_t launch_init(const Geom &g, uint64_t seed, float *dens, float *dens_prev, float *u, float *u_prev, float *v,
                        float *v_prev, cudaStream_t st)
{
    const dim3 block(64, 4);
    bool vec = (g.G % 4 == 0);
    for (const float *p : {dens, dens_prev, u, u_prev, v, v_prev}) vec = vec && ((uintptr_t)p % 16 == 0);
    if (vec) {
        const dim3 grid4((g.G / 4 + 63) / 64, (g.own_hi - g.own_lo + 3) / 4);
        init4_kernel<<<grid4, block, 0, st>>>(dens, dens_prev, u, u_prev, v, v_prev, g, seed);
        return cudaGetLastError();
    }
    const dim3 grid((g.G + 63) / 64, (g.own_hi - g.own_lo + 3) / 4);
    init_kernel<<<grid, block, 0, st>>>(dens, dens_prev, u, u_prev, v, v_prev, g, seed);
    return cudaGetLastError();
}

cudaError_t launch_max_abs(const Geom &g, const float *x, float *dev_out, bool zero_first, cudaStream_t st)
{
    if (zero_first) {
        cudaError_t e = cudaMemsetAsync(dev_out, 0, sizeof(float), st);
        if (e != cudaSuccess) return e;
    }
    const size_t first = (size_t)(g.own_lo - g.row_base) * g.G, count = (size_t)(g.own_hi - g.own_lo) * g.G;
    size_t blocks = (count + 1023) / 1024;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    max_abs_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, first, count, dev_out);
    return cudaGetLastError();
}

cudaError_t launch_residual(const Geom &g, const float *x, const float *x0, float alpha, float beta, double *dev_out, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(dev_out, 0, sizeof(double), st);
    if (e != cudaSuccess) return e;
    const int rows = interior_row_count(g);
    if (rows == 0) return cudaSuccess;
    dim3 grid((g.N + 255) / 256 > 8 ? 8 : (g.N + 255) / 256, rows > 1024 ? 1024 : rows);
    residual_kernel<<<grid, 256, 0, st>>>(x, x0, g, alpha, beta, dev_out);
    return cudaGetLastError();
}

}  // namespace sf
