// Non-Jacobi stages of the stable-fluids step, each with set_bnd fused in (the thread that
// produces an interior cell next to a wall also writes the wall cell, and the four threads at
// the interior corners write the grid corners), so every field is read and written once per stage.
// Reference semantics: FluidSequential.c:62-82 (set_bnd, add_source), :107-141 (advect),
// :143-158 (computeDivergenceAndPressure), :161-173 (lastProject), :244-271 (initial condition).
// All arithmetic uses __f*_rn intrinsics so nvcc cannot contract mul+add into FMA: results are
// bit-identical to the reference's sequential build.
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
// float4 variant (G % 4 == 0): one thread writes four consecutive cells of each field
__global__ void __launch_bounds__(256) init4_kernel(float *dens, float *dens_prev, float *u, float *u_prev, float *v,
                                                    float *v_prev, Geom g, uint64_t seed)
{
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
        d[k] = inside ? __fdiv_rn((float)hash100(seed, 0, cell + k), 1000.0f) : 0.0f;
        a[k] = __fdiv_rn((float)hash100(seed, 1, cell + k), 100.0f);
        b[k] = __fdiv_rn((float)hash100(seed, 2, cell + k), 100.0f);
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

// one back-trace + bilinear gather (FluidSequential.c:114-137); NF source fields share the trace
template <int NF>
__device__ __forceinline__ void advect_cell(const float *__restrict__ srcA, const float *__restrict__ srcB, const Geom &g,
                                            int row, int col, float uu, float vv, float dt0, float hiC, float &oA, float &oB)
{
    float px = __fsub_rn((float)col, __fmul_rn(dt0, uu));
    float py = __fsub_rn((float)row, __fmul_rn(dt0, vv));
    if (px < 0.5f) px = 0.5f;
    if (px > hiC) px = hiC;
    if (py < 0.5f) py = 0.5f;
    if (py > hiC) py = hiC;
    const int c0 = (int)px, r0 = (int)py;
    const float wx1 = __fsub_rn(px, (float)c0), wx0 = __fsub_rn(1.0f, wx1);
    const float wy1 = __fsub_rn(py, (float)r0), wy0 = __fsub_rn(1.0f, wy1);
    const size_t G = (size_t)g.G;
    // a slab context stores rows [row_base, row_base + rows): a back-trace that leaves them (the caller's ghost rows were
    // too few -- slab.SlabSolver verifies the reach afterwards and raises) reads the nearest stored rows, never outside the array
    const int rs = min(max(r0, g.row_base), g.row_base + g.rows - 2);
    const size_t j = (size_t)(rs - g.row_base) * G + c0;
    {
        const float a00 = __ldg(srcA + j), a10 = __ldg(srcA + j + G), a01 = __ldg(srcA + j + 1), a11 = __ldg(srcA + j + G + 1);
        oA = __fadd_rn(__fmul_rn(wx0, __fadd_rn(__fmul_rn(wy0, a00), __fmul_rn(wy1, a10))),
                       __fmul_rn(wx1, __fadd_rn(__fmul_rn(wy0, a01), __fmul_rn(wy1, a11))));
    }
    if (NF == 2) {
        const float a00 = __ldg(srcB + j), a10 = __ldg(srcB + j + G), a01 = __ldg(srcB + j + 1), a11 = __ldg(srcB + j + G + 1);
        oB = __fadd_rn(__fmul_rn(wx0, __fadd_rn(__fmul_rn(wy0, a00), __fmul_rn(wy1, a10))),
                       __fmul_rn(wx1, __fadd_rn(__fmul_rn(wy0, a01), __fmul_rn(wy1, a11))));
    }
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
__device__ __forceinline__ void advect_cell_peer(const PeerView &sA, const PeerView &sB, const Geom &g, const PeerGeom &pg,
                                                 int row, int col, float uu, float vv, float dt0, float hiC, float &oA, float &oB)
{
    float px = __fsub_rn((float)col, __fmul_rn(dt0, uu));
    float py = __fsub_rn((float)row, __fmul_rn(dt0, vv));
    if (px < 0.5f) px = 0.5f;
    if (px > hiC) px = hiC;
    if (py < 0.5f) py = 0.5f;
    if (py > hiC) py = hiC;
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

}  // namespace

void preload_stage_kernels()
{
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, set_bnd_kernel); cudaFuncGetAttributes(&a, add_source_kernel);
    cudaFuncGetAttributes(&a, advect_kernel<1>); cudaFuncGetAttributes(&a, advect_kernel<2>);
    cudaFuncGetAttributes(&a, advect_lanes_kernel<1, false>); cudaFuncGetAttributes(&a, advect_lanes_kernel<2, false>);
    cudaFuncGetAttributes(&a, advect_lanes_kernel<1, true>); cudaFuncGetAttributes(&a, advect_lanes_kernel<2, true>);
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

cudaError_t launch_advect(const Geom &g, int b, float *d, const float *d0, const float *u, const float *v, float dt, cudaStream_t st)
{
    const int rows = interior_row_count(g);
    if (rows == 0) return cudaSuccess;
    const dim3 block(64, 4);
    const float dt0 = dt * (float)g.N;   // FluidSequential.c:111, rounded once in binary32
    if (g.G % 4 == 0) {
        const PeerView sA{d0, nullptr, nullptr};
        advect_lanes_kernel<1, false><<<lanes_grid(g, rows), dim3(32, 8), 0, st>>>(d, nullptr, sA, sA, u, v, g, PeerGeom(), dt0, b);
        return cudaGetLastError();
    }
    advect_kernel<1><<<cell_grid(g, block, rows), block, 0, st>>>(d, nullptr, d0, nullptr, u, v, g, dt0, b);
    return cudaGetLastError();
}

cudaError_t launch_advect_uv(const Geom &g, float *du, float *dv, const float *u0, const float *v0, float dt, cudaStream_t st)
{
    const int rows = interior_row_count(g);
    if (rows == 0) return cudaSuccess;
    const dim3 block(64, 4);
    const float dt0 = dt * (float)g.N;
    if (g.G % 4 == 0) {
        const PeerView sA{u0, nullptr, nullptr}, sB{v0, nullptr, nullptr};
        advect_lanes_kernel<2, false><<<lanes_grid(g, rows), dim3(32, 8), 0, st>>>(du, dv, sA, sB, u0, v0, g, PeerGeom(), dt0, 1);
        return cudaGetLastError();
    }
    advect_kernel<2><<<cell_grid(g, block, rows), block, 0, st>>>(du, dv, u0, v0, u0, v0, g, dt0, 1);
    return cudaGetLastError();
}

cudaError_t launch_advect_peer(const Geom &g, int b, float *d, const float *d0, const float *u, const float *v, float dt,
                               PeerSrc d0p, PeerGeom pg, cudaStream_t st)
{
    const int rows = interior_row_count(g);
    if (rows == 0) return cudaSuccess;
    if (g.G % 4 != 0) return cudaErrorInvalidValue;
    const float dt0 = dt * (float)g.N;   // FluidSequential.c:111
    const PeerView sA{d0, d0p.up, d0p.dn};
    advect_lanes_kernel<1, true><<<lanes_grid(g, rows), dim3(32, 8), 0, st>>>(d, nullptr, sA, sA, u, v, g, pg, dt0, b);
    return cudaGetLastError();
}

cudaError_t launch_advect_uv_peer(const Geom &g, float *du, float *dv, const float *u0, const float *v0, float dt,
                                  PeerSrc u0p, PeerSrc v0p, PeerGeom pg, cudaStream_t st)
{
    const int rows = interior_row_count(g);
    if (rows == 0) return cudaSuccess;
    if (g.G % 4 != 0) return cudaErrorInvalidValue;
    const float dt0 = dt * (float)g.N;
    const PeerView sA{u0, u0p.up, u0p.dn}, sB{v0, v0p.up, v0p.dn};
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
