// Shared device/host definitions for the stable-fluids kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
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

enum ArithMode {
    MODE_STRICT = 0,    // bit-identical; division by the constant beta via the exact FMA sequence below
    MODE_PRESSURE = 1,  // bit-identical; alpha == 1, beta == 4
    MODE_FAST = 2,      // opt-in: FMA contraction + reciprocal multiply (not bit-identical)
    MODE_IEEE = 3       // bit-identical; plain __fdiv_rn (used when beta has not been validated)
};

// ---- correctly rounded division by a constant ---------------------------------------------------
// a / b for a divisor that is constant over a launch, with y = RN(1/b) precomputed on the host.
//  fast path (binary32, 3 FMA-pipe instructions):  q0 = RN(a*y);  e = RN(b*q0 - a) (exact);
//     q1 = RN(q0 - e*y)  -- Markstein's FMA division step.  q1 == RN(a/b) whenever no intermediate
//     under/overflows, which the range guard ensures (|a| in [1e-30, 1e30], or a == +-0: writing
//     the residual as e = b*q0 - a makes zeros come out with the right sign).
//  slow path (numerators in the subnormal-result range or huge; also used for whole ticks once a
//     warp has met such a numerator): the same step in binary64 and one rounding to binary32.  The binary64 quotient is exact when a/b is representable and within
//     2^-52 otherwise, while a binary32 quotient is never closer than 2^-49 (relative) to a
//     rounding boundary, so the final rounding is the correct one.  Straight-line code: the decaying
//     diffusion front of a density field (values of 1e-30 .. 1e-45) costs a few DP instructions per
//     cell instead of the IEEE division's call-based special-case path, which made single warps
//     ~8x slower than their neighbours and stretched whole launches.
//  Neither path is trusted on paper: before a beta is used the library compares div_const with
//  __fdiv_rn for ALL 2^32 numerator bit patterns on the device (validate_division in sf_jacobi.cu,
//  ~5 ms, cached per beta); a beta that fails, or cannot be checked, runs MODE_IEEE.
#define SF_DIV_LO 1e-30f
#define SF_DIV_HI 1e30f
// a right-hand-side cell of at least this magnitude (2^-75) proves that no numerator it enters lies in (0, SF_DIV_LO):
// see row_flags in sf_jacobi.cu (2^-99 > 1e-30)
#define SF_RHS_LO 2.6469779601696886e-23f
struct DivConst {
    float b, y;      // divisor, RN32(1/b)
    float nz, pad;   // -0.0f as a RUN-TIME value (see mul2_exact)
    double bd, yd;   // (double)b, RN64(1/b)
};
inline DivConst make_div_const(float b)
{
    DivConst d;
    d.b = b; d.y = 1.0f / b;
    d.nz = -0.0f; d.pad = 0.0f;
    d.bd = (double)b; d.yd = 1.0 / (double)b;
    return d;
}
__device__ __forceinline__ float div_const_fast(float a, const DivConst &d)
{
    const float q0 = __fmul_rn(a, d.y);
    const float e = __fmaf_rn(d.b, q0, -a);
    return __fmaf_rn(-e, d.y, q0);
}
// lower side of the guard in two integer instructions: 2*bits - 1 drops the sign bit and wraps
// +-0 to 0xffffffff, so "zero or |a| >= LO" is one unsigned compare
__device__ __forceinline__ bool div_low_ok(float a)
{
    const unsigned t = 2u * __float_as_uint(a) - 1u;
    return t >= 2u * 0x0DA24260u - 1u;   // 0x0DA24260 = bits of 1e-30f
}
__device__ __forceinline__ bool div_high_ok(float a) { return fabsf(a) <= SF_DIV_HI; }
__device__ __forceinline__ bool div_in_range(float a) { return div_low_ok(a) && div_high_ok(a); }
// binary64 step: exact for EVERY binary32 numerator (validated exhaustively like the fast path), no
// range guard, straight-line.  Same residual form so +-0 keep their sign; +-inf passes through.
__device__ __forceinline__ float div_const_slow(float a, const DivConst &d)
{
    const double A = (double)a;
    const double q0 = __dmul_rn(A, d.yd);
    const double e = __fma_rn(d.bd, q0, -A);
    const float q = __double2float_rn(__fma_rn(-e, d.yd, q0));
    return (fabsf(a) <= 3.4028234664e38f) ? q : __fmul_rn(a, d.y);   // +-inf (and NaN) numerators
}
__device__ __forceinline__ float div_const(float a, const DivConst &d)
{
    const float q = div_const_fast(a, d);
    return div_in_range(a) ? q : div_const_slow(a, d);
}

// ---- packed binary32 pairs ------------------------------------------------------------------------
// sm_100 has two-wide binary32 instructions (PTX add/mul/fma.rn.f32x2 -> SASS FADD2 / FMUL2 / FFMA2) that
// work on an even-aligned register pair: each half is an ordinary IEEE round-to-nearest operation (no
// flush to zero), so results are bit-identical to the scalar instruction -- at half the issue slots, which
// is what bounds the temporally blocked Jacobi kernels.  ptxas folds the sign flips below into the
// instruction's operand modifiers and takes lane-invariant operands from uniform registers.
#ifndef SF_PACKED_F32
#define SF_PACKED_F32 1
#endif
__device__ __forceinline__ float2 add2_rn(float2 a, float2 b)
{
    float2 r;
    asm("{\n.reg .b64 ra, rb, rc;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\nadd.rn.f32x2 rc, ra, rb;\nmov.b64 {%0, %1}, rc;\n}\n"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 mul2_rn(float2 a, float2 b)
{
    float2 r;
    asm("{\n.reg .b64 ra, rb, rc;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\nmul.rn.f32x2 rc, ra, rb;\nmov.b64 {%0, %1}, rc;\n}\n"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 fma2_rn(float2 a, float2 b, float2 c)
{
    float2 r;
    asm("{\n.reg .b64 ra, rb, rc, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\nmov.b64 rc, {%6, %7};\n"
        "fma.rn.f32x2 rd, ra, rb, rc;\nmov.b64 {%0, %1}, rd;\n}\n"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
// A packed product that FEEDS A PACKED ADD must not be written with mul2_rn: ptxas 12.9 contracts
// mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 (one rounding instead of two) although both carry an explicit
// .rn, which forbids exactly that for the scalar forms -- and it does so with --fmad=false as well, and
// after first rewriting fma(a, b, -0.0) into a multiply.  RN(a*b + (-0)) == RN(a*b) bit for bit (signed
// zeros included), so the product is taken with an FFMA2 whose addend is -0.0f held in a register whose
// value ptxas cannot know (DivConst::nz, a kernel parameter): same instruction count, nothing to contract.
// tests/test_parity_gpu.py (alpha = 0.635 diffusion, every depth) fails by one ulp without this.
__device__ __forceinline__ float2 mul2_exact(float2 a, float2 b, float nz) { return fma2_rn(a, b, make_float2(nz, nz)); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 dup2(float a) { return make_float2(a, a); }
// div_const_fast on a pair: the same three roundings per half (residual as b*q0 - a, see above)
__device__ __forceinline__ float2 div_const_fast2(float2 a, const DivConst &d)
{
    const float2 y = dup2(d.y);
    const float2 q0 = mul2_rn(a, y);
    const float2 e = fma2_rn(dup2(d.b), q0, neg2(a));
    return fma2_rn(neg2(e), y, q0);
}

// One Jacobi cell update with the reference's operand order (FluidSequential.c:95-96):
//   ((left + right) + up) + down ;  x0 + alpha*sum ;  / beta.
// __fadd_rn/__fmul_rn are never contracted into FMAs by nvcc.
//   MODE_PRESSURE: alpha == 1, beta == 4 exactly: 1*sum == sum and /4 == *0.25f are exact
//                  identities in binary32 (also for subnormal results), so this is bit-identical
//                  to the general formula at a fraction of the instructions.
template <int MODE>
__device__ __forceinline__ float jacobi_numerator(float l, float r, float up, float dn, float b, float alpha)
{
    const float s = __fadd_rn(__fadd_rn(__fadd_rn(l, r), up), dn);
    if (MODE == MODE_PRESSURE) return __fadd_rn(b, s);
    if (MODE == MODE_FAST) return __fmaf_rn(alpha, s, b);
    return __fadd_rn(b, __fmul_rn(alpha, s));
}
// Two adjacent cells' numerators from h = left + right (formed by the caller with scalar adds: the operand
// pairs of that first addition straddle the register pairs): ((h + up) + down, then x0 + alpha * sum.
template <int MODE>
__device__ __forceinline__ float2 jacobi_numerator2(float2 h, float2 up, float2 dn, float2 b, float alpha, float nz)
{
    const float2 s = add2_rn(add2_rn(h, up), dn);
    if (MODE == MODE_PRESSURE) return add2_rn(b, s);
    if (MODE == MODE_FAST) return fma2_rn(dup2(alpha), s, b);
    return add2_rn(b, mul2_exact(dup2(alpha), s, nz));
}
template <int MODE>
__device__ __forceinline__ float jacobi_cell(float l, float r, float up, float dn, float b, float alpha,
                                             const DivConst &d)
{
    const float a = jacobi_numerator<MODE>(l, r, up, dn, b, alpha);
    if (MODE == MODE_PRESSURE) return __fmul_rn(a, 0.25f);
    if (MODE == MODE_FAST) return __fmul_rn(a, d.y);
    if (MODE == MODE_STRICT) return div_const(a, d);
    return __fdiv_rn(a, d.b);
}

__host__ __device__ __forceinline__ uint32_t hash100(uint64_t seed, uint64_t field, uint64_t cell)
{
    // 32-bit avalanche mixer (murmur3 finaliser) over (seed, field, global cell id), then a
    // multiply-shift reduction to 0..99; the test oracle uses the same formula for its synthetic ICs
    uint32_t h = (uint32_t)cell ^ (uint32_t)(cell >> 32) * 0x9E3779B1u;
    h ^= (uint32_t)seed * 0x85EBCA6Bu + (uint32_t)field * 0xC2B2AE35u + 0x27D4EB2Fu;
    h ^= h >> 16; h *= 0x85EBCA6Bu;
    h ^= h >> 13; h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return (uint32_t)(((uint64_t)h * 100ull) >> 32);
}

// ---- peer-memory slabs (sf_slab.cu) ------------------------------------------------------------
// Synchronisation words of one slab, in its own device memory; the neighbours write the inboxes
// through peer mappings.
//  * neighbour barrier on channel ch (nbr_barrier_kernel): bump epoch[ch], store it into both
//    neighbours' inbox[ch][side], spin until both of this slab's inbox[ch][*] have reached it;
//  * fused strip exchange (jacobi_stream_kernel): the boundary-strip warps of launch number k on side d
//    first wait until strip_inbox[d] >= k (the neighbour's k-th strip launch has stored its rows into
//    my ghost rows and is done reading its own), store their output rows into the neighbour's ghost
//    rows as they produce them, and the last strip warp to finish posts k+1 to the neighbour.
struct SlabFlags {
    unsigned long long inbox[2][2];       // [channel][0 = written by the up neighbour, 1 = by the down neighbour]
    unsigned long long epoch[2];          // barriers executed per channel (touched by this slab's barrier kernels only)
    unsigned long long strip_inbox[2];    // [side] strip launches the neighbour on that side has completed
    unsigned long long strip_seq[2];      // [side] strip launches this slab has completed
    unsigned long long strip_arrive[2];   // [side] strip warps finished, over all launches
    unsigned int error;                   // sticky SF_SLAB_ERR_* bits
    unsigned int pad;
};
// One side of a blocked Jacobi launch on a peer-memory slab: the `rows` owned rows next to the
// neighbour are produced by dedicated warps of the same kernel (scheduled first) that exchange them.
struct StripPort {
    int rows = 0;                              // strip height; 0 = no neighbour on this side
    float *xpeer = nullptr;                    // the neighbour's copy of the output field (peer-mapped)
    int peer_row_base = 0;                     // global row of the first row stored in xpeer
    unsigned long long *inbox = nullptr;       // local strip_inbox[side]
    unsigned long long *seq = nullptr;         // local strip_seq[side]
    unsigned long long *arrive = nullptr;      // local strip_arrive[side]
    unsigned long long *nbr_inbox = nullptr;   // neighbour's strip_inbox[other side] (peer-mapped)
};
// Everything the strip warps of one launch need, resident in DEVICE memory (one entry per output field
// and strip height, written when the neighbours are connected): the kernel receives a pointer, so the
// kernel parameter block of the single-GPU launches keeps its size.
struct StripArgs {
    int o_lo = 0, o_hi = 0;                    // all output rows of the launch (the slab's owned rows)
    StripPort port[2];                         // [0] side facing the up neighbour, [1] the down neighbour
    unsigned int *error = nullptr;             // device word for SF_SLAB_ERR_TIMEOUT
    unsigned long long timeout_ns = 0;
};

// ---- row-level work stealing between the warps of one Jacobi launch --------------------------------
// All work items of a launch cost the same EXCEPT where the exact division needs its guarded
// (binary64) ticks -- the decaying front of a density field -- and there a warp runs ~1.5-2x slower and
// would set the duration of the whole single-wave launch.  Every warp therefore publishes the row range
// it is working on; a warp that has finished samples a few slots, halves the largest remaining range
// with a compare-and-swap on its end and processes the upper half with a fresh pipeline (temporal
// blocking reads only level-0 rows, so a range can be split anywhere at the cost of 2T halo rows).
// Double coverage is harmless: both warps would write bit-identical values.
struct StealSlot {
    int pos;    // input row the owner has reached (published every few ticks)
    int tag;    // launch the slot belongs to (StealCtl::epoch + 1): stale slots are ignored
    // (end, band) form ONE aligned 64-bit word: a thief lowers `end` with a 64-bit compare-and-swap on the pair, so a
    // range republished with the same end but another band in between (chunk ends are aligned across bands) cannot be
    // mistaken for the range the thief sampled
    int end;    // one past the last output row the owner will produce; lowered by a thief
    int band;
};
static_assert(sizeof(StealSlot) == 16 && offsetof(StealSlot, end) == 8, "StealSlot: (end, band) is an aligned 64-bit word");
struct StealCtl {
    int epoch;      // launches completed
    int done;       // warps of the current launch that have finished
    int min_pct;    // a range is only halved while at least this percentage of a chunk remains (and >= 64 rows)
    int taken;      // statistics: ranges taken over
    StealSlot slots[1];   // [capacity]
};

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
    int staging;         // 0 = cp.async per lane (LDGSTS), 1 = bulk copies per warp row (cp.async.bulk / TMA unit)
    // Fused halo exchange (peer-memory slabs, sf_slab.cu): device-resident StripArgs of this launch and,
    // for the host-side launch geometry, the strip heights it holds (0 = no strip on that side)
    const StripArgs *strips = nullptr;
    int strip_rows[2] = {0, 0};
    // row-level work stealing (nullptr = off): device control block with room for steal_capacity items
    StealCtl *steal = nullptr;
    int steal_capacity = 0;
    // opt-in red-black Gauss-Seidel / SOR on the streaming pipeline: `sweeps` LEVELS (even: two per iteration)
    int rb = 0;
    float omega = 1.0f;
    // fused add_source (first launch of a solve in the step drivers): rhs = raw field, xin = source field = initial guess;
    // the launch forms raw + src_dt * xin on the fly and stores it to rhs_out (nullptr = off)
    float *rhs_out = nullptr;
    float src_dt = 0.0f;
    // unequal chunks for the three CTAs an SM holds (see chunk_range in sf_jacobi.cu): p0 * 1000 + p1 = rows of a chunk of the
    // first / second third of the items in percent of the mean chunk (e.g. 135106); 0 = equal chunks.  `ticket`: a device
    // word that is zero between launches (the context owns one).
    int strip_balance = 1;   // peer slabs: shorter chunks for the warps that computed a boundary strip first
    int wave_skew = 0;
    unsigned *ticket = nullptr;
};
// can launch_jacobi_stream fuse add_source into a launch of this depth and mode?  (the instantiations that exist)
inline bool jacobi_src_fusion_built(int sweeps, int mode) { return sweeps >= 5 && sweeps <= 7 && (mode == MODE_STRICT || mode == MODE_IEEE); }
cudaError_t launch_jacobi_stream(const Geom &g, const JacobiLaunch &L, int sm_count, cudaStream_t st);
cudaError_t launch_jacobi_generic(const Geom &g, const JacobiLaunch &L, cudaStream_t st);
bool jacobi_stream_supported(const Geom &g);
// opt-in red-black Gauss-Seidel / SOR (sf_solvers.cu): one in-place half-sweep over the cells of one colour
cudaError_t launch_rbgs_half_sweep(const Geom &g, float *x, const float *rhs, int colour, int mode, float alpha, float beta,
                                   float omega, cudaStream_t st);
// Exhaustive device check of div_const against __fdiv_rn for this beta (cached per process).
// Returns true when MODE_STRICT may be used.  Synchronises `st`; must not be called while `st`
// is being captured (pass allow_run = false to only consult the cache).
bool division_validated(float beta, bool allow_run, cudaStream_t st);

// the two neighbours' copies of one field plus their geometry: row r of the up neighbour's array is
// up + (r - up_row_base) * G, valid for r in [up_lo, own_lo); down likewise for r in [own_hi, dn_hi)
struct PeerSrc {
    const float *up = nullptr, *dn = nullptr;
};
struct PeerGeom {
    int up_row_base = 0, up_lo = 0, dn_row_base = 0, dn_hi = 0;
    unsigned int *error = nullptr;    // device word that receives SF_SLAB_ERR_REACH
};
cudaError_t launch_nbr_barrier(SlabFlags *me, SlabFlags *up, SlabFlags *dn, int channel, unsigned long long timeout_ns,
                               cudaStream_t st);
struct PushSegment {
    const float *src;
    float *dst;
    size_t count;   // floats, multiple of 4, both pointers 16-byte aligned
};
cudaError_t launch_push_rows(const PushSegment *segs, int nsegs, cudaStream_t st);
// advect whose gather may leave the slab: rows outside [own_lo, own_hi) are read from the neighbours
// `tile` (everywhere below): 0 = gather from global memory, > 0 = source tile staged by the TMA unit (SF_OPT_ADVECT_TILE);
// `tile_stats`: two device words that count the tiles served by the TMA box / by the gather fallback
cudaError_t launch_advect_peer(const Geom &g, int b, float *d, const float *d0, const float *u, const float *v, float dt,
                               PeerSrc d0p, PeerGeom pg, int tile, unsigned int *tile_stats, cudaStream_t st);
cudaError_t launch_advect_uv_peer(const Geom &g, float *du, float *dv, const float *u0, const float *v0, float dt,
                                  PeerSrc u0p, PeerSrc v0p, PeerGeom pg, int tile, unsigned int *tile_stats, cudaStream_t st);

// force-load every kernel of the library (see preload_jacobi_kernels)
void preload_jacobi_kernels();
void preload_stage_kernels();

cudaError_t launch_set_bnd(const Geom &g, int b, float *x, cudaStream_t st);
cudaError_t launch_add_source(const Geom &g, int nfields, float *const *x, const float *const *s, float dt,
                              cudaStream_t st);
cudaError_t launch_advect(const Geom &g, int b, float *d, const float *d0, const float *u, const float *v, float dt, int tile,
                          unsigned int *tile_stats, cudaStream_t st);
// both velocity components in one pass: d_u <- advect(b=1, u0), d_v <- advect(b=2, v0) by (u0, v0)
cudaError_t launch_advect_uv(const Geom &g, float *du, float *dv, const float *u0, const float *v0, float dt, int tile,
                             unsigned int *tile_stats, cudaStream_t st);
cudaError_t launch_divergence(const Geom &g, const float *u, const float *v, float *p, float *div, int write_p,
                              cudaStream_t st);
cudaError_t launch_last_project(const Geom &g, float *u, float *v, const float *p, cudaStream_t st);
cudaError_t launch_init(const Geom &g, uint64_t seed, float *dens, float *dens_prev, float *u, float *u_prev, float *v,
                        float *v_prev, cudaStream_t st);
cudaError_t launch_max_abs(const Geom &g, const float *x, float *dev_out, bool zero_first, cudaStream_t st);
cudaError_t launch_residual(const Geom &g, const float *x, const float *x0, float alpha, float beta, double *dev_out,
                            cudaStream_t st);

}  // namespace sf
