// Jacobi lin_solve kernels (replaces the reference's diffuse(): FluidSequential.c:85-104, and the
// one-sweep-per-launch diffuseOnGPU + host loop, naivePar/FluidParallelBlockPerElement-Naive.cu:121-144,
// 261-264).  NOT a port: the reference goes to DRAM once per sweep; here T sweeps are fused per
// launch (temporal blocking) so each field crosses HBM once per T sweeps.
//
// jacobi_stream_kernel<T, MODE>  -- the product path (grid width G % 4 == 0)
//   * Each WARP owns a band of 128 columns (one float4 per lane) and streams down the rows of a
//     row chunk.  Sweep level t+1 of row a is computed as soon as level t of row a+1 exists, so a
//     warp carries a T-deep register pipeline: per level two previous rows (2 x float4 per lane).
//     Left/right neighbours come from the adjacent lanes by warp shuffle; warps never synchronise
//     with each other (no __syncthreads, no shared-memory exchange): the band carries an 8-column
//     halo on each side that absorbs the T <= 8 columns invalidated by the missing neighbours.
//   * Rows of x (level 0) and of the right-hand side x0 are staged global -> shared with
//     cp.async (LDGSTS, 16 B per lane, L2-only) into per-warp rings several rows ahead, so DRAM
//     latency is hidden without holding prefetch registers; the rhs ring is read once per level.
//   * set_bnd is fused: a wall cell at level t+1 is +-(the adjacent interior cell at level t+1),
//     applied in registers (the wall columns 0 and N+1 share a float4 with columns 1 / N because
//     G % 4 == 0); wall rows are patched into the next level's window when row 1 / row N are
//     produced.  Corners are never read by the 5-point stencil and are written with the last level.
//     All of that lives in the WALLS=true instantiation of the tick, which only the two edge bands
//     and the <= T+1 ticks next to the top/bottom walls execute; everything else runs the
//     WALLS=false tick: per level 2 shuffles + 1 LDS.128 + the cell arithmetic.
//   * Arithmetic per cell is jacobi_cell<MODE> (sf_common.cuh): same operand order as the reference.
//
// jacobi_generic_kernel<MODE> -- one sweep per launch, one thread per cell, any G (fallback for
//   widths that are not a multiple of 4, e.g. the literal N=128 -> G=130).
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>

#include "sf_common.cuh"

namespace sf {

namespace {

constexpr int BAND_W = 128;   // columns per warp (32 lanes x float4)
constexpr int HALO_X = 8;     // band halo columns on each side (>= max T, multiple of 4)
constexpr int VALID_W = BAND_W - 2 * HALO_X;  // 112 output columns per band
constexpr int STEAL_MIN_ROWS = 16;            // work stealing: smallest remaining range worth halving, whatever the option says
#ifndef SF_WPC
#define SF_WPC 4
#define SF_RING_X 8
#define SF_RING_R 16
#define SF_PREFETCH 5
#endif
#ifndef SF_OFF32
#define SF_OFF32 1        // 32-bit cell offsets inside a field (see stream_rows)
#endif
// Adopted after the A/B of round 2 (profiles/r02/s1_86a0b52_ab_variants.txt, G = 8192, K = 40, one B200):
//  * consecutive wall-free groups of the modes WITHOUT a range check (pressure, fast, IEEE) run in an inner loop of their own,
//    and the pressure kernel (T >= 6) is register-bounded for 3 CTAs per SM: with one loop for both paths ptxas sank the
//    general tick's window rotation (8T register moves) into the common latch, 56 MOVs per group on the fast path; the
//    inner loop needs 142 registers, hence 12 instead of 16 warps per SM.  Pressure solve 0.916 -> 0.872 ms.
//  * rows inside a guarded span (the decaying front of a density field) run in wall-free groups of three ticks with the
//    binary64 division instead of one general tick per row.  Density solve 1.92 -> 1.67 ms, step 8.58 -> 8.33 ms.
// Measured and removed: an inner loop for the strict groups too (no change: ptxas already gives them their own loop), a
// separate group instantiation for bands without a wall column (no change at best, spills at worst).
#if SF_OFF32
typedef unsigned cell_t;
#else
typedef size_t cell_t;
#endif
constexpr int WPC = SF_WPC;            // warps per CTA (adjacent bands, same row chunk)
constexpr int RING_X = SF_RING_X;      // x-row ring slots per warp (power of two, >= PREFETCH + 3)
constexpr int RING_R = SF_RING_R;      // rhs-row ring slots per warp (power of two, >= T + PREFETCH + 3)
constexpr int PREFETCH = SF_PREFETCH;  // rows in flight ahead of the row being consumed
// CTAs per SM the register allocation is bounded for.  The branch-free strict tick keeps more values
// in flight: at T >= 6 it needs ~150 registers, so it runs 3 CTAs (12 warps) per SM instead of 4 --
// measured faster than spilling at 128 registers (1.78 vs 2.18 ms per 40-sweep solve at G=8192).
template <int T, int MODE>
constexpr int min_ctas() { return ((MODE == MODE_STRICT || MODE == MODE_IEEE || MODE == MODE_PRESSURE) && T >= 6) ? 3 : 4; }

struct StreamArgs {
    const float *xin, *rhs;
    float *xout;
    int G, N, row_base;
    int a_lo, a_hi;      // interior output rows [a_lo, a_hi), subset of [1, N+1)
    int write_top, write_bot;
    int chunk_rows, nchunks, nbands;
    int zero_guess;
    float alpha, sx, sy;
    float hi_in;         // level-0 magnitudes up to this keep every numerator of the launch <= SF_DIV_HI
    DivConst div;        // beta and its reciprocals
    const StripArgs *strips;   // peer-memory slabs: the fused exchange of the top / bottom strip (device memory)
    StealCtl *steal;           // row-level work stealing (VAR 3 / 4), device memory
    // fused add_source (VAR 6 / 7, on peer slabs 8 / 9; first launch of a solve inside the step drivers): rhs = raw + src_dt * xin is formed
    // as the rows land and stored to rhs_out for the later launches of the solve; `rhs` points at the raw field
    float *rhs_out;
    float src_dt;
    // age-ordered work items with unequal chunks (see chunk_range); ticket == nullptr: items by blockIdx, equal chunks
    unsigned *ticket;    // four device words, zero between launches: per-class item counters + arrivals (see the kernel)
    unsigned skew;       // rows of a chunk of the 1st third of the items | rows of a chunk of the 2nd third << 16 (0 = equal chunks)
    unsigned skew_cpw;   // chunks per third
};
// Output rows of chunk `chunk`.  All warps of a launch are resident at once (one wave of CTAs, three per SM at T >= 6), but
// they do not finish together: a warp scheduler favours the warp in its lowest hardware slot.  With equal chunks of 342 rows
// (G = 8192, strict T = 7) the warps of the three resident CTAs of an SM ran at about 2.4 / 1.9 / 1.2 rows per us while all
// three were active and finished after 144 / 173 / 216 us; a lone warp cannot use a scheduler's issue slots (~2.6 rows/us
// against ~5.4 for three), so the last third of the launch runs at a fraction of the machine (tools/warp_times.py).
// So (i) every warp draws its work item according to the hardware slot it runs in (see the kernel) and (ii) the chunks of the
// first / second / last third of the items get rows in proportion to the rates of the three slots classes; the total is
// unchanged, and so is every bit of the result (temporal blocking does not depend on where the chunks are cut).
// Peer-slab launches (SHORT = true; they never use the slot classes above): the warps of the first one or two chunks have
// computed a boundary strip before they get here -- (strip rows + 2T) general ticks, about 40 rows' worth -- so those chunks
// are that much shorter than the others (skew = rows of a short chunk | number of short chunks << 16, skew_cpw = 0,
// chunk_rows = rows of the others) and every warp of the launch finishes at about the same time.
template <bool SHORT = false>
__device__ __forceinline__ void chunk_range(const StreamArgs &A, int chunk, int &a_lo, int &a_hi)
{
    if constexpr (SHORT) {
        if (A.skew != 0u && A.skew_cpw == 0u) {
            const int c0 = (int)(A.skew & 0xffffu), ns = (int)(A.skew >> 16);
            const int rows = chunk < ns ? c0 : A.chunk_rows;
            a_lo = A.a_lo + (chunk < ns ? chunk * c0 : ns * c0 + (chunk - ns) * A.chunk_rows);
            a_hi = min(a_lo + rows, A.a_hi);
            a_lo = min(a_lo, A.a_hi);
            return;
        }
    }
    if (A.skew == 0u) {
        a_lo = A.a_lo + chunk * A.chunk_rows;
        a_hi = min(a_lo + A.chunk_rows, A.a_hi);
        return;
    }
    const int r0 = (int)(A.skew & 0xffffu), r1 = (int)(A.skew >> 16), r2 = 3 * A.chunk_rows - r0 - r1, cpw = (int)A.skew_cpw;
    const int w = chunk / cpw, j = chunk - w * cpw;
    const int rows = w == 0 ? r0 : (w == 1 ? r1 : r2);
    a_lo = A.a_lo + cpw * (w == 0 ? 0 : (w == 1 ? r0 : r0 + r1)) + j * rows;
    a_hi = min(a_lo + rows, A.a_hi);
    a_lo = min(a_lo, A.a_hi);
}

// GPU-scope relaxed accesses for the stealing words (plain `volatile` would be system-scope strong accesses)
__device__ __forceinline__ int ld_relaxed_gpu(const int *p)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_acquire_gpu(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu(int *p, int v)
{
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

// the hardware warp slot this warp runs in (%warpid: a hint -- it may change under preemption; used for load balance only)
__device__ __forceinline__ unsigned hw_warp_slot()
{
    unsigned wid;
    asm volatile("mov.u32 %0, %%warpid;\n" : "=r"(wid));
    return wid;
}

__device__ __forceinline__ unsigned hw_sm_id()
{
    unsigned id;
    asm volatile("mov.u32 %0, %%smid;\n" : "=r"(id));
    return id;
}

// ---- fused strip exchange (peer-memory slabs) -----------------------------------------------------
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;\n" : "=l"(t));
    return t;
}
// A strip warp of launch number k (= strip launches this side has completed) may read its ghost rows
// and overwrite the neighbour's once the neighbour has completed k strip launches of its own.
__device__ __forceinline__ void strip_wait(const unsigned long long *seq, const unsigned long long *inbox, unsigned int *error,
                                        unsigned long long timeout_ns, int lane)
{
    if (lane == 0) {
        const unsigned long long k = *(const volatile unsigned long long *)seq;
        const unsigned long long t0 = globaltimer_ns();
        unsigned spins = 0;
        while (ld_acquire_sys_u64(inbox) < k) {
            if ((++spins & 255u) == 0) {
                if (*(volatile unsigned int *)error & 1u) break;
                if (globaltimer_ns() - t0 > timeout_ns) { atomicOr(error, 1u); break; }   // SF_SLAB_ERR_TIMEOUT
            }
            __nanosleep(32);
        }
    }
    __syncwarp();
}
// Every strip warp publishes its peer stores; the last one of the launch posts the new count to the neighbour.
__device__ __forceinline__ void strip_post(unsigned long long *arrive, unsigned long long *seq, unsigned long long *nbr_inbox,
                                        int warps_per_launch, int lane)
{
    __threadfence_system();
    __syncwarp();
    if (lane == 0) {
        const unsigned long long n = atomicAdd(arrive, 1ull) + 1ull;
        if (n % (unsigned long long)warps_per_launch == 0) {
            const unsigned long long k1 = n / (unsigned long long)warps_per_launch;
            *(volatile unsigned long long *)seq = k1;
            __threadfence_system();
            st_release_sys_u64(nbr_inbox, k1);
        }
    }
}

__device__ __forceinline__ void cp_async16(float4 *smem_dst, const float *gmem_src, int src_bytes)
{
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N_) : "memory"); }

// ---- TMA-style staging: 1-D bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
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
__device__ __forceinline__ void bulk_copy_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

__device__ __forceinline__ float4 scale4(float4 v, float s)
{
    return make_float4(__fmul_rn(v.x, s), __fmul_rn(v.y, s), __fmul_rn(v.z, s), __fmul_rn(v.w, s));
}

// Four adjacent cells of one row at one level.
// MODE_STRICT, GUARDED = true : the binary64 division step for every cell (general tick): exact for
//                               all numerators, no range test at all.
// MODE_STRICT, GUARDED = false: optimistic -- always the 3-instruction division, and only the LOW end
//                               of the range is tested (2 integer instructions per cell) and AND-ed
//                               into `ok`; no branch sits between the sweep levels.  The caller votes
//                               once per tick and redoes the tick GUARDED if any lane failed (the high
//                               end is proved per fetched row, see row_is_big).
//                CHECK = false : no test at all -- the caller has proved from the right-hand-side rows that no numerator
//                               of this tick can lie in (0, SF_DIV_LO) (see row_flags in stream_rows).
template <int MODE, bool GUARDED, bool CHECK = true>
__device__ __forceinline__ float4 jacobi4(float lft, const float4 &mid, float rgt, const float4 &up, const float4 &dn,
                                          const float4 &r, float alpha, const DivConst &d, bool &ok)
{
    float4 o;
#if SF_PACKED_F32
    {
        // l + r with scalar adds, everything after that two cells per instruction (FADD2 / FMUL2 / FFMA2)
        const float2 h01 = make_float2(__fadd_rn(lft, mid.y), __fadd_rn(mid.x, mid.z));
        const float2 h23 = make_float2(__fadd_rn(mid.y, mid.w), __fadd_rn(mid.z, rgt));
        const float2 a01 = jacobi_numerator2<MODE>(h01, make_float2(up.x, up.y), make_float2(dn.x, dn.y), make_float2(r.x, r.y), alpha, d.nz);
        const float2 a23 = jacobi_numerator2<MODE>(h23, make_float2(up.z, up.w), make_float2(dn.z, dn.w), make_float2(r.z, r.w), alpha, d.nz);
        if (MODE == MODE_PRESSURE) {
            const float2 o01 = mul2_rn(a01, dup2(0.25f)), o23 = mul2_rn(a23, dup2(0.25f));
            return make_float4(o01.x, o01.y, o23.x, o23.y);
        }
        if (MODE == MODE_FAST) {
            const float2 o01 = mul2_rn(a01, dup2(d.y)), o23 = mul2_rn(a23, dup2(d.y));
            return make_float4(o01.x, o01.y, o23.x, o23.y);
        }
        if (MODE == MODE_STRICT && !GUARDED) {
            const float2 o01 = div_const_fast2(a01, d), o23 = div_const_fast2(a23, d);
            if (CHECK) ok = ok & div_low_ok(a01.x) & div_low_ok(a01.y) & div_low_ok(a23.x) & div_low_ok(a23.y);
            return make_float4(o01.x, o01.y, o23.x, o23.y);
        }
        if (MODE == MODE_STRICT) {
            return make_float4(div_const_slow(a01.x, d), div_const_slow(a01.y, d), div_const_slow(a23.x, d), div_const_slow(a23.y, d));
        }
        return make_float4(__fdiv_rn(a01.x, d.b), __fdiv_rn(a01.y, d.b), __fdiv_rn(a23.x, d.b), __fdiv_rn(a23.y, d.b));
    }
#endif
    if (MODE == MODE_STRICT) {
        const float a0 = jacobi_numerator<MODE>(lft, mid.y, up.x, dn.x, r.x, alpha);
        const float a1 = jacobi_numerator<MODE>(mid.x, mid.z, up.y, dn.y, r.y, alpha);
        const float a2 = jacobi_numerator<MODE>(mid.y, mid.w, up.z, dn.z, r.z, alpha);
        const float a3 = jacobi_numerator<MODE>(mid.z, rgt, up.w, dn.w, r.w, alpha);
        if (!GUARDED) {
            o.x = div_const_fast(a0, d);
            o.y = div_const_fast(a1, d);
            o.z = div_const_fast(a2, d);
            o.w = div_const_fast(a3, d);
        }
        if (GUARDED) {
            // binary64 step for all four cells: exact for every numerator, no branches, and the four
            // chains overlap (a lone warp in this mode runs about as fast as a warp on the fast tick
            // because the conversion/FP64 pipes are otherwise idle)
            o.x = div_const_slow(a0, d);
            o.y = div_const_slow(a1, d);
            o.z = div_const_slow(a2, d);
            o.w = div_const_slow(a3, d);
        } else {
            ok = ok & div_low_ok(a0) & div_low_ok(a1) & div_low_ok(a2) & div_low_ok(a3);
        }
    } else {
        o.x = jacobi_cell<MODE>(lft, mid.y, up.x, dn.x, r.x, alpha, d);
        o.y = jacobi_cell<MODE>(mid.x, mid.z, up.y, dn.y, r.y, alpha, d);
        o.z = jacobi_cell<MODE>(mid.y, mid.w, up.z, dn.z, r.z, alpha, d);
        o.w = jacobi_cell<MODE>(mid.z, rgt, up.w, dn.w, r.w, alpha, d);
    }
    return o;
}

// One pipeline tick: row `s` of level 0 enters; for t = 0..T-1 level t+1 of row s-t-1 is produced.
// Each level keeps a window of three row slots W[t][0..2] that rotate with period 3: at phase PH
// slot PH holds row a-1 (up), slot PH+1 row a (mid) and slot PH+2 receives row a+1 (dn), written by
// the level below in this same tick.  PH is a compile-time constant, so after three ticks every
// row is back in the register it started in and the hot loop contains no register moves.
// WALLS adds the fused set_bnd handling.  Returns level T of row s-T in `out`.
// RB = true (opt-in red-black Gauss-Seidel / SOR, SF_OPT_RBGS_BLOCKED; design checked in tools/models/rbgs_blocked_model.py):
// one red-black ITERATION is two LEVELS -- level t+1 with t even is the state after the red half-sweep ((row + col) even),
// with t odd after the black one -- a level updates the cells of its colour and copies the others through, and set_bnd acts
// on black levels only (wall columns and wall rows copy through on red levels).  A lane's first column is a multiple of 4,
// so which two of its four cells a level updates is warp-uniform.  omega travels in A.div.pad (1.0f = plain Gauss-Seidel).
// GUARD = true: the wall-free tick with the binary64 division for every cell (groups inside a guarded span).
// CHECK = false: the strict wall-free tick without the per-cell low-range test (see jacobi4).
template <int T, int MODE, int PH, bool WALLS, bool RB = false, bool GUARD = false, bool CHECK = true>
__device__ __forceinline__ bool pipeline_tick(const StreamArgs &A, int s, const float4 &row_in, float4 (&W)[T][3],
                                              const float4 *rring, bool ownsL, bool ownsR, float4 &out)
{
    bool ok = true;
    constexpr int UP = PH % 3, MID = (PH + 1) % 3, DN = (PH + 2) % 3;
    W[0][DN] = row_in;
#pragma unroll
    for (int t = 0; t < T; ++t) {
        const int a = s - t - 1;               // row produced at level t+1
        const float4 up = W[t][UP], mid = W[t][MID], dn = W[t][DN];
        const float4 r = rring[(a & (RING_R - 1)) * 32];
        const float lft = __shfl_up_sync(0xffffffffu, mid.w, 1);
        const float rgt = __shfl_down_sync(0xffffffffu, mid.x, 1);
        float4 o = jacobi4<MODE, WALLS || GUARD, CHECK>(lft, mid, rgt, up, dn, r, A.alpha, A.div, ok);
        if constexpr (RB) {
            const bool black = (t & 1) != 0;                 // compile-time: t is an unrolled index
            const float om = A.div.pad;
            if (om != 1.0f) {                                // SOR: x + omega*(gs - x), three roundings (warp-uniform branch)
                o.x = __fadd_rn(mid.x, __fmul_rn(om, __fsub_rn(o.x, mid.x)));
                o.y = __fadd_rn(mid.y, __fmul_rn(om, __fsub_rn(o.y, mid.y)));
                o.z = __fadd_rn(mid.z, __fmul_rn(om, __fsub_rn(o.z, mid.z)));
                o.w = __fadd_rn(mid.w, __fmul_rn(om, __fsub_rn(o.w, mid.w)));
            }
            // cell (a, c + k) belongs to this level iff ((a + k) & 1) == colour, colour = black ? 1 : 0 (c is even)
            const bool even_upd = (((a & 1) != 0) == black);
            o.x = even_upd ? o.x : mid.x;
            o.z = even_upd ? o.z : mid.z;
            o.y = even_upd ? mid.y : o.y;
            o.w = even_upd ? mid.w : o.w;
            if (black) {
                if (ownsL) o.x = __fmul_rn(A.sx, o.y);
                if (ownsR) o.w = __fmul_rn(A.sx, o.z);
            } else {
                if (ownsL) o.x = mid.x;
                if (ownsR) o.w = mid.w;
            }
            if (WALLS) {
                if (t + 1 < T) {
                    if (a == A.N + 1) o = black ? scale4(W[t + 1][MID], A.sy) : mid;     // row N+1 of level t+1
                    if (a == 1) W[t + 1][MID] = black ? scale4(o, A.sy) : up;            // row 0 of level t+1
                }
            }
        } else {
        // wall columns: x[row][0] = sx * x[row][1], x[row][N+1] = sx * x[row][N]  (two predicated
        // multiplies; only the lanes holding columns 0 / N+1 of the two edge bands execute them)
        if (ownsL) o.x = __fmul_rn(A.sx, o.y);
        if (ownsR) o.w = __fmul_rn(A.sx, o.z);
        if (WALLS) {
            if (t + 1 < T) {
                // wall rows of level t+1 live in the NEXT level's window: its MID slot is row a-1
                if (a == A.N + 1) o = scale4(W[t + 1][MID], A.sy);     // row N+1 = sy * row N
                if (a == 1) W[t + 1][MID] = scale4(o, A.sy);           // row 0   = sy * row 1
            }
        }
        }
        if (t + 1 < T) W[t + 1][DN] = o; else out = o;
    }
    return ok;
}

// TMA = false: rows are staged with per-lane cp.async (LDGSTS.128).  TMA = true: one elected lane
// issues a 1-D bulk copy (cp.async.bulk -> UBLKCP) of the warp's whole 512-byte row piece per field,
// completing on a per-slot mbarrier that all lanes wait on (SF_OPT_STAGING; measured in DESIGN.md).
// The streaming pipeline of one warp over output rows [a_lo, a_hi) of band `band`.
// STRIP = true (peer-memory slabs): general ticks only, and every produced row is also stored at
// peer + row * pitch, the neighbour GPU's copy of the field (pre-offset so that global row numbers index it).
// STEAL = true: the warp's slot (A.steal->slots[global warp]) holds its published range; the end may be
// lowered by another warp at any time.
// SRC = true: add_source fused into the first launch of a lin_solve (FluidSequential.c:78-82 + :85-104).  A.rhs is the RAW
// field x, A.xin the source field s -- which is also the solve's initial guess in dens_step / vel_step (the local SWAP at
// :181 / :201) -- and every level-0 row becomes rhs = x + dt * s in the warp's own ring slot as soon as it has landed
// (each lane reads and writes only its own 16 bytes of a slot: no synchronisation).  The rows of the warp's output range
// are also stored to A.rhs_out, a separate field: the later launches of the solve read their right-hand side there, and
// x itself stays untouched (in place, a neighbouring warp could read a halo row after it was updated and add the source
// twice).  One pass over x and s less per solve: 8 B per cell.
template <int T, int MODE, bool TMA, bool STRIP, bool STEAL, bool RB = false, bool SRC = false>
__device__ __forceinline__ void stream_rows(const StreamArgs &A, float4 *ring, const int lane, const int warp, const int band,
                                            const int a_lo, const int a_hi, float *peer)
{
    const int c = band * VALID_W - HALO_X + 4 * lane;     // first of this lane's 4 columns
    const bool indom = (c >= 0) && (c + 4 <= A.G);
    const bool ownsL = (c == 0), ownsR = (c + 4 == A.G);
    const bool st_ok = indom && lane >= HALO_X / 4 && lane < 32 - HALO_X / 4;
    const int cc = indom ? c : 0;
    const int nbytes = indom ? 16 : 0;

    float4 *xring = ring + (size_t)warp * (RING_X + RING_R) * 32 + lane;
    float4 *rring = xring + RING_X * 32;
    // TMA staging: per-warp mbarriers behind the rings, one per x-ring slot (row r uses slot r & 7)
    uint64_t *bars = reinterpret_cast<uint64_t *>(ring + (size_t)WPC * (RING_X + RING_R) * 32) + warp * RING_X;
    const int col0 = band * VALID_W - HALO_X;                 // band's first column (may be < 0)
    const int tma_lo = max(col0, 0), tma_hi = min(col0 + BAND_W, A.G);
    const unsigned tma_bytes = (unsigned)(tma_hi - tma_lo) * 4u;
    if (TMA) {
        if (!indom) {   // lanes outside the grid are never written by the bulk copies: zero them once
            for (int k = 0; k < RING_X; ++k) xring[k * 32] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int k = 0; k < RING_R; ++k) rring[k * 32] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (lane == 0) {
            for (int k = 0; k < RING_X; ++k) mbar_init(bars + k, 1);
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        }
        __syncwarp();
    }

    // `first`: the first output row that has not been produced yet.  It starts at a_lo and moves only when the
    // pipeline is restarted (GROUP_VOTE below); the pipeline enters T rows above it.
    int first = a_lo;
    int s_lo = max(first - T, 0);
    const int s_hi = a_hi - 1 + T;                 // inclusive
    const int load_hi = min(s_hi, A.G - 1);
    const size_t pitch = (size_t)A.G;
    // Cell offsets inside a field are taken in unsigned 32-bit arithmetic (jacobi_stream_supported keeps
    // rows * G below 2^32): one 32-bit multiply per row instead of a sign-extended 64-bit one, and the pointer
    // is one widening multiply-add from the field base.
    const cell_t Gu = (cell_t)A.G, ccu = (cell_t)cc;
    auto cell = [&](int row) -> cell_t { return (cell_t)(row - A.row_base) * Gu + ccu; };
    const bool zero_guess = A.zero_guess != 0;

    auto issue = [&](int row) {
        if (TMA) {
            if (row <= load_hi && lane == 0) {
                const size_t off = (size_t)(row - A.row_base) * pitch + tma_lo;
                uint64_t *bar = bars + (row & (RING_X - 1));
                float4 *xw = xring - lane, *rw = rring - lane;     // warp-level ring bases
                mbar_expect_tx(bar, zero_guess ? tma_bytes : 2u * tma_bytes);
                if (!zero_guess)
                    bulk_copy_g2s(reinterpret_cast<float *>(xw + (row & (RING_X - 1)) * 32) + (tma_lo - col0), A.xin + off, tma_bytes, bar);
                bulk_copy_g2s(reinterpret_cast<float *>(rw + (row & (RING_R - 1)) * 32) + (tma_lo - col0), A.rhs + off, tma_bytes, bar);
            }
            return;
        }
        if (row <= load_hi) {
            const cell_t off = cell(row);
            if (!zero_guess) cp_async16(xring + (row & (RING_X - 1)) * 32, A.xin + off, nbytes);
            cp_async16(rring + (row & (RING_R - 1)) * 32, A.rhs + off, nbytes);
        }
        cp_async_commit();
    };
    // three consecutive rows (one fast group): one address computation when all of them exist
    auto issue3 = [&](int row) {
        if (!TMA && row + 2 <= load_hi) {
            const cell_t off = cell(row);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (!zero_guess) cp_async16(xring + ((row + k) & (RING_X - 1)) * 32, A.xin + (off + (cell_t)k * Gu), nbytes);
                cp_async16(rring + ((row + k) & (RING_R - 1)) * 32, A.rhs + (off + (cell_t)k * Gu), nbytes);
                cp_async_commit();
            }
            return;
        }
        issue(row); issue(row + 1); issue(row + 2);
    };
    // rows [first, first + n) have been issued; block until they have landed
    auto landed = [&](int first, int n) {
        if (TMA) {
            for (int k = 0; k < n; ++k) {
                const int row = first + k;
                if (row <= load_hi) mbar_wait(bars + (row & (RING_X - 1)), (unsigned)((row - s_lo) >> 3) & 1u);
            }
        } else {
            cp_async_wait<PREFETCH>();
        }
    };
    // MODE_STRICT only: the wall-free tick guards just the LOW end of the exact division's range per
    // cell.  The high end follows from a maximum principle: |numerator| <= (1 + 4|alpha|) * max|input|,
    // so rows whose magnitudes stay below A.hi_in can never produce a numerator above SF_DIV_HI.
    // Every row is checked when it is fetched; one outlier switches the warp to the fully guarded
    // tick for the next rows (up to the next multiple of 64).
    // (both lambdas below are only used inside a fast group, whose rows s..s+2 <= fast_hi <= load_hi)
    auto xrow_in = [&](int row) -> float4 {
        return zero_guess ? make_float4(0.f, 0.f, 0.f, 0.f) : xring[(row & (RING_X - 1)) * 32];
    };
    // The LOW end needs no per-cell test where the right-hand side is not tiny.  A numerator is a = RN(x0 + w), w = RN(alpha *
    // sum), with the SAME x0 at every level.  If |x0| >= 2^-75: either |w| <= |x0|/2 or |w| >= 2|x0|, and then |a| >= |x0|/2;
    // or w is within a factor 2 of x0, both are integer multiples of 2^(floor(log2|x0|) - 24) >= 2^-99, and so is their exact
    // sum -- it is 0 or at least 2^-99 in magnitude, and rounding keeps that.  Either way a == 0 or |a| >= 2^-99 > SF_DIV_LO:
    // inside the range the exhaustive validation of div_const_fast covers.  So a row of the right-hand side whose cells all
    // have |x0| >= SF_RHS_LO proves every numerator that will ever use it, and a group all of whose right-hand-side rows are
    // proven (rows s-T .. s+1: `unproven_until`) runs the tick WITHOUT the test: ~125 of ~875 instructions per group less.
    // Rows with a zero or tiny right-hand side (outside the support of a density field, its decaying front) keep the tested
    // tick.  bit 0 = outlier (see above), bit 1 = a right-hand-side cell of this lane is zero / below SF_RHS_LO / NaN.
    auto row_flags = [&](int row) -> unsigned {
        if (MODE != MODE_STRICT) return 0u;
        const float4 a = xrow_in(row), b = rring[(row & (RING_R - 1)) * 32];
        const float m = fmaxf(fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))),
                              fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w))));
        const float lo = fminf(fminf(fabsf(b.x), fabsf(b.y)), fminf(fabsf(b.z), fabsf(b.w)));
        // (lanes outside the grid hold zeros: their cells are never stored and never reach a stored cell)
        unsigned f = (!(m <= A.hi_in) ? 1u : 0u) | ((indom && !(lo >= SF_RHS_LO)) ? 2u : 0u);   // NaN counts as big and as unproven
        if constexpr (STEAL) {   // bit 2: some bit of the row is set (-0.0 counts: the zero-row shortcut below must be bit-exact)
            const unsigned any = __float_as_uint(a.x) | __float_as_uint(a.y) | __float_as_uint(a.z) | __float_as_uint(a.w) |
                                 __float_as_uint(b.x) | __float_as_uint(b.y) | __float_as_uint(b.z) | __float_as_uint(b.w);
            f |= any ? 4u : 0u;
        }
        return f;
    };
    // Scalar fields (the STEAL variants: dens_step's solve, sf_diffuse with b = 0) have compact support in the reference's
    // own initial condition (a centred source square, FluidSequential.c:252-256) and exact zeros everywhere else.  Once the
    // last 2T+3 level-0 rows (iterate and right-hand side, all 128 columns of the band) were all-zero bits, every window
    // entry and the three rows the next group emits are +0.0 exactly (b = 0: no wall negation), so the group only stores
    // zeros: ~1/8 of the instructions.  The warps that finish early take over row ranges of the busy ones (steal_next).
    [[maybe_unused]] int zero_run = 0;
    [[maybe_unused]] const bool zero_skip_ok = STEAL && A.sx == 1.0f && A.sy == 1.0f;
    // first tick none of whose numerators uses an unproven right-hand-side row (row r is used by ticks r+1 .. r+T)
    [[maybe_unused]] int unproven_until = 0;

    [[maybe_unused]] auto fuse_src = [&](int row) {      // exactly once per landed row
        if constexpr (SRC) {
            if (row <= load_hi) {
                float4 *slot = rring + (row & (RING_R - 1)) * 32;
                const float4 raw = *slot, sv = xring[(row & (RING_X - 1)) * 32];
                const float2 r01 = add2_rn(make_float2(raw.x, raw.y), mul2_exact(dup2(A.src_dt), make_float2(sv.x, sv.y), A.div.nz));
                const float2 r23 = add2_rn(make_float2(raw.z, raw.w), mul2_exact(dup2(A.src_dt), make_float2(sv.z, sv.w), A.div.nz));
                const float4 r = make_float4(r01.x, r01.y, r23.x, r23.y);
                *slot = r;
                // (a strip warp of a peer slab also stores the rows above / below its strip: the T ghost rows beside it, which
                // the later launches' strips read and nobody else forms, and T interior rows that their owner stores with
                // the same bits)
                const bool mine = STRIP ? (row >= a_lo - T && row < a_hi + T) : (row >= a_lo && row < a_hi);
                if (mine && st_ok) *reinterpret_cast<float4 *>(A.rhs_out + cell(row)) = r;
            }
        }
    };

    auto fetch = [&](int row) -> float4 {   // level-0 row `row` (after it landed)
        issue(row + PREFETCH);
        landed(row, 1);
        fuse_src(row);
        if constexpr (MODE == MODE_STRICT && !TMA && !STRIP) {
            if (row <= load_hi && __any_sync(0xffffffffu, (row_flags(row) & 2u) != 0)) unproven_until = row + T + 1;
            zero_run = 0;          // (rows taken by general ticks are not looked at: conservative)
        }
        float4 in = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!zero_guess && row <= load_hi) in = xring[(row & (RING_X - 1)) * 32];
        return in;
    };
    auto xrow = [&](int row) -> float4 {    // level-0 row already landed
        float4 in = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!zero_guess && row <= load_hi) in = xring[(row & (RING_X - 1)) * 32];
        return in;
    };
    float4 W[T][3];

    // peer-memory slabs: a boundary strip stores the rows its neighbour needs straight into the
    // neighbour's ghost rows over NVLink (plain peer stores; the exchange is part of the compute kernel)
    // peer-memory slabs: a strip warp stores every row it produces into the neighbour's ghost rows as well
    // (plain peer stores over NVLink: the exchange is part of the compute kernel)
    auto push = [&](int a, const float4 &o) {
        if constexpr (STRIP) *reinterpret_cast<float4 *>(peer + cc + (size_t)a * pitch) = o;
    };
    auto emit_plain = [&](int a, const float4 &o) {
        if (a >= first && a < a_hi && st_ok) *reinterpret_cast<float4 *>(A.xout + cell(a)) = o;
    };
    auto emit_walls = [&](int a, const float4 &o) {
        if (a < first || a >= a_hi) return;
        if (st_ok) {
            *reinterpret_cast<float4 *>(A.xout + cell(a)) = o;
            push(a, o);
        }
        if (a == 1 && A.write_top) {
            float4 w = scale4(o, A.sy);
            if (ownsL) w.x = __fmul_rn(0.5f, __fadd_rn(w.y, o.x));   // x[0][0] = .5*(x[0][1] + x[1][0])
            if (ownsR) w.w = __fmul_rn(0.5f, __fadd_rn(w.z, o.w));   // x[0][N+1] = .5*(x[0][N] + x[1][N+1])
            if (st_ok) *reinterpret_cast<float4 *>(A.xout + cell(0)) = w;
        }
        if (a == A.N && A.write_bot) {
            float4 w = scale4(o, A.sy);
            if (ownsL) w.x = __fmul_rn(0.5f, __fadd_rn(w.y, o.x));   // x[N+1][0] = .5*(x[N+1][1] + x[N][0])
            if (ownsR) w.w = __fmul_rn(0.5f, __fadd_rn(w.z, o.w));   // x[N+1][N+1] = .5*(x[N+1][N] + x[N][N+1])
            if (st_ok) *reinterpret_cast<float4 *>(A.xout + cell(A.N + 1)) = w;
        }
    };
    // The row-wall logic is needed when some level produces row 1 or row N+1 (s-t-1 in {1, N+1}, t < T)
    // or when the last level emits row 1 or row N (s = T+1, s = N+T): for s <= T+1 and for s > N.
    // Everything else runs the wall-free tick in groups of three (one full rotation of the windows).
    const int fast_lo = T + 2;
    const int fast_hi = min(s_hi, A.N);
    [[maybe_unused]] int end_seen = a_hi;    // STEAL: the slot's end as of the previous poll
    // MODE_STRICT: an out-of-range numerator or an outlier row was seen -- guarded ticks for rows below
    // slow_until (the next multiple of 64: the numerators that need them, the decaying front of a density
    // field, occupy a band of rows, not the rest of a chunk of thousands; a retry that fails wastes one
    // optimistic group per 64 rows).  Strip warps (peer slabs) only ever take the general tick (which carries
    // the neighbour push): they live for ~3T ticks, and a strip copy with the fast tick was measured to cost
    // the interior path registers.
    int slow_until = STRIP ? 0x7fffffff : 0;
    [[maybe_unused]] int span = 64;
    auto next64 = [](int row) { return (row + 63) & ~63; };
    // GROUP_VOTE: the three ticks of a fast group run without a branch between them (one basic block: the
    // ticks' dependency chains overlap, which is what the packed arithmetic needs to stay issue-bound) and
    // the warp votes once per group.  A failed vote means rows s-T.. have been emitted with a division that
    // was not proved exact and the windows hold such values: the pipeline is restarted T rows above the first
    // of them with guarded ticks (level-0 rows are re-read from global memory, 2T rows of redundant work per
    // failure; the rows are written again with the exact values by this same warp).  The bulk-copy staging
    // variant keeps a vote per tick (its mbarrier phases are tied to the first row of the pipeline).
    constexpr bool GROUP_VOTE = (MODE == MODE_STRICT) && !TMA && !STRIP;
    // the modes without a range check (pressure, fast, IEEE division) run the same branch-free group
    constexpr bool BRANCH_FREE_GROUP = (MODE != MODE_STRICT);

    // general tick at phase 0 followed by the register rotation that restores phase 0
    auto general_tick = [&](int s_, const float4 &row_in) {
        float4 o;
        pipeline_tick<T, MODE, 0, true, RB>(A, s_, row_in, W, rring, ownsL, ownsR, o);
        emit_walls(s_ - T, o);
#pragma unroll
        for (int t = 0; t < T; ++t) { W[t][0] = W[t][1]; W[t][1] = W[t][2]; }
    };

    for (;;) {   // (re)start of the pipeline at `first`
    bool restart = false;
    unproven_until = 0;          // every row the restarted pipeline uses lands (and is looked at) again
    zero_run = 0;
#pragma unroll
    for (int t = 0; t < T; ++t) W[t][0] = W[t][1] = W[t][2] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < PREFETCH; ++k) issue(s_lo + k);

    int s = s_lo;
    while (s <= s_hi) {
        if constexpr (STEAL) {
            // once per 32 rows (the test hits exactly one s of any window of 32, in steps of 1 or 3): publish
            // the progress and look at the slot's end.  Rows leave the pipeline in increasing order, so once
            // every row below an end lowered by a thief has been emitted (s - T >= end) this warp is done;
            // nothing else about the loop changes.  The load is consumed one poll LATER: no stall.
            // (every 8 rows inside a guarded span: those rows cost ~3x as much, and the ranges worth taking there are short)
            if (((s - a_lo) & (slow_until > s ? 7 : 31)) < 3) {
                if (end_seen <= s - T) break;
                // (the slot address is recomputed here rather than kept in registers across the hot loop)
                StealSlot *slot = A.steal->slots + (blockIdx.x * WPC + (threadIdx.x >> 5));
                if (lane == 0) st_relaxed_gpu(&slot->pos, s);
                end_seen = ld_relaxed_gpu(&slot->end);
            }
        }
        const bool slow = s < slow_until;
        if (!slow && s >= fast_lo && s + 2 <= fast_hi) {
            // consecutive groups of the modes without a range check run in this inner loop (the strict group already gets a
            // loop of its own from ptxas): every path through the body ends in `continue` (-> the inner condition) or `break`
            constexpr bool INNER_RUN = (MODE != MODE_STRICT);
            [[maybe_unused]] bool stolen = false;
            [[maybe_unused]] const int s_run = s;
            do {
            if constexpr (STEAL) {   // the same poll as above, for the groups after the first of this run
                if (s != s_run && ((s - a_lo) & 31) < 3) {
                    if (end_seen <= s - T) { stolen = true; break; }
                    StealSlot *slot = A.steal->slots + (blockIdx.x * WPC + (threadIdx.x >> 5));
                    if (lane == 0) st_relaxed_gpu(&slot->pos, s);
                    end_seen = ld_relaxed_gpu(&slot->end);
                }
            }
            issue3(s + PREFETCH);
            landed(s, 3);                    // rows <= s+2 have landed
            fuse_src(s); fuse_src(s + 1); fuse_src(s + 2);
            bool big = false;
            if (MODE == MODE_STRICT) {
                // one warp-wide OR for both row properties of the three rows that have just landed
                const unsigned fl = __reduce_or_sync(0xffffffffu, row_flags(s) | row_flags(s + 1) | row_flags(s + 2));
                big = (fl & 1u) != 0;
                if (fl & 2u) unproven_until = s + 2 + T + 1;
                if constexpr (STEAL) zero_run = (fl & 4u) ? 0 : zero_run + 3;
            }
            if (!big) {
                float4 o;
                if constexpr (GROUP_VOTE || BRANCH_FREE_GROUP) {
                    bool ok;
                    // the three rows this group emits: one offset, then +G, +2G
                    const cell_t e0 = cell(s - T);
                    auto emit_plain = [&](int a, const float4 &ov) {
                        if (a >= first && a < a_hi && st_ok) *reinterpret_cast<float4 *>(A.xout + (e0 + (cell_t)(a - (s - T)) * Gu)) = ov;
                    };
                    if constexpr (STEAL) {
                        if (zero_skip_ok && zero_run >= 2 * T + 3) {
                            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                            emit_plain(s - T, z); emit_plain(s + 1 - T, z); emit_plain(s + 2 - T, z);
                            s += 3;
                            continue;
                        }
                    }
                    // (not in the work-stealing variants: measured, the extra copy of the group costs them more than it saves
                    // -- 299 -> 331 us per launch on a density field, most of which is outside the proven region anyway)
                    if constexpr (GROUP_VOTE && !STEAL) {
                        if (unproven_until <= s) {
                            // every right-hand-side row these three ticks use is proven: no range test, no vote, no restart
                            pipeline_tick<T, MODE, 0, false, RB, false, false>(A, s, xrow_in(s), W, rring, ownsL, ownsR, o);
                            emit_plain(s - T, o);
                            pipeline_tick<T, MODE, 1, false, RB, false, false>(A, s + 1, xrow_in(s + 1), W, rring, ownsL, ownsR, o);
                            emit_plain(s + 1 - T, o);
                            pipeline_tick<T, MODE, 2, false, RB, false, false>(A, s + 2, xrow_in(s + 2), W, rring, ownsL, ownsR, o);
                            emit_plain(s + 2 - T, o);
                            s += 3;
                            continue;
                        }
                    }
                    ok = pipeline_tick<T, MODE, 0, false, RB>(A, s, xrow_in(s), W, rring, ownsL, ownsR, o);
                    emit_plain(s - T, o);
                    ok &= pipeline_tick<T, MODE, 1, false, RB>(A, s + 1, xrow_in(s + 1), W, rring, ownsL, ownsR, o);
                    emit_plain(s + 1 - T, o);
                    ok &= pipeline_tick<T, MODE, 2, false, RB>(A, s + 2, xrow_in(s + 2), W, rring, ownsL, ownsR, o);
                    emit_plain(s + 2 - T, o);
                    if (GROUP_VOTE && !__all_sync(0xffffffffu, ok)) {
                        first = max(first, s - T);     // rows below it were emitted by groups that passed
                        // a retry that fails straight away doubles the guarded span (64 .. 512 rows): where the
                        // front of a density field runs ALONG a band, every retry would cost a restart
                        span = (s < slow_until + 6) ? min(2 * span, 512) : 64;
                        slow_until = next64(s + 3) + span - 64;
                        restart = true;
                        break;
                    }
                    s += 3;
                    continue;
                } else {
                bool ok = pipeline_tick<T, MODE, 0, false, RB>(A, s, xrow_in(s), W, rring, ownsL, ownsR, o);
                if (MODE == MODE_STRICT && !__all_sync(0xffffffffu, ok)) {
                    slow_until = next64(s + 3);        // windows are still at phase 0: redo guarded
                    general_tick(s, xrow(s)); s += 1;
                    general_tick(s, xrow(s)); s += 1;  // the two other rows of this group are in flight already
                    general_tick(s, xrow(s)); s += 1;
                    continue;
                }
                emit_plain(s - T, o);
                ok = pipeline_tick<T, MODE, 1, false, RB>(A, s + 1, xrow_in(s + 1), W, rring, ownsL, ownsR, o);
                if (MODE == MODE_STRICT && !__all_sync(0xffffffffu, ok)) {
                    slow_until = next64(s + 3);        // phase 1 -> phase 0: up = slot 1, mid = slot 2
#pragma unroll
                    for (int t = 0; t < T; ++t) { W[t][0] = W[t][1]; W[t][1] = W[t][2]; }
                    general_tick(s + 1, xrow(s + 1));
                    general_tick(s + 2, xrow(s + 2));
                    s += 3;
                    continue;
                }
                emit_plain(s + 1 - T, o);
                ok = pipeline_tick<T, MODE, 2, false, RB>(A, s + 2, xrow_in(s + 2), W, rring, ownsL, ownsR, o);
                if (MODE == MODE_STRICT && !__all_sync(0xffffffffu, ok)) {
                    slow_until = next64(s + 3);        // phase 2 -> phase 0: up = slot 2, mid = slot 0
#pragma unroll
                    for (int t = 0; t < T; ++t) { const float4 m = W[t][0]; W[t][0] = W[t][2]; W[t][1] = m; }
                    general_tick(s + 2, xrow(s + 2));
                    s += 3;
                    continue;
                }
                emit_plain(s + 2 - T, o);
                s += 3;
                continue;
                }
            }
            // outlier row: the three rows of this group are already in flight; run them guarded
            slow_until = next64(s + 3);
            general_tick(s, xrow(s)); s += 1;
            general_tick(s, xrow(s)); s += 1;
            general_tick(s, xrow(s)); s += 1;
            continue;
            } while (INNER_RUN && s >= slow_until && s + 2 <= fast_hi);     // (s >= fast_lo holds for the whole run)
            if (restart || stolen) break;
            continue;
        }
        // guarded span: wall-free groups of three ticks with the binary64 division (same windows, no register rotation)
        if constexpr (MODE == MODE_STRICT && !TMA && !STRIP) {
            if (slow && s >= fast_lo && s + 2 <= fast_hi && s + 2 < slow_until) {
                issue3(s + PREFETCH);
                landed(s, 3);
                fuse_src(s); fuse_src(s + 1); fuse_src(s + 2);
                if (__any_sync(0xffffffffu, ((row_flags(s) | row_flags(s + 1) | row_flags(s + 2)) & 2u) != 0)) unproven_until = s + 2 + T + 1;
                zero_run = 0;
                float4 o;
                pipeline_tick<T, MODE, 0, false, RB, true>(A, s, xrow_in(s), W, rring, ownsL, ownsR, o);
                emit_plain(s - T, o);
                pipeline_tick<T, MODE, 1, false, RB, true>(A, s + 1, xrow_in(s + 1), W, rring, ownsL, ownsR, o);
                emit_plain(s + 1 - T, o);
                pipeline_tick<T, MODE, 2, false, RB, true>(A, s + 2, xrow_in(s + 2), W, rring, ownsL, ownsR, o);
                emit_plain(s + 2 - T, o);
                s += 3;
                continue;
            }
        }
        general_tick(s, fetch(s));
        ++s;
    }
    if (!restart) break;
    cp_async_wait<0>();        // rows in flight land before their ring slots are reused
    s_lo = max(first - T, 0);
    }
    if (TMA) {   // drain: rows issued beyond the last one consumed must land before the CTA's smem is released
        for (int row = s_hi + 1; row <= min(s_hi + PREFETCH + 2, load_hi); ++row)
            mbar_wait(bars + (row & (RING_X - 1)), (unsigned)((row - s_lo) >> 3) & 1u);
    } else {
        cp_async_wait<0>();
    }
}

// ---- row-level work stealing (see StealSlot in sf_common.cuh) -------------------------------------
// Out of line and self-contained (they recompute the warp's item from blockIdx / threadIdx), so that
// nothing of the bookkeeping is live in registers across the streaming loop.
__device__ __noinline__ void steal_publish(StealCtl *ctl, int band, int lo, int hi)
{
    if ((threadIdx.x & 31) == 0) {
        StealSlot *slot = ctl->slots + (blockIdx.x * WPC + (threadIdx.x >> 5));
        st_relaxed_gpu(&slot->end, hi); st_relaxed_gpu(&slot->band, band);
        st_relaxed_gpu(&slot->tag, ld_relaxed_gpu(&ctl->epoch) + 1);
        __threadfence();
        st_relaxed_gpu(&slot->pos, lo);      // a stale slot has pos = 0x3fffffff: pos goes last
    }
    __syncwarp();
}
// The warp has finished its range: sample 32 slots per attempt and halve the largest remaining range
// (returns the upper half in band / lo / hi), or close the warp's part of the launch and return false.
// seg_lo / seg_hi: the launch's interior segment; a candidate that does not lie inside it is never touched.
__device__ __noinline__ bool steal_next(StealCtl *ctl, int nitems, int chunk_rows, int seg_lo, int seg_hi, int &band, int &lo, int &hi)
{
    const int lane = threadIdx.x & 31, item = blockIdx.x * WPC + (threadIdx.x >> 5);
    StealSlot *slot = ctl->slots + item;
    const int tag = ld_relaxed_gpu(&ctl->epoch) + 1;
    // only clear outliers are worth the 2T halo rows and the pipeline fill of a fresh start: warps of a
    // balanced launch finish up to ~20 % apart anyway (the schedulers favour the oldest warp)
    // (for chunks of thousands of rows a fresh start is cheap in proportion: the threshold stops growing at 256 rows)
    const int steal_min = max(STEAL_MIN_ROWS, min(chunk_rows * ctl->min_pct / 100, 256));
    if (lane == 0) st_relaxed_gpu(&slot->pos, 0x3fffffff);     // nothing left to take here
    for (int attempt = 0; attempt < 2; ++attempt) {
        const int cand = (int)(((unsigned)item + 1u + (unsigned)lane * 37u + (unsigned)attempt * 1187u) % (unsigned)nitems);
        const StealSlot *S = ctl->slots + cand;
        // pos is written last by its owner (after a fence) and read first here: a valid pos vouches for the rest
        const int cpos = ld_acquire_gpu(&S->pos);
        const int ctag = ld_relaxed_gpu(&S->tag), cend = ld_relaxed_gpu(&S->end), cband = ld_relaxed_gpu(&S->band);
        const bool valid = ctag == tag && cand != item && cpos >= seg_lo - HALO_X && cend <= seg_hi && cend > cpos;
        const int rem = valid ? cend - cpos : 0;    // rows the owner has not reached yet
        int best = rem, who = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int ob = __shfl_xor_sync(0xffffffffu, best, o), ow = __shfl_xor_sync(0xffffffffu, who, o);
            if (ob > best || (ob == best && ow < who)) { best = ob; who = ow; }
        }
        if (best < steal_min) continue;
        int mid = 0, ok = 0;
        if (lane == who) {
            mid = cend - rem / 2;                       // the thief takes the upper half [mid, cend)
            // 64-bit compare-and-swap on (end, band): succeeds only against the very range that was sampled
            const unsigned long long hi32 = (unsigned long long)(unsigned)cband << 32;
            const unsigned long long want = hi32 | (unsigned)cend, put = hi32 | (unsigned)mid;
            ok = (atomicCAS(reinterpret_cast<unsigned long long *>(&ctl->slots[cand].end), want, put) == want) ? 1 : 0;
            if (ok) atomicAdd(&ctl->taken, 1);
        }
        ok = __shfl_sync(0xffffffffu, ok, who);
        if (ok) {
            lo = __shfl_sync(0xffffffffu, mid, who);
            hi = __shfl_sync(0xffffffffu, cend, who);
            band = __shfl_sync(0xffffffffu, cband, who);
            if (lane == 0) {     // the taken range is this warp's published range now
                st_relaxed_gpu(&slot->end, hi); st_relaxed_gpu(&slot->band, band);
                __threadfence();
                st_relaxed_gpu(&slot->pos, lo);
            }
            __syncwarp();
            return true;
        }
    }
    if (lane == 0) {     // the last warp of the launch closes the epoch: the next launch ignores these slots
        if (atomicAdd(&ctl->done, 1) == nitems - 1) {
            st_relaxed_gpu(&ctl->done, 0);
            __threadfence();
            st_relaxed_gpu(&ctl->epoch, tag);
        }
    }
    return false;
}

// VAR = 2: cp.async staging plus the fused strip exchange of peer-memory slabs.  The strip warps run
// their own (out-of-line) copy of the pipeline, so the interior warps execute exactly the code of the
// single-GPU kernel.
template <int T, int MODE, bool SRC = false>
__device__ __noinline__ void strip_warp(const StreamArgs A, float4 *ring, const int lane, const int warp, const int item0, const int n_top)
{
    const StripArgs *S = A.strips;
    const bool top = item0 < n_top;
    const StripPort *P = &S->port[top ? 0 : 1];
    strip_wait(P->seq, P->inbox, S->error, S->timeout_ns, lane);
    const int a_lo = top ? S->o_lo : S->o_hi - P->rows;
    const int a_hi = top ? S->o_lo + P->rows : S->o_hi;
    float *peer = P->xpeer - (ptrdiff_t)P->peer_row_base * (ptrdiff_t)A.G;
    stream_rows<T, MODE, false, true, false, false, SRC>(A, ring, lane, warp, top ? item0 : item0 - n_top, a_lo, a_hi, peer);
    strip_post(P->arrive, P->seq, P->nbr_inbox, A.nbands, lane);
}

#ifdef SF_WARP_TIMES
// developer build (-DSF_WARP_TIMES): every range a warp streams is logged as (start ns, end ns, band, lo, hi, stolen) into a
// device buffer set with sf_debug_warp_times -- where a launch's time goes when its warps are not equally loaded
struct WarpTimeRec { unsigned long long t0, t1; int band, lo, hi, stolen; };
__device__ WarpTimeRec *g_warp_times = nullptr;
__device__ unsigned int g_warp_times_n = 0, g_warp_times_cap = 0;
__device__ __forceinline__ void log_range(unsigned long long t0, int band, int lo, int hi, int stolen)
{
    if ((threadIdx.x & 31) == 0 && g_warp_times != nullptr) {
        const unsigned k = atomicAdd(&g_warp_times_n, 1u);
        const unsigned smid = hw_sm_id(), warpid = hw_warp_slot();
        if (k < g_warp_times_cap) g_warp_times[k] = WarpTimeRec{t0, globaltimer_ns(), band, lo, hi, stolen | (int)(smid << 8) | (int)(warpid << 20)};
    }
}
#endif

template <int T, int MODE, int VAR>
__global__ void __launch_bounds__(WPC * 32, min_ctas<T, MODE>()) jacobi_stream_kernel(const StreamArgs A)
{
    constexpr bool TMA = (VAR == 1), STRIPS = (VAR == 2 || VAR == 4 || VAR == 8 || VAR == 9);
    constexpr bool STEALS = (VAR == 3 || VAR == 4 || VAR == 7 || VAR == 9), RB = (VAR == 5);
    // fused add_source (first launch of a solve): 6 / 7 without / with work stealing, 8 / 9 the same on a peer slab (strips)
    constexpr bool SRC = (VAR == 6 || VAR == 7 || VAR == 8 || VAR == 9);
    extern __shared__ float4 ring[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // work item = (band, row chunk); consecutive warps take consecutive bands of the same chunk.  warps take their items in
    // the order they start (see chunk_range); the warp that draws the last ticket re-arms the counter for the next launch.
    int item = blockIdx.x * WPC + warp;
    if (A.ticket != nullptr) {
        // Unequal chunks by scheduling priority (see chunk_range).  What a warp scheduler favours is the warp in the LOWEST
        // hardware slot: with three resident CTAs per SM, %warpid / WPC = 0, 1, 2 ran at 2.37 / 1.98 / 1.58 rows per us on
        // equal chunks, cleanly separated (tools/warp_times.py --hw).  blockIdx does not tell the slot for a few percent of
        // the CTAs, and neither does the order in which warps start -- so every warp reads its slot and draws its item
        // from that class's counter (class c owns the c-th third of the items; a class that ever runs dry -- %warpid is
        // only a hint -- falls through to the next one, so every item is drawn exactly once whatever the slots are).
        // The warp that arrives last re-arms the counters for the next launch.
        int it = 0;
        if (lane == 0) {
            const unsigned wid = hw_warp_slot();
            const unsigned ipc = A.skew_cpw * (unsigned)A.nbands;
            const int cls = min((int)(wid / WPC), 2);
            it = A.nbands * A.nchunks;                     // nothing left: retire
            for (int k = 0; k < 3; ++k) {
                const int c = cls + k < 3 ? cls + k : cls + k - 3;
                const unsigned t = atomicAdd(A.ticket + c, 1u);
                if (t < ipc) { it = c * (int)ipc + (int)t; break; }
            }
            if (atomicAdd(A.ticket + 3, 1u) == gridDim.x * WPC - 1) {
                A.ticket[0] = 0u; A.ticket[1] = 0u; A.ticket[2] = 0u; A.ticket[3] = 0u;
            }
        }
        item = __shfl_sync(0xffffffffu, it, 0);
    }
    if constexpr (STRIPS) {
        // the first warps of the grid compute a boundary strip BEFORE their interior item: the strips are
        // scheduled first and travel while everybody computes, and the grid still is one wave of warps
        const StripArgs *S = A.strips;
        const int n_top = S->port[0].rows > 0 ? A.nbands : 0, n_bot = S->port[1].rows > 0 ? A.nbands : 0;
        if (item < n_top + n_bot) strip_warp<T, MODE, SRC>(A, ring, lane, warp, item, n_top);
    }
    const int nitems = A.nbands * A.nchunks;
    if (item >= nitems) return;
    const int band = item % A.nbands, chunk = item / A.nbands;
    int a_lo, a_hi;
    chunk_range<STRIPS>(A, chunk, a_lo, a_hi);
    if constexpr (!STEALS) {
        if (a_lo >= a_hi) return;
#ifdef SF_WARP_TIMES
        const unsigned long long t0 = globaltimer_ns();
#endif
        stream_rows<T, MODE, TMA, false, false, RB, SRC>(A, ring, lane, warp, band, a_lo, a_hi, nullptr);
#ifdef SF_WARP_TIMES
        log_range(t0, band, a_lo, a_hi, 0);
#endif
    } else {
        int b = band, lo = a_lo, hi = a_hi;
        steal_publish(A.steal, b, lo, hi);
        [[maybe_unused]] int stolen = 0;
        for (;;) {
#ifdef SF_WARP_TIMES
            const unsigned long long t0 = globaltimer_ns();
#endif
            if (lo < hi) stream_rows<T, MODE, false, false, true, false, SRC>(A, ring, lane, warp, b, lo, hi, nullptr);
#ifdef SF_WARP_TIMES
            log_range(t0, b, lo, hi, stolen);      // (hi = the range as taken; the owner may have stopped earlier)
            stolen = 1;
#endif
            if (!steal_next(A.steal, nitems, A.chunk_rows, A.a_lo, A.a_hi, b, lo, hi)) break;
        }
    }
}

// ---- generic fallback: one sweep, one thread per interior cell, any G -----------------------
template <int MODE>
__global__ void jacobi_generic_kernel(const float *__restrict__ xin, const float *__restrict__ rhs,
                                      float *__restrict__ xout, Geom g, int a_lo, int a_hi, int write_top,
                                      int write_bot, float alpha, DivConst d, float sx, float sy)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int row = blockIdx.y * blockDim.y + threadIdx.y + a_lo;
    if (col > g.N || row >= a_hi) return;
    const size_t G = (size_t)g.G;
    const size_t i = (size_t)(row - g.row_base) * G + col;
    const float o = jacobi_cell<MODE>(xin[i - 1], xin[i + 1], xin[i - G], xin[i + G], rhs[i], alpha, d);
    xout[i] = o;
    const float wx = __fmul_rn(sx, o), wy = __fmul_rn(sy, o);
    const bool L = (col == 1), R = (col == g.N);
    const bool Tp = (row == 1) && write_top, Bt = (row == g.N) && write_bot;
    if (L) xout[i - 1] = wx;
    if (R) xout[i + 1] = wx;
    if (Tp) xout[i - G] = wy;
    if (Bt) xout[i + G] = wy;
    if (L && Tp) xout[i - G - 1] = __fmul_rn(0.5f, __fadd_rn(wy, wx));
    if (R && Tp) xout[i - G + 1] = __fmul_rn(0.5f, __fadd_rn(wy, wx));
    if (L && Bt) xout[i + G - 1] = __fmul_rn(0.5f, __fadd_rn(wy, wx));
    if (R && Bt) xout[i + G + 1] = __fmul_rn(0.5f, __fadd_rn(wy, wx));
}

// ---- exhaustive validation of div_const for one beta ----------------------------------------
__global__ void validate_division_kernel(DivConst d, unsigned long long *mismatches)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long bad = 0;
    for (unsigned long long u = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; u < (1ull << 32); u += stride) {
        const float a = __uint_as_float((unsigned)u);
        if (a != a) continue;   // NaN numerators: payload propagation is not part of the contract
        const unsigned want = __float_as_uint(__fdiv_rn(a, d.b));
        bad += (want != __float_as_uint(div_const(a, d)));        // guarded fast path + slow path outside the range
        bad += (want != __float_as_uint(div_const_slow(a, d)));   // the binary64 step on its own, every numerator
#if SF_PACKED_F32
        if (div_in_range(a)) {   // the two-wide form the streaming kernel uses (FMUL2 / FFMA2), both halves
            const float2 q = div_const_fast2(make_float2(a, __uint_as_float((unsigned)u ^ 0x80000000u)), d);
            bad += (want != __float_as_uint(q.x)) + ((want ^ 0x80000000u) != __float_as_uint(q.y));
        }
#endif
    }
    if (bad) atomicAdd(mismatches, bad);
}

template <int T, int MODE>
cudaError_t launch_stream_T(const StreamArgs &A, dim3 grid, size_t smem, bool tma, cudaStream_t st)
{
    static_assert((size_t)WPC * (RING_X + RING_R) * 32 * sizeof(float4) <= 48 * 1024,
                  "ring fits the default 48 KB dynamic shared memory limit (no attribute call needed)");
    // the bulk-copy variant is built for the depths the default launch plans use (5, 6, 7)
    if (A.rhs_out != nullptr) {
        // fused add_source: built for the depths a default launch plan starts with and the two bit-exact divisions
        if constexpr ((T == 5 || T == 6 || T == 7) && (MODE == MODE_STRICT || MODE == MODE_IEEE)) {
            if (tma) return cudaErrorNotSupported;
            if constexpr (MODE == MODE_STRICT) {
                if (A.steal != nullptr) {
                    if (A.strips != nullptr) jacobi_stream_kernel<T, MODE_STRICT, 9><<<grid, WPC * 32, smem, st>>>(A);
                    else jacobi_stream_kernel<T, MODE_STRICT, 7><<<grid, WPC * 32, smem, st>>>(A);
                    return cudaGetLastError();
                }
            }
            if (A.strips != nullptr) jacobi_stream_kernel<T, MODE, 8><<<grid, WPC * 32, smem, st>>>(A);
            else jacobi_stream_kernel<T, MODE, 6><<<grid, WPC * 32, smem, st>>>(A);
            return cudaGetLastError();
        } else {
            return cudaErrorNotSupported;
        }
    }
    // (the bulk-copy variant has no strip warps: a peer-slab launch with strips takes the cp.async kernels below)
    if (tma && A.strips == nullptr && (T == 5 || T == 6 || T == 7) && (MODE == MODE_STRICT || MODE == MODE_PRESSURE)) {
        constexpr int TT = (T == 5 || T == 6 || T == 7) ? T : 7;
        constexpr int MM = (MODE == MODE_STRICT || MODE == MODE_PRESSURE) ? MODE : MODE_PRESSURE;
        const size_t smem_tma = smem + (size_t)WPC * RING_X * sizeof(uint64_t);
        {   // above the 48 KB default; the attribute is per device, so it is set on every launch (a cheap host-side call)
            cudaError_t e = cudaFuncSetAttribute(jacobi_stream_kernel<TT, MM, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tma);
            if (e != cudaSuccess) return e;
        }
        jacobi_stream_kernel<TT, MM, 1><<<grid, WPC * 32, smem_tma, st>>>(A);
        return cudaGetLastError();
    }
    if constexpr (MODE == MODE_STRICT) {
        // work stealing is built for the one mode whose ticks have a data-dependent cost
        if (A.steal != nullptr) {
            if (A.strips != nullptr) jacobi_stream_kernel<T, MODE_STRICT, 4><<<grid, WPC * 32, smem, st>>>(A);
            else jacobi_stream_kernel<T, MODE_STRICT, 3><<<grid, WPC * 32, smem, st>>>(A);
            return cudaGetLastError();
        }
    }
    if (A.strips != nullptr) {
        // the peer-slab variant is built for the arithmetic modes a step uses
        if (MODE == MODE_FAST || MODE == MODE_IEEE) {
            constexpr int MM = (MODE == MODE_FAST || MODE == MODE_IEEE) ? MODE_IEEE : MODE;
            if (MODE == MODE_FAST) return cudaErrorNotSupported;
            jacobi_stream_kernel<T, MM, 2><<<grid, WPC * 32, smem, st>>>(A);
        } else {
            jacobi_stream_kernel<T, MODE, 2><<<grid, WPC * 32, smem, st>>>(A);
        }
        return cudaGetLastError();
    }
    jacobi_stream_kernel<T, MODE, 0><<<grid, WPC * 32, smem, st>>>(A);
    return cudaGetLastError();
}

// red-black levels (VAR = 5): an even number of levels per launch (whole iterations), built for depths 2, 4 and 6
template <int MODE>
cudaError_t launch_stream_rb(int T, const StreamArgs &A, dim3 grid, size_t smem, cudaStream_t st)
{
    switch (T) {
        case 2: jacobi_stream_kernel<2, MODE, 5><<<grid, WPC * 32, smem, st>>>(A); break;
        case 4: jacobi_stream_kernel<4, MODE, 5><<<grid, WPC * 32, smem, st>>>(A); break;
        case 6: jacobi_stream_kernel<6, MODE, 5><<<grid, WPC * 32, smem, st>>>(A); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

template <int MODE>
cudaError_t launch_stream_mode(int T, const StreamArgs &A, dim3 grid, size_t smem, bool tma, cudaStream_t st)
{
    switch (T) {
        case 1: return launch_stream_T<1, MODE>(A, grid, smem, tma, st);
        case 2: return launch_stream_T<2, MODE>(A, grid, smem, tma, st);
        case 3: return launch_stream_T<3, MODE>(A, grid, smem, tma, st);
        case 4: return launch_stream_T<4, MODE>(A, grid, smem, tma, st);
        case 5: return launch_stream_T<5, MODE>(A, grid, smem, tma, st);
        case 6: return launch_stream_T<6, MODE>(A, grid, smem, tma, st);
        case 7: return launch_stream_T<7, MODE>(A, grid, smem, tma, st);
        case 8: return launch_stream_T<8, MODE>(A, grid, smem, tma, st);
        default: return cudaErrorInvalidValue;
    }
}

std::mutex g_div_mutex;
std::map<uint32_t, bool> g_div_ok;   // beta bits -> div_const verified bit-identical to __fdiv_rn

}  // namespace

// CUDA loads kernels lazily, and loading one while other kernels run synchronises the context.  A peer
// slab must never hit that inside a step (its neighbour barrier kernels may be spinning on the device
// at that moment), so every kernel a step can launch is loaded up front.
namespace {
template <int T, int MODE>
void preload_T(cudaFuncAttributes &a)
{
    cudaFuncGetAttributes(&a, jacobi_stream_kernel<T, MODE, 0>);
    if (MODE != MODE_FAST) cudaFuncGetAttributes(&a, jacobi_stream_kernel<T, MODE == MODE_FAST ? MODE_IEEE : MODE, 2>);
    if (MODE == MODE_STRICT) {
        cudaFuncGetAttributes(&a, jacobi_stream_kernel<T, MODE_STRICT, 3>);
        cudaFuncGetAttributes(&a, jacobi_stream_kernel<T, MODE_STRICT, 4>);
    }
    if constexpr ((T == 5 || T == 6 || T == 7) && (MODE == MODE_STRICT || MODE == MODE_IEEE)) {
        cudaFuncGetAttributes(&a, jacobi_stream_kernel<T, MODE, 6>);
        cudaFuncGetAttributes(&a, jacobi_stream_kernel<T, MODE, 8>);
        if constexpr (MODE == MODE_STRICT) {
            cudaFuncGetAttributes(&a, jacobi_stream_kernel<T, MODE_STRICT, 7>);
            cudaFuncGetAttributes(&a, jacobi_stream_kernel<T, MODE_STRICT, 9>);
        }
    }
}
template <int MODE>
void preload_mode()
{
    cudaFuncAttributes a;
    preload_T<1, MODE>(a); preload_T<2, MODE>(a); preload_T<3, MODE>(a); preload_T<4, MODE>(a);
    preload_T<5, MODE>(a); preload_T<6, MODE>(a); preload_T<7, MODE>(a); preload_T<8, MODE>(a);
    cudaFuncGetAttributes(&a, jacobi_generic_kernel<MODE>);
}
}  // namespace
void preload_jacobi_kernels()
{
    preload_mode<MODE_STRICT>(); preload_mode<MODE_PRESSURE>(); preload_mode<MODE_FAST>(); preload_mode<MODE_IEEE>();
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, jacobi_stream_kernel<5, MODE_STRICT, 1>); cudaFuncGetAttributes(&a, jacobi_stream_kernel<6, MODE_STRICT, 1>);
    cudaFuncGetAttributes(&a, jacobi_stream_kernel<7, MODE_STRICT, 1>); cudaFuncGetAttributes(&a, jacobi_stream_kernel<5, MODE_PRESSURE, 1>);
    cudaFuncGetAttributes(&a, jacobi_stream_kernel<6, MODE_PRESSURE, 1>); cudaFuncGetAttributes(&a, jacobi_stream_kernel<7, MODE_PRESSURE, 1>);
    cudaFuncGetAttributes(&a, validate_division_kernel);
    (void)cudaGetLastError();
}

// widths that are a multiple of 4 (16-byte rows) and fields of fewer than 2^32 cells (32-bit cell offsets in the
// streaming kernel: up to G = 65532 on one GPU, any BASELINE size); everything else runs the generic kernel
bool jacobi_stream_supported(const Geom &g)
{
    return (g.G % 4) == 0 && g.G >= 4 && (!SF_OFF32 || (unsigned long long)g.rows * (unsigned long long)g.G < (1ull << 32));
}

bool division_validated(float beta, bool allow_run, cudaStream_t st)
{
    if (!(beta > 0.0f) || !(beta < 3.0e38f) || beta < 1.2e-38f) return false;
    uint32_t key;
    std::memcpy(&key, &beta, sizeof(key));
    std::lock_guard<std::mutex> lock(g_div_mutex);
    auto it = g_div_ok.find(key);
    if (it != g_div_ok.end()) return it->second;
    if (!allow_run) return false;
    unsigned long long *dev = nullptr, host = ~0ull;
    bool ok = false;
    if (cudaMalloc(&dev, sizeof(*dev)) == cudaSuccess) {
        if (cudaMemsetAsync(dev, 0, sizeof(*dev), st) == cudaSuccess) {
            validate_division_kernel<<<148 * 16, 256, 0, st>>>(make_div_const(beta), dev);
            if (cudaGetLastError() == cudaSuccess &&
                cudaMemcpyAsync(&host, dev, sizeof(host), cudaMemcpyDeviceToHost, st) == cudaSuccess &&
                cudaStreamSynchronize(st) == cudaSuccess)
                ok = (host == 0);
        }
        cudaFree(dev);
    }
    g_div_ok[key] = ok;
    return ok;
}

cudaError_t launch_jacobi_stream(const Geom &g, const JacobiLaunch &L, int sm_count, cudaStream_t st)
{
    if (L.sweeps < 1 || L.sweeps > HALO_X) return cudaErrorInvalidValue;
    StreamArgs A;
    A.xin = L.xin; A.rhs = L.rhs; A.xout = L.xout;
    A.G = g.G; A.N = g.N; A.row_base = g.row_base;
    const int st_top = L.strips ? L.strip_rows[0] : 0, st_bot = L.strips ? L.strip_rows[1] : 0;
    A.strips = (st_top > 0 || st_bot > 0) ? L.strips : nullptr;
    if (st_top < 0 || st_bot < 0 || st_top + st_bot > L.out_hi - L.out_lo) return cudaErrorInvalidValue;
    if ((st_top > 0 && L.out_lo < 1) || (st_bot > 0 && L.out_hi > g.N + 1)) return cudaErrorInvalidValue;   // a strip faces a neighbour, never a wall
    // interior segment: the output rows between the strips
    A.a_lo = max(L.out_lo + st_top, 1);
    A.a_hi = min(L.out_hi - st_bot, g.N + 1);
    A.write_top = (L.out_lo == 0);
    A.write_bot = (L.out_hi == g.G);
    A.nbands = (g.G + VALID_W - 1) / VALID_W;
    A.zero_guess = L.zero_guess;
    A.rhs_out = L.rhs_out; A.src_dt = L.src_dt;
    if (L.rhs_out != nullptr && (L.zero_guess || L.rb)) return cudaErrorInvalidValue;
    A.alpha = L.alpha; A.div = make_div_const(L.beta);
    {   // see row_is_big: bound on level-0 magnitudes that keeps all numerators of T sweeps in range
        const double F = 1.0 + 4.0 * fabs((double)L.alpha);
        double g = F / fabs((double)L.beta) * 1.000001;
        if (g < 1.0) g = 1.0;
        double hi = (double)SF_DIV_HI / (F * 1.01);
        for (int t = 0; t < L.sweeps; ++t) hi /= g;
        A.hi_in = (float)hi;
    }
    A.sx = (L.b == 1) ? -1.0f : 1.0f;
    A.sy = (L.b == 2) ? -1.0f : 1.0f;
    const int n_strip_items = (st_top > 0 ? A.nbands : 0) + (st_bot > 0 ? A.nbands : 0);
    const int rows = A.a_hi - A.a_lo;
    if (rows <= 0 && n_strip_items == 0) return cudaSuccess;
    int chunk = L.chunk_rows;
    if (chunk <= 0) {
        // One work item (band x chunk) per resident warp (12 or 16 warps per SM).  All items cost about
        // the same, so a single full wave has no tail.  Large grids get chunks of hundreds of rows (2T
        // redundant halo rows each: a few percent); small grids cannot fill the wave with such chunks and
        // are latency-bound, so there the chunks shrink (down to 2T rows) to put every SM to work.
        const bool heavy = (L.mode == MODE_STRICT || L.mode == MODE_IEEE || L.mode == MODE_PRESSURE) && L.sweeps >= 6;   // min_ctas<T, MODE>()
        const int slots = sm_count * (heavy ? 3 : 4) * WPC;
        int want_chunks = slots / A.nbands;
        if (want_chunks < 1) want_chunks = 1;
        chunk = (max(rows, 1) + want_chunks - 1) / want_chunks;
        const int min_chunk = 2 * L.sweeps > 8 ? 2 * L.sweeps : 8;
        if (chunk < min_chunk) chunk = min_chunk;
    }
    if (chunk > rows) chunk = max(rows, 1);
    A.chunk_rows = chunk;
    A.nchunks = rows > 0 ? (rows + chunk - 1) / chunk : 0;
    A.skew = 0u; A.skew_cpw = 0u; A.ticket = nullptr;
    if (L.wave_skew > 0 && L.ticket != nullptr && L.chunk_rows <= 0 && n_strip_items == 0 && A.nchunks > 0) {
        // three CTAs per SM (min_ctas<T, MODE>() == 3) and a grid that really is one full wave: thirds of the items = the
        // first / second / third CTA of every SM
        const bool heavy3 = (L.mode == MODE_STRICT || L.mode == MODE_IEEE || L.mode == MODE_PRESSURE) && L.sweeps >= 6;
        const int slots = sm_count * 3 * WPC, items = A.nbands * A.nchunks;
        const int r0 = chunk * (L.wave_skew / 1000) / 100, r1 = chunk * (L.wave_skew % 1000) / 100, r2 = 3 * chunk - r0 - r1;
        if (heavy3 && A.nchunks % 3 == 0 && items <= slots && items * 10 >= slots * 9 && r0 < 0x10000 && r1 < 0x10000 &&
            r0 >= r1 && r1 >= r2 && r2 >= 2 * L.sweeps) {
            A.skew = (unsigned)r0 | ((unsigned)r1 << 16);
            A.skew_cpw = (unsigned)(A.nchunks / 3);
            A.ticket = L.ticket;
        }
    }
    if (n_strip_items > 0 && L.strip_balance) {
        // strip warps go on to an interior item of chunk 0 (top strips) / the next chunk (bottom strips): shorten those chunks
        // by what a strip costs, lengthen the others, so that strip + short chunk = long chunk (see chunk_range<true>)
        const int n = A.nchunks, n_short = (st_top > 0 ? 1 : 0) + (st_bot > 0 ? 1 : 0);
        const int cost = 2 * (max(st_top, st_bot) + 2 * L.sweeps);
        const int c1 = n > 0 ? (rows + n_short * cost + n - 1) / n : 0, c0 = c1 - cost;
        const int min_chunk = 2 * L.sweeps > 8 ? 2 * L.sweeps : 8;
        if (n > n_short && c0 >= min_chunk && c0 < 0x10000 && n_short * c0 + (n - n_short) * c1 >= rows) {
            A.chunk_rows = c1;
            A.skew = (unsigned)c0 | ((unsigned)n_short << 16);
            A.skew_cpw = 0u;
        }
    }
    const int items = max(n_strip_items, A.nbands * A.nchunks);   // strip warps go on to an interior item
    A.steal = (L.steal != nullptr && L.mode == MODE_STRICT && L.staging != 1 && A.nbands * A.nchunks <= L.steal_capacity &&
               A.nbands * A.nchunks > 1) ? L.steal : nullptr;
    dim3 grid((items + WPC - 1) / WPC);
    const size_t smem = (size_t)WPC * (RING_X + RING_R) * 32 * sizeof(float4);
    if (L.rb) {
        // red-black levels: omega rides in the divisor block's spare word; the magnitude bound that keeps every numerator
        // in the exact division's range grows per level by |1 - omega| + omega * F / beta instead of F / beta
        if (A.strips != nullptr || L.staging == 1 || L.zero_guess) return cudaErrorNotSupported;
        A.steal = nullptr;
        A.div.pad = L.omega;
        const double F = 1.0 + 4.0 * fabs((double)L.alpha), om = (double)L.omega;
        double gl = (fabs(1.0 - om) + om * F / fabs((double)L.beta)) * 1.000001;
        if (gl < 1.0) gl = 1.0;
        double hi = (double)SF_DIV_HI / (F * 1.01) / (1.0 + om);    // (1 + omega): the relaxation step's own intermediate
        for (int t = 0; t < L.sweeps; ++t) hi /= gl;
        A.hi_in = (float)hi;
        switch (L.mode) {
            case MODE_PRESSURE: return launch_stream_rb<MODE_PRESSURE>(L.sweeps, A, grid, smem, st);
            case MODE_STRICT: return launch_stream_rb<MODE_STRICT>(L.sweeps, A, grid, smem, st);
            case MODE_IEEE: return launch_stream_rb<MODE_IEEE>(L.sweeps, A, grid, smem, st);
            default: return cudaErrorNotSupported;
        }
    }
    switch (L.mode) {
        case MODE_PRESSURE: return launch_stream_mode<MODE_PRESSURE>(L.sweeps, A, grid, smem, L.staging == 1, st);
        case MODE_FAST: return launch_stream_mode<MODE_FAST>(L.sweeps, A, grid, smem, L.staging == 1, st);
        case MODE_STRICT: return launch_stream_mode<MODE_STRICT>(L.sweeps, A, grid, smem, L.staging == 1, st);
        default: return launch_stream_mode<MODE_IEEE>(L.sweeps, A, grid, smem, L.staging == 1, st);
    }
}

cudaError_t launch_jacobi_generic(const Geom &g, const JacobiLaunch &L, cudaStream_t st)
{
    if (L.sweeps != 1) return cudaErrorInvalidValue;
    const int a_lo = max(L.out_lo, 1), a_hi = min(L.out_hi, g.N + 1);
    if (a_hi <= a_lo) return cudaSuccess;
    dim3 block(64, 4), grid((g.N + 63) / 64, (a_hi - a_lo + 3) / 4);
    const float sx = (L.b == 1) ? -1.0f : 1.0f, sy = (L.b == 2) ? -1.0f : 1.0f;
    const DivConst dc = make_div_const(L.beta);
    const int wt = (L.out_lo == 0), wb = (L.out_hi == g.G);
    switch (L.mode) {
        case MODE_PRESSURE:
            jacobi_generic_kernel<MODE_PRESSURE><<<grid, block, 0, st>>>(L.xin, L.rhs, L.xout, g, a_lo, a_hi, wt, wb, L.alpha, dc, sx, sy);
            break;
        case MODE_FAST:
            jacobi_generic_kernel<MODE_FAST><<<grid, block, 0, st>>>(L.xin, L.rhs, L.xout, g, a_lo, a_hi, wt, wb, L.alpha, dc, sx, sy);
            break;
        default:   // STRICT and IEEE: the fallback kernel is not a hot path, use the plain IEEE division
            jacobi_generic_kernel<MODE_IEEE><<<grid, block, 0, st>>>(L.xin, L.rhs, L.xout, g, a_lo, a_hi, wt, wb, L.alpha, dc, sx, sy);
    }
    return cudaGetLastError();
}

}  // namespace sf

#ifdef SF_WARP_TIMES
extern "C" int sf_debug_warp_times(void *dev_buffer, unsigned int capacity_records)
{
    sf::WarpTimeRec *p = static_cast<sf::WarpTimeRec *>(dev_buffer);
    unsigned int zero = 0;
    if (cudaMemcpyToSymbol(sf::g_warp_times, &p, sizeof(p)) != cudaSuccess) return -1;
    if (cudaMemcpyToSymbol(sf::g_warp_times_cap, &capacity_records, sizeof(unsigned int)) != cudaSuccess) return -1;
    if (cudaMemcpyToSymbol(sf::g_warp_times_n, &zero, sizeof(unsigned int)) != cudaSuccess) return -1;
    return 0;
}
extern "C" int sf_debug_warp_times_count(unsigned int *n)
{
    return cudaMemcpyFromSymbol(n, sf::g_warp_times_n, sizeof(unsigned int)) == cudaSuccess ? 0 : -1;
}
#endif
