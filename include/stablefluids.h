/*
 * stablefluids.h -- C ABI of the B200-native stable-fluids time step.
 *
 * Drop-in boundary for the solver entry points of ArbiterMob/FluidSimulationCuda.  The reference
 * has no library or FFI: its boundary is the C function-call surface inside each single-file
 * program.  Every entry point below names the reference function it replaces
 * (paths relative to the reference's project/ directory; "seq" = sequential/FluidSequential.c,
 * "gpu" = naivePar/FluidParallelBlockPerElement-Naive.cu).
 *
 * Conventions kept from the reference:
 *   - a field is a flat float array of (N+2)*(N+2) cells, index  col + row*(N+2)  (seq:24,95),
 *     interior 1..N, one-cell wall ring;
 *   - results land in the FIRST-named buffers; the *0 / *_prev buffers are clobbered scratch with
 *     the reference's post-conditions (vel_step: u0 = last pressure, v0 = last divergence;
 *     dens_step: x0 = diffused density) -- seq:176-241;
 *   - the caller owns every field; the context owns scratch, streams and captured graphs.
 * What the reference hard-codes and this ABI takes as run-time arguments (north star): grid size N
 * (seq:6), dt (seq:7), viscosity / diffusion (seq:8-9) and the Jacobi iteration count (seq:91,
 * literal 40).  Any iteration count >= 1 works (the reference is only correct for even counts).
 *
 * All field pointers are DEVICE pointers unless the function name ends in _host.  Calls are
 * asynchronous with respect to the host and ordered on the context's stream; a context is not
 * thread-safe, distinct contexts are independent.  Every function returns SF_OK (0) or a negative
 * sf_status; sf_last_error_string() describes the last failure.  Nothing here ever calls exit()
 * (the reference's CHECK macro does, gpu:26-35).
 *
 * There is NO CPU fallback: without a CUDA device every compute entry point fails with
 * SF_ERR_CUDA.
 */
#ifndef STABLEFLUIDS_H
#define STABLEFLUIDS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sf_context sf_context;

typedef enum sf_status {
    SF_OK = 0,
    SF_ERR_INVALID = -1,     /* bad argument (N < 1, null pointer, iters < 1, unknown option ...) */
    SF_ERR_CUDA = -2,        /* CUDA runtime error; text in sf_last_error_string */
    SF_ERR_NOMEM = -3,
    SF_ERR_UNSUPPORTED = -4
} sf_status;

/* Options for sf_set_option. */
enum {
    /* 0 (default) = STRICT: every operation rounded exactly like the reference's sequential
     * build (no FMA contraction, IEEE division): results are BIT-IDENTICAL to seq.
     * 1 = FAST: x0 + alpha*sum contracted to one FMA and the division replaced by a multiply with
     * 1/beta; rel-L2 <= 1e-5 per field per step against seq (tests/test_parity_gpu.py). */
    SF_OPT_ARITHMETIC = 1,
    /* Jacobi sweeps fused per kernel launch (temporal blocking depth), 1..8; 0 = automatic. */
    SF_OPT_SWEEPS_PER_LAUNCH = 2,
    /* 1 (default) = replay sf_vel_step/sf_dens_step/sf_step from a captured CUDA graph when the
     * same arguments recur; 0 = launch kernels directly. */
    SF_OPT_USE_GRAPH = 3,
    /* 1 = force the generic one-sweep-per-launch kernels even when the grid width allows the
     * streaming kernels (testing aid). */
    SF_OPT_FORCE_GENERIC = 4,
    /* interior rows per streaming chunk (0 = automatic). */
    SF_OPT_CHUNK_ROWS = 5,
    /* how the Jacobi kernel stages rows global -> shared: 0 (default) = cp.async per lane
     * (LDGSTS.128), 1 = one bulk copy per warp row through the TMA unit (cp.async.bulk + mbarrier). */
    SF_OPT_STAGING = 6,
    /* > 0 (default 30): the warps of a STRICT Jacobi launch on a scalar field (dens_step's solve,
     * sf_diffuse with b = 0) balance their load by row-level work stealing: a warp that has finished
     * halves the largest remaining row range it finds.  Evens out the guarded binary64 ticks that the
     * decaying front of a density field needs.  The value is the smallest remaining share of a chunk,
     * in percent, that is worth halving (1..100); 0 = off.  Results are unchanged. */
    SF_OPT_WORK_STEALING = 7,
    /* read-only (sf_get_option): row ranges taken over by another warp so far; synchronises. */
    SF_OPT_STEAL_COUNT = 8,
    /* which STRICT solves use work stealing: 0 (default) = scalar fields only (see above), 1 = every
     * STRICT lin_solve (the velocity solves of vel_step too).  Results are unchanged. */
    SF_OPT_STEAL_SCOPE = 9,
    /* 1: a lin_solve that starts from the implicit zero guess (the pressure solves of
     * sf_project) may take an odd number of launches (its first launch does not read x, so it may
     * write x); 2 (default): the same, and those solves fuse up to 8 sweeps per launch instead of 7
     * (40 iterations = 5 launches instead of 6; measured 7.31 -> 7.15 ms per step at G=8192);
     * 0 = the even-count plan of every other solve.  Results are unchanged. */
    SF_OPT_PRESSURE_PLAN = 10,
    /* which scheme lin_solve (sf_diffuse, sf_project and the solves inside the step functions) runs.
     * SF_SOLVER_JACOBI (default) = the reference's double-buffered Jacobi (seq:85-104): results
     * bit-identical to the reference.  SF_SOLVER_RBGS = red-black Gauss-Seidel, in place: per iteration
     * a half-sweep over the cells with (row + col) even, one over the odd cells, then set_bnd(b); same
     * cell formula and operand order.  NOT the reference's scheme -- it converges about twice as fast per
     * iteration, so results differ from the reference by design; they are bit-identical to a CPU build of
     * the same scheme (tests/test_zzz_solvers_gpu.py).  Full-grid contexts only. */
    SF_OPT_SOLVER = 11,
    /* SF_SOLVER_RBGS only: over-relaxation factor omega in 1/1000 (1..1999; default 1000 = plain
     * Gauss-Seidel).  omega != 1: x = x + omega*(gs - x), three separately rounded operations. */
    SF_OPT_SOR_OMEGA_MILLI = 12,
    /* SF_SOLVER_RBGS only: 0 (default) = one kernel launch per half-sweep; 1 = the red-black sweeps run on the
     * temporally blocked streaming pipeline of the Jacobi kernels (three iterations per launch; grid widths with
     * (N+2) % 4 == 0, other widths keep the one-launch-per-half-sweep path).  Results are unchanged. */
    SF_OPT_RBGS_BLOCKED = 13,
    /* 1 (default) = inside dens_step / vel_step / step the add_source pass (seq:78-82) is fused into the first temporally
     * blocked launch of the lin_solve that follows it (seq:177-182, :193-210): the right-hand side x + dt*s is formed as
     * the rows stream in and kept in a context-owned field, one pass over x and s less per solve.  The fields the
     * reference's step functions leave behind (u, v, dens and the clobbered *_prev buffers) are unchanged, bit for bit.
     * 0 = a separate add_source kernel, as sf_add_source + sf_diffuse would run it.  Full-grid contexts and connected
     * peer slabs (there the boundary-strip warps of that first launch also form the right-hand side's ghost rows). */
    SF_OPT_FUSE_SOURCES = 14,
    /* Temporally blocked Jacobi launches are one full wave of warps, each with one chunk of rows, three CTAs per SM at the
     * default depth.  A warp scheduler favours its oldest warp, so with equal chunks the warps of the CTA an SM received
     * first finish long before those of the third, and the tail of the launch runs at a fraction of the machine.
     * value = p0 * 1000 + p1: CTAs draw their work item in the order they start, and the chunks of the first / second third
     * of the items get p0 / p1 percent of the mean chunk's rows (the last third gets the rest), e.g. 135106.  Results are
     * unchanged (temporal blocking does not depend on where the chunks are cut).  0 = equal chunks in blockIdx order. */
    SF_OPT_WAVE_SKEW = 15,
    /* advect (seq:107-141).  A CTA owns a tile of 16 or 32 rows x 128 columns, traces its cells back, and when the bounding
     * box of the traces fits (<= 160 columns, <= 8 rows per copy below, inside the rows the context stores) the box of the
     * source field(s) is fetched into shared memory by the TMA unit (2-D tensor copies, cp.async.bulk.tensor) and the four
     * bilinear corners are read from there; tiles whose box does not fit gather from global memory as with 0.
     * 1 (default) = automatic: 32-row tiles with up to 8 copies of 8 rows for one field, 16-row tiles with up to 5 copies per
     * field for the u, v pair; when a step is captured into a graph the tile
     * counters of the direct run before it decide whether the captured step keeps the tiles (most fitted) or runs as with 0;
     * connected slabs run as with 0.  2..8 = 32-row tiles, at most that many 8-row copies per field; 12..18 = 16-row tiles,
     * (value - 10) copies; both always on.  0 = every cell gathers from global memory.
     * Results are unchanged, bit for bit.  Needs (N+2) % 4 == 0 and N+2 >= 320; other grids always run as with 0. */
    SF_OPT_ADVECT_TILE = 16,
    /* diagnostics (sf_get_option synchronises): tiles of the advect launches on this context's device that were served by
     * the TMA box / that fell back to global gathers, since the last sf_set_option(ctx, SF_OPT_ADVECT_TILE_COUNT, 0) */
    SF_OPT_ADVECT_TILE_COUNT = 17,
    SF_OPT_ADVECT_FALLBACK_COUNT = 18,
    /* 1 (default) = inside sf_step (and sf_run_steps) the three lin_solves that do not depend on each other -- the viscosity
     * solves of u and v (seq:193-210) and the density's diffusion solve (seq:177-182) -- are enqueued on three streams
     * (graph branches), so the launches of one fill the SMs that the tail of another's leaves idle (u beside v, then the density solve
     * beside the projections and the advection of vel_step).  Costs four more scratch fields.  0 = one after the other.  Full-grid contexts with the
     * Jacobi solver; results are unchanged (same kernels, same arguments). */
    SF_OPT_OVERLAP_SOLVES = 19,
    /* Connected peer slabs: the warps that compute a boundary strip go on to an interior work item of the first (top strip)
     * or next (bottom strip) row chunk.  1 (default) = those chunks are shorter than the others by what a strip costs
     * (2 * (strip rows + 2 * sweeps) rows), so every warp of a launch finishes at about the same time; 0 = equal chunks.
     * Results are unchanged (temporal blocking does not depend on where the chunks are cut). */
    SF_OPT_STRIP_BALANCE = 20
};
enum { SF_ARITH_STRICT = 0, SF_ARITH_FAST = 1 };
enum { SF_SOLVER_JACOBI = 0, SF_SOLVER_RBGS = 1 };

/* ---- context -------------------------------------------------------------------------------- */

/* Create a context for an (N+2)^2 grid on CUDA device `device`, with its own stream.
 * Replaces the reference's file-scope N / __constant__ N (seq:6, gpu:11-20, gpu:386-389). */
int sf_create(sf_context **out, int N, int device);
/* Same, but every kernel is enqueued on the caller's CUDA stream (a cudaStream_t passed as
 * void*), e.g. PyTorch's current stream. */
int sf_create_on_stream(sf_context **out, int N, int device, void *cuda_stream);
/* Slab context for domain decomposition along rows: this context owns global
 * rows [row_lo, row_hi) of the (N+2)-row grid and every field passed to it is a LOCAL array of
 * (row_hi - row_lo + 2*halo) rows x (N+2) columns whose first row is global row row_lo - halo.
 * The caller fills the halo rows (neighbour exchange) before each call that reads them; see
 * sf_halo_rows_needed.  row_lo = 0, row_hi = N+2, halo = 0 is the single-GPU layout. */
int sf_create_slab(sf_context **out, int N, int device, void *cuda_stream, int row_lo, int row_hi, int halo);
/* Same, on a non-blocking stream the context creates and owns (see sf_get_stream). */
int sf_create_slab_own_stream(sf_context **out, int N, int device, int row_lo, int row_hi, int halo);
int sf_destroy(sf_context *ctx);
const char *sf_last_error_string(const sf_context *ctx);
int sf_set_option(sf_context *ctx, int option, int value);
int sf_get_option(const sf_context *ctx, int option, int *value);
int sf_synchronize(sf_context *ctx);
/* Re-target the context to another CUDA stream (all later calls are ordered on it).  Captured
 * graphs are dropped, since a graph launch is bound to the stream it is replayed on only. */
int sf_set_stream(sf_context *ctx, void *cuda_stream);
/* The CUDA stream (cudaStream_t as void*) the context orders its work on. */
int sf_get_stream(const sf_context *ctx, void **cuda_stream);
/* Number of kernels this context has launched (graph replays count their kernel nodes). */
int sf_launch_count(const sf_context *ctx, unsigned long long *count);
size_t sf_field_bytes(const sf_context *ctx);      /* bytes of one (local) field */

/* ---- fields (replace malloc/cudaMalloc/cudaMemcpy in the reference's main, gpu:375-384) ------ */
int sf_alloc_field(sf_context *ctx, float **dev_field);
int sf_free_field(sf_context *ctx, float *dev_field);
int sf_upload(sf_context *ctx, float *dev_field, const float *host_field);
int sf_download(sf_context *ctx, float *host_field, const float *dev_field);

/* ---- stage functions ---------------------------------------------------------------------------
 * set_bnd(b, x)                          seq:62-75   gpu:83-104 */
int sf_set_bnd(sf_context *ctx, int b, float *x);
/* add_source(x, s): x += dt*s on all (N+2)^2 cells   seq:78-82   gpu:108-118 */
int sf_add_source(sf_context *ctx, float *x, const float *s, float dt);
/* diffuse(b, x, x0, alpha, beta) = lin_solve: `iters` Jacobi sweeps + set_bnd(b) each
 *                                        seq:85-104  gpu:121-144 + host loop gpu:261-264 */
int sf_diffuse(sf_context *ctx, int b, float *x, const float *x0, float alpha, float beta, int iters);
/* advect(b, d, d0, u, v)                 seq:107-141 gpu:147-196 */
int sf_advect(sf_context *ctx, int b, float *d, const float *d0, const float *u, const float *v, float dt);
/* advect(1, u, u0, u0, v0); advect(2, v, v0, u0, v0) -- the self-advection of vel_step, seq:228-237 -- in one pass sharing
 * the back-trace (what sf_vel_step runs); same bits as the two sf_advect calls.  Full-grid and unconnected slab contexts. */
int sf_advect_velocity(sf_context *ctx, float *u, float *v, const float *u0, const float *v0, float dt);
/* computeDivergenceAndPressure(u, v, p, div)   seq:143-158 gpu:199-225 */
int sf_compute_divergence_and_pressure(sf_context *ctx, const float *u, const float *v, float *p, float *div);
/* lastProject(u, v, p, div)              seq:161-173 gpu:228-252 */
int sf_last_project(sf_context *ctx, float *u, float *v, const float *p, const float *div);
/* The projection triple as the reference sequences it (seq:213-223): divergence, lin_solve(0,
 * p, div, 1, 4), gradient subtract. */
int sf_project(sf_context *ctx, float *u, float *v, float *p, float *div, int iters);

/* ---- step drivers ------------------------------------------------------------------------------
 * dens_step(x, x0, u, v, diff)           seq:176-186 gpu:255-268 */
int sf_dens_step(sf_context *ctx, float *x, float *x0, const float *u, const float *v, float diff, float dt, int iters);
/* vel_step(u, v, u0, v0, visc, z)        seq:189-241 gpu:271-311 (z only indexed timers) */
int sf_vel_step(sf_context *ctx, float *u, float *v, float *u0, float *v0, float visc, float dt, int iters);
/* One iteration of the reference's main loop body (seq:305-306): vel_step then dens_step. */
int sf_step(sf_context *ctx, float *dens, float *dens_prev, float *u, float *u_prev, float *v, float *v_prev,
            float visc, float diff, float dt, int iters);
/* Same loop body for a caller whose six fields live in HOST memory (the reference's CPU path,
 * seq:277-282): uploads the six fields, runs the step on the device, downloads dens, u, v (and,
 * if download_scratch != 0, the three clobbered *_prev fields) and returns when the host buffers
 * are valid.  Pinned host memory makes the copies overlap the compute.
 * On a connected peer slab (sf_slab_connect_*) the call is collective -- one caller per slab, all at the same time -- the
 * host arrays hold the slab's OWNED rows only ((row_hi - row_lo) x (N+2) floats) and the device fields are the first six
 * fields of the slab's arena in the order of the arguments (sf_slab_field 0..5). */
int sf_step_host(sf_context *ctx, float *dens, float *dens_prev, float *u, float *u_prev, float *v, float *v_prev,
                 float visc, float diff, float dt, int iters, int download_scratch);

/* ---- device-resident driver loop -------------------------------------------------------------------
 * The reference's main loop (seq:289-312; gpu:400-439 / optPar LOOPUNROLLED-Interleaved2.cu:680-724)
 * zeroes the three source fields on the HOST and re-uploads them before every step -- about a third
 * of its per-step time at 8192^2.  sf_run_steps keeps the whole loop on the device: `steps` iterations
 * of { source schedule; vel_step; dens_step } are enqueued (graph replays) without any host round
 * trip in between.  Source schedule for the *_prev fields, which every step clobbers:
 *   SF_SOURCES_REFERENCE  step 0 uses the *_prev fields as passed in; later steps see zeros
 *                         (the reference's rule, seq:298-302);
 *   SF_SOURCES_SYNTHETIC  every step k regenerates them from the counter-based hash with seed + k
 *                         (same distributions as seq:244-271; see sf_init_sources);
 *   SF_SOURCES_FIELDS     every step copies src_dens / src_u / src_v (device fields the caller owns,
 *                         never written) into dens_prev / u_prev / v_prev.
 * Works on full-grid contexts and (collectively) on connected peer slabs. */
enum { SF_SOURCES_REFERENCE = 0, SF_SOURCES_SYNTHETIC = 1, SF_SOURCES_FIELDS = 2 };
int sf_run_steps(sf_context *ctx, float *dens, float *dens_prev, float *u, float *u_prev, float *v, float *v_prev,
                 float visc, float diff, float dt, int iters, int steps, int source_mode, uint64_t seed,
                 const float *src_dens, const float *src_u, const float *src_v);
/* Binary dump of one (local) device field -- the replacement for the reference's printStateGrid
 * (seq:32-52): a 32-byte header { "SFLD", int32 version = 1, N, row_lo, row_hi, halo, reserved }
 * followed by the owned rows as raw little-endian float32, row-major.  Synchronises. */
int sf_dump_field(sf_context *ctx, const float *dev_field, const char *path);

/* ---- synthetic initial conditions (seq:244-271 value distributions, counter-based) ---------- */
int sf_init_synthetic(sf_context *ctx, uint64_t seed, float *dens, float *dens_prev, float *u, float *u_prev,
                      float *v, float *v_prev);
/* Only the three source fields (what the reference's loop refreshes every step, seq:298-302). */
int sf_init_sources(sf_context *ctx, uint64_t seed, float *dens_prev, float *u_prev, float *v_prev);

/* ---- diagnostics (warp-shuffle reductions) ------------------------------------------------- */
/* max |x| over the owned cells; result written to *host_out after an internal synchronize. */
int sf_reduce_max_abs(sf_context *ctx, const float *x, float *host_out);
/* Asynchronous form for pipelines that must not stall: *dev_out = max(*dev_out, max |x|) with
 * dev_out a DEVICE float the caller has initialised (e.g. to 0); no synchronisation. */
int sf_reduce_max_abs_async(sf_context *ctx, const float *x, float *dev_out);
/* || x0 - (beta*x - alpha*sum_nb(x)) ||_2 over the interior: the lin_solve residual. */
int sf_residual_l2(sf_context *ctx, const float *x, const float *x0, float alpha, float beta, double *host_out);

/* STRICT arithmetic divides by beta with a 3-instruction exact FMA sequence once that sequence has
 * been checked on the device against the IEEE division for ALL 2^32 numerators with this beta
 * (about 4 ms, cached per process); otherwise it uses the IEEE division.  This call runs (or looks
 * up) that check: *exact = 1 when the short sequence is in use for `beta`.  Results are
 * bit-identical either way; only the speed differs. */
int sf_division_check(sf_context *ctx, float beta, int *exact);

/* ---- slab support --------------------------------------------------------------------------- */
/* Halo rows each call reads beyond the owned rows: lin_solve reads `sweeps_per_launch` rows per
 * launch, divergence / gradient 1 row.  *rows = temporal-blocking depth currently configured. */
int sf_halo_rows_needed(const sf_context *ctx, int *rows);
/* One temporally blocked launch of the lin_solve: `sweeps` (1..8) Jacobi sweeps from xin into
 * xout (different buffers), restricted to owned rows [out_lo, out_hi) (global row numbers; pass
 * -1, -1 for all owned rows).  xin must hold valid rows [out_lo - sweeps, out_hi + sweeps).
 * Building block for halo-exchange overlap: boundary strips first, interior while halos fly. */
int sf_jacobi_launch(sf_context *ctx, int b, float *xout, const float *xin, const float *x0, float alpha, float beta,
                     int sweeps, int out_lo, int out_hi);

/* ---- peer-memory slabs: the multi-GPU path (the reference is single-GPU) -------
 * A slab context whose neighbours' memory is mapped (NVLink / NVSwitch peer access) runs the SAME entry
 * points -- sf_step, sf_vel_step, sf_dens_step, sf_diffuse, sf_project, sf_advect -- and performs the
 * halo traffic itself, on the device: boundary strips of every temporally blocked Jacobi launch store
 * their rows straight into the neighbour's ghost rows, advect reads out-of-slab rows through the peer
 * mapping, and GPUs order themselves with one-warp neighbour-barrier kernels (no host thread, no NCCL
 * call and no stream synchronisation inside a step; a whole step is one CUDA-graph replay per GPU).
 * These calls are COLLECTIVE over the connected slabs: every slab must issue the same sequence.
 * Results are bit-identical to the single-GPU path for any partition.
 *
 * Set-up, per slab context (created with sf_create_slab, halo >= sweeps per launch, (N+2) % 4 == 0):
 *   sf_slab_arena_create(ctx, nfields)      one device allocation holding nfields fields (zeroed), the
 *                                           lin_solve scratch and the synchronisation words;
 *   sf_slab_field(ctx, k, &ptr)             field k of the arena (pass these to the entry points);
 *   sf_slab_ipc_handle / sf_slab_connect_ipc    one process per GPU: exchange the 64-byte CUDA IPC
 *                                           handle of the arena with ranks r-1 / r+1 (any transport)
 *   sf_slab_connect_local                   several slabs in one process (peer access is enabled). */
enum { SF_SLAB_UP = 0, SF_SLAB_DOWN = 1 };
enum { SF_SLAB_ERR_TIMEOUT = 1, SF_SLAB_ERR_REACH = 2 };
int sf_slab_arena_create(sf_context *ctx, int nfields);
int sf_slab_field(sf_context *ctx, int k, float **dev_field);
int sf_slab_ipc_handle(sf_context *ctx, void *handle64);
/* dir = SF_SLAB_UP: the slab owning the rows just above (smaller row numbers); the neighbour owns
 * global rows [nbr_row_lo, nbr_row_hi) and was created with the same N, halo and nfields. */
int sf_slab_connect_ipc(sf_context *ctx, int dir, const void *handle64, int nbr_row_lo, int nbr_row_hi);
int sf_slab_connect_local(sf_context *ctx, int dir, sf_context *neighbour);
/* A neighbour barrier that waits longer than this sets SF_SLAB_ERR_TIMEOUT and stops waiting
 * (default 20 s): a missing neighbour is an error report, never a hung GPU. */
int sf_slab_set_timeout_ms(sf_context *ctx, int milliseconds);
/* Synchronises the context and returns the sticky device-side error bits: SF_SLAB_ERR_TIMEOUT, or
 * SF_SLAB_ERR_REACH when an advection back-trace left the neighbouring slab (|dt*N*v| larger than a
 * whole slab: results of that step are invalid). */
int sf_slab_status(sf_context *ctx, unsigned int *error_bits);

#ifdef __cplusplus
}
#endif
#endif /* STABLEFLUIDS_H */
