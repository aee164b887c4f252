// C ABI (include/stablefluids.h) and the host-side step drivers.
//
// The drivers mirror the reference's sequencing (FluidSequential.c:176-186 dens_step, :189-241
// vel_step) with the local pointer SWAPs written out, plan each lin_solve as a short list of
// temporally blocked launches, and replay whole steps from captured CUDA graphs.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "sf_internal.h"

using namespace sf;

namespace sf {

int ensure_scratch(sf_context *c)
{
    if (!c->scratch) SF_CUDA(c, cudaMalloc(&c->scratch, field_cells(c) * sizeof(float)));
    // (allocated here, on the first direct run of a step, never inside a stream capture)
    // (full grids and arena slabs: the right-hand side of a solve with the add_source fused into its first launch; it is only
    // ever written and read by its own GPU, so it need not be part of the arena the neighbours map)
    if (!c->scratch2 && (is_full_grid(c) || c->link.base)) SF_CUDA(c, cudaMalloc(&c->scratch2, field_cells(c) * sizeof(float)));
    if (!c->steal) {
        c->steal_capacity = 16384;
        const size_t bytes = sizeof(StealCtl) + (size_t)c->steal_capacity * sizeof(StealSlot);
        SF_CUDA(c, cudaMalloc(&c->steal, bytes));
        {   // every slot starts out as "nothing to take here" (pos beyond any row), tag 0 = no launch
            std::vector<char> init(bytes, 0);
            StealCtl *h = reinterpret_cast<StealCtl *>(init.data());
            h->min_pct = c->steal_opt;
            for (int k = 0; k < c->steal_capacity; ++k) h->slots[k].pos = 0x3fffffff;
            SF_CUDA(c, cudaMemcpyAsync(c->steal, init.data(), bytes, cudaMemcpyHostToDevice, c->stream));
            SF_CUDA(c, cudaStreamSynchronize(c->stream));   // the staging vector goes out of scope
        }
    }
    if (!c->red_f) {
        SF_CUDA(c, cudaMalloc(&c->red_f, sizeof(float)));
        SF_CUDA(c, cudaMalloc(&c->red_d, sizeof(double)));
    }
    if (!c->ticket) {
        SF_CUDA(c, cudaMalloc(&c->ticket, 64));
        SF_CUDA(c, cudaMemsetAsync(c->ticket, 0, 64, c->stream));
    }
    return SF_OK;
}

// can this context run the independent solves of a step side by side?  (the work-stealing block and the red-black solver's
// in-place sweeps have one user at a time; slabs sequence their exchanges on one stream)
bool overlap_ok(const sf_context *c)
{
    return c->overlap && is_full_grid(c) && c->solver == SF_SOLVER_JACOBI && c->steal_scope == 0 && stream_kernels_ok(c);
}

// lanes are created on the first direct run of a step, never inside a stream capture
void release_lanes(sf_context *c)
{
    for (auto &L : c->lanes) {
        if (L.stream) cudaStreamDestroy(L.stream);
        if (L.scratch) cudaFree(L.scratch);
        if (L.scratch2) cudaFree(L.scratch2);
        if (L.ticket) cudaFree(L.ticket);
        if (L.fork) cudaEventDestroy(L.fork);
        if (L.join) cudaEventDestroy(L.join);
        L = sf_context::SolveLane();
    }
}

int ensure_lanes(sf_context *c)
{
    if (!overlap_ok(c) || c->lanes[1].join) return SF_OK;
    bool ok = true;
    for (int k = 0; k < 2 && ok; ++k) {
        sf_context::SolveLane &L = c->lanes[k];
        ok = ok && cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking) == cudaSuccess;
        ok = ok && cudaMalloc(&L.scratch, field_cells(c) * sizeof(float)) == cudaSuccess;
        ok = ok && cudaMalloc(&L.scratch2, field_cells(c) * sizeof(float)) == cudaSuccess;
        ok = ok && cudaMalloc(&L.ticket, 64) == cudaSuccess && cudaMemset(L.ticket, 0, 64) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&L.fork, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&L.join, cudaEventDisableTiming) == cudaSuccess;
    }
    if (!ok) {
        // not enough device memory for four more fields (or no more streams): the solves run one after the other, as with
        // SF_OPT_OVERLAP_SOLVES = 0 -- same results, and the call that got here still succeeds
        (void)cudaGetLastError();
        release_lanes(c);
        c->overlap = 0;
    }
    return SF_OK;
}

// SF_OPT_ADVECT_TILE = 1 is automatic: advect_tile_kernel counts the tiles whose traces fitted the TMA box and those that fell
// back to gathers.  Where most tiles fall back (a velocity field whose back-traces scatter further than the box: the same
// initial condition on a much finer grid, dt0 = dt * N) the plain gather kernel is the better choice, since it keeps twice the
// warps in flight.  The counters are looked at when a step is about to be captured into a graph -- the one place where the
// library synchronises anyway -- and the choice is frozen into that graph.
int refresh_advect_policy(sf_context *c)
{
    if (c->advect_tile != 1 || !c->tile_stats || c->link.base != nullptr) return SF_OK;
    unsigned int now[2] = {0u, 0u};
    SF_CUDA(c, cudaStreamSynchronize(c->stream));
    SF_CUDA(c, cudaMemcpy(now, c->tile_stats, sizeof(now), cudaMemcpyDeviceToHost));
    const unsigned int tma = now[0] - c->tile_seen[0], fallback = now[1] - c->tile_seen[1];
    c->tile_seen[0] = now[0]; c->tile_seen[1] = now[1];
    if (tma + fallback > 0u) c->advect_tile_live = (fallback <= tma);
    return SF_OK;
}

int arith_mode(const sf_context *c, float alpha, float beta)
{
    if (alpha == 1.0f && beta == 4.0f) return MODE_PRESSURE;   // exact identity, see jacobi_cell
    if (c->arith == SF_ARITH_FAST) return MODE_FAST;
    // STRICT: the 3-instruction exact division is used only for a beta that has been checked
    // against __fdiv_rn over all 2^32 numerators (cached); the check cannot run inside a capture.
    bool can_run = !c->capturing;
    if (can_run) {
        cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(c->work, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) can_run = false;
    }
    return division_validated(beta, can_run, c->work) ? MODE_STRICT : MODE_IEEE;
}

int default_sweeps(const sf_context *c)
{
    if (c->sweeps_opt >= 1 && c->sweeps_opt <= 8) return c->sweeps_opt;
    // 7, not 8: at depth 8 the strict kernel spills (168-register cap for 3 CTAs/SM) and the pressure kernel
    // gains nothing (measured, K=200 at G=16384: strict 32.7 ms vs 34.1 ms, pressure 20.9 vs 21.1)
    return 7;
}

// Split `iters` sweeps into launches of at most T sweeps.  The lin_solve ping-pongs between x and
// the scratch field, so an EVEN number of launches leaves the result in x without a copy.
std::vector<int> plan_launches(int iters, int T, bool odd_ok)
{
    int L = (iters + T - 1) / T;
    if (!odd_ok && (L & 1) && L + 1 <= iters) ++L;
    std::vector<int> plan(L, iters / L);
    for (int k = 0; k < iters % L; ++k) ++plan[k];
    return plan;
}

int one_jacobi_launch(sf_context *c, cudaStream_t st, int b, float *xout, const float *xin, const float *x0, float alpha,
                      float beta, int sweeps, int out_lo, int out_hi, int zero_guess, int strip_rows, float *rhs_out, float src_dt)
{
    JacobiLaunch L;
    L.xin = xin; L.rhs = x0; L.xout = xout;
    L.rhs_out = rhs_out; L.src_dt = src_dt;
    L.wave_skew = c->wave_skew; L.ticket = c->ticket; L.strip_balance = c->strip_balance;
    L.alpha = alpha; L.beta = beta; L.b = b; L.sweeps = sweeps;
    L.mode = arith_mode(c, alpha, beta);
    L.out_lo = out_lo; L.out_hi = out_hi;
    L.chunk_rows = c->chunk_rows;
    L.zero_guess = zero_guess;
    L.staging = c->staging;
    if (c->steal_opt && (c->steal_now || c->steal_scope == 1) && c->steal) { L.steal = c->steal; L.steal_capacity = c->steal_capacity; }
    if (strip_rows > 0) {
        L.strips = slab_strip_args(c, xout, strip_rows);
        SF_REQUIRE(c, L.strips != nullptr, "peer slab: output field is not an arena field or strip too high");
        L.strip_rows[0] = c->link.nbr[0].present ? strip_rows : 0;
        L.strip_rows[1] = c->link.nbr[1].present ? strip_rows : 0;
    }
    {   // rows the launch reads must be stored locally (strip launches read the ghost rows the neighbour pushed: same bound)
        const int rd_lo = out_lo - sweeps > 0 ? out_lo - sweeps : 0;
        const int rd_hi = out_hi - 1 + sweeps < c->g.G - 1 ? out_hi - 1 + sweeps : c->g.G - 1;
        SF_REQUIRE(c, rd_lo >= c->g.row_base && rd_hi < c->g.row_base + c->g.rows, "Jacobi launch: fewer ghost rows than fused sweeps");
    }
    if (stream_kernels_ok(c)) {
        SF_CUDA(c, launch_jacobi_stream(c->g, L, c->sm_count, st));
    } else {
        SF_REQUIRE(c, sweeps == 1, "generic Jacobi kernel does one sweep per launch");
        SF_REQUIRE(c, rhs_out == nullptr, "generic Jacobi kernel has no fused add_source");
        SF_REQUIRE(c, strip_rows == 0, "generic Jacobi kernel has no fused strip exchange");
        if (zero_guess) {
            // generic kernel always reads xin
            SF_CUDA(c, cudaMemsetAsync(const_cast<float *>(xin), 0, field_cells(c) * sizeof(float), st));
        }
        SF_CUDA(c, launch_jacobi_generic(c->g, L, st));
    }
    ++c->launches;
    return SF_OK;
}

// SF_SOLVER_RBGS (opt-in, sf_solvers.cu): in place on x, iters x { red half-sweep, black half-sweep, set_bnd(b) }
int rbgs_solve(sf_context *c, int b, float *x, const float *x0, float alpha, float beta, int iters)
{
    if (!is_full_grid(c)) return fail(c, SF_ERR_UNSUPPORTED, "SF_SOLVER_RBGS: full-grid contexts only (no slabs yet)");
    const int mode = arith_mode(c, alpha, beta);
    const float omega = (float)c->omega_milli / 1000.0f;
    if (c->rbgs_blocked && stream_kernels_ok(c) && mode != MODE_FAST) {
        // SF_OPT_RBGS_BLOCKED: the same scheme on the streaming pipeline, three iterations (six levels) per launch,
        // ping-pong between x and the scratch field like the Jacobi path
        int rc = ensure_scratch(c);
        if (rc) return rc;
        float *cur = x, *nxt = c->scratch;
        for (int done = 0; done < iters;) {
            const int k = iters - done < 3 ? iters - done : 3;
            JacobiLaunch L;
            L.xin = cur; L.rhs = x0; L.xout = nxt;
            L.alpha = alpha; L.beta = beta; L.b = b; L.sweeps = 2 * k; L.mode = mode;
            L.out_lo = c->g.own_lo; L.out_hi = c->g.own_hi;
            L.chunk_rows = c->chunk_rows; L.zero_guess = 0; L.staging = 0;
            L.rb = 1; L.omega = omega;
            SF_CUDA(c, launch_jacobi_stream(c->g, L, c->sm_count, c->work));
            ++c->launches;
            float *t = cur; cur = nxt; nxt = t;
            done += k;
        }
        if (cur != x) SF_CUDA(c, cudaMemcpyAsync(x, cur, field_cells(c) * sizeof(float), cudaMemcpyDeviceToDevice, c->work));
        return SF_OK;
    }
    for (int k = 0; k < iters; ++k) {
        SF_CUDA(c, launch_rbgs_half_sweep(c->g, x, x0, 0, mode, alpha, beta, omega, c->work));
        SF_CUDA(c, launch_rbgs_half_sweep(c->g, x, x0, 1, mode, alpha, beta, omega, c->work));
        SF_CUDA(c, launch_set_bnd(c->g, b, x, c->work));
        c->launches += 3;
    }
    return SF_OK;
}

// lin_solve: result always ends in x.
int lin_solve(sf_context *c, int b, float *x, const float *x0, float alpha, float beta, int iters, int zero_guess)
{
    if (c->solver == SF_SOLVER_RBGS) return rbgs_solve(c, b, x, x0, alpha, beta, iters);   // (never called with zero_guess)
    if (is_linked_slab(c)) return slab_lin_solve(c, b, x, x0, alpha, beta, iters, zero_guess);
    int rc = ensure_scratch(c);
    if (rc) return rc;
    const bool stream_ok = jacobi_stream_supported(c->g) && !c->force_generic;
    // A solve from the implicit zero guess (the pressure solves of project) never reads x in its first launch, so
    // that launch may write x itself and the ping-pong may take an ODD number of launches: K = 20 runs as 3
    // launches (7,7,6) instead of 4 (5,5,5,5), K = 200 as 29 instead of 30.  (K = 40 stays at 6 launches: 5 launches
    // of 8 sweeps were measured SLOWER, 9.42 vs 8.87 ms per step at G=8192 -- at depth 8 the pressure kernel
    // spills under its 128-register cap.)
    const bool odd_ok = stream_ok && zero_guess && c->pressure_plan != 0;
    int depth = stream_ok ? default_sweeps(c) : 1;
    // SF_OPT_PRESSURE_PLAN = 2: the pressure kernel (alpha = 1, beta = 4) is built for 3 CTAs per SM since round 2 and no longer
    // spills at depth 8 (154 registers), so K = 40 can run as 5 launches of 8 sweeps instead of 6 of 7,7,7,7,6,6
    if (stream_ok && c->pressure_plan == 2 && c->sweeps_opt == 0 && alpha == 1.0f && beta == 4.0f) depth = 8;
    const std::vector<int> plan = plan_launches(iters, depth, odd_ok);
    float *cur = x, *nxt = c->scratch;
    if (odd_ok && (plan.size() & 1)) { cur = c->scratch; nxt = x; }   // `cur` is not read by the first launch
    for (size_t k = 0; k < plan.size(); ++k) {
        rc = one_jacobi_launch(c, c->work, b, nxt, cur, x0, alpha, beta, plan[k], c->g.own_lo, c->g.own_hi, zero_guess && k == 0);
        if (rc) return rc;
        float *t = cur; cur = nxt; nxt = t;
    }
    if (cur != x) SF_CUDA(c, cudaMemcpyAsync(x, cur, field_cells(c) * sizeof(float), cudaMemcpyDeviceToDevice, c->work));
    return SF_OK;
}

// add_source followed by the lin_solve that consumes it, as dens_step / vel_step run them (FluidSequential.c:177-182,
// :193-204, :197-210): x0 += dt * x; SWAP; solve with right-hand side x0 from the initial guess x, result in x.
// Fused form (SF_OPT_FUSE_SOURCES): the first launch forms x0 + dt * x itself while the rows stream in and stores it to the
// context's second scratch field, which the later launches read as their right-hand side; x0 is left as it was -- every
// caller below overwrites it before anything reads it again (vel_step: p / div of the projection; dens_step: advect's output).
int source_lin_solve(sf_context *c, int b, float *x, float *x0, float dt, float alpha, float beta, int iters)
{
    bool fuse = c->fuse_sources && is_full_grid(c) && c->solver == SF_SOLVER_JACOBI && stream_kernels_ok(c) && c->staging == 0;
    std::vector<int> plan;
    if (fuse) {
        plan = plan_launches(iters, default_sweeps(c), false);
        fuse = jacobi_src_fusion_built(plan[0], arith_mode(c, alpha, beta));
    }
    if (!fuse) {
        float *xs[1] = {x0};
        const float *ss[1] = {x};
        SF_CUDA(c, launch_add_source(c->g, 1, xs, ss, dt, c->work));
        ++c->launches;
        return lin_solve(c, b, x, x0, alpha, beta, iters, 0);
    }
    int rc = ensure_scratch(c);
    if (rc) return rc;
    float *cur = x, *nxt = c->scratch;
    for (size_t k = 0; k < plan.size(); ++k) {
        rc = one_jacobi_launch(c, c->work, b, nxt, cur, k == 0 ? x0 : c->scratch2, alpha, beta, plan[k], c->g.own_lo, c->g.own_hi, 0, 0,
                               k == 0 ? c->scratch2 : nullptr, dt);
        if (rc) return rc;
        float *t = cur; cur = nxt; nxt = t;
    }
    if (cur != x) SF_CUDA(c, cudaMemcpyAsync(x, cur, field_cells(c) * sizeof(float), cudaMemcpyDeviceToDevice, c->work));
    return SF_OK;
}

int check_multi_launch_ok(sf_context *c, int iters)
{
    if (c->solver == SF_SOLVER_RBGS && !is_full_grid(c))
        return fail(c, SF_ERR_UNSUPPORTED, "SF_SOLVER_RBGS: full-grid contexts only (no slabs yet)");
    if (is_full_grid(c) || is_linked_slab(c)) return SF_OK;
    const bool stream_ok = jacobi_stream_supported(c->g) && !c->force_generic;
    const int L = (int)plan_launches(iters, stream_ok ? default_sweeps(c) : 1).size();
    if (L > 1)
        return fail(c, SF_ERR_UNSUPPORTED,
                    "slab context: a lin_solve of several launches needs a halo exchange between launches; "
                    "drive it with sf_jacobi_launch");
    return SF_OK;
}

// ---- step bodies (enqueue only) --------------------------------------------------------------
int enqueue_dens_step(sf_context *c, float *x, float *x0, const float *u, const float *v, float diff, float dt, int iters)
{
    if (is_linked_slab(c)) return slab_dens_step(c, x, x0, u, v, diff, dt, iters);
    const float fN = (float)c->g.N;
    float alpha = dt * diff;      // FluidSequential.c:179, left to right in binary32
    alpha = alpha * fN;
    alpha = alpha * fN;
    float beta = 4.0f * alpha;    // :180
    beta = 1.0f + beta;
    // Density fields have compact support with a decaying front, where the exact division needs its
    // guarded ticks: the one solve of a step whose warps are worth balancing (velocity fields are dense).
    c->steal_now = true;
    int rc = source_lin_solve(c, 0, x0, x, dt, alpha, beta, iters);   // add_source(x, x0); SWAP; diffuse(0, x, x0): solves into the old x0
    c->steal_now = false;
    if (rc) return rc;
    SF_CUDA(c, launch_advect(c->g, 0, x, x0, u, v, dt, advect_tile_now(c), c->tile_stats, c->work));   // SWAP; advect(0, x, x0, u, v)
    ++c->launches;
    return SF_OK;
}

int enqueue_project(sf_context *c, float *u, float *v, float *p, float *div, int iters)
{
    if (is_linked_slab(c)) return slab_project(c, u, v, p, div, iters);
    // the streaming lin_solve can start from an implicit zero guess, so p need not be written here
    // (the in-place red-black solver reads its guess: p is zeroed like in the reference)
    const bool stream_ok = jacobi_stream_supported(c->g) && !c->force_generic && c->solver == SF_SOLVER_JACOBI;
    SF_CUDA(c, launch_divergence(c->g, u, v, p, div, stream_ok ? 0 : 1, c->work));
    ++c->launches;
    int rc = lin_solve(c, 0, p, div, 1.0f, 4.0f, iters, stream_ok ? 1 : 0);
    if (rc) return rc;
    SF_CUDA(c, launch_last_project(c->g, u, v, p, c->work));
    ++c->launches;
    return SF_OK;
}

// vel_step in three pieces, so that a caller whose fields arrive one by one (sf_step_host) can start the u
// solve while v is still in flight: viscosity solve of one component (add_source + lin_solve, :193-210) ...
int enqueue_vel_diffuse(sf_context *c, int b, float *x, float *x0, float visc, float dt, int iters)
{
    const float fN = (float)c->g.N;
    float alpha = dt * visc;      // :199
    alpha = alpha * fN;
    alpha = alpha * fN;
    float beta = 4.0f * alpha;    // :200
    beta = 1.0f + beta;
    return source_lin_solve(c, b, x0, x, dt, alpha, beta, iters);
}
// ... and everything after the two solves (:213-240)
int enqueue_vel_tail(sf_context *c, float *u, float *v, float *u0, float *v0, float dt, int iters)
{
    int rc = enqueue_project(c, u0, v0, u, v, iters);                  // :213-223 (p in u, div in v)
    if (rc) return rc;
    SF_CUDA(c, launch_advect_uv(c->g, u, v, u0, v0, dt, advect_tile_now(c), c->tile_stats, c->work));   // :228-237
    ++c->launches;
    return enqueue_project(c, u, v, u0, v0, iters);                    // :238-240 (p in u0, div in v0)
}

int enqueue_vel_step(sf_context *c, float *u, float *v, float *u0, float *v0, float visc, float dt, int iters)
{
    if (is_linked_slab(c)) return slab_vel_step(c, u, v, u0, v0, visc, dt, iters);
    const float fN = (float)c->g.N;
    float alpha = dt * visc;      // :199
    alpha = alpha * fN;
    alpha = alpha * fN;
    float beta = 4.0f * alpha;    // :200
    beta = 1.0f + beta;
    int rc = source_lin_solve(c, 1, u0, u, dt, alpha, beta, iters);    // :193, :201-204
    if (rc) return rc;
    rc = source_lin_solve(c, 2, v0, v, dt, alpha, beta, iters);        // :197, :209-210
    if (rc) return rc;
    rc = enqueue_project(c, u0, v0, u, v, iters);                      // :213-223 (p in u, div in v)
    if (rc) return rc;
    SF_CUDA(c, launch_advect_uv(c->g, u, v, u0, v0, dt, advect_tile_now(c), c->tile_stats, c->work));   // :228-237
    ++c->launches;
    return enqueue_project(c, u, v, u0, v0, iters);                    // :238-240 (p in u0, div in v0)
}

// ---- the independent solves of a step side by side (SF_OPT_OVERLAP_SOLVES) -----------------------------
// A temporally blocked Jacobi launch is one wave of warps with one chunk of rows each; its warps finish up to ~15 % apart, and
// the next launch of the same solve cannot start before the last one has (it reads what that one writes).  But a step holds
// three solves that do not depend on each other -- the viscosity solves of u and v (FluidSequential.c:193-210) and the diffusion
// solve of the density (:177-182, which only meets the velocity in the advect that ends dens_step) -- so their launches go to
// three streams (forked and joined by events, which a stream capture turns into graph edges): the CTAs of one solve's next
// launch take the SM slots the previous launch of another solve frees while its last warps are still running.
// Every lane has its own ping-pong partner, right-hand-side field and ticket word; the kernels and their arguments are those of
// the sequential order, so the bits cannot differ.
struct LaneScope {
    sf_context *c;
    cudaStream_t work;
    float *scratch, *scratch2;
    unsigned *ticket;
    LaneScope(sf_context *ctx, int k) : c(ctx), work(ctx->work), scratch(ctx->scratch), scratch2(ctx->scratch2), ticket(ctx->ticket)
    {
        const sf_context::SolveLane &L = c->lanes[k];
        c->work = L.stream; c->scratch = L.scratch; c->scratch2 = L.scratch2; c->ticket = L.ticket;
    }
    ~LaneScope() { c->work = work; c->scratch = scratch; c->scratch2 = scratch2; c->ticket = ticket; }
};
int lane_fork(sf_context *c, int k)
{
    SF_CUDA(c, cudaEventRecord(c->lanes[k].fork, c->work));
    SF_CUDA(c, cudaStreamWaitEvent(c->lanes[k].stream, c->lanes[k].fork, 0));
    return SF_OK;
}
int lane_join(sf_context *c, int k)
{
    SF_CUDA(c, cudaEventRecord(c->lanes[k].join, c->lanes[k].stream));
    SF_CUDA(c, cudaStreamWaitEvent(c->work, c->lanes[k].join, 0));
    return SF_OK;
}
void diffusion_coefficients(const sf_context *c, float coef, float dt, float &alpha, float &beta)
{
    const float fN = (float)c->g.N;
    alpha = dt * coef;            // FluidSequential.c:179 / :199, left to right in binary32
    alpha = alpha * fN;
    alpha = alpha * fN;
    beta = 4.0f * alpha;          // :180 / :200
    beta = 1.0f + beta;
}

// the two viscosity solves of vel_step (:193-210), v on lane 0 beside u
int enqueue_vel_solves_overlapped(sf_context *c, float *u, float *v, float *u0, float *v0, float visc, float dt, int iters)
{
    float av, bv;
    diffusion_coefficients(c, visc, dt, av, bv);
    (void)arith_mode(c, av, bv);                       // the divisor check synchronises: before anything is forked
    int rc = lane_fork(c, 0);
    if (rc) return rc;
    int r0, r1;
    { LaneScope lane(c, 0); r0 = source_lin_solve(c, 2, v0, v, dt, av, bv, iters); }              // :197, :209-210
    r1 = source_lin_solve(c, 1, u0, u, dt, av, bv, iters);                                        // :193, :201-204
    rc = lane_join(c, 0);                              // (always joined: an unjoined stream would poison a capture)
    return rc ? rc : (r0 ? r0 : r1);
}

int enqueue_vel_step_overlapped(sf_context *c, float *u, float *v, float *u0, float *v0, float visc, float dt, int iters)
{
    const int rc = enqueue_vel_solves_overlapped(c, u, v, u0, v0, visc, dt, iters);
    return rc ? rc : enqueue_vel_tail(c, u, v, u0, v0, dt, iters);
}

// vel_step + dens_step (FluidSequential.c:305-306)
int enqueue_step(sf_context *c, float *dens, float *dens_prev, float *u, float *u_prev, float *v, float *v_prev, float visc, float diff,
                 float dt, int iters)
{
    if (!overlap_ok(c) || !c->lanes[1].join) {
        int r = enqueue_vel_step(c, u, v, u_prev, v_prev, visc, dt, iters);
        if (r) return r;
        return enqueue_dens_step(c, dens, dens_prev, u, v, diff, dt, iters);
    }
    // Two phases of two streams each: u || v viscosity solves, then the density solve beside the projection / advection /
    // projection chain of vel_step, which has no other partner.  (Forking the density solve at the start as well, and stream
    // priorities for the three branches, made no measurable difference; forking it after the first projection was 1.5 % slower:
    // profiles/r02/b2_overlap_priorities.txt, b3_overlap_order.txt, b4_overlap_fork_point.txt.)
    float ad, bd;
    diffusion_coefficients(c, diff, dt, ad, bd);
    (void)arith_mode(c, ad, bd);
    int rc = enqueue_vel_solves_overlapped(c, u, v, u_prev, v_prev, visc, dt, iters);
    if (rc) return rc;
    if ((rc = lane_fork(c, 1))) return rc;
    int r2;
    {
        LaneScope lane(c, 1);
        c->steal_now = true;                                                                      // see enqueue_dens_step
        r2 = source_lin_solve(c, 0, dens_prev, dens, dt, ad, bd, iters);                          // :177-182
        c->steal_now = false;
    }
    rc = enqueue_vel_tail(c, u, v, u_prev, v_prev, dt, iters);                                    // :213-240
    const int rj = lane_join(c, 1);
    if (rc || r2 || rj) return rc ? rc : (r2 ? r2 : rj);
    SF_CUDA(c, launch_advect(c->g, 0, dens, dens_prev, u, v, dt, advect_tile_now(c), c->tile_stats, c->work));   // :185
    ++c->launches;
    return SF_OK;
}

GraphKey make_key(const sf_context *c, int kind, std::initializer_list<const void *> ptrs, float f0, float f1, float f2, int iters)
{
    GraphKey k;
    std::memset(&k, 0, sizeof(k));
    k.kind = kind;
    int n = 0;
    for (const void *p : ptrs) k.p[n++] = p;
    k.f[0] = f0; k.f[1] = f1; k.f[2] = f2;
    k.iters = iters;
    k.opts[0] = c->arith; k.opts[1] = c->sweeps_opt; k.opts[2] = c->force_generic; k.opts[3] = c->chunk_rows;
    k.opts[4] = c->staging * 2 + (c->steal_opt ? 1 : 0) + 4 * c->steal_scope + 8 * c->pressure_plan + 16 * c->solver +
                32 * c->omega_milli + 65536 * c->rbgs_blocked + 131072 * c->fuse_sources;
    k.opts[5] = c->wave_skew; k.opts[6] = c->advect_tile + 32 * c->overlap + 64 * c->strip_balance;
    return k;
}

}  // namespace sf

namespace {

int create_common(sf_context **out, int N, int device, void *stream, bool own_stream, int row_lo, int row_hi, int halo)
{
    if (!out) return SF_ERR_INVALID;
    *out = nullptr;
    if (N < 1 || N > (1 << 24) - 2) return SF_ERR_INVALID;
    const int G = N + 2;
    if (row_lo < 0 || row_hi > G || row_hi <= row_lo || halo < 0) return SF_ERR_INVALID;
    // a partial slab reads its neighbours' rows through ghost rows (every stage is at least a radius-1 stencil), and a wall
    // row belongs to the slab that owns the interior row it mirrors (set_bnd is fused into the kernels that write row 1 / N)
    if ((row_lo > 0 || row_hi < G) && halo < 1) return SF_ERR_INVALID;
    if (row_lo == 1 || row_hi == G - 1) return SF_ERR_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return SF_ERR_CUDA;
    sf_context *c = new (std::nothrow) sf_context();
    if (!c) return SF_ERR_NOMEM;
    c->g.N = N; c->g.G = G;
    c->g.own_lo = row_lo; c->g.own_hi = row_hi;
    c->g.row_base = row_lo - halo;
    c->g.rows = row_hi - row_lo + 2 * halo;
    c->halo = halo;
    c->device = device;
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; return SF_ERR_CUDA; }
    c->sm_count = prop.multiProcessorCount;
    if (own_stream) {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return SF_ERR_CUDA; }
        c->own_stream = true;
    } else {
        c->stream = (cudaStream_t)stream;
    }
    c->work = c->stream;
    // counters of the TMA-staged advect kernel (a context without them runs the gather kernels)
    if (cudaMalloc(&c->tile_stats, 2 * sizeof(unsigned int)) != cudaSuccess || cudaMemset(c->tile_stats, 0, 2 * sizeof(unsigned int)) != cudaSuccess) {
        (void)cudaGetLastError();
        if (c->tile_stats) cudaFree(c->tile_stats);
        c->tile_stats = nullptr;
    }
    *out = c;
    return SF_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

int sf_create(sf_context **out, int N, int device) { return create_common(out, N, device, nullptr, true, 0, N + 2, 0); }
int sf_create_on_stream(sf_context **out, int N, int device, void *cuda_stream)
{
    return create_common(out, N, device, cuda_stream, false, 0, N + 2, 0);
}
int sf_create_slab(sf_context **out, int N, int device, void *cuda_stream, int row_lo, int row_hi, int halo)
{
    return create_common(out, N, device, cuda_stream, false, row_lo, row_hi, halo);
}

int sf_create_slab_own_stream(sf_context **out, int N, int device, int row_lo, int row_hi, int halo)
{
    return create_common(out, N, device, nullptr, true, row_lo, row_hi, halo);
}

int sf_destroy(sf_context *c)
{
    if (!c) return SF_ERR_INVALID;
    DeviceGuard guard(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto &e : c->graphs) if (e.exec) { cudaGraphExecDestroy(e.exec); cudaGraphDestroy(e.graph); }
    if (c->scratch && !c->scratch_in_arena) cudaFree(c->scratch);
    if (c->scratch2) cudaFree(c->scratch2);
    slab_release(c);
    if (c->steal) cudaFree(c->steal);
    if (c->red_f) cudaFree(c->red_f);
    if (c->red_d) cudaFree(c->red_d);
    if (c->ticket) cudaFree(c->ticket);
    if (c->tile_stats) cudaFree(c->tile_stats);
    release_lanes(c);
    for (auto &s : c->stage) if (s) cudaFree(s);
    for (auto &e : c->ev) if (e) cudaEventDestroy(e);
    if (c->h2d) cudaStreamDestroy(c->h2d);
    if (c->d2h) cudaStreamDestroy(c->d2h);
    if (c->cap_stream) cudaStreamDestroy(c->cap_stream);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return SF_OK;
}

const char *sf_last_error_string(const sf_context *c) { return c ? c->err.c_str() : "null context"; }

int sf_set_option(sf_context *c, int option, int value)
{
    if (!c) return SF_ERR_INVALID;
    switch (option) {
        case SF_OPT_ARITHMETIC: SF_REQUIRE(c, value == 0 || value == 1, "arithmetic: 0 strict / 1 fast"); c->arith = value; break;
        case SF_OPT_SWEEPS_PER_LAUNCH: SF_REQUIRE(c, value >= 0 && value <= 8, "sweeps per launch: 0..8"); c->sweeps_opt = value; break;
        case SF_OPT_USE_GRAPH: c->use_graph = value ? 1 : 0; break;
        case SF_OPT_FORCE_GENERIC: c->force_generic = value ? 1 : 0; break;
        case SF_OPT_CHUNK_ROWS: SF_REQUIRE(c, value >= 0, "chunk rows >= 0"); c->chunk_rows = value; break;
        case SF_OPT_STAGING: SF_REQUIRE(c, value == 0 || value == 1, "staging: 0 cp.async / 1 bulk copy"); c->staging = value; break;
        case SF_OPT_WORK_STEALING: {
            SF_REQUIRE(c, value >= 0 && value <= 100, "work stealing: 0 = off, 1..100 = smallest remaining share of a chunk (percent) worth halving");
            DeviceGuard guard(c->device);
            int rc = ensure_scratch(c);
            if (rc) return rc;
            c->steal_opt = value;
            SF_CUDA(c, cudaMemcpyAsync(&c->steal->min_pct, &c->steal_opt, sizeof(int), cudaMemcpyHostToDevice, c->stream));
            SF_CUDA(c, cudaStreamSynchronize(c->stream));
            break;
        }
        case SF_OPT_PRESSURE_PLAN: SF_REQUIRE(c, value >= 0 && value <= 2, "pressure plan: 0, 1 or 2"); c->pressure_plan = value; break;
        case SF_OPT_SOLVER:
            SF_REQUIRE(c, value == SF_SOLVER_JACOBI || value == SF_SOLVER_RBGS, "solver: 0 Jacobi (reference) / 1 red-black Gauss-Seidel");
            if (value == SF_SOLVER_RBGS && !is_full_grid(c)) return fail(c, SF_ERR_UNSUPPORTED, "SF_SOLVER_RBGS: full-grid contexts only (no slabs yet)");
            c->solver = value;
            break;
        case SF_OPT_SOR_OMEGA_MILLI: SF_REQUIRE(c, value >= 1 && value <= 1999, "SOR omega in 1/1000: 1..1999"); c->omega_milli = value; break;
        case SF_OPT_RBGS_BLOCKED: c->rbgs_blocked = value ? 1 : 0; break;
        case SF_OPT_FUSE_SOURCES: c->fuse_sources = value ? 1 : 0; break;
        case SF_OPT_OVERLAP_SOLVES: c->overlap = value ? 1 : 0; break;
        case SF_OPT_ADVECT_TILE_COUNT:
        case SF_OPT_ADVECT_FALLBACK_COUNT: {
            SF_REQUIRE(c, value == 0, "advect tile counters: only 0 (reset) can be set");
            DeviceGuard guard(c->device);
            if (c->tile_stats) {
                SF_CUDA(c, cudaStreamSynchronize(c->stream));
                SF_CUDA(c, cudaMemset(c->tile_stats, 0, 2 * sizeof(unsigned int)));
            }
            c->tile_seen[0] = c->tile_seen[1] = 0u;
            break;
        }
        case SF_OPT_ADVECT_TILE:
            SF_REQUIRE(c, (value >= 0 && value <= 8) || (value >= 12 && value <= 18),
                       "advect tile: 0 off / 1 automatic / 2..8 on, 32-row tiles with that many 8-row copies of shared memory per field / 12..18 on, 16-row tiles");
            c->advect_tile = value; c->advect_tile_live = true;
            break;
        case SF_OPT_WAVE_SKEW:
            SF_REQUIRE(c, value == 0 || (value / 1000 >= 100 && value / 1000 <= 200 && value % 1000 >= 50 && value % 1000 <= value / 1000 &&
                                         300 - value / 1000 - value % 1000 >= 20),
                       "wave skew: 0, or p0 * 1000 + p1 with 200 >= p0 >= p1 >= 50 and p0 + p1 <= 280 (percent of the mean chunk)");
            c->wave_skew = value;
            break;
        case SF_OPT_STEAL_SCOPE: SF_REQUIRE(c, value == 0 || value == 1, "steal scope: 0 scalar fields / 1 every strict solve"); c->steal_scope = value; break;
        case SF_OPT_STRIP_BALANCE: c->strip_balance = value ? 1 : 0; break;
        default: return fail(c, SF_ERR_INVALID, "unknown option");
    }
    return SF_OK;
}

int sf_get_option(const sf_context *c, int option, int *value)
{
    if (!c || !value) return SF_ERR_INVALID;
    switch (option) {
        case SF_OPT_ARITHMETIC: *value = c->arith; break;
        case SF_OPT_SWEEPS_PER_LAUNCH: *value = c->sweeps_opt; break;
        case SF_OPT_USE_GRAPH: *value = c->use_graph; break;
        case SF_OPT_FORCE_GENERIC: *value = c->force_generic; break;
        case SF_OPT_CHUNK_ROWS: *value = c->chunk_rows; break;
        case SF_OPT_STAGING: *value = c->staging; break;
        case SF_OPT_WORK_STEALING: *value = c->steal_opt; break;
        case SF_OPT_STEAL_SCOPE: *value = c->steal_scope; break;
        case SF_OPT_PRESSURE_PLAN: *value = c->pressure_plan; break;
        case SF_OPT_SOLVER: *value = c->solver; break;
        case SF_OPT_SOR_OMEGA_MILLI: *value = c->omega_milli; break;
        case SF_OPT_RBGS_BLOCKED: *value = c->rbgs_blocked; break;
        case SF_OPT_FUSE_SOURCES: *value = c->fuse_sources; break;
        case SF_OPT_OVERLAP_SOLVES: *value = c->overlap; break;
        case SF_OPT_ADVECT_TILE: *value = c->advect_tile; break;
        case SF_OPT_ADVECT_TILE_COUNT:
        case SF_OPT_ADVECT_FALLBACK_COUNT: {
            DeviceGuard guard(c->device);
            unsigned int st[2] = {0u, 0u};
            if (c->tile_stats && (cudaStreamSynchronize(c->stream) != cudaSuccess ||
                                  cudaMemcpy(st, c->tile_stats, sizeof(st), cudaMemcpyDeviceToHost) != cudaSuccess)) return SF_ERR_CUDA;
            *value = (int)st[option == SF_OPT_ADVECT_TILE_COUNT ? 0 : 1];
            break;
        }
        case SF_OPT_WAVE_SKEW: *value = c->wave_skew; break;
        case SF_OPT_STRIP_BALANCE: *value = c->strip_balance; break;
        case SF_OPT_STEAL_COUNT: {   // diagnostics: row ranges taken over by another warp so far (synchronises)
            *value = 0;
            if (c->steal) {
                DeviceGuard guard(c->device);
                if (cudaStreamSynchronize(c->stream) != cudaSuccess ||
                    cudaMemcpy(value, &c->steal->taken, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return SF_ERR_CUDA;
            }
            break;
        }
        default: return SF_ERR_INVALID;
    }
    return SF_OK;
}

int sf_synchronize(sf_context *c)
{
    if (!c) return SF_ERR_INVALID;
    DeviceGuard guard(c->device);
    SF_CUDA(c, cudaStreamSynchronize(c->stream));
    return SF_OK;
}

int sf_set_stream(sf_context *c, void *cuda_stream)
{
    if (!c) return SF_ERR_INVALID;
    SF_REQUIRE(c, !c->capturing, "set_stream during capture");
    DeviceGuard guard(c->device);
    for (auto &e : c->graphs) if (e.exec) { cudaGraphExecDestroy(e.exec); cudaGraphDestroy(e.graph); }
    c->graphs.clear();
    if (c->own_stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); c->own_stream = false; }
    c->stream = (cudaStream_t)cuda_stream;
    c->work = c->stream;
    return SF_OK;
}

int sf_get_stream(const sf_context *c, void **cuda_stream)
{
    if (!c || !cuda_stream) return SF_ERR_INVALID;
    *cuda_stream = (void *)c->stream;
    return SF_OK;
}

int sf_launch_count(const sf_context *c, unsigned long long *count)
{
    if (!c || !count) return SF_ERR_INVALID;
    *count = c->launches;
    return SF_OK;
}

size_t sf_field_bytes(const sf_context *c) { return c ? field_cells(c) * sizeof(float) : 0; }

int sf_alloc_field(sf_context *c, float **dev_field)
{
    if (!c || !dev_field) return SF_ERR_INVALID;
    DeviceGuard guard(c->device);
    SF_CUDA(c, cudaMalloc(dev_field, field_cells(c) * sizeof(float)));
    SF_CUDA(c, cudaMemsetAsync(*dev_field, 0, field_cells(c) * sizeof(float), c->stream));
    return SF_OK;
}

int sf_free_field(sf_context *c, float *dev_field)
{
    if (!c) return SF_ERR_INVALID;
    DeviceGuard guard(c->device);
    SF_CUDA(c, cudaStreamSynchronize(c->stream));
    SF_CUDA(c, cudaFree(dev_field));
    return SF_OK;
}

int sf_upload(sf_context *c, float *dev_field, const float *host_field)
{
    if (!c || !dev_field || !host_field) return SF_ERR_INVALID;
    DeviceGuard guard(c->device);
    SF_CUDA(c, cudaMemcpyAsync(dev_field, host_field, field_cells(c) * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    return SF_OK;
}

int sf_download(sf_context *c, float *host_field, const float *dev_field)
{
    if (!c || !dev_field || !host_field) return SF_ERR_INVALID;
    DeviceGuard guard(c->device);
    SF_CUDA(c, cudaMemcpyAsync(host_field, dev_field, field_cells(c) * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    SF_CUDA(c, cudaStreamSynchronize(c->stream));
    return SF_OK;
}

int sf_set_bnd(sf_context *c, int b, float *x)
{
    if (!c) return SF_ERR_INVALID;
    SF_REQUIRE(c, x && b >= 0 && b <= 2, "set_bnd: null field or b not in 0..2");
    DeviceGuard guard(c->device);
    SF_CUDA(c, launch_set_bnd(c->g, b, x, c->stream));
    ++c->launches;
    return SF_OK;
}

int sf_add_source(sf_context *c, float *x, const float *s, float dt)
{
    if (!c) return SF_ERR_INVALID;
    SF_REQUIRE(c, x && s, "add_source: null field");
    DeviceGuard guard(c->device);
    float *xs[1] = {x};
    const float *ss[1] = {s};
    SF_CUDA(c, launch_add_source(c->g, 1, xs, ss, dt, c->stream));
    ++c->launches;
    return SF_OK;
}

int sf_diffuse(sf_context *c, int b, float *x, const float *x0, float alpha, float beta, int iters)
{
    if (!c) return SF_ERR_INVALID;
    SF_REQUIRE(c, x && x0 && x != x0, "diffuse: null or aliased fields");
    SF_REQUIRE(c, b >= 0 && b <= 2, "diffuse: b not in 0..2");
    SF_REQUIRE(c, iters >= 1, "diffuse: iters < 1");
    DeviceGuard guard(c->device);
    c->link.barrier_valid = false;   // barriers collapse only within one call (see slab_barrier)
    int rc = check_multi_launch_ok(c, iters);
    if (rc) return rc;
    if (is_linked_slab(c)) (void)arith_mode(c, alpha, beta);   // the divisor check synchronises: before anything is enqueued
    c->steal_now = (b == 0);     // scalar (density-like) fields: see enqueue_dens_step
    rc = lin_solve(c, b, x, x0, alpha, beta, iters, 0);
    c->steal_now = false;
    return rc;
}

int sf_advect(sf_context *c, int b, float *d, const float *d0, const float *u, const float *v, float dt)
{
    if (!c) return SF_ERR_INVALID;
    SF_REQUIRE(c, d && d0 && u && v && d != d0 && d != u && d != v, "advect: null fields or output aliases an input");
    SF_REQUIRE(c, b >= 0 && b <= 2, "advect: b not in 0..2");
    DeviceGuard guard(c->device);
    c->link.barrier_valid = false;   // barriers collapse only within one call (see slab_barrier)
    if (is_linked_slab(c)) return slab_advect(c, b, d, d0, u, v, dt, true);
    SF_CUDA(c, launch_advect(c->g, b, d, d0, u, v, dt, advect_tile_now(c), c->tile_stats, c->stream));
    ++c->launches;
    return SF_OK;
}

int sf_advect_velocity(sf_context *c, float *u, float *v, const float *u0, const float *v0, float dt)
{
    if (!c) return SF_ERR_INVALID;
    SF_REQUIRE(c, u && v && u0 && v0 && u != v && u != u0 && u != v0 && v != u0 && v != v0 && u0 != v0,
               "advect_velocity: null fields or an output aliases another field");
    if (is_linked_slab(c)) return fail(c, SF_ERR_UNSUPPORTED, "advect_velocity on a connected slab: use sf_vel_step (it is one pass there too)");
    DeviceGuard guard(c->device);
    SF_CUDA(c, launch_advect_uv(c->g, u, v, u0, v0, dt, advect_tile_now(c), c->tile_stats, c->stream));   // :228-237
    ++c->launches;
    return SF_OK;
}

int sf_compute_divergence_and_pressure(sf_context *c, const float *u, const float *v, float *p, float *div)
{
    if (!c) return SF_ERR_INVALID;
    SF_REQUIRE(c, u && v && p && div && p != div && p != u && p != v && div != u && div != v, "divergence: null or aliased fields");
    DeviceGuard guard(c->device);
    SF_CUDA(c, launch_divergence(c->g, u, v, p, div, 1, c->stream));
    ++c->launches;
    return SF_OK;
}

int sf_last_project(sf_context *c, float *u, float *v, const float *p, const float *div)
{
    (void)div;   // unused by the reference too (FluidSequential.c:161-173)
    if (!c) return SF_ERR_INVALID;
    SF_REQUIRE(c, u && v && p && u != v && p != u && p != v, "lastProject: null or aliased fields");
    DeviceGuard guard(c->device);
    SF_CUDA(c, launch_last_project(c->g, u, v, p, c->stream));
    ++c->launches;
    return SF_OK;
}

int sf_project(sf_context *c, float *u, float *v, float *p, float *div, int iters)
{
    if (!c) return SF_ERR_INVALID;
    SF_REQUIRE(c, u && v && p && div && iters >= 1, "project: null field or iters < 1");
    DeviceGuard guard(c->device);
    c->link.barrier_valid = false;   // barriers collapse only within one call (see slab_barrier)
    int rc = check_multi_launch_ok(c, iters);
    if (rc) return rc;
    return enqueue_project(c, u, v, p, div, iters);
}

int sf_dens_step(sf_context *c, float *x, float *x0, const float *u, const float *v, float diff, float dt, int iters)
{
    if (!c) return SF_ERR_INVALID;
    SF_REQUIRE(c, x && x0 && u && v && x != x0, "dens_step: null or aliased fields");
    SF_REQUIRE(c, iters >= 1, "dens_step: iters < 1");
    DeviceGuard guard(c->device);
    c->link.barrier_valid = false;   // barriers collapse only within one call (see slab_barrier)
    int rc = check_multi_launch_ok(c, iters);
    if (rc) return rc;
    if (is_linked_slab(c) && (rc = slab_prevalidate(c, diff, dt))) return rc;
    const GraphKey key = make_key(c, 1, {x, x0, u, v}, diff, dt, 0.f, iters);
    return run_graphed(c, key, [&] { return enqueue_dens_step(c, x, x0, u, v, diff, dt, iters); });
}

int sf_vel_step(sf_context *c, float *u, float *v, float *u0, float *v0, float visc, float dt, int iters)
{
    if (!c) return SF_ERR_INVALID;
    SF_REQUIRE(c, u && v && u0 && v0 && u != v && u != u0 && u != v0 && v != u0 && v != v0 && u0 != v0, "vel_step: null or aliased fields");
    SF_REQUIRE(c, iters >= 1, "vel_step: iters < 1");
    DeviceGuard guard(c->device);
    c->link.barrier_valid = false;   // barriers collapse only within one call (see slab_barrier)
    int rc = check_multi_launch_ok(c, iters);
    if (rc) return rc;
    if (is_linked_slab(c) && (rc = slab_prevalidate(c, visc, dt))) return rc;
    const GraphKey key = make_key(c, 2, {u, v, u0, v0}, visc, dt, 0.f, iters);
    if (!c->use_graph && ((rc = ensure_scratch(c)) || (rc = ensure_lanes(c)))) return rc;
    return run_graphed(c, key, [&] {
        return overlap_ok(c) && c->lanes[0].join ? enqueue_vel_step_overlapped(c, u, v, u0, v0, visc, dt, iters)
                                                 : enqueue_vel_step(c, u, v, u0, v0, visc, dt, iters);
    });
}

int sf_step(sf_context *c, float *dens, float *dens_prev, float *u, float *u_prev, float *v, float *v_prev, float visc,
            float diff, float dt, int iters)
{
    if (!c) return SF_ERR_INVALID;
    SF_REQUIRE(c, dens && dens_prev && u && u_prev && v && v_prev, "step: null field");
    SF_REQUIRE(c, iters >= 1, "step: iters < 1");
    DeviceGuard guard(c->device);
    c->link.barrier_valid = false;   // barriers collapse only within one call (see slab_barrier)
    int rc = check_multi_launch_ok(c, iters);
    if (rc) return rc;
    if (is_linked_slab(c) && ((rc = slab_prevalidate(c, visc, dt)) || (rc = slab_prevalidate(c, diff, dt)))) return rc;
    const GraphKey key = make_key(c, 3, {dens, dens_prev, u, u_prev, v, v_prev}, visc, diff, dt, iters);
    if (!c->use_graph && ((rc = ensure_scratch(c)) || (rc = ensure_lanes(c)))) return rc;
    return run_graphed(c, key, [&] { return enqueue_step(c, dens, dens_prev, u, u_prev, v, v_prev, visc, diff, dt, iters); });
}

int sf_step_host(sf_context *c, float *dens, float *dens_prev, float *u, float *u_prev, float *v, float *v_prev,
                 float visc, float diff, float dt, int iters, int download_scratch)
{
    if (!c) return SF_ERR_INVALID;
    SF_REQUIRE(c, dens && dens_prev && u && u_prev && v && v_prev, "step_host: null field");
    SF_REQUIRE(c, iters >= 1, "step_host: iters < 1");
    SF_REQUIRE(c, is_full_grid(c) || is_linked_slab(c), "step_host: full-grid contexts and connected peer slabs only");
    DeviceGuard guard(c->device);
    if (!c->h2d) {
        SF_CUDA(c, cudaStreamCreateWithFlags(&c->h2d, cudaStreamNonBlocking));
        SF_CUDA(c, cudaStreamCreateWithFlags(&c->d2h, cudaStreamNonBlocking));
        for (auto &e : c->ev) SF_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    if (is_linked_slab(c)) {
        // Connected peer slab (collective, one caller per slab): the host arrays hold this slab's OWNED rows
        // (own_rows x (N+2)); the device fields are the first six fields of the arena in the order of the arguments
        // (sf_slab_field 0..5).  Ghost rows need no upload: every stage that reads them refreshes them from the neighbour.
        SF_REQUIRE(c, c->link.nfields >= 6, "step_host on a peer slab: the arena needs at least six fields");
        c->link.barrier_valid = false;
        int rc = check_multi_launch_ok(c, iters);
        if (rc) return rc;
        if ((rc = slab_prevalidate(c, visc, dt)) || (rc = slab_prevalidate(c, diff, dt))) return rc;
        const size_t G = (size_t)c->g.G;
        const size_t own_bytes = (size_t)(c->g.own_hi - c->g.own_lo) * G * sizeof(float);
        const size_t own_off = (size_t)(c->g.own_lo - c->g.row_base) * G;
        float *dev[6];
        for (int k = 0; k < 6; ++k) dev[k] = reinterpret_cast<float *>(c->link.base + (size_t)k * c->link.field_bytes);
        float *host[6] = {dens, dens_prev, u, u_prev, v, v_prev};
        SF_CUDA(c, cudaEventRecord(c->ev[5], c->stream));
        SF_CUDA(c, cudaStreamWaitEvent(c->h2d, c->ev[5], 0));            // the previous step is done with the device fields
        for (int k : {2, 3, 4, 5}) SF_CUDA(c, cudaMemcpyAsync(dev[k] + own_off, host[k], own_bytes, cudaMemcpyHostToDevice, c->h2d));
        SF_CUDA(c, cudaEventRecord(c->ev[0], c->h2d));                   // velocity first: vel_step starts while the density fields travel
        for (int k : {0, 1}) SF_CUDA(c, cudaMemcpyAsync(dev[k] + own_off, host[k], own_bytes, cudaMemcpyHostToDevice, c->h2d));
        SF_CUDA(c, cudaEventRecord(c->ev[1], c->h2d));
        SF_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev[0], 0));
        rc = run_graphed(c, make_key(c, 2, {dev[2], dev[4], dev[3], dev[5]}, visc, dt, 0.f, iters),
                         [&] { return enqueue_vel_step(c, dev[2], dev[4], dev[3], dev[5], visc, dt, iters); });
        if (rc) return rc;
        SF_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
        SF_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev[1], 0));
        rc = run_graphed(c, make_key(c, 1, {dev[0], dev[1], dev[2], dev[4]}, diff, dt, 0.f, iters),
                         [&] { return enqueue_dens_step(c, dev[0], dev[1], dev[2], dev[4], diff, dt, iters); });
        if (rc) return rc;
        SF_CUDA(c, cudaEventRecord(c->ev[3], c->stream));
        SF_CUDA(c, cudaStreamWaitEvent(c->d2h, c->ev[2], 0));            // u, v drain while dens_step runs
        for (int k : {2, 4}) SF_CUDA(c, cudaMemcpyAsync(host[k], dev[k] + own_off, own_bytes, cudaMemcpyDeviceToHost, c->d2h));
        if (download_scratch)
            for (int k : {3, 5}) SF_CUDA(c, cudaMemcpyAsync(host[k], dev[k] + own_off, own_bytes, cudaMemcpyDeviceToHost, c->d2h));
        SF_CUDA(c, cudaStreamWaitEvent(c->d2h, c->ev[3], 0));
        SF_CUDA(c, cudaMemcpyAsync(host[0], dev[0] + own_off, own_bytes, cudaMemcpyDeviceToHost, c->d2h));
        if (download_scratch) SF_CUDA(c, cudaMemcpyAsync(host[1], dev[1] + own_off, own_bytes, cudaMemcpyDeviceToHost, c->d2h));
        SF_CUDA(c, cudaStreamSynchronize(c->d2h));
        return SF_OK;
    }
    const size_t bytes = field_cells(c) * sizeof(float);
    if (!c->stage[0])
        for (auto &s : c->stage) SF_CUDA(c, cudaMalloc(&s, bytes));
    int rc = ensure_scratch(c);
    if (rc) return rc;
    float *d_dens = c->stage[0], *d_dens0 = c->stage[1], *d_u = c->stage[2], *d_u0 = c->stage[3], *d_v = c->stage[4], *d_v0 = c->stage[5];
    // the previous call's work on the compute stream must be done before staging is overwritten
    SF_CUDA(c, cudaEventRecord(c->ev[5], c->stream));
    SF_CUDA(c, cudaStreamWaitEvent(c->h2d, c->ev[5], 0));
    // u first, then v, then the density fields: the u solve starts while v is still in flight, the rest of
    // vel_step while the density fields are
    SF_CUDA(c, cudaMemcpyAsync(d_u, u, bytes, cudaMemcpyHostToDevice, c->h2d));
    SF_CUDA(c, cudaMemcpyAsync(d_u0, u_prev, bytes, cudaMemcpyHostToDevice, c->h2d));
    SF_CUDA(c, cudaEventRecord(c->ev[4], c->h2d));
    SF_CUDA(c, cudaMemcpyAsync(d_v, v, bytes, cudaMemcpyHostToDevice, c->h2d));
    SF_CUDA(c, cudaMemcpyAsync(d_v0, v_prev, bytes, cudaMemcpyHostToDevice, c->h2d));
    SF_CUDA(c, cudaEventRecord(c->ev[0], c->h2d));
    SF_CUDA(c, cudaMemcpyAsync(d_dens, dens, bytes, cudaMemcpyHostToDevice, c->h2d));
    SF_CUDA(c, cudaMemcpyAsync(d_dens0, dens_prev, bytes, cudaMemcpyHostToDevice, c->h2d));
    SF_CUDA(c, cudaEventRecord(c->ev[1], c->h2d));

    SF_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev[4], 0));
    rc = run_graphed(c, make_key(c, 4, {d_u, d_u0}, visc, dt, 0.f, iters), [&] { return enqueue_vel_diffuse(c, 1, d_u, d_u0, visc, dt, iters); });
    if (rc) return rc;
    SF_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev[0], 0));
    rc = run_graphed(c, make_key(c, 5, {d_v, d_v0}, visc, dt, 0.f, iters), [&] { return enqueue_vel_diffuse(c, 2, d_v, d_v0, visc, dt, iters); });
    if (rc) return rc;
    rc = run_graphed(c, make_key(c, 6, {d_u, d_v, d_u0, d_v0}, dt, 0.f, 0.f, iters), [&] { return enqueue_vel_tail(c, d_u, d_v, d_u0, d_v0, dt, iters); });
    if (rc) return rc;
    SF_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
    SF_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev[1], 0));
    rc = sf_dens_step(c, d_dens, d_dens0, d_u, d_v, diff, dt, iters);
    if (rc) return rc;
    SF_CUDA(c, cudaEventRecord(c->ev[3], c->stream));

    // velocity results drain while dens_step runs
    SF_CUDA(c, cudaStreamWaitEvent(c->d2h, c->ev[2], 0));
    SF_CUDA(c, cudaMemcpyAsync(u, d_u, bytes, cudaMemcpyDeviceToHost, c->d2h));
    SF_CUDA(c, cudaMemcpyAsync(v, d_v, bytes, cudaMemcpyDeviceToHost, c->d2h));
    if (download_scratch) {
        SF_CUDA(c, cudaMemcpyAsync(u_prev, d_u0, bytes, cudaMemcpyDeviceToHost, c->d2h));
        SF_CUDA(c, cudaMemcpyAsync(v_prev, d_v0, bytes, cudaMemcpyDeviceToHost, c->d2h));
    }
    SF_CUDA(c, cudaStreamWaitEvent(c->d2h, c->ev[3], 0));
    SF_CUDA(c, cudaMemcpyAsync(dens, d_dens, bytes, cudaMemcpyDeviceToHost, c->d2h));
    if (download_scratch) SF_CUDA(c, cudaMemcpyAsync(dens_prev, d_dens0, bytes, cudaMemcpyDeviceToHost, c->d2h));
    SF_CUDA(c, cudaStreamSynchronize(c->d2h));
    return SF_OK;
}

int sf_run_steps(sf_context *c, float *dens, float *dens_prev, float *u, float *u_prev, float *v, float *v_prev, float visc,
                 float diff, float dt, int iters, int steps, int source_mode, uint64_t seed, const float *src_dens,
                 const float *src_u, const float *src_v)
{
    if (!c) return SF_ERR_INVALID;
    SF_REQUIRE(c, dens && dens_prev && u && u_prev && v && v_prev, "run_steps: null field");
    SF_REQUIRE(c, iters >= 1 && steps >= 0, "run_steps: iters < 1 or steps < 0");
    SF_REQUIRE(c, source_mode >= SF_SOURCES_REFERENCE && source_mode <= SF_SOURCES_FIELDS, "run_steps: unknown source mode");
    if (source_mode == SF_SOURCES_FIELDS) SF_REQUIRE(c, src_dens && src_u && src_v, "run_steps: SF_SOURCES_FIELDS needs the three source fields");
    DeviceGuard guard(c->device);
    const size_t bytes = field_cells(c) * sizeof(float);
    for (int k = 0; k < steps; ++k) {
        // the source schedule runs on the device, ordered on the context's stream like the step itself
        if (source_mode == SF_SOURCES_REFERENCE && k > 0) {
            SF_CUDA(c, cudaMemsetAsync(dens_prev, 0, bytes, c->stream));
            SF_CUDA(c, cudaMemsetAsync(u_prev, 0, bytes, c->stream));
            SF_CUDA(c, cudaMemsetAsync(v_prev, 0, bytes, c->stream));
        } else if (source_mode == SF_SOURCES_SYNTHETIC) {
            SF_CUDA(c, launch_init(c->g, seed + (uint64_t)k, nullptr, dens_prev, nullptr, u_prev, nullptr, v_prev, c->stream));
            ++c->launches;
        } else if (source_mode == SF_SOURCES_FIELDS) {
            SF_CUDA(c, cudaMemcpyAsync(dens_prev, src_dens, bytes, cudaMemcpyDeviceToDevice, c->stream));
            SF_CUDA(c, cudaMemcpyAsync(u_prev, src_u, bytes, cudaMemcpyDeviceToDevice, c->stream));
            SF_CUDA(c, cudaMemcpyAsync(v_prev, src_v, bytes, cudaMemcpyDeviceToDevice, c->stream));
        }
        int rc = sf_step(c, dens, dens_prev, u, u_prev, v, v_prev, visc, diff, dt, iters);
        if (rc) return rc;
    }
    return SF_OK;
}

int sf_dump_field(sf_context *c, const float *dev_field, const char *path)
{
    if (!c) return SF_ERR_INVALID;
    SF_REQUIRE(c, dev_field && path, "dump_field: null argument");
    DeviceGuard guard(c->device);
    const size_t G = (size_t)c->g.G, own = (size_t)(c->g.own_hi - c->g.own_lo);
    std::vector<float> host(own * G);
    SF_CUDA(c, cudaMemcpyAsync(host.data(), dev_field + (size_t)(c->g.own_lo - c->g.row_base) * G, own * G * sizeof(float),
                               cudaMemcpyDeviceToHost, c->stream));
    SF_CUDA(c, cudaStreamSynchronize(c->stream));
    FILE *f = std::fopen(path, "wb");
    if (!f) return fail(c, SF_ERR_INVALID, "dump_field: cannot open the output file");
    const int32_t hdr[8] = {0x444C4653 /* "SFLD" */, 1, c->g.N, c->g.own_lo, c->g.own_hi, c->halo, 0, 0};
    const bool ok = std::fwrite(hdr, sizeof(hdr), 1, f) == 1 && std::fwrite(host.data(), sizeof(float), host.size(), f) == host.size();
    std::fclose(f);
    if (!ok) return fail(c, SF_ERR_INVALID, "dump_field: short write");
    return SF_OK;
}

int sf_init_synthetic(sf_context *c, uint64_t seed, float *dens, float *dens_prev, float *u, float *u_prev, float *v, float *v_prev)
{
    if (!c) return SF_ERR_INVALID;
    SF_REQUIRE(c, dens && dens_prev && u && u_prev && v && v_prev, "init_synthetic: null field");
    DeviceGuard guard(c->device);
    SF_CUDA(c, launch_init(c->g, seed, dens, dens_prev, u, u_prev, v, v_prev, c->stream));
    ++c->launches;
    return SF_OK;
}

int sf_init_sources(sf_context *c, uint64_t seed, float *dens_prev, float *u_prev, float *v_prev)
{
    if (!c) return SF_ERR_INVALID;
    SF_REQUIRE(c, dens_prev && u_prev && v_prev, "init_sources: null field");
    DeviceGuard guard(c->device);
    SF_CUDA(c, launch_init(c->g, seed, nullptr, dens_prev, nullptr, u_prev, nullptr, v_prev, c->stream));
    ++c->launches;
    return SF_OK;
}

int sf_reduce_max_abs(sf_context *c, const float *x, float *host_out)
{
    if (!c) return SF_ERR_INVALID;
    SF_REQUIRE(c, x && host_out, "reduce_max_abs: null argument");
    DeviceGuard guard(c->device);
    int rc = ensure_scratch(c);
    if (rc) return rc;
    SF_CUDA(c, launch_max_abs(c->g, x, c->red_f, true, c->stream));
    ++c->launches;
    SF_CUDA(c, cudaMemcpyAsync(host_out, c->red_f, sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    SF_CUDA(c, cudaStreamSynchronize(c->stream));
    return SF_OK;
}

int sf_reduce_max_abs_async(sf_context *c, const float *x, float *dev_out)
{
    if (!c) return SF_ERR_INVALID;
    SF_REQUIRE(c, x && dev_out, "reduce_max_abs_async: null argument");
    DeviceGuard guard(c->device);
    SF_CUDA(c, launch_max_abs(c->g, x, dev_out, false, c->work));
    ++c->launches;
    return SF_OK;
}

int sf_residual_l2(sf_context *c, const float *x, const float *x0, float alpha, float beta, double *host_out)
{
    if (!c) return SF_ERR_INVALID;
    SF_REQUIRE(c, x && x0 && host_out, "residual_l2: null argument");
    DeviceGuard guard(c->device);
    int rc = ensure_scratch(c);
    if (rc) return rc;
    SF_CUDA(c, launch_residual(c->g, x, x0, alpha, beta, c->red_d, c->stream));
    ++c->launches;
    double sumsq = 0.0;
    SF_CUDA(c, cudaMemcpyAsync(&sumsq, c->red_d, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SF_CUDA(c, cudaStreamSynchronize(c->stream));
    *host_out = sumsq;   // caller takes sqrt after summing over slabs
    return SF_OK;
}

int sf_division_check(sf_context *c, float beta, int *exact)
{
    if (!c || !exact) return SF_ERR_INVALID;
    DeviceGuard guard(c->device);
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    const bool can_run = cudaStreamIsCapturing(c->stream, &st) == cudaSuccess && st == cudaStreamCaptureStatusNone;
    *exact = division_validated(beta, can_run, c->stream) ? 1 : 0;
    return SF_OK;
}

int sf_halo_rows_needed(const sf_context *c, int *rows)
{
    if (!c || !rows) return SF_ERR_INVALID;
    *rows = default_sweeps(c);
    return SF_OK;
}

int sf_jacobi_launch(sf_context *c, int b, float *xout, const float *xin, const float *x0, float alpha, float beta,
                     int sweeps, int out_lo, int out_hi)
{
    if (!c) return SF_ERR_INVALID;
    SF_REQUIRE(c, xout && xin && x0 && xout != xin && xout != x0, "jacobi_launch: null or aliased fields");
    SF_REQUIRE(c, b >= 0 && b <= 2, "jacobi_launch: b not in 0..2");
    SF_REQUIRE(c, sweeps >= 1 && sweeps <= 8, "jacobi_launch: sweeps not in 1..8");
    if (out_lo < 0 && out_hi < 0) { out_lo = c->g.own_lo; out_hi = c->g.own_hi; }
    SF_REQUIRE(c, out_lo >= c->g.own_lo && out_hi <= c->g.own_hi && out_lo < out_hi, "jacobi_launch: rows outside the owned range");
    {   // rows the launch reads: [max(out_lo - sweeps, 0), min(out_hi - 1 + sweeps, G - 1)]
        const int rd_lo = out_lo - sweeps > 0 ? out_lo - sweeps : 0;
        const int rd_hi = out_hi - 1 + sweeps < c->g.G - 1 ? out_hi - 1 + sweeps : c->g.G - 1;
        SF_REQUIRE(c, rd_lo >= c->g.row_base && rd_hi < c->g.row_base + c->g.rows, "jacobi_launch: halo rows too few for this many sweeps");
    }
    DeviceGuard guard(c->device);
    return one_jacobi_launch(c, c->stream, b, xout, xin, x0, alpha, beta, sweeps, out_lo, out_hi, 0);
}

}  // extern "C"
