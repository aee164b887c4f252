/*
 * stam_oracle.c -- CPU restatement of the reference's SEQUENTIAL stable-fluids path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (fluidsimulationcuda_b200/, include/)
 * may link, import or call this file.  Allowed callers: tests/, __graft_entry__.smoke(),
 * and bench.py's cpu_baseline / --impl reference legs.
 *
 * What it restates: /root/reference/project/sequential/FluidSequential.c, the parity target
 * named by SURVEY.md section 8(c).  The reference bakes N, DT, K(=40) into macros; here they are
 * run-time arguments so that every BASELINE.json config (N, iteration count, odd K) has an
 * oracle.  The arithmetic (operand association, true divisions, float/int conversions,
 * truncation) follows the reference line by line; each function cites the lines it follows.
 *
 * Parity pin: tests/test_oracle_pin.py proves this file BITWISE equal to the reference itself
 * (oracle/_ref/libref_seq_N*_K*.so, built from the reference's own source by oracle/Makefile)
 * for every stage function and for multi-step runs, and against the committed fixtures in
 * tests/golden/ (generated from that same reference build by tests/golden/make_golden.py).
 *
 * Build: gcc -O2 -ffp-contract=off (no -mfma / -march=native / -ffast-math): x86-64 SSE2 scalar
 * float arithmetic is IEEE-754 binary32 with round-to-nearest-even and no fused multiply-add, the
 * same arithmetic the reference gets from "gcc FluidSequential.c".
 * Optional -fopenmp only splits independent rows over threads; results are identical.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define AT(row, col) ((size_t)(row) * (size_t)G + (size_t)(col))
/* The solver the step functions call.  Always so_lin_solve (the reference's Jacobi) in the oracle itself;
 * oracle/rbgs_check.c re-compiles this file with its own dispatcher to check the product's OPT-IN red-black
 * solver against a CPU build of the same scheme. */
#ifndef SO_SOLVE
#define SO_SOLVE so_lin_solve
#endif

/* FluidSequential.c:62-75 -- wall cells mirror the adjacent interior cell (negated on the
 * left/right walls for b==1, on the top/bottom walls for b==2); corners average their two
 * wall neighbours.  Walls first, corners after. */
void so_set_bnd(int N, int b, float *x)
{
    const int G = N + 2;
    const float fx = (b == 1) ? -1.0f : 1.0f; /* negation == multiply by -1 exactly */
    const float fy = (b == 2) ? -1.0f : 1.0f;
    for (int k = 1; k <= N; ++k) {
        x[AT(k, 0)]     = fx * x[AT(k, 1)];
        x[AT(k, N + 1)] = fx * x[AT(k, N)];
        x[AT(0, k)]     = fy * x[AT(1, k)];
        x[AT(N + 1, k)] = fy * x[AT(N, k)];
    }
    x[AT(0, 0)]         = 0.5f * (x[AT(0, 1)] + x[AT(1, 0)]);
    x[AT(N + 1, 0)]     = 0.5f * (x[AT(N + 1, 1)] + x[AT(N, 0)]);
    x[AT(0, N + 1)]     = 0.5f * (x[AT(0, N)] + x[AT(1, N + 1)]);
    x[AT(N + 1, N + 1)] = 0.5f * (x[AT(N + 1, N)] + x[AT(N, N + 1)]);
}

/* FluidSequential.c:78-82 -- every cell of the (N+2)^2 array, ring included. */
void so_add_source(int N, float *x, const float *s, float dt)
{
    const size_t cells = (size_t)(N + 2) * (size_t)(N + 2);
#pragma omp parallel for schedule(static)
    for (size_t c = 0; c < cells; ++c) {
        float inc = dt * s[c];
        x[c] = x[c] + inc;
    }
}

/* One Jacobi sweep, FluidSequential.c:93-98: ((left + right) + up) + down, then
 * x0 + alpha*sum, then a true division by beta. */
static void jacobi_sweep(int N, float *dst, const float *src, const float *rhs, float alpha, float beta)
{
    const int G = N + 2;
#pragma omp parallel for schedule(static)
    for (int r = 1; r <= N; ++r) {
        const float *up = src + AT(r - 1, 0), *mid = src + AT(r, 0), *dn = src + AT(r + 1, 0);
        const float *b0 = rhs + AT(r, 0);
        float *out = dst + AT(r, 0);
        for (int c = 1; c <= N; ++c) {
            float nb = mid[c - 1] + mid[c + 1];
            nb = nb + up[c];
            nb = nb + dn[c];
            float scaled = alpha * nb;
            float num = b0[c] + scaled;
            out[c] = num / beta;
        }
    }
}

/* FluidSequential.c:85-104 (lin_solve): iters x { sweep into the other buffer; swap; set_bnd }.
 * The reference only returns the result in the caller's x for even iteration counts (it frees
 * whichever buffer the last swap left in x_new); here the result ALWAYS ends in x, which is
 * what the reference delivers for its own (even) K.  The sweep never writes ring cells, so the
 * scratch buffer's ring is whatever set_bnd put there -- same as the reference's malloc'ed
 * x_new, whose ring is written by set_bnd before anything reads it. */
void so_lin_solve(int N, int b, float *x, const float *x0, float alpha, float beta, int iters)
{
    const size_t cells = (size_t)(N + 2) * (size_t)(N + 2);
    float *other = (float *)malloc(cells * sizeof(float));
    float *cur = x, *nxt = other;
    for (int k = 0; k < iters; ++k) {
        jacobi_sweep(N, nxt, cur, x0, alpha, beta);
        float *t = cur; cur = nxt; nxt = t;
        so_set_bnd(N, b, cur);
    }
    if (cur != x) memcpy(x, cur, cells * sizeof(float));
    free(other);
}

/* FluidSequential.c:107-141 -- semi-Lagrangian back-trace with clamp to [0.5, N+0.5],
 * truncating int conversion, bilinear blend s0*(t0*a + t1*b) + s1*(t0*c + t1*d). */
void so_advect(int N, int b, float *d, const float *d0, const float *u, const float *v, float dt)
{
    const int G = N + 2;
    const float dt0 = dt * (float)N;            /* :111 */
    const float lo = 0.5f, hi = (float)N + 0.5f; /* :117-127; N+0.5 is exact in binary32 */
#pragma omp parallel for schedule(static)
    for (int r = 1; r <= N; ++r) {
        for (int c = 1; c <= N; ++c) {
            float px = (float)c - dt0 * u[AT(r, c)];
            float py = (float)r - dt0 * v[AT(r, c)];
            if (px < lo) px = lo;
            if (px > hi) px = hi;
            if (py < lo) py = lo;
            if (py > hi) py = hi;
            int c0 = (int)px, r0 = (int)py;
            int c1 = c0 + 1, r1 = r0 + 1;
            float wx1 = px - (float)c0, wx0 = 1.0f - wx1;
            float wy1 = py - (float)r0, wy0 = 1.0f - wy1;
            float colA = wy0 * d0[AT(r0, c0)] + wy1 * d0[AT(r1, c0)];
            float colB = wy0 * d0[AT(r0, c1)] + wy1 * d0[AT(r1, c1)];
            d[AT(r, c)] = wx0 * colA + wx1 * colB;
        }
    }
    so_set_bnd(N, b, d);
}

/* FluidSequential.c:143-158 -- div = (-0.5f*h) * (((u_r - u_l) + v_d) - v_u), p = 0. */
void so_compute_divergence_and_pressure(int N, const float *u, const float *v, float *p, float *div)
{
    const int G = N + 2;
    const float h = 1.0f / (float)N;
    const float scale = -0.5f * h;
#pragma omp parallel for schedule(static)
    for (int r = 1; r <= N; ++r) {
        for (int c = 1; c <= N; ++c) {
            float acc = u[AT(r, c + 1)] - u[AT(r, c - 1)];
            acc = acc + v[AT(r + 1, c)];
            acc = acc - v[AT(r - 1, c)];
            div[AT(r, c)] = scale * acc;
            p[AT(r, c)] = 0.0f;
        }
    }
    so_set_bnd(N, 0, div);
    so_set_bnd(N, 0, p);
}

/* FluidSequential.c:161-173 -- u -= (0.5f*(p_r - p_l)) / h ; v -= (0.5f*(p_d - p_u)) / h. */
void so_last_project(int N, float *u, float *v, const float *p, const float *div)
{
    (void)div;
    const int G = N + 2;
    const float h = 1.0f / (float)N;
#pragma omp parallel for schedule(static)
    for (int r = 1; r <= N; ++r) {
        for (int c = 1; c <= N; ++c) {
            float gx = 0.5f * (p[AT(r, c + 1)] - p[AT(r, c - 1)]);
            float gy = 0.5f * (p[AT(r + 1, c)] - p[AT(r - 1, c)]);
            u[AT(r, c)] = u[AT(r, c)] - gx / h;
            v[AT(r, c)] = v[AT(r, c)] - gy / h;
        }
    }
    so_set_bnd(N, 1, u);
    so_set_bnd(N, 2, v);
}

/* FluidSequential.c:176-186.  The reference swaps its local pointers; written out, the
 * diffusion solves INTO x0 (guess = x0's content, rhs = x) and the advection writes x. */
void so_dens_step(int N, float *x, float *x0, const float *u, const float *v, float diff, float dt, int iters)
{
    so_add_source(N, x, x0, dt);
    float alpha = dt * diff * (float)N * (float)N; /* :179 left-to-right */
    float beta = 1.0f + 4.0f * alpha;               /* :180 */
    SO_SOLVE(N, 0, x0, x, alpha, beta, iters);
    so_advect(N, 0, x, x0, u, v, dt);
}

/* FluidSequential.c:189-241 with the pointer swaps written out. */
void so_vel_step(int N, float *u, float *v, float *u0, float *v0, float visc, float dt, int iters)
{
    so_add_source(N, u, u0, dt);
    so_add_source(N, v, v0, dt);
    float alpha = dt * visc * (float)N * (float)N; /* :199 */
    float beta = 1.0f + 4.0f * alpha;               /* :200 */
    SO_SOLVE(N, 1, u0, u, alpha, beta, iters);  /* :201-204 */
    SO_SOLVE(N, 2, v0, v, alpha, beta, iters);  /* :209-210 */
    /* project #1 on (u0, v0); u holds p, v holds div  (:213-223) */
    so_compute_divergence_and_pressure(N, u0, v0, u, v);
    SO_SOLVE(N, 0, u, v, 1.0f, 4.0f, iters);
    so_last_project(N, u0, v0, u, v);
    /* :228-237 advect both components with the projected field */
    so_advect(N, 1, u, u0, u0, v0, dt);
    so_advect(N, 2, v, v0, u0, v0, dt);
    /* project #2 on (u, v); u0 holds p, v0 holds div (:238-240) */
    so_compute_divergence_and_pressure(N, u, v, u0, v0);
    SO_SOLVE(N, 0, u0, v0, 1.0f, 4.0f, iters);
    so_last_project(N, u, v, u0, v0);
}

/* FluidSequential.c:244-271 -- the reference's own initial condition: glibc rand() with its
 * default seed, density source on the centred square first (row-major), then per cell u, v. */
void so_init_reference_rand(int N, float *dens, float *dens_prev, float *u, float *u_prev, float *v, float *v_prev)
{
    const int G = N + 2;
    const int mid = G / 2, half = G / 8;
    srand(1); /* what an un-seeded process starts with */
    for (int r = 0; r < G; ++r)
        for (int c = 0; c < G; ++c) {
            int inside = (c < mid + half) && (c >= mid - half) && (r < mid + half) && (r >= mid - half);
            dens_prev[AT(r, c)] = inside ? (float)(rand() % 100) / 1000.0f : 0.0f;
            dens[AT(r, c)] = 0.0f;
        }
    for (int r = 0; r < G; ++r)
        for (int c = 0; c < G; ++c) {
            u_prev[AT(r, c)] = (float)(rand() % 100) / 100.0f;
            v_prev[AT(r, c)] = (float)(rand() % 100) / 100.0f;
            u[AT(r, c)] = 0.0f;
            v[AT(r, c)] = 0.0f;
        }
}

/* Counter-based synthetic initial condition with the reference's value distributions
 * (SURVEY.md section 8(d)): same formula as the device-side generator sf_init_synthetic, so big
 * grids never cross PCIe.  32-bit avalanche mixer over (seed, field, cell), reduced to 0..99. */
static inline uint32_t so_hash100(uint64_t seed, uint64_t field, uint64_t cell)
{
    uint32_t h = (uint32_t)cell ^ (uint32_t)(cell >> 32) * 0x9E3779B1u;
    h ^= (uint32_t)seed * 0x85EBCA6Bu + (uint32_t)field * 0xC2B2AE35u + 0x27D4EB2Fu;
    h ^= h >> 16; h *= 0x85EBCA6Bu;
    h ^= h >> 13; h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return (uint32_t)(((uint64_t)h * 100ull) >> 32);
}

void so_init_synthetic(int N, uint64_t seed, float *dens, float *dens_prev, float *u, float *u_prev, float *v, float *v_prev)
{
    const int G = N + 2;
    const int mid = G / 2, half = G / 8;
#pragma omp parallel for schedule(static)
    for (int r = 0; r < G; ++r)
        for (int c = 0; c < G; ++c) {
            size_t cell = AT(r, c);
            int inside = (c < mid + half) && (c >= mid - half) && (r < mid + half) && (r >= mid - half);
            dens_prev[cell] = inside ? (float)so_hash100(seed, 0, cell) / 1000.0f : 0.0f;
            u_prev[cell] = (float)so_hash100(seed, 1, cell) / 100.0f;
            v_prev[cell] = (float)so_hash100(seed, 2, cell) / 100.0f;
            dens[cell] = 0.0f; u[cell] = 0.0f; v[cell] = 0.0f;
        }
}

/* Reference main loop body, FluidSequential.c:289-312: sources are live in step 0 only and
 * zeroed before every later step. */
void so_run_steps(int N, int steps, int first_step, float *dens, float *dens_prev, float *u, float *u_prev,
                  float *v, float *v_prev, float visc, float diff, float dt, int iters)
{
    const size_t bytes = (size_t)(N + 2) * (size_t)(N + 2) * sizeof(float);
    for (int s = 0; s < steps; ++s) {
        if (first_step + s > 0) {
            memset(u_prev, 0, bytes); memset(v_prev, 0, bytes); memset(dens_prev, 0, bytes);
        }
        so_vel_step(N, u, v, u_prev, v_prev, visc, dt, iters);
        so_dens_step(N, dens, dens_prev, u, v, diff, dt, iters);
    }
}
