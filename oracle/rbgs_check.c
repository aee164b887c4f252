/*
 * rbgs_check.c -- CPU build of the product's OPT-IN red-black Gauss-Seidel / SOR solver.
 *
 * TEST INFRASTRUCTURE ONLY (same rules as stam_oracle.c).  This is NOT a restatement of the reference:
 * ArbiterMob/FluidSimulationCuda solves with double-buffered Jacobi everywhere (FluidSequential.c:85-104)
 * and stam_oracle.c restates exactly that.  The product additionally offers SF_OPT_SOLVER = SF_SOLVER_RBGS
 * (fluidsimulationcuda_b200/csrc/sf_solvers.cu); BASELINE.json's north star asks
 * that a red-black variant be "validated against a CPU red-black build" -- this file is that build.
 * A half-sweep only reads cells of the other colour, so the scheme is order-independent and the GPU must
 * match it BITWISE (tests/test_zzz_solvers_gpu.py).
 *
 * Scheme (one iteration): red half-sweep ((row + col) even), black half-sweep ((row + col) odd), set_bnd(b);
 * cell update gs = (x0 + alpha*(((l + r) + up) + dn)) / beta with the reference's operand order
 * (FluidSequential.c:95-96); omega == 1: x = gs; otherwise x = x + omega*(gs - x), three roundings.
 *
 * The step functions are the oracle's own (stam_oracle.c is compiled into this object a second time) with
 * their solver call routed through rb_dispatch: rb_set_solver(0, .) gives the reference path, (1, omega) the
 * red-black one.
 */
#include <stddef.h>

static void rb_dispatch(int N, int b, float *x, const float *x0, float alpha, float beta, int iters);
#define SO_SOLVE rb_dispatch
#include "stam_oracle.c"

static int g_solver = 0;
static float g_omega = 1.0f;

void rb_set_solver(int solver, float omega) { g_solver = solver; g_omega = omega; }

void rb_lin_solve(int N, int b, float *x, const float *x0, float alpha, float beta, int iters, float omega)
{
    const int G = N + 2;
    for (int k = 0; k < iters; ++k) {
        for (int colour = 0; colour < 2; ++colour) {
#pragma omp parallel for schedule(static)
            for (int r = 1; r <= N; ++r) {
                for (int c = 1 + ((r + 1 + colour) & 1); c <= N; c += 2) {
                    float nb = x[AT(r, c - 1)] + x[AT(r, c + 1)];
                    nb = nb + x[AT(r - 1, c)];
                    nb = nb + x[AT(r + 1, c)];
                    float scaled = alpha * nb;
                    float num = x0[AT(r, c)] + scaled;
                    float gs = num / beta;
                    if (omega == 1.0f) {
                        x[AT(r, c)] = gs;
                    } else {
                        float xo = x[AT(r, c)];
                        float d = gs - xo;
                        d = omega * d;
                        x[AT(r, c)] = xo + d;
                    }
                }
            }
        }
        so_set_bnd(N, b, x);
    }
}

static void rb_dispatch(int N, int b, float *x, const float *x0, float alpha, float beta, int iters)
{
    if (g_solver == 1) rb_lin_solve(N, b, x, x0, alpha, beta, iters, g_omega);
    else so_lin_solve(N, b, x, x0, alpha, beta, iters);
}

/* sum over the interior of (beta*x - alpha*(l + r + up + dn) - x0)^2 in binary64: how far x is from solving the system */
double rb_residual_sumsq(int N, const float *x, const float *x0, float alpha, float beta)
{
    const int G = N + 2;
    double acc = 0.0;
    for (int r = 1; r <= N; ++r)
        for (int c = 1; c <= N; ++c) {
            double nb = (double)x[AT(r, c - 1)] + (double)x[AT(r, c + 1)] + (double)x[AT(r - 1, c)] + (double)x[AT(r + 1, c)];
            double e = (double)beta * (double)x[AT(r, c)] - (double)alpha * nb - (double)x0[AT(r, c)];
            acc += e * e;
        }
    return acc;
}
