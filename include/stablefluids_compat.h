/*
 * stablefluids_compat.h -- the reference's exact solver names over libstablefluids_b200.so.
 *
 * ArbiterMob/FluidSimulationCuda defines these functions inside every program
 * (project/sequential/FluidSequential.c:62,78,85,107,143,161,176,189) with N, DT and the iteration
 * count baked in.  Include this header INSTEAD of those definitions, bind a context once, and the
 * reference's main-loop body (FluidSequential.c:305-306) compiles unchanged.  Pointers are DEVICE
 * pointers (sf_alloc_field / sf_upload / sf_download replace malloc and the printouts' reads).
 * Errors abort like the reference's CHECK macro does (naivePar/...Naive.cu:26-35) unless
 * SF_COMPAT_NO_ABORT is defined, in which case sf_compat_status holds the last status.
 */
#ifndef STABLEFLUIDS_COMPAT_H
#define STABLEFLUIDS_COMPAT_H

#include <stdio.h>
#include <stdlib.h>

#include "stablefluids.h"

#ifndef SF_COMPAT_ITERS
#define SF_COMPAT_ITERS 40 /* FluidSequential.c:91 */
#endif

static sf_context *sf_compat_ctx = NULL;
static float sf_compat_dt = 0.016f; /* FluidSequential.c:7 */
static int sf_compat_status = 0;

static inline void sf_compat_bind(sf_context *ctx, float dt) { sf_compat_ctx = ctx; sf_compat_dt = dt; }

static inline void sf_compat_check(int rc, const char *what)
{
    sf_compat_status = rc;
#ifndef SF_COMPAT_NO_ABORT
    if (rc != SF_OK) {
        fprintf(stderr, "%s failed (%d): %s\n", what, rc, sf_last_error_string(sf_compat_ctx));
        exit(EXIT_FAILURE);
    }
#else
    (void)what;
#endif
}

static inline void set_bnd(int b, float *x) { sf_compat_check(sf_set_bnd(sf_compat_ctx, b, x), "set_bnd"); }
static inline void add_source(float *x, float *s) { sf_compat_check(sf_add_source(sf_compat_ctx, x, s, sf_compat_dt), "add_source"); }
static inline void diffuse(int b, float *x, float *x0, float alpha, float beta)
{
    sf_compat_check(sf_diffuse(sf_compat_ctx, b, x, x0, alpha, beta, SF_COMPAT_ITERS), "diffuse");
}
static inline void advect(int b, float *d, float *d0, float *u, float *v)
{
    sf_compat_check(sf_advect(sf_compat_ctx, b, d, d0, u, v, sf_compat_dt), "advect");
}
static inline void computeDivergenceAndPressure(float *u, float *v, float *p, float *div)
{
    sf_compat_check(sf_compute_divergence_and_pressure(sf_compat_ctx, u, v, p, div), "computeDivergenceAndPressure");
}
static inline void lastProject(float *u, float *v, float *p, float *div)
{
    sf_compat_check(sf_last_project(sf_compat_ctx, u, v, p, div), "lastProject");
}
static inline void dens_step(float *x, float *x0, float *u, float *v, float diff)
{
    sf_compat_check(sf_dens_step(sf_compat_ctx, x, x0, u, v, diff, sf_compat_dt, SF_COMPAT_ITERS), "dens_step");
}
static inline void vel_step(float *u, float *v, float *u0, float *v0, float visc, int z)
{
    (void)z; /* the reference only uses z to index its timing arrays (FluidSequential.c:195) */
    sf_compat_check(sf_vel_step(sf_compat_ctx, u, v, u0, v0, visc, sf_compat_dt, SF_COMPAT_ITERS), "vel_step");
}

#endif /* STABLEFLUIDS_COMPAT_H */
