/*
 * A C driver in the shape of the reference's main() (FluidSequential.c:273-334: allocate six
 * fields, initialise sources, loop { refresh sources; vel_step; dens_step }, report the mean step
 * time) on top of libstablefluids_b200.so, using the reference's own function names through
 * stablefluids_compat.h.
 *
 *   gcc -O2 -Iinclude examples/fluid_main.c -Lfluidsimulationcuda_b200 -lstablefluids_b200 \
 *       -Wl,-rpath,$PWD/fluidsimulationcuda_b200 -o examples/fluid_main
 *   ./examples/fluid_main [N=1022] [steps=50]
 */
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include "stablefluids_compat.h"

#define DT 0.016f   /* FluidSequential.c:7-9 */
#define VIS 0.0025f
#define DIFF 0.1f

static double now(void)
{
    struct timespec ts;
    timespec_get(&ts, TIME_UTC);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int main(int argc, char **argv)
{
    const int N = argc > 1 ? atoi(argv[1]) : 1022;
    const int steps = argc > 2 ? atoi(argv[2]) : 50;
    const size_t cells = (size_t)(N + 2) * (size_t)(N + 2);
    sf_context *ctx;
    if (sf_create(&ctx, N, 0) != SF_OK) { fprintf(stderr, "no CUDA device / bad N\n"); return 1; }
    sf_compat_bind(ctx, DT);

    float *u, *v, *u_prev, *v_prev, *dens, *dens_prev;
    sf_alloc_field(ctx, &u); sf_alloc_field(ctx, &v); sf_alloc_field(ctx, &u_prev);
    sf_alloc_field(ctx, &v_prev); sf_alloc_field(ctx, &dens); sf_alloc_field(ctx, &dens_prev);
    sf_init_synthetic(ctx, 1, dens, dens_prev, u, u_prev, v, v_prev);   /* initializeParameters (:244-271) */

    double t0 = 0.0;
    for (int z = 0; z < steps; ++z) {
        if (z == 1) { sf_synchronize(ctx); t0 = now(); }                /* first step pays graph capture */
        if (z > 0) sf_init_sources(ctx, 1 + z, dens_prev, u_prev, v_prev);  /* the loop's source refresh (:298-302) */
        vel_step(u, v, u_prev, v_prev, VIS, z);                        /* :305 */
        dens_step(dens, dens_prev, u, v, DIFF);                        /* :306 */
    }
    sf_synchronize(ctx);
    const double per_step = steps > 1 ? (now() - t0) / (steps - 1) : 0.0;

    float *h = (float *)malloc(cells * sizeof(float));
    sf_download(ctx, h, dens);
    double sum = 0.0;
    for (size_t i = 0; i < cells; ++i) sum += h[i];
    printf("N %d steps %d  Tot %f s/step  sum(dens) %.9g\n", N, steps, per_step, sum);
    free(h);
    sf_free_field(ctx, u); sf_free_field(ctx, v); sf_free_field(ctx, u_prev);
    sf_free_field(ctx, v_prev); sf_free_field(ctx, dens); sf_free_field(ctx, dens_prev);
    sf_destroy(ctx);
    return 0;
}
