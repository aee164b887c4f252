// Host harness around the generated build of csrc/sf_stages.cu (tools/emu/gen_emu.py): the stage kernels' own source,
// launched with the geometry of the launch_* wrappers for a full-grid context, 32 host threads per warp.
#include "emu_prelude.h"

#include <thread>
#include <vector>

#include "_gen/emu_stages.inc"

namespace sf {
namespace {
template <class Body>
void run_kernel(dim3 grid, dim3 block, Body body)
{
    const int threads = (int)(block.x * block.y * block.z);
    emu::g_dim = {grid.x, grid.y, grid.z};
    emu::b_dim = {block.x, block.y, block.z};
    for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx)
            for (int w = 0; w * 32 < threads; ++w) {
                const int count = std::min(32, threads - 32 * w);
                std::barrier<> bar(count);
                emu::warp_barrier = &bar;
                std::vector<std::thread> lanes;
                for (int l = 0; l < count; ++l)
                    lanes.emplace_back([&, l] {
                        const unsigned t = (unsigned)(32 * w + l);
                        emu::t_idx = {t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
                        emu::b_idx = {bx, by, 0};
                        body();
                        bar.arrive_and_drop();      // a lane that has returned no longer takes part in collectives
                    });
                for (auto &t : lanes) t.join();
            }
}
Geom full(int N) { Geom g; g.N = N; g.G = N + 2; g.row_base = 0; g.own_lo = 0; g.own_hi = N + 2; g.rows = N + 2; return g; }
}  // namespace
}  // namespace sf

using namespace sf;
extern "C" {
void emu_set_bnd(int N, int b, float *x)
{
    const Geom g = full(N);
    run_kernel(dim3((N + 255) / 256), dim3(256), [&] { set_bnd_kernel(x, g, b == 1 ? -1.0f : 1.0f, b == 2 ? -1.0f : 1.0f); });
}
void emu_add_source(int N, float *x, const float *s, float dt)
{
    const Geom g = full(N);
    AddSrcArgs A;
    std::memset(&A, 0, sizeof(A));
    A.x[0] = x; A.s[0] = s;
    A.first = 0; A.count = (size_t)g.G * g.G; A.dt = dt;
    const bool vec = (A.count % 4 == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)s % 16 == 0);
    A.vec = vec ? 1 : 0;
    const size_t work = vec ? A.count / 4 : A.count;
    size_t blocks = (work + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    run_kernel(dim3((unsigned)blocks, 1), dim3(256), [&] { add_source_kernel(A); });
}
void emu_advect(int N, int b, float *d, const float *d0, const float *u, const float *v, float dt)
{
    const Geom g = full(N);
    const int rows = interior_row_count(g);
    const float dt0 = dt * (float)g.N;
    if (g.G % 4 == 0) {
        const PeerView sA{d0, nullptr, nullptr};
        run_kernel(lanes_grid(g, rows), dim3(32, 8), [&] { advect_lanes_kernel<1, false>(d, nullptr, sA, sA, u, v, g, PeerGeom(), dt0, b); });
    } else {
        const dim3 block(64, 4);
        run_kernel(cell_grid(g, block, rows), block, [&] { advect_kernel<1>(d, nullptr, d0, nullptr, u, v, g, dt0, b); });
    }
}
void emu_advect_uv(int N, float *du, float *dv, const float *u0, const float *v0, float dt)
{
    const Geom g = full(N);
    const int rows = interior_row_count(g);
    const float dt0 = dt * (float)g.N;
    if (g.G % 4 == 0) {
        const PeerView sA{u0, nullptr, nullptr}, sB{v0, nullptr, nullptr};
        run_kernel(lanes_grid(g, rows), dim3(32, 8), [&] { advect_lanes_kernel<2, false>(du, dv, sA, sB, u0, v0, g, PeerGeom(), dt0, 1); });
    } else {
        const dim3 block(64, 4);
        run_kernel(cell_grid(g, block, rows), block, [&] { advect_kernel<2>(du, dv, u0, v0, u0, v0, g, dt0, 1); });
    }
}
void emu_divergence(int N, const float *u, const float *v, float *p, float *div, int write_p)
{
    const Geom g = full(N);
    const int rows = interior_row_count(g);
    const float h = 1.0f / (float)g.N;
    const float scale = -0.5f * h;
    if (row4_ok(g, {u, v, p, div})) {
        const dim3 b4(32, 8);
        run_kernel(row4_grid(g, b4, rows), b4, [&] { divergence4_kernel(u, v, p, div, g, scale, write_p); });
    } else {
        const dim3 block(64, 4);
        run_kernel(cell_grid(g, block, rows), block, [&] { divergence_kernel(u, v, p, div, g, scale, write_p); });
    }
}
void emu_last_project(int N, float *u, float *v, const float *p)
{
    const Geom g = full(N);
    const int rows = interior_row_count(g);
    const float h = 1.0f / (float)g.N;
    if (row4_ok(g, {u, v, p})) {
        const dim3 b4(32, 8);
        run_kernel(row4_grid(g, b4, rows), b4, [&] { last_project4_kernel(u, v, p, g, h); });
    } else {
        const dim3 block(64, 4);
        run_kernel(cell_grid(g, block, rows), block, [&] { last_project_kernel(u, v, p, g, h); });
    }
}
// rbgs_solve of csrc/sf_api.cu with launch_rbgs_half_sweep of csrc/sf_solvers.cu: iters x { red, black, set_bnd(b) }, in place
void emu_rbgs(int N, int b, float *x, const float *x0, float alpha, float beta, int iters, float omega)
{
    const Geom g = full(N);
    const DivConst dc = make_div_const(beta);
    const int relax = (omega != 1.0f) ? 1 : 0;
    const bool pressure = (alpha == 1.0f && beta == 4.0f);
    const dim3 block(64, 4), grid(((g.N + 1) / 2 + 63) / 64, (g.N + 3) / 4);
    for (int k = 0; k < iters; ++k) {
        for (int colour = 0; colour < 2; ++colour) {
            if (pressure) run_kernel(grid, block, [&] { rbgs_half_sweep_kernel<MODE_PRESSURE>(x, x0, g, colour, alpha, dc, omega, relax); });
            else run_kernel(grid, block, [&] { rbgs_half_sweep_kernel<MODE_STRICT>(x, x0, g, colour, alpha, dc, omega, relax); });
        }
        run_kernel(dim3((N + 255) / 256), dim3(256), [&] { set_bnd_kernel(x, g, b == 1 ? -1.0f : 1.0f, b == 2 ? -1.0f : 1.0f); });
    }
}
}
