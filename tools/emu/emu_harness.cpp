// Host harness around the generated build of the Jacobi kernels' source (tools/emu/gen_emu.py): restates the launch
// geometry of launch_jacobi_stream and the launch plan of lin_solve (csrc/sf_jacobi.cu, csrc/sf_api.cu), and runs
// jacobi_stream_kernel<T, MODE, VAR> warp by warp with 32 host threads per warp.  Exposed to Python through ctypes
// (tests/test_emu_cpu.py).  Test infrastructure: nothing in the product uses it.
#include "emu_prelude.h"

#include <thread>
#include <vector>

#include "_gen/emu_jacobi.inc"

namespace sf {
namespace {
float4 ring[WPC * (RING_X + RING_R) * 32 + 256];     // the kernel's `extern __shared__ float4 ring[]` (+ mbarrier words)
int g_wave_skew = 0;                                  // emu_set_wave_skew: p0 * 1000 + p1, see chunk_range (0 = equal chunks)
unsigned g_ticket[4] = {0, 0, 0, 0};
bool g_steal_variant = false;                         // emu_set_steal_variant: strict launches run VAR 3 / 7 instead of 0 / 6

template <int T, int MODE, int VAR>
void run_grid(const StreamArgs &A, int ctas)
{
    std::barrier<> bar(32);
    emu::warp_barrier = &bar;
    emu::g_dim = {(unsigned)ctas, 1, 1};
    emu::b_dim = {(unsigned)(WPC * 32), 1, 1};
    for (int cta = 0; cta < ctas; ++cta)
        for (int warp = 0; warp < WPC; ++warp) {       // warps never synchronise with each other: one at a time
            for (auto &v : ring) v = make_float4(NAN, NAN, NAN, NAN);    // stale shared memory is poison
            std::vector<std::thread> lanes;
            for (int l = 0; l < 32; ++l)
                lanes.emplace_back([&, l] {
                    emu::t_idx = {(unsigned)(warp * 32 + l), 0, 0};
                    emu::b_idx = {(unsigned)cta, 0, 0};
                    jacobi_stream_kernel<T, MODE, VAR>(A);
                });
            for (auto &t : lanes) t.join();
        }
}

template <int MODE, int VAR>
int run_T(int T, const StreamArgs &A, int ctas)
{
    switch (T) {
        case 1: if constexpr (VAR != 5) { run_grid<1, MODE, VAR>(A, ctas); return 0; } break;
        case 2: run_grid<2, MODE, VAR>(A, ctas); return 0;
        case 3: if constexpr (VAR != 5) { run_grid<3, MODE, VAR>(A, ctas); return 0; } break;
        case 4: run_grid<4, MODE, VAR>(A, ctas); return 0;
        case 5: if constexpr (VAR != 5) { run_grid<5, MODE, VAR>(A, ctas); return 0; } break;
        case 6: run_grid<6, MODE, VAR>(A, ctas); return 0;
        case 7: if constexpr (VAR != 5) { run_grid<7, MODE, VAR>(A, ctas); return 0; } break;
        case 8: if constexpr (VAR != 5) { run_grid<8, MODE, VAR>(A, ctas); return 0; } break;
    }
    return -1;
}

// launch_jacobi_stream (csrc/sf_jacobi.cu) for a full-grid context: same geometry, same StreamArgs
int launch(float *xout, const float *xin, const float *rhs, int N, int b, int mode, float alpha, float beta, int sweeps,
           int zero_guess, int chunk_rows, int rb, float omega, float *rhs_out = nullptr, float src_dt = 0.0f)
{
    StreamArgs A;
    std::memset(&A, 0, sizeof(A));
    const int G = N + 2;
    A.xin = xin; A.rhs = rhs; A.xout = xout;
    A.G = G; A.N = N; A.row_base = 0;
    A.a_lo = 1; A.a_hi = N + 1;
    A.write_top = 1; A.write_bot = 1;
    A.nbands = (G + VALID_W - 1) / VALID_W;
    A.zero_guess = zero_guess;
    A.rhs_out = rhs_out; A.src_dt = src_dt;
    A.alpha = alpha; A.div = make_div_const(beta);
    const double F = 1.0 + 4.0 * fabs((double)alpha);
    if (!rb) {
        double g = F / fabs((double)beta) * 1.000001;
        if (g < 1.0) g = 1.0;
        double hi = (double)SF_DIV_HI / (F * 1.01);
        for (int t = 0; t < sweeps; ++t) hi /= g;
        A.hi_in = (float)hi;
    } else {
        A.div.pad = omega;
        const double om = (double)omega;
        double gl = (fabs(1.0 - om) + om * F / fabs((double)beta)) * 1.000001;
        if (gl < 1.0) gl = 1.0;
        double hi = (double)SF_DIV_HI / (F * 1.01) / (1.0 + om);
        for (int t = 0; t < sweeps; ++t) hi /= gl;
        A.hi_in = (float)hi;
    }
    A.sx = (b == 1) ? -1.0f : 1.0f;
    A.sy = (b == 2) ? -1.0f : 1.0f;
    const int rows = A.a_hi - A.a_lo;
    int chunk = chunk_rows;
    if (chunk <= 0) {
        const bool heavy = (mode == MODE_STRICT || mode == MODE_IEEE || mode == MODE_PRESSURE) && sweeps >= 6;
        const int slots = 148 * (heavy ? 3 : 4) * WPC;
        int want = slots / A.nbands;
        if (want < 1) want = 1;
        chunk = (std::max(rows, 1) + want - 1) / want;
        const int min_chunk = 2 * sweeps > 8 ? 2 * sweeps : 8;
        if (chunk < min_chunk) chunk = min_chunk;
    }
    if (chunk > rows) chunk = std::max(rows, 1);
    A.chunk_rows = chunk;
    A.nchunks = rows > 0 ? (rows + chunk - 1) / chunk : 0;
    const int items = A.nbands * A.nchunks;
    const int ctas = (items + WPC - 1) / WPC;
    if (g_wave_skew > 0 && !rb && A.nchunks % 3 == 0) {
        // as launch_jacobi_stream sets it up, minus the "one full wave of 148 SMs" condition (the emulated grid is whatever
        // the problem needs): what is checked here is that the slot-class tickets and the unequal chunks still partition the rows
        const int r0 = chunk * (g_wave_skew / 1000) / 100, r1 = chunk * (g_wave_skew % 1000) / 100, r2 = 3 * chunk - r0 - r1;
        if (r0 >= r1 && r1 >= r2 && r2 >= 2 * sweeps) {
            A.skew = (unsigned)r0 | ((unsigned)r1 << 16);
            A.skew_cpw = (unsigned)(A.nchunks / 3);
            A.ticket = g_ticket;
        }
    }
    if (rb) {
        if (mode == MODE_PRESSURE) return run_T<MODE_PRESSURE, 5>(sweeps, A, ctas);
        return run_T<MODE_STRICT, 5>(sweeps, A, ctas);
    }
    if (g_steal_variant && mode == MODE_STRICT && !rb && items > 1) {
        // the work-stealing variants (VAR 3 / 7: scalar fields).  Warps run one after the other here, so nobody ever finds
        // a range to take over -- what this exercises is the variant's own streaming code (zero-row shortcut, polls)
        static std::vector<char> ctl_mem;
        const int capacity = 16384;
        ctl_mem.assign(sizeof(StealCtl) + (size_t)capacity * sizeof(StealSlot), 0);
        StealCtl *ctl = reinterpret_cast<StealCtl *>(ctl_mem.data());
        ctl->min_pct = 30;
        for (int k = 0; k < capacity; ++k) ctl->slots[k].pos = 0x3fffffff;
        if (items > capacity) return -1;
        A.steal = ctl;
        if (rhs_out != nullptr) {
            switch (sweeps) {
                case 5: run_grid<5, MODE_STRICT, 7>(A, ctas); return 0;
                case 6: run_grid<6, MODE_STRICT, 7>(A, ctas); return 0;
                case 7: run_grid<7, MODE_STRICT, 7>(A, ctas); return 0;
            }
            return -1;
        }
        return run_T<MODE_STRICT, 3>(sweeps, A, ctas);
    }
    if (rhs_out != nullptr) {      // fused add_source: the depths launch_stream_T builds (5, 6, 7), strict arithmetic
        switch (sweeps) {
            case 5: run_grid<5, MODE_STRICT, 6>(A, ctas); return 0;
            case 6: run_grid<6, MODE_STRICT, 6>(A, ctas); return 0;
            case 7: run_grid<7, MODE_STRICT, 6>(A, ctas); return 0;
        }
        return -1;
    }
    if (mode == MODE_PRESSURE) return run_T<MODE_PRESSURE, 0>(sweeps, A, ctas);
    return run_T<MODE_STRICT, 0>(sweeps, A, ctas);
}

// ---- peer-memory slabs on the host: p slabs of one grid, each with its own rows (+ HALO ghost rows a side), exchanging
// boundary strips exactly as the device does (strip warps store into the neighbour's ghost rows, post / wait on counters).
// The slabs run their k-th launch one after the other, so every wait finds its counter already posted.
struct EmuSlab {
    int own_lo = 0, own_hi = 0, row_base = 0, rows = 0;
    std::vector<float> x, x0, scratch, rhs;     // local fields: rows x G, NaN where nothing was delivered
    SlabFlags flags;
    StripArgs strips;
};

// launch_jacobi_stream (csrc/sf_jacobi.cu) for a connected slab: strips, interior segment, short chunks behind the strips
int slab_launch(EmuSlab &S, EmuSlab *up, EmuSlab *dn, float *xout, float *up_xout, float *dn_xout, const float *xin,
                const float *rhs, int N, int b, int mode, float alpha, float beta, int sweeps, int strip, int chunk_rows,
                int balance, float *rhs_out, float src_dt, int zero_guess)
{
    StreamArgs A;
    std::memset(&A, 0, sizeof(A));
    const int G = N + 2;
    A.xin = xin; A.rhs = rhs; A.xout = xout;
    A.zero_guess = zero_guess;
    A.G = G; A.N = N; A.row_base = S.row_base;
    const int st_top = up ? strip : 0, st_bot = dn ? strip : 0;
    S.strips = StripArgs();
    S.strips.o_lo = S.own_lo; S.strips.o_hi = S.own_hi;
    S.strips.error = &S.flags.error; S.strips.timeout_ns = 2000000000ull;
    EmuSlab *nb[2] = {up, dn};
    float *nb_x[2] = {up_xout, dn_xout};
    for (int dir = 0; dir < 2; ++dir) {
        if (!nb[dir]) continue;
        StripPort &P = S.strips.port[dir];
        P.rows = strip; P.xpeer = nb_x[dir]; P.peer_row_base = nb[dir]->row_base;
        P.inbox = &S.flags.strip_inbox[dir]; P.seq = &S.flags.strip_seq[dir]; P.arrive = &S.flags.strip_arrive[dir];
        P.nbr_inbox = &nb[dir]->flags.strip_inbox[1 - dir];
    }
    A.strips = (st_top > 0 || st_bot > 0) ? &S.strips : nullptr;
    A.a_lo = std::max(S.own_lo + st_top, 1);
    A.a_hi = std::min(S.own_hi - st_bot, N + 1);
    A.write_top = (S.own_lo == 0); A.write_bot = (S.own_hi == G);
    A.nbands = (G + VALID_W - 1) / VALID_W;
    A.rhs_out = rhs_out; A.src_dt = src_dt;
    A.alpha = alpha; A.div = make_div_const(beta);
    {
        const double F = 1.0 + 4.0 * fabs((double)alpha);
        double g = F / fabs((double)beta) * 1.000001;
        if (g < 1.0) g = 1.0;
        double hi = (double)SF_DIV_HI / (F * 1.01);
        for (int t = 0; t < sweeps; ++t) hi /= g;
        A.hi_in = (float)hi;
    }
    A.sx = (b == 1) ? -1.0f : 1.0f;
    A.sy = (b == 2) ? -1.0f : 1.0f;
    const int n_strip_items = (st_top > 0 ? A.nbands : 0) + (st_bot > 0 ? A.nbands : 0);
    const int rows = A.a_hi - A.a_lo;
    int chunk = chunk_rows;
    if (chunk <= 0) {
        const bool heavy = sweeps >= 6;
        const int slots = 148 * (heavy ? 3 : 4) * WPC;
        int want = slots / A.nbands;
        if (want < 1) want = 1;
        chunk = (std::max(rows, 1) + want - 1) / want;
        const int min_chunk = 2 * sweeps > 8 ? 2 * sweeps : 8;
        if (chunk < min_chunk) chunk = min_chunk;
    }
    if (chunk > rows) chunk = std::max(rows, 1);
    A.chunk_rows = chunk;
    A.nchunks = rows > 0 ? (rows + chunk - 1) / chunk : 0;
    int balanced = 0;
    if (n_strip_items > 0 && balance) {
        const int n = A.nchunks, n_short = (st_top > 0 ? 1 : 0) + (st_bot > 0 ? 1 : 0);
        const int cost = 2 * (std::max(st_top, st_bot) + 2 * sweeps);
        const int c1 = n > 0 ? (rows + n_short * cost + n - 1) / n : 0, c0 = c1 - cost;
        const int min_chunk = 2 * sweeps > 8 ? 2 * sweeps : 8;
        if (n > n_short && c0 >= min_chunk && c0 < 0x10000 && n_short * c0 + (n - n_short) * c1 >= rows) {
            A.chunk_rows = c1;
            A.skew = (unsigned)c0 | ((unsigned)n_short << 16);
            A.skew_cpw = 0u;
            balanced = 1;
        }
    }
    const int items = std::max(n_strip_items, A.nbands * A.nchunks);
    const int ctas = (items + WPC - 1) / WPC;
    if (mode == MODE_PRESSURE) {
        if (rhs_out) return -1;
        if (run_T<MODE_PRESSURE, 2>(sweeps, A, ctas)) return -1;
        return balanced;
    }
    if (g_steal_variant && A.nbands * A.nchunks > 1) {
        // the work-stealing variants on slabs (VAR 4 / 9: the density solve).  As in launch(): warps run one after the other, nobody
        // finds a range to take over; what runs is the variant's own streaming code behind the strips
        static std::vector<char> ctl_mem;
        const int capacity = 16384;
        ctl_mem.assign(sizeof(StealCtl) + (size_t)capacity * sizeof(StealSlot), 0);
        StealCtl *ctl = reinterpret_cast<StealCtl *>(ctl_mem.data());
        ctl->min_pct = 30;
        for (int k = 0; k < capacity; ++k) ctl->slots[k].pos = 0x3fffffff;
        if (items > capacity) return -1;
        A.steal = ctl;
        if (rhs_out != nullptr) {
            switch (sweeps) {
                case 5: run_grid<5, MODE_STRICT, 9>(A, ctas); return balanced;
                case 6: run_grid<6, MODE_STRICT, 9>(A, ctas); return balanced;
                case 7: run_grid<7, MODE_STRICT, 9>(A, ctas); return balanced;
            }
            return -1;
        }
        if (run_T<MODE_STRICT, 4>(sweeps, A, ctas)) return -1;
        return balanced;
    }
    if (rhs_out != nullptr) {
        switch (sweeps) {
            case 5: run_grid<5, MODE_STRICT, 8>(A, ctas); return balanced;
            case 6: run_grid<6, MODE_STRICT, 8>(A, ctas); return balanced;
            case 7: run_grid<7, MODE_STRICT, 8>(A, ctas); return balanced;
        }
        return -1;
    }
    if (run_T<MODE_STRICT, 2>(sweeps, A, ctas)) return -1;
    return balanced;
}
}  // namespace
}  // namespace sf

extern "C" {
// slab_lin_solve of csrc/sf_slab.cu on `world` slabs of one (N+2)^2 grid, fields given and returned as full grids.
// fuse != 0: x is the source field (and the initial guess), x0 the RAW field, the first launch forms x0 + dt * x (VAR 8).
// Returns the number of launches that ran with short chunks behind the strips (>= 0), or -1.
int emu_slab_lin_solve(int N, int world, int b, float *x, const float *x0, float alpha, float beta, int iters, int T,
                       int zero_guess, int chunk_rows, int balance, int fuse, float dt)
{
    using namespace sf;
    const int G = N + 2, H = HALO_X;
    const int mode = (alpha == 1.0f && beta == 4.0f) ? MODE_PRESSURE : MODE_STRICT;
    int L = (iters + T - 1) / T;
    if ((L & 1) && L + 1 <= iters) ++L;
    std::vector<int> plan(L, iters / L);
    for (int k = 0; k < iters % L; ++k) ++plan[k];
    const int maxT = *std::max_element(plan.begin(), plan.end());
    if (fuse && (zero_guess || plan[0] != maxT)) return -1;
    std::vector<EmuSlab> S(world);
    for (int r = 0; r < world; ++r) {
        EmuSlab &s = S[r];
        s.own_lo = (int)((long long)r * G / world); s.own_hi = (int)((long long)(r + 1) * G / world);
        if (s.own_hi - s.own_lo < 2 * maxT) return -1;
        s.row_base = s.own_lo - H; s.rows = s.own_hi - s.own_lo + 2 * H;
        const size_t cells = (size_t)s.rows * G;
        s.x.assign(cells, NAN); s.x0.assign(cells, NAN); s.scratch.assign(cells, NAN); s.rhs.assign(cells, NAN);
        std::memset(&s.flags, 0, sizeof(s.flags));
        // owned rows + what the exchange in front of the solve delivers: plan[0] ghost rows of the guess, maxT of the rhs
        auto fill = [&](std::vector<float> &dst, const float *src, int ghost) {
            for (int row = std::max(s.own_lo - ghost, 0); row < std::min(s.own_hi + ghost, G); ++row)
                std::memcpy(&dst[(size_t)(row - s.row_base) * G], src + (size_t)row * G, (size_t)G * sizeof(float));
        };
        if (!zero_guess) fill(s.x, x, plan[0]);
        fill(s.x0, x0, maxT);
    }
    int balanced = 0;
    for (size_t k = 0; k < plan.size(); ++k) {
        const int need = (k + 1 < plan.size()) ? plan[k + 1] : 1;
        const int strip = std::max(need, plan[k]);
        for (int r = 0; r < world; ++r) {
            EmuSlab &s = S[r];
            EmuSlab *up = r > 0 ? &S[r - 1] : nullptr, *dn = r + 1 < world ? &S[r + 1] : nullptr;
            // every slab ping-pongs in step: the output field of this launch is the same one on all of them
            auto out_of = [&](EmuSlab &t) { return (k % 2 == 0) ? t.scratch.data() : t.x.data(); };
            auto in_of = [&](EmuSlab &t) { return (k % 2 == 0) ? t.x.data() : t.scratch.data(); };
            const float *rhs = fuse ? (k == 0 ? s.x0.data() : s.rhs.data()) : s.x0.data();
            const int rc = slab_launch(s, up, dn, out_of(s), up ? out_of(*up) : nullptr, dn ? out_of(*dn) : nullptr, in_of(s), rhs, N, b,
                                       mode, alpha, beta, plan[k], strip, chunk_rows, balance,
                                       (fuse && k == 0) ? s.rhs.data() : nullptr, dt, (zero_guess && k == 0) ? 1 : 0);
            if (rc < 0) return -1;
            balanced += rc;
            if (s.flags.error) return -1;
        }
    }
    for (int r = 0; r < world; ++r) {
        EmuSlab &s = S[r];
        const float *res = (plan.size() % 2 == 0) ? s.x.data() : s.scratch.data();
        for (int row = s.own_lo; row < s.own_hi; ++row)
            std::memcpy(x + (size_t)row * G, res + (size_t)(row - s.row_base) * G, (size_t)G * sizeof(float));
    }
    return balanced;
}
void emu_set_steal_variant(int on) { sf::g_steal_variant = on != 0; }
void emu_set_wave_skew(int code) { sf::g_wave_skew = code; }
int emu_ticket_words_nonzero() { return (sf::g_ticket[0] | sf::g_ticket[1] | sf::g_ticket[2] | sf::g_ticket[3]) != 0; }
// lin_solve of csrc/sf_api.cu (Jacobi: plan_launches; red-black: three iterations per launch); result ends in x
int emu_lin_solve(int N, int b, float *x, const float *x0, float alpha, float beta, int iters, int T, int zero_guess,
                  int chunk_rows, int rb, float omega)
{
    using namespace sf;
    const size_t cells = (size_t)(N + 2) * (N + 2);
    std::vector<float> scratch(cells, NAN);
    const int mode = (alpha == 1.0f && beta == 4.0f) ? MODE_PRESSURE : MODE_STRICT;
    std::vector<int> plan;
    if (rb) {
        for (int done = 0; done < iters;) { const int k = std::min(3, iters - done); plan.push_back(2 * k); done += k; }
    } else {
        int L = (iters + T - 1) / T;
        if (!zero_guess && (L & 1) && L + 1 <= iters) ++L;
        plan.assign(L, iters / L);
        for (int k = 0; k < iters % L; ++k) ++plan[k];
    }
    float *cur = x, *nxt = scratch.data();
    if (!rb && zero_guess && (plan.size() & 1)) { cur = scratch.data(); nxt = x; }
    for (size_t k = 0; k < plan.size(); ++k) {
        if (launch(nxt, cur, x0, N, b, mode, alpha, beta, plan[k], zero_guess && k == 0, chunk_rows, rb, omega)) return -1;
        std::swap(cur, nxt);
    }
    if (cur != x) std::memcpy(x, cur, cells * sizeof(float));
    return 0;
}

// source_lin_solve of csrc/sf_api.cu, fused form: x = source field and initial guess (receives the result), x0 = the raw
// field (left untouched); the first launch forms x0 + dt * x and stores it to a separate right-hand-side field
int emu_source_lin_solve(int N, int b, float *x, const float *x0, float dt, float alpha, float beta, int iters, int T, int chunk_rows)
{
    using namespace sf;
    const size_t cells = (size_t)(N + 2) * (N + 2);
    std::vector<float> scratch(cells, NAN), rhs(cells, NAN);
    int L = (iters + T - 1) / T;
    if ((L & 1) && L + 1 <= iters) ++L;
    std::vector<int> plan(L, iters / L);
    for (int k = 0; k < iters % L; ++k) ++plan[k];
    float *cur = x, *nxt = scratch.data();
    for (size_t k = 0; k < plan.size(); ++k) {
        if (launch(nxt, cur, k == 0 ? x0 : rhs.data(), N, b, MODE_STRICT, alpha, beta, plan[k], 0, chunk_rows, 0, 1.0f,
                   k == 0 ? rhs.data() : nullptr, dt)) return -1;
        std::swap(cur, nxt);
    }
    if (cur != x) std::memcpy(x, cur, cells * sizeof(float));
    return 0;
}
}
