// Host-side internals shared by the C ABI (sf_api.cu) and the peer-memory slab driver (sf_slab.cu):
// the context, error plumbing, launch planning and the CUDA-graph cache.  Not installed.
#pragma once
#include <cuda_runtime.h>

#include <cstring>
#include <initializer_list>
#include <string>
#include <vector>

#include "../../include/stablefluids.h"
#include "sf_common.cuh"

namespace sf {

struct GraphKey {   // laid out without padding so memcmp is a valid equality
    uint64_t kind;
    const void *p[6];
    float f[4];
    int iters;
    int opts[7];
    bool operator==(const GraphKey &o) const { return std::memcmp(this, &o, sizeof(GraphKey)) == 0; }
};
struct GraphEntry {
    GraphKey key;
    cudaGraphExec_t exec;
    cudaGraph_t graph;
    unsigned long long kernels;
    unsigned long long last_use;
};

// Peer-memory view of a slab and its two neighbours (sf_slab.cu).  All fields of a slab live in ONE
// device allocation (the arena: nfields caller fields, the lin_solve scratch field, then the
// synchronisation flags), so that one pointer -- or one CUDA IPC handle -- makes a neighbour's whole
// slab addressable: field k of a neighbour is nbr.base + k * nbr.field_bytes.
struct SlabLink {
    int nfields = 0;              // caller fields; index nfields is the scratch field
    char *base = nullptr;         // local arena (nullptr: no arena, context is not peer-capable)
    size_t field_bytes = 0;
    SlabFlags *flags = nullptr;   // inside the arena, after the fields
    struct Nbr {
        bool present = false;
        bool ipc = false;         // base came from cudaIpcOpenMemHandle (closed in sf_destroy)
        char *base = nullptr;
        size_t field_bytes = 0;
        SlabFlags *flags = nullptr;
        int row_lo = 0, row_hi = 0, row_base = 0;
    } nbr[2];                     // 0 = up (smaller rows), 1 = down
    unsigned long long timeout_ns = 20000000000ull;   // barrier spin limit before the error bit is set
    bool barrier_valid = false;                 // see slab_barrier: back-to-back barriers collapse into one
    unsigned long long launches_at_barrier = 0;
    StripArgs *strip_table = nullptr;   // device: entry [field * 9 + rows] = StripArgs of a launch writing `field`
    int world_links() const { return (nbr[0].present ? 1 : 0) + (nbr[1].present ? 1 : 0); }
};

}  // namespace sf

struct sf_context {
    sf::Geom g;
    int device = 0;
    int halo = 0;
    cudaStream_t stream = nullptr;   // the stream results are ordered on (own or caller's)
    cudaStream_t work = nullptr;     // where launches go: == stream, or cap_stream while capturing
    cudaStream_t cap_stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    int arith = SF_ARITH_STRICT;
    int sweeps_opt = 0;
    int use_graph = 1;
    int force_generic = 0;
    int chunk_rows = 0;
    int staging = 0;
    float *scratch2 = nullptr;       // right-hand side of a solve whose add_source is fused into its first launch
    int fuse_sources = 1;            // SF_OPT_FUSE_SOURCES
    int advect_tile = 1;             // SF_OPT_ADVECT_TILE (1 = automatic: see refresh_advect_policy)
    bool advect_tile_live = true;    // automatic mode: what the launches enqueued now use
    unsigned int *tile_stats = nullptr;        // device: tiles served by the TMA box / by the gather fallback
    unsigned int tile_seen[2] = {0u, 0u};      // their values when the policy last looked
    // SF_OPT_OVERLAP_SOLVES: independent lin_solves of a step on streams of their own (see enqueue_step in sf_api.cu); a lane
    // owns everything a solve scribbles on
    struct SolveLane {
        cudaStream_t stream = nullptr;
        float *scratch = nullptr, *scratch2 = nullptr;
        unsigned *ticket = nullptr;
        cudaEvent_t fork = nullptr, join = nullptr;
    } lanes[2];
    int overlap = 1;
    int strip_balance = 1;           // SF_OPT_STRIP_BALANCE
    int wave_skew = 131103;          // SF_OPT_WAVE_SKEW (p0 * 1000 + p1); swept in profiles/r02/s15_*_skew_sweep.txt
    unsigned *ticket = nullptr;      // device word: start-order tickets of the CTAs of a Jacobi launch
    float *scratch = nullptr;        // lin_solve ping-pong partner (inside the arena for peer slabs)
    bool scratch_in_arena = false;
    sf::StealCtl *steal = nullptr;   // row-level work stealing between the warps of a Jacobi launch
    int steal_capacity = 0;
    int steal_opt = 30;              // SF_OPT_WORK_STEALING (percent; 0 = off)
    int steal_scope = 0;             // SF_OPT_STEAL_SCOPE
    int pressure_plan = 2;           // SF_OPT_PRESSURE_PLAN (2: odd launch counts + depth 8 for the pressure solves)
    int solver = SF_SOLVER_JACOBI;   // SF_OPT_SOLVER
    int omega_milli = 1000;          // SF_OPT_SOR_OMEGA_MILLI
    int rbgs_blocked = 0;            // SF_OPT_RBGS_BLOCKED
    bool steal_now = false;          // set by the drivers around the solves that are worth it (see lin_solve)
    float *red_f = nullptr;          // reduction outputs
    double *red_d = nullptr;
    unsigned long long launches = 0;
    unsigned long long tick = 0;
    bool capturing = false;
    std::string err;
    std::vector<sf::GraphEntry> graphs;
    // sf_step_host resources
    float *stage[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaStream_t h2d = nullptr, d2h = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    sf::SlabLink link;
};

namespace sf {

inline int fail(sf_context *c, int code, const char *what, cudaError_t e = cudaSuccess)
{
    if (c) {
        c->err = what;
        if (e != cudaSuccess) { c->err += ": "; c->err += cudaGetErrorString(e); }
    }
    return code;
}
#define SF_CUDA(ctx, call)                                                       \
    do {                                                                         \
        cudaError_t e_ = (call);                                                 \
        if (e_ != cudaSuccess) return sf::fail(ctx, SF_ERR_CUDA, #call, e_);     \
    } while (0)
#define SF_REQUIRE(ctx, cond, msg)                                               \
    do {                                                                         \
        if (!(cond)) return sf::fail(ctx, SF_ERR_INVALID, msg);                  \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

inline size_t field_cells(const sf_context *c) { return (size_t)c->g.rows * (size_t)c->g.G; }
inline bool is_full_grid(const sf_context *c) { return c->g.own_lo == 0 && c->g.own_hi == c->g.G; }
// a slab whose neighbours are reachable through peer memory: the step drivers exchange halos themselves
inline bool is_linked_slab(const sf_context *c) { return !is_full_grid(c) && c->link.base && c->link.world_links() > 0; }
inline bool stream_kernels_ok(const sf_context *c) { return jacobi_stream_supported(c->g) && !c->force_generic; }

// ---- implemented in sf_api.cu ----
int ensure_scratch(sf_context *c);
int arith_mode(const sf_context *c, float alpha, float beta);
int default_sweeps(const sf_context *c);
std::vector<int> plan_launches(int iters, int T, bool odd_ok = false);
// strip_rows > 0 (peer-memory slabs): the launch exchanges boundary strips of that height with the
// neighbours (fused into the kernel; see StripArgs in sf_common.cuh); xout must be an arena field
int one_jacobi_launch(sf_context *c, cudaStream_t st, int b, float *xout, const float *xin, const float *x0, float alpha,
                      float beta, int sweeps, int out_lo, int out_hi, int zero_guess, int strip_rows = 0, float *rhs_out = nullptr,
                      float src_dt = 0.0f);
int lin_solve(sf_context *c, int b, float *x, const float *x0, float alpha, float beta, int iters, int zero_guess);
int enqueue_dens_step(sf_context *c, float *x, float *x0, const float *u, const float *v, float diff, float dt, int iters);
int enqueue_project(sf_context *c, float *u, float *v, float *p, float *div, int iters);
int enqueue_vel_step(sf_context *c, float *u, float *v, float *u0, float *v0, float visc, float dt, int iters);
int enqueue_vel_diffuse(sf_context *c, int b, float *x, float *x0, float visc, float dt, int iters);
int enqueue_vel_tail(sf_context *c, float *u, float *v, float *u0, float *v0, float dt, int iters);
GraphKey make_key(const sf_context *c, int kind, std::initializer_list<const void *> ptrs, float f0, float f1, float f2, int iters);

// ---- implemented in sf_slab.cu (peer-memory slabs; every rank must issue the same call sequence) ----
// fuse_dt != nullptr: x0 is the RAW field and x the source (= the initial guess): the first launch forms x0 + *fuse_dt * x
// itself (see slab_sources_fusable), x0 is left as it was
int slab_lin_solve(sf_context *c, int b, float *x, const float *x0, float alpha, float beta, int iters, int zero_guess,
                   const float *fuse_dt = nullptr);
int slab_project(sf_context *c, float *u, float *v, float *p, float *div, int iters);
int slab_advect(sf_context *c, int b, float *d, const float *d0, const float *u, const float *v, float dt, bool trailing_barrier);
int slab_vel_step(sf_context *c, float *u, float *v, float *u0, float *v0, float visc, float dt, int iters);
int slab_dens_step(sf_context *c, float *x, float *x0, const float *u, const float *v, float diff, float dt, int iters);
int slab_prevalidate(sf_context *c, float coef, float dt);
void slab_release(sf_context *c);
const StripArgs *slab_strip_args(const sf_context *c, const float *xout, int rows);   // device pointer, or nullptr

// ---- CUDA graph cache ------------------------------------------------------------------------
// what the advect launches enqueued next get as `tile`
// (automatic mode needs the counters of a finished step, i.e. a host synchronisation, which a connected slab must not do
// in the middle of a collective sequence -- one thread may be driving several slabs: slabs take the tiles only on request)
inline int advect_tile_now(const sf_context *c)
{
    if (c->advect_tile != 1) return c->advect_tile;
    return (c->advect_tile_live && c->link.base == nullptr) ? 1 : 0;
}
int refresh_advect_policy(sf_context *c);
int ensure_lanes(sf_context *c);

template <class Body>
int run_graphed(sf_context *c, const GraphKey &key, Body body)
{
    if (!c->use_graph || c->capturing) return body();
    ++c->tick;
    GraphEntry *seen = nullptr;
    for (auto &e : c->graphs)
        if (e.key == key) {
            e.last_use = c->tick;
            if (e.exec) {
                SF_CUDA(c, cudaGraphLaunch(e.exec, c->stream));
                c->launches += e.kernels;
                return SF_OK;
            }
            seen = &e;
        }
    int rc = ensure_scratch(c);
    if (rc) return rc;
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(c->stream, &st);
    if (st != cudaStreamCaptureStatusNone) return body();   // caller is capturing already (lanes that exist are used, none are made)
    if ((rc = ensure_lanes(c))) return rc;
    if (!seen) {
        // First sighting of these arguments: launch directly (this also loads the kernels' modules
        // outside of any capture); the second call with the same arguments captures the graph.
        if (c->graphs.size() >= 16) {   // evict least recently used
            size_t victim = 0;
            for (size_t k = 1; k < c->graphs.size(); ++k)
                if (c->graphs[k].last_use < c->graphs[victim].last_use) victim = k;
            if (c->graphs[victim].exec) { cudaGraphExecDestroy(c->graphs[victim].exec); cudaGraphDestroy(c->graphs[victim].graph); }
            c->graphs.erase(c->graphs.begin() + victim);
        }
        c->graphs.push_back(GraphEntry{key, nullptr, nullptr, 0, c->tick});
        return body();
    }
    if ((rc = refresh_advect_policy(c))) return rc;   // the first (direct) run has shown whether the advect tiles fit: freeze the choice
    // Capture on a private stream: the caller's stream may be the legacy default stream, which
    // cannot be captured.  The instantiated graph is then launched on the caller's stream.
    if (!c->cap_stream) SF_CUDA(c, cudaStreamCreateWithFlags(&c->cap_stream, cudaStreamNonBlocking));
    const unsigned long long before = c->launches;
    c->capturing = true;
    cudaError_t e = cudaStreamBeginCapture(c->cap_stream, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) { c->capturing = false; return fail(c, SF_ERR_CUDA, "cudaStreamBeginCapture", e); }
    c->work = c->cap_stream;
    rc = body();
    c->work = c->stream;
    cudaGraph_t graph = nullptr;
    e = cudaStreamEndCapture(c->cap_stream, &graph);
    c->capturing = false;
    const unsigned long long kernels = c->launches - before;
    c->launches = before;
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return fail(c, SF_ERR_CUDA, "cudaStreamEndCapture", e);
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    if (e != cudaSuccess) { cudaGraphDestroy(graph); return fail(c, SF_ERR_CUDA, "cudaGraphInstantiate", e); }
    seen->exec = exec; seen->graph = graph; seen->kernels = kernels;
    SF_CUDA(c, cudaGraphLaunch(exec, c->stream));
    c->launches += kernels;
    return SF_OK;
}

}  // namespace sf
