// Host stand-ins for what the Jacobi kernels' source uses from CUDA (see gen_emu.py).  A warp = 32 host threads; every
// shuffle / vote / __syncwarp is a rendezvous of all 32 at a barrier, which is exactly the contract of the *_sync
// intrinsics under a full mask.  Arithmetic intrinsics map to the host's IEEE binary32/binary64 operations (compile with
// -ffp-contract=off, no -ffast-math): __fadd_rn = +, __fmaf_rn = fmaf (correctly rounded), __fdiv_rn = /.
#pragma once
#include <cuda_runtime.h>   // float2/float4/uint3/dim3 and the make_* helpers (host-usable headers)

#include <algorithm>
#include <barrier>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#undef __device__
#undef __global__
#undef __host__
#undef __shared__
#undef __forceinline__
#undef __noinline__
#undef __launch_bounds__
#undef __restrict__
#define __device__
#define __global__
#define __host__
#define __shared__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __restrict__

namespace emu {
struct Idx { unsigned x, y, z; };
inline thread_local Idx t_idx{0, 0, 0}, b_idx{0, 0, 0};
inline Idx g_dim{1, 1, 1}, b_dim{128, 1, 1};
inline std::barrier<> *warp_barrier = nullptr;     // 32 participants: the lanes of the warp being run
inline uint64_t exchange[32];
inline int lane() { return (int)(t_idx.x & 31u); }
inline unsigned long long now_ns()
{
    return (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
template <class V>
inline V shfl_from(V v, int src)
{
    static_assert(sizeof(V) <= 8, "");
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(V));
    exchange[lane()] = raw;
    warp_barrier->arrive_and_wait();
    const uint64_t got = exchange[src];
    warp_barrier->arrive_and_wait();
    V r;
    std::memcpy(&r, &got, sizeof(V));
    return r;
}
}  // namespace emu

#define threadIdx (emu::t_idx)
#define blockIdx (emu::b_idx)
#define gridDim (emu::g_dim)
#define blockDim (emu::b_dim)

template <class V> inline V __shfl_up_sync(unsigned, V v, int d) { const int l = emu::lane(); return emu::shfl_from(v, l - d >= 0 ? l - d : l); }
template <class V> inline V __shfl_down_sync(unsigned, V v, int d) { const int l = emu::lane(); return emu::shfl_from(v, l + d <= 31 ? l + d : l); }
template <class V> inline V __shfl_xor_sync(unsigned, V v, int m) { return emu::shfl_from(v, emu::lane() ^ m); }
template <class V> inline V __shfl_sync(unsigned, V v, int src) { return emu::shfl_from(v, src & 31); }
inline int __any_sync(unsigned, int pred)
{
    emu::exchange[emu::lane()] = pred ? 1u : 0u;
    emu::warp_barrier->arrive_and_wait();
    int r = 0;
    for (int k = 0; k < 32; ++k) r |= (int)emu::exchange[k];
    emu::warp_barrier->arrive_and_wait();
    return r;
}
inline int __all_sync(unsigned, int pred)
{
    emu::exchange[emu::lane()] = pred ? 1u : 0u;
    emu::warp_barrier->arrive_and_wait();
    int r = 1;
    for (int k = 0; k < 32; ++k) r &= (int)emu::exchange[k];
    emu::warp_barrier->arrive_and_wait();
    return r;
}
inline unsigned __reduce_or_sync(unsigned, unsigned v)
{
    emu::exchange[emu::lane()] = v;
    emu::warp_barrier->arrive_and_wait();
    unsigned r = 0;
    for (int k = 0; k < 32; ++k) r |= (unsigned)emu::exchange[k];
    emu::warp_barrier->arrive_and_wait();
    return r;
}
inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier->arrive_and_wait(); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __nanosleep(unsigned) {}
inline size_t __cvta_generic_to_shared(const void *p) { return (size_t)p; }

inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
inline float __double2float_rn(double a) { return (float)a; }
inline unsigned __float_as_uint(float a) { unsigned u; std::memcpy(&u, &a, 4); return u; }
inline float __uint_as_float(unsigned u) { float a; std::memcpy(&a, &u, 4); return a; }
template <class V> inline V __ldg(const V *p) { return *p; }
inline int atomicCAS(int *p, int cmp, int val) { __atomic_compare_exchange_n(p, &cmp, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }
inline unsigned long long atomicCAS(unsigned long long *p, unsigned long long cmp, unsigned long long val) { __atomic_compare_exchange_n(p, &cmp, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }
inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline double atomicAdd(double *p, double v) { double o = *p; *p = o + v; return o; }       // (reductions are not emulated)
inline int atomicMax(int *p, int v) { int o = *p; if (v > o) *p = v; return o; }
inline int __float_as_int(float a) { int i; std::memcpy(&i, &a, 4); return i; }
inline void __syncthreads() { std::abort(); }                                               // CTA-wide barriers: not emulated
using std::max;
using std::min;
