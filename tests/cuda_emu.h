// cuda_emu.h -- TEST INFRASTRUCTURE: the warp-level CUDA intrinsics the product's device headers use, for HOST threads in
// lockstep.  32 std::threads are the lanes of one warp; every collective (__shfl_*_sync, __ballot_sync, __reduce_*_sync,
// __match_any_sync, __syncwarp) is "publish my value, barrier, read what I need, barrier", which is the lockstep the
// hardware provides; atomics are real atomics because the lanes really run concurrently.  Device headers compiled against
// this run exactly as written: a missing synchronisation becomes a data race between free-running threads, a wrong lane
// index a wrong result.  Not a product path; nothing under sibeliaz_b200/ includes this.
#pragma once
#include <algorithm>
#include <barrier>
#include <cstdint>
#include <cstring>
#include <type_traits>

// The device headers include <cuda_runtime.h>; on the host its include guard is pre-defined here, so that none of the
// toolkit's own definitions of __device__ & co. meet the ones below, and the two vector types are declared by hand.
#define __CUDA_RUNTIME_H__
struct alignas(8) int2 {
    int x, y;
};
struct alignas(16) int4 {
    int x, y, z, w;
};

namespace emu {
constexpr int kMaxWarps = 32;
inline std::barrier<> *warp_bar[kMaxWarps];
inline std::barrier<> *cta_bar;
inline uint64_t slot[kMaxWarps][32];
inline thread_local int lane, warp;

template <class T>
inline uint64_t to_bits(T v)
{
    uint64_t b = 0;
    static_assert(sizeof(T) <= 8, "");
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <class T>
inline T from_bits(uint64_t b)
{
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}
// publish, barrier, f(slots) -> result, barrier
template <class F>
inline auto collective(uint64_t mine, F f)
{
    slot[warp][lane] = mine;
    warp_bar[warp]->arrive_and_wait();
    auto r = f(slot[warp]);
    warp_bar[warp]->arrive_and_wait();
    return r;
}
} // namespace emu

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#ifndef __restrict__
#define __restrict__
#endif

inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
inline int2 make_int2(int x, int y) { return int2{x, y}; }

template <class T>
inline T __shfl_sync(unsigned, T v, int src)
{
    return emu::collective(emu::to_bits(v), [&](const uint64_t *s) { return emu::from_bits<T>(s[src & 31]); });
}
template <class T>
inline T __shfl_up_sync(unsigned, T v, unsigned d)
{
    return emu::collective(emu::to_bits(v), [&](const uint64_t *s) { return emu::lane >= (int)d ? emu::from_bits<T>(s[emu::lane - (int)d]) : v; });
}
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int m)
{
    return emu::collective(emu::to_bits(v), [&](const uint64_t *s) { return emu::from_bits<T>(s[(emu::lane ^ m) & 31]); });
}
inline unsigned __ballot_sync(unsigned, bool p)
{
    return emu::collective(p ? 1u : 0u, [&](const uint64_t *s) {
        unsigned m = 0;
        for (int l = 0; l < 32; l++) m |= (unsigned)(s[l] & 1u) << l;
        return m;
    });
}
inline bool __any_sync(unsigned mask, bool p) { return __ballot_sync(mask, p) != 0; }
inline bool __all_sync(unsigned mask, bool p) { return __ballot_sync(mask, p) == 0xFFFFFFFFu; }
template <class T>
inline unsigned __match_any_sync(unsigned, T v)
{
    return emu::collective(emu::to_bits(v), [&](const uint64_t *s) {
        unsigned m = 0;
        for (int l = 0; l < 32; l++) m |= (unsigned)(s[l] == s[emu::lane]) << l;
        return m;
    });
}
template <class T>
inline T __reduce_min_sync(unsigned, T v)
{
    return emu::collective(emu::to_bits(v), [&](const uint64_t *s) {
        T r = emu::from_bits<T>(s[0]);
        for (int l = 1; l < 32; l++) r = std::min(r, emu::from_bits<T>(s[l]));
        return r;
    });
}
template <class T>
inline T __reduce_max_sync(unsigned, T v)
{
    return emu::collective(emu::to_bits(v), [&](const uint64_t *s) {
        T r = emu::from_bits<T>(s[0]);
        for (int l = 1; l < 32; l++) r = std::max(r, emu::from_bits<T>(s[l]));
        return r;
    });
}
template <class T>
inline T __reduce_add_sync(unsigned, T v)
{
    return emu::collective(emu::to_bits(v), [&](const uint64_t *s) {
        T r = 0;
        for (int l = 0; l < 32; l++) r = (T)(r + emu::from_bits<T>(s[l]));
        return r;
    });
}
inline void __syncwarp(unsigned = 0xFFFFFFFFu) { emu::warp_bar[emu::warp]->arrive_and_wait(); }
inline void __syncthreads() { emu::cta_bar->arrive_and_wait(); }

template <class T>
inline T __ldg(const T *p)
{
    return *p;
}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }

template <class T>
inline T atomicCAS(T *p, T cmp, T val)
{
    __atomic_compare_exchange_n(p, &cmp, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
}
template <class T>
inline T atomicAdd(T *p, T v)
{
    return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
}
template <class T>
inline T atomicMax(T *p, T v)
{
    T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
    }
    return old;
}

// CUDA's min / max overloads on mixed integer types
template <class A, class B>
inline typename std::common_type<A, B>::type min(A a, B b)
{
    using C = typename std::common_type<A, B>::type;
    return (C)a < (C)b ? (C)a : (C)b;
}
template <class A, class B>
inline typename std::common_type<A, B>::type max(A a, B b)
{
    using C = typename std::common_type<A, B>::type;
    return (C)a > (C)b ? (C)a : (C)b;
}
