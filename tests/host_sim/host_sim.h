// host_sim.h -- TEST INFRASTRUCTURE ONLY. Plain C++ stand-ins for the CUDA intrinsics used by the
// device headers (chess.cuh, stream.cuh), so that the CPU suite can run the *device functions
// themselves* against the oracle and the golden vectors without a GPU. Nothing in the product
// (libnnuepack.so) is built with NNP_HOST_SIM; the shipped library has no CPU path.
#pragma once
#include <algorithm>
#include <cstdint>

#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__

using std::max;
using std::min;

inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __popc(unsigned int v) { return __builtin_popcount(v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
inline unsigned int __brev(unsigned int v)
{
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
    return __builtin_bswap32(v);
}
inline unsigned long long __brevll(unsigned long long v)
{
    return ((unsigned long long)__brev((unsigned)v) << 32) | __brev((unsigned)(v >> 32));
}
inline unsigned int __byte_perm(unsigned int x, unsigned int y, unsigned int s)
{
    const unsigned long long src = ((unsigned long long)y << 32) | x;
    unsigned int r = 0;
    for (int i = 0; i < 4; ++i) {
        const unsigned sel = (s >> (4 * i)) & 7;  // (the sign-replication bit 3 is not used by the device code)
        r |= (unsigned)((src >> (8 * sel)) & 0xFF) << (8 * i);
    }
    return r;
}
inline unsigned int __funnelshift_l(unsigned int lo, unsigned int hi, unsigned int s)
{
    s &= 31;
    return (unsigned int)(((((unsigned long long)hi << 32) | lo) << s) >> 32);
}
inline unsigned int __funnelshift_lc(unsigned int lo, unsigned int hi, unsigned int s)
{
    s = s > 32 ? 32 : s;
    if (s == 32) return lo;
    return (unsigned int)(((((unsigned long long)hi << 32) | lo) << s) >> 32);
}
inline unsigned int __funnelshift_r(unsigned int lo, unsigned int hi, unsigned int s)
{
    s &= 31;
    return (unsigned int)((((unsigned long long)hi << 32) | lo) >> s);
}
inline unsigned int __funnelshift_rc(unsigned int lo, unsigned int hi, unsigned int s)
{
    s = s > 32 ? 32 : s;
    if (s == 32) return hi;
    return (unsigned int)((((unsigned long long)hi << 32) | lo) >> s);
}

struct uint2 {
    unsigned int x, y;
};
struct uint4 {
    unsigned int x, y, z, w;
};
inline uint2 make_uint2(unsigned int x, unsigned int y) { return uint2{x, y}; }
inline uint4 make_uint4(unsigned int x, unsigned int y, unsigned int z, unsigned int w) { return uint4{x, y, z, w}; }
#define __restrict__
