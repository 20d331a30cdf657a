// Host shim so that gobblet_core.cuh compiles with g++ for tests/emul (TEST ONLY: lets the per-lane
// device functions and the warp staging be checked against the oracle on a box without a GPU).
#pragma once
#include <stdint.h>
#include <algorithm>
#define __device__
#define __host__
#define __forceinline__ inline
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
struct ulonglong2 { unsigned long long x, y; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return {x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return {x, y, z, w}; }
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return {x, y}; }
static inline int __ffs(uint32_t v) { return v ? __builtin_ctz(v) + 1 : 0; }
static inline int __ffsll(long long v) { return v ? __builtin_ctzll((unsigned long long)v) + 1 : 0; }
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t s) {
    s &= 31u;
    return s ? (hi << s) | (lo >> (32u - s)) : hi;
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) {
    s &= 31u;
    return s ? (lo >> s) | (hi << (32u - s)) : lo;
}
using std::max;
using std::min;
// two-pass lockstep emulation of __shfl_up_sync(.., 1): pass 0 records, pass 1 replays
struct ShflCtx { int pass, call, lane; uint32_t rec[8][32]; };
extern thread_local ShflCtx g_shfl;
static inline uint32_t __shfl_up_sync(uint32_t, uint32_t v, int delta) {
    int c = g_shfl.call++;
    if (g_shfl.pass == 0) { g_shfl.rec[c][g_shfl.lane] = v; return v; }
    return g_shfl.lane >= delta ? g_shfl.rec[c][g_shfl.lane - delta] : v;
}
static inline void __stcs(uint4 *p, uint4 v) { *p = v; }
static inline void __syncwarp() {}   // lanes are emulated one after the other
