// cuda_shim.h -- just enough of the CUDA C++ dialect to compile gym_b200/csrc/{rng,b2lite,lunar,walker}.cuh with a
// host compiler.  TEST INFRASTRUCTURE: tests/hostsim builds the device solver source for the CPU so that the
// `-m "not gpu"` suite can run the very code the kernels run against the C oracle (see hostsim.cpp).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#define __device__
#define __host__
#define __global__
#define __constant__
#define __forceinline__ inline
#define __noinline__

static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }

struct ulonglong2 { unsigned long long x, y; };
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { ulonglong2 r = {x, y}; return r; }

// warp votes of a one-lane "warp"
static inline int __any_sync(unsigned, int p) { return p; }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
