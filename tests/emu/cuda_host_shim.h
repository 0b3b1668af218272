// Host stand-ins for the handful of CUDA device intrinsics the ensemble kernel body uses, so that
// tests/emu/ens_emu.cpp can compile esp32-fluid-simulation_b200/csrc/ensemble_reg.cuh — the SAME source
// the GPU runs — with g++ and execute it on CPU threads.  TEST INFRASTRUCTURE ONLY (never part of the
// product library).  Build with -ffp-contract=off: every __f*_rn below must stay one IEEE operation.
#pragma once

#include <cuda_runtime.h>   // float2, make_float2, empty __host__/__device__, __forceinline__

#include <cmath>
#include <cstdint>
#include <cstring>

static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
// cvt.rzi.u32.f32 saturates; NaN -> 0
static inline unsigned int __float2uint_rz(float x)
{
    if (!(x > 0.0f)) return 0u;
    if (x >= 4294967296.0f) return 0xffffffffu;
    return (unsigned int)x;
}
static inline float __uint2float_rn(unsigned int u) { return (float)u; }
static inline unsigned int __float_as_uint(float f) { unsigned int u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned int u) { float f; std::memcpy(&f, &u, 4); return f; }
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline int atomicCAS(int *a, int cmp, int val) { const int old = *a; if (old == cmp) *a = val; return old; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
namespace fs { static inline void prefetch_l2(const void *) {} }
