// Shared device/host helpers for libmpc_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mpc_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libmpc_b200 is written for sm_100a (B200) only"
#endif

#define MPC_HD __host__ __device__ __forceinline__
// per-step (not per-stage-visit) pieces of the fused loops: kept out of line so that their register needs do not
// spill the interior-point sweeps they sit next to
#define MPC_HD_COLD __host__ __device__ __noinline__

namespace mpc {

// ---------------------------------------------------------------- error plumbing (host)
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

#define MPC_REQUIRE(cond, code, ...) \
  do {                               \
    if (!(cond)) return ::mpc::fail((code), __VA_ARGS__); \
  } while (0)

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// ---------------------------------------------------------------- vector global access (device)
// B200 has 256-bit global loads/stores (LDG.E.256 / STG.E.256).  CNT elements of T, contiguous.
// `VEC` = number of bytes guaranteed aligned (32, 16, or sizeof(T)).
template <typename T>
__device__ __forceinline__ T ldg(const T* p) {
  return __ldg(p);
}

__device__ __forceinline__ void ld256(const double* p, double* r) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(r[0]), "=d"(r[1]), "=d"(r[2]), "=d"(r[3])
               : "l"(p));
}
__device__ __forceinline__ void ld256(const float* p, float* r) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]),
                 "=f"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void st256(double* p, const double* r) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(r[0]), "d"(r[1]), "d"(r[2]),
               "d"(r[3])
               : "memory");
}
__device__ __forceinline__ void st256(float* p, const float* r) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(r[0]), "f"(r[1]),
               "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7])
               : "memory");
}
__device__ __forceinline__ void ld128(const double* p, double* r) {
  double2 v = __ldg(reinterpret_cast<const double2*>(p));
  r[0] = v.x;
  r[1] = v.y;
}
__device__ __forceinline__ void ld128(const float* p, float* r) {
  float4 v = __ldg(reinterpret_cast<const float4*>(p));
  r[0] = v.x;
  r[1] = v.y;
  r[2] = v.z;
  r[3] = v.w;
}
__device__ __forceinline__ void st128(double* p, const double* r) {
  *reinterpret_cast<double2*>(p) = make_double2(r[0], r[1]);
}
__device__ __forceinline__ void st128(float* p, const float* r) {
  *reinterpret_cast<float4*>(p) = make_float4(r[0], r[1], r[2], r[3]);
}

// Load CNT contiguous elements into registers; ALIGN = guaranteed byte alignment of p.
template <typename T, int CNT, int ALIGN>
MPC_HD void load_row(const T* __restrict__ p, T* r) {
#ifdef __CUDA_ARCH__
  constexpr int E32 = 32 / (int)sizeof(T);
  constexpr int E16 = 16 / (int)sizeof(T);
  int i = 0;
  if constexpr (ALIGN >= 32) {
#pragma unroll
    for (; i + E32 <= CNT; i += E32) ld256(p + i, r + i);
  }
  if constexpr (ALIGN >= 16) {
#pragma unroll
    for (; i + E16 <= CNT; i += E16) ld128(p + i, r + i);
  }
#pragma unroll
  for (; i < CNT; ++i) r[i] = __ldg(p + i);
#else
  for (int i = 0; i < CNT; ++i) r[i] = p[i];
#endif
}

template <typename T, int CNT, int ALIGN>
MPC_HD void store_row(T* __restrict__ p, const T* r) {
#ifdef __CUDA_ARCH__
  constexpr int E32 = 32 / (int)sizeof(T);
  constexpr int E16 = 16 / (int)sizeof(T);
  int i = 0;
  if constexpr (ALIGN >= 32) {
#pragma unroll
    for (; i + E32 <= CNT; i += E32) st256(p + i, r + i);
  }
  if constexpr (ALIGN >= 16) {
#pragma unroll
    for (; i + E16 <= CNT; i += E16) st128(p + i, r + i);
  }
#pragma unroll
  for (; i < CNT; ++i) p[i] = r[i];
#else
  for (int i = 0; i < CNT; ++i) p[i] = r[i];
#endif
}

template <typename T>
MPC_HD T fma_(T a, T b, T c);
template <>
MPC_HD double fma_<double>(double a, double b, double c) {
  return fma(a, b, c);
}
template <>
MPC_HD float fma_<float>(float a, float b, float c) {
  return fmaf(a, b, c);
}

// reciprocal: one MUFU seed + Newton steps on the device (correctly rounded), plain division on the host
MPC_HD double rcp_(double x) {
#ifdef __CUDA_ARCH__
  // MUFU.RCP64H seed (relative error <= 2^-23) + two Newton steps: within ~1 ulp of 1/x for normal, finite x -- every
  // reciprocal on these paths is of a positive, normal quantity (slacks, complementarity products, R + B'PB, det C).
  // __drcp_rn adds a fifth DFMA, a range check and a slow-path call for correct rounding and denormals: 11 instructions
  // instead of 5, and it was 11 % of the instructions / 16 % of the stall samples of the interior-point kernels (ncu).
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
#else
  return 1.0 / x;
#endif
}
MPC_HD float rcp_(float x) {
#ifdef __CUDA_ARCH__
  return __frcp_rn(x);
#else
  return 1.0f / x;
#endif
}

// compile-time alignment (bytes, <= 32) of a densely packed row of CNT elements of T whose
// base pointer is 32-byte aligned
template <typename T, int CNT>
struct RowAlign {
  static constexpr int bytes = CNT * (int)sizeof(T);
  static constexpr int value = (bytes % 32 == 0) ? 32 : (bytes % 16 == 0) ? 16 : (int)sizeof(T);
};

// Largest alignment (32/16/elem) shared by a base pointer and a per-scenario stride in elements.
template <typename T>
inline int common_align(const void* p, int64_t stride_elems) {
  int a = 32;
  while (a > (int)sizeof(T)) {
    if (aligned(p, a) && ((stride_elems * (int64_t)sizeof(T)) % a) == 0) break;
    a >>= 1;
  }
  return a;
}

}  // namespace mpc
