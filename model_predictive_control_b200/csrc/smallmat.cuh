// Register-resident small dense algebra for one scenario per thread.
// All loops are compile-time bounded and fully unrolled so every matrix lives in registers.
#pragma once

#include "common.cuh"

namespace mpc {

// C[MxN] (+)= A[MxK] * B[KxN], row-major register arrays.
template <typename T, int M, int K, int N, bool ACC>
MPC_HD void mm(const T* A, const T* B, T* C) {
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) {
      T acc = ACC ? C[i * N + j] : T(0);
#pragma unroll
      for (int k = 0; k < K; ++k) acc = fma_<T>(A[i * K + k], B[k * N + j], acc);
      C[i * N + j] = acc;
    }
}

// C[MxN] (+)= A'[MxK] * B[KxN] with A stored as [KxM].
template <typename T, int M, int K, int N, bool ACC>
MPC_HD void mtm(const T* A, const T* B, T* C) {
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) {
      T acc = ACC ? C[i * N + j] : T(0);
#pragma unroll
      for (int k = 0; k < K; ++k) acc = fma_<T>(A[k * M + i], B[k * N + j], acc);
      C[i * N + j] = acc;
    }
}

// y[M] (+)= A[MxN] x[N]
template <typename T, int M, int N, bool ACC>
MPC_HD void mv(const T* A, const T* x, T* y) {
#pragma unroll
  for (int i = 0; i < M; ++i) {
    T acc = ACC ? y[i] : T(0);
#pragma unroll
    for (int j = 0; j < N; ++j) acc = fma_<T>(A[i * N + j], x[j], acc);
    y[i] = acc;
  }
}

// x' A x
template <typename T, int N>
MPC_HD T quad(const T* A, const T* x) {
  T acc = T(0);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    T r = T(0);
#pragma unroll
    for (int j = 0; j < N; ++j) r = fma_<T>(A[i * N + j], x[j], r);
    acc = fma_<T>(x[i], r, acc);
  }
  return acc;
}

// X[MxN] = -S^-1 G with S[MxM] (symmetric positive definite in every use: R + B'PB), G[MxN].
// Unpivoted Gaussian elimination, unrolled; M is 1 or 2 on the fast path (any small M works).
template <typename T, int M, int N>
MPC_HD void neg_solve(T* S, T* G) {
#pragma unroll
  for (int p = 0; p < M; ++p) {
    const T inv = rcp_(S[p * M + p]);
#pragma unroll
    for (int j = p + 1; j < M; ++j) S[p * M + j] *= inv;
#pragma unroll
    for (int j = 0; j < N; ++j) G[p * N + j] *= inv;
#pragma unroll
    for (int r = 0; r < M; ++r) {
      if (r == p) continue;
      const T f = S[r * M + p];
#pragma unroll
      for (int j = p + 1; j < M; ++j) S[r * M + j] = fma_<T>(-f, S[p * M + j], S[r * M + j]);
#pragma unroll
      for (int j = 0; j < N; ++j) G[r * N + j] = fma_<T>(-f, G[p * N + j], G[r * N + j]);
    }
  }
#pragma unroll
  for (int i = 0; i < M * N; ++i) G[i] = -G[i];
}

// One backward Riccati stage (reference session_1/FHC.py:56-57), in the operation order
//   W = P A;  G = B'W;  PB = P B;  S = R + B'PB;  K = -S^-1 G;  W += PB K;  P = Q + A'W
// which is algebraically  K = -(R+B'PB)^-1 B'PA,  P = Q + A'PA + A'PB K  and costs
// 4n^3 + 6n^2 m + 4 n m^2 + O(m^3) flops (F_ric).
template <typename T, int NX, int NU, bool SYM = false>
MPC_HD void riccati_stage(const T* A, const T* B, const T* Q, const T* R, T* P,
                                              T* K) {
  T W[NX * NX];
  mm<T, NX, NX, NX, false>(P, A, W);
  mtm<T, NU, NX, NX, false>(B, W, K);  // K <- G = B' W   [NU x NX]
  T PB[NX * NU];
  mm<T, NX, NX, NU, false>(P, B, PB);
  T S[NU * NU];
#pragma unroll
  for (int i = 0; i < NU * NU; ++i) S[i] = R[i];
  mtm<T, NU, NX, NU, true>(B, PB, S);
  neg_solve<T, NU, NX>(S, K);
  mm<T, NX, NU, NX, true>(PB, K, W);
  if constexpr (SYM) {
    // P is symmetric in exact arithmetic: form the upper triangle of Q + A'W and mirror it.
    // (The reference does not symmetrise; the difference is rounding-level, ~1e-16 relative.)
#pragma unroll
    for (int i = 0; i < NX; ++i)
#pragma unroll
      for (int j = i; j < NX; ++j) {
        T acc = Q[i * NX + j];
#pragma unroll
        for (int k = 0; k < NX; ++k) acc = fma_<T>(A[k * NX + i], W[k * NX + j], acc);
        P[i * NX + j] = acc;
        P[j * NX + i] = acc;
      }
  } else {
#pragma unroll
    for (int i = 0; i < NX * NX; ++i) P[i] = Q[i];
    mtm<T, NX, NX, NX, true>(A, W, P);
  }
}

}  // namespace mpc
