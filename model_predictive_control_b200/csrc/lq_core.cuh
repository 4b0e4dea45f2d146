// Per-scenario bodies of the LQ path (K1 Riccati recursion, K2 rollout, fused K1+K2 solve).
//
// Every body is __host__ __device__: the kernels in lq.cu call them with one thread per scenario,
// and tests/harness/host_harness.cu calls the very same templates in a plain CPU loop so that the
// algebra and indexing are checked against the oracle on machines without a GPU.  The harness is
// test infrastructure; the product library has no CPU execution path.
#pragma once

#include "smallmat.cuh"

namespace mpc {

// ------------------------------------------------------------------------------------------ K1
// Replaces FHC.ricatti_recursion (reference session_1/FHC.py:51-61) and
// session1_sol.riccati_recursion (session_1/session1_sol.py:44-65).
template <typename T>
struct RiccatiArgs {
  const T *A, *B, *Q, *R, *Pf;
  int64_t sA, sB, sQ, sR, sPf;  // elements between scenarios, 0 = shared
  T* K;                         // [N][batch][m][n]
  T* P;                         // all_P ? [N+1][batch][n][n] : [batch][n][n]; may be null
  int all_P;
  int64_t batch;
  int N;
};

template <typename T, int NX, int NU, bool AL>
MPC_HD void riccati_body(const RiccatiArgs<T>& a, int64_t b) {
  constexpr int ES = (int)sizeof(T);
  constexpr int ANN = AL ? RowAlign<T, NX * NX>::value : ES;
  constexpr int ANM = AL ? RowAlign<T, NX * NU>::value : ES;
  constexpr int AMM = AL ? RowAlign<T, NU * NU>::value : ES;
  T A[NX * NX], B[NX * NU], Q[NX * NX], R[NU * NU], P[NX * NX], K[NU * NX];
  load_row<T, NX * NX, ANN>(a.A + b * a.sA, A);
  load_row<T, NX * NU, ANM>(a.B + b * a.sB, B);
  load_row<T, NX * NX, ANN>(a.Q + b * a.sQ, Q);
  load_row<T, NU * NU, AMM>(a.R + b * a.sR, R);
  load_row<T, NX * NX, ANN>(a.Pf + b * a.sPf, P);
  const int64_t pstage = a.batch * (NX * NX);
  const int64_t kstage = a.batch * (NU * NX);
  if (a.P && a.all_P) store_row<T, NX * NX, ANN>(a.P + a.N * pstage + b * (NX * NX), P);
  for (int k = a.N - 1; k >= 0; --k) {
    riccati_stage<T, NX, NU>(A, B, Q, R, P, K);
    store_row<T, NU * NX, ANM>(a.K + k * kstage + b * (NU * NX), K);
    if (a.P && a.all_P) store_row<T, NX * NX, ANN>(a.P + k * pstage + b * (NX * NX), P);
  }
  if (a.P && !a.all_P) store_row<T, NX * NX, ANN>(a.P + b * (NX * NX), P);
}

// ------------------------------------------------------------------------------------------ K2
// Replaces LinearSystem.simulate / prediction (reference session_1/LinearSystem.py:20-35) under
// the AutoCruising policies (session_1/FHC.py:25-29) and session1_sol.simulate (:68-91).
template <typename T>
struct RolloutArgs {
  const T *A, *B;  // [*][n][n], [*][n][m]
  int64_t sA, sB;
  const T* K;  // [ng][*][m][n]
  int64_t sK_stage, sK;
  int ng, gain_offset, gain_step;
  const T* x0;          // [n][batch]
  T* X;                 // [T][n][batch]
  T* U;                 // [T-1][m][batch] or null
  const T *Q, *R, *Pf;  // shared; needed when cost != null
  T* cost;              // [batch] or null
  uint8_t* unstable;    // [batch] or null
  T norm_limit2;        // squared limit
  int64_t batch;
  int Tn;               // number of states (incl. x0)
};

// Shared model and gains: `sm` points at [A | B | Q | R | Pf | K(all ng stages)] (shared memory on
// the device).  VEC consecutive scenarios per thread so that the batch-contiguous rows are moved
// with 16/32-byte accesses.
template <typename T, int NX, int NU>
struct RolloutSmem {
  static constexpr int oA = 0;
  static constexpr int oB = oA + NX * NX;
  static constexpr int oQ = oB + NX * NU;
  static constexpr int oR = oQ + NX * NX;
  static constexpr int oPf = oR + NU * NU;
  static constexpr int oK = oPf + NX * NX;
  static int total(int ng) { return oK + ng * NU * NX; }
};

template <typename T, int VEC>
MPC_HD void load_vec(const T* p, T* r) {
  if constexpr (VEC == 1) {
    r[0] = p[0];
  } else {
    load_row<T, VEC, VEC * (int)sizeof(T)>(p, r);
  }
}
template <typename T, int VEC>
MPC_HD void store_vec(T* p, const T* r) {
  if constexpr (VEC == 1) {
    p[0] = r[0];
  } else {
    store_row<T, VEC, VEC * (int)sizeof(T)>(p, r);
  }
}

template <typename T, int NX, int NU, int VEC>
MPC_HD void rollout_shared_body(const RolloutArgs<T>& a, const T* sm, int64_t b0) {
  using L = RolloutSmem<T, NX, NU>;
  T x[NX][VEC], u[NU][VEC], xn[NX][VEC], c[VEC];
  bool bad[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    c[v] = T(0);
    bad[v] = false;
  }
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    load_vec<T, VEC>(a.x0 + i * a.batch + b0, x[i]);
    store_vec<T, VEC>(a.X + i * a.batch + b0, x[i]);
  }
  const bool want_cost = a.cost != nullptr;
  for (int t = 0; t + 1 < a.Tn; ++t) {
    const T* Kt = sm + L::oK + (a.gain_offset + a.gain_step * t) * (NU * NX);
#pragma unroll
    for (int j = 0; j < NU; ++j) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) u[j][v] = T(0);
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        const T kji = Kt[j * NX + i];
#pragma unroll
        for (int v = 0; v < VEC; ++v) u[j][v] = fma_<T>(kji, x[i][v], u[j][v]);
      }
      if (a.U) store_vec<T, VEC>(a.U + ((int64_t)t * NU + j) * a.batch + b0, u[j]);
    }
    if (want_cost) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        T xv[NX], uv[NU];
#pragma unroll
        for (int i = 0; i < NX; ++i) xv[i] = x[i][v];
#pragma unroll
        for (int j = 0; j < NU; ++j) uv[j] = u[j][v];
        c[v] += quad<T, NX>(sm + L::oQ, xv) + quad<T, NU>(sm + L::oR, uv);
      }
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) xn[i][v] = T(0);
#pragma unroll
      for (int k = 0; k < NX; ++k) {
        const T aik = sm[L::oA + i * NX + k];
#pragma unroll
        for (int v = 0; v < VEC; ++v) xn[i][v] = fma_<T>(aik, x[k][v], xn[i][v]);
      }
#pragma unroll
      for (int j = 0; j < NU; ++j) {
        const T bij = sm[L::oB + i * NU + j];
#pragma unroll
        for (int v = 0; v < VEC; ++v) xn[i][v] = fma_<T>(bij, u[j][v], xn[i][v]);
      }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      T n2 = T(0);
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        x[i][v] = xn[i][v];
        n2 = fma_<T>(xn[i][v], xn[i][v], n2);
      }
      bad[v] = bad[v] || (n2 > a.norm_limit2);
    }
#pragma unroll
    for (int i = 0; i < NX; ++i)
      store_vec<T, VEC>(a.X + ((int64_t)(t + 1) * NX + i) * a.batch + b0, x[i]);
  }
  if (want_cost) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      T xv[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) xv[i] = x[i][v];
      c[v] += quad<T, NX>(sm + L::oPf, xv);
    }
    store_vec<T, VEC>(a.cost + b0, c);
  }
  if (a.unstable) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) a.unstable[b0 + v] = bad[v] ? 1 : 0;
  }
}

// Per-scenario model and/or gains (any of sA, sB, sK may still be 0).  One scenario per thread.
template <typename T, int NX, int NU>
MPC_HD void rollout_perscn_body(const RolloutArgs<T>& a, int64_t b) {
  T A[NX * NX], B[NX * NU], x[NX], u[NU], xn[NX];
#pragma unroll
  for (int i = 0; i < NX * NX; ++i) A[i] = a.A[b * a.sA + i];
#pragma unroll
  for (int i = 0; i < NX * NU; ++i) B[i] = a.B[b * a.sB + i];
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    x[i] = a.x0[i * a.batch + b];
    a.X[i * a.batch + b] = x[i];
  }
  T c = T(0);
  bool bad = false;
  for (int t = 0; t + 1 < a.Tn; ++t) {
    const T* Kt = a.K + (int64_t)(a.gain_offset + a.gain_step * t) * a.sK_stage + b * a.sK;
    T Kr[NU * NX];
#pragma unroll
    for (int i = 0; i < NU * NX; ++i) Kr[i] = Kt[i];
    mv<T, NU, NX, false>(Kr, x, u);
    if (a.U) {
#pragma unroll
      for (int j = 0; j < NU; ++j) a.U[((int64_t)t * NU + j) * a.batch + b] = u[j];
    }
    if (a.cost) c += quad<T, NX>(a.Q, x) + quad<T, NU>(a.R, u);
    mv<T, NX, NX, false>(A, x, xn);
    mv<T, NX, NU, true>(B, u, xn);
    T n2 = T(0);
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      x[i] = xn[i];
      n2 = fma_<T>(x[i], x[i], n2);
      a.X[((int64_t)(t + 1) * NX + i) * a.batch + b] = x[i];
    }
    bad = bad || (n2 > a.norm_limit2);
  }
  if (a.cost) a.cost[b] = c + quad<T, NX>(a.Pf, x);
  if (a.unstable) a.unstable[b] = bad ? 1 : 0;
}

// ------------------------------------------------------------------------------------- K1 + K2
// Fused per-scenario finite-horizon LQ solve: backward recursion (FHC.py:51-61), then the optimal
// plan u_k = K_k x_k, x_{k+1} = A x_k + B u_k (LinearSystem.py:16-18) and its cost
// (= x0' P_0 x0, FHC.py:123-124).  The N gains stay on chip between the two sweeps: `Ks` is this
// scenario's slice of shared memory, element e of stage k at Ks[(k * NU*NX + e) * kstride].
template <typename T>
struct LqSolveArgs {
  const T *A, *B, *Q, *R, *Pf;
  int64_t sA, sB, sQ, sR, sPf;
  const T* x0;  // [batch][n]
  T* X;         // [N+1][batch][n]
  T* U;         // [N][batch][m]
  T* V;         // [batch]
  T* K;         // optional [N][batch][m][n]
  T* P0;        // optional [batch][n][n]
  int64_t batch;
  int N;
};

template <typename T, int NX, int NU, bool AL>
MPC_HD void lq_solve_body(const LqSolveArgs<T>& a, int64_t b, T* Ks, int kstride) {
  constexpr int ES = (int)sizeof(T);
  constexpr int ANN = AL ? RowAlign<T, NX * NX>::value : ES;
  constexpr int ANM = AL ? RowAlign<T, NX * NU>::value : ES;
  constexpr int AMM = AL ? RowAlign<T, NU * NU>::value : ES;
  constexpr int AN = AL ? RowAlign<T, NX>::value : ES;
  constexpr int AM = AL ? RowAlign<T, NU>::value : ES;
  T A[NX * NX], B[NX * NU];
  load_row<T, NX * NX, ANN>(a.A + b * a.sA, A);
  load_row<T, NX * NU, ANM>(a.B + b * a.sB, B);
  T x[NX], xn[NX], u[NU];
  load_row<T, NX, AN>(a.x0 + b * NX, x);
  store_row<T, NX, AN>(a.X + b * NX, x);
  {
    T Q[NX * NX], R[NU * NU], P[NX * NX], K[NU * NX];
    load_row<T, NX * NX, ANN>(a.Q + b * a.sQ, Q);
    load_row<T, NU * NU, AMM>(a.R + b * a.sR, R);
    load_row<T, NX * NX, ANN>(a.Pf + b * a.sPf, P);
    for (int k = a.N - 1; k >= 0; --k) {
      riccati_stage<T, NX, NU, true>(A, B, Q, R, P, K);
#pragma unroll
      for (int e = 0; e < NU * NX; ++e) Ks[(k * (NU * NX) + e) * kstride] = K[e];
      if (a.K) store_row<T, NU * NX, ANM>(a.K + ((int64_t)k * a.batch + b) * (NU * NX), K);
    }
    if (a.P0) store_row<T, NX * NX, ANN>(a.P0 + b * (NX * NX), P);
    // optimal cost exactly as the reference evaluates it: V_N(x0) = x0' P_0 x0 (FHC.py:123-124)
    a.V[b] = quad<T, NX>(P, x);
  }
  for (int k = 0; k < a.N; ++k) {
    T K[NU * NX];
#pragma unroll
    for (int e = 0; e < NU * NX; ++e) K[e] = Ks[(k * (NU * NX) + e) * kstride];
    mv<T, NU, NX, false>(K, x, u);
    store_row<T, NU, AM>(a.U + ((int64_t)k * a.batch + b) * NU, u);
    mv<T, NX, NX, false>(A, x, xn);
    mv<T, NX, NU, true>(B, u, xn);
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = xn[i];
    store_row<T, NX, AN>(a.X + ((int64_t)(k + 1) * a.batch + b) * NX, x);
  }
}

}  // namespace mpc
