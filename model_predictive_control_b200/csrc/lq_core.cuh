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

// ----------------------------------------------------------------- K1 + K2, single input (m = 1)
// Same solve in Krylov coordinates.  For a single-input model the change of variables x = C z with
// the controllability matrix C = [b, Ab, ..., A^{n-1}b] turns (A, b) into
//   A_o = C^-1 A C = [e_2, ..., e_n, -a]   (shifts and ONE dense column),   C^-1 b = e_1,
// where a = -C^-1 A^n b holds the characteristic-polynomial coefficients.  In these coordinates a
// backward Riccati stage (FHC.py:56-57 applied to A_o, e_1, Q_c = C'QC) is O(n^2):
//   P b = P[:,0],  S = R + P_00,  W = P A_o = [P[:,1..n-1], w],  w = -P a,  K = -W[0,:]/S,
//   M = W + P[:,0] K,  P+ = Q_c + A_o' M  (row i < n-1 of A_o'M is row i+1 of M, the last row is -a'M),
// and so is a forward stage: u = K z, z+ = [u, z_0, .., z_{n-2}] - a z_{n-1}, x = C z (the plan u_k is
// coordinate-free).  Per solve at n=4, N=20: ~2 100 FP64 instructions instead of ~3 900.
// The transformation loses about cond(C)^2 * eps: the body measures cond_F(C)^2 and returns false
// (nothing but X[0] written) when it exceeds `cond2_max`; the caller then runs lq_solve_body.
// Symmetric Q, Pf assumed (the upper triangles of C'QC and C'PfC are used), as in lq_solve_body.

// adjugate and determinant of a small matrix (row-major); inverse = adj / det
template <int NX>
MPC_HD double adjugate(const double* m, double* adj);

template <>
MPC_HD double adjugate<2>(const double* m, double* adj) {
  adj[0] = m[3];
  adj[1] = -m[1];
  adj[2] = -m[2];
  adj[3] = m[0];
  return fma(m[0], m[3], -m[1] * m[2]);
}

template <>
MPC_HD double adjugate<4>(const double* m, double* adj) {
#define MPC_D2(a, b, c, d) fma(m[a], m[b], -(m[c] * m[d]))
  const double s0 = MPC_D2(0, 5, 4, 1), s1 = MPC_D2(0, 6, 4, 2), s2 = MPC_D2(0, 7, 4, 3);
  const double s3 = MPC_D2(1, 6, 5, 2), s4 = MPC_D2(1, 7, 5, 3), s5 = MPC_D2(2, 7, 6, 3);
  const double c5 = MPC_D2(10, 15, 14, 11), c4 = MPC_D2(9, 15, 13, 11), c3 = MPC_D2(9, 14, 13, 10);
  const double c2 = MPC_D2(8, 15, 12, 11), c1 = MPC_D2(8, 14, 12, 10), c0 = MPC_D2(8, 13, 12, 9);
#undef MPC_D2
  adj[0] = fma(m[5], c5, fma(-m[6], c4, m[7] * c3));
  adj[1] = fma(-m[1], c5, fma(m[2], c4, -(m[3] * c3)));
  adj[2] = fma(m[13], s5, fma(-m[14], s4, m[15] * s3));
  adj[3] = fma(-m[9], s5, fma(m[10], s4, -(m[11] * s3)));
  adj[4] = fma(-m[4], c5, fma(m[6], c2, -(m[7] * c1)));
  adj[5] = fma(m[0], c5, fma(-m[2], c2, m[3] * c1));
  adj[6] = fma(-m[12], s5, fma(m[14], s2, -(m[15] * s1)));
  adj[7] = fma(m[8], s5, fma(-m[10], s2, m[11] * s1));
  adj[8] = fma(m[4], c4, fma(-m[5], c2, m[7] * c0));
  adj[9] = fma(-m[0], c4, fma(m[1], c2, -(m[3] * c0)));
  adj[10] = fma(m[12], s4, fma(-m[13], s2, m[15] * s0));
  adj[11] = fma(-m[8], s4, fma(m[9], s2, -(m[11] * s0)));
  adj[12] = fma(-m[4], c3, fma(m[5], c1, -(m[6] * c0)));
  adj[13] = fma(m[0], c3, fma(-m[1], c1, m[2] * c0));
  adj[14] = fma(-m[12], s3, fma(m[13], s1, -(m[14] * s0)));
  adj[15] = fma(m[8], s3, fma(-m[9], s1, m[10] * s0));
  return fma(s0, c5, fma(-s1, c4, fma(s2, c3, fma(s3, c2, fma(-s4, c1, s5 * c0)))));
}

// upper triangle (mirrored storage: only entries i <= j are defined) of C' S C for a full S
template <int NX>
MPC_HD void congruence_upper(const double* C, const double* S, double* out) {
  double T[NX * NX];
  mm<double, NX, NX, NX, false>(S, C, T);
#pragma unroll
  for (int i = 0; i < NX; ++i)
#pragma unroll
    for (int j = i; j < NX; ++j) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < NX; ++k) acc = fma(C[k * NX + i], T[k * NX + j], acc);
      out[i * NX + j] = acc;
    }
}

#define MPC_SYM(P, i, j) ((i) <= (j) ? (P)[(i) * NX + (j)] : (P)[(j) * NX + (i)])

// One backward stage in Krylov coordinates: Pn = Q + A_o'(P A_o + P e_1 K), K = -(e_1'P A_o)/(R + P_00); K -> Ks.
// Upper triangles only (entries i <= j of P, Pn, Q).
template <int NX>
MPC_HD void krylov_stage(const double* P, double* Pn, const double* Q, const double* na, double R, double* Ks,
                         int kstride) {
  using T = double;
  const T nrS = -rcp_(R + P[0]);
  T w[NX], K[NX], mc[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) {  // w = -P a
    T s = T(0);
#pragma unroll
    for (int l = 0; l < NX; ++l) s = fma(MPC_SYM(P, i, l), na[l], s);
    w[i] = s;
  }
#pragma unroll
  for (int j = 0; j + 1 < NX; ++j) K[j] = P[j + 1] * nrS;  // G = W[0,:] = [P_01 .. P_0,n-1, w_0]
  K[NX - 1] = w[0] * nrS;
#pragma unroll
  for (int e = 0; e < NX; ++e) Ks[e * kstride] = K[e];
  // rows 0..n-2 of P+ = Q_c + rows 1..n-1 of M,  M(r,c) = W(r,c) + P_0r K_c
#pragma unroll
  for (int i = 0; i + 1 < NX; ++i)
#pragma unroll
    for (int j = i; j + 1 < NX; ++j)
      Pn[i * NX + j] = fma(P[i + 1], K[j], Q[i * NX + j] + MPC_SYM(P, i + 1, j + 1));
#pragma unroll
  for (int r = 0; r < NX; ++r) mc[r] = fma(P[r], K[NX - 1], w[r]);  // last column of M
#pragma unroll
  for (int i = 0; i + 1 < NX; ++i) Pn[i * NX + NX - 1] = Q[i * NX + NX - 1] + mc[i + 1];
  T last = Q[NX * NX - 1];
#pragma unroll
  for (int r = 0; r < NX; ++r) last = fma(na[r], mc[r], last);
  Pn[NX * NX - 1] = last;
}

template <int NX, bool AL>
MPC_HD bool lq_solve_krylov_body(const LqSolveArgs<double>& a, int64_t b, double* Ks, int kstride,
                                 double cond2_max) {
  using T = double;
  constexpr int ANN = AL ? RowAlign<T, NX * NX>::value : 8;
  constexpr int AN = AL ? RowAlign<T, NX>::value : 8;
  T C[NX * NX], na[NX], z[NX], x[NX];
  load_row<T, NX, AN>(a.x0 + b * NX, x);
  store_row<T, NX, AN>(a.X + b * NX, x);
  {
    T A[NX * NX], col[NX], nxt[NX], adj[NX * NX];
    load_row<T, NX * NX, ANN>(a.A + b * a.sA, A);
    load_row<T, NX, AN>(a.B + b * a.sB, col);
#pragma unroll
    for (int j = 0; j < NX; ++j) {
#pragma unroll
      for (int i = 0; i < NX; ++i) C[i * NX + j] = col[i];
      mv<T, NX, NX, false>(A, col, nxt);  // after the last column: nxt = A^n b
#pragma unroll
      for (int i = 0; i < NX; ++i) col[i] = nxt[i];
    }
    const T det = adjugate<NX>(C, adj);
    T nC = T(0), nAdj = T(0);
#pragma unroll
    for (int i = 0; i < NX * NX; ++i) {
      nC = fma(C[i], C[i], nC);
      nAdj = fma(adj[i], adj[i], nAdj);
    }
    const T rdet = rcp_(det);
    // cond_F(C)^2 = |C|_F^2 |adj|_F^2 / det^2; NaN/inf (singular C, non-finite data) compare false
    if (!(nC * nAdj * rdet * rdet <= cond2_max)) return false;
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      T s = T(0), t = T(0);
#pragma unroll
      for (int k = 0; k < NX; ++k) {
        s = fma(adj[i * NX + k], col[k], s);
        t = fma(adj[i * NX + k], x[k], t);
      }
      na[i] = s * rdet;  // -a = C^-1 A^n b
      z[i] = t * rdet;   // z_0 = C^-1 x_0
    }
  }
  T P[NX * NX], Q[NX * NX], R;
  {
    T S[NX * NX];
    load_row<T, NX * NX, ANN>(a.Q + b * a.sQ, S);
    congruence_upper<NX>(C, S, Q);
    load_row<T, NX * NX, ANN>(a.Pf + b * a.sPf, S);
    congruence_upper<NX>(C, S, P);
    R = a.R[b * a.sR];
  }
  // two stages per trip, ping-ponging between P and P2, so that no stage ends with a register copy
  int k = a.N - 1;
  T P2[NX * NX];
  double* Kb = Ks + (int64_t)k * NX * kstride;  // slot of stage k, walked downwards
  const int kdec = NX * kstride;
  for (; k >= 1; k -= 2) {
    krylov_stage<NX>(P, P2, Q, na, R, Kb, kstride);
    krylov_stage<NX>(P2, P, Q, na, R, Kb - kdec, kstride);
    Kb -= 2 * kdec;
  }
  if (k == 0) {
    krylov_stage<NX>(P, P2, Q, na, R, Ks, kstride);
#pragma unroll
    for (int i = 0; i < NX; ++i)
#pragma unroll
      for (int j = i; j < NX; ++j) P[i * NX + j] = P2[i * NX + j];
  }
  {
    // V_N(x0) = x0' P_0 x0 = z0' P_c z0 (FHC.py:123-124)
    T v = T(0);
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      T r = T(0);
#pragma unroll
      for (int j = 0; j < NX; ++j) r = fma(MPC_SYM(P, i, j), z[j], r);
      v = fma(z[i], r, v);
    }
    a.V[b] = v;
  }
  // forward sweep with running pointers; the next stage's gain leaves shared memory while this stage is computed
  // (the last trip re-reads its own slot instead of branching)
  T Kc[NX];
  const double* Kp = Ks;
  const int kinc = NX * kstride;
  double* Up = a.U + b;
  double* Xp = a.X + (a.batch + b) * NX;
  const int64_t xinc = a.batch * NX;
#pragma unroll
  for (int e = 0; e < NX; ++e) Kc[e] = Kp[e * kstride];
  for (int k = 0; k < a.N; ++k) {
    T u = T(0);
#pragma unroll
    for (int e = 0; e < NX; ++e) u = fma(Kc[e], z[e], u);
    Kp += (k + 1 < a.N) ? kinc : 0;
#pragma unroll
    for (int e = 0; e < NX; ++e) Kc[e] = Kp[e * kstride];
    *Up = u;
    Up += a.batch;
    const T zl = z[NX - 1];
#pragma unroll
    for (int i = NX - 1; i >= 1; --i) z[i] = fma(na[i], zl, z[i - 1]);
    z[0] = fma(na[0], zl, u);
    mv<T, NX, NX, false>(C, z, x);
    store_row<T, NX, AN>(Xp, x);
    Xp += xinc;
  }
  return true;
}
#undef MPC_SYM

}  // namespace mpc
