// K4: per-scenario body of the box-constrained LQ-MPC QP solver.
//
//   min  sum_{k<N} x_k'Q x_k + u_k'R u_k + x_N'Pf x_N
//   s.t. x_{k+1} = A_k x_k + B_k u_k + c_k,  u_lo <= u_k <= u_hi,  x_lo <= x_{k+1} <= x_hi
//
// (the QP posed by the reference's session_2/problem.py:8-24 and session_3/problem.py:12-28, and --
// with per-scenario stage matrices A_k, B_k, c_k -- the linearised QP of session 4's closed loop,
// session_4/session4_sol.py:158-217).  The reference ships no solver for it.
//
// Method: Mehrotra predictor-corrector interior point.  With z_k = (u_k, x_{k+1}), slacks s and
// multipliers lam for every finite bound, each Newton system (residual form, unknown dz)
//     dz = argmin 1/2 dz'(H + Sigma) dz - rhs'dz   s.t. dx_{k+1} = A_k dx_k + B_k du_k, dx_0 = 0
//     Sigma = lam_l/s_l + lam_u/s_u,   rhs = -H z + (sig mu - cc_l)/s_l - Sigma_l r_l - (sig mu - cc_u)/s_u + Sigma_u r_u
// is an unconstrained LQ problem with stage-varying diagonal weight updates, solved by one
// backward Riccati sweep (factorisation + feed-forward) and one forward rollout; the corrector
// reuses the stored gains.  The iterate z always satisfies the dynamics exactly (it starts from a
// rollout and moves along dynamics-consistent directions).  In residual form every term of rhs
// stays O(lam), so the rounding error of the huge barrier weights (Sigma ~ 1e12) scales with |dz|
// and vanishes at the solution.  A first-order (projected-gradient / ADMM) iteration was evaluated in
// oracle/boxqp.py (admm_riccati): 600-2000+ iterations for 1e-6 parity on the session-2 data versus
// ~15 here, so the interior-point iteration is the one that ships.
//
// One thread per scenario.  All per-scenario state lives in a caller-provided workspace laid out
// [stage][element][batch] (batch-contiguous), so every access of a warp is one coalesced row.
// The body is __host__ __device__: tests/harness runs it on the CPU against oracle/boxqp.py.
#pragma once

#include "smallmat.cuh"

namespace mpc {

constexpr double kBigBound = 1e19;  // |bound| >= kBigBound means "no bound"

template <typename T>
struct BoxQpArgs {
  const T *A, *B, *c;  // ltv == 0: shared A [n][n], B [n][m], c null
                       // ltv == 1: A [N][n*n][batch], B [N][n*m][batch], c [N][n][batch]
  int ltv;
  const T *Q, *R, *Pf;                  // shared
  const T *u_lo, *u_hi, *x_lo, *x_hi;   // shared [m], [m], [n], [n]
  const T* x0;                          // [n][batch]
  const T* warm_U;                      // optional [N][m][batch]
  T* U;                                 // [N][m][batch]
  T* X;                                 // [N+1][n][batch]
  T* cost;                              // [batch]
  int32_t* status;                      // [batch]
  int32_t* iters;                       // [batch]
  int8_t* sat_u;                        // optional [N][m][batch]: -1 lower, +1 upper, 0 free
  int8_t* sat_x;                        // optional [N][n][batch]
  // optional general stage rows  Cg_k x_{k+1} >= hg_k  (kernels instantiated with NC > 0 only):
  const T* Cg;                          // [N][NC*n][batch]
  const T* hg;                          // [N][NC][batch]
  int8_t* sat_c;                        // optional [N][NC][batch]: -1 = row active
  T* ws;                                // workspace, boxqp_ws_elems(n, m, N, nc) * batch elements
  int64_t batch;
  int N;
  int max_iter;
  T eps;
  int pf_dist = 0;  // stages of L2 prefetch ahead of each sweep's loads (device only; 0 = off)
};

// workspace elements per scenario
inline int64_t boxqp_ws_elems(int n, int m, int N, int nc = 0) {
  const int d = n + m;
  return (int64_t)N * (7 * d + m * n + m * m + m + 3 * nc);
}

// shared-parameter block (shared memory on the device)
template <int NX, int NU>
struct BoxQpShared {
  static constexpr int D = NX + NU;
  static constexpr int oA = 0;
  static constexpr int oB = oA + NX * NX;
  static constexpr int oQ = oB + NX * NU;
  static constexpr int oR = oQ + NX * NX;
  static constexpr int oPf = oR + NU * NU;
  static constexpr int oLo = oPf + NX * NX;  // [u_lo | x_lo]
  static constexpr int oHi = oLo + D;        // [u_hi | x_hi]
  static constexpr int total = oHi + D;
};

// MODEL = 0: generic dense model (shared LTI, or per-scenario LTV A [N][n*n][batch], B, c).
// MODEL = 1: forward-Euler kinematic bicycle (NX = 4, NU = 2), per-scenario LTV in PACKED form: only the 10 entries of
//            A = I + ts J_x and B = ts J_u that are not structurally 0 or 1, plus c: a.A -> [N][14][batch]
//            {a02, a03, a12, a13, a23, a33, b01, b11, b21, b30, c0..c3}.  The structural zeros and ones are written
//            as literals, so the unrolled register algebra drops the corresponding multiplications at compile time.
constexpr int kBicyclePack = 14;

template <typename T, int NX, int NU, int NC = 0, int MODEL = 0>
struct BoxQpIpm {
  static constexpr int D = NX + NU;
  // Loads of a stage visit are hidden by resident warps plus an L2 prefetch of the rows a few visits ahead (pf_stage).
  // Double-buffering the next stage in REGISTERS (an earlier version, for n + m <= 3) cost more than it hid: 168
  // registers with 245 M local-memory sectors of spill traffic per 65 536 solves (ncu); without it the (2,1) kernel
  // has no spills at 152 registers and runs 1.2x faster at 4 CTAs/SM.
  using SH = BoxQpShared<NX, NU>;

  const BoxQpArgs<T>& a;
  const T* sh;
  int64_t b, bs;
  T mu_scale;  // max(1, max|Q|, max|R|): scale of the complementarity tolerance
  T mu0;       // start value of the barrier parameter, per scenario: max(mu_scale, |H z0|_inf)
  // workspace sections, each [N][per][batch]
  T *z, *sl, *su, *ll, *lu, *dza, *dzw, *Kw, *Sw, *dw, *sc, *lc, *rc;  // general rows: slack, multiplier, residual C x - h - s

  MPC_HD BoxQpIpm(const BoxQpArgs<T>& args, const T* shared, int64_t scenario)
      : a(args), sh(shared), b(scenario), bs(args.batch) {
    const int64_t sec = (int64_t)a.N * D * bs;
    z = a.ws;
    sl = z + sec;
    su = sl + sec;
    ll = su + sec;
    lu = ll + sec;
    dza = lu + sec;   // affine (predictor) direction dz_aff
    dzw = dza + sec;  // corrector direction dz
    Kw = dzw + sec;
    Sw = Kw + (int64_t)a.N * NU * NX * bs;
    dw = Sw + (int64_t)a.N * NU * NU * bs;
    sc = dw + (int64_t)a.N * NU * bs;
    lc = sc + (int64_t)a.N * NC * bs;
    rc = lc + (int64_t)a.N * NC * bs;
    mu_scale = T(1);
    for (int i = 0; i < NX * NX; ++i) {
      const T v = sh[SH::oQ + i] < T(0) ? -sh[SH::oQ + i] : sh[SH::oQ + i];
      mu_scale = v > mu_scale ? v : mu_scale;
    }
    for (int i = 0; i < NU * NU; ++i) {
      const T v = sh[SH::oR + i] < T(0) ? -sh[SH::oR + i] : sh[SH::oR + i];
      mu_scale = v > mu_scale ? v : mu_scale;
    }
    mu0 = mu_scale;
  }

  MPC_HD int64_t ix(int k, int i, int per) const { return ((int64_t)k * per + i) * bs + b; }
  MPC_HD bool hasl(int i) const { return sh[SH::oLo + i] > T(-kBigBound); }
  MPC_HD bool hasu(int i) const { return sh[SH::oHi + i] < T(kBigBound); }
  MPC_HD T lo(int i) const { return sh[SH::oLo + i]; }
  MPC_HD T hi(int i) const { return sh[SH::oHi + i]; }

  // ---- one stage's iterate.  All loads are unconditional and issued together (entries without a
  // bound hold s = 1, lam = 0), so a stage visit costs one memory round trip instead of a chain.
  struct Stage {
    T z[D], sl[D], su[D], ll[D], lu[D];
  };
  MPC_HD void load(int k, Stage& s) const {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const int64_t o = ix(k, i, D);
      s.z[i] = z[o];
      s.sl[i] = sl[o];
      s.su[i] = su[o];
      s.ll[i] = ll[o];
      s.lu[i] = lu[o];
    }
  }
  // ---- L2 prefetch of the rows a later stage visit will load.  The workspace is streamed from HBM once per pass
  // (it is far larger than L2); a `prefetch.global.L2` per row, issued pf_dist stage visits early, turns the
  // ~1 us DRAM round trip of those loads into an L2 hit without holding registers for the data in flight.  It pays
  // while a kernel is latency-bound (the (2,1) kernel: pf_dist = 2) and costs once it is bandwidth-bound, because lines
  // evicted before use are fetched twice (the (4,2) kernels and the fused RTI loop: pf_dist = 0); the launchers decide.
  MPC_HD static void pf(const T* p) {
#ifdef __CUDA_ARCH__
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
  }
  template <int PER>
  MPC_HD void pf_rows(const T* base, int k) const {
#pragma unroll
    for (int i = 0; i < PER; ++i) pf(base + ix(k, i, PER));
  }
  // rows every pass reads: the iterate of stage k (+ the per-scenario model)
  MPC_HD void pf_stage(int k) const {
#ifdef __CUDA_ARCH__
    if constexpr (NC > 0) return;     // general rows: measured 6.1e5 QPs/s with, 6.2-6.9e5 without (obstacle workload)
    if (a.pf_dist <= 0 || k < 0 || k >= a.N) return;
    pf_rows<D>(z, k);
    pf_rows<D>(sl, k);
    pf_rows<D>(su, k);
    pf_rows<D>(ll, k);
    pf_rows<D>(lu, k);
    if constexpr (MODEL == 1) {
      pf_rows<kBicyclePack>(a.A, k);
    } else if (a.ltv) {
      pf_rows<NX * NX>(a.A, k);
      pf_rows<NX * NU>(a.B, k);
      pf_rows<NX>(a.c, k);
    }
#else
    (void)k;
#endif
  }
  MPC_HD void pf_extra(int k, bool gains, bool sinv, bool ff, bool aff, bool dir) const {
#ifdef __CUDA_ARCH__
    if constexpr (NC > 0) return;
    if (a.pf_dist <= 0 || k < 0 || k >= a.N) return;
    if (gains) pf_rows<NU * NX>(Kw, k);
    if (sinv) pf_rows<NU * NU>(Sw, k);
    if (ff) pf_rows<NU>(dw, k);
    if (aff) pf_rows<D>(dza, k);
    if (dir) pf_rows<D>(dzw, k);
#else
    (void)k; (void)gains; (void)sinv; (void)ff; (void)aff; (void)dir;
#endif
  }

  MPC_HD void load_vec(const T* base, int k, int per, T* v) const {
    for (int i = 0; i < per; ++i) v[i] = base[ix(k, i, per)];
  }
  template <int PER>
  MPC_HD void loadn(const T* base, int k, T* v) const {
#pragma unroll
    for (int i = 0; i < PER; ++i) v[i] = base[ix(k, i, PER)];
  }
  template <int PER>
  MPC_HD void storen(T* base, int k, const T* v) const {
#pragma unroll
    for (int i = 0; i < PER; ++i) base[ix(k, i, PER)] = v[i];
  }

  // general row j of stage k: coefficients C[NX] and right-hand side h
  MPC_HD T load_row_c(int k, int j, T* C) const {
#pragma unroll
    for (int i = 0; i < NX; ++i) C[i] = a.Cg[ix(k, j * NX + i, NC * NX)];
    return a.hg[ix(k, j, NC)];
  }
  MPC_HD static T dotx(const T* C, const T* x) {
    T acc = T(0);
#pragma unroll
    for (int i = 0; i < NX; ++i) acc = fma_<T>(C[i], x[i], acc);
    return acc;
  }

  MPC_HD void load_model(int k, T* A, T* B, T* c) const {
    if constexpr (MODEL == 1) {
      static_assert(MODEL == 0 || (NX == 4 && NU == 2), "packed bicycle model is 4 x 2");
      T v[kBicyclePack];
      loadn<kBicyclePack>(a.A, k, v);
#pragma unroll
      for (int i = 0; i < NX * NX; ++i) A[i] = T(0);
#pragma unroll
      for (int i = 0; i < NX * NU; ++i) B[i] = T(0);
      A[0] = T(1);
      A[2] = v[0];
      A[3] = v[1];
      A[5] = T(1);
      A[6] = v[2];
      A[7] = v[3];
      A[10] = T(1);
      A[11] = v[4];
      A[15] = v[5];
      B[1] = v[6];
      B[3] = v[7];
      B[5] = v[8];
      B[6] = v[9];
#pragma unroll
      for (int i = 0; i < NX; ++i) c[i] = v[10 + i];
      return;
    }
    if (a.ltv) {
      loadn<NX * NX>(a.A, k, A);
      loadn<NX * NU>(a.B, k, B);
      loadn<NX>(a.c, k, c);
    } else {
#pragma unroll
      for (int i = 0; i < NX * NX; ++i) A[i] = sh[SH::oA + i];
#pragma unroll
      for (int i = 0; i < NX * NU; ++i) B[i] = sh[SH::oB + i];
#pragma unroll
      for (int i = 0; i < NX; ++i) c[i] = T(0);
    }
  }

  // x+ = A x + B u + c
  MPC_HD static void step(const T* A, const T* B, const T* c, const T* x, const T* u, T* xn) {
#pragma unroll
    for (int i = 0; i < NX; ++i) xn[i] = c[i];
    mv<T, NX, NX, true>(A, x, xn);
    mv<T, NX, NU, true>(B, u, xn);
  }

  // ---- start point: inputs clamped into their box, states by rollout, slacks >= 1, lam = mu0/s
  // with mu0 = max(mu_scale, |H z0|_inf): multipliers start at the size of the cost gradient.
  // Pass 0 only measures |H z0|_inf, pass 1 writes the start point.
  MPC_HD void init() {
    T g0 = T(0);
    for (int pass = 0; pass < 2; ++pass) {
      T x[NX], xn[NX], u[NU], A[NX * NX], B[NX * NU], c[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = a.x0[i * bs + b];
      for (int k = 0; k < a.N; ++k) {
        load_model(k, A, B, c);
#pragma unroll
        for (int j = 0; j < NU; ++j) {
          T v = a.warm_U ? a.warm_U[ix(k, j, NU)] : T(0);
          v = v < lo(j) ? lo(j) : v;
          v = v > hi(j) ? hi(j) : v;
          u[j] = v;
        }
        step(A, B, c, x, u, xn);
        if (pass == 0) {
          const T* Qx = sh + (k == a.N - 1 ? SH::oPf : SH::oQ);
#pragma unroll
          for (int i = 0; i < NU; ++i) {
            T acc = T(0);
#pragma unroll
            for (int j = 0; j < NU; ++j) acc = fma_<T>(sh[SH::oR + i * NU + j], u[j], acc);
            acc = acc < T(0) ? -acc : acc;
            g0 = acc > g0 ? acc : g0;
          }
#pragma unroll
          for (int i = 0; i < NX; ++i) {
            T acc = T(0);
#pragma unroll
            for (int j = 0; j < NX; ++j) acc = fma_<T>(Qx[i * NX + j], xn[j], acc);
            acc = acc < T(0) ? -acc : acc;
            g0 = acc > g0 ? acc : g0;
          }
        } else {
          Stage st;
#pragma unroll
          for (int i = 0; i < D; ++i) {
            const T zi = i < NU ? u[i] : xn[i - NU];
            T s_l = T(1), s_u = T(1), l_l = T(0), l_u = T(0);
            if (hasl(i)) {
              s_l = zi - lo(i);
              s_l = s_l > T(1) ? s_l : T(1);
              l_l = mu0 / s_l;
            }
            if (hasu(i)) {
              s_u = hi(i) - zi;
              s_u = s_u > T(1) ? s_u : T(1);
              l_u = mu0 / s_u;
            }
            st.z[i] = zi;
            st.sl[i] = s_l;
            st.su[i] = s_u;
            st.ll[i] = l_l;
            st.lu[i] = l_u;
          }
          store_stage(k, st);
          if constexpr (NC > 0) {
#pragma unroll 1
            for (int j = 0; j < NC; ++j) {
              T C[NX];
              const T h = load_row_c(k, j, C);
              const T w = dotx(C, xn) - h;
              const T s = w > T(1) ? w : T(1);
              sc[ix(k, j, NC)] = s;
              lc[ix(k, j, NC)] = mu0 / s;
              // the row residual is carried, not recomputed: it decays exactly by (1 - alpha) per step, whereas
              // C x - h - s recomputed from a dot product keeps ~1e-16 of rounding noise that Sigma ~ 1e12 amplifies
              rc[ix(k, j, NC)] = w - s;
            }
          }
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) x[i] = xn[i];
      }
      if (pass == 0) mu0 = g0 > mu_scale ? g0 : mu_scale;
    }
  }

  MPC_HD void store_stage(int k, const Stage& s) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const int64_t o = ix(k, i, D);
      z[o] = s.z[i];
      sl[o] = s.sl[i];
      su[o] = s.su[i];
      ll[o] = s.ll[i];
      lu[o] = s.lu[i];
    }
  }

  // second-order term cc = ds_aff * dlam_aff of one bound, recomputed from the stored dz_aff
  // (affine direction: ds = +-dz + r, dlam = -lam - Sigma ds with Sigma = lam/s).
  // Divisions: ONE reciprocal per bound and stage visit, rinv = 1/(s lam); 1/s = rinv lam, 1/lam = rinv s.
  MPC_HD static T cc_of(T dz_signed, T r, T sig, T l) {
    const T ds = dz_signed + r;
    return ds * (-l - sig * ds);
  }

  // ---- backward sweep: (factorisation and) feed-forward terms of the Newton step.
  // rhs_i = -(H z)_i + (sig_mu - cc_l)/s_l - Sigma_l r_l - (sig_mu - cc_u)/s_u + Sigma_u r_u
  MPC_HD void backward(const bool FACTOR, T sig_mu) {
    T Pacc[NX * NX], pacc[NX];
#pragma unroll
    for (int i = 0; i < NX * NX; ++i) Pacc[i] = sh[SH::oPf + i];
#pragma unroll
    for (int i = 0; i < NX; ++i) pacc[i] = T(0);
    Stage cur;
    T da[D];
#pragma unroll
    for (int i = 0; i < D; ++i) da[i] = T(0);
    for (int k = a.N - 1; k >= 0; --k) {
      pf_stage(k - a.pf_dist);
      pf_extra(k - a.pf_dist, !FACTOR, !FACTOR, false, !FACTOR, false);
      load(k, cur);
      if (!FACTOR) loadn<D>(dza, k, da);
      T A[NX * NX], B[NX * NU], c[NX], K[NU * NX], Sinv[NU * NU];
      load_model(k, A, B, c);
      if (!FACTOR) {
        loadn<NU * NX>(Kw, k, K);
        loadn<NU * NU>(Sw, k, Sinv);
      }
      T sig[D], rhs[D];
      // -(H z): inputs weighted by R, state x_{k+1} by Q (Pf for the last stage)
      {
        const T* Qx = sh + (k == a.N - 1 ? SH::oPf : SH::oQ);
#pragma unroll
        for (int i = 0; i < NU; ++i) {
          T acc = T(0);
#pragma unroll
          for (int j = 0; j < NU; ++j) acc = fma_<T>(-sh[SH::oR + i * NU + j], cur.z[j], acc);
          rhs[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          T acc = T(0);
#pragma unroll
          for (int j = 0; j < NX; ++j) acc = fma_<T>(-Qx[i * NX + j], cur.z[NU + j], acc);
          rhs[NU + i] = acc;
        }
      }
#pragma unroll
      for (int i = 0; i < D; ++i) {
        T sg = T(0), r = rhs[i];
        if (hasl(i)) {
          const T s = cur.sl[i], l = cur.ll[i];
          const T inv = rcp_(s);
          const T sgl = l * inv;
          const T rl = cur.z[i] - lo(i) - s;
          const T cc = FACTOR ? T(0) : cc_of(da[i], rl, sgl, l);
          sg += sgl;
          r += (sig_mu - cc) * inv - sgl * rl;
        }
        if (hasu(i)) {
          const T s = cur.su[i], l = cur.lu[i];
          const T inv = rcp_(s);
          const T sgu = l * inv;
          const T ru = hi(i) - cur.z[i] - s;
          const T cc = FACTOR ? T(0) : cc_of(-da[i], ru, sgu, l);
          sg += sgu;
          r -= (sig_mu - cc) * inv - sgu * ru;
        }
        sig[i] = sg;
        rhs[i] = r;
      }
      if constexpr (NC > 0) {
        // general rows: value w = C x_{k+1}; adds C' Sigma_c C to the stage Hessian and C' rhs_c to the gradient
#pragma unroll 1
        for (int j = 0; j < NC; ++j) {
          T C[NX];
          const T h = load_row_c(k, j, C);
          const T s = sc[ix(k, j, NC)], l = lc[ix(k, j, NC)];
          const T inv = rcp_(s), sgc = l * inv;
          const T r = rc[ix(k, j, NC)];
          (void)h;
          const T cc = FACTOR ? T(0) : cc_of(dotx(C, da + NU), r, sgc, l);
          const T rhs_c = (sig_mu - cc) * inv - sgc * r;
#pragma unroll
          for (int i = 0; i < NX; ++i) rhs[NU + i] = fma_<T>(C[i], rhs_c, rhs[NU + i]);
          if (FACTOR) {
#pragma unroll
            for (int i = 0; i < NX; ++i)
#pragma unroll
              for (int i2 = 0; i2 < NX; ++i2) Pacc[i * NX + i2] = fma_<T>(sgc * C[i], C[i2], Pacc[i * NX + i2]);
          }
        }
      }
      if (FACTOR) {
        // P = Pacc + diag(Sigma_x) (+ rows' C' Sigma_c C, added above)
#pragma unroll
        for (int i = 0; i < NX; ++i) Pacc[i * NX + i] += sig[NU + i];
        // Joseph (symmetric) form of the Riccati update:
        //   PB = P B,  S = Rt + B'PB,  K = -S^-1 (PB)'A,  Acl = A + B K,  Pacc <- Q + Acl' P Acl + K' Rt K
        // with Rt = R + diag(Sigma_u).  Every term is a positive semidefinite sum: the huge barrier weights inside P
        // (lam/s ~ 1e12 on active states / rows) meet closed-loop rows Acl_i ~ 1/Sigma and drop out, whereas the short
        // form Q + A'(PA + PB K) subtracts two O(Sigma) products and loses eps * Sigma of absolute accuracy.
        T PB[NX * NU], S[NU * NU], Rt[NU * NU];
        mm<T, NX, NX, NU, false>(Pacc, B, PB);
#pragma unroll
        for (int i = 0; i < NU * NU; ++i) Rt[i] = sh[SH::oR + i];
#pragma unroll
        for (int i = 0; i < NU; ++i) Rt[i * NU + i] += sig[i];
#pragma unroll
        for (int i = 0; i < NU * NU; ++i) S[i] = Rt[i];
        mtm<T, NU, NX, NU, true>(B, PB, S);
        sym_inverse(S, Sinv);
        T G[NU * NX];
        mtm<T, NU, NX, NX, false>(PB, A, G);  // (PB)'A
#pragma unroll
        for (int i = 0; i < NU; ++i)
#pragma unroll
          for (int j = 0; j < NX; ++j) {
            T acc = T(0);
#pragma unroll
            for (int l = 0; l < NU; ++l) acc = fma_<T>(Sinv[i * NU + l], G[l * NX + j], acc);
            K[i * NX + j] = -acc;
          }
        T Acl[NX * NX], Tm[NX * NX];
#pragma unroll
        for (int i = 0; i < NX * NX; ++i) Acl[i] = A[i];
        mm<T, NX, NU, NX, true>(B, K, Acl);
        mm<T, NX, NX, NX, false>(Pacc, Acl, Tm);
        mm<T, NU, NU, NX, false>(Rt, K, G);  // G <- Rt K
#pragma unroll
        for (int i = 0; i < NX; ++i)
#pragma unroll
          for (int j = i; j < NX; ++j) {
            T acc = sh[SH::oQ + i * NX + j];
#pragma unroll
            for (int l = 0; l < NX; ++l) acc = fma_<T>(Acl[l * NX + i], Tm[l * NX + j], acc);
#pragma unroll
            for (int l = 0; l < NU; ++l) acc = fma_<T>(K[l * NX + i], G[l * NX + j], acc);
            Pacc[i * NX + j] = acc;
            Pacc[j * NX + i] = acc;
          }
        storen<NU * NX>(Kw, k, K);
        storen<NU * NU>(Sw, k, Sinv);
      }
      // h = -(rhs_x + pacc);  gu = rhs_u - B'h;  dff = Sinv gu;  pacc <- -A'h + K'gu
      T h[NX], gu[NU], dff[NU];
#pragma unroll
      for (int i = 0; i < NX; ++i) h[i] = -(rhs[NU + i] + pacc[i]);
#pragma unroll
      for (int j = 0; j < NU; ++j) {
        T acc = rhs[j];
#pragma unroll
        for (int i = 0; i < NX; ++i) acc = fma_<T>(-B[i * NU + j], h[i], acc);
        gu[j] = acc;
      }
      mv<T, NU, NU, false>(Sinv, gu, dff);
      storen<NU>(dw, k, dff);
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        T acc = T(0);
#pragma unroll
        for (int l = 0; l < NX; ++l) acc = fma_<T>(-A[l * NX + i], h[l], acc);
#pragma unroll
        for (int j = 0; j < NU; ++j) acc = fma_<T>(K[j * NX + i], gu[j], acc);
        pacc[i] = acc;
      }
    }
  }

  // inverse of a small symmetric positive definite matrix
  MPC_HD static void sym_inverse(const T* S, T* Si) {
    if constexpr (NU == 1) {
      Si[0] = rcp_(S[0]);
    } else if constexpr (NU == 2) {
      const T det = S[0] * S[3] - S[1] * S[2];
      const T id = rcp_(det);
      Si[0] = S[3] * id;
      Si[1] = -S[1] * id;
      Si[2] = -S[2] * id;
      Si[3] = S[0] * id;
    } else {
      // Gauss-Jordan without pivoting (SPD)
      T M[NU * NU];
#pragma unroll
      for (int i = 0; i < NU * NU; ++i) {
        M[i] = S[i];
        Si[i] = (i / NU == i % NU) ? T(1) : T(0);
      }
#pragma unroll
      for (int p = 0; p < NU; ++p) {
        const T inv = rcp_(M[p * NU + p]);
#pragma unroll
        for (int j = 0; j < NU; ++j) {
          M[p * NU + j] *= inv;
          Si[p * NU + j] *= inv;
        }
#pragma unroll
        for (int r = 0; r < NU; ++r) {
          if (r == p) continue;
          const T f = M[r * NU + p];
#pragma unroll
          for (int j = 0; j < NU; ++j) {
            M[r * NU + j] = fma_<T>(-f, M[p * NU + j], M[r * NU + j]);
            Si[r * NU + j] = fma_<T>(-f, Si[p * NU + j], Si[r * NU + j]);
          }
        }
      }
    }
  }

  struct Acc {
    T qmax, s0, s1, s2, dzmax, rp;  // qmax = max_i(-ds_i/s_i, -dlam_i/lam_i): largest feasible step = 1/qmax
    MPC_HD T amin() const { return qmax > T(0) ? T(1) / qmax : T(1e30); }
  };

  // ---- forward sweep: dz by rollout with the stored gains; per element the slack / multiplier
  // directions.  AFFINE: stores dz_aff and accumulates the sums that give mu_aff for any step
  // length.  Otherwise stores dz and accumulates the step ratios / norms.
  MPC_HD void forward(const bool AFFINE, T sig_mu, Acc& acc) {
    T x[NX], xn[NX], u[NU];
    acc.qmax = acc.s0 = acc.s1 = acc.s2 = acc.dzmax = acc.rp = T(0);
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = T(0);  // dx_0 = 0
    Stage cur;
    T da[D], K[NU * NX], dff[NU];
#pragma unroll
    for (int i = 0; i < D; ++i) da[i] = T(0);
    for (int k = 0; k < a.N; ++k) {
      pf_stage(k + a.pf_dist);
      pf_extra(k + a.pf_dist, true, false, true, !AFFINE, false);
      load(k, cur);
      loadn<NU * NX>(Kw, k, K);
      loadn<NU>(dw, k, dff);
      if (!AFFINE) loadn<D>(dza, k, da);
      T A[NX * NX], B[NX * NU], c[NX];
      load_model(k, A, B, c);
#pragma unroll
      for (int j = 0; j < NU; ++j) u[j] = dff[j];
      mv<T, NU, NX, true>(K, x, u);
      mv<T, NX, NX, false>(A, x, xn);
      mv<T, NX, NU, true>(B, u, xn);
      T dzv[D];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const T dz = i < NU ? u[i] : xn[i - NU];
        dzv[i] = dz;
        const T zi = cur.z[i];
        if (!AFFINE) {
          const T ad = dz < T(0) ? -dz : dz;
          acc.dzmax = ad > acc.dzmax ? ad : acc.dzmax;
        }
        if (hasl(i)) {
          const T s = cur.sl[i], l = cur.ll[i];
          const T r = zi - lo(i) - s;
          const T ds = dz + r;
          const T rinv = rcp_(s * l), inv_s = rinv * l, inv_l = rinv * s;
          const T sgl = l * inv_s;
          const T cc = AFFINE ? T(0) : cc_of(da[i], r, sgl, l);
          const T dl = (sig_mu - cc) * inv_s - l - sgl * ds;
          const T qs = -ds * inv_s, ql = -dl * inv_l;  // step is limited to 1 / max(q)
          acc.qmax = qs > acc.qmax ? qs : acc.qmax;
          acc.qmax = ql > acc.qmax ? ql : acc.qmax;
          acc.s0 += s * l;
          acc.s1 += s * dl + l * ds;
          acc.s2 += ds * dl;
          const T ar = r < T(0) ? -r : r;
          acc.rp = ar > acc.rp ? ar : acc.rp;
        }
        if (hasu(i)) {
          const T s = cur.su[i], l = cur.lu[i];
          const T r = hi(i) - zi - s;
          const T ds = -dz + r;
          const T rinv = rcp_(s * l), inv_s = rinv * l, inv_l = rinv * s;
          const T sgu = l * inv_s;
          const T cc = AFFINE ? T(0) : cc_of(-da[i], r, sgu, l);
          const T dl = (sig_mu - cc) * inv_s - l - sgu * ds;
          const T qs = -ds * inv_s, ql = -dl * inv_l;
          acc.qmax = qs > acc.qmax ? qs : acc.qmax;
          acc.qmax = ql > acc.qmax ? ql : acc.qmax;
          acc.s0 += s * l;
          acc.s1 += s * dl + l * ds;
          acc.s2 += ds * dl;
          const T ar = r < T(0) ? -r : r;
          acc.rp = ar > acc.rp ? ar : acc.rp;
        }
      }
      if constexpr (NC > 0) {
#pragma unroll 1
        for (int j = 0; j < NC; ++j) {
          T C[NX];
          const T h = load_row_c(k, j, C);
          const T s = sc[ix(k, j, NC)], l = lc[ix(k, j, NC)];
          const T r = rc[ix(k, j, NC)];
          (void)h;
          const T ds = dotx(C, xn) + r;
          const T rinv = rcp_(s * l), inv_s = rinv * l, inv_l = rinv * s;
          const T sgc = l * inv_s;
          const T cc = AFFINE ? T(0) : cc_of(dotx(C, da + NU), r, sgc, l);
          const T dl = (sig_mu - cc) * inv_s - l - sgc * ds;
          const T qs = -ds * inv_s, ql = -dl * inv_l;
          acc.qmax = qs > acc.qmax ? qs : acc.qmax;
          acc.qmax = ql > acc.qmax ? ql : acc.qmax;
          acc.s0 += s * l;
          acc.s1 += s * dl + l * ds;
          acc.s2 += ds * dl;
          const T ar = r < T(0) ? -r : r;
          acc.rp = ar > acc.rp ? ar : acc.rp;
        }
      }
      storen<D>(AFFINE ? dza : dzw, k, dzv);
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = xn[i];
    }
  }

  // ---- step: (z, s, lam) += alpha * direction; returns max |z|.  Stages are independent here.
  MPC_HD T update(T sig_mu, T alpha, bool second_order) {
    T zn = T(1);
    for (int k = 0; k < a.N; ++k) {
      pf_stage(k + a.pf_dist);
      pf_extra(k + a.pf_dist, false, false, false, second_order, true);
      Stage st;
      T da[D], dz[D];
      load(k, st);
      loadn<D>(dzw, k, dz);
      if (second_order) {
        loadn<D>(dza, k, da);
      } else {
#pragma unroll
        for (int i = 0; i < D; ++i) da[i] = T(0);
      }
      if constexpr (NC > 0) {
#pragma unroll 1
        for (int j = 0; j < NC; ++j) {
          T C[NX];
          const T h = load_row_c(k, j, C);
          const T s = sc[ix(k, j, NC)], l = lc[ix(k, j, NC)];
          const T r = rc[ix(k, j, NC)];
          (void)h;
          const T ds = dotx(C, dz + NU) + r;
          const T inv = rcp_(s), sgc = l * inv;
          const T cc = second_order ? cc_of(dotx(C, da + NU), r, sgc, l) : T(0);
          const T dl = (sig_mu - cc) * inv - l - sgc * ds;
          sc[ix(k, j, NC)] = s + alpha * ds;
          lc[ix(k, j, NC)] = l + alpha * dl;
          rc[ix(k, j, NC)] = (T(1) - alpha) * r;
        }
      }
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const T zi = st.z[i];
        if (hasl(i)) {
          const T s = st.sl[i], l = st.ll[i];
          const T r = zi - lo(i) - s;
          const T ds = dz[i] + r;
          const T inv = rcp_(s), sgl = l * inv;
          const T cc = second_order ? cc_of(da[i], r, sgl, l) : T(0);
          const T dl = (sig_mu - cc) * inv - l - sgl * ds;
          st.sl[i] = s + alpha * ds;
          st.ll[i] = l + alpha * dl;
        }
        if (hasu(i)) {
          const T s = st.su[i], l = st.lu[i];
          const T r = hi(i) - zi - s;
          const T ds = -dz[i] + r;
          const T inv = rcp_(s), sgu = l * inv;
          const T cc = second_order ? cc_of(-da[i], r, sgu, l) : T(0);
          const T dl = (sig_mu - cc) * inv - l - sgu * ds;
          st.su[i] = s + alpha * ds;
          st.lu[i] = l + alpha * dl;
        }
        const T zn_i = zi + alpha * dz[i];
        st.z[i] = zn_i;
        const T az = zn_i < T(0) ? -zn_i : zn_i;
        zn = az > zn ? az : zn;
      }
      store_stage(k, st);
    }
    return zn;
  }

  // ---- output: active set from the complementarity pairs, variables snapped onto active bounds,
  // states by rollout of the snapped inputs, cost as the reference defines it.
  MPC_HD void output(int status, int iters) {
    T x[NX], xn[NX], u[NU], A[NX * NX], B[NX * NU], c[NX];
    T cost = T(0);
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      x[i] = a.x0[i * bs + b];
      a.X[i * bs + b] = x[i];
    }
    for (int k = 0; k < a.N; ++k) {
      Stage st;
      load(k, st);
      load_model(k, A, B, c);
#pragma unroll
      for (int i = 0; i < D; ++i) {
        int sat = 0;
        if (hasl(i) && st.ll[i] > st.sl[i]) sat = -1;
        if (hasu(i) && st.lu[i] > st.su[i]) sat = 1;
        if (i < NU) {
          u[i] = sat < 0 ? lo(i) : (sat > 0 ? hi(i) : st.z[i]);
          a.U[ix(k, i, NU)] = u[i];
          if (a.sat_u) a.sat_u[ix(k, i, NU)] = (int8_t)sat;
        } else if (a.sat_x) {
          a.sat_x[ix(k, i - NU, NX)] = (int8_t)sat;
        }
      }
      if constexpr (NC > 0) {
        if (a.sat_c) {
#pragma unroll 1
          for (int j = 0; j < NC; ++j) a.sat_c[ix(k, j, NC)] = lc[ix(k, j, NC)] > sc[ix(k, j, NC)] ? (int8_t)-1 : (int8_t)0;
        }
      }
      cost += quad<T, NX>(sh + SH::oQ, x) + quad<T, NU>(sh + SH::oR, u);
      step(A, B, c, x, u, xn);
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        x[i] = xn[i];
        a.X[ix(k + 1, i, NX)] = x[i];
      }
    }
    cost += quad<T, NX>(sh + SH::oPf, x);
    a.cost[b] = cost;
    a.status[b] = status;
    a.iters[b] = iters;
  }

  MPC_HD void solve() {
    int ncons = 0;
#pragma unroll
    for (int i = 0; i < D; ++i) ncons += (hasl(i) ? 1 : 0) + (hasu(i) ? 1 : 0);
    ncons = (ncons + NC) * a.N;
    init();
    int status = MPC_UNSOLVED, it = 0;
    T rp = T(0), zn = T(1);
    // (no finite bound at all: the first iteration below is one exact Newton step onto the LQ optimum -- all barrier
    // terms vanish, alpha = 1 -- and the loop stops after it.  One call site per pass keeps the kernel's code, which
    // is far larger than the instruction caches, as small as it can be.)
    const T inv_nc = ncons ? T(1) / T(ncons) : T(0);
    while (status == MPC_UNSOLVED && it < a.max_iter) {
      ++it;
      Acc acc;
      T sig_mu = T(0);
      // predictor (phase 0: factorise, sigma = 0, no second-order term) and corrector (phase 1) run through the SAME
      // code: one copy of each sweep in the kernel instead of two
#pragma unroll 1
      for (int phase = 0; phase < 2; ++phase) {
        const bool predictor = (phase == 0);
        backward(predictor, sig_mu);
        forward(predictor, sig_mu, acc);
        if (predictor) {
          const T mu = acc.s0 * inv_nc;
          const T am_aff = acc.amin();
          const T a_aff = am_aff < T(1) ? am_aff : T(1);
          const T mu_aff = (acc.s0 + a_aff * (acc.s1 + a_aff * acc.s2)) * inv_nc;
          T ratio = mu_aff / (mu > T(1e-300) ? mu : T(1e-300));
          T sigma = ratio * ratio * ratio;
          sigma = sigma < T(1) ? sigma : T(1);
          // centring target; never below 1e-3 of the complementarity tolerance: driving mu further only inflates
          // the barrier weights (lam/s ~ lam^2/mu) and with them the rounding noise of the Newton step
          sig_mu = sigma * mu;
          const T mu_floor = T(1e-3) * a.eps * mu_scale;
          sig_mu = sig_mu > mu_floor ? sig_mu : mu_floor;
        }
      }
      T alpha = T(0.995) * acc.amin();
      alpha = alpha < T(1) ? alpha : T(1);
      zn = update(sig_mu, alpha, true);
      const T mu_new = (acc.s0 + alpha * (acc.s1 + alpha * acc.s2)) * inv_nc;
      rp = (T(1) - alpha) * acc.rp;
      const bool done = (ncons == 0) ||
                        ((mu_new <= a.eps * mu_scale) && (rp <= a.eps * zn) && (alpha * acc.dzmax <= T(1e-6) * zn));
      if (done) {
        status = MPC_SOLVED;
      } else if (!(alpha >= T(1e-6)) || !(mu_new <= T(100) * mu0)) {
        // stalled: step length collapsed / barrier parameter grew 100x above its start value.  With a bound residual that
        // cannot be closed the box and the dynamics do not meet: infeasible.
        status = (rp <= T(1e-6) * zn) ? MPC_MAX_ITER : MPC_INFEASIBLE;
      }
    }
    if (status == MPC_UNSOLVED) status = (rp <= T(1e-6) * zn) ? MPC_MAX_ITER : MPC_INFEASIBLE;
    output(status, it);
  }
};

}  // namespace mpc
