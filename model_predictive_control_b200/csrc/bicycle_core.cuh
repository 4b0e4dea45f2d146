// K5: kinematic bicycle (session 4) -- model, exact Jacobians, Euler / RK4 discretisation with
// chain-rule sensitivities, plant step, and the RTI preparation step (shift, roll out, linearise).
//
// The reference takes the model from rcracers.simulator.dynamics.KinematicBicycle (not vendored;
// call sites /root/reference/session_4/session4_sol.py:11,74,191,452) and the integrators from
// session4_sol.py:22-34.  OUR DEFINITION of the ODE (state [p_x, p_y, psi, v], input [a, delta]):
//     beta = atan(l_r tan(delta) / (l_r + l_f))
//     p_x' = v cos(psi + beta)   p_y' = v sin(psi + beta)   psi' = v sin(beta) / l_r
//     v'   = acceleration * a - friction * v
#pragma once

#include <math.h>

#include "boxqp_core.cuh"

namespace mpc {

template <typename T>
struct BicycleModel {
  T lr, lf, accel, ts;
  int rk4;  // OCP discretisation: 0 forward Euler (session4_sol.py:192), 1 RK4 (template.py:141)
};

template <typename T>
MPC_HD void bicycle_f(const BicycleModel<T>& p, T friction, const T* x, const T* u, T* f) {
  const T beta = atan(p.lr * tan(u[1]) / (p.lr + p.lf));
  T s, c;
  sincos(x[2] + beta, &s, &c);
  f[0] = x[3] * c;
  f[1] = x[3] * s;
  f[2] = x[3] * sin(beta) / p.lr;
  f[3] = p.accel * u[0] - friction * x[3];
}

// f, Jx = df/dx [4x4], Ju = df/du [4x2]
template <typename T>
MPC_HD void bicycle_jac(const BicycleModel<T>& p, T friction, const T* x, const T* u, T* f, T* Jx, T* Ju) {
  const T kap = p.lr / (p.lr + p.lf);
  const T td = tan(u[1]);
  const T beta = atan(kap * td);
  const T dbeta = kap * (T(1) + td * td) / (T(1) + kap * kap * td * td);
  T s, c, sb, cb;
  sincos(x[2] + beta, &s, &c);
  sincos(beta, &sb, &cb);
  const T v = x[3];
  f[0] = v * c;
  f[1] = v * s;
  f[2] = v * sb / p.lr;
  f[3] = p.accel * u[0] - friction * v;
#pragma unroll
  for (int i = 0; i < 16; ++i) Jx[i] = T(0);
#pragma unroll
  for (int i = 0; i < 8; ++i) Ju[i] = T(0);
  Jx[0 * 4 + 2] = -v * s;
  Jx[0 * 4 + 3] = c;
  Jx[1 * 4 + 2] = v * c;
  Jx[1 * 4 + 3] = s;
  Jx[2 * 4 + 3] = sb / p.lr;
  Jx[3 * 4 + 3] = -friction;
  Ju[0 * 2 + 1] = -v * s * dbeta;
  Ju[1 * 2 + 1] = v * c * dbeta;
  Ju[2 * 2 + 1] = v * cb * dbeta / p.lr;
  Ju[3 * 2 + 0] = p.accel;
}

// One RK4 stage's total derivatives: Dx = Jx (I + h Dpx), Du = Jx (h Dpu) + Ju
template <typename T>
MPC_HD void rk4_chain(const T* Jx, const T* Ju, T h, const T* Dpx, const T* Dpu, T* Dx, T* Du) {
  T M[16], Mu[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) M[i] = h * Dpx[i] + ((i / 4 == i % 4) ? T(1) : T(0));
#pragma unroll
  for (int i = 0; i < 8; ++i) Mu[i] = h * Dpu[i];
  mm<T, 4, 4, 4, false>(Jx, M, Dx);
#pragma unroll
  for (int i = 0; i < 8; ++i) Du[i] = Ju[i];
  mm<T, 4, 4, 2, true>(Jx, Mu, Du);
}

// x+ = f_d(x, u), A = d f_d / dx, B = d f_d / du
template <typename T>
MPC_HD void bicycle_discretize(const BicycleModel<T>& p, T friction, const T* x, const T* u, T* xn, T* A, T* B) {
  const T ts = p.ts;
  if (!p.rk4) {
    T f[4];
    bicycle_jac(p, friction, x, u, f, A, B);
#pragma unroll
    for (int i = 0; i < 4; ++i) xn[i] = fma_<T>(ts, f[i], x[i]);
#pragma unroll
    for (int i = 0; i < 16; ++i) A[i] = ts * A[i] + ((i / 4 == i % 4) ? T(1) : T(0));
#pragma unroll
    for (int i = 0; i < 8; ++i) B[i] = ts * B[i];
    return;
  }
  T k1[4], k2[4], k3[4], k4[4], xs[4], Jx[16], Ju[8];
  T D1x[16], D1u[8], D2x[16], D2u[8], D3x[16], D3u[8], D4x[16], D4u[8];
  bicycle_jac(p, friction, x, u, k1, D1x, D1u);
#pragma unroll
  for (int i = 0; i < 4; ++i) xs[i] = x[i] + T(0.5) * ts * k1[i];
  bicycle_jac(p, friction, xs, u, k2, Jx, Ju);
  rk4_chain(Jx, Ju, T(0.5) * ts, D1x, D1u, D2x, D2u);
#pragma unroll
  for (int i = 0; i < 4; ++i) xs[i] = x[i] + T(0.5) * ts * k2[i];
  bicycle_jac(p, friction, xs, u, k3, Jx, Ju);
  rk4_chain(Jx, Ju, T(0.5) * ts, D2x, D2u, D3x, D3u);
#pragma unroll
  for (int i = 0; i < 4; ++i) xs[i] = x[i] + ts * k3[i];
  bicycle_jac(p, friction, xs, u, k4, Jx, Ju);
  rk4_chain(Jx, Ju, ts, D3x, D3u, D4x, D4u);
  const T w = ts / T(6);
#pragma unroll
  for (int i = 0; i < 4; ++i) xn[i] = x[i] + w * (k1[i] + T(2) * k2[i] + T(2) * k3[i] + k4[i]);
#pragma unroll
  for (int i = 0; i < 16; ++i)
    A[i] = w * (D1x[i] + T(2) * D2x[i] + T(2) * D3x[i] + D4x[i]) + ((i / 4 == i % 4) ? T(1) : T(0));
#pragma unroll
  for (int i = 0; i < 8; ++i) B[i] = w * (D1u[i] + T(2) * D2u[i] + T(2) * D3u[i] + D4u[i]);
}

// Plant: forward Euler over ts (substeps == 0; the nominal model of session4_sol.py:453), classic
// RK4 with `substeps` equal sub-steps (substeps > 0), or -- substeps < 0 -- the adaptive
// Dormand-Prince 5(4) pair with rtol = atol = 10^substeps.  The adaptive mode is the counterpart of the
// reference's exact_integration (scipy odeint / LSODA, session4_sol.py:37-56); the tests check it
// against odeint itself.
template <typename T>
MPC_HD void bicycle_plant(const BicycleModel<T>& p, T friction, int substeps, T* x, const T* u) {
  if (substeps == 0) {
    T f[4];
    bicycle_f(p, friction, x, u, f);
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = fma_<T>(p.ts, f[i], x[i]);
    return;
  }
  if (substeps < 0) {
    T tol = T(1);
    for (int i = 0; i < -substeps; ++i) tol *= T(0.1);
    T t = T(0), h = p.ts;
    T k1[4], k2[4], k3[4], k4[4], k5[4], k6[4], k7[4], xs[4], xn[4];
    bicycle_f(p, friction, x, u, k1);
    for (int it = 0; it < 10000 && t < p.ts; ++it) {
      if (t + h > p.ts) h = p.ts - t;
#pragma unroll
      for (int i = 0; i < 4; ++i) xs[i] = x[i] + h * (T(1.0 / 5.0) * k1[i]);
      bicycle_f(p, friction, xs, u, k2);
#pragma unroll
      for (int i = 0; i < 4; ++i) xs[i] = x[i] + h * (T(3.0 / 40.0) * k1[i] + T(9.0 / 40.0) * k2[i]);
      bicycle_f(p, friction, xs, u, k3);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        xs[i] = x[i] + h * (T(44.0 / 45.0) * k1[i] - T(56.0 / 15.0) * k2[i] + T(32.0 / 9.0) * k3[i]);
      bicycle_f(p, friction, xs, u, k4);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        xs[i] = x[i] + h * (T(19372.0 / 6561.0) * k1[i] - T(25360.0 / 2187.0) * k2[i] + T(64448.0 / 6561.0) * k3[i] -
                            T(212.0 / 729.0) * k4[i]);
      bicycle_f(p, friction, xs, u, k5);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        xs[i] = x[i] + h * (T(9017.0 / 3168.0) * k1[i] - T(355.0 / 33.0) * k2[i] + T(46732.0 / 5247.0) * k3[i] +
                            T(49.0 / 176.0) * k4[i] - T(5103.0 / 18656.0) * k5[i]);
      bicycle_f(p, friction, xs, u, k6);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        xn[i] = x[i] + h * (T(35.0 / 384.0) * k1[i] + T(500.0 / 1113.0) * k3[i] + T(125.0 / 192.0) * k4[i] -
                            T(2187.0 / 6784.0) * k5[i] + T(11.0 / 84.0) * k6[i]);
      bicycle_f(p, friction, xn, u, k7);
      T err = T(0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const T e = h * (T(71.0 / 57600.0) * k1[i] - T(71.0 / 16695.0) * k3[i] + T(71.0 / 1920.0) * k4[i] -
                         T(17253.0 / 339200.0) * k5[i] + T(22.0 / 525.0) * k6[i] - T(1.0 / 40.0) * k7[i]);
        const T ax = x[i] < T(0) ? -x[i] : x[i], an = xn[i] < T(0) ? -xn[i] : xn[i];
        const T sc = tol + tol * (ax > an ? ax : an);
        const T r = (e < T(0) ? -e : e) / sc;
        err = r > err ? r : err;
      }
      if (err <= T(1)) {  // accept (first-same-as-last: k7 is the next k1)
        t += h;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          x[i] = xn[i];
          k1[i] = k7[i];
        }
      }
      T fac = err > T(1e-30) ? T(0.9) * (T)pow((double)(T(1) / err), 0.2) : T(5);
      fac = fac > T(5) ? T(5) : (fac < T(0.2) ? T(0.2) : fac);
      h *= fac;
    }
    return;
  }
  const T h = p.ts / T(substeps);
  for (int s = 0; s < substeps; ++s) {
    T k1[4], k2[4], k3[4], k4[4], xs[4];
    bicycle_f(p, friction, x, u, k1);
#pragma unroll
    for (int i = 0; i < 4; ++i) xs[i] = x[i] + T(0.5) * h * k1[i];
    bicycle_f(p, friction, xs, u, k2);
#pragma unroll
    for (int i = 0; i < 4; ++i) xs[i] = x[i] + T(0.5) * h * k2[i];
    bicycle_f(p, friction, xs, u, k3);
#pragma unroll
    for (int i = 0; i < 4; ++i) xs[i] = x[i] + h * k3[i];
    bicycle_f(p, friction, xs, u, k4);
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = x[i] + h / T(6) * (k1[i] + T(2) * k2[i] + T(2) * k3[i] + k4[i]);
  }
}

// Covering circles of the obstacle-avoidance variant (reference session_4/main.py:49-56,191-200): NCIRC
// circles of radius r along the vehicle axis at offsets a_i, the same for the parked obstacle at pose
// x_obs; constraint |c_i(x) - o_j|^2 >= (2 r)^2 for every pair.
constexpr int kObsCircles = 3;
template <typename T>
struct ObstacleParams {
  T a[kObsCircles];               // centre offsets along the vehicle axis
  T ox[kObsCircles], oy[kObsCircles];  // obstacle circle centres (world frame)
  T r2;                           // (r + r_p)^2
};

// Linearisation at xbar of the kObsCircles^2 collision constraints: rows C x >= h with
// C = grad g(xbar), h = r2 - g(xbar) + C xbar  (g_ij(x) = |p + a_i (cos psi, sin psi) - o_j|^2).
template <typename T, typename TIO>
MPC_HD void obstacle_rows(const ObstacleParams<T>& ob, const T* xbar, const StageRows<TIO>& Cg, const StageRows<TIO>& hg,
                          int k) {
  T sp, cp;
  sincos(xbar[2], &sp, &cp);
#pragma unroll
  for (int i = 0; i < kObsCircles; ++i) {
    const T cx = xbar[0] + ob.a[i] * cp, cy = xbar[1] + ob.a[i] * sp;
#pragma unroll
    for (int j = 0; j < kObsCircles; ++j) {
      const T dx = cx - ob.ox[j], dy = cy - ob.oy[j];
      const T g = dx * dx + dy * dy;
      const T c0 = T(2) * dx, c1 = T(2) * dy;
      const T c2 = T(2) * dx * (-ob.a[i] * sp) + T(2) * dy * (ob.a[i] * cp);
      const int row = i * kObsCircles + j;
      Cg.at(k, row * 4 + 0) = (TIO)c0;
      Cg.at(k, row * 4 + 1) = (TIO)c1;
      Cg.at(k, row * 4 + 2) = (TIO)c2;
      Cg.at(k, row * 4 + 3) = TIO(0);
      hg.at(k, row) = (TIO)(ob.r2 - g + c0 * xbar[0] + c1 * xbar[1] + c2 * xbar[2]);
    }
  }
}

// smallest clearance  min_ij |c_i(x) - o_j|^2 - r2  of a pose (negative = the covering circles overlap)
template <typename T>
MPC_HD T obstacle_clearance(const ObstacleParams<T>& ob, const T* x) {
  T sp, cp;
  sincos(x[2], &sp, &cp);
  T best = T(1e300);
#pragma unroll
  for (int i = 0; i < kObsCircles; ++i) {
    const T cx = x[0] + ob.a[i] * cp, cy = x[1] + ob.a[i] * sp;
#pragma unroll
    for (int j = 0; j < kObsCircles; ++j) {
      const T dx = cx - ob.ox[j], dy = cy - ob.oy[j];
      const T g = dx * dx + dy * dy - ob.r2;
      best = g < best ? g : best;
    }
  }
  return best;
}

// ---------------------------------------------------------------------------------------------
// RTI preparation for one scenario: shift the previous plan (first == 0), roll the nonlinear model
// out from the measured state y, linearise along the trajectory:
//   warm[k] = first ? Uprev[k] : Uprev[min(k+1, N-1)]
//   xbar_{k+1} = f_d(xbar_k, warm[k]);  A_k, B_k its Jacobians;  c_k = xbar_{k+1} - A_k xbar_k - B_k warm[k]
// Layouts: y [4][batch], Uprev / warm [N][2][batch], A [N][16][batch], B [N][8][batch], c [N][4][batch].
// Optional: obstacle rows Cg [N][9*4][batch], hg [N][9][batch] linearised at xbar_{k+1} (ob != nullptr).
// T = arithmetic type, TIO = element type of the arrays.
template <typename T, typename TIO>
MPC_HD void rti_prepare_views(const BicycleModel<T>& p, T friction, const TIO* y, const TIO* Uprev, int first, TIO* warm,
                              TIO* A, TIO* B, TIO* c, int N, int64_t bs, int64_t b, const ObstacleParams<T>* ob,
                              const StageRows<TIO>& Cg, const StageRows<TIO>& hg, const StageRows<TIO>& pack) {
  T x[4], xn[4], u[2], Ak[16], Bk[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = (T)y[i * bs + b];
  for (int k = 0; k < N; ++k) {
    const int ks = first ? k : (k + 1 < N ? k + 1 : N - 1);
    const TIO u0 = Uprev[((int64_t)ks * 2 + 0) * bs + b], u1 = Uprev[((int64_t)ks * 2 + 1) * bs + b];
    u[0] = (T)u0;
    u[1] = (T)u1;
    warm[((int64_t)k * 2 + 0) * bs + b] = u0;
    warm[((int64_t)k * 2 + 1) * bs + b] = u1;
    bicycle_discretize(p, friction, x, u, xn, Ak, Bk);
    if constexpr (sizeof(TIO) < sizeof(T)) {
      // the QP sees the stored (rounded) stage matrices: c must close the linearisation with THOSE
#pragma unroll
      for (int i = 0; i < 16; ++i) Ak[i] = (T)(TIO)Ak[i];
#pragma unroll
      for (int i = 0; i < 8; ++i) Bk[i] = (T)(TIO)Bk[i];
    }
    T ck[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      T acc = xn[i];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc = fma_<T>(-Ak[i * 4 + j], x[j], acc);
      acc = fma_<T>(-Bk[i * 2 + 0], u[0], acc);
      acc = fma_<T>(-Bk[i * 2 + 1], u[1], acc);
      ck[i] = acc;
    }
    if (pack.p) {
      // forward-Euler model only: the entries of A, B that are not structurally 0 or 1 (see BoxQpIpm, MODEL = 1)
      const T v[kBicyclePack] = {Ak[2], Ak[3], Ak[6], Ak[7], Ak[11], Ak[15], Bk[1], Bk[3], Bk[5], Bk[6],
                                 ck[0], ck[1], ck[2], ck[3]};
#pragma unroll
      for (int i = 0; i < kBicyclePack; ++i) pack.at(k, i) = (TIO)v[i];
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) A[((int64_t)k * 16 + i) * bs + b] = (TIO)Ak[i];
#pragma unroll
      for (int i = 0; i < 8; ++i) B[((int64_t)k * 8 + i) * bs + b] = (TIO)Bk[i];
#pragma unroll
      for (int i = 0; i < 4; ++i) c[((int64_t)k * 4 + i) * bs + b] = (TIO)ck[i];
    }
    if (ob) obstacle_rows<T, TIO>(*ob, xn, Cg, hg, k);
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = xn[i];
  }
}

// the same on arrays in the caller's [N][rows][batch] layout
template <typename T, typename TIO>
MPC_HD void rti_prepare_body(const BicycleModel<T>& p, T friction, const TIO* y, const TIO* Uprev, int first, TIO* warm,
                             TIO* A, TIO* B, TIO* c, int N, int64_t bs, int64_t b, const ObstacleParams<T>* ob = nullptr,
                             TIO* Cg = nullptr, TIO* hg = nullptr, TIO* pack = nullptr) {
  constexpr int R = kObsCircles * kObsCircles;
  rti_prepare_views<T, TIO>(p, friction, y, Uprev, first, warm, A, B, c, N, bs, b, ob, batch_rows(Cg, R * 4, bs, b),
                            batch_rows(hg, R, bs, b), batch_rows(pack, kBicyclePack, bs, b));
}

// out-of-line copy for the fused loop (see MPC_HD_COLD)
template <typename T, typename TIO>
MPC_HD_COLD void rti_prepare_cold(const BicycleModel<T>& p, T friction, const TIO* y, const TIO* Uprev, int first,
                                  TIO* warm, TIO* A, TIO* B, TIO* c, int N, int64_t bs, int64_t b,
                                  const ObstacleParams<T>* ob, StageRows<TIO> Cg, StageRows<TIO> hg,
                                  StageRows<TIO> pack) {
  rti_prepare_views<T, TIO>(p, friction, y, Uprev, first, warm, A, B, c, N, bs, b, ob, Cg, hg, pack);
}
template <typename T>
MPC_HD_COLD void bicycle_plant_cold(const BicycleModel<T>& p, T friction, int substeps, T* x, const T* u) {
  bicycle_plant<T>(p, friction, substeps, x, u);
}

// ---------------------------------------------------------------------------------------------
// Closed loop (session4_sol.py:443-465 / main.py:241-271 with the RTI controller): per control step
//   sqp_iters x (prepare -> LTV QP (K4)) -> apply u_0 -> plant step, with running summaries.
// sqp_iters = 1 is the real-time iteration (one linearised QP per control step, warm-started by the shifted plan);
// sqp_iters > 1 re-linearises at the new plan and solves again (full-step SQP towards the converged NLP solution the
// reference's IPOPT call returns, session4_sol.py:126-130); a round whose plan moves by <= sqp_tol ends the step.
template <typename T, typename TIO>
struct RtiLoopArgs {
  BicycleModel<T> model;    // prediction model (friction_model below)
  T friction_model;
  BicycleModel<T> plant;    // plant: its own axle distances and acceleration gain (ts shared)
  const TIO* friction_plant;  // [batch] per-scenario plant friction
  int plant_substeps;       // 0: Euler plant; > 0: RK4 sub-steps; < 0: adaptive Dormand-Prince, tol = 10^substeps
  int steps;
  int sqp_iters;
  T sqp_tol;
  int has_obstacle;
  ObstacleParams<T> ob;
  const TIO* x0;            // [4][batch]
  TIO* xcur;                // [4][batch] scratch: measured state of the current step
  TIO* Acur;                // [N][16][batch] scratch (packed: [N][14][batch])
  TIO* Bcur;                // [N][8][batch]
  TIO* ccur;                // [N][4][batch]
  TIO* warm;                // [N][2][batch]
  TIO* Cgcur;               // [N][36][batch] (obstacle rows)
  TIO* hgcur;               // [N][9][batch]
  TIO* X_cl;                // [steps+1][4][batch]
  TIO* U_cl;                // [steps][2][batch]
  TIO* cost_cl;             // [batch] sum_t x_t'Q x_t + u_t'R u_t
  TIO* viol_cl;             // [batch] max state-bound violation of the closed-loop trajectory
  TIO* clear_cl;            // optional [batch] smallest obstacle clearance |c_i - o_j|^2 - r2 along the closed loop
  int32_t* n_sat;           // [batch] number of applied inputs on a bound
  int32_t* n_fail;          // [batch] number of QPs that did not reach MPC_SOLVED
  int32_t* iters_total;     // [batch]
  TIO* X_bundle;            // optional [steps][N+1][4][batch]: the state prediction of every control step
  TIO* U_bundle;            // optional [steps][N][2][batch]: the input plan of every control step
  BoxQpArgs<TIO> qp;        // ltv = 1; A/B/c = Acur/Bcur/ccur; x0 = xcur; warm_U = warm; Cg/hg = Cgcur/hgcur
                            // U = plan buffer: initial plan on entry (zeros = cold start), last plan on exit
};

// l1 merit of the NONLINEAR OCP at the plan  warm + beta (U - warm):  cost of the nonlinear rollout from xcur plus
// rho times the summed violations of the state box (and of the collision constraints |c_i - o_j|^2 >= r2).  The
// globalisation of the SQP rounds (sqp_iters > 1): full Gauss-Newton steps cycle on this OCP (the steering weight is
// 0.01 and the curvature of the dynamics is not in the QP Hessian: the plan flips between the steering bounds), a
// backtracking line search on this merit makes every round a descent step.
constexpr double kSqpMeritRho = 100.0;
constexpr int kSqpMaxHalvings = 10;
template <typename T, typename TIO, int NC>
MPC_HD_COLD T rti_merit(const RtiLoopArgs<T, TIO>& a, const T* sh, int64_t b, T beta) {
  using SH = BoxQpShared<4, 2>;
  const int64_t bs = a.qp.batch;
  const int N = a.qp.N;
  T x[4], u[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = (T)a.xcur[i * bs + b];
  T cost = T(0), viol = T(0);
  for (int k = 0; k < N; ++k) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const T w = (T)a.warm[((int64_t)k * 2 + j) * bs + b], q = (T)a.qp.U[((int64_t)k * 2 + j) * bs + b];
      u[j] = fma_<T>(beta, q - w, w);
    }
    cost += quad<T, 4>(sh + SH::oQ, x) + quad<T, 2>(sh + SH::oR, u);
    bicycle_plant<T>(a.model, a.friction_model, a.model.rk4 ? 1 : 0, x, u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const T lo = sh[SH::oLo + 2 + i], hi = sh[SH::oHi + 2 + i];
      const T v = (lo - x[i]) > (x[i] - hi) ? (lo - x[i]) : (x[i] - hi);
      viol += v > T(0) ? v : T(0);
    }
    if constexpr (NC > 0) {
      T sp, cp;
      sincos(x[2], &sp, &cp);
#pragma unroll
      for (int i = 0; i < kObsCircles; ++i) {
        const T cx = x[0] + a.ob.a[i] * cp, cy = x[1] + a.ob.a[i] * sp;
#pragma unroll
        for (int j = 0; j < kObsCircles; ++j) {
          const T dx = cx - a.ob.ox[j], dy = cy - a.ob.oy[j];
          const T g = a.ob.r2 - (dx * dx + dy * dy);
          viol += g > T(0) ? g : T(0);
        }
      }
    }
  }
  cost += quad<T, 4>(sh + SH::oPf, x);
  return cost + T(kSqpMeritRho) * viol;
}

// PACKED = the prediction model is forward Euler: stage matrices in the packed 14-value form (a.Acur holds them).
// NC = 0 (box constraints) or 9 (obstacle rows).
template <typename T, typename TIO, bool PACKED, int NC, class ST>
MPC_HD void rti_closed_loop_body(const RtiLoopArgs<T, TIO>& a, const T* sh, int64_t b) {
  using SH = BoxQpShared<4, 2>;
  const int64_t bs = a.qp.batch;
  const int N = a.qp.N;
  T x[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    x[i] = (T)a.x0[i * bs + b];
    a.X_cl[i * bs + b] = (TIO)x[i];
  }
  const T fr_plant = (T)a.friction_plant[b];
  T cost = T(0), viol = T(0), clear = T(1e300);
  int nsat = 0, nfail = 0, itsum = 0;
  for (int t = 0; t < a.steps; ++t) {
#pragma unroll
    for (int i = 0; i < 4; ++i) a.xcur[i * bs + b] = (TIO)x[i];
    for (int round = 0; round < a.sqp_iters; ++round) {
      // forward-Euler prediction model: the packed stage model and the collision rows live in the QP's workspace tile
      // (MODEL = 4); RK4 prediction model: dense stage matrices in the side arrays (MODEL = 3)
      using Ipm = BoxQpIpm<T, TIO, 4, 2, NC, PACKED ? 4 : 3, ST>;
      Ipm ipm(a.qp, sh, b, b, bs);
      if constexpr (PACKED) {
        rti_prepare_cold<T, TIO>(a.model, a.friction_model, a.xcur, a.qp.U, (t == 0 || round > 0) ? 1 : 0, a.warm, a.Acur,
                                 a.Bcur, a.ccur, N, bs, b, NC > 0 ? &a.ob : nullptr,
                                 ipm.template view<typename Ipm::CGs>(), ipm.template view<typename Ipm::HGs>(),
                                 ipm.template view<typename Ipm::MDs>());
      } else {
        constexpr int R = kObsCircles * kObsCircles;
        rti_prepare_cold<T, TIO>(a.model, a.friction_model, a.xcur, a.qp.U, (t == 0 || round > 0) ? 1 : 0, a.warm, a.Acur,
                                 a.Bcur, a.ccur, N, bs, b, NC > 0 ? &a.ob : nullptr, batch_rows(a.Cgcur, R * 4, bs, b),
                                 batch_rows(a.hgcur, R, bs, b), StageRows<TIO>{nullptr, 0, 0});
      }
      ipm.solve();
      const bool solved = a.qp.status[b] == MPC_SOLVED;
      if (!solved) ++nfail;
      itsum += a.qp.iters[b];
      if (a.sqp_iters > 1) {
        // globalised SQP round: backtrack on the l1 merit of the nonlinear OCP, plan <- warm + beta (U_qp - warm)
        T beta = T(1);
        if (solved) {
          const T m0 = rti_merit<T, TIO, NC>(a, sh, b, T(0));
          int h = 0;
          while (h <= kSqpMaxHalvings && !(rti_merit<T, TIO, NC>(a, sh, b, beta) < m0)) {
            beta *= T(0.5);
            ++h;
          }
          if (h > kSqpMaxHalvings) beta = T(0);
        }
        T du = T(0), un = T(1);
        for (int i = 0; i < N * 2; ++i) {
          const T q = (T)a.qp.U[(int64_t)i * bs + b], w = (T)a.warm[(int64_t)i * bs + b];
          const T v = fma_<T>(beta, q - w, w);
          if (beta != T(1)) a.qp.U[(int64_t)i * bs + b] = (TIO)v;
          const T d = v > w ? v - w : w - v, av = v < T(0) ? -v : v;
          du = d > du ? d : du;
          un = av > un ? av : un;
        }
        if (du <= a.sqp_tol * un) break;   // converged (sqp_tol = 0: only an exactly stationary plan ends the rounds)
      }
    }
    if (a.X_bundle) {
      TIO* dst = a.X_bundle + (int64_t)t * (N + 1) * 4 * bs;
      for (int i = 0; i < (N + 1) * 4; ++i) dst[(int64_t)i * bs + b] = a.qp.X[(int64_t)i * bs + b];
    }
    if (a.U_bundle) {
      TIO* dst = a.U_bundle + (int64_t)t * N * 2 * bs;
      for (int i = 0; i < N * 2; ++i) dst[(int64_t)i * bs + b] = a.qp.U[(int64_t)i * bs + b];
    }
    T u[2];
    u[0] = (T)a.qp.U[(int64_t)0 * bs + b];
    u[1] = (T)a.qp.U[(int64_t)1 * bs + b];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (u[j] == (T)(TIO)sh[SH::oLo + j] || u[j] == (T)(TIO)sh[SH::oHi + j]) ++nsat;
      a.U_cl[((int64_t)t * 2 + j) * bs + b] = (TIO)u[j];
    }
    cost += quad<T, 4>(sh + SH::oQ, x) + quad<T, 2>(sh + SH::oR, u);
    bicycle_plant_cold<T>(a.plant, fr_plant, a.plant_substeps, x, u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a.X_cl[((int64_t)(t + 1) * 4 + i) * bs + b] = (TIO)x[i];
      const T lo = sh[SH::oLo + 2 + i], hi = sh[SH::oHi + 2 + i];
      const T v = (lo - x[i]) > (x[i] - hi) ? (lo - x[i]) : (x[i] - hi);
      viol = v > viol ? v : viol;
    }
    if constexpr (NC > 0) {
      const T g = obstacle_clearance<T>(a.ob, x);
      clear = g < clear ? g : clear;
    }
  }
  a.cost_cl[b] = (TIO)cost;
  a.viol_cl[b] = (TIO)viol;
  if (a.clear_cl) a.clear_cl[b] = (TIO)(NC > 0 ? clear : T(0));
  a.n_sat[b] = nsat;
  a.n_fail[b] = nfail;
  a.iters_total[b] = itsum;
}

}  // namespace mpc
