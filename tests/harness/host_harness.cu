// TEST INFRASTRUCTURE -- not part of the product library.
// Runs the __host__ __device__ per-scenario bodies of model_predictive_control_b200/csrc/*_core.cuh
// in plain CPU loops, so that `pytest -m "not gpu"` can check their algebra and indexing against
// the oracle on a machine without a GPU.  Nothing in the package links or loads this file.
#include <vector>

#include "../../model_predictive_control_b200/csrc/lq_core.cuh"
#include "../../model_predictive_control_b200/csrc/boxqp_core.cuh"
#include "../../model_predictive_control_b200/csrc/bicycle_core.cuh"

using namespace mpc;

namespace mpc {
void set_error(const char*, ...) {}
int fail(int code, const char*, ...) { return code; }
int check_launch(const char*) { return 0; }
}  // namespace mpc

template <int NX, int NU>
static void riccati_loop(const RiccatiArgs<double>& a) {
  for (int64_t b = 0; b < a.batch; ++b) riccati_body<double, NX, NU, false>(a, b);
}

extern "C" int hh_riccati(const double* A, int64_t sA, const double* B, int64_t sB, const double* Q,
                          int64_t sQ, const double* R, int64_t sR, const double* Pf, int64_t sPf,
                          double* K, double* P, int all_P, int64_t batch, int n, int m, int N) {
  RiccatiArgs<double> a{A, B, Q, R, Pf, sA, sB, sQ, sR, sPf, K, P, all_P, batch, N};
  if (n == 2 && m == 1) riccati_loop<2, 1>(a);
  else if (n == 4 && m == 1) riccati_loop<4, 1>(a);
  else if (n == 4 && m == 2) riccati_loop<4, 2>(a);
  else return -5;
  return 0;
}

template <int NX, int NU>
static void rollout_loop(const RolloutArgs<double>& a, int vec) {
  using L = RolloutSmem<double, NX, NU>;
  const bool shared = a.sA == 0 && a.sB == 0 && a.sK == 0;
  if (!shared) {
    for (int64_t b = 0; b < a.batch; ++b) rollout_perscn_body<double, NX, NU>(a, b);
    return;
  }
  std::vector<double> sm(L::total(a.ng), 0.0);
  for (int i = 0; i < NX * NX; ++i) {
    sm[L::oA + i] = a.A[i];
    sm[L::oQ + i] = a.Q ? a.Q[i] : 0.0;
    sm[L::oPf + i] = a.Pf ? a.Pf[i] : 0.0;
  }
  for (int i = 0; i < NX * NU; ++i) sm[L::oB + i] = a.B[i];
  for (int i = 0; i < NU * NU; ++i) sm[L::oR + i] = a.R ? a.R[i] : 0.0;
  for (int i = 0; i < a.ng * NU * NX; ++i) sm[L::oK + i] = a.K[(int64_t)(i / (NU * NX)) * a.sK_stage + i % (NU * NX)];
  if (vec == 2)
    for (int64_t b = 0; b < a.batch; b += 2) rollout_shared_body<double, NX, NU, 2>(a, sm.data(), b);
  else
    for (int64_t b = 0; b < a.batch; ++b) rollout_shared_body<double, NX, NU, 1>(a, sm.data(), b);
}

extern "C" int hh_lq_rollout(const double* A, int64_t sA, const double* B, int64_t sB, const double* K,
                             int64_t sK_stage, int64_t sK, int gain_offset, int gain_step,
                             const double* x0, double* X, double* U, const double* Q, const double* R,
                             const double* Pf, double* cost, uint8_t* unstable, double norm_limit,
                             int64_t batch, int n, int m, int T, int vec) {
  const int ng = (T >= 2) ? gain_offset + gain_step * (T - 2) + 1 : 0;
  RolloutArgs<double> a{A, B, sA, sB, K, sK_stage, sK, ng, gain_offset, gain_step, x0, X, U, Q, R, Pf,
                        cost, unstable, norm_limit * norm_limit, batch, T};
  if (n == 2 && m == 1) rollout_loop<2, 1>(a, vec);
  else if (n == 4 && m == 1) rollout_loop<4, 1>(a, vec);
  else if (n == 4 && m == 2) rollout_loop<4, 2>(a, vec);
  else return -5;
  return 0;
}

template <int NX, int NU>
static void lq_solve_loop(const LqSolveArgs<double>& a) {
  std::vector<double> Ks((size_t)a.N * NU * NX);
  for (int64_t b = 0; b < a.batch; ++b) lq_solve_body<double, NX, NU, false>(a, b, Ks.data(), 1);
}

extern "C" int hh_lq_solve(const double* A, int64_t sA, const double* B, int64_t sB, const double* Q,
                           int64_t sQ, const double* R, int64_t sR, const double* Pf, int64_t sPf,
                           const double* x0, double* X, double* U, double* V, double* K, double* P0,
                           int64_t batch, int n, int m, int N) {
  LqSolveArgs<double> a{A, B, Q, R, Pf, sA, sB, sQ, sR, sPf, x0, X, U, V, K, P0, batch, N};
  if (n == 2 && m == 1) lq_solve_loop<2, 1>(a);
  else if (n == 4 && m == 1) lq_solve_loop<4, 1>(a);
  else if (n == 4 && m == 2) lq_solve_loop<4, 2>(a);
  else return -5;
  return 0;
}

// Krylov-coordinate body of the single-input solve; `used[b]` = 1 where it accepted the scenario,
// 0 where it asked for the dense body (which is then run, exactly as the kernel does).
template <int NX>
static void lq_solve_krylov_loop(const LqSolveArgs<double>& a, double cond2_max, uint8_t* used) {
  std::vector<double> Ks((size_t)a.N * NX);
  for (int64_t b = 0; b < a.batch; ++b) {
    const bool ok = lq_solve_krylov_body<NX, false>(a, b, Ks.data(), 1, cond2_max);
    if (!ok) lq_solve_body<double, NX, 1, false>(a, b, Ks.data(), 1);
    if (used) used[b] = ok ? 1 : 0;
  }
}

extern "C" int hh_lq_solve_krylov(const double* A, int64_t sA, const double* B, int64_t sB, const double* Q,
                                  int64_t sQ, const double* R, int64_t sR, const double* Pf, int64_t sPf,
                                  const double* x0, double* X, double* U, double* V, int64_t batch, int n,
                                  int N, double cond_max, uint8_t* used) {
  LqSolveArgs<double> a{A, B, Q, R, Pf, sA, sB, sQ, sR, sPf, x0, X, U, V, nullptr, nullptr, batch, N};
  if (n == 2) lq_solve_krylov_loop<2>(a, cond_max * cond_max, used);
  else if (n == 4) lq_solve_krylov_loop<4>(a, cond_max * cond_max, used);
  else return -5;
  return 0;
}

template <int NX, int NU>
static void boxqp_loop(const BoxQpArgs<double>& a) {
  using SH = BoxQpShared<NX, NU>;
  std::vector<double> sh(SH::total, 0.0);
  if (!a.ltv) {
    for (int i = 0; i < NX * NX; ++i) sh[SH::oA + i] = a.A[i];
    for (int i = 0; i < NX * NU; ++i) sh[SH::oB + i] = a.B[i];
  }
  for (int i = 0; i < NX * NX; ++i) { sh[SH::oQ + i] = a.Q[i]; sh[SH::oPf + i] = a.Pf[i]; }
  for (int i = 0; i < NU * NU; ++i) sh[SH::oR + i] = a.R[i];
  for (int i = 0; i < NU; ++i) { sh[SH::oLo + i] = a.u_lo[i]; sh[SH::oHi + i] = a.u_hi[i]; }
  for (int i = 0; i < NX; ++i) { sh[SH::oLo + NU + i] = a.x_lo[i]; sh[SH::oHi + NU + i] = a.x_hi[i]; }
  for (int64_t b = 0; b < a.batch; ++b) {
    BoxQpIpm<double, NX, NU> ipm(a, sh.data(), b);
    ipm.solve();
  }
}

extern "C" int hh_boxqp_solve(const double* A, const double* B, const double* c, int ltv, const double* Q,
                              const double* R, const double* Pf, const double* u_lo, const double* u_hi,
                              const double* x_lo, const double* x_hi, const double* x0, const double* warm_U,
                              double* U, double* X, double* cost, int32_t* status, int32_t* iters,
                              int8_t* sat_u, int8_t* sat_x, int64_t batch, int n, int m, int N, int max_iter,
                              double eps) {
  std::vector<double> ws((size_t)(boxqp_ws_elems(n, m, N) * batch));
  BoxQpArgs<double> a{A, B, c, ltv, Q, R, Pf, u_lo, u_hi, x_lo, x_hi, x0, warm_U, U, X, cost, status, iters,
                      sat_u, sat_x, nullptr, nullptr, nullptr, ws.data(), batch, N, max_iter, eps};
  if (n == 2 && m == 1) boxqp_loop<2, 1>(a);
  else if (n == 4 && m == 1) boxqp_loop<4, 1>(a);
  else if (n == 4 && m == 2) boxqp_loop<4, 2>(a);
  else return -5;
  return 0;
}

extern "C" int hh_rti_prepare(double lr, double lf, double accel, double friction, double ts, int rk4, const double* y,
                              const double* Uprev, int first, double* warm, double* A, double* B, double* c,
                              int64_t batch, int N) {
  BicycleModel<double> m{lr, lf, accel, ts, rk4};
  for (int64_t b = 0; b < batch; ++b) rti_prepare_body<double>(m, friction, y, Uprev, first, warm, A, B, c, N, batch, b);
  return 0;
}

extern "C" int hh_plant_step(double lr, double lf, double accel, double ts, const double* friction, int substeps,
                             const double* x, const double* u, double* xn, int64_t batch) {
  BicycleModel<double> m{lr, lf, accel, ts, 0};
  for (int64_t b = 0; b < batch; ++b) {
    double xv[4], uv[2] = {u[b], u[batch + b]};
    for (int i = 0; i < 4; ++i) xv[i] = x[i * batch + b];
    bicycle_plant<double>(m, friction[b], substeps, xv, uv);
    for (int i = 0; i < 4; ++i) xn[i * batch + b] = xv[i];
  }
  return 0;
}

extern "C" int hh_rti_closed_loop(double lr, double lf, double accel, double friction_model, double ts, int rk4,
                                  const double* friction_plant, int plant_substeps, int steps, const double* Q,
                                  const double* R, const double* Pf, const double* u_lo, const double* u_hi,
                                  const double* x_lo, const double* x_hi, const double* x0, double* U_plan,
                                  double* X_pred, double* X_cl, double* U_cl, double* cost_cl, double* viol_cl,
                                  int32_t* n_sat, int32_t* n_fail, int32_t* iters_total, int32_t* last_status,
                                  int64_t batch, int N, int max_iter, double eps) {
  using SH = BoxQpShared<4, 2>;
  std::vector<double> sh(SH::total, 0.0);
  for (int i = 0; i < 16; ++i) { sh[SH::oQ + i] = Q[i]; sh[SH::oPf + i] = Pf[i]; }
  for (int i = 0; i < 4; ++i) sh[SH::oR + i] = R[i];
  for (int i = 0; i < 2; ++i) { sh[SH::oLo + i] = u_lo[i]; sh[SH::oHi + i] = u_hi[i]; }
  for (int i = 0; i < 4; ++i) { sh[SH::oLo + 2 + i] = x_lo[i]; sh[SH::oHi + 2 + i] = x_hi[i]; }
  std::vector<double> ws((size_t)((boxqp_ws_elems(4, 2, N) + 6 + (int64_t)N * 30) * batch));
  double* w = ws.data();
  double* qp_ws = w; w += boxqp_ws_elems(4, 2, N) * batch;
  double* xcur = w; w += 4 * batch;
  double* Acur = w; w += (int64_t)N * 16 * batch;
  double* Bcur = w; w += (int64_t)N * 8 * batch;
  double* ccur = w; w += (int64_t)N * 4 * batch;
  double* warm = w; w += (int64_t)N * 2 * batch;
  double* qp_cost = w; w += batch;
  int32_t* qp_iters = (int32_t*)w;
  RtiLoopArgs<double> a;
  a.model = BicycleModel<double>{lr, lf, accel, ts, rk4};
  a.friction_model = friction_model; a.friction_plant = friction_plant; a.plant_substeps = plant_substeps;
  a.steps = steps; a.x0 = x0; a.xcur = xcur; a.Acur = Acur; a.Bcur = Bcur; a.ccur = ccur; a.warm = warm;
  a.X_cl = X_cl; a.U_cl = U_cl; a.cost_cl = cost_cl; a.viol_cl = viol_cl; a.n_sat = n_sat; a.n_fail = n_fail;
  a.iters_total = iters_total;
  a.X_bundle = nullptr;
  a.U_bundle = nullptr;
  a.qp = BoxQpArgs<double>{Acur, Bcur, ccur, 1, Q, R, Pf, u_lo, u_hi, x_lo, x_hi, xcur, warm, U_plan, X_pred, qp_cost,
                           last_status, qp_iters, nullptr, nullptr, nullptr, nullptr, nullptr, qp_ws, batch, N, max_iter, eps};
  for (int64_t b = 0; b < batch; ++b) {
    if (rk4) rti_closed_loop_body<double, false>(a, sh.data(), b);
    else rti_closed_loop_body<double, true>(a, sh.data(), b);
  }
  return 0;
}

template <int NX, int NU, int NC>
static void boxqp_rows_loop(const BoxQpArgs<double>& a) {
  using SH = BoxQpShared<NX, NU>;
  std::vector<double> sh(SH::total, 0.0);
  if (!a.ltv) {
    for (int i = 0; i < NX * NX; ++i) sh[SH::oA + i] = a.A[i];
    for (int i = 0; i < NX * NU; ++i) sh[SH::oB + i] = a.B[i];
  }
  for (int i = 0; i < NX * NX; ++i) { sh[SH::oQ + i] = a.Q[i]; sh[SH::oPf + i] = a.Pf[i]; }
  for (int i = 0; i < NU * NU; ++i) sh[SH::oR + i] = a.R[i];
  for (int i = 0; i < NU; ++i) { sh[SH::oLo + i] = a.u_lo[i]; sh[SH::oHi + i] = a.u_hi[i]; }
  for (int i = 0; i < NX; ++i) { sh[SH::oLo + NU + i] = a.x_lo[i]; sh[SH::oHi + NU + i] = a.x_hi[i]; }
  for (int64_t b = 0; b < a.batch; ++b) {
    BoxQpIpm<double, NX, NU, NC> ipm(a, sh.data(), b);
    ipm.solve();
  }
}

extern "C" int hh_boxqp_solve_rows(const double* A, const double* B, const double* c, int ltv, const double* Q,
                                   const double* R, const double* Pf, const double* u_lo, const double* u_hi,
                                   const double* x_lo, const double* x_hi, const double* Cg, const double* hg, int nc,
                                   const double* x0, const double* warm_U, double* U, double* X, double* cost,
                                   int32_t* status, int32_t* iters, int8_t* sat_u, int8_t* sat_x, int8_t* sat_c,
                                   int64_t batch, int n, int m, int N, int max_iter, double eps) {
  std::vector<double> ws((size_t)(boxqp_ws_elems(n, m, N, nc) * batch));
  BoxQpArgs<double> a{A, B, c, ltv, Q, R, Pf, u_lo, u_hi, x_lo, x_hi, x0, warm_U, U, X, cost, status, iters,
                      sat_u, sat_x, Cg, hg, sat_c, ws.data(), batch, N, max_iter, eps};
  if (n == 4 && m == 2 && nc == 3) boxqp_rows_loop<4, 2, 3>(a);
  else if (n == 4 && m == 2 && nc == 9) boxqp_rows_loop<4, 2, 9>(a);
  else return -5;
  return 0;
}
