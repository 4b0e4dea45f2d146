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

template <typename TIO, int NX, int NU>
static std::vector<double> shared_block(const BoxQpArgs<TIO>& a) {
  using SH = BoxQpShared<NX, NU>;
  std::vector<double> sh(SH::total, 0.0);
  for (int i = 0; i < SH::total; ++i) sh[i] = boxqp_shared_elem<double, TIO, NX, NU>(a, i);
  return sh;
}

// store: 0 = all-float64 workspace, 1 = the float64 product's mixed workspace, 2 = all-float32 workspace
template <int NX, int NU, int NC, class ST>
static void boxqp_loop_st(BoxQpArgs<double> a) {
  std::vector<double> ws((size_t)(boxqp_ws_bytes<ST>(NX, NU, a.N, NC, a.batch) / 8 + 2));
  a.ws = ws.data();
  a.ws_lanes = a.batch;
  const std::vector<double> sh = shared_block<double, NX, NU>(a);
  for (int64_t b = 0; b < a.batch; ++b) {
    BoxQpIpm<double, double, NX, NU, NC, 0, ST> ipm(a, sh.data(), b, b, a.batch);
    ipm.solve();
  }
}

template <int NX, int NU, int NC>
static int boxqp_loop(const BoxQpArgs<double>& a, int store) {
  if (store == 0) boxqp_loop_st<NX, NU, NC, StoreF64>(a);
  else if (store == 1) boxqp_loop_st<NX, NU, NC, StoreMix>(a);
  else if (store == 2) boxqp_loop_st<NX, NU, NC, StoreF32>(a);
  else return -5;
  return 0;
}

extern "C" int hh_boxqp_solve(const double* A, const double* B, const double* c, int ltv, const double* Q,
                              const double* R, const double* Pf, const double* u_lo, const double* u_hi,
                              const double* x_lo, const double* x_hi, const double* x0, const double* warm_U,
                              double* U, double* X, double* cost, int32_t* status, int32_t* iters,
                              int8_t* sat_u, int8_t* sat_x, int64_t batch, int n, int m, int N, int max_iter,
                              double eps, int store) {
  BoxQpArgs<double> a{A, B, c, ltv, Q, R, Pf, u_lo, u_hi, x_lo, x_hi, x0, warm_U, U, X, cost, status, iters,
                      sat_u, sat_x, nullptr, nullptr, nullptr, nullptr, batch, N, max_iter, eps};
  if (n == 2 && m == 1) return boxqp_loop<2, 1, 0>(a, store);
  if (n == 4 && m == 1) return boxqp_loop<4, 1, 0>(a, store);
  if (n == 4 && m == 2) return boxqp_loop<4, 2, 0>(a, store);
  return -5;
}

// float32 product end to end: float arrays, float32 workspace, float64 arithmetic
extern "C" int hh_boxqp_solve_f32(const float* A, const float* B, const float* c, int ltv, const float* Q,
                                  const float* R, const float* Pf, const float* u_lo, const float* u_hi,
                                  const float* x_lo, const float* x_hi, const float* x0, const float* warm_U,
                                  float* U, float* X, float* cost, int32_t* status, int32_t* iters,
                                  int8_t* sat_u, int8_t* sat_x, int64_t batch, int n, int m, int N, int max_iter,
                                  double eps) {
  BoxQpArgs<float> a{A, B, c, ltv, Q, R, Pf, u_lo, u_hi, x_lo, x_hi, x0, warm_U, U, X, cost, status, iters,
                     sat_u, sat_x, nullptr, nullptr, nullptr, nullptr, batch, N, max_iter, eps};
  if (!(n == 2 && m == 1)) return -5;
  std::vector<double> ws((size_t)(boxqp_ws_bytes<StoreF32>(2, 1, N, 0, batch) / 8 + 2));
  a.ws = ws.data();
  a.ws_lanes = batch;
  const std::vector<double> sh = shared_block<float, 2, 1>(a);
  for (int64_t b = 0; b < batch; ++b) {
    BoxQpIpm<double, float, 2, 1, 0, 0, StoreF32> ipm(a, sh.data(), b, b, batch);
    ipm.solve();
  }
  return 0;
}

extern "C" int hh_rti_prepare(double lr, double lf, double accel, double friction, double ts, int rk4, const double* y,
                              const double* Uprev, int first, double* warm, double* A, double* B, double* c,
                              int64_t batch, int N) {
  BicycleModel<double> m{lr, lf, accel, ts, rk4};
  for (int64_t b = 0; b < batch; ++b)
    rti_prepare_body<double, double>(m, friction, y, Uprev, first, warm, A, B, c, N, batch, b);
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

// the fused closed loop; nc = 0 (box) or 9 (obstacle rows, x_obs / length / width), store as in hh_boxqp_solve
template <bool PACKED, int NC, class ST>
static void rti_loop_st(RtiLoopArgs<double, double> a, int N, int64_t batch) {
  std::vector<double> ws((size_t)(boxqp_ws_bytes<ST>(4, 2, N, NC, batch, (kBicyclePack + NC * 5) * 8) / 8 + 2));
  a.qp.ws = ws.data();
  a.qp.ws_lanes = batch;
  const std::vector<double> sh = shared_block<double, 4, 2>(a.qp);
  for (int64_t b = 0; b < batch; ++b) rti_closed_loop_body<double, double, PACKED, NC, ST>(a, sh.data(), b);
}

extern "C" int hh_rti_closed_loop(double lr, double lf, double accel, double friction_model, double ts, int rk4,
                                  double plant_lr, double plant_lf, double plant_accel,
                                  const double* friction_plant, int plant_substeps, int steps, int sqp_iters,
                                  double sqp_tol, const double* Q, const double* R, const double* Pf, const double* u_lo,
                                  const double* u_hi, const double* x_lo, const double* x_hi, int nc, double length,
                                  double width, const double* x_obs, const double* x0, double* U_plan,
                                  double* X_pred, double* X_cl, double* U_cl, double* cost_cl, double* viol_cl,
                                  double* clear_cl, int32_t* n_sat, int32_t* n_fail, int32_t* iters_total,
                                  int32_t* last_status, int64_t batch, int N, int max_iter, double eps, int store) {
  std::vector<double> side((size_t)((6 + (int64_t)N * (30 + nc * 5)) * batch));
  double* w = side.data();
  double* xcur = w; w += 4 * batch;
  double* Acur = w; w += (int64_t)N * 16 * batch;
  double* Bcur = w; w += (int64_t)N * 8 * batch;
  double* ccur = w; w += (int64_t)N * 4 * batch;
  double* warm = w; w += (int64_t)N * 2 * batch;
  double* Cgcur = nullptr; double* hgcur = nullptr;
  if (nc > 0) { Cgcur = w; w += (int64_t)N * nc * 4 * batch; hgcur = w; w += (int64_t)N * nc * batch; }
  double* qp_cost = w; w += batch;
  int32_t* qp_iters = (int32_t*)w;
  RtiLoopArgs<double, double> a;
  a.model = BicycleModel<double>{lr, lf, accel, ts, rk4};
  a.friction_model = friction_model;
  a.plant = BicycleModel<double>{plant_lr, plant_lf, plant_accel, ts, 0};
  a.friction_plant = friction_plant; a.plant_substeps = plant_substeps;
  a.steps = steps; a.sqp_iters = sqp_iters; a.sqp_tol = sqp_tol; a.has_obstacle = nc > 0;
  if (nc > 0) {
    const double d = length / (2.0 * kObsCircles);
    const double r = sqrt(d * d + width * width / 4.0);
    a.ob.r2 = (2.0 * r) * (2.0 * r);
    for (int k = 0; k < kObsCircles; ++k) {
      a.ob.a[k] = (2 * k + 1) * d - length / 2.0;
      a.ob.ox[k] = x_obs[0] + a.ob.a[k] * cos(x_obs[2]);
      a.ob.oy[k] = x_obs[1] + a.ob.a[k] * sin(x_obs[2]);
    }
  }
  a.x0 = x0; a.xcur = xcur; a.Acur = Acur; a.Bcur = Bcur; a.ccur = ccur; a.warm = warm; a.Cgcur = Cgcur; a.hgcur = hgcur;
  a.X_cl = X_cl; a.U_cl = U_cl; a.cost_cl = cost_cl; a.viol_cl = viol_cl; a.clear_cl = clear_cl; a.n_sat = n_sat;
  a.n_fail = n_fail; a.iters_total = iters_total;
  a.X_bundle = nullptr;
  a.U_bundle = nullptr;
  a.qp = BoxQpArgs<double>{Acur, Bcur, ccur, 1, Q, R, Pf, u_lo, u_hi, x_lo, x_hi, xcur, warm, U_plan, X_pred, qp_cost,
                           last_status, qp_iters, nullptr, nullptr, Cgcur, hgcur, nullptr, nullptr, batch, N, max_iter, eps};
  if (nc == 9) {
    if (rk4) return -5;
    if (store == 0) rti_loop_st<true, 9, StoreF64>(a, N, batch);
    else if (store == 1) rti_loop_st<true, 9, StoreMix>(a, N, batch);
    else rti_loop_st<true, 9, StoreF32>(a, N, batch);
    return 0;
  }
  if (nc != 0) return -5;
  if (rk4) {
    if (store == 0) rti_loop_st<false, 0, StoreF64>(a, N, batch);
    else if (store == 1) rti_loop_st<false, 0, StoreMix>(a, N, batch);
    else rti_loop_st<false, 0, StoreF32>(a, N, batch);
  } else {
    if (store == 0) rti_loop_st<true, 0, StoreF64>(a, N, batch);
    else if (store == 1) rti_loop_st<true, 0, StoreMix>(a, N, batch);
    else rti_loop_st<true, 0, StoreF32>(a, N, batch);
  }
  return 0;
}

extern "C" int hh_boxqp_solve_rows(const double* A, const double* B, const double* c, int ltv, const double* Q,
                                   const double* R, const double* Pf, const double* u_lo, const double* u_hi,
                                   const double* x_lo, const double* x_hi, const double* Cg, const double* hg, int nc,
                                   const double* x0, const double* warm_U, double* U, double* X, double* cost,
                                   int32_t* status, int32_t* iters, int8_t* sat_u, int8_t* sat_x, int8_t* sat_c,
                                   int64_t batch, int n, int m, int N, int max_iter, double eps, int store) {
  BoxQpArgs<double> a{A, B, c, ltv, Q, R, Pf, u_lo, u_hi, x_lo, x_hi, x0, warm_U, U, X, cost, status, iters,
                      sat_u, sat_x, Cg, hg, sat_c, nullptr, batch, N, max_iter, eps};
  if (n == 4 && m == 2 && nc == 3) return boxqp_loop<4, 2, 3>(a, store);
  if (n == 4 && m == 2 && nc == 9) return boxqp_loop<4, 2, 9>(a, store);
  return -5;
}
