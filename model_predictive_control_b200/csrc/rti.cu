// K5 kernels + C ABI: bicycle RTI preparation, plant step and the fused closed loop (session 4).
#include <stdlib.h>
#include <string.h>

#include "bicycle_core.cuh"

namespace mpc {

constexpr int kRtiThreads = 128;

template <typename T, typename TIO>
__global__ void __launch_bounds__(kRtiThreads) rti_prepare_kernel(BicycleModel<T> model, T friction, const TIO* y,
                                                                  const TIO* Uprev, int first, TIO* warm, TIO* A, TIO* B,
                                                                  TIO* c, int N, int64_t batch) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b < batch) rti_prepare_body<T, TIO>(model, friction, y, Uprev, first, warm, A, B, c, N, batch, b);
}

template <typename T, typename TIO>
__global__ void __launch_bounds__(kRtiThreads) rti_prepare_obstacle_kernel(BicycleModel<T> model, T friction,
                                                                           ObstacleParams<T> ob, const TIO* y,
                                                                           const TIO* Uprev, int first, TIO* warm, TIO* A,
                                                                           TIO* B, TIO* c, TIO* Cg, TIO* hg, int N,
                                                                           int64_t batch) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b < batch) rti_prepare_body<T, TIO>(model, friction, y, Uprev, first, warm, A, B, c, N, batch, b, &ob, Cg, hg);
}

template <typename T, typename TIO>
__global__ void __launch_bounds__(256) bicycle_plant_kernel(BicycleModel<T> model, const TIO* friction, int64_t sfr,
                                                            int substeps, const TIO* x, const TIO* u, TIO* xn,
                                                            int64_t batch) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  T xv[4], uv[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) xv[i] = (T)x[i * batch + b];
  uv[0] = (T)u[b];
  uv[1] = (T)u[batch + b];
  bicycle_plant<T>(model, (T)friction[b * sfr], substeps, xv, uv);
#pragma unroll
  for (int i = 0; i < 4; ++i) xn[i * batch + b] = (TIO)xv[i];
}

template <typename TIO, bool PACKED, int NC, class ST, int MINB>
__global__ void __launch_bounds__(kRtiThreads, MINB) rti_closed_loop_kernel(RtiLoopArgs<double, TIO> a) {
  using SH = BoxQpShared<4, 2>;
  __shared__ double sh[SH::total];
  for (int i = threadIdx.x; i < SH::total; i += blockDim.x) sh[i] = boxqp_shared_elem<double, TIO, 4, 2>(a.qp, i);
  __syncthreads();
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b < a.qp.batch) rti_closed_loop_body<double, TIO, PACKED, NC, ST>(a, sh, b);
}

// one globalised SQP round of the step-wise controller path: backtracking on the l1 merit (rti_merit) along
// warm -> U_qp, plan <- warm + beta (U_qp - warm); scenarios whose QP did not solve keep the full step
template <typename TIO, int NC>
__global__ void __launch_bounds__(kRtiThreads) sqp_linesearch_kernel(RtiLoopArgs<double, TIO> a, TIO* beta_out) {
  using SH = BoxQpShared<4, 2>;
  __shared__ double sh[SH::total];
  for (int i = threadIdx.x; i < SH::total; i += blockDim.x) sh[i] = boxqp_shared_elem<double, TIO, 4, 2>(a.qp, i);
  __syncthreads();
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.qp.batch) return;
  const int64_t bs = a.qp.batch;
  double beta = 1.0;
  if (a.qp.status[b] == MPC_SOLVED) {
    const double m0 = rti_merit<double, TIO, NC>(a, sh, b, 0.0);
    int h = 0;
    while (h <= kSqpMaxHalvings && !(rti_merit<double, TIO, NC>(a, sh, b, beta) < m0)) {
      beta *= 0.5;
      ++h;
    }
    if (h > kSqpMaxHalvings) beta = 0.0;
  }
  if (beta != 1.0) {
    for (int i = 0; i < a.qp.N * 2; ++i) {
      const double q = (double)a.qp.U[(int64_t)i * bs + b], w = (double)a.warm[(int64_t)i * bs + b];
      a.qp.U[(int64_t)i * bs + b] = (TIO)fma(beta, q - w, w);
    }
  }
  if (beta_out) beta_out[b] = (TIO)beta;
}

// covering circles of vehicle and obstacle (reference session_4/main.py:49-56,191-200); x_obs is a HOST pointer to
// the obstacle pose [p_x, p_y, psi, ...]
static ObstacleParams<double> make_obstacle(double length, double width, const double* x_obs) {
  ObstacleParams<double> ob;
  const double d = length / (2.0 * kObsCircles);
  const double r = sqrt(d * d + width * width / 4.0);
  ob.r2 = (2.0 * r) * (2.0 * r);
  for (int k = 0; k < kObsCircles; ++k) {
    ob.a[k] = (2 * k + 1) * d - length / 2.0;
    ob.ox[k] = x_obs[0] + ob.a[k] * cos(x_obs[2]);
    ob.oy[k] = x_obs[1] + ob.a[k] * sin(x_obs[2]);
  }
  return ob;
}

static bool al(size_t es, std::initializer_list<const void*> ps) {
  for (const void* p : ps)
    if (p && !aligned(p, es)) return false;
  return true;
}

template <typename TIO>
static int prepare_impl(const BicycleModel<double>& m, double friction, const ObstacleParams<double>* ob, const void* y,
                        const void* U_prev, int first, void* warm_U, void* A, void* B, void* c, void* Cg, void* hg,
                        int64_t batch, int N, cudaStream_t st) {
  const unsigned grid = (unsigned)((batch + kRtiThreads - 1) / kRtiThreads);
  if (ob) {
    rti_prepare_obstacle_kernel<double, TIO><<<grid, kRtiThreads, 0, st>>>(
        m, friction, *ob, (const TIO*)y, (const TIO*)U_prev, first, (TIO*)warm_U, (TIO*)A, (TIO*)B, (TIO*)c, (TIO*)Cg,
        (TIO*)hg, N, batch);
    return check_launch("rti_prepare_obstacle_kernel");
  }
  rti_prepare_kernel<double, TIO><<<grid, kRtiThreads, 0, st>>>(m, friction, (const TIO*)y, (const TIO*)U_prev, first,
                                                               (TIO*)warm_U, (TIO*)A, (TIO*)B, (TIO*)c, N, batch);
  return check_launch("rti_prepare_kernel");
}

}  // namespace mpc

using namespace mpc;

extern "C" int mpc_bicycle_rti_prepare(double lr, double lf, double accel, double friction, double ts, int rk4,
                                       const void* y, const void* U_prev, int first, void* warm_U, void* A,
                                       void* B, void* c, int64_t batch, int N, int dtype, mpc_stream_t stream) {
  MPC_REQUIRE(dtype == MPC_F64 || dtype == MPC_F32, MPC_ERR_DTYPE, "mpc_bicycle_rti_prepare: unknown dtype %d", dtype);
  if (batch == 0) return MPC_OK;  // nothing to do; pointers of an empty batch may be null
  MPC_REQUIRE(y && U_prev && warm_U && A && B && c, MPC_ERR_NULL, "mpc_bicycle_rti_prepare: null pointer");
  MPC_REQUIRE(N >= 1 && batch >= 0 && lr > 0 && lf >= 0 && ts > 0, MPC_ERR_SHAPE, "mpc_bicycle_rti_prepare: bad argument");
  MPC_REQUIRE(warm_U != U_prev, MPC_ERR_UNSUPPORTED, "mpc_bicycle_rti_prepare: warm_U must not alias U_prev");
  MPC_REQUIRE(al(dtype == MPC_F32 ? 4 : 8, {y, U_prev, warm_U, A, B, c}), MPC_ERR_ALIGN,
              "mpc_bicycle_rti_prepare: misaligned pointer");
  BicycleModel<double> m{lr, lf, accel, ts, rk4 ? 1 : 0};
  if (dtype == MPC_F32)
    return prepare_impl<float>(m, friction, nullptr, y, U_prev, first, warm_U, A, B, c, nullptr, nullptr, batch, N,
                               (cudaStream_t)stream);
  return prepare_impl<double>(m, friction, nullptr, y, U_prev, first, warm_U, A, B, c, nullptr, nullptr, batch, N,
                              (cudaStream_t)stream);
}

extern "C" int mpc_bicycle_rti_prepare_obstacle(double lr, double lf, double accel, double friction, double ts, int rk4,
                                                double length, double width, const double* x_obs, const void* y,
                                                const void* U_prev, int first, void* warm_U, void* A, void* B, void* c,
                                                void* Cg, void* hg, int64_t batch, int N, int dtype,
                                                mpc_stream_t stream) {
  MPC_REQUIRE(dtype == MPC_F64 || dtype == MPC_F32, MPC_ERR_DTYPE, "mpc_bicycle_rti_prepare_obstacle: unknown dtype %d",
              dtype);
  if (batch == 0) return MPC_OK;
  MPC_REQUIRE(x_obs && y && U_prev && warm_U && A && B && c && Cg && hg, MPC_ERR_NULL,
              "mpc_bicycle_rti_prepare_obstacle: null pointer");
  MPC_REQUIRE(N >= 1 && batch >= 0 && lr > 0 && lf >= 0 && ts > 0 && length > 0 && width > 0, MPC_ERR_SHAPE,
              "mpc_bicycle_rti_prepare_obstacle: bad argument");
  MPC_REQUIRE(warm_U != U_prev, MPC_ERR_UNSUPPORTED, "mpc_bicycle_rti_prepare_obstacle: warm_U must not alias U_prev");
  MPC_REQUIRE(al(dtype == MPC_F32 ? 4 : 8, {y, U_prev, warm_U, A, B, c, Cg, hg}), MPC_ERR_ALIGN,
              "mpc_bicycle_rti_prepare_obstacle: misaligned pointer");
  BicycleModel<double> m{lr, lf, accel, ts, rk4 ? 1 : 0};
  const ObstacleParams<double> ob = make_obstacle(length, width, x_obs);
  if (dtype == MPC_F32)
    return prepare_impl<float>(m, friction, &ob, y, U_prev, first, warm_U, A, B, c, Cg, hg, batch, N, (cudaStream_t)stream);
  return prepare_impl<double>(m, friction, &ob, y, U_prev, first, warm_U, A, B, c, Cg, hg, batch, N, (cudaStream_t)stream);
}

extern "C" int mpc_bicycle_plant_step(double lr, double lf, double accel, double ts, const void* friction,
                                      int64_t s_friction, int substeps, const void* x, const void* u, void* xn,
                                      int64_t batch, int dtype, mpc_stream_t stream) {
  MPC_REQUIRE(dtype == MPC_F64 || dtype == MPC_F32, MPC_ERR_DTYPE, "mpc_bicycle_plant_step: unknown dtype %d", dtype);
  if (batch == 0) return MPC_OK;  // nothing to do; pointers of an empty batch may be null
  MPC_REQUIRE(friction && x && u && xn, MPC_ERR_NULL, "mpc_bicycle_plant_step: null pointer");
  MPC_REQUIRE(batch >= 0 && lr > 0 && ts > 0 && substeps >= -15 && (s_friction == 0 || s_friction == 1), MPC_ERR_SHAPE,
              "mpc_bicycle_plant_step: bad argument");
  MPC_REQUIRE(al(dtype == MPC_F32 ? 4 : 8, {friction, x, u, xn}), MPC_ERR_ALIGN, "mpc_bicycle_plant_step: misaligned pointer");
  BicycleModel<double> m{lr, lf, accel, ts, 0};
  const unsigned grid = (unsigned)((batch + 255) / 256);
  if (dtype == MPC_F32)
    bicycle_plant_kernel<double, float><<<grid, 256, 0, (cudaStream_t)stream>>>(
        m, (const float*)friction, s_friction, substeps, (const float*)x, (const float*)u, (float*)xn, batch);
  else
    bicycle_plant_kernel<double, double><<<grid, 256, 0, (cudaStream_t)stream>>>(
        m, (const double*)friction, s_friction, substeps, (const double*)x, (const double*)u, (double*)xn, batch);
  return check_launch("bicycle_plant_kernel");
}

// elements of the caller's dtype the loop keeps beside the box-QP scratch, per scenario:
// measured state (4) + QP cost (1) + QP iteration count (1, int32 in an element slot of >= 4 bytes)
// + stage model (16 + 8 + 4) + warm plan (2) + obstacle rows (9 * 4 + 9 when nc > 0), the last three per stage
static int64_t rti_side_elems(int N, int nc) { return 6 + (int64_t)N * (16 + 8 + 4 + 2 + (nc > 0 ? nc * 5 : 0)); }

extern "C" int64_t mpc_rti_workspace_bytes(int64_t batch, int N, int nc, int dtype) {
  if (batch < 0 || N < 1 || nc < 0) return 0;
  const int64_t es = dtype == MPC_F32 ? 4 : 8;
  // the QP tile also holds the packed stage model and the collision rows of the forward-Euler prediction model
  const int64_t extra = (kBicyclePack + nc * 5) * es;
  const int64_t qp = dtype == MPC_F32 ? boxqp_ws_bytes<StoreF32>(4, 2, N, nc, batch, extra)
                                      : boxqp_ws_bytes<StoreF64>(4, 2, N, nc, batch, extra);
  return qp + ws_round16(rti_side_elems(N, nc) * batch * es);
}

template <typename TIO, class ST>
static int rti_loop_impl(const BicycleModel<double>& model, double friction_model, const BicycleModel<double>& plant,
                         const void* friction_plant, int plant_substeps, int steps, int sqp_iters, double sqp_tol,
                         const void* Q, const void* R, const void* Pf, const void* u_lo, const void* u_hi, const void* x_lo,
                         const void* x_hi, int nc, const ObstacleParams<double>* ob, const void* x0, void* U_plan,
                         void* X_pred, void* X_cl, void* U_cl, void* cost_cl, void* viol_cl, void* clear_cl,
                         int32_t* n_sat, int32_t* n_fail, int32_t* iters_total, int32_t* last_status, void* X_bundle,
                         void* U_bundle, void* ws, int64_t batch, int N, int max_iter, double eps, cudaStream_t st) {
  char* p = static_cast<char*>(ws);
  void* qp_ws = p;
  p += boxqp_ws_bytes<ST>(4, 2, N, nc, batch, (int64_t)(kBicyclePack + nc * 5) * (int64_t)sizeof(TIO));
  TIO* w = reinterpret_cast<TIO*>(p);
  TIO* xcur = w;
  w += 4 * batch;
  TIO* Acur = w;
  w += (int64_t)N * 16 * batch;
  TIO* Bcur = w;
  w += (int64_t)N * 8 * batch;
  TIO* ccur = w;
  w += (int64_t)N * 4 * batch;
  TIO* warm = w;
  w += (int64_t)N * 2 * batch;
  TIO* Cgcur = nullptr;
  TIO* hgcur = nullptr;
  if (nc > 0) {
    Cgcur = w;
    w += (int64_t)N * nc * 4 * batch;
    hgcur = w;
    w += (int64_t)N * nc * batch;
  }
  TIO* qp_cost = w;
  w += batch;
  int32_t* qp_iters = reinterpret_cast<int32_t*>(w);
  RtiLoopArgs<double, TIO> a;
  a.model = model;
  a.friction_model = friction_model;
  a.plant = plant;
  a.friction_plant = (const TIO*)friction_plant;
  a.plant_substeps = plant_substeps;
  a.steps = steps;
  a.sqp_iters = sqp_iters;
  a.sqp_tol = sqp_tol;
  a.has_obstacle = nc > 0;
  if (ob) a.ob = *ob;
  a.x0 = (const TIO*)x0;
  a.xcur = xcur;
  a.Acur = Acur;
  a.Bcur = Bcur;
  a.ccur = ccur;
  a.warm = warm;
  a.Cgcur = Cgcur;
  a.hgcur = hgcur;
  a.X_cl = (TIO*)X_cl;
  a.U_cl = (TIO*)U_cl;
  a.cost_cl = (TIO*)cost_cl;
  a.viol_cl = (TIO*)viol_cl;
  a.clear_cl = (TIO*)clear_cl;
  a.n_sat = n_sat;
  a.n_fail = n_fail;
  a.iters_total = iters_total;
  a.X_bundle = (TIO*)X_bundle;
  a.U_bundle = (TIO*)U_bundle;
  a.qp = BoxQpArgs<TIO>{Acur, Bcur, ccur, 1, (const TIO*)Q, (const TIO*)R, (const TIO*)Pf, (const TIO*)u_lo,
                        (const TIO*)u_hi, (const TIO*)x_lo, (const TIO*)x_hi, xcur, warm, (TIO*)U_plan, (TIO*)X_pred,
                        qp_cost, last_status, qp_iters, nullptr, nullptr, Cgcur, hgcur, nullptr, qp_ws, batch, N, max_iter,
                        eps};
  // bulk L2 prefetch of the next stage's sweep range: one visit ahead pays with box constraints (cfg 4: 2.16 -> 2.08 s per
  // 13.1 M QPs; two / three ahead 2.12 / 2.25 s), not with the collision rows (3.34 -> 3.90 s at distance 2)
  a.qp.pf_dist = nc > 0 ? 0 : 1;
  a.qp.ws_lanes = batch;
  if (const char* env = getenv("MPC_QP_PREFETCH")) a.qp.pf_dist = atoi(env);
  int threads = kRtiThreads;   // env MPC_RTI_THREADS: 32 / 64 / 128 threads per CTA (same resident warps)
  if (const char* env = getenv("MPC_RTI_THREADS")) {
    const int t = atoi(env);
    if (t == 32 || t == 64 || t == 128) threads = t;
  }
  const unsigned grid = (unsigned)((batch + threads - 1) / threads);
  // MINB = resident CTAs per SM the register allocation must allow: 2 (255 registers).  Measured at cfg 4 in round 1
  // (tools/prof/exp_rti_minb.sh): 2 CTAs/SM 2.74 s per 13.1 M QPs, 4 CTAs/SM (128 registers, spills) 3.15 s, 3: 3.96 s.
  if (nc > 0) {
    if (model.rk4) return fail(MPC_ERR_UNSUPPORTED, "mpc_rti_closed_loop: obstacle rows need the forward-Euler prediction model");
    rti_closed_loop_kernel<TIO, true, 9, ST, 2><<<grid, threads, 0, st>>>(a);
  } else if (model.rk4) {
    rti_closed_loop_kernel<TIO, false, 0, ST, 2><<<grid, threads, 0, st>>>(a);
  } else {  // forward-Euler prediction model: packed sparse stage matrices
    // MPC_RTI_PAD_SMEM=<bytes>: unused dynamic shared memory per CTA, an occupancy experiment (fewer resident CTAs with
    // the same code): tells a latency-bound kernel from a bandwidth-bound one
    int pad = 0;
    if (const char* env = getenv("MPC_RTI_PAD_SMEM")) pad = atoi(env);
    if (pad > 48 * 1024)
      cudaFuncSetAttribute(rti_closed_loop_kernel<TIO, true, 0, ST, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, pad);
    rti_closed_loop_kernel<TIO, true, 0, ST, 2><<<grid, threads, pad, st>>>(a);
  }
  return check_launch("rti_closed_loop_kernel");
}

extern "C" int mpc_rti_closed_loop(double lr, double lf, double accel, double friction_model, double ts, int rk4,
                                   double plant_lr, double plant_lf, double plant_accel, const void* friction_plant,
                                   int plant_substeps, int steps, int sqp_iters, double sqp_tol, const void* Q,
                                   const void* R, const void* Pf, const void* u_lo, const void* u_hi, const void* x_lo,
                                   const void* x_hi, int nc, double length, double width, const double* x_obs,
                                   const void* x0, void* U_plan, void* X_pred, void* X_cl, void* U_cl, void* cost_cl,
                                   void* viol_cl, void* clear_cl, int32_t* n_sat, int32_t* n_fail, int32_t* iters_total,
                                   int32_t* last_status, void* X_bundle, void* U_bundle, void* ws, int64_t ws_bytes,
                                   int64_t batch, int N, int max_iter, double eps, int dtype, mpc_stream_t stream) {
  MPC_REQUIRE(dtype == MPC_F64 || dtype == MPC_F32, MPC_ERR_DTYPE, "mpc_rti_closed_loop: unknown dtype %d", dtype);
  if (batch == 0) return MPC_OK;  // nothing to do; pointers of an empty batch may be null
  MPC_REQUIRE(friction_plant && Q && R && Pf && u_lo && u_hi && x_lo && x_hi && x0 && U_plan && X_pred && X_cl && U_cl &&
                  cost_cl && viol_cl && n_sat && n_fail && iters_total && last_status,
              MPC_ERR_NULL, "mpc_rti_closed_loop: null pointer");
  MPC_REQUIRE(N >= 1 && batch >= 0 && steps >= 0 && max_iter >= 1 && lr > 0 && ts > 0 && plant_lr > 0 &&
                  plant_substeps >= -15 && sqp_iters >= 1 && sqp_tol >= 0,
              MPC_ERR_SHAPE, "mpc_rti_closed_loop: bad argument");
  MPC_REQUIRE(nc == 0 || (nc == kObsCircles * kObsCircles && x_obs && length > 0 && width > 0), MPC_ERR_UNSUPPORTED,
              "mpc_rti_closed_loop: nc must be 0 or 9 (with x_obs, length, width)");
  MPC_REQUIRE(ws && ws_bytes >= mpc_rti_workspace_bytes(batch, N, nc, dtype), MPC_ERR_WORKSPACE,
              "mpc_rti_closed_loop: workspace too small (%lld < %lld bytes)", (long long)ws_bytes,
              (long long)mpc_rti_workspace_bytes(batch, N, nc, dtype));
  MPC_REQUIRE(al(dtype == MPC_F32 ? 4 : 8, {friction_plant, Q, R, Pf, u_lo, u_hi, x_lo, x_hi, x0, U_plan, X_pred, X_cl, U_cl,
                                           cost_cl, viol_cl, clear_cl, X_bundle, U_bundle}) && aligned(ws, 16),
              MPC_ERR_ALIGN, "mpc_rti_closed_loop: misaligned pointer");
  const BicycleModel<double> model{lr, lf, accel, ts, rk4 ? 1 : 0};
  const BicycleModel<double> plant{plant_lr, plant_lf, plant_accel, ts, 0};
  ObstacleParams<double> ob;
  if (nc > 0) ob = make_obstacle(length, width, x_obs);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MPC_F32)
    return rti_loop_impl<float, StoreF32>(model, friction_model, plant, friction_plant, plant_substeps, steps, sqp_iters,
                                          sqp_tol, Q, R, Pf, u_lo, u_hi, x_lo, x_hi, nc, nc > 0 ? &ob : nullptr, x0, U_plan,
                                          X_pred, X_cl, U_cl, cost_cl, viol_cl, clear_cl, n_sat, n_fail, iters_total,
                                          last_status, X_bundle, U_bundle, ws, batch, N, max_iter, eps, st);
  // float64 product: the all-float64 workspace is the default HERE (the (4,2) stages at 8 warps per SM are bound by
  // latency and instruction issue, not by bytes: measured at cfg 4 (tools/prof/r2_gpu10.sh) 2.91 s per 13.1 M QPs with
  // it against 3.15 s with the mixed float32/float64 workspace of mpc_boxqp_solve, whose conversions sit on every
  // load-to-use chain); MPC_QP_STORE=mix selects the mixed one (0.6x the DRAM bytes)
  const char* env = getenv("MPC_QP_STORE");
  if (env && strcmp(env, "mix") == 0)
    return rti_loop_impl<double, StoreMix>(model, friction_model, plant, friction_plant, plant_substeps, steps, sqp_iters,
                                           sqp_tol, Q, R, Pf, u_lo, u_hi, x_lo, x_hi, nc, nc > 0 ? &ob : nullptr, x0,
                                           U_plan, X_pred, X_cl, U_cl, cost_cl, viol_cl, clear_cl, n_sat, n_fail,
                                           iters_total, last_status, X_bundle, U_bundle, ws, batch, N, max_iter, eps, st);
  return rti_loop_impl<double, StoreF64>(model, friction_model, plant, friction_plant, plant_substeps, steps, sqp_iters,
                                         sqp_tol, Q, R, Pf, u_lo, u_hi, x_lo, x_hi, nc, nc > 0 ? &ob : nullptr, x0, U_plan,
                                         X_pred, X_cl, U_cl, cost_cl, viol_cl, clear_cl, n_sat, n_fail, iters_total,
                                         last_status, X_bundle, U_bundle, ws, batch, N, max_iter, eps, st);
}

extern "C" int mpc_bicycle_sqp_linesearch(double lr, double lf, double accel, double friction, double ts, int rk4,
                                          const void* Q, const void* R, const void* Pf, const void* x_lo,
                                          const void* x_hi, int nc, double length, double width, const double* x_obs,
                                          const void* y, const void* warm_U, void* U, const int32_t* status,
                                          void* beta, int64_t batch, int N, int dtype, mpc_stream_t stream) {
  MPC_REQUIRE(dtype == MPC_F64 || dtype == MPC_F32, MPC_ERR_DTYPE, "mpc_bicycle_sqp_linesearch: unknown dtype %d", dtype);
  if (batch == 0) return MPC_OK;
  MPC_REQUIRE(Q && R && Pf && x_lo && x_hi && y && warm_U && U && status, MPC_ERR_NULL,
              "mpc_bicycle_sqp_linesearch: null pointer");
  MPC_REQUIRE(N >= 1 && batch >= 0 && lr > 0 && ts > 0, MPC_ERR_SHAPE, "mpc_bicycle_sqp_linesearch: bad argument");
  MPC_REQUIRE(nc == 0 || (nc == kObsCircles * kObsCircles && x_obs && length > 0 && width > 0), MPC_ERR_UNSUPPORTED,
              "mpc_bicycle_sqp_linesearch: nc must be 0 or 9 (with x_obs, length, width)");
  MPC_REQUIRE(al(dtype == MPC_F32 ? 4 : 8, {Q, R, Pf, x_lo, x_hi, y, warm_U, U, beta}), MPC_ERR_ALIGN,
              "mpc_bicycle_sqp_linesearch: misaligned pointer");
  const unsigned grid = (unsigned)((batch + kRtiThreads - 1) / kRtiThreads);
  cudaStream_t st = (cudaStream_t)stream;
  auto launch = [&](auto tag) {
    using TIO = decltype(tag);
    RtiLoopArgs<double, TIO> a{};
    a.model = BicycleModel<double>{lr, lf, accel, ts, rk4 ? 1 : 0};
    a.friction_model = friction;
    if (nc > 0) a.ob = make_obstacle(length, width, x_obs);
    a.xcur = (TIO*)y;
    a.warm = (TIO*)warm_U;
    a.qp = BoxQpArgs<TIO>{};
    a.qp.ltv = 1;
    a.qp.Q = (const TIO*)Q;
    a.qp.R = (const TIO*)R;
    a.qp.Pf = (const TIO*)Pf;
    // the input box is not needed by the merit (the QP solution and the warm plan are inside it): reuse the state box
    a.qp.u_lo = (const TIO*)x_lo;
    a.qp.u_hi = (const TIO*)x_hi;
    a.qp.x_lo = (const TIO*)x_lo;
    a.qp.x_hi = (const TIO*)x_hi;
    a.qp.U = (TIO*)U;
    a.qp.status = const_cast<int32_t*>(status);
    a.qp.batch = batch;
    a.qp.N = N;
    if (nc > 0) sqp_linesearch_kernel<TIO, 9><<<grid, kRtiThreads, 0, st>>>(a, (TIO*)beta);
    else sqp_linesearch_kernel<TIO, 0><<<grid, kRtiThreads, 0, st>>>(a, (TIO*)beta);
  };
  if (dtype == MPC_F32) launch(float{});
  else launch(double{});
  return check_launch("sqp_linesearch_kernel");
}
