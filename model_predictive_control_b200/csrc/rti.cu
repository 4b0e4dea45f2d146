// K5 kernels + C ABI: bicycle RTI preparation, plant step and the fused closed loop (session 4).
#include <stdlib.h>

#include "bicycle_core.cuh"

namespace mpc {

constexpr int kRtiThreads = 128;

template <typename T>
__global__ void __launch_bounds__(kRtiThreads) rti_prepare_kernel(BicycleModel<T> model, T friction, const T* y,
                                                                  const T* Uprev, int first, T* warm, T* A, T* B, T* c,
                                                                  int N, int64_t batch) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b < batch) rti_prepare_body<T>(model, friction, y, Uprev, first, warm, A, B, c, N, batch, b);
}

template <typename T>
__global__ void __launch_bounds__(kRtiThreads) rti_prepare_obstacle_kernel(BicycleModel<T> model, T friction,
                                                                           ObstacleParams<T> ob, const T* y,
                                                                           const T* Uprev, int first, T* warm, T* A, T* B,
                                                                           T* c, T* Cg, T* hg, int N, int64_t batch) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b < batch) rti_prepare_body<T>(model, friction, y, Uprev, first, warm, A, B, c, N, batch, b, &ob, Cg, hg);
}

template <typename T>
__global__ void __launch_bounds__(256) bicycle_plant_kernel(BicycleModel<T> model, const T* friction, int64_t sfr,
                                                            int substeps, const T* x, const T* u, T* xn, int64_t batch) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  T xv[4], uv[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) xv[i] = x[i * batch + b];
  uv[0] = u[b];
  uv[1] = u[batch + b];
  bicycle_plant<T>(model, friction[b * sfr], substeps, xv, uv);
#pragma unroll
  for (int i = 0; i < 4; ++i) xn[i * batch + b] = xv[i];
}

template <typename T, bool PACKED, int MINB>
__global__ void __launch_bounds__(kRtiThreads, MINB) rti_closed_loop_kernel(RtiLoopArgs<T> a) {
  using SH = BoxQpShared<4, 2>;
  __shared__ T sh[SH::total];
  for (int i = threadIdx.x; i < SH::total; i += blockDim.x) {
    T v;
    if (i < SH::oQ) v = T(0);
    else if (i < SH::oR) v = a.qp.Q[i - SH::oQ];
    else if (i < SH::oPf) v = a.qp.R[i - SH::oR];
    else if (i < SH::oLo) v = a.qp.Pf[i - SH::oPf];
    else if (i < SH::oLo + 2) v = a.qp.u_lo[i - SH::oLo];
    else if (i < SH::oHi) v = a.qp.x_lo[i - SH::oLo - 2];
    else if (i < SH::oHi + 2) v = a.qp.u_hi[i - SH::oHi];
    else v = a.qp.x_hi[i - SH::oHi - 2];
    sh[i] = v;
  }
  __syncthreads();
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b < a.qp.batch) rti_closed_loop_body<T, PACKED>(a, sh, b);
}

}  // namespace mpc

using namespace mpc;

static bool al8(std::initializer_list<const void*> ps) {
  for (const void* p : ps)
    if (p && !aligned(p, 8)) return false;
  return true;
}

extern "C" int mpc_bicycle_rti_prepare(double lr, double lf, double accel, double friction, double ts, int rk4,
                                       const void* y, const void* U_prev, int first, void* warm_U, void* A,
                                       void* B, void* c, int64_t batch, int N, int dtype, mpc_stream_t stream) {
  MPC_REQUIRE(dtype == MPC_F64, dtype == MPC_F32 ? MPC_ERR_UNSUPPORTED : MPC_ERR_DTYPE,
              "mpc_bicycle_rti_prepare: float64 only (dtype %d)", dtype);
  if (batch == 0) return MPC_OK;  // nothing to do; pointers of an empty batch may be null
  MPC_REQUIRE(y && U_prev && warm_U && A && B && c, MPC_ERR_NULL, "mpc_bicycle_rti_prepare: null pointer");
  MPC_REQUIRE(N >= 1 && batch >= 0 && lr > 0 && lf >= 0 && ts > 0, MPC_ERR_SHAPE, "mpc_bicycle_rti_prepare: bad argument");
  MPC_REQUIRE(warm_U != U_prev, MPC_ERR_UNSUPPORTED, "mpc_bicycle_rti_prepare: warm_U must not alias U_prev");
  MPC_REQUIRE(al8({y, U_prev, warm_U, A, B, c}), MPC_ERR_ALIGN, "mpc_bicycle_rti_prepare: misaligned pointer");
  BicycleModel<double> m{lr, lf, accel, ts, rk4 ? 1 : 0};
  rti_prepare_kernel<double><<<(unsigned)((batch + kRtiThreads - 1) / kRtiThreads), kRtiThreads, 0, (cudaStream_t)stream>>>(
      m, friction, (const double*)y, (const double*)U_prev, first, (double*)warm_U, (double*)A, (double*)B, (double*)c, N,
      batch);
  return check_launch("rti_prepare_kernel");
}

extern "C" int mpc_bicycle_rti_prepare_obstacle(double lr, double lf, double accel, double friction, double ts, int rk4,
                                                double length, double width, const double* x_obs, const void* y,
                                                const void* U_prev, int first, void* warm_U, void* A, void* B, void* c,
                                                void* Cg, void* hg, int64_t batch, int N, int dtype,
                                                mpc_stream_t stream) {
  MPC_REQUIRE(dtype == MPC_F64, dtype == MPC_F32 ? MPC_ERR_UNSUPPORTED : MPC_ERR_DTYPE,
              "mpc_bicycle_rti_prepare_obstacle: float64 only (dtype %d)", dtype);
  if (batch == 0) return MPC_OK;
  MPC_REQUIRE(x_obs && y && U_prev && warm_U && A && B && c && Cg && hg, MPC_ERR_NULL,
              "mpc_bicycle_rti_prepare_obstacle: null pointer");
  MPC_REQUIRE(N >= 1 && batch >= 0 && lr > 0 && lf >= 0 && ts > 0 && length > 0 && width > 0, MPC_ERR_SHAPE,
              "mpc_bicycle_rti_prepare_obstacle: bad argument");
  MPC_REQUIRE(warm_U != U_prev, MPC_ERR_UNSUPPORTED, "mpc_bicycle_rti_prepare_obstacle: warm_U must not alias U_prev");
  MPC_REQUIRE(al8({y, U_prev, warm_U, A, B, c, Cg, hg}), MPC_ERR_ALIGN, "mpc_bicycle_rti_prepare_obstacle: misaligned pointer");
  BicycleModel<double> m{lr, lf, accel, ts, rk4 ? 1 : 0};
  // covering circles (x_obs is a HOST pointer to the obstacle pose [p_x, p_y, psi, v])
  ObstacleParams<double> ob;
  const double d = length / (2.0 * kObsCircles);
  const double r = sqrt(d * d + width * width / 4.0);
  ob.r2 = (2.0 * r) * (2.0 * r);
  for (int k = 0; k < kObsCircles; ++k) {
    ob.a[k] = (2 * k + 1) * d - length / 2.0;
    ob.ox[k] = x_obs[0] + ob.a[k] * cos(x_obs[2]);
    ob.oy[k] = x_obs[1] + ob.a[k] * sin(x_obs[2]);
  }
  rti_prepare_obstacle_kernel<double><<<(unsigned)((batch + kRtiThreads - 1) / kRtiThreads), kRtiThreads, 0, (cudaStream_t)stream>>>(
      m, friction, ob, (const double*)y, (const double*)U_prev, first, (double*)warm_U, (double*)A, (double*)B, (double*)c,
      (double*)Cg, (double*)hg, N, batch);
  return check_launch("rti_prepare_obstacle_kernel");
}

extern "C" int mpc_bicycle_plant_step(double lr, double lf, double accel, double ts, const void* friction,
                                      int64_t s_friction, int substeps, const void* x, const void* u, void* xn,
                                      int64_t batch, int dtype, mpc_stream_t stream) {
  MPC_REQUIRE(dtype == MPC_F64, dtype == MPC_F32 ? MPC_ERR_UNSUPPORTED : MPC_ERR_DTYPE,
              "mpc_bicycle_plant_step: float64 only (dtype %d)", dtype);
  if (batch == 0) return MPC_OK;  // nothing to do; pointers of an empty batch may be null
  MPC_REQUIRE(friction && x && u && xn, MPC_ERR_NULL, "mpc_bicycle_plant_step: null pointer");
  MPC_REQUIRE(batch >= 0 && lr > 0 && ts > 0 && substeps >= -15 && (s_friction == 0 || s_friction == 1), MPC_ERR_SHAPE,
              "mpc_bicycle_plant_step: bad argument");
  MPC_REQUIRE(al8({friction, x, u, xn}), MPC_ERR_ALIGN, "mpc_bicycle_plant_step: misaligned pointer");
  BicycleModel<double> m{lr, lf, accel, ts, 0};
  bicycle_plant_kernel<double><<<(unsigned)((batch + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      m, (const double*)friction, s_friction, substeps, (const double*)x, (const double*)u, (double*)xn, batch);
  return check_launch("bicycle_plant_kernel");
}

extern "C" int64_t mpc_rti_workspace_bytes(int64_t batch, int N, int dtype) {
  if (batch < 0 || N < 1) return 0;
  const int64_t es = dtype == MPC_F32 ? 4 : 8;
  // box-QP scratch + measured state (4) + per-step QP cost and iteration count (2) + A, B, c, warm plan
  return (boxqp_ws_elems(4, 2, N) + 6 + (int64_t)N * (16 + 8 + 4 + 2)) * batch * es;
}

extern "C" int mpc_rti_closed_loop(double lr, double lf, double accel, double friction_model, double ts, int rk4,
                                   const void* friction_plant, int plant_substeps, int steps, const void* Q,
                                   const void* R, const void* Pf, const void* u_lo, const void* u_hi,
                                   const void* x_lo, const void* x_hi, const void* x0, void* U_plan, void* X_pred,
                                   void* X_cl, void* U_cl, void* cost_cl, void* viol_cl, int32_t* n_sat,
                                   int32_t* n_fail, int32_t* iters_total, int32_t* last_status, void* X_bundle,
                                   void* U_bundle, void* ws, int64_t ws_bytes, int64_t batch, int N, int max_iter,
                                   double eps, int dtype, mpc_stream_t stream) {
  MPC_REQUIRE(dtype == MPC_F64, dtype == MPC_F32 ? MPC_ERR_UNSUPPORTED : MPC_ERR_DTYPE,
              "mpc_rti_closed_loop: float64 only (dtype %d)", dtype);
  if (batch == 0) return MPC_OK;  // nothing to do; pointers of an empty batch may be null
  MPC_REQUIRE(friction_plant && Q && R && Pf && u_lo && u_hi && x_lo && x_hi && x0 && U_plan && X_pred && X_cl && U_cl &&
                  cost_cl && viol_cl && n_sat && n_fail && iters_total && last_status,
              MPC_ERR_NULL, "mpc_rti_closed_loop: null pointer");
  MPC_REQUIRE(N >= 1 && batch >= 0 && steps >= 0 && max_iter >= 1 && lr > 0 && ts > 0 && plant_substeps >= -15, MPC_ERR_SHAPE,
              "mpc_rti_closed_loop: bad argument");
  MPC_REQUIRE(ws && ws_bytes >= mpc_rti_workspace_bytes(batch, N, dtype), MPC_ERR_WORKSPACE,
              "mpc_rti_closed_loop: workspace too small (%lld < %lld bytes)", (long long)ws_bytes,
              (long long)mpc_rti_workspace_bytes(batch, N, dtype));
  MPC_REQUIRE(al8({friction_plant, Q, R, Pf, u_lo, u_hi, x_lo, x_hi, x0, U_plan, X_pred, X_cl, U_cl, cost_cl, viol_cl, ws}),
              MPC_ERR_ALIGN, "mpc_rti_closed_loop: misaligned pointer");
  double* w = (double*)ws;
  double* qp_ws = w;
  w += boxqp_ws_elems(4, 2, N) * batch;
  double* xcur = w;
  w += 4 * batch;
  double* Acur = w;
  w += (int64_t)N * 16 * batch;
  double* Bcur = w;
  w += (int64_t)N * 8 * batch;
  double* ccur = w;
  w += (int64_t)N * 4 * batch;
  double* warm = w;
  w += (int64_t)N * 2 * batch;
  double* qp_cost = w;
  w += batch;
  int32_t* qp_iters = (int32_t*)w;
  RtiLoopArgs<double> a;
  a.model = BicycleModel<double>{lr, lf, accel, ts, rk4 ? 1 : 0};
  a.friction_model = friction_model;
  a.friction_plant = (const double*)friction_plant;
  a.plant_substeps = plant_substeps;
  a.steps = steps;
  a.x0 = (const double*)x0;
  a.xcur = xcur;
  a.Acur = Acur;
  a.Bcur = Bcur;
  a.ccur = ccur;
  a.warm = warm;
  a.X_cl = (double*)X_cl;
  a.U_cl = (double*)U_cl;
  a.cost_cl = (double*)cost_cl;
  a.viol_cl = (double*)viol_cl;
  a.n_sat = n_sat;
  a.n_fail = n_fail;
  a.iters_total = iters_total;
  a.X_bundle = (double*)X_bundle;
  a.U_bundle = (double*)U_bundle;
  a.qp = BoxQpArgs<double>{Acur, Bcur, ccur, 1, (const double*)Q, (const double*)R, (const double*)Pf,
                           (const double*)u_lo, (const double*)u_hi, (const double*)x_lo, (const double*)x_hi, xcur, warm,
                           (double*)U_plan, (double*)X_pred, qp_cost, last_status, qp_iters, nullptr, nullptr, nullptr, nullptr, nullptr,
                           qp_ws, batch, N, max_iter, eps};
  a.qp.pf_dist = 0;
  if (const char* env = getenv("MPC_QP_PREFETCH")) a.qp.pf_dist = atoi(env);
  const unsigned grid = (unsigned)((batch + kRtiThreads - 1) / kRtiThreads);
  // MINB = resident CTAs per SM the register allocation must allow: 2 (255 registers) is the default.  Measured at
  // cfg 4 with the final code of round 1 (tools/prof/exp_rti_minb.sh, 13.1 M QPs): 2 CTAs/SM 2.74 s, 4 CTAs/SM (128
  // registers, spills, one wave instead of 1.7) 3.15 s, 3 CTAs/SM 3.96 s; with the L2 prefetch of the workspace rows
  // 3.44 s at 2 CTAs/SM -- the kernel streams its workspace at ~70 % of the DRAM peak, so prefetches that are evicted
  // before use only add traffic.  (Earlier in the round, with 25 % more code and an 11-instruction reciprocal, the
  // same kernel was latency-bound and both the prefetch and the one-wave variant paid.)
  int minb = 2;
  if (const char* env = getenv("MPC_RTI_MINB")) minb = atoi(env);
  cudaStream_t st = (cudaStream_t)stream;
  if (rk4) {
    rti_closed_loop_kernel<double, false, 2><<<grid, kRtiThreads, 0, st>>>(a);
  } else {  // forward-Euler prediction model: packed sparse stage matrices
    if (minb >= 4) rti_closed_loop_kernel<double, true, 4><<<grid, kRtiThreads, 0, st>>>(a);
    else if (minb == 3) rti_closed_loop_kernel<double, true, 3><<<grid, kRtiThreads, 0, st>>>(a);
    else rti_closed_loop_kernel<double, true, 2><<<grid, kRtiThreads, 0, st>>>(a);
  }
  return check_launch("rti_closed_loop_kernel");
}
