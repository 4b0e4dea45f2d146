// K4 kernel + C ABI: batched box-constrained LQ-MPC QP (interior point with Riccati Newton solves).
#include <stdlib.h>

#include "boxqp_core.cuh"

namespace mpc {

// csrc/boxqp_coop.cu: warp-per-scenario variant for state dimensions beyond one thread's registers
int64_t coop_ws_elems(int n, int m, int N, int64_t batch);
bool coop_supported(int n, int m, int ltv);
int launch_boxqp_coop(const BoxQpArgs<double>& a, int n, int m, cudaStream_t st);

constexpr int kQpThreads = 128;

// MINB = resident CTAs per SM the register allocation must allow (latency hiding for the streamed
// workspace matters more than a few spills).
template <typename T, int NX, int NU, int MINB>
__global__ void __launch_bounds__(kQpThreads, MINB) boxqp_ipm_kernel(BoxQpArgs<T> a) {
  using SH = BoxQpShared<NX, NU>;
  __shared__ T sh[SH::total];
  for (int i = threadIdx.x; i < SH::total; i += blockDim.x) {
    T v;
    if (i < SH::oB) v = a.ltv ? T(0) : a.A[i - SH::oA];
    else if (i < SH::oQ) v = a.ltv ? T(0) : a.B[i - SH::oB];
    else if (i < SH::oR) v = a.Q[i - SH::oQ];
    else if (i < SH::oPf) v = a.R[i - SH::oR];
    else if (i < SH::oLo) v = a.Pf[i - SH::oPf];
    else if (i < SH::oLo + NU) v = a.u_lo[i - SH::oLo];
    else if (i < SH::oHi) v = a.x_lo[i - SH::oLo - NU];
    else if (i < SH::oHi + NU) v = a.u_hi[i - SH::oHi];
    else v = a.x_hi[i - SH::oHi - NU];
    sh[i] = v;
  }
  __syncthreads();
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.batch) return;
  BoxQpIpm<T, NX, NU> ipm(a, sh, b);
  ipm.solve();
}

template <typename T, int NX, int NU>
static int launch_boxqp(const BoxQpArgs<T>& a_in, cudaStream_t st) {
  const unsigned grid = (unsigned)((a_in.batch + kQpThreads - 1) / kQpThreads);
  // measured on B200 (tools/prof/exp_q3.sh, cfg 3): (2,1) best at 4 CTAs/SM (128 registers) with the L2 prefetch two
  // stage visits ahead (20.1 ms per 2^18 solves; 3 CTAs: 20.4, without prefetch 21.9-22.1); (4,x) at 2 CTAs (255
  // registers) WITHOUT prefetch: those kernels stream their workspace at ~70 % of the DRAM peak, and lines prefetched
  // and evicted before use are extra traffic (tools/prof/exp_rti_minb.sh)
  int minb = (NX + NU <= 3) ? 4 : 2;
  if (const char* env = getenv("MPC_QP_MINB")) minb = atoi(env);
  BoxQpArgs<T> a = a_in;
  if (NX + NU <= 3 && !getenv("MPC_QP_PREFETCH")) a.pf_dist = 2;
  if constexpr (NX + NU > 8) {
    boxqp_ipm_kernel<T, NX, NU, 1><<<grid, kQpThreads, 0, st>>>(a);
  } else {
    if (minb >= 6) boxqp_ipm_kernel<T, NX, NU, 6><<<grid, kQpThreads, 0, st>>>(a);
    else if (minb >= 4) boxqp_ipm_kernel<T, NX, NU, 4><<<grid, kQpThreads, 0, st>>>(a);
    else if (minb >= 3) boxqp_ipm_kernel<T, NX, NU, 3><<<grid, kQpThreads, 0, st>>>(a);
    else boxqp_ipm_kernel<T, NX, NU, 2><<<grid, kQpThreads, 0, st>>>(a);
  }
  return check_launch("boxqp_ipm_kernel");
}

// K4 with general stage rows (polytopic constraints): same body, NC > 0
template <typename T, int NX, int NU, int NC>
__global__ void __launch_bounds__(kQpThreads, 2) boxqp_ipm_rows_kernel(BoxQpArgs<T> a) {
  using SH = BoxQpShared<NX, NU>;
  __shared__ T sh[SH::total];
  for (int i = threadIdx.x; i < SH::total; i += blockDim.x) {
    T v;
    if (i < SH::oB) v = a.ltv ? T(0) : a.A[i - SH::oA];
    else if (i < SH::oQ) v = a.ltv ? T(0) : a.B[i - SH::oB];
    else if (i < SH::oR) v = a.Q[i - SH::oQ];
    else if (i < SH::oPf) v = a.R[i - SH::oR];
    else if (i < SH::oLo) v = a.Pf[i - SH::oPf];
    else if (i < SH::oLo + NU) v = a.u_lo[i - SH::oLo];
    else if (i < SH::oHi) v = a.x_lo[i - SH::oLo - NU];
    else if (i < SH::oHi + NU) v = a.u_hi[i - SH::oHi];
    else v = a.x_hi[i - SH::oHi - NU];
    sh[i] = v;
  }
  __syncthreads();
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.batch) return;
  BoxQpIpm<T, NX, NU, NC> ipm(a, sh, b);
  ipm.solve();
}

template <typename T, int NX, int NU, int NC>
static int launch_boxqp_rows(const BoxQpArgs<T>& a, cudaStream_t st) {
  const unsigned grid = (unsigned)((a.batch + kQpThreads - 1) / kQpThreads);
  boxqp_ipm_rows_kernel<T, NX, NU, NC><<<grid, kQpThreads, 0, st>>>(a);
  return check_launch("boxqp_ipm_rows_kernel");
}

}  // namespace mpc

using namespace mpc;

extern "C" int64_t mpc_boxqp_rows_workspace_bytes(int64_t batch, int n, int m, int N, int nc, int dtype);

extern "C" int64_t mpc_boxqp_rows_workspace_bytes(int64_t batch, int n, int m, int N, int nc, int dtype) {
  if (batch < 0 || n < 1 || m < 1 || N < 1 || nc < 0) return 0;
  return boxqp_ws_elems(n, m, N, nc) * batch * (dtype == MPC_F32 ? 4 : 8);
}

extern "C" int64_t mpc_boxqp_workspace_bytes(int64_t batch, int n, int m, int N, int dtype) {
  if (batch < 0 || n < 1 || m < 1 || N < 1) return 0;
  const int64_t es = dtype == MPC_F32 ? 4 : 8;
  // (12,4) runs on the persistent warp-per-scenario kernel: one slot per resident warp, not per scenario
  if (coop_supported(n, m, 0)) return coop_ws_elems(n, m, N, batch) * es;
  return boxqp_ws_elems(n, m, N) * batch * es;
}

static int boxqp_solve_impl(const void* A, const void* B, const void* c, int ltv, const void* Q,
                               const void* R, const void* Pf, const void* u_lo, const void* u_hi,
                               const void* x_lo, const void* x_hi, const void* x0, const void* warm_U,
                               void* U, void* X, void* cost, int32_t* status, int32_t* iters,
                               int8_t* sat_u, int8_t* sat_x, const void* Cg, const void* hg, int nc, int8_t* sat_c,
                               void* ws, int64_t ws_bytes, int64_t batch, int n, int m, int N, int max_iter,
                               double eps, int dtype, mpc_stream_t stream) {
  MPC_REQUIRE(dtype == MPC_F64 || dtype == MPC_F32, MPC_ERR_DTYPE, "mpc_boxqp_solve: unknown dtype %d", dtype);
  MPC_REQUIRE(dtype == MPC_F64, MPC_ERR_UNSUPPORTED,
              "mpc_boxqp_solve: the interior-point iteration runs in float64 only (barrier weights span > 1e10)");
  MPC_REQUIRE(n >= 1 && n <= MPC_MAX_NX && m >= 1 && m <= MPC_MAX_NU, MPC_ERR_SHAPE, "mpc_boxqp_solve: bad (n=%d, m=%d)", n, m);
  MPC_REQUIRE(N >= 1 && batch >= 0 && max_iter >= 1, MPC_ERR_SHAPE, "mpc_boxqp_solve: bad N / batch / max_iter");
  if (batch == 0) return MPC_OK;  // nothing to do; pointers of an empty batch may be null
  MPC_REQUIRE(A && B && Q && R && Pf && u_lo && u_hi && x_lo && x_hi && x0 && U && X && cost && status && iters,
              MPC_ERR_NULL, "mpc_boxqp_solve: null pointer");
  MPC_REQUIRE(!ltv || c, MPC_ERR_NULL, "mpc_boxqp_solve: ltv model needs c");
  MPC_REQUIRE(nc >= 0 && (nc == 0 || (Cg && hg)), MPC_ERR_NULL, "mpc_boxqp_solve_rows: rows need Cg and hg");
  const int64_t need = nc ? mpc_boxqp_rows_workspace_bytes(batch, n, m, N, nc, dtype) : mpc_boxqp_workspace_bytes(batch, n, m, N, dtype);
  MPC_REQUIRE(ws && ws_bytes >= need, MPC_ERR_WORKSPACE, "mpc_boxqp_solve: workspace too small (%lld < %lld bytes)",
              (long long)ws_bytes, (long long)need);
  for (const void* p : {A, B, c, Q, R, Pf, u_lo, u_hi, x_lo, x_hi, x0, warm_U, (const void*)U, (const void*)X,
                        (const void*)cost, (const void*)ws, Cg, hg})
    MPC_REQUIRE(!p || aligned(p, 8), MPC_ERR_ALIGN, "mpc_boxqp_solve: misaligned pointer");
  BoxQpArgs<double> a{(const double*)A, (const double*)B, (const double*)c, ltv ? 1 : 0, (const double*)Q,
                      (const double*)R, (const double*)Pf, (const double*)u_lo, (const double*)u_hi,
                      (const double*)x_lo, (const double*)x_hi, (const double*)x0, (const double*)warm_U,
                      (double*)U, (double*)X, (double*)cost, status, iters, sat_u, sat_x, (const double*)Cg, (const double*)hg,
                      sat_c, (double*)ws, batch, N,
                      max_iter, eps};
  cudaStream_t st = (cudaStream_t)stream;
  a.pf_dist = 0;  // (4,x) stages: the kernels are bandwidth-bound on their workspace, see launch_boxqp and rti.cu
  if (const char* env = getenv("MPC_QP_PREFETCH")) a.pf_dist = atoi(env);
  if (nc > 0) {
    if (n == 4 && m == 2 && nc == 9) return launch_boxqp_rows<double, 4, 2, 9>(a, st);
    if (n == 4 && m == 2 && nc == 3) return launch_boxqp_rows<double, 4, 2, 3>(a, st);
    return fail(MPC_ERR_UNSUPPORTED, "mpc_boxqp_solve_rows: no kernel instantiated for n=%d m=%d nc=%d", n, m, nc);
  }
  if (n == 2 && m == 1) return launch_boxqp<double, 2, 1>(a, st);
  if (n == 4 && m == 1) return launch_boxqp<double, 4, 1>(a, st);
  if (n == 4 && m == 2) return launch_boxqp<double, 4, 2>(a, st);
  if (coop_supported(n, m, ltv)) return launch_boxqp_coop(a, n, m, st);
  return fail(MPC_ERR_UNSUPPORTED, "mpc_boxqp_solve: no kernel instantiated for n=%d m=%d", n, m);
}

extern "C" int mpc_boxqp_solve(const void* A, const void* B, const void* c, int ltv, const void* Q, const void* R,
                               const void* Pf, const void* u_lo, const void* u_hi, const void* x_lo, const void* x_hi,
                               const void* x0, const void* warm_U, void* U, void* X, void* cost, int32_t* status,
                               int32_t* iters, int8_t* sat_u, int8_t* sat_x, void* ws, int64_t ws_bytes, int64_t batch,
                               int n, int m, int N, int max_iter, double eps, int dtype, mpc_stream_t stream) {
  return boxqp_solve_impl(A, B, c, ltv, Q, R, Pf, u_lo, u_hi, x_lo, x_hi, x0, warm_U, U, X, cost, status, iters, sat_u,
                          sat_x, nullptr, nullptr, 0, nullptr, ws, ws_bytes, batch, n, m, N, max_iter, eps, dtype, stream);
}

extern "C" int mpc_boxqp_solve_rows(const void* A, const void* B, const void* c, int ltv, const void* Q, const void* R,
                                    const void* Pf, const void* u_lo, const void* u_hi, const void* x_lo,
                                    const void* x_hi, const void* Cg, const void* hg, int nc, const void* x0,
                                    const void* warm_U, void* U, void* X, void* cost, int32_t* status, int32_t* iters,
                                    int8_t* sat_u, int8_t* sat_x, int8_t* sat_c, void* ws, int64_t ws_bytes,
                                    int64_t batch, int n, int m, int N, int max_iter, double eps, int dtype,
                                    mpc_stream_t stream) {
  MPC_REQUIRE(nc >= 1, MPC_ERR_SHAPE, "mpc_boxqp_solve_rows: nc must be >= 1");
  return boxqp_solve_impl(A, B, c, ltv, Q, R, Pf, u_lo, u_hi, x_lo, x_hi, x0, warm_U, U, X, cost, status, iters, sat_u,
                          sat_x, Cg, hg, nc, sat_c, ws, ws_bytes, batch, n, m, N, max_iter, eps, dtype, stream);
}
