// K4 kernel + C ABI: batched box-constrained LQ-MPC QP (interior point with Riccati Newton solves).
#include <stdlib.h>
#include <string.h>

#include "boxqp_core.cuh"

namespace mpc {

// csrc/boxqp_coop.cu: warp-per-scenario variant for state dimensions beyond one thread's registers
int64_t coop_ws_elems(int n, int m, int N, int64_t batch);
bool coop_supported(int n, int m, int ltv);
int launch_boxqp_coop(const BoxQpArgs<double>& a, int n, int m, cudaStream_t st);

constexpr int kQpThreads = 128;
// the first bytes of the caller's workspace hold the scenario queue of the refill kernel; a full 256 bytes, so that every
// workspace row stays aligned to the 128-byte lines a warp reads and writes
constexpr int64_t kWsHeader = 256;

// MINB = resident CTAs per SM the register allocation must allow (latency hiding for the streamed
// workspace matters more than a few spills).  NC > 0: general stage rows (polytopic constraints), same body.
// LTV = the stage model: false = shared (A, B), true = per-scenario (A_k, B_k, c_k) arrays; a compile-time choice, so
// that the other case's loads and address arithmetic are not in the instruction stream
template <typename TIO, class ST, int NX, int NU, int NC, int MINB, bool LTV>
__global__ void __launch_bounds__(kQpThreads, MINB) boxqp_ipm_kernel(BoxQpArgs<TIO> a) {
  using SH = BoxQpShared<NX, NU>;
  __shared__ double sh[SH::total];
  for (int i = threadIdx.x; i < SH::total; i += blockDim.x) sh[i] = boxqp_shared_elem<double, TIO, NX, NU>(a, i);
  __syncthreads();
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.batch) return;
  // lane b of the workspace solves scenario order[b] (or b): with scenarios ordered along a space-filling curve of
  // their initial states, the lanes of a warp see similar active sets and finish after similar iteration counts
  const int64_t scn = a.order ? (int64_t)a.order[b] : b;
  BoxQpIpm<double, TIO, NX, NU, NC, LTV ? 3 : 2, ST> ipm(a, sh, scn, b, a.batch);
  ipm.solve();
}

// The same solve with the warp's workspace tile STREAMED THROUGH SHARED MEMORY (shared LTI model, no general rows):
// the read set of every sweep is one contiguous byte range of a tile stage (boxqp_core.cuh, section order), which one
// lane copies kDepth stage visits ahead with cp.async.bulk (completion on the warp's mbarrier); the sweeps read shared
// memory (29-cycle loads instead of L2 / HBM round trips on the critical path of every stage visit) and store to
// global memory directly.  Dynamic shared memory: kDepth buffers x kBufBytes per warp.
template <typename TIO, class ST, int NX, int NU, int MINB>
__global__ void __launch_bounds__(kQpThreads, MINB) boxqp_ipm_staged_kernel(BoxQpArgs<TIO> a) {
  using SH = BoxQpShared<NX, NU>;
  using Ipm = BoxQpIpm<double, TIO, NX, NU, 0, 2, ST, true>;
  extern __shared__ __align__(128) char qp_stage_buffers[];
  __shared__ double sh[SH::total];
  __shared__ unsigned long long bars[Ipm::kDepth * kQpThreads / 32];
  for (int i = threadIdx.x; i < SH::total; i += blockDim.x) sh[i] = boxqp_shared_elem<double, TIO, NX, NU>(a, i);
  const int warp = threadIdx.x / 32;
  if (threadIdx.x % 32 == 0) {
    for (int j = 0; j < Ipm::kDepth; ++j)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(bars + Ipm::kDepth * warp + j)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lanes = __ballot_sync(0xffffffffu, b < a.batch);
  if (b >= a.batch) return;
  const int64_t scn = a.order ? (int64_t)a.order[b] : b;
  Ipm ipm(a, sh, scn, b, a.batch);
  ipm.stage_setup(qp_stage_buffers + (size_t)warp * Ipm::kDepth * Ipm::kBufBytes, bars + Ipm::kDepth * warp, lanes);
  ipm.solve();
}

template <typename TIO, class ST, int NX, int NU, int MINB>
static int launch_staged(const BoxQpArgs<TIO>& a, unsigned grid, int threads, cudaStream_t st) {
  using Ipm = BoxQpIpm<double, TIO, NX, NU, 0, 2, ST, true>;
  auto kern = boxqp_ipm_staged_kernel<TIO, ST, NX, NU, MINB>;
  int smem = threads / 32 * Ipm::kDepth * Ipm::kBufBytes;
  // MPC_QP_PAD_SMEM=<bytes>: unused shared memory per CTA, an occupancy experiment (fewer resident CTAs, same code):
  // fewer resident warps = a smaller streamed working set against the 126 MB L2
  int pad = 0;
  if (const char* env = getenv("MPC_QP_PAD_SMEM")) pad = atoi(env);
  if (pad < 0 || pad > 180 * 1024) pad = 0;
  smem += pad;
  // function attributes are per device: one flag per device of the process (not thread-safe beyond a repeated,
  // idempotent attribute call)
  static int configured[64];
  static bool init_done = false;
  if (!init_done) {
    for (int& c : configured) c = -1;
    init_done = true;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || configured[dev] != pad) {
    const int want = kQpThreads / 32 * Ipm::kDepth * Ipm::kBufBytes + pad;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
    if (e != cudaSuccess)
      return fail((int)e, "mpc_boxqp_solve: %d bytes of shared memory per CTA for the staged kernel: %s", want, cudaGetErrorString(e));
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (dev >= 0 && dev < 64) configured[dev] = pad;
  }
  kern<<<grid, threads, smem, st>>>(a);
  return check_launch("boxqp_ipm_staged_kernel");
}

// staged sweeps (shared LTI model, box constraints only): on unless MPC_QP_STAGED=0
static bool staged_on() {
  const char* env = getenv("MPC_QP_STAGED");
  return !(env && atoi(env) == 0);
}

// Morton (Z-order) key of the initial state: 8 bits per coordinate, coordinates scaled by the batch's own range
// (lohi = [min_0..min_{n-1}, max_0..max_{n-1}] on the device).  Sorting the scenarios by this key puts neighbouring
// initial states -- similar active sets, similar interior-point iteration counts -- into the same warp.
template <typename TIO>
__global__ void __launch_bounds__(256) state_order_key_kernel(const TIO* x0, const TIO* lohi, int n, int64_t batch,
                                                              int32_t* keys) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  unsigned key = 0;
  const int bits = n <= 4 ? 8 : 2;
  for (int j = 0; j < n && j < 16; ++j) {
    const double lo = (double)lohi[j], hi = (double)lohi[n + j];
    const double t = hi > lo ? ((double)x0[(int64_t)j * batch + b] - lo) / (hi - lo) : 0.0;
    int q = (int)(t * (double)((1 << bits) - 1) + 0.5);
    q = q < 0 ? 0 : (q > (1 << bits) - 1 ? (1 << bits) - 1 : q);
    for (int k = 0; k < bits; ++k) key |= (unsigned)((q >> k) & 1) << (k * n + j);
  }
  keys[b] = (int32_t)(key & 0x7fffffffu);
}

// Persistent variant with LANE REFILL.  Interior-point iteration counts differ between scenarios (cfg 3: mean 10.8,
// max 22), and a warp of the kernel above runs until its slowest lane has converged: 21 of 32 lanes were active on
// average (ncu, round 1), and because a 32-byte sector is moved whole, the idle lanes' share of every workspace row
// still crossed HBM.  Here the grid is the resident set, every thread owns ONE workspace lane and pulls scenarios from
// a global queue: when a lane's scenario has converged it is parked, and as soon as `refill_min` lanes of the warp
// are parked (or the queue is empty) they write their outputs and start their next scenarios together -- the
// divergent output/init section then serves several lanes at once.  The workspace is one lane per resident thread,
// independent of the batch.
template <typename TIO, class ST, int NX, int NU, int MINB>
__global__ void __launch_bounds__(kQpThreads, MINB) boxqp_ipm_refill_kernel(BoxQpArgs<TIO> a, unsigned long long* queue,
                                                                            int refill_min) {
  using SH = BoxQpShared<NX, NU>;
  __shared__ double sh[SH::total];
  for (int i = threadIdx.x; i < SH::total; i += blockDim.x) sh[i] = boxqp_shared_elem<double, TIO, NX, NU>(a, i);
  __syncthreads();
  const int64_t lane = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool has_lane = lane < a.ws_lanes;
  BoxQpIpm<double, TIO, NX, NU, 0, 0, ST> ipm(a, sh, 0, has_lane ? lane : 0, a.ws_lanes);
  auto fetch = [&]() -> int64_t {
    if (!has_lane) return -1;
    const unsigned long long t = atomicAdd(queue, 1ULL);
    return t < (unsigned long long)a.batch ? (int64_t)t : -1;
  };
  int64_t scn = fetch();
  if (scn >= 0) ipm.begin(scn);
  bool parked = false;      // converged, output not written yet
  bool drained = false;     // this warp has seen the end of the queue
  while (true) {
    const bool running = scn >= 0 && !parked;
    if (running) parked = ipm.iterate();
    const unsigned m_parked = __ballot_sync(0xffffffffu, parked);
    const unsigned m_idle = __ballot_sync(0xffffffffu, scn < 0);
    const int n_parked = __popc(m_parked), n_idle = __popc(m_idle);
    if (n_parked == 0) {
      if (n_idle == 32) break;
      continue;
    }
    // refill when enough lanes wait (parked or out of work), when nothing else runs in this warp, or at the tail
    if (n_parked + n_idle >= refill_min || drained || n_parked + n_idle == 32) {
      if (parked) {
        ipm.finish();
        parked = false;
        scn = fetch();
        if (scn >= 0) ipm.begin(scn);
      }
      drained = drained || __any_sync(0xffffffffu, scn < 0 && has_lane);
    }
  }
}

// which storage policy the float64 product uses: "mix" (default; slacks, multipliers and dz_aff in float32) or
// "f64" (everything float64, for A/B comparisons): env MPC_QP_STORE
static bool store_all_f64() {
  const char* env = getenv("MPC_QP_STORE");
  return env && strcmp(env, "f64") == 0;
}

// lanes of the refill kernel: the resident threads (occupancy query), capped by the batch
template <typename K>
static int64_t resident_threads(K kern) {
  int dev = 0, sms = kNumSMs, occ = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kQpThreads, 0) != cudaSuccess || occ < 1) occ = 1;
  return (int64_t)sms * occ * kQpThreads;
}

template <typename TIO, class ST, int NX, int NU, int MINB>
static int launch_refill(BoxQpArgs<TIO> a, int refill_min, cudaStream_t st) {
  auto kern = boxqp_ipm_refill_kernel<TIO, ST, NX, NU, MINB>;
  int64_t lanes = resident_threads(kern);
  if (lanes > a.batch) lanes = a.batch;
  a.ws_lanes = lanes;
  // the queue counter lives in the header of the caller's workspace
  unsigned long long* queue = static_cast<unsigned long long*>(a.ws);
  a.ws = static_cast<char*>(a.ws) + kWsHeader;
  cudaError_t e = cudaMemsetAsync(queue, 0, 16, st);  // (the header is kWsHeader bytes; the counter uses the first 8)
  if (e != cudaSuccess) return fail((int)e, "mpc_boxqp_solve: %s", cudaGetErrorString(e));
  const unsigned grid = (unsigned)((lanes + kQpThreads - 1) / kQpThreads);
  kern<<<grid, kQpThreads, 0, st>>>(a, queue, refill_min);
  return check_launch("boxqp_ipm_refill_kernel");
}

// threads per CTA of the thread-per-scenario kernels (env MPC_QP_THREADS: 32, 64 or 128).  A CTA holds its registers
// until its slowest warp has converged; smaller CTAs return them earlier (same resident warps per SM).
static int qp_threads(int dflt) {
  if (const char* env = getenv("MPC_QP_THREADS")) {
    const int t = atoi(env);
    if (t == 32 || t == 64 || t == 128) return t;
  }
  return dflt;
}

template <typename TIO, class ST, int NX, int NU, int NC>
static int launch_boxqp_st(const BoxQpArgs<TIO>& a_in, cudaStream_t st) {
  const int kQpThreadsRt = qp_threads(kQpThreads);
  const unsigned grid = (unsigned)((a_in.batch + kQpThreadsRt - 1) / kQpThreadsRt);
  BoxQpArgs<TIO> a = a_in;
  a.ws_lanes = a.batch;
  // lane refill (persistent kernel): lanes of a warp waiting before they restart together; 0 (default) = one thread
  // per scenario, no refill.  Measured on B200 at cfg 3 (tools/prof/r2_gpu6.sh, 2^18 scenarios): 17.7 ms without,
  // 18.1-20.2 ms with refill thresholds 16 / 8 / 4 -- the divergent output + start section and the per-lane iteration
  // state cost more than the recovered lanes bring, so it stays an opt-in experiment.
  int refill = 0;
  if (const char* env = getenv("MPC_QP_REFILL")) refill = atoi(env);
  if constexpr (NC > 0) {
    a.ws = static_cast<char*>(a.ws) + kWsHeader;
    if (a.ltv) boxqp_ipm_kernel<TIO, ST, NX, NU, NC, 2, true><<<grid, kQpThreadsRt, 0, st>>>(a);
    else boxqp_ipm_kernel<TIO, ST, NX, NU, NC, 2, false><<<grid, kQpThreadsRt, 0, st>>>(a);
    return check_launch("boxqp_ipm_rows_kernel");
  } else if constexpr (NX + NU <= 3) {
    // (2,1): residency against registers, measured on B200 (tools/prof/exp_q3.sh); default 4 CTAs/SM
    int minb = 4;
    if (const char* env = getenv("MPC_QP_MINB")) minb = atoi(env);
    if (!getenv("MPC_QP_PREFETCH")) a.pf_dist = 2;  // kernels without staging (LTV, MPC_QP_STAGED=0): bulk L2 prefetch 2 visits ahead
    if (refill > 0 && a.batch > 4096) {
      if (minb >= 4) return launch_refill<TIO, ST, NX, NU, 4>(a, refill, st);
      return launch_refill<TIO, ST, NX, NU, 3>(a, refill, st);
    }
    a.ws = static_cast<char*>(a.ws) + kWsHeader;
    if (!a.ltv && staged_on()) {
      if (minb >= 6) return launch_staged<TIO, ST, NX, NU, 6>(a, grid, kQpThreadsRt, st);
      if (minb == 5) return launch_staged<TIO, ST, NX, NU, 5>(a, grid, kQpThreadsRt, st);
      return launch_staged<TIO, ST, NX, NU, 4>(a, grid, kQpThreadsRt, st);
    }
    if (a.ltv) boxqp_ipm_kernel<TIO, ST, NX, NU, 0, 4, true><<<grid, kQpThreadsRt, 0, st>>>(a);
    else if (minb >= 6) boxqp_ipm_kernel<TIO, ST, NX, NU, 0, 6, false><<<grid, kQpThreadsRt, 0, st>>>(a);
    else if (minb == 5) boxqp_ipm_kernel<TIO, ST, NX, NU, 0, 5, false><<<grid, kQpThreadsRt, 0, st>>>(a);
    else if (minb == 4) boxqp_ipm_kernel<TIO, ST, NX, NU, 0, 4, false><<<grid, kQpThreadsRt, 0, st>>>(a);
    else boxqp_ipm_kernel<TIO, ST, NX, NU, 0, 3, false><<<grid, kQpThreadsRt, 0, st>>>(a);
  } else {
    if (refill > 0 && a.batch > 4096) return launch_refill<TIO, ST, NX, NU, 2>(a, refill, st);
    a.ws = static_cast<char*>(a.ws) + kWsHeader;
    if (!a.ltv && staged_on()) return launch_staged<TIO, ST, NX, NU, 2>(a, grid, kQpThreadsRt, st);
    if (a.ltv) boxqp_ipm_kernel<TIO, ST, NX, NU, 0, 2, true><<<grid, kQpThreadsRt, 0, st>>>(a);
    else boxqp_ipm_kernel<TIO, ST, NX, NU, 0, 2, false><<<grid, kQpThreadsRt, 0, st>>>(a);
  }
  return check_launch("boxqp_ipm_kernel");
}

template <int NX, int NU, int NC>
static int launch_boxqp(const BoxQpArgs<double>& a, cudaStream_t st) {
  if (store_all_f64()) return launch_boxqp_st<double, StoreF64, NX, NU, NC>(a, st);
  return launch_boxqp_st<double, StoreMix, NX, NU, NC>(a, st);
}

}  // namespace mpc

using namespace mpc;

extern "C" int64_t mpc_boxqp_rows_workspace_bytes(int64_t batch, int n, int m, int N, int nc, int dtype) {
  if (batch < 0 || n < 1 || m < 1 || N < 1 || nc < 0) return 0;
  // kWsHeader bytes for the scenario queue of the refill kernel + one lane per scenario (the refill kernel uses one per
  // resident thread only).  MPC_F64: the size of the all-float64 layout (the default mixed layout is smaller;
  // MPC_QP_STORE=f64 needs this)
  return kWsHeader + (dtype == MPC_F32 ? boxqp_ws_bytes<StoreF32>(n, m, N, nc, batch) : boxqp_ws_bytes<StoreF64>(n, m, N, nc, batch));
}

extern "C" int64_t mpc_boxqp_workspace_bytes(int64_t batch, int n, int m, int N, int dtype) {
  if (batch < 0 || n < 1 || m < 1 || N < 1) return 0;
  // (12,4) runs on the persistent warp-per-scenario kernel: one slot per resident warp, not per scenario
  if (coop_supported(n, m, 0)) return coop_ws_elems(n, m, N, batch) * 8;
  return mpc_boxqp_rows_workspace_bytes(batch, n, m, N, 0, dtype);
}

template <typename TIO>
static BoxQpArgs<TIO> make_args(const void* A, const void* B, const void* c, int ltv, const void* Q, const void* R,
                                const void* Pf, const void* u_lo, const void* u_hi, const void* x_lo, const void* x_hi,
                                const void* x0, const void* warm_U, void* U, void* X, void* cost, int32_t* status,
                                int32_t* iters, int8_t* sat_u, int8_t* sat_x, const void* Cg, const void* hg,
                                int8_t* sat_c, void* ws, int64_t batch, int N, int max_iter, double eps) {
  BoxQpArgs<TIO> a{(const TIO*)A, (const TIO*)B, (const TIO*)c, ltv ? 1 : 0, (const TIO*)Q, (const TIO*)R, (const TIO*)Pf,
                   (const TIO*)u_lo, (const TIO*)u_hi, (const TIO*)x_lo, (const TIO*)x_hi, (const TIO*)x0,
                   (const TIO*)warm_U, (TIO*)U, (TIO*)X, (TIO*)cost, status, iters, sat_u, sat_x, (const TIO*)Cg,
                   (const TIO*)hg, sat_c, ws, batch, N, max_iter, eps};
  return a;
}

static int boxqp_solve_impl(const void* A, const void* B, const void* c, int ltv, const void* Q,
                               const void* R, const void* Pf, const void* u_lo, const void* u_hi,
                               const void* x_lo, const void* x_hi, const void* x0, const void* warm_U,
                               void* U, void* X, void* cost, int32_t* status, int32_t* iters,
                               int8_t* sat_u, int8_t* sat_x, const void* Cg, const void* hg, int nc, int8_t* sat_c,
                               void* ws, int64_t ws_bytes, int64_t batch, int n, int m, int N, int max_iter,
                               double eps, int dtype, mpc_stream_t stream, const int32_t* order = nullptr) {
  MPC_REQUIRE(dtype == MPC_F64 || dtype == MPC_F32, MPC_ERR_DTYPE, "mpc_boxqp_solve: unknown dtype %d", dtype);
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
  const size_t es = dtype == MPC_F32 ? 4 : 8;
  for (const void* p : {A, B, c, Q, R, Pf, u_lo, u_hi, x_lo, x_hi, x0, warm_U, (const void*)U, (const void*)X,
                        (const void*)cost, Cg, hg})
    MPC_REQUIRE(!p || aligned(p, es), MPC_ERR_ALIGN, "mpc_boxqp_solve: misaligned pointer");
  MPC_REQUIRE(aligned(ws, 16), MPC_ERR_ALIGN, "mpc_boxqp_solve: workspace must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  int pf_dist = 0;  // (4,x) stages: the kernels are bandwidth-bound on their workspace, see launch_boxqp_st and rti.cu
  if (const char* env = getenv("MPC_QP_PREFETCH")) pf_dist = atoi(env);
  if (dtype == MPC_F32) {
    // float32 product: caller's arrays and the whole workspace in float32, arithmetic in float64
    BoxQpArgs<float> a = make_args<float>(A, B, c, ltv, Q, R, Pf, u_lo, u_hi, x_lo, x_hi, x0, warm_U, U, X, cost, status,
                                          iters, sat_u, sat_x, Cg, hg, sat_c, ws, batch, N, max_iter, eps);
    a.pf_dist = pf_dist;
    a.order = order;
    if (nc > 0) {
      if (n == 4 && m == 2 && nc == 9) return launch_boxqp_st<float, StoreF32, 4, 2, 9>(a, st);
      return fail(MPC_ERR_UNSUPPORTED, "mpc_boxqp_solve_rows: no float32 kernel instantiated for n=%d m=%d nc=%d", n, m, nc);
    }
    if (n == 2 && m == 1) return launch_boxqp_st<float, StoreF32, 2, 1, 0>(a, st);
    if (n == 4 && m == 1) return launch_boxqp_st<float, StoreF32, 4, 1, 0>(a, st);
    if (n == 4 && m == 2) return launch_boxqp_st<float, StoreF32, 4, 2, 0>(a, st);
    return fail(MPC_ERR_UNSUPPORTED, "mpc_boxqp_solve: no float32 kernel instantiated for n=%d m=%d", n, m);
  }
  BoxQpArgs<double> a = make_args<double>(A, B, c, ltv, Q, R, Pf, u_lo, u_hi, x_lo, x_hi, x0, warm_U, U, X, cost, status,
                                          iters, sat_u, sat_x, Cg, hg, sat_c, ws, batch, N, max_iter, eps);
  a.pf_dist = pf_dist;
  a.order = order;
  MPC_REQUIRE(!order || !coop_supported(n, m, ltv) || n <= 4, MPC_ERR_UNSUPPORTED,
              "mpc_boxqp_solve_ordered: the warp-per-scenario kernel takes no order (its lanes share one scenario)");
  if (nc > 0) {
    if (n == 4 && m == 2 && nc == 9) return launch_boxqp<4, 2, 9>(a, st);
    if (n == 4 && m == 2 && nc == 3) return launch_boxqp<4, 2, 3>(a, st);
    return fail(MPC_ERR_UNSUPPORTED, "mpc_boxqp_solve_rows: no kernel instantiated for n=%d m=%d nc=%d", n, m, nc);
  }
  if (n == 2 && m == 1) return launch_boxqp<2, 1, 0>(a, st);
  if (n == 4 && m == 1) return launch_boxqp<4, 1, 0>(a, st);
  if (n == 4 && m == 2) return launch_boxqp<4, 2, 0>(a, st);
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

extern "C" int mpc_state_order_keys(const void* x0, const void* lohi, int32_t* keys, int64_t batch, int n, int dtype,
                                    mpc_stream_t stream) {
  MPC_REQUIRE(dtype == MPC_F64 || dtype == MPC_F32, MPC_ERR_DTYPE, "mpc_state_order_keys: unknown dtype %d", dtype);
  if (batch == 0) return MPC_OK;
  MPC_REQUIRE(x0 && lohi && keys, MPC_ERR_NULL, "mpc_state_order_keys: null pointer");
  MPC_REQUIRE(n >= 1 && n <= 16 && batch >= 0, MPC_ERR_SHAPE, "mpc_state_order_keys: bad (n=%d)", n);
  const unsigned grid = (unsigned)((batch + 255) / 256);
  if (dtype == MPC_F32)
    state_order_key_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x0, (const float*)lohi, n, batch, keys);
  else
    state_order_key_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>((const double*)x0, (const double*)lohi, n, batch, keys);
  return check_launch("state_order_key_kernel");
}

extern "C" int mpc_boxqp_solve_ordered(const void* A, const void* B, const void* c, int ltv, const void* Q, const void* R,
                                       const void* Pf, const void* u_lo, const void* u_hi, const void* x_lo,
                                       const void* x_hi, const void* x0, const void* warm_U, void* U, void* X, void* cost,
                                       int32_t* status, int32_t* iters, int8_t* sat_u, int8_t* sat_x, const int32_t* order,
                                       void* ws, int64_t ws_bytes, int64_t batch, int n, int m, int N, int max_iter,
                                       double eps, int dtype, mpc_stream_t stream) {
  return boxqp_solve_impl(A, B, c, ltv, Q, R, Pf, u_lo, u_hi, x_lo, x_hi, x0, warm_U, U, X, cost, status, iters, sat_u,
                          sat_x, nullptr, nullptr, 0, nullptr, ws, ws_bytes, batch, n, m, N, max_iter, eps, dtype, stream,
                          order);
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
