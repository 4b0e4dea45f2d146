// LQ path kernels and their C ABI: mpc_riccati (K1), mpc_lq_rollout (K2), mpc_lq_solve (K1+K2).
// sm_100a only.  See include/mpc_b200.h for the contract and the reference lines each entry
// point replaces.
#include <stdlib.h>

#include <type_traits>

#include "lq_core.cuh"

namespace mpc {

constexpr int kLqThreads = 128;

// ================================================================== K1, register-resident
template <typename T, int NX, int NU, bool AL>
__global__ void __launch_bounds__(kLqThreads) riccati_reg_kernel(RiccatiArgs<T> a) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b < a.batch) riccati_body<T, NX, NU, AL>(a, b);
}

// ================================================================== K1, generic (runtime n, m)
// One CTA per scenario, matrices in shared memory, threads over matrix entries.  Used for shapes
// without a register-resident instantiation (e.g. n = 12, m = 4); with a shared model this is a
// single CTA, so it is a latency-only kernel.
template <typename T>
__global__ void __launch_bounds__(128) riccati_generic_kernel(RiccatiArgs<T> a, int n, int m) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);
  T* A = sm;
  T* B = A + n * n;
  T* Q = B + n * m;
  T* R = Q + n * n;
  T* P = R + m * m;
  T* W = P + n * n;    // P A, later P A + P B K
  T* PB = W + n * n;   // n x m
  T* S = PB + n * m;   // m x m
  T* G = S + m * m;    // m x n  -> K
  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < n * n; i += nt) {
    A[i] = a.A[b * a.sA + i];
    Q[i] = a.Q[b * a.sQ + i];
    P[i] = a.Pf[b * a.sPf + i];
  }
  for (int i = tid; i < n * m; i += nt) B[i] = a.B[b * a.sB + i];
  for (int i = tid; i < m * m; i += nt) R[i] = a.R[b * a.sR + i];
  __syncthreads();
  const int64_t pstage = a.batch * (int64_t)(n * n), kstage = a.batch * (int64_t)(m * n);
  if (a.P && a.all_P)
    for (int i = tid; i < n * n; i += nt) a.P[a.N * pstage + b * (n * n) + i] = P[i];
  for (int k = a.N - 1; k >= 0; --k) {
    // W = P A ; PB = P B
    for (int e = tid; e < n * n; e += nt) {
      const int i = e / n, j = e % n;
      T acc = T(0);
      for (int l = 0; l < n; ++l) acc = fma_<T>(P[i * n + l], A[l * n + j], acc);
      W[e] = acc;
    }
    for (int e = tid; e < n * m; e += nt) {
      const int i = e / m, j = e % m;
      T acc = T(0);
      for (int l = 0; l < n; ++l) acc = fma_<T>(P[i * n + l], B[l * m + j], acc);
      PB[e] = acc;
    }
    __syncthreads();
    // G = B' W ; S = R + B' PB
    for (int e = tid; e < m * n; e += nt) {
      const int i = e / n, j = e % n;
      T acc = T(0);
      for (int l = 0; l < n; ++l) acc = fma_<T>(B[l * m + i], W[l * n + j], acc);
      G[e] = acc;
    }
    for (int e = tid; e < m * m; e += nt) {
      const int i = e / m, j = e % m;
      T acc = R[e];
      for (int l = 0; l < n; ++l) acc = fma_<T>(B[l * m + i], PB[l * m + j], acc);
      S[e] = acc;
    }
    __syncthreads();
    // Gauss-Jordan on [S | G] (S symmetric positive definite: no pivoting), then K = -G
    for (int p = 0; p < m; ++p) {
      const T inv = T(1) / S[p * m + p];
      __syncthreads();
      for (int j = tid; j < m + n; j += nt) {
        if (j < m) S[p * m + j] *= inv;
        else G[p * n + (j - m)] *= inv;
      }
      __syncthreads();
      for (int e = tid; e < m * (m + n); e += nt) {
        const int r = e / (m + n), j = e % (m + n);
        if (r == p) continue;
        const T f = S[r * m + p];
        if (j < m) {
          if (j != p) S[r * m + j] = fma_<T>(-f, S[p * m + j], S[r * m + j]);
        } else {
          G[r * n + (j - m)] = fma_<T>(-f, G[p * n + (j - m)], G[r * n + (j - m)]);
        }
      }
      __syncthreads();
      for (int r = tid; r < m; r += nt)
        if (r != p) S[r * m + p] = T(0);
      __syncthreads();
    }
    for (int e = tid; e < m * n; e += nt) {
      G[e] = -G[e];
      a.K[k * kstage + b * (m * n) + e] = G[e];
    }
    __syncthreads();
    // W += PB K
    for (int e = tid; e < n * n; e += nt) {
      const int i = e / n, j = e % n;
      T acc = W[e];
      for (int l = 0; l < m; ++l) acc = fma_<T>(PB[i * m + l], G[l * n + j], acc);
      W[e] = acc;
    }
    __syncthreads();
    // P = Q + A' W
    for (int e = tid; e < n * n; e += nt) {
      const int i = e / n, j = e % n;
      T acc = Q[e];
      for (int l = 0; l < n; ++l) acc = fma_<T>(A[l * n + i], W[l * n + j], acc);
      P[e] = acc;
      if (a.P && a.all_P) a.P[k * pstage + b * (n * n) + e] = acc;
    }
    __syncthreads();
  }
  if (a.P && !a.all_P)
    for (int i = tid; i < n * n; i += nt) a.P[b * (n * n) + i] = P[i];
}

// ================================================================== K2
template <typename T, int NX, int NU, int VEC>
__global__ void __launch_bounds__(256) rollout_shared_kernel(RolloutArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);
  using L = RolloutSmem<T, NX, NU>;
  for (int i = threadIdx.x; i < NX * NX; i += blockDim.x) {
    sm[L::oA + i] = a.A[i];
    sm[L::oQ + i] = a.Q ? a.Q[i] : T(0);
    sm[L::oPf + i] = a.Pf ? a.Pf[i] : T(0);
  }
  for (int i = threadIdx.x; i < NX * NU; i += blockDim.x) sm[L::oB + i] = a.B[i];
  for (int i = threadIdx.x; i < NU * NU; i += blockDim.x) sm[L::oR + i] = a.R ? a.R[i] : T(0);
  for (int i = threadIdx.x; i < a.ng * NU * NX; i += blockDim.x) {
    const int g = i / (NU * NX), e = i % (NU * NX);
    sm[L::oK + i] = a.K[(int64_t)g * a.sK_stage + e];
  }
  __syncthreads();
  const int64_t b0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (b0 < a.batch) rollout_shared_body<T, NX, NU, VEC>(a, sm, b0);
}

template <typename T, int NX, int NU>
__global__ void __launch_bounds__(128) rollout_perscn_kernel(RolloutArgs<T> a) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b < a.batch) rollout_perscn_body<T, NX, NU>(a, b);
}

// Generic runtime-(n, m) rollout: one thread per scenario, state in local memory.
template <typename T>
__global__ void __launch_bounds__(128) rollout_generic_kernel(RolloutArgs<T> a, int n, int m) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.batch) return;
  T x[MPC_MAX_NX], xn[MPC_MAX_NX], u[MPC_MAX_NU];
  const T* A = a.A + b * a.sA;
  const T* B = a.B + b * a.sB;
  for (int i = 0; i < n; ++i) {
    x[i] = a.x0[i * a.batch + b];
    a.X[i * a.batch + b] = x[i];
  }
  T c = T(0);
  bool bad = false;
  for (int t = 0; t + 1 < a.Tn; ++t) {
    const T* Kt = a.K + (int64_t)(a.gain_offset + a.gain_step * t) * a.sK_stage + b * a.sK;
    for (int j = 0; j < m; ++j) {
      T acc = T(0);
      for (int i = 0; i < n; ++i) acc = fma_<T>(Kt[j * n + i], x[i], acc);
      u[j] = acc;
      if (a.U) a.U[((int64_t)t * m + j) * a.batch + b] = acc;
    }
    if (a.cost) {
      for (int i = 0; i < n; ++i) {
        T r = T(0);
        for (int j = 0; j < n; ++j) r = fma_<T>(a.Q[i * n + j], x[j], r);
        c = fma_<T>(x[i], r, c);
      }
      for (int i = 0; i < m; ++i) {
        T r = T(0);
        for (int j = 0; j < m; ++j) r = fma_<T>(a.R[i * m + j], u[j], r);
        c = fma_<T>(u[i], r, c);
      }
    }
    T n2 = T(0);
    for (int i = 0; i < n; ++i) {
      T acc = T(0);
      for (int l = 0; l < n; ++l) acc = fma_<T>(A[i * n + l], x[l], acc);
      for (int j = 0; j < m; ++j) acc = fma_<T>(B[i * m + j], u[j], acc);
      xn[i] = acc;
      n2 = fma_<T>(acc, acc, n2);
    }
    for (int i = 0; i < n; ++i) {
      x[i] = xn[i];
      a.X[((int64_t)(t + 1) * n + i) * a.batch + b] = x[i];
    }
    bad = bad || (n2 > a.norm_limit2);
  }
  if (a.cost) {
    for (int i = 0; i < n; ++i) {
      T r = T(0);
      for (int j = 0; j < n; ++j) r = fma_<T>(a.Pf[i * n + j], x[j], r);
      c = fma_<T>(x[i], r, c);
    }
    a.cost[b] = c;
  }
  if (a.unstable) a.unstable[b] = bad ? 1 : 0;
}

// One plant step x+ = A x + B u for an externally supplied input (LinearSystem.f, reference
// session_1/LinearSystem.py:16-18), batch-contiguous x [n][batch], u [m][batch]; shared A, B.
template <typename T>
__global__ void __launch_bounds__(256) linear_step_kernel(const T* __restrict__ A, const T* __restrict__ B,
                                                          const T* __restrict__ x, const T* __restrict__ u,
                                                          T* __restrict__ xn, int64_t batch, int n, int m) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  for (int i = 0; i < n; ++i) {
    T acc = T(0);
    for (int l = 0; l < n; ++l) acc = fma_<T>(__ldg(A + i * n + l), __ldg(x + l * batch + b), acc);
    for (int j = 0; j < m; ++j) acc = fma_<T>(__ldg(B + i * m + j), __ldg(u + j * batch + b), acc);
    xn[i * batch + b] = acc;
  }
}

// ================================================================== K1 + K2 fused
// L2 prefetch of `count` elements starting at p (16-byte aligned start and size): one instruction, no
// registers or shared memory held while the data is in flight.
template <typename T>
__device__ __forceinline__ void prefetch_l2(const T* p, int64_t count) {
  if (count > 0)
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"((unsigned)(count * sizeof(T))) : "memory");
}

// Thread 0 of a CTA pulls the per-scenario inputs of the CTA `dist` positions later in the grid into L2 (see
// launch_lq_solve for the distance): that CTA's first loads are then L2 hits instead of ~1 us DRAM round trips with
// nothing else to do.  Needs 32-byte aligned, densely packed inputs (the AL kernels).
template <typename T, int NX, int NU>
__device__ __forceinline__ void prefetch_next_inputs(const LqSolveArgs<T>& a, int dist) {
  if (dist <= 0 || threadIdx.x != 0) return;
  const int64_t b0 = ((int64_t)blockIdx.x + dist) * blockDim.x;
  const int64_t cnt = min((int64_t)blockDim.x, a.batch - b0) & ~(int64_t)3;  // multiples of 16 bytes for every array
  if (cnt <= 0) return;
  if (a.sA) prefetch_l2(a.A + b0 * a.sA, cnt * (NX * NX));
  if (a.sQ) prefetch_l2(a.Q + b0 * a.sQ, cnt * (NX * NX));
  if (a.sPf) prefetch_l2(a.Pf + b0 * a.sPf, cnt * (NX * NX));
  if (a.sB) prefetch_l2(a.B + b0 * a.sB, cnt * (NX * NU));
  if (a.sR) prefetch_l2(a.R + b0 * a.sR, cnt * (NU * NU));
  prefetch_l2(a.x0 + b0 * NX, cnt * NX);
}

template <typename T, int NX, int NU, bool AL>
__global__ void __launch_bounds__(kLqThreads) lq_solve_kernel(LqSolveArgs<T> a, int prefetch_dist) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* Ks = reinterpret_cast<T*>(smem_raw) + threadIdx.x;
  if (AL) prefetch_next_inputs<T, NX, NU>(a, prefetch_dist);
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b < a.batch) lq_solve_body<T, NX, NU, AL>(a, b, Ks, blockDim.x);
}

// Single-input models, fp64: the solve in Krylov coordinates (lq_solve_krylov_body); scenarios whose
// controllability matrix is too ill-conditioned for the 1e-6 parity bar take the dense body instead.
template <int NX, bool AL>
__global__ void __launch_bounds__(kLqThreads) lq_solve_krylov_kernel(LqSolveArgs<double> a, double cond2_max,
                                                                     int prefetch_dist) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Ks = reinterpret_cast<double*>(smem_raw) + threadIdx.x;
  if (AL) prefetch_next_inputs<double, NX, 1>(a, prefetch_dist);
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.batch) return;
  if (!lq_solve_krylov_body<NX, AL>(a, b, Ks, blockDim.x, cond2_max))
    lq_solve_body<double, NX, 1, AL>(a, b, Ks, blockDim.x);
}

// ================================================================== host dispatch
template <typename T>
static bool rows_aligned32(std::initializer_list<const void*> ptrs) {
  for (const void* p : ptrs)
    if (p && !aligned(p, 32)) return false;
  return true;
}

static bool stride_ok(int64_t s, int cnt) { return s == 0 || s == cnt; }

template <typename T, int NX, int NU>
static int launch_riccati_reg(const RiccatiArgs<T>& a, cudaStream_t st) {
  const bool al = rows_aligned32<T>({a.A, a.B, a.Q, a.R, a.Pf, a.K, a.P}) &&
                  stride_ok(a.sA, NX * NX) && stride_ok(a.sB, NX * NU) && stride_ok(a.sQ, NX * NX) &&
                  stride_ok(a.sR, NU * NU) && stride_ok(a.sPf, NX * NX);
  const unsigned grid = (unsigned)((a.batch + kLqThreads - 1) / kLqThreads);
  if (al)
    riccati_reg_kernel<T, NX, NU, true><<<grid, kLqThreads, 0, st>>>(a);
  else
    riccati_reg_kernel<T, NX, NU, false><<<grid, kLqThreads, 0, st>>>(a);
  return check_launch("riccati_reg_kernel");
}

template <typename T>
static int riccati_dispatch(RiccatiArgs<T> a, int n, int m, cudaStream_t st) {
  if (n == 2 && m == 1) return launch_riccati_reg<T, 2, 1>(a, st);
  if (n == 4 && m == 1) return launch_riccati_reg<T, 4, 1>(a, st);
  if (n == 4 && m == 2) return launch_riccati_reg<T, 4, 2>(a, st);
  const size_t smem = sizeof(T) * (size_t)(4 * n * n + 3 * n * m + 2 * m * m);
  MPC_REQUIRE(a.batch <= 0x7fffffff, MPC_ERR_SHAPE, "mpc_riccati: generic kernel batch too large");
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(riccati_generic_kernel<T>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "mpc_riccati: %s", cudaGetErrorString(e));
  }
  riccati_generic_kernel<T><<<(unsigned)a.batch, 128, smem, st>>>(a, n, m);
  return check_launch("riccati_generic_kernel");
}

template <typename T, int NX, int NU>
static int launch_rollout(const RolloutArgs<T>& a, cudaStream_t st) {
  const bool shared = a.sA == 0 && a.sB == 0 && a.sK == 0;
  if (shared) {
    const size_t smem = sizeof(T) * (size_t)RolloutSmem<T, NX, NU>::total(a.ng);
    MPC_REQUIRE(smem <= 200 * 1024, MPC_ERR_SHAPE, "mpc_lq_rollout: %d gain stages do not fit shared memory", a.ng);
    constexpr int VEC = 16 / (int)sizeof(T);
    const bool vec = (a.batch % VEC == 0) && aligned(a.x0, 16) && aligned(a.X, 16) &&
                     (!a.U || aligned(a.U, 16)) && (!a.cost || aligned(a.cost, 16));
    const int threads = 256;
    if (vec) {
      auto kern = rollout_shared_kernel<T, NX, NU, VEC>;
      if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      const int64_t nthr = a.batch / VEC;
      kern<<<(unsigned)((nthr + threads - 1) / threads), threads, smem, st>>>(a);
    } else {
      auto kern = rollout_shared_kernel<T, NX, NU, 1>;
      if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      kern<<<(unsigned)((a.batch + threads - 1) / threads), threads, smem, st>>>(a);
    }
    return check_launch("rollout_shared_kernel");
  }
  rollout_perscn_kernel<T, NX, NU><<<(unsigned)((a.batch + 127) / 128), 128, 0, st>>>(a);
  return check_launch("rollout_perscn_kernel");
}

template <typename T>
static int rollout_dispatch(RolloutArgs<T> a, int n, int m, cudaStream_t st) {
  if (n == 2 && m == 1) return launch_rollout<T, 2, 1>(a, st);
  if (n == 4 && m == 1) return launch_rollout<T, 4, 1>(a, st);
  if (n == 4 && m == 2) return launch_rollout<T, 4, 2>(a, st);
  rollout_generic_kernel<T><<<(unsigned)((a.batch + 127) / 128), 128, 0, st>>>(a, n, m);
  return check_launch("rollout_generic_kernel");
}

// cond_F(C) <= 1e3 keeps the Krylov-coordinate solve within ~1e-7 of the dense recursion (measured:
// <= 1.5e-10 on the cfg-2b distribution, <= 2e-8 on 10x wider model spreads).
// MPC_LQ_KRYLOV_COND overrides the bound; 0 disables the path.
static double krylov_cond_max() {
  if (const char* env = getenv("MPC_LQ_KRYLOV_COND")) return atof(env);
  return 1e3;
}

extern "C" int mpc_lq_solve_variant(int n, int m, int dtype, int wants_gains_or_P0) {
  const bool shape = (m == 1) && (n == 2 || n == 4);
  return (shape && dtype == MPC_F64 && !wants_gains_or_P0 && krylov_cond_max() > 0) ? 1 : 0;
}

template <typename T, int NX, int NU>
static int launch_lq_solve(const LqSolveArgs<T>& a, cudaStream_t st) {
  const bool al = rows_aligned32<T>({a.A, a.B, a.Q, a.R, a.Pf, a.x0, a.X, a.U, a.K, a.P0}) &&
                  stride_ok(a.sA, NX * NX) && stride_ok(a.sB, NX * NU) && stride_ok(a.sQ, NX * NX) &&
                  stride_ok(a.sR, NU * NU) && stride_ok(a.sPf, NX * NX);
  int threads = kLqThreads;
  size_t per_thread = sizeof(T) * (size_t)a.N * NU * NX;
  // the on-chip gains bound the resident warps: keep each CTA's slice <= ~44 KB so that >= 5 CTAs
  // share the 227 KB of an SM
  while (threads > 32 && per_thread * threads > 44 * 1024) threads >>= 1;
  if (const char* env = getenv("MPC_LQ_THREADS")) {
    const int t = atoi(env);
    if (t >= 32 && t <= kLqThreads && t % 32 == 0) threads = t;
  }
  const size_t smem = per_thread * threads;
  MPC_REQUIRE(smem <= 220 * 1024, MPC_ERR_SHAPE, "mpc_lq_solve: horizon %d too long for on-chip gains", a.N);
  const unsigned grid = (unsigned)((a.batch + threads - 1) / threads);
  // L2 prefetch distance in CTAs: 48 scenarios per SM ahead of the CTA that issues it, i.e. about
  // 1/7 of the scenarios in flight (5 CTAs x 64 per SM).  Measured on B200 (tools/prof/
  // exp_lq_variants.py, cfg 2b, sustained clocks): 0.288 ms without prefetch, 0.2505 ms for 74..148 CTAs
  // of 64, 0.265 ms at 370, no gain at the full resident set (740): lines prefetched too early are
  // evicted by the write stream before they are used.
  int pf = (48 * kNumSMs) / threads;
  if (const char* env = getenv("MPC_LQ_PREFETCH")) pf = atoi(env);
  if constexpr (std::is_same<T, double>::value && NU == 1 && (NX == 2 || NX == 4)) {
    const double cond_max = krylov_cond_max();
    if (cond_max > 0 && !a.K && !a.P0) {
      const double c2 = cond_max * cond_max;
      if (al) {
        auto kern = lq_solve_krylov_kernel<NX, true>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<grid, threads, smem, st>>>(a, c2, pf);
      } else {
        auto kern = lq_solve_krylov_kernel<NX, false>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<grid, threads, smem, st>>>(a, c2, 0);
      }
      return check_launch("lq_solve_krylov_kernel");
    }
  }
  if (al) {
    auto kern = lq_solve_kernel<T, NX, NU, true>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, threads, smem, st>>>(a, pf);
  } else {
    auto kern = lq_solve_kernel<T, NX, NU, false>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, threads, smem, st>>>(a, 0);
  }
  return check_launch("lq_solve_kernel");
}

template <typename T>
static int lq_solve_dispatch(LqSolveArgs<T> a, int n, int m, cudaStream_t st) {
  if (n == 2 && m == 1) return launch_lq_solve<T, 2, 1>(a, st);
  if (n == 4 && m == 1) return launch_lq_solve<T, 4, 1>(a, st);
  if (n == 4 && m == 2) return launch_lq_solve<T, 4, 2>(a, st);
  return fail(MPC_ERR_UNSUPPORTED, "mpc_lq_solve: no register-resident kernel for n=%d m=%d (use mpc_riccati + mpc_lq_rollout)", n, m);
}

}  // namespace mpc

using namespace mpc;

#define CHECK_COMMON(fn)                                                                          \
  MPC_REQUIRE(dtype == MPC_F64 || dtype == MPC_F32, MPC_ERR_DTYPE, fn ": unknown dtype %d", dtype); \
  MPC_REQUIRE(n >= 1 && n <= MPC_MAX_NX && m >= 1 && m <= MPC_MAX_NU, MPC_ERR_SHAPE,              \
              fn ": (n=%d, m=%d) outside 1..%d x 1..%d", n, m, MPC_MAX_NX, MPC_MAX_NU);           \
  MPC_REQUIRE(batch >= 0, MPC_ERR_SHAPE, fn ": negative batch");

template <typename T>
static bool elem_aligned(std::initializer_list<const void*> ptrs) {
  for (const void* p : ptrs)
    if (p && !aligned(p, sizeof(T))) return false;
  return true;
}

extern "C" int mpc_riccati(const void* A, int64_t sA, const void* B, int64_t sB, const void* Q,
                           int64_t sQ, const void* R, int64_t sR, const void* Pf, int64_t sPf,
                           void* K, void* P, int all_P, int64_t batch, int n, int m, int N,
                           int dtype, mpc_stream_t stream) {
  CHECK_COMMON("mpc_riccati");
  MPC_REQUIRE(N >= 0, MPC_ERR_SHAPE, "mpc_riccati: negative horizon");
  if (batch == 0) return MPC_OK;  // nothing to do; pointers of an empty batch may be null
  MPC_REQUIRE(A && B && Q && R && Pf, MPC_ERR_NULL, "mpc_riccati: null model pointer");
  MPC_REQUIRE(K || N == 0, MPC_ERR_NULL, "mpc_riccati: null K");
  MPC_REQUIRE(sA >= 0 && sB >= 0 && sQ >= 0 && sR >= 0 && sPf >= 0, MPC_ERR_SHAPE, "mpc_riccati: negative stride");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MPC_F64) {
    MPC_REQUIRE(elem_aligned<double>({A, B, Q, R, Pf, K, P}), MPC_ERR_ALIGN, "mpc_riccati: misaligned pointer");
    RiccatiArgs<double> a{(const double*)A, (const double*)B, (const double*)Q, (const double*)R, (const double*)Pf,
                          sA, sB, sQ, sR, sPf, (double*)K, (double*)P, all_P, batch, N};
    return riccati_dispatch<double>(a, n, m, st);
  }
  MPC_REQUIRE(elem_aligned<float>({A, B, Q, R, Pf, K, P}), MPC_ERR_ALIGN, "mpc_riccati: misaligned pointer");
  RiccatiArgs<float> a{(const float*)A, (const float*)B, (const float*)Q, (const float*)R, (const float*)Pf,
                       sA, sB, sQ, sR, sPf, (float*)K, (float*)P, all_P, batch, N};
  return riccati_dispatch<float>(a, n, m, st);
}

extern "C" int mpc_lq_rollout(const void* A, int64_t sA, const void* B, int64_t sB, const void* K,
                              int64_t sK_stage, int64_t sK, int gain_offset, int gain_step,
                              const void* x0, void* X, void* U, const void* Q, const void* R,
                              const void* Pf, void* cost, uint8_t* unstable, double norm_limit,
                              int64_t batch, int n, int m, int T, int dtype, mpc_stream_t stream) {
  CHECK_COMMON("mpc_lq_rollout");
  MPC_REQUIRE(T >= 1, MPC_ERR_SHAPE, "mpc_lq_rollout: need at least one state (T=%d)", T);
  if (batch == 0) return MPC_OK;  // nothing to do; pointers of an empty batch may be null
  MPC_REQUIRE(A && B && x0 && X, MPC_ERR_NULL, "mpc_lq_rollout: null pointer");
  MPC_REQUIRE(K || T == 1, MPC_ERR_NULL, "mpc_lq_rollout: null gains");
  MPC_REQUIRE(!cost || (Q && R && Pf), MPC_ERR_NULL, "mpc_lq_rollout: cost needs Q, R, Pf");
  MPC_REQUIRE(gain_offset >= 0 && gain_step >= 0 && sA >= 0 && sB >= 0 && sK >= 0 && sK_stage >= 0,
              MPC_ERR_SHAPE, "mpc_lq_rollout: negative stride / gain index");
  const int ng = (T >= 2) ? gain_offset + gain_step * (T - 2) + 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MPC_F64) {
    MPC_REQUIRE(elem_aligned<double>({A, B, K, x0, X, U, Q, R, Pf, cost}), MPC_ERR_ALIGN, "mpc_lq_rollout: misaligned pointer");
    RolloutArgs<double> a{(const double*)A, (const double*)B, sA, sB, (const double*)K, sK_stage, sK, ng,
                          gain_offset, gain_step, (const double*)x0, (double*)X, (double*)U,
                          (const double*)Q, (const double*)R, (const double*)Pf, (double*)cost, unstable,
                          norm_limit * norm_limit, batch, T};
    return rollout_dispatch<double>(a, n, m, st);
  }
  MPC_REQUIRE(elem_aligned<float>({A, B, K, x0, X, U, Q, R, Pf, cost}), MPC_ERR_ALIGN, "mpc_lq_rollout: misaligned pointer");
  RolloutArgs<float> a{(const float*)A, (const float*)B, sA, sB, (const float*)K, sK_stage, sK, ng,
                       gain_offset, gain_step, (const float*)x0, (float*)X, (float*)U,
                       (const float*)Q, (const float*)R, (const float*)Pf, (float*)cost, unstable,
                       (float)(norm_limit * norm_limit), batch, T};
  return rollout_dispatch<float>(a, n, m, st);
}

extern "C" int mpc_lq_solve(const void* A, int64_t sA, const void* B, int64_t sB, const void* Q,
                            int64_t sQ, const void* R, int64_t sR, const void* Pf, int64_t sPf,
                            const void* x0, void* X, void* U, void* V, void* K, void* P0,
                            int64_t batch, int n, int m, int N, int dtype, mpc_stream_t stream) {
  CHECK_COMMON("mpc_lq_solve");
  MPC_REQUIRE(N >= 1, MPC_ERR_SHAPE, "mpc_lq_solve: horizon must be >= 1");
  if (batch == 0) return MPC_OK;  // nothing to do; pointers of an empty batch may be null
  MPC_REQUIRE(A && B && Q && R && Pf && x0 && X && U && V, MPC_ERR_NULL, "mpc_lq_solve: null pointer");
  MPC_REQUIRE(sA >= 0 && sB >= 0 && sQ >= 0 && sR >= 0 && sPf >= 0, MPC_ERR_SHAPE, "mpc_lq_solve: negative stride");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MPC_F64) {
    MPC_REQUIRE(elem_aligned<double>({A, B, Q, R, Pf, x0, X, U, V, K, P0}), MPC_ERR_ALIGN, "mpc_lq_solve: misaligned pointer");
    LqSolveArgs<double> a{(const double*)A, (const double*)B, (const double*)Q, (const double*)R, (const double*)Pf,
                          sA, sB, sQ, sR, sPf, (const double*)x0, (double*)X, (double*)U, (double*)V,
                          (double*)K, (double*)P0, batch, N};
    return lq_solve_dispatch<double>(a, n, m, st);
  }
  MPC_REQUIRE(elem_aligned<float>({A, B, Q, R, Pf, x0, X, U, V, K, P0}), MPC_ERR_ALIGN, "mpc_lq_solve: misaligned pointer");
  LqSolveArgs<float> a{(const float*)A, (const float*)B, (const float*)Q, (const float*)R, (const float*)Pf,
                       sA, sB, sQ, sR, sPf, (const float*)x0, (float*)X, (float*)U, (float*)V,
                       (float*)K, (float*)P0, batch, N};
  return lq_solve_dispatch<float>(a, n, m, st);
}

extern "C" int mpc_linear_step(const void* A, const void* B, const void* x, const void* u, void* xn,
                               int64_t batch, int n, int m, int dtype, mpc_stream_t stream) {
  CHECK_COMMON("mpc_linear_step");
  if (batch == 0) return MPC_OK;  // nothing to do; pointers of an empty batch may be null
  MPC_REQUIRE(A && B && x && u && xn, MPC_ERR_NULL, "mpc_linear_step: null pointer");
  MPC_REQUIRE(xn != x, MPC_ERR_UNSUPPORTED, "mpc_linear_step: in-place step not supported");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((batch + 255) / 256);
  if (dtype == MPC_F64) {
    MPC_REQUIRE(elem_aligned<double>({A, B, x, u, xn}), MPC_ERR_ALIGN, "mpc_linear_step: misaligned pointer");
    linear_step_kernel<double><<<grid, 256, 0, st>>>((const double*)A, (const double*)B, (const double*)x,
                                                     (const double*)u, (double*)xn, batch, n, m);
  } else {
    MPC_REQUIRE(elem_aligned<float>({A, B, x, u, xn}), MPC_ERR_ALIGN, "mpc_linear_step: misaligned pointer");
    linear_step_kernel<float><<<grid, 256, 0, st>>>((const float*)A, (const float*)B, (const float*)x,
                                                    (const float*)u, (float*)xn, batch, n, m);
  }
  return check_launch("linear_step_kernel");
}
