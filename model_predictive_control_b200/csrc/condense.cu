// K3: condensed prediction matrices of the LTI box-QP, and K6: summary reduction.
//
// K3 builds, for x_{k+1} = A x_k + B u_k and X = (x_1..x_N), U = (u_0..u_{N-1}):
//     X = Phi x0 + Gamma U,   Phi = [A; A^2; ...; A^N],   Gamma_{ij} = A^{i-j} B (i >= j)
//     H = Gamma' Qbar Gamma + Rbar,   F = Gamma' Qbar Phi,   Qbar = blkdiag(Q, .., Q, Pf)
// so that the MPC cost of the reference's Problem data (session_2/problem.py:8-24) is
// J(U) = U'HU + 2 x0'F'U + const ("condensed form", BASELINE.json configs[2]).
// One CTA per model; the A-power blocks G_d = A^d B, Phi_d = A^(d+1) and their Q-weighted copies
// are staged in shared memory, every output entry is then a short dot product over staged blocks.
#include "common.cuh"

namespace mpc {

template <typename T>
struct CondenseArgs {
  const T *A, *B, *Q, *R, *Pf;
  int64_t sA, sB, sQ, sR, sPf;
  T *Phi, *Gamma, *H, *F;  // any may be null
  int n, m, N;
};

template <typename T>
__global__ void __launch_bounds__(256) condense_kernel(CondenseArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = a.n, m = a.m, N = a.N;
  T* G = reinterpret_cast<T*>(smem_raw);  // [N][n][m]   A^d B
  T* P = G + N * n * m;                   // [N][n][n]   A^(d+1)
  T* QG = P + N * n * n;                  // [N][n][m]   Q A^d B
  T* TG = QG + N * n * m;                 // [N][n][m]   Pf A^d B
  T* sA = TG + N * n * m;                 // [n][n]
  T* sQ = sA + n * n;
  T* sPf = sQ + n * n;
  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  const T* Bm = a.B + b * a.sB;
  for (int i = tid; i < n * n; i += nt) {
    sA[i] = a.A[b * a.sA + i];
    sQ[i] = a.Q[b * a.sQ + i];
    sPf[i] = a.Pf[b * a.sPf + i];
    P[i] = sA[i];
  }
  for (int i = tid; i < n * m; i += nt) G[i] = Bm[i];
  __syncthreads();
  for (int d = 1; d < N; ++d) {  // G_d = A G_{d-1},  Phi_d = A Phi_{d-1}
    for (int e = tid; e < n * m; e += nt) {
      const int r = e / m, c = e % m;
      T acc = T(0);
      for (int l = 0; l < n; ++l) acc = fma_<T>(sA[r * n + l], G[(d - 1) * n * m + l * m + c], acc);
      G[d * n * m + e] = acc;
    }
    for (int e = tid; e < n * n; e += nt) {
      const int r = e / n, c = e % n;
      T acc = T(0);
      for (int l = 0; l < n; ++l) acc = fma_<T>(sA[r * n + l], P[(d - 1) * n * n + l * n + c], acc);
      P[d * n * n + e] = acc;
    }
    __syncthreads();
  }
  for (int e = tid; e < N * n * m; e += nt) {
    const int d = e / (n * m), r = (e / m) % n, c = e % m;
    T q = T(0), t = T(0);
    for (int l = 0; l < n; ++l) {
      const T g = G[d * n * m + l * m + c];
      q = fma_<T>(sQ[r * n + l], g, q);
      t = fma_<T>(sPf[r * n + l], g, t);
    }
    QG[e] = q;
    TG[e] = t;
  }
  __syncthreads();
  const int nz = N * m, nc = N * n;
  if (a.Phi) {
    T* out = a.Phi + b * (int64_t)nc * n;
    for (int e = tid; e < nc * n; e += nt) out[e] = P[e];
  }
  if (a.Gamma) {
    T* out = a.Gamma + b * (int64_t)nc * nz;
    for (int e = tid; e < nc * nz; e += nt) {
      const int row = e / nz, col = e % nz;
      const int i = row / n, r = row % n, j = col / m, c = col % m;
      out[e] = i >= j ? G[(i - j) * n * m + r * m + c] : T(0);
    }
  }
  if (a.H) {
    T* out = a.H + b * (int64_t)nz * nz;
    const T* Rm = a.R + b * a.sR;
    for (int e = tid; e < nz * nz; e += nt) {
      const int row = e / nz, col = e % nz;
      const int j = row / m, c = row % m, j2 = col / m, c2 = col % m;
      const int i0 = j > j2 ? j : j2;
      T acc = (j == j2) ? Rm[c * m + c2] : T(0);
      for (int i = i0; i < N; ++i) {
        const T* g = G + (i - j) * n * m;
        const T* w = (i == N - 1 ? TG : QG) + (i - j2) * n * m;
        for (int l = 0; l < n; ++l) acc = fma_<T>(g[l * m + c], w[l * m + c2], acc);
      }
      out[e] = acc;
    }
  }
  if (a.F) {
    T* out = a.F + b * (int64_t)nz * n;
    for (int e = tid; e < nz * n; e += nt) {
      const int row = e / n, c2 = e % n;
      const int j = row / m, c = row % m;
      T acc = T(0);
      for (int i = j; i < N; ++i) {
        const T* w = (i == N - 1 ? TG : QG) + (i - j) * n * m;  // (Qbar_i G_{i-j})[:, c]
        const T* ph = P + i * n * n;
        for (int l = 0; l < n; ++l) acc = fma_<T>(w[l * m + c], ph[l * n + c2], acc);
      }
      out[e] = acc;
    }
  }
}

// ------------------------------------------------------------------------------------------ K6
// out[0] = scenarios, out[1] = sum cost, out[2] = max violation, out[3] = sum saturated,
// out[4] = # MPC_INFEASIBLE, out[5] = # MPC_MAX_ITER, out[6] = sum iterations, out[7] = # MPC_SOLVED
template <typename T>
__global__ void __launch_bounds__(256) summary_kernel(const T* cost, const T* viol, const int32_t* n_sat,
                                                      const int32_t* status, const int32_t* iters, int64_t batch,
                                                      double* out) {
  double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < batch; b += (int64_t)gridDim.x * blockDim.x) {
    v[0] += 1.0;
    if (cost) v[1] += (double)cost[b];
    if (viol) v[2] = fmax(v[2], (double)viol[b]);
    if (n_sat) v[3] += (double)n_sat[b];
    if (status) {
      const int s = status[b];
      v[4] += s == MPC_INFEASIBLE;
      v[5] += s == MPC_MAX_ITER;
      v[7] += s == MPC_SOLVED;
    }
    if (iters) v[6] += (double)iters[b];
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    for (int off = 16; off > 0; off >>= 1) {
      const double o = __shfl_down_sync(0xffffffffu, v[k], off);
      v[k] = (k == 2) ? fmax(v[k], o) : v[k] + o;
    }
  }
  __shared__ double part[8][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0)
    for (int k = 0; k < 8; ++k) part[warp][k] = v[k];
  __syncthreads();
  if (threadIdx.x < 8) {
    const int k = threadIdx.x;
    double r = part[0][k];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = (k == 2) ? fmax(r, part[w][k]) : r + part[w][k];
    if (k == 2) {
      // non-negative doubles order like their bit patterns
      atomicMax(reinterpret_cast<unsigned long long*>(out + 2), (unsigned long long)__double_as_longlong(r));
    } else {
      atomicAdd(out + k, r);
    }
  }
}

}  // namespace mpc

using namespace mpc;

extern "C" int mpc_condense(const void* A, int64_t sA, const void* B, int64_t sB, const void* Q, int64_t sQ,
                            const void* R, int64_t sR, const void* Pf, int64_t sPf, void* Phi, void* Gamma, void* H,
                            void* F, int64_t batch, int n, int m, int N, int dtype, mpc_stream_t stream) {
  MPC_REQUIRE(dtype == MPC_F64 || dtype == MPC_F32, MPC_ERR_DTYPE, "mpc_condense: unknown dtype %d", dtype);
  MPC_REQUIRE(n >= 1 && n <= MPC_MAX_NX && m >= 1 && m <= MPC_MAX_NU && N >= 1 && batch >= 0, MPC_ERR_SHAPE,
              "mpc_condense: bad shape n=%d m=%d N=%d", n, m, N);
  if (batch == 0) return MPC_OK;  // nothing to do; pointers of an empty batch may be null
  MPC_REQUIRE(A && B && Q && R && Pf, MPC_ERR_NULL, "mpc_condense: null model pointer");
  MPC_REQUIRE(sA >= 0 && sB >= 0 && sQ >= 0 && sR >= 0 && sPf >= 0, MPC_ERR_SHAPE, "mpc_condense: negative stride");
  const size_t es = dtype == MPC_F64 ? 8 : 4;
  for (const void* p : {A, B, Q, R, Pf, (const void*)Phi, (const void*)Gamma, (const void*)H, (const void*)F})
    MPC_REQUIRE(!p || aligned(p, es), MPC_ERR_ALIGN, "mpc_condense: misaligned pointer");
  const size_t smem = es * ((size_t)N * n * (3 * m + n) + 3 * (size_t)n * n);
  MPC_REQUIRE(smem <= 220 * 1024, MPC_ERR_SHAPE, "mpc_condense: N*n*(3m+n) = %d does not fit shared memory",
              N * n * (3 * m + n));
  MPC_REQUIRE(batch <= 0x7fffffff, MPC_ERR_SHAPE, "mpc_condense: batch too large");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MPC_F64) {
    CondenseArgs<double> a{(const double*)A, (const double*)B, (const double*)Q, (const double*)R, (const double*)Pf,
                           sA, sB, sQ, sR, sPf, (double*)Phi, (double*)Gamma, (double*)H, (double*)F, n, m, N};
    if (smem > 48 * 1024) cudaFuncSetAttribute(condense_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    condense_kernel<double><<<(unsigned)batch, 256, smem, st>>>(a);
  } else {
    CondenseArgs<float> a{(const float*)A, (const float*)B, (const float*)Q, (const float*)R, (const float*)Pf,
                          sA, sB, sQ, sR, sPf, (float*)Phi, (float*)Gamma, (float*)H, (float*)F, n, m, N};
    if (smem > 48 * 1024) cudaFuncSetAttribute(condense_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    condense_kernel<float><<<(unsigned)batch, 256, smem, st>>>(a);
  }
  return check_launch("condense_kernel");
}

extern "C" int mpc_summary(const void* cost, const void* viol, const int32_t* n_sat, const int32_t* status,
                           const int32_t* iters, int64_t batch, double* out8, int dtype, mpc_stream_t stream) {
  MPC_REQUIRE(dtype == MPC_F64 || dtype == MPC_F32, MPC_ERR_DTYPE, "mpc_summary: unknown dtype %d", dtype);
  if (batch == 0) return MPC_OK;  // nothing to do; pointers of an empty batch may be null
  MPC_REQUIRE(out8, MPC_ERR_NULL, "mpc_summary: null output");
  MPC_REQUIRE(batch >= 0, MPC_ERR_SHAPE, "mpc_summary: negative batch");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(out8, 0, 8 * sizeof(double), st);
  if (e != cudaSuccess) return fail((int)e, "mpc_summary: %s", cudaGetErrorString(e));
  int64_t blocks = (batch + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  if (dtype == MPC_F64)
    summary_kernel<double><<<(unsigned)blocks, 256, 0, st>>>((const double*)cost, (const double*)viol, n_sat, status, iters, batch, out8);
  else
    summary_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((const float*)cost, (const float*)viol, n_sat, status, iters, batch, out8);
  return check_launch("summary_kernel");
}
