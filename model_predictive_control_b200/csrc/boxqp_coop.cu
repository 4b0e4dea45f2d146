// K4, cooperative variant: one WARP per scenario for state dimensions whose Riccati matrices do not
// fit one thread's registers (n = 12, m = 4: BASELINE config 5).  Same algorithm, same passes and
// the same arithmetic order per element as BoxQpIpm (csrc/boxqp_core.cuh); what changes is the
// mapping:
//   * the n x n / n x m / m x m matrices of the factorisation live in the warp's slice of shared
//     memory and every product is computed lane-parallel over its output entries;
//   * lane i < n + m owns element i of the stage vectors (z, slacks, multipliers, directions);
//   * the kernel is persistent: grid = a fixed number of warps, each warp strides over scenarios and
//     reuses ONE workspace slot, so the workspace does not grow with the batch (8M scenarios of
//     72 KB state each would not fit any HBM) and no lane ever waits for another scenario.
// Shared LTI model only (A, B staged once per CTA).
#include <stdlib.h>

#include "boxqp_core.cuh"

namespace mpc {

constexpr int kCoopWarps = 8;            // warps per CTA
constexpr int kCoopMaxSlots = 148 * 32;  // upper bound of resident warps (workspace slots)

template <int NX, int NU>
struct CoopLayout {
  static constexpr int D = NX + NU;
  // per CTA
  static constexpr int oA = 0;
  static constexpr int oB = oA + NX * NX;
  static constexpr int oQ = oB + NX * NU;
  static constexpr int oR = oQ + NX * NX;
  static constexpr int oPf = oR + NU * NU;
  static constexpr int oLo = oPf + NX * NX;
  static constexpr int oHi = oLo + D;
  static constexpr int shared_total = oHi + D;
  // per warp
  static constexpr int wP = 0;
  static constexpr int wW = wP + NX * NX;
  static constexpr int wAcl = wW + NX * NX;  // closed-loop matrix A + B K
  static constexpr int wPB = wAcl + NX * NX;
  static constexpr int wAug = wPB + NX * NU;  // [NU][2 NU] augmented matrix of the S inverse
  static constexpr int wSi = wAug + NU * 2 * NU;
  static constexpr int wK = wSi + NU * NU;
  static constexpr int wG = wK + NU * NX;
  static constexpr int wRhs = wG + NU * NX;
  static constexpr int wSig = wRhs + D;
  static constexpr int wZ = wSig + D;
  static constexpr int wH = wZ + D;
  static constexpr int wGu = wH + NX;
  static constexpr int wDff = wGu + NU;
  static constexpr int wPacc = wDff + NU;
  static constexpr int wX = wPacc + NX;
  static constexpr int wXn = wX + NX;
  static constexpr int wU = wXn + NX;
  static constexpr int warp_total = wU + NU;
  // workspace slot (global), elements: STAGE-MAJOR -- one stage of the slot is kStage contiguous doubles with every
  // section at a compile-time offset, so a stage visit addresses [slot + k * kStage + immediate + lane] from ONE pointer
  // (round 1: ten section pointers = 20 of the kernel's 64 registers, a 64-bit multiply-add per access)
  static constexpr int sZ = 0, sSl = D, sSu = 2 * D, sLl = 3 * D, sLu = 4 * D, sDa = 5 * D, sDz = 6 * D;
  static constexpr int sK = 7 * D, sS = sK + NU * NX, sDw = sS + NU * NU;
  static constexpr int kStage = sDw + NU;
  __host__ __device__ static int64_t slot_elems(int N) { return (int64_t)N * kStage; }
};

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// C[M x Nn] (+)= op(A) B, lane-parallel over the output entries; A is [M x K] (TA: stored [K x M]).
// Shared-memory loads, not FMAs, bound this kernel (one wavefront per load, one per clock and SM,
// against two FP64 warp instructions per clock), so each lane computes TWO adjacent outputs of a row
// from one scalar load of A and one 16-byte load of B per k: 1 load per FMA instead of 2.
template <int M, int K, int Nn, bool TA, bool ACC>
__device__ __forceinline__ void wmm(const double* __restrict__ A, const double* __restrict__ B, double* C, int lane) {
  static_assert(Nn % 2 == 0, "pairs of adjacent columns");
  constexpr int H = Nn / 2;
  for (int e = lane; e < M * H; e += 32) {
    const int i = e / H, j = 2 * (e % H);
    double2 acc = ACC ? *reinterpret_cast<const double2*>(C + i * Nn + j) : make_double2(0.0, 0.0);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const double av = TA ? A[k * M + i] : A[i * K + k];
      const double2 bv = *reinterpret_cast<const double2*>(B + k * Nn + j);
      acc.x = fma(av, bv.x, acc.x);
      acc.y = fma(av, bv.y, acc.y);
    }
    *reinterpret_cast<double2*>(C + i * Nn + j) = acc;
  }
}

// D[M x Nn] = Cin + op(A) B on the FP64 TENSOR pipe: warp-wide DMMA m8n8k4 (mma.sync, f64), operands in shared memory,
// zero-padded to the 8 x 8 x 4 tile.  Element (i, k) of op(A) is A[i * sai + k * sak], element (k, j) of B is
// B[k * sbk + j * sbj]; Cin (row-major, ldc; nullptr = 0) may alias D: a lane reads and writes the same two entries.
// Fragment layout (PTX ISA, mma.m8n8k4 f64): A: row = lane / 4, col = lane % 4;  B: row = lane % 4, col = lane / 4;
// C / D: row = lane / 4, cols = 2 (lane % 4) + {0, 1}.
// Per 8 x 8 x 4 tile step a lane issues 2 shared loads and 1 DMMA (512 FMAs warp-wide), where the scalar form needs
// 16 DFMAs and ~16 loads: the 12 x 12 products of the factorisation are no longer what the kernel waits for.
template <int M, int K, int Nn>
__device__ __forceinline__ void dmma_mm(const double* __restrict__ A, int sai, int sak, const double* __restrict__ B, int sbk,
                                        int sbj, const double* Cin, double* D, int ldc, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < (M + 7) / 8; ++mt) {
#pragma unroll
    for (int nt = 0; nt < (Nn + 7) / 8; ++nt) {
      const int ci = mt * 8 + g, cj = nt * 8 + 2 * t;
      const bool ok0 = ci < M && cj < Nn, ok1 = ci < M && cj + 1 < Nn;
      double c0 = (Cin && ok0) ? Cin[ci * ldc + cj] : 0.0;
      double c1 = (Cin && ok1) ? Cin[ci * ldc + cj + 1] : 0.0;
#pragma unroll
      for (int kt = 0; kt < (K + 3) / 4; ++kt) {
        const int ak = kt * 4 + t;           // A fragment: (row g, col t)
        const int bj = nt * 8 + g;           // B fragment: (row t, col g)
        const double av = (ci < M && ak < K) ? A[ci * sai + ak * sak] : 0.0;
        const double bv = (ak < K && bj < Nn) ? B[ak * sbk + bj * sbj] : 0.0;
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0), "+d"(c1)
                     : "d"(av), "d"(bv));
      }
      if (ok0) D[ci * ldc + cj] = c0;
      if (ok1) D[ci * ldc + cj + 1] = c1;
    }
  }
}

// env MPC_COOP_DMMA=0 selects the scalar (lane-parallel DFMA) products, for A/B measurements
template <int NX, int NU, bool DMMA = true>
struct CoopIpm {
  using L = CoopLayout<NX, NU>;
  static constexpr int D = NX + NU;
  const BoxQpArgs<double>& a;
  const double* sh;  // CTA-shared model block
  double* w;         // this warp's shared-memory slice
  int lane;
  int64_t b, bs;     // scenario, batch (I/O stride)
  double* slot;      // this warp's workspace slot, [N][kStage]
  __device__ __forceinline__ double* st(int k) const { return slot + (int64_t)k * L::kStage; }
  // one bulk L2 prefetch of a whole stage of the slot (contiguous kStage doubles), a.pf_dist stage visits ahead
  __device__ __forceinline__ void pf(int k) const {
    if (a.pf_dist > 0 && k >= 0 && k < a.N && lane == 0)
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(st(k)), "r"(L::kStage * 8) : "memory");
  }
  double mu_scale, mu0;
  bool hl, hu;   // this lane's element has a finite lower / upper bound
  double lo, hi;

  __device__ CoopIpm(const BoxQpArgs<double>& args, const double* shared, double* wsm, double* slot_, int ln)
      : a(args), sh(shared), w(wsm), lane(ln), b(0), bs(args.batch), slot(slot_) {
    mu_scale = 1.0;
    for (int i = 0; i < NX * NX; ++i) mu_scale = fmax(mu_scale, fabs(sh[L::oQ + i]));
    for (int i = 0; i < NU * NU; ++i) mu_scale = fmax(mu_scale, fabs(sh[L::oR + i]));
    mu0 = mu_scale;
    const int i = lane < D ? lane : 0;
    lo = sh[L::oLo + i];
    hi = sh[L::oHi + i];
    hl = lane < D && lo > -kBigBound;
    hu = lane < D && hi < kBigBound;
  }

  // (With the section-major slot of the first half of round 2 an L2 prefetch of the next stage's rows -- one lane per
  // section, `prefetch.global.L2` -- cost more than it brought: 4.23 s instead of 3.53 s per 2^20 cfg-5 solves.  With the
  // stage-major slot it is ONE bulk prefetch of a contiguous 1 440-byte stage: pf() above.)
  // x+ = A x + B u (lanes < NX), from / to the warp's vectors
  __device__ void step_vec(const double* x, const double* u, double* xn) {
    // 2 NX lanes: lane = 2 i + h does half of row i (NX / 2 columns of A, NU / 2 of B), the halves meet by one shuffle
    // (half the dependent FMA chain of the row-per-lane form; every lane of the warp executes the shuffle)
    static_assert(2 * NX <= 32 && NX % 2 == 0 && NU % 2 == 0, "split rows over lane pairs");
    const int i = lane >> 1, h = lane & 1;
    double acc = 0.0;
    if (lane < 2 * NX) {
#pragma unroll
      for (int j = 0; j < NX / 2; ++j) acc = fma(sh[L::oA + i * NX + h * (NX / 2) + j], x[h * (NX / 2) + j], acc);
#pragma unroll
      for (int j = 0; j < NU / 2; ++j) acc = fma(sh[L::oB + i * NU + h * (NU / 2) + j], u[h * (NU / 2) + j], acc);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (lane < 2 * NX && h == 0) xn[i] = acc;
  }

  // (H z)_lane for the stage vector in w[wZ]: R u for inputs, Q x (Pf x at the last stage) for states
  __device__ double hz(int k) const {
    double acc = 0.0;
    if (lane < NU) {
#pragma unroll
      for (int j = 0; j < NU; ++j) acc = fma(sh[L::oR + lane * NU + j], w[L::wZ + j], acc);
    } else if (lane < D) {
      const double* Qx = sh + (k == a.N - 1 ? L::oPf : L::oQ);
      const int i = lane - NU;
#pragma unroll
      for (int j = 0; j < NX; ++j) acc = fma(Qx[i * NX + j], w[L::wZ + NU + j], acc);
    }
    return acc;
  }

  __device__ void init() {
    double g0 = 0.0;
    for (int pass = 0; pass < 2; ++pass) {
      if (lane < NX) w[L::wX + lane] = a.x0[lane * bs + b];
      __syncwarp();
      for (int k = 0; k < a.N; ++k) {
        if (lane < NU) {
          double v = a.warm_U ? a.warm_U[((int64_t)k * NU + lane) * bs + b] : 0.0;
          v = v < lo ? lo : v;
          v = v > hi ? hi : v;
          w[L::wU + lane] = v;
        }
        __syncwarp();
        step_vec(w + L::wX, w + L::wU, w + L::wXn);
        __syncwarp();
        const double zi = lane < NU ? w[L::wU + lane] : (lane < D ? w[L::wXn + lane - NU] : 0.0);
        if (lane < D) w[L::wZ + lane] = zi;
        __syncwarp();
        if (pass == 0) {
          g0 = fmax(g0, fabs(hz(k)));
        } else if (lane < D) {
          double s_l = 1.0, s_u = 1.0, l_l = 0.0, l_u = 0.0;
          if (hl) {
            s_l = fmax(zi - lo, 1.0);
            l_l = mu0 / s_l;
          }
          if (hu) {
            s_u = fmax(hi - zi, 1.0);
            l_u = mu0 / s_u;
          }
          double* const sk = st(k);
          sk[L::sZ + lane] = zi;
          sk[L::sSl + lane] = s_l;
          sk[L::sSu + lane] = s_u;
          sk[L::sLl + lane] = l_l;
          sk[L::sLu + lane] = l_u;
        }
        if (lane < NX) w[L::wX + lane] = w[L::wXn + lane];
        __syncwarp();
      }
      if (pass == 0) {
        g0 = warp_max(g0);
        mu0 = g0 > mu_scale ? g0 : mu_scale;
      }
    }
  }

  static __device__ __forceinline__ double cc_of(double dz_signed, double r, double sig, double l) {
    const double ds = dz_signed + r;
    return ds * (-l - sig * ds);
  }

  template <bool FACTOR>
  __device__ void backward(double sig_mu) {
    for (int e = lane; e < NX * NX; e += 32) w[L::wP + e] = sh[L::oPf + e];  // Pacc
    if (lane < NX) w[L::wPacc + lane] = 0.0;
    __syncwarp();
    for (int k = a.N - 1; k >= 0; --k) {
      pf(k - a.pf_dist);
      double* const sk = st(k);
      double zi = 0.0, s_l = 1.0, s_u = 1.0, l_l = 0.0, l_u = 0.0, da = 0.0;
      if (lane < D) {
        zi = sk[L::sZ + lane];
        s_l = sk[L::sSl + lane];
        s_u = sk[L::sSu + lane];
        l_l = sk[L::sLl + lane];
        l_u = sk[L::sLu + lane];
        if (!FACTOR) da = sk[L::sDa + lane];
        w[L::wZ + lane] = zi;
      }
      if (!FACTOR) {
        for (int e = lane; e < NU * NX; e += 32) w[L::wK + e] = sk[L::sK + e];
        for (int e = lane; e < NU * NU; e += 32) w[L::wSi + e] = sk[L::sS + e];
      }
      __syncwarp();
      double r = -hz(k), sg = 0.0;
      if (hl) {
        const double inv = rcp_(s_l), sgl = l_l * inv, rl = zi - lo - s_l;
        const double cc = FACTOR ? 0.0 : cc_of(da, rl, sgl, l_l);
        sg += sgl;
        r += (sig_mu - cc) * inv - sgl * rl;
      }
      if (hu) {
        const double inv = rcp_(s_u), sgu = l_u * inv, ru = hi - zi - s_u;
        const double cc = FACTOR ? 0.0 : cc_of(-da, ru, sgu, l_u);
        sg += sgu;
        r -= (sig_mu - cc) * inv - sgu * ru;
      }
      if (lane < D) {
        w[L::wRhs + lane] = r;
        w[L::wSig + lane] = sg;
      }
      __syncwarp();
      if (FACTOR) {
        // P = Pacc + diag(Sigma_x)
        if (lane < NX) w[L::wP + lane * NX + lane] += w[L::wSig + NU + lane];
        __syncwarp();
        // Joseph form (see BoxQpIpm::backward): PB = P B, S = Rt + B'PB, K = -S^-1 (PB)'A, Acl = A + B K,
        // Pacc <- Q + Acl' P Acl + K' Rt K with Rt = R + diag(Sigma_u)
        if constexpr (DMMA) dmma_mm<NX, NX, NU>(w + L::wP, NX, 1, sh + L::oB, NU, 1, nullptr, w + L::wPB, NU, lane);
        else wmm<NX, NX, NU, false, false>(w + L::wP, sh + L::oB, w + L::wPB, lane);  // PB = P B
        __syncwarp();
        // augmented [S | I], S = Rt + B'PB ;  G = (PB)'A
        for (int e = lane; e < NU * NU; e += 32) {
          const int i = e / NU, j = e % NU;
          double acc = sh[L::oR + e] + (i == j ? w[L::wSig + i] : 0.0);
#pragma unroll
          for (int l = 0; l < NX; ++l) acc = fma(sh[L::oB + l * NU + i], w[L::wPB + l * NU + j], acc);
          w[L::wAug + i * 2 * NU + j] = acc;
          w[L::wAug + i * 2 * NU + NU + j] = (i == j) ? 1.0 : 0.0;
        }
        if constexpr (DMMA) dmma_mm<NU, NX, NX>(w + L::wPB, 1, NU, sh + L::oA, NX, 1, nullptr, w + L::wG, NX, lane);  // (PB)'A
        else wmm<NU, NX, NX, true, false>(w + L::wPB, sh + L::oA, w + L::wG, lane);
        __syncwarp();
        // Gauss-Jordan without pivoting (S is symmetric positive definite); lanes over [NU][2 NU]
        for (int p = 0; p < NU; ++p) {
          const double inv = rcp_(w[L::wAug + p * 2 * NU + p]);
          double newv = 0.0;
          const int rr = lane / (2 * NU), cc = lane % (2 * NU);
          const bool act = lane < NU * 2 * NU;
          if (act) {
            const double prow = w[L::wAug + p * 2 * NU + cc] * inv;
            newv = (rr == p) ? prow : fma(-w[L::wAug + rr * 2 * NU + p], prow, w[L::wAug + rr * 2 * NU + cc]);
          }
          __syncwarp();
          if (act) w[L::wAug + rr * 2 * NU + cc] = newv;
          __syncwarp();
        }
        for (int e = lane; e < NU * NU; e += 32) w[L::wSi + e] = w[L::wAug + (e / NU) * 2 * NU + NU + e % NU];
        __syncwarp();
        // K = -Sinv G
        for (int e = lane; e < NU * NX; e += 32) {
          const int i = e / NX, j = e % NX;
          double acc = 0.0;
#pragma unroll
          for (int l = 0; l < NU; ++l) acc = fma(w[L::wSi + i * NU + l], w[L::wG + l * NX + j], acc);
          w[L::wK + e] = -acc;
        }
        __syncwarp();
        // Acl = A + B K (into wAcl = the PB.. no: its own buffer wW2), G <- Rt K
        if constexpr (DMMA) {
          dmma_mm<NX, NU, NX>(sh + L::oB, NU, 1, w + L::wK, NX, 1, sh + L::oA, w + L::wAcl, NX, lane);
        } else {
          for (int e = lane; e < NX * NX; e += 32) {
            const int i = e / NX, j = e % NX;
            double acc = sh[L::oA + e];
#pragma unroll
            for (int l = 0; l < NU; ++l) acc = fma(sh[L::oB + i * NU + l], w[L::wK + l * NX + j], acc);
            w[L::wAcl + e] = acc;
          }
        }
        for (int e = lane; e < NU * NX; e += 32) {
          const int i = e / NX, j = e % NX;
          double acc = 0.0;
#pragma unroll
          for (int l = 0; l < NU; ++l) acc = fma(sh[L::oR + i * NU + l] + (i == l ? w[L::wSig + i] : 0.0), w[L::wK + l * NX + j], acc);
          w[L::wG + e] = acc;
        }
        __syncwarp();
        if constexpr (DMMA) dmma_mm<NX, NX, NX>(w + L::wP, NX, 1, w + L::wAcl, NX, 1, nullptr, w + L::wW, NX, lane);
        else wmm<NX, NX, NX, false, false>(w + L::wP, w + L::wAcl, w + L::wW, lane);  // T = P Acl
        __syncwarp();
        if constexpr (DMMA) {
          // Pacc <- Q + Acl'T + K'(Rt K): two accumulating tile products, then the upper triangle mirrored so that P
          // stays exactly symmetric (as the scalar form, which only computes that triangle)
          dmma_mm<NX, NX, NX>(w + L::wAcl, 1, NX, w + L::wW, NX, 1, sh + L::oQ, w + L::wP, NX, lane);
          __syncwarp();
          dmma_mm<NX, NU, NX>(w + L::wK, 1, NX, w + L::wG, NX, 1, w + L::wP, w + L::wP, NX, lane);
          __syncwarp();
          for (int e = lane; e < NX * NX; e += 32) {
            const int i = e / NX, j = e % NX;
            if (i > j) w[L::wP + e] = w[L::wP + j * NX + i];
          }
        } else
        // Pacc <- Q + Acl'T + K'(Rt K) on the upper triangle (NX (NX+1)/2 entries over the lanes), mirrored
        for (int e = lane; e < NX * (NX + 1) / 2; e += 32) {
          int i = 0, rem = e;
          while (rem >= NX - i) {
            rem -= NX - i;
            ++i;
          }
          const int j = i + rem;
          double acc = sh[L::oQ + i * NX + j];
#pragma unroll
          for (int l = 0; l < NX; ++l) acc = fma(w[L::wAcl + l * NX + i], w[L::wW + l * NX + j], acc);
#pragma unroll
          for (int l = 0; l < NU; ++l) acc = fma(w[L::wK + l * NX + i], w[L::wG + l * NX + j], acc);
          w[L::wP + i * NX + j] = acc;
          w[L::wP + j * NX + i] = acc;
        }
        for (int e = lane; e < NU * NX; e += 32) sk[L::sK + e] = w[L::wK + e];
        for (int e = lane; e < NU * NU; e += 32) sk[L::sS + e] = w[L::wSi + e];
        __syncwarp();
      }
      // h = -(rhs_x + pacc);  gu = rhs_u - B'h;  dff = Sinv gu;  pacc <- -A'h + K'gu
      if (lane < NX) w[L::wH + lane] = -(w[L::wRhs + NU + lane] + w[L::wPacc + lane]);
      __syncwarp();
      if (lane < NU) {
        double acc = w[L::wRhs + lane];
#pragma unroll
        for (int i = 0; i < NX; ++i) acc = fma(-sh[L::oB + i * NU + lane], w[L::wH + i], acc);
        w[L::wGu + lane] = acc;
      }
      __syncwarp();
      if (lane < NU) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < NU; ++j) acc = fma(w[L::wSi + lane * NU + j], w[L::wGu + j], acc);
        sk[L::sDw + lane] = acc;
      }
      if (lane < NX) {
        double acc = 0.0;
#pragma unroll
        for (int l = 0; l < NX; ++l) acc = fma(-sh[L::oA + l * NX + lane], w[L::wH + l], acc);
#pragma unroll
        for (int j = 0; j < NU; ++j) acc = fma(w[L::wK + j * NX + lane], w[L::wGu + j], acc);
        w[L::wPacc + lane] = acc;
      }
      __syncwarp();
    }
  }

  struct Acc {
    double qmax, s0, s1, s2, dzmax, rp;
    __device__ double amin() const { return qmax > 0.0 ? 1.0 / qmax : 1e30; }
  };

  template <bool AFFINE>
  __device__ void forward(double sig_mu, Acc& acc) {
    double qmax = 0.0, s0 = 0.0, s1 = 0.0, s2 = 0.0, dzmax = 0.0, rp = 0.0;
    if (lane < NX) w[L::wX + lane] = 0.0;  // dx_0 = 0
    __syncwarp();
    for (int k = 0; k < a.N; ++k) {
      pf(k + a.pf_dist);
      double* const sk = st(k);
      double zi = 0.0, s_l = 1.0, s_u = 1.0, l_l = 0.0, l_u = 0.0, da = 0.0;
      if (lane < D) {
        zi = sk[L::sZ + lane];
        s_l = sk[L::sSl + lane];
        s_u = sk[L::sSu + lane];
        l_l = sk[L::sLl + lane];
        l_u = sk[L::sLu + lane];
        if (!AFFINE) da = sk[L::sDa + lane];
      }
      // u = d + K x on 4 NU lanes: lane = 4 i + q multiplies NX / 4 entries of row i of K, read straight from the slot
      // (no shared-memory copy of the gains in this pass), two shuffles add the quarters
      static_assert(4 * NU <= 32 && NX % 4 == 0, "quarter rows of K over the lanes");
      {
        const int i = lane >> 2, q = lane & 3;
        double kx = 0.0;
        if (lane < 4 * NU) {
          const double* kr = sk + L::sK + i * NX + q * (NX / 4);
#pragma unroll
          for (int j = 0; j < NX / 4; ++j) kx = fma(kr[j], w[L::wX + q * (NX / 4) + j], kx);
          if (q == 0) kx += sk[L::sDw + i];
        }
        kx += __shfl_xor_sync(0xffffffffu, kx, 1);
        kx += __shfl_xor_sync(0xffffffffu, kx, 2);
        if (lane < 4 * NU && q == 0) w[L::wU + i] = kx;
      }
      __syncwarp();
      step_vec(w + L::wX, w + L::wU, w + L::wXn);
      __syncwarp();
      const double dz = lane < NU ? w[L::wU + lane] : (lane < D ? w[L::wXn + lane - NU] : 0.0);
      if (lane < D) {
        sk[(AFFINE ? L::sDa : L::sDz) + lane] = dz;
        if (!AFFINE) dzmax = fmax(dzmax, fabs(dz));
      }
      if (hl) {
        const double r = zi - lo - s_l, ds = dz + r;
        const double rinv = rcp_(s_l * l_l), inv_s = rinv * l_l, inv_l = rinv * s_l, sgl = l_l * inv_s;
        const double cc = AFFINE ? 0.0 : cc_of(da, r, sgl, l_l);
        const double dl = (sig_mu - cc) * inv_s - l_l - sgl * ds;
        qmax = fmax(qmax, fmax(-ds * inv_s, -dl * inv_l));
        s0 += s_l * l_l;
        s1 += s_l * dl + l_l * ds;
        s2 += ds * dl;
        rp = fmax(rp, fabs(r));
      }
      if (hu) {
        const double r = hi - zi - s_u, ds = -dz + r;
        const double rinv = rcp_(s_u * l_u), inv_s = rinv * l_u, inv_l = rinv * s_u, sgu = l_u * inv_s;
        const double cc = AFFINE ? 0.0 : cc_of(-da, r, sgu, l_u);
        const double dl = (sig_mu - cc) * inv_s - l_u - sgu * ds;
        qmax = fmax(qmax, fmax(-ds * inv_s, -dl * inv_l));
        s0 += s_u * l_u;
        s1 += s_u * dl + l_u * ds;
        s2 += ds * dl;
        rp = fmax(rp, fabs(r));
      }
      if (lane < NX) w[L::wX + lane] = w[L::wXn + lane];
      __syncwarp();
    }
    acc.qmax = warp_max(qmax);
    acc.s0 = warp_sum(s0);
    acc.s1 = warp_sum(s1);
    acc.s2 = warp_sum(s2);
    acc.dzmax = warp_max(dzmax);
    acc.rp = warp_max(rp);
  }

  __device__ double update(double sig_mu, double alpha, bool second_order) {
    double zn = 1.0;
    if (lane < D) {
      for (int k = 0; k < a.N; ++k) {
        pf(k + a.pf_dist);
        double* const sk = st(k);
        const double zi = sk[L::sZ + lane], dz = sk[L::sDz + lane], da = second_order ? sk[L::sDa + lane] : 0.0;
        if (hl) {
          const double s = sk[L::sSl + lane], l = sk[L::sLl + lane], r = zi - lo - s, ds = dz + r;
          const double inv = rcp_(s), sgl = l * inv;
          const double cc = second_order ? cc_of(da, r, sgl, l) : 0.0;
          const double dl = (sig_mu - cc) * inv - l - sgl * ds;
          sk[L::sSl + lane] = s + alpha * ds;
          sk[L::sLl + lane] = l + alpha * dl;
        }
        if (hu) {
          const double s = sk[L::sSu + lane], l = sk[L::sLu + lane], r = hi - zi - s, ds = -dz + r;
          const double inv = rcp_(s), sgu = l * inv;
          const double cc = second_order ? cc_of(-da, r, sgu, l) : 0.0;
          const double dl = (sig_mu - cc) * inv - l - sgu * ds;
          sk[L::sSu + lane] = s + alpha * ds;
          sk[L::sLu + lane] = l + alpha * dl;
        }
        const double zn_i = zi + alpha * dz;
        sk[L::sZ + lane] = zn_i;
        zn = fmax(zn, fabs(zn_i));
      }
    }
    return warp_max(zn);
  }

  __device__ void output(int status, int iters) {
    double cost = 0.0;
    if (lane < NX) {
      const double x = a.x0[lane * bs + b];
      w[L::wX + lane] = x;
      a.X[lane * bs + b] = x;
    }
    __syncwarp();
    for (int k = 0; k < a.N; ++k) {
      pf(k + a.pf_dist);
      double* const sk = st(k);
      int sat = 0;
      double zi = 0.0;
      if (lane < D) {
        zi = sk[L::sZ + lane];
        if (hl && sk[L::sLl + lane] > sk[L::sSl + lane]) sat = -1;
        if (hu && sk[L::sLu + lane] > sk[L::sSu + lane]) sat = 1;
      }
      if (lane < NU) {
        const double u = sat < 0 ? lo : (sat > 0 ? hi : zi);
        w[L::wU + lane] = u;
        a.U[((int64_t)k * NU + lane) * bs + b] = u;
        if (a.sat_u) a.sat_u[((int64_t)k * NU + lane) * bs + b] = (int8_t)sat;
      } else if (lane < D && a.sat_x) {
        a.sat_x[((int64_t)k * NX + lane - NU) * bs + b] = (int8_t)sat;
      }
      __syncwarp();
      // stage cost x'Qx + u'Ru: lane i contributes x_i (Qx)_i or u_i (Ru)_i
      if (lane < NX) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < NX; ++j) acc = fma(sh[L::oQ + lane * NX + j], w[L::wX + j], acc);
        cost = fma(w[L::wX + lane], acc, cost);
      }
      if (lane < NU) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < NU; ++j) acc = fma(sh[L::oR + lane * NU + j], w[L::wU + j], acc);
        cost = fma(w[L::wU + lane], acc, cost);
      }
      step_vec(w + L::wX, w + L::wU, w + L::wXn);
      __syncwarp();
      if (lane < NX) {
        const double xn = w[L::wXn + lane];
        w[L::wX + lane] = xn;
        a.X[((int64_t)(k + 1) * NX + lane) * bs + b] = xn;
      }
      __syncwarp();
    }
    if (lane < NX) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < NX; ++j) acc = fma(sh[L::oPf + lane * NX + j], w[L::wX + j], acc);
      cost = fma(w[L::wX + lane], acc, cost);
    }
    cost = warp_sum(cost);
    if (lane == 0) {
      a.cost[b] = cost;
      a.status[b] = status;
      a.iters[b] = iters;
    }
  }

  __device__ void solve(int64_t scenario) {
    b = scenario;
    mu0 = mu_scale;
    int ncons = 0;
    for (int i = 0; i < D; ++i) ncons += (sh[L::oLo + i] > -kBigBound ? 1 : 0) + (sh[L::oHi + i] < kBigBound ? 1 : 0);
    ncons *= a.N;
    init();
    int status = MPC_UNSOLVED, it = 0;
    double rp = 0.0, zn = 1.0;
    Acc acc;
    if (ncons == 0) {
      backward<true>(0.0);
      forward<false>(0.0, acc);
      update(0.0, 1.0, false);
      status = MPC_SOLVED;
      it = 1;
    }
    const double inv_nc = ncons ? 1.0 / (double)ncons : 0.0;
    while (status == MPC_UNSOLVED && it < a.max_iter) {
      ++it;
      backward<true>(0.0);
      forward<true>(0.0, acc);
      const double mu = acc.s0 * inv_nc;
      const double am_aff = acc.amin();
      const double a_aff = am_aff < 1.0 ? am_aff : 1.0;
      const double mu_aff = (acc.s0 + a_aff * (acc.s1 + a_aff * acc.s2)) * inv_nc;
      const double ratio = mu_aff / (mu > 1e-300 ? mu : 1e-300);
      double sigma = ratio * ratio * ratio;
      sigma = sigma < 1.0 ? sigma : 1.0;
      double sig_mu = sigma * mu;
      const double mu_floor = 1e-3 * a.eps * mu_scale;  // see BoxQpIpm::solve
      sig_mu = sig_mu > mu_floor ? sig_mu : mu_floor;
      backward<false>(sig_mu);
      forward<false>(sig_mu, acc);
      double alpha = 0.995 * acc.amin();
      alpha = alpha < 1.0 ? alpha : 1.0;
      zn = update(sig_mu, alpha, true);
      const double mu_new = (acc.s0 + alpha * (acc.s1 + alpha * acc.s2)) * inv_nc;
      rp = (1.0 - alpha) * acc.rp;
      const bool done = (mu_new <= a.eps * mu_scale) && (rp <= a.eps * zn) && (alpha * acc.dzmax <= 1e-6 * zn);
      if (done) {
        status = MPC_SOLVED;
      } else if (!(alpha >= 1e-6) || !(mu_new <= 100.0 * mu0)) {
        status = (rp <= 1e-6 * zn) ? MPC_MAX_ITER : MPC_INFEASIBLE;
      }
    }
    if (status == MPC_UNSOLVED) status = (rp <= 1e-6 * zn) ? MPC_MAX_ITER : MPC_INFEASIBLE;
    output(status, it);
    __syncwarp();
  }
};

template <int NX, int NU, int MINB, bool DMMA>
__global__ void __launch_bounds__(kCoopWarps * 32, MINB) boxqp_ipm_coop_kernel(BoxQpArgs<double> a, int nslots) {
  using L = CoopLayout<NX, NU>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sh = reinterpret_cast<double*>(smem_raw);
  for (int i = threadIdx.x; i < L::shared_total; i += blockDim.x) {
    double v;
    if (i < L::oB) v = a.A[i - L::oA];
    else if (i < L::oQ) v = a.B[i - L::oB];
    else if (i < L::oR) v = a.Q[i - L::oQ];
    else if (i < L::oPf) v = a.R[i - L::oR];
    else if (i < L::oLo) v = a.Pf[i - L::oPf];
    else if (i < L::oLo + NU) v = a.u_lo[i - L::oLo];
    else if (i < L::oHi) v = a.x_lo[i - L::oLo - NU];
    else if (i < L::oHi + NU) v = a.u_hi[i - L::oHi];
    else v = a.x_hi[i - L::oHi - NU];
    sh[i] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = blockIdx.x * kCoopWarps + warp;
  if (slot >= nslots) return;
  double* wsm = sh + L::shared_total + warp * L::warp_total;
  double* ws_slot = static_cast<double*>(a.ws) + (int64_t)slot * L::slot_elems(a.N);
  CoopIpm<NX, NU, DMMA> ipm(a, sh, wsm, ws_slot, lane);
  for (int64_t b = slot; b < a.batch; b += nslots) ipm.solve(b);
}

int64_t coop_ws_elems(int n, int m, int N, int64_t batch) {
  const int64_t slots = batch < kCoopMaxSlots ? batch : kCoopMaxSlots;
  return (int64_t)N * (7 * (n + m) + m * n + m * m + m) * slots;
}

bool coop_supported(int n, int m, int ltv) { return n == 12 && m == 4 && !ltv; }

template <int NX, int NU, int MINB, bool DMMA>
static int launch_coop_variant(const BoxQpArgs<double>& a, cudaStream_t st) {
  using L = CoopLayout<NX, NU>;
  const size_t smem = sizeof(double) * (size_t)(L::shared_total + kCoopWarps * L::warp_total);
  auto kern = boxqp_ipm_coop_kernel<NX, NU, MINB, DMMA>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail((int)e, "mpc_boxqp_solve: %s", cudaGetErrorString(e));
  // persistent grid: exactly the CTAs that are resident at once (one workspace slot per warp)
  int dev = 0, sms = kNumSMs, occ = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kCoopWarps * 32, smem);
  if (e != cudaSuccess || occ < 1) return fail((int)e, "mpc_boxqp_solve: occupancy query failed");
  int64_t slots = (int64_t)sms * occ * kCoopWarps;
  if (slots > kCoopMaxSlots) slots = kCoopMaxSlots;
  if (slots > a.batch) slots = a.batch;
  const unsigned grid = (unsigned)((slots + kCoopWarps - 1) / kCoopWarps);
  kern<<<grid, kCoopWarps * 32, smem, st>>>(a, (int)slots);
  return check_launch("boxqp_ipm_coop_kernel");
}

int launch_boxqp_coop(const BoxQpArgs<double>& a_in, int n, int m, cudaStream_t st) {
  BoxQpArgs<double> a = a_in;
  a.pf_dist = 1;  // stage visits of bulk L2 prefetch ahead of each pass (env MPC_COOP_PREFETCH; 0 / 1 / 2: 1.578 / 1.514 /
                  // 1.518 s per 2^19 cfg-5 solves)
  if (const char* env = getenv("MPC_COOP_PREFETCH")) a.pf_dist = atoi(env);
  if (n == 12 && m == 4) {
    int minb = 4;  // measured on B200: 4 CTAs/SM (64 registers) is marginally the fastest
    if (const char* env = getenv("MPC_COOP_MINB")) minb = atoi(env);
    bool dmma = true;  // products of the factorisation on the FP64 tensor pipe (DMMA); 0 = lane-parallel DFMA
    if (const char* env = getenv("MPC_COOP_DMMA")) dmma = atoi(env) != 0;
    if (!dmma) {
      if (minb >= 4) return launch_coop_variant<12, 4, 4, false>(a, st);
      return launch_coop_variant<12, 4, 2, false>(a, st);
    }
    if (minb >= 4) return launch_coop_variant<12, 4, 4, true>(a, st);
    if (minb == 3) return launch_coop_variant<12, 4, 3, true>(a, st);
    return launch_coop_variant<12, 4, 2, true>(a, st);
  }
  return fail(MPC_ERR_UNSUPPORTED, "mpc_boxqp_solve: no cooperative kernel for n=%d m=%d", n, m);
}

}  // namespace mpc
