// K4: per-scenario body of the box-constrained LQ-MPC QP solver.
//
//   min  sum_{k<N} x_k'Q x_k + u_k'R u_k + x_N'Pf x_N
//   s.t. x_{k+1} = A_k x_k + B_k u_k + c_k,  u_lo <= u_k <= u_hi,  x_lo <= x_{k+1} <= x_hi
//
// (the QP posed by the reference's session_2/problem.py:8-24 and session_3/problem.py:12-28, and --
// with per-scenario stage matrices A_k, B_k, c_k -- the linearised QP of session 4's closed loop,
// session_4/session4_sol.py:158-217).  The reference ships no solver for it.
//
// Method: Mehrotra predictor-corrector interior point.  With z_k = (u_k, x_{k+1}), slacks s and
// multipliers lam for every finite bound, each Newton system (residual form, unknown dz)
//     dz = argmin 1/2 dz'(H + Sigma) dz - rhs'dz   s.t. dx_{k+1} = A_k dx_k + B_k du_k, dx_0 = 0
//     Sigma = lam_l/s_l + lam_u/s_u,   rhs = -H z + (tau - cc_l)/s_l - Sigma_l r_l - (tau - cc_u)/s_u + Sigma_u r_u
// is an unconstrained LQ problem with stage-varying diagonal weight updates, solved by one
// backward Riccati sweep (factorisation + feed-forward) and one forward rollout.  The iterate z always satisfies
// the dynamics exactly (it starts from a rollout and moves along dynamics-consistent directions).  In residual form
// every term of rhs stays O(lam), so the rounding error of the huge barrier weights (Sigma ~ 1e12) scales with |dz|
// and vanishes at the solution.  A first-order (projected-gradient / ADMM) iteration was evaluated in
// oracle/boxqp.py (admm_riccati): 600-2000+ iterations for 1e-6 parity on the session-2 data versus
// ~11 here, so the interior-point iteration is the one that ships.
//
// One thread per scenario.  All per-scenario state lives in a caller-provided workspace laid out in tiles of 32 lanes,
// [tile][stage][section row][lane] (lane-contiguous), so every access of a warp is one coalesced row at a compile-time
// offset from the thread's stage pointer.  The kernel is bound by
// the bytes of that workspace it streams per iteration and by the instructions of its per-bound algebra, so one
// interior-point iteration is organised in FOUR sweeps that touch as little of it as possible (round 1 had five,
// each reading the whole iterate):
//   A  backward  apply the previous iteration's step to the iterate (fused: no separate update pass), factorise
//                (Joseph-form Riccati), feed-forward of the AFFINE right-hand side          -> K, S^-1, d_aff
//   B  forward   affine direction dz_aff by rollout; per bound the affine step ratio and the second-order term cc;
//                stores dz_aff and, per element, e = 1/s_l - 1/s_u and g = cc_l/s_l - cc_u/s_u
//   C  backward  the Newton system is linear in its right-hand side and the corrector's differs from the
//                predictor's by  tau e - g  only: feed-forward of THAT difference with the stored gains.  Reads e, g,
//                K, S^-1 -- not the iterate                                                   -> d_aff + d_cor
//   D  forward   dz by rollout with d_aff + d_cor; per bound ds, dlam, the step ratio and the sums that give the new
//                barrier parameter; stores dz.  The step itself is applied by the next iteration's sweep A (or by
//                the output pass after the last iteration).
// Storage precision is a template policy (BoxQpStore): the float64 product keeps z, the gains and the step dz in
// float64 and the slacks, multipliers, dz_aff and the corrector data e, g in float32 (relative quantities: they enter
// through ratios/products only, every sweep reads the SAME rounded value, and e, g only perturb the direction, so the
// Newton system stays consistent and its fixed point does not move); the float32 product stores everything in float32.
// All arithmetic is float64 in both.
// The body is __host__ __device__: tests/harness runs it on the CPU against oracle/boxqp.py.
#pragma once

#include "smallmat.cuh"

namespace mpc {

constexpr double kBigBound = 1e19;  // |bound| >= kBigBound means "no bound"

// TIO = element type of the caller's arrays (double for MPC_F64, float for MPC_F32)
template <typename TIO>
struct BoxQpArgs {
  const TIO *A, *B, *c;  // ltv == 0: shared A [n][n], B [n][m], c null
                         // ltv == 1: A [N][n*n][batch], B [N][n*m][batch], c [N][n][batch]
  int ltv;
  const TIO *Q, *R, *Pf;                  // shared
  const TIO *u_lo, *u_hi, *x_lo, *x_hi;   // shared [m], [m], [n], [n]
  const TIO* x0;                          // [n][batch]
  const TIO* warm_U;                      // optional [N][m][batch]
  TIO* U;                                 // [N][m][batch]
  TIO* X;                                 // [N+1][n][batch]
  TIO* cost;                              // [batch]
  int32_t* status;                        // [batch]
  int32_t* iters;                         // [batch]
  int8_t* sat_u;                          // optional [N][m][batch]: -1 lower, +1 upper, 0 free
  int8_t* sat_x;                          // optional [N][n][batch]
  // optional general stage rows  Cg_k x_{k+1} >= hg_k  (kernels instantiated with NC > 0 only):
  const TIO* Cg;                          // [N][NC*n][batch]
  const TIO* hg;                          // [N][NC][batch]
  int8_t* sat_c;                          // optional [N][NC][batch]: -1 = row active
  void* ws;                               // workspace, boxqp_ws_bytes<ST>(n, m, N, nc, lanes) bytes
  int64_t batch;
  int N;
  int max_iter;
  double eps;
  const int32_t* order = nullptr;  // optional [batch]: lane b solves scenario order[b] (difficulty-sorted batches)
  int pf_dist = 0;       // stages of L2 prefetch ahead of each sweep's loads (device only; 0 = off)
  int64_t ws_lanes = 0;  // lanes of the workspace (0 = one per scenario: lane b = scenario b)
};

// ---- storage policy: element types of the workspace sections
template <typename TZ, typename TSL, typename TDA, typename TDZ, typename TEG, typename TGN>
struct BoxQpStore {
  using Z = TZ;    // iterate z; carried residuals of general rows
  using SL = TSL;  // slacks and multipliers
  using DA = TDA;  // dz_aff (enters the second-order term only)
  using DZ = TDZ;  // dz (the step the iterate takes: float64 keeps z on the dynamics to float64 accuracy)
  using EG = TEG;  // corrector right-hand-side ingredients e, g
  using GN = TGN;  // gains K, S^-1, feed-forward d
};
using StoreF64 = BoxQpStore<double, double, double, double, double, double>;  // everything float64
using StoreMix = BoxQpStore<double, float, float, double, float, double>;    // float64 product (float32 gains were
// tried: the closed-loop rows A + B K must cancel to 1/Sigma ~ 1e-12 on active states, rounded gains stall hard solves)
using StoreF32 = BoxQpStore<float, float, float, float, float, float>;        // float32 product

inline int64_t ws_round16(int64_t bytes) { return (bytes + 15) / 16 * 16; }

// workspace bytes for `lanes` resident scenarios: tiles of 32 lanes, each [N stages][section rows][32 lanes]
// (BoxQpIpm: kStage bytes per stage and tile)
template <class ST>
inline int64_t boxqp_ws_bytes(int n, int m, int N, int nc, int64_t lanes, int64_t extra_stage_bytes_per_lane = 0) {
  const int64_t d = n + m;
  int64_t per_lane = d * (int64_t)(sizeof(typename ST::Z) + 4 * sizeof(typename ST::SL) + sizeof(typename ST::DA) +
                                   sizeof(typename ST::DZ) + 2 * sizeof(typename ST::EG));
  per_lane += (int64_t)(m * n + m * m + m) * (int64_t)sizeof(typename ST::GN);
  per_lane += (int64_t)nc * (int64_t)(2 * sizeof(typename ST::SL) + sizeof(typename ST::Z));
  per_lane += extra_stage_bytes_per_lane;  // stage model / rows kept in the tile (MODEL = 4, the fused RTI loop)
  const int64_t tiles = (lanes + 31) / 32;
  return tiles * N * per_lane * 32;
}

// shared-parameter block (shared memory on the device), always in the compute type
template <int NX, int NU>
struct BoxQpShared {
  static constexpr int D = NX + NU;
  static constexpr int oA = 0;
  static constexpr int oB = oA + NX * NX;
  static constexpr int oQ = oB + NX * NU;
  static constexpr int oR = oQ + NX * NX;
  static constexpr int oPf = oR + NU * NU;
  static constexpr int oLo = oPf + NX * NX;  // [u_lo | x_lo]
  static constexpr int oHi = oLo + D;        // [u_hi | x_hi]
  static constexpr int total = oHi + D;
};

// fill the shared block from the caller's arrays; element i by the calling thread (host: loop over i)
template <typename T, typename TIO, int NX, int NU>
MPC_HD T boxqp_shared_elem(const BoxQpArgs<TIO>& a, int i) {
  using SH = BoxQpShared<NX, NU>;
  if (i < SH::oB) return a.ltv ? T(0) : T(a.A[i - SH::oA]);
  if (i < SH::oQ) return a.ltv ? T(0) : T(a.B[i - SH::oB]);
  if (i < SH::oR) return T(a.Q[i - SH::oQ]);
  if (i < SH::oPf) return T(a.R[i - SH::oR]);
  if (i < SH::oLo) return T(a.Pf[i - SH::oPf]);
  if (i < SH::oLo + NU) return T(a.u_lo[i - SH::oLo]);
  if (i < SH::oHi) return T(a.x_lo[i - SH::oLo - NU]);
  if (i < SH::oHi + NU) return T(a.u_hi[i - SH::oHi]);
  return T(a.x_hi[i - SH::oHi - NU]);
}

// MODEL = 0: generic dense model (shared LTI, or per-scenario LTV A [N][n*n][batch], B, c), chosen by a.ltv at run time;
// MODEL = 2 / 3: the same with the choice made at compile time (2 = shared LTI, 3 = LTV): the kernels instantiate
//            these, so the body of the other case is not in their instruction stream.
// MODEL = 1: forward-Euler kinematic bicycle (NX = 4, NU = 2), per-scenario LTV in PACKED form: only the 10 entries of
//            A = I + ts J_x and B = ts J_u that are not structurally 0 or 1, plus c: a.A -> [N][14][batch]
//            {a02, a03, a12, a13, a23, a33, b01, b11, b21, b30, c0..c3}.  The structural zeros and ones are written
//            as literals, so the unrolled register algebra drops the corresponding multiplications at compile time.
// MODEL = 4: as 1, with the packed stage model (and the NC general rows) stored IN the workspace tile, written there by
//            the fused RTI loop's preparation: the model loads of every sweep are immediate-offset accesses too.
constexpr int kBicyclePack = 14;

// a [stage][row] array of one scenario: element (k, i) at p[k * sk + i * si]
template <typename S>
struct StageRows {
  S* p;
  int64_t sk, si;
  MPC_HD S& at(int k, int i) const { return p[(int64_t)k * sk + (int64_t)i * si]; }
};
template <typename S>
MPC_HD StageRows<S> batch_rows(S* base, int rows, int64_t bs, int64_t b) {  // caller layout [N][rows][batch]
  return StageRows<S>{base ? base + b : nullptr, (int64_t)rows * bs, bs};
}

template <typename S, typename T>
MPC_HD T round_to(T v) {   // the value a later sweep will read back from a section stored as S
  return (T)(S)v;
}

// STAGED (device only): the warp streams its tile through shared memory -- every sweep's read set is ONE contiguous
// byte range of a stage, copied kDepth stage visits ahead by a bulk asynchronous copy (cp.async.bulk + mbarrier, issued
// by one lane); the sweeps read shared memory, their stores go to global memory directly.
template <typename T, typename TIO, int NX, int NU, int NC = 0, int MODEL = 0, class ST = StoreMix, bool STAGED = false>
struct BoxQpIpm {
  static constexpr int D = NX + NU;
  using SH = BoxQpShared<NX, NU>;
  using TZ = typename ST::Z;
  using TSL = typename ST::SL;
  using TDA = typename ST::DA;
  using TDZ = typename ST::DZ;
  using TEG = typename ST::EG;
  using TGN = typename ST::GN;
  static constexpr bool kNarrowSL = sizeof(TSL) < sizeof(T);
  static constexpr bool kNarrowZ = sizeof(TZ) < sizeof(T);
  static constexpr bool kNarrowDir = sizeof(TDZ) < sizeof(T);

  const BoxQpArgs<TIO>& a;
  const T* sh;
  int64_t b, bs;    // scenario and batch stride of the caller's arrays
  T mu_scale;       // max(1, max|Q|, max|R|): scale of the complementarity tolerance
  T mu0;            // start value of the barrier parameter, per scenario: max(mu_scale, |H z0|_inf)

  // ---- workspace: TILES of kTile = 32 lanes (one warp).  A tile holds [stage][section row][32 lanes]; one stage of a
  // tile is kStage bytes with every section row at a COMPILE-TIME byte offset, so each access of a stage visit is
  // [per-thread base + stage * kStage + immediate]: two base pointers per thread (8- and 4-byte rows) instead of
  // fifteen section pointers and a 64-bit multiply-add per access (a third of the executed instructions with the
  // [stage][row][batch] layout of round 1, whose lane stride is a runtime value).  Every row of a warp is one
  // contiguous 256- or 128-byte segment, as before.
  static constexpr int kTile = 32;
  template <typename S, int OFF>
  struct Sec {
    using type = S;
    static constexpr int off = OFF;
    static constexpr int rowb = kTile * (int)sizeof(S);
  };
  // Section order: the read set of every sweep is one contiguous range of a stage --
  //   C: [e g S K d]   B: [K d z s lam rows model]   D: [K d z s lam rows model dz_aff]   A: [z s lam rows model dz_aff dz]
  using Es = Sec<TEG, 0>;                                // corrector data e, g
  using Gs = Sec<TEG, Es::off + D * Es::rowb>;
  using Ss = Sec<TGN, Gs::off + D * Gs::rowb>;           // gains S^-1, K, feed-forward d
  using Ks = Sec<TGN, Ss::off + NU * NU * Ss::rowb>;
  using Ds = Sec<TGN, Ks::off + NU * NX * Ks::rowb>;
  using Zs = Sec<TZ, Ds::off + NU * Ds::rowb>;           // iterate z
  using SLs = Sec<TSL, Zs::off + D * Zs::rowb>;          // slacks, multipliers
  using SUs = Sec<TSL, SLs::off + D * SLs::rowb>;
  using LLs = Sec<TSL, SUs::off + D * SUs::rowb>;
  using LUs = Sec<TSL, LLs::off + D * LLs::rowb>;
  using SCs = Sec<TSL, LUs::off + D * LUs::rowb>;        // general rows: slack, multiplier
  using LCs = Sec<TSL, SCs::off + NC * SCs::rowb>;
  using RCs = Sec<TZ, LCs::off + NC * LCs::rowb>;        // general rows: residual C x - h - s (carried, see init)
  static constexpr bool kPacked = MODEL == 1 || MODEL == 4;
  static constexpr bool kTileModel = MODEL == 4;
  static constexpr int kMdRows = kTileModel ? kBicyclePack : 0;
  using MDs = Sec<TIO, RCs::off + NC * RCs::rowb>;       // MODEL = 4: packed stage model, general rows C, h
  using CGs = Sec<TIO, MDs::off + kMdRows * MDs::rowb>;
  using HGs = Sec<TIO, CGs::off + (kTileModel ? NC * NX : 0) * CGs::rowb>;
  using DAs = Sec<TDA, HGs::off + (kTileModel ? NC : 0) * HGs::rowb>;  // dz_aff
  using DZs = Sec<TDZ, DAs::off + D * DAs::rowb>;        // dz
  static constexpr int kStage = DZs::off + D * DZs::rowb;
  // read ranges of the four sweeps [lo, hi) and the staging buffer that holds the largest
  static constexpr int kLoA = Zs::off, kHiA = kStage;
  static constexpr int kLoB = Ks::off, kHiB = DAs::off;
  static constexpr int kLoC = 0, kHiC = Zs::off;
  static constexpr int kLoD = Ks::off, kHiD = DZs::off;
  static constexpr int kBufBytes = kHiA - kLoA > kHiD - kLoD ? (kHiA - kLoA > kHiC - kLoC ? kHiA - kLoA : kHiC - kLoC)
                                                             : (kHiD - kLoD > kHiC - kLoC ? kHiD - kLoD : kHiC - kLoC);
  static constexpr int kModelBytesPerLane = kTileModel ? (kBicyclePack + NC * NX + NC) * (int)sizeof(TIO) : 0;
  char *t8, *t4;    // tile base + lane * 8 / lane * 4
  char* tile_;      // tile base
  template <class SEC>
  MPC_HD StageRows<typename SEC::type> view() const {  // a section as a strided array (for code outside the sweeps)
    return StageRows<typename SEC::type>{row<SEC>(0, 0), kStage / (int)sizeof(typename SEC::type), kTile};
  }

  template <class SEC>
  MPC_HD typename SEC::type* row(int k, int i) const {
    char* base = sizeof(typename SEC::type) == 8 ? t8 : t4;
    return reinterpret_cast<typename SEC::type*>(base + (int64_t)k * kStage + (SEC::off + i * SEC::rowb));
  }
  // read of section row i of stage k; SM = inside a sweep: from the staged copy of the current stage visit (STAGED)
  template <class SEC, bool SM = false>
  MPC_HD typename SEC::type rd(int k, int i) const {
#ifdef __CUDA_ARCH__
    if constexpr (SM && STAGED) {
      const char* base = sizeof(typename SEC::type) == 8 ? rb8 : rb4;
      return *reinterpret_cast<const typename SEC::type*>(base + (SEC::off + i * SEC::rowb));
    }
#endif
    return *row<SEC>(k, i);
  }

  // ---- staging pipeline (STAGED, device).  kDepth buffers + mbarriers per warp.  pipe_begin: every lane makes its
  // global stores of the previous sweep visible to the asynchronous proxy, then the leader starts the copies of the
  // sweep's first kDepth stages; visit_begin waits for the current stage; visit_end hands the buffer back (all lanes have
  // read it) and the leader refills it with the stage kDepth visits ahead.  The release sits at the END of the visit, when
  // every value read from the buffer has been consumed: a warp barrier does not wait for shared-memory loads in flight,
  // and a release right after the loads let the refill overwrite a buffer whose loads were still queued behind other
  // warps' scattered global accesses (measured: 0.6 % wrong solutions with permuted batches).  Per interior-point iteration each buffer is
  // used an even number of times (4 sweeps), so lanes that sit an iteration out keep the right barrier parities.
  const char *rb8 = nullptr, *rb4 = nullptr;  // staged copy of the current stage: buffer - range lo + lane * 8 / 4
  char* sbuf = nullptr;                       // the warp's kDepth buffers (shared memory), kBufBytes each
  unsigned long long* sbar = nullptr;         // the warp's kDepth mbarriers
  unsigned wmask = 0xffffffffu;               // lanes taking part in the current iteration
  unsigned par = 0;                           // bit j: parity the next wait on barrier j uses
  int vis = 0;                                // buffer of the current visit
  static constexpr int kDepth = 3;            // buffers per warp = stage visits a copy is issued ahead
  bool leader = false;
#ifdef __CUDA_ARCH__
  __device__ __forceinline__ static unsigned sa(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
  template <int LO, int HI>
  __device__ __forceinline__ void issue_copy(int k, int j) const {
    const unsigned bar = sa(sbar + j), dst = sa(sbuf + j * kBufBytes);
    const char* src = tile_ + (int64_t)k * kStage + LO;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(HI - LO) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(HI - LO), "r"(bar)
                 : "memory");
  }
#endif
  template <int LO, int HI>
  MPC_HD void pipe_begin(int k0, int dir) {
#ifdef __CUDA_ARCH__
    if constexpr (STAGED) {
      asm volatile("fence.proxy.async.global;" ::: "memory");
      __syncwarp(wmask);
      vis = 0;
      if (leader) {
#pragma unroll
        for (int j = 0; j < kDepth; ++j)
          if (j < a.N) issue_copy<LO, HI>(k0 + j * dir, j);
      }
    }
#endif
  }
  template <int LO>
  MPC_HD void visit_begin() {
#ifdef __CUDA_ARCH__
    if constexpr (STAGED) {
      const int j = vis;
      const unsigned bar = sa(sbar + j), ph = (par >> j) & 1u;
      unsigned ok;
      do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok)
                     : "r"(bar), "r"(ph)
                     : "memory");
      } while (!ok);
      par ^= 1u << j;
      const char* buf = sbuf + j * kBufBytes - LO;
      rb8 = buf + (threadIdx.x % kTile) * 8;
      rb4 = buf + (threadIdx.x % kTile) * 4;
    }
#endif
  }
  template <int LO, int HI>
  MPC_HD void visit_end(int k_refill) {
#ifdef __CUDA_ARCH__
    if constexpr (STAGED) {
      __syncwarp(wmask);
      if (leader && k_refill >= 0 && k_refill < a.N) issue_copy<LO, HI>(k_refill, vis);
      vis = vis + 1 == kDepth ? 0 : vis + 1;
    }
#endif
  }

  MPC_HD BoxQpIpm(const BoxQpArgs<TIO>& args, const T* shared, int64_t scenario, int64_t lane, int64_t /*lanes*/)
      : a(args), sh(shared), b(scenario), bs(args.batch) {
    static_assert(sizeof(TZ) == 8 || sizeof(TZ) == 4, "4- or 8-byte sections");
    char* tile = static_cast<char*>(a.ws) + (lane / kTile) * ((int64_t)a.N * kStage);
    const int l = (int)(lane % kTile);
    tile_ = tile;
    t8 = tile + l * 8;
    t4 = tile + l * 4;
    mu_scale = T(1);
    for (int i = 0; i < NX * NX; ++i) {
      const T v = sh[SH::oQ + i] < T(0) ? -sh[SH::oQ + i] : sh[SH::oQ + i];
      mu_scale = v > mu_scale ? v : mu_scale;
    }
    for (int i = 0; i < NU * NU; ++i) {
      const T v = sh[SH::oR + i] < T(0) ? -sh[SH::oR + i] : sh[SH::oR + i];
      mu_scale = v > mu_scale ? v : mu_scale;
    }
    mu0 = mu_scale;
  }

  MPC_HD int64_t ix(int k, int i, int per) const { return ((int64_t)k * per + i) * bs + b; }     // caller's arrays
  MPC_HD bool is_ltv() const {
    if constexpr (MODEL == 2) return false;
    else if constexpr (MODEL == 3) return true;
    else return a.ltv != 0;
  }
  MPC_HD bool hasl(int i) const { return sh[SH::oLo + i] > T(-kBigBound); }
  MPC_HD bool hasu(int i) const { return sh[SH::oHi + i] < T(kBigBound); }
  MPC_HD T lo(int i) const { return sh[SH::oLo + i]; }
  MPC_HD T hi(int i) const { return sh[SH::oHi + i]; }
  MPC_HD static T abs_(T v) { return v < T(0) ? -v : v; }
  MPC_HD static T max_(T x, T y) { return x > y ? x : y; }

  // ---- one stage's iterate.  All loads are unconditional and issued together (entries without a
  // bound hold s = 1, lam = 0), so a stage visit costs one memory round trip instead of a chain.
  struct Stage {
    T z[D], sl[D], su[D], ll[D], lu[D];
  };
  template <bool SM = false>
  MPC_HD void load(int k, Stage& s) const {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      s.z[i] = (T)rd<Zs, SM>(k, i);
      s.sl[i] = (T)rd<SLs, SM>(k, i);
      s.su[i] = (T)rd<SUs, SM>(k, i);
      s.ll[i] = (T)rd<LLs, SM>(k, i);
      s.lu[i] = (T)rd<LUs, SM>(k, i);
    }
  }
  MPC_HD void store_stage(int k, const Stage& s) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      *row<Zs>(k, i) = (TZ)s.z[i];
      *row<SLs>(k, i) = (TSL)s.sl[i];
      *row<SUs>(k, i) = (TSL)s.su[i];
      *row<LLs>(k, i) = (TSL)s.ll[i];
      *row<LUs>(k, i) = (TSL)s.lu[i];
    }
  }
  // the values the next sweeps will read back (identity for float64 storage)
  MPC_HD static void round_stage(Stage& s) {
    if constexpr (kNarrowZ || kNarrowSL) {
#pragma unroll
      for (int i = 0; i < D; ++i) {
        s.z[i] = round_to<TZ>(s.z[i]);
        s.sl[i] = round_to<TSL>(s.sl[i]);
        s.su[i] = round_to<TSL>(s.su[i]);
        s.ll[i] = round_to<TSL>(s.ll[i]);
        s.lu[i] = round_to<TSL>(s.lu[i]);
      }
    }
  }

  // ---- L2 prefetch of the rows a later stage visit will load.  The workspace is streamed from HBM once per sweep
  // (it is far larger than L2); a `prefetch.global.L2` per row, issued pf_dist stage visits early, turns the
  // ~1 us DRAM round trip of those loads into an L2 hit without holding registers for the data in flight.  It pays
  // while a kernel is latency-bound and costs once it is bandwidth-bound, because lines evicted before use are
  // fetched twice; the launchers decide (pf_dist = 0: off).
  template <typename S>
  MPC_HD static void pf(const S* p) {
#ifdef __CUDA_ARCH__
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
  }
  template <class SEC, int PER>
  MPC_HD void pf_rows(int k) const {
#pragma unroll
    for (int i = 0; i < PER; ++i) pf(row<SEC>(k, i));
  }
  template <int PER, typename S>
  MPC_HD void pf_rows_io(const S* base, int k) const {
#pragma unroll
    for (int i = 0; i < PER; ++i) pf(base + ix(k, i, PER));
  }
  // one bulk L2 prefetch of a contiguous byte range of the warp's tile stage (the read set of a sweep), issued by the
  // lowest active lane: 2 instructions per stage visit instead of one prefetch per row and lane
  template <int LO, int HI>
  MPC_HD void pf_range(int k) const {
#ifdef __CUDA_ARCH__
    if ((threadIdx.x % kTile) == (unsigned)(__ffs(__activemask()) - 1))
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(tile_ + (int64_t)k * kStage + LO), "r"(HI - LO)
                   : "memory");
#else
    (void)k;
#endif
  }
  MPC_HD void pf_model(int k) const {
    if constexpr (kTileModel) {
      pf_rows<MDs, kBicyclePack>(k);
    } else if constexpr (MODEL == 1) {
      pf_rows_io<kBicyclePack>(a.A, k);
    } else if (is_ltv()) {
      pf_rows_io<NX * NX>(a.A, k);
      pf_rows_io<NX * NU>(a.B, k);
    }
  }
  MPC_HD bool pf_on(int k) const {
#ifdef __CUDA_ARCH__
    // compiled out with general rows: the prefetch does not pay there (obstacle loop 3.34 -> 3.90 s at distance 2), and
    // even switched off at run time its code in the four sweeps cost that register-starved kernel 13 % (3.38 -> 3.81 s)
    if constexpr (STAGED || NC > 0) return false;
    return a.pf_dist > 0 && k >= 0 && k < a.N;
#else
    (void)k;
    return false;
#endif
  }

  template <class SEC, int PER, bool SM = false>
  MPC_HD void loadn(int k, T* v) const {
#pragma unroll
    for (int i = 0; i < PER; ++i) v[i] = (T)rd<SEC, SM>(k, i);
  }
  template <class SEC, int PER>
  MPC_HD void storen(int k, const T* v) const {
#pragma unroll
    for (int i = 0; i < PER; ++i) *row<SEC>(k, i) = (typename SEC::type)v[i];
  }
  template <int PER>
  MPC_HD void loadn_io(const TIO* base, int k, T* v) const {
#pragma unroll
    for (int i = 0; i < PER; ++i) v[i] = (T)base[ix(k, i, PER)];
  }

  // general row j of stage k: coefficients C[NX] and right-hand side h
  template <bool SM = false>
  MPC_HD T load_row_c(int k, int j, T* C) const {
    if constexpr (kTileModel) {
#pragma unroll
      for (int i = 0; i < NX; ++i) C[i] = (T)rd<CGs, SM>(k, j * NX + i);
      return (T)rd<HGs, SM>(k, j);
    } else {
#pragma unroll
      for (int i = 0; i < NX; ++i) C[i] = (T)a.Cg[ix(k, j * NX + i, NC * NX)];
      return (T)a.hg[ix(k, j, NC)];
    }
  }
  MPC_HD static T dotx(const T* C, const T* x) {
    T acc = T(0);
#pragma unroll
    for (int i = 0; i < NX; ++i) acc = fma_<T>(C[i], x[i], acc);
    return acc;
  }

  template <bool SM = false>
  MPC_HD void load_model(int k, T* A, T* B, T* c) const {
    if constexpr (kPacked) {
      static_assert(!kPacked || (NX == 4 && NU == 2), "packed bicycle model is 4 x 2");
      T v[kBicyclePack];
      if constexpr (kTileModel) loadn<MDs, kBicyclePack, SM>(k, v);
      else loadn_io<kBicyclePack>(a.A, k, v);
#pragma unroll
      for (int i = 0; i < NX * NX; ++i) A[i] = T(0);
#pragma unroll
      for (int i = 0; i < NX * NU; ++i) B[i] = T(0);
      A[0] = T(1);
      A[2] = v[0];
      A[3] = v[1];
      A[5] = T(1);
      A[6] = v[2];
      A[7] = v[3];
      A[10] = T(1);
      A[11] = v[4];
      A[15] = v[5];
      B[1] = v[6];
      B[3] = v[7];
      B[5] = v[8];
      B[6] = v[9];
#pragma unroll
      for (int i = 0; i < NX; ++i) c[i] = v[10 + i];
      return;
    }
    if (is_ltv()) {
      loadn_io<NX * NX>(a.A, k, A);
      loadn_io<NX * NU>(a.B, k, B);
      loadn_io<NX>(a.c, k, c);
    } else {
#pragma unroll
      for (int i = 0; i < NX * NX; ++i) A[i] = sh[SH::oA + i];
#pragma unroll
      for (int i = 0; i < NX * NU; ++i) B[i] = sh[SH::oB + i];
#pragma unroll
      for (int i = 0; i < NX; ++i) c[i] = T(0);
    }
  }

  // x+ = A x + B u + c
  MPC_HD static void step(const T* A, const T* B, const T* c, const T* x, const T* u, T* xn) {
#pragma unroll
    for (int i = 0; i < NX; ++i) xn[i] = c[i];
    mv<T, NX, NX, true>(A, x, xn);
    mv<T, NX, NU, true>(B, u, xn);
  }

  // ---- start point: inputs clamped into their box, states by rollout, slacks >= 1, lam = mu0/s
  // with mu0 = max(mu_scale, |H z0|_inf): multipliers start at the size of the cost gradient.
  // Pass 0 only measures |H z0|_inf, pass 1 writes the start point.
  MPC_HD void init() {
    T g0 = T(0);
    for (int pass = 0; pass < 2; ++pass) {
      T x[NX], xn[NX], u[NU], A[NX * NX], B[NX * NU], c[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = (T)a.x0[i * bs + b];
      for (int k = 0; k < a.N; ++k) {
        load_model(k, A, B, c);
#pragma unroll
        for (int j = 0; j < NU; ++j) {
          T v = a.warm_U ? (T)a.warm_U[ix(k, j, NU)] : T(0);
          v = v < lo(j) ? lo(j) : v;
          v = v > hi(j) ? hi(j) : v;
          u[j] = v;
        }
        step(A, B, c, x, u, xn);
        if (pass == 0) {
          const T* Qx = sh + (k == a.N - 1 ? SH::oPf : SH::oQ);
#pragma unroll
          for (int i = 0; i < NU; ++i) {
            T acc = T(0);
#pragma unroll
            for (int j = 0; j < NU; ++j) acc = fma_<T>(sh[SH::oR + i * NU + j], u[j], acc);
            g0 = max_(abs_(acc), g0);
          }
#pragma unroll
          for (int i = 0; i < NX; ++i) {
            T acc = T(0);
#pragma unroll
            for (int j = 0; j < NX; ++j) acc = fma_<T>(Qx[i * NX + j], xn[j], acc);
            g0 = max_(abs_(acc), g0);
          }
        } else {
          Stage st;
#pragma unroll
          for (int i = 0; i < D; ++i) {
            const T zi = round_to<TZ>(i < NU ? u[i] : xn[i - NU]);
            T s_l = T(1), s_u = T(1), l_l = T(0), l_u = T(0);
            if (hasl(i)) {
              s_l = zi - lo(i);
              s_l = s_l > T(1) ? s_l : T(1);
              l_l = mu0 / s_l;
            }
            if (hasu(i)) {
              s_u = hi(i) - zi;
              s_u = s_u > T(1) ? s_u : T(1);
              l_u = mu0 / s_u;
            }
            st.z[i] = zi;
            st.sl[i] = s_l;
            st.su[i] = s_u;
            st.ll[i] = l_l;
            st.lu[i] = l_u;
          }
          store_stage(k, st);
          if constexpr (NC > 0) {
#pragma unroll 1
            for (int j = 0; j < NC; ++j) {
              T C[NX];
              const T h = load_row_c(k, j, C);
              const T w = dotx(C, xn) - h;
              const T s = round_to<TSL>(w > T(1) ? w : T(1));
              *row<SCs>(k, j) = (TSL)s;
              *row<LCs>(k, j) = (TSL)(mu0 / s);
              // the row residual is carried, not recomputed: it decays exactly by (1 - alpha) per step, whereas
              // C x - h - s recomputed from a dot product keeps ~1e-16 of rounding noise that Sigma ~ 1e12 amplifies
              *row<RCs>(k, j) = (TZ)(w - s);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) x[i] = xn[i];
      }
      if (pass == 0) mu0 = g0 > mu_scale ? g0 : mu_scale;
    }
  }

  // second-order term cc = ds_aff * dlam_aff of one bound, recomputed from the stored dz_aff
  // (affine direction: ds = +-dz + r, dlam = -lam - Sigma ds with Sigma = lam/s).
  MPC_HD static T cc_of(T dz_signed, T r, T sig, T l) {
    const T ds = dz_signed + r;
    return ds * (-l - sig * ds);
  }

  // primal residual of a bound as the convergence test sees it: with narrow storage the recomputed residual carries
  // the storage rounding of z and s (a few ulp of the stored type), which is not an infeasibility
  MPC_HD static T rp_of(T r, T zi, T bound, T s) {
    T ar = abs_(r);
    if constexpr (kNarrowZ || kNarrowSL) {
      const T ulp = T(2.4e-7);  // 4 ulp of float32
      T allow = T(0);
      if constexpr (kNarrowZ) allow += ulp * (abs_(zi) + abs_(bound));
      if constexpr (kNarrowSL) allow += ulp * s;
      ar = ar > allow ? ar - allow : T(0);
    }
    return ar;
  }

  // step of one bound: (s, lam) += alpha * (ds, dlam), rounded to the stored type.  r = residual of the bound at the
  // OLD iterate, dzs / das = the signed direction (+dz for a lower bound, -dz for an upper one)
  MPC_HD static void step_bound(T& s, T& l, T r, T dzs, T das, T tau, T alpha) {
    const T ds = dzs + r;
    const T inv = rcp_(s), sg = l * inv;
    const T cc = cc_of(das, r, sg, l);
    const T dl = (tau - cc) * inv - l - sg * ds;
    T sn = s + alpha * ds, ln = l + alpha * dl;
    if constexpr (kNarrowSL || kNarrowDir) {  // the ratio test ran on unrounded values: keep the margin
      sn = max_(sn, T(1e-3) * s);
      ln = max_(ln, T(1e-3) * l);
    }
    s = round_to<TSL>(sn);
    l = round_to<TSL>(ln);
  }

  // ---- the step of one stage: (z, s, lam) += alpha * direction, slack / multiplier directions recomputed from the
  // stored dz and dz_aff exactly as sweep D computed them.  Used by sweep A (fused update) and by the output pass.
  MPC_HD void apply_step(Stage& st, const T* dz, const T* da, T tau, T alpha) const {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const T zi = st.z[i];
      if (hasl(i)) step_bound(st.sl[i], st.ll[i], zi - lo(i) - st.sl[i], dz[i], da[i], tau, alpha);
      if (hasu(i)) step_bound(st.su[i], st.lu[i], hi(i) - zi - st.su[i], -dz[i], -da[i], tau, alpha);
      st.z[i] = round_to<TZ>(zi + alpha * dz[i]);
    }
  }
  // step of the general rows of stage k (in place in the workspace)
  template <bool SM = false>
  MPC_HD void apply_step_rows(int k, const T* dz, const T* da, T tau, T alpha) {
    static_assert(!(STAGED && NC > 0), "the staged sweeps do not carry general rows yet (their in-place update)");
    if constexpr (NC > 0) {
#pragma unroll 1
      for (int j = 0; j < NC; ++j) {
        T C[NX];
        (void)load_row_c<SM>(k, j, C);
        const T s = (T)rd<SCs, SM>(k, j), l = (T)rd<LCs, SM>(k, j);
        const T r = (T)rd<RCs, SM>(k, j);
        const T ds = dotx(C, dz + NU) + r;
        const T inv = rcp_(s), sgc = l * inv;
        const T cc = cc_of(dotx(C, da + NU), r, sgc, l);
        const T dl = (tau - cc) * inv - l - sgc * ds;
        T sn = s + alpha * ds, ln = l + alpha * dl;
        if constexpr (kNarrowSL || kNarrowDir) {
          sn = max_(sn, T(1e-3) * s);
          ln = max_(ln, T(1e-3) * l);
        }
        // the residual is carried: it must absorb the storage rounding of the slack, or C x - h = s + r drifts by an
        // ulp of s per iteration (2e-7 on the active rows after ~15 iterations with float32 slacks)
        const T sr = round_to<TSL>(sn);
        *row<SCs>(k, j) = (TSL)sr;
        *row<LCs>(k, j) = (TSL)ln;
        *row<RCs>(k, j) = (TZ)((T(1) - alpha) * r + (sn - sr));
      }
    }
  }

  // inverse of a small symmetric positive definite matrix
  MPC_HD static void sym_inverse(const T* S, T* Si) {
    if constexpr (NU == 1) {
      Si[0] = rcp_(S[0]);
    } else if constexpr (NU == 2) {
      const T det = S[0] * S[3] - S[1] * S[2];
      const T id = rcp_(det);
      Si[0] = S[3] * id;
      Si[1] = -S[1] * id;
      Si[2] = -S[2] * id;
      Si[3] = S[0] * id;
    } else {
      // Gauss-Jordan without pivoting (SPD)
      T M[NU * NU];
#pragma unroll
      for (int i = 0; i < NU * NU; ++i) {
        M[i] = S[i];
        Si[i] = (i / NU == i % NU) ? T(1) : T(0);
      }
#pragma unroll
      for (int p = 0; p < NU; ++p) {
        const T inv = rcp_(M[p * NU + p]);
#pragma unroll
        for (int j = 0; j < NU; ++j) {
          M[p * NU + j] *= inv;
          Si[p * NU + j] *= inv;
        }
#pragma unroll
        for (int r = 0; r < NU; ++r) {
          if (r == p) continue;
          const T f = M[r * NU + p];
#pragma unroll
          for (int j = 0; j < NU; ++j) {
            M[r * NU + j] = fma_<T>(-f, M[p * NU + j], M[r * NU + j]);
            Si[r * NU + j] = fma_<T>(-f, Si[p * NU + j], Si[r * NU + j]);
          }
        }
      }
    }
  }

  // feed-forward recursion of one stage, shared by sweeps A and C:
  //   h = -(rhs_x + pacc);  gu = rhs_u - B'h;  d = Sinv gu;  pacc <- -A'h + K'gu
  MPC_HD static void ff_stage(const T* A, const T* B, const T* K, const T* Sinv, const T* rhs, T* pacc, T* dff) {
    T h[NX], gu[NU];
#pragma unroll
    for (int i = 0; i < NX; ++i) h[i] = -(rhs[NU + i] + pacc[i]);
#pragma unroll
    for (int j = 0; j < NU; ++j) {
      T acc = rhs[j];
#pragma unroll
      for (int i = 0; i < NX; ++i) acc = fma_<T>(-B[i * NU + j], h[i], acc);
      gu[j] = acc;
    }
    mv<T, NU, NU, false>(Sinv, gu, dff);
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      T acc = T(0);
#pragma unroll
      for (int l = 0; l < NX; ++l) acc = fma_<T>(-A[l * NX + i], h[l], acc);
#pragma unroll
      for (int j = 0; j < NU; ++j) acc = fma_<T>(K[j * NX + i], gu[j], acc);
      pacc[i] = acc;
    }
  }

  // ---- sweep A (backward): apply the previous step (have_step), factorise, affine feed-forward.
  // Affine right-hand side: rhs_i = -(H z)_i - Sigma_l r_l + Sigma_u r_u
  MPC_HD void sweep_a(const bool have_step, T tau, T alpha) {
    T Pacc[NX * NX], pacc[NX];
#pragma unroll
    for (int i = 0; i < NX * NX; ++i) Pacc[i] = sh[SH::oPf + i];
#pragma unroll
    for (int i = 0; i < NX; ++i) pacc[i] = T(0);
    pipe_begin<kLoA, kHiA>(a.N - 1, -1);
    for (int k = a.N - 1; k >= 0; --k) {
      visit_begin<kLoA>();
      if (pf_on(k - a.pf_dist)) {
        pf_range<kLoA, kHiA>(k - a.pf_dist);
        if constexpr (!kTileModel) pf_model(k - a.pf_dist);
      }
      T znew[D], sig[D], rhs[D];
      T A[NX * NX], B[NX * NU], c[NX];
      if constexpr (D <= 3) {
        // (2,1): the whole iterate in double registers, one step / store, then the bound terms (the variant ptxas fits
        // into 128 registers without spills; the element-wise form below costs ~100 bytes of spills there)
        Stage cur;
        load<true>(k, cur);
        load_model<true>(k, A, B, c);
        T dz[D], da[D];
        if (have_step) {
          loadn<DZs, D, true>(k, dz);
          loadn<DAs, D, true>(k, da);
        }
        if (have_step) {
          apply_step_rows<true>(k, dz, da, tau, alpha);
          apply_step(cur, dz, da, tau, alpha);
          store_stage(k, cur);
        }
#pragma unroll
        for (int i = 0; i < D; ++i) {
          T sg = T(0), r = T(0);
          if (hasl(i)) {
            const T s = cur.sl[i], l = cur.ll[i];
            const T sgl = l * rcp_(s);
            sg += sgl;
            r = fma_<T>(-sgl, cur.z[i] - lo(i) - s, r);
          }
          if (hasu(i)) {
            const T s = cur.su[i], l = cur.lu[i];
            const T sgu = l * rcp_(s);
            sg += sgu;
            r = fma_<T>(sgu, hi(i) - cur.z[i] - s, r);
          }
          znew[i] = cur.z[i];
          sig[i] = sg;
          rhs[i] = r;
        }
      } else {
        // ALL loads of the stage visit are issued first (one memory round trip; the stores below would otherwise fence
        // the later loads, the compiler cannot prove that the sections do not alias); slacks and multipliers stay in
        // their stored type until they are used, which keeps the (4,2) instantiation inside the register file.
        TZ zr[D];
        TSL slr[D], sur[D], llr[D], lur[D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
          zr[i] = rd<Zs, true>(k, i);
          slr[i] = rd<SLs, true>(k, i);
          sur[i] = rd<SUs, true>(k, i);
          llr[i] = rd<LLs, true>(k, i);
          lur[i] = rd<LUs, true>(k, i);
        }
        T dz[D], da[D];
        if (have_step) {
          loadn<DZs, D, true>(k, dz);
          loadn<DAs, D, true>(k, da);
        }
        load_model<true>(k, A, B, c);
        if (have_step) apply_step_rows<true>(k, dz, da, tau, alpha);
        // per element: (previous step applied,) Sigma and the bound part of the affine right-hand side
#pragma unroll
        for (int i = 0; i < D; ++i) {
          const T zi = (T)zr[i];
          const T zn = have_step ? round_to<TZ>(zi + alpha * dz[i]) : zi;
          T sg = T(0), r = T(0);
          if (hasl(i)) {
            T s = (T)slr[i], l = (T)llr[i];
            if (have_step) {
              step_bound(s, l, zi - lo(i) - s, dz[i], da[i], tau, alpha);
              *row<SLs>(k, i) = (TSL)s;
              *row<LLs>(k, i) = (TSL)l;
            }
            const T sgl = l * rcp_(s);
            sg += sgl;
            r = fma_<T>(-sgl, zn - lo(i) - s, r);
          }
          if (hasu(i)) {
            T s = (T)sur[i], l = (T)lur[i];
            if (have_step) {
              step_bound(s, l, hi(i) - zi - s, -dz[i], -da[i], tau, alpha);
              *row<SUs>(k, i) = (TSL)s;
              *row<LUs>(k, i) = (TSL)l;
            }
            const T sgu = l * rcp_(s);
            sg += sgu;
            r = fma_<T>(sgu, hi(i) - zn - s, r);
          }
          if (have_step) *row<Zs>(k, i) = (TZ)zn;
          znew[i] = zn;
          sig[i] = sg;
          rhs[i] = r;
        }
      }
      // -(H z): inputs weighted by R, state x_{k+1} by Q (Pf for the last stage)
      {
        const T* Qx = sh + (k == a.N - 1 ? SH::oPf : SH::oQ);
#pragma unroll
        for (int i = 0; i < NU; ++i) {
          T acc = rhs[i];
#pragma unroll
          for (int j = 0; j < NU; ++j) acc = fma_<T>(-sh[SH::oR + i * NU + j], znew[j], acc);
          rhs[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          T acc = rhs[NU + i];
#pragma unroll
          for (int j = 0; j < NX; ++j) acc = fma_<T>(-Qx[i * NX + j], znew[NU + j], acc);
          rhs[NU + i] = acc;
        }
      }
      if constexpr (NC > 0) {
        // general rows: value w = C x_{k+1}; adds C' Sigma_c C to the stage Hessian and C' rhs_c to the gradient
#pragma unroll 1
        for (int j = 0; j < NC; ++j) {
          T C[NX];
          (void)load_row_c<true>(k, j, C);
          const T s = (T)rd<SCs, true>(k, j), l = (T)rd<LCs, true>(k, j);
          const T sgc = l * rcp_(s);
          const T r = (T)rd<RCs, true>(k, j);
          const T rhs_c = -sgc * r;
#pragma unroll
          for (int i = 0; i < NX; ++i) rhs[NU + i] = fma_<T>(C[i], rhs_c, rhs[NU + i]);
#pragma unroll
          for (int i = 0; i < NX; ++i)
#pragma unroll
            for (int i2 = 0; i2 < NX; ++i2) Pacc[i * NX + i2] = fma_<T>(sgc * C[i], C[i2], Pacc[i * NX + i2]);
        }
      }
      // P = Pacc + diag(Sigma_x) (+ rows' C' Sigma_c C, added above)
#pragma unroll
      for (int i = 0; i < NX; ++i) Pacc[i * NX + i] += sig[NU + i];
      // Joseph (symmetric) form of the Riccati update:
      //   PB = P B,  S = Rt + B'PB,  K = -S^-1 (PB)'A,  Acl = A + B K,  Pacc <- Q + Acl' P Acl + K' Rt K
      // with Rt = R + diag(Sigma_u).  Every term is a positive semidefinite sum: the huge barrier weights inside P
      // (lam/s ~ 1e12 on active states / rows) meet closed-loop rows Acl_i ~ 1/Sigma and drop out, whereas the short
      // form Q + A'(PA + PB K) subtracts two O(Sigma) products and loses eps * Sigma of absolute accuracy.
      T K[NU * NX], Sinv[NU * NU];
      {
        T PB[NX * NU], S[NU * NU], Rt[NU * NU];
        mm<T, NX, NX, NU, false>(Pacc, B, PB);
#pragma unroll
        for (int i = 0; i < NU * NU; ++i) Rt[i] = sh[SH::oR + i];
#pragma unroll
        for (int i = 0; i < NU; ++i) Rt[i * NU + i] += sig[i];
#pragma unroll
        for (int i = 0; i < NU * NU; ++i) S[i] = Rt[i];
        mtm<T, NU, NX, NU, true>(B, PB, S);
        sym_inverse(S, Sinv);
        T G[NU * NX];
        mtm<T, NU, NX, NX, false>(PB, A, G);  // (PB)'A
#pragma unroll
        for (int i = 0; i < NU; ++i)
#pragma unroll
          for (int j = 0; j < NX; ++j) {
            T acc = T(0);
#pragma unroll
            for (int l = 0; l < NU; ++l) acc = fma_<T>(Sinv[i * NU + l], G[l * NX + j], acc);
            K[i * NX + j] = -acc;
          }
        if constexpr (sizeof(TGN) < sizeof(T)) {  // later sweeps read the stored gains: use the same values here
#pragma unroll
          for (int i = 0; i < NU * NX; ++i) K[i] = round_to<TGN>(K[i]);
#pragma unroll
          for (int i = 0; i < NU * NU; ++i) Sinv[i] = round_to<TGN>(Sinv[i]);
        }
        T Acl[NX * NX], Tm[NX * NX];
#pragma unroll
        for (int i = 0; i < NX * NX; ++i) Acl[i] = A[i];
        mm<T, NX, NU, NX, true>(B, K, Acl);
        mm<T, NX, NX, NX, false>(Pacc, Acl, Tm);
        mm<T, NU, NU, NX, false>(Rt, K, G);  // G <- Rt K
#pragma unroll
        for (int i = 0; i < NX; ++i)
#pragma unroll
          for (int j = i; j < NX; ++j) {
            T acc = sh[SH::oQ + i * NX + j];
#pragma unroll
            for (int l = 0; l < NX; ++l) acc = fma_<T>(Acl[l * NX + i], Tm[l * NX + j], acc);
#pragma unroll
            for (int l = 0; l < NU; ++l) acc = fma_<T>(K[l * NX + i], G[l * NX + j], acc);
            Pacc[i * NX + j] = acc;
            Pacc[j * NX + i] = acc;
          }
        storen<Ks, NU * NX>(k, K);
        storen<Ss, NU * NU>(k, Sinv);
      }
      T dff[NU];
      ff_stage(A, B, K, Sinv, rhs, pacc, dff);
      storen<Ds, NU>(k, dff);
      visit_end<kLoA, kHiA>(k - kDepth);
    }
  }

  struct Acc {
    T qmax, s0, s1, s2, dzmax, rp, zn;  // qmax = max_i(-ds_i/s_i, -dlam_i/lam_i): largest feasible step = 1/qmax
    MPC_HD T amin() const { return qmax > T(0) ? T(1) / qmax : T(1e30); }
  };

  // ---- sweep B (forward): affine direction by rollout with the stored gains.  For the affine direction
  // dlam = -lam (1 + ds/s), so per bound the step ratio is max(-t, 1 + t) with t = ds/s, sum(s dlam + lam ds) = -sum(s lam)
  // and cc = ds dlam; one reciprocal per bound.  Stores dz_aff, e = 1/s_l - 1/s_u, g = cc_l/s_l - cc_u/s_u
  // (general rows add C'(1/s_c) and C'(cc_c/s_c) to the state part of e and g).
  MPC_HD void sweep_b(Acc& acc) {
    T x[NX], xn[NX], u[NU];
    acc.qmax = acc.s0 = acc.s1 = acc.s2 = acc.dzmax = acc.rp = T(0);
    acc.zn = T(1);
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = T(0);  // dx_0 = 0
    pipe_begin<kLoB, kHiB>(0, 1);
    for (int k = 0; k < a.N; ++k) {
      visit_begin<kLoB>();
      if (pf_on(k + a.pf_dist)) {
        pf_range<kLoB, kHiB>(k + a.pf_dist);
        if constexpr (!kTileModel) pf_model(k + a.pf_dist);
      }
      Stage cur;
      load<true>(k, cur);
      T K[NU * NX], dff[NU];
      loadn<Ks, NU * NX, true>(k, K);
      loadn<Ds, NU, true>(k, dff);
      T A[NX * NX], B[NX * NU], c[NX];
      load_model<true>(k, A, B, c);
#pragma unroll
      for (int j = 0; j < NU; ++j) u[j] = dff[j];
      mv<T, NU, NX, true>(K, x, u);
      mv<T, NX, NX, false>(A, x, xn);
      mv<T, NX, NU, true>(B, u, xn);
      T dzv[D], ev[D], gv[D];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const T dz = round_to<TDA>(i < NU ? u[i] : xn[i - NU]);  // cc must use the value sweeps D and A read back
        dzv[i] = dz;
        const T zi = cur.z[i];
        T e = T(0), g = T(0);
        if (hasl(i)) {
          const T s = cur.sl[i], l = cur.ll[i];
          const T r = zi - lo(i) - s;
          const T ds = dz + r;
          const T inv = rcp_(s);
          const T t = ds * inv;
          const T cc = -ds * l * (T(1) + t);
          acc.qmax = max_(acc.qmax, max_(-t, T(1) + t));
          acc.s0 = fma_<T>(s, l, acc.s0);
          acc.s2 += cc;
          e += inv;
          g = fma_<T>(cc, inv, g);
        }
        if (hasu(i)) {
          const T s = cur.su[i], l = cur.lu[i];
          const T r = hi(i) - zi - s;
          const T ds = -dz + r;
          const T inv = rcp_(s);
          const T t = ds * inv;
          const T cc = -ds * l * (T(1) + t);
          acc.qmax = max_(acc.qmax, max_(-t, T(1) + t));
          acc.s0 = fma_<T>(s, l, acc.s0);
          acc.s2 += cc;
          e -= inv;
          g = fma_<T>(-cc, inv, g);
        }
        ev[i] = e;
        gv[i] = g;
      }
      if constexpr (NC > 0) {
#pragma unroll 1
        for (int j = 0; j < NC; ++j) {
          T C[NX];
          (void)load_row_c<true>(k, j, C);
          const T s = (T)rd<SCs, true>(k, j), l = (T)rd<LCs, true>(k, j);
          const T r = (T)rd<RCs, true>(k, j);
          const T ds = dotx(C, dzv + NU) + r;
          const T inv = rcp_(s);
          const T t = ds * inv;
          const T cc = -ds * l * (T(1) + t);
          acc.qmax = max_(acc.qmax, max_(-t, T(1) + t));
          acc.s0 = fma_<T>(s, l, acc.s0);
          acc.s2 += cc;
          const T gi = cc * inv;
#pragma unroll
          for (int i = 0; i < NX; ++i) {
            ev[NU + i] = fma_<T>(C[i], inv, ev[NU + i]);
            gv[NU + i] = fma_<T>(C[i], gi, gv[NU + i]);
          }
        }
      }
      storen<DAs, D>(k, dzv);
      storen<Es, D>(k, ev);
      storen<Gs, D>(k, gv);
      visit_end<kLoB, kHiB>(k + kDepth);
      // the rollout continues with the UNROUNDED state direction (the stored copy is only used for cc)
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = xn[i];
    }
    acc.s1 = -acc.s0;
  }

  // ---- sweep C (backward): feed-forward of the corrector's extra right-hand side  tau e - g, added to the affine one
  MPC_HD void sweep_c(T tau) {
    T pacc[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) pacc[i] = T(0);
    pipe_begin<kLoC, kHiC>(a.N - 1, -1);
    for (int k = a.N - 1; k >= 0; --k) {
      visit_begin<kLoC>();
      if (pf_on(k - a.pf_dist)) {
        pf_range<kLoC, kHiC>(k - a.pf_dist);
        if constexpr (kTileModel) pf_range<MDs::off, MDs::off + kMdRows * MDs::rowb>(k - a.pf_dist);
        else pf_model(k - a.pf_dist);
      }
      T ev[D], gv[D], K[NU * NX], Sinv[NU * NU];
      loadn<Es, D, true>(k, ev);
      loadn<Gs, D, true>(k, gv);
      loadn<Ks, NU * NX, true>(k, K);
      loadn<Ss, NU * NU, true>(k, Sinv);
      T A[NX * NX], B[NX * NU], c[NX];
      load_model<true>(k, A, B, c);
      T rhs[D];
#pragma unroll
      for (int i = 0; i < D; ++i) rhs[i] = fma_<T>(tau, ev[i], -gv[i]);
      T dff[NU], daff[NU];
      loadn<Ds, NU, true>(k, daff);
      ff_stage(A, B, K, Sinv, rhs, pacc, dff);
#pragma unroll
      for (int j = 0; j < NU; ++j) dff[j] += daff[j];  // d_aff + d_cor: sweep D rolls the whole direction out at once
      storen<Ds, NU>(k, dff);
      visit_end<kLoC, kHiC>(k - kDepth);
    }
  }

  // ---- sweep D (forward): dz by rollout with the summed feed-forward (exactly on the linearised dynamics); per bound
  // the slack / multiplier directions (second-order term from the stored dz_aff), the step ratios and the sums that
  // give the new barrier parameter for any step length.
  MPC_HD void sweep_d(T tau, Acc& acc) {
    T x[NX], xn[NX], u[NU];
    acc.qmax = acc.s0 = acc.s1 = acc.s2 = acc.dzmax = acc.rp = T(0);
    acc.zn = T(1);
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = T(0);
    pipe_begin<kLoD, kHiD>(0, 1);
    for (int k = 0; k < a.N; ++k) {
      visit_begin<kLoD>();
      if (pf_on(k + a.pf_dist)) {
        pf_range<kLoD, kHiD>(k + a.pf_dist);
        if constexpr (!kTileModel) pf_model(k + a.pf_dist);
      }
      Stage cur;
      load<true>(k, cur);
      T K[NU * NX], dff[NU], da[D];
      loadn<Ks, NU * NX, true>(k, K);
      loadn<Ds, NU, true>(k, dff);
      loadn<DAs, D, true>(k, da);
      T A[NX * NX], B[NX * NU], c[NX];
      load_model<true>(k, A, B, c);
#pragma unroll
      for (int j = 0; j < NU; ++j) u[j] = dff[j];
      mv<T, NU, NX, true>(K, x, u);
      mv<T, NX, NX, false>(A, x, xn);
      mv<T, NX, NU, true>(B, u, xn);
      T dzv[D];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const T dz = round_to<TDZ>(i < NU ? u[i] : xn[i - NU]);
        dzv[i] = dz;
        const T zi = cur.z[i];
        acc.dzmax = max_(acc.dzmax, abs_(dz));
        acc.zn = max_(acc.zn, abs_(zi));
        if (hasl(i)) {
          const T s = cur.sl[i], l = cur.ll[i];
          const T r = zi - lo(i) - s;
          const T ds = dz + r;
          const T rinv = rcp_(s * l), inv_s = rinv * l, inv_l = rinv * s;
          const T sgl = l * inv_s;
          const T cc = cc_of(da[i], r, sgl, l);
          const T dl = (tau - cc) * inv_s - l - sgl * ds;
          acc.qmax = max_(acc.qmax, max_(-ds * inv_s, -dl * inv_l));  // step is limited to 1 / max(q)
          acc.s0 = fma_<T>(s, l, acc.s0);
          acc.s1 += s * dl + l * ds;
          acc.s2 = fma_<T>(ds, dl, acc.s2);
          acc.rp = max_(acc.rp, rp_of(r, zi, lo(i), s));
        }
        if (hasu(i)) {
          const T s = cur.su[i], l = cur.lu[i];
          const T r = hi(i) - zi - s;
          const T ds = -dz + r;
          const T rinv = rcp_(s * l), inv_s = rinv * l, inv_l = rinv * s;
          const T sgu = l * inv_s;
          const T cc = cc_of(-da[i], r, sgu, l);
          const T dl = (tau - cc) * inv_s - l - sgu * ds;
          acc.qmax = max_(acc.qmax, max_(-ds * inv_s, -dl * inv_l));
          acc.s0 = fma_<T>(s, l, acc.s0);
          acc.s1 += s * dl + l * ds;
          acc.s2 = fma_<T>(ds, dl, acc.s2);
          acc.rp = max_(acc.rp, rp_of(r, zi, hi(i), s));
        }
      }
      if constexpr (NC > 0) {
#pragma unroll 1
        for (int j = 0; j < NC; ++j) {
          T C[NX];
          (void)load_row_c<true>(k, j, C);
          const T s = (T)rd<SCs, true>(k, j), l = (T)rd<LCs, true>(k, j);
          const T r = (T)rd<RCs, true>(k, j);
          const T ds = dotx(C, dzv + NU) + r;
          const T rinv = rcp_(s * l), inv_s = rinv * l, inv_l = rinv * s;
          const T sgc = l * inv_s;
          const T cc = cc_of(dotx(C, da + NU), r, sgc, l);
          const T dl = (tau - cc) * inv_s - l - sgc * ds;
          acc.qmax = max_(acc.qmax, max_(-ds * inv_s, -dl * inv_l));
          acc.s0 = fma_<T>(s, l, acc.s0);
          acc.s1 += s * dl + l * ds;
          acc.s2 = fma_<T>(ds, dl, acc.s2);
          acc.rp = max_(acc.rp, abs_(r));
        }
      }
      storen<DZs, D>(k, dzv);
      visit_end<kLoD, kHiD>(k + kDepth);
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = xn[i];
    }
  }

  // ---- output: last step applied on the fly, active set from the complementarity pairs, variables snapped onto
  // active bounds, states by rollout of the snapped inputs, cost as the reference defines it.
  MPC_HD void output(int status, int iters, const bool have_step, T tau, T alpha) {
    T x[NX], xn[NX], u[NU], A[NX * NX], B[NX * NU], c[NX];
    T cost = T(0);
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      x[i] = (T)a.x0[i * bs + b];
      a.X[i * bs + b] = (TIO)x[i];
    }
    for (int k = 0; k < a.N; ++k) {
      Stage st;
      load(k, st);
      load_model(k, A, B, c);
      if (have_step) {
        T dz[D], da[D];
        loadn<DZs, D>(k, dz);
        loadn<DAs, D>(k, da);
        apply_step_rows(k, dz, da, tau, alpha);
        apply_step(st, dz, da, tau, alpha);
      }
#pragma unroll
      for (int i = 0; i < D; ++i) {
        int sat = 0;
        if (hasl(i) && st.ll[i] > st.sl[i]) sat = -1;
        if (hasu(i) && st.lu[i] > st.su[i]) sat = 1;
        if (i < NU) {
          u[i] = sat < 0 ? lo(i) : (sat > 0 ? hi(i) : st.z[i]);
          if constexpr (sizeof(TIO) < sizeof(T)) u[i] = (T)(TIO)u[i];  // the rollout uses the input that is returned
          a.U[ix(k, i, NU)] = (TIO)u[i];
          if (a.sat_u) a.sat_u[ix(k, i, NU)] = (int8_t)sat;
        } else if (a.sat_x) {
          a.sat_x[ix(k, i - NU, NX)] = (int8_t)sat;
        }
      }
      if constexpr (NC > 0) {
        if (a.sat_c) {
#pragma unroll 1
          for (int j = 0; j < NC; ++j)
            a.sat_c[ix(k, j, NC)] = (T)*row<LCs>(k, j) > (T)*row<SCs>(k, j) ? (int8_t)-1 : (int8_t)0;
        }
      }
      cost += quad<T, NX>(sh + SH::oQ, x) + quad<T, NU>(sh + SH::oR, u);
      step(A, B, c, x, u, xn);
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        x[i] = xn[i];
        a.X[ix(k + 1, i, NX)] = (TIO)x[i];
      }
    }
    cost += quad<T, NX>(sh + SH::oPf, x);
    a.cost[b] = (TIO)cost;
    a.status[b] = status;
    a.iters[b] = iters;
  }

  // number of slack / multiplier pairs of one scenario
  MPC_HD int count_constraints() const {
    int nc = 0;
#pragma unroll
    for (int i = 0; i < D; ++i) nc += (hasl(i) ? 1 : 0) + (hasu(i) ? 1 : 0);
    return (nc + NC) * a.N;
  }

  // ---- one interior-point iteration on the caller's iteration state; true when the scenario is finished (status set)
  MPC_HD bool iterate_on(int ncons, T inv_nc, int& it, int& status, T& tau, T& alpha, bool& have_step) {
    ++it;
    const T eps = (T)a.eps;
    Acc acc;
    sweep_a(have_step, tau, alpha);
    sweep_b(acc);
    {
      const T mu = acc.s0 * inv_nc;
      const T am_aff = acc.amin();
      const T a_aff = am_aff < T(1) ? am_aff : T(1);
      const T mu_aff = (acc.s0 + a_aff * (acc.s1 + a_aff * acc.s2)) * inv_nc;
      T ratio = mu_aff / (mu > T(1e-300) ? mu : T(1e-300));
      T sigma = ratio * ratio * ratio;
      sigma = sigma < T(1) ? sigma : T(1);
      // centring target; never below 1e-3 of the complementarity tolerance: driving mu further only inflates
      // the barrier weights (lam/s ~ lam^2/mu) and with them the rounding noise of the Newton step
      tau = sigma * mu;
      const T mu_floor = T(1e-3) * eps * mu_scale;
      tau = tau > mu_floor ? tau : mu_floor;
    }
    sweep_c(tau);
    sweep_d(tau, acc);
    alpha = T(0.995) * acc.amin();
    alpha = alpha < T(1) ? alpha : T(1);
    have_step = true;
    const T zn = acc.zn;
    const T mu_new = (acc.s0 + alpha * (acc.s1 + alpha * acc.s2)) * inv_nc;
    const T rp = (T(1) - alpha) * acc.rp;
    // no finite bound: one exact Newton step -- exact when the gains are stored in the arithmetic type; with narrower
    // gains the step carries their rounding (1e-7 of the step), and one more iteration removes it
    const bool small_step = alpha * acc.dzmax <= T(1e-6) * zn;
    const bool done = (ncons == 0) ? (sizeof(TGN) == sizeof(T) || small_step)
                                   : ((mu_new <= eps * mu_scale) && (rp <= eps * zn) && small_step);
    if (done) {
      status = MPC_SOLVED;
    } else {
      // stalled: step length collapsed / barrier parameter grew 100x above its start value.  With a bound residual
      // that cannot be closed the box and the dynamics do not meet: infeasible.  Running out of iterations without
      // that signature is reported as MPC_MAX_ITER unless the barrier parameter has grown above its start value
      // while the residual is still open (multipliers diverging: the Farkas-type signature of an empty feasible set).
      const bool stalled = !(alpha >= T(1e-6)) || !(mu_new <= T(100) * mu0);
      const bool open = !(rp <= T(1e-6) * zn);
      if (stalled) status = open ? MPC_INFEASIBLE : MPC_MAX_ITER;
      else if (it >= a.max_iter) status = (open && mu_new > mu0) ? MPC_INFEASIBLE : MPC_MAX_ITER;
    }
    return status != MPC_UNSOLVED;
  }

  // ---- one scenario, start to end: the iteration state lives in locals (registers).
  // (no finite bound at all: the first iteration is one exact Newton step onto the LQ optimum -- all barrier
  // terms vanish, alpha = 1 -- and the loop stops after it.)
  MPC_HD void solve() {
    const int ncons = count_constraints();
    const T inv_nc = ncons ? T(1) / T(ncons) : T(0);
    init();
    int status = MPC_UNSOLVED, it = 0;
    T tau = T(0), alpha = T(0);
    bool have_step = false;
#ifdef __CUDA_ARCH__
    if constexpr (STAGED) {
      // warp-uniform iteration loop: the lanes still running form the mask of the staging pipeline's warp barriers,
      // their lowest lane issues the copies
      static_assert(!STAGED || !kTileModel, "staged sweeps with the model in the tile: sweep C's range lacks it");
      const unsigned all = wmask;
      bool run = true;
      while (true) {
        const unsigned m = __ballot_sync(all, run);
        if (m == 0) break;
        if (run) {
          wmask = m;
          leader = (threadIdx.x % kTile) == (unsigned)(__ffs(m) - 1);
          run = !iterate_on(ncons, inv_nc, it, status, tau, alpha, have_step);
        }
      }
      wmask = all;
      output(status, it, have_step, tau, alpha);
      return;
    }
#endif
    while (!iterate_on(ncons, inv_nc, it, status, tau, alpha, have_step)) {
    }
    output(status, it, have_step, tau, alpha);
  }
  // STAGED: the warp's staging buffers (kDepth x kBufBytes, 128-byte aligned), its kDepth mbarriers (initialised to one
  // arrival each by the caller) and the lanes of the warp that own a scenario
  MPC_HD void stage_setup(char* buffers, unsigned long long* barriers, unsigned lanes_mask) {
    sbuf = buffers;
    sbar = barriers;
    wmask = lanes_mask;
    par = 0;
  }

  // ---- the same, split for a persistent kernel that refills a lane with the next scenario as soon as its current one
  // has converged (csrc/boxqp.cu, boxqp_ipm_refill_kernel): begin(); while (!iterate()) {}; finish();
  int ncons_m = 0, status_m = MPC_UNSOLVED, it_m = 0;
  T tau_m = T(0), alpha_m = T(0), inv_nc_m = T(0);
  bool have_step_m = false;

  MPC_HD void begin(int64_t scenario) {
    b = scenario;
    ncons_m = count_constraints();
    inv_nc_m = ncons_m ? T(1) / T(ncons_m) : T(0);
    mu0 = mu_scale;
    init();
    status_m = MPC_UNSOLVED;
    it_m = 0;
    tau_m = alpha_m = T(0);
    have_step_m = false;
  }
  MPC_HD bool iterate() { return iterate_on(ncons_m, inv_nc_m, it_m, status_m, tau_m, alpha_m, have_step_m); }
  MPC_HD void finish() { output(status_m, it_m, have_step_m, tau_m, alpha_m); }
};

}  // namespace mpc
