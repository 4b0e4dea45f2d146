/*
 * mpc_b200.h -- C ABI of libmpc_b200.so: batched finite-horizon MPC solves on NVIDIA B200 (sm_100a).
 *
 * The reference (konnpaku-youmu/Model_Predictive_Control) is pure Python and has no FFI; its
 * "plugin interface" for this path is the Python call surface of
 *   session_1/FHC.py            ricatti_recursion (:51-61), AutoCruising (:20-29)
 *   session_1/LinearSystem.py   LinearSystem.f/simulate/prediction (:16-35)
 *   session_1/session1_sol.py   riccati_recursion (:44-65), simulate (:68-91)
 *   session_2|3/problem.py      Problem (:4-32 / :8-36), session_2/log.py ControllerLog (:8-12)
 *   session_4/session4_sol.py   MPCController (:113-230), integrators (:22-56)
 * Each entry point below names the reference interface it replaces.  INTEGRATION.md shows the
 * ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch); the library never keeps
 *     memory between calls.  Kernels are asynchronous on `stream` (a cudaStream_t passed as void*);
 *     no call synchronises the device.
 *   - return value: 0 = ok, < 0 = mpc_error (invalid argument), > 0 = cudaError_t.  Never aborts.
 *     mpc_last_error() returns a thread-local message for the last non-zero return.
 *   - per-scenario solver outcomes (status, iterations) are DATA, not errors.
 *   - dtype: MPC_F64 (reference arithmetic, numpy default) or MPC_F32.
 *   - "batch stride" arguments (s*) are in ELEMENTS between consecutive scenarios; 0 = the array
 *     is shared by every scenario.
 *   - matrices are row-major.
 */
#ifndef MPC_B200_H
#define MPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPC_B200_VERSION 100 /* 0.1.0 */

typedef void* mpc_stream_t; /* cudaStream_t */

enum mpc_dtype { MPC_F64 = 0, MPC_F32 = 1 };

enum mpc_error {
  MPC_OK = 0,
  MPC_ERR_NULL = -1,        /* required pointer is NULL */
  MPC_ERR_SHAPE = -2,       /* dimension out of the supported range */
  MPC_ERR_DTYPE = -3,       /* unknown dtype enum */
  MPC_ERR_ALIGN = -4,       /* pointer / stride not aligned to the element size */
  MPC_ERR_UNSUPPORTED = -5, /* valid request that this build has no kernel for */
  MPC_ERR_WORKSPACE = -6    /* workspace missing or too small */
};

/* per-scenario status codes written by the constrained solvers (ControllerLog.solver_success,
 * reference session_2/log.py:10, is `status == MPC_SOLVED`) */
enum mpc_solve_status {
  MPC_SOLVED = 1,
  MPC_MAX_ITER = 2,
  MPC_INFEASIBLE = 3,
  MPC_UNSOLVED = 0
};

int mpc_version(void);
const char* mpc_last_error(void);

/* Largest (n, m) the generic (CTA-per-scenario) kernels accept. */
#define MPC_MAX_NX 32
#define MPC_MAX_NU 16

/* ---------------------------------------------------------------------------------------------
 * K1  batched backward Riccati recursion.
 * Replaces FHC.ricatti_recursion (session_1/FHC.py:51-61) and session1_sol.riccati_recursion
 * (session_1/session1_sol.py:44-65):  P_N = Pf;  K_k = -(R + B'PB)^-1 B'PA;  P_k = Q + A'PA + A'PB K_k.
 * Sign convention u = +K x.  P is NOT symmetrised between stages (as the reference).
 *   A [*,n,n]  B [*,n,m]  Q [*,n,n]  R [*,m,m]  Pf [*,n,n]   (batch stride s?, 0 = shared)
 *   K  out [N][batch][m][n]    K[0] = first-stage gain (the reference returns the list reversed)
 *   P  out, optional: all_P != 0 -> [N+1][batch][n][n] (P[0] = cost-to-go at stage 0, P[N] = Pf)
 *                     all_P == 0 -> [batch][n][n] = P[0] only.  NULL = not written.
 */
int mpc_riccati(const void* A, int64_t sA, const void* B, int64_t sB, const void* Q, int64_t sQ,
                const void* R, int64_t sR, const void* Pf, int64_t sPf, void* K, void* P, int all_P,
                int64_t batch, int n, int m, int N, int dtype, mpc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K2  batched linear rollout under time-varying state feedback, batch-contiguous layout.
 * Replaces LinearSystem.simulate / prediction (session_1/LinearSystem.py:20-35) driven by
 * AutoCruising.control_law / pred (session_1/FHC.py:25-29) and session1_sol.simulate (:68-91).
 *   x0 [n][batch]  ->  X [T][n][batch] (X[0] = x0), U [T-1][m][batch] (optional)
 *   transition i (0-based) applies gain index g = gain_offset + gain_step * i:
 *       (0,0) gains[0] every step          AutoCruising.control_law  (FHC.py:25-26)
 *       (1,1) gains[t], t = 1..T-1         AutoCruising.pred through LinearSystem.prediction
 *                                          (FHC.py:28-29, LinearSystem.py:30-31 -- skips gains[0])
 *       (0,1) gains[t], t = 0..T-2         session1_sol.py:121-126
 *   K [ng][*][m][n] with stage stride sK_stage and batch stride sK (0 = shared gains)
 *   A, B shared (sA = sB = 0) or per scenario.
 *   cost   optional [batch]: sum_i x_i'Q x_i + u_i'R u_i  +  x_{T-1}' Pf x_{T-1}  (Q, R, Pf shared;
 *          all three required when cost != NULL)
 *   unstable optional [batch] uint8: 1 when any |x_t|_2 > norm_limit (session1_sol.py:86-89 uses 100)
 */
int mpc_lq_rollout(const void* A, int64_t sA, const void* B, int64_t sB, const void* K,
                   int64_t sK_stage, int64_t sK, int gain_offset, int gain_step, const void* x0,
                   void* X, void* U, const void* Q, const void* R, const void* Pf, void* cost,
                   uint8_t* unstable, double norm_limit, int64_t batch, int n, int m, int T,
                   int dtype, mpc_stream_t stream);

/* One plant step x+ = A x + B u with a caller-supplied input: LinearSystem.f
 * (session_1/LinearSystem.py:16-18) for policies that are arbitrary Python callables.
 *   A [n][n], B [n][m] shared;  x [n][batch], u [m][batch] -> xn [n][batch] (xn != x). */
int mpc_linear_step(const void* A, const void* B, const void* x, const void* u, void* xn,
                    int64_t batch, int n, int m, int dtype, mpc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K1+K2 fused  per-scenario finite-horizon LQ solve (one thread per scenario):
 * backward recursion of FHC.ricatti_recursion (FHC.py:51-61) followed by the optimal open-loop
 * plan u_k = K_k x_k, x_{k+1} = A x_k + B u_k (LinearSystem.f, LinearSystem.py:16-18) and the cost
 * V = sum x'Qx + u'Ru + x_N' Pf x_N  (= x0' P_0 x0, FHC.py:123-124).
 *   x0 [batch][n];  outputs stage-major:  X [N+1][batch][n], U [N][batch][m], V [batch];
 *   optional K [N][batch][m][n], P0 [batch][n][n].
 * Supported (n, m): n <= 4, m <= 2 (register-resident); others -> MPC_ERR_UNSUPPORTED.
 * Q and Pf must be symmetric (their upper triangles are used).
 * Single-input fp64 solves (m = 1, n = 2 or 4) that do not ask for K / P0 run the same recursion
 * in Krylov coordinates x = [b, Ab, .., A^{n-1}b] z, where a stage costs O(n^2) instead of O(n^3);
 * a scenario whose controllability matrix has cond_F > 1e3 (env MPC_LQ_KRYLOV_COND, 0 = never)
 * takes the dense recursion inside the same launch, so the result stays within the 1e-6 parity bar
 * for every input.  mpc_lq_solve_variant tells which kernel a call with these arguments launches:
 * 0 = lq_solve_kernel (dense stages), 1 = lq_solve_krylov_kernel.
 */
int mpc_lq_solve_variant(int n, int m, int dtype, int wants_gains_or_P0);
int mpc_lq_solve(const void* A, int64_t sA, const void* B, int64_t sB, const void* Q, int64_t sQ,
                 const void* R, int64_t sR, const void* Pf, int64_t sPf, const void* x0, void* X,
                 void* U, void* V, void* K, void* P0, int64_t batch, int n, int m, int N, int dtype,
                 mpc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K3  condensed prediction matrices of the LTI MPC problem ("condensed form" of the QP posed by the
 * reference's Problem data, session_2/problem.py:8-24):  X = Phi x0 + Gamma U,
 *   Phi = [A; ..; A^N] [N n][n],  Gamma_{ij} = A^(i-j) B [N n][N m],  H = Gamma'Qbar Gamma + Rbar
 *   [N m][N m],  F = Gamma'Qbar Phi [N m][n],  Qbar = blkdiag(Q,..,Q,Pf);  J(U) = U'HU + 2 x0'F'U + c.
 * One CTA per model (batch stride s?, 0 = shared); outputs [batch][rows][cols] row-major, any NULL.
 */
int mpc_condense(const void* A, int64_t sA, const void* B, int64_t sB, const void* Q, int64_t sQ, const void* R,
                 int64_t sR, const void* Pf, int64_t sPf, void* Phi, void* Gamma, void* H, void* F, int64_t batch,
                 int n, int m, int N, int dtype, mpc_stream_t stream);

/* K6  per-rank summary of a batch of solves / closed loops (the payload of the final NCCL gather):
 *   out8 = {scenarios, sum cost, max violation, sum saturated inputs, #infeasible, #max_iter,
 *           sum iterations, #solved};  every input array is optional (NULL). */
int mpc_summary(const void* cost, const void* viol, const int32_t* n_sat, const int32_t* status,
                const int32_t* iters, int64_t batch, double* out8, int dtype, mpc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K4  batched box-constrained linear MPC QP
 *     min  sum_{k<N} x_k'Q x_k + u_k'R u_k + x_N'Pf x_N
 *     s.t. x_{k+1} = A_k x_k + B_k u_k + c_k,  u_lo <= u_k <= u_hi (k<N),  x_lo <= x_k <= x_hi (1<=k<=N)
 * i.e. the QP posed by the reference's Problem data (session_2/problem.py:8-24,
 * session_3/problem.py:12-28; the reference has no solver for it) and, with ltv = 1, the linearised
 * QP of session 4's closed loop (session_4/session4_sol.py:158-217 cost/bounds).  Outputs follow the
 * reference's per-solve log schema ControllerLog (session_2/log.py:8-12): status (solver_success =
 * status == MPC_SOLVED), X (state_prediction), U (input_prediction).
 *   ltv = 0: A [n][n], B [n][m] shared by all scenarios and stages, c ignored
 *   ltv = 1: A [N][n*n][batch], B [N][n*m][batch], c [N][n][batch]
 *   Q, R, Pf, u_lo [m], u_hi [m], x_lo [n], x_hi [n] shared; |bound| >= 1e19 = unbounded
 *   x0 [n][batch]; warm_U optional [N][m][batch] (start point; clamped into the box)
 *   U [N][m][batch], X [N+1][n][batch], cost [batch], status [batch], iters [batch],
 *   sat_u [N][m][batch] / sat_x [N][n][batch] optional int8: -1 at lower bound, +1 at upper, 0 free.
 *   Inputs at an active bound are returned exactly equal to the bound.
 *   ws: caller-owned workspace of mpc_boxqp_workspace_bytes(...) bytes; opaque (tiles of 32 scenarios,
 *   [tile][stage][section row][lane]: the size depends on the batch through ceil(batch / 32) only).
 * Method: Mehrotra predictor-corrector interior point, Newton systems solved by Riccati sweeps; arithmetic is float64
 * for both dtypes.  MPC_F64: float64 arrays; the workspace keeps the iterate, gains and steps in float64 and slacks /
 * multipliers / affine direction / corrector data in float32 (env MPC_QP_STORE=f64: everything float64).  A shared
 * model (ltv = 0) without general rows runs on the staged kernel (tile stages copied into shared memory by
 * cp.async.bulk; env MPC_QP_STAGED=0: the kernel without staging; results are bitwise the same).  MPC_F32: float32 arrays and workspace
 * (north-star tolerance 1e-4).  Supported (n, m): (2,1), (4,1), (4,2) with one thread per
 * scenario (register-resident matrices); (12,4), MPC_F64, with a shared model (ltv = 0) on a persistent
 * warp-per-scenario kernel whose workspace is one slot per resident warp, independent of the batch.
 * status: MPC_SOLVED; MPC_INFEASIBLE when the iteration stalls (step length < 1e-6 or barrier parameter 100x above its
 * start) with a bound residual that does not close; MPC_MAX_ITER otherwise (iteration limit without that signature).
 * ws must be 16-byte aligned.
 */
int64_t mpc_boxqp_workspace_bytes(int64_t batch, int n, int m, int N, int dtype);
int mpc_boxqp_solve(const void* A, const void* B, const void* c, int ltv, const void* Q, const void* R,
                    const void* Pf, const void* u_lo, const void* u_hi, const void* x_lo,
                    const void* x_hi, const void* x0, const void* warm_U, void* U, void* X, void* cost,
                    int32_t* status, int32_t* iters, int8_t* sat_u, int8_t* sat_x, void* ws,
                    int64_t ws_bytes, int64_t batch, int n, int m, int N, int max_iter, double eps,
                    int dtype, mpc_stream_t stream);

/* Difficulty-sorted batches for the thread-per-scenario kernels.  Interior-point iteration counts differ between
 * scenarios (cfg 3: mean 10.8, max 22) and a warp runs until its slowest lane has converged.  Neighbouring initial
 * states have similar active sets and iteration counts, so:
 *   mpc_state_order_keys  writes keys[b] = Morton (Z-order) code of x0[:, b] (8 bits per coordinate, scaled by
 *                         lohi = [min_0..min_{n-1}, max_0..max_{n-1}] on the device); the caller sorts them (any stable
 *                         device sort) into order [batch] (int32);
 *   mpc_boxqp_solve_ordered = mpc_boxqp_solve where lane b of the workspace solves scenario order[b].  Every scenario's
 *                         arithmetic is independent of its lane: results are bitwise those of mpc_boxqp_solve and are
 *                         written at the scenario's own index. */
int mpc_state_order_keys(const void* x0, const void* lohi, int32_t* keys, int64_t batch, int n, int dtype,
                         mpc_stream_t stream);
int mpc_boxqp_solve_ordered(const void* A, const void* B, const void* c, int ltv, const void* Q, const void* R,
                            const void* Pf, const void* u_lo, const void* u_hi, const void* x_lo, const void* x_hi,
                            const void* x0, const void* warm_U, void* U, void* X, void* cost, int32_t* status,
                            int32_t* iters, int8_t* sat_u, int8_t* sat_x, const int32_t* order, void* ws,
                            int64_t ws_bytes, int64_t batch, int n, int m, int N, int max_iter, double eps, int dtype,
                            mpc_stream_t stream);

/* K4 with general stage rows (polytopic constraints):  additionally  Cg_k x_{k+1} >= hg_k,  k < N.
 * Replaces, after linearisation, the collision constraints of the obstacle-avoidance controller
 * (session_4/main.py:95-104: nine squared-distance constraints per stage).
 *   Cg [N][nc*n][batch] (row-major rows), hg [N][nc][batch]; sat_c optional [N][nc][batch] (-1 = active).
 * Instantiated: (n, m, nc) = (4, 2, 9) and (4, 2, 3). */
int64_t mpc_boxqp_rows_workspace_bytes(int64_t batch, int n, int m, int N, int nc, int dtype);
int mpc_boxqp_solve_rows(const void* A, const void* B, const void* c, int ltv, const void* Q, const void* R,
                         const void* Pf, const void* u_lo, const void* u_hi, const void* x_lo, const void* x_hi,
                         const void* Cg, const void* hg, int nc, const void* x0, const void* warm_U, void* U, void* X,
                         void* cost, int32_t* status, int32_t* iters, int8_t* sat_u, int8_t* sat_x, int8_t* sat_c,
                         void* ws, int64_t ws_bytes, int64_t batch, int n, int m, int N, int max_iter, double eps,
                         int dtype, mpc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K5  session-4 kinematic bicycle, real-time-iteration (RTI) MPC.
 * Replaces, for a batch of scenarios, the per-step work of MPCController.__call__/solve
 * (session_4/session4_sol.py:113-230: OCP cost/bounds :166-181, Euler model :191-192; RK4 variant
 * template.py:141) and of the closed-loop driver exercise5 (:443-465).  The reference solves the
 * nonlinear OCP with CasADi + IPOPT; here ONE linearised QP (K4, ltv = 1) is solved per control
 * step.  The bicycle ODE is our definition (rcracers is not vendored), see csrc/bicycle_core.cuh:
 *   state [p_x, p_y, psi, v], input [a, delta], parameters lr = axis_rear, lf = axis_front,
 *   friction, accel = acceleration (session_4/parameters.py:7-8,47-48).  MPC_F64 or MPC_F32 arrays (float64 arithmetic).
 *
 * mpc_bicycle_rti_prepare: shift the previous plan (first == 0) or keep it (first != 0), roll the
 *   model out from y [4][batch] and linearise:  U_prev [N][2][batch] -> warm_U [N][2][batch],
 *   A [N][16][batch], B [N][8][batch], c [N][4][batch]  (inputs of mpc_boxqp_solve with ltv = 1).
 * mpc_bicycle_plant_step: x [4][batch], u [2][batch] -> xn;  substeps = 0: forward Euler over ts
 *   (session4_sol.py:22-25), substeps > 0: RK4 sub-steps, substeps < 0: adaptive Dormand-Prince 5(4) with
 *   rtol = atol = 10^substeps (the counterpart of the reference's odeint plant, :37-56);
 *   friction [batch] (s_friction = 1) or one shared value (s_friction = 0).
 * mpc_rti_closed_loop: `steps` control steps of  sqp_iters x (prepare -> QP) -> apply u_0 -> plant  in ONE kernel.
 *   Prediction model (lr, lf, accel, friction_model, ts, rk4) and plant (plant_lr, plant_lf, plant_accel, per-scenario
 *   friction_plant [batch], plant_substeps as in mpc_bicycle_plant_step) are separate, as in the reference's mismatch
 *   study (session4_sol.py:461-465).  sqp_iters = 1: real-time iteration.  sqp_iters > 1: the OCP is re-linearised at
 *   the new plan and solved again (full-step SQP towards the converged solution IPOPT returns in the reference,
 *   session4_sol.py:126-130); a round that moves the plan by <= sqp_tol * max(1, |U|) ends the step (0 = run all).
 *   nc = 0: box constraints.  nc = 9: additionally the nine linearised collision rows of the obstacle-avoidance
 *   controller (session_4/main.py:95-104) for an obstacle at the HOST pose x_obs, vehicle length / width
 *   (forward-Euler prediction model only); clear_cl optional [batch] = smallest |c_i - o_j|^2 - (2r)^2 along the loop.
 *   U_plan [N][2][batch] in/out (initial plan; zeros = cold start), X_pred [N+1][4][batch] (last
 *   prediction), X_cl [steps+1][4][batch], U_cl [steps][2][batch], cost_cl [batch] (sum of
 *   x'Qx + u'Ru along the closed loop), viol_cl [batch] (max state-bound violation), n_sat (applied
 *   inputs on a bound), n_fail (QPs that did not reach MPC_SOLVED), iters_total, last_status.
 *   Optional prediction bundles (NULL = not written), the (time step x horizon x state) layout that
 *   AnimateParking.bundle consumes (session_4/animation.py:75-83): X_bundle [steps][N+1][4][batch],
 *   U_bundle [steps][N][2][batch].  ws: mpc_rti_workspace_bytes(batch, N, nc, dtype) bytes, 16-byte aligned.
 */
int mpc_bicycle_rti_prepare(double lr, double lf, double accel, double friction, double ts, int rk4,
                            const void* y, const void* U_prev, int first, void* warm_U, void* A, void* B,
                            void* c, int64_t batch, int N, int dtype, mpc_stream_t stream);
/* Obstacle-avoidance variant (session_4/main.py:29-129): as mpc_bicycle_rti_prepare, and additionally the nine
 * collision constraints between the three covering circles of the vehicle and of the parked obstacle
 * (main.py:49-56,95-104,191-200), linearised at the rolled-out states:  Cg [N][9*4][batch], hg [N][9][batch],
 * the general rows of mpc_boxqp_solve_rows.  x_obs is a HOST pointer to the obstacle pose [p_x, p_y, psi, v]. */
int mpc_bicycle_rti_prepare_obstacle(double lr, double lf, double accel, double friction, double ts, int rk4,
                                     double length, double width, const double* x_obs, const void* y,
                                     const void* U_prev, int first, void* warm_U, void* A, void* B, void* c, void* Cg,
                                     void* hg, int64_t batch, int N, int dtype, mpc_stream_t stream);
int mpc_bicycle_plant_step(double lr, double lf, double accel, double ts, const void* friction,
                           int64_t s_friction, int substeps, const void* x, const void* u, void* xn,
                           int64_t batch, int dtype, mpc_stream_t stream);
/* One globalised SQP round of the step-wise controller path (MPCController.solve / __call__ with sqp_iters > 1): after
 * mpc_bicycle_rti_prepare + mpc_boxqp_solve, backtrack on the l1 merit of the NONLINEAR OCP (cost of the nonlinear
 * rollout from y + 100 x summed state-box / collision violations) along warm_U -> U:  U <- warm_U + beta (U - warm_U),
 * beta = 1, 1/2, .., 2^-10, else 0; scenarios with status != MPC_SOLVED keep the full step.  beta optional [batch].
 * The fused loop (mpc_rti_closed_loop) runs the same rule inside its kernel. */
int mpc_bicycle_sqp_linesearch(double lr, double lf, double accel, double friction, double ts, int rk4, const void* Q,
                               const void* R, const void* Pf, const void* x_lo, const void* x_hi, int nc, double length,
                               double width, const double* x_obs, const void* y, const void* warm_U, void* U,
                               const int32_t* status, void* beta, int64_t batch, int N, int dtype, mpc_stream_t stream);
int64_t mpc_rti_workspace_bytes(int64_t batch, int N, int nc, int dtype);
int mpc_rti_closed_loop(double lr, double lf, double accel, double friction_model, double ts, int rk4,
                        double plant_lr, double plant_lf, double plant_accel, const void* friction_plant,
                        int plant_substeps, int steps, int sqp_iters, double sqp_tol, const void* Q, const void* R,
                        const void* Pf, const void* u_lo, const void* u_hi, const void* x_lo, const void* x_hi, int nc,
                        double length, double width, const double* x_obs, const void* x0, void* U_plan, void* X_pred,
                        void* X_cl, void* U_cl, void* cost_cl, void* viol_cl, void* clear_cl, int32_t* n_sat,
                        int32_t* n_fail, int32_t* iters_total, int32_t* last_status, void* X_bundle, void* U_bundle,
                        void* ws, int64_t ws_bytes, int64_t batch, int N, int max_iter, double eps, int dtype,
                        mpc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Exact stability test of a linear closed loop: rho[b] = spectral radius of A_b + B_b K_b (stable iff < 1).
 * The reference only flags |x| > 100 while simulating and leaves the exact test as an exercise
 * (session_1/session1_sol.py:86-89, :114-116).  A [*,n,n], B [*,n,m], K [*,m,n] (batch stride s?, 0 = shared),
 * rho [batch].  (n, m) in {(2,1), (4,1), (4,2)}.
 */
int mpc_spectral_radius(const void* A, int64_t sA, const void* B, int64_t sB, const void* K, int64_t sK, void* rho,
                        int64_t batch, int n, int m, int dtype, mpc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Device FP pipe probe: runs a register-resident FMA chain kernel and reports achieved
 * FLOP/s (2 flops per FMA).  Used by bench.py as the measured FP64 / FP32 vector-pipe roofline
 * denominator (MEASURED_PEAKS.json only carries HBM and bf16 tensor peaks).  Synchronises.
 */
int mpc_fma_peak_probe(int dtype, double* flops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* MPC_B200_H */
