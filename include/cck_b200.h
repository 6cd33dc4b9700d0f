/* cck_b200.h — C ABI of libcck_b200.so: the B200 (sm_100a) drop-in for the
 * belief-propagation + chance-constraint-checking hot path of CCK-CBS.
 *
 * Plain C, POD only, caller-owned buffers, int status codes, no exceptions and
 * no torch/OMPL/Eigen types cross this boundary.  Each entry point names the
 * reference interface it replaces (paths relative to the reference repo).
 * There is NO CPU fallback: every compute entry point fails with
 * CCK_ERR_NODEVICE / CCK_ERR_CUDA when no sm_100-class GPU is usable.
 *
 * Belief layout (replaces RealVectorBeliefSpace::StateType's three heap blocks,
 * include/Spaces/RealVectorBeliefSpace.h:21-96): structure of arrays.  A belief
 * batch is `W = dim + 2*dim*dim` planes of `n` doubles; plane f of belief i is
 * `base[f*ld + i]` (ld >= n, in elements).  Plane order: values[0..dim),
 * Sigma row-major, Lambda row-major (linear model: W=10; unicycle: W=36).
 * Controls are `dim` planes with their own ld.  Trajectory buffers hold one
 * belief batch per step: plane f of step t is `traj[(t*W + f)*ld + i]`.
 *
 * Host entry points take HOST pointers and do H2D / kernel / D2H on the
 * context's stream (pinned staging owned by the context).  `_dev` entry points
 * take DEVICE pointers plus a cudaStream_t (as void*; NULL = context stream),
 * launch asynchronously and never synchronise.
 */
#ifndef CCK_B200_H
#define CCK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CCK_OK 0
#define CCK_ERR_INVALID 1   /* bad argument / missing table */
#define CCK_ERR_CUDA 2      /* a CUDA call failed; see cck_last_error */
#define CCK_ERR_NOMEM 3
#define CCK_ERR_NODEVICE 4  /* no usable GPU: the library never falls back to the CPU */

#define CCK_MODEL_LINEAR2D 0 /* R2_UncertainLinearStatePropagator  (src/StatePropogators/UncertainLinearSP.cpp:58-111) */
#define CCK_MODEL_UNICYCLE 1 /* DynUnicycleControlSpace            (src/StatePropogators/UncertainUnicycleSP.cpp:146-227) */

#define CCK_SVC_BLACKMORE 0  /* PCCBlackmoreSVC          (src/StateValidityCheckers/PCCBlackmoreSVC.cpp:29-75) */
#define CCK_SVC_ADAPTIVE 1   /* AdaptiveRiskBlackmoreSVC (src/StateValidityCheckers/AdaptiveRiskBlackmoreSVC.cpp:24-92) */
#define CCK_SVC_CHISQUARED 2 /* ChiSquaredBoundarySVC    (src/StateValidityCheckers/ChiSquaredBoundarySVC.cpp:19-69) */

#define CCK_PVC_BLACKMORE 0            /* MinkowskiSumBlackmorePVC   */
#define CCK_PVC_BOUNDINGBOX 1          /* BoundingBoxBlackmorePVC    */
#define CCK_PVC_CHISQUARED 2           /* ChiSquaredBoundaryPVC      */
#define CCK_PVC_ADAPTIVE_BLACKMORE 3   /* AdaptiveRiskBlackmorePVC   */
#define CCK_PVC_ADAPTIVE_BOUNDINGBOX 4 /* AdaptiveRiskBoundingBoxPVC */
#define CCK_PVC_BLACKMORE2 5           /* Blackmore2PVC  (src/PlanValidityCheckers/Blackmore2PVC.cpp:169-192) */
#define CCK_PVC_CDFGRID 6              /* CDFGridPVC(n)  (src/PlanValidityCheckers/CDFGridPVC.cpp:182-291; `--pvc CDFGrid-n`) */

typedef struct cck_ctx cck_ctx;

/* Model constants: the fields of the reference's propagator constructors
 * (UncertainLinearSP.cpp:4-45, UncertainUnicycleSP.cpp:11-130) as one POD. */
typedef struct cck_model_params {
    int32_t model;        /* CCK_MODEL_* */
    int32_t dim;          /* 2 or 4 */
    double dt;            /* si->getPropagationStepSize() (OmplSetUp.cpp:186,224) */
    double lin_Q[4], lin_R[4], lin_Acl[4];
    double uni_F[16], uni_Acl[16], uni_Q[16], uni_R[16];
    double uni_gains[8];  /* controller_parameters_ */
    double uni_acc[2];    /* forward_acceleration_bounds_ */
    double uni_turn[2];   /* turning_rate_bounds_ */
} cck_model_params;

/* State-validity-checker constants (what the reference SVC constructors compute once). */
typedef struct cck_svc_params {
    int32_t kind;          /* CCK_SVC_* */
    int32_t dim;           /* state dimension whose bounds are checked */
    double low[4], high[4];/* RealVectorBounds (OmplSetUp.cpp:195-200, 275-289) */
    double erf_inv_result; /* Blackmore: erf_inv(1 - 2*p_coll/M)  (PCCBlackmoreSVC.cpp:10) */
    double p_collision;    /* Adaptive:  1 - accep_prob           (AdaptiveRiskBlackmoreSVC.cpp:6) */
    double sc;             /* ChiSquared: chi2_2 quantile         (ChiSquaredBoundarySVC.cpp:7) */
    double max_rad;        /* ChiSquared: robot bounding radius   (ChiSquaredBoundarySVC.cpp:4) */
} cck_svc_params;

/* Plan-validity-checker constants (MinkowskiSumBlackmorePVC.cpp:4-30 and variants). */
typedef struct cck_pvc_params {
    int32_t kind;          /* CCK_PVC_* */
    int32_t k;             /* number of robots (<= CCK_MAX_ROBOTS) */
    int32_t dim;           /* 2 or 4: selects getDistribution_'s sub-block (BeliefPVC.cpp:84-113) */
    int32_t grid_n;        /* CDFGridPVC: disk_steps_ (grid points per axis, >= 2); ignored by the other kinds */
    double c_pair;         /* erf_inv(1 - 2*p_coll_dist) (MinkowskiSumBlackmorePVC.cpp:178) */
    double p_coll_agnts;   /* Adaptive PVCs: 1 - p_safe_agnts (BeliefPVC.cpp:7) */
    double sc;             /* ChiSquared PVC: chi2_2 quantile (ChiSquaredBoundaryPVC.cpp:19) */
    const double* radii;      /* [k] bounding radii */
    const double* pair_A;     /* [k*k][4][2] half-plane normals, entry [a*k+b] valid for a<b (Blackmore2: Minkowski sum of the
                                 bounding squares; CDFGrid does not read them: its box is +-(r_a + r_b)) */
    const double* pair_B;     /* [k*k][4] */
    const int32_t* map_order; /* [k] robot indices in std::map<std::string,...> name order (quirk 6) */
} cck_pvc_params;
#define CCK_MAX_ROBOTS 64

/* One conflict run = the reference's std::vector<ConflictPtr> compressed:
 * Conflict{a,b,first_step}, ..., Conflict{a,b,first_step+run_len-1}. */
typedef struct cck_conflict_run {
    int32_t a, b;          /* Conflict::agent1Idx_/agent2Idx_ (include/utils/Conflict.h:6-13) */
    int32_t first_step;    /* -1 when the plan is conflict free */
    int32_t run_len;
} cck_conflict_run;

/* ---- context ------------------------------------------------------------ */
int cck_ctx_create(int device, cck_ctx** out);
int cck_ctx_destroy(cck_ctx* ctx);
const char* cck_last_error(const cck_ctx* ctx);   /* valid until the next call on ctx; ctx==NULL: last create error */
int cck_ctx_synchronize(cck_ctx* ctx);
int cck_num_sms(const cck_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t cck_launch_count(const cck_ctx* ctx);

/* ---- constants ---------------------------------------------------------- */
/* The reference constructors' constants for `model` at step size dt. */
int cck_model_params_default(int model, double dt, cck_model_params* out);
int cck_set_model(cck_ctx* ctx, const cck_model_params* p);
/* Replaces the tables PCCBlackmoreSVC / AdaptiveRiskBlackmoreSVC / ChiSquaredBoundarySVC
 * build in their constructors: A_list_ [M][4][2], B_list_ [M][4], obstacle centres [M][2]
 * (Adaptive; may be NULL otherwise), inflated rectangles [M][4] = xlo,ylo,xhi,yhi
 * (ChiSquared; may be NULL otherwise).  M may be 0. */
int cck_set_obstacles(cck_ctx* ctx, const cck_svc_params* p, const double* A, const double* B,
                      const double* centres, const double* rects, int M);
int cck_set_pvc(cck_ctx* ctx, const cck_pvc_params* p);

/* Host-side builders restating the reference's constant arithmetic (no GPU needed):
 * Boost erf_inv / chi2 quantile call sites, Instance risk split, Minkowski half-planes. */
double cck_erf_inv(double z);
double cck_chi2_quantile_2dof(double p);
/* Instance.cpp:38-44 */
void cck_split_risk(double p_safe, int n_obstacles, int n_robots, double* p_safe_obs, double* p_safe_agents);
/* RectangularRobot + createBoundingShape (common.cpp:117-136, 61-108) */
double cck_rect_robot_bounding_radius(double sx, double sy, double len, double width);
/* Unit-cell style rectangular obstacles (common.cpp:25-42) inflated by the robot's bounding
 * square (PCCBlackmoreSVC.cpp:110-145) and converted to half-planes (:77-108). */
int cck_build_obstacle_tables(int M, const double* centres, double len, double width, double robot_rad,
                              double* A, double* B, double* rects);
/* Robot-pair half-planes: Minkowski sum of the two bounding squares (MinkowskiSumBlackmorePVC.cpp:214-248)
 * or the bounding box of r1+r2 (BoundingBoxBlackmorePVC.cpp:174-195). */
int cck_build_pair_halfplanes(int bounding_box, double r1, double r2, double* A8, double* B4);

/* ---- StatePropagator::propagate, batched --------------------------------
 * replaces R2_UncertainLinearStatePropagator::propagate / DynUnicycleControlSpace::propagate
 * for n independent (state, control) pairs; `duration` as in the reference signature
 * (the linear model ignores it, quirk 1). */
int cck_propagate(cck_ctx* ctx, int64_t n, const double* bel, int64_t ld, const double* ctrl, int64_t ldc,
                  double duration, double* out, int64_t ldo);
int cck_propagate_dev(cck_ctx* ctx, void* stream, int64_t n, const double* bel, int64_t ld, const double* ctrl,
                      int64_t ldc, double duration, double* out, int64_t ldo);

/* ---- StateValidityChecker::isValid, batched ----------------------------- */
int cck_is_valid(cck_ctx* ctx, int64_t n, const double* bel, int64_t ld, uint8_t* valid);
int cck_is_valid_dev(cck_ctx* ctx, void* stream, int64_t n, const double* bel, int64_t ld, uint8_t* valid);

/* ---- SpaceInformation::propagateWhileValid, batched (the fused hot kernel) -
 * replaces the OMPL loop called at ConstraintRespectingBSST.cpp:248: for each i,
 * hold ctrl[i] for steps[i] steps of size dt, checking each new state with the
 * installed SVC.  valid_steps[i] = number of leading valid steps (= index of the
 * first violating step, or steps[i]); out = last valid state.  traj (optional,
 * NULL to skip) receives every valid intermediate state, traj_T >= max(steps). */
int cck_propagate_while_valid(cck_ctx* ctx, int64_t n, const double* bel, int64_t ld, const double* ctrl,
                              int64_t ldc, const int32_t* steps, double* out, int64_t ldo, int32_t* valid_steps,
                              double* traj, int64_t traj_ld, int32_t traj_T);
int cck_propagate_while_valid_dev(cck_ctx* ctx, void* stream, int64_t n, const double* bel, int64_t ld,
                                  const double* ctrl, int64_t ldc, const int32_t* steps, double* out, int64_t ldo,
                                  int32_t* valid_steps, double* traj, int64_t traj_ld, int32_t traj_T);

/* ---- CentralizedChiSquaredBoundarySVC's agent-agent test (src/StateValidityCheckers/CentralizedChiSquaredBoundarySVC.cpp:70-99) ----
 * for n pairs: 10-gon of radius lambda_max(cov_a) * sc + rad_a around mu_a vs the one of agent b; out[i] = 1 iff disjoint
 * (touching is not).  SoA host arrays: mu [2][n], cov [4][n] (row-major 2x2), rad [n].  (The agent-obstacle half of that
 * checker is cck_is_valid with CCK_SVC_CHISQUARED on the agents' 2-D blocks; the centralized propagator,
 * CentralizedUncertainLinearSP.cpp:42-102, is cck_propagate over the agents' 2-D blocks - same constants as the
 * single-agent linear model.) */
int cck_chisq_pairs_disjoint(cck_ctx* ctx, int64_t n, const double* mu_a, const double* cov_a, const double* rad_a, const double* mu_b,
                             const double* cov_b, const double* rad_b, double sc, uint8_t* out);

/* Host-only test hook (no GPU): the pair test (*PVC::isSafe_) of plan-validity-checker `kind` evaluated on the host by the same
 * source functions the validatePlan / satisfiesConstraints kernels call; lets CPU-only tests compare the kernels' logic with
 * the oracle.  1 safe, 0 unsafe, -1 unknown kind. */
int cck_host_pair_safe(int kind, const double* mu_a, const double* cov_a, const double* mu_b, const double* cov_b, const double* A8,
                       const double* B4, double rad_a, double rad_b, double c_pair, double p_coll_agnts, double sc, int grid_n);

int cck_host_chisq_pair_disjoint(const double* mu_a, const double* cov_a, double rad_a, const double* mu_b, const double* cov_b, double rad_b,
                                 double sc);   /* same, for cck_chisq_pairs_disjoint's kernel source (one pair, row-major cov) */

/* ---- PlanValidityChecker::validatePlan ----------------------------------
 * replaces {MinkowskiSumBlackmore,BoundingBoxBlackmore,ChiSquaredBoundary,AdaptiveRisk*}PVC::validatePlan
 * on already-interpolated trajectories (one belief per dt).  traj: robot r, plane f,
 * step t at traj[(r*W + f)*ld + t]; lens[r] = number of states of robot r. */
int cck_validate_plan(cck_ctx* ctx, int k, const int32_t* lens, const double* traj, int64_t ld, cck_conflict_run* out);
int cck_validate_plan_dev(cck_ctx* ctx, void* stream, int k, const int32_t* lens_dev, int max_len, const double* traj,
                          int64_t ld, cck_conflict_run* out_dev);
/* Sharded conflict search (the pair x step tests are independent up to the final "earliest (step, pair order)" argmin;
 * one rank scans only the steps [t_begin, t_end) of the replicated plan; t_end < 0 = to the end).  out = the first conflict
 * INSIDE the range with the follow-up run over the WHOLE plan (a = -1 when the range is conflict free); key_out (optional) =
 * first_step << 32 | rank of the pair in std::map order, all ones when there is none: the record with the smallest key over
 * the ranks is what validatePlan returns (sharding.gather_conflict all-gathers {key, a, b, first_step, run_len}). */
int cck_validate_plan_range(cck_ctx* ctx, int k, const int32_t* lens, const double* traj, int64_t ld, int t_begin, int t_end,
                            cck_conflict_run* out, unsigned long long* key_out);
int cck_validate_plan_range_dev(cck_ctx* ctx, void* stream, int k, const int32_t* lens_dev, int max_len, int t_begin, int t_end,
                                const double* traj, int64_t ld, cck_conflict_run* out_dev, unsigned long long* key_out_dev);

/* ---- PlanValidityChecker::satisfiesConstraints, batched ------------------
 * Constraints (BeliefConstraint, include/Constraints/BeliefConstraint.h:8-16) are
 * flattened into entries: entry e applies at absolute step ent_step[e] and carries the
 * constraining robot's belief as (mu[2], cov[4]) = getDistribution_ of the stored state,
 * plus pair_ab[e] = constrained*k + constraining (the planning robot first; the half-planes of the
 * unordered pair are used, the bounding radii pair up with the two beliefs as in
 * ChiSquaredBoundaryPVC.cpp:81).  Every PVC kind checks its own isSafe_ at the constraint times,
 * ChiSquaredBoundaryPVC included (ChiSquaredBoundaryPVC.cpp:54-87).
 * Paths: n_paths candidate paths in one SoA trajectory buffer (plane f of step t of
 * path i at path[(t*W+f)*ld + i]); path i has path_len[i] states and its first state
 * sits at absolute step path_step0[i].  ok[i] = 1 iff every entry whose step falls on
 * a state of path i is safe. */
int cck_set_constraints(cck_ctx* ctx, int n_entries, const int32_t* ent_step, const int32_t* pair_ab,
                        const double* mu2, const double* cov4);
int cck_satisfies_constraints(cck_ctx* ctx, int64_t n_paths, const double* path, int64_t ld, int32_t max_len,
                              const int32_t* path_len, const int32_t* path_step0, uint8_t* ok);
int cck_satisfies_constraints_dev(cck_ctx* ctx, void* stream, int64_t n_paths, const double* path, int64_t ld,
                                  int32_t max_len, const int32_t* path_len, const int32_t* path_step0, uint8_t* ok);

/* ---- batched low-level expansion (ConstraintRespectingBSST.cpp:245-326 for K candidates) --
 * For candidate i: parent belief, control, steps, parent's absolute step index and
 * accumulated cost.  One launch does propagateWhileValid, the constraint check on the new
 * segment (entries set by cck_set_constraints), the EuclideanPathLengthObjective increment
 * and ChanceConstrainedGoal::isSatisfied.  flags bit0: propCd==cd, bit1: constraints ok,
 * bit2: goal satisfied, bit3 (diagnostic): the covariances came from the tree's covariance table
 * (cck_set_cov_table) instead of the Kalman step - same bits either way. */
typedef struct cck_goal_params {
    double gx, gy;      /* goal location */
    double threshold;   /* goalTollorance = 2.0 (OmplSetUp.cpp:185) */
    double sc;          /* chi2_2 quantile of 0.95 (BeliefSpaceGoals.cpp:110) */
} cck_goal_params;
/* Optional: the covariance table of ONE low-level tree.  The covariance recursion of both models does not depend on the mean
 * or the control (UncertainLinearSP.cpp:98-104, UncertainUnicycleSP.cpp:210-220), so every tree node at absolute step s
 * carries the same (Sigma_s, Lambda_s), determined by the root covariance root_cov = [Sigma row-major, Lambda row-major]
 * (the start state's 0.01 I, OmplSetUp.cpp:232-236).  With a table installed, cck_expand_batch reads a candidate's
 * covariances from it instead of running the Kalman step - but ONLY for a candidate whose parent covariance is bit-identical
 * to the table's entry for parent_step and whose segment ends inside the table; every other candidate is computed as usual,
 * so results never depend on the table.  Entries are produced by the kernels' own step functions.  n_steps <= 0 removes it;
 * cck_set_model removes it too. */
int cck_set_cov_table(cck_ctx* ctx, const double* root_cov, int n_steps);
int cck_expand_batch_dev(cck_ctx* ctx, void* stream, int64_t n, const double* bel, int64_t ld, const double* ctrl,
                         int64_t ldc, const int32_t* steps, const int32_t* parent_step, const double* parent_cost,
                         const cck_goal_params* goal, double* out, int64_t ldo, int32_t* valid_steps,
                         double* cost, double* goal_dist, uint8_t* flags);
int cck_expand_batch(cck_ctx* ctx, int64_t n, const double* bel, int64_t ld, const double* ctrl, int64_t ldc,
                     const int32_t* steps, const int32_t* parent_step, const double* parent_cost,
                     const cck_goal_params* goal, double* out, int64_t ldo, int32_t* valid_steps, double* cost,
                     double* goal_dist, uint8_t* flags);

/* ---- result summary (what a rank contributes to the NCCL gather) ---------------------
 * out_dev[0] = number of candidates whose whole segment was valid (valid_steps == steps),
 * out_dev[1] = executed belief-steps = sum_i min(valid_steps[i]+1, steps[i]),
 * out_dev[2] = sum_i valid_steps[i].  One warp-shuffle reduction kernel. */
int cck_rollout_summary_dev(cck_ctx* ctx, void* stream, int64_t n, const int32_t* steps, const int32_t* valid_steps,
                            uint64_t* out_dev);

/* Winner of a batched expansion (the candidate the planner would commit first by cost):
 * among candidates with (flags & required_mask) == required_mask the one with the smallest
 * cost, ties to the smallest index.  out_dev[0] = cost (+inf if none), out_dev[1] = global
 * index (index_offset + i, as a double; -1 if none).  One per-rank record to all-gather. */
int cck_select_winner_dev(cck_ctx* ctx, void* stream, int64_t n, const double* cost, const uint8_t* flags,
                          int required_mask, int64_t index_offset, double* out_dev);

/* ---- measurement helper ---------------------------------------------------
 * FP64 pipe peak (DFMA chains, all SMs) in TFLOP/s measured on the context's device;
 * the roofline denominator for the obstacle-check kernels (MEASURED_PEAKS.json has no
 * FP64 entry). */
int cck_measure_fp64_peak(cck_ctx* ctx, double* tflops);
/* out[0..2]: NON-fused FP64 instruction rates (1e12 instructions/s per lane): alternating DMUL/DADD, DADD only, DMUL only.
 * The decision arithmetic may not be contracted into FMAs, so these are the attainable rates of the path's FP64 work.
 * out[3..5]: the FP64 rate of the alternating stream when 1, 2, 3 independent fp32 instructions are issued per FP64
 * instruction (is the issue port shared?  it tells how much fp32/int work the FP64 pipe hides). */
int cck_measure_fp64_rates(cck_ctx* ctx, double* out6);

/* Host-only (no GPU, no context): the serialised fp32 box + cell-grid table that cck_set_obstacles builds for the
 * PCCBlackmoreSVC broad phase (PCCBlackmoreSVC.cpp:54-75 is what it must never contradict).  Exposed so that the
 * table's invariants — every obstacle binned into its own cell, boxes rounded outward / inward by at least one ulp —
 * can be tested without a device.  `out` may be NULL to query `*bytes`; order[j] = obstacle stored in row j. */
int cck_build_grid_table(int M, const double* A, const double* B, double erf_inv_result, unsigned char* out, int64_t cap,
                         int64_t* bytes, int32_t* order);

/* Host-only: the clearance field built next to the grid table - CCK_DF_N x CCK_DF_N (128 x 128) 32-bit words over the fp32
 * state bounds: bits 0-7 floor(16 x a lower bound of the Chebyshev clearance between any mean mapped to the cell and the
 * NEAREST obstacle's outer box), bits 8-15 the same for the second nearest box, bits 16-31 the grid-table row of the nearest
 * one (order[] of cck_build_grid_table).  It is the FIRST test of a belief-step: one load + one compare proves "safe w.r.t.
 * every obstacle" for a belief whose sqrt2 c sqrt(P) interval is smaller than the clearance, a second compare reduces the
 * decision to the nearest obstacle alone; everything else goes on to the cell-grid broad phase and the exact decision.  params[0..2] = cells per unit in x, in y, and the factor that turns a
 * variance into the squared 1/16-unit interval half-width.  *has = 0: no field (M = 0, an obstacle without an axis-aligned
 * box, unusable bounds).  For CPU tests of the invariant. */
int cck_build_clearance_field(int M, const double* A, const double* B, double erf_inv_result, const double* low, const double* high,
                              uint32_t* out, float* params, int* has);

/* ---- belief-space nearest neighbours over a device-resident tree ----
 * RealVectorBeliefSpace::distance (src/Spaces/RealVectorBeliefSpace.cpp:16-62):
 *   | ||mu1-mu2||^2 + trace(C1 + C2 - 2 (C2^1/2 C1 C2^1/2)^1/2) |, 0 when the means coincide, C = Sigma + Lambda.
 * A cck_tree holds tree nodes (or witnesses): belief record, accumulated cost, active flag.
 *   cck_tree_select   replaces ConstraintRespectingBSST::selectNode (:119-148) for K samples at once:
 *                     best-cost ACTIVE node within selection_radius of the sample, else the nearest active node.
 *   cck_tree_nearest  replaces the nearest-witness query of findClosestWitness (:150-178): argmin distance
 *                     over all items (ties -> lowest index), the caller compares out_dist with the pruning radius.
 * Belief records are SoA [W][ld] host arrays as everywhere else; indices are item positions in append order. */
typedef struct cck_tree cck_tree;
int cck_tree_create(cck_ctx* ctx, int dim, cck_tree** out);
int cck_tree_destroy(cck_tree* tree);
int cck_tree_clear(cck_tree* tree);
int64_t cck_tree_size(const cck_tree* tree);
int cck_tree_append(cck_tree* tree, int64_t n, const double* bel, int64_t ld, const double* cost, int64_t* first_index);
int cck_tree_update(cck_tree* tree, int64_t n, const int32_t* idx, const double* cost /* or NULL */, const uint8_t* active /* or NULL */);
int cck_tree_select(cck_tree* tree, int64_t K, const double* samples, int64_t lds, double selection_radius, int32_t* out_idx,
                    double* out_dist /* or NULL */);
int cck_tree_nearest(cck_tree* tree, int64_t K, const double* queries, int64_t ldq, int32_t* out_idx, double* out_dist /* or NULL */);
/* The witness rules of one launch (ConstraintRespectingBSST.cpp:150-178, :250-326) for n <= 1024 survivors in sample order,
 * against `tree` = the existing witnesses.  out_widx / out_wdist: nearest existing witness (findClosestWitness; -1 if none).
 * A survivor with no existing witness within the pruning radius, and eligible[s] != 0 (NULL = all: not rejected by the
 * constraints), becomes a new witness unless an earlier NEW witness of the same launch is within the radius: out_fresh.
 * out_near / out_near_dist: the nearest new witness created by an EARLIER survivor whose distance is below
 * min(out_wdist, radius) (first one on ties; -1 if none) - what the sequential scan would have found.  One upload, five
 * kernels, one download. */
int cck_tree_witness_batch(cck_tree* tree, int64_t n, const double* queries, int64_t ldq, double pruning_radius, const uint8_t* eligible /* or NULL */,
                           int32_t* out_widx, double* out_wdist, int32_t* out_near, double* out_near_dist, uint8_t* out_fresh);
/* distance(a_i, b_i) for n pairs (RealVectorBeliefSpace::distance(state1 = a, state2 = b)) */
int cck_belief_distance(cck_ctx* ctx, int dim, int64_t n, const double* a, int64_t lda, const double* b, int64_t ldb, double* out);
/* out[q*n + i] = distance(b_i, b_q) for i < q (n <= 8192; other entries unspecified): the witness tests between
 * candidates committed from the same launch, so the sequential witness rules can be replayed on the host exactly */
int cck_belief_distance_lower(cck_ctx* ctx, int dim, int64_t n, const double* bel, int64_t ld, double* out);
/* the same matrix restricted to nf columns: out[q*nf + a] = distance(b_cols[a], b_q) for cols[a] < q.  Only candidates farther
 * than the pruning radius from every existing witness can become witnesses, so the planner asks for those columns only */
int cck_belief_distance_cols(cck_ctx* ctx, int dim, int64_t n, const double* bel, int64_t ld, int32_t nf, const int32_t* cols, double* out);
/* the part of that matrix the sequential witness rules (ConstraintRespectingBSST.cpp:150-178, :250-326) can use: for row q
 * the pairs (a, distance(b_cols[a], b_q)) with cols[a] < q and distance < limit[q], in increasing a.  out_count[q] is the
 * full number of such pairs; the first min(out_count[q], cap) are stored at out_idx / out_dist [q*cap ...].  cols must be
 * ascending.  A candidate only ever adopts a same-launch witness that is nearer than its nearest existing witness and
 * within the pruning radius, so limit[q] = min(that distance, nextafter(pruning radius)) makes the rows nearly empty. */
int cck_belief_distance_cols_near(cck_ctx* ctx, int dim, int64_t n, const double* bel, int64_t ld, int32_t nf, const int32_t* cols,
                                  const double* limit, int32_t cap, int32_t* out_idx, double* out_dist, int32_t* out_count);


#ifdef __cplusplus
}
#endif
#endif /* CCK_B200_H */
