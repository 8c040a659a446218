/*
 * ppo_ba.h — C-ABI of the B200-native local bundle-adjustment engine for the mixed
 * point / plane / cuboid factor graph of benchun123/point-plane-object-SLAM.
 *
 * The reference has no FFI layer: its boundary is the C++ static-method API
 *   Optimizer::LocalBundleAdjustment      (include/Optimizer.h:45,  src/Optimizer.cc:461-786)
 *   Optimizer::LocalBACameraPlaneCuboids  (include/Optimizer.h:62,  src/Optimizer.cc:1994-2967)
 * which build a g2o graph, call SparseOptimizer::optimize() twice and write the map back.
 * This header is what a replacement Optimizer.cc binds instead of g2o (INTEGRATION.md shows
 * the stub).  Each entry point cites the reference interface it replaces.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; the caller owns every input
 * array (copied at ppo_ba_set_graph); outputs are copied into caller buffers; all functions
 * return PPO_OK (0) or a negative error code; a handle is single-caller (one CUDA stream).
 * All solver arithmetic is IEEE double ("f64"), as in g2o; observations/intrinsics arrive as
 * float32 because that is what KeyFrame stores (KeyFrame.h:184,191-192,209).
 */
#ifndef PPO_BA_H
#define PPO_BA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PPO_OK 0
#define PPO_E_INVALID (-1)  /* bad argument / inconsistent graph                              */
#define PPO_E_CUDA (-2)     /* CUDA runtime error (ppo_ba_last_error gives the string)        */
#define PPO_E_NCCL (-3)     /* collective error in the multi-GPU variants                     */
#define PPO_E_NOGPU (-4)    /* no CUDA device: the engine has no CPU fallback by design       */
#define PPO_E_EMPTY (-5)    /* no active vertex (g2o: "0 vertices to optimize",
                               core/sparse_optimizer.cpp:356-359)                             */

/* Edge families, in the order Optimizer.cc adds them. */
enum ppo_edge_kind {
  PPO_EDGE_POINT = 0,        /* EdgeSE3ProjectXYZ / EdgeStereoSE3ProjectXYZ  (Optimizer.cc:2357-2424) */
  PPO_EDGE_PLANE = 1,        /* EdgePlane / EdgeVerticalPlane / EdgeParallelPlane (:2222-2309)        */
  PPO_EDGE_CUBOID_CAM = 2,   /* EdgeSE3CuboidProj / EdgeSE3CuboidCornerProj  (:2433-2551)             */
  PPO_EDGE_POINT_CUBOID = 3, /* EdgePointCuboidOnlyObject                    (:2556-2655)             */
  PPO_EDGE_CUBOID_PLANE = 4, /* EdgeCuboidPlane (constant residual, G2O_Plane3D.h:470-473)            */
  PPO_EDGE_KINDS = 5
};

/* ple_kind values */
#define PPO_PLANE_OBS 0 /* EdgePlane         3-D, Plane3D::ominus     */
#define PPO_PLANE_VER 1 /* EdgeVerticalPlane 2-D, Plane3D::ominus_ver */
#define PPO_PLANE_PAR 2 /* EdgeParallelPlane 2-D, Plane3D::ominus_par */
/* cbe_kind values */
#define PPO_CUBOID_BBOX 0   /* EdgeSE3CuboidProj       4-D  */
#define PPO_CUBOID_CORNER 1 /* EdgeSE3CuboidCornerProj 16-D */
#define PPO_CUBOID_SE3 2    /* EdgeSE3Cuboid 9-D: min_log_error of the measured cuboid moved to the world frame, over the four yaw
                             * rotations of the ambiguous front face (g2o_cuboid.h:72-109,322-340); only added by
                             * LocalBACameraPointCuboids2D (Optimizer.cc:1764-1804).  cbe_meas = the measured cuboid in the CAMERA
                             * frame, laid out like cu_state: [tx ty tz qx qy qz qw sx sy sz] (+ 6 unused) */
/* cu_flags bits (VertexCuboid, g2o_cuboid.h:281-285) */
#define PPO_CU_FIXROLLPITCH 1u
#define PPO_CU_FIXHEIGHT 2u
/* per-edge flag bits (ppo_ba_set_levels / ppo_ba_get_edge_flags) */
#define PPO_EF_LEVEL1 1u /* edge->setLevel(1): inactive in optimize() of level 0      */
#define PPO_EF_ROBUST 2u /* a RobustKernelHuber is attached                            */

/* Solver flavour: which g2o stack the call mirrors.  Both factorise the Schur-reduced pose system with the same
 * dense tile Cholesky on the GPU.  They differ where the reference's linear solvers differ: LinearSolverDense
 * (Eigen::LDLT + isPositive(), solvers/linear_solver_dense.h:65-113) reports a failed solve on a system that is not
 * positive definite; LinearSolverEigen (Eigen::SimplicialLDLT, solvers/linear_solver_eigen.h:94-124) is an LDL^T
 * without pivoting that fails only on a zero pivot -- for PPO_SOLVER_6_3 a failed Cholesky is followed by such an
 * LDL^T on a copy of the system. */
#define PPO_SOLVER_DENSE_X 0 /* BlockSolverX + LinearSolverDense   (Optimizer.cc:2108-2113) */
#define PPO_SOLVER_6_3 1     /* BlockSolver_6_3 + LinearSolverEigen (Optimizer.cc:516-522)   */

/* BA-wide configuration: the globals of include/Parameters.h:45-76 and the literals inside
 * Optimizer.cc, snapshotted per call by the host shim.  ppo_ba_default_params() fills the
 * reference defaults (Parameters.cc:58-74; Optimizer.cc:2194-2204,2326-2327). */
typedef struct ppo_ba_params {
  /* Huber deltas exactly as the reference forms them: (double)(float)sqrt(threshold) */
  double huber_mono;         /* sqrt(5.991)              Optimizer.cc:2326 */
  double huber_stereo;       /* sqrt(7.815)              :2327 */
  double huber_plane;        /* sqrt(plane_chi=500)      :2202-2203 */
  double huber_vp_plane;     /* sqrt(200)                :2204-2205 */
  double huber_bbox;         /* sqrt(thHuberBbox2d=80)   :2470 */
  double huber_corner;       /* sqrt(thHuberConer2d=10)  :2536 */
  double huber_cuboid_plane; /* sqrt(cuboid_plane_chi=500) :2671 */
  /* outlier thresholds of the re-levelling pass (Optimizer.cc:2736-2833), used by
   * ppo_ba_outlier_pass / ppo_ba_local_ba */
  double chi2_mono;        /* 5.991 */
  double chi2_stereo;      /* 7.815 */
  double chi2_plane;       /* plane_chi      */
  double chi2_vp_plane;    /* 200            */
  double norm_bbox;        /* thHuberBbox2d: compared with ||error|| (:2774) */
  double norm_corner;      /* thHuberConer2d (:2783) */
  /* Levenberg-Marquardt constants (optimization_algorithm_levenberg.cpp:42-55) */
  double lm_tau;           /* 1e-5 */
  double lm_good_upper;    /* 2/3  */
  double lm_good_lower;    /* 1/3  */
  int32_t lm_max_trials;   /* 10   */
  int32_t solver;          /* PPO_SOLVER_* */
  int32_t iters_round1;    /* 5  (Optimizer.cc:2728) */
  int32_t iters_round2;    /* 10 (Optimizer.cc:2837) */
  /* point-cuboid edge constants (Optimizer.cc:2647, g2o_cuboid.cc:147) */
  double ptcu_max_outside_margin_ratio; /* 1.0 */
  double ptcu_prior_weight;             /* 0.2 */
  /* EdgeSE3Cuboid (PPO_CUBOID_SE3, Optimizer.cc:1792-1794,1875-1882; Parameters.cc:65) */
  double huber_se3;                     /* thHuberSE3 = 900 (set as the delta itself, no square root) */
  double norm_se3;                      /* thHuberSE3: compared with ||error|| in the outlier pass      */
} ppo_ba_params;

/* Flat SoA factor graph. Index spaces are typed (KF slot, point, plane, cuboid) instead of
 * g2o's single colliding id space. KF slots must be sorted by KeyFrame::mnId
 * (g2o orders the Hessian by vertex id, core/sparse_optimizer.cpp:166-190,482-487); cuboids
 * follow the KFs in the pose block; planes and points are the marginalised landmark block. */
typedef struct ppo_ba_graph {
  /* -- vertices -------------------------------------------------------------------------- */
  int32_t n_kf;            /* local + fixed KeyFrames                                          */
  const double *kf_pose;   /* n_kf x 7  Tcw as [qx qy qz qw tx ty tz] (SE3Quat, se3quat.h:46-47) */
  const uint8_t *kf_fixed; /* n_kf      1 = setFixed(true)  (Optimizer.cc:2126-2128,2141)       */
  const float *kf_intr;    /* n_kf x 5  fx fy cx cy mbf  (KeyFrame.h:184)                       */
  int32_t n_pt;
  const double *pt_xyz;    /* n_pt x 3  (VertexSBAPointXYZ)                                     */
  const uint8_t *pt_fixed; /* n_pt or NULL: fixPoint (Optimizer.cc:2343-2346)                   */
  int32_t n_pl;
  const double *pl_coef;   /* n_pl x 4  Plane3D coeffs (normalised on entry like fromVector)    */
  int32_t n_cu;
  const double *cu_state;  /* n_cu x 10 [tx ty tz qx qy qz qw sx sy sz] (cuboid::toVector)      */
  const uint8_t *cu_flags; /* n_cu      PPO_CU_* bits                                           */
  /* -- point edges, CSR by point (rows = points) ----------------------------------------- */
  const int32_t *pt_rowptr; /* n_pt + 1 */
  int32_t n_pe;
  const int32_t *pe_kf;      /* n_pe   KF slot                                                  */
  const float *pe_obs;       /* n_pe x 3  u v u_right; u_right < 0 => monocular edge            */
  const float *pe_invsigma2; /* n_pe   mvInvLevelSigma2[octave]                                 */
  /* -- plane edges ------------------------------------------------------------------------ */
  int32_t n_ple;
  const int32_t *ple_plane;
  const int32_t *ple_kf;
  const uint8_t *ple_kind;   /* PPO_PLANE_*                                                     */
  const double *ple_meas;    /* n_ple x 4  Converter::toPlane3D(kf->mvPlaneCoefficients[idx])   */
  const double *ple_info;    /* n_ple x 3  diagonal of the information matrix (3rd unused for 2-D) */
  /* -- camera-cuboid edges ---------------------------------------------------------------- */
  int32_t n_cbe;
  const int32_t *cbe_kf;
  const int32_t *cbe_cuboid;
  const uint8_t *cbe_kind;   /* PPO_CUBOID_*                                                    */
  const double *cbe_meas;    /* n_cbe x 16 (bbox uses the first 4: cx cy w h)                   */
  const double *cbe_info;    /* n_cbe   (weight*meas_quality)^2, scalar times identity          */
  /* -- point-cuboid unary edges ----------------------------------------------------------- */
  int32_t n_pce;
  const int32_t *pce_cuboid;
  const int32_t *pce_rowptr; /* n_pce + 1 into pce_pts                                          */
  const double *pce_pts;     /* x 3  constant world points captured at graph build              */
  /* -- cuboid-plane edges (constant residual) --------------------------------------------- */
  int32_t n_cpe;
  const int32_t *cpe_cuboid;
  const int32_t *cpe_plane;
  const double *cpe_meas;    /* n_cpe x 3 */
  const double *cpe_info;    /* n_cpe x 3 diagonal */
} ppo_ba_graph;

/* Mutable state returned by ppo_ba_get_state (same layouts as the graph's vertex arrays). */
typedef struct ppo_ba_state {
  double *kf_pose;  /* n_kf x 7  */
  double *pt_xyz;   /* n_pt x 3  */
  double *pl_coef;  /* n_pl x 4  */
  double *cu_state; /* n_cu x 10 */
} ppo_ba_state;

#define PPO_TRACE_MAX 64
/* One record per outer LM iteration (OptimizationAlgorithmLevenberg::solve). */
typedef struct ppo_ba_iter {
  double chi2_before; /* currentChi at entry                        */
  double chi2_after;  /* currentChi after the accepted trial (or unchanged) */
  double lambda;      /* _currentLambda after the iteration          */
  double rho;         /* last rho                                    */
  int32_t trials;     /* qmax                                        */
  int32_t accepted;   /* 1 if the last trial was accepted            */
} ppo_ba_iter;

typedef struct ppo_ba_stats {
  int32_t iterations;       /* outer iterations executed (return of optimize())             */
  int32_t terminated;       /* 1: LM returned Terminate; 2: stop flag seen                  */
  int32_t n_pose_dim;       /* scalar size of the reduced system                             */
  int32_t n_landmarks;      /* active marginalised landmarks                                 */
  int32_t n_active_edges;
  int32_t total_trials;
  double chi2_initial, chi2_final;
  double ms_total;          /* device time of the call, CUDA events                          */
  double ms_linearize, ms_schur, ms_solve, ms_update; /* filled when profiling is enabled     */
  ppo_ba_iter trace[PPO_TRACE_MAX];
} ppo_ba_stats;

/* Result of the fused whole-call schedule. */
typedef struct ppo_ba_result {
  ppo_ba_stats round1, round2;
  int32_t n_outlier_point_edges; /* edges levelled out between the rounds                    */
  int32_t n_outlier_plane_edges;
  int32_t n_outlier_cuboid_edges;
  int32_t skipped;               /* 1: stop flag was set on entry -> nothing done (:2723-2725) */
} ppo_ba_result;

typedef struct ppo_ba_handle ppo_ba_handle;

/* Reference defaults. */
void ppo_ba_default_params(ppo_ba_params *p);

/* Replaces the construction of SparseOptimizer + BlockSolver + LinearSolver + Levenberg
 * (Optimizer.cc:2108-2116 / :516-522). device = CUDA ordinal. */
int ppo_ba_create(const ppo_ba_params *params, int device, ppo_ba_handle **out);
void ppo_ba_destroy(ppo_ba_handle *h);
const char *ppo_ba_last_error(const ppo_ba_handle *h);
/* Replaces the configuration of an existing handle (the reference re-reads its globals, include/Parameters.h:45-76, on every
 * call); cheaper than destroy + create: streams, pinned staging and the device memory pool are kept. */
int ppo_ba_set_params(ppo_ba_handle *h, const ppo_ba_params *params);

/* Replaces every optimizer.addVertex / addEdge of Optimizer.cc:2120-2714 (and :526-650).
 * Copies the graph to the device; all edges start at level 0 with their Huber kernel on
 * (point, plane, cuboid-cam, cuboid-plane) or off (point-cuboid), as the reference builds them.
 * The caller's arrays are only read during the call.  Pageable arrays go through the handle's pinned staging arena (one host copy);
 * an array that is page-locked already (cudaMallocHost, cudaHostRegister, ppo_ba_host_register) is copied from where it lies. */
int ppo_ba_set_graph(ppo_ba_handle *h, const ppo_ba_graph *g);
/* Page-locks / releases a host range (cudaHostRegister / cudaHostUnregister) for callers that do not link the CUDA runtime themselves:
 * a host that keeps its flat arrays across calls registers them once and saves the staging copy of every ppo_ba_set_graph. */
int ppo_ba_host_register(void *ptr, size_t bytes);
int ppo_ba_host_unregister(void *ptr);

/* One SparseOptimizer::initializeOptimization(0) + optimize(iters)
 * (core/sparse_optimizer.cpp:199-267,354-420): rebuilds the index mapping from the current
 * edge levels, re-initialises lambda, runs <= iters LM iterations.  stop_flag (may be NULL)
 * is polled where g2o polls forceStopFlag (sparse_optimizer.cpp:376, levenberg.cpp:149). */
int ppo_ba_optimize(ppo_ba_handle *h, int iters, const volatile unsigned char *stop_flag,
                    ppo_ba_stats *stats);

/* Batch variants: n independent windows (one handle each, possibly on different devices) optimised concurrently;
 * every handle replays its own LM graph on its own stream, so the latency-bound phases of one window
 * overlap the others (BASELINE configs[3]: many independent key-frame windows).  The windows of one device
 * share its SMs for the persistent factorisation of the reduced system (SMs / windows CTAs each), so that
 * their dependency chains run side by side.  stats / res are arrays of n.
 * Returns the first non-zero error code of any window (all windows are always run to completion). */
int ppo_ba_optimize_batch(ppo_ba_handle **h, int n, int iters, const volatile unsigned char *stop_flag, ppo_ba_stats *stats);
int ppo_ba_local_ba_batch(ppo_ba_handle **h, int n, const volatile unsigned char *stop_flag, ppo_ba_result *res);

/* e->chi2() of every edge of a kind as g2o holds it after the last computeActiveErrors
 * (stale for level-1 edges and after a rejected last trial), and
 * e->isDepthPositive() from the current estimates (point edges; plane edges: distance()>0).
 * err_norm (may be NULL) receives ||e->error()|| (used for cuboid edges, Optimizer.cc:2774). */
int ppo_ba_edge_chi2(ppo_ba_handle *h, int kind, double *chi2, unsigned char *depth_positive,
                     double *err_norm);

/* The point edges LocalBundleAdjustment / LocalBACameraPlaneCuboids erase after the optimisation (Optimizer.cc:2840-2852, :719-741):
 * ascending indices of the point edges with e->chi2() > chi2_mono (monocular) / chi2_stereo (stereo) or !e->isDepthPositive() --
 * the test of ppo_ba_edge_chi2's outputs done on the device, so that a few KB come back instead of 9 bytes per edge.
 * *idx points into the handle and stays valid until the next call of this function or ppo_ba_set_graph. */
int ppo_ba_point_edge_outliers(ppo_ba_handle *h, double chi2_mono, double chi2_stereo, const int32_t **idx, int32_t *n);

/* e->computeError() at the current estimates for every LEVEL-1 edge of a kind (the optimiser skips those, so their
 * stored error is stale); afterwards ppo_ba_edge_chi2 reports a fresh chi2 for them.  This is what
 * Optimizer::PoseOptimization does before re-classifying its outliers (Optimizer.cc:400-403,431-434).
 * Point edges only (PPO_EDGE_POINT); other kinds return PPO_E_INVALID. */
int ppo_ba_recompute_edge_errors(ppo_ba_handle *h, int kind);

/* e->setLevel(level) / e->setRobustKernel(0) for all edges of a kind.  flags[i] = PPO_EF_* bits. */
int ppo_ba_set_edge_flags(ppo_ba_handle *h, int kind, const unsigned char *flags);
int ppo_ba_get_edge_flags(ppo_ba_handle *h, int kind, unsigned char *flags);
int ppo_ba_edge_count(const ppo_ba_handle *h, int kind);

/* The outlier pass of Optimizer.cc:2736-2833 executed on the device (no per-edge D2H). */
int ppo_ba_outlier_pass(ppo_ba_handle *h, int32_t n_out[3]);

/* Whole call: [stop?] optimize(iters_round1) -> [stop?] outlier pass -> optimize(iters_round2)
 * = stages C-E of Optimizer.cc:2727-2837 (or :668-715 for the points-only graph). */
int ppo_ba_local_ba(ppo_ba_handle *h, const volatile unsigned char *stop_flag, ppo_ba_result *res);

/* Reads the vertex estimates back (stage G inputs, Optimizer.cc:2913-2966). NULL members skipped. */
int ppo_ba_get_state(ppo_ba_handle *h, ppo_ba_state *out);

/* Restores the estimates given at ppo_ba_set_graph and the initial edge flags (benchmark reuse). */
int ppo_ba_reset(ppo_ba_handle *h);

/* -- instrumentation -------------------------------------------------------------------- */
/* Device timing of individual phases of the last linearisation (CUDA events on the handle's
 * stream).  enable != 0 adds per-phase events to ppo_ba_optimize. */
int ppo_ba_set_profiling(ppo_ba_handle *h, int enable);
/* Number of kernel launches issued by this handle since creation (kernels replayed by the captured LM graph are counted per replay). */
long long ppo_ba_launch_count(const ppo_ba_handle *h);
/* Number of times the host blocked on the device inside ppo_ba_optimize since creation.  With the LM controller on the device
 * (default: the whole loop of core/optimization_algorithm_levenberg.cpp:61-164 runs as one CUDA graph with conditional WHILE nodes)
 * an optimize() call blocks twice: once for the sizes of the index mapping, once for the result. */
long long ppo_ba_host_sync_count(const ppo_ba_handle *h);
/* enable = 0: drive the same device-side controller kernels from a host loop (one blocking read of two loop flags per damped trial)
 * instead of the captured graph.  Profiling and sharded windows always use the host loop.  Default: enabled (env PPO_BA_NO_GRAPH=1 disables). */
int ppo_ba_set_graph_mode(ppo_ba_handle *h, int enable);
/* Runs ONLY the point-edge Jacobian/assembly kernel `reps` times on the current state and
 * returns its mean device time in ms (CUDA events on the handle's stream) and the algorithmic
 * bytes one launch moves.  Used by bench.py for the roofline object. */
int ppo_ba_time_assembly(ppo_ba_handle *h, int reps, double *ms_mean, double *algo_bytes);

/* Same for the dense solve of the reduced pose system (factorisation + substitutions); flops = n_p^3/3 + 2 n_p^2. */
int ppo_ba_time_solve(ppo_ba_handle *h, int reps, double *ms_mean, double *flops, int *n_p);
/* Device-side stopwatch on the handle's stream: mark(0) ... mark(1), elapsed = ms between the two
 * CUDA events (includes every gap in which the stream idles waiting for the host LM controller). */
int ppo_ba_mark(ppo_ba_handle *h, int which);
int ppo_ba_elapsed_ms(ppo_ba_handle *h, double *ms);
/* Writes a scratch buffer larger than L2 (256 MiB) on the handle's stream: L2 flush between timed steps. */
int ppo_ba_flush_l2(ppo_ba_handle *h);

/* -- parity / debugging exports (the oracle offers the same two calls) ---------------------- */
/* One linearisation at the current estimates: Hpp n_p x n_p row-major (upper blocks filled),
 * b = [pose gradient | landmark gradients of the ACTIVE landmarks, planes first], Hll n_l x 9.
 * dims = {n_p, n_l}; call with NULL outputs to size the buffers. */
int ppo_ba_debug_linearize(ppo_ba_handle *h, int32_t dims[2], double *Hpp, double *b, double *Hll,
                           double *chi2);
/* After debug_linearize: damped Schur system and its solution for a given lambda. */
int ppo_ba_debug_solve(ppo_ba_handle *h, double lambda, double *Hschur_upper, double *bschur,
                       double *x, int32_t *ok);
/* The linear solver of the handle's stack alone, on a caller-supplied dense symmetric system (A: n x n row-major, upper
 * triangle used): LinearSolver::solve(A, x, b).  PPO_SOLVER_DENSE_X = LinearSolverDense (solvers/linear_solver_dense.h:65-113:
 * *ok = 0 unless A is positive definite); PPO_SOLVER_6_3 = LinearSolverEigen (solvers/linear_solver_eigen.h:94-124: LDL^T
 * without pivoting, *ok = 0 only on a zero pivot, so an indefinite A is solved).  Does not need or touch a resident window. */
int ppo_ba_debug_dense_solve(ppo_ba_handle *h, int32_t n, const double *A_upper, const double *b, double *x, int32_t *ok);

/* -- multi-GPU: one window, landmarks sharded over ranks ------------------------ */
/* Call before ppo_ba_set_graph, on every rank (one process per GPU of ONE node; world <= 8 for the
 * distributed solve).  Each rank then sets a graph with the SAME key-frames, cuboids, camera-cuboid and
 * point-cuboid edges and ITS OWN landmarks: a slice of the points with their edges and a slice of the
 * planes with their plane edges and the cuboid-plane edges that name them (sharding.py: shard_graph).
 * Rank 0 alone accumulates the replicated edges.  Per linearisation the per-key-frame pose blocks of the
 * landmark edges are all-reduced (NCCL, n_kf x 27 doubles); per damped trial every rank accumulates the
 * Schur complement of ITS landmarks into its copy of the reduced system, which lives in peer-mapped
 * (cudaIpc) memory: tile column j is summed by its owner, rank j mod world, straight from the other ranks'
 * copies, the Cholesky factorisation is spread over the ranks by tile column with the finished panel
 * tiles pushed to every rank over NVLink, and every rank back-substitutes and updates its own landmarks
 * (csrc/cuda/ppo_dense.cu, "distributed factorisation").  {chi2, scale, failed-pivot flag} are all-reduced
 * per trial, so all ranks take the same LM decisions.  ppo_ba_set_graph / ppo_ba_optimize / ppo_ba_local_ba
 * are COLLECTIVE in this mode.  PPO_DIST_SOLVE=0 in the environment falls back to an ncclAllReduce of the
 * whole reduced system with a replicated factorisation.  nccl_comm is an ncclComm_t passed as void*. */
int ppo_ba_set_shard(ppo_ba_handle *h, void *nccl_comm, int rank, int world);
/* NCCL plumbing without a link-time dependency (libnccl.so.2 is dlopen'ed; the torch wheel already maps it):
 * rank 0 creates the 128-byte unique id, the host program broadcasts it (e.g. torch.distributed), every rank
 * creates its communicator for `device`. */
int ppo_ba_nccl_unique_id(char out[128]);
int ppo_ba_nccl_init(const char id_bytes[128], int rank, int world, int device, void **comm);
int ppo_ba_nccl_destroy(void *comm);
long long ppo_ba_collective_count(const ppo_ba_handle *h);

#ifdef __cplusplus
}
#endif
#endif /* PPO_BA_H */
