// Device-resident layout of one local-BA window.
#pragma once
#include <cstdint>

namespace ppo {

// packed point-edge record, 16 B, loaded as one int4 (coalesced, 2 edges per 32-B sector)
struct __align__(16) PointEdgeRec {
  int kf;          // key-frame slot
  float u, v, ur;  // observation; ur < 0 => monocular
};

// Estimates of every vertex. Two copies exist (current / trial) and are swapped on LM accept —
// the device equivalent of g2o's push / pop / discardTop estimate stack.
struct DevState {
  double *kf_pose;  // n_kf x 7   [qx qy qz qw tx ty tz]
  double *kf_Rt;    // n_kf x 12  rotation (row-major) + translation, derived cache of kf_pose
  double *pt;       // n_pt x 3
  double *pl;       // n_pl x 4
  double *cu;       // n_cu x 10
};

struct DevGraph {
  int n_kf, n_pt, n_pl, n_cu;
  int n_pe, n_ple, n_cbe, n_pce, n_cpe;
  int n_slots;  // unique (plane, key-frame) pairs: the Hpl blocks of plane landmarks
  int n_ent;    // n_slots + n_pe : Hpl blocks ("entries") of all landmarks
  int n_lm;     // n_pl + n_pt   : landmarks, planes first
  // vertices (constant part)
  const uint8_t *kf_fixed, *pt_fixed, *cu_flags;
  const float *kf_intr;  // n_kf x 5
  // point edges (CSR by point; entry id of edge e is n_slots + e)
  const int *pt_rowptr;
  const PointEdgeRec *pe_rec;
  const float *pe_is2;
  const int *pe_pt;
  uint8_t *pe_flags;
  double *pe_chi2;
  // linearisation work units: consecutive points packed so that a warp owns <= 32 edges
  int n_units;
  const int *unit_pt0;  // n_units + 1 (first point of each unit)
  const int *unit_e0;   // n_units + 1 (first edge of each unit)
  // point edges grouped by key-frame, cut into chunks of <= 256 edges
  int n_chunks;
  const int *chunk_kf, *chunk_begin, *chunk_end, *kfe_edge;
  // plane edges
  const int *ple_plane, *ple_kf, *ple_slot;
  const uint8_t *ple_kind;
  const double *ple_meas, *ple_info;
  uint8_t *ple_flags;
  double *ple_chi2;
  double *ple_J;  // n_ple x 9 columns x 3
  // camera-cuboid edges
  const int *cbe_kf, *cbe_cuboid;
  const uint8_t *cbe_kind;
  const double *cbe_meas, *cbe_info;
  uint8_t *cbe_flags;
  double *cbe_chi2, *cbe_norm;
  double *cbe_J;  // n_cbe x 15 columns x 16
  double *cbe_err, *cbe_w;  // n_cbe x 16 residual, n_cbe robust weight x information (linearisation scratch)
  // point-cuboid edges
  const int *pce_cuboid, *pce_rowptr;
  const double *pce_pts;
  uint8_t *pce_flags;
  double *pce_chi2;
  double *pce_J;  // n_pce x 9 columns x 3
  // fixed-order assembly of the non-point edges: per-edge contributions + gather lists (built once per window); no atomics
  double *ple_part;  // n_ple x 54: [0..26] key-frame block (21 upper + 6 gradient), [27..35] plane Hll (6) + bl (3), [36..53] Hpl block 6 x 3
  double *cbe_part;  // n_cbe x 81: [0..26] key-frame block, [27..80] cuboid block (45 upper + 9 gradient)
  double *pce_part;  // n_pce x 54: cuboid block
  const int *kf_ple_ptr, *kf_ple_idx;      // plane edges of a key-frame
  const int *kf_cbe_ptr, *kf_cbe_idx;      // camera-cuboid edges of a key-frame
  const int *cu_cbe_ptr, *cu_cbe_idx;      // camera-cuboid edges of a cuboid
  const int *cu_pce_ptr, *cu_pce_idx;      // point-cuboid edges of a cuboid
  const int *pl_ple_ptr, *pl_ple_idx;      // plane edges of a plane
  const int *slot_ple_ptr, *slot_ple_idx;  // plane edges of a (plane, key-frame) slot
  // landmark -> entries CSR (planes: slots; points: edges)
  const int *lm_rowptr;  // n_lm + 1
  const int *slot_kf;    // n_slots
  // index mapping of the current optimize()
  int *kf_act, *cu_act, *pl_act, *pt_act;
  int *kf_idx;    // pose-block index among free active key-frames, -1 otherwise
  int *cu_off;    // scalar offset of the cuboid in the pose block, -1 otherwise
  int *ent_pidx;  // per entry: key-frame block index, -1 fixed KF, -2 inactive
  int *dims;      // [0] free active KFs, [1] n_p
  // normal equations
  double *Hpp_kf;  // n_kf x 36 (indexed by kf_idx)
  double *Hpp_cu;  // n_cu x 81
  double *Hpc;     // n_cbe x 54 : (KF, cuboid) off-diagonal contribution of each camera-cuboid edge
  double *bp;      // pose-block gradient, n_p
  double *Hll;     // n_lm x 6 (xx xy xz yy yz zz)
  double *bl;      // n_lm x 3
  double *Hpl;     // n_ent x 18 (6 x 3 row-major)
  double *BD;      // n_ent x 18 : Y = Hpl * C of the current damped trial, C C^T = (Hll + lambda)^-1
  double *Zent;    // n_ent x 3  : z = C^T bl of the entry's landmark (reduced-gradient operand of k_schur_pairs)
  double *Dinv;    // n_lm x 6
  double *xl;      // n_lm x 3
  double *S;       // (n_p + 1) x ld, row-major upper = column-major lower; last column = reduced rhs
  double *xp;      // n_p
  // huber deltas / constants
  double huber_mono, huber_stereo, huber_plane, huber_vp, huber_bbox, huber_corner, huber_se3;
  double ptcu_ratio, ptcu_prior;
};

}  // namespace ppo
