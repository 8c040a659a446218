/*
 * ppo_synth.h — seeded synthetic local-BA windows: the flat graph a
 * replacement Optimizer.cc would hand to ppo_ba_set_graph for a window of N_kf key-frames,
 * N_pt map points, N_pl planes and N_cu cuboids.  Input generation only: no solver code.
 */
#ifndef PPO_SYNTH_H
#define PPO_SYNTH_H
#include <stdint.h>

#include "ppo_ba.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ppo_synth_cfg {
  int32_t n_kf;        /* local key-frames (id 0 fixed)                          */
  int32_t n_fixed;     /* extra fixed key-frames; <0 => max(2, n_kf/10)          */
  int32_t n_pt, n_pl, n_cu;
  uint64_t seed;
  int32_t cuboid_2d;   /* optimize_with_cuboid_2d  : bbox edges                   */
  int32_t corners_2d;  /* optimize_with_corners_2d : corner edges                 */
  int32_t pt_obj_3d;   /* optimize_with_pt_obj_3d                                 */
  int32_t cuboid_plane;/* optimize_with_cuboid_plane                              */
  int32_t plane_3d;    /* optimize_with_plane_3d                                  */
  double outlier_frac; /* gross outliers among point observations (0.03)         */
  double stereo_frac;  /* fraction of stereo observations (0.7)                   */
  int32_t sort_points; /* 1: order points by their first observing KF (banded CSR) */
  int32_t cuboid_3d;   /* optimize_with_cuboid_3d: EdgeSE3Cuboid edges (PPO_CUBOID_SE3), LocalBACameraPointCuboids2D only */
} ppo_synth_cfg;

typedef struct ppo_synth ppo_synth;

/* BASELINE.json configs[i], i = 0..4 (config 3 = one of its 64 windows; pass window w). */
void ppo_synth_config(int config_index, int window, ppo_synth_cfg *cfg);
ppo_synth *ppo_synth_create(const ppo_synth_cfg *cfg);
const ppo_ba_graph *ppo_synth_graph(const ppo_synth *s);
/* ground truth the estimates were perturbed from (same layouts as ppo_ba_state) */
void ppo_synth_truth(const ppo_synth *s, ppo_ba_state *out);
void ppo_synth_destroy(ppo_synth *s);

#ifdef __cplusplus
}
#endif
#endif
