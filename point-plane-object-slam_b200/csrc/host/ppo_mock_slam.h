// Minimal stand-ins for the reference's map state types, with IDENTICAL member names, types and access
// patterns for everything the local BA touches (SURVEY.md section 8a "State layout"):
//   KeyFrame.h:76-77,93,112-116,133,151,167-168,184,191-192,209,216,221-222,234-242
//   MapPoint.h:45-61,77,89,96,113     MapPlane.h:34-68     MapCuboid.h:41-116     Map.h:67
// The reference headers need OpenCV / PCL / Eigen, none of which exist in this image, so the Optimizer shim
// (ppo_optimizer_shim.cpp) is compiled and tested against these; with the real headers on the include path
// (-DPPO_WITH_ORB_SLAM2) the same shim source binds to the real classes.
#pragma once
#include <cstddef>
#include <cstring>
#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

#ifndef CV_32F
#define CV_32F 5
#endif

namespace cv {
// float-only matrix of at most 4x4: what Converter reads/writes with at<float>(r,c)
class Mat {
 public:
  Mat() : rows(0), cols(0) { std::memset(d, 0, sizeof d); }
  Mat(int r, int c, int /*type*/) : rows(r), cols(c) { std::memset(d, 0, sizeof d); }
  template <typename T> T &at(int r, int c = 0) { return d[r * cols + c]; }
  template <typename T> const T &at(int r, int c = 0) const { return d[r * cols + c]; }
  Mat clone() const { return *this; }
  bool empty() const { return rows == 0; }
  void create(int r, int c, int /*type*/) { rows = r, cols = c; }
  void copyTo(Mat &dst) const { dst = *this; }
  int rows, cols;
  float d[16];
};
struct Point2f { float x, y; };
struct KeyPoint { Point2f pt; int octave; };
struct Rect { int x, y, width, height; };
}  // namespace cv

namespace Eigen {  // the few fixed-size types the BA reads from MapCuboid / MapPlane
struct Vector3d { double v[3]; double &operator()(int i) { return v[i]; } double operator()(int i) const { return v[i]; } };
struct Vector4d { double v[4]; double &operator()(int i) { return v[i]; } double operator()(int i) const { return v[i]; } };
struct Matrix2Xd { double v[2][8]; double &operator()(int r, int c) { return v[r][c]; } double operator()(int r, int c) const { return v[r][c]; } };
}  // namespace Eigen

namespace g2o {
// g2o::cuboid as the BA sees it: SE3Quat pose (object -> world) + half scale  (include/g2o_cuboid.h:33-34)
struct cuboid {
  double pose7[7];  // [qx qy qz qw tx ty tz]
  double scale[3];
};
}  // namespace g2o

namespace ORB_SLAM2 {
class KeyFrame;
class MapCuboid;

class MapPoint {
 public:
  void SetWorldPos(const cv::Mat &Pos) { mWorldPos = Pos.clone(); }
  cv::Mat GetWorldPos() { return mWorldPos.clone(); }
  std::map<KeyFrame *, size_t> GetObservations() { return mObservations; }
  int Observations() { return nObs; }
  void AddObservation(KeyFrame *pKF, size_t idx) { mObservations[pKF] = idx; mnObsVersion++; }
  void EraseObservation(KeyFrame *pKF) { erased.push_back(pKF); mObservations.erase(pKF); mnObsVersion++; }
  bool isBad() { return mbBad; }
  void UpdateNormalAndDepth() { n_updates++; }
  // Observation version (INTEGRATION.md "observation mirror"): the ONE member the shim asks the reference to add -- incremented wherever
  // MapPoint.cc changes mObservations (AddObservation :95-107, EraseObservation :109-139, SetBadFlag :151-167, Replace :176-213).  With it the
  // shim keeps the flattened observation row of a map point across local-BA calls instead of copying its std::map every time.
  long unsigned int mnObsVersion = 1;
#define PPO_HAVE_OBS_VERSION 1
  std::map<MapCuboid *, int> MapObjObservations;
  long unsigned int mnId = 0;
  long unsigned int mnBALocalForKF = 0;
  cv::Mat mPosGBA;                        // MapPoint.h:120-121 (global BA results when nLoopKF != 0)
  long unsigned int mnBAGlobalForKF = 0;
  // test bookkeeping
  cv::Mat mWorldPos;
  std::map<KeyFrame *, size_t> mObservations;
  int nObs = 0, n_updates = 0;
  bool mbBad = false;
  std::vector<KeyFrame *> erased;
};

class MapPlane {
 public:
  void SetWorldPos(const cv::Mat &Pos) { mWorldPos = Pos.clone(); }
  cv::Mat GetWorldPos() { return mWorldPos.clone(); }
  void EraseObservation(KeyFrame *pKF) { erased.push_back(pKF); mObservations.erase(pKF); }
  std::map<KeyFrame *, int> GetObservations() { return mObservations; }
  std::map<KeyFrame *, int> GetParObservations() { return mParObservations; }
  std::map<KeyFrame *, int> GetVerObservations() { return mVerObservations; }
  bool isBad() { return mbBad; }
  long unsigned int mnId = 0;
  long unsigned int mnBALocalForKF = 0;
  long unsigned int asso_cuboid_id = 999;
  Eigen::Vector3d asso_cuboid_meas{};
  cv::Mat mWorldPos;
  std::map<KeyFrame *, int> mObservations, mParObservations, mVerObservations;
  bool mbBad = false;
  std::vector<KeyFrame *> erased;
};

class MapCuboid {
 public:
  void SetWorldPos(const cv::Mat &Pos) { mWorldPos = Pos.clone(); }
  std::unordered_map<KeyFrame *, size_t> GetObservations() { return mObservations; }
  bool isBad() { return mbBad; }
  std::vector<MapPoint *> GetUniqueMapPoints() { return mappoints_unique_own; }
  long int mnId = 0;
  int object_graph_id = 0;                 // MapCuboid.h:97
  g2o::cuboid cuboid_local_meas{};         // MapCuboid.h:100: local measurement in the camera frame
  g2o::cuboid cuboid_global_data{};
  double meas_quality = 0.7;
  Eigen::Vector4d bbox_vec{};
  cv::Rect bbox_2d{};
  Eigen::Matrix2Xd box_corners_2d{};
  long unsigned int mnBALocalForKF = 0;
  bool obj_been_optimized = false;
  g2o::cuboid cuboid_global_opti{};
  std::vector<MapPoint *> used_points_in_BA_filtered;
  std::vector<MapPoint *> mappoints_unique_own;
  std::unordered_map<KeyFrame *, size_t> mObservations;
  cv::Mat mWorldPos;
  bool mbBad = false;
};

class KeyFrame {
 public:
  KeyFrame(float fx_, float fy_, float cx_, float cy_, float bf_) : fx(fx_), fy(fy_), cx(cx_), cy(cy_), mbf(bf_) {}
  void SetPose(const cv::Mat &Tcw_) { Tcw = Tcw_.clone(); n_setpose++; }
  cv::Mat GetPose() { return Tcw.clone(); }
  std::vector<KeyFrame *> GetVectorCovisibleKeyFrames() { return mvpOrderedConnectedKeyFrames; }
  void EraseMapPointMatch(MapPoint *pMP) { erased_points.push_back(pMP); }
  void EraseMapPlaneMatch(MapPlane *pMP) { erased_planes.push_back(pMP); }
  std::vector<MapPoint *> GetMapPointMatches() { return mvpMapPoints; }
  bool isBad() { return mbBad; }
  long unsigned int mnId = 0;
  long unsigned int mnBALocalForKF = 0;
  long unsigned int mnBAFixedForKF = 0;
  cv::Mat mTcwGBA;                        // KeyFrame.h:179-181
  long unsigned int mnBAGlobalForKF = 0;
  const float fx, fy, cx, cy, mbf;
  std::vector<cv::KeyPoint> mvKeysUn;
  std::vector<float> mvuRight;
  std::vector<float> mvInvLevelSigma2;
  std::vector<MapCuboid *> local_cuboids;
  std::vector<MapCuboid *> mvpMapCuboid;
  std::vector<cv::Mat> mvPlaneCoefficients;
  std::vector<MapPlane *> mvpMapPlanes;
  // test bookkeeping
  cv::Mat Tcw;
  std::vector<KeyFrame *> mvpOrderedConnectedKeyFrames;
  std::vector<MapPoint *> mvpMapPoints;
  std::vector<MapPoint *> erased_points;
  std::vector<MapPlane *> erased_planes;
  bool mbBad = false;
  int n_setpose = 0;
};

// Frame.h:100-190 (only what Optimizer::PoseOptimization touches)
class Frame {
 public:
  Frame(float fx_, float fy_, float cx_, float cy_, float bf_) : fx(fx_), fy(fy_), cx(cx_), cy(cy_), mbf(bf_) {}
  void SetPose(cv::Mat Tcw) { mTcw = Tcw.clone(); n_setpose++; }
  int N = 0;
  std::vector<cv::KeyPoint> mvKeysUn;
  std::vector<float> mvuRight;
  std::vector<float> mvInvLevelSigma2;
  std::vector<MapPoint *> mvpMapPoints;  // NULL: no association
  std::vector<bool> mvbOutlier;
  float fx, fy, cx, cy, mbf;
  cv::Mat mTcw;
  int n_setpose = 0;  // test bookkeeping
};

class Map {
 public:
  std::vector<KeyFrame *> GetAllKeyFrames() { return mvpKeyFrames; }  // Map.h:54-55
  std::vector<MapPoint *> GetAllMapPoints() { return mvpMapPoints; }
  std::mutex mMutexMapUpdate;
  // test bookkeeping
  std::vector<KeyFrame *> mvpKeyFrames;
  std::vector<MapPoint *> mvpMapPoints;
};

// include/Parameters.h:45-76 (only what the local BA reads)
extern bool optimize_with_cuboid_plane, optimize_with_plane_3d, optimize_with_cuboid_2d, optimize_with_corners_2d, optimize_with_pt_obj_3d, optimize_with_cuboid_3d;
extern double ba_weight_bbox, ba_weight_corner, thHuberBbox2d, thHuberConer2d, ba_weight_SE3, thHuberSE3;
extern double plane_angle_info, plane_dist_info, plane_chi, cuboid_plane_angle_info, cuboid_plane_dist_info, cuboid_plane_chi;

// include/Optimizer.h:40-45,62 — the entry points the shim re-implements, signatures unchanged
class Optimizer {
 public:
  void static BundleAdjustment(const std::vector<KeyFrame *> &vpKF, const std::vector<MapPoint *> &vpMP, int nIterations = 5, bool *pbStopFlag = NULL,
                               const unsigned long nLoopKF = 0, const bool bRobust = true);
  void static GlobalBundleAdjustemnt(Map *pMap, int nIterations = 5, bool *pbStopFlag = NULL, const unsigned long nLoopKF = 0, const bool bRobust = true);
  int static PoseOptimization(Frame *pFrame);
  void static LocalBundleAdjustment(KeyFrame *pKF, bool *pbStopFlag, Map *pMap);
  void static LocalBACameraPlaneCuboids(KeyFrame *pKF, bool *pbStopFlag, Map *pMap, bool fixCamera = false, bool fixPoint = false);
  void static LocalBACameraPointCuboids2D(KeyFrame *pKF, bool *pbStopFlag, Map *pMap, bool fixCamera = false, bool fixPoint = false);  // Optimizer.h:60
};
}  // namespace ORB_SLAM2
