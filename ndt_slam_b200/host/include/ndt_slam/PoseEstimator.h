// PoseEstimator.h -- NDT pose estimation behind the reference's class interface
// [REF include/ndt_slam/PoseEstimator.h:36-133, src/PoseEstimator.cpp:4-69].
//
// Where the reference owns a pcl::NDT object and calls ApproximateVoxelGrid / setInputSource /
// setInputTarget / align / getFitnessScore / getHessian on it, this class owns one ndt_handle of the
// C ABI (include/ndt_b200.h): the grid build and the whole Newton / More-Thuente match run as CUDA
// kernels on the B200. Constructor parameters, setScanPair overloads, estimatePose signature, the
// theta extraction, the cost sentinel and the covariance post-processing are the reference's.
// There is no CPU fallback: estimatePose throws std::runtime_error if the device path is unavailable.
#ifndef NDT_SLAM_B200_POSEESTIMATOR_H_
#define NDT_SLAM_B200_POSEESTIMATOR_H_

#include <Eigen/Core>
#include <pcl/point_cloud.h>
#include <ros/ros.h>

#include "MyUtil.h"
#include "Pose2D.h"
#include "Scan2D.h"
#include "Timer.h"
#include "ndt_b200.h"

class PoseEstimator {
 private:
  pcl::PointCloud<pcl::PointXYZ>::Ptr source_cloud;   // current scan, float32 (z = 0)
  pcl::PointCloud<pcl::PointXYZ>::Ptr target_cloud;   // reference map (aliased, not copied)

  const Scan2D *curScan;
  const Scan2D *refScan;
  double coeNDTCov;

  double TransformationEpsilon;
  double StepSize;
  double Resolution;
  int MaximumIterations;

  double LeafSize;

  uint64_t hintEpoch = 0, uploadedEpoch = 0;     // see hintTargetPrefix
  size_t hintPrefix = 0, hintSettled = 0;
  const void *uploadedCloud = nullptr;
  uint64_t prefetchedEpoch = 0;                  // see prefetchTarget
  size_t prefetchedPoints = 0;
  bool prefetchTargetEnabled = true;             // parameter prefetch_target (default true)

  bool incrementalTarget;   // parameter incremental_target (default true): use ndt_set_target_incremental when the map hints a settled prefix
  ndt_handle ndt;       // stands where the reference has  pcl::NDT<pcl::PointXYZ, pcl::PointXYZ> ndt;
  Timer timer;

  void ensureHandle();
  static void fillFromScan(const Scan2D *scan, pcl::PointCloud<pcl::PointXYZ> &cloud);

 public:
  double totalError;

  // what the last estimatePose did (instrumentation; not in the reference)
  ndt_result lastResult;
  double lastGridMs, lastMatchMs, lastFilterMs;            // device grid kernels, device match kernel, host voxel filter
  double lastSetSourceWallMs, lastSetTargetWallMs, lastAlignWallMs;   // host wall time of the three ABI calls
  int lastSourcePoints, lastTargetPoints;

  PoseEstimator();
  ~PoseEstimator();
  PoseEstimator(const PoseEstimator &) = delete;
  PoseEstimator &operator=(const PoseEstimator &) = delete;

  // Optional hint (not in the reference): the reference cloud handed to the next estimatePose is version `epoch` of a
  // cloud whose first `stablePoints` points did not change since version epoch - 1. When the previous estimatePose
  // uploaded exactly that previous version, only the changed tail is copied to the device (ndt_set_target_prefix).
  // `settledPoints`: the first settledPoints points of this version stay a prefix of every later version (until a version
  // arrives whose stablePoints is smaller): the device then keeps running per-cell sums of that part and re-derives only the
  // cells a scan touches (ndt_set_target_incremental).
  void hintTargetPrefix(uint64_t epoch, size_t stablePoints, size_t settledPoints = 0) {
    hintEpoch = epoch; hintPrefix = stablePoints; hintSettled = settledPoints;
  }

  // Optional, right after the map has produced the cloud the NEXT estimatePose will match against (ScanMatcher::growMap ->
  // makeLocalMap): queue the device grid update now (ndt_set_target_incremental_async) so that it runs while the host
  // resamples and filters the next scan; estimatePose then finds its target in place. Same arguments as hintTargetPrefix.
  void prefetchTarget(pcl::PointCloud<pcl::PointXYZ>::Ptr refScan, uint64_t epoch, size_t stablePoints, size_t settledPoints);

  void setScanPair(const Scan2D *curScan, pcl::PointCloud<pcl::PointXYZ>::Ptr refScan);
  void setScanPair(const Scan2D *curScan, const Scan2D *refScan);

  double estimatePose(Pose2D &initPose, Pose2D &estPose, Eigen::Matrix3d &cov);
};

#endif
