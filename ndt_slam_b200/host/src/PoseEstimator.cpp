// PoseEstimator.cpp -- estimatePose on the B200 through the C ABI [REF src/PoseEstimator.cpp:4-69].
#include "ndt_slam/PoseEstimator.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>

#include "ndt_slam/VoxelFilter.h"

namespace {
inline double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
[[noreturn]] void fail(ndt_handle h, const char *what) {
  const char *msg = ndt_last_error(h);
  throw std::runtime_error(std::string("PoseEstimator: ") + what + ": " + (msg ? msg : "") + " (no CPU fallback)");
}
}  // namespace

PoseEstimator::PoseEstimator()
    : curScan(nullptr), refScan(nullptr), coeNDTCov(1.0), TransformationEpsilon(0.01), StepSize(0.1), Resolution(1.0),
      MaximumIterations(35), LeafSize(0.1), incrementalTarget(true), ndt(nullptr), totalError(0.0), lastGridMs(0), lastMatchMs(0), lastFilterMs(0),
      lastSetSourceWallMs(0), lastSetTargetWallMs(0), lastAlignWallMs(0),
      lastSourcePoints(0), lastTargetPoints(0) {
  ros::param::get("coeNDTCov", coeNDTCov);
  ros::param::get("TransformationEpsilon", TransformationEpsilon);
  ros::param::get("StepSize", StepSize);
  ros::param::get("Resolution", Resolution);
  ros::param::get("MaximumIterations", MaximumIterations);
  ros::param::get("LeafSize", LeafSize);
  ros::param::get("incremental_target", incrementalTarget);
  ros::param::get("prefetch_target", prefetchTargetEnabled);
  source_cloud = std::make_shared<pcl::PointCloud<pcl::PointXYZ>>();
  target_cloud = std::make_shared<pcl::PointCloud<pcl::PointXYZ>>();
  std::memset(&lastResult, 0, sizeof(lastResult));
}

PoseEstimator::~PoseEstimator() {
  if (ndt) ndt_destroy(ndt);
}

// The handle takes the values the reference hands to ndt.setTransformationEpsilon / setStepSize /
// setResolution / setMaximumIterations in its constructor (PoseEstimator.h:77-83). Created on first
// use so that constructing the object does not need a device.
void PoseEstimator::ensureHandle() {
  if (ndt) return;
  ndt_params prm;
  ndt_params_default(&prm);
  prm.resolution = static_cast<float>(Resolution);
  prm.step_size = StepSize;
  prm.trans_eps = TransformationEpsilon;
  prm.max_iter = MaximumIterations;
  int dev = 0;
  ros::param::get("cuda_device", dev);
  prm.device = dev;
  if (ndt_create(&prm, &ndt) != NDT_OK) fail(nullptr, "ndt_create");
}

void PoseEstimator::fillFromScan(const Scan2D *scan, pcl::PointCloud<pcl::PointXYZ> &cloud) {
  const size_t n = scan->lps.size();
  cloud.points.resize(n);
  for (size_t i = 0; i < n; ++i) {
    pcl::PointXYZ &p = cloud.points[i];
    p.x = static_cast<float>(scan->lps[i].x);      // double -> float32 exactly where the reference casts
    p.y = static_cast<float>(scan->lps[i].y);
    p.z = 0.f;
  }
  cloud.width = static_cast<uint32_t>(n);
  cloud.height = 1;
  cloud.is_dense = false;
}

void PoseEstimator::setScanPair(const Scan2D *cur, pcl::PointCloud<pcl::PointXYZ>::Ptr ref) {
  curScan = cur;
  fillFromScan(cur, *source_cloud);
  target_cloud = ref;                              // aliased like the reference (PoseEstimator.h:103)
}

void PoseEstimator::setScanPair(const Scan2D *cur, const Scan2D *ref) {
  curScan = cur;
  refScan = ref;
  fillFromScan(cur, *source_cloud);
  target_cloud = std::make_shared<pcl::PointCloud<pcl::PointXYZ>>();
  fillFromScan(ref, *target_cloud);
}

void PoseEstimator::prefetchTarget(pcl::PointCloud<pcl::PointXYZ>::Ptr ref, uint64_t epoch, size_t stablePoints, size_t settledPoints) {
  if (!prefetchTargetEnabled || !incrementalTarget || epoch == 0 || !ref) return;
  ensureHandle();
  int64_t n_same = 0;
  if (uploadedEpoch != 0 && epoch == uploadedEpoch + 1 && uploadedCloud == ref.get()) n_same = (int64_t)std::min(stablePoints, ref->points.size());
  const int64_t n_target = (int64_t)ref->points.size();
  if (ndt_set_target_incremental_async(ndt, reinterpret_cast<const float *>(ref->points.data()), n_target, n_same,
                                       (int64_t)std::min(settledPoints, ref->points.size()), NDT_MEM_HOST) != NDT_OK)
    fail(ndt, "ndt_set_target_incremental_async");
  uploadedEpoch = epoch; uploadedCloud = ref.get();
  prefetchedEpoch = epoch; prefetchedPoints = ref->points.size();
}

double PoseEstimator::estimatePose(Pose2D &initPose, Pose2D &estPose, Eigen::Matrix3d &cov) {
  ensureHandle();

  // source pre-filter (ApproximateVoxelGrid with LeafSize on every axis)
  double t0 = now_ms();
  pcl::PointCloud<pcl::PointXYZ> filtered;
  ndt_host::approximate_voxel_grid(*source_cloud, static_cast<float>(LeafSize), filtered);
  lastFilterMs = now_ms() - t0;
  lastSourcePoints = static_cast<int>(filtered.points.size());
  lastTargetPoints = static_cast<int>(target_cloud->points.size());

  timer.start_timer();
  // ndt.setInputSource(filtered) ; ndt.setInputTarget(target_cloud) -> device grid build
  t0 = now_ms();
  if (ndt_set_source(ndt, reinterpret_cast<const float *>(filtered.points.data()), (int64_t)filtered.points.size(), NDT_MEM_HOST) != NDT_OK)
    fail(ndt, "ndt_set_source");
  lastSetSourceWallMs = now_ms() - t0;
  t0 = now_ms();
  // a local map that only changed at its end since the cloud uploaded last time: copy the tail only
  int64_t n_same = 0;
  if (hintEpoch != 0 && uploadedEpoch != 0 && hintEpoch == uploadedEpoch + 1 && uploadedCloud == target_cloud.get())
    n_same = (int64_t)std::min(hintPrefix, target_cloud->points.size());
  const int64_t n_target = (int64_t)target_cloud->points.size();
  const float *target_pts = reinterpret_cast<const float *>(target_cloud->points.data());
  int rc = NDT_OK;
  if (hintEpoch != 0 && prefetchedEpoch == hintEpoch && uploadedCloud == target_cloud.get() && prefetchedPoints == target_cloud->points.size()) {
    // the grid of exactly this cloud was queued by prefetchTarget (ndt_set_source above already waited for it)
  } else if (hintEpoch != 0 && incrementalTarget) {
    rc = ndt_set_target_incremental(ndt, target_pts, n_target, n_same, (int64_t)std::min(hintSettled, target_cloud->points.size()), NDT_MEM_HOST);
  } else {
    rc = ndt_set_target_prefix(ndt, target_pts, n_target, n_same, NDT_MEM_HOST);
  }
  if (rc != NDT_OK) fail(ndt, "ndt_set_target");
  uploadedEpoch = hintEpoch; uploadedCloud = target_cloud.get();
  hintEpoch = 0; hintPrefix = 0; hintSettled = 0;
  prefetchedEpoch = 0;
  lastSetTargetWallMs = now_ms() - t0;
  float ms = 0.f;
  ndt_last_kernel_ms(ndt, &ms);
  lastGridMs = ms;

  // ndt.align(output, Translation3f(tx, ty, 0) * AngleAxisf(DEG2RAD(th), Z)): the ABI rounds the guess through float
  const double guess[3] = {initPose.tx, initPose.ty, DEG2RAD(initPose.th)};
  ndt_result res;
  t0 = now_ms();
  if (ndt_align(ndt, guess, &res) != NDT_OK) fail(ndt, "ndt_align");
  lastAlignWallMs = now_ms() - t0;
  ndt_last_kernel_ms(ndt, &ms);
  lastMatchMs = ms;
  lastResult = res;
  timer.end_timer();

  // yaw from the first column of the final float transform, by quadrant (PoseEstimator.cpp:31-35)
  const float r00 = res.T[0], r10 = res.T[1];
  double theta;
  if (r00 > 0 && r10 > 0) theta = std::asin(r10);
  else if (r00 > 0 && r10 < 0) theta = std::asin(r10);
  else if (r00 < 0 && r10 > 0) theta = std::acos(r00);
  else theta = std::acos(r00) * (-1.0);
  estPose.setPose(res.T[12], res.T[13], RAD2DEG(theta));

  double cost = res.fitness;
  if (!res.converged) cost = 10000000;             // failed match: the caller falls back to odometry

  // covariance = (-H)^-1 * coeNDTCov on (x, y, yaw)
  Eigen::Matrix3d negH;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) negH(r, c) = -res.hess[3 * r + c];
  cov = negH.inverse() * coeNDTCov;
  return cost;
}
