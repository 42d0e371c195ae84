// Relocalizer.h -- multi-start global relocalisation on every GPU of the box, driven from C++ through the C ABI alone
// (BASELINE config 4). Not a class of the reference (its matcher only ever tracks from odometry,
// ScanMatcher.cpp:22-45): this is the batched use of PoseEstimator's primitive -- n x ndt.align from n hypotheses
// [REF src/PoseEstimator.cpp:17-29] -- that the north star shards. The data path is the one DESIGN.md section 7 describes:
// the grid is built once on the first device, replicated once (ndt_replicate_grid: peer copies over NVLink), the
// hypotheses are block-partitioned, every device matches its shard with no collective, and only one best result per
// device comes back (ndt_best_of_multi). No NCCL, no torch: a C++ host needs nothing but libndt_b200.so.
#ifndef NDT_SLAM_B200_RELOCALIZER_H_
#define NDT_SLAM_B200_RELOCALIZER_H_

#include <cstdint>
#include <vector>
#include <pcl/point_cloud.h>
#include "Pose2D.h"
#include "ndt_b200.h"

class Relocalizer {
 public:
  // one handle per entry of `devices` (a device may appear more than once: several handles on one GPU)
  explicit Relocalizer(const std::vector<int> &devices, double resolution = 0.5, double stepSize = 0.1, double transEps = 0.01,
                       int maxIter = 35);
  ~Relocalizer();
  Relocalizer(const Relocalizer &) = delete;
  Relocalizer &operator=(const Relocalizer &) = delete;

  // build the NDT grid of the map on the first device and replicate it to the others (no target points: ranking by score)
  void setMap(const pcl::PointCloud<pcl::PointXYZ> &map);
  // the scan to localise (already resampled / voxel-filtered like estimatePose does), sent to every device
  void setScan(const pcl::PointCloud<pcl::PointXYZ> &scan);
  // match every hypothesis (x, y, yaw [rad]) on its device; returns the index of the best converged one (-1: none) and its
  // result. results (optional, n entries) receives every match.
  int64_t relocalize(const double *hypotheses, int64_t n, ndt_result *best, ndt_result *results = nullptr);

  size_t devicesUsed() const { return handles.size(); }
  double lastDeviceMs = 0.0;      // slowest shard, device time

 private:
  std::vector<ndt_handle> handles;
  std::vector<int> device_of;
  std::vector<void *> d_guess, d_res;          // per handle: device buffers for its shard
  std::vector<int64_t> cap;
  void reserve(size_t k, int64_t n);
};

#endif
