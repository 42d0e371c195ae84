// PCFilter.h -- moving-object removal used by Submap::makeMap when removeMoving is set
// [REF include/ndt_slam/PCFilter.h:17-107; the launch default is removeMoving = true, ndt_mapping.launch:20].
// Map-side pre-filtering, not NDT arithmetic, but it decides which points reach the NDT target,
// so it has to make the reference's decisions point for point.
//
// difference_extraction is pcl::octree::OctreePointCloudChangeDetector: "which points of cloud_test fall into voxels that
// cloud_base does not occupy". PCL's voxels are NOT an absolute floor(p / resol) lattice: the octree's bounding box is
// anchored at the first point inserted (lower corner = p0 - resol once the key range is set up) and doubles towards the
// side a later point violates, in double precision, so voxel faces pass through p0's coordinates and a voxel's identity
// is its integer key at the moment of insertion, carried along when the box grows downwards. VoxelAnchor below follows
// that bookkeeping (PCL 1.10.0 octree_pointcloud.hpp: adoptBoundingBoxToPoint / getKeyBitSize / genOctreeKeyforPoint).
#ifndef NDT_SLAM_B200_PCFILTER_H_
#define NDT_SLAM_B200_PCFILTER_H_

#include <cmath>
#include <cstdint>
#include <limits>
#include <string>
#include <unordered_set>
#include <pcl/point_cloud.h>
#include <ros/ros.h>

namespace ndt_host {

// The growing bounding box of a PCL point-cloud octree, reduced to what voxel identity needs: the lower corner, the
// current depth, and how many voxels the corner has moved down per axis since the box was defined.
class VoxelAnchor {
 public:
  explicit VoxelAnchor(double resolution) : res(resolution) {}
  // 64-bit id of the voxel a point falls into (21 bits per axis of shift-invariant key, offset to stay positive)
  uint64_t voxel_of(float x, float y, float z) {
    const double p[3] = {x, y, z};
    grow_to(p);
    uint64_t id = 0;
    for (int a = 0; a < 3; ++a) {
      const int64_t key = (int64_t) static_cast<unsigned>((p[a] - lo[a]) / res) - moved[a];
      id = (id << 21) | (uint64_t)((key + (1 << 20)) & ((1 << 21) - 1));
    }
    return id;
  }

 private:
  double res;
  double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
  int64_t moved[3] = {0, 0, 0};
  unsigned depth = 0;
  bool defined = false;

  void grow_to(const double p[3]) {
    const float tiny = std::numeric_limits<float>::epsilon();
    for (;;) {
      bool below[3], above[3], outside = false;
      for (int a = 0; a < 3; ++a) { below[a] = p[a] < lo[a]; above[a] = p[a] >= hi[a]; outside = outside || below[a] || above[a]; }
      if (defined && !outside) return;
      if (!defined) {
        // first point: a box of one voxel around it, then widened to the smallest key range PCL allows (two voxels per
        // axis, depth 1), half of the extra on either side
        for (int a = 0; a < 3; ++a) { lo[a] = p[a] - res / 2; hi[a] = p[a] + res / 2; }
        unsigned widest = 2;
        for (int a = 0; a < 3; ++a) widest = std::max(widest, static_cast<unsigned>(std::ceil((hi[a] - lo[a] - tiny) / res)));
        depth = std::min(32u, static_cast<unsigned>(std::ceil(std::log2((double)widest) - tiny)));
        const double side = static_cast<double>(1u << depth) * res;
        for (int a = 0; a < 3; ++a) { const double extra = (side - (hi[a] - lo[a])) / 2.0; lo[a] -= extra; hi[a] += extra; }
        defined = true;
        continue;
      }
      // one more tree level: the box doubles, downwards on every axis whose upper bound the point respects
      double side = static_cast<double>(1u << depth) * res;
      for (int a = 0; a < 3; ++a)
        if (!above[a]) { lo[a] -= side; moved[a] += (int64_t)1 << depth; }
      ++depth;
      side = static_cast<double>(1u << depth) * res - tiny;
      for (int a = 0; a < 3; ++a) hi[a] = lo[a] + side;
    }
  }
};

}  // namespace ndt_host

class PCFilter {
 public:
  double resol;
  double thre_neighbor;

  PCFilter() : resol(0.05), thre_neighbor(0.1) {
    ros::param::get("resol", resol);
    ros::param::get("thre_neighbor", thre_neighbor);
  }

  // cloud_base without the points closer than thre_neighbor to any point of point_list
  pcl::PointCloud<pcl::PointXYZ>::Ptr remove_neighborPoint(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud_base,
                                                           pcl::PointCloud<pcl::PointXYZ>::Ptr point_list) {
    auto kept = std::make_shared<pcl::PointCloud<pcl::PointXYZ>>();
    for (const auto &p : cloud_base->points) {
      bool near_any = false;
      for (const auto &q : point_list->points) {
        // PCLUtil::distance_points [REF include/ndt_slam/PCLUtil.h:21-23]: float differences, sqrt of their squares
        const float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
        if (std::sqrt(dx * dx + dy * dy + dz * dz) < thre_neighbor) { near_any = true; break; }
      }
      if (!near_any) kept->points.push_back(pcl::PointXYZ(p.x, p.y, p.z));
    }
    kept->width = static_cast<uint32_t>(kept->points.size());
    kept->height = 1;
    return kept;
  }

  // points of cloud_test that fall into voxels cloud_base does not occupy (voxels anchored like PCL's octree: see above)
  pcl::PointCloud<pcl::PointXYZ>::Ptr difference_extraction(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud_base,
                                                            pcl::PointCloud<pcl::PointXYZ>::Ptr cloud_test) {
    ndt_host::VoxelAnchor anchor(resol);
    std::unordered_set<uint64_t> occupied;
    occupied.reserve(cloud_base->points.size() * 2);
    for (const auto &p : cloud_base->points)
      if (std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z)) occupied.insert(anchor.voxel_of(p.x, p.y, p.z));
    auto diff = std::make_shared<pcl::PointCloud<pcl::PointXYZ>>();
    for (const auto &p : cloud_test->points)
      if (std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z) && !occupied.count(anchor.voxel_of(p.x, p.y, p.z)))
        diff->points.push_back(pcl::PointXYZ(p.x, p.y, p.z));
    diff->width = static_cast<uint32_t>(diff->points.size());
    diff->height = 1;
    return diff;
  }

  pcl::PointCloud<pcl::PointXYZ>::Ptr pass_through(std::string axis, double lo, double hi,
                                                   pcl::PointCloud<pcl::PointXYZ>::Ptr cloud) {
    auto out = std::make_shared<pcl::PointCloud<pcl::PointXYZ>>();
    for (const auto &p : cloud->points) {
      const float v = axis == "x" ? p.x : (axis == "y" ? p.y : p.z);
      if (v >= lo && v <= hi) out->push_back(p);
    }
    return out;
  }
};

#endif
