// pcl::octree::OctreePointCloudChangeDetector<PointT> -- restated from PCL 1.10.0's published algorithm
// (octree/include/pcl/octree/impl/octree_pointcloud.hpp: addPointsFromInputCloud, addPointIdx, adoptBoundingBoxToPoint,
// getKeyBitSize, genOctreeKeyforPoint; octree2buf_base.h: switchBuffers / serializeNewLeafs). PCL itself is absent from this
// machine. TEST INFRASTRUCTURE: lets the reference's own include/ndt_slam/PCFilter.h compile into oracle/_ref unmodified.
//
// What matters for PCFilter::difference_extraction is WHICH voxel a point falls into, and PCL's voxels are not an absolute
// floor(p / resolution) lattice: the bounding box is anchored at the FIRST point added (min = p0 - resolution after
// getKeyBitSize) and doubles towards whichever side a later point violates, all in double precision, so the voxel faces
// pass through p0's coordinates. A leaf is identified by its integer key (x - min) / resolution at the CURRENT box; when the
// box grows downwards on an axis every existing key on that axis gains 2^depth (the old root becomes a child of the new
// one). The restatement keeps, per axis, that cumulative shift and stores shift-invariant keys, so leaves keep their
// identity exactly as the tree's structure does -- no re-derivation from coordinates after the fact.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <set>
#include <tuple>
#include <vector>
#include <pcl/point_cloud.h>
namespace pcl { namespace octree {
template <class PointT> class OctreePointCloudChangeDetector {
  typedef std::tuple<int64_t, int64_t, int64_t> Key;
  double resolution_;
  typename PointCloud<PointT>::ConstPtr input_;
  double min_[3] = {0, 0, 0}, max_[3] = {0, 0, 0};
  bool bounding_box_defined_ = false;
  unsigned octree_depth_ = 0;
  int64_t shift_[3] = {0, 0, 0};          // voxels the lower corner has moved down since the box was defined
  std::size_t leaf_count_ = 0;
  std::set<Key> cur_, prev_;              // leaves of the current / previous buffer (shift-invariant keys)
  std::vector<std::pair<Key, int>> cur_points_;   // (leaf, point index) of the current buffer, insertion order

  void getKeyBitSize() {
    const float minValue = std::numeric_limits<float>::epsilon();
    unsigned max_key[3];
    for (int a = 0; a < 3; ++a) max_key[a] = static_cast<unsigned>(std::ceil((max_[a] - min_[a] - minValue) / resolution_));
    const unsigned max_voxels = std::max(std::max(std::max(max_key[0], max_key[1]), max_key[2]), 2u);
    octree_depth_ = std::max(std::min(32u, static_cast<unsigned>(std::ceil(std::log2((double)max_voxels) - minValue))), 0u);
    const double octree_side_len = static_cast<double>(1u << octree_depth_) * resolution_;
    if (leaf_count_ == 0) {
      for (int a = 0; a < 3; ++a) {
        const double oversize = (octree_side_len - (max_[a] - min_[a])) / 2.0;
        min_[a] -= oversize; max_[a] += oversize;
      }
    } else {
      for (int a = 0; a < 3; ++a) max_[a] = min_[a] + octree_side_len;
    }
  }
  void adoptBoundingBoxToPoint(const double p[3]) {
    const float minValue = std::numeric_limits<float>::epsilon();
    while (true) {
      bool lower[3], upper[3];
      bool any = false;
      for (int a = 0; a < 3; ++a) { lower[a] = p[a] < min_[a]; upper[a] = p[a] >= max_[a]; any = any || lower[a] || upper[a]; }
      if (any || !bounding_box_defined_) {
        if (bounding_box_defined_) {
          double octreeSideLen = static_cast<double>(1u << octree_depth_) * resolution_;
          for (int a = 0; a < 3; ++a)
            if (!upper[a]) { min_[a] -= octreeSideLen; shift_[a] += (int64_t)1 << octree_depth_; }   // the old root becomes the upper child
          octree_depth_++;
          octreeSideLen = static_cast<double>(1u << octree_depth_) * resolution_ - minValue;
          for (int a = 0; a < 3; ++a) max_[a] = min_[a] + octreeSideLen;
        } else {
          for (int a = 0; a < 3; ++a) { min_[a] = p[a] - resolution_ / 2; max_[a] = p[a] + resolution_ / 2; }
          getKeyBitSize();
          bounding_box_defined_ = true;
        }
      } else {
        break;
      }
    }
  }
 public:
  explicit OctreePointCloudChangeDetector(double resolution) : resolution_(resolution) {}
  void setInputCloud(const typename PointCloud<PointT>::ConstPtr &c) { input_ = c; }
  void addPointsFromInputCloud() {
    for (std::size_t i = 0; i < input_->points.size(); ++i) {
      const PointT &pt = input_->points[i];
      if (!std::isfinite(pt.x) || !std::isfinite(pt.y) || !std::isfinite(pt.z)) continue;
      const double p[3] = {pt.x, pt.y, pt.z};
      adoptBoundingBoxToPoint(p);
      int64_t k[3];
      for (int a = 0; a < 3; ++a) k[a] = (int64_t) static_cast<unsigned>((p[a] - min_[a]) / resolution_) - shift_[a];   // genOctreeKeyforPoint
      const Key key(k[0], k[1], k[2]);
      if (cur_.insert(key).second) ++leaf_count_;
      cur_points_.emplace_back(key, (int)i);
    }
  }
  void switchBuffers() { prev_.swap(cur_); cur_.clear(); cur_points_.clear(); }
  // point indices of the current buffer that live in leaves the previous buffer does not have (PCL returns them in tree
  // traversal order; PCFilter only uses them as a set)
  std::size_t getPointIndicesFromNewVoxels(std::vector<int> &out, int = 0) {
    out.clear();
    for (const auto &kp : cur_points_) if (!prev_.count(kp.first)) out.push_back(kp.second);
    return out.size();
  }
};
}}  // namespace pcl::octree
