// OctreePointCloudChangeDetector, APPROXIMATE: voxels are an absolute floor(p / resolution) lattice,
// whereas PCL's octree anchors its voxels at a bounding box that grows with the data. Only PCFilter
// (moving-object removal, SURVEY.md row f1, outside the hot path) uses it.
#pragma once
#include <cmath>
#include <set>
#include <tuple>
#include <vector>
#include <pcl/point_cloud.h>
namespace pcl { namespace octree {
template <class PointT> class OctreePointCloudChangeDetector {
  double res_;
  typename PointCloud<PointT>::ConstPtr input_;
  typedef std::tuple<long, long, long> Key;
  std::set<Key> cur_, prev_;
  std::vector<int> new_idx_;
  Key key(const PointT &p) const { return Key((long)std::floor(p.x / res_), (long)std::floor(p.y / res_), (long)std::floor(p.z / res_)); }
 public:
  explicit OctreePointCloudChangeDetector(double r) : res_(r) {}
  void setInputCloud(const typename PointCloud<PointT>::ConstPtr &c) { input_ = c; }
  void addPointsFromInputCloud() {
    new_idx_.clear();
    for (std::size_t i = 0; i < input_->points.size(); ++i) {
      Key k = key(input_->points[i]);
      if (!prev_.empty() || !cur_.empty()) { if (!prev_.count(k)) new_idx_.push_back((int)i); }
      cur_.insert(k);
    }
  }
  void switchBuffers() { prev_.swap(cur_); cur_.clear(); }
  std::size_t getPointIndicesFromNewVoxels(std::vector<int> &out, int = 0) { out = new_idx_; return out.size(); }
};
}}  // namespace pcl::octree
