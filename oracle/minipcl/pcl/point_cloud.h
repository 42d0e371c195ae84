#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include <boost/shared_ptr.hpp>
#include <Eigen/Core>
#include <Eigen/StdVector>
#include <pcl/point_types.h>
namespace pcl {
struct PCLHeader { uint32_t seq = 0; uint64_t stamp = 0; std::string frame_id; };
template <class PointT> class PointCloud {
 public:
  typedef boost::shared_ptr<PointCloud<PointT>> Ptr;
  typedef boost::shared_ptr<const PointCloud<PointT>> ConstPtr;
  PCLHeader header;
  std::vector<PointT, Eigen::aligned_allocator<PointT>> points;
  uint32_t width = 0, height = 0;
  bool is_dense = true;
  PointCloud &operator+=(const PointCloud &rhs) {
    points.insert(points.end(), rhs.points.begin(), rhs.points.end());
    width = static_cast<uint32_t>(points.size());
    height = 1;
    is_dense = (rhs.is_dense && is_dense);
    return *this;
  }
  void clear() { points.clear(); width = 0; height = 0; }
  void push_back(const PointT &p) { points.push_back(p); width = static_cast<uint32_t>(points.size()); height = 1; }
  std::size_t size() const { return points.size(); }
  PointT &operator[](std::size_t i) { return points[i]; }
  const PointT &operator[](std::size_t i) const { return points[i]; }
};
}  // namespace pcl
